"""Minimal legacy-VTK (DataFile 5.1, BINARY or ASCII, UNSTRUCTURED_GRID) reader / writer.

The reference reads its test meshes with meshio (`meshio.read(get_mesh(...))`, e.g.
/root/reference/tests/test_maps.py:8, tests/test_fea_solve.py); meshio is not installable here, so this is the
on-disk side of the hot path's callers.  `read_vtk` returns the points and, like
meshio's `mesh.cells`, one block per run of consecutive cells of the same type.
"""
from __future__ import annotations

import numpy as np

VTK_NAMES = {1: "vertex", 3: "line", 5: "triangle", 9: "quad", 10: "tetra", 12: "hexahedron", 22: "triangle6",
             24: "tetra10", 29: "hexahedron27", 28: "quad9"}
_DT = {"double": ">f8", "float": ">f4", "vtktypeint64": ">i8", "vtktypeint32": ">i4", "int": ">i4", "long": ">i8"}


class _Reader:
    def __init__(self, data: bytes):
        self.b, self.pos = data, 0

    def line(self):
        while True:
            end = self.b.find(b"\n", self.pos)
            if end < 0:
                end = len(self.b)
            s = self.b[self.pos:end].decode("ascii", "replace").strip()
            self.pos = min(end + 1, len(self.b))
            if s or self.pos >= len(self.b):
                return s

    def array(self, n, dtype, binary):
        if binary:
            dt = np.dtype(_DT[dtype])
            out = np.frombuffer(self.b, dtype=dt, count=n, offset=self.pos)
            self.pos += n * dt.itemsize
            return out.astype(dt.newbyteorder("="))
        vals = []
        while len(vals) < n:
            vals += self.line().split()
        return np.array(vals[:n], dtype=np.dtype(_DT[dtype]).newbyteorder("="))


def read_vtk(path):
    """-> (points (n,3) float64, [(cell_type_name, connectivity (m, nodes) int64), ...])"""
    r = _Reader(open(path, "rb").read())
    assert r.line().startswith("# vtk DataFile")
    r.line()
    binary = r.line().upper() == "BINARY"
    assert r.line().upper() == "DATASET UNSTRUCTURED_GRID"
    points = offsets = conn = types = None
    while r.pos < len(r.b):
        head = r.line().split()
        if not head:
            continue
        key = head[0].upper()
        if key == "POINTS":
            points = r.array(3 * int(head[1]), head[2], binary).reshape(-1, 3).astype(np.float64)
        elif key == "CELLS":
            n_off, n_conn = int(head[1]), int(head[2])
            save = r.pos
            h = r.line().split()
            if h and h[0].upper() == "OFFSETS":              # DataFile 5.x: offsets + connectivity
                offsets = r.array(n_off, h[1], binary).astype(np.int64)
                h = r.line().split()
                assert h[0].upper() == "CONNECTIVITY"
                conn = r.array(n_conn, h[1], binary).astype(np.int64)
            else:                                            # DataFile <= 4.x: "count id id ..." records
                r.pos = save
                flat = r.array(n_conn, "int", binary).astype(np.int64)
                offsets, conn, i = [0], [], 0
                while i < flat.shape[0]:
                    c = int(flat[i])
                    conn.append(flat[i + 1:i + 1 + c])
                    offsets.append(offsets[-1] + c)
                    i += 1 + c
                offsets, conn = np.array(offsets, dtype=np.int64), np.concatenate(conn)
        elif key == "CELL_TYPES":
            types = r.array(int(head[1]), "int", binary).astype(np.int64)
        elif key in ("POINT_DATA", "CELL_DATA", "FIELD", "METADATA"):
            break
    assert points is not None and offsets is not None and types is not None
    blocks, start = [], 0
    for i in range(1, len(types) + 1):
        if i == len(types) or types[i] != types[start]:
            n_nodes = int(offsets[start + 1] - offsets[start])
            block = conn[offsets[start]:offsets[i]].reshape(i - start, n_nodes)
            blocks.append((VTK_NAMES.get(int(types[start]), str(int(types[start]))), block))
            start = i
    return points, blocks


def write_vtk(path, points, blocks, binary=True):
    """Inverse of read_vtk (used by the tests and to dump results)."""
    points = np.asarray(points, dtype=np.float64)
    if points.shape[1] < 3:
        points = np.hstack([points, np.zeros((points.shape[0], 3 - points.shape[1]))])
    code = {v: k for k, v in VTK_NAMES.items()}
    offsets, conn, types = [0], [], []
    for name, c in blocks:
        c = np.asarray(c, dtype=np.int64)
        for row in c:
            conn.append(row)
            offsets.append(offsets[-1] + row.shape[0])
            types.append(code[name])
    conn = np.concatenate(conn) if conn else np.zeros(0, dtype=np.int64)
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 5.1\nwritten by feb200\n" + (b"BINARY" if binary else b"ASCII") +
                b"\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {points.shape[0]} double\n".encode())

        def put(a, dt):
            if binary:
                f.write(np.asarray(a).astype(dt).tobytes())
                f.write(b"\n")
            else:
                f.write((" ".join(repr(x) if isinstance(x, float) else str(x) for x in np.asarray(a).ravel().tolist()) + "\n").encode())
        put(points, ">f8")
        f.write(f"CELLS {len(offsets)} {conn.shape[0]}\nOFFSETS vtktypeint64\n".encode())
        put(np.array(offsets), ">i8")
        f.write(b"CONNECTIVITY vtktypeint64\n")
        put(conn, ">i8")
        f.write(f"CELL_TYPES {len(types)}\n".encode())
        put(np.array(types), ">i4")
