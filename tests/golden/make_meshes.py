"""Convert a few of the reference's test meshes (tests/meshes/*.vtk, read-only under /root/reference) into
tests/golden/reference_meshes.npz with the product's own legacy-VTK reader, so that GPU tests can use the
reference's real (unstructured) meshes on a box where /root/reference does not exist.
Run:  python tests/golden/make_meshes.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "fea-in-jax_b200"))
from fe_jax.vtk_io import read_vtk   # noqa: E402

SRC = "/root/reference/tests/meshes"
out = {}
for name in ("polygon_mesh_0.5", "polygon_mesh_0.1", "polygon_mesh_0.05", "rectangle_with_hole", "microscale_2D_r0"):
    pts, blocks = read_vtk(os.path.join(SRC, name + ".vtk"))
    key = name.replace(".", "p")
    out[key + "__points"] = pts
    for btype, conn in blocks:
        out[f"{key}__{btype}"] = conn
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_meshes.npz"), **out)
print({k: v.shape for k, v in out.items()})
