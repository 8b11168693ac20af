"""Shared-memory wavefronts per instruction from an ncu report: python profiles/ncu_smem_report.py rep.ncu-rep [unit_count]
Prints every shared-memory instruction (program order) with executed count, wavefronts (actual / ideal), and samples;
then totals per opcode.  unit_count (e.g. vertices) scales the totals to per-unit numbers."""
import csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, rows = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
W, WI, IE, SM = ci["L1 Wavefronts Shared"], ci["L1 Wavefronts Shared Ideal"], ci["Instructions Executed"], ci["# Samples"]
tot = {}
tw = ti = 0
tinst = sum(int(r[IE]) for r in rows)
tsamp = sum(int(r[SM]) for r in rows)
for i, r in enumerate(rows):
    w = int(r[W] or 0)
    if w == 0:
        continue
    ins = r[1].strip()
    op = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
    wi = int(r[WI] or 0)
    print(f"{i:5d} {ins[:58]:58s} exec {int(r[IE]):>11d} wf {w:>11d} ideal {wi:>11d} ({w / max(int(r[IE]), 1):5.2f}/inst) samples {r[SM]}")
    t = tot.setdefault(op, [0, 0, 0])
    t[0] += int(r[IE]); t[1] += w; t[2] += wi
    tw += w; ti += wi
print("--- per opcode: executed, wavefronts, ideal", "(per unit)" if units != 1.0 else "")
for op, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{op:20s} {t[0] / units:14.2f} {t[1] / units:14.2f} {t[2] / units:14.2f}")
print(f"total wavefronts {tw / units:.2f} ideal {ti / units:.2f}; instructions {tinst / units:.2f}; samples {tsamp}")
