"""Phase / stall breakdown of one kernel from an ncu report: python profiles/ncu_phase_report.py rep.ncu-rep
Segments the SASS at barriers and prints samples and the top stall reasons per segment, plus hot instructions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, rows = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[2]) for r in rows)
print("total samples", tot, "instructions", len(rows))
start, acc, st = 0, 0, {}
def flush(i, why):
    global start, acc, st
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    print(f"rows {start:5d}-{i:5d} {acc:8d} ({100*acc/max(tot,1):5.1f}%) {why[:40]:40s} " + " ".join(f"{k[6:]}={v}" for k, v in top))
    start, acc, st = i + 1, 0, {}
for i, r in enumerate(rows):
    acc += int(r[2])
    for c in cols:
        v = int(r[c] or 0)
        if v: st[hdr[c]] = st.get(hdr[c], 0) + v
    ins = r[1].strip()
    if "BAR.SYNC" in ins or ("ATOMS" in ins and "ADD" in ins) or "EXIT" in ins or "WARPSYNC" in ins:
        flush(i, ins)
flush(len(rows) - 1, "tail")
print("--- hot instructions")
for i in sorted(sorted(range(len(rows)), key=lambda i: -int(rows[i][2]))[:top_n]):
    r = rows[i]
    st = sorted(((hdr[c][6:], int(r[c] or 0)) for c in cols), key=lambda kv: -kv[1])[:2]
    print(i, r[2], r[5], r[1].strip()[:80], st)
