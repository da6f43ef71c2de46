"""Top stalled SASS instructions of an ncu report's source page.
usage: python scripts/ncu_hot.py report.ncu-rep [top_n] [kernel-index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
start = hdr_i[k]; end = hdr_i[k + 1] - 1 if k + 1 < len(hdr_i) else len(rows)
hdr = rows[start]; body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]; E = ci["Instructions Executed"]
tot = sum(int(r[S]) for r in body)
print("kernel", rows[start - 1][:2], "total samples", tot, "instructions", len(body))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall")]
idx = sorted(range(len(body)), key=lambda i: -int(body[i][S]))[:top]
for i in sorted(idx):
    r = body[i]
    reasons = sorted(((int(r[c]), hdr[c]) for c in range(len(hdr)) if hdr[c].startswith("stall_") and r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[S]):6d} {100*int(r[S])/max(tot,1):5.1f}% exec {r[E]:>8s}  {r[1].strip()[:90]:90s} {reasons}")
