"""Warp-stall samples per CUDA source line (needs -lineinfo and --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = ""
lines = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; S = hdr.index("# Samples"); E = hdr.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        key = (fname, int(r[0]))
        stalls = {hdr[c]: int(r[c]) for c in range(len(hdr)) if hdr[c].startswith("stall_") and "Not Issued" not in hdr[c] and r[c].isdigit() and int(r[c])}
        s, e, src, st = lines.get(key, (0, 0, r[1], collections.Counter()))
        st.update(stalls)
        lines[key] = (s + int(r[S]), e + int(r[E]), r[1], st)
tot = sum(v[0] for v in lines.values())
print("total samples", tot)
for (f, ln), (s, e, src, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} {s:6d} {100*s/max(tot,1):5.1f}% exec {e:9d} | {src.strip()[:70]:70s} {st.most_common(3)}")
