"""Aggregate warp-stall samples of an ncu source page over SASS instruction index ranges.
usage: python scripts/ncu_ranges.py report.ncu-rep a:b [a:b ...]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]; body = [r for r in rows[start + 1:] if len(r) == len(hdr)]
S = hdr.index("# Samples")
for rg in sys.argv[2:]:
    a, b = [int(x) for x in rg.split(":")]
    agg = collections.Counter(); tot = 0
    for r in body[a:b]:
        tot += int(r[S])
        for c, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and r[c].isdigit():
                agg[h] += int(r[c])
    print(rg, "samples", tot, agg.most_common(8))
