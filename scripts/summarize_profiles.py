"""Turns the ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/:
  r1_dram_traffic.json  per kernel family: launches, time, DRAM bytes of ONE step (ncu, cold-cache, serialised)
  r1_<name>_raw.csv     selected metrics of the full captures (ncu --set full) of the top kernels
usage: python scripts/summarize_profiles.py [tag]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles"); os.makedirs(out_dir, exist_ok=True)

rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{tag}_traffic.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, body = rows[hi], rows[hi + 1:]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = {}
for r in body:
    d = launches.setdefault(int(r[0]), {"name": r[ki]})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "second": 1e6}.get(u, 1)
    d[r[mi]] = v * scale
ids = sorted(launches)
last_logmel = [i for i in ids if "logmel" in launches[i]["name"]][-1]
step = [launches[i] for i in ids if i >= last_logmel]
fam_of = lambda n: ("gemm" if "gemm_tc" in n or "gemm_tn" in n else "dwconv" if "dwconv" in n else "se" if "se_kernel" in n
                    else "logmel" if "logmel" in n else "stem" if "stem" in n else "trm_greedy" if "greedy" in n else "other")
fams = {}
for l in step:
    f = fams.setdefault(fam_of(l["name"]), {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
    f["launches"] += 1
    f["time_us"] += l.get("gpu__time_duration.sum", 0.0)
    f["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
total = sum(f["time_us"] for f in fams.values())
for f in fams.values():
    f["share_of_step"] = f["time_us"] / total
json.dump({"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                     f"python scripts/one_step.py 1 (last step; cold-cache, serialised launches)",
           "step_time_us": total, "families": fams}, open(os.path.join(out_dir, f"{tag}_dram_traffic.json"), "w"), indent=1)
print(json.dumps(fams, indent=1))

want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic", "launch__cluster_size")
for name in sys.argv[2:] or [f"{tag}_gemm_b2", f"{tag}_dw_b2"]:
    rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h = r[0]
    keep = [i for i, c in enumerate(h) if c in want or c in ("Kernel Name", "ID")]
    with open(os.path.join(out_dir, name + "_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        for row in r:
            w.writerow([row[i] for i in keep])
    print(name, "->", [dict(zip([h[i] for i in keep], row_)) for row_ in ([ [row[i] for i in keep] for row in r[2:] ])][:2])
