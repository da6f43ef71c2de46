"""Round-2 ncu outputs (gpurun_out/r2_*) -> tracked summaries under profiles/:
  r2_train_step_kernels.txt / .json   per kernel: launches, time, DRAM bytes of ONE training step (ncu, cold-cache, serialised)
  r2_dram_traffic.json                the same per family for one inference step (bench.py's roofline `traffic`)
  r2_<capture>_raw.csv                selected metrics of the `ncu --set full` captures
usage: python scripts/summarize_profiles_r2.py [tag]"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out_dir = os.path.join(ROOT, "profiles")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, body = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    L = {}
    for r in body:
        d = L.setdefault(int(r[0]), {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "second": 1e6}.get(r[ui], 1)
        d[r[mi]] = v * scale
    ids = sorted(L)
    last = [i for i in ids if "logmel" in L[i]["name"]][-1]
    return [L[i] for i in ids if i >= last]


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("ac::", "")
    return n if not n.startswith("at::") else "torch:" + n.split("<")[0][4:]


# ---- training step
step = launches(os.path.join(ROOT, "gpurun_out", f"{tag}_train_traffic.csv"))
fam = collections.OrderedDict()
for l in step:
    f = fam.setdefault(short(l["name"]), {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
    f["launches"] += 1
    f["time_us"] += l.get("gpu__time_duration.sum", 0.0)
    f["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
tot = sum(f["time_us"] for f in fam.values())
lines = [f"# one training step (32 clips x 10 s, bf16 frozen CNN, 2 sampled steps), ncu --metrics gpu__time_duration.sum,dram__bytes_*"
         f" --clock-control none python scripts/train_one_step.py 1 bf16: {len(step)} launches, {tot / 1e3:.3f} ms (cold-cache, serialised)",
         f"{'kernel':58s} {'n':>4s} {'time us':>10s} {'share':>7s} {'DRAM MB':>9s}"]
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["time_us"]):
    f["share_of_step"] = f["time_us"] / tot
    lines.append(f"{k[:58]:58s} {f['launches']:4d} {f['time_us']:10.1f} {100 * f['share_of_step']:6.1f}% {f['dram_bytes'] / 1e6:9.1f}")
open(os.path.join(out_dir, f"{tag}_train_step_kernels.txt"), "w").write("\n".join(lines) + "\n")
json.dump({"step_time_us": tot, "launches": len(step), "kernels": fam}, open(os.path.join(out_dir, f"{tag}_train_step_kernels.json"), "w"), indent=1)
print("\n".join(lines[:14]))

# ---- inference step (families as bench.py's roofline leg names them)
p = os.path.join(ROOT, "gpurun_out", f"{tag}_traffic.csv")
if os.path.exists(p):
    step = launches(p)
    fam_of = lambda n: ("gemm" if "gemm_tc" in n or "gemm_tn" in n else "dwconv" if "dwconv" in n else "se" if short(n).startswith(("se_kernel", "se_solo_kernel")) else "other" if "transpose" in n
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
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                         "python scripts/one_step.py 1 (last step; cold-cache, serialised launches)",
               "step_time_us": total, "families": fams}, open(os.path.join(out_dir, f"{tag}_dram_traffic.json"), "w"), indent=1)
    print(json.dumps({k: (v["launches"], round(v["time_us"]), round(v["dram_bytes"] / 1e6)) for k, v in fams.items()}))

# ---- full captures
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio")
summary = []
for name in ("greedy_heads", "beam_heads", "stem", "logmel", "se", "conv_bf16", "bigru_bwd", "bigru_fwd", "attn_bwd", "conv_tf32", "ls_ce", "adam"):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h = r[0]
    keep = [i for i, c in enumerate(h) if c in want or c in ("Kernel Name", "ID")]
    with open(os.path.join(out_dir, f"{tag}_{name}_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        for row in r:
            w.writerow([row[i] for i in keep])
    d = dict(zip([h[i] for i in keep], [r[2][i] for i in keep]))
    summary.append((name, d))
with open(os.path.join(out_dir, f"{tag}_captures_summary.txt"), "w") as f:
    for name, d in summary:
        f.write(f"== {name}: {d.get('Kernel Name', '')[:90]}\n")
        for k, v in d.items():
            if k not in ("Kernel Name", "ID"):
                f.write(f"   {k:70s} {v}\n")
print(open(os.path.join(out_dir, f"{tag}_captures_summary.txt")).read()[:3000])
