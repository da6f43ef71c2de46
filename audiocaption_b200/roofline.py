"""Algorithmic work of the hot-path kernels (fp32), used by bench.py's roofline leg.

Bytes are the compulsory HBM traffic of each kernel family for one batch: every operand read
once, every result written once (weights once per launch).  DESIGN.md states the same figures
per clip."""
import ctypes
import json
import os

from . import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm, src = 6650.0, "fallback"
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            hbm, src = float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return hbm, src


def effb2_geometry(n_mels=64, n_frames=1001):
    """[(block plan tuple, (Hin, Win), (Hout, Wout))] + stem/head dims."""
    l = _lib.lib()
    info = (ctypes.c_int * 9)()
    n = l.ac_effb2_block_info(0, info)
    H, W = (n_mels + 1 - 3) // 2 + 1, (n_frames + 1 - 3) // 2 + 1
    stem = (H, W)
    out = []
    for i in range(n):
        l.ac_effb2_block_info(i, info)
        cin, cout, e, k, s, lo, hi, nsq, skip = tuple(info)
        Ho, Wo = (H + lo + hi - k) // s + 1, (W + lo + hi - k) // s + 1
        out.append(((cin, cout, e, k, s, lo, hi, nsq, skip), (H, W), (Ho, Wo)))
        H, W = Ho, Wo
    return stem, out, (H, W)


def per_clip_work(n_samples=160000, hop=160, n_mels=64, t_mem=32, d_enc=1408, d=256, nlayers=2):
    """{family: (bytes per clip, bytes per launch-independent weights, flops per clip)}"""
    T = 1 + n_samples // hop
    stem, blocks, last = effb2_geometry(n_mels, T)
    f = 4
    gemm_b = gemm_w = gemm_fl = 0
    dw_b = dw_fl = 0
    for (cin, cout, e, k, s, lo, hi, nsq, skip), (H, W), (Ho, Wo) in blocks:
        ce, pin, pout = cin * e, H * W, Ho * Wo
        if e != 1:
            gemm_b += pin * (cin + ce) * f; gemm_w += cin * ce * f; gemm_fl += 2 * pin * cin * ce
        dw_b += (pin + pout) * ce * f; dw_fl += 2 * pout * ce * k * k
        gemm_b += pout * (ce + cout + (cout if skip else 0)) * f; gemm_w += ce * cout * f
        gemm_fl += 2 * pout * ce * cout
    px = last[0] * last[1]
    gemm_b += px * (352 + 1408) * f; gemm_w += 352 * 1408 * f; gemm_fl += 2 * px * 352 * 1408
    # decoder memory preparation GEMMs (attn_proj, cross-attention K/V of both layers)
    gemm_b += t_mem * (d_enc + d) * f + nlayers * t_mem * (d + 2 * d) * f
    gemm_w += d_enc * d * f + nlayers * 2 * d * d * f
    gemm_fl += 2 * t_mem * d_enc * d + nlayers * 2 * t_mem * d * 2 * d
    return {
        "gemm": dict(bytes=gemm_b, weight_bytes=gemm_w, flops=gemm_fl),
        "dwconv": dict(bytes=dw_b, weight_bytes=0, flops=dw_fl),
        "logmel": dict(bytes=4 * n_samples + 4 * n_mels * T, weight_bytes=0, flops=0),
        "stem": dict(bytes=4 * n_mels * T + 4 * stem[0] * stem[1] * 32, weight_bytes=0, flops=2 * stem[0] * stem[1] * 32 * 9),
    }


def kernel_families(report, n_steps, batch):
    """report = _lib.timing_report(); returns per-family achieved GB/s against the HBM peak."""
    hbm, _ = peaks()
    work = per_clip_work()
    fam = {}
    for name, w in work.items():
        sel = {k: v for k, v in report.items() if k.startswith(name)}
        if not sel:
            continue
        ms = sum(v[1] for v in sel.values()) / n_steps
        launches = sum(v[0] for v in sel.values()) / n_steps
        nbytes = w["bytes"] * batch + w["weight_bytes"]
        fam[name] = {"name": name, "bound": "hbm", "unit": "GB/s", "peak": hbm,
                     "achieved": nbytes / (ms * 1e-3) / 1e9, "ms_per_step": ms, "launches_per_step": launches,
                     "algorithmic_bytes_per_step": nbytes, "gflop_per_step": w["flops"] * batch / 1e9,
                     "tflops": w["flops"] * batch / (ms * 1e-3) / 1e12}
    for name in ("trm_greedy", "trm_beam", "se", "freq_mean", "layernorm_rows"):
        sel = {k: v for k, v in report.items() if k.startswith(name)}
        if sel:
            ms = sum(v[1] for v in sel.values()) / n_steps
            fam[name] = {"name": name, "bound": "latency", "unit": "GB/s", "peak": hbm, "achieved": 0.0,
                         "ms_per_step": ms, "launches_per_step": sum(v[0] for v in sel.values()) / n_steps}
    return fam


def measured_traffic(family):
    """dram__bytes_read.sum + dram__bytes_write.sum of the family's launches in one step, from the committed ncu
    capture (profiles/r2_dram_traffic.json, produced by scripts/gpu_profile_r2.sh + scripts/summarize_profiles_r2.py); None when absent."""
    for name in ("r2_dram_traffic.json", "r1_dram_traffic.json"):          # newest committed capture first
        p = os.path.join(ROOT, "profiles", name)
        try:
            d = json.load(open(p))
            f = d["families"][family]
            return {"bytes_per_step": f["dram_bytes"], "launches_per_step": f["launches"],
                    "bytes_per_launch": f["dram_bytes"] / max(f["launches"], 1), "source": f"profiles/{name}: " + d.get("source", "")}
        except Exception:
            continue
    return None
