#!/usr/bin/env python
"""Headline benchmark: EffB2-Transformer batched greedy inference (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the hot path (log-mel -> EfficientNet-B2 -> KV-cached greedy decode,
max_length 20) over one batch of 64 synthetic 10 s clips (0.1*randn, the model's 16 kHz input =
a 10 s @ 32 kHz clip after the caller's resample, demo.py:36 of the reference).  Metric:
clips/sec.  One process per GPU; clips are independent so ranks shard the clips with no
data-path collective ("scaling": "weak", 64 clips per rank).

Printed JSON (one line, rank 0): value = device-timed throughput with inputs resident in HBM;
e2e = the same through the public module API with HOST (pinned) inputs and the token ids read
back to the host; roofline = the dominant kernel's algorithmic bytes/flops over its CUDA-event
time, against MEASURED_PEAKS.json; cpu_baseline = the CPU oracle port on a bounded sample.

`--impl reference` times the reference's CPU algorithm (the oracle port: the reference is pure
Python/PyTorch and /root/reference does not travel to the GPU box) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
N_SAMPLES = 160000
MAX_LEN = 20
METRIC = "clips/sec (10s clips) EffB2-Trm greedy inference"
WORKLOAD = "EffB2-Transformer batched greedy inference, batch=64x10s synthetic clips per GPU (configs[1])"


def build_models(device=None):
    import numpy as np
    from oracle import caption_model as cm     # weights + CPU baseline only; never on the product path
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "effb2_trm.npz")))
    orc = cm.build_effb2_trm(int(g["seed"]), bn_stats=g["bn_stats"])
    if device is None:
        return orc, None
    from audiocaption_b200.captioning.models.hf_wrapper import Effb2TrmCaptioningModel
    m = Effb2TrmCaptioningModel().eval()
    m.load_state_dict(orc.state_dict(), strict=True)
    return orc, m.to(device)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("hbm_gbs", 6650.0)), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def cpu_sample(orc, n_clips, max_len=MAX_LEN, reps=1, budget_s=None):
    """The CPU oracle port on `n_clips` clips of the workload; returns clips/s, cores used, seconds per pass.
    With `budget_s` the pass is repeated until about that much CPU time has been spent (bounded sample)."""
    import torch
    from oracle import caption_model as cm
    torch.set_num_threads(os.cpu_count())
    wav, lens = cm.synth_wav(n_clips, N_SAMPLES, seed=0)
    with torch.no_grad():
        orc(wav[:1], lens[:1], sample_method="greedy", max_length=max_len)     # warm-up
        t0 = time.perf_counter()
        orc(wav, lens, sample_method="greedy", max_length=max_len)
        first = time.perf_counter() - t0
        if budget_s is not None:
            reps = max(1, min(400, int(budget_s / max(first, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(reps):
            orc(wav, lens, sample_method="greedy", max_length=max_len)
        dt = (time.perf_counter() - t0) / reps
    cpu_sample.last_reps = reps
    return n_clips / dt, torch.get_num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, _ = build_models(None)
    n = 8
    times = []
    for i in range(args.warmup + args.steps):
        v, cores, dt = cpu_sample(orc, n)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = n / (ms / 1000.0)
    sample = f"{n} of the {BATCH} clips per step (full 10 s clips, greedy, max_length {MAX_LEN})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_native(args):
    import torch
    import torch.distributed as dist
    from oracle import caption_model as cm
    from audiocaption_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on the C-level stdout at communicator creation: keep stdout clean (the one
        # JSON line) by routing fd 1 to stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    orc, model = build_models(dev)
    enc, dec = model.model.model.encoder, model.model.model.decoder
    lib = _lib.lib()

    # rotating input set larger than L2 (8 x 41 MB = 328 MB > 126 MB): every step reads its clips from HBM
    n_rot = 8
    host = [cm.synth_wav(BATCH, N_SAMPLES, seed=100 * rank + i)[0].pin_memory() for i in range(n_rot)]
    devb = [h.to(dev) for h in host]
    lens = torch.full((BATCH,), N_SAMPLES, dtype=torch.long)

    def step_resident(i):
        e = enc({"wav": devb[i % n_rot], "wav_len": lens, "specaug": False})
        return dec.greedy(e["attn_emb"], e["attn_emb_len"], MAX_LEN, 1, 2, 0, need_logit=False)["seq"]

    def step_e2e(i):
        return model(host[i % n_rot], lens, sample_method="greedy", max_length=MAX_LEN)

    def run_e2e_pipelined(steps, first):
        """`steps` batches through the public serving call `submit()`: every batch's pinned-host -> device upload,
        its kernels and the read-back of its token ids are inside the timed region; the upload of batch i+1
        overlaps the kernels of batch i (copy stream), the host never blocks except on the results."""
        pending, seq = None, None
        for i in range(steps):
            nxt = model.submit(host[(first + i) % n_rot], lens, sample_method="greedy", max_length=MAX_LEN)
            if pending is not None:
                seq = pending.result()
            pending = nxt
        return pending.result()

    from audiocaption_b200 import sharding

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        l0 = lib.ac_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.ac_launch_count() - l0
        ms = sharding.max_over_ranks(ms, device=dev)      # multi-GPU: the slowest rank, never wall clock
        return ms / steps, launches

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.summary() if sampler else None
    ms_e2e_sync, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))
    # pipelined end-to-end: same public API family (submit / result), uploads overlapped with the previous batch
    run_e2e_pipelined(max(3, args.warmup // 2), 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run_e2e_pipelined(args.steps, 3)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    ms_e2e = sharding.max_over_ranks(max(e0.elapsed_time(e1), wall_ms), device=dev) / args.steps

    # ---- roofline leg: the same steps with every launch bracketed by CUDA events
    roof = None
    if rank == 0:
        lib.ac_timing_enable(1)
        n_prof = min(args.steps, 5)
        for i in range(n_prof):
            step_resident(i)
        rep = _lib.timing_report()
        lib.ac_timing_enable(0)
        tot = sum(ms for _, ms in rep.values())
        shares = {k: round(ms / tot, 4) for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
        peak, which = measured_peaks()
        # dominant kernel family = the pointwise-convolution GEMM (all tile shapes); algorithmic HBM bytes of
        # one encoder pass: every 1x1 conv reads its input activations and weights once and writes its output
        # once (fp32) -- computed from the block plan in audiocaption_b200.roofline
        from audiocaption_b200 import roofline as rl
        fam = rl.kernel_families(rep, n_prof, BATCH)
        hbm_fams = [f for f in fam.values() if f["bound"] == "hbm"]
        top = max(hbm_fams, key=lambda f: f["ms_per_step"])      # dominant roofline-bounded kernel family
        roof = {"bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
                "frac": top["achieved"] / top["peak"], "traffic": rl.measured_traffic(top["name"]), "kernel": top["name"],
                "launches_per_step": top["launches_per_step"], "ms_per_step": top["ms_per_step"],
                "share_of_step": top["ms_per_step"] / (tot / n_prof), "peak_source": which,
                "kernel_shares": shares, "families": fam}

    if rank == 0:
        cpu_v, cores, cpu_dt = cpu_sample(orc, 8, budget_s=12.0)
        total = world * BATCH
        out = {
            "metric": METRIC, "value": total / (ms_step / 1000.0), "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": BATCH, "samples_per_clip": N_SAMPLES,
                       "sample_rate": 16000, "max_length": MAX_LEN, "decode": "greedy, KV cache, all steps on device",
                       "l2": f"rotating {n_rot} input batches (328 MB > 126 MB L2)", "parallelism": f"clips sharded x{world}"},
            "clocks": clocks,
            "e2e": {"value": total / (ms_e2e / 1000.0), "unit": "clips/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": BATCH * N_SAMPLES * 4, "d2h_bytes_per_step": BATCH * MAX_LEN * 8,
                    "api": "Effb2TrmCaptioningModel.submit()/result(): pinned host input, upload of batch i+1 overlaps "
                           "the kernels of batch i; timed as max(CUDA events, host wall clock) over all batches",
                    "synchronous_forward_ms_per_step": ms_e2e_sync,
                    "synchronous_forward_value": total / (ms_e2e_sync / 1000.0)},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": {"value": cpu_v, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"8 of the {BATCH} clips per pass, {cpu_sample.last_reps} passes ({cpu_dt * cpu_sample.last_reps:.1f} s of CPU work), "
                                       "oracle port of the reference CPU path"},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
