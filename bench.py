#!/usr/bin/env python
"""Headline benchmark: EffB2-Transformer batched greedy inference (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload effb2_trm|train]

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
Python/PyTorch and /root/reference does not travel to the GPU box) on all host cores, on the SAME 64 clips per step.

The second half of BASELINE.json's metric -- "tokens/sec Cnn14-Trm train @1/2/4/8 B200" (configs[2..3]) -- is measured
by `--workload train` (one step = one optimizer step of the fused TrainStep on 32 synthetic 10 s @ 32 kHz clips per
GPU, captions of 8..22 tokens, one NCCL all-reduce of the flat gradient when N > 1); the default run attaches the same
measurement to its JSON line under "train", so the driver's bench and scaling runs carry both halves of the metric.
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
TRAIN_METRIC = "tokens/sec Cnn14Rnn-Trm training step"
TRAIN_WORKLOAD = ("Cnn14_Rnn-Transformer training step (clotho_v2/waveform/cnn14rnn_trm.yaml), random-init, synthetic "
                  "batch=32x10s @ 32 kHz per GPU, captions 8..22 tokens (configs[2]; configs[3] = the same on 8 GPUs)")
TRAIN_BATCH, TRAIN_SAMPLES, TRAIN_VOCAB = 32, 320000, 4368
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
    if args.workload == "train":
        return run_reference_train(args)
    orc, _ = build_models(None)
    n = BATCH
    times = []
    for i in range(args.warmup + args.steps):
        v, cores, dt = cpu_sample(orc, n)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = n / (ms / 1000.0)
    sample = f"all {n} clips of a step (full 10 s clips, greedy, max_length {MAX_LEN})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ====================================================================================================== training workload
def build_train_model(device):
    """TransformerModel(CrnnEncoder(Cnn14Encoder, RnnEncoder), TransformerDecoder) as the Clotho YAML builds it, seeded
    random-init weights (oracle builders: BatchNorm statistics randomised so that folding bugs cannot hide)."""
    from oracle import cnn14 as oc, crnn
    from audiocaption_b200.captioning.models.cnn_encoder import Cnn14Encoder
    from audiocaption_b200.captioning.models.crnn_trm_encoder import CrnnEncoder
    from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
    from audiocaption_b200.captioning.models.transformer_decoder import TransformerDecoder
    from audiocaption_b200.captioning.models.transformer_model import TransformerModel
    enc = CrnnEncoder(Cnn14Encoder(sample_rate=32000),
                      RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True, hidden_size=256,
                                 dropout=0.5, num_layers=3), freeze_cnn=True, freeze_cnn_bn=True)
    dec = TransformerDecoder(emb_dim=256, vocab_size=TRAIN_VOCAB, fc_emb_dim=512, attn_emb_dim=512, nlayers=2, dropout=0.2)
    m = TransformerModel(enc, dec)
    m.load_state_dict(crnn.model_state_dict(oc.build_state_dict(3), crnn.build_gru_state_dict(4),
                                            crnn.build_decoder(6, vocab_size=TRAIN_VOCAB)), strict=True)
    return m.to(device)


def train_batches(rank, n_rot):
    """wav = 0.1 * randn(B, 320000) (full length), captions seed 1 with cap_len ~ U{8..22}, tokens U{4..V-1},
    pad 0, sorted by length."""
    from oracle import caption_model as cm, train_step as ts
    out = []
    for i in range(n_rot):
        wav, lens = cm.synth_wav(TRAIN_BATCH, TRAIN_SAMPLES, seed=1000 * rank + i, sample_rate=32000)
        cap, cap_len = ts.synth_captions(TRAIN_BATCH, 22, TRAIN_VOCAB, seed=1 + 17 * rank + i, min_len=8)
        out.append({"wav": wav.pin_memory(), "wav_len": lens, "cap": cap.pin_memory(), "cap_len": cap_len.numpy()})
    return out


def run_train_leg(dev, rank, world, steps, warmup, timed, lib, ss_ratio=0.99):
    """K optimizer steps; returns the "train" record (rank 0) or None."""
    import random
    import torch
    from audiocaption_b200 import _lib
    from audiocaption_b200.train_step import TrainStep
    model = build_train_model(dev)
    step = TrainStep(model, total_iters=10 ** 9, lr=5e-4, warmup_iters=3000)     # ss_ratio and lr effectively constant
    n_rot = 4                                                                      # 4 x 41 MB of waveforms > 126 MB L2
    host = train_batches(rank, n_rot)
    devb = [dict(b, wav=b["wav"].to(dev), cap=b["cap"].to(dev)) for b in host]
    torch.manual_seed(1 + rank)
    random.seed(1)                       # the same coin sequence on every rank keeps the step shapes in lock-step
    step.ss_ratio = ss_ratio
    tokens_per_step = [int((b["cap_len"] - 1).sum()) for b in host]
    keep = {}

    def step_resident(i):
        keep["loss"] = step.step(devb[i % n_rot])["loss"]

    def step_pipelined(i):
        # device-resident inputs, look-ahead on: the frozen encoder of batch i+1 runs on a second stream beside the trainable
        # part of step i (TrainStep.prefetch); every timed step still contains one full encoder pass and one optimizer step
        cur = keep.pop("staged_dev", None) or step.prefetch(devb[i % n_rot])
        keep["staged_dev"] = step.prefetch(devb[(i + 1) % n_rot])
        keep["loss"] = step.step(cur)["loss"]

    def step_e2e(i):
        # every step: one pinned H2D upload (of the NEXT batch, on the copy stream, overlapping this step's kernels -- the
        # prefetch a DataLoader does for the reference's loop) and one D2H read of this step's loss
        cur = keep.pop("staged", None) or step.prefetch(host[i % n_rot])
        keep["staged"] = step.prefetch(host[(i + 1) % n_rot])
        keep["loss_host"] = step.step(cur)["loss"].item()

    def step_e2e_sync(i):
        keep["loss_host"] = step.step(host[i % n_rot])["loss"].item()          # upload, step and loss read strictly in order

    model.encoder.cnn.conv_precision = "fp32"
    ms_fp32, _ = timed(step_resident, steps, warmup)
    model.encoder.cnn.conv_precision = "tf32"
    ms_tf32, _ = timed(step_resident, steps, warmup)
    # BASELINE configs[2..3] are stated in bf16: the headline runs the frozen CNN in bf16 (activations and weights, fp32
    # accumulation; csrc/conv_bf16.cu); the TF32 and the fp32-exact (3xTF32) modes are timed beside it
    model.encoder.cnn.conv_precision = "bf16"
    ms_inline, launches = timed(step_resident, steps, warmup)
    ms_step, _ = timed(step_pipelined, steps, warmup)
    keep.pop("staged_dev", None)
    torch.cuda.synchronize()
    schedule = "look-ahead"
    if ms_step > ms_inline:              # no SM partition on this driver (ordinary streams do not pay): report the plain schedule
        ms_step, schedule = ms_inline, "back to back (the look-ahead schedule was slower here)"
    # host time to enqueue ONE step on an idle GPU (a loop of many steps is throttled by the launch queue, not the host)
    host_ms = 0.0
    for i in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_resident(i)
        host_ms += (time.perf_counter() - t0) * 1e3 / 5
    torch.cuda.synchronize()
    ms_e2e, _ = timed(step_e2e, steps, 3)
    keep.pop("staged", None)
    ms_e2e_sync, _ = timed(step_e2e_sync, steps, 3)
    # share of the sampled (two-row) path in real training: the YAML's ratio goes 1.0 -> 0.7, mean 0.85
    step.ss_ratio = 0.85
    ms_ss085, _ = timed(step_resident, steps, 3)
    step.ss_ratio = ss_ratio
    rec = None
    # per-kernel CUDA-event timing: EVERY rank runs the same steps (each contains the gradient all-reduce); rank 0 records
    n_prof = min(steps, 5)
    if rank == 0:
        lib.ac_timing_enable(1)
    for i in range(n_prof):
        step_resident(i)
    torch.cuda.synchronize()
    if rank == 0:
        rep = _lib.timing_report()
        lib.ac_timing_enable(0)
        spans = {k: round(ms / n_prof, 4) for k, (n, ms) in rep.items() if k.startswith("span_")}
        rep = {k: v for k, v in rep.items() if not k.startswith("span_")}      # spans contain the per-kernel timers
        tot = sum(ms for _, ms in rep.values())
        shares = {k: round(ms / tot, 4) for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
        per_kernel = {k: {"launches_per_step": n / n_prof, "ms_per_step": round(ms / n_prof, 4)} for k, (n, ms) in rep.items()}
        conv_ms = (rep.get("conv3x3_bf16", (0, 0.0))[1] + rep.get("conv3x3_tc", (0, 0.0))[1]) / n_prof
        bf16_burst, bf16_sust = measured_tensor_peaks()
        flop = 40.07e9 * TRAIN_BATCH                       # Cnn14 = 40.07 GFLOP per clip (2 x MAC)
        achieved = flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        tok = sum(tokens_per_step) / len(tokens_per_step)
        rec = {
            "metric": TRAIN_METRIC, "value": world * tok / (ms_step / 1000.0), "unit": "tokens/s", "n_gpus": world,
            "ms_per_step": ms_step, "steps": steps, "clips_per_s": world * TRAIN_BATCH / (ms_step / 1000.0),
            "tokens_per_step_per_gpu": tok,
            "dtype": "bf16 (frozen Cnn14: bf16 activations and weights, fp32 accumulate; trainable path: 3xTF32 GEMMs = fp32-level, "
                     "fp32 master weights and optimizer)",
            "tf32_mode": {"value": world * tok / (ms_tf32 / 1000.0), "unit": "tokens/s", "ms_per_step": ms_tf32,
                          "what": "the same step with the convolutions on fp32 activations with plain TF32 operands"},
            "fp32_mode": {"value": world * tok / (ms_fp32 / 1000.0), "unit": "tokens/s", "ms_per_step": ms_fp32,
                          "what": "the same step with the convolutions in 3xTF32 (fp32-level accuracy everywhere)"},
            "scaling": "weak", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD, "clips_per_gpu": TRAIN_BATCH, "vocab": TRAIN_VOCAB,
                       "ss_ratio": ss_ratio, "dropout": "on (YAML values)", "trainable_params": step.n_trainable,
                       "collective": "one all_reduce(sum) of the flat fp32 gradient per step" if world > 1 else "none (1 GPU)",
                       "l2": f"rotating {n_rot} batches (4 x 41 MB of waveforms > 126 MB L2)"},
            "ms_per_step_ss_ratio_0.85": ms_ss085,
            "pipeline": {"what": "value / ms_per_step: TrainStep.prefetch(next batch) + TrainStep.step(staged batch) -- the frozen "
                                 "Cnn14 encoder of batch i+1 runs on a second stream (convolutions capped at cnn_sms persistent CTAs) "
                                 "beside the bi-GRU / decoder forward + backward + optimizer of batch i; unpipelined = "
                                 "TrainStep.step(batch) alone, encoder and trainable part back to back on one stream",
                         "schedule": schedule, "cnn_sms": step.cnn_sms, "trainable_sms": getattr(step, "train_sms", None),
                         "sm_partition": "green contexts" if getattr(step, "partition", None) is not None else
                                         f"none ({getattr(step, 'partition_error', None)})",
                         "unpipelined_ms_per_step": ms_inline,
                         "unpipelined_value": world * (sum(tokens_per_step) / len(tokens_per_step)) / (ms_inline / 1000.0)},
            "e2e": {"value": world * tok / (ms_e2e / 1000.0), "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": TRAIN_BATCH * TRAIN_SAMPLES * 4 + int(host[0]["cap"].numel()) * 8,
                    "d2h_bytes_per_step": 4,
                    "api": "TrainStep.prefetch(batch i+1) + TrainStep.step(staged batch i): pinned host waveforms + captions in "
                           "(upload of the next batch overlaps the step), loss.item() out every step",
                    "unpipelined_ms_per_step": ms_e2e_sync},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms,
            "roofline": {"bound": "tensor", "kernel": "conv3x3_bf16 (frozen Cnn14 forward, 11 launches)", "achieved": achieved,
                         "peak": bf16_sust, "unit": "TFLOP/s", "frac": achieved / bf16_sust if bf16_sust else None,
                         "traffic": None, "ms_per_step": conv_ms, "share_of_step": conv_ms / ms_inline,
                         "note": "achieved = algorithmic flops (40.07 GFLOP/clip) of the bf16 convolutions over their CUDA-event "
                                 "time; peak = measured dense bf16",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained", "kernel_shares": shares,
                         "per_kernel": per_kernel, "spans_ms_per_step": spans},
        }
    del step, model
    torch.cuda.empty_cache()
    return rec


def measured_tensor_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d.get("bf16_tflops", 1600.0)), float(d.get("bf16_tflops_sustained", 1400.0))
    except Exception:
        return 1600.0, 1400.0


def cpu_train_sample(n_clips, reps=1):
    """The oracle's literal training step (oracle/train_step.py: step-by-step decoder loop, torch autograd, Adam) on the
    host cores; returns tokens/s, cores, seconds per step."""
    import random
    import torch
    from oracle import caption_model as cm, cnn14 as oc, crnn, train_step as ts
    torch.set_num_threads(os.cpu_count())
    wav, lens = cm.synth_wav(n_clips, TRAIN_SAMPLES, seed=0, sample_rate=32000)
    cap, cap_len = ts.synth_captions(n_clips, 22, TRAIN_VOCAB, seed=1, min_len=8)
    cnn_sd, rnn_sd = oc.build_state_dict(3), crnn.build_gru_state_dict(4)
    dec = crnn.build_decoder(6, vocab_size=TRAIN_VOCAB)
    random.seed(1)
    L = cap.size(1) - 1
    t0 = time.perf_counter()
    for _ in range(reps):
        coins = [random.random() < 0.99 for _ in range(L)]
        ts.train_step(cnn_sd, rnn_sd, dec, wav, lens, cap, cap_len, coins, 5e-4)
    dt = (time.perf_counter() - t0) / reps
    return float((cap_len - 1).sum()) / dt, torch.get_num_threads(), dt


def run_reference_train(args):
    times, toks = [], None
    n = 8
    for i in range(args.warmup + args.steps):
        v, cores, dt = cpu_train_sample(n)
        if i >= args.warmup:
            times.append(dt)
            toks = v * dt
        if sum(times) > 150:          # bounded: the whole run stays within a few minutes
            break
    ms = 1000.0 * sum(times) / len(times)
    value = toks / (ms / 1000.0)
    sample = f"{n} of the {TRAIN_BATCH} clips per step, {len(times)} timed steps (full 10 s clips, oracle port of the reference step)"
    print(json.dumps({
        "impl": "reference", "metric": TRAIN_METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": TRAIN_WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ====================================================================================================== other BASELINE configs
def run_config1_leg(model, dev, timed):
    """configs[0]: EffB2-Transformer greedy, batch = 1, one 10 s 32 kHz clip, the reference's demo.py path (demo.py:36
    resamples to the model's 16 kHz first) -- here the resample runs on the device inside the timed call."""
    import torch
    from oracle import caption_model as cm
    from audiocaption_b200.resample import resample
    clips = [cm.synth_wav(1, 320000, seed=50 + i, sample_rate=32000)[0].pin_memory() for i in range(4)]
    lens = torch.tensor([160000])

    def one(i):
        wav16 = resample(clips[i % 4].to(dev, non_blocking=True), 32000, 16000)
        return model(wav16, lens, sample_method="greedy", max_length=MAX_LEN)           # host -> ... -> token ids on the host

    ms, launches = timed(one, 30, 5)
    return {"workload": "EffB2-Transformer greedy inference, batch=1, single 10 s 32 kHz clip, resample 32->16 kHz on the device "
                        "(configs[0], demo.py path)", "latency_ms": ms, "value": 1000.0 / ms, "unit": "clips/s",
            "launches_per_call": launches / 30, "h2d_bytes": 320000 * 4, "d2h_bytes": MAX_LEN * 8}


def run_config5_leg(dev, rank, world, timed, steps=10):
    """configs[4]: Cnn14Rnn-TempAttnGru (HF `Cnn14RnnTempAttnGruModel.forward`: log-mel -> SED tagger -> Cnn14 -> bi-GRU ->
    GRU-attention beam search), beam 4, 16 clips per GPU (128 over 8 GPUs), sharded over clips, no collective."""
    import torch
    from oracle import bah_decoder as bd, caption_model as cm, cnn14 as oc, crnn, sed as sed_o
    from audiocaption_b200.captioning.models import hf_wrapper as hw
    m = hw.Cnn14RnnTempAttnGruModel().eval()
    cnn_sd = oc.build_state_dict(3)
    sd = {f"cap_model.encoder.cnn.{k}": v for k, v in cnn_sd.items()}
    sd.update({f"cap_model.encoder.rnn.{k}": v for k, v in crnn.build_gru_state_dict(4).items()})
    sd.update({f"cap_model.decoder.{k}": v for k, v in bd.build_state_dict(8).items()})
    sd.update({f"sed_model.{k}": v for k, v in sed_o.build_state_dict(12).items()})
    sd.update({k: v for k, v in cnn_sd.items() if k.startswith("melspec")})
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    B = 16
    host = [cm.synth_wav(B, 320000, seed=200 + 10 * rank + i, sample_rate=32000)[0].pin_memory() for i in range(8)]   # 164 MB > L2
    devb = [h.to(dev) for h in host]
    lens = torch.full((B,), 320000, dtype=torch.long)
    out = {}
    for prec in ("fp32", "tf32", "bf16"):
        m.cap_model.encoder.cnn.conv_precision = prec
        m.sed_model.conv_precision = prec
        ms, launches = timed(lambda i: m(devb[i % 8], lens, sample_method="beam", beam_size=4, max_length=MAX_LEN), steps, 3)
        ms_e2e, _ = timed(lambda i: m(host[i % 8], lens, sample_method="beam", beam_size=4, max_length=MAX_LEN), steps, 3)
        out[prec] = {"value": world * B / (ms / 1000.0), "unit": "clips/s", "ms_per_step": ms,
                     "e2e": {"value": world * B / (ms_e2e / 1000.0), "unit": "clips/s", "ms_per_step": ms_e2e,
                             "h2d_bytes_per_step": B * 320000 * 4, "d2h_bytes_per_step": B * MAX_LEN * 8},
                     "launches_per_step": launches / steps}
    del m, devb
    torch.cuda.empty_cache()
    return {"workload": "Cnn14Rnn-TempGRU temporal model, beam 4, SED tagger on, 16 clips x 10 s @ 32 kHz per GPU "
                        f"(configs[4]: 128 clips over 8 GPUs), {world} GPU(s), sharded over clips, no collective",
            "n_gpus": world, "bf16": out["bf16"], "tf32": out["tf32"], "fp32": out["fp32"],
            "note": "bf16 = Cnn14 encoder and SED tagger convolutions on bf16 activations / weights; tf32 = plain-TF32 "
                    "convolutions; fp32 = 3xTF32 everywhere"}


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
            import datetime
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from audiocaption_b200 import sharding
    lib = _lib.lib()

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
        t0 = time.perf_counter()
        for i in range(steps):
            fn(warmup + i)
        timed.host_ms_per_step = (time.perf_counter() - t0) * 1e3 / steps     # host time to ENQUEUE a step (no sync inside fn)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.ac_launch_count() - l0
        ms = sharding.max_over_ranks(ms, device=dev)      # multi-GPU: the slowest rank, never wall clock
        return ms / steps, launches

    if args.workload == "train":
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        rec = run_train_leg(dev, rank, world, args.steps, args.warmup, timed, lib)
        clocks = sampler.summary() if sampler else None
        if rank == 0:
            cpu_v, cores, cpu_dt = cpu_train_sample(4)
            rec.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "clocks": clocks,
                        "cpu_baseline": {"value": cpu_v, "unit": "tokens/s", "cores": cores, "kind": "port",
                                         "sample": f"4 of the {TRAIN_BATCH} clips, one step ({cpu_dt:.1f} s of CPU work), oracle port "
                                                   "of the reference training step"}})
            print(json.dumps(rec))
        if world > 1:
            dist.destroy_process_group()
        return

    orc, model = build_models(dev)
    enc, dec = model.model.model.encoder, model.model.model.decoder

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


    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
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

    # sustained figure: the same step looped for >= 2 s (the K-step region above is a ~0.1 s burst)
    n_sust = max(args.steps, int(2000.0 / max(ms_step, 0.1)) + 1)
    ms_sust, _ = timed(step_resident, n_sust, 0)
    clocks = sampler.summary() if sampler else None        # sampled over the K-step region AND the 2 s sustained loop

    # like-for-like GPU baseline: the oracle port (stock PyTorch eager: cuFFT / cuDNN / cuBLAS kernels) on the same B200
    eager = None
    if rank == 0:
        try:
            orc_dev = orc.to(dev).eval()
            with torch.no_grad():
                def eager_step(i):
                    return orc_dev(devb[i % n_rot], lens, sample_method="greedy", max_length=MAX_LEN)
                for i in range(2):
                    eager_step(i)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n_eager = 5
                for i in range(n_eager):
                    eager_step(i)
                torch.cuda.synchronize()
                ms_eager = (time.perf_counter() - t0) * 1000.0 / n_eager
            eager = {"value": BATCH / (ms_eager / 1000.0), "unit": "clips/s", "ms_per_step": ms_eager, "steps": n_eager,
                     "what": "oracle port of the reference moved to the same GPU with .to('cuda'): stock PyTorch eager "
                             "(cuFFT, cuDNN, cuBLAS, python decode loop); fp32 with torch's default TF32 settings"}
        except Exception as e:                 # a baseline leg must never take the bench line down
            eager = {"unavailable": repr(e)[:200]}
        finally:
            orc.to("cpu")

    # ---- the other BASELINE configs, same process, same GPUs: configs[0] (B = 1 demo path), configs[2..3] (training step,
    # tokens/s: the second half of the metric), configs[4] (temporal captioner, beam 4)
    train = config1 = config5 = None
    if not args.no_train:
        config1 = run_config1_leg(model, dev, timed)
        del devb
        torch.cuda.empty_cache()
        train = run_train_leg(dev, rank, world, min(args.steps, 30), args.warmup, timed, lib)
        config5 = run_config5_leg(dev, rank, world, timed)
        devb = [h.to(dev) for h in host]

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
            "sustained": {"value": total / (ms_sust / 1000.0), "unit": "clips/s", "ms_per_step": ms_sust, "steps": n_sust,
                          "seconds": ms_sust * n_sust / 1000.0},
            "gpu_eager_baseline": eager,
            "train": train,
            "config1_single_clip": config1,
            "config5_tempgru_beam4": config5,
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
    ap.add_argument("--workload", default="effb2_trm", choices=["effb2_trm", "train"])
    ap.add_argument("--no-train", action="store_true", help="effb2_trm workload: skip the attached training measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
