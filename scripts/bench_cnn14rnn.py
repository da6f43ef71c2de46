#!/usr/bin/env python
"""Secondary benchmark (not the driver's bench line): Cnn14Rnn-Transformer greedy inference
(eg_configs/audiocaps/waveform/cnn14rnn_trm.yaml), batch = 64 x 10 s @ 32 kHz synthetic clips, 1 GPU.

    python scripts/bench_cnn14rnn.py [--steps K] [--warmup W] [--batch B] [--out file.json]

Same JSON shape as bench.py: value = device-resident clips/s, e2e = through TransformerModel.forward with pinned host
input and the token ids read back, roofline = the 3x3 convolutions (tensor-bound: 2*MACs x 3 tf32 MMAs per fp32 product)
against the measured dense bf16 peak / 2 (tf32 runs at half the bf16 rate), cpu_baseline = the oracle port on one clip."""
import argparse, json, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--out", default=None)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--decoder", default="trm", choices=["trm", "tempgru"], help="trm: cnn14rnn_trm.yaml; tempgru: HF Cnn14RnnTempAttnGru (config 5)")
ap.add_argument("--beam", type=int, default=0, help="0 = greedy")
ap.add_argument("--sed", action="store_true", help="tempgru only: full HF forward, temporal tags from the SED tagger")
args = ap.parse_args()

from oracle import cnn14 as oc, crnn, caption_model as cm          # weights + CPU baseline only
from audiocaption_b200 import _lib
from audiocaption_b200.captioning.models.cnn_encoder import Cnn14Encoder
from audiocaption_b200.captioning.models.crnn_trm_encoder import CrnnEncoder
from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
from audiocaption_b200.captioning.models.transformer_decoder import TransformerDecoder
from audiocaption_b200.captioning.models.transformer_model import TransformerModel

dev = torch.device("cuda", 0)
B, N, MAX_LEN = args.batch, 320000, 20
cnn_sd, rnn_sd, dec_o = oc.build_state_dict(3), crnn.build_gru_state_dict(4), crnn.build_decoder(6)
enc = CrnnEncoder(Cnn14Encoder(sample_rate=32000),
                  RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True, hidden_size=256,
                             dropout=0.5, num_layers=3), freeze_cnn=True, freeze_cnn_bn=True)
model = TransformerModel(enc, TransformerDecoder(emb_dim=256, vocab_size=4981, fc_emb_dim=512, attn_emb_dim=512,
                                                 nlayers=2, dropout=0.2)).eval()
if args.decoder == "trm":
    model.load_state_dict(crnn.model_state_dict(cnn_sd, rnn_sd, dec_o), strict=True)
else:
    from oracle import bah_decoder as bd
    from audiocaption_b200.captioning.models import hf_wrapper as hw
    dsd = bd.build_state_dict(8)
    dec = hw.TemporalBahAttnDecoder(emb_dim=512, vocab_size=4981, fc_emb_dim=512, attn_emb_dim=512, rnn_type="GRU",
                                    num_layers=1, d_model=512, dropout=0.5)
    dec.load_state_dict(dsd, strict=True)
    enc.cnn.load_state_dict(cnn_sd, strict=True)
    enc.rnn.load_state_dict(rnn_sd, strict=True)
    model = hw.TemporalSeq2SeqAttnModel(enc, dec).eval()
model = model.to(dev)
lib = _lib.lib()
n_rot = 4                                                   # 4 x 82 MB of input > 126 MB L2
host = [cm.synth_wav(B, N, seed=i, sample_rate=32000)[0].pin_memory() for i in range(n_rot)]
devb = [h.to(dev) for h in host]
lens = torch.full((B,), N, dtype=torch.long)
base = {"wav_len": lens, "specaug": False, "mode": "inference", "sample_method": "beam" if args.beam else "greedy",
        "max_length": MAX_LEN, "temp": 1.0}
if args.beam:
    base["beam_size"] = args.beam
if args.decoder == "tempgru":
    base["temporal_tag"] = torch.arange(B) % 4


hf_model = None
if args.sed:
    from oracle import sed as sed_o
    hf_model = hw.Cnn14RnnTempAttnGruModel().eval()
    sd = {f"cap_model.encoder.cnn.{k}": v for k, v in cnn_sd.items()}
    sd.update({f"cap_model.encoder.rnn.{k}": v for k, v in rnn_sd.items()})
    sd.update({f"cap_model.decoder.{k}": v for k, v in dsd.items()})
    sd.update({f"sed_model.{k}": v for k, v in sed_o.build_state_dict(12).items()})
    sd.update({k: v for k, v in cnn_sd.items() if k.startswith("melspec")})
    hf_model.load_state_dict(sd, strict=True)
    hf_model = hf_model.to(dev)


def step_resident(i):
    if hf_model is not None:
        return hf_model(devb[i % n_rot], lens, sample_method="beam" if args.beam else "greedy", beam_size=args.beam or 3, max_length=MAX_LEN)
    return model(dict(base, wav=devb[i % n_rot], need_logit=False, _device_seq=True))["seq"]


def step_e2e(i):
    if hf_model is not None:
        return hf_model(host[i % n_rot], lens, sample_method="beam" if args.beam else "greedy", beam_size=args.beam or 3, max_length=MAX_LEN)
    return model(dict(base, wav=host[i % n_rot].to(dev, non_blocking=True), need_logit=False))["seq"]


def timed(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    l0 = lib.ac_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (lib.ac_launch_count() - l0) // steps


ms, launches = timed(step_resident, args.steps, args.warmup)
ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
lib.ac_timing_enable(1)
n_prof = min(args.steps, 3)
for i in range(n_prof):
    step_resident(i)
rep = _lib.timing_report()
lib.ac_timing_enable(0)
per = {k: v[1] / n_prof for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])}
chans = (1, 64, 128, 256, 512, 1024, 2048)
H, W, flops = 1001, 64, 0.0
for i in range(6):
    flops += 2.0 * H * W * 9 * (chans[i] * chans[i + 1] + chans[i + 1] * chans[i + 1]) if i else 2.0 * H * W * 9 * chans[1] * chans[1]
    if i < 5:
        H, W = H // 2, W // 2
if getattr(args, "sed", False):
    # the SED tagger's convolutions run on the same kernel and are inside conv_ms: CNN8 blocks 64-128-256-512, time pooled
    # x4 only (hf_wrapper.py:1791-1859) = 33.1 GFLOP per clip
    sH, sW, sch = 1001, 64, (1, 64, 128, 256, 512)
    for i in range(4):
        flops += 2.0 * sH * sW * 9 * (sch[i] * sch[i + 1] + sch[i + 1] * sch[i + 1])
        sH, sW = (sH // 2, sW // 2) if i < 2 else (sH, sW // 2)
conv_ms = sum(v for k, v in per.items() if k.startswith("conv3x3"))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
bf16 = float(peaks.get("bf16_tflops", peaks.get("bf16_dense_tflops", 1593.5)))
tf32_peak = bf16 / 2
achieved = 3 * flops * B / (conv_ms * 1e-3) / 1e12          # tf32 tensor-pipe work actually issued (3 MMAs per product)
out = {
    "metric": "clips/sec (10s@32kHz clips) Cnn14Rnn-%s %s inference" % ("Trm" if args.decoder == "trm" else "TempAttnGru", "beam-%d" % args.beam if args.beam else "greedy"), "value": B / (ms / 1e3), "unit": "clips/s",
    "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "f32",
    "data": "synthetic", "gpu_launches": launches,
    "config": {"workload": "Cnn14Rnn-%s %s inference, batch=%d x 10 s @ 32 kHz synthetic clips" % ("Transformer (cnn14rnn_trm.yaml)" if args.decoder == "trm" else ("TempAttnGru (HF forward incl. SED tagger)" if args.sed else "TempAttnGru (HF config, temporal tags given)"), "beam-%d" % args.beam if args.beam else "greedy", B),
               "l2": "rotating %d input batches (%d MB)" % (n_rot, n_rot * B * N * 4 // 2 ** 20)},
    "e2e": {"value": B / (ms_e2e / 1e3), "unit": "clips/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * N * 4,
            "d2h_bytes_per_step": B * MAX_LEN * 8, "api": "TransformerModel.forward(input_dict), pinned host wav"},
    "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel<conv3x3> (11 launches)", "achieved": achieved, "peak": tf32_peak,
                 "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": None, "ms_per_step": conv_ms,
                 "share_of_step": conv_ms / sum(per.values()),
                 "fp32_equivalent_tflops": flops * B / (conv_ms * 1e-3) / 1e12,
                 "peak_source": "MEASURED_PEAKS.json dense bf16 / 2 (kind::tf32 runs at half the bf16 rate)",
                 "kernel_ms_per_step": {k: round(v, 4) for k, v in per.items()}},
}
if not args.no_cpu and args.decoder == "trm" and not args.beam:
    torch.set_num_threads(os.cpu_count())
    w1, l1 = cm.synth_wav(1, N, seed=0, sample_rate=32000)
    crnn.caption(cnn_sd, rnn_sd, dec_o, w1[:, :32000], torch.tensor([32000]))       # warm-up
    t0 = time.perf_counter()
    crnn.caption(cnn_sd, rnn_sd, dec_o, w1, l1)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "1 clip of the batch (%.1f s of CPU work), oracle port of the reference CPU path" % dt}
print(json.dumps(out))
if args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
