"""Sweep of the L2-resident chunked schedule of the EfficientNet-B2 head (AC_EFFB2_CHUNK="blocks,clips"): per setting the
encoder's wall time per 64-clip batch (CUDA events over 20 batches) and the sum of its kernels' own times (per-launch
events), so host-enqueue limits are visible separately from device time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from audiocaption_b200 import _lib
from oracle import caption_model as cm

dev = torch.device("cuda", 0)
orc, model = bench.build_models(dev)
enc = model.model.model.encoder
lib = _lib.lib()
n_rot = 8
devb = [cm.synth_wav(64, 160000, seed=i)[0].to(dev) for i in range(n_rot)]
lens = torch.full((64,), 160000, dtype=torch.long)
ref = None
for setting in sys.argv[1:] or ["0,0", "5,8", "5,16", "5,4", "3,8", "9,8", "9,16"]:
    os.environ["AC_EFFB2_CHUNK"] = setting
    for i in range(3):
        out = enc({"wav": devb[i % n_rot], "wav_len": lens, "specaug": False})["attn_emb"]
    torch.cuda.synchronize()
    if ref is None:
        ref = enc({"wav": devb[0], "wav_len": lens, "specaug": False})["attn_emb"].clone()
    same = (enc({"wav": devb[0], "wav_len": lens, "specaug": False})["attn_emb"] == ref).all().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        enc({"wav": devb[i % n_rot], "wav_len": lens, "specaug": False})
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / 20
    lib.ac_timing_enable(1)
    for i in range(5):
        enc({"wav": devb[i % n_rot], "wav_len": lens, "specaug": False})
    rep = _lib.timing_report()
    lib.ac_timing_enable(0)
    ksum = sum(ms for _, ms in rep.values()) / 5
    fam = {k: round(ms / 5, 3) for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    print(json.dumps({"chunk": setting, "identical": same, "wall_ms": round(wall, 3), "kernel_sum_ms": round(ksum, 3), "families": fam}))
