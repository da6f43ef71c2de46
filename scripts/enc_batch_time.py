"""EfficientNet-B2 encoder kernel time by family at small batch sizes (per-launch CUDA events).
usage: [AC_TC_NARROW=n] python scripts/enc_batch_time.py [batch ...]"""
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
for B in [int(x) for x in sys.argv[1:]] or [1, 4, 16]:
    wav = cm.synth_wav(B, 160000, seed=1)[0].to(dev)
    lens = torch.full((B,), 160000, dtype=torch.long)
    for i in range(3):
        enc({"wav": wav, "wav_len": lens, "specaug": False})
    torch.cuda.synchronize()
    lib.ac_timing_enable(1)
    for i in range(5):
        enc({"wav": wav, "wav_len": lens, "specaug": False})
    rep = _lib.timing_report()
    lib.ac_timing_enable(0)
    fam = {k: round(ms / 5, 3) for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    print(json.dumps({"narrow": os.environ.get("AC_TC_NARROW", "default"), "batch": B,
                      "kernel_sum_ms": round(sum(ms for _, ms in rep.values()) / 5, 3), "families": fam}))
