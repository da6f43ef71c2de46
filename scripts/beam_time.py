"""Transformer beam-search decode time by beam size and batch: head-split kernel (default) vs AC_BEAM_HEADS=0 (column split).
usage: [AC_BEAM_HEADS=0|1|2] python scripts/beam_time.py"""
import os, sys, warnings; warnings.filterwarnings("ignore"); sys.path.insert(0, ".")
import torch, bench
from audiocaption_b200 import _lib
orc, model = bench.build_models(torch.device("cuda", 0))
dec = model.model.model.decoder
for B in (64, 16, 1):
    attn = torch.randn(B, 32, 1408, device="cuda"); lens = torch.full((B,), 31, dtype=torch.long)
    for beam in (2, 3, 5):
        for i in range(2): dec.beam_search(attn, lens, 20, beam, 1.0, 1, 2, 0)
        _lib.lib().ac_timing_enable(1)
        for i in range(3): out = dec.beam_search(attn, lens, 20, beam, 1.0, 1, 2, 0)["seq"]
        rep = _lib.timing_report(); _lib.lib().ac_timing_enable(0)
        print("heads", os.environ.get("AC_BEAM_HEADS", "default"), "B", B, "beam", beam,
              {k: round(v[1]/v[0], 3) for k, v in rep.items() if "trm" in k}, "checksum", int(out.sum()))
