#!/bin/bash
# greedy decode time at 64 clips: head-split kernel (G = 1, 2, 4) against the column-split kernel (AC_GREEDY=2,1)
for cfg in "AC_GREEDY_HEADS=2" "AC_GREEDY_HEADS=1" "AC_GREEDY_HEADS=4" "AC_GREEDY=2,1"; do
  env $cfg timeout 300 python - <<PY
import os, sys, warnings; warnings.filterwarnings("ignore"); sys.path.insert(0, ".")
import torch, bench
from audiocaption_b200 import _lib
orc, model = bench.build_models(torch.device("cuda", 0))
dec = model.model.model.decoder
for B in (64, 16, 1):
    attn = torch.randn(B, 32, 1408, device="cuda"); lens = torch.full((B,), 31, dtype=torch.long)
    for i in range(3): out = dec.greedy(attn, lens, 20, 1, 2, 0, need_logit=False)["seq"]
    _lib.lib().ac_timing_enable(1)
    for i in range(5): out = dec.greedy(attn, lens, 20, 1, 2, 0, need_logit=False)["seq"]
    rep = _lib.timing_report(); _lib.lib().ac_timing_enable(0)
    print("$cfg", "B", B, {k: round(v[1]/v[0], 3) for k, v in rep.items() if "greedy" in k}, "checksum", int(out.sum()))
PY
done
