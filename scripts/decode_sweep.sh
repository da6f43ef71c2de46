#!/bin/bash
# greedy decode time for several (CTAs per cluster, clips per cluster) shapes
for cfg in 2,1 4,1 4,2 8,2 8,4 2,2 1,1; do
  AC_GREEDY=$cfg timeout 300 python - <<PY
import os, sys, warnings; warnings.filterwarnings("ignore"); sys.path.insert(0, ".")
import torch, bench
from audiocaption_b200 import _lib
from oracle import caption_model as cm
orc, model = bench.build_models(torch.device("cuda", 0))
dec = model.model.model.decoder
attn = torch.randn(64, 32, 1408, device="cuda"); lens = torch.full((64,), 31, dtype=torch.long)
ref = None
for i in range(3): out = dec.greedy(attn, lens, 20, 1, 2, 0, need_logit=False)["seq"]
_lib.lib().ac_timing_enable(1)
for i in range(5): out = dec.greedy(attn, lens, 20, 1, 2, 0, need_logit=False)["seq"]
rep = _lib.timing_report()
print("AC_GREEDY=$cfg", {k: round(v[1]/v[0], 3) for k, v in rep.items() if "greedy" in k}, "checksum", int(out.sum()))
PY
done
