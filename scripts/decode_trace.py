"""Phase timeline (clock64) of one greedy decode step of CTA 0.  usage: python scripts/decode_trace.py"""
import os, sys, warnings; warnings.filterwarnings("ignore"); sys.path.insert(0, ".")
import numpy as np, torch, bench
from audiocaption_b200 import _lib
orc, model = bench.build_models(torch.device("cuda", 0))
dec = model.model.model.decoder
attn = torch.randn(64, 32, 1408, device="cuda"); lens = torch.full((64,), 31, dtype=torch.long)
for i in range(2): dec.greedy(attn, lens, 20, 1, 2, 0, need_logit=False)
l = _lib.lib(); l.ac_trm_trace(1, None)
dec.greedy(attn, lens, 20, 1, 2, 0, need_logit=False)
buf = np.zeros(64, dtype=np.int64); l.ac_trm_trace(0, buf.ctypes.data)
t0 = buf[0]
names = {0: "step start", 40: "classifier start", 41: "classifier done", 42: "argmax done"}
for l_ in range(2):
    for k, n in [(1, "sa_in start"), (2, "sa_in gemv done"), (3, "sa_in synced"), (4, "self-attn done"), (5, "sa_out done"), (6, "LN1 done"),
                 (7, "ca_q done"), (8, "cross-attn done"), (9, "ca_out done"), (10, "LN2 done"), (11, "ff1 done"), (12, "ff2 done")]:
        names[k + 12 * l_] = f"L{l_} {n}"
prev = t0
for i in sorted(names):
    if buf[i]:
        print(f"{names[i]:24s} {int(buf[i]-t0):8d} (+{int(buf[i]-prev)})"); prev = buf[i]
