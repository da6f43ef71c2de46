"""Per-step time of the bi-GRU recurrence kernel (3 layers x 31 steps) at 16 and 64 clips."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiocaption_b200 import _lib
from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
from oracle import crnn
sd = crnn.build_gru_state_dict(21)
m = RnnEncoder(-1, 2048, 2048, bidirectional=True, hidden_size=256, dropout=0.5, num_layers=3).eval()
m.load_state_dict(sd, strict=True); m = m.to("cuda:0")
lib = _lib.lib()
for B in (16, 64):
    x = torch.randn(B, 31, 2048, device="cuda:0").abs(); lens = torch.full((B,), 31)
    for _ in range(2): m({"attn": x, "attn_len": lens})
    lib.ac_timing_enable(1)
    for _ in range(5): m({"attn": x, "attn_len": lens})
    rep = _lib.timing_report(); lib.ac_timing_enable(0)
    print("B", B, {k: round(v[1] / v[0] * 1000 / 31, 2) for k, v in rep.items() if k.startswith("bigru")}, "us per step")
