"""Kernel time of the Transformer decode (greedy / beam-3) at small batches: cluster-size sweep via AC_TRM_CLUSTER."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audiocaption_b200 import _lib
from audiocaption_b200.captioning.models.transformer_decoder import TransformerDecoder
from oracle import caption_model as cm
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "effb2_trm.npz")))
orc = cm.build_effb2_trm(int(g["seed"]), bn_stats=g["bn_stats"])
dec = TransformerDecoder(emb_dim=256, vocab_size=4981, fc_emb_dim=1408, attn_emb_dim=1408, dropout=0.2, nlayers=2, tie_weights=True).eval()
dec.load_state_dict(orc.decoder.state_dict(), strict=True)
dec = dec.to("cuda:0")
lib = _lib.lib()
for B in [int(x) for x in sys.argv[1].split(",")]:
    mem = torch.randn(B, 32, 1408, device="cuda:0"); lens = torch.full((B,), 31)
    for name, f in (("greedy", lambda: dec.greedy(mem, lens, 20, 1, 2, 0, need_logit=True)), ("beam3", lambda: dec.beam_search(mem, lens, 20, 3, 1.0, 1, 2, 0))):
        for _ in range(2): f()
        lib.ac_timing_enable(1)
        for _ in range(5): f()
        rep = _lib.timing_report(); lib.ac_timing_enable(0)
        print(f"P={os.environ.get('AC_TRM_CLUSTER', 'auto')} B={B} {name}:", {k: round(v[1] / v[0], 3) for k, v in rep.items() if k.startswith("trm")}, flush=True)
