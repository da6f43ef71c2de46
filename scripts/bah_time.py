"""Kernel time of the GRU-attention decode (greedy / beam) on seeded memories: cluster-size and beam sweeps."""
import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audiocaption_b200 import _lib
from audiocaption_b200.captioning.models import hf_wrapper as hw
from oracle import bah_decoder as bd
dev = "cuda:0"
sd = bd.build_state_dict(8)
sd["classifier.bias"][2] = -50.0          # never end: every call runs all 20 steps
dec = hw.TemporalBahAttnDecoder(emb_dim=512, vocab_size=4981, fc_emb_dim=512, attn_emb_dim=512, rnn_type="GRU", num_layers=1,
                                d_model=512, dropout=0.5).eval()
dec.load_state_dict(sd, strict=True)
dec = dec.to(dev)
lib = _lib.lib()
for B in [int(x) for x in sys.argv[1].split(",")]:
    fc, attn, lens, tags = bd.synth_memory(3, B, 31)
    fc, attn = fc.to(dev), attn.to(dev)
    for beam in [int(x) for x in sys.argv[2].split(",")]:
        f = (lambda: dec.greedy(fc, attn, lens, tags, 20, 1, 2, need_logit=False)) if beam == 0 else \
            (lambda: dec.beam_search(fc, attn, lens, tags, 20, beam, 1.0, 1, 2))
        for _ in range(2):
            f()
        lib.ac_timing_enable(1)
        for _ in range(5):
            f()
        rep = _lib.timing_report()
        lib.ac_timing_enable(0)
        ms = {k: v[1] / v[0] for k, v in rep.items() if k.startswith("bah")}
        print(f"P={os.environ.get('AC_BAH_CLUSTER', 'auto')} B={B} beam={beam}: {ms}", flush=True)
