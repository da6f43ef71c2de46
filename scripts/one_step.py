"""One warm step + N profiled steps of the headline workload (for ncu launch lists / captures).
usage: python scripts/one_step.py [n_steps] [batch]"""
import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import caption_model as cm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
orc, model = bench.build_models(torch.device("cuda", 0))
wav, lens = cm.synth_wav(batch, bench.N_SAMPLES, seed=0)
wd = wav.cuda()
enc, dec = model.model.model.encoder, model.model.model.decoder
for _ in range(1 + n):
    e = enc({"wav": wd, "wav_len": lens, "specaug": False})
    seq = dec.greedy(e["attn_emb"], e["attn_emb_len"], bench.MAX_LEN, 1, 2, 0, need_logit=False)["seq"]
torch.cuda.synchronize()
print("done", seq[0].tolist())
