"""Host-side enqueue time of one step (no synchronisation) vs device time."""
import os, sys, time, warnings; warnings.filterwarnings("ignore"); sys.path.insert(0, ".")
import torch, bench
from oracle import caption_model as cm
orc, model = bench.build_models(torch.device("cuda", 0))
enc, dec = model.model.model.encoder, model.model.model.decoder
wav, lens = cm.synth_wav(64, bench.N_SAMPLES, seed=0); wd = wav.cuda()
def step():
    e = enc({"wav": wd, "wav_len": lens, "specaug": False})
    return dec.greedy(e["attn_emb"], e["attn_emb_len"], 20, 1, 2, 0, need_logit=False)["seq"]
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1000*(t1-t0)/20:.3f} ms/step, total {1000*(t2-t0)/20:.3f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
