"""Host-side cost of ONE training step (enqueue only): cProfile over a few steps, each started on an idle GPU.
usage: python scripts/train_host_profile.py [n_steps]"""
import cProfile, os, pstats, random, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from audiocaption_b200.train_step import TrainStep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model = bench.build_train_model(dev)
model.encoder.cnn.conv_precision = "bf16"
step = TrainStep(model, total_iters=10 ** 9, lr=5e-4, warmup_iters=3000)
host = bench.train_batches(0, 2)
devb = [dict(b, wav=b["wav"].to(dev), cap=b["cap"].to(dev)) for b in host]
random.seed(1)
step.ss_ratio = 1.0
for i in range(5):
    step.step(devb[i % 2])
torch.cuda.synchronize()
pr = cProfile.Profile()
tot = 0.0
for i in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr.enable()
    step.step(devb[i % 2])
    pr.disable()
    tot += time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host enqueue: {tot / n * 1e3:.3f} ms per step (profiler on)")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
