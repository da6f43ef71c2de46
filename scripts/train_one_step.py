"""One warm training step + N profiled steps of the training workload (for ncu launch lists / captures).
usage: python scripts/train_one_step.py [n_steps] [precision fp32|tf32|bf16] [ss_ratio]"""
import os, random, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from audiocaption_b200.train_step import TrainStep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
ss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.85
dev = torch.device("cuda", 0)
model = bench.build_train_model(dev)
model.encoder.cnn.conv_precision = prec
step = TrainStep(model, total_iters=10 ** 9, lr=5e-4, warmup_iters=3000)
batch = bench.train_batches(0, 1)[0]
batch = dict(batch, wav=batch["wav"].to(dev), cap=batch["cap"].to(dev))
torch.manual_seed(1); random.seed(1)
step.ss_ratio = ss
L = batch["cap"].shape[1] - 1
coins = [t not in (3, 11) for t in range(L)]          # fixed coin pattern: two sampled steps (both token rows run)
for _ in range(1 + n):
    loss = step.step(batch, coins=coins)["loss"]
torch.cuda.synchronize()
print("done loss", loss.item())
