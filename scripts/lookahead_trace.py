"""Timeline of the look-ahead training step: CUDA events around the encoder pass (look-ahead stream) and the trainable part
(high-priority stream) of a few steps, relative to one origin.  usage: python scripts/lookahead_trace.py [cnn_sms]"""
import os, random, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from audiocaption_b200.train_step import TrainStep

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model = bench.build_train_model(dev)
model.encoder.cnn.conv_precision = "bf16"
step = TrainStep(model, total_iters=10 ** 9, lr=5e-4, warmup_iters=3000)
if len(sys.argv) > 1:
    step.cnn_sms = int(sys.argv[1])
host = bench.train_batches(0, 4)
devb = [dict(b, wav=b["wav"].to(dev), cap=b["cap"].to(dev)) for b in host]
random.seed(1)
step.ss_ratio = 1.0
staged = step.prefetch(devb[0])
for i in range(6):
    nxt = step.prefetch(devb[(i + 1) % 4])
    step.step(staged)
    staged = nxt
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
origin = ev(); origin.record()
marks = []
for i in range(6, 12):
    c0, c1, s0, s1 = ev(), ev(), ev(), ev()
    c0.record(step._cnn_stream)
    nxt = step.prefetch(devb[(i + 1) % 4])
    c1.record(step._cnn_stream)
    s0.record(step._hi_stream)
    step.step(staged)
    s1.record(step._hi_stream)
    staged = nxt
    marks.append((c0, c1, s0, s1))
torch.cuda.synchronize()
print("partition:", getattr(step, "partition", None) is not None, getattr(step, "partition_error", None), "cnn sms", step.cnn_sms,
      "trainable sms", getattr(step, "train_sms", None))
print(f"period {(origin.elapsed_time(marks[-1][3]) - origin.elapsed_time(marks[1][3])) / (len(marks) - 2):.3f} ms per step")
for i, (c0, c1, s0, s1) in enumerate(marks[:3]):
    print(f"iter {i}: cnn(i+1) [{origin.elapsed_time(c0):7.3f} .. {origin.elapsed_time(c1):7.3f}] ms   "
          f"trainable(i) [{origin.elapsed_time(s0):7.3f} .. {origin.elapsed_time(s1):7.3f}] ms")
