"""configs[4] leg of bench.py on its own (TempGRU beam-4 + SED, 16 clips): python scripts/c5_time.py [steps]"""
import json, os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from audiocaption_b200 import _lib

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
lib = _lib.lib()


def timed(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    l0 = lib.ac_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, lib.ac_launch_count() - l0


rec = bench.run_config5_leg(dev, 0, 1, timed, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
print(json.dumps({k: (round(rec[k]["ms_per_step"], 3), round(rec[k]["e2e"]["ms_per_step"], 3)) for k in ("bf16", "tf32", "fp32")}))
