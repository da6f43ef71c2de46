"""GPU check of the tcgen05 3x3 convolution (ac_conv3x3) against a float64 torch reference + timing on the
Cnn14 layer shapes.   usage: python scripts/conv_check.py [quick|shapes]"""
import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from audiocaption_b200 import _lib

lib = _lib.lib()
dev = "cuda:0"


def run(B, H, W, Cin, Cout, act=2, seed=0, check=True, reps=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev)
    sc = (torch.rand(Cout, generator=g) + 0.5).to(dev)
    bi = torch.randn(Cout, generator=g).to(dev)
    out = torch.full((B, H, W, Cout), float("nan"), device=dev)
    st = _lib.current_stream()
    passes = int(os.environ.get("CONV_PASSES", "3"))       # 1 = plain TF32 mode
    call = lambda: _lib.check(lib.ac_conv3x3_p(_lib.ptr(x), _lib.ptr(w), _lib.ptr(sc), _lib.ptr(bi), _lib.ptr(out), B, H, W,
                                               Cin, Cout, act, passes, st), "ac_conv3x3")
    call()
    torch.cuda.synchronize()
    err = None
    if check:
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=1) * sc.double().view(1, -1, 1, 1) + bi.double().view(1, -1, 1, 1)
        if act == 2:
            ref = ref.clamp_min(0)
        ref = ref.permute(0, 2, 3, 1)
        assert not torch.isnan(out).any(), "NaN / unwritten output"
        err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    ms = None
    if reps:
        lib.ac_timing_enable(1)
        for _ in range(reps):
            call()
        rep = _lib.timing_report()
        lib.ac_timing_enable(0)
        ms = min(v[1] / v[0] for k, v in rep.items() if k.startswith("conv3x3"))
    return err, ms


mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
print("device", torch.cuda.get_device_name(0))
if mode in ("quick", "all"):
    for c in [dict(B=1, H=4, W=32, Cin=32, Cout=32), dict(B=2, H=9, W=64, Cin=64, Cout=64), dict(B=3, H=31, W=2, Cin=64, Cout=128),
              dict(B=2, H=62, W=4, Cin=96, Cout=160, act=0), dict(B=2, H=125, W=8, Cin=64, Cout=256),
              dict(B=1, H=21, W=16, Cin=32, Cout=96), dict(B=5, H=7, W=2, Cin=128, Cout=64)]:
        err, _ = run(**c)
        print(c, f"rel err {err:.3e}", "OK" if err < 2e-5 else "BAD", flush=True)
if mode in ("shapes", "all"):
    B = 64
    tot = 0.0
    flops = 0.0
    for (H, W, Cin, Cout) in [(1001, 64, 64, 64), (500, 32, 64, 128), (500, 32, 128, 128), (250, 16, 128, 256), (250, 16, 256, 256),
                              (125, 8, 256, 512), (125, 8, 512, 512), (62, 4, 512, 1024), (62, 4, 1024, 1024), (31, 2, 1024, 2048),
                              (31, 2, 2048, 2048)]:
        _, ms = run(B, H, W, Cin, Cout, check=False, reps=2)
        fl = 2.0 * B * H * W * 9 * Cin * Cout
        tot += ms; flops += fl
        print(f"H={H:5d} W={W:3d} Cin={Cin:5d} Cout={Cout:5d}: {ms:8.3f} ms  {fl/ms/1e9:8.1f} TFLOP/s fp32-equivalent "
              f"({3*fl/ms/1e9:7.1f} tf32)", flush=True)
    print(f"total {tot:.2f} ms for {B} clips = {B/tot*1e3:.0f} clips/s, {flops/tot/1e9:.1f} TFLOP/s fp32-equivalent")
if mode == "one":
    B, H, W, Cin, Cout = [int(x) for x in sys.argv[2:7]]
    run(B, H, W, Cin, Cout, check=False, reps=2)
    print("one done")
if mode == "trace":
    import numpy as np
    B, H, W, Cin, Cout = [int(x) for x in sys.argv[2:7]]
    run(B, H, W, Cin, Cout, check=False)
    lib.ac_gemm_trace(1, None)
    run(B, H, W, Cin, Cout, check=False)
    buf = np.zeros((8, 256), dtype=np.int64)
    lib.ac_gemm_trace(0, buf.ctypes.data)
    t0 = buf[buf > 0].min()
    n = int((buf[3] > 0).sum())
    print(f"conv trace B={B} H={H} W={W} Cin={Cin} Cout={Cout}: {n} chunks")
    prev_commit = None
    for i in range(min(n, 60)):
        st, iss, com = int(buf[3][i] - t0), int(buf[7][i] - t0), int(buf[4][i] - t0)
        gap = st - prev_commit if prev_commit is not None else 0
        print(f"{i:3d} landed={int(buf[1][i]-t0):7d} xf_done={int(buf[2][i]-t0):7d} mma_start={st:7d} issue={iss-st:5d} commits={com-iss:5d} wait_gap={gap:5d}")
        prev_commit = com
