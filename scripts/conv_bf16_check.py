"""GPU check of the bf16 3x3 convolution (ac_conv3x3_bf16) against a float64 torch reference evaluated on the same
bf16-rounded operands, + timing on the Cnn14 layer shapes.   usage: python scripts/conv_bf16_check.py [quick|shapes]"""
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
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev)
    sc = (torch.rand(Cout, generator=g) + 0.5).to(dev)
    bi = torch.randn(Cout, generator=g).to(dev)
    out = torch.full((B, H, W, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    st = _lib.current_stream()
    call = lambda: _lib.check(lib.ac_conv3x3_bf16(x.data_ptr(), _lib.ptr(w), _lib.ptr(sc), _lib.ptr(bi), out.data_ptr(), B, H, W,
                                                  Cin, Cout, act, st), "ac_conv3x3_bf16")
    call()
    torch.cuda.synchronize()
    err = None
    if check:
        wq = (w * sc.view(-1, 1, 1, 1)).to(torch.bfloat16).double()
        ref = F.conv2d(x.double().permute(0, 3, 1, 2), wq, padding=1) + bi.double().view(1, -1, 1, 1)
        if act == 2:
            ref = ref.clamp_min(0)
        ref = ref.permute(0, 2, 3, 1)
        assert not torch.isnan(out.float()).any(), "NaN / unwritten output"
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
    for c in [dict(B=1, H=4, W=32, Cin=64, Cout=64), dict(B=2, H=9, W=64, Cin=64, Cout=128), dict(B=3, H=31, W=2, Cin=128, Cout=256),
              dict(B=2, H=62, W=4, Cin=192, Cout=128, act=0), dict(B=2, H=125, W=8, Cin=64, Cout=256),
              dict(B=5, H=7, W=2, Cin=128, Cout=64), dict(B=70, H=31, W=2, Cin=256, Cout=512)]:
        err, _ = run(**c)
        print(c, f"rel err {err:.3e}", "OK" if err < 6e-3 else "BAD", flush=True)      # output rounding to bf16: 2^-9 = 2e-3 relative
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
        print(f"H={H:5d} W={W:3d} Cin={Cin:5d} Cout={Cout:5d}: {ms:8.3f} ms  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    print(f"total {tot:.2f} ms for {B} clips = {B/tot*1e3:.0f} clips/s, {flops/tot/1e9:.1f} TFLOP/s bf16")
