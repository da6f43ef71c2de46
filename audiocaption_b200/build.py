"""Builds the in-tree CUDA library (sm_100a) with nvcc.  `python -m audiocaption_b200.build`.

Every csrc/*.cu is compiled to its own object (in parallel, only when stale) and the objects are linked into
lib/libaudiocaption_b200.so; objects live under lib/obj/ (git-ignored like the .so)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libaudiocaption_b200.so")
SOURCES = ["capi.cu", "logmel.cu", "gemm.cu", "gemm_tc.cu", "dwconv_tma.cu", "effb2.cu", "cnn14.cu", "bigru.cu",
           "trm_decode.cu", "bah_decode.cu", "train_ops.cu", "trm_train.cu", "bigru_train.cu", "resample.cu", "conv_bf16.cu", "sm_partition.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "audiocaption_b200.h")]


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, src[:-3] + ".o")


def _stale_sources():
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    out = []
    for s in SOURCES:
        o = _obj(s)
        if not os.path.exists(o) or os.path.getmtime(o) < max(hdr_t, os.path.getmtime(os.path.join(CSRC, s))):
            out.append(s)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH):       # the .so alone travels to the GPU box (objects do not)
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= t for d in _headers() + [os.path.join(CSRC, s) for s in SOURCES]):
            return LIB_PATH
    stale = list(SOURCES) if force else _stale_sources()
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_one(s):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", _obj(s), os.path.join(CSRC, s)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(len(stale), os.cpu_count() or 4) or 1) as ex:
        logs = list(ex.map(compile_one, stale))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + [_obj(s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
