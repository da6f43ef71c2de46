"""Import the REAL reference (read-only at /root/reference) with import stubs.

TEST INFRASTRUCTURE, build container only: /root/reference does not exist on the GPU
box, so nothing that runs there may import this module.  It is used by
oracle/gen_golden.py to produce tests/golden/*.npz and by the ``not gpu`` tests that
pin the oracle restatement when the reference happens to be present.

Stubs: ``efficientnet_pytorch`` -> oracle.efficientnet_b2 (the package is
absent; everything else in hf_wrapper.py is the reference's own code), ``h5py`` and
``torchlibrosa`` -> empty modules (import-time only; not executed on the hot path).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("AUDIOCAPTION_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "captioning"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    from . import efficientnet_b2 as eb
    if "efficientnet_pytorch" not in sys.modules:
        utils = _stub("efficientnet_pytorch.utils", get_model_params=eb.get_model_params)
        _stub("efficientnet_pytorch", EfficientNet=eb.EfficientNet, utils=utils)
    if "h5py" not in sys.modules:
        _stub("h5py")
    if "torchlibrosa" not in sys.modules:
        class SpecAugmentation:  # constructed at import/ctor time only; never called (specaug=False)
            def __init__(self, *a, **k):
                pass
        aug = _stub("torchlibrosa.augmentation", SpecAugmentation=SpecAugmentation)
        _stub("torchlibrosa", augmentation=aug)


def load(module: str):
    """e.g. load('captioning.models.hf_wrapper')"""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module(module)
