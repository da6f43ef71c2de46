"""audiocaption_b200 -- B200 (sm_100a) implementation of the AudioCaption hot path.

waveform -> log-mel -> EfficientNet-B2 -> Transformer caption decoder (greedy / beam), as
hand-written CUDA kernels behind a C ABI (include/audiocaption_b200.h), with host-side
mirrors of the reference's module API under ``audiocaption_b200.captioning`` (same class
names, constructor kwargs, ``forward(input_dict) -> dict`` contract and ``state_dict`` keys).
"""
from . import _lib  # noqa: F401
from ._lib import AudioCaptionB200Error, lib  # noqa: F401

__version__ = "0.1.0"


def install_as_captioning():
    """Make ``import captioning.models...`` resolve to the B200 mirrors, so the reference's
    YAML ``type:`` strings (captioning/utils/train_util.py:63-68) work unchanged."""
    import sys
    from . import captioning
    sys.modules.setdefault("captioning", captioning)
    for name, mod in list(sys.modules.items()):
        if name.startswith(__name__ + ".captioning."):
            sys.modules.setdefault(name[len(__name__) + 1:], mod)
