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
    """Make ``import captioning.models...`` resolve to the B200 mirrors, so the reference's YAML ``type:`` strings
    (captioning/utils/train_util.py:63-68 resolves them with importlib) work unchanged.

    Every mirror module is imported under its real name first (their relative imports reach up to this package) and
    then registered under the ``captioning.*`` alias; importing them lazily under the alias would re-execute the files
    as a different top-level package.  Returns the list of aliased module names."""
    import importlib
    import pkgutil
    import sys
    from . import captioning
    prefix = captioning.__name__ + "."
    names = [captioning.__name__] + [m.name for m in pkgutil.walk_packages(captioning.__path__, prefix)]
    aliased = []
    for name in names:
        mod = importlib.import_module(name)
        alias = name[len(__name__) + 1:]
        other = sys.modules.get(alias)
        if other is not None and other is not mod:
            raise ImportError(f"install_as_captioning: a different module named {alias!r} is already imported "
                              f"({getattr(other, '__file__', '?')}); install the alias before importing the reference")
        sys.modules[alias] = mod
        aliased.append(alias)
    return aliased
