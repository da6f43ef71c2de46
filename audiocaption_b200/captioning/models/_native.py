"""Shared plumbing between the mirror modules and the C ABI."""
import torch

from ... import _lib


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.AudioCaptionB200Error(
            f"{what}: tensor is on {t.device}; audiocaption_b200 runs on CUDA only (no CPU fallback)")


class Workspace:
    """Grow-only device scratch buffer owned by the calling module (the caller owns all
    buffers; the library never allocates per call)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return self.buf


def params_signature(tensors):
    """Changes whenever a parameter is re-assigned, moved or written in place."""
    return tuple((t.data_ptr(), t._version) for t in tensors)
