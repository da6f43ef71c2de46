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


def to_device_async(t: torch.Tensor, device, dtype=None) -> torch.Tensor:
    """Small host tensor -> device without synchronising the stream: a pageable `.to(device)` makes torch wait for
    everything queued so far (the host then cannot run ahead of the GPU); staging through pinned memory (torch's
    caching host allocator makes this cheap) keeps the copy asynchronous."""
    if t.is_cuda:
        return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous().pin_memory().to(device, non_blocking=True)
