"""On-device replacement of `torchaudio.functional.resample` for the input pipeline (demo.py:36, inference.py:37,
captioning/datasets/caption_dataset.py:110-120): 32 kHz -> 16 kHz in front of the EfficientNet-B2 captioner, 44.1 kHz ->
32 kHz for Clotho.  The polyphase coefficient table is built once per (orig, new) pair with the closed form of
torchaudio's `_get_sinc_resample_kernel` (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99); the FIR itself runs in
csrc/resample.cu."""
import math

import torch

from . import _lib

_TABLES = {}


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """-> (coef [new, taps] float32 (CPU), orig, new, width) with orig / new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64) / orig
    t = (torch.arange(0, -new, -1, dtype=torch.float64)[:, None] / new + idx[None, :]) * base_freq
    t = t.clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    coef = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * window * scale
    return coef.to(torch.float32).contiguous(), orig, new, width


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """waveform [..., N] fp32 on a CUDA device -> [..., ceil(N * new / orig)] (same device, asynchronous on the current
    stream).  Equal sample rates return the input unchanged, as torchaudio does."""
    if int(orig_freq) == int(new_freq):
        return waveform
    if not waveform.is_cuda:
        raise _lib.AudioCaptionB200Error("resample: the waveform must live on a CUDA device (no CPU fallback)")
    key = (int(orig_freq), int(new_freq), waveform.device)
    if key not in _TABLES:
        coef, orig, new, width = sinc_resample_kernel(orig_freq, new_freq)
        _TABLES[key] = (coef.to(waveform.device), orig, new, width)
    coef, orig, new, width = _TABLES[key]
    shape = waveform.shape
    x = waveform.reshape(-1, shape[-1]).float().contiguous()
    l = _lib.lib()
    n_out = l.ac_resample_out_len(x.shape[1], orig, new)
    out = torch.empty(x.shape[0], n_out, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        for b0 in range(0, x.shape[0], 65535):
            xb, ob = x[b0:b0 + 65535], out[b0:b0 + 65535]
            _lib.check(l.ac_resample(_lib.ptr(xb), xb.shape[0], x.shape[1], _lib.ptr(coef), orig, new, coef.shape[1], width,
                                     _lib.ptr(ob), _lib.current_stream()), "ac_resample")
    return out.reshape(*shape[:-1], n_out)
