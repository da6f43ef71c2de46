"""Oracle restatement of the log-mel front-end.  TEST INFRASTRUCTURE (oracle/__init__.py).

Follows the reference's torchaudio call sites:

* EffB2 (16 kHz): ``MelSpectrogram(sample_rate=16000, n_fft=512, win_length=512,
  hop_length=160, f_min=0, n_mels=64)`` + ``AmplitudeToDB(top_db=120)``
  -- captioning/models/hf_wrapper.py:269-279, applied at :292-293.
* Cnn14 (32 kHz): ``MelSpectrogram(sample_rate=32000, n_fft=1024, win_length=1024,
  hop_length=320, f_min=50, f_max=14000, n_mels=64, norm="slaney",
  mel_scale="slaney")`` + ``AmplitudeToDB()`` -- captioning/models/cnn_encoder.py:338-350,
  applied at :418-419.

torchaudio semantics restated (torchaudio 2.11 ``functional.spectrogram``,
``melscale_fbanks``, ``amplitude_to_DB``): centre/reflect padding by n_fft//2,
periodic Hann window, one-sided rFFT, |X|^2, ``matmul(spec^T, fb)``,
``10*log10(clamp(x, 1e-10))``; ``top_db`` clamps at (max - top_db) where the max is
taken over the WHOLE batch for a 3-D ``[B, F, T]`` input.

Pinned in tests/test_oracle_cpu.py against torchaudio itself (installed in the image)
and against the golden vectors produced from the imported reference.
"""
import math

import numpy as np
import torch


def hann_periodic(n: int) -> torch.Tensor:
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).float()


def _hz_to_mel(f: float, scale: str) -> float:
    if scale == "htk":
        return 2595.0 * math.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    if f >= min_log_hz:
        mels = min_log_hz / f_sp + math.log(f / min_log_hz) / (math.log(6.4) / 27.0)
    return mels


def _mel_to_hz(m: torch.Tensor, scale: str) -> torch.Tensor:
    if scale == "htk":
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_mel = 1000.0 / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = m >= min_log_mel
    freqs[log_t] = 1000.0 * torch.exp(logstep * (m[log_t] - min_log_mel))
    return freqs


def mel_filterbank(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, scale="htk") -> torch.Tensor:
    """[n_freqs, n_mels] triangular filters (fp32), as torchaudio builds them."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min, scale), _hz_to_mel(f_max, scale), n_mels + 2)
    f_pts = _mel_to_hz(m_pts, scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.min(down, up), min=0.0)
    if norm == "slaney":
        enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
        fb = fb * enorm.unsqueeze(0)
    return fb


FRONTENDS = {
    # name: (sample_rate, n_fft, hop, f_min, f_max, norm, scale, top_db)
    "effb2": dict(sample_rate=16000, n_fft=512, hop=160, f_min=0.0, f_max=8000.0,
                  norm=None, scale="htk", top_db=120.0),
    "cnn14": dict(sample_rate=32000, n_fft=1024, hop=320, f_min=50.0, f_max=14000.0,
                  norm="slaney", scale="slaney", top_db=None),
}


def frontend_buffers(kind: str):
    c = FRONTENDS[kind]
    window = hann_periodic(c["n_fft"])
    fb = mel_filterbank(c["n_fft"] // 2 + 1, c["f_min"], c["f_max"], 64, c["sample_rate"],
                        c["norm"], c["scale"])
    return window, fb


def power_spectrogram(wav: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """[B, N] -> [B, n_fft/2+1, 1 + N//hop] by explicit framing + rFFT (fp32)."""
    pad = n_fft // 2
    x = torch.nn.functional.pad(wav.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = x.unfold(-1, n_fft, hop)                      # [B, T, n_fft]
    spec = torch.fft.rfft(frames * window, dim=-1)         # [B, T, F]
    return (spec.real ** 2 + spec.imag ** 2).transpose(1, 2)


def log_mel(wav: torch.Tensor, window: torch.Tensor, fb: torch.Tensor, n_fft: int, hop: int,
            top_db=None) -> torch.Tensor:
    """[B, N] fp32 -> [B, n_mels, T] dB."""
    p = power_spectrogram(wav.float(), window, n_fft, hop)
    mel = torch.matmul(p.transpose(1, 2), fb).transpose(1, 2)
    db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    if top_db is not None:
        db = torch.max(db, db.amax() - top_db)            # batch-global max (3-D input)
    return db


def log_mel_numpy_small(wav: np.ndarray, window: np.ndarray, fb: np.ndarray, n_fft: int, hop: int):
    """float64 direct-DFT restatement for tiny inputs (second opinion on the fp32 path)."""
    pad = n_fft // 2
    x = np.pad(wav.astype(np.float64), ((0, 0), (pad, pad)), mode="reflect")
    T = 1 + wav.shape[1] // hop
    k = np.arange(n_fft // 2 + 1)[:, None] * np.arange(n_fft)[None, :]
    dft = np.exp(-2j * np.pi * k / n_fft)
    out = np.empty((wav.shape[0], fb.shape[1], T))
    for b in range(wav.shape[0]):
        for t in range(T):
            fr = x[b, t * hop:t * hop + n_fft] * window.astype(np.float64)
            pw = np.abs(dft @ fr) ** 2
            out[b, :, t] = 10.0 * np.log10(np.maximum(pw @ fb.astype(np.float64), 1e-10))
    return out
