"""Datasets for the entry points.  The reference reads fp16 waveforms from HDF5 (captioning/datasets/caption_dataset.py:
20-36,131-145; `h5py` is not part of this image) -- here: waveform files (`.wav` PCM through the standard library, `.npy`,
`.pt`) and a seeded synthetic caption dataset for smoke runs and benchmarks."""
import wave
from pathlib import Path

import numpy as np
import torch


def load_waveform(path):
    """-> (float32 mono waveform [N], sample_rate or None when the container does not carry one)"""
    path = str(path)
    suffix = Path(path).suffix.lower()
    if suffix == ".npy":
        return np.load(path).astype(np.float32).reshape(-1), None
    if suffix in (".pt", ".pth"):
        return torch.load(path).float().reshape(-1).numpy(), None
    if suffix == ".wav":
        with wave.open(path, "rb") as f:
            sr, n_ch, width = f.getframerate(), f.getnchannels(), f.getsampwidth()
            raw = f.readframes(f.getnframes())
        if width != 2:
            raise ValueError(f"{path}: only 16-bit PCM wav files are supported (sample width {width})")
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        return x.reshape(-1, n_ch).mean(1), sr
    raise ValueError(f"{path}: unsupported audio container (wav / npy / pt); HDF5 needs h5py, which this image lacks")


class InferenceDataset(torch.utils.data.Dataset):
    """aid -> file name (python_scripts/inference/inference.py:47-82); resampling happens later, on the device."""

    def __init__(self, aid_to_fname):
        self.aid_to_fname = dict(aid_to_fname)
        self.aids = list(self.aid_to_fname)

    def __len__(self):
        return len(self.aids)

    def __getitem__(self, index):
        aid = self.aids[index]
        wav, sr = load_waveform(self.aid_to_fname[aid])
        return {"audio_id": aid, "wav": wav, "sample_rate": sr}


class SyntheticCaptionDataset(torch.utils.data.Dataset):
    """Seeded random clips (0.1 * randn) with random captions over a synthetic vocabulary `w4 .. w{V-1}`."""

    def __init__(self, size=64, n_samples=320000, vocab_size=4368, min_words=6, max_words=20, seed=0, ragged=False):
        self.size, self.n_samples, self.vocab_size = size, n_samples, vocab_size
        self.min_words, self.max_words, self.seed, self.ragged = min_words, max_words, seed, ragged

    def __len__(self):
        return self.size

    def vocabulary(self):
        return [f"w{i}" for i in range(4, self.vocab_size)]

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + index)
        n = self.n_samples
        if self.ragged:
            n = int(torch.randint(self.n_samples // 2, self.n_samples + 1, (1,), generator=g))
        wav = 0.1 * torch.randn(n, generator=g)
        n_words = int(torch.randint(self.min_words, self.max_words + 1, (1,), generator=g))
        words = torch.randint(4, self.vocab_size, (n_words,), generator=g).tolist()
        return {"audio_id": f"synthetic_{index}", "wav": wav.numpy(), "caption": " ".join(f"w{w}" for w in words)}
