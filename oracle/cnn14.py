"""Oracle restatement of the Cnn14 (PANNs) audio encoder.  TEST INFRASTRUCTURE (oracle/__init__.py).

Follows captioning/models/cnn_encoder.py:
  * `ConvBlock` :32-75      -- relu(bn1(conv1(x))), relu(bn2(conv2(x))), avg_pool2d(pool_size); 3x3 convolutions,
                               stride 1, padding 1, no bias.
  * `Cnn14Encoder.__init__` :326-366 -- MelSpectrogram(32 kHz, n_fft 1024, hop 320, 50-14000 Hz, 64 slaney mels),
                               AmplitudeToDB() (no top_db), bn0 = BatchNorm2d(64) over the MEL axis, six ConvBlocks
                               1->64->128->256->512->1024->2048, fc1 = Linear(2048, 2048), downsample_ratio 32.
  * `Cnn14Encoder.forward` :414-464 -- (B, mel, T) -> transpose -> (B, 1, T, mel); bn0 applied with mel as the channel
                               axis; blocks 1-5 pool (2, 2), block 6 pool (1, 1); mean over mel; attn_emb (B, T', 2048);
                               feat_length = (wav_len // hop + 1) // 32; fc_emb = relu(fc1(max_with_lens + mean_with_lens)).
    (HF copy: captioning/models/hf_wrapper.py:1259-1304.)  Dropout is the identity in eval mode.
  * `max_with_lens` / `mean_with_lens`: captioning/utils/model_util.py:41-84.

Pinned against the imported reference class (tests/test_oracle_cpu.py, build container) and against
tests/golden/cnn14.npz produced from it (oracle/gen_golden.py).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import audio_frontend as fe

CHANNELS = (1, 64, 128, 256, 512, 1024, 2048)
HOP, DOWNSAMPLE = 320, 32


def state_dict_keys():
    """Cnn14Encoder.state_dict() keys in module order (the SpecAugmentation member holds no tensors)."""
    keys = ["melspec_extractor.spectrogram.window", "melspec_extractor.mel_scale.fb"]
    bn = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")
    keys += [f"bn0.{k}" for k in bn]
    for i in range(1, 7):
        keys += [f"conv_block{i}.conv1.weight", f"conv_block{i}.conv2.weight"]
        keys += [f"conv_block{i}.bn1.{k}" for k in bn] + [f"conv_block{i}.bn2.{k}" for k in bn]
    return keys + ["fc1.weight", "fc1.bias"]


def build_state_dict(seed: int = 3):
    """Seeded 'trained-like' weights: He-scaled convolutions (activations keep their scale through the ReLUs) and
    non-trivial BatchNorm statistics / affines (identity BN would hide folding bugs).  Deterministic on the CPU
    generator, so the GPU box rebuilds the same tensors."""
    g = torch.Generator().manual_seed(seed)
    window, fb = fe.frontend_buffers("cnn14")
    sd = {"melspec_extractor.spectrogram.window": window, "melspec_extractor.mel_scale.fb": fb}

    def bn(prefix, c, mean_scale=0.1):
        sd[prefix + ".weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = mean_scale * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.long)

    # log-mel values sit around -40 dB with a spread of tens of dB: bn0 statistics of that order
    bn("bn0", 64)
    sd["bn0.running_mean"] = -30.0 + 10.0 * torch.randn(64, generator=g)
    sd["bn0.running_var"] = 200.0 + 200.0 * torch.rand(64, generator=g)
    for i in range(1, 7):
        ci, co = CHANNELS[i - 1], CHANNELS[i]
        sd[f"conv_block{i}.conv1.weight"] = torch.randn(co, ci, 3, 3, generator=g) * math.sqrt(2.0 / (9 * ci))
        sd[f"conv_block{i}.conv2.weight"] = torch.randn(co, co, 3, 3, generator=g) * math.sqrt(2.0 / (9 * co))
        bn(f"conv_block{i}.bn1", co)
        bn(f"conv_block{i}.bn2", co)
    sd["fc1.weight"] = torch.randn(2048, 2048, generator=g) * math.sqrt(1.0 / 2048)
    sd["fc1.bias"] = 0.1 * torch.randn(2048, generator=g)
    return {k: sd[k] for k in state_dict_keys()}


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=1e-5)


def conv_block(x, sd, i, pool):
    """cnn_encoder.py:60-75 with pool_type='avg'."""
    x = F.relu(_bn(F.conv2d(x, sd[f"conv_block{i}.conv1.weight"], padding=1), sd, f"conv_block{i}.bn1"))
    x = F.relu(_bn(F.conv2d(x, sd[f"conv_block{i}.conv2.weight"], padding=1), sd, f"conv_block{i}.bn2"))
    return F.avg_pool2d(x, kernel_size=pool)


def length_mask(lens, max_len):
    return torch.arange(max_len).unsqueeze(0) < torch.as_tensor(lens).view(-1, 1)


def max_with_lens(x, lens):
    m = length_mask(lens, x.size(1))
    y = x.clone()
    y[~m] = float("-inf")
    return y.max(1)[0]


def mean_with_lens(x, lens):
    m = length_mask(lens, x.size(1)).unsqueeze(-1)
    return (x * m).sum(1) / torch.as_tensor(lens).view(-1, 1)


def log_mel(sd, wav):
    """cnn_encoder.py:418-419: db_transform(melspec_extractor(wav)) -> [B, 64, T] (no top_db clamp)."""
    c = fe.FRONTENDS["cnn14"]
    return fe.log_mel(wav, sd["melspec_extractor.spectrogram.window"], sd["melspec_extractor.mel_scale.fb"],
                      c["n_fft"], c["hop"], None)


def body(sd, lms, feat_length):
    """cnn_encoder.py:420-464 from the dB log-mel [B, 64, T] on."""
    x = lms.transpose(1, 2).unsqueeze(1)                   # (B, 1, T, mel)
    x = _bn(x.transpose(1, 3), sd, "bn0").transpose(1, 3)  # BatchNorm over mel
    for i in range(1, 7):
        x = conv_block(x, sd, i, (2, 2) if i < 6 else (1, 1))
    attn_emb = x.mean(dim=3).transpose(1, 2)               # (B, T', 2048)
    pooled = max_with_lens(attn_emb, feat_length) + mean_with_lens(attn_emb, feat_length)
    fc_emb = F.relu(F.linear(pooled, sd["fc1.weight"], sd["fc1.bias"]))
    return attn_emb, fc_emb


def feat_lengths(wav_len):
    wl = torch.as_tensor(wav_len)
    fl = torch.div(wl, HOP, rounding_mode="floor") + 1
    return torch.div(fl, DOWNSAMPLE, rounding_mode="floor")


@torch.no_grad()
def forward(sd, wav, wav_len):
    fl = feat_lengths(wav_len)
    attn_emb, fc_emb = body(sd, log_mel(sd, wav), fl)
    return {"attn_emb": attn_emb, "fc_emb": fc_emb, "attn_emb_len": fl}
