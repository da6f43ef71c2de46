"""Oracle restatement of the sound-event tagger behind the temporal captioner's tags.
TEST INFRASTRUCTURE (oracle/__init__.py).

Follows captioning/models/hf_wrapper.py:
  * `Cnn8rnnSedModel.forward_prob` :1823-1859 -- bn0 over mel; ConvBlocks 1->64->128->256->512 with pool_type 'avg+max'
    (avg_pool2d + max_pool2d) and pool sizes (2,2) (2,2) (1,2) (1,2); mean over mel; relu(fc1); bidirectional GRU(512, 256)
    over the whole sequence (no packing); sigmoid(fc_audioset).clamp(1e-7, 1); `interpolate` (repeat x4 along time, :54-68)
    and `pad_framewise_output` (repeat the last frame up to the input length, :70-87);
  * `Cnn8rnnSedModel.forward` :1811-1821 -- double_threshold(framewise, 0.75, 0.25) then decode_with_timestamps(., 0.01);
  * `double_threshold` / `_double_threshold` / `find_contiguous_regions` / `connect_` :89-178 -- per (clip, class) column: maximal
    runs of x > low that contain an x > high (the membership test is inclusive of the run's end index), runs closer than
    n_connect = 1 frame merged, result as a 0/1 array;
  * `decode_with_timestamps` / `segments_to_temporal_tag` :180-216 -- every run of every class becomes (class, onset s,
    offset s); over ordered pairs of runs of different classes: overlap = end_j - start_k; overlap < 0.5 * min duration ->
    "after" flag 2; start_j < start_k and overlap > 0.5 * min duration -> "while" flag 1; tag = sum of the flags.

Pinned against the imported reference (tests/test_oracle_cpu.py) and tests/golden/sed.npz.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

CHANNELS = (1, 64, 128, 256, 512)
POOLS = ((2, 2), (2, 2), (1, 2), (1, 2))
CLASSES = 447


def state_dict_keys():
    bn = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")
    keys = [f"bn0.{k}" for k in bn]
    for i in range(1, 5):
        keys += [f"conv_block{i}.conv1.weight", f"conv_block{i}.conv2.weight"]
        keys += [f"conv_block{i}.bn1.{k}" for k in bn] + [f"conv_block{i}.bn2.{k}" for k in bn]
    keys += ["fc1.weight", "fc1.bias"]
    for sfx in ("", "_reverse"):
        keys += [f"rnn.{n}_l0{sfx}" for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    return keys + ["fc_audioset.weight", "fc_audioset.bias"]


def build_state_dict(seed=12, classes=CLASSES):
    """Seeded 'trained-like' tagger: He-scaled convolutions, non-trivial BN statistics, and an output layer strong enough
    that class probabilities cross both thresholds (events of different classes start and stop inside a clip)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def bn(prefix, c):
        sd[prefix + ".weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.long)

    bn("bn0", 64)
    sd["bn0.running_mean"] = -30.0 + 10.0 * torch.randn(64, generator=g)
    sd["bn0.running_var"] = 200.0 + 200.0 * torch.rand(64, generator=g)
    for i in range(1, 5):
        ci, co = CHANNELS[i - 1], CHANNELS[i]
        sd[f"conv_block{i}.conv1.weight"] = torch.randn(co, ci, 3, 3, generator=g) * math.sqrt(2.0 / (9 * ci))
        sd[f"conv_block{i}.conv2.weight"] = torch.randn(co, co, 3, 3, generator=g) * math.sqrt(2.0 / (9 * co))
        bn(f"conv_block{i}.bn1", co)
        bn(f"conv_block{i}.bn2", co)
    sd["fc1.weight"] = torch.randn(512, 512, generator=g) * math.sqrt(1.0 / 512)
    sd["fc1.bias"] = 0.1 * torch.randn(512, generator=g)
    for sfx in ("", "_reverse"):
        sd[f"rnn.weight_ih_l0{sfx}"] = torch.randn(768, 512, generator=g) * (1.5 / math.sqrt(512))
        sd[f"rnn.weight_hh_l0{sfx}"] = torch.randn(768, 256, generator=g) * (1.0 / math.sqrt(256))
        sd[f"rnn.bias_ih_l0{sfx}"] = 0.1 * torch.randn(768, generator=g)
        sd[f"rnn.bias_hh_l0{sfx}"] = 0.1 * torch.randn(768, generator=g)
    g = torch.Generator().manual_seed(seed + 87)
    sd["fc_audioset.weight"] = torch.randn(classes, 512, generator=g) * (2.0 / math.sqrt(512))
    sd["fc_audioset.bias"] = -5.0 + 0.5 * torch.randn(classes, generator=g)      # a handful of classes fire per clip
    return {k: sd[k] for k in state_dict_keys()}


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=1e-5)


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H)
    out = torch.zeros(B, T, H)
    gi_all = x @ w_ih.t() + b_ih
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gi, gh = gi_all[:, t], h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        out[:, t] = h
    return out


@torch.no_grad()
def forward_prob(sd, lms):
    """lms [B, 64, T] -> segmentwise [B, T//4, classes], framewise [B, T, classes]."""
    frames = lms.shape[2]
    x = lms.transpose(1, 2).unsqueeze(1)
    x = _bn(x.transpose(1, 3), sd, "bn0").transpose(1, 3)
    for i in range(1, 5):
        x = F.relu(_bn(F.conv2d(x, sd[f"conv_block{i}.conv1.weight"], padding=1), sd, f"conv_block{i}.bn1"))
        x = F.relu(_bn(F.conv2d(x, sd[f"conv_block{i}.conv2.weight"], padding=1), sd, f"conv_block{i}.bn2"))
        x = F.avg_pool2d(x, POOLS[i - 1]) + F.max_pool2d(x, POOLS[i - 1])
    x = x.mean(dim=3).transpose(1, 2)
    x = F.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))
    fwd = _gru_direction(x, sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"], False)
    bwd = _gru_direction(x, sd["rnn.weight_ih_l0_reverse"], sd["rnn.weight_hh_l0_reverse"], sd["rnn.bias_ih_l0_reverse"],
                         sd["rnn.bias_hh_l0_reverse"], True)
    seg = torch.sigmoid(F.linear(torch.cat((fwd, bwd), dim=-1), sd["fc_audioset.weight"], sd["fc_audioset.bias"])).clamp(1e-7, 1.0)
    frame = seg.repeat_interleave(4, dim=1)
    if frame.shape[1] < frames:
        frame = torch.cat((frame, frame[:, -1:].expand(-1, frames - frame.shape[1], -1)), dim=1)
    return seg, frame


def runs(mask):
    """[(start, end)] of the maximal True runs of a 1-D bool array (end exclusive)."""
    out, start = [], None
    for i, v in enumerate(mask):
        if v and start is None:
            start = i
        elif not v and start is not None:
            out.append((start, i))
            start = None
    if start is not None:
        out.append((start, len(mask)))
    return out


def double_threshold_column(x, high=0.75, low=0.25, n_connect=1):
    """One (clip, class) column of frame probabilities -> 0/1 int array."""
    highs = np.nonzero(x > high)[0]
    kept = [(a, b) for a, b in runs(x > low) if ((a <= highs) & (highs <= b)).any()]
    merged = []
    for a, b in kept:                                   # merge runs whose gap is <= n_connect
        if merged and a - merged[-1][1] <= n_connect:
            merged[-1] = (merged[-1][0], b)
        else:
            merged.append((a, b))
    y = np.zeros(len(x), dtype=int)
    for a, b in merged:
        y[a:b] = 1
    return y


def temporal_tag(labels, resolution=0.01, thre=0.5):
    """labels [T, classes] 0/1 -> tag in {0, 1, 2, 3}."""
    segs = []
    for c in range(labels.shape[1]):
        for a, b in runs(labels[:, c] != 0):
            segs.append((c, a * resolution, b * resolution))
    after, whil = 0, 0
    for sj in segs:
        for sk in segs:
            if sj[0] == sk[0]:
                continue
            min_dur = min(sj[2] - sj[1], sk[2] - sk[1])
            overlap = sj[2] - sk[1]
            if overlap < thre * min_dur:
                after = 2
            if sj[1] < sk[1] and overlap > thre * min_dur:
                whil = 1
    return after + whil


def tags(sd, lms, high=0.75, low=0.25):
    """Cnn8rnnSedModel.forward: list of temporal tags, one per clip."""
    _, frame = forward_prob(sd, lms)
    f = frame.numpy()
    out = []
    for b in range(f.shape[0]):
        lab = np.stack([double_threshold_column(f[b, :, c], high, low) for c in range(f.shape[2])], axis=1)
        out.append(temporal_tag(lab))
    return out
