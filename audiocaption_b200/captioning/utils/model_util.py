"""Length-mask helpers with the reference's semantics (captioning/utils/model_util.py:29-84)."""
import torch


def generate_length_mask(lens, max_length=None):
    """[N, max_length] bool, True where index < length (model_util.py:29-39)."""
    lens = torch.as_tensor(lens)
    if max_length is None:
        max_length = int(lens.max())
    return torch.arange(max_length, device=lens.device).unsqueeze(0) < lens.view(-1, 1)


def mean_with_lens(features, lens):
    """Masked mean over dim 1, divided by the given lengths (model_util.py:41-63)."""
    lens = torch.as_tensor(lens).to(features.device)
    mask = generate_length_mask(lens, features.size(1))
    while mask.ndim < features.ndim:
        mask = mask.unsqueeze(-1)
    out = (features * mask).sum(1)
    return out / lens.view(-1, *([1] * (out.ndim - 1)))


def max_with_lens(features, lens):
    """Masked max over dim 1 (model_util.py:65-81)."""
    lens = torch.as_tensor(lens).to(features.device)
    mask = generate_length_mask(lens, features.size(1))
    f = features.clone()
    f[~mask] = float("-inf")
    return f.max(1)[0]


def repeat_tensor(x, n):
    return x.unsqueeze(0).repeat(n, *([1] * len(x.shape)))
