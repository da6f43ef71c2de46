"""Mirror of captioning/losses/loss.py:40-74 `LabelSmoothingLoss`: length-masked label-smoothing cross entropy over the
logits of a training forward.  The forward and the gradient w.r.t. the logits come from ONE fused kernel pass
(csrc/train_ops.cu `ac_ls_ce_fwd_bwd`: log-softmax, smoothed target distribution, mask, mean over valid tokens and
(softmax - true_dist) * mask / n_tokens), wrapped in a ``torch.autograd.Function``."""
from typing import Dict

import torch
import torch.nn as nn

from ... import _lib
from ..models._native import require_cuda, to_device_async


def ls_ce_fwd_bwd(logit, tgt, tgt_len_dev, smoothing, want_grad=True):
    """logit [B, L, V] fp32 cuda (last-dim stride 1; rows may be padded), tgt [B, L] int64 cuda (a strided view such as
    cap[:, 1:] is fine), tgt_len_dev [B] int64 cuda -> (loss [1], dlogit with logit's row layout or None)."""
    B, L, V = logit.shape
    if logit.stride(2) != 1 or logit.stride(0) != L * logit.stride(1):
        logit = logit.contiguous()
    ld = logit.stride(1)
    if tgt.stride(1) != 1:
        tgt = tgt.contiguous()
    dev = logit.device
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dfull = torch.empty(B, L, ld, dtype=torch.float32, device=dev) if want_grad else None
        ws = torch.empty(B * L, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().ac_ls_ce_fwd_bwd(_lib.ptr(logit), ld, _lib.ptr(tgt), tgt.stride(0), _lib.ptr(tgt_len_dev), B, L, V,
                                               float(smoothing), 1.0, _lib.ptr(loss), _lib.ptr(dfull), _lib.ptr(ws),
                                               ws.numel() * 4, _lib.current_stream()), "ac_ls_ce_fwd_bwd")
    return loss, dfull


class _LabelSmoothingFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logit, tgt, tgt_len_dev, smoothing):
        loss, dfull = ls_ce_fwd_bwd(logit.detach(), tgt, tgt_len_dev, smoothing, want_grad=True)
        ctx.save_for_backward(dfull)
        ctx.V = logit.shape[-1]
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, dloss):
        (dfull,) = ctx.saved_tensors
        return dfull[:, :, :ctx.V] * dloss, None, None, None


class LabelSmoothingLoss(nn.Module):

    def __init__(self, smoothing=0.0, dim=-1, reduction="mean", logit_name="logit", target_name="tgt"):
        super().__init__()
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing
        self.dim = dim
        self.reduction = reduction
        self.logit_name = logit_name
        self.target_name = target_name
        if dim not in (-1, 2) or reduction != "mean":
            raise NotImplementedError("the B200 loss implements dim=-1, reduction='mean' (the training configs' setting)")

    def forward(self, output: Dict):
        logit = output[self.logit_name]                       # [bs, max_len, c]
        tgt = output[self.target_name]                        # [bs, max_len]
        tgt_len = output[f"{self.target_name}_len"]           # [bs]
        require_cuda(logit, "LabelSmoothingLoss")
        tgt = tgt.to(logit.device)
        tgt_len_dev = to_device_async(torch.as_tensor(tgt_len), logit.device, torch.int64)
        return _LabelSmoothingFn.apply(logit.float(), tgt, tgt_len_dev, self.smoothing)
