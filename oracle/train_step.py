"""Oracle restatement of ONE training step of the Cnn14Rnn-Transformer captioner.  TEST INFRASTRUCTURE (oracle/__init__.py).

Follows, in the reference's order (python_scripts/train_eval/run.py:77-148 `Runner._train_epoch`):
  * run.py:21-47 `_forward(training=True)`: model({"mode": "train", wav, wav_len, cap, cap_len, ss_ratio, specaug}),
    output["tgt"] = cap[:, 1:], output["tgt_len"] = cap_len - 1;
  * captioning/models/base.py:48-137 `CaptionModel.forward` / `train_forward`: ss_ratio != 1 -> `stepwise_forward` in train
    mode (:152-170, no early stop), else `TransformerModel.seq_forward` (transformer_model.py:20-32);
  * captioning/models/transformer_model.py:34-57 `prepare_decoder_input`: ONE python coin per step for the whole batch,
    GT prefix cap[:, :t+1] or <start> + own samples seq[:, :t]; `cap_padding_mask = word == pad`;
  * base.py:172-218 `decode_step` / `sample_next_word("greedy")`: last position's logits, arg-max of the log-softmax;
  * captioning/losses/loss.py:51-74 `LabelSmoothingLoss`;
  * run.py:123-127: `loss.backward()`, `clip_grad_norm_(model.parameters(), max_grad_norm)`, `optimizer.step()` with
    torch.optim.Adam(lr, weight_decay) (eg_configs/*/waveform/cnn14rnn_trm.yaml:42-46);
  * captioning/utils/lr_scheduler.py:5-45 `ExponentialDecayScheduler` (closed form in `exponential_decay_lr`).

The decoder loop below is the reference's literal O(L^2) one (a full-prefix decoder call per step); the bi-GRU is an
explicit per-clip, per-step loop with packed-sequence semantics (oracle/crnn.py), here differentiable.  Dropout is OFF in
the oracle: parity of the deterministic arithmetic is what is pinned (tests compare the CUDA path with its dropout
probabilities set to 0); the frozen CNN (oracle/cnn14.py) only produces the GRU's input.  Gradients come from torch
autograd on the CPU; Adam and the clip are restated explicitly and pinned against torch.optim.Adam / clip_grad_norm_.
"""
import math

import torch

from . import caption_model as cm
from . import cnn14 as oc
from . import crnn


# ------------------------------------------------------------------------------------------------ schedules
def exponential_decay_lr(step_count, base_lr, final_lr, total_iters, warmup_iters):
    """lr_scheduler.py:22-42 with current_iter = the scheduler's `_step_count` (1 after construction; the k-th training
    iteration, 1-based, steps it to k + 1 before the optimizer step, run.py:105)."""
    if step_count < warmup_iters:
        return (step_count / warmup_iters) * base_lr
    if step_count == warmup_iters:
        return base_lr
    base = (final_lr / base_lr) ** (1 / (total_iters - warmup_iters))
    return base_lr * (base ** (step_count - warmup_iters))


def ss_ratio_after(iteration_1based, total_iters, final_ratio=0.7, mode="linear"):
    """run.py:55-65 applied `iteration_1based` times starting from 1.0."""
    r = 1.0
    for _ in range(iteration_1based):
        r = r * 0.01 ** (1.0 / total_iters) if mode == "exponential" else r - (1.0 - final_ratio) / total_iters
    return r


# ------------------------------------------------------------------------------------------------ dropout masks
# The reference draws its dropout masks from torch's global RNG; no implementation can reproduce those bit for bit, so the
# parity bar for the stochastic mode is: (i) with p = 0 everything equals the reference, (ii) with p > 0 the CUDA path
# equals THIS restatement of the same network evaluated with the SAME masks.  The masks are a pure function
# keep(seed, site, element index) -- the counter-based generator of audiocaption_b200/csrc/train_ops.cuh, restated here.
_M64 = (1 << 64) - 1


def _mix64(z):
    import numpy as np
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def drop_scale(seed, site, n, p, dtype=torch.float64):
    """Multipliers of elements 0..n-1 of dropout site `site`: 0 with probability p, else 1 / (1 - p)."""
    import numpy as np
    if p <= 0:
        return torch.ones(n, dtype=dtype)
    key = _mix64(np.array([(seed ^ (site << 40)) & _M64], dtype=np.uint64))[0]
    with np.errstate(over="ignore"):
        h = _mix64(key + np.arange(n, dtype=np.uint64))
    u = (h >> np.uint64(40)).astype(np.float64) / 16777216.0
    return torch.from_numpy(np.where(u.astype(np.float32) < np.float32(p), 0.0, 1.0 / (1.0 - np.float32(p)))).to(dtype)


SITE_EMB_IN, SITE_EMB_PE, SITE_MEM, SITE_LAYER0, SITE_GRU0 = 0, 1, 2, 8, 64


def decoder_forward_masked(dec, word, attn_emb, attn_emb_len, pad, p, seed, row0=0):
    """`TransformerDecoder.forward` (transformer_decoder.py:80-103) written out op by op -- nn.TransformerDecoderLayer
    (post-norm, ReLU), nn.MultiheadAttention -- with an explicit dropout multiplier at every place the reference applies
    dropout in train mode.  Returns logits [B, L, V].  Works in the dtype of `dec` (tests use float64)."""
    import torch.nn.functional as F
    B, L = word.shape
    T = attn_emb.shape[1]
    D, H = dec.d_model, dec.model.layers[0].self_attn.num_heads
    dt = dec.classifier.weight.dtype
    ms = lambda site, shape: drop_scale(seed, site, int(torch.tensor(shape).prod()), p, dt).view(*shape)
    lin0, ln = dec.attn_proj[0], dec.attn_proj[3]
    U = torch.relu(attn_emb.to(dt) @ lin0.weight.T + lin0.bias) * ms(SITE_MEM, (B, T, D))
    mem = F.layer_norm(U, (D,), ln.weight, ln.bias, 1e-5)
    X = dec.word_embedding(word) * ms(SITE_EMB_IN, (B, L, D)) * math.sqrt(D) + dec.pos_encoder.pe[:L, 0].unsqueeze(0)
    X = X * ms(SITE_EMB_PE, (B, L, D))
    causal = torch.full((L, L), float("-inf"), dtype=dt).triu(1)
    key_bias = torch.zeros(B, 1, 1, L, dtype=dt).masked_fill(pad.view(B, 1, 1, L), float("-inf"))
    mem_bias = torch.zeros(B, 1, 1, T, dtype=dt).masked_fill(~cm.length_mask(attn_emb_len, T).view(B, 1, 1, T), float("-inf"))
    heads = lambda t, n: t.view(B, n, H, D // H).transpose(1, 2)
    for l, layer in enumerate(dec.model.layers):
        base = SITE_LAYER0 + 8 * l
        sa, ca = layer.self_attn, layer.multihead_attn
        q, k, v = (X @ sa.in_proj_weight.T + sa.in_proj_bias).split(D, dim=-1)
        P = torch.softmax(heads(q, L) @ heads(k, L).transpose(-1, -2) / math.sqrt(D // H) + causal + key_bias, dim=-1)
        A = ((P * ms(base + 0, (B, H, L, L))) @ heads(v, L)).transpose(1, 2).reshape(B, L, D)
        O = A @ sa.out_proj.weight.T + sa.out_proj.bias
        X1 = F.layer_norm(X + O * ms(base + 1, (B, L, D)), (D,), layer.norm1.weight, layer.norm1.bias, 1e-5)
        qc = X1 @ ca.in_proj_weight[:D].T + ca.in_proj_bias[:D]
        kc, vc = (mem @ ca.in_proj_weight[D:].T + ca.in_proj_bias[D:]).split(D, dim=-1)
        Pc = torch.softmax(heads(qc, L) @ heads(kc, T).transpose(-1, -2) / math.sqrt(D // H) + mem_bias, dim=-1)
        Ac = ((Pc * ms(base + 2, (B, H, L, T))) @ heads(vc, T)).transpose(1, 2).reshape(B, L, D)
        Oc = Ac @ ca.out_proj.weight.T + ca.out_proj.bias
        X2 = F.layer_norm(X1 + Oc * ms(base + 3, (B, L, D)), (D,), layer.norm2.weight, layer.norm2.bias, 1e-5)
        Hf = torch.relu(X2 @ layer.linear1.weight.T + layer.linear1.bias) * ms(base + 4, (B, L, layer.linear1.out_features))
        Ff = Hf @ layer.linear2.weight.T + layer.linear2.bias
        X = F.layer_norm(X2 + Ff * ms(base + 5, (B, L, D)), (D,), layer.norm3.weight, layer.norm3.bias, 1e-5)
    return X @ dec.classifier.weight.T


# ------------------------------------------------------------------------------------------------ differentiable bi-GRU
def gru_params(sd, requires_grad=True, dtype=torch.float32):
    return {k: v.clone().to(dtype).requires_grad_(requires_grad) for k, v in sd.items()}


def bigru(p, x, lens, hidden=crnn.HIDDEN, layers=crnn.LAYERS, p_drop=0.0, seed=0):
    """Same arithmetic as oracle/crnn.py `bigru`, built from out-of-place ops so autograd can differentiate it.
    p_drop: nn.GRU's inter-layer dropout with the masks of `drop_scale` (site SITE_GRU0 + layer)."""
    lens = torch.as_tensor(lens)
    B, T = x.shape[0], int(lens.max())
    inp = x[:, :T]
    dt = x.dtype
    for l in range(layers):
        dirs = []
        for d, suffix in enumerate(("", "_reverse")):
            w_ih, w_hh = p[f"network.weight_ih_l{l}{suffix}"], p[f"network.weight_hh_l{l}{suffix}"]
            b_ih, b_hh = p[f"network.bias_ih_l{l}{suffix}"], p[f"network.bias_hh_l{l}{suffix}"]
            rows = []
            for b in range(B):
                n = int(lens[b])
                h = torch.zeros(hidden, dtype=dt)
                outs = [None] * n
                for t in (range(n) if d == 0 else range(n - 1, -1, -1)):
                    h = crnn._cell(inp[b, t], h, w_ih, w_hh, b_ih, b_hh)
                    outs[t] = h
                outs += [torch.zeros(hidden, dtype=dt)] * (T - n)
                rows.append(torch.stack(outs))
            dirs.append(torch.stack(rows))
        inp = torch.cat(dirs, dim=-1)
        if l + 1 < layers and p_drop > 0:
            inp = inp * drop_scale(seed, SITE_GRU0 + l, inp.numel(), p_drop, dt).view(inp.shape)
    return inp


# ------------------------------------------------------------------------------------------------ decoder train forward
def stepwise_train_forward(dec, attn_emb, attn_emb_len, cap, coins):
    """base.py:152-170 (mode == "train") over transformer_model.py:34-57.  coins[t] True = GT prefix at step t.
    Returns logit [B, L, V], embed [B, L, D], seq [B, L], sampled_logprob [B, L]."""
    B, L = cap.size(0), cap.size(1) - 1
    seq = torch.full((B, L), cm.END, dtype=torch.long)
    logits, embeds, logprobs = [], [], []
    for t in range(L):
        if coins[t]:
            word = cap[:, :t + 1]
        else:
            start = torch.full((B, 1), cm.START, dtype=torch.long)
            word = start if t == 0 else torch.cat((start, seq[:, :t]), dim=-1)
        out = dec(word, attn_emb, attn_emb_len, word == cm.PAD)
        logit_t, embed_t = out["logit"][:, -1, :], out["embed"][:, -1, :]
        logprob = torch.log_softmax(logit_t, dim=1)
        lp, w = torch.max(logprob.detach(), 1)
        seq[:, t] = w
        logits.append(logit_t); embeds.append(embed_t); logprobs.append(lp)
    return {"logit": torch.stack(logits, 1), "embed": torch.stack(embeds, 1), "seq": seq,
            "sampled_logprob": torch.stack(logprobs, 1)}


def seq_forward(dec, attn_emb, attn_emb_len, cap):
    """transformer_model.py:20-32 (ss_ratio == 1)."""
    return dec(cap[:, :-1], attn_emb, attn_emb_len, (cap == cm.PAD)[:, :-1])


def label_smoothing_loss(logit, tgt, tgt_len, smoothing=0.1):
    """loss.py:51-74, reduction 'mean'."""
    preds = logit.log_softmax(dim=-1)
    V = logit.size(-1)
    true_dist = torch.full_like(preds, smoothing / (V - 1))
    true_dist.scatter_(-1, tgt.unsqueeze(-1), 1.0 - smoothing)
    loss = torch.sum(-true_dist * preds, dim=-1)
    mask = cm.length_mask(torch.as_tensor(tgt_len), logit.size(1)).to(loss.dtype)
    return (loss * mask).sum() / mask.sum()


# ------------------------------------------------------------------------------------------------ optimizer
def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_: total 2-norm over all gradients, coefficient clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam (amsgrad False, L2 weight decay added to the gradient).  Returns (p', m', v')."""
    g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


# ------------------------------------------------------------------------------------------------ one step
def synth_captions(batch, max_len, vocab, seed=1, min_len=5):
    """Seeded captions as TextCollate would deliver them: <start> ... <end>, pad 0, sorted by length (longest first)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, max_len + 1, (batch,), generator=g).sort(descending=True).values
    cap = torch.zeros(batch, int(lens.max()), dtype=torch.long)
    for b in range(batch):
        n = int(lens[b])
        cap[b, 0] = cm.START
        cap[b, 1:n - 1] = torch.randint(4, vocab, (n - 2,), generator=g)
        cap[b, n - 1] = cm.END
    return cap, lens


def train_step(cnn_sd, rnn_sd, dec, wav, wav_len, cap, cap_len, coins, lr, smoothing=0.1, max_grad_norm=1.0,
               weight_decay=1e-6, adam_state=None, step=1, dtype=torch.float32, cnn_noise=0.0):
    """One full step.  Returns dict(loss, output, grads {name: tensor}, grad_norm, new_params {name: tensor}, state).
    dtype=torch.float64 evaluates the trainable part (bi-GRU, decoder, loss, optimizer) in double precision: the yardstick
    against which BOTH fp32 implementations (the reference's CPU kernels and the CUDA path) are measured -- the fp32 CPU
    autograd is itself up to 2e-3 away from it on some gradients."""
    with torch.no_grad():
        c = oc.forward(cnn_sd, wav, wav_len)                 # frozen CNN, BatchNorm eval, (dropout off in the oracle)
        c["attn_emb"] = c["attn_emb"].to(dtype)
        if cnn_noise > 0:      # sensitivity probe: the frozen CNN's features perturbed at the level of fp32 re-ordering
            g = torch.Generator().manual_seed(77)
            c["attn_emb"] = c["attn_emb"] * (1 + cnn_noise * torch.randn(c["attn_emb"].shape, generator=g, dtype=dtype))
    rp = gru_params(rnn_sd, dtype=dtype)
    dec = dec.to(dtype)
    for p in dec.parameters():
        p.requires_grad_(True)
    dec.pos_encoder.pe.requires_grad_(False)
    dec.zero_grad()
    mem = bigru(rp, c["attn_emb"], c["attn_emb_len"])
    if coins is None:
        out = seq_forward(dec, mem, c["attn_emb_len"], cap)
    else:
        out = stepwise_train_forward(dec, mem, c["attn_emb_len"], cap, coins)
    loss = label_smoothing_loss(out["logit"], cap[:, 1:], torch.as_tensor(cap_len) - 1, smoothing)
    loss.backward()
    named = {f"encoder.rnn.{k}": v for k, v in rp.items()}
    named.update({f"decoder.{k}": v for k, v in dec.named_parameters() if v.requires_grad})
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach().clone() for k, v in named.items()}
    total, coef = clip_coef(list(grads.values()), max_grad_norm)
    state = adam_state or {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in named.items()}
    new_params, new_state = {}, {}
    for k, v in named.items():
        p2, m2, v2 = adam_update(v.detach(), grads[k] * coef, state[k][0], state[k][1], step, lr, weight_decay=weight_decay)
        new_params[k], new_state[k] = p2, (m2, v2)
    return {"loss": loss.detach(), "output": {k: v.detach() for k, v in out.items()}, "grads": grads, "grad_norm": total,
            "new_params": new_params, "state": new_state, "mem": mem.detach(), "attn_emb_len": c["attn_emb_len"]}
