"""GPU parity tests of the training step (`-m gpu`): the CUDA forward / backward / loss / optimizer through the C ABI against
the CPU oracle (oracle/train_step.py, pinned to the reference's own training step) and the golden vectors
(tests/golden/train_step.npz), on identical seeded inputs with every dropout probability set to 0; plus stochastic-mode
checks (mask determinism, forward / backward mask agreement by finite differences).  fp32; tolerances next to each check."""
import ctypes
import os
import random
import warnings

import numpy as np
import pytest
import torch

from oracle import caption_model as cm
from oracle import cnn14 as oc
from oracle import crnn
from oracle import train_step as ts

warnings.filterwarnings("ignore")
pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


# ------------------------------------------------------------------ loss kernel (row A15)
@pytest.mark.parametrize("B,L,V,pad", [(3, 5, 37, 3), (4, 8, 4368, 0), (2, 21, 4981, 3)])
def test_label_smoothing_ce_matches_oracle(B, L, V, pad):
    from audiocaption_b200.captioning.losses.loss import LabelSmoothingLoss, ls_ce_fwd_bwd
    g = torch.Generator().manual_seed(B * 100 + L)
    full = torch.randn(B, L, V + pad, generator=g) * 3
    logit = full[:, :, :V]
    tgt_full = torch.randint(0, V, (B, L + 1), generator=g)
    tgt = tgt_full[:, 1:]                                   # a strided view, as cap[:, 1:] in run.py:46
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[0] = L
    ref_logit = logit.clone().requires_grad_(True)
    want = ts.label_smoothing_loss(ref_logit, tgt, lens, 0.1)
    want.backward()
    loss, dl = ls_ce_fwd_bwd(full.to(DEV)[:, :, :V], tgt_full.to(DEV)[:, 1:], lens.to(DEV), 0.1)
    assert abs(loss.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert (dl[:, :, :V].cpu() - ref_logit.grad).abs().max() < 1e-7 + 1e-5 * ref_logit.grad.abs().max()
    assert (dl[:, :, V:] == 0).all()
    # the module face (captioning/losses/loss.py:40-74), differentiable
    lg = full.to(DEV)[:, :, :V].clone().requires_grad_(True)
    out = LabelSmoothingLoss(smoothing=0.1)({"logit": lg, "tgt": tgt.to(DEV), "tgt_len": lens})
    (out * 2.0).backward()
    assert abs(out.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert (lg.grad.cpu() - 2.0 * ref_logit.grad).abs().max() < 1e-7 + 1e-5 * ref_logit.grad.abs().max()


# ------------------------------------------------------------------ clip + Adam (row A16)
def test_clip_adam_matches_oracle_including_nan_skip():
    from audiocaption_b200 import _lib
    l = _lib.lib()
    n = 100_003
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(n, generator=g)
    p = p0.to(DEV).clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    norm = torch.zeros(1, device=DEV)
    ws = torch.empty(l.ac_clip_adam_workspace_bytes(), dtype=torch.uint8, device=DEV)
    rp, rm, rv = p0.clone(), torch.zeros(n), torch.zeros(n)
    lrs, scales = [1e-4, 5e-4, 3e-4, 2e-4], [30.0, 1e-3, 1.0, 1.0]
    k = 0
    for it, (lr, sc) in enumerate(zip(lrs, scales)):
        grad = torch.randn(n, generator=g) * sc
        loss = torch.tensor([float("nan") if it == 2 else 1.0], device=DEV)
        _lib.check(l.ac_clip_adam(_lib.ptr(p), _lib.ptr(grad.to(DEV)), _lib.ptr(m), _lib.ptr(v), n, lr, 0.9, 0.999, 1e-8, 1e-6, 1.0,
                                  0.5, _lib.ptr(loss), _lib.ptr(step), _lib.ptr(norm), _lib.ptr(ws), ws.numel(), None), "ac_clip_adam")
        torch.cuda.synchronize()
        if it == 2:                                      # NaN loss: run.py:123 skips backward + step entirely
            assert step.item() == k and (p.cpu() == rp_last).all()
            continue
        k += 1
        total, coef = ts.clip_coef([grad * 0.5], 1.0)
        rp, rm, rv = ts.adam_update(rp, grad * 0.5 * coef, rm, rv, k, lr, weight_decay=1e-6)
        rp_last = p.cpu().clone()
        assert step.item() == k
        assert abs(norm.item() - total.item()) < 1e-4 * total.item()
        assert (p.cpu() - rp).abs().max() < 1e-6, it          # 1-2 ulp of |p| ~ 4
        assert (m.cpu() - rm).abs().max() < 1e-6 * max(1.0, rm.abs().max().item())
        assert (v.cpu() - rv).abs().max() < 1e-6 * max(1.0, rv.abs().max().item())


# ------------------------------------------------------------------ decoder full-prefix forward (row A7) and its gradients
def _decoder_pair(vocab, attn_emb_dim=512, seed=6):
    from audiocaption_b200.captioning.models.transformer_decoder import TransformerDecoder
    orc = crnn.build_decoder(seed, attn_emb_dim=attn_emb_dim, vocab_size=vocab)
    dec = TransformerDecoder(emb_dim=256, vocab_size=vocab, fc_emb_dim=512, attn_emb_dim=attn_emb_dim, nlayers=2, dropout=0.2)
    dec.load_state_dict(orc.state_dict(), strict=True)
    return dec.to(DEV), orc


def _decoder_inputs(B, T, L, vocab, E=512, seed=0):
    g = torch.Generator().manual_seed(seed)
    mem = 0.5 * torch.tanh(torch.randn(B, T, E, generator=g))
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    word = torch.randint(3, vocab, (B, L), generator=g)
    word[:, 0] = cm.START
    for b in range(1, B):                                   # ragged captions: trailing <pad>
        n = int(torch.randint(2, L + 1, (1,), generator=g))
        word[b, n:] = cm.PAD
    return mem, lens, word


@pytest.mark.parametrize("B,T,L,vocab", [(4, 9, 8, 4981), (3, 31, 21, 4368), (1, 1, 1, 4368), (5, 128, 40, 520)])
def test_decoder_forward_input_dict_matches_oracle(B, T, L, vocab):
    """`TransformerDecoder.forward(input_dict)` (transformer_decoder.py:80-103), eval mode: logits / embed of every
    prefix position vs the oracle decoder (pinned to the imported reference at 1e-5): <= 1e-4 of the tensor scale."""
    dec, orc = _decoder_pair(vocab)
    dec.eval()
    mem, lens, word = _decoder_inputs(B, T, L, vocab, seed=B + T)
    pad = word == cm.PAD
    with torch.no_grad():
        want = orc(word, mem, lens, pad)
        got = dec({"word": word, "attn_emb": mem.to(DEV), "attn_emb_len": lens, "cap_padding_mask": pad})
    assert got["logit"].shape == (B, L, vocab) and got["embed"].shape == (B, L, 256)
    valid = ~pad                                            # rows whose query is a padding token are not used by anyone
    assert _relerr(got["embed"].cpu()[valid], want["embed"][valid]) < 1e-4
    assert _relerr(got["logit"].cpu()[valid], want["logit"][valid]) < 1e-4


@pytest.mark.parametrize("B,T,L,vocab,tied", [(4, 9, 8, 4981, False), (3, 31, 21, 4368, False), (2, 5, 6, 520, True)])
def test_decoder_gradients_match_oracle_autograd(B, T, L, vocab, tied):
    """d(sum(logit * R)) w.r.t. every decoder parameter and the audio memory: CUDA backward vs torch autograd on the oracle
    evaluated in float64 (the fp32 CPU autograd of the same oracle is itself 1e-3 .. 2e-2 away from float64 on these
    gradients -- scripts/train_debug.py -- so it cannot serve as the yardstick): <= 3e-4 of each tensor's scale."""
    from audiocaption_b200.captioning.models.transformer_decoder import TransformerDecoder
    if tied:
        orc = cm.TransformerDecoder(emb_dim=256, vocab_size=vocab, attn_emb_dim=512, tie_weights=True)
        torch.manual_seed(3)
        for p in orc.parameters():
            if p.dim() > 1:
                torch.nn.init.xavier_uniform_(p, gain=2.0)
        orc.eval()
        dec = TransformerDecoder(emb_dim=256, vocab_size=vocab, fc_emb_dim=512, attn_emb_dim=512, nlayers=2, dropout=0.2,
                                 tie_weights=True)
        dec.load_state_dict(orc.state_dict(), strict=True)
        dec = dec.to(DEV)
    else:
        dec, orc = _decoder_pair(vocab)
    dec.train()
    for mod in dec.modules():                               # deterministic parity: dropout off, train-mode code path
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    mem, lens, word = _decoder_inputs(B, T, L, vocab, seed=7 * B + L)
    pad = word == cm.PAD
    R = torch.randn(B, L, vocab, generator=torch.Generator().manual_seed(1)) * (~pad).unsqueeze(-1)
    orc = orc.double()
    for p in orc.parameters():
        p.requires_grad_(True)
    orc.pos_encoder.pe.requires_grad_(False)
    orc.zero_grad()
    mem_ref = mem.double().requires_grad_(True)
    (orc(word, mem_ref, lens, pad)["logit"] * R.double()).sum().backward()
    mem_dev = mem.to(DEV).requires_grad_(True)
    out = dec({"word": word, "attn_emb": mem_dev, "attn_emb_len": lens, "cap_padding_mask": pad})
    (out["logit"] * R.to(DEV)).sum().backward()
    assert _relerr(mem_dev.grad, mem_ref.grad) < 3e-4
    ref = dict(orc.named_parameters())
    for name, p in dec.named_parameters():
        if name == "pos_encoder.pe":
            assert p.grad is None
            continue
        assert p.grad is not None, name
        assert _relerr(p.grad, ref[name].grad) < 3e-4, name


# ------------------------------------------------------------------ bi-GRU training forward / backward (row A5 + A16)
@pytest.mark.parametrize("B,T,lens", [(3, 6, [6, 4, 1]), (9, 5, [5, 5, 3, 2, 5, 1, 4, 5, 2]), (1, 3, [3])])
def test_bigru_train_forward_backward_matches_oracle(B, T, lens):
    from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
    sd = crnn.build_gru_state_dict(4)
    enc = RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True, hidden_size=256, dropout=0.0,
                     num_layers=3)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).train()
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, T, 2048, generator=g) * 0.3
    R = torch.randn(B, T, 512, generator=g)
    lens_t = torch.tensor(lens)
    rp = ts.gru_params(sd, dtype=torch.float64)
    x_ref = x.double().requires_grad_(True)
    want = ts.bigru(rp, x_ref, lens_t)
    (want * R.double()).sum().backward()
    x_dev = x.to(DEV).requires_grad_(True)
    out = enc({"attn": x_dev, "attn_len": lens_t})
    assert (out["attn_emb"].cpu() - want.detach()).abs().max() < 5e-5
    (out["attn_emb"] * R.to(DEV)).sum().backward()
    assert _relerr(x_dev.grad, x_ref.grad) < 3e-4
    for name, p in enc.named_parameters():
        assert _relerr(p.grad, rp[name].grad) < 3e-4, name
    # fc_emb = mean over the valid frames, differentiable too
    assert (out["fc_emb"].cpu().detach() - oc.mean_with_lens(want.detach().float(), lens_t)).abs().max() < 5e-5


# ------------------------------------------------------------------ the whole training step (rows A9, A15, A16)
def _train_model(vocab):
    from audiocaption_b200.captioning.models.cnn_encoder import Cnn14Encoder
    from audiocaption_b200.captioning.models.crnn_trm_encoder import CrnnEncoder
    from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
    from audiocaption_b200.captioning.models.transformer_decoder import TransformerDecoder
    from audiocaption_b200.captioning.models.transformer_model import TransformerModel
    enc = CrnnEncoder(Cnn14Encoder(sample_rate=32000),
                      RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True, hidden_size=256,
                                 dropout=0.5, num_layers=3), freeze_cnn=True, freeze_cnn_bn=True)
    dec = TransformerDecoder(emb_dim=256, vocab_size=vocab, fc_emb_dim=512, attn_emb_dim=512, nlayers=2, dropout=0.2)
    m = TransformerModel(enc, dec)
    m.load_state_dict(crnn.model_state_dict(oc.build_state_dict(3), crnn.build_gru_state_dict(4),
                                            crnn.build_decoder(6, vocab_size=vocab)), strict=True)
    return m.to(DEV)


def _no_dropout(m):
    m.encoder.cnn.conv_dropout = 0.0
    m.encoder.cnn.fc_dropout = 0.0
    m.encoder.rnn.dropout = 0.0
    for mod in m.decoder.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def _golden_batch(g):
    wav, lens = cm.synth_wav(int(g["batch"]), int(g["n_samples"]), seed=13, ragged=True, varied=True, sample_rate=32000)
    cap, cap_len = ts.synth_captions(int(g["batch"]), int(g["cap_max"]), int(g["vocab"]), seed=1)
    return wav, lens, cap, cap_len


def test_train_step_matches_reference_golden():
    """One fused step (TrainStep) on the inputs of the reference's own step (tests/golden/train_step.npz, produced by
    oracle/gen_golden.py from the imported reference: TransformerModel + LabelSmoothingLoss + clip + Adam): loss,
    sampled tokens, gradient norm, every trainable gradient and the parameter update."""
    from audiocaption_b200.train_step import TrainStep
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "train_step.npz")))
    m = _no_dropout(_train_model(int(g["vocab"])))
    before = {k: v.detach().cpu().clone() for k, v in m.named_parameters() if v.requires_grad}
    wav, lens, cap, cap_len = _golden_batch(g)
    total, warm = 1000, 10
    # pick the iteration whose scheduled learning rate is the golden one: lr(k + 2) = base * (k + 2) / warm
    base = float(g["lr"]) * warm / 2
    step = TrainStep(m, total_iters=total, lr=base, warmup_iters=warm, final_lr=base * 1e-3)
    res = step.step({"wav": wav, "wav_len": lens, "cap": cap, "cap_len": cap_len.numpy()}, coins=[bool(c) for c in g["coins"]])
    torch.cuda.synchronize()
    assert abs(res["lr"] - float(g["lr"])) < 1e-12
    assert abs(res["loss"].item() - float(g["loss"])) < 2e-4
    assert res["tokens"] == int((cap_len - 1).sum())
    assert (step.last_output["seq"].cpu().numpy() == g["seq"]).all()
    assert np.abs(step.last_output["logit"][:, :, :16].cpu().numpy() - g["logit_head"]).max() < 2e-3
    assert abs(step.grad_norm.item() - float(g["gnorm"])) < 2e-3 * float(g["gnorm"])
    names = [str(n) for n in g["names"]]
    params = dict(m.named_parameters())
    for i, k in enumerate(names):
        gr = params[k].grad
        assert abs(gr.norm().item() - g["grad_norms"][i]) <= 2e-3 * g["grad_norms"][i] + 1e-6, k
        head = np.resize(gr.flatten()[:16].cpu().numpy(), 16)
        assert np.abs(head - g["grad_heads"][i]).max() <= 2e-3 * max(np.abs(g["grad_heads"][i]).max(), 1e-3), k
        upd = (params[k].detach().cpu() - before[k]).norm().item()
        assert abs(upd - g["update_norms"][i]) <= 0.03 * g["update_norms"][i] + 1e-7, k


@pytest.mark.parametrize("pattern", ["all_gt", "all_sampled", "last_sampled", "seq_forward", "mixed"])
def test_train_step_matches_oracle_for_every_coin_pattern(pattern):
    """Loss, every gradient and the updated parameters vs the oracle's literal step-by-step loop (V = 4368, the Clotho
    vocabulary of configs[2]) for the corner cases of the scheduled-sampling coin sequence; `seq_forward` is ss_ratio == 1
    (transformer_model.py:20-32)."""
    from audiocaption_b200.train_step import TrainStep
    vocab = 4368
    m = _no_dropout(_train_model(vocab))
    wav, lens = cm.synth_wav(3, 64000, seed=21, ragged=True, varied=True, sample_rate=32000)
    cap, cap_len = ts.synth_captions(3, 7, vocab, seed=4)
    L = cap.size(1) - 1
    coins = {"all_gt": [True] * L, "all_sampled": [False] * L, "last_sampled": [True] * (L - 1) + [False], "seq_forward": None,
             "mixed": [t % 2 == 1 for t in range(L)]}[pattern]
    lr = 3e-4
    step = TrainStep(m, total_iters=1000, lr=lr * 5, warmup_iters=10, final_lr=1e-7, use_ss=pattern != "seq_forward")
    if pattern != "seq_forward":
        step.ss_ratio = 0.5
    dec = crnn.build_decoder(6, vocab_size=vocab)
    o = ts.train_step(oc.build_state_dict(3), crnn.build_gru_state_dict(4), dec, wav, lens, cap, cap_len, coins, lr,
                      dtype=torch.float64)            # the yardstick (see test_decoder_gradients_match_oracle_autograd)
    res = step.step({"wav": wav, "wav_len": lens, "cap": cap, "cap_len": cap_len.numpy()}, coins=coins)
    torch.cuda.synchronize()
    assert abs(res["lr"] - lr) < 1e-12
    assert abs(res["loss"].item() - float(o["loss"])) < 2e-4
    if coins is not None:
        # the sampled tokens feed back into the prefix: exact unless a logit pair is tied at the fp32 rounding level
        lg = o["output"]["logit"]
        top2 = lg.topk(2, dim=-1).values
        stable = ((top2[..., 0] - top2[..., 1]) > 1e-3).all(1)
        assert stable.any()
        assert (step.last_output["seq"].cpu()[stable] == o["output"]["seq"][stable]).all()
        if not stable.all():
            return                      # a flipped sample changes the rest of that row; gradients are compared on stable batches
    # Conditioning: the device's Cnn14 features differ from the CPU's by ~1e-4 relative (3xTF32 convolutions, summation
    # order).  Where a batch parks many FFN pre-activations on the ReLU kink (degenerate sampled rows: the same token at
    # every step) such a difference moves some gradients by percents in ANY implementation; the oracle's own response to a
    # 1e-4 perturbation of the features measures that, and the tolerance is 5e-4 or 3x that response, whichever is larger.
    dec2 = crnn.build_decoder(6, vocab_size=vocab)
    o2 = ts.train_step(oc.build_state_dict(3), crnn.build_gru_state_dict(4), dec2, wav, lens, cap, cap_len, coins, lr,
                       dtype=torch.float64, cnn_noise=1e-4)
    sens = {k: _relerr(o2["grads"][k], o["grads"][k]) for k in o["grads"]}
    assert abs(step.grad_norm.item() - float(o["grad_norm"])) < max(5e-4, 3 * max(sens.values())) * float(o["grad_norm"])
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue
        tol = max(5e-4, 3 * sens[k])
        assert _relerr(p.grad, o["grads"][k]) < tol, (k, sens[k])
        # the first Adam step is lr * g / (|g| + eps) ~ lr * sign(g): compared where g is well above its own uncertainty
        big = o["grads"][k].abs() > max(1e-4, 20 * tol * o["grads"][k].abs().max().item())
        if big.any():
            assert (p.detach().cpu().double() - o["new_params"][k])[big].abs().max() < 0.03 * lr, k


def test_module_api_training_loop_matches_fused_step():
    """The reference's own loop shape (run.py:116-127): model(input_dict) -> LabelSmoothingLoss(output) -> loss.backward()
    -> clip_grad_norm_ -> torch.optim.Adam.step(), through the mirrors' autograd wrappers, gives the same gradients and
    the same updated parameters as the fused TrainStep."""
    from audiocaption_b200.captioning.losses.loss import LabelSmoothingLoss
    from audiocaption_b200.train_step import TrainStep
    vocab = 4368
    wav, lens = cm.synth_wav(3, 64000, seed=22, ragged=True, varied=True, sample_rate=32000)
    cap, cap_len = ts.synth_captions(3, 8, vocab, seed=5)
    L = cap.size(1) - 1
    coin_seed, ss = 3, 0.6
    # ---- module API
    m = _no_dropout(_train_model(vocab)).train()
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=3e-4, weight_decay=1e-6)
    random.seed(coin_seed)
    out = m({"mode": "train", "wav": wav.to(DEV), "wav_len": lens, "specaug": False, "cap": cap.to(DEV),
             "cap_len": cap_len.numpy(), "ss_ratio": ss})
    assert not out["seq"].is_cuda and out["seq"].shape == (3, L) and out["logit"].shape == (3, L, vocab)
    assert out["logit"].requires_grad and out["embed"].shape == (3, L, 256)
    out["tgt"], out["tgt_len"] = cap.to(DEV)[:, 1:], torch.as_tensor(cap_len.numpy() - 1)
    loss = LabelSmoothingLoss(smoothing=0.1)(out)
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.requires_grad}
    gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    opt.step()
    # ---- fused step on a fresh copy of the model, same coins
    m2 = _no_dropout(_train_model(vocab))
    step = TrainStep(m2, total_iters=1000, lr=3e-4 * 5, warmup_iters=10, final_lr=1e-7)
    random.seed(coin_seed)
    coins = [random.random() < ss for _ in range(L)]
    res = step.step({"wav": wav, "wav_len": lens, "cap": cap, "cap_len": cap_len.numpy()}, coins=coins)
    torch.cuda.synchronize()
    assert abs(res["loss"].item() - loss.item()) < 1e-5
    assert abs(step.grad_norm.item() - gnorm.item()) < 1e-4 * gnorm.item()
    p2 = dict(m2.named_parameters())
    for k, gr in grads.items():
        assert _relerr(p2[k].grad, gr) < 1e-5, k
        assert (p2[k].detach() - dict(m.named_parameters())[k].detach()).abs().max() < 2e-6, k


def test_train_mode_dropout_matches_the_oracle_evaluated_with_the_same_masks():
    """Train mode with the YAML's dropout probabilities (decoder 0.2 at all 15 places the reference applies it, GRU 0.5
    between layers).  The masks are a pure function of (seed, site, element index); the oracle restates the generator and
    evaluates the same network in float64 with the SAME masks: logits / GRU output and every gradient must agree, which
    only holds if each kernel -- forward and backward -- regenerates exactly its site's mask.  Also: same seed ->
    bit-identical, other seed -> different."""
    vocab = 520
    from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
    dec, orc = _decoder_pair(vocab)
    dec.train()
    B, T, L = 4, 7, 8
    mem, lens, word = _decoder_inputs(B, T, L, vocab, seed=17)
    pad = word == cm.PAD
    len_dev = lens.to(DEV)
    R = torch.randn(B, L, vocab, generator=torch.Generator().manual_seed(1)) * (~pad).unsqueeze(-1)
    eng = dec.train_engine
    seed = 1234567
    out = eng.forward(mem.to(DEV), len_dev, word.to(DEV), key_pad=pad.to(DEV), p_drop=0.2, seed=seed, grads="scratch")
    again = eng.forward(mem.to(DEV), len_dev, word.to(DEV), key_pad=pad.to(DEV), p_drop=0.2, seed=seed, grads="scratch")
    other = eng.forward(mem.to(DEV), len_dev, word.to(DEV), key_pad=pad.to(DEV), p_drop=0.2, seed=seed + 1, grads="scratch")
    assert (out["logit"] == again["logit"]).all() and not (out["logit"] == other["logit"]).all()
    out = eng.forward(mem.to(DEV), len_dev, word.to(DEV), key_pad=pad.to(DEV), p_drop=0.2, seed=seed, grads="scratch")
    dl = torch.zeros_like(out["logit_padded"])
    dl[:, :, :vocab] = R.to(DEV)
    dmem = eng.backward(dl)
    orc = orc.double()
    for p in orc.parameters():
        p.requires_grad_(True)
    orc.pos_encoder.pe.requires_grad_(False)
    mem_ref = mem.double().requires_grad_(True)
    want = ts.decoder_forward_masked(orc, word, mem_ref, lens, pad, 0.2, seed)
    assert (want.detach() - orc(word, mem.double(), lens, pad)["logit"].detach()).abs().max() > 0.1     # dropout is really on
    valid = ~pad
    assert _relerr(out["logit"].cpu()[valid], want.detach()[valid]) < 1e-4
    (want * R.double()).sum().backward()
    assert _relerr(dmem, mem_ref.grad) < 3e-4
    ref = dict(orc.named_parameters())
    names = [n for n, _ in dec.named_parameters()]
    for p, g in zip(dec._tensors(), eng._grads):
        name = next(n for n, q in dec.named_parameters() if q is p)
        if g is None:
            continue
        assert _relerr(g, ref[name].grad) < 3e-4, name
    # ---- bi-GRU, inter-layer dropout 0.5
    sd = crnn.build_gru_state_dict(4)
    rnn = RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True, hidden_size=256, dropout=0.5,
                     num_layers=3)
    rnn.load_state_dict(sd, strict=True)
    rnn = rnn.to(DEV).train()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 7, 2048, generator=g) * 0.3
    Rg = torch.randn(4, 7, 512, generator=g)
    glens = torch.tensor([7, 5, 7, 2])
    ge = rnn.train_engine
    y = ge.forward(x.to(DEV), glens.to(DEV), p_drop=0.5, seed=seed, grads="scratch", need_dx=True)
    dx = ge.backward(Rg.to(DEV), need_dx=True)
    rp = ts.gru_params(sd, dtype=torch.float64)
    x_ref = x.double().requires_grad_(True)
    want = ts.bigru(rp, x_ref, glens, p_drop=0.5, seed=seed)
    assert (want.detach() - ts.bigru(rp, x.double(), glens).detach()).abs().max() > 0.05
    assert (y.cpu() - want.detach()).abs().max() < 5e-5
    (want * Rg.double()).sum().backward()
    assert _relerr(dx, x_ref.grad) < 3e-4
    for p, gr in zip(rnn._tensors(), ge._grads):
        name = next(n for n, q in rnn.named_parameters() if q is p)
        assert _relerr(gr, rp[name].grad) < 3e-4, name


def test_cnn14_train_mode_dropout():
    """Frozen Cnn14 in train mode (cnn_encoder.py:432-456): BatchNorm folded, dropout active -- seeded, ~p of the block
    outputs zeroed and the rest scaled by 1/(1-p) (checked on the mean of attn_emb over many elements), and eval mode is
    untouched."""
    from audiocaption_b200.captioning.models.cnn_encoder import Cnn14Encoder
    enc = Cnn14Encoder(sample_rate=32000, freeze=True)
    enc.load_state_dict(oc.build_state_dict(3), strict=True)
    enc = enc.to(DEV)
    wav, lens = cm.synth_wav(2, 64000, seed=5, varied=True, sample_rate=32000)
    inp = {"wav": wav.to(DEV), "wav_len": lens, "specaug": False}
    with torch.no_grad():
        ev = enc.eval()(dict(inp))["attn_emb"]
        tr1 = enc.train()(dict(inp, _dropout_seed=5))["attn_emb"]
        tr1b = enc.train()(dict(inp, _dropout_seed=5))["attn_emb"]
        tr2 = enc.train()(dict(inp, _dropout_seed=6))["attn_emb"]
        enc.conv_dropout, enc.fc_dropout = 0.0, 0.0
        tr0 = enc.train()(dict(inp))["attn_emb"]
    assert (tr1 == tr1b).all() and not (tr1 == tr2).all()
    assert (tr0 == ev).all()                                  # p = 0 is the eval arithmetic
    assert (tr1 - ev).abs().max() > 1e-3


def test_training_reduces_the_loss_on_a_fixed_batch():
    """Ten fused steps on one batch with the YAML's hyper-parameters (dropout on, scheduled sampling on): the loss goes down
    and everything stays finite -- an end-to-end sanity check that the update direction is a descent direction."""
    from audiocaption_b200.train_step import TrainStep
    vocab = 4368
    m = _train_model(vocab)
    wav, lens = cm.synth_wav(4, 64000, seed=31, ragged=True, varied=True, sample_rate=32000)
    cap, cap_len = ts.synth_captions(4, 10, vocab, seed=6)
    torch.manual_seed(0)
    random.seed(0)
    step = TrainStep(m, total_iters=100, lr=5e-4, warmup_iters=2)
    batch = {"wav": wav.pin_memory(), "wav_len": lens, "cap": cap, "cap_len": cap_len.numpy()}
    losses = [step.step(batch)["loss"].item() for _ in range(10)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0] - 0.5, losses
    assert step._step_dev.item() == 10


@pytest.mark.parametrize("host_input", [True, False])
def test_look_ahead_steps_equal_inline_steps(host_input):
    """TrainStep.prefetch(next batch) runs the upload and the frozen encoder of batch i+1 on other streams while step i runs
    (capped convolution grid, trainable chain on a high-priority stream).  With dropout off the schedule cannot change any
    number beyond summation order: four look-ahead steps over three different batches give the losses, gradient norms and
    parameters of four inline steps -- a stale / overwritten staging buffer or a missed stream dependency would not."""
    from audiocaption_b200.train_step import TrainStep
    vocab = 520
    batches = []
    for i, (B, n, L) in enumerate([(4, 64000, 10), (3, 96000, 7), (4, 48000, 12)]):
        wav, lens = cm.synth_wav(B, n, seed=40 + i, ragged=True, varied=True, sample_rate=32000)
        cap, cap_len = ts.synth_captions(B, L, vocab, seed=9 + i)
        wav = wav.pin_memory() if host_input else wav.to(DEV)
        batches.append({"wav": wav, "wav_len": lens, "cap": cap if host_input else cap.to(DEV), "cap_len": cap_len.numpy()})
    order = [0, 1, 2, 0]
    coins = [[True] * 32, [True, False] * 16, [False] * 32, [True] * 32]
    runs = {}
    for mode in ("inline", "look_ahead"):
        m = _no_dropout(_train_model(vocab))
        step = TrainStep(m, total_iters=1000, lr=1e-3, warmup_iters=10)
        step.cnn_sms = 100
        losses, norms = [], []
        staged = step.prefetch(batches[order[0]]) if mode == "look_ahead" else None
        for k, bi in enumerate(order):
            L = batches[bi]["cap"].shape[1] - 1
            if mode == "look_ahead":
                nxt = step.prefetch(batches[order[k + 1]]) if k + 1 < len(order) else None
                res = step.step(staged, coins=coins[k][:L])
                staged = nxt
            else:
                res = step.step(batches[bi], coins=coins[k][:L])
            losses.append(res["loss"].item())
            norms.append(step.grad_norm.item())
        torch.cuda.synchronize()
        runs[mode] = (losses, norms, step.flat_param.detach().cpu().clone())
    # (not bit-identical: the embedding gradient is accumulated with fp32 atomics, whose order depends on what else runs)
    # step 0 sees identical parameters: equal to rounding; later steps inherit the noise-level parameter differences below
    # (a stale staging buffer or a missed dependency changes the loss in the first digits, or makes it NaN)
    assert abs(runs["inline"][0][0] - runs["look_ahead"][0][0]) <= 2e-6 * abs(runs["inline"][0][0])
    assert np.allclose(runs["inline"][0], runs["look_ahead"][0], rtol=2e-4, atol=0), (runs["inline"][0], runs["look_ahead"][0])
    assert np.allclose(runs["inline"][1], runs["look_ahead"][1], rtol=2e-3, atol=0), (runs["inline"][1], runs["look_ahead"][1])
    # parameters: Adam divides by sqrt(v) + eps, so an element whose gradient is itself at the atomics' noise level (sums that
    # cancel to ~1e-9) may move by a fraction of the learning rate either way; everything else agrees to fp32 rounding
    diff = (runs["inline"][2] - runs["look_ahead"][2]).abs()
    assert diff.max().item() < 2e-3, diff.max().item()
    assert (diff > 1e-6).float().mean().item() < 1e-3, (diff > 1e-6).float().mean().item()
    assert all(np.isfinite(runs["inline"][0]))


def test_tf32_convolution_mode():
    """The "tf32" precision mode of the Cnn14 convolutions (plain TF32 operands, fp32 accumulation; BASELINE configs[2..4]
    are stated in bf16): one convolution vs float64 (<= 2e-3 of the output scale, against 2e-5 in the default 3xTF32
    mode), the whole encoder vs its own fp32-mode output, and one training step's loss vs the fp32-mode step."""
    from audiocaption_b200 import _lib
    from audiocaption_b200.train_step import TrainStep
    l = _lib.lib()
    g = torch.Generator().manual_seed(0)
    B, H, W, Cin, Cout = 2, 31, 8, 256, 128
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1)
    errs = {}
    for passes in (3, 1):
        out = torch.empty(B, H, W, Cout, device=DEV)
        _lib.check(l.ac_conv3x3_p(_lib.ptr(x.to(DEV)), _lib.ptr(w.to(DEV)), None, None, _lib.ptr(out), B, H, W, Cin, Cout, 0, passes,
                                  None), "ac_conv3x3_p")
        errs[passes] = ((out.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    assert errs[3] < 2e-5 and 2e-5 < errs[1] < 2e-3, errs
    # the bf16 convolution against float64 on the SAME bf16-rounded operands: only the fp32 accumulation and the bf16 store
    sc, bi = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    xb = x.to(DEV).to(torch.bfloat16)
    outb = torch.empty(B, H, W, Cout, device=DEV, dtype=torch.bfloat16)
    wd, scd, bid = w.to(DEV), sc.to(DEV), bi.to(DEV)          # (named: the device copies must outlive the call)
    _lib.check(l.ac_conv3x3_bf16(xb.data_ptr(), _lib.ptr(wd), _lib.ptr(scd), _lib.ptr(bid), outb.data_ptr(), B, H, W,
                                 Cin, Cout, 2, None), "ac_conv3x3_bf16")
    wq = (w * sc.view(-1, 1, 1, 1)).to(torch.bfloat16).double()
    refb = (torch.nn.functional.conv2d(xb.cpu().double().permute(0, 3, 1, 2), wq, padding=1) + bi.double().view(1, -1, 1, 1))
    refb = refb.clamp_min(0).permute(0, 2, 3, 1)
    errb = ((outb.cpu().double() - refb).abs().max() / refb.abs().max()).item()
    assert errb < 6e-3, errb                           # 2^-9 relative from the bf16 store
    vocab = 4368
    m = _no_dropout(_train_model(vocab))
    wav, lens = cm.synth_wav(3, 64000, seed=21, ragged=True, varied=True, sample_rate=32000)
    inp = {"wav": wav.to(DEV), "wav_len": lens, "specaug": False}
    with torch.no_grad():
        m.eval()
        a = m.encoder.cnn(dict(inp))["attn_emb"]
        m.encoder.cnn.conv_precision = "tf32"
        b = m.encoder.cnn(dict(inp))["attn_emb"]
    rel = ((a - b).abs().max() / a.abs().max()).item()
    assert 1e-6 < rel < 1e-2, rel
    with torch.no_grad():
        m.encoder.cnn.conv_precision = "bf16"          # bf16 activations + weights: 8-bit mantissas through 12 layers
        c = m.encoder.cnn(dict(inp))
        m.encoder.cnn.conv_precision = "fp32"
    relb = ((a - c["attn_emb"]).abs().max() / a.abs().max()).item()
    assert 1e-4 < relb < 5e-2, relb
    cap, cap_len = ts.synth_captions(3, 7, vocab, seed=4)
    losses = {}
    for prec in ("fp32", "tf32", "bf16"):
        m2 = _no_dropout(_train_model(vocab))
        m2.encoder.cnn.conv_precision = prec
        step = TrainStep(m2, total_iters=1000, lr=1e-3, warmup_iters=10)
        losses[prec] = step.step({"wav": wav, "wav_len": lens, "cap": cap, "cap_len": cap_len.numpy()}, coins=None)["loss"].item()
    assert abs(losses["fp32"] - losses["tf32"]) < 2e-3 * abs(losses["fp32"]), losses
    assert abs(losses["fp32"] - losses["bf16"]) < 2e-2 * abs(losses["fp32"]), losses


def test_specaugment_stripes():
    """SpecAugment on the dB log-mel in train mode (cnn_encoder.py:352-353,424-425): the device kernel zeroes exactly the
    host-drawn frame / mel stripes (vs torch indexing), the stripes follow torchlibrosa's draw rule (width < drop width,
    inside the spectrogram, seeded by torch's generator), eval mode ignores `specaug`."""
    from audiocaption_b200 import _lib
    from audiocaption_b200.captioning.models.cnn_encoder import Cnn14Encoder, draw_specaug_stripes
    torch.manual_seed(3)
    st = draw_specaug_stripes(5, 301, 64)
    torch.manual_seed(3)
    assert (draw_specaug_stripes(5, 301, 64) == st).all() and st.shape == (5, 4, 2)
    assert (st[:, :2, 1] < 64).all() and (st[:, 2:, 1] < 8).all() and (st[:, :2, 0] + st[:, :2, 1] <= 301).all()
    assert (st[:, 2:, 0] + st[:, 2:, 1] <= 64).all()
    x = torch.randn(5, 64, 301, generator=torch.Generator().manual_seed(1)) - 40.0
    want = x.clone()
    for b in range(5):
        for k in range(2):
            want[b, :, st[b, k, 0]:st[b, k, 0] + st[b, k, 1]] = 0
            want[b, st[b, 2 + k, 0]:st[b, 2 + k, 0] + st[b, 2 + k, 1], :] = 0
    got = x.to(DEV)
    _lib.check(_lib.lib().ac_specaug_apply(_lib.ptr(got), 5, 64, 301, _lib.ptr(st.to(DEV)), 2, None), "ac_specaug_apply")
    assert (got.cpu() == want).all() and (want == 0).any()
    enc = Cnn14Encoder(sample_rate=32000, freeze=True)
    enc.load_state_dict(oc.build_state_dict(3), strict=True)
    enc = enc.to(DEV)
    enc.conv_dropout = enc.fc_dropout = 0.0
    wav, lens = cm.synth_wav(2, 64000, seed=5, varied=True, sample_rate=32000)
    inp = {"wav": wav.to(DEV), "wav_len": lens}
    with torch.no_grad():
        plain = enc.train()(dict(inp, specaug=False))["attn_emb"]
        aug = enc.train()(dict(inp, specaug=True))["attn_emb"]
        ev = enc.eval()(dict(inp, specaug=True))["attn_emb"]
    assert (ev == plain).all() and not (aug == plain).all()
