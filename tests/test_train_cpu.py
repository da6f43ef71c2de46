"""CPU tests of the training-step oracle (oracle/train_step.py) and of the host-side schedules: the oracle against the
golden vectors of the reference's own training step, against the imported reference (build container), and the explicit
Adam / clip / learning-rate restatements against torch."""
import math
import os
import random
import warnings

import numpy as np
import pytest
import torch

from oracle import cnn14 as oc
from oracle import crnn
from oracle import ref_import
from oracle import train_step as ts

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "train_step.npz")))


@pytest.fixture(scope="module")
def oracle_step(golden):
    from oracle import caption_model as cm
    g = golden
    wav, lens = cm.synth_wav(int(g["batch"]), int(g["n_samples"]), seed=13, ragged=True, varied=True, sample_rate=32000)
    assert (lens.numpy() == g["wav_len"]).all()
    cap, cap_len = ts.synth_captions(int(g["batch"]), int(g["cap_max"]), int(g["vocab"]), seed=1)
    assert (cap.numpy() == g["cap"]).all() and (cap_len.numpy() == g["cap_len"]).all()
    dec = crnn.build_decoder(6, vocab_size=int(g["vocab"]))
    return ts.train_step(oc.build_state_dict(3), crnn.build_gru_state_dict(4), dec, wav, lens, cap, cap_len,
                         [bool(c) for c in g["coins"]], float(g["lr"]))


def test_golden_coins_are_the_reference_rng_sequence(golden):
    random.seed(int(golden["coin_seed"]))
    coins = [random.random() < float(golden["ss_ratio"]) for _ in range(golden["cap"].shape[1] - 1)]
    assert coins == [bool(c) for c in golden["coins"]]
    assert not all(coins) and any(coins)               # both branches of transformer_model.py:44-51 are exercised


def test_oracle_train_step_matches_golden(golden, oracle_step):
    """loss, sampled tokens, logits, every trainable gradient and the Adam update of the reference's own step."""
    g, o = golden, oracle_step
    assert abs(float(o["loss"]) - float(g["loss"])) < 2e-5
    assert (o["output"]["seq"].numpy() == g["seq"]).all()
    assert np.abs(o["output"]["logit"][:, :, :16].numpy() - g["logit_head"]).max() < 1e-4
    assert abs(float(o["grad_norm"]) - float(g["gnorm"])) < 1e-3 * float(g["gnorm"])
    names = [str(n) for n in g["names"]]
    assert sorted(o["grads"]) == names
    init = {f"encoder.rnn.{k}": v for k, v in crnn.build_gru_state_dict(4).items()}
    init.update({f"decoder.{k}": v for k, v in crnn.build_decoder(6, vocab_size=int(g["vocab"])).state_dict().items()})
    for i, k in enumerate(names):
        gr = o["grads"][k]
        assert abs(gr.norm().item() - g["grad_norms"][i]) <= 1e-4 * g["grad_norms"][i] + 1e-7, k
        head = np.resize(gr.flatten()[:16].numpy(), 16)
        assert np.abs(head - g["grad_heads"][i]).max() <= 2e-5 * max(np.abs(g["grad_heads"][i]).max(), 1e-3), k
        upd = (o["new_params"][k] - init[k]).norm().item()
        assert abs(upd - g["update_norms"][i]) <= 0.02 * g["update_norms"][i] + 1e-7, k     # first Adam step ~ lr * sign(g)


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_oracle_train_step_matches_imported_reference(oracle_step):
    from oracle import gen_golden as gg
    r = gg.reference_train_step()
    o = oracle_step
    assert abs(float(o["loss"]) - float(r["loss"])) < 1e-5
    assert (o["output"]["seq"] == r["seq"]).all()
    assert (o["output"]["logit"] - r["logit"]).abs().max() < 1e-4
    for k, gr in r["grads"].items():
        err = (gr - o["grads"][k]).abs().max().item() / (gr.abs().max().item() + 1e-12)
        assert err < 1e-4, (k, err)
        big = gr.abs() > 1e-5                  # the first Adam step is lr * g / (|g| + eps): ill-conditioned for tiny g
        if big.any():
            assert (r["new_params"][k] - o["new_params"][k])[big].abs().max() < 0.02 * r["lr"], k


def test_adam_and_clip_restatements_match_torch():
    torch.manual_seed(0)
    ps = [torch.randn(37, 5), torch.randn(11)]
    gs = [[torch.randn_like(p) * s for p in ps] for s in (3.0, 0.01, 1.0)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref, lr=5e-4, weight_decay=1e-6)
    mine = [p.clone() for p in ps]
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in ps]
    for step, g in enumerate(gs, 1):
        for r, gi in zip(ref, g):
            r.grad = gi.clone()
        total_ref = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        total, coef = ts.clip_coef(g, 1.0)
        assert abs(float(total) - float(total_ref)) < 1e-5 * float(total_ref)
        for i in range(len(ps)):
            mine[i], m, v = ts.adam_update(mine[i], g[i] * coef, state[i][0], state[i][1], step, 5e-4, weight_decay=1e-6)
            state[i] = (m, v)
            assert (mine[i] - ref[i].detach()).abs().max() < 1e-7


def test_lr_schedule_package_vs_oracle_vs_reference_formula():
    from audiocaption_b200.captioning.utils.lr_scheduler import ExponentialDecayScheduler, exponential_decay_lr
    total, warm, base, final = 1000, 200, 5e-4, 5e-7
    for k in (1, 2, 50, 199, 200, 201, 500, 1000):
        a = exponential_decay_lr(k, base, final, total, warm)
        b = ts.exponential_decay_lr(k, base, final, total, warm)
        assert abs(a - b) <= 1e-12 + 1e-9 * b
    assert exponential_decay_lr(200, base, final, total, warm) == base
    assert abs(exponential_decay_lr(1000, base, final, total, warm) - final) < 1e-12
    # the torch-protocol wrapper: constructor performs the first step; run.py:105 steps before the optimizer
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=base)
    sched = ExponentialDecayScheduler(opt, total_iters=total, final_lrs=final, warmup_iters=warm)
    assert abs(opt.param_groups[0]["lr"] - base / warm) < 1e-12
    for it in range(1, 300):
        sched.step()
        assert abs(opt.param_groups[0]["lr"] - ts.exponential_decay_lr(it + 1, base, final, total, warm)) < 1e-12
    if ref_import.available():               # the reference's own closed form (its constructor does not run on torch 2.11)
        mod = ref_import.load("captioning.utils.lr_scheduler")
        obj = object.__new__(mod.ExponentialDecayScheduler)
        obj.total_iters, obj.warmup_iters, obj.base_lrs, obj.final_lrs = total, warm, [base], [final]
        obj.bases = [(final / base) ** (1 / (total - warm))]
        for k in (1, 2, 199, 200, 201, 777):
            obj._step_count = k
            assert abs(obj._get_closed_form_lr()[0] - exponential_decay_lr(k, base, final, total, warm)) < 1e-15


def test_ss_ratio_schedule():
    assert abs(ts.ss_ratio_after(1000, 1000) - 0.7) < 1e-9
    assert abs(ts.ss_ratio_after(1, 1000) - (1 - 0.3 / 1000)) < 1e-12
    assert abs(ts.ss_ratio_after(10, 10, mode="exponential") - 0.01) < 1e-9


def test_label_smoothing_loss_restatement():
    torch.manual_seed(1)
    logit = torch.randn(3, 5, 17)
    tgt = torch.randint(0, 17, (3, 5))
    lens = torch.tensor([5, 3, 1])
    got = ts.label_smoothing_loss(logit, tgt, lens, 0.1)
    logp = logit.log_softmax(-1)
    want = 0.0
    for b in range(3):
        for t in range(int(lens[b])):
            td = torch.full((17,), 0.1 / 16)
            td[tgt[b, t]] = 0.9
            want += -(td * logp[b, t]).sum()
    assert abs(float(got) - float(want / lens.sum())) < 1e-6
    if ref_import.available():
        ref = ref_import.load("captioning.losses.loss").LabelSmoothingLoss(smoothing=0.1)
        assert abs(float(ref({"logit": logit, "tgt": tgt, "tgt_len": lens})) - float(got)) < 1e-6
