"""Debug aid (GPU box): (1) decoder gradient error vs a float64 oracle, next to the fp32 oracle's own error;
(2) per-parameter finite-difference check of the train engines with and without dropout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import copy
import torch
from oracle import caption_model as cm, crnn, train_step as ts
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_train_gpu as T

DEV = "cuda:0"


def f64_vs_f32(B=3, Tm=31, L=21, vocab=4368):
    dec, orc = T._decoder_pair(vocab)
    dec.train()
    for mod in dec.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    mem, lens, word = T._decoder_inputs(B, Tm, L, vocab, seed=7 * B + L)
    pad = word == cm.PAD
    R = torch.randn(B, L, vocab, generator=torch.Generator().manual_seed(1)) * (~pad).unsqueeze(-1)
    res = {}
    for name, o, dt in (("f32", orc, torch.float32), ("f64", copy.deepcopy(orc).double(), torch.float64)):
        for p in o.parameters():
            p.requires_grad_(True)
        o.pos_encoder.pe.requires_grad_(False)
        o.zero_grad()
        m = mem.to(dt).clone().requires_grad_(True)
        (o(word, m, lens, pad)["logit"] * R.to(dt)).sum().backward()
        res[name] = {"mem": m.grad.double(), **{k: v.grad.double() for k, v in o.named_parameters() if v.grad is not None}}
    mem_dev = mem.to(DEV).requires_grad_(True)
    out = dec({"word": word, "attn_emb": mem_dev, "attn_emb_len": lens, "cap_padding_mask": pad})
    (out["logit"] * R.to(DEV)).sum().backward()
    mine = {"mem": mem_dev.grad.double().cpu(), **{k: v.grad.double().cpu() for k, v in dec.named_parameters() if v.grad is not None}}
    print(f"{'param':55s} {'mine-f64':>10s} {'f32-f64':>10s}  (relative to max |f64|)")
    for k in res["f64"]:
        sc = res["f64"][k].abs().max().item() + 1e-30
        print(f"{k:55s} {(mine[k] - res['f64'][k]).abs().max().item() / sc:10.2e} {(res['f32'][k] - res['f64'][k]).abs().max().item() / sc:10.2e}")


def fd(p_dec, p_rnn):
    from audiocaption_b200.captioning.losses.loss import ls_ce_fwd_bwd
    from audiocaption_b200.captioning.models.rnn_encoder import RnnEncoder
    vocab = 520
    dec, _ = T._decoder_pair(vocab)
    dec.train()
    rnn = RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True, hidden_size=256, dropout=0.5, num_layers=3)
    rnn.load_state_dict(crnn.build_gru_state_dict(4), strict=True)
    rnn = rnn.to(DEV).train()
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(4, 7, 2048, generator=g) * 0.3).to(DEV)
    len_dev = torch.tensor([7, 5, 7, 2]).to(DEV)
    cap, cap_len = ts.synth_captions(4, 9, vocab, seed=3)
    cap = cap.to(DEV)
    tl = (cap_len - 1).to(DEV)

    def run(seed=11):
        mem = rnn.train_engine.forward(x, len_dev, p_drop=p_rnn, seed=seed, grads="param")
        out = dec.train_engine.forward(mem, len_dev, cap[:, :-1].contiguous(), coins=None, p_drop=p_dec, seed=seed + 1, grads="param")
        loss, dl = ls_ce_fwd_bwd(out["logit_padded"][:, :, :vocab], cap[:, 1:], tl, 0.1)
        return loss, dl
    _, dl = run()
    dmem = dec.train_engine.backward(dl)
    rnn.train_engine.backward(dmem)
    torch.cuda.synchronize()
    print(f"--- p_dec {p_dec} p_rnn {p_rnn}")
    torch.manual_seed(0)
    for mod in (dec, rnn):
        for k, p in mod.named_parameters():
            if p.grad is None or p.grad.abs().max() == 0:
                continue
            gr = p.grad.clone()
            d = torch.randn_like(p)
            analytic = (gr * d).sum().item()
            eps = 1e-3
            vals = []
            for sgn in (1.0, -1.0):
                with torch.no_grad():
                    p.add_(d, alpha=sgn * eps)
                vals.append(run()[0].item())
                with torch.no_grad():
                    p.add_(d, alpha=-sgn * eps)
            numeric = (vals[0] - vals[1]) / (2 * eps)
            flag = "" if abs(numeric - analytic) < 0.05 * abs(analytic) + 2e-3 else "   <<<<"
            print(f"{k:50s} analytic {analytic:10.4f} numeric {numeric:10.4f}{flag}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["f64", "fd"]
    if "f64" in which:
        f64_vs_f32()
    if "fd" in which:
        fd(0.0, 0.0)
        fd(0.2, 0.0)
        fd(0.0, 0.5)
