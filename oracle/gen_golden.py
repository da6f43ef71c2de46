"""Regenerate tests/golden/*.npz from the IMPORTED reference (build container only).

TEST INFRASTRUCTURE.  Run:  python -m oracle.gen_golden
The reference's own code (captioning/models/hf_wrapper.py at /root/reference) is executed on
CPU with seeded synthetic clips and the seeded 'trained-like' weights of
oracle.caption_model.build_effb2_trm; only ``efficientnet_pytorch`` is substituted by the
restatement in oracle/efficientnet_b2.py (the package is absent from the image).  The outputs
are the golden vectors the oracle restatement AND the CUDA path are checked against.
"""
import os

import numpy as np
import torch

from . import caption_model as cm
from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def effb2_trm(seed=1, batch=8, n=160000):
    hf = ref_import.load("captioning.models.hf_wrapper")
    orc = cm.build_effb2_trm(seed)
    ref = hf.Effb2TrmCaptioningModel(hf.Effb2TrmConfig()).eval()
    ref.load_state_dict(orc.state_dict(), strict=True)
    wav, lens = cm.synth_wav(batch, n, seed=7, ragged=True, varied=True)
    base = {"wav": wav, "wav_len": lens, "specaug": False, "mode": "inference", "temp": 1.0, "max_length": 20}
    with torch.no_grad():
        enc = ref.model.model.encoder
        lms = enc.db_transform(enc.melspec_extractor(wav))
        g = ref.model(dict(base, sample_method="greedy"))
        b3 = ref(wav, lens, sample_method="beam", beam_size=3)
        b2 = ref(wav, lens, sample_method="beam", beam_size=2, max_length=12)
    # which rows are numerically well-conditioned?  re-decode (reference code) with the audio
    # memory perturbed by 1e-3 relative noise; rows whose caption changes sit on a near-tie and
    # are excluded from exact-match checks (fp32 summation order alone can flip them).
    rd = ref.model.model
    stable = {k: torch.ones(batch, dtype=torch.bool) for k in ("greedy", "beam3", "beam2")}
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for _ in range(8):
            e = {"attn_emb": g["attn_emb"] * (1 + 1e-3 * torch.randn(g["attn_emb"].shape, generator=gen)),
                 "attn_emb_len": g["attn_emb_len"], "fc_emb": g["fc_emb"]}
            pg = rd.forward_decoder({"mode": "inference", "sample_method": "greedy", "max_length": 20, "temp": 1.0}, dict(e))
            p3 = rd.forward_decoder({"mode": "inference", "sample_method": "beam", "beam_size": 3, "max_length": 20, "temp": 1.0}, dict(e))
            p2 = rd.forward_decoder({"mode": "inference", "sample_method": "beam", "beam_size": 2, "max_length": 12, "temp": 1.0}, dict(e))
            stable["greedy"] &= (pg["seq"] == g["seq"]).all(1)
            stable["beam3"] &= (p3["seq"] == b3).all(1)
            stable["beam2"] &= (p2["seq"] == b2).all(1)
    print("stable rows:", {k: v.tolist() for k, v in stable.items()})
    np.savez_compressed(
        os.path.join(OUT, "effb2_trm.npz"),
        seed=seed, batch=batch, n_samples=n, wav_seed=7,
        bn_stats=cm.bn_stats_vector(orc.encoder).numpy(),
        wav_len=lens.numpy(),
        lms_stride=np.array([1, 7]),                       # lms[:, ::1, ::7] keeps the file small
        lms=lms[:, :, ::7].numpy(),
        lms_max=lms.max().item(),
        attn_emb=g["attn_emb"].numpy(),  attn_emb_len=g["attn_emb_len"].numpy(), fc_emb=g["fc_emb"].numpy(),
        greedy_seq=g["seq"].numpy(), greedy_logit0=g["logit"][:, :2].numpy(),
        greedy_logprob=g["sampled_logprob"].numpy(), greedy_embed0=g["embed"][:, :2].numpy(),
        beam3_seq=b3.numpy(), beam2_len12_seq=b2.numpy(),
        greedy_stable=stable["greedy"].numpy(), beam3_stable=stable["beam3"].numpy(),
        beam2_stable=stable["beam2"].numpy(),
    )
    print("effb2_trm greedy\n", g["seq"], "\nbeam3\n", b3, "\nbeam2/12\n", b2)


def cnn14(seed=3, batch=3, n=64000):
    """Golden vectors of the reference's own `Cnn14Encoder` (captioning/models/cnn_encoder.py:326-464) on seeded
    weights (oracle.cnn14.build_state_dict) and 2 s ragged clips."""
    from . import cnn14 as oc
    ce = ref_import.load("captioning.models.cnn_encoder")
    ref = ce.Cnn14Encoder(sample_rate=32000).eval()
    ref.load_state_dict(oc.build_state_dict(seed), strict=True)
    wav, lens = cm.synth_wav(batch, n, seed=9, ragged=True, varied=True, sample_rate=32000)
    with torch.no_grad():
        lms = ref.db_transform(ref.melspec_extractor(wav))
        out = ref({"wav": wav, "wav_len": lens, "specaug": False})
    print("cnn14 feat lengths", out["attn_emb_len"].tolist(), "attn_emb", tuple(out["attn_emb"].shape))
    np.savez_compressed(
        os.path.join(OUT, "cnn14.npz"), seed=seed, batch=batch, n_samples=n, wav_seed=9, wav_len=lens.numpy(),
        lms=lms[:, :, ::5].numpy(), attn_emb=out["attn_emb"].numpy(), fc_emb=out["fc_emb"].numpy(),
        attn_emb_len=out["attn_emb_len"].numpy())


def build_reference_cnn14rnn_trm():
    """The reference's Cnn14Rnn-Transformer exactly as eg_configs/audiocaps/waveform/cnn14rnn_trm.yaml:7-38 builds it."""
    ce = ref_import.load("captioning.models.cnn_encoder")
    re_ = ref_import.load("captioning.models.rnn_encoder")
    cte = ref_import.load("captioning.models.crnn_trm_encoder")
    td = ref_import.load("captioning.models.transformer_decoder")
    tm = ref_import.load("captioning.models.transformer_model")
    enc = cte.CrnnEncoder(ce.Cnn14Encoder(sample_rate=32000),
                          re_.RnnEncoder(spec_dim=-1, fc_feat_dim=2048, attn_feat_dim=2048, bidirectional=True,
                                         hidden_size=256, dropout=0.5, num_layers=3),
                          freeze_cnn=True, freeze_cnn_bn=True)
    dec = td.TransformerDecoder(emb_dim=256, vocab_size=4981, fc_emb_dim=512, attn_emb_dim=512, nlayers=2, dropout=0.2)
    return tm.TransformerModel(enc, dec).eval()


def cnn14rnn_trm(cnn_seed=3, rnn_seed=4, dec_seed=6, batch=4, n=96000):
    """Golden vectors of the reference's Cnn14Rnn-Transformer (greedy + beam 3) on seeded weights and 3 s ragged clips."""
    from . import cnn14 as oc
    from . import crnn
    ref = build_reference_cnn14rnn_trm()
    dec = crnn.build_decoder(dec_seed)
    ref.load_state_dict(crnn.model_state_dict(oc.build_state_dict(cnn_seed), crnn.build_gru_state_dict(rnn_seed), dec), strict=True)
    wav, lens = cm.synth_wav(batch, n, seed=13, ragged=True, varied=True, sample_rate=32000)
    base = {"wav": wav, "wav_len": lens, "specaug": False, "mode": "inference", "temp": 1.0, "max_length": 20}
    with torch.no_grad():
        g = ref(dict(base, sample_method="greedy"))
        b3 = ref(dict(base, sample_method="beam", beam_size=3))
    stable = {k: torch.ones(batch, dtype=torch.bool) for k in ("greedy", "beam3")}
    gen = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for _ in range(8):
            e = {"attn_emb": g["attn_emb"] * (1 + 1e-3 * torch.randn(g["attn_emb"].shape, generator=gen)),
                 "attn_emb_len": g["attn_emb_len"], "fc_emb": g["fc_emb"]}
            pg = ref.forward_decoder({"mode": "inference", "sample_method": "greedy", "max_length": 20, "temp": 1.0}, dict(e))
            p3 = ref.forward_decoder({"mode": "inference", "sample_method": "beam", "beam_size": 3, "max_length": 20, "temp": 1.0}, dict(e))
            stable["greedy"] &= (pg["seq"] == g["seq"]).all(1)
            stable["beam3"] &= (p3["seq"] == b3["seq"]).all(1)
    print("cnn14rnn_trm lens", g["attn_emb_len"].tolist(), "stable", {k: v.tolist() for k, v in stable.items()})
    print("greedy\n", g["seq"], "\nbeam3\n", b3["seq"])
    np.savez_compressed(
        os.path.join(OUT, "cnn14rnn_trm.npz"), cnn_seed=cnn_seed, rnn_seed=rnn_seed, dec_seed=dec_seed, batch=batch,
        n_samples=n, wav_seed=13, wav_len=lens.numpy(), attn_emb=g["attn_emb"].numpy(), fc_emb=g["fc_emb"].numpy(),
        attn_emb_len=g["attn_emb_len"].numpy(), greedy_seq=g["seq"].numpy(), greedy_logit0=g["logit"][:, :2].numpy(),
        beam3_seq=b3["seq"].numpy(), greedy_stable=stable["greedy"].numpy(), beam3_stable=stable["beam3"].numpy())


def temp_gru(seed=8, mem_seed=1, batch=8, T=9):
    """Golden token ids of the reference's TemporalSeq2SeqAttnModel + TemporalBahAttnDecoder (hf_wrapper.py:1502-1788) on
    seeded decoder weights and seeded encoder outputs: greedy, beam 3, beam 4."""
    import torch.nn as nn
    from . import bah_decoder as bd
    hf = ref_import.load("captioning.models.hf_wrapper")
    dec = hf.TemporalBahAttnDecoder(emb_dim=512, vocab_size=4981, fc_emb_dim=512, attn_emb_dim=512, rnn_type="GRU",
                                    num_layers=1, d_model=512, dropout=0.5)
    dec.load_state_dict(bd.build_state_dict(seed), strict=True)
    model = hf.TemporalSeq2SeqAttnModel(nn.Identity(), dec).eval()
    fc, attn, lens, tags = bd.synth_memory(mem_seed, batch, T)

    def run(attn_, fc_, method, beam=None):
        d = {"mode": "inference", "sample_method": method, "max_length": 20, "temp": 1.0, "temporal_tag": tags}
        if beam:
            d["beam_size"] = beam
        with torch.no_grad():
            return model.forward_decoder(d, {"fc_emb": fc_, "attn_emb": attn_, "attn_emb_len": lens})
    g, b3, b4 = run(attn, fc, "greedy"), run(attn, fc, "beam", 3), run(attn, fc, "beam", 4)
    stable = {k: torch.ones(batch, dtype=torch.bool) for k in ("greedy", "beam3", "beam4")}
    gen = torch.Generator().manual_seed(23)
    for _ in range(8):
        pa = attn * (1 + 1e-3 * torch.randn(attn.shape, generator=gen))
        pf = fc * (1 + 1e-3 * torch.randn(fc.shape, generator=gen))
        stable["greedy"] &= (run(pa, pf, "greedy")["seq"] == g["seq"]).all(1)
        stable["beam3"] &= (run(pa, pf, "beam", 3)["seq"] == b3["seq"]).all(1)
        stable["beam4"] &= (run(pa, pf, "beam", 4)["seq"] == b4["seq"]).all(1)
    print("temp_gru stable", {k: v.tolist() for k, v in stable.items()})
    print("greedy\n", g["seq"], "\nbeam4\n", b4["seq"])
    np.savez_compressed(
        os.path.join(OUT, "temp_gru.npz"), seed=seed, mem_seed=mem_seed, batch=batch, T=T, lens=lens.numpy(), tags=tags.numpy(),
        greedy_seq=g["seq"].numpy(), greedy_logit0=g["logit"][:, :2].numpy(), greedy_logprob=g["sampled_logprob"].numpy(),
        beam3_seq=b3["seq"].numpy(), beam4_seq=b4["seq"].numpy(), greedy_stable=stable["greedy"].numpy(),
        beam3_stable=stable["beam3"].numpy(), beam4_stable=stable["beam4"].numpy())


def sed(seed=12, batch=8, n=96000):
    """Golden outputs of the reference's Cnn8rnnSedModel (hf_wrapper.py:1791-1859) on seeded weights and 3 s clips:
    temporal tags, a sample of the segment-wise probabilities, and which clips keep their tag under input noise."""
    from . import cnn14 as oc
    from . import sed as so
    hf = ref_import.load("captioning.models.hf_wrapper")
    ref = hf.Cnn8rnnSedModel(classes_num=447).eval()
    ref.load_state_dict(so.build_state_dict(seed), strict=True)
    wav, lens = cm.synth_wav(batch, n, seed=31, ragged=True, varied=True, sample_rate=32000)
    lms = oc.log_mel(oc.build_state_dict(3), wav)
    with torch.no_grad():
        seg = ref.forward_prob(lms)["segmentwise_output"]
        tags = ref(lms)
    stable = torch.ones(batch, dtype=torch.bool)
    gen = torch.Generator().manual_seed(29)
    with torch.no_grad():
        for _ in range(6):       # a tag that survives 1e-2 dB of input noise cannot flip on fp32 reordering
            stable &= torch.tensor(ref(lms + 1e-2 * torch.randn(lms.shape, generator=gen))) == torch.tensor(tags)
    print("sed tags", tags, "stable", stable.tolist())
    np.savez_compressed(os.path.join(OUT, "sed.npz"), seed=seed, batch=batch, n_samples=n, wav_seed=31, tags=np.array(tags),
                        seg=seg[:, ::3, ::7].numpy(), stable=stable.numpy())


def reference_train_step(ss_ratio=0.6, coin_seed=22, batch=4, n=96000, cap_max=9, lr=2e-4, vocab=4981):
    """ONE training step of the reference's own classes (python_scripts/train_eval/run.py:116-127 on
    TransformerModel(CrnnEncoder(Cnn14Encoder, RnnEncoder), TransformerDecoder) + LabelSmoothingLoss + Adam) with every
    dropout disabled: returns (inputs, loss, grads, updated parameters)."""
    import random
    from . import cnn14 as oc
    from . import crnn
    from . import train_step as ts
    ref = build_reference_cnn14rnn_trm() if vocab == 4981 else None
    assert ref is not None
    loss_mod = ref_import.load("captioning.losses.loss")
    dec = crnn.build_decoder(6, vocab_size=vocab)
    sd = crnn.model_state_dict(oc.build_state_dict(3), crnn.build_gru_state_dict(4), dec)
    ref.load_state_dict(sd, strict=True)
    ref.train()
    for m in ref.modules():                                  # dropout off everywhere (deterministic parity)
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, (torch.nn.MultiheadAttention, torch.nn.GRU)):
            m.dropout = 0.0
    wav, lens = cm.synth_wav(batch, n, seed=13, ragged=True, varied=True, sample_rate=32000)
    cap, cap_len = ts.synth_captions(batch, cap_max, vocab, seed=1)
    real_dropout = torch.nn.functional.dropout
    torch.nn.functional.dropout = lambda x, p=0.5, training=True, inplace=False: x     # the frozen CNN's F.dropout calls
    try:
        random.seed(coin_seed)
        out = ref({"mode": "train", "wav": wav, "wav_len": lens, "specaug": False, "cap": cap, "cap_len": cap_len.numpy(),
                   "ss_ratio": ss_ratio})
    finally:
        torch.nn.functional.dropout = real_dropout
    out["tgt"], out["tgt_len"] = cap[:, 1:], torch.as_tensor(cap_len.numpy() - 1)
    loss = loss_mod.LabelSmoothingLoss(smoothing=0.1)(out)
    opt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=lr, weight_decay=1e-6)
    opt.zero_grad()
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in ref.named_parameters() if v.requires_grad}
    gnorm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    opt.step()
    new_params = {k: v.detach().clone() for k, v in ref.named_parameters() if v.requires_grad}
    random.seed(coin_seed)
    coins = [random.random() < ss_ratio for _ in range(cap.size(1) - 1)]
    return dict(wav=wav, wav_len=lens, cap=cap, cap_len=cap_len, coins=coins, loss=loss.detach(), grads=grads, gnorm=gnorm,
                new_params=new_params, seq=out["seq"], logit=out["logit"].detach(), lr=lr, ss_ratio=ss_ratio,
                coin_seed=coin_seed, vocab=vocab, batch=batch, n=n, cap_max=cap_max)


def train_step():
    """tests/golden/train_step.npz: loss, total gradient norm, per-parameter gradient norms and leading entries, and the
    per-parameter update norms of one reference training step (reference_train_step above)."""
    r = reference_train_step()
    names = sorted(r["grads"])
    init = {k: v for k, v in build_reference_state(r["vocab"]).items()}
    np.savez_compressed(
        os.path.join(OUT, "train_step.npz"), names=np.array(names), loss=r["loss"].numpy(), gnorm=r["gnorm"].numpy(),
        coins=np.array(r["coins"]), cap=r["cap"].numpy(), cap_len=r["cap_len"].numpy(), wav_len=r["wav_len"].numpy(),
        seq=r["seq"].numpy(), lr=r["lr"], ss_ratio=r["ss_ratio"], coin_seed=r["coin_seed"], vocab=r["vocab"], batch=r["batch"],
        n_samples=r["n"], cap_max=r["cap_max"], logit_head=r["logit"][:, :, :16].numpy(),
        grad_norms=np.array([r["grads"][k].norm().item() for k in names]),
        grad_heads=np.stack([np.resize(r["grads"][k].flatten()[:16].numpy(), 16) for k in names]),
        update_norms=np.array([(r["new_params"][k] - init[k]).norm().item() for k in names]))
    print("train_step loss", float(r["loss"]), "gnorm", float(r["gnorm"]), "coins", r["coins"], "seq\n", r["seq"])


def build_reference_state(vocab=4981):
    from . import cnn14 as oc
    from . import crnn
    return crnn.model_state_dict(oc.build_state_dict(3), crnn.build_gru_state_dict(4), crnn.build_decoder(6, vocab_size=vocab))


def eg_configs():
    """The `model:` sections of the two Cnn14Rnn-Transformer training YAMLs (eg_configs/{audiocaps,clotho_v2}/waveform/
    cnn14rnn_trm.yaml:7-38) as fixtures for the factory tests (tests/test_boundary_cpu.py)."""
    import yaml
    os.makedirs(os.path.join(OUT, "eg_configs"), exist_ok=True)
    for ds in ("audiocaps", "clotho_v2"):
        with open(os.path.join(ref_import.REF_ROOT, "eg_configs", ds, "waveform", "cnn14rnn_trm.yaml")) as f:
            cfg = yaml.load(f, Loader=yaml.FullLoader)
        with open(os.path.join(OUT, "eg_configs", f"{ds}_cnn14rnn_trm.yaml"), "w") as f:
            f.write(f"# model section of the reference's eg_configs/{ds}/waveform/cnn14rnn_trm.yaml (oracle/gen_golden.py eg_configs)\n")
            yaml.safe_dump({"model": cfg["model"]}, f, default_flow_style=False)


if __name__ == "__main__":
    import sys
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["effb2_trm", "cnn14", "cnn14rnn_trm", "temp_gru", "sed", "eg_configs", "train_step"]
    if "eg_configs" in which:
        eg_configs()
    if "train_step" in which:
        train_step()
    if "effb2_trm" in which:
        effb2_trm()
    if "cnn14" in which:
        cnn14()
    if "cnn14rnn_trm" in which:
        cnn14rnn_trm()
    if "temp_gru" in which:
        temp_gru()
    if "sed" in which:
        sed()
