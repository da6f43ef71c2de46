"""Stage-by-stage GPU-vs-oracle report (development aid; writes gpurun_out/debug.txt)."""
import os, sys, warnings, traceback
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import caption_model as cm, audio_frontend as fe

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "debug.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); log.write(s + "\n"); log.flush()

try:
    from audiocaption_b200.captioning.models.hf_wrapper import Effb2TrmCaptioningModel
    from audiocaption_b200.captioning.models.cnn_encoder import MelSpectrogram
    g = dict(np.load(os.path.join(ROOT, "tests/golden/effb2_trm.npz")))
    orc = cm.build_effb2_trm(int(g["seed"]), bn_stats=g["bn_stats"])
    m = Effb2TrmCaptioningModel().eval(); m.load_state_dict(orc.state_dict()); m = m.cuda()
    wav, lens = cm.synth_wav(2, 32000, seed=3, ragged=True, varied=True)
    enc_o, enc_m = orc.encoder, m.model.model.encoder
    with torch.no_grad():
        lo = enc_o.log_mel(wav)
    lm = enc_m.log_mel(wav.cuda()).cpu()
    P("logmel max abs diff", (lo - lm).abs().max().item(), "ref range", lo.min().item(), lo.max().item())
    d = (lo - lm).abs()
    P("  worst idx", np.unravel_index(d.argmax().item(), d.shape), "frame-wise max", d.amax(dim=(0, 1))[:8].tolist(), d.amax(dim=(0, 1))[-4:].tolist())
    for kind in ("cnn14",):
        c = fe.FRONTENDS[kind]
        mel = MelSpectrogram(c["sample_rate"], c["n_fft"], c["hop"], c["f_min"], c["f_max"], 64, "slaney", "slaney").cuda()
        w2, _ = cm.synth_wav(2, 64000, seed=4, varied=True, sample_rate=32000)
        window, fb = fe.frontend_buffers(kind)
        r = fe.log_mel(w2, window, fb, c["n_fft"], c["hop"], None)
        got, _ = mel(w2.cuda()); got = got.cpu()
        P("cnn14 logmel max abs diff", (r - got).abs().max().item())
    # encoder stage by stage through torch hooks on the oracle
    with torch.no_grad():
        ro = enc_o({"wav": wav, "wav_len": lens})
    rm = enc_m({"wav": wav.cuda(), "wav_len": lens, "specaug": False})
    a, r = rm["attn_emb"].cpu(), ro["attn_emb"]
    P("attn_emb rel err per clip", ((a - r).abs().amax(dim=(1, 2)) / r.abs().amax(dim=(1, 2))).tolist(), "shape", tuple(a.shape))
    P("fc_emb max diff", (rm["fc_emb"].cpu() - ro["fc_emb"]).abs().max().item())
    # decoder
    attn = torch.from_numpy(g["attn_emb"]); alen = torch.from_numpy(g["attn_emb_len"])
    dec = m.model.model.decoder
    out = dec.greedy(attn.cuda(), alen, 20, 1, 2, 0)
    P("greedy seq equal", (out["seq"].cpu().numpy() == g["greedy_seq"]).all())
    P(out["seq"].cpu().numpy()); P(g["greedy_seq"])
    P("logit0 diff", np.abs(out["logit"][:, :2].cpu().numpy() - g["greedy_logit0"]).max())
    P("embed0 diff", np.abs(out["embed"][:, :2].cpu().numpy() - g["greedy_embed0"]).max())
    ob = dec.beam_search(attn.cuda(), alen, 20, 3, 1.0, 1, 2, 0)
    P("beam3 equal rows", (ob["seq"].cpu().numpy() == g["beam3_seq"]).all(1))
    P(ob["seq"].cpu().numpy())
    # timing, B=64
    wav64, l64 = cm.synth_wav(64, 160000, seed=0)
    wd = wav64.cuda()
    for name, fn in [("encoder", lambda: enc_m({"wav": wd, "wav_len": l64, "specaug": False})),
                     ("model greedy", lambda: m(wd, l64, sample_method="greedy")),
                     ("model beam3", lambda: m(wd, l64, sample_method="beam", beam_size=3))]:
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        P(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms per batch of 64")
    e = enc_m({"wav": wd, "wav_len": l64, "specaug": False})
    for name, fn in [("logmel", lambda: enc_m.melspec_extractor(wd, want_max=True)),
                     ("greedy decode only", lambda: dec.greedy(e["attn_emb"], e["attn_emb_len"], 20, 1, 2, 0, need_logit=False))]:
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        P(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms per batch of 64")
except Exception:
    P(traceback.format_exc())
