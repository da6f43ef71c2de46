"""Oracle restatement of the EffB2-Transformer captioning model.  TEST INFRASTRUCTURE.

Restates (fp32, CPU, plain PyTorch) the reference's
captioning/models/hf_wrapper.py:

* ``EfficientNetB2.forward``                    :287-315  (mel, dB, backbone, lengths, fc_emb)
* ``TransformerDecoder``                        :976-1068 (attn_proj, embedding*sqrt(d)+PE,
                                                           nn.TransformerDecoder, classifier)
* ``CaptionModel.stepwise_forward`` / greedy    :517-617
* ``CaptionModel.beam_search``                  :619-726
* ``TransformerModel.prepare_*decoder_input``   :868-920
* ``Effb2TrmCaptioningModel``                   :1144-1181 (state-dict prefix ``model.model.``)

``nn.TransformerDecoder`` itself is the library the reference calls, so the oracle
calls it too; everything around it is restated.  Module nesting mirrors the reference
so that ``state_dict`` keys are interchangeable.  Pinned by oracle/gen_golden.py
against the imported reference (decoder + decode loops + mel exactly; the EfficientNet
body only through the restated third-party package -- see efficientnet_b2.py).
"""
import math

import torch
import torch.nn as nn

from . import audio_frontend as fe
from . import efficientnet_b2 as effnet

PAD, START, END = 0, 1, 2


def length_mask(lens, max_len):
    lens = torch.as_tensor(lens)
    return torch.arange(max_len).unsqueeze(0) < lens.view(-1, 1)


def mean_with_lens(x, lens):
    lens = torch.as_tensor(lens)
    m = length_mask(lens, x.size(1)).to(x.device).unsqueeze(-1)
    return (x * m).sum(1) / lens.view(-1, 1).to(x.device)


# --------------------------------------------------------------------------- encoder
class _Spectrogram(nn.Module):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer("window", fe.hann_periodic(n_fft))


class _MelScale(nn.Module):
    def __init__(self, fb):
        super().__init__()
        self.register_buffer("fb", fb)


class _MelSpectrogram(nn.Module):
    def __init__(self, kind):
        super().__init__()
        c = fe.FRONTENDS[kind]
        self.cfg = c
        window, fb = fe.frontend_buffers(kind)
        self.spectrogram = _Spectrogram(c["n_fft"])
        self.mel_scale = _MelScale(fb)

    def forward(self, wav, top_db):
        return fe.log_mel(wav, self.spectrogram.window, self.mel_scale.fb,
                          self.cfg["n_fft"], self.cfg["hop"], top_db)


class _EffiNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.eff_net = effnet.EfficientNet()
        self.eff_net._change_in_channels(1)

    def forward(self, x):                      # [B, F, T]
        x = self.eff_net.extract_features(x.unsqueeze(1))
        return x.mean(dim=2).transpose(1, 2)   # 'b c f t -> b t c' mean over f


class EfficientNetB2Encoder(nn.Module):
    hop_length = 160
    downsample_ratio = 32

    def __init__(self):
        super().__init__()
        self.melspec_extractor = _MelSpectrogram("effb2")
        self.backbone = _EffiNet()

    def log_mel(self, wav):
        return self.melspec_extractor(wav, 120.0)

    def forward(self, input_dict):
        x = self.log_mel(input_dict["wav"])
        attn_emb = self.backbone(x)
        wl = torch.as_tensor(input_dict["wav_len"])
        fl = torch.div(wl, self.hop_length, rounding_mode="floor") + 1
        fl = torch.div(fl, self.downsample_ratio, rounding_mode="floor")
        return {"fc_emb": mean_with_lens(attn_emb, fl), "attn_emb": attn_emb, "attn_emb_len": fl}


# --------------------------------------------------------------------------- decoder
class _PositionalEncoding(nn.Module):
    def __init__(self, d_model, max_len=100):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.pe = nn.Parameter(pe.unsqueeze(1), requires_grad=False)   # [max_len, 1, d]


class TransformerDecoder(nn.Module):
    def __init__(self, emb_dim=256, vocab_size=4981, attn_emb_dim=1408, dropout=0.2,
                 nlayers=2, nhead=None, dim_feedforward=None, tie_weights=True):
        super().__init__()
        self.d_model, self.vocab_size = emb_dim, vocab_size
        self.word_embedding = nn.Embedding(vocab_size, emb_dim)
        self.pos_encoder = _PositionalEncoding(emb_dim)
        layer = nn.TransformerDecoderLayer(emb_dim, nhead or emb_dim // 64,
                                           dim_feedforward or emb_dim * 4, dropout)
        self.model = nn.TransformerDecoder(layer, nlayers)
        self.classifier = nn.Linear(emb_dim, vocab_size, bias=False)
        if tie_weights:
            self.classifier.weight = self.word_embedding.weight
        self.attn_proj = nn.Sequential(nn.Linear(attn_emb_dim, emb_dim), nn.ReLU(),
                                       nn.Dropout(dropout), nn.LayerNorm(emb_dim))
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, word, attn_emb, attn_emb_len, cap_padding_mask):
        dev = attn_emb.device                  # (device-agnostic so that bench.py can time the same code as GPU eager)
        word, cap_padding_mask = word.to(dev), cap_padding_mask.to(dev)
        mem = self.attn_proj(attn_emb).transpose(0, 1)
        emb = self.word_embedding(word) * math.sqrt(self.d_model)
        emb = emb.transpose(0, 1)
        emb = emb + self.pos_encoder.pe[:emb.size(0)]
        L = emb.size(0)
        causal = torch.full((L, L), float("-inf"), device=dev).triu(1)
        mem_pad = ~length_mask(attn_emb_len, attn_emb.size(1)).to(dev)
        out = self.model(emb, mem, tgt_mask=causal, tgt_key_padding_mask=cap_padding_mask,
                         memory_key_padding_mask=mem_pad).transpose(0, 1)
        return {"embed": out, "logit": self.classifier(out)}


def greedy_decode(decoder, attn_emb, attn_emb_len, max_length=20):
    """stepwise_forward(mode=inference, sample_method=greedy): full-prefix recompute per step,
    rows that emitted <end> stay <end>, stop when every row has finished."""
    B = attn_emb.size(0)
    dev = attn_emb.device
    seq = torch.full((B, max_length), END, dtype=torch.long)           # host tensor, as base.py:122
    logit = torch.zeros(B, max_length, decoder.vocab_size, device=dev)
    logprob = torch.zeros(B, max_length)
    embed = torch.zeros(B, max_length, decoder.d_model, device=dev)
    steps = 0
    unfinished = None
    for t in range(max_length):
        start = torch.full((B, 1), START, dtype=torch.long)
        word = start if t == 0 else torch.cat((start, seq[:, :t]), dim=-1)
        out = decoder(word, attn_emb, attn_emb_len, word == PAD)
        lg, em = out["logit"][:, -1], out["embed"][:, -1]
        lp, w = torch.max(torch.log_softmax(lg, dim=1), 1)
        logit[:, t], embed[:, t], logprob[:, t], seq[:, t] = lg, em, lp.cpu(), w.cpu()
        steps = t + 1
        un_t = seq[:, t] != END
        unfinished = un_t if t == 0 else unfinished * un_t
        seq[:, t][~unfinished] = END
        if unfinished.sum() == 0:
            break
    return {"seq": seq, "logit": logit, "sampled_logprob": logprob, "embed": embed, "steps": steps}


def beam_search(decoder, attn_emb, attn_emb_len, beam_size=3, max_length=20, temp=1.0, n_best_size=None):
    """Per-sample beam search with the reference's exact bookkeeping: double log-softmax,
    step-0 top-k from row 0, finished beams stay in the beam with a -1000 penalty, stop on
    ``len(done) == beam_size`` (equality), score = logprob / (t+1), stable best-first sort."""
    B, V = attn_emb.size(0), decoder.vocab_size
    seq_out = torch.full((B, max_length), END, dtype=torch.long)
    nbest_out = torch.full((B, n_best_size or 1, max_length), END, dtype=torch.long)     # base.py:259-263 (`n_best`)
    nbest_score = torch.full((B, n_best_size or 1), float("-inf"))
    for i in range(B):
        mem = attn_emb[i].unsqueeze(0).repeat(beam_size, 1, 1)
        mlen = torch.as_tensor(attn_emb_len)[i].repeat(beam_size)
        scores = torch.zeros(beam_size)
        seq, done = None, []
        for t in range(max_length):
            start = torch.full((beam_size, 1), START, dtype=torch.long)
            word = start if t == 0 else torch.cat((start, seq), dim=-1)
            lg = decoder(word, mem, mlen, word == PAD)["logit"][:, -1]
            lp = torch.log_softmax(torch.log_softmax(lg, dim=1) / temp, dim=1)
            lp = scores.unsqueeze(1) + lp
            if t == 0:
                scores, idx = lp[0].topk(beam_size, 0, True, True)
            else:
                scores, idx = lp.view(-1).topk(beam_size, 0, True, True)
            prev = torch.div(idx, V, rounding_mode="trunc")
            nxt = idx % V
            seq = nxt.unsqueeze(1) if t == 0 else torch.cat([seq[prev], nxt.unsqueeze(1)], dim=1)
            is_end = nxt == END
            if t == max_length - 1:
                is_end.fill_(True)
            for b in range(beam_size):
                if is_end[b]:
                    done.append({"seq": seq[b].clone(), "score": scores[b].item() / (t + 1)})
            scores[is_end] -= 1000
            if len(done) == beam_size:
                break
        ranked = sorted(done, key=lambda x: -x["score"])
        best = ranked[0]["seq"]
        seq_out[i, :len(best)] = best
        for j, d in enumerate(ranked[:n_best_size or 1]):                # base.py:351-358
            nbest_out[i, j, :len(d["seq"])] = d["seq"]
            nbest_score[i, j] = d["score"]
    out = {"seq": seq_out}
    if n_best_size:
        out["n_best_seq"], out["n_best_score"] = nbest_out, nbest_score
    return out


def sample_next_word(logit, method, temp, generator=None):
    """base.py:214-252.  Returns (word [N], sampled_logprob [N]).  `generator` seeds the random draws (the reference uses
    torch's global generator)."""
    logprob = torch.log_softmax(logit, dim=1)
    if method == "greedy":
        lp, word = torch.max(logprob, 1)
        return word, lp
    if method == "gumbel":
        u = torch.rand(logprob.shape, generator=generator)
        y = logprob - torch.log(-torch.log(u + 1e-20) + 1e-20)
        word = torch.log_softmax(y / temp, dim=-1).argmax(1)
        return word, logprob.gather(1, word.unsqueeze(-1)).squeeze(1)
    logprob = logprob / temp
    if method.startswith("top"):
        top_num = float(method[3:])
        if 0 < top_num < 1:                                   # nucleus: keep the smallest prefix reaching top_num
            probs = torch.softmax(logit, dim=1)
            sp, si = torch.sort(probs, descending=True, dim=1)
            keep = sp.cumsum(1) < top_num
            keep = torch.cat([torch.ones_like(keep[:, :1]), keep[:, :-1]], 1)
            sp = sp * keep.to(sp)
            sp = sp / sp.sum(1, keepdim=True)
            logprob = logprob.scatter(1, si, sp.log())
        else:                                                 # top-k
            k = int(top_num)
            tk, ti = torch.topk(logprob, k, dim=1)
            logprob = torch.full_like(logprob, float("-inf")).scatter(1, ti, tk)
    p = torch.softmax(logprob, dim=1)
    word = torch.multinomial(p, 1, generator=generator).squeeze(1)
    return word, logprob.gather(1, word.unsqueeze(-1)).squeeze(1)


# --------------------------------------------------------------------------- full model
class _TransformerModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = EfficientNetB2Encoder()
        self.decoder = TransformerDecoder()


class _KdWrapper(nn.Module):
    def __init__(self, shared_dim=1024, tchr_dim=768):
        super().__init__()
        self.model = _TransformerModel()
        self.stdnt_proj = nn.Linear(1408, shared_dim)
        self.tchr_proj = nn.Linear(tchr_dim, shared_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))


class Effb2TrmOracle(nn.Module):
    """state_dict-compatible with hf_wrapper.Effb2TrmCaptioningModel."""

    def __init__(self):
        super().__init__()
        self.model = _KdWrapper()

    @property
    def encoder(self):
        return self.model.model.encoder

    @property
    def decoder(self):
        return self.model.model.decoder

    @torch.no_grad()
    def forward(self, audio, audio_length, sample_method="beam", beam_size=3, max_length=20, temp=1.0):
        enc = self.encoder({"wav": audio, "wav_len": audio_length})
        if sample_method == "beam":
            out = beam_search(self.decoder, enc["attn_emb"], enc["attn_emb_len"], beam_size, max_length, temp)
        else:
            out = greedy_decode(self.decoder, enc["attn_emb"], enc["attn_emb_len"], max_length)
        out.update(enc)
        return out


def _calibrate_batchnorm(enc: "EfficientNetB2Encoder", seed: int):
    """Set every BatchNorm's running statistics to the float64 batch statistics of four
    synthetic clips, the way a trained network's statistics match its activations.  Without
    this a random-init EfficientNet is degenerate (the signal dies within a few blocks and
    the output depends on the BN betas only), which would let kernel bugs hide."""
    cal, _ = synth_wav(4, 160000, seed, varied=True)
    bns = [m for m in enc.modules() if isinstance(m, nn.BatchNorm2d)]
    enc.double()
    for m in bns:
        m.reset_running_stats()
        m.momentum = None        # cumulative average == the batch statistics of the single pass
        m.train()
    with torch.no_grad():
        enc.backbone(enc.log_mel(cal.double()))
    for m in bns:
        m.eval()
    enc.float()


def _randomize_affines(module: nn.Module, seed: int):
    """Non-trivial BN/LN gammas, betas and linear biases so folding bugs cannot hide."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
                n = m.weight.numel()
                m.weight.copy_(torch.rand(n, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
        for n, p in module.named_parameters():
            if n.endswith("bias") and p.dim() == 1 and "norm" not in n and "bn" not in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def bn_stats_vector(module: nn.Module) -> torch.Tensor:
    return torch.cat([torch.cat([m.running_mean, m.running_var]) for m in module.modules()
                      if isinstance(m, nn.BatchNorm2d)])


def load_bn_stats_vector(module: nn.Module, vec: torch.Tensor):
    o = 0
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(vec[o:o + n]); m.running_var.copy_(vec[o + n:o + 2 * n])
            o += 2 * n
    assert o == vec.numel()


def build_effb2_trm(seed: int = 1, calibrate: bool = True, bn_stats=None) -> Effb2TrmOracle:
    """Seeded random-init EffB2-Trm (no checkpoint is reachable offline), made
    'trained-like': calibrated BN statistics, random affines/biases, and a decoder whose
    attention/FFN branches are strong enough that greedy/beam captions vary with the audio
    and terminate at different lengths (exercises <end>, early stop and the beam rules)."""
    torch.manual_seed(seed)
    m = Effb2TrmOracle()
    if bn_stats is not None:      # stats calibrated elsewhere (golden fixture): bit-identical weights
        load_bn_stats_vector(m.encoder, torch.as_tensor(bn_stats))
    elif calibrate:
        _calibrate_batchnorm(m.encoder, seed + 3000)
    _randomize_affines(m, seed + 1000)
    dec = m.decoder
    with torch.no_grad():
        for l in dec.model.layers:
            for w in (l.self_attn.out_proj.weight, l.multihead_attn.out_proj.weight,
                      l.linear2.weight, l.multihead_attn.in_proj_weight):
                w.mul_(DEC_BRANCH_GAIN)
        dec.word_embedding.weight[PAD].mul_(DEC_PAD_GAIN)
        # <end> gets weight on a slowly varying positional-encoding dimension so that rows
        # terminate at different, position-dependent steps
        dec.word_embedding.weight[END, DEC_END_DIM] += DEC_END_BUMP
    return m.eval()


DEC_BRANCH_GAIN, DEC_PAD_GAIN, DEC_END_DIM, DEC_END_BUMP = 4.0, 1.6, 70, 1.5


def synth_wav(batch: int, n_samples: int, seed: int = 0, ragged: bool = False, varied: bool = False,
              sample_rate: int = 16000):
    """Synthetic clips.  Default: 0.1*randn (the benchmark input).  varied=True:
    per-clip gain over 2.5 decades + an AM chirp, so clips differ audibly from each other.
    ragged=True zero-pads rows to random lengths (row 0 keeps the full length)."""
    g = torch.Generator().manual_seed(seed)
    if not varied:
        wav = 0.1 * torch.randn(batch, n_samples, generator=g)
    else:
        t = torch.arange(n_samples, dtype=torch.float64) / sample_rate
        wav = torch.zeros(batch, n_samples)
        for b in range(batch):
            u = torch.rand(3, generator=g).double()
            gain = 10.0 ** (-2.5 * u[0])
            f0 = 100.0 + 3000.0 * u[1]
            am = 0.5 + 0.5 * torch.sin(2 * math.pi * (0.2 + 2.0 * u[2]) * t)
            tone = am * torch.sin(2 * math.pi * f0 * t * (1.0 + 0.02 * t * b / max(batch, 1)))
            wav[b] = (gain * (0.3 * torch.randn(n_samples, generator=g).double() + tone)).float()
    lens = torch.full((batch,), n_samples, dtype=torch.long)
    if ragged and batch > 1:
        lens = torch.randint(n_samples // 2, n_samples + 1, (batch,), generator=g)
        lens[0] = n_samples
        wav = wav * (torch.arange(n_samples).unsqueeze(0) < lens.unsqueeze(1))
    return wav, lens
