"""Oracle restatement of the temporal Bahdanau-attention GRU caption decoder and its decode loops.
TEST INFRASTRUCTURE (oracle/__init__.py).

Follows captioning/models/hf_wrapper.py (HF release copy; training twins in captioning/models/rnn_decoder.py):
  * `Seq2SeqAttention.forward` :1377-1414 -- score = v . tanh(W [h_dec ; h_enc] + b), masked (-1e10) past src_lens,
    softmax, ctx = sum_s w_s h_enc_s;
  * `BahAttnCatFcDecoder` :1444-1499 / `TemporalBahAttnDecoder.forward` :1502-1554 -- t = 0: input embedding =
    temporal_embedding[tag], else word_embedding[word]; query = GRU state; rnn_input = [embed ; ctx_proj(ctx) ;
    fc_proj(fc_emb)]; one GRU step (1 layer, hidden 512); logit = classifier(out) WITH bias;
  * `Seq2SeqAttnModel` / `TemporalSeq2SeqAttnModel` :1557-1788 -- decode glue: the GRU state is carried between steps;
    beam search re-orders it by `prev_words_beam` (:1656-1661);
  * `CaptionModel.stepwise_forward` / `beam_search` (base.py:152-218, 254-361; HF :517-726) -- same loops as the
    Transformer model (oracle/caption_model.py restates their bookkeeping).

Eval mode (dropout = identity).  Pinned against the imported reference classes and tests/golden/temp_gru.npz.
"""
import math

import torch

from . import caption_model as cm

EMB = HID = ATT = 512
VOCAB = 4981

KEYS = ["word_embedding.weight", "classifier.weight", "classifier.bias", "model.weight_ih_l0", "model.weight_hh_l0",
        "model.bias_ih_l0", "model.bias_hh_l0", "attn.v", "attn.h2attn.weight", "attn.h2attn.bias", "fc_proj.weight",
        "fc_proj.bias", "ctx_proj.weight", "ctx_proj.bias", "temporal_embedding.weight"]


def build_state_dict(seed=8, vocab=VOCAB, attn_emb_dim=512, fc_emb_dim=512):
    """Seeded 'trained-like' decoder: strong enough recurrent / attention paths that the caption depends on the audio,
    the temporal tag and the position; <end> gets a bias that grows with the GRU state's drift so clips stop at
    different lengths."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    sd = {
        "word_embedding.weight": r(vocab, EMB) * 0.5,
        "classifier.weight": r(vocab, HID) * (2.0 / math.sqrt(HID)),
        "classifier.bias": 0.1 * r(vocab),
        "model.weight_ih_l0": r(3 * HID, 3 * EMB) * (1.5 / math.sqrt(3 * EMB)),
        "model.weight_hh_l0": r(3 * HID, HID) * (1.0 / math.sqrt(HID)),
        "model.bias_ih_l0": 0.1 * r(3 * HID),
        "model.bias_hh_l0": 0.1 * r(3 * HID),
        "attn.v": r(ATT),
        "attn.h2attn.weight": r(ATT, HID + attn_emb_dim) * (2.0 / math.sqrt(HID + attn_emb_dim)),
        "attn.h2attn.bias": 0.1 * r(ATT),
        "fc_proj.weight": r(EMB, fc_emb_dim) * (1.0 / math.sqrt(fc_emb_dim)),
        "fc_proj.bias": 0.1 * r(EMB),
        "ctx_proj.weight": r(EMB, attn_emb_dim) * (3.0 / math.sqrt(attn_emb_dim)),
        "ctx_proj.bias": 0.1 * r(EMB),
        "temporal_embedding.weight": r(4, EMB),
    }
    sd["classifier.bias"][cm.END] = 3.0        # some captions end early, at clip-dependent steps
    return {k: sd[k] for k in KEYS}


def attention(sd, h_dec, h_enc, src_lens):
    """hf_wrapper.py:1390-1414.  h_dec [N, H], h_enc [N, T, E] -> ctx [N, E], weights [N, T]."""
    N, T = h_enc.shape[:2]
    x = torch.cat((h_dec.unsqueeze(1).expand(N, T, -1), h_enc), dim=-1)
    a = torch.tanh(x @ sd["attn.h2attn.weight"].t() + sd["attn.h2attn.bias"])
    score = a @ sd["attn.v"]
    mask = torch.arange(T).unsqueeze(0) < torch.as_tensor(src_lens).view(-1, 1)
    score = score.masked_fill(~mask, -1e10)
    w = torch.softmax(score, dim=-1)
    return torch.bmm(w.unsqueeze(1), h_enc).squeeze(1), w


def gru_cell(sd, x, h):
    gi = x @ sd["model.weight_ih_l0"].t() + sd["model.bias_ih_l0"]
    gh = h @ sd["model.weight_hh_l0"].t() + sd["model.bias_hh_l0"]
    H = h.shape[1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def step(sd, t, word, tag, state, fc_emb, attn_emb, attn_emb_len):
    """One decoder call (hf_wrapper.py:1513-1554).  word [N] int64, tag [N] int64, state [N, H] -> logit, state, w."""
    embed = sd["temporal_embedding.weight"][tag] if t == 0 else sd["word_embedding.weight"][word]
    ctx, w = attention(sd, state, attn_emb, attn_emb_len)
    p_ctx = ctx @ sd["ctx_proj.weight"].t() + sd["ctx_proj.bias"]
    p_fc = fc_emb @ sd["fc_proj.weight"].t() + sd["fc_proj.bias"]
    h = gru_cell(sd, torch.cat((embed, p_ctx, p_fc), dim=-1), state)
    return h @ sd["classifier.weight"].t() + sd["classifier.bias"], h, w


@torch.no_grad()
def greedy_decode(sd, fc_emb, attn_emb, attn_emb_len, tags, max_length=20):
    B = fc_emb.size(0)
    tags = torch.as_tensor(tags).long()
    seq = torch.full((B, max_length), cm.END, dtype=torch.long)
    logit = torch.zeros(B, max_length, sd["classifier.weight"].shape[0])
    logprob = torch.zeros(B, max_length)
    state = torch.zeros(B, HID)
    unfinished = None
    for t in range(max_length):
        word = torch.full((B,), cm.START, dtype=torch.long) if t == 0 else seq[:, t - 1]
        lg, state, _ = step(sd, t, word, tags, state, fc_emb, attn_emb, attn_emb_len)
        lp, w = torch.max(torch.log_softmax(lg, dim=1), 1)
        logit[:, t], logprob[:, t], seq[:, t] = lg, lp, w
        un_t = seq[:, t] != cm.END
        unfinished = un_t if t == 0 else unfinished * un_t
        seq[:, t][~unfinished] = cm.END
        if unfinished.sum() == 0:
            break
    return {"seq": seq, "logit": logit, "sampled_logprob": logprob}


@torch.no_grad()
def beam_search(sd, fc_emb, attn_emb, attn_emb_len, tags, beam_size=3, max_length=20, temp=1.0):
    B, V = fc_emb.size(0), sd["classifier.weight"].shape[0]
    tags = torch.as_tensor(tags).long()
    lens = torch.as_tensor(attn_emb_len)
    seq_out = torch.full((B, max_length), cm.END, dtype=torch.long)
    for i in range(B):
        mem = attn_emb[i].unsqueeze(0).repeat(beam_size, 1, 1)
        fc = fc_emb[i].unsqueeze(0).repeat(beam_size, 1)
        mlen, tag = lens[i].repeat(beam_size), tags[i].repeat(beam_size)
        scores = torch.zeros(beam_size)
        state = torch.zeros(beam_size, HID)
        seq, done, nxt, prev = None, [], None, None
        for t in range(max_length):
            word = torch.full((beam_size,), cm.START, dtype=torch.long) if t == 0 else nxt
            if t > 0:
                state = state[prev]                       # hf_wrapper.py:1656-1661
            lg, state, _ = step(sd, t, word, tag, state, fc, mem, mlen)
            lp = torch.log_softmax(torch.log_softmax(lg, dim=1) / temp, dim=1)
            lp = scores.unsqueeze(1) + lp
            if t == 0:
                scores, idx = lp[0].topk(beam_size, 0, True, True)
            else:
                scores, idx = lp.view(-1).topk(beam_size, 0, True, True)
            prev = torch.div(idx, V, rounding_mode="trunc")
            nxt = idx % V
            seq = nxt.unsqueeze(1) if t == 0 else torch.cat([seq[prev], nxt.unsqueeze(1)], dim=1)
            is_end = nxt == cm.END
            if t == max_length - 1:
                is_end.fill_(True)
            for b in range(beam_size):
                if is_end[b]:
                    done.append({"seq": seq[b].clone(), "score": scores[b].item() / (t + 1)})
            scores[is_end] -= 1000
            if len(done) == beam_size:
                break
        best = sorted(done, key=lambda x: -x["score"])[0]["seq"]
        seq_out[i, :len(best)] = best
    return {"seq": seq_out}


def synth_memory(seed=1, batch=8, T=9):
    """Seeded encoder outputs shaped like the bi-GRU encoder's (values in [-1, 1], zeros past each clip's length,
    fc_emb = masked mean) + temporal tags."""
    g = torch.Generator().manual_seed(seed)
    attn = torch.tanh(torch.randn(batch, T, 512, generator=g))
    lens = torch.randint(max(1, T // 3), T + 1, (batch,), generator=g)
    lens[0] = T
    attn = attn * (torch.arange(T).view(1, T, 1) < lens.view(batch, 1, 1))
    fc = attn.sum(1) / lens.view(-1, 1)
    tags = torch.randint(0, 4, (batch,), generator=g)
    return fc, attn, lens, tags
