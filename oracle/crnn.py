"""Oracle restatement of the Cnn14 + bi-GRU encoder and the Cnn14Rnn-Transformer captioner.
TEST INFRASTRUCTURE (oracle/__init__.py).

Follows
  * captioning/models/rnn_encoder.py:10-49 `RnnEncoder` (nn.GRU(attn_feat_dim, hidden, num_layers, bidirectional,
    batch_first) driven through `pack_wrapper`; fc_emb = embedding_pooling(out, lens, "mean")),
  * captioning/utils/model_util.py:10-27 `sort_pack_padded_sequence` / `pad_unsort_packed_sequence` / `pack_wrapper`
    (packed-sequence semantics: clip b is processed over t < lens[b] only, the reverse direction starting at
    lens[b]-1 from a zero state; padding positions of the output are zero; the output has max(lens) frames),
  * torch.nn.GRU's documented cell:  r = s(W_ir x + b_ir + W_hr h + b_hr), z = s(W_iz x + b_iz + W_hz h + b_hz),
    n = tanh(W_in x + b_in + r*(W_hn h + b_hn)), h' = (1-z)*n + z*h, gate order (r, z, n) in the stacked weights,
  * captioning/models/crnn_trm_encoder.py:179-211 `CrnnEncoder` (cnn -> rename attn_emb/attn_emb_len -> rnn; the
    CNN's fc_emb is replaced by the RNN's),
  * eg_configs/audiocaps/waveform/cnn14rnn_trm.yaml:7-38 (3 layers, hidden 256, bidirectional; decoder
    attn_emb_dim 512, vocab 4981, no weight tying).

The explicit per-clip loops below are the restatement; tests pin them against the imported reference classes
(nn.GRU + pack_wrapper) and the golden file produced from them.
"""
import math

import torch

from . import caption_model as cm
from . import cnn14 as oc

HIDDEN, LAYERS, INPUT_DIM = 256, 3, 2048


def gru_state_dict_keys(layers=LAYERS):
    keys = []
    for l in range(layers):
        for suffix in ("", "_reverse"):
            keys += [f"network.{n}_l{l}{suffix}" for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    return keys


def build_gru_state_dict(seed=4, input_dim=INPUT_DIM, hidden=HIDDEN, layers=LAYERS):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for l in range(layers):
        din = input_dim if l == 0 else 2 * hidden
        for suffix in ("", "_reverse"):
            sd[f"network.weight_ih_l{l}{suffix}"] = torch.randn(3 * hidden, din, generator=g) * (1.5 / math.sqrt(din))
            sd[f"network.weight_hh_l{l}{suffix}"] = torch.randn(3 * hidden, hidden, generator=g) * (1.0 / math.sqrt(hidden))
            sd[f"network.bias_ih_l{l}{suffix}"] = 0.1 * torch.randn(3 * hidden, generator=g)
            sd[f"network.bias_hh_l{l}{suffix}"] = 0.1 * torch.randn(3 * hidden, generator=g)
    return sd


def _cell(x, h, w_ih, w_hh, b_ih, b_hh):
    H = h.numel()
    gi = w_ih @ x + b_ih
    gh = w_hh @ h + b_hh
    r = torch.sigmoid(gi[:H] + gh[:H])
    z = torch.sigmoid(gi[H:2 * H] + gh[H:2 * H])
    n = torch.tanh(gi[2 * H:] + r * gh[2 * H:])
    return (1 - z) * n + z * h


@torch.no_grad()
def bigru(sd, x, lens, hidden=HIDDEN, layers=LAYERS):
    """x [B, T, D], lens [B] -> [B, max(lens), 2*hidden] (zeros past each clip's length)."""
    lens = torch.as_tensor(lens)
    B, T = x.shape[0], int(lens.max())
    inp = x[:, :T]
    for l in range(layers):
        out = torch.zeros(B, T, 2 * hidden)
        for d, suffix in enumerate(("", "_reverse")):
            w_ih, w_hh = sd[f"network.weight_ih_l{l}{suffix}"], sd[f"network.weight_hh_l{l}{suffix}"]
            b_ih, b_hh = sd[f"network.bias_ih_l{l}{suffix}"], sd[f"network.bias_hh_l{l}{suffix}"]
            for b in range(B):
                h = torch.zeros(hidden)
                n = int(lens[b])
                for t in (range(n) if d == 0 else range(n - 1, -1, -1)):
                    h = _cell(inp[b, t], h, w_ih, w_hh, b_ih, b_hh)
                    out[b, t, d * hidden:(d + 1) * hidden] = h
        inp = out
    return inp


def rnn_encoder(sd, attn, attn_len):
    """rnn_encoder.py:34-49 with pooling='mean'."""
    attn_len = torch.as_tensor(attn_len)
    out = bigru(sd, attn, attn_len)
    return {"attn_emb": out, "fc_emb": oc.mean_with_lens(out, attn_len), "attn_emb_len": attn_len}


def crnn_encoder(cnn_sd, rnn_sd, wav, wav_len):
    """crnn_trm_encoder.py:203-210."""
    c = oc.forward(cnn_sd, wav, wav_len)
    return rnn_encoder(rnn_sd, c["attn_emb"], c["attn_emb_len"])


def build_decoder(seed=6, attn_emb_dim=2 * HIDDEN, vocab_size=4981):
    """Seeded 'trained-like' TransformerDecoder (cnn14rnn_trm.yaml:28-36: emb 256, 2 layers, no weight tying); same
    strengthening as caption_model.build_effb2_trm so captions vary with the audio and end at different lengths."""
    torch.manual_seed(seed)
    dec = cm.TransformerDecoder(emb_dim=256, vocab_size=vocab_size, attn_emb_dim=attn_emb_dim, tie_weights=False)
    cm._randomize_affines(dec, seed + 1000)
    with torch.no_grad():
        for l in dec.model.layers:
            for w in (l.self_attn.out_proj.weight, l.multihead_attn.out_proj.weight, l.linear2.weight,
                      l.multihead_attn.in_proj_weight):
                w.mul_(cm.DEC_BRANCH_GAIN)
        dec.word_embedding.weight[cm.PAD].mul_(cm.DEC_PAD_GAIN)
        dec.word_embedding.weight[cm.END, cm.DEC_END_DIM] += cm.DEC_END_BUMP
        # untied classifier (yaml: no weight tying) that still behaves like a trained one: the embedding plus noise,
        # so a kernel that read the embedding instead of the classifier would be caught
        g = torch.Generator().manual_seed(seed + 2000)
        dec.classifier.weight.copy_(dec.word_embedding.weight + 0.01 * torch.randn(dec.classifier.weight.shape, generator=g))
        # remove the decoder output's common-mode direction from the classifier (three rounds on seeded random
        # memory): otherwise one token wins for every clip and the argmax never depends on the audio
        dec.eval()
        cal = 0.5 * torch.tanh(torch.randn(8, 9, attn_emb_dim, generator=g))
        cal_len = torch.full((8,), 9)
        for _ in range(3):
            seq = cm.greedy_decode(dec, cal, cal_len, 6)["seq"]
            word = torch.cat([torch.full((8, 1), cm.START), seq[:, :5]], 1)
            emb = dec(word, cal, cal_len, torch.zeros(8, 6, dtype=torch.bool))["embed"].reshape(-1, 256)
            m = emb.mean(0)
            m = m / m.norm()
            dec.classifier.weight.sub_(torch.outer(dec.classifier.weight @ m, m))
    return dec.eval()


def model_state_dict(cnn_sd, rnn_sd, dec):
    """state_dict of the reference's TransformerModel(CrnnEncoder(Cnn14Encoder, RnnEncoder), TransformerDecoder)."""
    sd = {f"encoder.cnn.{k}": v for k, v in cnn_sd.items()}
    sd.update({f"encoder.rnn.{k}": v for k, v in rnn_sd.items()})
    sd.update({f"decoder.{k}": v for k, v in dec.state_dict().items()})
    return sd


@torch.no_grad()
def caption(cnn_sd, rnn_sd, dec, wav, wav_len, sample_method="greedy", beam_size=3, max_length=20, temp=1.0):
    enc = crnn_encoder(cnn_sd, rnn_sd, wav, wav_len)
    if sample_method == "beam":
        out = cm.beam_search(dec, enc["attn_emb"], enc["attn_emb_len"], beam_size, max_length, temp)
    else:
        out = cm.greedy_decode(dec, enc["attn_emb"], enc["attn_emb_len"], max_length)
    out.update(enc)
    return out
