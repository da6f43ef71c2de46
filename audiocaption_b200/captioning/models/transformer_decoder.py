"""Mirror of captioning/models/transformer_decoder.py:11-103 (HF copy hf_wrapper.py:976-1068).

``TransformerDecoder`` holds the parameters under the reference's state_dict names (the
torch.nn transformer modules are used purely as parameter containers) and exposes
  * the KV-cached decode entry points of the C ABI (csrc/trm_decode.cu): ``greedy`` / ``beam_search``;
  * ``forward(input_dict)`` -- the reference's full-prefix call (transformer_decoder.py:80-103), eval or train mode,
    differentiable (a ``torch.autograd.Function`` over csrc/trm_train.cu's dense forward / backward);
  * ``train_engine`` -- the same dense kernels without autograd, used by the scheduled-sampling forward of
    ``TransformerModel`` and by the fused train step (audiocaption_b200/train_step.py).
"""
import ctypes
import math

import torch
import torch.nn as nn

from ... import _lib
from . import BaseDecoder
from ._native import Workspace, params_signature, require_cuda, to_device_async


class PositionalEncoding(nn.Module):
    """captioning/utils/model_util.py:167-186: sinusoid table stored as a frozen Parameter `pe`."""

    def __init__(self, d_model, dropout=0.1, max_len=100):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_parameter("pe", nn.Parameter(pe.unsqueeze(0).transpose(0, 1), requires_grad=False))


class TransformerDecoder(BaseDecoder):

    def __init__(self, emb_dim, vocab_size, fc_emb_dim, attn_emb_dim, dropout, freeze=False,
                 tie_weights=False, **kwargs):
        super().__init__(emb_dim, vocab_size, fc_emb_dim, attn_emb_dim, dropout=dropout, tie_weights=tie_weights)
        self.d_model = emb_dim
        self.nhead = kwargs.get("nhead", self.d_model // 64)
        self.nlayers = kwargs.get("nlayers", 2)
        self.dim_feedforward = kwargs.get("dim_feedforward", self.d_model * 4)
        self.pos_encoder = PositionalEncoding(self.d_model, dropout)
        layer = nn.TransformerDecoderLayer(d_model=self.d_model, nhead=self.nhead,
                                           dim_feedforward=self.dim_feedforward, dropout=dropout)
        self.model = nn.TransformerDecoder(layer, self.nlayers)
        self.classifier = nn.Linear(self.d_model, vocab_size, bias=False)
        if tie_weights:
            self.classifier.weight = self.word_embedding.weight
        self.attn_proj = nn.Sequential(nn.Linear(self.attn_emb_dim, self.d_model), nn.ReLU(),
                                       nn.Dropout(dropout), nn.LayerNorm(self.d_model))
        for p in self.parameters():          # init_params, transformer_decoder.py:50-53
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.freeze = freeze
        if freeze:
            for p in self.parameters():
                p.requires_grad = False
        self._ws = Workspace()
        self._handle = None
        self._sig = None
        self._engine = None

    # ---- weight pack -----------------------------------------------------------------------
    def _tensors(self):
        ts = [self.word_embedding.weight, self.pos_encoder.pe]
        for l in self.model.layers:
            ts += [l.self_attn.in_proj_weight, l.self_attn.in_proj_bias, l.self_attn.out_proj.weight,
                   l.self_attn.out_proj.bias, l.multihead_attn.in_proj_weight, l.multihead_attn.in_proj_bias,
                   l.multihead_attn.out_proj.weight, l.multihead_attn.out_proj.bias, l.linear1.weight,
                   l.linear1.bias, l.linear2.weight, l.linear2.bias, l.norm1.weight, l.norm1.bias,
                   l.norm2.weight, l.norm2.bias, l.norm3.weight, l.norm3.bias]
        ts += [self.classifier.weight, self.attn_proj[0].weight, self.attn_proj[0].bias,
               self.attn_proj[3].weight, self.attn_proj[3].bias]
        return ts

    def _dec(self):
        tensors = self._tensors()
        sig = params_signature(tensors)
        if self._handle is None or sig != self._sig:
            self._release_decode()
            ts = [t.detach().float().contiguous() for t in tensors]
            for t in ts:
                require_cuda(t, "TransformerDecoder parameters")
            ptrs, numels, n = _lib.tensor_table(ts)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_trm_create(ptrs, numels, n, self.d_model, self.nhead, self.nlayers,
                                                self.dim_feedforward, self.vocab_size, self.attn_emb_dim,
                                                self.pos_encoder.pe.shape[0], _lib.current_stream(),
                                                ctypes.byref(h)), "ac_trm_create")
            self._handle, self._sig = h, sig
        return self._handle

    def sync_decode_weights(self):
        """Re-read the live parameters into the KV-cached decode handle without re-allocating it (the fused optimizer
        writes parameters through raw pointers, which `params_signature` cannot see)."""
        handle = self._dec()
        ts = [t.detach() for t in self._tensors()]
        ptrs, _, n = _lib.tensor_table(ts)
        _lib.check(_lib.lib().ac_trm_update(handle, ptrs, n, _lib.current_stream()), "ac_trm_update")
        return handle

    def _release_decode(self):
        if self._handle is not None:
            _lib.lib().ac_trm_destroy(self._handle)
            self._handle = None

    def release(self):
        self._release_decode()
        if self._engine is not None:
            self._engine.release()
            self._engine = None

    @property
    def train_engine(self):
        if self._engine is None:
            self._engine = DecoderTrainEngine(self)
        return self._engine

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ---- decode entry points ------------------------------------------------------------------
    def _prep(self, attn_emb, attn_emb_len):
        require_cuda(attn_emb, "TransformerDecoder")
        attn_emb = attn_emb.float().contiguous()
        lens = to_device_async(torch.as_tensor(attn_emb_len), attn_emb.device, torch.int64).contiguous()
        return attn_emb, lens

    def greedy(self, attn_emb, attn_emb_len, max_length, start_idx, end_idx, pad_idx, need_logit=True):
        """KV-cached equivalent of CaptionModel.stepwise_forward(greedy) over this decoder.
        Returns dict(seq [B,L] i64, sampled_logprob [B,L], logit [B,L,V] | None, embed [B,L,D] | None),
        all on the device."""
        attn_emb, lens = self._prep(attn_emb, attn_emb_len)
        B, T, _ = attn_emb.shape
        dev = attn_emb.device
        l = _lib.lib()
        with torch.cuda.device(dev):
            dec = self._dec()
            seq = torch.empty(B, max_length, dtype=torch.int64, device=dev)
            logprob = torch.zeros(B, max_length, dtype=torch.float32, device=dev)
            logit = torch.zeros(B, max_length, self.vocab_size, device=dev) if need_logit else None
            embed = torch.zeros(B, max_length, self.d_model, device=dev) if need_logit else None
            nbytes = l.ac_trm_workspace_bytes(dec, B, T, max_length)
            ws = self._ws.get(nbytes, dev)
            _lib.check(l.ac_trm_greedy(dec, _lib.ptr(attn_emb), _lib.ptr(lens), B, T, max_length, start_idx, end_idx,
                                       pad_idx, _lib.ptr(seq), _lib.ptr(logprob), _lib.ptr(logit), _lib.ptr(embed),
                                       _lib.ptr(ws), nbytes, _lib.current_stream()), "ac_trm_greedy")
        return {"seq": seq, "sampled_logprob": logprob, "logit": logit, "embed": embed}

    def beam_search(self, attn_emb, attn_emb_len, max_length, beam_size, temp, start_idx, end_idx, pad_idx):
        """All clips' beam searches in one launch; reference bookkeeping per clip (base.py:254-361)."""
        attn_emb, lens = self._prep(attn_emb, attn_emb_len)
        B, T, _ = attn_emb.shape
        dev = attn_emb.device
        l = _lib.lib()
        with torch.cuda.device(dev):
            dec = self._dec()
            seq = torch.empty(B, max_length, dtype=torch.int64, device=dev)
            nbytes = l.ac_trm_workspace_bytes(dec, B * beam_size, T, max_length)
            ws = self._ws.get(nbytes, dev)
            _lib.check(l.ac_trm_beam(dec, _lib.ptr(attn_emb), _lib.ptr(lens), B, T, max_length, beam_size, float(temp),
                                     start_idx, end_idx, pad_idx, _lib.ptr(seq), _lib.ptr(ws), nbytes,
                                     _lib.current_stream()), "ac_trm_beam")
        return {"seq": seq}

    def scheduled_sampling_forward(self, cap, attn_emb, attn_emb_len, coins, start_idx, end_idx, pad_idx):
        """The whole train-mode `stepwise_forward` (captioning/models/base.py:152-170 + transformer_model.py:34-57) in one
        differentiable call: cap [B, L+1] ground-truth tokens, coins[t] = True where step t's prefix is the ground truth.
        Returns the reference's output dict entries: logit [B, L, V], embed [B, L, D], seq [B, L], sampled_logprob [B, L]
        (all on the device)."""
        attn_emb, lens = self._prep(attn_emb, attn_emb_len)
        words = cap[:, :-1].to(attn_emb.device, torch.int64).contiguous()
        params = [p for p in self._tensors()]
        p_drop = float(self.in_dropout.p) if self.training else 0.0
        idx = (start_idx, end_idx, pad_idx)
        if torch.is_grad_enabled() and (attn_emb.requires_grad or any(p.requires_grad for p in params)):
            logit, embed, seq, logprob = _DecoderForwardFn.apply(attn_emb, self, words, lens, None, list(coins), p_drop, idx,
                                                                 *params)
        else:
            with torch.no_grad():
                out = self.train_engine.forward(attn_emb, lens, words, coins=list(coins), p_drop=p_drop, grads="none",
                                                start_idx=start_idx, end_idx=end_idx, pad_idx=pad_idx)
            logit, embed, seq, logprob = out["logit"], out["embed"], out["seq"], out["sampled_logprob"]
        return {"logit": logit, "embed": embed, "seq": seq, "sampled_logprob": logprob}

    def forward(self, input_dict):
        """transformer_decoder.py:80-103: word [N, t+1], attn_emb [N, T, E], attn_emb_len [N], cap_padding_mask [N, t+1]
        -> {"embed" [N, t+1, d_model], "logit" [N, t+1, vocab]}.  Train mode applies the module's dropouts; gradients
        flow to the parameters and to attn_emb."""
        word = input_dict["word"]
        attn_emb, lens = self._prep(input_dict["attn_emb"], input_dict["attn_emb_len"])
        word = word.to(attn_emb.device, torch.int64).contiguous()
        pad = input_dict["cap_padding_mask"].to(attn_emb.device)
        params = [p for p in self._tensors()]
        need_grad = torch.is_grad_enabled() and (attn_emb.requires_grad or any(p.requires_grad for p in params))
        p_drop = float(self.in_dropout.p) if self.training else 0.0
        if need_grad:
            logit, embed, _, _ = _DecoderForwardFn.apply(attn_emb, self, word, lens, pad, None, p_drop, (1, 2, 0), *params)
        else:
            with torch.no_grad():
                out = self.train_engine.forward(attn_emb, lens, word, key_pad=pad, p_drop=p_drop, grads="none")
            logit, embed = out["logit"], out["embed"]
        return {"embed": embed, "logit": logit}


class DecoderTrainEngine:
    """Dense full-prefix decoder forward / backward on csrc/trm_train.cu, without autograd.

    `forward` runs the ground-truth token rows and -- for scheduled sampling -- builds and runs the model's own sampled
    rows (one KV-cached decode launch, csrc/trm_decode.cu `ac_trm_sample_forced`), then takes each step's hidden state
    from the row its coin selected.  `backward` turns d(logit) into every parameter gradient and d(attn_emb).
    Gradients are written either into the parameters' own `.grad` tensors (grads="param": the fused train step, whose
    `.grad`s are views of one flat buffer) or into a scratch buffer owned by the engine (grads="scratch": the autograd
    wrapper copies them out)."""

    def __init__(self, dec: "TransformerDecoder"):
        self.dec = dec
        self._handle = None
        self._key = None
        self.layout_pinned = None
        self._handle_mode = None
        self._ws = Workspace()
        self._dws = Workspace()
        self._scratch = None
        self._ctx = None

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_trm_train_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _grad_tensors(self, params, mode):
        if mode == "none":
            return [None] * len(params)
        if mode == "param":
            out = []
            for p in params:
                if not p.requires_grad:
                    out.append(None)
                    continue
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                out.append(p.grad)
            return out
        if self._scratch is None or self._scratch[0].device != params[0].device:
            seen = {}
            self._scratch = []
            for p in params:                       # tied weights share one gradient buffer, as they share storage
                if p.data_ptr() not in seen:
                    seen[p.data_ptr()] = torch.zeros_like(p, dtype=torch.float32)
                self._scratch.append(seen[p.data_ptr()])
        return [g if p.requires_grad else None for p, g in zip(params, self._scratch)]

    def handle(self, grads="param"):
        # `layout_pinned` (set by TrainStep, which owns the flat parameter / gradient buffers): the storage of every
        # parameter and gradient is fixed, so the per-call pointer signature (~0.2 ms of host time) is skipped
        if self._handle is not None and self.layout_pinned == grads and self._handle_mode == grads:
            return self._handle
        dec = self.dec
        params = dec._tensors()
        for p in params:
            require_cuda(p, "TransformerDecoder parameters")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.AudioCaptionB200Error("TransformerDecoder training needs contiguous fp32 parameters")
        gts = self._grad_tensors(params, grads)
        key = (tuple(p.data_ptr() for p in params), tuple(0 if g is None else g.data_ptr() for g in gts))
        if self._handle is None or key != self._key:
            self.release()
            pp, numels, n = _lib.tensor_table([p.detach() for p in params])
            gp, _, _ = _lib.pointer_table(gts)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_trm_train_create(pp, gp, numels, n, dec.d_model, dec.nhead, dec.nlayers,
                                                      dec.dim_feedforward, dec.vocab_size, dec.attn_emb_dim,
                                                      dec.pos_encoder.pe.shape[0], _lib.current_stream(), ctypes.byref(h)),
                       "ac_trm_train_create")
            self._handle, self._key = h, key
            self._grads = gts
        self._handle_mode = grads          # the mode the CURRENT handle serves (a "none"-mode call in between re-creates it)
        return self._handle

    @staticmethod
    def new_seed():
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def refresh(self, grads="param"):
        """Re-pack the live weights on the CURRENT stream (once per optimizer step; `forward` does it itself unless told
        that the caller already has)."""
        with torch.cuda.device(self.dec.classifier.weight.device):
            _lib.check(_lib.lib().ac_trm_train_refresh(self.handle(grads), _lib.current_stream()), "ac_trm_train_refresh")

    def forward(self, attn_emb, attn_len_dev, words, key_pad=None, coins=None, p_drop=0.0, seed=None, grads="param",
                start_idx=1, pad_idx=0, end_idx=2, refresh=True):
        """attn_emb [B, T, E] fp32 cuda; attn_len_dev [B] int64 cuda; words [B, L] int64 cuda (ground-truth prefix tokens
        `cap[:, :-1]`, or the given prefix for a plain decoder call); key_pad [B, L] bool (default: words == pad_idx);
        coins: None (teacher forcing) or a list of L bools, True = this step's prefix is the ground truth.
        Returns {"logit" [B, L, V], "embed" [B, L, D], "seq" [B, L] i64, "sampled_logprob" [B, L]} on the device."""
        l = _lib.lib()
        dec = self.dec
        dev = attn_emb.device
        B, T, _ = attn_emb.shape
        R, L = words.shape                    # R token rows: B, or k * B (row r attends to the memory of clip r % B: beams)
        seed = self.new_seed() if seed is None else seed
        sampled = [] if coins is None else [t for t, c in enumerate(coins) if not c]
        if R % B != 0 or (sampled and R != B):
            raise _lib.AudioCaptionB200Error(f"DecoderTrainEngine: {R} token rows for {B} clips")
        n_seq = 2 * B if sampled else R
        with torch.cuda.device(dev):
            h = self.handle(grads)
            st = _lib.current_stream()
            Vp = l.ac_trm_train_vocab_padded(h)
            nbytes = l.ac_trm_train_workspace_bytes(h, n_seq, L, B, T)
            ws = self._ws.get(nbytes, dev)
            if refresh:
                _lib.check(l.ac_trm_train_refresh(h, st), "ac_trm_train_refresh")
            _lib.check(l.ac_trm_train_memory_fwd(h, _lib.ptr(attn_emb), B, T, n_seq, L, p_drop, seed, _lib.ptr(ws), nbytes, st),
                       "ac_trm_train_memory_fwd")
            word_rows = torch.full((n_seq, L), pad_idx, dtype=torch.int64, device=dev)
            word_rows[:R] = words
            pad_rows = torch.ones(n_seq, L, dtype=torch.uint8, device=dev)
            pad_rows[:R] = (words == pad_idx) if key_pad is None else key_pad
            _lib.check(l.ac_trm_train_seq_fwd(h, _lib.ptr(word_rows), _lib.ptr(pad_rows), 0, R, n_seq, L, _lib.ptr(attn_len_dev),
                                              B, T, p_drop, seed, _lib.ptr(ws), nbytes, st), "ac_trm_train_seq_fwd")
            rows = None
            logits = torch.empty(R * L, Vp, dtype=torch.float32, device=dev)
            embed = torch.empty(R, L, dec.d_model, dtype=torch.float32, device=dev)
            if sampled:
                t_star = sampled[-1]
                word_rows[B:, 0] = start_idx
                if t_star > 0:
                    # arg-max of the ground-truth-prefix logits: what the reference's `seq` holds after a GT-coin step
                    _lib.check(l.ac_trm_train_logits(h, None, B * L, n_seq, L, B, T, _lib.ptr(logits), None, 0, _lib.ptr(ws),
                                                     nbytes, st), "ac_trm_train_logits")
                    gt_arg = torch.empty(B, L, dtype=torch.int64, device=dev)
                    _lib.check(l.ac_argmax_rows(_lib.ptr(logits), Vp, B * L, dec.vocab_size, _lib.ptr(gt_arg), None, st),
                               "ac_argmax_rows")
                    coin_dev = to_device_async(torch.tensor(coins[:t_star], dtype=torch.bool), dev)
                    forced = torch.where(coin_dev.unsqueeze(0), gt_arg[:, :t_star], torch.full_like(gt_arg[:, :t_star], -1))
                    forced = forced.contiguous()
                    dh = dec.sync_decode_weights()
                    seq_s = torch.empty(B, t_star, dtype=torch.int64, device=dev)
                    dn = l.ac_trm_workspace_bytes(dh, B, T, t_star)
                    dws = self._dws.get(dn, dev)
                    _lib.check(l.ac_trm_sample_forced(dh, _lib.ptr(attn_emb), _lib.ptr(attn_len_dev), B, T, t_star, start_idx,
                                                      end_idx, pad_idx, _lib.ptr(forced), _lib.ptr(seq_s), None, _lib.ptr(dws),
                                                      dn, st), "ac_trm_sample_forced")
                    word_rows[B:, 1:t_star + 1] = seq_s
                pad_rows[B:] = word_rows[B:] == pad_idx
                _lib.check(l.ac_trm_train_seq_fwd(h, _lib.ptr(word_rows), _lib.ptr(pad_rows), B, B, n_seq, L,
                                                  _lib.ptr(attn_len_dev), B, T, p_drop, seed, _lib.ptr(ws), nbytes, st),
                           "ac_trm_train_seq_fwd")
                sel = torch.arange(B * L, dtype=torch.int32).view(B, L)
                sel[:, sampled] += B * L
                rows = to_device_async(sel.reshape(-1).contiguous(), dev)
            _lib.check(l.ac_trm_train_logits(h, _lib.ptr(rows), R * L, n_seq, L, B, T, _lib.ptr(logits), _lib.ptr(embed), 1,
                                             _lib.ptr(ws), nbytes, st), "ac_trm_train_logits")
            seq = torch.empty(R, L, dtype=torch.int64, device=dev)
            logprob = torch.empty(R, L, dtype=torch.float32, device=dev)
            _lib.check(l.ac_argmax_rows(_lib.ptr(logits), Vp, R * L, dec.vocab_size, _lib.ptr(seq), _lib.ptr(logprob), st),
                       "ac_argmax_rows")
        self._ctx = dict(attn_emb=attn_emb, attn_len=attn_len_dev, words=word_rows, pads=pad_rows, rows=rows, n_seq=n_seq,
                         B=B, T=T, L=L, R=R, p_drop=p_drop, seed=seed, nbytes=nbytes, Vp=Vp)
        logit3 = logits.view(R, L, Vp)
        return {"logit": logit3 if Vp == dec.vocab_size else logit3[:, :, :dec.vocab_size], "logit_padded": logit3,
                "embed": embed, "seq": seq, "sampled_logprob": logprob}

    def backward(self, dlogits, need_dattn=True):
        """dlogits [B, L, Vp] (padded layout of forward()["logit_padded"], contiguous).  Parameter gradients are written to
        the tensors chosen by forward(grads=...); returns d(attn_emb) [B, T, E] or None."""
        c = self._ctx
        if c is None:
            raise _lib.AudioCaptionB200Error("DecoderTrainEngine.backward without a forward")
        l = _lib.lib()
        dev = dlogits.device
        with torch.cuda.device(dev):
            ws = self._ws.get(c["nbytes"], dev)
            dattn = torch.empty_like(c["attn_emb"]) if need_dattn else None
            _lib.check(l.ac_trm_train_bwd(self._handle, _lib.ptr(dlogits), _lib.ptr(c["rows"]), c["R"] * c["L"],
                                          _lib.ptr(c["words"]), _lib.ptr(c["pads"]), c["n_seq"], c["n_seq"], c["L"],
                                          _lib.ptr(c["attn_emb"]), _lib.ptr(c["attn_len"]), c["B"], c["T"], c["p_drop"],
                                          c["seed"], _lib.ptr(dattn), _lib.ptr(ws), c["nbytes"], _lib.current_stream()),
                       "ac_trm_train_bwd")
        return dattn


def _pad_logit_grad(dlogit, Vp):
    B, L, V = dlogit.shape
    if V == Vp and dlogit.is_contiguous():
        return dlogit
    out = torch.zeros(B, L, Vp, dtype=torch.float32, device=dlogit.device)
    out[:, :, :V] = dlogit
    return out


class _DecoderForwardFn(torch.autograd.Function):
    """Autograd face of DecoderTrainEngine: a plain full-prefix decoder call (coins None) or the whole scheduled-sampling
    forward (coins = one bool per step; idx = (start, end, pad) token ids).  Returns (logit, embed, seq, sampled_logprob);
    only `logit` is differentiable (w.r.t. the parameters and attn_emb)."""

    @staticmethod
    def forward(ctx, attn_emb, dec, word, lens, pad, coins, p_drop, idx, *params):
        eng = dec.train_engine
        out = eng.forward(attn_emb.detach(), lens, word, key_pad=pad, coins=coins, p_drop=p_drop, grads="scratch",
                          start_idx=idx[0], end_idx=idx[1], pad_idx=idx[2])
        ctx.dec, ctx.Vp = dec, out["logit_padded"].shape[-1]
        ctx.need_dattn = attn_emb.requires_grad
        ctx.mark_non_differentiable(out["embed"], out["seq"], out["sampled_logprob"])
        return out["logit"], out["embed"], out["seq"], out["sampled_logprob"]

    @staticmethod
    def backward(ctx, dlogit, _dembed, _dseq, _dlogprob):
        eng = ctx.dec.train_engine
        dattn = eng.backward(_pad_logit_grad(dlogit.float(), ctx.Vp), ctx.need_dattn)
        grads = [None if g is None else g.clone() for g in eng._grads]
        seen = set()
        for i, g in enumerate(eng._grads):          # tied weights: report the shared gradient once
            if g is not None:
                if g.data_ptr() in seen:
                    grads[i] = None
                seen.add(g.data_ptr())
        return (dattn, None, None, None, None, None, None, None, *grads)
