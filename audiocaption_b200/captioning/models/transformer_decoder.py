"""Mirror of captioning/models/transformer_decoder.py:11-103 (HF copy hf_wrapper.py:976-1068).

``TransformerDecoder`` holds the parameters under the reference's state_dict names (the
torch.nn transformer modules are used purely as parameter containers) and exposes the
KV-cached decode entry points of the C ABI (csrc/trm_decode.cu).  Eval mode only.
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
            self.release()
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

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_trm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ---- decode entry points ------------------------------------------------------------------
    def _prep(self, attn_emb, attn_emb_len):
        require_cuda(attn_emb, "TransformerDecoder")
        if self.training:
            raise NotImplementedError("the B200 decoder implements the eval-mode (inference) path")
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

    def forward(self, input_dict):
        raise NotImplementedError(
            "full-prefix TransformerDecoder.forward is the training-time call; inference goes through "
            "greedy()/beam_search() (KV-cached).  Training on B200 is not built yet.")
