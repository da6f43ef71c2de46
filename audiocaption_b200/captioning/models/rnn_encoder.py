"""Mirror of captioning/models/rnn_encoder.py:10-49 `RnnEncoder` (HF copy hf_wrapper.py:1307-1347).

``self.network`` is a torch ``nn.GRU`` used purely as the parameter container (state_dict keys
``network.weight_ih_l0`` ... ``network.bias_hh_l2_reverse`` as in the reference); the arithmetic runs in
csrc/bigru.cu through the C ABI (tensor-core input projections + cluster-resident recurrence).  Eval mode only."""
import ctypes

import torch
import torch.nn as nn

from ... import _lib
from . import BaseEncoder
from ._native import Workspace, params_signature, require_cuda, to_device_async


class RnnEncoder(BaseEncoder):

    def __init__(self, spec_dim, fc_feat_dim, attn_feat_dim, pooling="mean", **kwargs):
        super().__init__(spec_dim, fc_feat_dim, attn_feat_dim)
        self.pooling = pooling
        self.hidden_size = kwargs.get("hidden_size", 512)
        self.bidirectional = kwargs.get("bidirectional", False)
        self.num_layers = kwargs.get("num_layers", 1)
        self.dropout = kwargs.get("dropout", 0.2)
        self.rnn_type = kwargs.get("rnn_type", "GRU")
        self.in_bn = kwargs.get("in_bn", False)
        if self.rnn_type != "GRU" or not self.bidirectional or self.hidden_size != 256 or self.in_bn:
            raise NotImplementedError("the B200 path implements the bidirectional GRU encoder with hidden_size 256 "
                                      "(eg_configs/*/waveform/cnn14rnn_trm.yaml); other variants are not built")
        if pooling != "mean":
            raise NotImplementedError(f"pooling {pooling!r} is not built (mean is)")
        self.embed_dim = self.hidden_size * 2
        self.network = nn.GRU(attn_feat_dim, self.hidden_size, num_layers=self.num_layers, bidirectional=True,
                              dropout=self.dropout, batch_first=True)
        self._ws = Workspace()
        self._handle = None
        self._sig = None

    def _tensors(self):
        ts = []
        for l in range(self.num_layers):
            for suffix in ("", "_reverse"):
                ts += [getattr(self.network, f"{n}_l{l}{suffix}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return ts

    def _net(self):
        tensors = self._tensors()
        sig = params_signature(tensors)
        if self._handle is None or sig != self._sig:
            self.release()
            ts = [t.detach().float().contiguous() for t in tensors]
            for t in ts:
                require_cuda(t, "RnnEncoder parameters")
            ptrs, numels, n = _lib.tensor_table(ts)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_bigru_create(ptrs, numels, n, self.attn_feat_dim, self.hidden_size, self.num_layers,
                                                  _lib.current_stream(), ctypes.byref(h)), "ac_bigru_create")
            self._handle, self._sig = h, sig
        return self._handle

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_bigru_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def forward(self, input_dict):
        x = input_dict["attn"]
        lens = torch.as_tensor(input_dict["attn_len"])
        require_cuda(x, "RnnEncoder.forward")
        if self.training:
            raise NotImplementedError("the B200 GRU encoder implements the eval-mode (inference) path")
        x = x.float().contiguous()
        B, T, D = x.shape
        t_out = int(lens.max()) if B > 0 else 0       # pad_packed_sequence: max(lens) frames (lens lives on the host)
        if t_out > T or (B > 0 and int(lens.min()) < 1):
            raise _lib.AudioCaptionB200Error(f"RnnEncoder: lengths must be in 1..{T}, got {lens.tolist()}")
        l = _lib.lib()
        dev = x.device
        with torch.cuda.device(dev):
            net = self._net()
            len_dev = to_device_async(lens, dev, torch.int64)
            out = torch.empty(B, t_out, self.embed_dim, device=dev, dtype=torch.float32)
            fc_emb = torch.empty(B, self.embed_dim, device=dev, dtype=torch.float32)
            nbytes = l.ac_bigru_workspace_bytes(net, B, T)
            ws = self._ws.get(nbytes, dev)
            _lib.check(l.ac_bigru_fwd(net, _lib.ptr(x), _lib.ptr(len_dev), B, T, t_out, _lib.ptr(out), _lib.ptr(ws), nbytes,
                                      _lib.current_stream()), "ac_bigru_fwd")
            _lib.check(l.ac_masked_mean(_lib.ptr(out), _lib.ptr(len_dev), B, t_out, self.embed_dim, _lib.ptr(fc_emb),
                                        _lib.current_stream()), "ac_masked_mean")
        return {"attn_emb": out, "fc_emb": fc_emb, "attn_emb_len": lens}
