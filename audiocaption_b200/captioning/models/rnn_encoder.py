"""Mirror of captioning/models/rnn_encoder.py:10-49 `RnnEncoder` (HF copy hf_wrapper.py:1307-1347).

``self.network`` is a torch ``nn.GRU`` used purely as the parameter container (state_dict keys
``network.weight_ih_l0`` ... ``network.bias_hh_l2_reverse`` as in the reference); the arithmetic runs in
csrc/bigru.cu through the C ABI (tensor-core input projections + cluster-resident recurrence).  In train mode the forward
saves its gates and the backward runs csrc/bigru_train.cu (back-propagation through time on the same cluster layout),
wrapped in a ``torch.autograd.Function``; ``train_engine`` exposes the same kernels without autograd for the fused train
step (audiocaption_b200/train_step.py)."""
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
        self._engine = None

    @property
    def train_engine(self):
        if self._engine is None:
            self._engine = GruTrainEngine(self)
        return self._engine

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
            if self._handle is not None:
                _lib.lib().ac_bigru_destroy(self._handle)
                self._handle = None
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
        if self._engine is not None:
            self._engine.release()
            self._engine = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def forward(self, input_dict):
        x = input_dict["attn"]
        lens = torch.as_tensor(input_dict["attn_len"])
        require_cuda(x, "RnnEncoder.forward")
        x = x.float().contiguous()
        B, T, D = x.shape
        t_out = int(lens.max()) if B > 0 else 0       # pad_packed_sequence: max(lens) frames (lens lives on the host)
        if t_out > T or (B > 0 and int(lens.min()) < 1):
            raise _lib.AudioCaptionB200Error(f"RnnEncoder: lengths must be in 1..{T}, got {lens.tolist()}")
        params = self._tensors()
        if self.training and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            with torch.cuda.device(x.device):
                len_dev = to_device_async(lens, x.device, torch.int64)
                p_drop = float(self.dropout) if self.num_layers > 1 else 0.0
                out = _GruForwardFn.apply(x[:, :t_out].contiguous(), self, len_dev, p_drop, *params)
                fc_emb = (out.sum(1) / len_dev.unsqueeze(1).to(out.dtype))      # mean_with_lens: padded frames are zero
            return {"attn_emb": out, "fc_emb": fc_emb, "attn_emb_len": lens}
        if self.training and self.num_layers > 1 and self.dropout > 0:
            raise _lib.AudioCaptionB200Error("RnnEncoder in train mode without gradients: call .eval() for inference "
                                             "(the inter-layer dropout only exists on the training path)")
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


class GruTrainEngine:
    """bi-GRU training forward / backward on csrc/bigru_train.cu without autograd (see DecoderTrainEngine)."""

    def __init__(self, enc: "RnnEncoder"):
        self.enc = enc
        self._handle = None
        self._key = None
        self.layout_pinned = None
        self._handle_mode = None
        self._ws = Workspace()
        self._scratch = None
        self._grads = None
        self._ctx = None

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_bigru_train_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def handle(self, grads="param", need_dx=False):
        # `layout_pinned` (set by TrainStep, which owns the flat parameter / gradient buffers): skip the pointer signature
        if self._handle is not None and self.layout_pinned == (grads, bool(need_dx)) and self._handle_mode == (grads, bool(need_dx)):
            return self._handle
        enc = self.enc
        params = enc._tensors()
        for p in params:
            require_cuda(p, "RnnEncoder parameters")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.AudioCaptionB200Error("RnnEncoder training needs contiguous fp32 parameters")
        if grads == "param":
            gts = []
            for p in params:
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)
                gts.append(p.grad if p.requires_grad else None)
        else:
            if self._scratch is None or self._scratch[0].device != params[0].device:
                self._scratch = [torch.zeros_like(p) for p in params]
            gts = [g if p.requires_grad else None for p, g in zip(params, self._scratch)]
        key = (tuple(p.data_ptr() for p in params), tuple(0 if g is None else g.data_ptr() for g in gts), bool(need_dx))
        if self._handle is None or key != self._key:
            self.release()
            pp, numels, n = _lib.tensor_table([p.detach() for p in params])
            gp, _, _ = _lib.pointer_table(gts)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_bigru_train_create(pp, gp, numels, n, enc.attn_feat_dim, enc.hidden_size, enc.num_layers,
                                                        int(need_dx), _lib.current_stream(), ctypes.byref(h)),
                       "ac_bigru_train_create")
            self._handle, self._key, self._grads = h, key, gts
        self._handle_mode = (grads, bool(need_dx))     # the mode the CURRENT handle serves
        return self._handle

    def refresh(self, grads="param", need_dx=False):
        """Re-pack the live weights into the tensor core's layout on the CURRENT stream (once per optimizer step; `forward`
        does it itself unless told that the caller already has, e.g. on a side stream next to the frozen CNN)."""
        with torch.cuda.device(self.enc.network.weight_hh_l0.device):
            _lib.check(_lib.lib().ac_bigru_train_refresh(self.handle(grads, need_dx), _lib.current_stream()), "ac_bigru_train_refresh")

    def forward(self, x, len_dev, p_drop=0.0, seed=None, grads="param", need_dx=False, refresh=True):
        """x [B, T, D] fp32 cuda (T = max length), len_dev [B] int64 cuda -> out [B, T, 512].  need_dx: the backward call
        will be asked for the gradient w.r.t. x (never the case when the CNN below is frozen)."""
        l = _lib.lib()
        dev = x.device
        B, T, _ = x.shape
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else seed
        with torch.cuda.device(dev):
            h = self.handle(grads, need_dx)
            st = _lib.current_stream()
            nbytes = l.ac_bigru_train_workspace_bytes(h, B, T)
            ws = self._ws.get(nbytes, dev)
            out = torch.empty(B, T, self.enc.embed_dim, dtype=torch.float32, device=dev)
            if refresh:
                _lib.check(l.ac_bigru_train_refresh(h, st), "ac_bigru_train_refresh")
            _lib.check(l.ac_bigru_train_fwd(h, _lib.ptr(x), _lib.ptr(len_dev), B, T, p_drop, seed, _lib.ptr(out), _lib.ptr(ws),
                                            nbytes, st), "ac_bigru_train_fwd")
        self._ctx = dict(x=x, lens=len_dev, B=B, T=T, p_drop=p_drop, seed=seed, nbytes=nbytes)
        return out

    def backward(self, dout, need_dx=False):
        c = self._ctx
        if c is None:
            raise _lib.AudioCaptionB200Error("GruTrainEngine.backward without a forward")
        l = _lib.lib()
        dev = dout.device
        with torch.cuda.device(dev):
            ws = self._ws.get(c["nbytes"], dev)
            dx = torch.empty_like(c["x"]) if need_dx else None
            dout = dout.float().contiguous()
            _lib.check(l.ac_bigru_train_bwd(self._handle, _lib.ptr(c["x"]), _lib.ptr(c["lens"]), _lib.ptr(dout), c["B"], c["T"],
                                            c["p_drop"], c["seed"], _lib.ptr(dx), _lib.ptr(ws), c["nbytes"],
                                            _lib.current_stream()), "ac_bigru_train_bwd")
        return dx


class _GruForwardFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, enc, len_dev, p_drop, *params):
        eng = enc.train_engine
        out = eng.forward(x.detach(), len_dev, p_drop=p_drop, grads="scratch", need_dx=x.requires_grad)
        ctx.enc, ctx.need_dx = enc, x.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.enc.train_engine
        dx = eng.backward(dout, ctx.need_dx)
        grads = [None if g is None else g.clone() for g in eng._grads]
        return (dx, None, None, None, *grads)
