"""The training step of python_scripts/train_eval/run.py:77-148 (`Runner._train_epoch` body) as one fused device pipeline.

Per iteration, in the reference's order: scheduled-sampling ratio update (:55-65), learning-rate step (:105,
captioning/utils/lr_scheduler.py:5-45), zero_grad, forward (:21-47 -> captioning/models/base.py:48-137), label-smoothing
loss (captioning/losses/loss.py:51-74), backward, `clip_grad_norm_(model.parameters(), max_grad_norm)` (:125-126),
Adam step (:127), skipped when the loss is NaN (:123).  Here:

  * the trainable parameters (bi-GRU + Transformer decoder, 10.7 M) are re-homed as views of ONE flat fp32 buffer, their
    `.grad`s as views of a second one; the backward kernels write gradients straight into it (no autograd, no zero_grad:
    every gradient is overwritten each step);
  * data parallel: ONE `all_reduce(sum)` over the flat gradient buffer (NCCL over NVLink / NVSwitch,
    python_scripts/train_eval/run_ddp.py:105-107 wraps the model in DDP for the same exchange), the 1/world scale is
    folded into the optimizer kernel;
  * global-norm clip + Adam (+ L2 weight decay, bias correction, NaN skip) run as two launches over the flat buffers
    (csrc/train_ops.cu `ac_clip_adam`); nothing synchronises with the host -- the loss stays on the device until read.

The frozen Cnn14 runs forward only (BatchNorm in eval mode, dropout active: `freeze_cnn`, `freeze_cnn_bn`)."""
import random

import contextlib
import ctypes
import os

import torch

from . import _lib
from .captioning.models._native import to_device_async
from .captioning.utils.lr_scheduler import exponential_decay_lr


def flatten_trainable(model, device=None):
    """Re-home every trainable parameter of `model` as a view of ONE flat fp32 buffer and its `.grad` as a view of a
    second one (each tensor 128-byte aligned: gradients are written by TMA stores).  Returns (params, flat_param,
    flat_grad).  After this, one `all_reduce(flat_grad)` is the whole data-parallel gradient exchange and one fused kernel
    over the two buffers is the whole optimizer step; `state_dict()` keeps working (the views ARE the parameters)."""
    params, seen = [], set()
    for p in model.parameters():
        if p.requires_grad and id(p) not in seen:
            seen.add(id(p))
            params.append(p)
    offs, total = [], 0
    for p in params:
        if p.dtype != torch.float32:
            raise _lib.AudioCaptionB200Error("flatten_trainable needs fp32 master parameters")
        offs.append(total)
        total += (p.numel() + 31) // 32 * 32
    device = params[0].device if device is None else device
    flat_param = torch.zeros(total, dtype=torch.float32, device=device)
    flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
    for p, o in zip(params, offs):
        view = flat_param[o:o + p.numel()].view_as(p)
        view.copy_(p.data)
        p.data = view
        p.grad = flat_grad[o:o + p.numel()].view_as(p)
    return params, flat_param, flat_grad


def allreduce_gradients(flat_grad, group=None):
    """The data-parallel exchange (python_scripts/train_eval/run_ddp.py:105-107 gets it from DistributedDataParallel): one
    sum all-reduce over the flat gradient; returns the factor (1 / world) the optimizer kernel folds in."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return 1.0
    world = torch.distributed.get_world_size(group)
    if world > 1:
        torch.distributed.all_reduce(flat_grad, group=group)
    return 1.0 / world


def allreduce_bucket_async(bucket, group=None):
    """Start the sum all-reduce of one contiguous slice of the flat gradient (it runs on the collective's own stream, behind
    everything enqueued on the current stream so far); returns the work handle, or None when not distributed."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return None
    if torch.distributed.get_world_size(group) == 1 or bucket.numel() == 0:
        return None
    return torch.distributed.all_reduce(bucket, group=group, async_op=True)


def bucket_boundary(model, params):
    """Offset (in floats of the flat buffers) where the decoder's parameters start, when the flat order is
    [encoder..., decoder...] (it is: `model.parameters()` yields the encoder first) -- else None.  The decoder's gradients
    are complete before the GRU's backward pass starts, so their all-reduce overlaps it."""
    dec_ids = {id(p) for p in model.decoder.parameters()}
    off, boundary = 0, None
    for p in params:
        if id(p) in dec_ids:
            if boundary is None:
                boundary = off
        elif boundary is not None:
            return None                      # an encoder parameter after a decoder one: no clean split
        off += (p.numel() + 31) // 32 * 32
    return boundary


class TrainStep:
    """`step(batch)` = one optimizer step of the reference's training loop on the Cnn14Rnn-Transformer captioner
    (eg_configs/*/waveform/cnn14rnn_trm.yaml): TransformerModel(CrnnEncoder(Cnn14Encoder, RnnEncoder), TransformerDecoder)."""

    def __init__(self, model, total_iters, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, max_grad_norm=1.0,
                 smoothing=0.1, final_lr=5e-7, warmup_iters=None, ss_mode="linear", ss_final_ratio=0.7, use_ss=True,
                 process_group=None, specaug=False):
        from .captioning.models.crnn_trm_encoder import CrnnEncoder
        from .captioning.models.transformer_model import TransformerModel
        if not isinstance(model, TransformerModel) or not isinstance(model.encoder, CrnnEncoder):
            raise NotImplementedError("TrainStep is built for TransformerModel(CrnnEncoder(...), TransformerDecoder)")
        if any(p.requires_grad for p in model.encoder.cnn.parameters()):
            raise NotImplementedError("the CNN has no backward pass on the B200 path: build it with freeze_cnn: True")
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.AudioCaptionB200Error("TrainStep needs the model on a CUDA device (no CPU fallback)")
        self.total_iters = int(total_iters)
        self.warmup_iters = self.total_iters // 5 if warmup_iters is None else int(warmup_iters)     # run.py:249-251
        self.base_lr, self.final_lr = float(lr), float(final_lr)
        self.betas, self.eps, self.weight_decay = betas, float(eps), float(weight_decay)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.smoothing = float(smoothing)
        self.use_ss, self.ss_mode, self.ss_final_ratio = use_ss, ss_mode, float(ss_final_ratio)
        self.ss_ratio = 1.0
        self.specaug = bool(specaug)                 # SpecAugment on the log-mel (the YAML's top-level `specaug:`)
        # persistent CTAs the look-ahead encoder's convolutions may hold (prefetch(encode=True)); the rest of the 148 SMs
        # stay free for the trainable chain of the step in flight
        self.cnn_sms = int(os.environ.get("AC_TRAIN_CNN_SMS", "84"))
        self.iteration = 0
        self.group = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self._flatten()
        l = _lib.lib()
        with torch.cuda.device(self.device):
            self._adam_ws = torch.empty(l.ac_clip_adam_workspace_bytes(), dtype=torch.uint8, device=self.device)
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.grad_norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.lr = 0.0
        model.train()

    # ---- flat parameter / gradient storage -------------------------------------------------------------------------
    def _flatten(self):
        self.params, self.flat_param, self.flat_grad = flatten_trainable(self.model, self.device)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.n_trainable = sum(p.numel() for p in self.params)
        self.dec_offset = bucket_boundary(self.model, self.params)      # flat_grad[dec_offset:] = the decoder's gradients
        # the flat buffers fix every parameter's and gradient's storage, and nothing writes the frozen CNN: the engines skip
        # their per-call pointer / version signatures (~1 ms of host time per step); re-run _flatten() after moving the model
        m = self.model
        for eng, pin in ((m.encoder.rnn.train_engine, ("param", False)), (m.decoder.train_engine, "param")):
            eng.release()
            eng.layout_pinned = pin
        m.encoder.cnn.release()
        m.encoder.cnn.weights_pinned = True

    # ---- schedules (host) ---------------------------------------------------------------------------------------------
    def _update_ss_ratio(self):
        """run.py:55-65"""
        if not self.use_ss:
            return
        if self.ss_mode == "exponential":
            self.ss_ratio *= 0.01 ** (1.0 / self.total_iters)
        elif self.ss_mode == "linear":
            self.ss_ratio -= (1.0 - self.ss_final_ratio) / self.total_iters
        else:
            raise Exception(f"mode {self.ss_mode} not supported")

    # ---- input staging / look-ahead ---------------------------------------------------------------------------------------
    def prefetch(self, batch, encode=True):
        """Look-ahead for the NEXT step; returns the staged batch to hand to `step`.

        1. Upload: the host -> device copy of the waveforms and captions runs on a copy stream (asynchronous when the host
           tensors are pinned), as a DataLoader prefetcher does for the reference's loop.  Two staging slots; a slot is
           overwritten only after the step that read it has finished with it.  Device-resident batches skip this.
        2. `encode`: the FROZEN CNN's forward pass of that batch (log-mel, Cnn14 with its train-mode dropouts) runs on a
           second stream.  It does not depend on the optimizer step in flight, and the trainable part of a step (bi-GRU,
           decoder, backward passes: ~250 latency-bound launches of a few CTAs each) leaves most of the chip idle, so the
           tensor-core-bound encoder of batch i+1 fills it while step i runs.  The two run on DISJOINT SM sets (CUDA green
           contexts, `_make_look_ahead_streams`): `cnn_sms` SMs (AC_TRAIN_CNN_SMS, default 84) for the encoder, whose
           convolutions size their persistent grid to that set, the other 64 for the trainable chain.  Results equal the
           inline schedule's (only the order in which dropout seeds are drawn changes)."""
        dev = self.device
        m = self.model
        with torch.cuda.device(dev), torch.no_grad():
            user = torch.cuda.current_stream()
            staged = dict(batch)
            ready = None
            if not batch["wav"].is_cuda:
                if getattr(self, "_copy_stream", None) is None:
                    self._copy_stream = torch.cuda.Stream(device=dev)
                    self._stage = [{"free": None, "enc": None, "wav": None, "cap": None},
                                   {"free": None, "enc": None, "wav": None, "cap": None}]
                    self._stage_i = 0
                k = self._stage_i
                self._stage_i ^= 1
                s = self._stage[k]
                with torch.cuda.stream(self._copy_stream):
                    for ev in (s["free"], s["enc"]):        # the step / the look-ahead encoder pass that last read this slot
                        if ev is not None:
                            self._copy_stream.wait_event(ev)
                    for name, dt in (("wav", torch.float32), ("cap", torch.int64)):
                        src = batch[name]
                        if s[name] is None or s[name].shape != src.shape:
                            s[name] = torch.empty(src.shape, dtype=dt, device=dev)
                        s[name].copy_(src, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self._copy_stream)
                staged.update(wav=s["wav"], cap=s["cap"], _ready=ready, _slot=k)
            if encode:
                if getattr(self, "_cnn_stream", None) is None:
                    self._make_look_ahead_streams()
                cnn = m.encoder.cnn
                if ready is not None:
                    self._cnn_stream.wait_event(ready)
                else:
                    self._cnn_stream.wait_stream(user)               # a device-resident batch: whatever produced it
                cnn.sm_limit = self.cnn_sms                          # (only this pass: other callers get the whole chip)
                try:
                    with torch.cuda.stream(self._cnn_stream):
                        wav = staged["wav"].to(dev, torch.float32)
                        staged["_cnn_out"] = cnn({"wav": wav, "wav_len": batch["wav_len"], "specaug": self.specaug})
                        done = torch.cuda.Event()
                        done.record(self._cnn_stream)
                finally:
                    cnn.sm_limit = 0
                staged["_cnn_ready"] = done
                if "_slot" in staged:
                    self._stage[staged["_slot"]]["enc"] = done      # (prefetching more than one batch ahead stays safe)
        return staged

    def wait_look_ahead(self):
        """Make the current stream wait for an encoder pass `prefetch` may have in flight (the CNN has ONE workspace: call
        this before running the encoder yourself, e.g. for validation between epochs)."""
        if getattr(self, "_cnn_stream", None) is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._cnn_stream)

    def _make_look_ahead_streams(self):
        """Two streams on DISJOINT SM sets (csrc/sm_partition.cu: CUDA green contexts): `cnn_sms` SMs for the look-ahead
        encoder, the rest for the trainable chain.  On ordinary streams the convolutions' pending CTAs take every SM that
        frees up and the small kernels starve (measured: no gain).  AC_TRAIN_PARTITION=0, or a driver without green
        contexts: ordinary streams (trainable chain at high priority)."""
        dev = self.device
        self.partition, self.partition_error = None, None
        if os.environ.get("AC_TRAIN_PARTITION", "1") != "0":
            l = _lib.lib()
            h = ctypes.c_void_p()
            if l.ac_sm_partition_create(int(self.cnn_sms), ctypes.byref(h)) == 0:
                self.partition = h
                self._cnn_stream = torch.cuda.ExternalStream(l.ac_sm_partition_stream(h, 0), device=dev)
                self._hi_stream = torch.cuda.ExternalStream(l.ac_sm_partition_stream(h, 1), device=dev)
                self.cnn_sms = int(l.ac_sm_partition_sms(h, 0))
                self.train_sms = int(l.ac_sm_partition_sms(h, 1))
                return
            self.partition_error = l.ac_last_error().decode()
        self._cnn_stream = torch.cuda.Stream(device=dev)
        self._hi_stream = torch.cuda.Stream(device=dev, priority=-1)
        self.train_sms = None

    # ---- one step -------------------------------------------------------------------------------------------------------
    def step(self, batch, coins=None):
        """batch: {"wav" [B, N] fp32 (host, ideally pinned, or cuda), "wav_len" [B], "cap" [B, Lc] int64, "cap_len" [B]},
        or what `prefetch(batch)` returned.  Returns {"loss": device scalar tensor [1], "tokens": int, "lr": float,
        "ss_ratio": float}."""
        m = self.model
        dev = self.device
        l = _lib.lib()
        self._update_ss_ratio()
        # the scheduler was stepped once by its constructor; iteration k (0-based) steps it for the (k+2)-th time (run.py:105)
        self.lr = exponential_decay_lr(self.iteration + 2, self.base_lr, self.final_lr, self.total_iters, self.warmup_iters)
        look_ahead = "_cnn_out" in batch
        with contextlib.ExitStack() as stack, torch.cuda.device(dev), torch.no_grad():
            user = torch.cuda.current_stream()
            # with the encoder of the next batch running beside this step (prefetch(encode=True)), the trainable chain runs
            # on a high-priority stream: its few CTAs win the SMs the look-ahead's kernels free up
            main = self._hi_stream if look_ahead else user
            if main is not user:
                main.wait_stream(user)
            stack.enter_context(torch.cuda.stream(main))
            if "_ready" in batch:
                main.wait_event(batch["_ready"])
            # the weight re-pack of both trainable engines only depends on the previous optimizer step: it runs on a side
            # stream next to the frozen CNN's forward pass
            rnn, dec = m.encoder.rnn, m.decoder
            if getattr(self, "_side_stream", None) is None:
                self._side_stream = torch.cuda.Stream(device=dev)
            self._side_stream.wait_stream(main)
            with torch.cuda.stream(self._side_stream):
                rnn.train_engine.refresh(grads="param")
                dec.train_engine.refresh(grads="param")
            wav = batch["wav"]
            wav = wav.to(dev, torch.float32, non_blocking=True)
            cap = batch["cap"].to(dev, torch.int64, non_blocking=True)
            cap_len = torch.as_tensor(batch["cap_len"]).to(torch.int64)
            tgt_len_dev = to_device_async(cap_len - 1, dev, torch.int64)
            # frozen CNN (dropout on, BatchNorm eval) -> frames; bi-GRU; decoder
            if look_ahead:
                main.wait_event(batch["_cnn_ready"])
                cnn_out = batch["_cnn_out"]
                for t in (cnn_out["attn_emb"], cnn_out["fc_emb"]):          # allocated on the look-ahead stream, read here
                    t.record_stream(main)
            else:
                if getattr(self, "_cnn_stream", None) is not None:         # never two encoder passes at once (one workspace)
                    main.wait_stream(self._cnn_stream)
                m.encoder.cnn.sm_limit = 0
                cnn_out = m.encoder.cnn({"wav": wav, "wav_len": batch["wav_len"], "specaug": self.specaug})
            lens = cnn_out["attn_emb_len"]
            t_out = int(lens.max())
            x = cnn_out["attn_emb"][:, :t_out].contiguous()
            len_dev = to_device_async(lens, dev, torch.int64)
            main.wait_stream(self._side_stream)
            mem = rnn.train_engine.forward(x, len_dev, p_drop=float(rnn.dropout) if rnn.num_layers > 1 else 0.0, grads="param",
                                           refresh=False)
            L = cap.size(1) - 1
            if coins is None:
                coins = [random.random() < self.ss_ratio for _ in range(L)] if self.ss_ratio != 1 else None
            out = dec.train_engine.forward(mem, len_dev, cap[:, :-1].contiguous(), coins=coins,
                                           p_drop=float(dec.in_dropout.p), grads="param", start_idx=m.start_idx,
                                           end_idx=m.end_idx, pad_idx=m.pad_idx, refresh=False)
            from .captioning.losses.loss import ls_ce_fwd_bwd
            loss, dlogit = ls_ce_fwd_bwd(out["logit_padded"][:, :, :dec.vocab_size], cap[:, 1:], tgt_len_dev, self.smoothing)
            dmem = dec.train_engine.backward(dlogit, need_dattn=True)
            # data parallel: the decoder's gradients are final here -- their all-reduce overlaps the GRU's backward pass
            work = None
            if self.world > 1 and self.dec_offset is not None:
                work = allreduce_bucket_async(self.flat_grad[self.dec_offset:], self.group)
            rnn.train_engine.backward(dmem, need_dx=False)
            if work is not None:
                allreduce_gradients(self.flat_grad[:self.dec_offset], self.group)
                work.wait()
                grad_scale = 1.0 / self.world
            else:
                grad_scale = allreduce_gradients(self.flat_grad, self.group) if self.world > 1 else 1.0
            _lib.check(l.ac_clip_adam(_lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                      _lib.ptr(self.exp_avg_sq), self.flat_param.numel(), self.lr, self.betas[0], self.betas[1],
                                      self.eps, self.weight_decay, self.max_grad_norm, grad_scale, _lib.ptr(loss),
                                      _lib.ptr(self._step_dev), _lib.ptr(self.grad_norm), _lib.ptr(self._adam_ws),
                                      self._adam_ws.numel(), _lib.current_stream()), "ac_clip_adam")
            if "_slot" in batch:                     # the staging slot may be overwritten once this step's kernels are done
                ev = torch.cuda.Event()
                ev.record(main)
                self._stage[batch["_slot"]]["free"] = ev
            if main is not user:
                loss.record_stream(user)
                user.wait_stream(main)
        self.iteration += 1
        self.last_output = out
        return {"loss": loss, "tokens": int((cap_len - 1).sum()), "lr": self.lr, "ss_ratio": self.ss_ratio}
