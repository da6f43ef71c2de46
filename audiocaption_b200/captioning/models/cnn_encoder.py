"""Mirror of the reference's waveform encoders (captioning/models/cnn_encoder.py).

``Cnn14Encoder`` -- cnn_encoder.py:326-464 / hf_wrapper.py:1185-1304: 32 kHz log-mel front-end (n_fft 1024, hop 320,
64 slaney mels, 50-14000 Hz, no top_db) + bn0 + six 3x3 ConvBlocks + fc1; arithmetic in csrc/logmel.cu and
csrc/cnn14.cu (tcgen05 implicit-GEMM convolutions).

``EfficientNetB2`` -- cnn_encoder.py:769-839 / hf_wrapper.py:260-315: 16 kHz log-mel front-end
(n_fft 512, hop 160, 64 HTK mels, top_db 120) + EfficientNet-B2 backbone + mean over
frequency.  The modules below only HOLD parameters under the reference's state_dict names
(``melspec_extractor.spectrogram.window``, ``melspec_extractor.mel_scale.fb``,
``backbone.eff_net._blocks.N._expand_conv.weight`` ...); the arithmetic runs in
csrc/logmel.cu and csrc/effb2.cu through the C ABI.
"""
import ctypes
import math

import torch
import torch.nn as nn

from ... import _lib
from ._native import Workspace, params_signature, require_cuda, to_device_async


# ----------------------------------------------------------------------------- parameter holders
class _Conv(nn.Module):
    def __init__(self, cin, cout, k, groups=1, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(cin // groups * k * k)
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        self.out_channels = cout


class _BatchNormParams(nn.Module):
    """Holder of eval-mode BatchNorm statistics (the class name contains "BatchNorm" so that
    crnn_trm_encoder.py:195-203's `freeze_cnn_bn` walk finds it)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


_BN = _BatchNormParams


class _MBConv(nn.Module):
    def __init__(self, cin, cout, expand, k, nsq):
        super().__init__()
        ce = cin * expand
        if expand != 1:
            self._expand_conv = _Conv(cin, ce, 1)
            self._bn0 = _BN(ce)
        self._depthwise_conv = _Conv(ce, ce, k, groups=ce)
        self._bn1 = _BN(ce)
        self._se_reduce = _Conv(ce, nsq, 1, bias=True)
        self._se_expand = _Conv(nsq, ce, 1, bias=True)
        self._project_conv = _Conv(ce, cout, 1)
        self._bn2 = _BN(cout)


def _block_plan():
    l = _lib.lib()
    info = (ctypes.c_int * 9)()
    n = l.ac_effb2_block_info(0, info)
    plan = []
    for i in range(n):
        l.ac_effb2_block_info(i, info)
        plan.append(tuple(info))
    return plan


class _EffNetHolder(nn.Module):
    """state_dict-compatible with efficientnet_pytorch.EfficientNet('efficientnet-b2', include_top=False)
    after ``_change_in_channels(1)`` (hf_wrapper.py:235-241)."""

    def __init__(self):
        super().__init__()
        plan = _block_plan()
        self._conv_stem = _Conv(1, 32, 3)
        self._bn0 = _BN(32)
        self._blocks = nn.ModuleList([_MBConv(p[0], p[1], p[2], p[3], p[7]) for p in plan])
        self._conv_head = _Conv(plan[-1][1], _lib.lib().ac_effb2_out_dim(), 1)
        self._bn1 = _BN(self._conv_head.out_channels)


class _EffiNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.eff_net = _EffNetHolder()


class _Spectrogram(nn.Module):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft, periodic=True))


class _MelScale(nn.Module):
    def __init__(self, fb):
        super().__init__()
        self.register_buffer("fb", fb)


def _melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm, mel_scale):
    """Same construction as torchaudio.functional.melscale_fbanks (the reference gets its `fb`
    buffer from torchaudio.transforms.MelScale); used only to initialise the buffer -- a loaded
    state_dict overrides it."""
    import torchaudio
    return torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm, mel_scale)


class MelSpectrogram(nn.Module):
    """Parameter holder + launcher for the fused log-mel kernel (csrc/logmel.cu)."""

    def __init__(self, sample_rate, n_fft, hop_length, f_min, f_max, n_mels, norm=None, mel_scale="htk"):
        super().__init__()
        self.n_fft, self.hop_length, self.n_mels = n_fft, hop_length, n_mels
        self.spectrogram = _Spectrogram(n_fft)
        self.mel_scale = _MelScale(_melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels,
                                                    sample_rate, norm, mel_scale))
        self._handle = None
        self._sig = None

    def _frontend(self):
        bufs = [self.spectrogram.window, self.mel_scale.fb]
        sig = params_signature(bufs)
        if self._handle is None or sig != self._sig:
            self.release()
            win = self.spectrogram.window.detach().float().cpu().contiguous()
            fb = self.mel_scale.fb.detach().float().cpu().contiguous()
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_frontend_create(_lib.ptr(win), self.n_fft, self.hop_length, _lib.ptr(fb),
                                                     fb.shape[0], fb.shape[1], ctypes.byref(h)), "ac_frontend_create")
            self._handle, self._sig = h, sig
        return self._handle

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_frontend_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def forward(self, wav: torch.Tensor, want_max: bool = False):
        """wav [B, N] fp32 cuda -> (lms [B, n_mels, T] in dB (un-clamped), gmax [1] or None)"""
        require_cuda(wav, "MelSpectrogram")
        wav = wav.float().contiguous()
        B, N = wav.shape
        T = 1 + N // self.hop_length
        out = torch.empty(B, self.n_mels, T, device=wav.device, dtype=torch.float32)
        gmax = torch.empty(1, device=wav.device, dtype=torch.float32) if want_max else None
        with torch.cuda.device(wav.device):
            _lib.check(_lib.lib().ac_logmel_fwd(self._frontend(), _lib.ptr(wav), B, N, _lib.ptr(out), _lib.ptr(gmax),
                                                _lib.current_stream()), "ac_logmel_fwd")
        return out, gmax


class AmplitudeToDB(nn.Module):
    """AmplitudeToDB(stype="power", top_db) on an already-dB tensor produced by the fused kernel:
    only the batch-global top_db clamp is left to do."""

    def __init__(self, top_db=None):
        super().__init__()
        self.top_db = top_db

    def forward(self, lms, gmax):
        if self.top_db is not None:
            _lib.check(_lib.lib().ac_db_clamp(_lib.ptr(lms), lms.numel(), _lib.ptr(gmax), float(self.top_db),
                                              _lib.current_stream()), "ac_db_clamp")
        return lms


class EfficientNetB2(nn.Module):
    """Drop-in for captioning.models.cnn_encoder.EfficientNetB2 (cnn_encoder.py:769-839,
    hf_wrapper.py:260-315).  Inference only (eval-mode BatchNorm, no SpecAugment)."""

    def __init__(self, n_mels: int = 64, win_length: int = 32, hop_length: int = 10, f_min: int = 0,
                 pretrained: bool = False, prune_ratio: float = 0.0, prune_se: bool = True,
                 prune_start_layer: int = 0, freeze: bool = False):
        super().__init__()
        if prune_ratio > 0 or pretrained:
            raise NotImplementedError("pruned / downloaded EfficientNet-B2 variants are out of scope")
        sample_rate = 16000
        n_fft = win_length * sample_rate // 1000
        self.melspec_extractor = MelSpectrogram(sample_rate, n_fft, hop_length * sample_rate // 1000,
                                                f_min, sample_rate // 2, n_mels)
        self.hop_length = 10 * sample_rate // 1000
        self.db_transform = AmplitudeToDB(top_db=120)
        self.backbone = _EffiNet()
        self.fc_emb_size = self.backbone.eff_net._conv_head.out_channels
        self.downsample_ratio = 32
        self.db_max_reduce = None       # optional callable(gmax [1] device tensor) -> tensor, see forward()
        self._ws = Workspace()
        self._handle = None
        self._sig = None
        if freeze:
            for p in self.parameters():
                p.requires_grad = False

    # -- weight pack (re-done whenever a parameter changed)
    def _backbone_tensors(self):
        return [v for k, v in self.backbone.eff_net.state_dict(keep_vars=True).items()
                if not k.endswith("num_batches_tracked")]

    def _net(self):
        tensors = self._backbone_tensors()
        sig = params_signature(tensors)
        if self._handle is None or sig != self._sig:
            self.release()
            ts = [t.detach().float().contiguous() for t in tensors]
            for t in ts:
                require_cuda(t, "EfficientNetB2 parameters")
            ptrs, numels, n = _lib.tensor_table(ts)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_effb2_create(ptrs, numels, n, _lib.current_stream(), ctypes.byref(h)),
                       "ac_effb2_create")
            self._handle, self._sig = h, sig
        return self._handle

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_effb2_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def log_mel(self, wav):
        """The reference's `db_transform(melspec_extractor(wav))` (clamped), for inspection."""
        lms, gmax = self.melspec_extractor(wav, want_max=True)
        return self.db_transform(lms, gmax)

    def forward(self, input_dict):
        wav = input_dict["wav"]
        wav_len = input_dict["wav_len"]
        if self.training and input_dict.get("specaug", False):
            raise NotImplementedError("SpecAugment (training) is out of scope of the B200 inference path")
        require_cuda(wav, "EfficientNetB2.forward")
        l = _lib.lib()
        with torch.cuda.device(wav.device):
            lms, gmax = self.melspec_extractor(wav, want_max=True)   # clamp is fused into the stem conv
            if self.db_max_reduce is not None:
                # a batch split across ranks: one scalar all-reduce(max) (audiocaption_b200.sharding.global_db_max) restores
                # the batch-global top_db reference of hf_wrapper.py:292-293; runs between the two kernels, on the stream
                gmax = self.db_max_reduce(gmax)
            B, F, T = lms.shape
            Tp = l.ac_effb2_out_frames(T)
            attn_emb = torch.empty(B, Tp, self.fc_emb_size, device=wav.device, dtype=torch.float32)
            nbytes = l.ac_effb2_workspace_bytes(B, F, T)
            ws = self._ws.get(nbytes, wav.device)
            _lib.check(l.ac_effb2_fwd(self._net(), _lib.ptr(lms), _lib.ptr(gmax), float(self.db_transform.top_db),
                                      B, F, T, _lib.ptr(attn_emb), _lib.ptr(ws), nbytes, _lib.current_stream()),
                       "ac_effb2_fwd")
            wave_length = torch.as_tensor(wav_len)
            feat_length = torch.div(wave_length, self.hop_length, rounding_mode="floor") + 1
            feat_length = torch.div(feat_length, self.downsample_ratio, rounding_mode="floor")
            len_dev = to_device_async(feat_length, wav.device, torch.int64)
            fc_emb = torch.empty(B, self.fc_emb_size, device=wav.device, dtype=torch.float32)
            _lib.check(l.ac_masked_mean(_lib.ptr(attn_emb), _lib.ptr(len_dev), B, Tp, self.fc_emb_size,
                                        _lib.ptr(fc_emb), _lib.current_stream()), "ac_masked_mean")
        # (lengths given as a CUDA tensor stay on the device: nothing here synchronises, so the call can be captured in a
        #  CUDA graph -- Effb2TrmCaptioningModel.forward does that for repeated shapes)
        return {"fc_emb": fc_emb, "attn_emb": attn_emb, "attn_emb_len": len_dev if feat_length.is_cuda else feat_length.cpu()}


def draw_specaug_stripes(batch, n_frames, n_mels, time_drop_width=64, time_stripes_num=2, freq_drop_width=8,
                         freq_stripes_num=2):
    """Stripe positions of torchlibrosa's `SpecAugmentation` (`DropStripes`: per clip and stripe, width ~ U{0..drop_width-1},
    begin ~ U{0..total - width - 1}; all time stripes of the batch first, then all mel stripes), drawn from torch's global
    CPU generator in that order.  torchlibrosa is a third-party dependency absent from this image: its published algorithm
    is restated, unpinned.  Returns int32 [batch, 2 * stripes, 2] (begin, width): frame ranges first, then mel ranges."""
    assert time_stripes_num == freq_stripes_num, "one stripe count for both axes"
    ns = time_stripes_num
    out = torch.zeros(batch, 2 * ns, 2, dtype=torch.int32)
    for axis, (total, width) in enumerate(((n_frames, time_drop_width), (n_mels, freq_drop_width))):
        for b in range(batch):
            for k in range(ns):
                distance = int(torch.randint(low=0, high=width, size=(1,))[0])
                bgn = int(torch.randint(low=0, high=total - distance, size=(1,))[0])
                out[b, axis * ns + k, 0], out[b, axis * ns + k, 1] = bgn, distance
    return out


# ----------------------------------------------------------------------------- Cnn14
class _ConvBlock(nn.Module):
    """Parameter holder with the state_dict layout of cnn_encoder.py:32-50 `ConvBlock`."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _Conv(cin, cout, 3)
        self.conv2 = _Conv(cout, cout, 3)
        self.bn1 = _BN(cout)
        self.bn2 = _BN(cout)
        nn.init.xavier_uniform_(self.conv1.weight)       # init_layer, cnn_encoder.py:17-23
        nn.init.xavier_uniform_(self.conv2.weight)


class Cnn14Encoder(nn.Module):
    """Drop-in for captioning.models.cnn_encoder.Cnn14Encoder (cnn_encoder.py:326-464; HF copy
    hf_wrapper.py:1185-1304).  Forward only: BatchNorm always uses its running statistics (the training configs freeze
    the CNN and its BatchNorm, eg_configs/*/waveform/cnn14rnn_trm.yaml:11-12); in train mode the functional dropouts of
    cnn_encoder.py:432-456 are applied (p = 0.2 after every ConvBlock, 0.5 around fc1).  No SpecAugment."""

    conv_dropout = 0.2
    fc_dropout = 0.5
    # "fp32": 3xTF32 tensor-core convolutions (fp32-level accuracy, every fp32 parity claim); "tf32": plain TF32 operands
    # with fp32 accumulation; "bf16": bf16 activations and weights between bn0 and the pooled output, fp32 accumulation
    # (csrc/conv_bf16.cu, 3.7x the TF32 convolutions) -- the last two for the configurations BASELINE.json states in bf16
    # (training with the frozen CNN, temporal captioner)
    conv_precision = "fp32"

    def __init__(self, sample_rate: int = 32000, freeze: bool = False):
        super().__init__()
        sr_to_fmax = {32000: 14000, 16000: 8000}
        self.melspec_extractor = MelSpectrogram(sample_rate, 32 * sample_rate // 1000, 10 * sample_rate // 1000,
                                                50, sr_to_fmax[sample_rate], 64, norm="slaney", mel_scale="slaney")
        self.hop_length = 10 * sample_rate // 1000
        self.db_transform = AmplitudeToDB()
        self.bn0 = _BN(64)
        chans = (1, 64, 128, 256, 512, 1024, 2048)
        for i in range(6):
            setattr(self, f"conv_block{i + 1}", _ConvBlock(chans[i], chans[i + 1]))
        self.downsample_ratio = 32
        self.fc1 = nn.Linear(2048, 2048, bias=True)
        nn.init.xavier_uniform_(self.fc1.weight)
        nn.init.zeros_(self.fc1.bias)
        self.fc_emb_size = 2048
        self.freeze = freeze
        self._ws = Workspace()
        self._handle = None
        self._sig = None
        if freeze:
            for p in self.parameters():
                p.requires_grad = False

    def load_pretrained(self, pretrained, output_fn=print):
        """cnn_encoder.py:376-412: accepts the three checkpoint layouts the reference knows -- PANNs ({"model": state_dict
        with this module's key names plus spectrogram / logmel extractor and fc_audioset entries that do not match and
        are skipped}), COLA ({"model": {"backbone.<key>": ...}}) and BLAT ({"state_dict": {"...audio_encoder.<key>": ...}})
        -- and merge-loads the entries whose key and shape match.  With `freeze`, exactly the loaded parameters are
        frozen.  `pretrained` may also be an already-loaded checkpoint dict."""
        from ..utils.train_util import merge_load_state_dict
        checkpoint = pretrained if isinstance(pretrained, dict) else torch.load(pretrained, map_location="cpu")
        if "model" in checkpoint:
            model_sd = checkpoint["model"]
            if any(key.startswith("backbone.") for key in model_sd):          # COLA
                state_dict = {key.replace("backbone.", ""): value for key, value in model_sd.items()
                              if key.startswith("backbone.")}
            else:                                                              # PANNs
                state_dict = model_sd
        elif "state_dict" in checkpoint:                                       # BLAT
            state_dict = {key.replace("audio_encoder.", ""): value for key, value in checkpoint["state_dict"].items()
                          if "audio_encoder" in key}
        else:
            raise Exception("Unkown checkpoint format")
        loaded_keys = merge_load_state_dict(state_dict, self, output_fn)
        if self.freeze:
            for name, param in self.named_parameters():
                param.requires_grad = name not in loaded_keys

    def _body_tensors(self):
        ts = [self.bn0.weight, self.bn0.bias, self.bn0.running_mean, self.bn0.running_var]
        for i in range(1, 7):
            blk = getattr(self, f"conv_block{i}")
            ts += [blk.conv1.weight, blk.conv2.weight]
            for bn in (blk.bn1, blk.bn2):
                ts += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        return ts + [self.fc1.weight, self.fc1.bias]

    def _net(self):
        # `weights_pinned` (set by TrainStep for the frozen CNN): nothing writes the weights, skip the per-call signature
        if self._handle is not None and getattr(self, "weights_pinned", False) and \
                self._pinned_precision == (self.conv_precision, getattr(self, "sm_limit", 0)):
            return self._handle
        tensors = self._body_tensors()
        sig = params_signature(tensors)
        if self._handle is None or sig != self._sig:
            self.release()
            ts = [t.detach().float().contiguous() for t in tensors]
            for t in ts:
                require_cuda(t, "Cnn14Encoder parameters")
            ptrs, numels, n = _lib.tensor_table(ts)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().ac_cnn14_create(ptrs, numels, n, _lib.current_stream(), ctypes.byref(h)),
                       "ac_cnn14_create")
            self._handle, self._sig = h, sig
        modes = {"fp32": 3, "tf32": 1, "bf16": 16}
        if self.conv_precision not in modes:
            raise ValueError(f"conv_precision {self.conv_precision!r}: expected 'fp32', 'tf32' or 'bf16'")
        _lib.check(_lib.lib().ac_cnn14_set_precision(self._handle, modes[self.conv_precision]), "ac_cnn14_set_precision")
        _lib.check(_lib.lib().ac_cnn14_set_sm_limit(self._handle, int(getattr(self, "sm_limit", 0))), "ac_cnn14_set_sm_limit")
        self._pinned_precision = (self.conv_precision, getattr(self, "sm_limit", 0))
        return self._handle

    def release(self):
        if self._handle is not None:
            _lib.lib().ac_cnn14_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def log_mel(self, wav):
        """The reference's `db_transform(melspec_extractor(wav))` [B, 64, T]."""
        return self.melspec_extractor(wav)[0]

    def forward(self, input_dict):
        # the HF copy (hf_wrapper.py:1259-1261) receives the log-mel as `lms`, the training class the waveform
        wav = input_dict["lms"] if "lms" in input_dict else input_dict["wav"]
        wav_len = input_dict["wav_len"]
        require_cuda(wav, "Cnn14Encoder.forward")
        l = _lib.lib()
        with torch.cuda.device(wav.device):
            lms = wav.float().contiguous() if "lms" in input_dict else self.log_mel(wav)
            B, F, T = lms.shape
            if self.training and input_dict.get("specaug", False):          # cnn_encoder.py:424-425
                stripes = input_dict.get("_specaug_stripes")
                stripes = draw_specaug_stripes(B, T, F) if stripes is None else stripes
                if "lms" in input_dict:
                    lms = lms.clone()
                st_dev = to_device_async(stripes.to(torch.int32).contiguous(), wav.device)
                _lib.check(l.ac_specaug_apply(_lib.ptr(lms), B, F, T, _lib.ptr(st_dev), stripes.shape[1] // 2,
                                              _lib.current_stream()), "ac_specaug_apply")
            Tp = l.ac_cnn14_out_frames(T)
            wave_length = torch.as_tensor(wav_len)
            feat_length = torch.div(wave_length, self.hop_length, rounding_mode="floor") + 1
            feat_length = torch.div(feat_length, self.downsample_ratio, rounding_mode="floor")
            len_dev = to_device_async(feat_length, wav.device, torch.int64)
            attn_emb = torch.empty(B, Tp, self.fc_emb_size, device=wav.device, dtype=torch.float32)
            fc_emb = torch.empty(B, self.fc_emb_size, device=wav.device, dtype=torch.float32)
            nbytes = l.ac_cnn14_workspace_bytes(B, F, T)
            ws = self._ws.get(nbytes, wav.device)
            if self.training:
                if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                    raise _lib.AudioCaptionB200Error(
                        "Cnn14Encoder has no backward pass on the B200 path: freeze it (freeze_cnn: True, as in "
                        "eg_configs/*/waveform/cnn14rnn_trm.yaml) or call .eval()")
                seed = input_dict.get("_dropout_seed")
                seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else seed
                _lib.check(l.ac_cnn14_fwd_train(self._net(), _lib.ptr(lms), B, F, T, _lib.ptr(len_dev), float(self.conv_dropout),
                                                float(self.fc_dropout), seed, _lib.ptr(attn_emb), _lib.ptr(fc_emb), _lib.ptr(ws),
                                                nbytes, _lib.current_stream()), "ac_cnn14_fwd_train")
            else:
                _lib.check(l.ac_cnn14_fwd(self._net(), _lib.ptr(lms), B, F, T, _lib.ptr(len_dev), _lib.ptr(attn_emb),
                                          _lib.ptr(fc_emb), _lib.ptr(ws), nbytes, _lib.current_stream()), "ac_cnn14_fwd")
        return {"fc_emb": fc_emb, "attn_emb": attn_emb, "attn_emb_len": feat_length.cpu()}
