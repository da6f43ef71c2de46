"""Mirror of the HF release wrapper captioning/models/hf_wrapper.py:1071-1181
(`ContraEncoderKdWrapper`, `Effb2TrmConfig`, `Effb2TrmCaptioningModel`).

state_dict keys equal the reference's (`model.model.encoder...`, `model.model.decoder...`,
`model.stdnt_proj`, `model.tchr_proj`, `model.logit_scale`).  `transformers.PreTrainedModel`
is not required: hub download is impossible offline, weights arrive through `load_state_dict`."""
from typing import List, Union

import numpy as np
import torch
import torch.nn as nn

from .base import CaptionMetaMixin
from .cnn_encoder import EfficientNetB2
from .transformer_decoder import TransformerDecoder
from .transformer_model import TransformerModel


class ContraEncoderKdWrapper(nn.Module, CaptionMetaMixin):
    """hf_wrapper.py:1071-1112; without `tchr_output` it is a pass-through (the inference path)."""

    def __init__(self, model: nn.Module, shared_dim: int, tchr_dim: int):
        super().__init__()
        self.model = model
        self.tchr_dim = tchr_dim
        self.stdnt_proj = nn.Linear(model.encoder.fc_emb_size, shared_dim)
        self.tchr_proj = nn.Linear(tchr_dim, shared_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def forward(self, input_dict):
        if "tchr_output" in input_dict or input_dict.get("unsup", False):
            raise NotImplementedError("encoder-KD training branch is out of scope")
        return self.model(input_dict)


class Effb2TrmConfig:
    def __init__(self, sample_rate: int = 16000, tchr_dim: int = 768, shared_dim: int = 1024,
                 fc_emb_dim: int = 1408, attn_emb_dim: int = 1408, decoder_n_layers: int = 2,
                 decoder_we_tie_weights: bool = True, decoder_emb_dim: int = 256,
                 decoder_dropout: float = 0.2, vocab_size: int = 4981, **kwargs):
        self.sample_rate = sample_rate
        self.tchr_dim = tchr_dim
        self.shared_dim = shared_dim
        self.fc_emb_dim = fc_emb_dim
        self.attn_emb_dim = attn_emb_dim
        self.decoder_n_layers = decoder_n_layers
        self.decoder_we_tie_weights = decoder_we_tie_weights
        self.decoder_emb_dim = decoder_emb_dim
        self.decoder_dropout = decoder_dropout
        self.vocab_size = vocab_size


class PendingCaptions:
    """Handle returned by `Effb2TrmCaptioningModel.submit`: the token ids land in pinned host memory when the
    recorded event completes."""

    def __init__(self, seq_host: torch.Tensor, done: "torch.cuda.Event"):
        self._seq, self._done = seq_host, done

    def done(self) -> bool:
        return self._done.query()

    def result(self) -> torch.Tensor:
        self._done.synchronize()
        return self._seq


class Effb2TrmCaptioningModel(nn.Module):
    config_class = Effb2TrmConfig

    def __init__(self, config=None):
        super().__init__()
        config = config or Effb2TrmConfig()
        self.config = config
        encoder = EfficientNetB2()
        decoder = TransformerDecoder(emb_dim=config.decoder_emb_dim, vocab_size=config.vocab_size,
                                     fc_emb_dim=config.fc_emb_dim, attn_emb_dim=config.attn_emb_dim,
                                     dropout=config.decoder_dropout, nlayers=config.decoder_n_layers,
                                     tie_weights=config.decoder_we_tie_weights)
        model = TransformerModel(encoder, decoder)
        self.model = ContraEncoderKdWrapper(model, config.shared_dim, config.tchr_dim)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, audio: torch.Tensor, audio_length: Union[List, np.ndarray, torch.Tensor],
                sample_method: str = "beam", beam_size: int = 3, max_length: int = 20, temp: float = 1.0):
        """hf_wrapper.py:1162-1181: returns LongTensor[B, max_length] on the CPU."""
        input_dict = {
            "wav": audio.to(self.device, non_blocking=True),
            "wav_len": audio_length,
            "specaug": False,
            "mode": "inference",
            "sample_method": sample_method,
            "max_length": max_length,
            "temp": temp,
            "need_logit": False,
        }
        if sample_method == "beam":
            input_dict["beam_size"] = beam_size
        return self.model(input_dict)["seq"].cpu()

    def submit(self, audio: torch.Tensor, audio_length, sample_method: str = "beam", beam_size: int = 3,
               max_length: int = 20, temp: float = 1.0) -> PendingCaptions:
        """Asynchronous `forward` for back-to-back batches (serving): the host->device copy of `audio` runs on a
        private copy stream, the kernels on the current stream, the token ids come back into pinned memory, and
        nothing blocks the host -- so batch i+1's upload overlaps batch i's kernels.  `result()` of the returned
        handle gives the same LongTensor[B, max_length] (CPU) as `forward`.  `audio` should be pinned."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        if audio.is_cuda:
            wav = audio
        else:
            # (the buffer is allocated in the copy stream's pool and handed to the compute stream with
            #  record_stream, so the caching allocator keeps reuse of earlier batches ordered)
            with torch.cuda.stream(self._copy_stream):
                wav = audio.to(dev, non_blocking=True)
            cur.wait_stream(self._copy_stream)
            wav.record_stream(cur)
        input_dict = {"wav": wav, "wav_len": audio_length, "specaug": False, "mode": "inference",
                      "sample_method": sample_method, "max_length": max_length, "temp": temp,
                      "need_logit": False, "_device_seq": True}
        if sample_method == "beam":
            input_dict["beam_size"] = beam_size
        seq_dev = self.model(input_dict)["seq"]
        seq_host = torch.empty(seq_dev.shape, dtype=seq_dev.dtype, pin_memory=True)
        seq_host.copy_(seq_dev, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        return PendingCaptions(seq_host, done)
