"""Mirror of captioning/models/base.py: CaptionMetaMixin (:11-21) and CaptionModel (:24-361): the train-mode forward
(teacher forcing / scheduled sampling, :131-170) and the inference modes greedy and beam.  The python per-token loops of
the reference are replaced by single launches of the decode kernels (inference) and by dense full-prefix passes
(training); the dict-in / dict-out contract is unchanged."""
from typing import Dict

import torch
import torch.nn as nn


class CaptionMetaMixin:
    pad_idx = 0
    start_idx = 1
    end_idx = 2
    max_length = 20

    @classmethod
    def set_index(cls, start_idx, end_idx, pad_idx):
        cls.start_idx = start_idx
        cls.end_idx = end_idx
        cls.pad_idx = pad_idx


class CaptionModel(nn.Module, CaptionMetaMixin):
    """Encoder-decoder captioning model (base.py:24-477), inference modes greedy and beam."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module, **kwargs):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.vocab_size = decoder.vocab_size
        self.train_forward_keys = ["cap", "cap_len", "ss_ratio"]
        self.inference_forward_keys = ["sample_method", "max_length", "temp"]
        if kwargs.get("freeze_encoder", False):
            for param in self.encoder.parameters():
                param.requires_grad = False
        self.check_decoder_compatibility()

    def check_decoder_compatibility(self):
        names = [x.__name__ for x in self.compatible_decoders]
        assert isinstance(self.decoder, self.compatible_decoders), \
            f"{self.decoder.__class__.__name__} is incompatible with " \
            f"{self.__class__.__name__}, please use decoder in {names} "

    def forward(self, input_dict: Dict):
        encoder_output_dict = self.encoder(input_dict)
        return self.forward_decoder(input_dict, encoder_output_dict)

    def forward_decoder(self, input_dict: Dict, encoder_output_dict: Dict):
        if input_dict["mode"] == "train":
            forward_dict = {"mode": "train", "sample_method": "greedy", "temp": 1.0}
            for key in self.train_forward_keys:
                forward_dict[key] = input_dict[key]
            forward_dict.update(encoder_output_dict)
            output = self.train_forward(forward_dict)
        elif input_dict["mode"] == "inference":
            forward_dict = {"mode": "inference"}
            default_args = {"sample_method": "greedy", "max_length": self.max_length, "temp": 1.0}
            for key in self.inference_forward_keys:
                forward_dict[key] = input_dict[key] if key in input_dict else default_args[key]
            if forward_dict["sample_method"] == "beam":
                forward_dict["beam_size"] = input_dict.get("beam_size", 3)
                if input_dict.get("n_best", False):
                    raise NotImplementedError("n_best beam output is not built")
            for key in ("need_logit", "_device_seq"):      # B200-side options (not in the reference)
                if key in input_dict:
                    forward_dict[key] = input_dict[key]
            forward_dict.update(encoder_output_dict)
            output = self.inference_forward(forward_dict)
        else:
            raise Exception("mode should be either 'train' or 'inference'")
        output.update(encoder_output_dict)
        return output

    def train_forward(self, input_dict):
        """base.py:131-137: scheduled sampling whenever ss_ratio != 1 (always, from the first iteration on:
        run.py:55-65 lowers it before the forward), plain teacher forcing otherwise."""
        if input_dict["ss_ratio"] != 1:
            input_dict["mode"] = "train"
            return self.stepwise_forward(input_dict)
        output = self.seq_forward(input_dict)
        self.train_process(output, input_dict)
        return output

    def seq_forward(self, input_dict):
        raise NotImplementedError

    def train_process(self, output, input_dict):
        pass

    def inference_forward(self, input_dict):
        method = input_dict["sample_method"]
        if method == "beam":
            return self.beam_search(input_dict)
        if method == "greedy":
            return self.stepwise_forward(input_dict)
        raise NotImplementedError(f"sample_method {method!r} is not built (greedy and beam are)")

    def stepwise_forward(self, input_dict):
        raise NotImplementedError

    def beam_search(self, input_dict):
        raise NotImplementedError
