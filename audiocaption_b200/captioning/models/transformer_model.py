"""Mirror of captioning/models/transformer_model.py:11-86 (HF copy hf_wrapper.py:845-920)."""
import torch

from .base import CaptionModel
from .transformer_decoder import TransformerDecoder


class TransformerModel(CaptionModel):

    def __init__(self, encoder, decoder, **kwargs):
        if not hasattr(self, "compatible_decoders"):
            self.compatible_decoders = (TransformerDecoder,)
        super().__init__(encoder, decoder, **kwargs)

    def stepwise_forward(self, input_dict):
        """Greedy decode.  Output keys/shapes as base.py:112-129: `seq` and `sampled_logprob` are
        CPU tensors, `logit` / `embed` stay on the device."""
        out = self.decoder.greedy(input_dict["attn_emb"], input_dict["attn_emb_len"], input_dict["max_length"],
                                  self.start_idx, self.end_idx, self.pad_idx,
                                  need_logit=input_dict.get("need_logit", True))
        if not input_dict.get("_device_seq", False):      # internal: the async submit() path copies later
            out["seq"] = out["seq"].cpu()
            out["sampled_logprob"] = out["sampled_logprob"].cpu()
        return out

    def beam_search(self, input_dict):
        out = self.decoder.beam_search(input_dict["attn_emb"], input_dict["attn_emb_len"], input_dict["max_length"],
                                       input_dict["beam_size"], input_dict["temp"],
                                       self.start_idx, self.end_idx, self.pad_idx)
        if not input_dict.get("_device_seq", False):
            out["seq"] = out["seq"].cpu()
        return out
