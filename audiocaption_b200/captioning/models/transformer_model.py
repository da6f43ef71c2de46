"""Mirror of captioning/models/transformer_model.py:11-86 (HF copy hf_wrapper.py:845-920)."""
import random

import torch

from .base import CaptionModel
from .transformer_decoder import TransformerDecoder


class TransformerModel(CaptionModel):

    def __init__(self, encoder, decoder, **kwargs):
        if not hasattr(self, "compatible_decoders"):
            self.compatible_decoders = (TransformerDecoder,)
        super().__init__(encoder, decoder, **kwargs)

    def seq_forward(self, input_dict):
        """transformer_model.py:20-32: teacher forcing, one dense pass over cap[:, :-1]."""
        cap = input_dict["cap"]
        cap_padding_mask = (cap == self.pad_idx)[:, :-1]
        return self.decoder({"word": cap[:, :-1], "attn_emb": input_dict["attn_emb"],
                             "attn_emb_len": input_dict["attn_emb_len"], "cap_padding_mask": cap_padding_mask})

    def _train_stepwise_forward(self, input_dict):
        """Scheduled-sampling training forward (base.py:152-170 with mode == "train", transformer_model.py:34-57).

        The reference runs L decoder calls, step t on the full prefix [:t+1] -- the ground-truth prefix when that step's
        coin `random.random() < ss_ratio` says so, else <start> + its own samples -- and keeps the last position.  With a
        causal decoder, the last position of the step-t call equals position t of ONE pass over the whole token row, so
        two dense passes (ground-truth row, sampled row) give every step's logits; the sampled row is built by a
        KV-cached decode launch in between.  The coins are drawn here, one per step from python's `random`, in the
        reference's order, so a seeded run takes the same GT / sample decisions."""
        cap = input_dict["cap"]
        L = cap.size(1) - 1
        ss_ratio = input_dict["ss_ratio"]
        coins = [random.random() < ss_ratio for _ in range(L)]
        out = self.decoder.scheduled_sampling_forward(cap, input_dict["attn_emb"], input_dict["attn_emb_len"], coins,
                                                      self.start_idx, self.end_idx, self.pad_idx)
        out["seq"] = out["seq"].cpu()                           # base.py:122,126: CPU tensors
        out["sampled_logprob"] = out["sampled_logprob"].cpu()
        return out

    def stepwise_forward(self, input_dict):
        """Greedy decode.  Output keys/shapes as base.py:112-129: `seq` and `sampled_logprob` are
        CPU tensors, `logit` / `embed` stay on the device."""
        if input_dict["mode"] == "train":
            return self._train_stepwise_forward(input_dict)
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
