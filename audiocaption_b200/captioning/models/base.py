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
                forward_dict["n_best"] = input_dict.get("n_best", False)
                forward_dict["n_best_size"] = input_dict.get("n_best_size", forward_dict["beam_size"])
            elif forward_dict["sample_method"] == "dbs":
                raise NotImplementedError("diverse beam search (base.py:363-471) is not built on the B200 path")
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
        if input_dict["sample_method"] == "beam":
            return self.beam_search(input_dict)
        return self.stepwise_forward(input_dict)          # greedy (one fused launch) or a sampling method (step loop)

    def sample_next_word(self, logit, method, temp):
        """base.py:214-252 on device tensors: greedy arg-max, Gumbel-max, top-k ("top5"), nucleus ("top0.9") and plain
        temperature sampling.  logit [N, V] -> {"word" [N] i64, "probs" [N] log-probability of the drawn word}."""
        logprob = torch.log_softmax(logit, dim=1)
        if method == "greedy":
            sampled_logprob, word = torch.max(logprob, 1)
            return {"word": word, "probs": sampled_logprob}
        if method == "gumbel":
            u = torch.rand_like(logprob)
            y = logprob - torch.log(-torch.log(u + 1e-20) + 1e-20)
            word = torch.log_softmax(y / temp, dim=-1).argmax(1)
            return {"word": word, "probs": logprob.gather(1, word.unsqueeze(-1)).squeeze(1)}
        logprob = logprob / temp
        if method.startswith("top"):
            top_num = float(method[3:])
            if 0 < top_num < 1:                        # nucleus: the shortest prefix of the sorted distribution reaching top_num
                sorted_probs, sorted_indices = torch.sort(torch.softmax(logit, dim=1), descending=True, dim=1)
                keep = sorted_probs.cumsum(1) < top_num
                keep = torch.cat([torch.ones_like(keep[:, :1]), keep[:, :-1]], 1)
                sorted_probs = sorted_probs * keep.to(sorted_probs)
                sorted_probs = sorted_probs / sorted_probs.sum(1, keepdim=True)
                logprob = logprob.scatter(1, sorted_indices, sorted_probs.log())
            else:                                      # top-k
                topk, indices = torch.topk(logprob, int(top_num), dim=1)
                logprob = torch.full_like(logprob, float("-inf")).scatter(1, indices, topk)
        word = torch.multinomial(torch.softmax(logprob, dim=1), 1).squeeze(1)
        return {"word": word, "probs": logprob.gather(1, word.unsqueeze(-1)).squeeze(1)}

    def stepwise_forward(self, input_dict):
        raise NotImplementedError

    def beam_search(self, input_dict):
        raise NotImplementedError
