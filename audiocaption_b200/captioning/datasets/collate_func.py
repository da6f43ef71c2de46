"""Mirror of captioning/datasets/collate_func.py: `VarLenPadCollate`-style padding of the waveform entries of a batch and
`TextCollate` on top of it (tokenise the captions, optionally sort the batch by caption length, longest first)."""
import numpy as np
import torch

from ..utils.train_util import pad_sequence


class VarLenPadCollate:
    """data_batch: list of dicts.  Entries named in `pad_keys` are padded to the longest item and get a `<key>_len`
    companion; everything else is stacked (arrays) or listed (collate_func.py:8-43)."""

    def __init__(self, pad_keys=("wav",), sort_key=None):
        self.pad_keys = list(pad_keys)
        self.sort_key = sort_key

    def __call__(self, data_batch):
        if self.sort_key:
            data_batch = sorted(data_batch, key=lambda x: len(x[self.sort_key]), reverse=True)
        out = {}
        for key in data_batch[0]:
            values = [item[key] for item in data_batch]
            if key in self.pad_keys:
                padded, lens = pad_sequence(values)
                out[key], out[f"{key}_len"] = padded, lens
            elif isinstance(values[0], (np.ndarray, torch.Tensor)):
                out[key] = torch.as_tensor(np.array(values))
            else:
                out[key] = values
        return out


class TextCollate(VarLenPadCollate):
    """collate_func.py:46-84: the `text_key` entries are tokenised into `cap` / `cap_len`."""

    def __init__(self, tokenizer, text_key="caption", pad_keys=("wav",), sort_key="caption"):
        super().__init__(pad_keys, sort_key)
        self.tokenizer = tokenizer
        self.text_key = text_key

    def __call__(self, data_batch):
        if self.sort_key:
            data_batch = sorted(data_batch, key=lambda x: len(x[self.sort_key].split()), reverse=True)
            sort_key, self.sort_key = self.sort_key, None
            try:
                out = super().__call__(data_batch)
            finally:
                self.sort_key = sort_key
        else:
            out = super().__call__(data_batch)
        out.update(self.tokenizer(out[self.text_key]))
        return out
