"""Mirror of captioning/datasets/text_tokenizer.py:9-83 `DictTokenizer`: whitespace tokens <-> ids with the four special
words first (<pad> 0, <start> 1, <end> 2, <unk> 3); its `state_dict` (word2idx) travels inside every checkpoint
(python_scripts/train_eval/base.py:231-244)."""
import pickle
from pathlib import Path

import numpy as np

from ..utils.train_util import pad_sequence


class DictTokenizer:

    def __init__(self, tokenizer_path: str = None, max_length: int = 20) -> None:
        self.word2idx, self.idx2word = {}, {}
        for word in ("<pad>", "<start>", "<end>", "<unk>"):
            self.add_word(word)
        self.loaded = False
        if tokenizer_path is not None and Path(tokenizer_path).exists():
            with open(tokenizer_path, "rb") as f:
                self.load_state_dict(pickle.load(f))
            self.loaded = True
        self.bos, self.eos, self.pad = self.word2idx["<start>"], self.word2idx["<end>"], self.word2idx["<pad>"]
        self.max_length = max_length

    def add_word(self, word):
        if word not in self.word2idx:
            idx = len(self.word2idx)
            self.word2idx[word] = idx
            self.idx2word[idx] = word

    def encode_word(self, word):
        return self.word2idx.get(word, self.word2idx["<unk>"])

    def __call__(self, texts):
        assert isinstance(texts, list), "the input must be List[str]"
        batch = []
        for text in texts:
            ids = [self.encode_word(tok) for tok in text.split()][:self.max_length]
            batch.append(np.array([self.bos] + ids + [self.eos]))
        caps, cap_lens = pad_sequence(batch, self.pad)
        return {"cap": caps, "cap_len": cap_lens}

    def decode(self, batch_token_ids):
        out = []
        for token_ids in batch_token_ids:
            words = []
            for token_id in token_ids:
                token_id = int(token_id)
                if token_id == self.eos:
                    break
                if token_id != self.bos:
                    words.append(self.idx2word[token_id])
            out.append(" ".join(words))
        return out

    def __len__(self):
        return len(self.word2idx)

    def state_dict(self):
        return self.word2idx

    def load_state_dict(self, state_dict):
        self.word2idx = dict(state_dict)
        self.idx2word = {idx: word for word, idx in self.word2idx.items()}
