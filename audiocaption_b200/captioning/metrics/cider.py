"""CIDEr, the score `Runner._eval_epoch` monitors (python_scripts/train_eval/run.py:12,150-155 of the reference:
`from pycocoevalcap.cider.cider import Cider`; `captioning/utils/model_util.py:134-135` uses the same scorer).

pycocoevalcap is a third-party dependency of the reference that is absent from this image (and its PTB tokenizer needs
Java), so this file restates the published algorithm of `pycocoevalcap.cider` (Vedantam et al., CIDEr-D as shipped in
coco-caption: n-grams up to 4, TF-IDF weights with document frequencies over the reference sets, clipped cosine
similarity, Gaussian length penalty sigma = 6, mean over n and over references, x 10) -- including its quirks: the
caption "length" is the number of BIGRAMS, and log(df) is clamped at 0.  Host-side Python, as in the reference: it is not
part of the device path.  PARITY UNPINNED: nothing executable of pycocoevalcap is available to check against; the tests
pin the properties the definition implies (tests/test_boundary_cpu.py).

`simple_tokenize` stands in for the PTB tokenizer (lower-case, punctuation tokens dropped -- the same punctuation list
pycocoevalcap removes); the captions of AudioCaps / Clotho are plain lower-case sentences, for which the two agree."""
import math
import re
from collections import defaultdict

import numpy as np

_PUNCT = {"''", "'", "``", "`", "-lrb-", "-rrb-", "-lcb-", "-rcb-", "(", ")", "{", "}", ".", "?", "!", ",", ":", "-", "--", "...", ";",
          '"'}
_TOKEN = re.compile(r"[a-z0-9]+(?:'[a-z]+)?|\.\.\.|--|[^\sa-z0-9]")


def simple_tokenize(caption: str) -> str:
    """Lower-case, split words from punctuation, drop the punctuation tokens; returns the space-joined token string."""
    return " ".join(t for t in _TOKEN.findall(caption.lower()) if t not in _PUNCT)


def _ngrams(sentence: str, n: int):
    words = sentence.split()
    counts = defaultdict(int)
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            counts[tuple(words[i:i + k])] += 1
    return counts


class Cider:
    """`compute_score(gts, res)`: gts {key: [reference strings]}, res {key: [one candidate string]} (already tokenized,
    space-separated) -> (corpus score, per-key scores in the order of gts' keys)."""

    def __init__(self, n: int = 4, sigma: float = 6.0):
        self._n = n
        self._sigma = sigma

    def method(self):
        return "CIDEr"

    def compute_score(self, gts, res):
        if gts.keys() != res.keys():
            raise ValueError("Cider.compute_score: references and candidates must have the same keys")
        n = self._n
        keys = list(gts.keys())
        tests, refsets = [], []
        for key in keys:
            hypo, refs = res[key], gts[key]
            if not (isinstance(hypo, (list, tuple)) and len(hypo) == 1 and isinstance(refs, (list, tuple)) and len(refs) >= 1):
                raise ValueError(f"Cider.compute_score: key {key!r} needs exactly one candidate and at least one reference")
            tests.append(_ngrams(hypo[0], n))
            refsets.append([_ngrams(r, n) for r in refs])
        # document frequency of an n-gram = number of reference SETS that contain it
        df = defaultdict(float)
        for refs in refsets:
            for ngram in {g for ref in refs for g in ref}:
                df[ngram] += 1.0
        log_docs = np.log(float(len(refsets)))

        def vectorize(counts):
            vec = [defaultdict(float) for _ in range(n)]
            norm = [0.0] * n
            length = 0
            for ngram, tf in counts.items():
                k = len(ngram) - 1
                vec[k][ngram] = float(tf) * (log_docs - np.log(max(1.0, df[ngram])))
                norm[k] += vec[k][ngram] ** 2
                if k == 1:
                    length += tf                      # the shipped scorer counts bigrams here
            return vec, [np.sqrt(x) for x in norm], length

        def similarity(vh, vr, nh, nr, lh, lr):
            delta = float(lh - lr)
            val = np.zeros(n)
            for k in range(n):
                for ngram, w in vh[k].items():
                    val[k] += min(w, vr[k][ngram]) * vr[k][ngram]            # clipped
                if nh[k] != 0 and nr[k] != 0:
                    val[k] /= nh[k] * nr[k]
                assert not math.isnan(val[k])
                val[k] *= np.e ** (-(delta ** 2) / (2 * self._sigma ** 2))
            return val

        scores = []
        for test, refs in zip(tests, refsets):
            vh, nh, lh = vectorize(test)
            score = np.zeros(n)
            for ref in refs:
                vr, nr, lr = vectorize(ref)
                score += similarity(vh, vr, nh, nr, lh, lr)
            scores.append(np.mean(score) / len(refs) * 10.0)
        scores = np.array(scores)
        return float(np.mean(scores)) if len(scores) else 0.0, scores
