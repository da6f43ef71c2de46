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
        if input_dict["sample_method"] != "greedy":
            return self._sampling_stepwise_forward(input_dict)
        out = self.decoder.greedy(input_dict["attn_emb"], input_dict["attn_emb_len"], input_dict["max_length"],
                                  self.start_idx, self.end_idx, self.pad_idx,
                                  need_logit=input_dict.get("need_logit", True))
        if not input_dict.get("_device_seq", False):      # internal: the async submit() path copies later
            out["seq"] = out["seq"].cpu()
            out["sampled_logprob"] = out["sampled_logprob"].cpu()
        return out

    def _prefix_last(self, word, input_dict, rows_per_clip=1):
        """The reference's per-step decoder call (transformer_model.py:34-86): full prefix in, last position out.  `word`
        [rows_per_clip * B, t + 1] with row r belonging to clip r % B."""
        out = self.decoder({"word": word, "attn_emb": input_dict["attn_emb"], "attn_emb_len": input_dict["attn_emb_len"],
                            "cap_padding_mask": word == self.pad_idx})
        return out["logit"][:, -1], out["embed"][:, -1]

    @torch.no_grad()
    def _sampling_stepwise_forward(self, input_dict):
        """`stepwise_forward` (base.py:152-170) for the stochastic samplers (gumbel / topK / topP / temperature): the
        reference's step loop, each step one dense full-prefix decoder pass on the device followed by `sample_next_word`.
        (Greedy does not come here: it is one fused KV-cached launch.)"""
        attn_emb = input_dict["attn_emb"]
        B, dev, L = attn_emb.size(0), attn_emb.device, input_dict["max_length"]
        seq = torch.full((B, L), self.end_idx, dtype=torch.long, device=dev)
        logit = torch.zeros(B, L, self.vocab_size, device=dev)
        embed = torch.zeros(B, L, self.decoder.d_model, device=dev)
        logprob = torch.zeros(B, L, device=dev)
        start = torch.full((B, 1), self.start_idx, dtype=torch.long, device=dev)
        unfinished = torch.ones(B, dtype=torch.bool, device=dev)
        for t in range(L):
            word = start if t == 0 else torch.cat((start, seq[:, :t]), dim=-1)
            lg, em = self._prefix_last(word, input_dict)
            sampled = self.sample_next_word(lg, input_dict["sample_method"], input_dict["temp"])
            logit[:, t], embed[:, t], logprob[:, t] = lg, em, sampled["probs"]
            unfinished = unfinished & (sampled["word"] != self.end_idx)
            seq[:, t] = torch.where(unfinished, sampled["word"], torch.full_like(sampled["word"], self.end_idx))
            if not bool(unfinished.any()):
                break
        return {"seq": seq.cpu(), "logit": logit, "sampled_logprob": logprob.cpu(), "embed": embed}

    @torch.no_grad()
    def _beam_search_n_best(self, input_dict):
        """`beam_search` with `n_best` (base.py:254-361): every clip's `n_best_size` best finished beams.  All clips advance
        in lock-step through dense full-prefix decoder passes over [beam * B] rows (row r = beam r // B of clip r % B); the
        per-clip bookkeeping -- flattened top-k, -1000 penalty, `== beam_size` stop, score / (t + 1), stable sort -- is the
        reference's, kept on the host per clip."""
        attn_emb = input_dict["attn_emb"]
        B, dev = attn_emb.size(0), attn_emb.device
        L, R, temp, V = input_dict["max_length"], input_dict["beam_size"], input_dict["temp"], self.vocab_size
        n_best = input_dict["n_best_size"]
        out_seq = torch.full((B, n_best, L), self.end_idx, dtype=torch.long)
        scores = torch.zeros(R, B, device=dev)
        seq = None                                                  # [R, B, t]
        done = [[] for _ in range(B)]
        active = [True] * B
        start = torch.full((R * B, 1), self.start_idx, dtype=torch.long, device=dev)
        for t in range(L):
            word = start if t == 0 else torch.cat((start, seq.reshape(R * B, t)), dim=-1)
            lg, _ = self._prefix_last(word, input_dict, R)
            lp = torch.log_softmax(torch.log_softmax(lg, dim=1) / temp, dim=1).view(R, B, V) + scores.unsqueeze(-1)
            flat = lp[0] if t == 0 else lp.permute(1, 0, 2).reshape(B, R * V)          # per clip: [V] or [R * V]
            top, idx = flat.topk(R, 1, True, True)                                      # [B, R]
            prev, nxt = torch.div(idx, V, rounding_mode="trunc").t(), (idx % V).t()     # [R, B]
            new = nxt.unsqueeze(-1) if t == 0 else torch.cat(
                [seq.gather(0, prev.unsqueeze(-1).expand(R, B, t)), nxt.unsqueeze(-1)], dim=-1)
            is_end = (nxt == self.end_idx) | (t == L - 1)
            top_t = top.t().contiguous()
            end_h, score_h, new_h = is_end.cpu(), top_t.cpu(), new.cpu()
            for b in range(B):
                if not active[b]:
                    continue
                for r in range(R):
                    if end_h[r, b]:
                        done[b].append({"seq": new_h[r, b].clone(), "score": score_h[r, b].item() / (t + 1)})
                if len(done[b]) == R:
                    active[b] = False
            scores = torch.where(is_end, top_t - 1000, top_t)
            seq = new
            if not any(active):
                break
        for b in range(B):
            for j, beam in enumerate(sorted(done[b], key=lambda x: -x["score"])[:n_best]):
                out_seq[b, j, :len(beam["seq"])] = beam["seq"]
        return {"seq": out_seq}

    def beam_search(self, input_dict):
        if input_dict.get("n_best", False):
            return self._beam_search_n_best(input_dict)
        out = self.decoder.beam_search(input_dict["attn_emb"], input_dict["attn_emb_len"], input_dict["max_length"],
                                       input_dict["beam_size"], input_dict["temp"],
                                       self.start_idx, self.end_idx, self.pad_idx)
        if not input_dict.get("_device_seq", False):
            out["seq"] = out["seq"].cpu()
        return out
