"""Mirror of captioning/models/__init__.py: encoder / decoder contracts (:26-92)."""
import torch.nn as nn


class BaseEncoder(nn.Module):
    """forward({"wav","wav_len","specaug"}) -> {"fc_emb","attn_emb","attn_emb_len"}
    (captioning/models/__init__.py:41-61)."""

    def __init__(self, spec_dim, fc_feat_dim, attn_feat_dim):
        super().__init__()
        self.spec_dim, self.fc_feat_dim, self.attn_feat_dim = spec_dim, fc_feat_dim, attn_feat_dim


class BaseDecoder(nn.Module):
    """Word / audio embeddings in, next-word logits out (captioning/models/__init__.py:64-92)."""

    def __init__(self, emb_dim, vocab_size, fc_emb_dim, attn_emb_dim, dropout=0.2, tie_weights=False):
        super().__init__()
        self.emb_dim, self.vocab_size = emb_dim, vocab_size
        self.fc_emb_dim, self.attn_emb_dim = fc_emb_dim, attn_emb_dim
        self.tie_weights = tie_weights
        self.word_embedding = nn.Embedding(vocab_size, emb_dim)
        self.in_dropout = nn.Dropout(dropout)
