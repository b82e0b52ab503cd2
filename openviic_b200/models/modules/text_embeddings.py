"""Word embedding (reference: models/modules/text_embeddings.py:7-31)."""

from __future__ import annotations

from torch import nn

from ...builders.text_embedding_builder import META_TEXT_EMBEDDING
from ..utils import generate_padding_mask, generate_sequential_mask


@META_TEXT_EMBEDDING.register()
class UsualEmbedding(nn.Module):
    def __init__(self, config, vocab):
        super().__init__()
        self.padding_idx = vocab.padding_idx
        if config.WORD_EMBEDDING is not None:
            raise NotImplementedError("pretrained word vectors need the reference's download/cache layer (out of scope)")
        self.components = nn.Embedding(len(vocab), config.D_MODEL, vocab.padding_idx)

    def forward(self, tokens):
        padding_masks = generate_padding_mask(tokens, padding_idx=self.padding_idx).to(tokens.device)
        sequential_masks = generate_sequential_mask(tokens.shape[-1]).to(tokens.device)
        return self.components(tokens), (padding_masks, sequential_masks)
