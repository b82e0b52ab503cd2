"""Standard encoder-decoder captioners over region or grid features
(reference: models/standard_stransformer.py:11-76)."""

from __future__ import annotations

import torch

from ..builders.decoder_builder import build_decoder
from ..builders.encoder_builder import build_encoder
from ..builders.model_builder import META_ARCHITECTURE
from ..builders.vision_embedding_builder import build_vision_embedding
from .base_transformer import BaseTransformer


class _SingleStreamTransformer(BaseTransformer):
    feature_field = "region_features"

    def __init__(self, config, vocab):
        super().__init__(vocab)
        self.model_config = config
        self.device = torch.device(config.DEVICE)
        self.vision_embedding = build_vision_embedding(config.VISION_EMBEDDING)
        self.encoder = build_encoder(config.ENCODER)
        self.decoder = build_decoder(config.DECODER, vocab)

    def engine_inputs(self, input_features):
        return getattr(input_features, self.feature_field), None

    def encoder_forward(self, input_features):
        vision_features, vision_padding_mask = self.vision_embedding(getattr(input_features, self.feature_field))
        encoder_features = self.encoder(features=vision_features, padding_mask=vision_padding_mask)
        return encoder_features, vision_padding_mask


@META_ARCHITECTURE.register()
class StandardTransformerUsingRegion(_SingleStreamTransformer):
    feature_field = "region_features"


@META_ARCHITECTURE.register()
class StandardTransformerUsingGrid(_SingleStreamTransformer):
    feature_field = "grid_features"
