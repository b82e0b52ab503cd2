"""Dual-path (region + grid) captioner: GeometricDualFeatureEmbedding -> DualCollaborativeLevelEncoder -> Decoder.

The reference ships the two modules (vision_embeddings.py:45-71, encoders.py:114-211) and two YAMLs named after the
DLCT / RSTNet papers, but no architecture class wires them together and the YAMLs name an unregistered class
(configs/rstnet.yaml:36): this class is that wiring, with the reference's conventions
(``encoder_forward`` / ``forward`` / ``step`` / ``beam_search`` of models/base_transformer.py).  Inputs are
``region_features``, ``region_boxes``, ``grid_features``, ``grid_boxes`` of an InstanceList.  The whole-path engine does
not cover the dual encoder; ``beam_search`` runs on the registered modules (the module-level CUDA path)."""

from __future__ import annotations

import torch

from ..builders.decoder_builder import build_decoder
from ..builders.encoder_builder import build_encoder
from ..builders.model_builder import META_ARCHITECTURE
from ..builders.vision_embedding_builder import build_vision_embedding
from .base_transformer import BaseTransformer


@META_ARCHITECTURE.register()
class DualCollaborativeTransformer(BaseTransformer):
    def __init__(self, config, vocab):
        super().__init__(vocab)
        self.model_config = config
        self.device = torch.device(config.DEVICE)
        self.vision_embedding = build_vision_embedding(config.VISION_EMBEDDING)
        self.encoder = build_encoder(config.ENCODER)
        self.decoder = build_decoder(config.DECODER, vocab)

    def engine_supported(self) -> bool:
        return False

    def encoder_forward(self, input_features):
        f = input_features
        (region, r_mask), (grid, g_mask), (region2all, grid2all) = self.vision_embedding(
            f.region_features, f.region_boxes, f.grid_features, f.grid_boxes)
        return self.encoder(region, f.region_boxes, r_mask, region2all, grid, f.grid_boxes, g_mask, grid2all)
