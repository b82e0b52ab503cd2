"""META_ENCODER registry + builder (reference: builders/encoder_builder.py:3-8)."""

from .registry import Registry

META_ENCODER = Registry("META_ENCODER")


def build_encoder(config):
    return META_ENCODER.get(config.ARCHITECTURE)(config)
