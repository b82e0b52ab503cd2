"""META_DECODER registry + builder (reference: builders/decoder_builder.py:3-8)."""

from .registry import Registry

META_DECODER = Registry("META_DECODER")


def build_decoder(config, vocab):
    return META_DECODER.get(config.ARCHITECTURE)(config, vocab)
