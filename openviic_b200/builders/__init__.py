"""Registries and ``build_*`` functions selected by the YAML ``ARCHITECTURE`` strings
(reference: builders/registry.py:8-90 and the six ``builders/*_builder.py`` files).

One table defines the six plug-in points -- registry name (as it appears in the reference's error messages), and
whether the constructor also receives the vocabulary; the registries and builder functions are generated from it.
``builders/<point>_builder.py`` re-export them under the reference's import paths.  Importing ``openviic_b200.models``
fills the registries (the reference gets the same effect from the star-imports in its builders/__init__.py:1-2).
"""

import torch

from .registry import Registry  # noqa: F401

#  plug-in point      registry name            constructor takes vocab
_POINTS = {
    "attention":        ("META_ATTENTION",        False),
    "encoder":          ("ENCODER_LAYER",         False),
    "decoder":          ("DECODER_LAYER",         True),
    "vision_embedding": ("META_VISION_EMBEDDING", False),
    "text_embedding":   ("TEXT_EMBEDDING",        True),
    "model":            ("ARCHITECTURE",          True),
}


def _make_builder(registry: Registry, point: str, with_vocab: bool):
    if with_vocab:
        def build(config, vocab):
            return registry.get(config.ARCHITECTURE)(config, vocab)
    else:
        def build(config):
            return registry.get(config.ARCHITECTURE)(config)
    build.__name__ = build.__qualname__ = f"build_{point}"
    build.__doc__ = f"Instantiate the class that ``config.ARCHITECTURE`` names in the {registry._name} registry."
    return build


_REGISTRIES = {point: Registry(name) for point, (name, _) in _POINTS.items()}
_BUILDERS = {point: _make_builder(_REGISTRIES[point], point, with_vocab) for point, (_, with_vocab) in _POINTS.items()}

META_ATTENTION, build_attention = _REGISTRIES["attention"], _BUILDERS["attention"]
META_ENCODER, build_encoder = _REGISTRIES["encoder"], _BUILDERS["encoder"]
META_DECODER, build_decoder = _REGISTRIES["decoder"], _BUILDERS["decoder"]
META_VISION_EMBEDDING, build_vision_embedding = _REGISTRIES["vision_embedding"], _BUILDERS["vision_embedding"]
META_TEXT_EMBEDDING, build_text_embedding = _REGISTRIES["text_embedding"], _BUILDERS["text_embedding"]
META_ARCHITECTURE = _REGISTRIES["model"]


def build_model(config, vocab):
    """The architecture named by ``config.ARCHITECTURE``, moved to ``config.DEVICE`` (builders/model_builder.py:6-10)."""
    return _BUILDERS["model"](config, vocab).to(torch.device(config.DEVICE))
