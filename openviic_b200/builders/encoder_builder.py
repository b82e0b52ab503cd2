"""Reference import path ``builders.encoder_builder``; defined in ``builders/__init__.py``."""

from . import META_ENCODER, build_encoder  # noqa: F401
