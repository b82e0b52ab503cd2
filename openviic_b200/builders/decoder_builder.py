"""Reference import path ``builders.decoder_builder``; defined in ``builders/__init__.py``."""

from . import META_DECODER, build_decoder  # noqa: F401
