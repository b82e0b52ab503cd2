"""openviic_b200 -- B200-native caption-generation hot path behind OpenViIC's registry/builder API.

    from openviic_b200 import build_model, get_config
    cfg = get_config("standard_transformer.yaml")
    model = build_model(cfg.MODEL, vocab)               # classes chosen by the YAML ARCHITECTURE strings
    ids, log_probs = model.beam_search(items, batch_size=B, beam_size=5, out_size=1)

All compute runs in hand-written sm_100a CUDA kernels reached through the C ABI in
``include/openviic_cap.h`` (``openviic_b200/lib/libopenviic_cap.so``); there is no CPU fallback.
"""

import os as _os

# Several engines on several streams keep independent batches in flight (bench.py, serving loops).  With the driver's
# default of 8 hardware work queues, 32 streams serialise falsely; the setting only takes effect if it is in the
# environment before the CUDA context is created, so it is set at import (an explicit user value wins).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import models  # noqa: F401,E402  (fills the registries)
from .builders import (META_ARCHITECTURE, META_ATTENTION, META_DECODER, META_ENCODER,  # noqa: F401
                       META_TEXT_EMBEDDING, META_VISION_EMBEDDING, build_attention, build_decoder,
                       build_encoder, build_model, build_text_embedding, build_vision_embedding)
from .configs import CfgNode, get_config  # noqa: F401
from .engine import CaptionEngine  # noqa: F401
from .utils.instance import Instance, InstanceList  # noqa: F401

__version__ = "0.1.0"
