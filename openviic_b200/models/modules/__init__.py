from .attentions import *  # noqa: F401,F403
from .encoders import *  # noqa: F401,F403
from .decoders import *  # noqa: F401,F403
from .text_embeddings import *  # noqa: F401,F403
from .vision_embeddings import *  # noqa: F401,F403
