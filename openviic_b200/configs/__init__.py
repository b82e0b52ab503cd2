from .utils import CfgNode, get_config  # noqa: F401
