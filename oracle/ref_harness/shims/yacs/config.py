"""Import shim: minimal attribute-dict stand-in for yacs.config.CfgNode (absent in this image)."""


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        for key, value in (init_dict or {}).items():
            self[key] = CfgNode(value) if isinstance(value, dict) else value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value
