from .instance import Instance, InstanceList  # noqa: F401
