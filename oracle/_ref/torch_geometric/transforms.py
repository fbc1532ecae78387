from . import _Any


class BaseTransform(_Any):
    pass
