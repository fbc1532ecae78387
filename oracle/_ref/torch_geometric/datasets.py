from . import _Any


class ZINC(_Any):
    pass
