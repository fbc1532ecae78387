from . import _Any


class Data(_Any):
    pass


class Batch(_Any):
    pass


class InMemoryDataset:
    pass


class Dataset:
    pass
