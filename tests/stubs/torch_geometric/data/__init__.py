import torch


class Data(object):
    """Attribute bag (PyG ``Data`` is used by the reference only as such)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__.keys()]


class InMemoryDataset(object):  # placeholder so that wrapper.py imports
    def __init__(self, *a, **k):
        pass


class Batch(Data):
    pass
