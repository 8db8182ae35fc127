"""Ordered, attribute- and index-addressable result records (the behaviour the reference gets
from ``diffusers.utils.BaseOutput``: src/SegDiffEditPipeline.py:33-37, src/metrics.py:99)."""
from collections import OrderedDict
from dataclasses import fields


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                OrderedDict.__setitem__(self, f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        if name in self.keys() and value is not None:
            OrderedDict.__setitem__(self, name, value)
        super().__setattr__(name, value)

    def __reduce__(self):
        return (type(self), tuple(getattr(self, f.name) for f in fields(self)))

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())
