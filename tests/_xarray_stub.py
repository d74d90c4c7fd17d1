"""A minimal stand-in for the parts of xarray the layer accessors touch (xarray is not installed
in this image): ``register_dataset_accessor`` and a Dataset with ``.coords[name] = (dims, array)``,
attribute access to coordinates / variables (``.values``, ``.size``), ``data_vars`` and
``__getitem__``. Test infrastructure only."""

import types

import numpy as np


class _Var:
    def __init__(self, values):
        self.values = np.asarray(values)

    @property
    def size(self):
        return self.values.size


class _Coords(dict):
    def __setitem__(self, name, value):
        if isinstance(value, tuple):  # (dims, array)
            value = value[1]
        super().__setitem__(name, _Var(value))


class Dataset:
    _accessors = {}

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.coords = _Coords()
        for k, v in (coords or {}).items():
            self.coords[k] = v
        self.data_vars = {k: _Var(v[1] if isinstance(v, tuple) else v) for k, v in (data_vars or {}).items()}
        self.attrs = dict(attrs or {})

    def __getitem__(self, name):
        return self.data_vars[name] if name in self.data_vars else self.coords[name]

    def __getattr__(self, name):
        if name in type(self)._accessors:
            acc = type(self)._accessors[name](self)
            object.__setattr__(self, name, acc)
            return acc
        for table in (self.__dict__.get("coords", {}), self.__dict__.get("data_vars", {})):
            if name in table:
                return table[name]
        raise AttributeError(name)


def register_dataset_accessor(name):
    def decorator(cls):
        Dataset._accessors[name] = cls
        return cls

    return decorator


def module():
    m = types.ModuleType("xarray")
    m.Dataset = Dataset
    m.register_dataset_accessor = register_dataset_accessor
    return m
