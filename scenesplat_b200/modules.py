"""Containers of the Point pipeline.

`PointSequential` chains three kinds of children and hands each the view of the data it works on -- the dispatch
rule of pointcept/models/modules.py:57-91, restated:

  PointModule       gets the Point itself and returns a Point
  sparse-conv layer gets `point.sparse_conv_feat` with fp32 features (NOT `point.feat`; this is what makes the
                    first decoder Block's xCPE conv read the skip projection only); its
                    output becomes both `sparse_conv_feat` and `feat`
  any other module  gets the feature matrix: `point.feat` (mirrored into `sparse_conv_feat` when one exists), the
                    features of a bare SparseConvTensor (skipped when it has no rows), or the input as it is

Children are registered under "0", "1", .. or under the names given (keyword arguments, an OrderedDict, or
`add(module, name)`), so `state_dict` keys come out exactly as in the reference.
"""
from __future__ import annotations

import itertools
from collections import OrderedDict

import torch.nn as nn

from . import spconv_compat as spconv
from .structure import Point


class PointModule(nn.Module):
    """Marker base class: a module whose forward takes and returns a `Point`."""


def _apply(module, data):
    """One dispatch step of PointSequential.forward."""
    if isinstance(module, PointModule):
        return module(data)
    is_point = isinstance(data, Point)
    if spconv.is_spconv_module(module):
        if not is_point:
            return module(data)
        sp = data.sparse_conv_feat
        data.sparse_conv_feat = module(sp.replace_feature(sp.features.float()))
        data.feat = data.sparse_conv_feat.features
        return data
    if is_point:
        data.feat = module(data.feat)
        if "sparse_conv_feat" in data.keys():
            data.sparse_conv_feat = data.sparse_conv_feat.replace_feature(data.feat)
        return data
    if isinstance(data, spconv.SparseConvTensor):
        return data.replace_feature(module(data.features)) if data.features.shape[0] else data
    return module(data)


class PointSequential(PointModule):
    def __init__(self, *modules, **named):
        super().__init__()
        if len(modules) == 1 and isinstance(modules[0], OrderedDict):
            named = {**modules[0], **named}
            modules = ()
        for module in modules:
            self.add(module)
        for name, module in named.items():
            if name in self._modules:  # the reference raises ValueError here and KeyError in add()
                raise ValueError(f"PointSequential already has a child named {name!r}")
            self.add(module, name)

    def add(self, module, name=None):
        name = str(len(self._modules)) if name is None else name
        if name in self._modules:
            raise KeyError(f"PointSequential already has a child named {name!r}")
        self.add_module(name, module)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, idx):
        n = len(self)
        if not -n <= idx < n:
            raise IndexError(f"index {idx} is out of range for {n} children")
        return next(itertools.islice(self._modules.values(), idx % n, None))

    def forward(self, input):
        for module in self._modules.values():
            input = _apply(module, input)
        return input
