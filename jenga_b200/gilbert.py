"""Drop-in for the table builders of the reference's gilbert.py.

Same names, arguments and return types as gilbert.py:442 gilbert_mapping, :332
sliced_gilbert_mapping, :597 gilbert_block_neighbor_mapping and :679
sliced_gilbert_block_neighbor_mapping (lists of Python ints / a bool torch tensor), computed by
the O(N) curve walker in libjenga_b200.so (csrc/gilbert.cpp) instead of one recursive Python
call per voxel.  `transpose_order` (never passed by the reference's scripts) is not built.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib


def _mapping_np(t: int, h: int, w: int, sliced: bool):
    n = t * h * w
    l2h = np.empty(n, dtype=np.int64)
    h2l = np.empty(n, dtype=np.int64)
    check(lib.jenga_gilbert_mapping_host(t, h, w, int(sliced), l2h.ctypes.data, h2l.ctypes.data),
          "gilbert_mapping")
    return l2h, h2l


def mapping_tensors(t: int, h: int, w: int, sliced: bool = False):
    """(linear_to_hilbert, hilbert_order) as int64 tensors."""
    l2h, h2l = _mapping_np(t, h, w, sliced)
    return torch.from_numpy(l2h), torch.from_numpy(h2l)


def block_neighbor_mapping(t: int, h: int, w: int, block_size: int = 128, sliced: bool = False) -> torch.Tensor:
    n = t * h * w
    nb = (n + block_size - 1) // block_size
    out = np.empty((nb, nb), dtype=np.uint8)
    check(lib.jenga_gilbert_block_neighbors_host(t, h, w, block_size, int(sliced), out.ctypes.data),
          "gilbert_block_neighbors")
    return torch.from_numpy(out).bool()


def _no_transpose(transpose_order):
    if transpose_order is not None:
        raise NotImplementedError("transpose_order is not used by any Jenga script and is not built")


def gilbert_xyz2d(x, y, z, width, height, depth):
    return int(lib.jenga_gilbert_xyz2d(x, y, z, width, height, depth))


def gilbert_mapping(t, h, w, transpose_order=None):
    _no_transpose(transpose_order)
    l2h, h2l = _mapping_np(t, h, w, False)
    return l2h.tolist(), h2l.tolist()


def sliced_gilbert_mapping(t, h, w, transpose_order=None):
    _no_transpose(transpose_order)
    l2h, h2l = _mapping_np(t, h, w, True)
    return l2h.tolist(), h2l.tolist()


def gilbert_block_neighbor_mapping(t, h, w, block_size=128, transpose_order=None):
    _no_transpose(transpose_order)
    return block_neighbor_mapping(t, h, w, block_size, False)


def sliced_gilbert_block_neighbor_mapping(t, h, w, block_size=128, transpose_order=None):
    _no_transpose(transpose_order)
    return block_neighbor_mapping(t, h, w, block_size, True)
