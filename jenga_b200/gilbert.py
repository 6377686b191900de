"""Drop-in for the table builders of the reference's gilbert.py.

Same names, arguments and return types as gilbert.py:442 gilbert_mapping, :332
sliced_gilbert_mapping, :597 gilbert_block_neighbor_mapping and :679
sliced_gilbert_block_neighbor_mapping (lists of Python ints / a bool torch tensor), computed by
the O(N) curve walker in libjenga_b200.so (csrc/gilbert.cpp) instead of one recursive Python
call per voxel.  `transpose_gilbert_mapping` / `transpose_order` (gilbert.py:274) are axis
permutations of the walker's table.
"""
from __future__ import annotations

import ctypes as C

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


def _check_order(order):
    order = [0, 1, 2] if order is None else [int(o) for o in order]
    if len(order) != 3 or set(order) != {0, 1, 2}:
        raise ValueError("order must be a permutation of 0,1,2")   # gilbert.py:294-295
    return order


def _transposed_np(dims, order):
    """gilbert.py:274-330: the curve is drawn in the box dims[order] and voxel `coords` of the
    original box reads the index at coords[order]."""
    if len(dims) != 3:
        raise ValueError("Dimensions must be three-dimensional")       # gilbert.py:287-288
    order = _check_order(order)
    tt, hh, ww = (int(dims[o]) for o in order)
    l2h_t, _ = _mapping_np(tt, hh, ww, False)
    inv = np.argsort(order)
    l2h = np.ascontiguousarray(l2h_t.reshape(tt, hh, ww).transpose(inv)).reshape(-1)
    h2l = np.empty_like(l2h)
    h2l[l2h] = np.arange(l2h.size, dtype=np.int64)
    return l2h, h2l


def transpose_gilbert_mapping(dims, order=None):
    l2h, h2l = _transposed_np(dims, order)
    return l2h.tolist(), h2l.tolist()


def gilbert_xyz2d(x, y, z, width, height, depth):
    return int(lib.jenga_gilbert_xyz2d(x, y, z, width, height, depth))


def gilbert_mapping(t, h, w, transpose_order=None):
    if transpose_order is not None:
        return transpose_gilbert_mapping([t, h, w], transpose_order)    # gilbert.py:484-486
    l2h, h2l = _mapping_np(t, h, w, False)
    return l2h.tolist(), h2l.tolist()


def sliced_gilbert_mapping(t, h, w, transpose_order=None):
    if transpose_order is not None:
        return transpose_gilbert_mapping([t, h, w], transpose_order)    # gilbert.py:436-438
    l2h, h2l = _mapping_np(t, h, w, True)
    return l2h.tolist(), h2l.tolist()


def gilbert_block_neighbor_mapping(t, h, w, block_size=128, transpose_order=None):
    # the reference accepts transpose_order here and never reads it (gilbert.py:597-677)
    return block_neighbor_mapping(t, h, w, block_size, False)


def sliced_gilbert_block_neighbor_mapping(t, h, w, block_size=128, transpose_order=None):
    if transpose_order is None:
        return block_neighbor_mapping(t, h, w, block_size, True)
    l2h, _ = _transposed_np([t, h, w], transpose_order)                 # gilbert.py:704
    n = t * h * w
    nb = (n + block_size - 1) // block_size
    out = np.empty((nb, nb), dtype=np.uint8)
    check(lib.jenga_block_neighbors_from_mapping_host(t, h, w, block_size, l2h.ctypes.data, out.ctypes.data),
          "block_neighbors_from_mapping")
    return torch.from_numpy(out).bool()


def block_neighbor_csr(t: int, h: int, w: int, block_size: int = 128, sliced: bool = False):
    """The adjacency as CSR (row_ptr int32 [nb+1], col_idx int32 [nnz]) — SURVEY §8 f-4."""
    n = t * h * w
    nb = (n + block_size - 1) // block_size
    row_ptr = np.empty(nb + 1, dtype=np.int32)
    nnz = C.c_int64(0)
    check(lib.jenga_gilbert_block_neighbors_csr_host(t, h, w, block_size, int(sliced), row_ptr.ctypes.data,
                                                     None, 0, C.byref(nnz)), "block_neighbors_csr")
    col = np.empty(nnz.value, dtype=np.int32)
    check(lib.jenga_gilbert_block_neighbors_csr_host(t, h, w, block_size, int(sliced), row_ptr.ctypes.data,
                                                     col.ctypes.data, nnz.value, C.byref(nnz)),
          "block_neighbors_csr")
    return torch.from_numpy(row_ptr), torch.from_numpy(col)


def mapping_tensors_device(t: int, h: int, w: int, sliced: bool = False, device="cuda"):
    """(linear_to_hilbert, hilbert_order) int64 tensors built ON `device` (no host walk, no H2D)."""
    import torch.cuda
    dev = torch.device(device)
    n = t * h * w
    l2h = torch.empty(n, dtype=torch.int64, device=dev)
    h2l = torch.empty(n, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.jenga_gilbert_mapping_device(t, h, w, int(sliced), l2h.data_ptr(), h2l.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "gilbert_mapping_device")
    return l2h, h2l


def block_neighbor_bits_device(t: int, h: int, w: int, linear_to_hilbert: torch.Tensor, block_size: int = 128):
    """Packed adjacency bit rows int32 [nb, ceil(nb/32)] on the device of `linear_to_hilbert` — what
    attention.neighbour_bits() produces from the dense bool matrix, without the matrix."""
    if not linear_to_hilbert.is_cuda or linear_to_hilbert.dtype != torch.int64 or linear_to_hilbert.numel() != t * h * w:
        raise ValueError("linear_to_hilbert must be a CUDA int64 tensor of t*h*w entries")
    dev = linear_to_hilbert.device
    nb = (t * h * w + block_size - 1) // block_size
    words = (nb + 31) // 32
    bits = torch.empty((nb, words), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.jenga_block_neighbor_bits_device(t, h, w, block_size, linear_to_hilbert.contiguous().data_ptr(),
                                                   bits.data_ptr(), words, torch.cuda.current_stream(dev).cuda_stream),
              "block_neighbor_bits_device")
    return bits


def _not_on_the_hot_path(name):
    def f(*a, **k):
        raise NotImplementedError(f"gilbert.{name} is a plotting/experiment helper no Jenga script calls")
    f.__name__ = name
    return f


# imported by name in no script, defined so that `from gilbert import *`-style uses fail at CALL time
block_wise_mapping = _not_on_the_hot_path("block_wise_mapping")
visualize_gilbert_curve = _not_on_the_hot_path("visualize_gilbert_curve")
visualize_gilbert_curves_comparison = _not_on_the_hot_path("visualize_gilbert_curves_comparison")
