"""f-4: curve tables and block adjacency generated on the device == the host walker's tables
(which are SHA-pinned against the reference, tests/golden/gilbert.json)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

GRIDS = [(4, 6, 8, False), (5, 7, 9, False), (7, 3, 1, False), (1, 1, 5, False), (3, 5, 4, True), (8, 11, 13, True),
         (2, 9, 9, True), (32, 22, 40, False), (32, 33, 60, False), (32, 45, 80, False), (21, 30, 52, True),
         (21, 45, 80, True), (21, 22, 39, True)]


@pytest.mark.parametrize("t,h,w,sliced", GRIDS)
def test_device_tables_equal_host_tables(t, h, w, sliced):
    from jenga_b200 import gilbert as G
    from jenga_b200.attention import mask_onehot_to_bits
    l2h, h2l = G.mapping_tensors(t, h, w, sliced)
    dl2h, dh2l = G.mapping_tensors_device(t, h, w, sliced)
    torch.cuda.synchronize()
    assert torch.equal(dl2h.cpu(), l2h) and torch.equal(dh2l.cpu(), h2l)
    dense = G.block_neighbor_mapping(t, h, w, 128, sliced)
    bits = G.block_neighbor_bits_device(t, h, w, dl2h)
    want = mask_onehot_to_bits(dense.cuda())
    assert torch.equal(bits, want)
    if t * h * w >= 64:                         # another block size
        bits16 = G.block_neighbor_bits_device(t, h, w, dl2h, block_size=16)
        want16 = mask_onehot_to_bits(G.block_neighbor_mapping(t, h, w, 16, sliced).cuda())
        assert torch.equal(bits16, want16)
