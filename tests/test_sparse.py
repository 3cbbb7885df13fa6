"""Brick-sparse grids (molvoxel_b200/sparse.py): host-side scatter, and on the GPU the compaction kernel itself."""
import numpy as np
import pytest
import torch

import molvoxel_b200 as mv
from molvoxel_b200.sparse import SparseGrids
from tests.helpers import ligand_batch


def _bricks_of(dense):
    """Reference compaction on the host: every 8x8x8 brick (zero-padded at the border) that holds a non-zero value."""
    B, C, D = dense.shape[:3]
    nb = -(-D // 8)
    pad = np.zeros((B, C, nb * 8, nb * 8, nb * 8), np.float32)
    pad[:, :, :D, :D, :D] = dense
    blocks = pad.reshape(B * C, nb, 8, nb, 8, nb, 8).transpose(0, 1, 3, 5, 2, 4, 6).reshape(-1, 512)
    keep = np.flatnonzero((blocks != 0).any(1))
    return keep.astype(np.int64), blocks[keep]


@pytest.mark.parametrize("D", [8, 20, 24])
def test_to_dense_rebuilds_the_grid(D):
    rng = np.random.default_rng(D)
    dense = np.zeros((3, 2, D, D, D), np.float32)
    for _ in range(6):
        m, c = rng.integers(0, 3), rng.integers(0, 2)
        x, y, z = rng.integers(0, D - 3, size=3)
        dense[m, c, x:x + 3, y:y + 2, z:z + 3] = rng.uniform(0.5, 2, size=(3, 2, 3))
    ids, vals = _bricks_of(dense)
    perm = rng.permutation(len(ids))   # slot order is arbitrary
    sp = SparseGrids(ids[perm].astype(np.int32), vals[perm], (3, 2, D))
    assert np.array_equal(sp.to_dense(), dense) and sp.num_bricks == len(ids)
    spt = SparseGrids(torch.from_numpy(ids[perm].astype(np.int32)), torch.from_numpy(vals[perm]), (3, 2, D))
    assert np.array_equal(spt.to_dense().numpy(), dense)
    mol, ch, bx, by, bz = sp.split_ids()
    assert mol.max() < 3 and ch.max() < 2 and max(bx.max(), by.max(), bz.max()) < -(-D // 8)


@pytest.mark.gpu
@pytest.mark.parametrize("D,res", [(64, 0.5), (36, 0.5), (21, 0.8)], ids=["d64", "d36_vec_not_brick_multiple", "d21_scalar_loads"])
def test_compaction_is_lossless_and_minimal(D, res):
    """compact() keeps exactly the non-empty bricks of the dense output (same set, same values), for grids whose size
    is / is not a multiple of the brick and of the 16-byte load width."""
    rng = np.random.default_rng(5)
    B, C = 24, 5
    offs, coords, types = ligand_batch(rng, B, C)
    vox = mv.create_voxelizer(res, D, "scalar", "gaussian", library="b200")
    out = vox.forward_types_batch(coords, offs, None, types, 1.0, C)
    sp = vox.compact(out)
    dense = out.cpu().numpy()
    assert np.array_equal(sp.to_dense().cpu().numpy(), dense)
    want_ids, want_vals = _bricks_of(dense)
    host = sp.cpu()
    order = np.argsort(host.ids.astype(np.int64) & 0xFFFFFFFF)
    assert np.array_equal((host.ids.astype(np.int64) & 0xFFFFFFFF)[order], want_ids)
    assert np.array_equal(host.vals[order], want_vals)
    if D == 64:   # a ligand in a 32 A box: a few per cent of the bricks
        assert sp.nbytes < 0.1 * dense.nbytes
    # a too small capacity is detected and the buffers grow
    assert vox.compact(out, capacity=7).num_bricks == sp.num_bricks
    # a grid that is not the last call's output (no column occupancy to lean on): every column is scanned
    other = out.clone()
    other[3, 2, 1, 1, 1] = 5.0
    sp2 = vox.compact(other)
    assert torch.equal(sp2.to_dense(), other)


@pytest.mark.gpu
def test_compaction_of_dense_pockets_and_empty_batches():
    rng = np.random.default_rng(6)
    V = 1500
    coords = rng.uniform(-11, 11, size=(2 * V, 3))
    feats = rng.uniform(size=(2 * V, 16)).astype(np.float32)
    vox = mv.create_voxelizer(0.5, 48, "scalar", "gaussian", library="b200")
    out = vox.forward_features_batch(coords, np.array([0, V, 2 * V], dtype=np.int32), None, feats, 1.0)
    assert torch.equal(vox.compact(out).to_dense(), out)
    far = vox.forward_single_batch(coords + 500.0, np.array([0, V, 2 * V], dtype=np.int32), None, 1.0)
    sp = vox.compact(far)
    assert sp.num_bricks == 0 and not sp.to_dense().any()
