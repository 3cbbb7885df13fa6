"""Brick-sparse grids for host-side consumers.

A ligand grid is >= 97 % zeros (994 non-zero voxels of 1,048,576 for the 10gs ligand), so a dense
device->host copy of finished grids measures PCIe, not the voxelizer.  `Voxelizer.compact()` (csrc:
mvx_compact_bricks_kernel) keeps only the 8 x 8 x 8-voxel bricks that hold a non-zero value: `ids[n]` says where brick n
belongs, `vals[n]` is its 512 values.  `to_dense()` rebuilds the (B, C, D, D, D) grid exactly (tests compare it bit for
bit with the dense output); host consumers that can work on bricks never materialise it.
"""
from __future__ import annotations

import numpy as np
import torch

BRICK = 8


class SparseGrids:
    def __init__(self, ids, vals, shape):
        """ids (n,) int64-compatible brick ids, vals (n, 512) float32 — torch tensors or numpy arrays; shape (B, C, D)."""
        self.ids, self.vals = ids, vals
        self.num_mols, self.channels, self.dimension = (int(v) for v in shape)

    @property
    def num_bricks(self) -> int:
        return int(self.ids.shape[0])

    @property
    def nbytes(self) -> int:
        return self.num_bricks * (512 * 4 + 4)

    def cpu(self, pin: bool = False):
        """Host copy (numpy arrays)."""
        if isinstance(self.ids, np.ndarray):
            return self
        ids, vals = self.ids.cpu(), self.vals.cpu()
        if pin:
            ids, vals = ids.pin_memory(), vals.pin_memory()
        return SparseGrids(ids.numpy(), vals.numpy(), (self.num_mols, self.channels, self.dimension))

    def split_ids(self):
        """(molecule, channel, bx, by, bz) of every brick."""
        lib = torch if isinstance(self.ids, torch.Tensor) else np
        ids = self.ids.to(torch.int64) & 0xFFFFFFFF if lib is torch else self.ids.astype(np.int64) & 0xFFFFFFFF
        nb = -(-self.dimension // BRICK)
        bz = ids % nb
        col = (ids // nb) % (nb * nb)
        mc = ids // (nb * nb * nb)
        return mc // self.channels, mc % self.channels, col // nb, col % nb, bz

    def to_dense(self):
        """The (B, C, D, D, D) float32 grid these bricks came from (same container type as `vals`)."""
        B, C, D = self.num_mols, self.channels, self.dimension
        nb = -(-D // BRICK)
        mol, ch, bx, by, bz = self.split_ids()
        if isinstance(self.vals, torch.Tensor):
            padded = torch.zeros((B * C, nb, nb, nb, BRICK, BRICK, BRICK), dtype=torch.float32, device=self.vals.device)
            padded[mol * C + ch, bx, by, bz] = self.vals.reshape(-1, BRICK, BRICK, BRICK)
            full = padded.permute(0, 1, 4, 2, 5, 3, 6).reshape(B, C, nb * BRICK, nb * BRICK, nb * BRICK)
            return full[:, :, :D, :D, :D].contiguous()
        padded = np.zeros((B * C, nb, nb, nb, BRICK, BRICK, BRICK), dtype=np.float32)
        padded[mol * C + ch, bx, by, bz] = self.vals.reshape(-1, BRICK, BRICK, BRICK)
        full = padded.transpose(0, 1, 4, 2, 5, 3, 6).reshape(B, C, nb * BRICK, nb * BRICK, nb * BRICK)
        return np.ascontiguousarray(full[:, :, :D, :D, :D])
