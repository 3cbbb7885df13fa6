"""Multi-GPU sharding of a molecule batch: one process per GPU, one contiguous slice per rank.

Molecules are independent (no cross-molecule term in any reference forward_*), so there is no
collective on the data path.  `gather_grids` is the optional NCCL all-gather of
finished grids; `torch.distributed` is plumbing only.
"""
from __future__ import annotations

import numpy as np
import torch


def shard_bounds(num_mols: int, rank: int, world_size: int) -> tuple[int, int]:
    """Rank k of R owns molecules [k*ceil(N/R), min(N, (k+1)*ceil(N/R)))."""
    per = -(-num_mols // world_size) if world_size > 0 else num_mols
    lo = min(num_mols, rank * per)
    return lo, min(num_mols, lo + per)


def shard_batch(mol_offsets, rank: int, world_size: int, *per_atom, per_mol=()):
    """Slice a CSR batch for this rank.  Returns (local_offsets, [per-atom slices], [per-mol slices])."""
    offs = np.asarray(mol_offsets.cpu() if isinstance(mol_offsets, torch.Tensor) else mol_offsets, dtype=np.int64)
    lo, hi = shard_bounds(len(offs) - 1, rank, world_size)
    a0, a1 = int(offs[lo]), int(offs[hi])
    local = (offs[lo:hi + 1] - a0).astype(np.int32)
    return local, [None if a is None else a[a0:a1] for a in per_atom], [None if a is None else a[lo:hi] for a in per_mol]


def gather_grids(local_grids: torch.Tensor, num_mols: int, group=None) -> torch.Tensor:
    """Optional: all-gather the (B_k, C, D, H, W) slices into (num_mols, C, D, H, W) on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = -(-num_mols // world)
    padded = local_grids
    if local_grids.shape[0] < per:   # last rank may be short: pad so every rank contributes equally
        pad = torch.zeros((per - local_grids.shape[0],) + tuple(local_grids.shape[1:]), dtype=local_grids.dtype,
                          device=local_grids.device)
        padded = torch.cat([local_grids, pad], 0)
    out = torch.empty((world * per,) + tuple(local_grids.shape[1:]), dtype=local_grids.dtype, device=local_grids.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return out[:num_mols]
