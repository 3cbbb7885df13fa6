"""Vectorised point-cloud makers and CSR collation for the batched driver.

The reference builds a molecule's point cloud atom by atom through rdkit objects
(molvoxel/etc/rdkit/pointcloud.py): atom positions plus, optionally, one pseudo-atom at every bond
midpoint (:79-89), a type index or a feature row per point with the bond channels placed after the atom
channels (:163-182, :103-135), and, for a system of several molecules (ligand + protein), a channel offset
per molecule (:191-211, :236-248, :303-312).  rdkit is a host-side chemistry toolkit and stays outside this
package; the makers here take plain arrays (positions, per-atom type index or feature rows, bond index
pairs, per-bond type index or feature rows) and produce the same arrays with numpy only, for whole
molecules at once.  `from_rdkit` adapts an rdkit Mol when rdkit is installed.

`collate` packs many point clouds into the CSR batch `Voxelizer.forward_*_batch` takes (one concatenated
coordinate / channel array, molecule offsets, one centre per molecule), optionally into reusable pinned
host buffers so that the H2D copies of `non_blocking=True` calls are asynchronous.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class PointCloud:
    coords: np.ndarray          # (P, 3) float64: atoms, then bond midpoints
    channels: np.ndarray        # (P,) int16 types, or (P, C) float32 features
    num_channels: int


def bond_midpoints(atom_coords, bonds):
    """(atom_coords[begin] + atom_coords[end]) / 2 — reference pointcloud.py:85, same operation order."""
    atom_coords = np.asarray(atom_coords, dtype=np.float64)
    bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
    return (atom_coords[bonds[:, 0]] + atom_coords[bonds[:, 1]]) / 2


def mol_point_cloud(atom_coords, atom_channels, num_atom_channels, bonds=None, bond_channels=None,
                    num_bond_channels=0, channel_type="features", channel_offset=0, total_channels=None):
    """One molecule -> PointCloud (reference MolPointCloudMaker.run, pointcloud.py:72-182).

    atom_channels / bond_channels: integer type indices (N,) / (M,) for channel_type "types" (or for
    "features", where they are one-hot encoded like TypeGetter.get_feature, base.py:40-50), or float
    feature rows (N, Ca) / (M, Cb) for "features".  Bond points use channels
    [channel_offset + num_atom_channels, ... + num_bond_channels): the reference's bond_start_index
    (:150-153) and bond_st (:98-101)."""
    assert channel_type in ("features", "types"), f"channel_type(input: {channel_type}) should be 'features' or 'types'"
    atom_coords = np.asarray(atom_coords, dtype=np.float64).reshape(-1, 3)
    n_atoms = atom_coords.shape[0]
    use_bond = bonds is not None
    if use_bond:
        bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        coords = np.concatenate([atom_coords, bond_midpoints(atom_coords, bonds)], axis=0)
        n_bonds = bonds.shape[0]
    else:
        coords, n_bonds = atom_coords, 0
    width = num_atom_channels + (num_bond_channels if use_bond else 0)
    total = channel_offset + width if total_channels is None else int(total_channels)
    atom_channels = np.asarray(atom_channels)
    if channel_type == "types":
        assert atom_channels.ndim == 1, "types point clouds need integer type indices"
        out = np.empty(n_atoms + n_bonds, dtype=np.int16)
        out[:n_atoms] = atom_channels + channel_offset
        if use_bond:
            out[n_atoms:] = np.asarray(bond_channels) + (channel_offset + num_atom_channels)
        return PointCloud(coords, out, total)
    out = np.zeros((n_atoms + n_bonds, total), dtype=np.float32)
    _place(out[:n_atoms], atom_channels, channel_offset, num_atom_channels)
    if use_bond:
        _place(out[n_atoms:], np.asarray(bond_channels), channel_offset + num_atom_channels, num_bond_channels)
    return PointCloud(coords, out, total)


def _place(rows, channels, start, width):
    if rows.shape[0] == 0:
        return
    if channels.ndim == 1:      # type index -> one-hot row (base.py:40-50)
        rows[np.arange(rows.shape[0]), start + channels.astype(np.int64)] = 1.0
    else:
        assert channels.shape[1] == width, f"feature rows have {channels.shape[1]} channels, expected {width}"
        rows[:, start:start + width] = channels


def system_point_cloud(molecules, channel_type="features"):
    """Several molecules of one system (e.g. ligand + protein) -> one PointCloud whose channel blocks follow
    each other, reference MolSystemPointCloudMaker (pointcloud.py:185-312).  `molecules` is a list of dicts
    with the keyword arguments of mol_point_cloud (without the offsets)."""
    widths = [m["num_atom_channels"] + (m.get("num_bond_channels", 0) if m.get("bonds") is not None else 0)
              for m in molecules]
    total = int(sum(widths))
    clouds, off = [], 0
    for m, w in zip(molecules, widths):
        clouds.append(mol_point_cloud(channel_type=channel_type, channel_offset=off, total_channels=total, **m))
        off += w
    coords = np.concatenate([c.coords for c in clouds], axis=0) if clouds else np.zeros((0, 3))
    channels = np.concatenate([c.channels for c in clouds], axis=0)
    return PointCloud(coords, channels, total)


def from_rdkit(rdmol, symbols, bondtypes=None, unknown=False):
    """rdkit Mol -> keyword arguments of mol_point_cloud (AtomTypeGetter / BondTypeGetter semantics,
    getter.py:16-44).  Needs rdkit, which this package does not depend on."""
    sym = {s: i for i, s in enumerate(symbols)}
    n_sym = len(symbols) + (1 if unknown else 0)
    get = (lambda a: sym.get(a.GetSymbol(), n_sym - 1)) if unknown else (lambda a: sym[a.GetSymbol()])
    out = dict(atom_coords=rdmol.GetConformer().GetPositions(),
               atom_channels=np.array([get(a) for a in rdmol.GetAtoms()], dtype=np.int16), num_atom_channels=n_sym)
    if bondtypes is not None:
        bt = {b: i for i, b in enumerate(bondtypes)}
        out.update(bonds=np.array([(b.GetBeginAtomIdx(), b.GetEndAtomIdx()) for b in rdmol.GetBonds()], dtype=np.int64).reshape(-1, 2),
                   bond_channels=np.array([bt[b.GetBondType()] for b in rdmol.GetBonds()], dtype=np.int16),
                   num_bond_channels=len(bondtypes))
    return out


class Collator:
    """Packs point clouds into the CSR batch of Voxelizer.forward_*_batch.  With pinned=True the arrays are
    views of reusable page-locked buffers (two sets, used alternately, so a batch can be filled while the
    previous one is still being copied).  A set is handed out again only after the copy that last read it has
    finished: pass the event of that copy (`Voxelizer.last_copy_event` after a non_blocking forward) to `in_flight()`.

    compact=True sends fewer bytes over PCIe where that is LOSSLESS: feature rows whose values are all exactly
    representable as uint8 (one-hot / flag / count features) are packed as uint8 (widened exactly on the device,
    mvx_batch.features_dtype), and coordinates that are exactly representable in float32 are packed as float32 when
    centres are given as float64 — numpy's promotion then still centres in fp64 (reference numpy/voxelizer.py:263),
    so the grids are bit-identical.  The checks run here, in the loader, where the data is touched anyway."""

    def __init__(self, pinned: bool = False, compact: bool = False):
        self.pinned = pinned
        self.compact = compact
        self._bufs = [{}, {}]
        self._events = [None, None]
        self._turn = 0

    def in_flight(self, event):
        """The set returned by the last call is being read by an asynchronous copy that `event` (a torch.cuda.Event
        recorded after it) completes; the set is not refilled before that."""
        self._events[self._turn ^ 1] = event

    def _array(self, name, shape, dtype):
        n = int(np.prod(shape))
        if not self.pinned:
            return np.empty(shape, dtype=dtype)
        import torch
        slot = self._bufs[self._turn]
        buf = slot.get((name, np.dtype(dtype).str))
        tdt = torch.from_numpy(np.empty(0, dtype=dtype)).dtype
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1) * 5 // 4 + 16, dtype=tdt).pin_memory()
            slot[(name, np.dtype(dtype).str)] = buf
        return buf[:n].view(*shape).numpy()

    def __call__(self, clouds, centers=None, radii=None):
        """clouds: PointClouds of equal channel type / count.  centers: None (no centring), "mean" (centroid of
        every cloud, fp64) or (B, 3).  radii: None, or per-cloud (P_i,) arrays (atom-wise radii).
        Returns dict(coords, mol_offsets, centers, channels, radii, num_channels)."""
        B = len(clouds)
        if self._events[self._turn] is not None:   # the copy that last read this set of pinned buffers has finished
            self._events[self._turn].synchronize()
            self._events[self._turn] = None
        counts = np.array([c.coords.shape[0] for c in clouds], dtype=np.int64)
        offs = self._array("offs", (B + 1,), np.int32)
        offs[0] = 0
        np.cumsum(counts, out=offs[1:])
        N = int(offs[-1])
        is_types = B == 0 or clouds[0].channels.ndim == 1
        C = clouds[0].num_channels if B else 0
        cdt, fdt = np.float64, np.float32
        if self.compact and B:
            if centers is not None and all(np.array_equal(c.coords, np.asarray(c.coords, dtype=np.float32)) for c in clouds):
                cdt = np.float32
            if not is_types and all(np.array_equal(c.channels, np.asarray(c.channels).astype(np.uint8)) for c in clouds):
                fdt = np.uint8
        coords = self._array("coords", (N, 3), cdt)
        channels = self._array("chan", (N,) if is_types else (N, C), np.int32 if is_types else fdt)
        for c, a, b in zip(clouds, offs[:-1], offs[1:]):
            assert c.num_channels == C and (c.channels.ndim == 1) == is_types, "point clouds of one batch must agree in channels"
            coords[a:b] = c.coords
            channels[a:b] = c.channels
        out = dict(coords=coords, mol_offsets=offs, channels=channels, num_channels=C, centers=None, radii=None)
        if isinstance(centers, str):
            assert centers == "mean"
            cen = self._array("centers", (B, 3), np.float64)
            for m, (c, a, b) in enumerate(zip(clouds, offs[:-1], offs[1:])):
                cen[m] = np.asarray(c.coords, dtype=np.float64).mean(axis=0) if b > a else 0.0
            out["centers"] = cen
        elif centers is not None:
            cen = self._array("centers", (B, 3), np.float64)
            cen[:] = np.asarray(centers, dtype=np.float64).reshape(B, 3)
            out["centers"] = cen
        if radii is not None:
            r = self._array("radii", (N,), np.float32)
            for rr, a, b in zip(radii, offs[:-1], offs[1:]):
                r[a:b] = rr
            out["radii"] = r
        self._turn ^= 1
        return out


def collate(clouds, centers=None, radii=None):
    """One-shot Collator()(...) into fresh (pageable) arrays."""
    return Collator(False)(clouds, centers, radii)
