"""CPU tests of the host-side rows f2 / f4: vectorised point clouds + CSR collation against a loop restatement
of the reference makers (molvoxel/etc/rdkit/pointcloud.py, rdkit-free), and the .dx writer against a file written
by the reference's own writer (tests/golden/make_dx_golden.py)."""
import os

import numpy as np
import pytest

from molvoxel_b200.dx import write_grid_to_dx_file
from molvoxel_b200.pointcloud import Collator, collate, mol_point_cloud, system_point_cloud

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _loop_maker(mol, channel_type, atom_offset, total):
    """Atom-by-atom restatement of _MolElementPointCloudMaker (pointcloud.py:72-182): coordinates (:79-89),
    features (:118-135) with blocks [atom_st, atom_end) / [bond_st, bond_end) (:98-101), types with
    atom_start_index / bond_start_index (:150-153, :178-182)."""
    xyz = np.asarray(mol["atom_coords"], dtype=np.float64)
    bonds = mol.get("bonds")
    coords = [xyz[i] for i in range(len(xyz))]
    if bonds is not None:
        for b, e in bonds:
            coords.append((xyz[b] + xyz[e]) / 2)
    na, nbc = mol["num_atom_channels"], mol.get("num_bond_channels", 0)
    if channel_type == "types":
        out = [int(t) + atom_offset for t in mol["atom_channels"]]
        if bonds is not None:
            out += [int(t) + atom_offset + na for t in mol["bond_channels"]]
        return np.array(coords).reshape(-1, 3), np.array(out, dtype=np.int16)
    rows = []
    for t in mol["atom_channels"]:
        r = np.zeros(total, dtype=np.float32)
        if np.ndim(t) == 0:
            r[atom_offset + int(t)] = 1
        else:
            r[atom_offset:atom_offset + na] = t
        rows.append(r)
    if bonds is not None:
        for t in mol["bond_channels"]:
            r = np.zeros(total, dtype=np.float32)
            if np.ndim(t) == 0:
                r[atom_offset + na + int(t)] = 1
            else:
                r[atom_offset + na:atom_offset + na + nbc] = t
            rows.append(r)
    return np.array(coords).reshape(-1, 3), np.array(rows, dtype=np.float32).reshape(-1, total)


def _random_mol(rng, n_atoms, n_bonds, na, nb, feature_rows=False):
    m = dict(atom_coords=rng.normal(scale=4.0, size=(n_atoms, 3)), num_atom_channels=na)
    m["atom_channels"] = rng.uniform(size=(n_atoms, na)).astype(np.float32) if feature_rows else rng.integers(0, na, size=n_atoms)
    if nb:
        m["bonds"] = rng.integers(0, max(n_atoms, 1), size=(n_bonds, 2))
        m["bond_channels"] = rng.uniform(size=(n_bonds, nb)).astype(np.float32) if feature_rows else rng.integers(0, nb, size=n_bonds)
        m["num_bond_channels"] = nb
    return m


@pytest.mark.parametrize("channel_type", ["types", "features"])
@pytest.mark.parametrize("with_bonds", [False, True])
def test_mol_point_cloud_matches_loop_restatement(channel_type, with_bonds):
    rng = np.random.default_rng(3)
    m = _random_mol(rng, 23, 25, 5, 4 if with_bonds else 0)
    pc = mol_point_cloud(channel_type=channel_type, **m)
    total = 5 + (4 if with_bonds else 0)
    coords, chan = _loop_maker(m, channel_type, 0, total)
    assert pc.num_channels == total
    assert np.array_equal(pc.coords, coords) and pc.coords.dtype == np.float64
    assert np.array_equal(pc.channels, chan) and pc.channels.dtype == chan.dtype


@pytest.mark.parametrize("channel_type", ["types", "features"])
def test_system_point_cloud_offsets_channels_per_molecule(channel_type):
    """ligand (atoms + bonds) then protein (atoms only): ComplexPointCloudMaker, pointcloud.py:315-326."""
    rng = np.random.default_rng(4)
    lig, prot = _random_mol(rng, 17, 18, 4, 3), _random_mol(rng, 60, 0, 6, 0)
    pc = system_point_cloud([lig, prot], channel_type)
    total = 4 + 3 + 6
    c0, f0 = _loop_maker(lig, channel_type, 0, total)
    c1, f1 = _loop_maker(prot, channel_type, 7, total)
    assert pc.num_channels == total
    assert np.array_equal(pc.coords, np.concatenate([c0, c1])) and np.array_equal(pc.channels, np.concatenate([f0, f1]))


def test_feature_rows_and_empty_molecule():
    rng = np.random.default_rng(5)
    m = _random_mol(rng, 9, 7, 3, 2, feature_rows=True)
    pc = mol_point_cloud(channel_type="features", **m)
    _, ref = _loop_maker(m, "features", 0, 5)
    assert np.array_equal(pc.channels, ref)
    e = mol_point_cloud(np.zeros((0, 3)), np.zeros(0, dtype=np.int64), 4, channel_type="types")
    assert e.coords.shape == (0, 3) and e.channels.shape == (0,)
    with pytest.raises(AssertionError):
        mol_point_cloud(np.zeros((2, 3)), np.zeros(2), 4, channel_type="bogus")


def test_collate_builds_the_csr_batch():
    rng = np.random.default_rng(6)
    clouds = [mol_point_cloud(channel_type="types", **_random_mol(rng, n, n, 4, 2)) for n in (5, 0, 12, 1)]
    radii = [rng.uniform(1, 2, size=c.coords.shape[0]).astype(np.float32) for c in clouds]
    b = collate(clouds, centers="mean", radii=radii)
    assert b["mol_offsets"].dtype == np.int32 and b["mol_offsets"].tolist() == [0, 10, 10, 34, 36]
    assert b["num_channels"] == 6 and b["channels"].dtype == np.int32 and b["coords"].dtype == np.float64
    for m, c in enumerate(clouds):
        a, e = b["mol_offsets"][m], b["mol_offsets"][m + 1]
        assert np.array_equal(b["coords"][a:e], c.coords) and np.array_equal(b["channels"][a:e], c.channels)
        assert np.array_equal(b["radii"][a:e], radii[m])
        if e > a:
            assert np.array_equal(b["centers"][m], c.coords.mean(axis=0))
    feats = [mol_point_cloud(channel_type="features", **_random_mol(rng, n, 0, 3, 0)) for n in (4, 6)]
    fb = Collator(pinned=False)(feats, centers=np.zeros((2, 3)))
    assert fb["channels"].shape == (10, 3) and fb["channels"].dtype == np.float32 and fb["radii"] is None


def test_compact_collation_is_lossless_or_declined():
    """Collator(compact=True): uint8 feature rows only when every value is exactly a uint8, float32 coordinates only when
    they are fp32-representable AND fp64 centres are given (numpy then still centres in fp64); otherwise nothing narrows."""
    rng = np.random.default_rng(8)
    mols = [_random_mol(rng, n, 0, 5, 0) for n in (7, 9)]
    for m in mols:
        m["atom_coords"] = m["atom_coords"].astype(np.float32).astype(np.float64)
    feats = [mol_point_cloud(channel_type="features", **m) for m in mols]        # one-hot rows: exact in uint8
    col = Collator(compact=True)
    b = col(feats, centers="mean")
    assert b["channels"].dtype == np.uint8 and b["coords"].dtype == np.float32 and b["centers"].dtype == np.float64
    plain = Collator()(feats, centers="mean")
    assert np.array_equal(b["channels"].astype(np.float32), plain["channels"]) and np.array_equal(b["coords"].astype(np.float64), plain["coords"])
    assert np.array_equal(b["centers"], plain["centers"])
    assert col(feats, centers=None)["coords"].dtype == np.float64                  # no centre: the dtype is semantics
    mols[0]["atom_coords"] = mols[0]["atom_coords"] + 1e-9                          # not fp32-representable any more
    feats2 = [mol_point_cloud(channel_type="features", **m) for m in mols]
    assert col(feats2, centers="mean")["coords"].dtype == np.float64
    feats2[1].channels[0, 0] = 0.5                                                  # not a uint8 any more
    assert col(feats2, centers="mean")["channels"].dtype == np.float32


def test_dx_writer_is_byte_identical_to_the_reference_writer(tmp_path):
    vals = np.load(os.path.join(GOLDEN, "dx_small_values.npy"))
    p = tmp_path / "out.dx"
    write_grid_to_dx_file(str(p), vals, (1.25, -2.5, 0.125), 0.375)
    assert p.read_text() == open(os.path.join(GOLDEN, "dx_small.dx")).read()
