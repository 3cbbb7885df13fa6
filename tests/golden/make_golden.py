"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference, numpy, scipy):

    python tests/golden/make_golden.py

Each fixture is an .npz holding the exact inputs of one reference call and the output of
``molvoxel.create_voxelizer(..., library="numpy")`` (precision=32) for it.  The reference's own
tests pin no numeric values for this path, so these vectors — produced by the
unmodified reference with the library versions recorded in ``versions.json`` — are what pins
the oracle (oracle/mvx_oracle.c) and, through it, the CUDA path.

Outputs are stored sparsely (flat indices + values of the non-zero voxels) to stay small.
The 10gs ligand is hand-parsed from the reference's SDF fixture (rdkit is absent).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = os.environ.get("MOLVOXEL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import molvoxel  # noqa: E402
import scipy  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def parse_sdf_heavy(path):
    """V2000 SDF -> heavy-atom coords (V,3) f64 and element symbols (Chem.SDMolSupplier default removeHs)."""
    lines = open(path).read().splitlines()
    natoms = int(lines[3][:3])
    xyz, elem = [], []
    for ln in lines[4:4 + natoms]:
        sym = ln[31:34].strip()
        if sym == "H":
            continue
        xyz.append((float(ln[0:10]), float(ln[10:20]), float(ln[20:30])))
        elem.append(sym)
    return np.array(xyz, dtype=np.float64), elem


def sparse(out):
    flat = out.reshape(-1)
    idx = np.flatnonzero(flat).astype(np.int64)
    return idx, flat[idx].copy()


def run_ref(cfg, coords, center, channels, radii):
    kw = {}
    if cfg.get("blockdim") is not None:
        kw["blockdim"] = cfg["blockdim"]
    if "sigma" in cfg:
        kw["sigma"] = cfg["sigma"]
    vox = molvoxel.create_voxelizer(cfg["resolution"], cfg["dimension"], cfg["radii_type"], cfg["density_type"],
                                    library="numpy", **kw)
    if cfg["mode"] == "single":
        return vox.forward_single(coords, center, radii)
    if cfg["mode"] == "types":
        return vox.forward_types(coords, center, channels, radii)
    return vox.forward_features(coords, center, channels, radii)


def save_case(name, cfg, coords, center, channels, radii, sample_stride=None):
    out = run_ref(cfg, coords, center, channels, radii)
    assert out.dtype == np.float32
    payload = {"cfg": json.dumps(cfg), "coords": coords, "shape": np.array(out.shape, dtype=np.int64)}
    if center is not None:
        payload["center"] = center
    if channels is not None:
        payload["channels"] = channels
    payload["radii"] = np.asarray(radii)
    if sample_stride is None:
        idx, val = sparse(out)
        payload["nz_idx"], payload["nz_val"] = idx, val
    else:  # big dense outputs: strided sample + per-channel checksums
        flat = out.reshape(-1)
        sidx = np.arange(0, flat.shape[0], sample_stride, dtype=np.int64)
        payload["sample_idx"], payload["sample_val"] = sidx, flat[sidx].copy()
        payload["chan_sum"] = out.reshape(out.shape[0], -1).astype(np.float64).sum(1)
        payload["chan_nnz"] = (out.reshape(out.shape[0], -1) != 0).sum(1).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print(f"{name}: shape {out.shape} sum {float(out.sum()):.6f} max {float(out.max()):.6f} nnz {int((out != 0).sum())}")


def synth(rng, V, dim, res, spread=2.0, f32=False):
    half = res * (dim - 1) / 2.0
    c = rng.uniform(-half - spread, half + spread, size=(V, 3))
    if f32:
        c = c.astype(np.float32)
    return c


def main():
    rng = np.random.default_rng(20261018)
    lig_xyz, lig_elem = parse_sdf_heavy(os.path.join(REF, "test", "10gs", "10gs_ligand.sdf"))
    tmap = {"C": 0, "N": 1, "O": 2, "S": 3}
    lig_types = np.array([tmap[e] for e in lig_elem], dtype=np.int16)
    lig_center = lig_xyz.mean(0)

    # cfg 1 of BASELINE.json: 10gs ligand, types C/N/O/S, 64^3, res 0.5, r 1.0 (gaussian) + binary twin
    for dens in ("gaussian", "binary"):
        save_case(f"lig10gs_types_{dens}",
                  dict(mode="types", resolution=0.5, dimension=64, radii_type="scalar", density_type=dens, blockdim=None),
                  lig_xyz, lig_center, lig_types, 1.0)
    save_case("lig10gs_single_gaussian",
              dict(mode="single", resolution=0.5, dimension=64, radii_type="scalar", density_type="gaussian", blockdim=None),
              lig_xyz, lig_center, None, 1.0)
    # reference test_run_numpy.py:31-32 shapes: small single-block grid and high-resolution grid
    save_case("lig10gs_types_small16_bd16",
              dict(mode="types", resolution=0.5, dimension=16, radii_type="scalar", density_type="gaussian", blockdim=16),
              lig_xyz, lig_center, lig_types, 1.0)
    save_case("lig10gs_types_res04",
              dict(mode="types", resolution=0.4, dimension=64, radii_type="scalar", density_type="gaussian", blockdim=None),
              lig_xyz, lig_center, lig_types, 1.0)
    # channel-wise / atom-wise radii on the ligand (test_run_numpy.py:63-75)
    save_case("lig10gs_types_channelwise",
              dict(mode="types", resolution=0.5, dimension=64, radii_type="channel-wise", density_type="gaussian", blockdim=None),
              lig_xyz, lig_center, lig_types, np.array([1.0, 1.2, 1.4, 1.8], dtype=np.float32))
    save_case("lig10gs_types_atomwise",
              dict(mode="types", resolution=0.5, dimension=64, radii_type="atom-wise", density_type="binary", blockdim=None),
              lig_xyz, lig_center, lig_types, rng.uniform(0.8, 2.0, size=lig_xyz.shape[0]).astype(np.float32))

    # dim=20 (blocks 8+8+4): every radii type x mode x density, compat (bd=8) and exact (bd=20)
    V = 300
    for bd in (None, 20):
        for dens in ("gaussian", "binary"):
            coords = synth(rng, V, 20, 0.5)
            center = rng.uniform(-0.3, 0.3, size=3)
            types = rng.integers(0, 5, size=V).astype(np.int16)
            feats = rng.uniform(0, 1, size=(V, 6)).astype(np.float32)
            feats[rng.uniform(size=feats.shape) < 0.3] = 0.0
            tag = f"d20_bd{bd or 8}_{dens}"
            base = dict(resolution=0.5, dimension=20, density_type=dens, blockdim=bd)
            save_case(f"{tag}_types_scalar", dict(mode="types", radii_type="scalar", **base), coords, center, types, 1.3)
            save_case(f"{tag}_types_channel", dict(mode="types", radii_type="channel-wise", **base), coords, center, types,
                      rng.uniform(0.8, 1.9, size=5).astype(np.float32))
            save_case(f"{tag}_types_atom", dict(mode="types", radii_type="atom-wise", **base), coords, center, types,
                      rng.uniform(0.8, 1.9, size=V).astype(np.float32))
            save_case(f"{tag}_feat_scalar", dict(mode="features", radii_type="scalar", **base), coords, center, feats, 1.0)
            save_case(f"{tag}_feat_channel", dict(mode="features", radii_type="channel-wise", **base), coords, center, feats,
                      rng.uniform(0.8, 1.9, size=6).astype(np.float32))
            save_case(f"{tag}_feat_atom", dict(mode="features", radii_type="atom-wise", **base), coords, center, feats,
                      rng.uniform(0.8, 1.9, size=V).astype(np.float32))
            save_case(f"{tag}_single_scalar", dict(mode="single", radii_type="scalar", **base), coords, None, None, 1.5)
            save_case(f"{tag}_single_atom", dict(mode="single", radii_type="atom-wise", **base), coords, center, None,
                      rng.uniform(0.8, 1.9, size=V).astype(np.float32))

    # fp32 coords + fp32 centre (fp32 centring), sigma != default, non-cubic-friendly dim 21 (scalar stores)
    coords = synth(rng, 200, 21, 0.375, f32=True)
    save_case("d21_f32coords_types",
              dict(mode="types", resolution=0.375, dimension=21, radii_type="scalar", density_type="gaussian", blockdim=None, sigma=0.8),
              coords, rng.uniform(-0.2, 0.2, size=3).astype(np.float32), rng.integers(0, 3, size=200).astype(np.int16), 1.1)

    # adversarial ties: atoms exactly on grid points / on fp32-rounded cutoff, clip edge, cull window
    dim, res = 24, 0.5
    half = res * (dim - 1) / 2.0
    axis = np.arange(dim) * res - half
    pts = []
    for _ in range(60):  # exactly on grid points: d == r ties for r = 1.0, 1.5
        pts.append(axis[rng.integers(0, dim, size=3)])
    for _ in range(60):  # on voxel-plane midpoints
        pts.append(axis[rng.integers(0, dim, size=3)] + np.array([0.25, 0.0, 0.0]))
    for b in (8, 16):    # cull window: first plane of a non-first block minus [r - res/2, r]
        for t in np.linspace(0.75, 1.0, 9):
            p = axis[rng.integers(0, dim, size=3)].copy()
            p[rng.integers(0, 3)] = axis[b] - t
            pts.append(p)
    for sgn in (-1, 1):  # clip edge: exactly r outside, and one ulp inside
        for k in range(3):
            p = axis[rng.integers(0, dim, size=3)].copy()
            p[k] = sgn * (half + 1.0)
            pts.append(p)
            q = p.copy()
            q[k] = np.nextafter(p[k], 0.0)
            pts.append(q)
    pts = np.array(pts, dtype=np.float64)
    types = rng.integers(0, 4, size=pts.shape[0]).astype(np.int16)
    for dens in ("gaussian", "binary"):
        save_case(f"ties_d24_{dens}",
                  dict(mode="types", resolution=0.5, dimension=24, radii_type="scalar", density_type=dens, blockdim=None),
                  pts, None, types, 1.0)
    save_case("ties_d24_binary_r15_exact",
              dict(mode="types", resolution=0.5, dimension=24, radii_type="scalar", density_type="binary", blockdim=24),
              pts, None, types, 1.5)

    # cfg 2 shape: 2,000 atoms, features C=16, 48^3 (sampled + checksummed: the dense output is 7 MB)
    V = 2000
    half = 0.5 * 47 / 2.0
    coords = rng.uniform(-half, half, size=(V, 3)).astype(np.float32).astype(np.float64)
    feats = np.zeros((V, 16), dtype=np.float32)
    feats[np.arange(V), rng.integers(0, 8, size=V)] = 1.0
    feats[:, 8:] = (rng.uniform(size=(V, 8)) < 0.25).astype(np.float32)
    save_case("cfg2_feat16_d48",
              dict(mode="features", resolution=0.5, dimension=48, radii_type="scalar", density_type="gaussian", blockdim=None),
              coords, np.zeros(3), feats, 1.0, sample_stride=61)
    save_case("cfg2_types_binary_d48",
              dict(mode="types", resolution=0.5, dimension=48, radii_type="scalar", density_type="binary", blockdim=None),
              coords, np.zeros(3), rng.integers(0, 4, size=V).astype(np.int16), 1.0)

    # other blockdim values: the cull emulation must hold for any block size (non-divisors, 1-voxel blocks)
    rng2 = np.random.default_rng(77)
    for bd in (5, 16, 3, 1):
        coords = synth(rng2, 250, 24, 0.5, spread=1.5)
        types = rng2.integers(0, 4, size=250).astype(np.int16)
        save_case(f"d24_bd{bd}_types_binary",
                  dict(mode="types", resolution=0.5, dimension=24, radii_type="scalar", density_type="binary", blockdim=bd),
                  coords, None, types, 1.2)
        save_case(f"d24_bd{bd}_feat_atom_gaussian",
                  dict(mode="features", resolution=0.5, dimension=24, radii_type="atom-wise", density_type="gaussian", blockdim=bd),
                  coords, rng2.uniform(-0.2, 0.2, size=3), rng2.uniform(0, 1, size=(250, 5)).astype(np.float32),
                  rng2.uniform(0.7, 2.2, size=250).astype(np.float32))
    # resolution that is not exactly representable in fp32, 96-wide grid split in two z chunks, large radii
    coords = synth(rng2, 600, 96, 0.3, spread=1.0)
    save_case("d96_res03_types_atom_binary",
              dict(mode="types", resolution=0.3, dimension=96, radii_type="atom-wise", density_type="binary", blockdim=None),
              coords, None, rng2.integers(0, 3, size=600).astype(np.int16), rng2.uniform(0.6, 2.5, size=600).astype(np.float32))

    json.dump({"molvoxel": molvoxel.__version__, "numpy": np.__version__, "scipy": scipy.__version__,
               "python": sys.version.split()[0], "backend": "numpy precision=32"},
              open(os.path.join(HERE, "versions.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
