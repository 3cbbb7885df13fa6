"""precision=64 golden fixtures (tests/golden/p64/*.npz) from the LIVE reference:

    python tests/golden/make_golden_p64.py      (build container only: needs /root/reference)

Same layout as make_golden.py, but the voxelizer is created with precision=64
(reference molvoxel/voxelizer/numpy/voxelizer.py:28-34) and the stored values are float64.
Array radii are float32 in the fixtures (what the C ABI carries); the reference widens them (:130, :271).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, molvoxel, parse_sdf_heavy, sparse, synth  # noqa: E402

OUT = os.path.join(HERE, "p64")


def save_case(name, cfg, coords, center, channels, radii):
    kw = {"precision": 64}
    if cfg.get("blockdim") is not None:
        kw["blockdim"] = cfg["blockdim"]
    vox = molvoxel.create_voxelizer(cfg["resolution"], cfg["dimension"], cfg["radii_type"], cfg["density_type"],
                                    library="numpy", **kw)
    if cfg["mode"] == "single":
        out = vox.forward_single(coords, center, radii)
    elif cfg["mode"] == "types":
        out = vox.forward_types(coords, center, channels, radii)
    else:
        out = vox.forward_features(coords, center, channels, radii)
    assert out.dtype == np.float64
    idx, val = sparse(out)
    payload = {"cfg": json.dumps(cfg), "coords": coords, "shape": np.array(out.shape, dtype=np.int64),
               "radii": np.asarray(radii), "nz_idx": idx, "nz_val": val}
    if center is not None:
        payload["center"] = center
    if channels is not None:
        payload["channels"] = channels
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
    print(f"{name}: shape {out.shape} sum {float(out.sum()):.12f} max {float(out.max()):.12f} nnz {int((out != 0).sum())}")


def main():
    rng = np.random.default_rng(64)
    os.makedirs(OUT, exist_ok=True)
    lig_xyz, lig_elem = parse_sdf_heavy(os.path.join(REF, "test", "10gs", "10gs_ligand.sdf"))
    tmap = {"C": 0, "N": 1, "O": 2, "S": 3}
    lig_types = np.array([tmap[e] for e in lig_elem], dtype=np.int16)
    base = dict(resolution=0.5, dimension=32, density_type="gaussian", blockdim=None)
    save_case("lig10gs_types_gaussian_d32", dict(mode="types", radii_type="scalar", **base), lig_xyz, lig_xyz.mean(0), lig_types, 1.0)
    V = 250
    for bd in (None, 20):
        for dens in ("gaussian", "binary"):
            b = dict(resolution=0.5, dimension=20, density_type=dens, blockdim=bd)
            tag = f"d20_bd{bd or 8}_{dens}"
            coords = synth(rng, V, 20, 0.5)
            center = rng.uniform(-0.3, 0.3, size=3)
            types = rng.integers(0, 5, size=V).astype(np.int16)
            feats = rng.uniform(0, 1, size=(V, 6)).astype(np.float32)
            save_case(f"{tag}_types_scalar", dict(mode="types", radii_type="scalar", **b), coords, center, types, 1.3)
            save_case(f"{tag}_types_channel", dict(mode="types", radii_type="channel-wise", **b), coords, center, types,
                      rng.uniform(0.8, 1.8, size=5).astype(np.float32))
            save_case(f"{tag}_feat_atom", dict(mode="features", radii_type="atom-wise", **b), coords, center, feats,
                      rng.uniform(0.8, 1.8, size=V).astype(np.float32))
            save_case(f"{tag}_feat_channel", dict(mode="features", radii_type="channel-wise", **b), coords, center, feats,
                      rng.uniform(0.8, 1.8, size=6).astype(np.float32))
            save_case(f"{tag}_single_scalar", dict(mode="single", radii_type="scalar", **b), coords, None, None, 1.5)


if __name__ == "__main__":
    main()
