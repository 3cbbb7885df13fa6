"""Shared test helpers: golden-fixture loading, synthetic inputs, comparison."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.cfg = json.loads(str(z["cfg"]))
        self.coords = z["coords"]
        self.center = z["center"] if "center" in z.files else None
        self.channels = z["channels"] if "channels" in z.files else None
        r = z["radii"]
        self.radii = float(r) if r.ndim == 0 else r
        if r.ndim == 0 and self.cfg.get("radius_f64"):     # an np.float64 scalar radius: fp64 division in the reference
            self.radii = np.float64(r)
        elif r.ndim == 0 and self.cfg.get("radius_f32"):
            self.radii = np.float32(r)
        self.shape = tuple(int(v) for v in z["shape"])
        self.sampled = "sample_idx" in z.files
        if self.sampled:
            self.sample_idx, self.sample_val = z["sample_idx"], z["sample_val"]
            self.chan_sum, self.chan_nnz = z["chan_sum"], z["chan_nnz"]
        else:
            self.nz_idx, self.nz_val = z["nz_idx"], z["nz_val"]

    def dense(self):
        out = np.zeros(int(np.prod(self.shape)), dtype=np.float32)
        out[self.nz_idx] = self.nz_val
        return out.reshape(self.shape)

    def check(self, got, gauss_tol=2e-6):
        """Binary density: bit-exact.  Gaussian: max-abs <= gauss_tol * max(1, peak)."""
        got = np.asarray(got)
        assert got.shape == self.shape, (got.shape, self.shape)
        assert got.dtype == np.float32
        binary_exact = self.cfg["density_type"] == "binary" and self.cfg["mode"] != "features"
        if self.sampled:
            flat = got.reshape(-1)
            ref = self.sample_val
            peak = max(1.0, float(np.abs(ref).max()))
            err = float(np.abs(flat[self.sample_idx] - ref).max())
            assert err <= gauss_tol * peak, f"{self.name}: sampled max-abs {err}"
            nnz = (got.reshape(got.shape[0], -1) != 0).sum(1)
            assert np.array_equal(nnz, self.chan_nnz), f"{self.name}: per-channel nnz differs"
            csum = got.reshape(got.shape[0], -1).astype(np.float64).sum(1)
            assert np.allclose(csum, self.chan_sum, rtol=1e-6, atol=1e-3)
            return err
        ref = self.dense()
        if binary_exact:
            assert np.array_equal(got, ref), f"{self.name}: binary occupancy differs in {(got != ref).sum()} voxels"
            return 0.0
        # occupancy pattern must be identical (the cutoff is a 0.135-high step: SURVEY hazard 3)
        assert np.array_equal(got != 0, ref != 0), f"{self.name}: support differs in {((got != 0) != (ref != 0)).sum()} voxels"
        peak = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(got - ref).max())
        assert err <= gauss_tol * peak, f"{self.name}: max-abs {err} > {gauss_tol * peak}"
        return err


def random_walk_ligand(rng, n_atoms, step=1.5):
    """SURVEY §8d cfg3/cfg4 recipe: 3-D random walk with 1.5 A steps recentred to the origin."""
    d = rng.normal(size=(n_atoms, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = np.cumsum(d * step, axis=0)
    xyz -= xyz.mean(0, keepdims=True)
    return xyz.astype(np.float32).astype(np.float64)


def ligand_batch(rng, B, num_types, vmin=40, vmax=60):
    counts = rng.integers(vmin, vmax + 1, size=B)
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(counts)
    coords = np.concatenate([random_walk_ligand(rng, int(c)) for c in counts], axis=0)
    types = rng.integers(0, num_types, size=int(offs[-1])).astype(np.int32)
    return offs, coords, types


def import_reference():
    """The live, unmodified reference package (offline install under baseline/_ref, which travels to the GPU box;
    /root/reference in the build container).  Returns the module or None."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in (os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(path, "molvoxel")):
            if path not in sys.path:
                sys.path.insert(0, path)
            try:
                return importlib.import_module("molvoxel")
            except Exception:
                continue
    return None


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) in numpy:
    counter (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32.  Host restatement of csrc/mvx_rigid.cuh."""
    c = [np.asarray(counter)[..., i].astype(np.uint64) for i in range(4)]
    k = [np.asarray(key)[..., i].astype(np.uint64) for i in range(2)]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return np.stack(c, -1).astype(np.uint32)


def philox_transform_uniforms(seed, mol_index):
    """The six 53-bit uniforms (u1, u2, u3, tx, ty, tz) csrc/mvx_rigid.cuh:draw_rigid forms for one molecule."""
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    w = [philox4x32_10(np.array([mol_index & 0xFFFFFFFF, (mol_index >> 32) & 0xFFFFFFFF, d, 0x6D767874], dtype=np.uint32), key)
         for d in range(3)]
    words = np.concatenate(w).astype(np.uint64)
    u = [float(((words[2 * i] >> np.uint64(5)) << np.uint64(26)) | (words[2 * i + 1] >> np.uint64(6))) / 9007199254740992.0 for i in range(6)]
    return u
