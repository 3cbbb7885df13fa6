"""Golden fixtures for an np.float64 SCALAR radius (numpy/voxelizer.py:546-548), from the LIVE reference.

    python tests/golden/make_golden_r64.py

An np.float64 scalar is strongly typed under NEP 50: ``np.divide(dist32, radii)`` promotes to fp64, so the
cutoff compares in fp64 and the Gaussian is evaluated in fp64, unlike a python float / np.float32 radius
(fp32 arithmetic).  The occupancy differs for voxels whose fp32 distance lies between the radius and its
fp32 rounding; the adversarial case below puts atoms exactly there.  Fixtures are written next to the
others as r64_*.npz with ``"radius_f64": true`` in their cfg (tests/helpers.py turns the stored radius
back into an np.float64).
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import save_case, synth  # noqa: E402  (imports the live reference)


def main():
    rng = np.random.default_rng(6464)
    dim, res = 24, 0.5
    half = res * (dim - 1) / 2.0
    axis = np.arange(dim) * res - half
    # adversarial: atoms at an axis-aligned distance of exactly fp32(r) from some voxel centres.  fp32(1.1) > 1.1 and
    # fp32(1.3) < 1.3: d32 == fp32(r) is a hit for a python-float radius and, for np.float64(1.1), a miss.
    for r in (1.1, 1.3, 0.7):
        r32 = float(np.float32(r))
        pts = []
        for _ in range(40):
            p = axis[rng.integers(4, dim - 4, size=3)].copy()
            k = int(rng.integers(0, 3))
            p[k] += r32 * (1 if rng.uniform() < 0.5 else -1)
            pts.append(p)
        for _ in range(40):   # generic positions as well
            pts.append(rng.uniform(-half, half, size=3))
        pts = np.array(pts, dtype=np.float64)
        types = rng.integers(0, 3, size=pts.shape[0]).astype(np.int16)
        tag = str(r).replace(".", "")
        for dens in ("binary", "gaussian"):
            base = dict(resolution=res, dimension=dim, radii_type="scalar", density_type=dens, blockdim=None, radius_f64=True)
            save_case(f"r64_ties_r{tag}_{dens}_types", dict(mode="types", **base), pts, None, types, np.float64(r))
        # the python-float twin of the binary case: pins that the two differ exactly where expected
        save_case(f"r64_ties_r{tag}_binary_types_pyfloat",
                  dict(mode="types", resolution=res, dimension=dim, radii_type="scalar", density_type="binary", blockdim=None),
                  pts, None, types, float(r))
    # every mode with a generic np.float64 radius (dim 20: blocks 8 + 8 + 4)
    V = 300
    coords = synth(rng, V, 20, 0.5)
    center = rng.uniform(-0.3, 0.3, size=3)
    types = rng.integers(0, 5, size=V).astype(np.int16)
    feats = rng.uniform(0, 1, size=(V, 6)).astype(np.float32)
    for dens in ("gaussian", "binary"):
        base = dict(resolution=0.5, dimension=20, radii_type="scalar", density_type=dens, blockdim=None, radius_f64=True)
        save_case(f"r64_d20_{dens}_types", dict(mode="types", **base), coords, center, types, np.float64(1.3))
        save_case(f"r64_d20_{dens}_feat", dict(mode="features", **base), coords, center, feats, np.float64(1.15))
        save_case(f"r64_d20_{dens}_single", dict(mode="single", **base), coords, center, None, np.float64(1.7))
    # np.float32 scalar: fp32 arithmetic like a python float, radius used as is
    save_case("r32scalar_d20_binary_types",
              dict(mode="types", resolution=0.5, dimension=20, radii_type="scalar", density_type="binary", blockdim=None, radius_f32=True),
              coords, center, types, np.float32(1.3))
    # ... and its clip thresholds are fp32 results (python float -/+ np.float32): atoms just inside / outside both the
    # fp32 and the fp64 threshold, on every face
    dim, res = 24, 0.5
    half = res * (dim - 1) / 2.0
    axis = np.arange(dim) * res - half
    r = np.float32(1.3)
    pts = []
    for sgn in (-1.0, 1.0):
        t32 = float(np.float32(sgn * half) + np.float32(sgn) * r)   # fp32 arithmetic, like lower - r / upper + r
        t64 = sgn * half + sgn * float(r)
        for k in range(3):
            for t in (t32, t64):
                for v in (t, np.nextafter(t, 0.0), np.nextafter(t, sgn * 100.0)):
                    p = axis[rng.integers(2, dim - 2, size=3)].copy()
                    p[k] = v
                    pts.append(p)
    pts = np.array(pts, dtype=np.float64)
    save_case("r32scalar_clip_d24_binary_types",
              dict(mode="types", resolution=res, dimension=dim, radii_type="scalar", density_type="binary", blockdim=None, radius_f32=True),
              pts, None, rng.integers(0, 3, size=pts.shape[0]).astype(np.int16), r)
    save_case("r32scalar_clip_d24_binary_types_pyfloat",
              dict(mode="types", resolution=res, dimension=dim, radii_type="scalar", density_type="binary", blockdim=None),
              pts, None, rng.integers(0, 3, size=pts.shape[0]).astype(np.int16), float(r))


if __name__ == "__main__":
    main()
