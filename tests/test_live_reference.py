"""Differential tests against the LIVE, unmodified reference (baseline/_ref; /root/reference in the build container).

CPU part (this file, not marked gpu): the oracle (oracle/mvx_oracle.c) and the host-side transform restatement
(molvoxel_b200/transform.py) against the reference's numpy backend on random inputs — every mode x radii type x
density x blockdim x input dtype x scalar-radius typing.  The GPU twin lives in tests/test_gpu_parity.py
(test_cuda_matches_live_reference_*).  Skipped when the reference is not importable.
"""
import os

import numpy as np
import pytest

from oracle import OracleVoxelizer
from tests.helpers import import_reference, philox4x32_10

molvoxel = import_reference()
needs_ref = pytest.mark.skipif(molvoxel is None, reason="reference package not available (baseline/_ref)")

N_FUZZ = int(os.environ.get("MVX_LIVE_FUZZ", "240"))


def random_case(seed):
    rng = np.random.default_rng(50_000 + seed)
    dim = int(rng.choice([6, 8, 11, 16, 20, 24]))
    res = float(rng.choice([0.25, 0.375, 0.4, 0.5, 0.8]))
    mode = ["types", "features", "single"][seed % 3]
    density = ["gaussian", "binary"][(seed // 3) % 2]
    radii_type = str(rng.choice(["scalar", "atom-wise"] if mode == "single" else ["scalar", "atom-wise", "channel-wise"]))
    bd = [None, 3, 4, 8, dim][int(rng.integers(0, 5))]
    sigma = float(rng.choice([0.5, 0.35, 1.0]))
    C = 1 if mode == "single" else int(rng.integers(1, 7))
    V = int(rng.integers(1, 120))
    half = res * (dim - 1) / 2
    coords = rng.uniform(-half - 1.5, half + 1.5, size=(V, 3))
    if rng.uniform() < 0.25:   # atoms exactly on grid points: d == r ties
        axis = np.arange(dim) * res - half
        coords[: V // 2] = axis[rng.integers(0, dim, size=(V // 2, 3))]
    if rng.uniform() < 0.3:
        coords = coords.astype(np.float32)
    center = None if rng.uniform() < 0.3 else rng.normal(scale=0.4, size=3).astype(coords.dtype if rng.uniform() < 0.5 else np.float64)
    rmax = float(rng.uniform(0.6, 2.2))
    if radii_type == "scalar":
        kind = int(rng.integers(0, 4))
        radii = [rmax, np.float64(rmax), np.float32(rmax), float(np.float32(rmax))][kind]
    elif radii_type == "atom-wise":
        radii = rng.uniform(0.4 * rmax, rmax, size=V).astype(np.float32)
    else:
        radii = rng.uniform(0.4 * rmax, rmax, size=C).astype(np.float32)
    types = feats = None
    if mode == "types":
        types = rng.integers(0, C, size=V).astype(np.int16)
        types[int(rng.integers(0, V))] = C - 1
    elif mode == "features":
        feats = rng.uniform(0, 1, size=(V, C)).astype(np.float32)
    return dict(dim=dim, res=res, mode=mode, density=density, radii_type=radii_type, bd=bd, sigma=sigma, C=C,
                coords=coords, center=center, radii=radii, types=types, feats=feats)


def run(vox, c):
    if c["mode"] == "types":
        return vox.forward_types(c["coords"], c["center"], c["types"], c["radii"])
    if c["mode"] == "features":
        return vox.forward_features(c["coords"], c["center"], c["feats"], c["radii"])
    return vox.forward_single(c["coords"], c["center"], c["radii"])


@needs_ref
@pytest.mark.parametrize("seed", range(N_FUZZ))
def test_oracle_matches_live_reference(seed):
    """numpy/voxelizer.py:97-169, :240-315, :370-436 run live vs the C restatement: binary bit-exact (types / single),
    Gaussian identical support and <= 2e-6 of the peak."""
    c = random_case(seed)
    kw = {} if c["bd"] is None else {"blockdim": c["bd"]}
    ref = run(molvoxel.create_voxelizer(c["res"], c["dim"], c["radii_type"], c["density"], library="numpy", sigma=c["sigma"], **kw), c)
    got = run(OracleVoxelizer(c["res"], c["dim"], c["radii_type"], c["density"], blockdim=c["bd"], sigma=c["sigma"]), c)
    assert got.shape == ref.shape and ref.dtype == np.float32
    if c["density"] == "binary" and c["mode"] != "features":
        assert np.array_equal(got, ref), f"{(got != ref).sum()} voxels differ"
    else:
        assert np.array_equal(got != 0, ref != 0)
        assert float(np.abs(got - ref).max()) <= 2e-6 * max(1.0, float(np.abs(ref).max()))


@needs_ref
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rt,rr", [(0.5, True), (0.0, True), (0.5, False)])
def test_host_transform_is_the_reference_transform_bitwise(rt, rr, dtype):
    """molvoxel_b200.transform vs numpy/transform.py:10-80 under the same np.random.seed: T.create (translation drawn
    first), RandomTransform.forward (quaternion drawn first), do_transform arithmetic incl. the double translation."""
    import molvoxel_b200.transform as ours
    from molvoxel.voxelizer.numpy import transform as theirs
    rng = np.random.default_rng(5)
    coords = rng.normal(scale=4.0, size=(200, 3)).astype(dtype)
    center = coords.mean(0)
    for cen in (center, None):
        np.random.seed(11)
        t_ref = theirs.RandomTransform(rt, rr).get_transform()
        np.random.seed(11)
        t_our = ours.RandomTransform(rt, rr).get_transform()
        assert (t_ref.quaternion is None) == (t_our.quaternion is None)
        if rr:
            assert tuple(t_ref.quaternion) == tuple(t_our.quaternion)
        if rt > 0:
            assert np.array_equal(t_ref.translation, t_our.translation) and t_our.translation.dtype == np.float32
        a, b = t_ref(coords.copy(), cen), t_our(coords.copy(), cen)
        assert a.dtype == b.dtype and np.array_equal(a, b)
        np.random.seed(12)
        a = theirs.RandomTransform(rt, rr).forward(coords.copy(), cen)
        np.random.seed(12)
        b = ours.RandomTransform(rt, rr).forward(coords.copy(), cen)
        assert a.dtype == b.dtype and np.array_equal(a, b)


@needs_ref
def test_host_transform_rows_follow_the_reference_draw_order():
    """B consecutive reference forward_* calls draw (quaternion, translation) per molecule from the global RNG."""
    import molvoxel_b200.transform as ours
    from molvoxel.voxelizer.numpy import _quaternion as rq
    np.random.seed(3)
    rows = ours.host_transform_rows(5, 0.5, True)
    np.random.seed(3)
    for m in range(5):
        q = rq.random_quaternion()
        t = np.random.uniform(-0.5, 0.5, size=(1, 3)).astype(np.float32)
        assert tuple(rows[m, :4]) == tuple(q) and np.array_equal(rows[m, 4:], t.reshape(3).astype(np.float64))


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (pins the host restatement the GPU tests compare the device
    generator with)."""
    def hx(x):
        return [int(v) for v in x]
    assert hx(philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.full(4, 0xFFFFFFFF, np.uint32)
    assert hx(philox4x32_10(f, f[:2])) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert hx(philox4x32_10(np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32),
                            np.array([0xA4093822, 0x299F31D0], np.uint32))) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
