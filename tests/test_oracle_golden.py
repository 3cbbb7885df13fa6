"""The oracle (oracle/mvx_oracle.c) against the golden vectors produced by the live reference."""
import os

import numpy as np
import pytest

from oracle import OracleVoxelizer, oracle_forward_batch
from tests.helpers import GoldenCase, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g = GoldenCase(name)
    cfg = g.cfg
    vox = OracleVoxelizer(cfg["resolution"], cfg["dimension"], cfg["radii_type"], cfg["density_type"],
                          blockdim=cfg.get("blockdim"), sigma=cfg.get("sigma", 0.5))
    if cfg["mode"] == "types":
        out = vox.forward_types(g.coords, g.center, g.channels, g.radii)
    elif cfg["mode"] == "features":
        out = vox.forward_features(g.coords, g.center, g.channels, g.radii)
    else:
        out = vox.forward_single(g.coords, g.center, g.radii)
    g.check(out)


def test_known_answers_from_survey():
    """SURVEY.md §8c smoke KATs measured on the live numpy backend."""
    g = GoldenCase("lig10gs_types_gaussian")
    ref = g.dense()
    assert abs(float(ref.sum()) - 378.185120) < 1e-3 and int((ref != 0).sum()) == 994
    vox = OracleVoxelizer(0.5, 8, "scalar", "gaussian", blockdim=8)
    out = vox.forward_types(np.array([[0.25, 0.25, 0.25]]), None, np.array([0]), 1.0)
    assert int((out != 0).sum()) == 33 and abs(float(out[out != 0].min()) - 0.135335) < 1e-6
    # atom exactly r outside the box is dropped by the strict clip (numpy/voxelizer.py:487-488)
    vox = OracleVoxelizer(0.5, 8, "scalar", "binary", blockdim=8)
    out = vox.forward_types(np.array([[1.75 + 1.0, 0.25, 0.25]]), None, np.array([0]), 1.0)
    assert int((out != 0).sum()) == 0


def _p64_names():
    import glob
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "p64")
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(d, "*.npz")))


def load_p64(name):
    import json
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "p64", name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    ref = np.zeros(int(np.prod(z["shape"])), dtype=np.float64)
    ref[z["nz_idx"]] = z["nz_val"]
    r = z["radii"]
    return dict(cfg=cfg, coords=z["coords"], center=z["center"] if "center" in z.files else None,
                channels=z["channels"] if "channels" in z.files else None,
                radii=float(r) if r.ndim == 0 else r, ref=ref.reshape(tuple(int(v) for v in z["shape"])))


@pytest.mark.parametrize("name", _p64_names())
def test_oracle_precision64_matches_reference_golden(name):
    """precision=64 fixtures from the live reference (tests/golden/make_golden_p64.py): binary bit-exact for
    types/single, everything else to 1e-12 of the peak (libm exp vs numpy exp; dgemm summation order)."""
    g = load_p64(name)
    cfg = g["cfg"]
    V = g["coords"].shape[0]
    mode = cfg["mode"]
    types = g["channels"] if mode == "types" else None
    feats = g["channels"] if mode == "features" else None
    if mode == "types":
        C = g["radii"].shape[0] if cfg["radii_type"] == "channel-wise" else int(types.max()) + 1
    else:
        C = feats.shape[1] if mode == "features" else 1
    out = oracle_forward_batch(cfg["resolution"], cfg["dimension"], cfg["radii_type"], cfg["density_type"], 0.5,
                               cfg.get("blockdim"), mode, np.array([0, V], dtype=np.int32), g["coords"],
                               None if g["center"] is None else g["center"].reshape(1, 3), types, feats, C, g["radii"],
                               precision=64)[0]
    assert out.dtype == np.float64 and out.shape == g["ref"].shape
    if cfg["density_type"] == "binary" and mode != "features":
        assert np.array_equal(out, g["ref"])
    else:
        assert np.array_equal(out != 0, g["ref"] != 0)
        assert float(np.abs(out - g["ref"]).max()) <= 1e-12 * max(1.0, float(np.abs(g["ref"]).max()))
