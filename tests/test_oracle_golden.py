"""The oracle (oracle/mvx_oracle.c) against the golden vectors produced by the live reference."""
import numpy as np
import pytest

from oracle import OracleVoxelizer
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
