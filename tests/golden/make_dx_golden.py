"""Writes tests/golden/dx_small.dx with the REFERENCE's own writer (pure Python, importable here):
python tests/golden/make_dx_golden.py   (needs /root/reference; run in the build container only)"""
import importlib.util
import os

import numpy as np

spec = importlib.util.spec_from_file_location("ref_dx", "/root/reference/molvoxel/etc/pymol/dx.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
rng = np.random.default_rng(42)
vals = rng.normal(size=(3, 4, 5)).astype(np.float32)
here = os.path.dirname(os.path.abspath(__file__))
np.save(os.path.join(here, "dx_small_values.npy"), vals)
ref.write_grid_to_dx_file(os.path.join(here, "dx_small.dx"), vals, (1.25, -2.5, 0.125), 0.375)
