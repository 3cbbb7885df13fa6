"""cProfile of the Python side of one small forward() (numpy inputs, reference calling pattern)."""
import cProfile, pstats, sys, os, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molvoxel_b200 as mv
rng = np.random.default_rng(0)
coords = rng.normal(scale=3.0, size=(44, 3)); types = rng.integers(0, 9, size=44); center = np.zeros(3)
vox = mv.create_voxelizer(0.5, 64, "scalar", "gaussian", library="b200")
grid = vox.get_empty_grid(9)
for _ in range(50): vox.forward(coords, center, types, 1.0, out_grid=grid)
pr = cProfile.Profile()
pr.enable()
for _ in range(3000): vox.forward(coords, center, types, 1.0, out_grid=grid)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
