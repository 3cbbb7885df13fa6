"""Where the time of one small forward() goes: Python wrapper vs the C-ABI call (mvx_voxelize_host), run on the GPU box."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molvoxel_b200 as mv
from molvoxel_b200 import _lib
rng = np.random.default_rng(0)
coords = rng.normal(scale=3.0, size=(44, 3)); types = rng.integers(0, 9, size=44); center = np.zeros(3)
vox = mv.create_voxelizer(0.5, 64, "scalar", "gaussian", library="b200")
grid = vox.get_empty_grid(9)
for _ in range(20): vox.forward(coords, center, types, 1.0, out_grid=grid); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): vox.forward(coords, center, types, 1.0, out_grid=grid); torch.cuda.synchronize()
print("per call us", (time.perf_counter() - t0) / 300 * 1e6)
L = _lib.lib()
orig = L.mvx_voxelize_host
acc = [0.0, 0]
class Wrap:
    def __call__(self, *a):
        t = time.perf_counter(); r = orig(*a); acc[0] += time.perf_counter() - t; acc[1] += 1; return r
L.mvx_voxelize_host = Wrap()
t0 = time.perf_counter()
for _ in range(300): vox.forward(coords, center, types, 1.0, out_grid=grid)
tot = time.perf_counter() - t0
print("total us", tot / 300 * 1e6, " C call us", acc[0] / max(1, acc[1]) * 1e6, " python us", (tot - acc[0]) / 300 * 1e6)
# device-input path, no sync inside: enqueue cost and GPU time
L.mvx_voxelize_host = orig
tc, tt, tz = torch.from_numpy(coords).cuda(), torch.from_numpy(types).cuda().int(), torch.zeros(3, dtype=torch.float64, device="cuda")
for _ in range(20): vox.forward(tc, tz, tt, 1.0, out_grid=grid)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(300): vox.forward(tc, tz, tt, 1.0, out_grid=grid)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("device inputs: enqueue us/call", (t1 - t0) / 300 * 1e6, " GPU us/call", e0.elapsed_time(e1) / 300 * 1e3)
