"""GPU-side split of one small host call: CUDA-event times of prep / bin / voxelize (mvx_profile_*), and the wall time of
the C call's pieces measured by bracketing variants (device inputs without sync vs host inputs with sync)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molvoxel_b200 as mv
from molvoxel_b200 import _lib
rng = np.random.default_rng(0)
coords = rng.normal(scale=3.0, size=(44, 3)); types = rng.integers(0, 9, size=44); center = np.zeros(3)
vox = mv.create_voxelizer(0.5, 64, "scalar", "gaussian", library="b200")
grid = vox.get_empty_grid(9)
L = _lib.lib()
for _ in range(50): vox.forward(coords, center, types, 1.0, out_grid=grid)
torch.cuda.synchronize()
n = 500
L.mvx_profile_begin(n)
t0 = time.perf_counter()
for _ in range(n): vox.forward(coords, center, types, 1.0, out_grid=grid)
wall = (time.perf_counter() - t0) / n * 1e6
a, b, c, k = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
L.mvx_profile_end(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(k))
print(f"host call (with profile events): wall {wall:.1f} us; GPU prep {a.value/k.value*1e3:.1f} us, bin {b.value/k.value*1e3:.1f} us, voxelize {c.value/k.value*1e3:.1f} us over {k.value} calls")
# the same kernels back to back on the device path, one sync at the end: GPU time per call without host round trips
tc, tt, tz = torch.from_numpy(coords).cuda(), torch.from_numpy(types).cuda().int(), torch.zeros(3, dtype=torch.float64, device="cuda")
offs = torch.tensor([0, 44], dtype=torch.int32, device="cuda")
out5 = grid.unsqueeze(0)
call = lambda: vox.forward_types_batch(tc, offs, tz.reshape(1, 3), tt, 1.0, 9, out=out5)
for _ in range(20): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    call()
torch.cuda.synchronize()
e0.record()
for _ in range(n): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"device path captured in a CUDA graph: {e0.elapsed_time(e1)/n*1e3:.1f} us per replay (GPU-side chain of memset + prep + bin + voxelize)")
t0 = time.perf_counter()
for _ in range(n):
    g.replay(); torch.cuda.synchronize()
print(f"graph replay + synchronize: {(time.perf_counter()-t0)/n*1e6:.1f} us wall per call")
