"""Debug: cfg2 pool batch vs oracle per kernel form (run on the GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import molvoxel_b200 as mv
from oracle import oracle_forward_batch

b = bench.make_batch("cfg2", 256, seed=1000)
for m in (211, 256, 64):
    offs = b["offs"][:m + 1]; na = int(offs[-1])
    coords, feats = b["coords"][:na], b["feats"][:na]
    ref = oracle_forward_batch(0.5, 48, "scalar", "gaussian", 0.5, 8, "features", offs, coords, np.zeros((m, 3)), None, feats, 16, 1.0, num_threads=32)
    for form in ("pipe", "tiles", "cells", "rows"):
        os.environ["MVX_KERNEL"] = form
        vox = mv.create_voxelizer(0.5, 48, "scalar", "gaussian", library="b200")
        for dev_in in (True, False):
            if dev_in:
                got = vox.forward_features_batch(torch.from_numpy(coords).cuda(), torch.from_numpy(offs).cuda(), torch.zeros((m, 3), dtype=torch.float64, device="cuda"), torch.from_numpy(feats).cuda(), 1.0)
            else:
                got = vox.forward_features_batch(coords, offs, np.zeros((m, 3)), feats, 1.0)
            vox.check_status()
            got = got.cpu().numpy()
            bad = np.abs(got - ref) > 1e-5
            mols = np.unique(np.argwhere(bad)[:, 0]) if bad.any() else []
            print(f"B={m} form={form} dev_in={dev_in}: mismatching voxels {int(bad.sum())}, molecules {list(mols)[:10]}, max err {float(np.abs(got-ref).max()):.3e}", flush=True)
            if bad.any():
                idx = np.argwhere(bad)[:5]
                for i in idx:
                    print("   ", tuple(i), got[tuple(i)], ref[tuple(i)])
