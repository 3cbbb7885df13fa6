"""Debug: warp-specialised pipelined form vs the classic one on a cfg5-like case; where do they differ?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molvoxel_b200 as mv

def run(ws, V, C, dim, res, rmax, seed=5):
    os.environ["MVX_WS"] = str(ws)
    os.environ["MVX_KERNEL"] = "pipe"
    rng = np.random.default_rng(seed)
    B = 2
    offs = (np.arange(B + 1) * V).astype(np.int32)
    w = res * (dim - 1)
    coords = rng.uniform(-w / 2, w / 2, size=(B * V, 3)).astype(np.float32).astype(np.float64)
    feats = rng.uniform(size=(B * V, C)).astype(np.float32)
    radii = rng.uniform(1.0, rmax, size=B * V).astype(np.float32)
    vox = mv.create_voxelizer(res, dim, "atom-wise", "gaussian", library="b200")
    out = vox.forward_features_batch(coords, offs, None, feats, radii)
    torch.cuda.synchronize()
    return out

for (V, C, dim, res, rmax) in [(10000, 32, 96, 0.375, 2.0), (3000, 32, 96, 0.375, 2.0), (10000, 16, 96, 0.375, 2.0), (10000, 32, 96, 0.375, 1.2)]:
    a = run(0, V, C, dim, res, rmax)
    for ws in (4, 8):
        b = run(ws, V, C, dim, res, rmax)
        d = (a != b)
        n = int(d.sum())
        print(f"V={V} C={C} rmax={rmax} ws={ws}: differing voxels {n}")
        if n:
            idx = d.nonzero()[:12].cpu().numpy()
            chans = torch.unique(d.nonzero()[:, 1]).cpu().numpy()
            print("   channels:", chans[:40])
            print("   first:", idx.tolist())
            mols = d.nonzero()
            # cells: (x//2, y//4, z//16)
            cells = torch.unique(torch.stack([mols[:, 0], mols[:, 2] // 2, mols[:, 3] // 4, mols[:, 4] // 16], 1), dim=0)
            print("   distinct (mol, cell) =", cells.shape[0], cells[:8].cpu().numpy().tolist())
            bb = b[d]; aa = a[d]
            print("   sample a:", aa[:6].cpu().numpy(), " b:", bb[:6].cpu().numpy())
