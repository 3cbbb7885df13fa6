"""OpenDX export of one grid channel: the reference's visualisation hand-off,
molvoxel/etc/pymol/dx.py:2-39, byte-compatible output (5 decimals, three values per line)."""
from __future__ import annotations

import numpy as np


def write_grid_to_dx_file(dx_path, values, center, resolution):
    """values: (X, Y, Z) array-like (a torch tensor on any device is accepted); center: (3,)."""
    if hasattr(values, "detach"):
        values = values.detach().float().cpu().numpy()
    values = np.asarray(values)
    assert len(values.shape) == 3
    assert len(center) == 3
    size = values.shape
    origin = tuple(float(c) - resolution * (s - 1) / 2.0 for c, s in zip(center, size))
    head = [
        "object 1 class gridpositions counts {:d} {:d} {:d}\n".format(*size),
        "origin {:.5f} {:.5f} {:.5f}\n".format(*origin),
        f"delta {resolution:.5f} 0 0\n",
        f"delta 0 {resolution:.5f} 0\n",
        f"delta 0 0 {resolution:.5f}\n",
        "object 2 class gridconnections counts {:d} {:d} {:d}\n".format(*size),
        f"object 3 class array type double rank 0 items [ {size[0] * size[1] * size[2]:d} ] data follows\n",
    ]
    flat = values.reshape(-1).astype(np.float64).tolist()
    body = "".join(f"{v:.5f}\n" if i % 3 == 2 else f"{v:.5f} " for i, v in enumerate(flat))
    with open(dx_path, "w") as f:
        f.write("".join(head) + body)


def write_channels_to_dx(prefix, grid, center, resolution, names=None):
    """One .dx file per channel of a (C, X, Y, Z) grid; returns the paths."""
    paths = []
    for c in range(grid.shape[0]):
        p = f"{prefix}_{names[c] if names else c}.dx"
        write_grid_to_dx_file(p, grid[c], center, resolution)
        paths.append(p)
    return paths
