"""Small invocations of every kernel form, for compute-sanitizer (memcheck / racecheck) runs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import molvoxel_b200 as mv  # noqa: E402


def run(form):
    os.environ["MVX_KERNEL"] = form
    rng = np.random.default_rng(0)
    B = 3
    counts = np.array([40, 0, 700])
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(counts)
    N = int(offs[-1])
    coords = rng.uniform(-7, 7, size=(N, 3))
    for dim in (24, 21) if form == "rows" else (24, 40):
        vox = mv.create_voxelizer(0.5, dim, "atom-wise", "gaussian", library="b200")
        radii = rng.uniform(0.8, 1.8, size=N).astype(np.float32)
        t = vox.forward_types_batch(coords, offs, None, rng.integers(0, 5, size=N), radii, 5)
        f = vox.forward_features_batch(coords, offs, None, rng.uniform(size=(N, 20)).astype(np.float32), radii)
        s = vox.forward_single_batch(coords, offs, None, radii, random_translation=0.3, random_rotation=True)
        vox.check_status()
        print(form, dim, float(t.sum()), float(f.sum()), float(s.sum()))


for form in sys.argv[1:] or ["cells", "tiles", "rows"]:
    run(form)
print("sanitize_case ok")
