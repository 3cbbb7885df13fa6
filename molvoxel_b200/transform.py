"""Random rigid transform (rotation about the centre + translation) applied before voxelization.

Interface of reference molvoxel/voxelizer/base/transform.py:6-33 and numpy/transform.py:10-80
(`RandomTransform.forward`, `get_transform() -> T`, `T(coords, center)`).  Random numbers come
from numpy's global RNG in the reference's draw order (3 uniforms for the quaternion,
numpy/_quaternion.py:13-21, then 3 for the translation, numpy/transform.py:74-76), so
`np.random.seed(s)` reproduces the reference's transforms.  The rotation is applied as a 3x3
matrix built from the unit quaternion.  Deliberate deviation: the numpy backend
adds the translation twice when a rotation is also requested (numpy/transform.py:56-59); this
backend applies it once, like the reference's torch backend (torch/transform.py:56-60).
The numerics of this step are outside the parity metric (RNG-dependent).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def random_unit_quaternion():
    u1, u2, u3 = np.random.rand(3)
    a, b = math.sqrt(1.0 - u1), math.sqrt(u1)
    return (a * math.sin(2 * math.pi * u2), a * math.cos(2 * math.pi * u2),
            b * math.sin(2 * math.pi * u3), b * math.cos(2 * math.pi * u3))


def quaternion_to_matrix(q) -> np.ndarray:
    """Rotation matrix of v -> q v q^-1 for a unit quaternion q = (w, x, y, z)."""
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def _draw(random_translation, random_rotation):
    rot = quaternion_to_matrix(random_unit_quaternion()) if random_rotation else None
    if random_translation is not None and random_translation > 0.0:
        tr = np.random.uniform(-random_translation, random_translation, size=(1, 3)).astype(np.float32).astype(np.float64)
    else:
        tr = None
    return rot, tr


def random_transform_params(num_mols, random_translation, random_rotation):
    """One (rotation | None, translation | None) pair per molecule."""
    return [_draw(random_translation, random_rotation) for _ in range(num_mols)]


def transform_matrix_array(params) -> np.ndarray:
    """(B, 12) float64 for the C ABI: row-major rotation (identity if None) followed by the translation."""
    out = np.zeros((len(params), 12), dtype=np.float64)
    for m, (rot, tr) in enumerate(params):
        out[m, :9] = (np.eye(3) if rot is None else rot).reshape(-1)
        if tr is not None:
            out[m, 9:] = np.asarray(tr, dtype=np.float64).reshape(-1)
    return out


def _apply_one(xyz, center, rot, tr):
    lib = torch if isinstance(xyz, torch.Tensor) else np
    if isinstance(xyz, torch.Tensor):
        conv = lambda a: torch.as_tensor(a, dtype=xyz.dtype, device=xyz.device)   # noqa: E731
    else:
        conv = lambda a: np.asarray(a, dtype=xyz.dtype)   # noqa: E731
    if rot is not None:
        if center is not None:
            xyz = lib.matmul(xyz - center, conv(rot).T) + center
        else:
            xyz = lib.matmul(xyz, conv(rot).T)
    if tr is not None:
        xyz = xyz + conv(tr)
    return xyz


def apply_transform(coords, mol_offsets, centers, params):
    """Centre each molecule, rotate about the origin, translate.  Returns (coords', None): the result
    is already centred, like the reference which transforms after subtracting the centre."""
    is_t = isinstance(coords, torch.Tensor)
    x = coords.to(torch.float64) if is_t else np.asarray(coords, dtype=np.float64)
    offs = mol_offsets.tolist() if hasattr(mol_offsets, "tolist") else list(mol_offsets)
    parts = []
    for m, (rot, tr) in enumerate(params):
        seg = x[offs[m]:offs[m + 1]]
        if centers is not None:
            c = centers[m].reshape(1, 3)
            c = (c.to(x.device, torch.float64) if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c, dtype=np.float64), device=x.device)) if is_t \
                else np.asarray(c.detach().cpu().numpy() if isinstance(c, torch.Tensor) else c, dtype=np.float64)
            seg = seg - c
        parts.append(_apply_one(seg, None, rot, tr))
    out = (torch.cat(parts, 0) if is_t else np.concatenate(parts, 0)) if parts else x
    return out, None


class T:
    """A frozen transform (numpy/transform.py:10-33): reusable across calls."""

    def __init__(self, translation, rotation):
        self.translation = translation
        self.rotation = rotation

    def __call__(self, coords, center):
        if isinstance(center, torch.Tensor) or isinstance(center, np.ndarray):
            center = center.reshape(1, 3)
        return _apply_one(coords, center, self.rotation, self.translation)

    @classmethod
    def create(cls, random_translation: float = 0.0, random_rotation: bool = False):
        rot, tr = _draw(random_translation, random_rotation)
        return cls(tr, rot)


class RandomTransform:
    class_T = T

    def __init__(self, random_translation: float = 0.0, random_rotation: bool = False):
        self.random_translation = random_translation
        self.random_rotation = random_rotation

    def forward(self, coords, center):
        return self.get_transform()(coords, center)

    __call__ = forward

    def get_transform(self) -> T:
        return self.class_T.create(self.random_translation, self.random_rotation)
