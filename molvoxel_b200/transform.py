"""Random rigid transform (rotation about the centre + translation) applied before voxelization.

Interface of reference molvoxel/voxelizer/base/transform.py:6-33 and numpy/transform.py:10-80
(`RandomTransform.forward`, `get_transform() -> T`, `T(coords, center)`, `T.create`).

Two sources of randomness:

* host draws from numpy's global RNG in the reference's order — `do_random_transform` draws the quaternion
  first (3 uniforms, numpy/_quaternion.py:13-21) and then the translation (3 uniforms, numpy/transform.py:74-76);
  `T.create` draws the translation first (numpy/transform.py:19-33) — so `np.random.seed(s)` reproduces the
  reference's transforms exactly.  `Voxelizer(rng="numpy")` uses these and hands them to the kernels as
  explicit (quaternion, translation) rows;
* the device generator (`Voxelizer(rng="philox")`, the default): Philox4x32-10 keyed by (seed, molecule index),
  drawn inside the per-atom prep kernel (csrc/mvx_rigid.cuh) — no host work per molecule.

Either way the transform itself is applied by the prep kernel with the reference's arithmetic, operation for
operation: the two quaternion products of numpy/_quaternion.py:28-54 and then the translation, added TWICE when
a rotation is also requested (numpy/transform.py:56-59 — numpy and numba backends) unless
`translate_once=True` asks for the torch backend's behaviour (torch/transform.py:56-60).  `do_transform` below is
the same arithmetic on the host (numpy arrays or torch tensors) for callers that use `T` / `RandomTransform`
directly, as the reference's tests do (test/test_run_numpy.py:34-40).
"""
from __future__ import annotations

import math

import numpy as np
import torch

PI2 = 2 * math.pi


def random_quaternion():
    """Uniform random rotation as a unit quaternion from three uniforms (numpy/_quaternion.py:13-21)."""
    u1, u2, u3 = np.random.rand(3)
    lo, hi = math.sqrt(1 - u1), math.sqrt(u1)
    return (lo * math.sin(PI2 * u2), lo * math.cos(PI2 * u2), hi * math.sin(PI2 * u3), hi * math.cos(PI2 * u3))


def random_translation_vector(random_translation: float):
    """(1, 3) float32, each component ~ U(-t, t) (numpy/transform.py:26, :76)."""
    return np.random.uniform(-random_translation, random_translation, size=(1, 3)).astype(np.float32)


def rotate(xyz, quaternion):
    """q (0, p) q^-1 with the reference's operation order (numpy/_quaternion.py:28-54); xyz is (V, 3), numpy or torch.
    Python-float quaternion components are weak scalars, so the arithmetic runs in xyz's dtype like the reference's."""
    q0, q1, q2, q3 = quaternion
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    o = x * 0
    a0 = q0 * o - q1 * x - q2 * y - q3 * z
    a1 = q0 * x + q1 * o + q2 * z - q3 * y
    a2 = q0 * y - q1 * z + q2 * o + q3 * x
    a3 = q0 * z + q1 * y - q2 * x + q3 * o
    b0, b1, b2, b3 = q0, q1 * -1, q2 * -1, q3 * -1
    rx = a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2
    ry = a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1
    rz = a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0
    lib = torch if isinstance(xyz, torch.Tensor) else np
    return lib.stack([rx, ry, rz], -1)


def do_transform(coords, center=None, translation=None, quaternion=None, translate_once: bool = False):
    """numpy/transform.py:43-60 on the host: rotation about `center`, then the translation — twice when rotating
    (the numpy backend's behaviour) unless translate_once."""
    is_t = isinstance(coords, torch.Tensor)
    if translation is not None and is_t:
        translation = torch.as_tensor(np.asarray(translation), device=coords.device)
    if quaternion is not None:
        if center is not None:
            center = center.reshape(1, 3)
            coords = rotate(coords - center, quaternion) + center
        else:
            coords = rotate(coords, quaternion)
        if translation is not None and not translate_once:
            coords = coords + translation
    if translation is not None:
        coords = coords + translation
    return coords


def do_random_transform(coords, center=None, random_translation=0.0, random_rotation=False, translate_once: bool = False):
    """numpy/transform.py:63-80: quaternion drawn first, then the translation."""
    quaternion = random_quaternion() if random_rotation else None
    translation = None
    if random_translation is not None and random_translation > 0.0:
        translation = random_translation_vector(random_translation)
    return do_transform(coords, center, translation, quaternion, translate_once)


class T:
    """A frozen transform (numpy/transform.py:10-33): reusable across calls and voxelizers."""

    translate_once = False

    def __init__(self, translation, quaternion):
        self.translation = translation
        self.quaternion = quaternion

    def __call__(self, coords, center):
        return do_transform(coords, center, self.translation, self.quaternion, self.translate_once)

    @classmethod
    def create(cls, random_translation: float = 0.0, random_rotation: bool = False):
        # the reference draws the translation FIRST here (numpy/transform.py:19-33), unlike do_random_transform
        translation = random_translation_vector(random_translation) if random_translation > 0.0 else None
        quaternion = random_quaternion() if random_rotation else None
        return cls(translation, quaternion)

    def as_row(self) -> np.ndarray:
        """(7,) float64 for the C ABI (mvx_batch.transforms): quaternion (identity if None), translation (0 if None)."""
        row = np.zeros(7, dtype=np.float64)
        row[:4] = (1.0, 0.0, 0.0, 0.0) if self.quaternion is None else self.quaternion
        if self.translation is not None:
            row[4:] = np.asarray(self.translation, dtype=np.float64).reshape(3)
        return row


class RandomTransform:
    class_T = T

    def __init__(self, random_translation: float = 0.0, random_rotation: bool = False):
        self.random_translation = random_translation
        self.random_rotation = random_rotation

    def forward(self, coords, center):
        return do_random_transform(coords, center, self.random_translation, self.random_rotation, self.class_T.translate_once)

    __call__ = forward

    def get_transform(self) -> T:
        return self.class_T.create(self.random_translation, self.random_rotation)


def host_transform_rows(num_mols: int, random_translation, random_rotation) -> np.ndarray:
    """(B, 7) float64 rows drawn from numpy's global RNG exactly as B consecutive reference forward_* calls would
    (quaternion first, then translation: numpy/transform.py:63-80)."""
    rows = np.zeros((num_mols, 7), dtype=np.float64)
    rows[:, 0] = 1.0   # identity quaternion (scalar part first)
    translate = random_translation is not None and random_translation > 0.0
    for m in range(num_mols):
        if random_rotation:
            rows[m, :4] = random_quaternion()
        if translate:
            rows[m, 4:] = random_translation_vector(random_translation).reshape(3)
    return rows


def transform_rows(transforms, num_mols: int) -> np.ndarray:
    """Explicit per-molecule transforms for the C ABI: a (B, 7) array, one T for all molecules, or a sequence of T."""
    if isinstance(transforms, T):
        return np.tile(transforms.as_row(), (num_mols, 1))
    if isinstance(transforms, (list, tuple)) and len(transforms) > 0 and isinstance(transforms[0], T):
        assert len(transforms) == num_mols, f"one transform per molecule: {len(transforms)} vs {num_mols}"
        return np.stack([t.as_row() for t in transforms])
    rows = np.ascontiguousarray(np.asarray(transforms, dtype=np.float64))
    assert rows.shape == (num_mols, 7), f"transforms should be (B, 7): {rows.shape} vs {(num_mols, 7)}"
    return rows
