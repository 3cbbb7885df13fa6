"""ctypes front-end of oracle/mvx_oracle.c (CPU restatement of molvoxel's numpy backend).

TEST INFRASTRUCTURE ONLY — see the header of mvx_oracle.c.  The class mirrors the reference's
numpy ``Voxelizer`` (reference molvoxel/voxelizer/numpy/voxelizer.py:18-35) closely enough
that parity tests read like calls into the reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmvx_oracle.so")
_lib = None


class _Spec(ctypes.Structure):
    _fields_ = [
        ("resolution", ctypes.c_double),
        ("dimension", ctypes.c_int32),
        ("binary", ctypes.c_int32),
        ("sigma", ctypes.c_double),
        ("radii_mode", ctypes.c_int32),
        ("blockdim", ctypes.c_int32),
    ]


def build_oracle(force: bool = False) -> str:
    """Compile mvx_oracle.c with gcc (no GPU, no torch)."""
    src = os.path.join(_HERE, "mvx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = ctypes.CDLL(_SO)
        _lib.mvxo_forward_batch.restype = ctypes.c_int
        _lib.mvxo_forward.restype = ctypes.c_int
        _lib.mvxo_forward64.restype = ctypes.c_int
    return _lib


_RADII = {"scalar": 0, "channel-wise": 1, "atom-wise": 2}
_MODE = {"single": 0, "types": 1, "features": 2}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def oracle_forward_batch(resolution, dimension, radii_type, density_type, sigma, blockdim, mode,
                         mol_offsets, coords, centers, types, features, num_channels, radii,
                         out_channels=None, num_threads=1, precision=32):
    """B independent reference calls over a CSR batch; returns (B, C, D, D, D) float32 (float64 with precision=64,
    the reference's `precision` constructor argument, numpy/voxelizer.py:28-34)."""
    lib = _load()
    spec = _Spec(float(resolution), int(dimension), int(density_type == "binary"), float(sigma),
                 _RADII[radii_type], int(blockdim) if blockdim else 8)
    mol_offsets = np.ascontiguousarray(mol_offsets, dtype=np.int32)
    B = mol_offsets.shape[0] - 1
    coords = np.ascontiguousarray(coords)
    if coords.dtype not in (np.float32, np.float64):
        coords = coords.astype(np.float64)
    if centers is not None:
        centers = np.ascontiguousarray(centers)
        if centers.dtype not in (np.float32, np.float64):
            centers = centers.astype(np.float64)
    if types is not None:
        types = np.ascontiguousarray(np.asarray(types).astype(np.int16).astype(np.int32))
    if features is not None:
        features = np.ascontiguousarray(features, dtype=np.float32)
    radius = 0.0
    radii_arr = None
    # an np.float64 scalar is strongly typed (NEP 50): the reference divides in fp64 (numpy/voxelizer.py:546-548);
    # np.float32 / np.float16 scalars and python floats give fp32 arithmetic
    radius_kind = 1 if isinstance(radii, np.float64) else (2 if isinstance(radii, np.float32) else 0)
    if np.isscalar(radii):
        radius = float(radii)
    else:
        radii_arr = np.ascontiguousarray(radii, dtype=np.float32)
    C = int(num_channels)
    oc = int(out_channels) if out_channels is not None else C
    if precision == 64:   # molecule by molecule through mvxo_forward64
        out = np.empty((B, oc, dimension, dimension, dimension), dtype=np.float64)
        csz, zsz = coords.dtype.itemsize * 3, (centers.dtype.itemsize * 3 if centers is not None else 0)
        for m in range(B):
            a, b = int(mol_offsets[m]), int(mol_offsets[m + 1])
            cen = None if centers is None else ctypes.c_void_p(centers.ctypes.data + zsz * m)
            rad = radii_arr
            if radii_arr is not None and radii_type == "atom-wise":
                rad = radii_arr[a:b]
            rc = lib.mvxo_forward64(
                ctypes.byref(spec), _MODE[mode], b - a, ctypes.c_void_p(coords.ctypes.data + csz * a),
                int(coords.dtype == np.float64), cen, int(centers is not None and centers.dtype == np.float64),
                None if types is None else _ptr(types[a:b]), None if features is None else _ptr(features[a:b]),
                C, ctypes.c_double(radius), _ptr(rad), _ptr(out[m]), oc)
            if rc != 0:
                raise AssertionError(f"oracle rejected the arguments (code {rc})")
        return out
    out = np.empty((B, oc, dimension, dimension, dimension), dtype=np.float32)
    rc = lib.mvxo_forward_batch(
        ctypes.byref(spec), _MODE[mode], B, _ptr(mol_offsets), _ptr(coords), int(coords.dtype == np.float64),
        _ptr(centers), int(centers is not None and centers.dtype == np.float64), _ptr(types), _ptr(features),
        C, ctypes.c_double(radius), _ptr(radii_arr), _ptr(out), oc, int(num_threads), radius_kind)
    if rc != 0:
        raise AssertionError(f"oracle rejected the arguments (code {rc})")
    return out


class OracleVoxelizer:
    """Single-molecule front-end with the reference's constructor and forward_* signatures."""

    def __init__(self, resolution=0.5, dimension=64, radii_type="scalar", density_type="gaussian",
                 blockdim=None, sigma=0.5):
        self.resolution, self.dimension = resolution, dimension
        self.radii_type, self.density_type = radii_type, density_type
        self.blockdim = blockdim if blockdim is not None else 8
        self.sigma = sigma

    def _run(self, mode, coords, center, types, features, C, radii, out_channels=None):
        coords = np.asarray(coords)
        V = coords.shape[0]
        offs = np.array([0, V], dtype=np.int32)
        centers = None if center is None else np.asarray(center).reshape(1, 3)
        return oracle_forward_batch(self.resolution, self.dimension, self.radii_type, self.density_type,
                                    self.sigma, self.blockdim, mode, offs, coords, centers, types, features,
                                    C, radii, out_channels)[0]

    def forward_types(self, coords, center, types, radii, out_channels=None):
        types = np.asarray(types)
        if self.radii_type == "channel-wise":
            C = int(np.asarray(radii).shape[0])
        else:
            C = int(types.max()) + 1  # V == 0 raises ValueError like the reference (numpy/voxelizer.py:325)
        return self._run("types", coords, center, types, None, C, radii, out_channels)

    def forward_features(self, coords, center, features, radii):
        features = np.asarray(features, dtype=np.float32)
        return self._run("features", coords, center, None, features, features.shape[1], radii)

    def forward_single(self, coords, center, radii):
        return self._run("single", coords, center, None, None, 1, radii)
