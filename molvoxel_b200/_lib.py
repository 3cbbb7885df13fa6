"""Build + ctypes binding of libmolvoxel_b200.so (C ABI: include/molvoxel_b200.h).

The shared library is compiled in-tree by nvcc for sm_100a only.  There is no CPU fallback: if the
library is missing it is built; if it cannot be built or loaded, importing the binding raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("MVX_SO") or os.path.join(CSRC, "libmolvoxel_b200.so")   # MVX_SO: an experimental build
OBJ_DIR = os.path.join(CSRC, "_obj")
API_SOURCE = os.path.join(CSRC, "mvx_api.cu")
INST_SOURCE = os.path.join(CSRC, "mvx_vox_inst.cu")
SOURCES = [API_SOURCE, INST_SOURCE]
HEADERS = [os.path.join(CSRC, h) for h in ("mvx_common.cuh", "mvx_rigid.cuh", "mvx_bin_kernels.cuh", "mvx_vox_kernels.cuh", "mvx_vox_ws.cuh", "mvx_launch.cuh")] + \
          [os.path.join(os.path.dirname(_HERE), "include", "molvoxel_b200.h")]
# (mode, channel chunk) pairs the voxelize kernels are instantiated for (x binary / gaussian): mvx_api.cu:launch_vox
INSTANCES = [(0, 1)] + [(m, ch) for m in (1, 2) for ch in (1, 4, 8, 12, 16)]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]

MVX_OK = 0
MVX_ERR_NULL_POINTER, MVX_ERR_BAD_ENUM, MVX_ERR_BAD_SHAPE = -1, -2, -3
MVX_ERR_WORKSPACE, MVX_ERR_CUDA, MVX_ERR_UNSUPPORTED, MVX_ERR_DEVICE_FLAG = -4, -5, -6, -7
MVX_F32, MVX_F64, MVX_U8, MVX_F16 = 0, 1, 2, 3
DENSITY = {"gaussian": 0, "binary": 1}
RADII = {"scalar": 0, "channel-wise": 1, "atom-wise": 2}
MODE = {"single": 0, "types": 1, "features": 2}
FORM_KERNEL = {0: "mvx_voxelize_kernel", 1: "mvx_voxelize_cells_kernel", 3: "mvx_voxelize_tiles_kernel", 4: "mvx_voxelize_pipe_kernel"}
OUT_DTYPE = {"float32": 0, "bfloat16": 1, "float16": 2, "float64": 3}
RADIUS_PYFLOAT, RADIUS_NP_F64, RADIUS_NP_F32 = 0, 1, 2
LAYOUT_CDHW, LAYOUT_DHWC = 0, 1
TF_ROTATE, TF_TRANSLATE, TF_TRANSLATE_ONCE = 1, 2, 4


class GridSpec(ctypes.Structure):
    _fields_ = [
        ("resolution", ctypes.c_double),
        ("dimension", ctypes.c_int32),
        ("density_type", ctypes.c_int32),
        ("sigma", ctypes.c_double),
        ("radii_type", ctypes.c_int32),
        ("compat_blockdim", ctypes.c_int32),
    ]


class Batch(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int32),
        ("num_mols", ctypes.c_int32),
        ("total_atoms", ctypes.c_int64),
        ("mol_offsets", ctypes.c_void_p),
        ("coords", ctypes.c_void_p),
        ("coords_dtype", ctypes.c_int32),
        ("centers", ctypes.c_void_p),
        ("centers_dtype", ctypes.c_int32),
        ("types", ctypes.c_void_p),
        ("features", ctypes.c_void_p),
        ("num_channels", ctypes.c_int32),
        ("out_channels", ctypes.c_int32),
        ("radius", ctypes.c_double),
        ("radii", ctypes.c_void_p),
        ("max_radius", ctypes.c_double),
        ("transforms", ctypes.c_void_p),
        ("out_dtype", ctypes.c_int32),
        ("features_dtype", ctypes.c_int32),
        ("radius_kind", ctypes.c_int32),
        ("transform_flags", ctypes.c_int32),
        ("rng_seed", ctypes.c_uint64),
        ("rng_offset", ctypes.c_uint64),
        ("random_translation", ctypes.c_double),
        ("out_layout", ctypes.c_int32),
    ]


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    if os.environ.get("MVX_SO"):   # an experimental build is taken as it is
        return False
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("molvoxel_b200: nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return proc.stderr


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU).

    The API unit and one unit per (mode, channel chunk, density) instantiation of the voxelize kernels are compiled
    in parallel, linked into a temporary file and moved into place atomically; a file lock serialises concurrent
    builders (one process per GPU may import the package at the same time)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("molvoxel_b200: nvcc not found and libmolvoxel_b200.so is missing/stale")
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():   # another process built it while this one waited
            return SO_PATH
        # MVX_NVCC_FLAGS: extra flags of an experimental build (e.g. -DMVX_WITH_WS together with MVX_SO=<other path>)
        flags = NVCC_FLAGS + list(extra_flags) + os.environ.get("MVX_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else [])
        jobs = [([nvcc] + flags + ["-c", API_SOURCE, "-o", os.path.join(OBJ_DIR, "mvx_api.o")])]
        for mode, ch in INSTANCES:
            for binary in (0, 1):
                obj = os.path.join(OBJ_DIR, f"mvx_vox_m{mode}_c{ch}_b{binary}.o")
                jobs.append([nvcc] + flags + [f"-DMVX_INST_MODE={mode}", f"-DMVX_INST_CH={ch}", f"-DMVX_INST_BINARY={binary}",
                                              "-c", INST_SOURCE, "-o", obj])
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
            logs = list(pool.map(_run, jobs))
        tmp = SO_PATH + f".tmp{os.getpid()}"
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + [j[-1] for j in jobs])
        os.replace(tmp, SO_PATH)
        if verbose:
            print("\n".join(logs))
    return SO_PATH


_lib = None


def lib():
    """The loaded C-ABI library (built on first use when nvcc is present)."""
    global _lib
    if _lib is None:
        if _stale():
            build()
        L = ctypes.CDLL(SO_PATH)
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.mvx_version.restype = ctypes.c_int
        L.mvx_last_error.restype = ctypes.c_char_p
        L.mvx_workspace_bytes.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch), ctypes.POINTER(sz)]
        L.mvx_host_staging_bytes.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch), ctypes.POINTER(sz)]
        L.mvx_voxelize.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch), vp, vp, sz, vp]
        L.mvx_voxelize_host.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch), vp, vp, sz, vp]
        L.mvx_voxelize_split.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch), vp, vp, sz, vp, vp, vp]
        L.mvx_voxelize_split.restype = ctypes.c_int
        L.mvx_check_status.argtypes = [vp, vp]
        L.mvx_launches_per_call.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch)]
        L.mvx_voxelize_form.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch)]
        L.mvx_random_transforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, vp, vp]
        L.mvx_random_transforms.restype = ctypes.c_int
        L.mvx_synth_ligands.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_double, vp, vp, vp, ctypes.c_int32, vp, vp]
        L.mvx_synth_ligands.restype = ctypes.c_int
        L.mvx_compact_bricks.argtypes = [ctypes.POINTER(GridSpec), ctypes.POINTER(Batch), vp, vp, vp, vp, ctypes.c_uint32, vp, vp]
        L.mvx_compact_bricks.restype = ctypes.c_int
        L.mvx_profile_begin.argtypes = [ctypes.c_int]
        dp = ctypes.POINTER(ctypes.c_double)
        L.mvx_profile_end.argtypes = [dp, dp, dp, ctypes.POINTER(ctypes.c_int)]
        L.mvx_profile_begin.restype = L.mvx_profile_end.restype = ctypes.c_int
        for fn in (L.mvx_workspace_bytes, L.mvx_host_staging_bytes, L.mvx_voxelize, L.mvx_voxelize_host,
                   L.mvx_check_status, L.mvx_launches_per_call, L.mvx_voxelize_form):
            fn.restype = ctypes.c_int
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "mvx_version", "mvx_last_error", "mvx_workspace_bytes", "mvx_host_staging_bytes", "mvx_voxelize",
    "mvx_voxelize_host", "mvx_check_status", "mvx_launches_per_call", "mvx_voxelize_form", "mvx_profile_begin",
    "mvx_profile_end", "mvx_random_transforms", "mvx_synth_ligands", "mvx_compact_bricks", "mvx_voxelize_split",
]


def raise_for_status(rc: int):
    """Map a C-ABI status to the reference's exception conventions."""
    if rc == MVX_OK:
        return
    msg = lib().mvx_last_error().decode("utf-8", "replace")
    if rc in (MVX_ERR_BAD_SHAPE, MVX_ERR_UNSUPPORTED):
        raise AssertionError(msg)          # the reference raises AssertionError from bare asserts
    if rc in (MVX_ERR_BAD_ENUM, MVX_ERR_DEVICE_FLAG, MVX_ERR_NULL_POINTER):
        raise ValueError(msg)
    raise RuntimeError(f"molvoxel_b200 (status {rc}): {msg}")
