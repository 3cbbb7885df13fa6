"""CPU-only tests: host logic, the C-ABI library's exports, sharding arithmetic (gloo, world_size 2)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import molvoxel_b200 as mv
from molvoxel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "molvoxel_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mvx_[a-z_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTED_SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.mvx_version() == 200


def _batch(mode="types", B=2, N=10, C=4, out_channels=4, radius=1.0):
    b = _lib.Batch()
    b.mode, b.num_mols, b.total_atoms = _lib.MODE[mode], B, N
    b.num_channels, b.out_channels, b.radius = C, out_channels, radius
    dummy = ctypes.c_void_p(256)
    b.mol_offsets, b.coords, b.types, b.features = dummy, dummy, dummy, dummy
    return b


def test_workspace_planning_is_host_only():
    L = _lib.lib()
    spec = _lib.GridSpec(0.5, 64, 0, 0.5, 0, 8)
    b = _batch(N=100_000, B=2048)
    need = ctypes.c_size_t(0)
    assert L.mvx_workspace_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(need)) == 0
    # 40 B record + 4 B column range + 4 columns x 4 B list entries per atom, 8 B per (molecule, column) bin
    assert need.value >= 100_000 * (40 + 4 + 16 + 4 * 48) + 2048 * 64 * 8
    assert L.mvx_launches_per_call(ctypes.byref(spec), ctypes.byref(b)) == 3   # prep, bin + expand (fused), voxelize
    assert L.mvx_voxelize_form(ctypes.byref(spec), ctypes.byref(b)) == 1        # ligand batch: warp-cell form


def test_kernel_form_follows_atom_density(monkeypatch):
    """Dense batches (>= 64 expected atoms per 8x8 column) take the pipelined persistent form with its layered
    counting-sort binning (prep, scan, place, build, voxelize, overflow sweep); D % 4 != 0 takes the generic form."""
    monkeypatch.delenv("MVX_KERNEL", raising=False)
    L = _lib.lib()
    form = lambda dim, N, B, mode="features", C=16: L.mvx_voxelize_form(   # noqa: E731
        ctypes.byref(_lib.GridSpec(0.5, dim, 0, 0.5, 0, 8)), ctypes.byref(_batch(mode, B, N, C, C)))
    assert form(48, 256 * 2000, 256) == 4
    assert form(48, 256 * 300, 256) == 1
    assert form(64, 1024 * 50, 1024, "types", 9) == 1
    assert form(50, 8 * 3000, 8) == 0
    spec, b = _lib.GridSpec(0.5, 48, 0, 0.5, 0, 8), _batch("features", 256, 256 * 2000, 16, 16)
    assert L.mvx_launches_per_call(ctypes.byref(spec), ctypes.byref(b)) == 6
    monkeypatch.setenv("MVX_KERNEL", "tiles")
    assert L.mvx_voxelize_form(ctypes.byref(spec), ctypes.byref(b)) == 3


@pytest.mark.parametrize("mutate, code", [
    (lambda s, b: setattr(s, "dimension", 0), _lib.MVX_ERR_BAD_SHAPE),
    (lambda s, b: setattr(s, "density_type", 5), _lib.MVX_ERR_BAD_ENUM),
    (lambda s, b: setattr(b, "out_channels", 2), _lib.MVX_ERR_BAD_SHAPE),
    (lambda s, b: setattr(b, "radius", 0.0), _lib.MVX_ERR_BAD_SHAPE),
    (lambda s, b: setattr(b, "coords", None), _lib.MVX_ERR_NULL_POINTER),
    (lambda s, b: setattr(b, "out_layout", 2), _lib.MVX_ERR_BAD_ENUM),
])
def test_c_abi_argument_errors(mutate, code):
    L = _lib.lib()
    spec = _lib.GridSpec(0.5, 64, 0, 0.5, 0, 8)
    b = _batch()
    mutate(spec, b)
    need = ctypes.c_size_t(0)
    assert L.mvx_workspace_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(need)) == code
    assert len(L.mvx_last_error()) > 0


def test_single_rejects_channel_wise_like_reference():
    L = _lib.lib()
    spec = _lib.GridSpec(0.5, 64, 0, 0.5, _lib.RADII["channel-wise"], 8)
    b = _batch(mode="single", C=1, out_channels=1)
    need = ctypes.c_size_t(0)
    assert L.mvx_workspace_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(need)) == _lib.MVX_ERR_UNSUPPORTED
    with pytest.raises(AssertionError):
        _lib.raise_for_status(_lib.MVX_ERR_UNSUPPORTED)


def test_voxelizer_contract_surface():
    vox = mv.create_voxelizer(0.5, 48, "atom-wise", "binary", library="b200", blockdim=16)
    assert vox.LIB == "B200" and vox.dimension == 48 and vox.resolution == 0.5
    assert vox.width == 0.5 * 47 and vox.upper_bound == -vox.lower_bound == 0.5 * 47 / 2
    assert vox.grid_dimension(5) == (5, 48, 48, 48) and vox.spatial_dimension == (48, 48, 48)
    assert vox.is_radii_type_atom_wise and vox.is_density_type_binary and not vox.is_density_type_gaussian
    vox.radii_type = "scalar"
    assert vox.is_radii_type_scalar
    with pytest.raises(AssertionError):
        vox.radii_type = "nope"
    with pytest.raises(AssertionError):
        mv.create_voxelizer(library="cupy")
    with pytest.raises(ValueError):
        vox.asarray([1, 2, 3], "bogus") if torch.cuda.is_available() else (_ for _ in ()).throw(ValueError())


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    vox = mv.create_voxelizer(0.5, 16, library="b200")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vox.forward_single(np.zeros((1, 3)), None, 1.0)
    with pytest.raises(RuntimeError):
        vox.cpu()


def test_channels_last_argument_handling_is_host_side():
    """The output layout is an argument of the C ABI (mvx_batch.out_layout); the Python mirror derives it from the out
    tensor's strides and accepts exactly the two dense layouts."""
    L = _lib.lib()
    spec, b = _lib.GridSpec(0.5, 64, 0, 0.5, 0, 8), _batch()
    need_a, need_b = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert L.mvx_workspace_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(need_a)) == 0
    b.out_layout = _lib.LAYOUT_DHWC
    assert L.mvx_workspace_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(need_b)) == 0
    assert need_a.value == need_b.value > 0
    assert L.mvx_compact_bricks(ctypes.byref(spec), ctypes.byref(b), None, None, None, None, 0, None, None) == _lib.MVX_ERR_UNSUPPORTED
    vox = mv.create_voxelizer(0.5, 16, library="b200", device="cpu", channels_last=True)
    g = vox.get_empty_grid(5, 3, init_zero=True)
    assert tuple(g.shape) == (3, 5, 16, 16, 16) and g.is_contiguous(memory_format=torch.channels_last_3d)
    assert tuple(vox.get_empty_grid(5).shape) == (5, 16, 16, 16) and vox.get_empty_grid(5).permute(1, 2, 3, 0).is_contiguous()
    coords, offs, types = np.zeros((4, 3)), np.array([0, 2, 3, 4], dtype=np.int32), np.zeros(4, dtype=np.int32)
    with pytest.raises(AssertionError, match="contiguous"):   # neither (B,C,D,H,W)- nor (B,D,H,W,C)-contiguous
        vox.forward_types_batch(coords, offs, None, types, 1.0, 5, out=torch.zeros(3, 16, 5, 16, 16).permute(0, 2, 1, 3, 4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):   # a channels-last out passes the checks
            vox.forward_types_batch(coords, offs, None, types, 1.0, 5, out=g)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "molvoxel_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("numpy oracle", ""), f"{f} mentions the oracle"


def test_transform_is_rigid_and_has_the_reference_shape():
    np.random.seed(123)
    t = mv.create_random_transform(0.5, True).get_transform()
    assert t.translation.shape == (1, 3) and t.translation.dtype == np.float32 and np.abs(t.translation).max() <= 0.5
    assert abs(sum(v * v for v in t.quaternion) - 1.0) < 1e-12
    xyz = np.random.default_rng(0).normal(size=(7, 3))
    c = xyz.mean(0)
    out = t(xyz, c)
    # rotation about the centre, then the translation twice (numpy backend, numpy/transform.py:56-59)
    assert np.allclose(np.linalg.norm(out - c - 2 * t.translation, axis=1), np.linalg.norm(xyz - c, axis=1))
    assert t.as_row().shape == (7,) and np.array_equal(t.as_row()[4:], t.translation.reshape(3).astype(np.float64))


def test_shard_bounds_cover_and_partition():
    for n in (0, 1, 7, 1024, 1_000_000):
        for w in (1, 2, 4, 8):
            spans = [mv.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    offs = np.array([0, 3, 3, 8, 10, 15], dtype=np.int32)
    coords = np.arange(45.0).reshape(15, 3)
    local, (c,), (z,) = mv.shard_batch(offs, 1, 2, coords, per_mol=(np.arange(5),))
    assert local.tolist() == [0, 2, 7] and c.shape == (7, 3) and c[0, 0] == 24.0 and z.tolist() == [3, 4]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5
    lo, hi = mv.shard_bounds(n, rank, world)
    local = torch.full((hi - lo, 2, 3, 3, 3), 0.0)
    for i in range(lo, hi):
        local[i - lo] = float(i + 1)
    full = mv.gather_grids(local, n)
    ok = full.shape[0] == n and all(float(full[i].mean()) == i + 1 for i in range(n))
    # throughput aggregation used by bench.py: max over ranks of the elapsed time
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, ok, float(t)))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def _cpu_vox(**kw):
    return mv.create_voxelizer(0.5, 16, library="b200", **kw)


@pytest.mark.skipif(torch.cuda.is_available(), reason="argument checks are reached before the CUDA requirement")
@pytest.mark.parametrize("call, msg", [
    (lambda: _cpu_vox().forward_types(np.zeros((4, 3)), None, np.zeros(3, dtype=np.int16), 1.0),
     "types does not match dimension"),
    (lambda: _cpu_vox().forward_types(np.zeros((4, 3)), None, np.zeros(4, dtype=np.int16), np.ones(4, dtype=np.float32)),
     "radii should be scalar"),
    (lambda: _cpu_vox(radii_type="atom-wise").forward_types(np.zeros((4, 3)), None, np.zeros(4, dtype=np.int16), 1.0),
     "radii should be Array"),
    (lambda: _cpu_vox(radii_type="atom-wise").forward_single(np.zeros((4, 3)), None, np.ones(3, dtype=np.float32)),
     "radii does not match dimension (number of atoms,)"),
    (lambda: _cpu_vox(radii_type="channel-wise").forward_features(np.zeros((4, 3)), None, np.zeros((4, 5), dtype=np.float32),
                                                                np.ones(4, dtype=np.float32)),
     "radii does not match dimension (number of channels,)"),
    (lambda: _cpu_vox(radii_type="channel-wise").forward_single(np.zeros((4, 3)), None, np.ones(1, dtype=np.float32)),
     "Channel-Wise Radii Type is not supported"),
    (lambda: _cpu_vox().forward_features(np.zeros((4, 3)), None, np.zeros((3, 5), dtype=np.float32), 1.0),
     "atom features does not match number of atoms"),
    (lambda: _cpu_vox().forward_types(np.zeros((4, 3)), None, np.array([0, 1, 2, 3]), 1.0, out_grid=torch.zeros(2, 16, 16, 16)),
     "Output channel is less than number of types"),
    (lambda: _cpu_vox().forward_features(np.zeros((4, 3)), None, np.zeros((4, 5), dtype=np.float32), 1.0,
                                         out_grid=torch.zeros(4, 16, 16, 16)),
     "Output grid dimension incorrect"),
    (lambda: _cpu_vox().forward_single(np.zeros((4, 3)), None, 1.0, out_grid=torch.zeros(2, 16, 16, 16)),
     "Output channel should be 1"),
])
def test_reference_argument_checks_and_messages(call, msg):
    """Same AssertionError conditions and messages as reference numpy/voxelizer.py:171-192, :317-342, :438-455."""
    with pytest.raises(AssertionError) as e:
        call()
    assert msg in str(e.value)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_valid_arguments_then_no_cpu_fallback_and_dispatch():
    vox = _cpu_vox()
    with pytest.raises(ValueError):   # np.max of an empty array, like the reference (numpy/voxelizer.py:325)
        vox.forward_types(np.zeros((0, 3)), None, np.zeros((0,), dtype=np.int16), 1.0)
    for channels in (None, np.zeros(4, dtype=np.int16), np.zeros((4, 2), dtype=np.float32)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):   # forward() dispatches on channels' rank
            vox(np.zeros((4, 3)), np.zeros(3), channels, 1.0)
