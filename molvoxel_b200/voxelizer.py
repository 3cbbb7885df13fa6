"""Host-side mirror of the reference Voxelizer interface for the `library="b200"` backend.

Mirrors reference molvoxel/voxelizer/base/voxelizer.py:9-176 (contract) and the torch flavour's
device handling (molvoxel/voxelizer/torch/voxelizer.py:59-88, :569-582); argument checks repeat the
numpy backend's asserts (molvoxel/voxelizer/numpy/voxelizer.py:171-192, :317-342, :438-455) with the
same messages.  All compute happens in libmolvoxel_b200.so (hand-written sm_100a kernels) through
the C ABI in include/molvoxel_b200.h; torch is used for device memory and streams only.  There is
no CPU path: without a CUDA device every forward_* raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib
from .sparse import SparseGrids
from .transform import RandomTransform, host_transform_rows, transform_rows


def _norm(x):
    """Lists/tuples become numpy arrays; tensors, arrays, scalars and None pass through."""
    if x is None or isinstance(x, (torch.Tensor, np.ndarray)) or np.isscalar(x):
        return x
    return np.asarray(x)


def _is_scalar(x) -> bool:
    if isinstance(x, torch.Tensor):
        return x.ndim == 0
    return np.isscalar(x) or (isinstance(x, np.ndarray) and x.ndim == 0)


class Voxelizer:
    LIB = "B200"
    transform_class = RandomTransform
    RADII_TYPE_LIST = ["scalar", "channel-wise", "atom-wise"]
    DENSITY_TYPE_LIST = ["gaussian", "binary"]

    def __init__(
        self,
        resolution: float = 0.5,
        dimension: int = 64,
        radii_type: str = "scalar",
        density_type: str = "gaussian",
        device: str | torch.device = "cuda",
        blockdim: int | None = None,
        out_dtype: torch.dtype = torch.float32,
        **kwargs,
    ):
        assert radii_type in self.RADII_TYPE_LIST
        assert density_type in self.DENSITY_TYPE_LIST
        self._resolution = resolution
        self._dimension = dimension
        self._width = resolution * (dimension - 1)
        self._radii_type = radii_type
        self._density_type = density_type
        self.upper_bound = self.width / 2.0
        self.lower_bound = -1 * self.upper_bound
        self._spatial_dimension = (dimension, dimension, dimension)
        self._sigma = kwargs.get("sigma", 0.5)
        # `blockdim` keeps the reference meaning: the block size whose half-voxel cull the result must
        # reproduce (numpy default 8).  blockdim >= dimension gives the exact mathematics.
        self.blockdim = 8 if blockdim is None else int(blockdim)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        # fp32 is the reference (precision=32).  bfloat16 / float16 grids are computed in fp32 exactly like the
        # fp32 grid and rounded once on the store: half the bytes of the write-bound op.
        # precision=64 (reference numpy/voxelizer.py:28-34) = a float64 grid computed in fp64 (untuned).
        precision = kwargs.get("precision", 32)
        assert precision in [32, 64]
        if precision == 64:
            out_dtype = torch.float64
        assert out_dtype in (torch.float32, torch.bfloat16, torch.float16, torch.float64), \
            "out_dtype must be float32, bfloat16, float16 or float64"
        self.out_dtype = out_dtype
        # channels_last=True: grids are allocated in the channels-last memory format (torch.channels_last_3d: logical shape
        # (B, C, D, H, W), physical (B, D, H, W, C) — the layout the reference README writes its formulas in, README.md:138-142,
        # and the one a channels-last 3-D CNN reads without a transpose).  Same values bit for bit.  An
        # `out=` / `out_grid=` tensor decides by its own strides, whatever this default says.
        self.channels_last = bool(kwargs.get("channels_last", False))
        # random rigid transform (random_translation / random_rotation of every forward_*): where the per-molecule
        # parameters come from.  "philox" (default): drawn on the device inside the prep kernel from a counter-based
        # generator keyed by (seed, molecule index) — molecule indices continue across calls, `rng_offset=` pins them
        # for sharded sweeps.  "numpy": drawn on the host from numpy's global RNG in the reference's order, so
        # np.random.seed(s) reproduces the reference's augmentation bit for bit (a Python loop per molecule: slow).
        self.rng = kwargs.get("rng", "philox")
        assert self.rng in ("philox", "numpy"), "rng must be 'philox' or 'numpy'"
        self._seed = int(kwargs.get("seed", 0)) & 0xFFFFFFFFFFFFFFFF
        self._mol_counter = 0
        # the numpy / numba backends add the translation twice when a rotation is also requested
        # (numpy/transform.py:56-59); translate_once=True gives the torch backend's single addition
        self.translate_once = bool(kwargs.get("translate_once", False))
        # overlap=True (default): host batches passed with non_blocking=True run their prep / binning on a second stream next to
        # the previous call's voxelize kernel where that pays (ligand sweeps); device inputs opt in with inputs_ready=
        self.overlap = bool(kwargs.get("overlap", True))
        self._overlap = None
        self._last_ws = None
        self._ws = None
        self._ws_need = {}
        self._pipe = None
        self.last_copy_event = None
        self._sticky_flags = 0
        _lib.lib()   # fail loudly here if the CUDA library cannot be built/loaded

    # ---- properties of the reference contract (base/voxelizer.py:40-97) ----
    @property
    def radii_type(self) -> str:
        return self._radii_type

    @radii_type.setter
    def radii_type(self, radii_type: str):
        assert radii_type in self.RADII_TYPE_LIST
        self._radii_type = radii_type

    @property
    def is_radii_type_scalar(self):
        return self._radii_type == "scalar"

    @property
    def is_radii_type_channel_wise(self):
        return self._radii_type == "channel-wise"

    @property
    def is_radii_type_atom_wise(self):
        return self._radii_type == "atom-wise"

    @property
    def density_type(self) -> str:
        return self._density_type

    @density_type.setter
    def density_type(self, density_type: str):
        assert density_type in self.DENSITY_TYPE_LIST
        self._density_type = density_type
        if density_type == "gaussian":
            self._sigma = 0.5   # the reference setter cannot receive sigma either (base/voxelizer.py:65-70)

    @property
    def is_density_type_binary(self):
        return self._density_type == "binary"

    @property
    def is_density_type_gaussian(self):
        return self._density_type == "gaussian"

    def grid_dimension(self, num_channels: int):
        return (num_channels, self._dimension, self._dimension, self._dimension)

    @property
    def spatial_dimension(self):
        return self._spatial_dimension

    @property
    def resolution(self) -> float:
        return self._resolution

    @property
    def dimension(self) -> int:
        return self._dimension

    @property
    def width(self) -> float:
        return self._width

    # ---- device handling (torch/voxelizer.py:73-88; returns self even when unchanged) ----
    def to(self, device, update_blockdim: bool = True, blockdim: int | None = None):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("molvoxel_b200 has no CPU path; device must be a CUDA device")
        if device.index is None and torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
        if device != self.device:
            self.device = device
            self._ws = None
            self._overlap = None
            self._pipe = None
        if blockdim is not None:
            self.blockdim = int(blockdim)
        return self

    def cuda(self, update_blockdim: bool = True, blockdim: int | None = None):
        return self.to("cuda", update_blockdim, blockdim)

    def cpu(self, *args, **kwargs):
        raise RuntimeError("molvoxel_b200 has no CPU path")

    def get_empty_grid(self, num_channels: int, batch_size: int | None = None, init_zero: bool = False):
        shape = self.grid_dimension(num_channels)
        if batch_size is not None:
            shape = (batch_size,) + shape
        fn = torch.zeros if init_zero else torch.empty
        if self.channels_last:   # physical (.., D, H, W, C), logical (.., C, D, H, W)
            g = fn(shape[:-4] + shape[-3:] + (num_channels,), dtype=self.out_dtype, device=self.device)
            return g.permute(0, 4, 1, 2, 3) if batch_size is not None else g.permute(3, 0, 1, 2)
        return fn(shape, dtype=self.out_dtype, device=self.device)

    def asarray(self, array, obj: str):
        """torch/voxelizer.py:569-582; coords/center keep fp64 (the numpy oracle's centring precision)."""
        if obj in ("coords", "center"):
            dt = torch.float64
        elif obj in ("features", "radii"):
            dt = torch.float32
        elif obj == "types":
            dt = torch.int32
        else:
            raise ValueError("obj should be ['coords', 'center', 'radii', types', 'features']")
        if isinstance(array, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(array)).to(self.device, dt)
        return torch.as_tensor(array).to(self.device, dt)

    # ---- reference forward API ----
    def forward(self, coords, center, channels, radii, random_translation: float = 0.0,
                random_rotation: bool = False, out_grid=None):
        if channels is None:
            return self.forward_single(coords, center, radii, random_translation, random_rotation, out_grid)
        elif np.ndim(channels) == 1 if not isinstance(channels, torch.Tensor) else channels.ndim == 1:
            return self.forward_types(coords, center, channels, radii, random_translation, random_rotation, out_grid)
        else:
            return self.forward_features(coords, center, channels, radii, random_translation, random_rotation, out_grid)

    __call__ = forward

    def forward_types(self, coords, center, types, radii, random_translation: float = 0.0,
                      random_rotation: bool = False, out_grid=None):
        """(V,3), (3,)|None, (V,), scalar|(C,)|(V,) -> (C,D,H,W); numpy/voxelizer.py:240-315."""
        coords, center, types, radii = _norm(coords), _norm(center), _norm(types), _norm(radii)
        V = int(coords.shape[0])
        offs = [0, V]
        centers = None if center is None else center.reshape(1, 3)
        if self.is_radii_type_channel_wise and not _is_scalar(radii):
            C = int(radii.shape[0])
            Cmax = None
        else:
            C = None
            Cmax = True
        out = self._forward_batch("types", coords, offs, centers, types, radii, C, random_translation,
                                  random_rotation, None if out_grid is None else out_grid.unsqueeze(0),
                                  infer_types_channels=Cmax is True)
        return out_grid if out_grid is not None else out[0]

    def forward_features(self, coords, center, features, radii, random_translation: float = 0.0,
                         random_rotation: bool = False, out_grid=None):
        """(V,3), (3,)|None, (V,C), scalar|(C,)|(V,) -> (C,D,H,W); numpy/voxelizer.py:97-169."""
        coords, center, features, radii = _norm(coords), _norm(center), _norm(features), _norm(radii)
        V = int(coords.shape[0])
        centers = None if center is None else center.reshape(1, 3)
        assert features.ndim == 2, f"atom features does not match dimension: {tuple(features.shape)} vs {(V, '*')}"
        out = self._forward_batch("features", coords, [0, V], centers, features, radii, int(features.shape[1]),
                                  random_translation, random_rotation,
                                  None if out_grid is None else out_grid.unsqueeze(0))
        return out_grid if out_grid is not None else out[0]

    def forward_single(self, coords, center, radii, random_translation: float = 0.0,
                       random_rotation: bool = False, out_grid=None):
        """(V,3), (3,)|None, scalar|(V,) -> (1,D,H,W); numpy/voxelizer.py:370-436."""
        coords, center, radii = _norm(coords), _norm(center), _norm(radii)
        V = int(coords.shape[0])
        centers = None if center is None else center.reshape(1, 3)
        out = self._forward_batch("single", coords, [0, V], centers, None, radii, 1, random_translation,
                                  random_rotation, None if out_grid is None else out_grid.unsqueeze(0))
        return out_grid if out_grid is not None else out[0]

    # ---- batched driver (new, additive; per-molecule semantics = B independent reference calls) ----
    def forward_types_batch(self, coords, mol_offsets, centers, types, radii, num_channels,
                            random_translation: float = 0.0, random_rotation: bool = False, out=None,
                            non_blocking: bool = False, max_radius=None, transforms=None, rng_offset=None,
                            inputs_ready=None):
        """CSR batch -> (B, C, D, H, W).  non_blocking=True with HOST inputs pipelines the H2D copies of this
        call behind the kernels of the previous one (pinned inputs; call check_status() to synchronise).
        max_radius: a host-known bound of an array `radii` (saves a device->host read on the device path).
        transforms: explicit rigid transforms ((B, 7) rows, a T, or a list of T) instead of random draws.
        rng_offset: global index of the batch's first molecule for the device generator (sharded / chunked sweeps).
        inputs_ready: DEVICE inputs only — True, or the torch.cuda.Event that completes them: the inputs do not depend on
        work pending on the current stream, so this call's prep / binning may overlap the previous call's voxelize kernel."""
        return self._forward_batch("types", coords, mol_offsets, centers, types, radii, int(num_channels),
                                   random_translation, random_rotation, out, non_blocking=non_blocking,
                                   max_radius=max_radius, transforms=transforms, rng_offset=rng_offset,
                                   inputs_ready=inputs_ready)

    def forward_features_batch(self, coords, mol_offsets, centers, features, radii,
                               random_translation: float = 0.0, random_rotation: bool = False, out=None,
                               non_blocking: bool = False, max_radius=None, transforms=None, rng_offset=None):
        return self._forward_batch("features", coords, mol_offsets, centers, features, radii,
                                   int(features.shape[1]), random_translation, random_rotation, out,
                                   non_blocking=non_blocking, max_radius=max_radius, transforms=transforms,
                                   rng_offset=rng_offset)

    def forward_single_batch(self, coords, mol_offsets, centers, radii,
                             random_translation: float = 0.0, random_rotation: bool = False, out=None,
                             non_blocking: bool = False, max_radius=None, transforms=None, rng_offset=None):
        return self._forward_batch("single", coords, mol_offsets, centers, None, radii, 1,
                                   random_translation, random_rotation, out, non_blocking=non_blocking,
                                   max_radius=max_radius, transforms=transforms, rng_offset=rng_offset)

    def seed(self, seed: int, first_molecule: int = 0):
        """Key of the device generator and the molecule index the next call starts at."""
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._mol_counter = int(first_molecule)

    def random_transforms(self, num_mols: int, random_translation: float = 0.0, random_rotation: bool = False,
                          rng_offset: int = 0) -> torch.Tensor:
        """The (B, 7) rows (quaternion, translation) the device generator gives molecules
        [rng_offset, rng_offset + num_mols) under the current seed — what a forward_* call with the same arguments
        applies.  Passing them back as `transforms=` gives bit-identical grids."""
        flags = self._transform_flags(random_translation, random_rotation)
        out = torch.zeros((num_mols, 7), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.raise_for_status(_lib.lib().mvx_random_transforms(
                ctypes.c_uint64(self._seed), ctypes.c_uint64(int(rng_offset)), int(num_mols), flags,
                float(random_translation or 0.0), ctypes.c_void_p(out.data_ptr()), stream))
        return out

    def _transform_flags(self, random_translation, random_rotation) -> int:
        flags = 0
        if random_rotation:
            flags |= _lib.TF_ROTATE
        if random_translation is not None and random_translation > 0.0:
            flags |= _lib.TF_TRANSLATE
        if flags and self.translate_once:
            flags |= _lib.TF_TRANSLATE_ONCE
        return flags

    # ---- implementation ----
    def _spec(self):
        """mvx_grid_spec of the current settings (rebuilt only when one of them changed: the setters are plain attributes)."""
        key = (self._resolution, self._dimension, self._density_type, self._sigma, self._radii_type, self.blockdim)
        cached = getattr(self, "_spec_cache", None)
        if cached is None or cached[0] != key:
            cached = (key, _lib.GridSpec(float(self._resolution), int(self._dimension), _lib.DENSITY[self._density_type],
                                         float(self._sigma), _lib.RADII[self._radii_type], int(self.blockdim)))
            self._spec_cache = cached
        return cached[1]

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check_radii(self, mode, radii, V, C):
        if mode == "single":
            assert not self.is_radii_type_channel_wise, "Channel-Wise Radii Type is not supported"
        if self.is_radii_type_scalar:
            assert _is_scalar(radii), "the radii type of voxelizer is `scalar`, radii should be scalar"
        elif self.is_radii_type_channel_wise:
            assert not _is_scalar(radii), f"the radii type of voxelizer is `channel-wise`, radii should be Array[{C},]"
            assert tuple(radii.shape) == (C,), \
                f"radii does not match dimension (number of channels,): {tuple(radii.shape)} vs {(C,)}"
        else:
            assert not _is_scalar(radii), f"the radii type of voxelizer is `atom-wise`, radii should be Array[{V},]"
            assert tuple(radii.shape) == (V,), \
                f"radii does not match dimension (number of atoms,): {tuple(radii.shape)} vs {(V,)}"

    def _upload_pipelined(self, arrays):
        """H2D of one call's host inputs on the copy stream into this call's slot of a 2-deep ring of device
        staging tensors; the compute stream waits on the copy, and the slot is reused only after the kernels
        that read it have finished.  Returns device tensors in the order given."""
        if self._pipe is None:
            self._pipe = {"copy": torch.cuda.Stream(self.device), "idx": 0,
                          "slots": [{"bufs": {}, "copied": torch.cuda.Event(), "done": torch.cuda.Event()} for _ in range(2)],
                          "status": torch.zeros(2, dtype=torch.int32).pin_memory()}
        pipe = self._pipe
        slot = pipe["slots"][pipe["idx"]]
        slot["index"] = pipe["idx"]
        pipe["idx"] ^= 1
        slot["done"].synchronize()
        # the status word of the call that last used this slot has landed: keep its flags (sticky until check_status)
        self._sticky_flags |= int(pipe["status"][slot["index"]])
        pipe["status"][slot["index"]] = 0
        cur = torch.cuda.current_stream(self.device)
        outs = []
        with torch.cuda.stream(pipe["copy"]):
            for name, a in arrays:
                if a is None:
                    outs.append(None)
                    continue
                src = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
                buf = slot["bufs"].get(name)
                if buf is None or buf.dtype != src.dtype or buf.numel() < src.numel():
                    buf = torch.empty(max(src.numel(), 1), dtype=src.dtype, device=self.device)
                    slot["bufs"][name] = buf
                dst = buf[:src.numel()].view(src.shape)
                dst.copy_(src, non_blocking=True)
                outs.append(dst)
            slot["copied"].record(pipe["copy"])
        cur.wait_event(slot["copied"])
        self.last_copy_event = slot["copied"]   # host buffers of this call may be refilled once it has completed
        return outs, slot

    def _forward_batch(self, mode, coords, mol_offsets, centers, channels, radii, C, random_translation,
                       random_rotation, out, infer_types_channels=False, max_radius=None, non_blocking=False,
                       transforms=None, rng_offset=None, inputs_ready=None):
        have_cuda = self.device.type == "cuda" and torch.cuda.is_available()
        coords, centers, channels, radii = _norm(coords), _norm(centers), _norm(channels), _norm(radii)
        on_device = isinstance(coords, torch.Tensor) and coords.is_cuda
        if non_blocking and not on_device and have_cuda:
            # pipelined host path: async H2D on the copy stream, then the no-sync device path
            if not _is_scalar(radii) and max_radius is None:
                max_radius = float(np.asarray(radii).max()) if np.asarray(radii).size else 1.0
            if mode == "types" and C is None:
                C = int(np.asarray(channels).max()) + 1
            (coords_d, offs_d, centers_d, chan_d, radii_d), slot = self._upload_pipelined([
                ("coords", coords), ("offs", np.asarray(mol_offsets, dtype=np.int32) if not isinstance(mol_offsets, torch.Tensor) else mol_offsets),
                ("centers", centers), ("chan", channels), ("radii", None if _is_scalar(radii) else radii)])
            # the copy's own event says when the inputs are complete: prep / binning of this call may then run on the
            # binning stream next to the previous call's voxelize kernel (mvx_voxelize_split) where that applies
            res = self._forward_batch(mode, coords_d, offs_d, centers_d, chan_d, radii if _is_scalar(radii) else radii_d, C,
                                      random_translation, random_rotation, out, False, max_radius,
                                      transforms=transforms, rng_offset=rng_offset,
                                      inputs_ready=slot["copied"] if self.overlap else None)
            cur = torch.cuda.current_stream(self.device)
            wst = self._last_ws
            ws_ptr_off = (-wst.data_ptr()) % 256
            self._pipe["status"][slot["index"]:slot["index"] + 1].copy_(
                wst[ws_ptr_off:ws_ptr_off + 4].view(torch.int32), non_blocking=True)
            slot["done"].record(cur)
            return res
        D = self._dimension
        N = int(coords.shape[0])
        assert coords.ndim == 2 and coords.shape[1] == 3, f"coords should be (V, 3): {tuple(coords.shape)}"

        # molecule offsets
        if isinstance(mol_offsets, torch.Tensor):
            B = int(mol_offsets.shape[0]) - 1
        else:
            mol_offsets = np.ascontiguousarray(np.asarray(mol_offsets), dtype=np.int32)
            B = int(mol_offsets.shape[0]) - 1
            assert int(mol_offsets[0]) == 0 and int(mol_offsets[-1]) == N, "mol_offsets must span [0, N]"

        # channel bookkeeping + the reference's argument checks
        if mode == "types":
            assert tuple(channels.shape) == (N,), f"types does not match dimension: {tuple(channels.shape)} vs {(N,)}"
            if infer_types_channels:
                if out is not None and on_device:
                    C = int(out.shape[1])        # no host sync: the device validates types < C
                else:
                    tmax = channels.max()        # V == 0 raises like np.max on an empty array
                    C = int(tmax) + 1
            self._check_radii(mode, radii, N, C)
        elif mode == "features":
            assert channels.shape[0] == N, f"atom features does not match number of atoms: {channels.shape[0]} vs {N}"
            self._check_radii(mode, radii, N, C)
        else:
            self._check_radii(mode, radii, N, C)

        clast = self.channels_last
        if out is not None:
            assert isinstance(out, torch.Tensor) and out.ndim == 5, "out_grid must be a (C, D, H, W) / (B, C, D, H, W) tensor"
            clast = not out.is_contiguous() and out.permute(0, 2, 3, 4, 1).is_contiguous()
            assert out.dtype == self.out_dtype and (out.is_contiguous() or clast) and \
                (out.is_cuda or not have_cuda), f"out_grid must be a contiguous (or channels-last) {self.out_dtype} CUDA tensor"
            if mode == "types":
                assert out.shape[1] >= C, f"Output channel is less than number of types: {out.shape[1]} < {C}"
                assert tuple(out.shape[2:]) == (D, D, D), \
                    f'Output grid dimension incorrect: {tuple(out.shape[1:])} vs {("*", D, D, D)}'
            elif mode == "features":
                assert tuple(out.shape[1:]) == (C, D, D, D), \
                    f"Output grid dimension incorrect: {tuple(out.shape[1:])} vs {(C, D, D, D)}"
            else:
                assert out.shape[1] == 1, "Output channel should be 1"
                assert tuple(out.shape[2:]) == (D, D, D), \
                    f'Output grid dimension incorrect: {tuple(out.shape[1:])} vs {("*", D, D, D)}'
            assert out.shape[0] == B
            out_channels = int(out.shape[1])
        else:
            out_channels = C
        # every argument check of the reference has passed; from here on a CUDA device is required
        if not have_cuda:
            raise RuntimeError("molvoxel_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if out is None:
            out = self.get_empty_grid(out_channels, B)

        # centring and the optional rigid transform stay inside the prep kernel, in numpy's promoted dtype
        # (numpy/voxelizer.py:263-265).  Explicit transforms, or parameters drawn on the host in the reference's
        # order (rng="numpy"), travel as (B, 7) rows; otherwise the device generator draws them.
        tf_flags = 0
        tf_rows = None
        if transforms is not None:
            tf_rows = transform_rows(transforms, B)
            # which parts of the rows apply: the call's random_* arguments if any is set, else both
            tf_flags = self._transform_flags(random_translation, random_rotation) or \
                (_lib.TF_ROTATE | _lib.TF_TRANSLATE | (_lib.TF_TRANSLATE_ONCE if self.translate_once else 0))
            if isinstance(transforms, RandomTransform.class_T) or (isinstance(transforms, (list, tuple)) and len(transforms) > 0):
                ts = [transforms] if isinstance(transforms, RandomTransform.class_T) else list(transforms)
                if all(t.quaternion is None for t in ts):   # translation only: one addition (numpy/transform.py:58-59)
                    tf_flags &= ~_lib.TF_ROTATE
                if all(t.translation is None for t in ts):
                    tf_flags &= ~_lib.TF_TRANSLATE
        else:
            tf_flags = self._transform_flags(random_translation, random_rotation)
            if tf_flags and self.rng == "numpy":
                tf_rows = host_transform_rows(B, random_translation, random_rotation)

        keep = []   # keeps converted arrays alive until the call returns

        converted = [False]   # an input had to be cast / moved / compacted: it is produced on the current stream

        def dev(t, dt):
            t2 = t.to(self.device, dt).contiguous()
            if t2 is not t and t2.data_ptr() != t.data_ptr():
                converted[0] = True
            keep.append(t2)
            return t2

        def host(a, dt=None):
            a = np.ascontiguousarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a), dtype=dt)
            keep.append(a)
            return a

        b = _lib.Batch()
        b.mode, b.num_mols, b.total_atoms = _lib.MODE[mode], B, N
        b.num_channels, b.out_channels = C, out_channels
        b.radius, b.max_radius = 0.0, 0.0
        b.out_dtype = _lib.OUT_DTYPE[str(self.out_dtype).replace("torch.", "")]
        b.transform_flags = tf_flags
        b.out_layout = _lib.LAYOUT_DHWC if clast else _lib.LAYOUT_CDHW
        if tf_flags and tf_rows is None:   # device-drawn: key + global molecule index
            b.rng_seed = self._seed
            b.random_translation = float(random_translation or 0.0)
            if rng_offset is None:
                rng_offset = self._mol_counter
                self._mol_counter += B
            b.rng_offset = int(rng_offset)
        if self.is_radii_type_scalar:   # numpy's promotion depends on how the scalar is typed (NEP 50)
            b.radius_kind = (_lib.RADIUS_NP_F64 if isinstance(radii, np.float64) else
                             _lib.RADIUS_NP_F32 if isinstance(radii, np.float32) else _lib.RADIUS_PYFLOAT)

        def fdtype_of(x):
            if isinstance(x, torch.Tensor):
                return torch.float32 if x.dtype == torch.float32 else torch.float64
            return np.float32 if x.dtype == np.float32 else np.float64

        if on_device:
            ptr = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
            offs_t = mol_offsets if isinstance(mol_offsets, torch.Tensor) else torch.from_numpy(mol_offsets)
            b.mol_offsets = ptr(dev(offs_t, torch.int32))
            cdt = fdtype_of(coords)
            b.coords, b.coords_dtype = ptr(dev(coords, cdt)), int(cdt == torch.float64)
            if centers is not None:
                centers = torch.as_tensor(centers) if not isinstance(centers, torch.Tensor) else centers
                zdt = fdtype_of(centers)
                b.centers, b.centers_dtype = ptr(dev(centers.reshape(B, 3), zdt)), int(zdt == torch.float64)
            if mode == "types":
                b.types = ptr(dev(torch.as_tensor(channels), torch.int32))
            elif mode == "features":
                ft = torch.as_tensor(channels)
                if ft.dtype in (torch.uint8, torch.float16):   # compact rows: widened exactly on the device
                    b.features, b.features_dtype = ptr(dev(ft, ft.dtype)), (_lib.MVX_U8 if ft.dtype == torch.uint8 else _lib.MVX_F16)
                else:
                    b.features = ptr(dev(ft, torch.float32))
            if self.is_radii_type_scalar:
                b.radius = float(radii)
            else:
                r = dev(torch.as_tensor(radii), torch.float32)
                b.radii = ptr(r)
                b.max_radius = float(max_radius) if max_radius is not None else (float(r.max()) if r.numel() else 1.0)
            if tf_rows is not None:
                b.transforms = ptr(dev(torch.from_numpy(tf_rows), torch.float64))
        else:
            ptr = lambda a: ctypes.c_void_p(a.ctypes.data)   # noqa: E731
            b.mol_offsets = ptr(host(mol_offsets, np.int32))
            coords = coords.detach().cpu().numpy() if isinstance(coords, torch.Tensor) else np.asarray(coords)
            cdt = fdtype_of(coords)
            b.coords, b.coords_dtype = ptr(host(coords, cdt)), int(cdt == np.float64)
            if centers is not None:
                centers = centers.detach().cpu().numpy() if isinstance(centers, torch.Tensor) else np.asarray(centers)
                zdt = fdtype_of(centers)
                b.centers, b.centers_dtype = ptr(host(centers.reshape(B, 3), zdt)), int(zdt == np.float64)
            if mode == "types":
                t = channels.detach().cpu().numpy() if isinstance(channels, torch.Tensor) else np.asarray(channels)
                b.types = ptr(host(t.astype(np.int16), np.int32))   # the reference narrows to int16 (:269)
            elif mode == "features":
                fa = channels.detach().cpu().numpy() if isinstance(channels, torch.Tensor) else np.asarray(channels)
                if fa.dtype in (np.uint8, np.float16):   # compact rows cross PCIe as they are, widened exactly on the device
                    b.features, b.features_dtype = ptr(host(fa)), (_lib.MVX_U8 if fa.dtype == np.uint8 else _lib.MVX_F16)
                else:
                    b.features = ptr(host(fa, np.float32))
            if self.is_radii_type_scalar:
                b.radius = float(radii)
            else:
                r = host(radii, np.float32)
                b.radii = ptr(r)
                b.max_radius = float(max_radius) if max_radius is not None else (float(r.max()) if r.size else 1.0)
            if tf_rows is not None:
                b.transforms = ptr(host(tf_rows, np.float64))

        L = _lib.lib()
        spec = self._spec()
        # workspace size of this call shape: cached (two C calls, each planning the batch, per forward otherwise)
        key = (mode, B, N, C, out_channels, on_device, int(b.coords_dtype), int(b.centers_dtype), int(b.features_dtype),
               int(b.out_dtype), int(b.out_layout), float(b.radius), float(b.max_radius), int(b.transform_flags), tf_rows is not None,
               centers is None, self._radii_type, self._density_type, self.blockdim, float(self._sigma))
        total = self._ws_need.get(key)
        if total is None:
            need = ctypes.c_size_t(0)
            _lib.raise_for_status(L.mvx_workspace_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(need)))
            total = need.value
            if not on_device:
                stg = ctypes.c_size_t(0)
                _lib.raise_for_status(L.mvx_host_staging_bytes(ctypes.byref(spec), ctypes.byref(b), ctypes.byref(stg)))
                total += stg.value
            if len(self._ws_need) > 256:
                self._ws_need.clear()
            self._ws_need[key] = total
        # Two-stream form (mvx_voxelize_split): the caller vouches that the device inputs are complete (`inputs_ready` =
        # True or the event that completes them), nothing had to be converted on the current stream, and the kernel is
        # the HBM-bound ligand form with a register-capped instance (types, 9..16 channels): prep + binning then run on
        # a second stream in the shadow of the previous call's voxelize kernel, on alternating workspaces.
        split = (on_device and inputs_ready is not None and inputs_ready is not False and not converted[0] and tf_rows is None
                 and mode == "types" and 9 <= out_channels <= 16 and self.out_dtype != torch.float64
                 and L.mvx_voxelize_form(ctypes.byref(spec), ctypes.byref(b)) == 1)
        cur = torch.cuda.current_stream(self.device) if split else None
        # the caller's current stream as a raw handle (what the C ABI takes); the Stream object is only needed by the two-stream form
        raw_stream = cur.cuda_stream if split else torch._C._cuda_getCurrentRawStream(self.device.index)
        if split:
            ov = self._overlap
            if ov is None:
                ov = self._overlap = {"bin": torch.cuda.Stream(self.device, priority=-1), "idx": 0,
                                      "slots": [{"ws": None, "bin_done": torch.cuda.Event(), "free": torch.cuda.Event()} for _ in range(2)]}
                for sl in ov["slots"]:
                    sl["bin_done"].record(cur)   # creates the underlying cudaEvent_t
            sl = ov["slots"][ov["idx"]]
            ov["idx"] ^= 1
            if sl["ws"] is None or sl["ws"].numel() < total + 256:
                sl["ws"] = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8, device=self.device)
            ws = sl["ws"]
        else:
            ws = self._workspace(total)
        self._last_ws = ws
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        ws_bytes = ws.numel() - (ws_ptr - ws.data_ptr())

        def launch():
            args = (ctypes.byref(spec), ctypes.byref(b), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws_ptr), ctypes.c_size_t(ws_bytes))
            if split:
                bs = ov["bin"]
                if isinstance(inputs_ready, torch.cuda.Event):
                    bs.wait_event(inputs_ready)
                bs.wait_event(sl["free"])        # the voxelize kernel that last read this workspace has finished
                rc = L.mvx_voxelize_split(*args, ctypes.c_void_p(bs.cuda_stream), ctypes.c_void_p(cur.cuda_stream),
                                          ctypes.c_void_p(sl["bin_done"].cuda_event))
                sl["free"].record(cur)
                for t in keep:
                    if isinstance(t, torch.Tensor):
                        t.record_stream(bs)
                return rc
            fn = L.mvx_voxelize if on_device else L.mvx_voxelize_host
            return fn(*args, ctypes.c_void_p(raw_stream))
        if torch.cuda.current_device() == self.device.index:
            rc = launch()
        else:
            with torch.cuda.device(self.device):
                rc = launch()
        _lib.raise_for_status(rc)
        # shapes / flags of this call (pointers are not dereferenced again): compact() finds the column occupancy the
        # binning pass left in the workspace
        self._last_call = (spec, b, ws_ptr, out.data_ptr())   # b is this call's own struct: never written again
        if on_device:
            for t in keep:   # inputs converted on the fly must outlive the asynchronous kernels
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(cur if cur is not None else torch.cuda.current_stream(self.device))
        return out

    # ---- brick-sparse grids for host-side consumers (molvoxel_b200/sparse.py) ----
    def compact_into(self, grids: torch.Tensor, ids: torch.Tensor, vals: torch.Tensor, count: torch.Tensor):
        """Enqueue the brick compaction of `grids` (the output of the LAST forward_* call of this voxelizer, float32) on
        the current stream, no synchronisation: non-empty 8x8x8 bricks -> vals (cap, 512) float32 / ids (cap,) int32,
        their number -> count (1,) int32 (it may exceed the capacity: the surplus was dropped)."""
        assert self.out_dtype == torch.float32 and grids.dtype == torch.float32 and grids.is_contiguous(), "compact needs float32 grids"
        assert getattr(self, "_last_call", None) is not None, "compact() follows a forward_* call"
        spec, last, ws_ptr, out_ptr = self._last_call
        use_ws = out_ptr == grids.data_ptr() and int(last.num_mols) == int(grids.shape[0])
        if not use_ws:   # any other grid of this voxelizer's geometry: scan every column
            last = _lib.Batch()
            ctypes.memmove(ctypes.byref(last), ctypes.byref(self._last_call[1]), ctypes.sizeof(last))
            last.num_mols, last.out_channels = int(grids.shape[0]), int(grids.shape[1])
            last.num_channels = min(int(last.num_channels), int(grids.shape[1]))
        cap = min(int(ids.shape[0]), int(vals.shape[0]))
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = _lib.lib().mvx_compact_bricks(ctypes.byref(spec), ctypes.byref(last), ctypes.c_void_p(grids.data_ptr()),
                                               ctypes.c_void_p(ws_ptr if use_ws else 0), ctypes.c_void_p(ids.data_ptr()),
                                               ctypes.c_void_p(vals.data_ptr()), ctypes.c_uint32(cap),
                                               ctypes.c_void_p(count.data_ptr()), stream)
        _lib.raise_for_status(rc)

    def compact(self, grids: torch.Tensor, capacity: int | None = None) -> SparseGrids:
        """Brick-sparse form of the grids of the last forward_* call (synchronises once to learn the brick count;
        grows the buffers and repeats if `capacity` was too small)."""
        B, C, D = int(grids.shape[0]), int(grids.shape[1]), self._dimension
        nb = -(-D // 8)
        total = B * C * nb ** 3
        cap = int(capacity) if capacity else max(1024, total // 8)
        while True:
            ids = torch.empty(cap, dtype=torch.int32, device=self.device)
            vals = torch.empty((cap, 512), dtype=torch.float32, device=self.device)
            count = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.compact_into(grids, ids, vals, count)
            n = int(count.item())
            if n <= cap:
                return SparseGrids(ids[:n], vals[:n], (B, C, D))
            cap = n

    def check_status(self):
        """Synchronise and raise if a device-path call since the last check flagged bad types / radii (device-side validation)."""
        spaces = [w for w in [self._ws] + ([sl["ws"] for sl in self._overlap["slots"]] if self._overlap else []) if w is not None]
        if not spaces:
            return
        err = None
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            for w in spaces:
                ws_ptr = (w.data_ptr() + 255) // 256 * 256
                try:
                    _lib.raise_for_status(_lib.lib().mvx_check_status(ctypes.c_void_p(ws_ptr), stream))
                except ValueError as e:
                    err = e
        if err is not None:
            raise err
        if self._pipe is not None:   # status words copied back by the pipelined (non_blocking) calls: sticky
            flags = self._sticky_flags | int(self._pipe["status"][0]) | int(self._pipe["status"][1])
            self._pipe["status"].zero_()
            self._sticky_flags = 0
            if flags & 1:
                raise ValueError("a type index is outside [0, num_channels)")
            if flags & 2:
                raise ValueError("a radius exceeds max_radius")
