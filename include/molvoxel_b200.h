/*
 * molvoxel_b200.h — C ABI of the B200-native voxelization backend (libmolvoxel_b200.so).
 *
 * Drop-in boundary: this is what a `library="b200"` branch of the reference factory
 * (reference molvoxel/__init__.py:25-40) binds.  One call voxelizes a CSR batch of molecules
 * — the batched form of reference numpy/voxelizer.py forward_types (:240-315),
 * forward_features (:97-169) and forward_single (:370-436); B = 1 is the reference's
 * per-molecule call.  Plain pointers and sizes only: no torch / C++ types cross this boundary.
 *
 * All functions return MVX_OK (0) or a negative mvx_status; none throws.  mvx_last_error()
 * returns a thread-local message for the last failure on the calling thread.
 */
#ifndef MOLVOXEL_B200_H
#define MOLVOXEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVX_VERSION 200 /* 0.2.0 */

typedef enum mvx_status {
    MVX_OK = 0,
    MVX_ERR_NULL_POINTER = -1,
    MVX_ERR_BAD_ENUM = -2,
    MVX_ERR_BAD_SHAPE = -3,      /* the reference's shape asserts, numpy/voxelizer.py:171-192,317-342,438-455 */
    MVX_ERR_WORKSPACE = -4,      /* workspace too small / misaligned */
    MVX_ERR_CUDA = -5,           /* a CUDA runtime call failed */
    MVX_ERR_UNSUPPORTED = -6,    /* e.g. forward_single with channel-wise radii (numpy/voxelizer.py:443) */
    MVX_ERR_DEVICE_FLAG = -7     /* device-side validation failed (type out of range, radius > max_radius) */
} mvx_status;

enum { MVX_DENSITY_GAUSSIAN = 0, MVX_DENSITY_BINARY = 1 };          /* base/voxelizer.py:13 */
enum { MVX_RADII_SCALAR = 0, MVX_RADII_CHANNEL_WISE = 1, MVX_RADII_ATOM_WISE = 2 }; /* base/voxelizer.py:12 */
enum { MVX_MODE_SINGLE = 0, MVX_MODE_TYPES = 1, MVX_MODE_FEATURES = 2 };            /* base/voxelizer.py:121-128 */
enum { MVX_F32 = 0, MVX_F64 = 1, MVX_U8 = 2, MVX_F16 = 3 };   /* U8 / F16: compact feature rows only */
enum { MVX_OUT_F32 = 0, MVX_OUT_BF16 = 1, MVX_OUT_F16 = 2, MVX_OUT_F64 = 3 };   /* element type of the output grid */
/* Memory layout of the output grid.  CDHW is the reference's (numpy/voxelizer.py:60-70: channel, x, y, z).  DHWC is the
 * channels-last form the reference README writes its formulas in (README.md:138-142) and 3-D CNNs in
 * torch.channels_last_3d consume: (B, D, H, W, C), the C channels of one voxel contiguous.  Same values, bit for bit. */
enum { MVX_LAYOUT_CDHW = 0, MVX_LAYOUT_DHWC = 1 };
/* How the caller held the SCALAR radius.  numpy's promotion rules (NEP 50) make the reference's arithmetic depend on it:
 * a python float is weak (fp32 division dist32 / r, fp64 clip bounds); an np.float64 scalar is strong (fp64 division and
 * Gaussian, numpy/voxelizer.py:546-548); an np.float32 scalar demotes the python-float clip bounds to fp32 (:487-488). */
enum { MVX_RADIUS_PYFLOAT = 0, MVX_RADIUS_NP_F64 = 1, MVX_RADIUS_NP_F32 = 2 };
/* Random rigid transform of every reference forward_* (random_translation / random_rotation, numpy/voxelizer.py:265,
 * numpy/transform.py:43-80), applied per molecule to the centred coordinates inside the per-atom prep kernel. */
enum {
    MVX_TF_ROTATE = 1,           /* rotate by a unit quaternion, with the reference's quaternion products (numpy/_quaternion.py:28-54) */
    MVX_TF_TRANSLATE = 2,        /* add the (fp32-valued) translation (numpy/transform.py:56-59) */
    MVX_TF_TRANSLATE_ONCE = 4    /* with ROTATE: add it once (the reference's torch backend, torch/transform.py:56-60) instead of
                                    twice (its numpy and numba backends, numpy/transform.py:56-59) */
};

/* Constructor arguments of the reference Voxelizer (base/voxelizer.py:15-38, numpy/voxelizer.py:22-35). */
typedef struct mvx_grid_spec {
    double  resolution;       /* Angstrom per voxel */
    int32_t dimension;        /* D = H = W, 1..512 */
    int32_t density_type;     /* MVX_DENSITY_* */
    double  sigma;            /* Gaussian sigma (kwarg `sigma`, default 0.5) */
    int32_t radii_type;       /* MVX_RADII_* */
    int32_t compat_blockdim;  /* the reference `blockdim` whose half-voxel block cull is emulated
                                 (numpy/voxelizer.py:55,496-527; default 8).  <= 0 or >= dimension:
                                 exact mathematics (the reference with blockdim = dimension). */
} mvx_grid_spec;

/*
 * One batch of molecules in CSR form.  Every pointer is a DEVICE pointer for mvx_voxelize and a
 * HOST pointer for mvx_voxelize_host.  Argument meaning follows the reference forward_* calls:
 *   coords   (N,3)  f32|f64  atom coordinates                  (numpy/voxelizer.py:251)
 *   centers  (B,3)  f32|f64  or NULL = no centring             (:263)
 *   types    (N,)   int32    channel index per atom, TYPES     (:253)
 *   features (N,C)  f32      feature rows, FEATURES            (:110); u8 | f16 rows with features_dtype
 *   radius   python-float scalar when radii_type is SCALAR
 *   radii    (C,) f32 channel-wise | (N,) f32 atom-wise        (:111)
 *   transforms (B,7) f64 optional explicit rigid transform per molecule (:265): quaternion q0..q3, translation
 */
typedef struct mvx_batch {
    int32_t        mode;            /* MVX_MODE_* */
    int32_t        num_mols;        /* B */
    int64_t        total_atoms;     /* N = mol_offsets[B] */
    const int32_t *mol_offsets;     /* (B+1,) */
    const void    *coords;
    int32_t        coords_dtype;    /* MVX_F32 | MVX_F64 */
    const void    *centers;
    int32_t        centers_dtype;
    const int32_t *types;
    const void    *features;        /* element type by features_dtype (f32 unless stated) */
    int32_t        num_channels;    /* C: channels the inputs address (types < C; features row length) */
    int32_t        out_channels;    /* channels of `out` (>= C; surplus channels are zero, :337) */
    double         radius;
    const float   *radii;
    double         max_radius;      /* host-known upper bound of every radius in `radii` (array
                                       radii types).  For FEATURES + channel-wise it must be the
                                       exact max: the reference clips with radii.max() (:138). */
    const double  *transforms;      /* (B,7) f64 or NULL: per molecule a unit quaternion (q0, q1, q2, q3) and a
                                       translation (tx, ty, tz) — explicit parameters of the rigid transform
                                       selected by transform_flags (e.g. drawn on the host from numpy's global
                                       RNG in the reference's order).  NULL with transform_flags != 0: the
                                       parameters are drawn ON THE DEVICE from a counter-based generator
                                       (Philox4x32-10) keyed by (rng_seed, rng_offset + molecule index), so any
                                       sharding or chunking of a sweep gives the same augmentation. */
    int32_t        out_dtype;       /* MVX_OUT_F32 (the reference's precision=32 layout, default) or a
                                       reduced-precision grid: every voxel is computed in fp32 exactly as
                                       for MVX_OUT_F32 and rounded once (nearest-even) on the store.
                                       MVX_OUT_F64 is the reference's precision=64 (numpy/voxelizer.py:28-34):
                                       distances, kernel and accumulation in fp64, float64 grid (untuned path). */
    int32_t        features_dtype;  /* MVX_F32 (default) | MVX_U8 | MVX_F16: element type of `features`.  Compact rows
                                       (one-hot / flag / count features) are widened to fp32 on the device, exactly —
                                       the reference's features.astype(float32) (numpy/voxelizer.py:127-128) — so a
                                       host caller moves 1/4 or 1/2 of the bytes over PCIe. */
    int32_t        radius_kind;     /* MVX_RADIUS_*: how the scalar `radius` was typed by the caller (default python float) */
    int32_t        transform_flags; /* MVX_TF_* bits; 0 = no transform (transforms is then ignored) */
    uint64_t       rng_seed;        /* device-drawn transforms: generator key ... */
    uint64_t       rng_offset;      /* ... and global index of this batch's first molecule */
    double         random_translation;   /* device-drawn transforms: translation ~ U(-t, t)^3, rounded to fp32
                                            (numpy/transform.py:74-76) */
    int32_t        out_layout;      /* MVX_LAYOUT_CDHW (default, the reference's) | MVX_LAYOUT_DHWC (channels-last) */
} mvx_batch;

/* Bytes of device workspace mvx_voxelize needs for this spec/batch (256-byte aligned base). */
int mvx_workspace_bytes(const mvx_grid_spec *spec, const mvx_batch *batch, size_t *out_bytes);

/*
 * Voxelize a batch: out is a DEVICE buffer (B, out_channels, D, H, W) — or (B, D, H, W, out_channels) with
 * batch->out_layout = MVX_LAYOUT_DHWC — of batch->out_dtype (float32 by default), contiguous,
 * written exactly once per voxel (zeros included).  Work is enqueued on `stream`
 * (a cudaStream_t; NULL = legacy default stream); no host synchronisation happens inside.
 */
int mvx_voxelize(const mvx_grid_spec *spec, const mvx_batch *batch, void *out, void *workspace,
                 size_t workspace_bytes, void *stream);

/*
 * mvx_voxelize with the work split over two streams, for a caller that voxelizes a sequence of batches: the per-atom prep and
 * the binning kernels (4 % of a ligand step) go to `bin_stream`, `bin_done_event` (a cudaEvent_t of the caller) is recorded
 * there, and the voxelize kernel goes to `vox_stream` behind that event.  On `bin_stream` the work is sized to fit on the SMs
 * NEXT TO the HBM-bound voxelize kernel of the previous batch still running on `vox_stream` (128-thread CTAs; the ligand
 * voxelize kernel runs register-capped), so batch k+1's binning hides behind batch k's output writes.  The caller owns the
 * ordering: `bin_stream` must wait (cudaStreamWaitEvent) until this batch's inputs are complete and until the voxelize kernel
 * that last used `workspace` has finished — i.e. use two workspaces alternately.  Results are those of mvx_voxelize, bit for bit.
 */
int mvx_voxelize_split(const mvx_grid_spec *spec, const mvx_batch *batch, void *out, void *workspace,
                       size_t workspace_bytes, void *bin_stream, void *vox_stream, void *bin_done_event);

/*
 * Same, with HOST input buffers (what a numpy caller of the reference holds): copies the inputs
 * to the device, voxelizes into the DEVICE buffer `out`, and copies the device status word back
 * (one synchronisation).  `workspace` must hold mvx_workspace_bytes() + mvx_host_staging_bytes().
 * Pinned host memory makes the copies asynchronous but is not required.
 */
int mvx_host_staging_bytes(const mvx_grid_spec *spec, const mvx_batch *batch, size_t *out_bytes);
int mvx_voxelize_host(const mvx_grid_spec *spec, const mvx_batch *host_batch, void *out, void *workspace,
                      size_t workspace_bytes, void *stream);

/* Reads the device status word of the last mvx_voxelize on this workspace (synchronises `stream`); the next
 * mvx_voxelize on the workspace resets it. */
int mvx_check_status(void *workspace, void *stream);

/* Number of kernels one mvx_voxelize call launches for this spec/batch (bench.py's gpu_launches). */
int mvx_launches_per_call(const mvx_grid_spec *spec, const mvx_batch *batch);

/* Which voxelize kernel form mvx_voxelize picks for this spec/batch (by atom density; all forms give the same
 * results): 0 generic rows, 1 warp cells (ligand batches), 3 tiles, 4 pipelined persistent form (dense batches). */
int mvx_voxelize_form(const mvx_grid_spec *spec, const mvx_batch *batch);

/*
 * Per-kernel device timing for bench.py's roofline: between begin and end every mvx_voxelize call on
 * this thread records CUDA events around its launches on the caller's stream (no synchronisation until
 * end).  end returns the summed milliseconds of the prep, bin and voxelize launches over the recorded calls.
 */
int mvx_profile_begin(int max_calls);
int mvx_profile_end(double *ms_prep, double *ms_bin, double *ms_voxelize, int *num_calls);

/*
 * Brick compaction for host-side consumers (a ligand grid is >= 97 % zeros; copying it to the host densely measures
 * PCIe).  After mvx_voxelize(spec, batch, grids, workspace, ...) — float32 grids — on the same stream: every brick of
 * 8 x 8 x 8 voxels of one channel that holds a non-zero value is copied to brick_vals[slot] (512 floats, brick-local
 * [x][y][z], zero-padded where it sticks out of the grid) and brick_ids[slot] = ((mol * out_channels + c) * ncol + col) *
 * nbz + bz, with ncol = ceil(D/8)^2 columns (col = bx * ceil(D/8) + by) and nbz = ceil(D/8).  Slot order is arbitrary.
 * *count receives the number of non-empty bricks; if it exceeds `capacity` the surplus was dropped (call again with
 * larger buffers).  `workspace` = the voxelize call's workspace (its column occupancy lets empty columns be skipped
 * without reading them) or NULL (every column is read).  All pointers are DEVICE pointers.
 */
int mvx_compact_bricks(const mvx_grid_spec *spec, const mvx_batch *batch, const void *grids, const void *workspace,
                       uint32_t *brick_ids, float *brick_vals, uint32_t capacity, uint32_t *count, void *stream);

/*
 * The rigid transforms mvx_voxelize draws on the device for molecules [rng_offset, rng_offset + num_mols) with these
 * (rng_seed, transform_flags, random_translation): writes (num_mols, 7) f64 rows (quaternion, translation) to the
 * DEVICE buffer `out` on `stream`.  Passing them back as mvx_batch.transforms gives bit-identical grids.
 */
int mvx_random_transforms(uint64_t rng_seed, uint64_t rng_offset, int32_t num_mols, int32_t transform_flags,
                          double random_translation, double *out, void *stream);

/*
 * Synthetic ligands of the virtual-screening sweep (V ~ U{vmin..vmax} atoms, a 3-D random walk
 * with `step` Angstrom steps recentred to the origin, fp32-representable coordinates, types uniform over num_types),
 * generated on the DEVICE from the same counter-based generator keyed by (seed, first_mol + molecule index): any
 * sharding or chunking of a sweep sees the same molecules.  Benchmark / test input, not part of the hot path.
 * Pass 1 (mol_offsets == NULL): writes counts[num_mols].  Pass 2 (mol_offsets = exclusive prefix sums of the counts,
 * num_mols + 1 entries, relative to the first molecule): writes coords (N,3) of coords_dtype and types (N) (or NULL).
 */
int mvx_synth_ligands(uint64_t seed, uint64_t first_mol, int32_t num_mols, int32_t vmin, int32_t vmax, int32_t num_types,
                      double step, const int32_t *mol_offsets, int32_t *counts, void *coords, int32_t coords_dtype,
                      int32_t *types, void *stream);

const char *mvx_last_error(void);
int mvx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOLVOXEL_B200_H */
