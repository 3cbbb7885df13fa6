// mvx_common.cuh — shared types and geometry of the sm_100a kernels of the voxelization hot path
// (kernels: mvx_bin_kernels.cuh, mvx_vox_kernels.cuh, mvx_rigid.cuh; experiment: mvx_vox_ws.cuh).
//
// Path restated (semantics only; the structure is new): reference molvoxel/voxelizer/numpy/voxelizer.py
//   prologue + clip + block cull  :263-295, :481-527   -> mvx_prep_kernel   (one thread per atom, fp64)
//   per-block atom lists          :496-527             -> ligand batches: mvx_bin_* (CSR column lists, prefix sum) + mvx_expand_kernel
//                                                         dense batches:  counting sort by (molecule, column, z layer):
//                                                                         prep counts, mvx_lscan / mvx_lplace / mvx_lbuild_kernel
//   distance / density / channel accumulation  :531-560, :344-366, :194-236, :457-477
//                                                      -> gather kernels, every voxel written once:
//                                                         mvx_voxelize_cells_kernel  ligand batches, one CTA per tile, HBM-write-bound
//                                                         mvx_voxelize_pipe_kernel   dense batches, persistent, cp.async.bulk + mbarrier ring
//                                                         mvx_voxelize_tiles_kernel / _sweep_kernel  non-persistent form on the same entries
//                                                         mvx_voxelize_kernel        generic (any D);  mvx_voxelize_f64_kernel  precision=64
//
// Data layout in HBM
//   out      (B, Cout, D, H, W) fp32 (bf16 / fp16 / fp64 by out_dtype), W contiguous (reference layout, numpy/voxelizer.py:60-70);
//            (B, D, H, W, Cout) with out_layout = DHWC (channels-last instances of the same kernels)
//   AtomRec  40 B per atom: centred fp64 position, fp32 radius, cull "forbidden planes", z voxel range
//   lists    uint32 atom ids per (molecule, 8x8 voxel column), ascending = the reference's atom order (ligand batches)
//   lent     layered entries per (molecule, column, 16-voxel z layer), record + feature row, ascending atom order (dense batches)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvx {

constexpr int kTile = 8;          // column footprint in x and y (voxels)
constexpr int kThreads = 256;     // voxelize CTA size
constexpr int kMaxCand = 128;     // candidates staged per round
constexpr float kLog2e = 1.4426950408889634f;
// cell shape of the warp-cell kernels: 2 x 4 x 16 voxels (4 lanes of one float4 along z per row)
constexpr int kCellX = 2, kCellY = 4, kCellZ = 16;
constexpr int kCellsXY = (kTile / kCellX) * (kTile / kCellY);   // 8 cells per z layer of a tile

enum : int { kFlagBadType = 1, kFlagRadiusOverMax = 2 };

struct __align__(8) AtomRec {
    double px, py, pz;   // coords - center, fp64 (numpy/voxelizer.py:263-268)
    float r;             // fp32 radius of the density kernel (max radius for channel-wise features)
    int16_t fx, fy, fz;  // voxel plane the reference's block cull removes this atom from, or -1
    int16_t zlo, zhi;    // conservative voxel z range
    int16_t pad;
};
static_assert(sizeof(AtomRec) == 40, "AtomRec layout");

struct Geo {
    double res, inv_res, half_width, res_half, lower, upper;
    double clip_lo, clip_hi;     // scalar-form clip thresholds (lower - r, upper + r)
    double size_scalar;          // scalar-form atom_size for the cull
    double sigma;
    int dim, bd, nb;             // nb == 1: exact mode (no cull)
    int ncx;                     // columns per axis = ceil(dim / 8)
    int scalar_form;             // clip/cull use the scalar thresholds
    int radii_src;               // 0 scalar, 1 radii[n], 2 radii[types[n]], 3 channel-wise features
    int cols_axis_max;           // capacity: columns an atom may span per axis
    float r_scalar32;
};

struct PrepParams {
    Geo g;
    int mode, B, C;
    int64_t N;
    const int32_t* mol_offsets;
    const void* coords; int coords_f64;
    const void* centers; int centers_f64;
    const int32_t* types;
    const float* radii;
    const double* transforms;   // (B,7) explicit (quaternion, translation) per molecule, or nullptr = drawn here
    int tf_flags;               // kTf* bits; 0 = no rigid transform
    unsigned long long rng_seed, rng_offset;
    double rng_translation;
    AtomRec* recs;
    uint32_t* colrange;
    int* status;
    // layered forms: per atom, the 16-voxel z layers (global layer index = z chunk * layers-per-chunk + layer) its
    // cutoff sphere reaches; nullptr otherwise
    uint32_t* alayers;
    uint32_t* kcnt;     // layered forms: atoms per (molecule, column, layer), counted here with fire-and-forget atomics
    int nzc, tz, ncol, nl, zl;   // zl: layers reserved per atom
    float tau_lin, tau_quad;
};

struct BinParams {
    int B, ncol, ncx, maxcols;
    const int32_t* mol_offsets;
    const uint32_t* colrange;
    uint2* bins;        // (start relative to the molecule's segment, count) per (mol, column)
    uint32_t* lists;    // molecule m owns [mol_offsets[m]*maxcols, mol_offsets[m+1]*maxcols)
};

// One atom as seen from one 8x8 voxel column, ready for the voxelize kernel's shared-memory staging (48 B).
struct __align__(16) ColEntry {
    float ax, ay, az, r2hi;     // position relative to (column x0, column y0, grid z = 0);  r^2 + tau
    float r2lo, kc;             // r^2 - tau;  -0.5*log2(e)/(r*sigma)^2
    uint32_t forb;              // forbidden planes: fx | fy << 8 (column-relative, 0xFF none) | fz << 16 (absolute, 0xFFFF none)
    uint32_t type_or_r;         // TYPES: channel index;  else: fp32 radius bits
    uint32_t n;                 // global atom id (features row, exact recheck)
    uint32_t mask_lo, mask_hi;  // cells (layer * 8 + cx * 2 + cy) whose voxel-centre box the cutoff sphere reaches
    uint32_t pad;
};
static_assert(sizeof(ColEntry) == 48, "ColEntry layout");

// One tile (8 x 8 x tz voxels of one molecule) as the pipelined kernel sees it: where its layered entries start, how
// many there are, where each 16-voxel z layer ends, and which voxels it covers (so that no warp has to take the tile
// index apart with integer divisions).  32 bytes, fetched by one bulk copy.
struct __align__(16) TileDesc {
    unsigned long long start;   // first layered entry of the tile (index into lent, in entries)
    uint32_t total;             // entries of the tile (the layers of a z chunk are consecutive)
    uint32_t mol;               // molecule
    uint16_t lend[4];           // end of layer k relative to start (tz <= 64: at most 4 layers); saturates at 65535 —
                                // such tiles exceed what the pipelined form takes and go to the tile form, which does not read it
    uint32_t origin;            // first voxel: x0 | y0 << 10 | z0 << 20
    uint32_t pad;
};
static_assert(sizeof(TileDesc) == 32, "TileDesc layout");


struct ExpandParams {
    double res, half_width, sigma;
    float tau_lin, tau_quad;
    int dim, ncx, ncol, nzc, tz, maxcols, mode, masks, B;
    const int32_t* mol_offsets;
    const AtomRec* recs;
    const uint2* bins;
    const uint32_t* lists;
    const int32_t* types;
    ColEntry* entries;   // CELLS form: one 48-byte record per (column, atom)
};

struct VoxParams {
    double res, half_width, sigma;
    float tau_lin, tau_quad;
    int dim, ncx, ncol, nzc, tz;
    int C, Cout, c_begin, c_end;
    int maxcols;
    int cull;                  // 1: the reference's block cull is emulated (compat_blockdim < dimension)
    const int32_t* mol_offsets;
    const AtomRec* recs;
    const uint2* bins;
    const uint32_t* lists;
    const int32_t* types;
    const float* features;
    const float* chan_radii;   // channel-wise features: kernel radius of channel c_begin
    const ColEntry* entries;   // expanded column lists (warp-cell kernel)
    int masks;                 // 1: entries carry precomputed cell masks
    int nlayers, zl, es4;      // layered entries (see ExpandParams)
    const float4* lent;
    const uint2* lbins;
    const TileDesc* tdesc;     // pipelined form; for the tile form: non-null = only tiles with more than pipe_sc entries
    int pipe_sc;               // pipelined form: largest tile (entries) it takes; larger ones go to the tile form
    int pipe_q;                // pipelined form: float4 words of the shared-memory entry ring
    int ws_nb;                 // pipelined form: 0 = every warp does whole cells; 4 | 8 | 20 = warp-specialised (see ws_builders / ws_walkers)
    int lean;                  // cells form: the register-capped instance (the next batch's binning shares the SMs)
    void* out;                 // (B, Cout, D, D, D), element type by out_kind
    int out_kind;              // 0 fp32 (reference), 1 bf16, 2 fp16 (reduced-precision output)
    int clast;                 // 1: channels-last output (B, D, D, D, Cout) — the channels of a voxel are contiguous (row f3)
};

// ---- shared-memory geometry of the voxelize forms (the host plan needs it as well) ----
constexpr int kWarpList = 64;    // warp-private list capacity (two 32-bit hit masks)
#ifndef MVX_PIPE_THREADS
#define MVX_PIPE_THREADS 384
#endif
constexpr int kPipeThreads = MVX_PIPE_THREADS;   // 12 warps: 168 registers per thread, no spills
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeSlots = 8;    // tiles in flight (descriptor + barrier slots)
#ifndef MVX_PIPE_SMEM
#define MVX_PIPE_SMEM 232448
#endif
constexpr int kPipeSmemBytes = MVX_PIPE_SMEM;   // 227 KB: the whole SM
constexpr int kPipeFixedBytes = kPipeSlots * ((int)sizeof(TileDesc) + 3 * 8 + 2 * 4) +
                                kPipeWarps * kWarpList * ((int)sizeof(float4) + (int)sizeof(uint16_t));
// With several channel chunks per cell (C > 16) the weights of a cell's hits are computed once, cached per lane
// (4 weights + the entry's shared address) and reused by the later chunks: kPipeHitCache hits per lane.
constexpr int kPipeHitCache = 12;
constexpr int kPipeCacheBytes = kPipeWarps * kPipeHitCache * 32 * ((int)sizeof(float4) + (int)sizeof(uint32_t));
constexpr int pipe_ring_q(bool multi) { return (kPipeSmemBytes - kPipeFixedBytes - (multi ? kPipeCacheBytes : 0)) / 16; }


// ---- warp-specialised pipelined form: NB list-builder warps (few registers: cell filter + near test) hand
// (entry indices, per-lane hit masks) through a shared-memory job ring to NWK accumulator warps (many registers: hit walk + stores)
constexpr int kWsJobs = 32;                                   // slots of the job ring (a power of two)
constexpr int kWsList = 128;                                  // list capacity of a job (four 32-bit hit masks per lane)
constexpr int kWsJobBytes = 16 + kWsList * 2 + 32 * 16;       // header, entry indices (u16), hit masks (16 B per lane)
constexpr int ws_fixed_bytes(int nb) {   // tile slots (descriptor, 3 barriers, 3 words), tickets, job ring (+ full barrier, generation), builder lists
    return kPipeSlots * ((int)sizeof(TileDesc) + 3 * 8 + 3 * 4) + 32 + kWsJobs * (kWsJobBytes + 16) + nb * kWsList * (int)sizeof(float4);
}
constexpr int ws_cache_bytes(int nwk) { return nwk * kPipeHitCache * 32 * ((int)sizeof(float4) + (int)sizeof(uint32_t)); }
constexpr int ws_ring_q(int nb, int nwk, bool multi) { return (kPipeSmemBytes - ws_fixed_bytes(nb) - (multi ? ws_cache_bytes(nwk) : 0)) / 16; }
// role split by the experiment knob MVX_WS / VoxParams::ws_nb: 4 -> 4 + 12 warps, 8 -> 8 + 8, 20 -> 8 + 12 (640 threads)
constexpr int ws_builders(int knob) { return knob == 4 ? 4 : 8; }
constexpr int ws_walkers(int knob) { return knob == 8 ? 8 : 12; }

}  // namespace mvx
