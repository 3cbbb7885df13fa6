// mvx_api.cu — C ABI of libmolvoxel_b200.so (declared in include/molvoxel_b200.h).
// Host-side plumbing only: argument checks mirroring the reference's asserts, workspace carving,
// kernel dispatch.  No torch, no CPU compute fallback: without a CUDA device every entry point
// that would launch work fails with MVX_ERR_CUDA.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/molvoxel_b200.h"
#include "mvx_bin_kernels.cuh"
#include "mvx_launch.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define MVX_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(MVX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
    } while (0)

constexpr size_t kAlign = 256;

// Optional per-kernel timing (bench.py's roofline): CUDA events recorded around each launch on the
// caller's stream while a profile is open.  No synchronisation until mvx_profile_end.
struct Profile {
    bool active = false;
    int max_calls = 0, calls = 0;
    cudaEvent_t* ev = nullptr;   // 5 events per call: before prep, after prep, after bin (binning stream); before / after voxelize
};
thread_local Profile g_prof;

void prof_mark(cudaStream_t st, int slot) {
    if (g_prof.active && g_prof.calls < g_prof.max_calls) cudaEventRecord(g_prof.ev[g_prof.calls * 5 + slot], st);
}
size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Plan {
    mvx::Geo geo;
    int ncol, nzc, tz, maxcols, nv;
    float tau_lin, tau_quad;
    // workspace offsets (bytes)
    size_t off_status, off_recs, off_colrange, off_alayers, off_bins, off_lists, off_entries, total;
    int masks;   // expand pass precomputes the cell masks (<= 64 cells per column)
    int form;        // voxelize kernel form, see enum Form
    int ncell;
    int nlayers, zl, es4;
    size_t off_lent, off_lbins, off_tdesc, off_kcnt, off_lids, off_wide;   // off_wide: compact features widened to fp32   // kcnt: key counts, then key cursors
    int pipe_sc;   // pipelined form: largest tile (entries) it takes
    int pipe_q;    // ... float4 words of its shared-memory ring
    bool pipe_multi;   // ... with the hit-weight cache (several channel chunks per cell)
    int ws_nb;         // ... warp-specialised: list-builder warps (0 = the classic form)
};

int check_args(const mvx_grid_spec* s, const mvx_batch* b) {
    if (!s || !b) return fail(MVX_ERR_NULL_POINTER, "spec/batch is NULL");
    if (s->dimension < 1 || s->dimension > 512) return fail(MVX_ERR_BAD_SHAPE, "dimension must be in 1..512");
    if (!(s->resolution > 0.0)) return fail(MVX_ERR_BAD_SHAPE, "resolution must be positive");
    if (s->density_type != MVX_DENSITY_GAUSSIAN && s->density_type != MVX_DENSITY_BINARY)
        return fail(MVX_ERR_BAD_ENUM, "density_type");
    if (s->radii_type < MVX_RADII_SCALAR || s->radii_type > MVX_RADII_ATOM_WISE) return fail(MVX_ERR_BAD_ENUM, "radii_type");
    if (b->mode < MVX_MODE_SINGLE || b->mode > MVX_MODE_FEATURES) return fail(MVX_ERR_BAD_ENUM, "mode");
    if (s->density_type == MVX_DENSITY_GAUSSIAN && !(s->sigma > 0.0)) return fail(MVX_ERR_BAD_SHAPE, "sigma must be positive");
    if (b->num_mols < 0 || b->total_atoms < 0) return fail(MVX_ERR_BAD_SHAPE, "negative batch size");
    if (b->total_atoms > 0x7fffffffLL) return fail(MVX_ERR_BAD_SHAPE, "more than 2^31-1 atoms in one batch");
    if (b->mode == MVX_MODE_SINGLE && s->radii_type == MVX_RADII_CHANNEL_WISE)
        return fail(MVX_ERR_UNSUPPORTED, "Channel-Wise Radii Type is not supported");   // numpy/voxelizer.py:443
    const int C = b->mode == MVX_MODE_SINGLE ? 1 : b->num_channels;
    if (C < 1) return fail(MVX_ERR_BAD_SHAPE, "num_channels must be >= 1");
    if (b->out_channels < C)   // numpy/voxelizer.py:337 (types), :192 (features), :450 (single)
        return fail(MVX_ERR_BAD_SHAPE, "Output channel is less than number of types");
    if (b->mode != MVX_MODE_TYPES && b->out_channels != C) return fail(MVX_ERR_BAD_SHAPE, "Output grid dimension incorrect");
    if (b->out_dtype < MVX_OUT_F32 || b->out_dtype > MVX_OUT_F64) return fail(MVX_ERR_BAD_ENUM, "out_dtype");
    if (b->features_dtype != MVX_F32 && b->features_dtype != MVX_U8 && b->features_dtype != MVX_F16) return fail(MVX_ERR_BAD_ENUM, "features_dtype");
    if (b->out_layout != MVX_LAYOUT_CDHW && b->out_layout != MVX_LAYOUT_DHWC) return fail(MVX_ERR_BAD_ENUM, "out_layout");
    if (b->radius_kind < MVX_RADIUS_PYFLOAT || b->radius_kind > MVX_RADIUS_NP_F32) return fail(MVX_ERR_BAD_ENUM, "radius_kind");
    if (b->transform_flags & ~(MVX_TF_ROTATE | MVX_TF_TRANSLATE | MVX_TF_TRANSLATE_ONCE)) return fail(MVX_ERR_BAD_ENUM, "transform_flags");
    if ((b->transform_flags & MVX_TF_TRANSLATE) && !b->transforms && !(b->random_translation > 0.0))
        return fail(MVX_ERR_BAD_SHAPE, "random_translation must be positive when the translation is drawn on the device");
    if (b->num_mols > 0 && !b->mol_offsets) return fail(MVX_ERR_NULL_POINTER, "mol_offsets");
    if (b->total_atoms > 0) {
        if (!b->coords) return fail(MVX_ERR_NULL_POINTER, "coords");
        if (b->mode == MVX_MODE_TYPES && !b->types) return fail(MVX_ERR_NULL_POINTER, "types");
        if (b->mode == MVX_MODE_FEATURES && !b->features) return fail(MVX_ERR_NULL_POINTER, "features");
    }
    if (s->radii_type == MVX_RADII_SCALAR) {
        if (!(b->radius > 0.0)) return fail(MVX_ERR_BAD_SHAPE, "the radii type of voxelizer is `scalar`, radii should be a positive scalar");
    } else {
        if (!b->radii && (b->total_atoms > 0 || s->radii_type == MVX_RADII_CHANNEL_WISE))
            return fail(MVX_ERR_NULL_POINTER, "radii");
        if (!(b->max_radius > 0.0)) return fail(MVX_ERR_BAD_SHAPE, "max_radius must bound the radii array");
    }
    return MVX_OK;
}

int pick_chunk(int mode, int nchan);

using mvx::FORM_ROWS; using mvx::FORM_CELLS; using mvx::FORM_TILES; using mvx::FORM_PIPE;
using mvx::DeviceSet; using mvx::set_smem;

bool layered(int form) { return form == FORM_TILES || form == FORM_PIPE; }

int make_plan(const mvx_grid_spec* s, const mvx_batch* b, Plan* pl) {
    int rc = check_args(s, b);
    if (rc != MVX_OK) return rc;
    mvx::Geo& g = pl->geo;
    const int D = s->dimension;
    g.res = s->resolution;
    g.inv_res = 1.0 / s->resolution;
    const double width = s->resolution * (double)(D - 1);   // base/voxelizer.py:28
    g.half_width = width / 2.0;                            // numpy/voxelizer.py:42
    g.res_half = s->resolution / 2.0;                      // :55
    g.upper = width / 2.0;                                 // base/voxelizer.py:33
    g.lower = -1 * g.upper;                                // base/voxelizer.py:34
    g.sigma = s->sigma;
    g.dim = D;
    g.bd = (s->compat_blockdim <= 0 || s->compat_blockdim >= D) ? D : s->compat_blockdim;
    g.nb = (D + g.bd - 1) / g.bd;
    g.ncx = (D + mvx::kTile - 1) / mvx::kTile;

    const bool chan_feat = b->mode == MVX_MODE_FEATURES && s->radii_type == MVX_RADII_CHANNEL_WISE;
    g.scalar_form = (s->radii_type == MVX_RADII_SCALAR) || chan_feat;
    double reach;
    if (s->radii_type == MVX_RADII_SCALAR) {
        g.radii_src = 0;
        g.size_scalar = b->radius;
        g.r_scalar32 = (float)b->radius;
        g.clip_lo = g.lower - b->radius;     // numpy/voxelizer.py:487
        g.clip_hi = g.upper + b->radius;     // :488
        if (b->radius_kind == MVX_RADIUS_NP_F64) {
            // np.float64 scalar: dr = fp64(dist32) / r64 <= 1.0 in fp64 (numpy/voxelizer.py:546-555), i.e. dist32 <= r64,
            // i.e. dist32 <= the largest fp32 not above r64 — the kernels' fp32 radius is r64 rounded DOWN
            if ((double)g.r_scalar32 > b->radius) g.r_scalar32 = std::nextafterf(g.r_scalar32, 0.f);
        } else if (b->radius_kind == MVX_RADIUS_NP_F32) {
            // np.float32 scalar: python-float bound -/+ np.float32 is fp32 arithmetic (NEP 50, :487-488)
            g.clip_lo = (double)((float)g.lower - g.r_scalar32);
            g.clip_hi = (double)((float)g.upper + g.r_scalar32);
        }
        reach = std::fmax(b->radius, (double)g.r_scalar32);
    } else if (chan_feat) {
        // atom_size = radii.max() is an np.float32 scalar: the python-float bounds are demoted (NEP 50),
        // so the clip thresholds are fp32 results (numpy/voxelizer.py:138, :487-488)
        g.radii_src = 3;
        const float rmax = (float)b->max_radius;
        g.size_scalar = (double)rmax;
        g.r_scalar32 = rmax;
        g.clip_lo = (double)((float)g.lower - rmax);
        g.clip_hi = (double)((float)g.upper + rmax);
        if (b->out_dtype == MVX_OUT_F64) {   // precision=64: radii.astype(float64).max() keeps the bounds in fp64
            g.clip_lo = g.lower - (double)rmax;
            g.clip_hi = g.upper + (double)rmax;
        }
        reach = (double)rmax;
    } else {
        g.radii_src = (s->radii_type == MVX_RADII_ATOM_WISE) ? 1 : 2;
        g.size_scalar = 0.0;
        g.r_scalar32 = 0.f;
        g.clip_lo = g.clip_hi = 0.0;
        reach = b->max_radius;
    }
    const int span = (int)std::ceil(2.0 * reach * (1.0 + 1e-6) / s->resolution + 0.02) + 2;
    g.cols_axis_max = span / mvx::kTile + 2;
    if (g.cols_axis_max > g.ncx) g.cols_axis_max = g.ncx;
    pl->maxcols = g.cols_axis_max * g.cols_axis_max;
    pl->ncol = g.ncx * g.ncx;
    pl->nv = (D % 4 == 0) ? 4 : 1;
    pl->nzc = (D + 63) / 64;
    int tz = (D + pl->nzc - 1) / pl->nzc;
    tz = (tz + 3) / 4 * 4;
    if (const char* e = std::getenv("MVX_TZ")) {   // experiments: z extent of a tile (whole 16-voxel layers)
        const int v = std::atoi(e);
        if (v >= 16 && v <= 64 && v % 16 == 0 && (D + v - 1) / v * (v / 16) <= 32) tz = v;
    }
    pl->tz = tz;
    pl->nzc = (D + tz - 1) / tz;
    // tolerance band of the fp32 cutoff test (see DESIGN.md "cutoff decisions")
    // coordinates are column-relative in x, y and grid-relative in z (ColEntry), so the extent is the grid's
    const double ext = (double)D * s->resolution + reach;
    pl->tau_lin = (float)(21.0 * std::ldexp(1.0, -24) * ext);
    pl->tau_quad = (float)(12.0 * std::ldexp(1.0, -24));

    const size_t N = (size_t)b->total_atoms, B = (size_t)b->num_mols;
    size_t off = 0;
    pl->off_status = off;   off += kAlign;
    pl->off_recs = off;     off += align_up(N * sizeof(mvx::AtomRec));
    pl->off_colrange = off; off += align_up(N * sizeof(uint32_t));
    pl->off_bins = off;     off += align_up(B * (size_t)pl->ncol * sizeof(uint2));
    const int layers = pl->nzc * ((tz + mvx::kCellZ - 1) / mvx::kCellZ);
    pl->ncell = layers * mvx::kCellsXY;
    pl->masks = pl->nv == 4 && pl->ncell <= 64;
    // z layers one cutoff sphere can reach (prep enforces it): a span of S voxels touches at most floor(S / 16) + 2 full
    // layers, plus one more per z-chunk end inside the span when chunks end in a short layer (tz % 16 != 0)
    const double span_vox = 2.0 * reach * (1.0 + 1e-3) / s->resolution;
    int zl = (int)std::floor(span_vox / mvx::kCellZ) + 2 + ((tz % mvx::kCellZ) != 0 ? (int)std::floor(span_vox / tz) + 1 : 0);
    if (zl > layers) zl = layers;
    // Kernel form by density: expected entries per column = atoms * columns-per-atom / columns.
    {
        const double cols_per_atom = std::pow(2.0 * reach / (mvx::kTile * s->resolution) + 1.0, 2.0);
        const double per_col = (B > 0) ? (double)N * cols_per_atom / ((double)B * pl->ncol) : 0.0;
        pl->form = pl->nv != 4 ? FORM_ROWS : (per_col >= 64.0 ? FORM_PIPE : FORM_CELLS);
        if (const char* e = std::getenv("MVX_KERNEL")) {   // experiments / tests
            if (std::strcmp(e, "rows") == 0) pl->form = FORM_ROWS;
            else if (pl->nv == 4 && std::strcmp(e, "cells") == 0) pl->form = FORM_CELLS;
            else if (pl->nv == 4 && std::strcmp(e, "tiles") == 0) pl->form = FORM_TILES;
            else if (pl->nv == 4 && std::strcmp(e, "pipe") == 0) pl->form = FORM_PIPE;
        }
        if (b->out_dtype == MVX_OUT_F64) pl->form = FORM_ROWS;   // precision=64 runs on the generic column lists
        // channel-wise features patch the staged radii per channel pass, which needs the CTA-synchronous staging
        if (pl->form == FORM_PIPE && chan_feat) pl->form = per_col >= 200.0 ? FORM_TILES : FORM_CELLS;
    }
    {   // layered entries of the tile kernel
        const int C = b->mode == MVX_MODE_FEATURES ? b->num_channels : 0;
        int chunk = pick_chunk(b->mode, b->out_channels);
        if (chunk < 4) chunk = 4;
        const int cs = (C + chunk - 1) / chunk * chunk;
        pl->es4 = 3 + cs / 4;
        pl->nlayers = layers;
        pl->zl = zl;
        const size_t nle = layered(pl->form) ? N * (size_t)pl->maxcols * (size_t)zl : 0;
        pl->off_lent = off;  off += align_up(nle * (size_t)pl->es4 * 16);
        pl->off_lbins = off; off += align_up(layered(pl->form) ? B * (size_t)pl->ncol * (size_t)layers * sizeof(uint2) : 0);
        pl->off_tdesc = off; off += align_up(pl->form == FORM_PIPE ? B * (size_t)pl->ncol * (size_t)pl->nzc * sizeof(mvx::TileDesc) : 0);
        const size_t nkeys = layered(pl->form) ? B * (size_t)pl->ncol * (size_t)layers : 0;
        pl->off_kcnt = off;  off += align_up(2 * nkeys * sizeof(uint32_t));
        pl->off_lids = off;  off += align_up(nle * sizeof(uint32_t));
        // several channel chunks per cell: the pipelined kernel caches the hit weights, its ring is smaller
        pl->pipe_multi = pick_chunk(b->mode, b->out_channels) == 16 && b->out_channels > 16;
        pl->pipe_q = mvx::pipe_ring_q(pl->pipe_multi);
        pl->ws_nb = 0;
#ifdef MVX_WITH_WS   // experiment build: the warp-specialised variant (mvx_vox_ws.cuh)
        if (pl->form == FORM_PIPE && b->mode == MVX_MODE_FEATURES && pick_chunk(b->mode, b->out_channels) == 16) {
            if (const char* e = std::getenv("MVX_WS")) {   // experiments: 4 | 8 list-builder warps
                const int v = std::atoi(e);
                if ((v == 4 || v == 8 || v == 20) && b->out_layout == MVX_LAYOUT_CDHW) pl->ws_nb = v;
            }
            if (pl->ws_nb) pl->pipe_q = mvx::ws_ring_q(mvx::ws_builders(pl->ws_nb), mvx::ws_walkers(pl->ws_nb), pl->pipe_multi);
        }
#endif
        pl->pipe_sc = pl->pipe_q / 2 / pl->es4;   // the pipelined form takes tiles up to half its ring
    }
    pl->off_alayers = off;  off += align_up(layered(pl->form) ? N * sizeof(uint32_t) : 0);
    pl->off_lists = off;    off += align_up(layered(pl->form) ? 0 : N * (size_t)pl->maxcols * sizeof(uint32_t));
    pl->off_entries = off;  off += align_up(pl->form == FORM_CELLS ? N * (size_t)pl->maxcols * sizeof(mvx::ColEntry) : 0);
    // compact feature rows: the layered forms widen them inside the entry build; the others read a widened fp32 copy
    pl->off_wide = off;     off += align_up((b->mode == MVX_MODE_FEATURES && b->features_dtype != MVX_F32 && !layered(pl->form)) ? N * (size_t)b->num_channels * sizeof(float) : 0);
    pl->total = off;
    return MVX_OK;
}

template <int MODE, int CH>
cudaError_t launch_density(const mvx::VoxParams& vp, bool binary, int form, int nv, unsigned grid, cudaStream_t st) {
    return binary ? mvx::launch_form<MODE, CH, true>(vp, form, nv, grid, st) : mvx::launch_form<MODE, CH, false>(vp, form, nv, grid, st);
}

// column groups per molecule for the bin pass: 1 (fused kernel) once the batch alone gives two waves of
// CTAs on the 148 SMs, otherwise enough groups to get there (at least 8 columns = one per warp each).
int bin_groups(int B, int ncol, long long N = -1) {
    if (B >= 296) return 1;
    if (N >= 0 && N <= 512) return 1;   // a few small molecules (the one-ligand-per-call pattern): latency, one launch
    int g = (1184 + B - 1) / (B > 0 ? B : 1);   // aim at ~8 CTAs per SM so the latency-bound scans overlap
    int gmax = (ncol + 7) / 8;
    if (g > gmax) g = gmax;
    return g < 1 ? 1 : g;
}

int pick_chunk(int mode, int nchan) {
    if (mode == MVX_MODE_SINGLE) return 1;
    if (const char* e = std::getenv("MVX_CH")) {   // experiments: force the channel chunk
        int v = std::atoi(e);
        if (v == 1 || v == 4 || v == 8 || v == 12 || v == 16) return v;
    }
    if (nchan <= 1) return 1;
    if (nchan <= 4) return 4;
    if (nchan <= 8) return 8;
    if (nchan <= 12) return 12;   // e.g. the 9-channel ligand sweep: no dead accumulator rows
    return 16;
}

cudaError_t launch_vox(int mode, int ch, const mvx::VoxParams& vp, bool binary, int form, int nv, unsigned grid, cudaStream_t st) {
    if (mode == MVX_MODE_SINGLE) return launch_density<0, 1>(vp, binary, form, nv, grid, st);
    if (mode == MVX_MODE_TYPES) {
        switch (ch) {
            case 1: return launch_density<1, 1>(vp, binary, form, nv, grid, st);
            case 4: return launch_density<1, 4>(vp, binary, form, nv, grid, st);
            case 8: return launch_density<1, 8>(vp, binary, form, nv, grid, st);
            case 12: return launch_density<1, 12>(vp, binary, form, nv, grid, st);
            default: return launch_density<1, 16>(vp, binary, form, nv, grid, st);
        }
    }
    switch (ch) {
        case 1: return launch_density<2, 1>(vp, binary, form, nv, grid, st);
        case 4: return launch_density<2, 4>(vp, binary, form, nv, grid, st);
        case 8: return launch_density<2, 8>(vp, binary, form, nv, grid, st);
        case 12: return launch_density<2, 12>(vp, binary, form, nv, grid, st);
        default: return launch_density<2, 16>(vp, binary, form, nv, grid, st);
    }
}

}  // namespace

extern "C" {

int mvx_version(void) { return MVX_VERSION; }

const char* mvx_last_error(void) { return g_last_error.c_str(); }

int mvx_workspace_bytes(const mvx_grid_spec* spec, const mvx_batch* batch, size_t* out_bytes) {
    if (!out_bytes) return fail(MVX_ERR_NULL_POINTER, "out_bytes is NULL");
    Plan pl;
    int rc = make_plan(spec, batch, &pl);
    if (rc != MVX_OK) return rc;
    *out_bytes = pl.total;
    return MVX_OK;
}

int mvx_launches_per_call(const mvx_grid_spec* spec, const mvx_batch* batch) {
    Plan pl;
    int rc = make_plan(spec, batch, &pl);
    if (rc != MVX_OK) return rc;
    if (batch->num_mols == 0) return 0;
    const bool chan_feat = batch->mode == MVX_MODE_FEATURES && spec->radii_type == MVX_RADII_CHANNEL_WISE;
    int nvox = (chan_feat && batch->out_dtype != MVX_OUT_F64) ? batch->num_channels : 1;
    const int nbin = layered(pl.form) ? (batch->total_atoms > 0 ? 3 : 1) : (bin_groups(batch->num_mols, pl.ncol, batch->total_atoms) <= 1 ? 1 : 2);   // scan, place, build
    const int nexp = (pl.form == FORM_CELLS && batch->total_atoms > 0 && bin_groups(batch->num_mols, pl.ncol, batch->total_atoms) > 1) ? 1 : 0;   // fused into bin for big batches
    const int nwide = (batch->mode == MVX_MODE_FEATURES && batch->features_dtype != MVX_F32 && batch->total_atoms > 0 && !layered(pl.form)) ? 1 : 0;
    return (batch->total_atoms > 0 ? 1 : 0) + nwide + nbin + nexp + nvox * (pl.form == FORM_PIPE ? 2 : 1);   // prep + bin + expand + voxelize
}

int mvx_voxelize_form(const mvx_grid_spec* spec, const mvx_batch* batch) {
    Plan pl;
    int rc = make_plan(spec, batch, &pl);
    return rc != MVX_OK ? rc : pl.form;
}

}  // extern "C"

namespace {
// One call's launches.  st: the stream of the per-atom prep and the binning kernels; st_vox: the stream of the voxelize
// kernel(s).  When they differ (mvx_voxelize_split) `bin_done` orders the voxelize kernel after the binning, prep / binning
// use 128-thread CTAs and the cells form its register-capped instance, so that both fit on an SM next to the voxelize
// CTAs of the PREVIOUS call that are still running on st_vox.
int enqueue(const mvx_grid_spec* spec, const mvx_batch* batch, void* out, void* workspace, size_t workspace_bytes,
            cudaStream_t st, cudaStream_t st_vox, cudaEvent_t bin_done) {
    const bool split = st != st_vox;
    const int pt = split ? 128 : 256;   // threads per CTA of prep / bin
    Plan pl;
    int rc = make_plan(spec, batch, &pl);
    if (rc != MVX_OK) return rc;
    if (batch->num_mols == 0) return MVX_OK;
    if (!out) return fail(MVX_ERR_NULL_POINTER, "out is NULL");
    if (!workspace || workspace_bytes < pl.total) return fail(MVX_ERR_WORKSPACE, "workspace too small");
    if ((uintptr_t)workspace % kAlign != 0) return fail(MVX_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if ((uintptr_t)out % 16 != 0) return fail(MVX_ERR_BAD_SHAPE, "out must be 16-byte aligned");
    char* ws = (char*)workspace;
    int* status = (int*)(ws + pl.off_status);
    mvx::AtomRec* recs = (mvx::AtomRec*)(ws + pl.off_recs);
    uint32_t* colrange = (uint32_t*)(ws + pl.off_colrange);
    uint2* bins = (uint2*)(ws + pl.off_bins);
    uint32_t* lists = (uint32_t*)(ws + pl.off_lists);
    mvx::ColEntry* entries = (mvx::ColEntry*)(ws + pl.off_entries);
    const bool legacy_masks = pl.masks;   // CELLS form: cell masks precomputed by expand (<= 64 cells per column)

    const int B = batch->num_mols;
    const int64_t N = batch->total_atoms;
    const int C = batch->mode == MVX_MODE_SINGLE ? 1 : batch->num_channels;
    mvx_batch wide;   // compact feature rows: widened to fp32 once, everything downstream reads the fp32 copy
    const int feat_dtype = batch->features_dtype;   // as the caller passed them (the layered entry build widens compact rows itself)
    if (batch->mode == MVX_MODE_FEATURES && batch->features_dtype != MVX_F32 && N > 0 && !layered(pl.form)) {
        float* dst = (float*)(ws + pl.off_wide);
        const size_t n = (size_t)N * (size_t)C;
        mvx::mvx_widen_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(batch->features, batch->features_dtype == MVX_F16, n, dst);
        MVX_CUDA_OK(cudaGetLastError());
        wide = *batch;
        wide.features = dst;
        wide.features_dtype = MVX_F32;
        batch = &wide;
    }
    MVX_CUDA_OK(cudaMemsetAsync(status, 0, kAlign, st));
    if (layered(pl.form))   // key counts + key cursors of the layered binning
        MVX_CUDA_OK(cudaMemsetAsync(ws + pl.off_kcnt, 0, 2 * (size_t)batch->num_mols * pl.ncol * pl.nlayers * sizeof(uint32_t), st));
    prof_mark(st, 0);

    if (N > 0) {
        mvx::PrepParams pp;
        pp.g = pl.geo;
        pp.mode = batch->mode; pp.B = B; pp.C = C; pp.N = N;
        pp.mol_offsets = batch->mol_offsets;
        pp.coords = batch->coords; pp.coords_f64 = batch->coords_dtype == MVX_F64;
        pp.centers = batch->centers; pp.centers_f64 = batch->centers_dtype == MVX_F64;
        pp.types = batch->types; pp.radii = batch->radii;
        pp.tf_flags = batch->transform_flags; pp.transforms = batch->transform_flags ? batch->transforms : nullptr;
        pp.rng_seed = batch->rng_seed; pp.rng_offset = batch->rng_offset; pp.rng_translation = batch->random_translation;
        pp.recs = recs; pp.colrange = colrange; pp.status = status;
        pp.alayers = layered(pl.form) ? (uint32_t*)(ws + pl.off_alayers) : nullptr;
        pp.kcnt = layered(pl.form) ? (uint32_t*)(ws + pl.off_kcnt) : nullptr;
        pp.nzc = pl.nzc; pp.tz = pl.tz; pp.ncol = pl.ncol; pp.nl = pl.nlayers; pp.zl = pl.zl; pp.tau_lin = pl.tau_lin; pp.tau_quad = pl.tau_quad;
        const unsigned grid = (unsigned)((N + pt - 1) / pt);
        mvx::mvx_prep_kernel<<<grid, pt, 0, st>>>(pp);
        MVX_CUDA_OK(cudaGetLastError());
    }
    prof_mark(st, 1);
    if (layered(pl.form)) {   // per-layer entries (+ tile descriptors) straight from the per-atom column / layer words
        mvx::LBinParams lp;
        lp.res = pl.geo.res; lp.half_width = pl.geo.half_width; lp.sigma = spec->sigma;
        lp.tau_lin = pl.tau_lin; lp.tau_quad = pl.tau_quad;
        lp.B = B; lp.ncol = pl.ncol; lp.ncx = pl.geo.ncx; lp.maxcols = pl.maxcols; lp.zl = pl.zl; lp.nl = pl.nlayers;
        lp.nzc = pl.nzc; lp.tz = pl.tz; lp.dim = spec->dimension; lp.mode = batch->mode;
        lp.C = batch->mode == MVX_MODE_FEATURES ? C : 0; lp.es4 = pl.es4;
        lp.feat_dtype = batch->mode == MVX_MODE_FEATURES ? feat_dtype : MVX_F32;
        {   // four elements per load: 16 / 4 / 8-byte aligned rows for f32 / u8 / f16
            const uintptr_t al = lp.feat_dtype == MVX_U8 ? 4 : (lp.feat_dtype == MVX_F16 ? 8 : 16);
            lp.feat_vec = (C % 4 == 0) && ((uintptr_t)batch->features % al == 0);
        }
        lp.mol_offsets = batch->mol_offsets; lp.colrange = colrange; lp.alayers = (const uint32_t*)(ws + pl.off_alayers);
        lp.recs = recs; lp.types = batch->types; lp.features = (const float*)batch->features;
        lp.bins = bins; lp.lbins = (uint2*)(ws + pl.off_lbins);
        lp.tdesc = pl.form == FORM_PIPE ? (mvx::TileDesc*)(ws + pl.off_tdesc) : nullptr;
        lp.lent = (float4*)(ws + pl.off_lent);
        const size_t nkeys = (size_t)B * pl.ncol * pl.nlayers;
        lp.N = N; lp.kcnt = (const uint32_t*)(ws + pl.off_kcnt); lp.cursor = (uint32_t*)(ws + pl.off_kcnt) + nkeys;
        lp.lids = (uint32_t*)(ws + pl.off_lids);
        if ((nkeys + 7) / 8 > 0x7fffffffULL) return fail(MVX_ERR_BAD_SHAPE, "batch too large for one launch; split it");
        const size_t smem = 2 * (size_t)pl.ncol * sizeof(uint32_t);
        int sgroups = (pl.ncol + 7) / 8;   // small batches: spread each molecule's output over several CTAs
        if ((long long)B * sgroups > 1184) sgroups = (int)std::fmax(1.0, 1184.0 / B);
        mvx::mvx_lscan_kernel<<<(unsigned)(B * sgroups), 256, smem, st>>>(lp, sgroups);
        MVX_CUDA_OK(cudaGetLastError());
        if (N > 0) {
            mvx::mvx_lplace_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(lp);
            MVX_CUDA_OK(cudaGetLastError());
            static DeviceSet cfg_b;
            MVX_CUDA_OK(set_smem(mvx::mvx_lbuild_kernel, mvx::lbuild_smem_bytes(mvx::kLBuildStageQ), &cfg_b));
            mvx::mvx_lbuild_kernel<<<(unsigned)((nkeys + 7) / 8), 256, mvx::lbuild_smem_bytes(pl.es4), st>>>(lp);
            MVX_CUDA_OK(cudaGetLastError());
        }
    } else {
        mvx::BinParams bp;
        bp.B = B; bp.ncol = pl.ncol; bp.ncx = pl.geo.ncx; bp.maxcols = pl.maxcols;
        bp.mol_offsets = batch->mol_offsets; bp.colrange = colrange; bp.bins = bins; bp.lists = lists;
        const size_t smem = 2 * (size_t)pl.ncol * sizeof(uint32_t);
        const int groups = bin_groups(B, pl.ncol, N);
        mvx::ExpandParams ep;   // column lists -> staged-ready entries (CELLS form)
        ep.res = pl.geo.res; ep.half_width = pl.geo.half_width; ep.sigma = spec->sigma;
        ep.tau_lin = pl.tau_lin; ep.tau_quad = pl.tau_quad;
        ep.dim = spec->dimension; ep.ncx = pl.geo.ncx; ep.ncol = pl.ncol; ep.nzc = pl.nzc; ep.tz = pl.tz;
        ep.maxcols = pl.maxcols; ep.mode = batch->mode; ep.masks = legacy_masks; ep.B = B;
        ep.mol_offsets = batch->mol_offsets; ep.recs = recs; ep.bins = bins; ep.lists = lists;
        ep.types = batch->types; ep.entries = entries;
        const bool expand = pl.form == FORM_CELLS && N > 0;
        if (groups <= 1 && expand) {   // many small molecules: bin and expand in one launch
            // a handful of molecules (the reference's one-molecule-per-call pattern): one CTA each cannot fill the GPU, so
            // give it 32 warps — 2 of a 64^3 grid's 64 columns per warp instead of 8 (the call is latency, not throughput)
            const int ptb = (!split && B <= 74) ? 1024 : pt;
            mvx::mvx_bin_expand_kernel<<<(unsigned)B, ptb, smem, st>>>(bp, ep);
        } else if (groups <= 1) {
            mvx::mvx_bin_kernel<<<(unsigned)B, pt, smem, st>>>(bp);
        } else {   // few large molecules: spread each molecule's columns over several CTAs
            mvx::mvx_bin_count_kernel<<<(unsigned)(B * groups), 256, 0, st>>>(bp, groups);
            MVX_CUDA_OK(cudaGetLastError());
            mvx::mvx_bin_fill_kernel<<<(unsigned)(B * groups), 256, smem, st>>>(bp, groups);
        }
        MVX_CUDA_OK(cudaGetLastError());
        if (expand && groups > 1) {
            const long long warps = (long long)B * pl.ncol;
            mvx::mvx_expand_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(ep);
            MVX_CUDA_OK(cudaGetLastError());
        }
    }
    prof_mark(st, 2);
    if (split) {
        MVX_CUDA_OK(cudaEventRecord(bin_done, st));
        MVX_CUDA_OK(cudaStreamWaitEvent(st_vox, bin_done, 0));
        st = st_vox;
    }
    prof_mark(st, 3);
    {
        mvx::VoxParams vp;
        vp.lean = split ? 1 : 0;
        vp.res = pl.geo.res; vp.half_width = pl.geo.half_width; vp.sigma = spec->sigma;
        vp.tau_lin = pl.tau_lin; vp.tau_quad = pl.tau_quad;
        vp.dim = spec->dimension; vp.ncx = pl.geo.ncx; vp.ncol = pl.ncol; vp.nzc = pl.nzc; vp.tz = pl.tz;
        vp.C = C; vp.Cout = batch->out_channels; vp.maxcols = pl.maxcols; vp.cull = pl.geo.nb > 1;
        vp.mol_offsets = batch->mol_offsets; vp.recs = recs; vp.bins = bins; vp.lists = lists;
        vp.types = batch->types; vp.features = (const float*)batch->features; vp.chan_radii = nullptr; vp.out = out; vp.out_kind = batch->out_dtype;
        vp.clast = batch->out_layout == MVX_LAYOUT_DHWC;
        vp.entries = entries; vp.masks = legacy_masks;
        vp.nlayers = pl.nlayers; vp.zl = pl.zl; vp.es4 = pl.es4;
        vp.lent = (const float4*)(ws + pl.off_lent);
        vp.lbins = (const uint2*)(ws + pl.off_lbins);
        vp.tdesc = pl.form == FORM_PIPE ? (const mvx::TileDesc*)(ws + pl.off_tdesc) : nullptr;
        vp.pipe_sc = pl.pipe_sc; vp.pipe_q = pl.pipe_q; vp.ws_nb = pl.ws_nb;
        const unsigned long long nblk = (unsigned long long)B * pl.ncol * pl.nzc;
        if (nblk > 0x7fffffffULL) return fail(MVX_ERR_BAD_SHAPE, "batch too large for one launch; split it");
        const bool binary = spec->density_type == MVX_DENSITY_BINARY;
        const bool chan_feat = batch->mode == MVX_MODE_FEATURES && spec->radii_type == MVX_RADII_CHANNEL_WISE;
        if (batch->out_dtype == MVX_OUT_F64) {   // precision=64 (untuned): fp64 arithmetic and output
            mvx::VoxF64Params fp;
            fp.res = pl.geo.res; fp.half_width = pl.geo.half_width; fp.sigma = spec->sigma; fp.radius = batch->radius;
            fp.dim = spec->dimension; fp.ncx = pl.geo.ncx; fp.ncol = pl.ncol; fp.mode = batch->mode; fp.C = C;
            fp.Cout = batch->out_channels; fp.maxcols = pl.maxcols; fp.binary = binary;
            fp.scalar_radius = spec->radii_type == MVX_RADII_SCALAR;
            fp.clast = batch->out_layout == MVX_LAYOUT_DHWC;
            fp.mol_offsets = batch->mol_offsets; fp.recs = recs; fp.bins = bins; fp.lists = lists;
            fp.types = batch->types; fp.features = (const float*)batch->features; fp.chan_radii = chan_feat ? batch->radii : nullptr;
            fp.out = (double*)out;
            const unsigned long long nb64 = (unsigned long long)B * pl.ncol;
            if (nb64 > 0x7fffffffULL) return fail(MVX_ERR_BAD_SHAPE, "batch too large for one launch; split it");
            mvx::mvx_voxelize_f64_kernel<<<(unsigned)nb64, 256, 0, st>>>(fp);
            MVX_CUDA_OK(cudaGetLastError());
        } else if (chan_feat) {   // per-channel radius: one pass per channel (numpy/voxelizer.py:213-224)
            for (int c = 0; c < C; ++c) {
                vp.c_begin = c; vp.c_end = c + 1; vp.chan_radii = batch->radii;
                MVX_CUDA_OK(launch_vox(batch->mode, 1, vp, binary, pl.form, pl.nv, (unsigned)nblk, st));
            }
        } else {
            vp.c_begin = 0; vp.c_end = batch->out_channels;
            MVX_CUDA_OK(launch_vox(batch->mode, pick_chunk(batch->mode, batch->out_channels), vp, binary, pl.form, pl.nv,
                                   (unsigned)nblk, st));
        }
    }
    prof_mark(st, 4);
    if (g_prof.active && g_prof.calls < g_prof.max_calls) ++g_prof.calls;
    return MVX_OK;
}
}  // namespace

extern "C" {

int mvx_voxelize(const mvx_grid_spec* spec, const mvx_batch* batch, void* out, void* workspace,
                 size_t workspace_bytes, void* stream) {
    return enqueue(spec, batch, out, workspace, workspace_bytes, (cudaStream_t)stream, (cudaStream_t)stream, nullptr);
}

int mvx_voxelize_split(const mvx_grid_spec* spec, const mvx_batch* batch, void* out, void* workspace,
                       size_t workspace_bytes, void* bin_stream, void* vox_stream, void* bin_done_event) {
    if (bin_stream == vox_stream) return enqueue(spec, batch, out, workspace, workspace_bytes, (cudaStream_t)vox_stream, (cudaStream_t)vox_stream, nullptr);
    if (!bin_done_event) return fail(MVX_ERR_NULL_POINTER, "bin_done_event is NULL");
    return enqueue(spec, batch, out, workspace, workspace_bytes, (cudaStream_t)bin_stream, (cudaStream_t)vox_stream, (cudaEvent_t)bin_done_event);
}

int mvx_random_transforms(uint64_t rng_seed, uint64_t rng_offset, int32_t num_mols, int32_t transform_flags,
                          double random_translation, double* out, void* stream) {
    if (num_mols < 0) return fail(MVX_ERR_BAD_SHAPE, "negative batch size");
    if (num_mols == 0) return MVX_OK;
    if (!out) return fail(MVX_ERR_NULL_POINTER, "out is NULL");
    if (transform_flags & ~(MVX_TF_ROTATE | MVX_TF_TRANSLATE | MVX_TF_TRANSLATE_ONCE)) return fail(MVX_ERR_BAD_ENUM, "transform_flags");
    mvx::DrawParams dp;
    dp.seed = rng_seed; dp.offset = rng_offset; dp.B = num_mols; dp.flags = transform_flags; dp.rt = random_translation; dp.out = out;
    mvx::mvx_draw_transforms_kernel<<<(unsigned)((num_mols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dp);
    MVX_CUDA_OK(cudaGetLastError());
    return MVX_OK;
}

int mvx_synth_ligands(uint64_t seed, uint64_t first_mol, int32_t num_mols, int32_t vmin, int32_t vmax, int32_t num_types,
                      double step, const int32_t* mol_offsets, int32_t* counts, void* coords, int32_t coords_dtype,
                      int32_t* types, void* stream) {
    if (num_mols < 0 || vmin < 1 || vmax < vmin || num_types < 1) return fail(MVX_ERR_BAD_SHAPE, "bad synthetic-ligand shape");
    if (num_mols == 0) return MVX_OK;
    if (!mol_offsets && !counts) return fail(MVX_ERR_NULL_POINTER, "counts / mol_offsets");
    if (mol_offsets && !coords) return fail(MVX_ERR_NULL_POINTER, "coords");
    if (coords_dtype != MVX_F32 && coords_dtype != MVX_F64) return fail(MVX_ERR_BAD_ENUM, "coords_dtype");
    mvx::SynthParams sp;
    sp.seed = seed; sp.first_mol = first_mol; sp.B = num_mols; sp.vmin = vmin; sp.vmax = vmax; sp.num_types = num_types;
    sp.coords_f64 = coords_dtype == MVX_F64; sp.step = step; sp.mol_offsets = mol_offsets; sp.counts = counts;
    sp.coords = coords; sp.types = types;
    mvx::mvx_synth_ligands_kernel<<<(unsigned)((num_mols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(sp);
    MVX_CUDA_OK(cudaGetLastError());
    return MVX_OK;
}

int mvx_compact_bricks(const mvx_grid_spec* spec, const mvx_batch* batch, const void* grids, const void* workspace,
                       uint32_t* brick_ids, float* brick_vals, uint32_t capacity, uint32_t* count, void* stream) {
    Plan pl;
    int rc = make_plan(spec, batch, &pl);
    if (rc != MVX_OK) return rc;
    if (batch->out_dtype != MVX_OUT_F32) return fail(MVX_ERR_UNSUPPORTED, "brick compaction needs float32 grids");
    if (batch->out_layout != MVX_LAYOUT_CDHW) return fail(MVX_ERR_UNSUPPORTED, "brick compaction needs the (B, C, D, H, W) layout");
    if (!count) return fail(MVX_ERR_NULL_POINTER, "count is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    MVX_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(uint32_t), st));
    if (batch->num_mols == 0) return MVX_OK;
    if (!grids || (capacity > 0 && (!brick_ids || !brick_vals))) return fail(MVX_ERR_NULL_POINTER, "grids / brick buffers");
    if ((uintptr_t)grids % 16 != 0 || (uintptr_t)brick_vals % 16 != 0) return fail(MVX_ERR_BAD_SHAPE, "grids and brick_vals must be 16-byte aligned");
    mvx::CompactParams cp;
    cp.dim = spec->dimension; cp.ncx = pl.geo.ncx; cp.ncol = pl.ncol; cp.Cout = batch->out_channels;
    cp.nbz = (spec->dimension + mvx::kBrick - 1) / mvx::kBrick; cp.B = batch->num_mols;
    cp.bins = workspace ? (const uint2*)((const char*)workspace + pl.off_bins) : nullptr;
    cp.grid = (const float*)grids; cp.ids = brick_ids; cp.vals = brick_vals; cp.cap = capacity; cp.count = count;
    const unsigned long long nblk = (unsigned long long)batch->num_mols * pl.ncol;
    const unsigned long long nbricks = nblk * cp.Cout * cp.nbz;
    if (nblk > 0x7fffffffULL || nbricks > 0xffffffffULL) return fail(MVX_ERR_BAD_SHAPE, "batch too large for one launch; split it");
    mvx::mvx_compact_bricks_kernel<<<(unsigned)nblk, 256, 0, st>>>(cp);
    MVX_CUDA_OK(cudaGetLastError());
    return MVX_OK;
}

int mvx_profile_begin(int max_calls) {
    if (g_prof.active) return fail(MVX_ERR_UNSUPPORTED, "a profile is already open on this thread");
    if (max_calls < 1) return fail(MVX_ERR_BAD_SHAPE, "max_calls must be >= 1");
    g_prof.ev = new cudaEvent_t[(size_t)max_calls * 5];
    for (int i = 0; i < max_calls * 5; ++i) MVX_CUDA_OK(cudaEventCreate(&g_prof.ev[i]));
    g_prof.max_calls = max_calls; g_prof.calls = 0; g_prof.active = true;
    return MVX_OK;
}

int mvx_profile_end(double* ms_prep, double* ms_bin, double* ms_voxelize, int* num_calls) {
    if (!g_prof.active) return fail(MVX_ERR_UNSUPPORTED, "no profile is open on this thread");
    double t[3] = {0, 0, 0};
    int rc = MVX_OK;
    if (g_prof.calls > 0) {
        if (cudaEventSynchronize(g_prof.ev[(g_prof.calls - 1) * 5 + 4]) != cudaSuccess) rc = MVX_ERR_CUDA;
        static const int first[3] = {0, 1, 3};   // prep = e0..e1, bin = e1..e2, voxelize = e3..e4
        for (int c = 0; c < g_prof.calls && rc == MVX_OK; ++c)
            for (int k = 0; k < 3; ++k) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, g_prof.ev[c * 5 + first[k]], g_prof.ev[c * 5 + first[k] + 1]) != cudaSuccess) rc = MVX_ERR_CUDA;
                t[k] += ms;
            }
    }
    for (int i = 0; i < g_prof.max_calls * 5; ++i) cudaEventDestroy(g_prof.ev[i]);
    delete[] g_prof.ev;
    if (ms_prep) *ms_prep = t[0];
    if (ms_bin) *ms_bin = t[1];
    if (ms_voxelize) *ms_voxelize = t[2];
    if (num_calls) *num_calls = g_prof.calls;
    g_prof = Profile();
    if (rc != MVX_OK) return fail(rc, "cudaEvent timing failed");
    return MVX_OK;
}

namespace {
// one pinned word per calling thread for the status read-back (a pageable destination makes the copy a staged,
// synchronous driver round trip)
struct PinnedWord {
    int* p = nullptr;
    int* get() {
        if (p == nullptr && cudaHostAlloc((void**)&p, sizeof(int), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
        return p;
    }
    ~PinnedWord() { if (p != nullptr) cudaFreeHost(p); }
};
}  // namespace

int mvx_check_status(void* workspace, void* stream) {
    if (!workspace) return fail(MVX_ERR_NULL_POINTER, "workspace is NULL");
    thread_local PinnedWord pw;
    int pageable = 0;
    int* dst = pw.get() != nullptr ? pw.get() : &pageable;
    MVX_CUDA_OK(cudaMemcpyAsync(dst, workspace, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MVX_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    const int flags = *dst;
    if (flags & mvx::kFlagBadType) return fail(MVX_ERR_DEVICE_FLAG, "a type index is outside [0, num_channels)");
    if (flags & mvx::kFlagRadiusOverMax) return fail(MVX_ERR_DEVICE_FLAG, "a radius exceeds max_radius");
    return MVX_OK;
}

namespace {
// Pinned host block for packing the inputs of small host calls (one per calling thread, grown on demand; falls back to
// per-array copies when page-locking fails).
constexpr size_t kPackBytes = 1u << 20;
struct HostPack {
    char* p = nullptr;
    size_t cap = 0;
    char* get(size_t bytes) {
        if (bytes > cap) {
            if (p != nullptr) cudaFreeHost(p);
            p = nullptr; cap = 0;
            const size_t want = bytes < 65536 ? 65536 : bytes;
            if (cudaHostAlloc((void**)&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
            cap = want;
        }
        return p;
    }
    ~HostPack() { if (p != nullptr) cudaFreeHost(p); }   // at thread exit; an error after context teardown is ignored
};
size_t feature_bytes(int dtype) { return dtype == MVX_U8 ? 1 : (dtype == MVX_F16 ? 2 : 4); }
struct Staging { size_t off_offs, off_coords, off_centers, off_types, off_features, off_radii, off_transforms, total; };
void plan_staging(const mvx_grid_spec* s, const mvx_batch* b, Staging* sg) {
    const size_t N = (size_t)b->total_atoms, B = (size_t)b->num_mols;
    const size_t C = b->mode == MVX_MODE_SINGLE ? 1 : (size_t)b->num_channels;
    size_t off = 0;
    sg->off_offs = off;     off += align_up((B + 1) * sizeof(int32_t));
    sg->off_coords = off;   off += align_up(N * 3 * (b->coords_dtype == MVX_F64 ? 8 : 4));
    sg->off_centers = off;  off += align_up(b->centers ? B * 3 * (b->centers_dtype == MVX_F64 ? 8 : 4) : 0);
    sg->off_types = off;    off += align_up(b->mode == MVX_MODE_TYPES ? N * sizeof(int32_t) : 0);
    sg->off_features = off; off += align_up(b->mode == MVX_MODE_FEATURES ? N * C * feature_bytes(b->features_dtype) : 0);
    size_t nr = s->radii_type == MVX_RADII_ATOM_WISE ? N : (s->radii_type == MVX_RADII_CHANNEL_WISE ? C : 0);
    sg->off_radii = off;    off += align_up(nr * sizeof(float));
    sg->off_transforms = off; off += align_up((b->transforms && b->transform_flags) ? B * 7 * sizeof(double) : 0);
    sg->total = off;
}
}  // namespace

int mvx_host_staging_bytes(const mvx_grid_spec* spec, const mvx_batch* batch, size_t* out_bytes) {
    if (!out_bytes) return fail(MVX_ERR_NULL_POINTER, "out_bytes is NULL");
    int rc = check_args(spec, batch);
    if (rc != MVX_OK) return rc;
    Staging sg;
    plan_staging(spec, batch, &sg);
    *out_bytes = sg.total;
    return MVX_OK;
}

int mvx_voxelize_host(const mvx_grid_spec* spec, const mvx_batch* hb, void* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
    Plan pl;
    int rc = make_plan(spec, hb, &pl);
    if (rc != MVX_OK) return rc;
    if (hb->num_mols == 0) return MVX_OK;
    Staging sg;
    plan_staging(spec, hb, &sg);
    if (!workspace || workspace_bytes < pl.total + sg.total) return fail(MVX_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* dv = (char*)workspace + pl.total;
    const size_t N = (size_t)hb->total_atoms, B = (size_t)hb->num_mols;
    const size_t C = hb->mode == MVX_MODE_SINGLE ? 1 : (size_t)hb->num_channels;
    mvx_batch db = *hb;
    // Small calls (the reference's one-molecule-per-call pattern): the inputs are packed into one pinned staging block
    // and cross PCIe as ONE copy instead of up to seven pageable ones (each a synchronous driver round trip).
    // The block is reused: this function ends with a stream synchronisation (mvx_check_status).
    char* pack = nullptr;
    if (sg.total <= kPackBytes) {
        thread_local HostPack hp;
        pack = hp.get(sg.total);
    }
    auto put = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        if (bytes == 0) return cudaSuccess;
        if (pack != nullptr) { std::memcpy(pack + off, src, bytes); return cudaSuccess; }
        return cudaMemcpyAsync(dv + off, src, bytes, cudaMemcpyHostToDevice, st);
    };
    MVX_CUDA_OK(put(sg.off_offs, hb->mol_offsets, (B + 1) * sizeof(int32_t)));
    db.mol_offsets = (const int32_t*)(dv + sg.off_offs);
    if (N > 0) {
        MVX_CUDA_OK(put(sg.off_coords, hb->coords, N * 3 * (hb->coords_dtype == MVX_F64 ? 8 : 4)));
        db.coords = dv + sg.off_coords;
    }
    if (hb->centers) {
        MVX_CUDA_OK(put(sg.off_centers, hb->centers, B * 3 * (hb->centers_dtype == MVX_F64 ? 8 : 4)));
        db.centers = dv + sg.off_centers;
    }
    if (hb->mode == MVX_MODE_TYPES && N > 0) {
        MVX_CUDA_OK(put(sg.off_types, hb->types, N * sizeof(int32_t)));
        db.types = (const int32_t*)(dv + sg.off_types);
    }
    if (hb->mode == MVX_MODE_FEATURES && N > 0) {
        MVX_CUDA_OK(put(sg.off_features, hb->features, N * C * feature_bytes(hb->features_dtype)));
        db.features = (const float*)(dv + sg.off_features);
    }
    if (spec->radii_type != MVX_RADII_SCALAR && hb->radii) {
        size_t nr = spec->radii_type == MVX_RADII_ATOM_WISE ? N : C;
        if (nr > 0) {
            MVX_CUDA_OK(put(sg.off_radii, hb->radii, nr * sizeof(float)));
            db.radii = (const float*)(dv + sg.off_radii);
        }
    }
    if (hb->transforms && hb->transform_flags) {
        MVX_CUDA_OK(put(sg.off_transforms, hb->transforms, B * 7 * sizeof(double)));
        db.transforms = (const double*)(dv + sg.off_transforms);
    }
    if (pack != nullptr) MVX_CUDA_OK(cudaMemcpyAsync(dv, pack, sg.total, cudaMemcpyHostToDevice, st));
    rc = mvx_voxelize(spec, &db, out, workspace, pl.total, stream);
    if (rc != MVX_OK) return rc;
    return mvx_check_status(workspace, stream);
}

}  // extern "C"
