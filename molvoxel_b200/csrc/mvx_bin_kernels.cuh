// mvx_bin_kernels.cuh — per-atom prep, column / layer binning, entry build, precision=64 and feature widening kernels
// (compiled into the API translation unit only; see mvx_common.cuh for the path map and data layout).
#pragma once
#include "mvx_common.cuh"
#include "mvx_rigid.cuh"

namespace mvx {

// ---------------------------------------------------------------------------------------------
// prep: one thread per atom.  Everything the reference decides per atom in fp64 is decided here,
// with explicit round-to-nearest intrinsics so nvcc cannot contract a*b+c into an FMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double load_coord(const void* base, int is_f64, int64_t i) {
    return is_f64 ? reinterpret_cast<const double*>(base)[i] : (double)reinterpret_cast<const float*>(base)[i];
}

__global__ void __launch_bounds__(256) mvx_prep_kernel(const PrepParams P) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.N) return;
    const Geo& g = P.g;

    int lo = 0, hi = P.B;   // offs[lo] <= n < offs[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((int64_t)P.mol_offsets[mid] <= n) lo = mid; else hi = mid;
    }
    const int mol = lo;

    // centring in numpy's promoted dtype (fp32 - fp32 stays fp32, numpy/voxelizer.py:263), then the optional rigid
    // transform in that same dtype (:265), then fp64 (:268)
    double p[3];
    const bool all_f32 = !P.coords_f64 && (P.centers == nullptr || !P.centers_f64);
    Rigid R;
    if (P.tf_flags != 0) {
        if (P.transforms != nullptr) {
            const double* T = P.transforms + 7 * (int64_t)mol;
            R.q[0] = T[0]; R.q[1] = T[1]; R.q[2] = T[2]; R.q[3] = T[3];
            R.t[0] = T[4]; R.t[1] = T[5]; R.t[2] = T[6];
        } else {
            draw_rigid(P.rng_seed, P.rng_offset + (unsigned long long)mol, P.tf_flags, P.rng_translation, R);
        }
    }
    if (all_f32) {
        float pf[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pf[k] = reinterpret_cast<const float*>(P.coords)[3 * n + k];
            if (P.centers != nullptr) pf[k] = __fsub_rn(pf[k], reinterpret_cast<const float*>(P.centers)[3 * (int64_t)mol + k]);
        }
        if (P.tf_flags != 0) apply_rigid<float>(pf, R, P.tf_flags);
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = (double)pf[k];
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            p[k] = load_coord(P.coords, P.coords_f64, 3 * n + k);
            if (P.centers != nullptr) p[k] = __dsub_rn(p[k], load_coord(P.centers, P.centers_f64, 3 * (int64_t)mol + k));
        }
        if (P.tf_flags != 0) apply_rigid<double>(p, R, P.tf_flags);
    }

    bool keep = true;
    int type = 0;
    if (P.mode == 1) {
        type = P.types[n];
        if (type < 0 || type >= P.C) { atomicOr(P.status, kFlagBadType); keep = false; type = 0; }
    }
    float r32;
    if (g.radii_src == 0) r32 = g.r_scalar32;
    else if (g.radii_src == 1) r32 = P.radii[n];
    else if (g.radii_src == 2) r32 = P.radii[type];                // radii[types] (numpy/voxelizer.py:284-285)
    else r32 = (float)g.size_scalar;
    const double a = g.scalar_form ? g.size_scalar : (double)r32;  // "atom_size"

    // global clip, strict (numpy/voxelizer.py:481-494)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (g.scalar_form) keep = keep && (p[k] > g.clip_lo) && (p[k] < g.clip_hi);
        else keep = keep && (__dadd_rn(p[k], a) > g.lower) && (__dsub_rn(p[k], a) < g.upper);
    }

    // block cull (numpy/voxelizer.py:55, :504-511): block b >= 1 keeps the atom iff p > bounds[b-1] - a.
    // The test is monotone in b; the first failing block's first voxel plane is the only place where a
    // true hit is lost, so the cull reduces to one forbidden plane per axis.
    int16_t forb[3] = {-1, -1, -1};
    if (g.nb > 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            for (int b = 1; b < g.nb; ++b) {
                double bound = __dadd_rn(__dsub_rn(__dmul_rn((double)(b * g.bd), g.res), g.half_width), g.res_half);
                double thr = __dsub_rn(bound, a);
                if (!(p[k] > thr)) { forb[k] = (int16_t)(b * g.bd); break; }
            }
        }
    }

    // conservative voxel ranges (speed only; 0.01 voxel of slack covers every rounding in play)
    double reach = a > (double)r32 ? a : (double)r32;
    reach = reach * (1.0 + 1e-6);
    int v0[3], v1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double t0 = floor((p[k] - reach + g.half_width) * g.inv_res - 0.01);
        double t1 = ceil((p[k] + reach + g.half_width) * g.inv_res + 0.01);
        t0 = t0 < 0.0 ? 0.0 : t0;
        t1 = t1 > (double)(g.dim - 1) ? (double)(g.dim - 1) : t1;
        if (!(t0 <= t1)) keep = false;   // also catches NaN
        v0[k] = keep ? (int)t0 : 0;
        v1[k] = keep ? (int)t1 : 0;
    }

    uint32_t cr = 0x000000FFu;   // cx0 = 255 > cx1 = 0: overlaps nothing
    if (keep) {
        int cx0 = v0[0] / kTile, cx1 = v1[0] / kTile, cy0 = v0[1] / kTile, cy1 = v1[1] / kTile;
        if (cx1 - cx0 + 1 > g.cols_axis_max) { atomicOr(P.status, kFlagRadiusOverMax); cx1 = cx0 + g.cols_axis_max - 1; }
        if (cy1 - cy0 + 1 > g.cols_axis_max) { atomicOr(P.status, kFlagRadiusOverMax); cy1 = cy0 + g.cols_axis_max - 1; }
        cr = (uint32_t)cx0 | ((uint32_t)cx1 << 8) | ((uint32_t)cy0 << 16) | ((uint32_t)cy1 << 24);
    }
    AtomRec rec;
    rec.px = p[0]; rec.py = p[1]; rec.pz = p[2];
    rec.r = r32;
    rec.fx = forb[0]; rec.fy = forb[1]; rec.fz = forb[2];
    rec.zlo = (int16_t)v0[2]; rec.zhi = (int16_t)v1[2];
    rec.pad = 0;
    P.recs[n] = rec;
    P.colrange[n] = cr;
    if (P.alayers != nullptr) {
        uint32_t m = 0u;
        if (keep) {
            const float resf = (float)g.res;
            const float r2 = r32 * r32;
            const float r2hi = r2 + (r32 * P.tau_lin + r2 * P.tau_quad);
            const float lim = r2hi + 1e-4f * (1.f + r2hi);
            const float az = (float)(p[2] + g.half_width);
            const int ncz_max = (P.tz + kCellZ - 1) / kCellZ;
            for (int zc = 0; zc < P.nzc; ++zc) {
                const int z0 = zc * P.tz, z1 = min(g.dim, z0 + P.tz);
                for (int cz = 0; cz < ncz_max; ++cz) {
                    const int lo = z0 + cz * kCellZ, hi = min(lo + kCellZ, z1);
                    if (lo >= hi || v1[2] < lo || v0[2] >= hi) continue;
                    const float bhz = 0.5f * (hi - lo - 1) * resf;
                    const float ez = fmaxf(fabsf(az - (lo * resf + bhz)) - bhz, 0.f);
                    if (ez * ez <= lim) m |= 1u << (zc * ncz_max + cz);
                }
            }
        }
        if (__popc(m) > P.zl) {   // more layers than the workspace reserves per atom: a radius above max_radius
            atomicOr(P.status, kFlagRadiusOverMax);
            while (__popc(m) > P.zl) m &= m - 1u;
        }
        P.alayers[n] = m;
        if (m != 0u) {   // keep is true: count this atom under every (column, layer) key it belongs to
            const int cx0 = cr & 0xFF, cx1 = (cr >> 8) & 0xFF, cy0 = (cr >> 16) & 0xFF, cy1 = (cr >> 24) & 0xFF;
            for (int cx = cx0; cx <= cx1; ++cx)
                for (int cy = cy0; cy <= cy1; ++cy) {
                    uint32_t* k = P.kcnt + ((size_t)mol * P.ncol + (size_t)(cx * g.ncx + cy)) * P.nl;
                    for (uint32_t b = m; b != 0u; b &= b - 1u) atomicAdd(k + (__ffs((int)b) - 1), 1u);
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bin: one CTA per molecule.  Warp w owns columns w, w+nwarps, ...; lanes stride over the atoms
// 32 at a time and compact with ballot + popc, so each list keeps ascending atom order (the
// reference's fp32 summation order, numpy/voxelizer.py:364-365) without atomics.
// pass 1 counts, a CTA-wide exclusive prefix sum places the lists, pass 2 fills.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool col_overlaps(uint32_t cr, int cx, int cy) {
    int cx0 = cr & 0xFF, cx1 = (cr >> 8) & 0xFF, cy0 = (cr >> 16) & 0xFF, cy1 = (cr >> 24) & 0xFF;
    return cx >= cx0 && cx <= cx1 && cy >= cy0 && cy <= cy1;
}

// Counts (and, when `seg` is given, fills) one column's list.  128 atoms per iteration: four
// independent loads in flight per lane, ballots keep ascending atom order.
__device__ __forceinline__ uint32_t bin_scan_column(const uint32_t* __restrict__ cr, int V, int cx, int cy, int lane,
                                                    uint32_t* seg, uint32_t pos, int a0) {
    uint32_t cnt = 0;
    for (int base = 0; base < V; base += 128) {
        uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * 32 + lane;
            c[u] = i < V ? __ldg(cr + i) : 0x000000FFu;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool in = col_overlaps(c[u], cx, cy);
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (seg != nullptr && in) seg[pos + cnt + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(a0 + base + u * 32 + lane);
            cnt += __popc(m);
        }
    }
    return cnt;
}

// exclusive scan of s_cnt[0..n) into s_off[0..n) by warp 0
__device__ __forceinline__ void bin_scan_counts(const uint32_t* s_cnt, uint32_t* s_off, int n, int lane) {
    uint32_t carry = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const uint32_t v = i < n ? s_cnt[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (i < n) s_off[i] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
}

// Fused form (many molecules): one CTA per molecule does count, scan and fill.
__global__ void mvx_bin_kernel(const BinParams P) {
    extern __shared__ uint32_t s_u32[];
    uint32_t* s_cnt = s_u32;            // [ncol]
    uint32_t* s_off = s_u32 + P.ncol;   // [ncol]
    const int mol = blockIdx.x;
    const int a0 = P.mol_offsets[mol], V = P.mol_offsets[mol + 1] - a0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t* cr = P.colrange + a0;

    for (int col = warp; col < P.ncol; col += nwarps) {
        const uint32_t cnt = bin_scan_column(cr, V, col / P.ncx, col % P.ncx, lane, nullptr, 0, a0);
        if (lane == 0) s_cnt[col] = cnt;
    }
    __syncthreads();
    if (warp == 0) bin_scan_counts(s_cnt, s_off, P.ncol, lane);
    __syncthreads();
    uint32_t* seg = P.lists + (size_t)a0 * (size_t)P.maxcols;
    for (int col = warp; col < P.ncol; col += nwarps) {
        if (s_cnt[col] != 0) bin_scan_column(cr, V, col / P.ncx, col % P.ncx, lane, seg, s_off[col], a0);
        if (lane == 0) P.bins[(size_t)mol * P.ncol + col] = make_uint2(s_off[col], s_cnt[col]);
    }
}

// Split form (few, large molecules): grid = (molecule, column group) so that small batches still fill
// the 148 SMs.  Pass 1 counts into bins[].y; pass 2 rescans the molecule's counts and fills its group.
__global__ void mvx_bin_count_kernel(const BinParams P, int groups) {
    const int mol = blockIdx.x / groups, grp = blockIdx.x % groups;
    const int a0 = P.mol_offsets[mol], V = P.mol_offsets[mol + 1] - a0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t* cr = P.colrange + a0;
    for (int col = grp * nwarps + warp; col < P.ncol; col += groups * nwarps) {
        const uint32_t cnt = bin_scan_column(cr, V, col / P.ncx, col % P.ncx, lane, nullptr, 0, a0);
        if (lane == 0) P.bins[(size_t)mol * P.ncol + col].y = cnt;
    }
}

__global__ void mvx_bin_fill_kernel(const BinParams P, int groups) {
    extern __shared__ uint32_t s_u32[];
    uint32_t* s_cnt = s_u32;
    uint32_t* s_off = s_u32 + P.ncol;
    const int mol = blockIdx.x / groups, grp = blockIdx.x % groups;
    const int a0 = P.mol_offsets[mol], V = P.mol_offsets[mol + 1] - a0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t* cr = P.colrange + a0;
    for (int col = threadIdx.x; col < P.ncol; col += blockDim.x) s_cnt[col] = P.bins[(size_t)mol * P.ncol + col].y;
    __syncthreads();
    if (warp == 0) bin_scan_counts(s_cnt, s_off, P.ncol, lane);
    __syncthreads();
    uint32_t* seg = P.lists + (size_t)a0 * (size_t)P.maxcols;
    for (int col = grp * nwarps + warp; col < P.ncol; col += groups * nwarps) {
        if (s_cnt[col] != 0) bin_scan_column(cr, V, col / P.ncx, col % P.ncx, lane, seg, s_off[col], a0);
        if (lane == 0) P.bins[(size_t)mol * P.ncol + col].x = s_off[col];
    }
}

// ---------------------------------------------------------------------------------------------
// expand: one warp per (molecule, column), one lane per list entry.  Turns the atom ids of a column list
// into ColEntry records (column-relative fp32 position from the fp64 record, cutoff band, Gaussian
// coefficient, packed forbidden planes, and the mask of warp cells the cutoff sphere reaches), so that the
// voxelize kernel stages a column with one coalesced copy and no per-atom arithmetic.
// ---------------------------------------------------------------------------------------------
constexpr int kExpandSmemMasks = 512;   // masks of the first 512 entries of a column stay in shared memory

// One column's list -> ColEntry records, by one warp (lane = list entry).
__device__ __forceinline__ void expand_column(const ExpandParams& P, const int mol, const int col, const uint2 bin, const int lane) {
    if (bin.y == 0) return;
    const size_t base = (size_t)P.mol_offsets[mol] * (size_t)P.maxcols + bin.x;
    const int x0 = (col / P.ncx) * kTile, y0 = (col % P.ncx) * kTile;
    const double ox0 = (double)x0 * P.res - P.half_width, oy0 = (double)y0 * P.res - P.half_width;
    const float resf = (float)P.res;
    const float bhx = 0.5f * (kCellX - 1) * resf, bhy = 0.5f * (kCellY - 1) * resf;
    const int ncz_max = (P.tz + kCellZ - 1) / kCellZ;
    for (uint32_t i = lane; i < bin.y; i += 32) {
        const uint32_t n = P.lists[base + i];
        const AtomRec rec = P.recs[n];
        const float r = rec.r;
        const float r2 = r * r;
        const float tau = r * P.tau_lin + r2 * P.tau_quad;
        ColEntry e;
        e.ax = (float)(rec.px - ox0); e.ay = (float)(rec.py - oy0); e.az = (float)(rec.pz + P.half_width);
        e.r2hi = r2 + tau; e.r2lo = r2 - tau;
        const double rs = (double)r * P.sigma;
        e.kc = (float)(-0.5 * 1.4426950408889634 / (rs * rs));
        const int fx = rec.fx - x0, fy = rec.fy - y0;
        e.forb = (uint32_t)((rec.fx >= 0 && fx >= 0 && fx < kTile) ? fx : 0xFF) |
                 ((uint32_t)((rec.fy >= 0 && fy >= 0 && fy < kTile) ? fy : 0xFF) << 8) |
                 ((uint32_t)(rec.fz >= 0 ? rec.fz : 0xFFFF) << 16);
        e.type_or_r = (P.mode == 1) ? (uint32_t)P.types[n] : __float_as_uint(r);
        e.n = n; e.pad = 0;
        unsigned long long mask = 0ull;
        if (P.masks) {
            const float lim = e.r2hi + 1e-4f * (1.f + e.r2hi);
            for (int zc = 0; zc < P.nzc; ++zc) {
                const int z0 = zc * P.tz, z1 = min(P.dim, z0 + P.tz);
                if (rec.zhi < z0 || rec.zlo >= z1) continue;
                for (int cz = 0; cz < ncz_max; ++cz) {
                    const int lo = z0 + cz * kCellZ;
                    if (lo >= z1) break;
                    const float bhz = 0.5f * (min(kCellZ, z1 - lo) - 1) * resf;
                    const float ez = fmaxf(fabsf(e.az - (lo * resf + bhz)) - bhz, 0.f);
                    const float ez2 = ez * ez;
                    if (ez2 > lim) continue;
                    const int layer = zc * ncz_max + cz;
#pragma unroll
                    for (int ix = 0; ix < kTile / kCellX; ++ix) {
                        const float ex = fmaxf(fabsf(e.ax - ((ix * kCellX) * resf + bhx)) - bhx, 0.f);
                        const float exz = fmaf(ex, ex, ez2);
                        if (exz > lim) continue;
#pragma unroll
                        for (int iy = 0; iy < kTile / kCellY; ++iy) {
                            const float ey = fmaxf(fabsf(e.ay - ((iy * kCellY) * resf + bhy)) - bhy, 0.f);
                            if (fmaf(ey, ey, exz) <= lim) mask |= 1ull << (layer * kCellsXY + ix * (kTile / kCellY) + iy);
                        }
                    }
                }
            }
        }
        e.mask_lo = (uint32_t)mask; e.mask_hi = (uint32_t)(mask >> 32);
        float4* dst = reinterpret_cast<float4*>(P.entries + base + i);
        const float4* src = reinterpret_cast<const float4*>(&e);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

__global__ void __launch_bounds__(256) mvx_expand_kernel(const ExpandParams P) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (long long)P.B * P.ncol) return;
    expand_column(P, (int)(gw / P.ncol), (int)(gw % P.ncol), P.bins[gw], lane);
}

// Fused form for many small molecules (ligand batches): one CTA per molecule counts, scans, fills its column lists and
// expands each of them right away (the warp that filled a list turns it into entries), one launch instead of two.
__global__ void __launch_bounds__(1024) mvx_bin_expand_kernel(const BinParams P, const ExpandParams E) {
    extern __shared__ uint32_t s_u32[];
    uint32_t* s_cnt = s_u32;            // [ncol]
    uint32_t* s_off = s_u32 + P.ncol;   // [ncol]
    const int mol = blockIdx.x;
    const int a0 = P.mol_offsets[mol], V = P.mol_offsets[mol + 1] - a0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t* cr = P.colrange + a0;
    for (int col = warp; col < P.ncol; col += nwarps) {
        const uint32_t cnt = bin_scan_column(cr, V, col / P.ncx, col % P.ncx, lane, nullptr, 0, a0);
        if (lane == 0) s_cnt[col] = cnt;
    }
    __syncthreads();
    if (warp == 0) bin_scan_counts(s_cnt, s_off, P.ncol, lane);
    __syncthreads();
    uint32_t* seg = P.lists + (size_t)a0 * (size_t)P.maxcols;
    for (int col = warp; col < P.ncol; col += nwarps) {
        const uint2 bin = make_uint2(s_off[col], s_cnt[col]);
        if (bin.y != 0) bin_scan_column(cr, V, col / P.ncx, col % P.ncx, lane, seg, bin.x, a0);
        if (lane == 0) P.bins[(size_t)mol * P.ncol + col] = bin;
        __syncwarp();   // the list written by this warp's lanes is read back by other lanes below
        expand_column(E, mol, col, bin, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// layered binning (feeds the tile and pipelined kernels): the atoms of a column are grouped per 16-voxel z
// layer (an atom reaching two layers is written twice), each as a record that is ready to use from shared
// memory — tile-relative fp32 position, cutoff band, Gaussian coefficient, packed forbidden planes, type /
// radius, atom id, the 8-bit mask of the layer's 2 x 4 x 16-voxel cells its cutoff sphere reaches — followed
// by its feature row.  Order inside a layer = ascending atom id (the reference's summation order).
//
// A counting sort keyed by (molecule, column, layer), with work proportional to the (atom, key) pairs:
//   prep   counts the pairs per key (atomics without return value);
//   scan   one CTA per molecule: prefix sums over layers and columns -> segment offsets, tile descriptors;
//   place  one thread per atom: claims a slot in each of its keys' segments (any order) and drops its id there;
//   build  one warp per key: ranks the segment's ids (rank = number of smaller ids, so the final order is the
//          ascending atom order whatever order the slots were claimed in) and writes the entries, one lane per entry.
// ---------------------------------------------------------------------------------------------
struct LBinParams {
    double res, half_width, sigma;
    float tau_lin, tau_quad;
    int B, ncol, ncx, maxcols, zl, nl, nzc, tz, dim, mode, C, es4;
    int feat_vec;   // feature rows can be read four elements at a time (C % 4 == 0 and a base aligned to four elements)
    int feat_dtype; // element type of `features`: MVX_F32 0 | MVX_U8 2 | MVX_F16 3 — compact rows are widened here, exactly
    int64_t N;
    const int32_t* mol_offsets;
    const uint32_t* colrange;
    const uint32_t* alayers;
    const AtomRec* recs;
    const int32_t* types;
    const float* features;
    const uint32_t* kcnt;   // per key: pairs counted by prep
    uint32_t* cursor;       // per key: slots claimed so far (zeroed per call)
    uint32_t* lids;         // per layered slot: atom id (unordered inside a segment)
    uint2* bins;        // per (molecule, column): (offset in the molecule's layered segment, layered entries)
    uint2* lbins;       // per (molecule, column, layer): (offset in the column's segment, count)
    TileDesc* tdesc;    // per (molecule, column, z chunk), or nullptr
    float4* lent;       // molecule m owns entries [mol_offsets[m] * maxcols * zl, ...), es4 float4 words each
};

__global__ void __launch_bounds__(256) mvx_lscan_kernel(const LBinParams P, const int groups) {
    extern __shared__ uint32_t s_u32[];
    uint32_t* s_cnt = s_u32;            // [ncol] layered totals of the molecule's columns
    uint32_t* s_off = s_u32 + P.ncol;   // [ncol] their exclusive prefix sums
    const int mol = blockIdx.x / groups, grp = blockIdx.x % groups;   // every CTA of a molecule scans all its columns, writes its share
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nl = P.nl;
    const int ncz_max = (P.tz + kCellZ - 1) / kCellZ;
    const size_t lseg = (size_t)P.mol_offsets[mol] * (size_t)P.maxcols * (size_t)P.zl;
    for (int col = threadIdx.x; col < P.ncol; col += blockDim.x) {
        const uint32_t* k = P.kcnt + ((size_t)mol * P.ncol + col) * nl;
        uint32_t tot = 0;
        for (int L = 0; L < nl; ++L) tot += k[L];
        s_cnt[col] = tot;
    }
    __syncthreads();
    if (warp == 0) bin_scan_counts(s_cnt, s_off, P.ncol, lane);
    __syncthreads();
    for (int col = grp * nwarps + warp; col < P.ncol; col += groups * nwarps) {
        const size_t gcol = (size_t)mol * P.ncol + col;
        const uint32_t my_cnt = lane < nl ? P.kcnt[gcol * nl + lane] : 0u;
        uint32_t xs = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, xs, d); if (lane >= d) xs += y; }
        const uint32_t my_off = xs - my_cnt;   // lane L: first slot of layer L in the column's segment
        if (lane < nl) P.lbins[gcol * nl + lane] = make_uint2(my_off, my_cnt);
        if (lane == 0) P.bins[gcol] = make_uint2(s_off[col], s_cnt[col]);
        if (P.tdesc != nullptr) {   // tile descriptors of this column's z chunks (the layers of a chunk are consecutive)
            for (int zc = 0; zc < P.nzc; ++zc) {
                const int L0 = zc * ncz_max;
                const int ncz = (min(P.dim, (zc + 1) * P.tz) - zc * P.tz + kCellZ - 1) / kCellZ;
                const uint32_t o0 = __shfl_sync(0xffffffffu, my_off, L0);
                uint32_t e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    e[k] = __shfl_sync(0xffffffffu, my_off + my_cnt, min(L0 + min(k, ncz - 1), 31)) - o0;
                if (lane == 0) {
                    TileDesc d;
                    d.start = (unsigned long long)lseg + s_off[col] + o0; d.total = e[3]; d.mol = (uint32_t)mol;
#pragma unroll
                    for (int k = 0; k < 4; ++k) d.lend[k] = (uint16_t)min(e[k], 65535u);
                    d.origin = (uint32_t)((col / P.ncx) * kTile) | ((uint32_t)((col % P.ncx) * kTile) << 10) | ((uint32_t)(zc * P.tz) << 20);
                    d.pad = 0u;
                    P.tdesc[gcol * P.nzc + zc] = d;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) mvx_lplace_kernel(const LBinParams P) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.N) return;
    const uint32_t m = P.alayers[n];
    if (m == 0u) return;
    int lo = 0, hi = P.B;   // offs[lo] <= n < offs[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)P.mol_offsets[mid] <= n) lo = mid; else hi = mid;
    }
    const int mol = lo;
    const size_t lseg = (size_t)P.mol_offsets[mol] * (size_t)P.maxcols * (size_t)P.zl;
    const uint32_t cr = P.colrange[n];
    const int cx0 = cr & 0xFF, cx1 = (cr >> 8) & 0xFF, cy0 = (cr >> 16) & 0xFF, cy1 = (cr >> 24) & 0xFF;
    for (int cx = cx0; cx <= cx1; ++cx)
        for (int cy = cy0; cy <= cy1; ++cy) {
            const size_t gcol = (size_t)mol * P.ncol + (size_t)(cx * P.ncx + cy);
            const uint32_t coff = P.bins[gcol].x;
            for (uint32_t b = m; b != 0u; b &= b - 1u) {
                const size_t key = gcol * P.nl + (__ffs((int)b) - 1);
                const uint32_t pos = atomicAdd(P.cursor + key, 1u);
                P.lids[lseg + coff + P.lbins[key].x + pos] = (uint32_t)n;
            }
        }
}

constexpr int kLBuildIds = 256;     // ids of a segment ranked from shared memory per round
constexpr int kLBuildStageQ = 12;   // float4 words per entry the coalescing stage holds (C <= 36 channels)
inline size_t lbuild_smem_bytes(int es4) { return 8 * (2 * kLBuildIds * sizeof(uint32_t) + 32 * (size_t)(es4 <= kLBuildStageQ ? es4 : 0) * sizeof(float4)); }

__global__ void __launch_bounds__(256) mvx_lbuild_kernel(const LBinParams P) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* buf = reinterpret_cast<uint32_t*>(s_raw) + warp * 2 * kLBuildIds;   // the segment's ids as claimed
    uint32_t* srt = buf + kLBuildIds;                                             // ... in ascending order
    float4* stage = reinterpret_cast<float4*>(s_raw + 8 * 2 * kLBuildIds * sizeof(uint32_t)) + warp * 32 * P.es4;   // sized for es4 <= kLBuildStageQ
    const long long key = (long long)blockIdx.x * 8 + warp;
    if (key >= (long long)P.B * P.ncol * P.nl) return;
    const uint2 lb = P.lbins[key];
    const int S = (int)lb.y;
    if (S == 0) return;
    const int L = (int)(key % P.nl);
    const long long gcol = key / P.nl;
    const int col = (int)(gcol % P.ncol), mol = (int)(gcol / P.ncol);
    const size_t seg = (size_t)P.mol_offsets[mol] * (size_t)P.maxcols * (size_t)P.zl + P.bins[gcol].x + lb.x;
    const uint32_t* ids = P.lids + seg;
    const int ES4 = P.es4;

    const int x0 = (col / P.ncx) * kTile, y0 = (col % P.ncx) * kTile;
    const double ox0 = (double)x0 * P.res - P.half_width, oy0 = (double)y0 * P.res - P.half_width;
    const float resf = (float)P.res;
    const float bhx = 0.5f * (kCellX - 1) * resf, bhy = 0.5f * (kCellY - 1) * resf;
    const int ncz_max = (P.tz + kCellZ - 1) / kCellZ;
    const int zc = L / ncz_max, cz = L - zc * ncz_max;
    const int z0 = zc * P.tz, z1 = min(P.dim, z0 + P.tz);
    const int zlo = z0 + cz * kCellZ, zhi = min(zlo + kCellZ, z1);
    const float bhz = 0.5f * (zhi - zlo - 1) * resf;
    const float zmid = zlo * resf + bhz;

    // words 0..2 of atom n's entry and its 8-bit cell mask
    auto make_entry = [&](const uint32_t n, float4& e0, float4& e1, float4& e2) -> uint32_t {
        const AtomRec rec = P.recs[n];
        const float r = rec.r;
        const float r2 = r * r;
        const float tau = r * P.tau_lin + r2 * P.tau_quad;
        const float r2hi = r2 + tau, r2lo = r2 - tau;
        const double rs = (double)r * P.sigma;
        const float kc = (float)(-0.5 * 1.4426950408889634 / (rs * rs));
        const float ax = (float)(rec.px - ox0), ay = (float)(rec.py - oy0), az = (float)(rec.pz + P.half_width);
        // forbidden planes (block-cull emulation) are kept only where the cutoff sphere can reach them, so
        // that most entries carry "none" and the voxelize kernels skip the cull arithmetic
        const int fx = rec.fx - x0, fy = rec.fy - y0;
        const float rr = r * 1.001f + 0.01f * resf;
        const bool kx = rec.fx >= 0 && fx >= 0 && fx < kTile && fabsf(ax - fx * resf) <= rr;
        const bool ky = rec.fy >= 0 && fy >= 0 && fy < kTile && fabsf(ay - fy * resf) <= rr;
        const bool kz = rec.fz >= 0 && fabsf(az - rec.fz * resf) <= rr;
        const uint32_t forb = (uint32_t)(kx ? fx : 0xFF) | ((uint32_t)(ky ? fy : 0xFF) << 8) |
                              ((uint32_t)(kz ? rec.fz : 0xFFFF) << 16);
        // cells of this layer reached by the cutoff sphere: exact sphere / voxel-centre-box test
        const float ez = fmaxf(fabsf(az - zmid) - bhz, 0.f);
        const float ez2 = ez * ez;
        const float lim = r2hi + 1e-4f * (1.f + r2hi);
        uint32_t cm = 0u;
#pragma unroll
        for (int ix = 0; ix < kTile / kCellX; ++ix) {
            const float ex = fmaxf(fabsf(ax - ((ix * kCellX) * resf + bhx)) - bhx, 0.f);
            const float exz = fmaf(ex, ex, ez2);
#pragma unroll
            for (int iy = 0; iy < kTile / kCellY; ++iy) {
                const float ey = fmaxf(fabsf(ay - ((iy * kCellY) * resf + bhy)) - bhy, 0.f);
                if (fmaf(ey, ey, exz) <= lim) cm |= 1u << (ix * (kTile / kCellY) + iy);
            }
        }
        e0 = make_float4(ax, ay, az, r2hi);
        e1 = make_float4(r2lo, kc, __uint_as_float(forb), P.mode == 1 ? __uint_as_float((uint32_t)P.types[n]) : r);
        e2 = make_float4(__uint_as_float(n), __uint_as_float(cm), 0.f, 0.f);
        return cm;
    };
    // word k (4 channels) of atom n's padded feature row, as fp32.  Compact rows (u8 / f16) are widened here, exactly — the
    // reference's features.astype(float32), numpy/voxelizer.py:127-128 — so dense batches need no separate widening pass.
    auto feature_word = [&](const size_t n, const int k) -> float4 {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (4 * k >= P.C) return make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t at = n * (size_t)P.C + (size_t)(4 * k);
        if (P.feat_dtype == 0) {
            const float* f = P.features + at;
            if (P.feat_vec) return __ldg(reinterpret_cast<const float4*>(f));
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (4 * k + c < P.C) ? __ldg(f + c) : 0.f;
        } else if (P.feat_dtype == 2) {
            const unsigned char* f = reinterpret_cast<const unsigned char*>(P.features) + at;
            if (P.feat_vec) {
                const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(f));
                return make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (4 * k + c < P.C) ? (float)__ldg(f + c) : 0.f;
        } else {
            const __half* f = reinterpret_cast<const __half*>(P.features) + at;
            if (P.feat_vec) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(f));
                const __half2 a = *reinterpret_cast<const __half2*>(&u.x), b = *reinterpret_cast<const __half2*>(&u.y);
                const float2 fa = __half22float2(a), fb = __half22float2(b);
                return make_float4(fa.x, fa.y, fb.x, fb.y);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (4 * k + c < P.C) ? __half2float(f[c]) : 0.f;
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    };

    if (S <= kLBuildIds && ES4 <= kLBuildStageQ) {
        // common case.  rank = ids of the segment smaller than mine (ids are distinct) = the entry's place in
        // ascending atom order; sorted ids go to shared memory, then each round builds 32 CONSECUTIVE entries in
        // the stage and copies them out as one contiguous, fully coalesced block
        for (int j = lane; j < S; j += 32) buf[j] = ids[j];
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const uint32_t n = buf[i];
            uint32_t rank = 0;
            int j = 0;
            for (; j + 4 <= S; j += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(buf + j);
                rank += (v.x < n) + (v.y < n) + (v.z < n) + (v.w < n);
            }
            for (; j < S; ++j) rank += buf[j] < n;
            srt[rank] = n;
        }
        __syncwarp();
        for (int i0 = 0; i0 < S; i0 += 32) {
            const int i = i0 + lane, nv = min(32, S - i0);
            if (i < S) {
                const uint32_t n = srt[i];
                float4* e = stage + lane * ES4;
                // the first four feature words are requested before the record arithmetic (independent loads in
                // flight together); wider rows follow in groups of four
                const int nf = P.mode == 2 ? ES4 - 3 : 0;
                float4 fw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) fw[k] = k < nf ? feature_word(n, k) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 e0, e1, e2;
                make_entry(n, e0, e1, e2);
                e[0] = e0; e[1] = e1; e[2] = e2;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nf) e[3 + k] = fw[k];
                for (int k0 = 4; k0 < nf; k0 += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) fw[k] = k0 + k < nf ? feature_word(n, k0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k0 + k < nf) e[3 + k0 + k] = fw[k];
                }
            }
            __syncwarp();
            float4* dst = P.lent + (seg + i0) * (size_t)ES4;
            for (int q = lane; q < nv * ES4; q += 32) dst[q] = stage[q];
            __syncwarp();
        }
        return;
    }

    // large segments / wide feature rows: ids ranked in rounds from shared memory, entries stored in place
    for (int i0 = 0; i0 < S; i0 += 32) {   // 32 entries per round, one per lane
        const int i = i0 + lane;
        const uint32_t n = i < S ? ids[i] : 0xFFFFFFFFu;
        uint32_t rank = 0;
        for (int j0 = 0; j0 < S; j0 += kLBuildIds) {
            const int nj = min(kLBuildIds, S - j0);
            if (S > kLBuildIds || i0 == 0) {   // a segment that fits is staged once
                __syncwarp();
                for (int j = lane; j < nj; j += 32) buf[j] = ids[j0 + j];
                __syncwarp();
            }
            int j = 0;
            for (; j + 4 <= nj; j += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(buf + j);
                rank += (v.x < n) + (v.y < n) + (v.z < n) + (v.w < n);
            }
            for (; j < nj; ++j) rank += buf[j] < n;
        }
        if (i >= S) continue;
        const size_t slot = seg + rank;
        float4 e0, e1, e2;
        make_entry(n, e0, e1, e2);
        float4* e = P.lent + slot * (size_t)ES4;
        e[0] = e0; e[1] = e1; e[2] = e2;
        if (P.mode == 2) {
            for (int k = 0; k < ES4 - 3; ++k) e[3 + k] = feature_word(n, k);
        }
    }
}

// The transforms prep draws for molecules [offset, offset + B): same device function, one thread per molecule.
__global__ void __launch_bounds__(128) mvx_draw_transforms_kernel(const DrawParams P) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.B) return;
    Rigid R;
    draw_rigid(P.seed, P.offset + (unsigned long long)m, P.flags, P.rt, R);
    double* o = P.out + 7 * (size_t)m;
    o[0] = R.q[0]; o[1] = R.q[1]; o[2] = R.q[2]; o[3] = R.q[3];
    o[4] = R.t[0]; o[5] = R.t[1]; o[6] = R.t[2];
}

// Synthetic sweep ligands (see mvx_rigid.cuh): one thread per molecule.
__global__ void __launch_bounds__(128) mvx_synth_ligands_kernel(const SynthParams P) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.B) return;
    const unsigned long long gmol = P.first_mol + (unsigned long long)m;
    const int V = synth_count(P.seed, gmol, P.vmin, P.vmax);
    if (P.counts != nullptr) P.counts[m] = V;
    if (P.mol_offsets == nullptr) return;
    const int a0 = P.mol_offsets[m];
    const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
    auto step_of = [&](int a, double (&d)[3], uint32_t& tw) {
        uint32_t w[4];
        philox4x32_10((uint32_t)gmol, (uint32_t)(gmol >> 32), (uint32_t)(a + 1), kSynthDomain, k0, k1, w);
        const double z = 2.0 * ((double)w[0] * (1.0 / 4294967296.0)) - 1.0;   // uniform direction on the sphere
        const double rho = sqrt(fmax(0.0, 1.0 - z * z));
        double sn, cs;
        sincospi(2.0 * ((double)w[1] * (1.0 / 4294967296.0)), &sn, &cs);
        d[0] = P.step * rho * cs; d[1] = P.step * rho * sn; d[2] = P.step * z;
        tw = w[2];
    };
    double sum[3] = {0.0, 0.0, 0.0}, pos[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < V; ++a) {   // pass 1: centroid of the walk
        double d[3]; uint32_t tw;
        step_of(a, d, tw);
#pragma unroll
        for (int k = 0; k < 3; ++k) { pos[k] += d[k]; sum[k] += pos[k]; }
    }
    const double mean[3] = {sum[0] / V, sum[1] / V, sum[2] / V};
    pos[0] = pos[1] = pos[2] = 0.0;
    for (int a = 0; a < V; ++a) {   // pass 2: recentred, rounded to fp32-representable values
        double d[3]; uint32_t tw;
        step_of(a, d, tw);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pos[k] += d[k];
            const float v = (float)(pos[k] - mean[k]);
            if (P.coords_f64) reinterpret_cast<double*>(P.coords)[3 * (size_t)(a0 + a) + k] = (double)v;
            else reinterpret_cast<float*>(P.coords)[3 * (size_t)(a0 + a) + k] = v;
        }
        if (P.types != nullptr) P.types[a0 + a] = (int32_t)(tw % (uint32_t)P.num_types);
    }
}

// ---------------------------------------------------------------------------------------------
// Brick compaction (host-side consumers): a finished fp32 grid is >= 97 % zeros for ligands, so copying it to the host
// densely measures PCIe.  One CTA per (molecule, 8x8 column) — empty columns are known from the binning pass and exit
// at once —, one warp per (channel, 8-voxel z range): a brick of 8 x 8 x 8 voxels (2 KB) that holds any non-zero
// value claims a slot of the compact buffer (atomic counter; slot order is arbitrary, the id says where the brick
// belongs) and is copied there.  Bricks that stick out of the grid (D % 8 != 0) are zero-padded.
// ---------------------------------------------------------------------------------------------
constexpr int kBrick = 8;
struct CompactParams {
    int dim, ncx, ncol, Cout, nbz, B;
    const uint2* bins;      // per (molecule, column): count in .y (nullptr: scan every column)
    const float* grid;      // (B, Cout, D, D, D)
    uint32_t* ids;          // (cap) brick id = ((mol * Cout + c) * ncol + col) * nbz + bz
    float* vals;            // (cap, 512), brick-local [x][y][z]
    uint32_t cap;
    uint32_t* count;        // bricks found (may exceed cap: the caller re-runs with a larger buffer)
};

__global__ void __launch_bounds__(256) mvx_compact_bricks_kernel(const CompactParams P) {
    const int col = blockIdx.x % P.ncol, mol = blockIdx.x / P.ncol;
    if (P.bins != nullptr && P.bins[(size_t)mol * P.ncol + col].y == 0u) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = P.dim;
    const int x0 = (col / P.ncx) * kBrick, y0 = (col % P.ncx) * kBrick;
    const size_t plane = (size_t)D * D * D;
    const bool vec = (D & 3) == 0;
    for (int b = warp; b < P.Cout * P.nbz; b += 8) {
        const int c = b / P.nbz, bz = b - c * P.nbz;
        const int z0 = bz * kBrick;
        const float* g = P.grid + ((size_t)mol * P.Cout + c) * plane;
        float4 v[4];   // lane owns rows 2 * lane, 2 * lane + 1 (row = (x, y), 8 voxels of z)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = 2 * lane + r, x = x0 + (row >> 3), y = y0 + (row & 7);
            const float* src = g + ((size_t)x * D + y) * D + z0;
            const bool in = x < D && y < D;
            if (vec) {
                v[2 * r] = (in && z0 < D) ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[2 * r + 1] = (in && z0 + 4 < D) ? __ldg(reinterpret_cast<const float4*>(src) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) t[k] = (in && z0 + k < D) ? __ldg(src + k) : 0.f;
                v[2 * r] = make_float4(t[0], t[1], t[2], t[3]);
                v[2 * r + 1] = make_float4(t[4], t[5], t[6], t[7]);
            }
        }
        bool nz = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) nz = nz || v[q].x != 0.f || v[q].y != 0.f || v[q].z != 0.f || v[q].w != 0.f;
        if (__ballot_sync(0xffffffffu, nz) == 0u) continue;
        uint32_t slot = 0u;
        if (lane == 0) slot = atomicAdd(P.count, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= P.cap) continue;
        if (lane == 0) P.ids[slot] = (uint32_t)((((size_t)mol * P.Cout + c) * P.ncol + col) * P.nbz + bz);
        float4* dst = reinterpret_cast<float4*>(P.vals + (size_t)slot * 512) + 4 * lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) __stcs(dst + q, v[q]);
    }
}

// ---------------------------------------------------------------------------------------------
// voxelize, precision = 64 (the reference's `precision=64` constructor argument, numpy/voxelizer.py:28-34; SURVEY
// row f4): distances, dr = dist / r, the Gaussian and the accumulation all in fp64, (B, Cout, D, D, D) float64 out.
// API completeness, not a tuned path: one CTA per (molecule, 8x8 column) on the column lists of the generic form,
// one thread per voxel, atoms in ascending order.
// ---------------------------------------------------------------------------------------------
struct VoxF64Params {
    double res, half_width, sigma, radius;
    int dim, ncx, ncol, mode, C, Cout, maxcols, binary;
    int scalar_radius;          // 1: every atom uses `radius` (python float); 0: the atom record's fp32 radius, widened
    int clast;                  // 1: channels-last output (B, D, D, D, Cout)
    const int32_t* mol_offsets;
    const AtomRec* recs;
    const uint2* bins;
    const uint32_t* lists;
    const int32_t* types;
    const float* features;
    const float* chan_radii;    // features + channel-wise radii: (C,) fp32, else nullptr
    double* out;
};

__device__ __forceinline__ double term_f64(double s, double r, double sigma, int binary) {
    const double dr = __ddiv_rn(__dsqrt_rn(s), r);
    if (dr > 1.0) return 0.0;
    if (binary) return 1.0;
    const double q = __ddiv_rn(dr, sigma);
    return exp(-0.5 * __dmul_rn(q, q));
}

__global__ void __launch_bounds__(256) mvx_voxelize_f64_kernel(const VoxF64Params P) {
    const int col = blockIdx.x % P.ncol, mol = blockIdx.x / P.ncol;
    const int x0 = (col / P.ncx) * kTile, y0 = (col % P.ncx) * kTile;
    const int D = P.dim;
    const size_t plane = (size_t)D * D * D;
    const uint2 bin = P.bins[(size_t)mol * P.ncol + col];
    const int cnt = (int)bin.y;
    const uint32_t* list = P.lists + (size_t)P.mol_offsets[mol] * (size_t)P.maxcols + bin.x;
    double* out_mol = P.out + (size_t)mol * P.Cout * plane;
    for (int item = threadIdx.x; item < kTile * kTile * D; item += blockDim.x) {
        const int row = item / D, z = item - row * D;
        const int x = x0 + (row >> 3), y = y0 + (row & 7);
        if (x >= D || y >= D) continue;
        const double gx = __dsub_rn(__dmul_rn((double)x, P.res), P.half_width);
        const double gy = __dsub_rn(__dmul_rn((double)y, P.res), P.half_width);
        const double gz = __dsub_rn(__dmul_rn((double)z, P.res), P.half_width);
        const size_t vox = ((size_t)x * D + y) * D + z;
        for (int c0 = 0; c0 < P.Cout; c0 += 4) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int j = 0; j < cnt; ++j) {
                const uint32_t n = list[j];
                const AtomRec rec = P.recs[n];
                if (rec.fx == x || rec.fy == y || rec.fz == z) continue;   // the reference's block cull (forbidden planes)
                const double dx = __dsub_rn(rec.px, gx), dy = __dsub_rn(rec.py, gy), dz = __dsub_rn(rec.pz, gz);
                const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (P.chan_radii != nullptr) {   // features, channel-wise radii: the kernel radius is the channel's
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c0 + c < P.C)
                            acc[c] += (double)P.features[(size_t)n * P.C + c0 + c] * term_f64(s, (double)P.chan_radii[c0 + c], P.sigma, P.binary);
                    continue;
                }
                const double r = P.scalar_radius ? P.radius : (double)rec.r;
                if (s > r * r * 1.000001) continue;   // clearly outside (dr > 1): contributes exactly 0
                const double t = term_f64(s, r, P.sigma, P.binary);
                if (P.mode == 0) {
                    if (c0 == 0) acc[0] += t;
                } else if (P.mode == 1) {
                    const int ch = P.types[n] - c0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] += (ch == c) ? t : 0.0;
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c0 + c < P.C) acc[c] += (double)P.features[(size_t)n * P.C + c0 + c] * t;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c0 + c < P.Cout) out_mol[P.clast ? vox * (size_t)P.Cout + (size_t)(c0 + c) : (size_t)(c0 + c) * plane + vox] = acc[c];
        }
    }
}

// Compact feature rows (u8 / f16) -> fp32, exactly: the reference's features.astype(float32) (numpy/voxelizer.py:127-128).
__global__ void __launch_bounds__(256) mvx_widen_features_kernel(const void* __restrict__ src, int is_f16, size_t n, float* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = is_f16 ? __half2float(reinterpret_cast<const __half*>(src)[i]) : (float)reinterpret_cast<const unsigned char*>(src)[i];
}

}  // namespace mvx
