// mvx_vox_kernels.cuh — the voxelize kernel forms (instantiated per (mode, channel chunk, density) in mvx_vox_inst.cu).
#pragma once
#include "mvx_common.cuh"

namespace mvx {

// ---------------------------------------------------------------------------------------------
// voxelize: one CTA per (molecule, 8x8 column, z chunk).  Each thread owns NV consecutive z voxels of
// one (x, y) row for CH channels, gathers over the column's staged atoms in ascending order and
// writes every output voxel exactly once (zeros included) with 128-bit stores along W.
//
// Cutoff decisions (the 0.135-high step of SURVEY.md hazard 3): fp32 d^2 from tile-relative
// coordinates decides everything outside a tolerance band around r^2; inside the band the
// reference's arithmetic is replayed exactly — fp64 sqrt((dx*dx+dy*dy)+dz*dz) with no FMA, rounded
// to fp32, divided by the fp32 radius, compared with 1.0f (numpy/voxelizer.py:544-559 + scipy cdist).
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ bool exact_hit(const AtomRec* __restrict__ rec, float r32, int x, int y, int z,
                                       double res, double half_width) {
    double gx = __dsub_rn(__dmul_rn((double)x, res), half_width);
    double gy = __dsub_rn(__dmul_rn((double)y, res), half_width);
    double gz = __dsub_rn(__dmul_rn((double)z, res), half_width);
    double dx = __dsub_rn(rec->px, gx), dy = __dsub_rn(rec->py, gy), dz = __dsub_rn(rec->pz, gz);
    double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    float d32 = __double2float_rn(__dsqrt_rn(s));
    return __fdiv_rn(d32, r32) <= 1.0f;
}

// 2^x for x <= 0 by the SFU (ex2.approx: max rel. error 2^-22, inside the 1e-5 parity tolerance)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Two fp32 FMAs in one instruction (Blackwell FFMA2, PTX fma.rn.f32x2): d0 += a0*b, d1 += a1*b.  Each half is an
// ordinary IEEE fp32 FMA, so results equal two scalar fmaf calls bit for bit.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b) {
    unsigned long long a, bb, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

// Output element size and stores.  fp32 is the reference layout; bf16 / fp16 round each finished voxel once
// (round-to-nearest-even), halving the bytes that bound the op.
template <bool O16>
__device__ __forceinline__ void store_vox(char* p, const float (&v)[4], int kind) {
    if (!O16) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else if (kind == 1) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        __stcs(reinterpret_cast<uint2*>(p), make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b)));
    } else {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        __stcs(reinterpret_cast<uint2*>(p), make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b)));
    }
}
template <bool O16>
__device__ __forceinline__ void store_vox(char* p, const float (&v)[1], int kind) {
    if (!O16) __stcs(reinterpret_cast<float*>(p), v[0]);
    else if (kind == 1) { const __nv_bfloat16 h = __float2bfloat16_rn(v[0]); *reinterpret_cast<__nv_bfloat16*>(p) = h; }
    else { const __half h = __float2half_rn(v[0]); *reinterpret_cast<__half*>(p) = h; }
}

// Division-free zero fill of one tile (8 x 8 x [z0, z1) voxels, channels [c_begin, c_end)), 16-byte stores.
// VPI voxels of ES bytes per thread item; the fp32 instance (4 x 4 B) is the hot path of ligand batches.
template <int VPI, int ES, int NT>
__device__ __forceinline__ void zero_fill_items(char* out_mol, size_t plane, int D, int x0, int y0, int z0, int z1,
                                                int c_begin, int c_end, int tid) {
    const int lz = (z1 - z0) / VPI;
    const int nitems = kTile * kTile * lz;
    const int qstep = NT / lz, rstep = NT - qstep * lz;
    int row = tid / lz, lzi = tid - row * lz;
    const size_t pstride = plane * ES;
    for (int item = tid; item < nitems; item += NT) {
        const int x = x0 + (row >> 3), y = y0 + (row & 7);
        if (x < D && y < D) {
            char* p = out_mol + ((size_t)c_begin * plane + ((size_t)x * D + y) * D + z0 + lzi * VPI) * ES;
            if (VPI * ES == 16) {
                for (int ch = c_begin; ch < c_end; ++ch, p += pstride) __stcs(reinterpret_cast<float4*>(p), make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
                for (int ch = c_begin; ch < c_end; ++ch, p += pstride) __stcs(reinterpret_cast<uint2*>(p), make_uint2(0u, 0u));
            }
        }
        row += qstep; lzi += rstep;
        if (lzi >= lz) { lzi -= lz; ++row; }
    }
}

template <bool O16, int NT = kThreads>
__device__ __forceinline__ void zero_fill_tile(char* out_mol, size_t plane, int D, int x0, int y0, int z0, int z1,
                                               int c_begin, int c_end, int tid) {
    if (!O16) zero_fill_items<4, 4, NT>(out_mol, plane, D, x0, y0, z0, z1, c_begin, c_end, tid);
    else if (((z1 - z0) & 7) == 0 && (D & 7) == 0) zero_fill_items<8, 2, NT>(out_mol, plane, D, x0, y0, z0, z1, c_begin, c_end, tid);
    else zero_fill_items<4, 2, NT>(out_mol, plane, D, x0, y0, z0, z1, c_begin, c_end, tid);
}

// ---- channels-last output (B, D, H, W, Cout), SURVEY row f3: the Cout channels of one voxel are contiguous ----
// Plain (write-back) stores: a 32-byte sector of a channels-last voxel is completed by consecutive store instructions,
// so the lines should stay in L2 until they are whole (the streaming hint of the reference-layout stores evicts first).
template <bool O16>
__device__ __forceinline__ void store_chan1(char* p, float v, int kind) {
    if (!O16) *reinterpret_cast<float*>(p) = v;
    else if (kind == 1) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
    else *reinterpret_cast<__half*>(p) = __float2half_rn(v);
}
template <bool O16>
__device__ __forceinline__ void store_chan4(char* p, float a, float b, float c, float d, int kind) {   // four consecutive channels of one voxel
    if (!O16) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    } else if (kind == 1) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    } else {
        const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}
// Channels-last stores of a lane PAIR (lanes l, l ^ 1: z neighbours of one (x, y) row).  Each lane holds NQ chunks of 4
// consecutive elements (a chunk = 16 B fp32 / 8 B 16-bit), contiguous from its base address; the partner's run starts
// `pdelta` bytes away.  Adjacent chunks (2j, 2j + 1) of ONE lane are written by the two lanes in the same instruction —
// the even lane's pair first, then the odd lane's — so a store instruction fills whole 32-byte sectors instead of
// halves (one shuffle per element moves the other half across).  Must be called by the whole warp.
template <int NQ, bool O16>
__device__ __forceinline__ void store_chunk_pairs(char* p, const ptrdiff_t pdelta, const bool valid, const bool odd,
                                                  const float (&v)[NQ][4], const int kind) {
    constexpr int CB = O16 ? 8 : 16;
    const bool pvalid = __shfl_xor_sync(0xffffffffu, valid ? 1 : 0, 1) != 0;
    const bool valid_even = odd ? pvalid : valid, valid_odd = odd ? valid : pvalid;
    char* const pp = p + pdelta;
#pragma unroll
    for (int j = 0; j + 1 < NQ; j += 2) {
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = __shfl_xor_sync(0xffffffffu, odd ? v[j][e] : v[j + 1][e], 1);
        if (valid_even) {   // the even lane's chunks 2j, 2j + 1
            if (odd) store_chan4<O16>(pp + (j + 1) * CB, r[0], r[1], r[2], r[3], kind);
            else store_chan4<O16>(p + j * CB, v[j][0], v[j][1], v[j][2], v[j][3], kind);
        }
        if (valid_odd) {    // the odd lane's chunks 2j, 2j + 1
            if (odd) store_chan4<O16>(p + (j + 1) * CB, v[j + 1][0], v[j + 1][1], v[j + 1][2], v[j + 1][3], kind);
            else store_chan4<O16>(pp + j * CB, r[0], r[1], r[2], r[3], kind);
        }
    }
    if ((NQ & 1) && valid) store_chan4<O16>(p + (NQ - 1) * CB, v[NQ - 1][0], v[NQ - 1][1], v[NQ - 1][2], v[NQ - 1][3], kind);
}
// A lane's 4 voxels x COUT channels are 4 * COUT consecutive elements in the channels-last layout when one channel chunk
// covers the grid: written as COUT 4-element chunks whatever COUT is (register order fixed at compile time).
template <int CH, int COUT, bool O16>
__device__ __forceinline__ void store_lane_flat(char* p, const ptrdiff_t pdelta, const bool valid, const bool odd,
                                                const float (&acc)[CH][4], const int kind) {
    float v[COUT][4];
#pragma unroll
    for (int q = 0; q < COUT; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) v[q][e] = acc[(4 * q + e) % COUT][(4 * q + e) / COUT];
    store_chunk_pairs<COUT, O16>(p, pdelta, valid, odd, v, kind);
}
// One lane's CH channels x 4 z voxels -> out, either layout.  (x, y, z) is the lane's first voxel, c0 the chunk's first
// channel.  Called by the whole warp (the channels-last forms exchange data between lanes); `valid`: this lane's voxels exist.
template <int CH, bool O16, bool CL>
__device__ __forceinline__ void store_lane(const VoxParams& P, char* out_mol, size_t plane, int D, int x, int y, int z,
                                           int c0, const float (&acc)[CH][4], const bool valid) {
    constexpr int es = O16 ? 2 : 4;
    if (!CL) {   // compile-time: the reference layout's instances carry none of the channels-last code
        if (!valid) return;
        char* p = out_mol + ((size_t)c0 * plane + ((size_t)x * D + y) * D + z) * es;
        const size_t pstride = plane * es;
        if (c0 + CH <= P.c_end) {
#pragma unroll
            for (int c = 0; c < CH; ++c, p += pstride) store_vox<O16>(p, acc[c], P.out_kind);
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c, p += pstride)
                if (c0 + c < P.c_end) store_vox<O16>(p, acc[c], P.out_kind);
        }
        return;
    }
    const size_t vstride = (size_t)P.Cout * es;
    char* p = out_mol + (((size_t)x * D + y) * D + z) * vstride + (size_t)c0 * es;
    const bool odd = (threadIdx.x & 1u) != 0u;   // lane = row * 4 + zq: the partner lane holds the next / previous 4 z voxels
    const ptrdiff_t pdelta = odd ? -(ptrdiff_t)(4 * vstride) : (ptrdiff_t)(4 * vstride);
    const bool whole = c0 == 0 && P.c_begin == 0 && P.c_end == P.Cout;   // one channel chunk covers the grid
    if (CH >= 4 && whole && P.Cout <= CH && P.Cout > CH - 4) {   // warp-uniform
        switch (CH - P.Cout) {
            case 0: store_lane_flat<CH, CH, O16>(p, pdelta, valid, odd, acc, P.out_kind); break;
            case 1: store_lane_flat<CH, (CH > 1 ? CH - 1 : 1), O16>(p, pdelta, valid, odd, acc, P.out_kind); break;
            case 2: store_lane_flat<CH, (CH > 2 ? CH - 2 : 1), O16>(p, pdelta, valid, odd, acc, P.out_kind); break;
            default: store_lane_flat<CH, (CH > 3 ? CH - 3 : 1), O16>(p, pdelta, valid, odd, acc, P.out_kind); break;
        }
        return;
    }
    if (CH % 4 == 0 && (P.Cout & 3) == 0 && (c0 & 3) == 0 && c0 + CH <= P.c_end) {   // warp-uniform: 4-channel chunks per voxel
#pragma unroll
        for (int k = 0; k < 4; ++k, p += vstride) {
            float v[CH / 4 > 0 ? CH / 4 : 1][4];
#pragma unroll
            for (int q = 0; q < CH / 4; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[q][e] = acc[4 * q + e][k];
            store_chunk_pairs<(CH / 4 > 0 ? CH / 4 : 1), O16>(p, pdelta, valid, odd, v, P.out_kind);
        }
        return;
    }
    if (!valid) return;
#pragma unroll
    for (int k = 0; k < 4; ++k, p += vstride) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
            if (c0 + c < P.c_end) store_chan1<O16>(p + c * es, acc[c][k], P.out_kind);
    }
}
// Zero fill of one tile in the channels-last layout: per (x, y) row the z range x all channels is one contiguous run.
template <int ES, int NT>
__device__ __forceinline__ void zero_fill_tile_clast(char* out_mol, int D, int Cout, int x0, int y0, int z0, int z1,
                                                     int c_begin, int c_end, int tid) {
    const int nz = z1 - z0;
    const size_t vstride = (size_t)Cout * ES;
    const size_t rowbytes = (size_t)nz * vstride;
    if (c_begin == 0 && c_end == Cout && (rowbytes & 15) == 0 && (((size_t)D * vstride) & 15) == 0 && (((size_t)z0 * vstride) & 15) == 0) {
        const int q = (int)(rowbytes >> 4);
        const int total = kTile * kTile * q;
        const int qstep = NT / q, rstep = NT - qstep * q;
        int row = tid / q, w = tid - row * q;
        for (int i = tid; i < total; i += NT) {
            const int x = x0 + (row >> 3), y = y0 + (row & 7);
            if (x < D && y < D)
                __stcs(reinterpret_cast<float4*>(out_mol + (((size_t)x * D + y) * D + z0) * vstride) + w, make_float4(0.f, 0.f, 0.f, 0.f));
            row += qstep; w += rstep;
            if (w >= q) { w -= q; ++row; }
        }
    } else {
        const int nc = c_end - c_begin, per_row = nz * nc, total = kTile * kTile * per_row;
        for (int i = tid; i < total; i += NT) {
            const int row = i / per_row, r = i - row * per_row, zz = r / nc, c = c_begin + (r - zz * nc);
            const int x = x0 + (row >> 3), y = y0 + (row & 7);
            if (x < D && y < D) {
                char* p = out_mol + (((size_t)x * D + y) * D + z0 + zz) * vstride + (size_t)c * ES;
                if (ES == 4) *reinterpret_cast<float*>(p) = 0.f;
                else *reinterpret_cast<uint16_t*>(p) = (uint16_t)0;
            }
        }
    }
}
// Zero fill of a tile, either layout.
template <bool O16, int NT, bool CL>
__device__ __forceinline__ void zero_tile(const VoxParams& P, char* out_mol, size_t plane, int D, int x0, int y0, int z0, int z1, int tid) {
    if (!CL) zero_fill_tile<O16, NT>(out_mol, plane, D, x0, y0, z0, z1, P.c_begin, P.c_end, tid);
    else zero_fill_tile_clast<(O16 ? 2 : 4), NT>(out_mol, D, P.Cout, x0, y0, z0, z1, P.c_begin, P.c_end, tid);
}

template <int MODE, int CH, bool BINARY, int NV, bool O16, bool CL>
__global__ void __launch_bounds__(kThreads) mvx_voxelize_kernel(const VoxParams P) {
    __shared__ float4 sA[kMaxCand];   // rel x, rel y, rel z, r^2 + tau
    __shared__ float4 sB[kMaxCand];   // r^2 - tau, gaussian coefficient, forbidden planes, type | radius
    __shared__ int sIdx[kMaxCand];    // global atom id (exact recheck)
    __shared__ float sF[MODE == 2 ? kMaxCand * CH : 1];

    const int tid = threadIdx.x;
    int t = blockIdx.x;
    const int zc = t % P.nzc; t /= P.nzc;
    const int col = t % P.ncol;
    const int mol = t / P.ncol;
    const int x0 = (col / P.ncx) * kTile, y0 = (col % P.ncx) * kTile, z0 = zc * P.tz;
    const int D = P.dim;
    const int z1 = min(D, z0 + P.tz);
    const int lz = (z1 - z0 + NV - 1) / NV;   // thread items per (x, y) row
    const int nitems = kTile * kTile * lz;
    const size_t plane = (size_t)D * D * D;
    constexpr int es = O16 ? 2 : 4;
    char* out_mol = reinterpret_cast<char*>(P.out) + (size_t)mol * P.Cout * plane * es;

    const uint2 bin = P.bins[(size_t)mol * P.ncol + col];
    const int cnt = (int)bin.y;
    const uint32_t* list = P.lists + (size_t)P.mol_offsets[mol] * (size_t)P.maxcols + bin.x;

    if (cnt == 0 && CL) {
        zero_fill_tile_clast<es, kThreads>(out_mol, D, P.Cout, x0, y0, z0, z1, P.c_begin, P.c_end, tid);
        return;
    }
    if (cnt == 0) {   // empty column: pure zero fill
        float zero[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) zero[k] = 0.f;
        for (int ch = P.c_begin; ch < P.c_end; ++ch) {
            for (int item = tid; item < nitems; item += kThreads) {
                int row = item / lz, x = x0 + (row >> 3), y = y0 + (row & 7), z = z0 + (item - row * lz) * NV;
                if (x < D && y < D) store_vox<O16>(out_mol + ((size_t)ch * plane + ((size_t)x * D + y) * D + z) * es, zero, P.out_kind);
            }
        }
        return;
    }

    const double ox0 = (double)x0 * P.res - P.half_width;
    const double oy0 = (double)y0 * P.res - P.half_width;
    const double oz0 = (double)z0 * P.res - P.half_width;
    const bool single_round = cnt <= kMaxCand;
    int staged_c0 = -1;

    for (int c0 = P.c_begin; c0 < P.c_end; c0 += CH) {
        for (int item0 = 0; item0 < nitems; item0 += kThreads) {
            const int item = item0 + tid;
            const int row = item / lz;
            const int x = x0 + (row >> 3), y = y0 + (row & 7), z = z0 + (item - row * lz) * NV;
            const bool valid = item < nitems && x < D && y < D;
            const float ox = (float)((double)(x - x0) * P.res), oy = (float)((double)(y - y0) * P.res);
            float oz[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) oz[k] = (float)((double)(z + k - z0) * P.res);
            float acc[CH][NV];
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int k = 0; k < NV; ++k) acc[c][k] = 0.f;

            for (int r0 = 0; r0 < cnt; r0 += kMaxCand) {
                const int nc = min(kMaxCand, cnt - r0);
                if (!(single_round && staged_c0 == c0)) {
                    __syncthreads();
                    if (tid < nc) {
                        const int n = (int)list[r0 + tid];
                        const AtomRec rec = P.recs[n];
                        float r = rec.r;
                        if (MODE == 2 && P.chan_radii != nullptr) r = P.chan_radii[c0];
                        const float r2 = r * r;
                        const float tau = r * P.tau_lin + r2 * P.tau_quad;
                        float r2hi = r2 + tau;
                        if (rec.zhi < z0 || rec.zlo >= z1) r2hi = -1.f;   // outside this z chunk: never hits
                        const int fx = rec.fx - x0, fy = rec.fy - y0, fz = rec.fz - z0;
                        const uint32_t forb = (uint32_t)((rec.fx >= 0 && fx >= 0 && fx < kTile) ? fx : 0xFF) |
                                              ((uint32_t)((rec.fy >= 0 && fy >= 0 && fy < kTile) ? fy : 0xFF) << 8) |
                                              ((uint32_t)((rec.fz >= 0 && fz >= 0 && fz < 0xFFFF) ? fz : 0xFFFF) << 16);
                        const double rs = (double)r * P.sigma;
                        const float kc = (float)(-0.5 * 1.4426950408889634 / (rs * rs));
                        sA[tid] = make_float4((float)(rec.px - ox0), (float)(rec.py - oy0), (float)(rec.pz - oz0), r2hi);
                        sB[tid] = make_float4(r2 - tau, kc, __uint_as_float(forb),
                                              MODE == 1 ? __int_as_float(P.types[n]) : r);
                        sIdx[tid] = n;
                    }
                    if (MODE == 2) {
                        for (int i = tid; i < nc * CH; i += kThreads) {
                            int j = i / CH, c = i - j * CH;
                            int n = (int)list[r0 + j];
                            sF[i] = (c0 + c < P.C) ? P.features[(size_t)n * P.C + c0 + c] : 0.f;
                        }
                    }
                    __syncthreads();
                    staged_c0 = c0;
                }
                if (!valid) continue;
                for (int j = 0; j < nc; ++j) {
                    const float4 A = sA[j];
                    const float dx = A.x - ox, dy = A.y - oy;
                    const float dxy = dx * dx + dy * dy;
                    if (dxy > A.w) continue;
                    const float4 Bv = sB[j];
                    const uint32_t forb = __float_as_uint(Bv.z);
                    if ((int)(forb & 0xFF) == x - x0 || (int)((forb >> 8) & 0xFF) == y - y0) continue;
                    const int fz = (int)(forb >> 16);
                    float w[NV];
                    bool any = false;
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        const float dz = A.z - oz[k];
                        const float s = dxy + dz * dz;
                        bool hit = s < Bv.x;
                        if (!hit && s <= A.w) {
                            float r32 = (MODE == 1) ? P.recs[sIdx[j]].r : Bv.w;
                            hit = exact_hit(P.recs + sIdx[j], r32, x, y, z + k, P.res, P.half_width);
                        }
                        if (z + k - z0 == fz) hit = false;
                        w[k] = hit ? (BINARY ? 1.0f : exp2f(s * Bv.y)) : 0.f;
                        any = any || hit;
                    }
                    if (!any) continue;
                    if (MODE == 0) {
#pragma unroll
                        for (int k = 0; k < NV; ++k) acc[0][k] += w[k];
                    } else if (MODE == 1) {
                        const int ct = __float_as_int(Bv.w) - c0;
#pragma unroll
                        for (int c = 0; c < CH; ++c)
#pragma unroll
                            for (int k = 0; k < NV; ++k) acc[c][k] += (ct == c) ? w[k] : 0.f;
                    } else {
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            const float f = sF[j * CH + c];
#pragma unroll
                            for (int k = 0; k < NV; ++k) acc[c][k] = fmaf(f, w[k], acc[c][k]);
                        }
                    }
                }
            }
            if (valid && !CL) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int ch = c0 + c;
                    if (ch < P.c_end) store_vox<O16>(out_mol + ((size_t)ch * plane + ((size_t)x * D + y) * D + z) * es, acc[c], P.out_kind);
                }
            } else if (valid) {   // channels-last: element stores
#pragma unroll
                for (int k = 0; k < NV; ++k)
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        if (c0 + c < P.c_end)
                            store_chan1<O16>(out_mol + ((((size_t)x * D + y) * D + z + k) * P.Cout + c0 + c) * es, acc[c][k], P.out_kind);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// voxelize, "warp-cell" form (the main path; needs D % 4 == 0).  Same CTA tile as above, but each
// warp owns a compact cell of RX x RY x (4*LPR) voxels (LPR lanes along z per row):
//   0. staging (once per tile): tile-relative fp32 atom records in shared memory plus, per atom, a
//      32-bit mask of the tile's cells whose voxel-centre box its cutoff sphere reaches (exact test),
//   1. a warp compacts the atoms whose mask names its cell into a short warp-private list
//      (32 staged atoms per ballot; ascending order is kept),
//   2. every lane tests that short list for its 4 voxels and records candidate hits in a bitmask,
//   3. each lane walks ITS OWN set bits in ascending order (= the reference's atom order), so the
//      channel accumulation runs max-hits-per-lane times per warp instead of once per atom that touches
//      any lane — this removes the SIMT divergence that bounded the dense (protein) workloads.
// Zero fill of empty columns is division-free: one address computation per thread item.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kNoForb = 0xFFFFFFFFu;

// padded feature row (words): an odd number of 16-byte words, so lane-private LDS.128 of a quarter warp are conflict-free
template <int CH>
__host__ __device__ constexpr int feat_stride() { return 4 * ((CH / 4 + 1) | 1); }

template <int MODE, int CH>
constexpr size_t cells_smem_bytes() {
    constexpr int NT = kThreads;
    return (size_t)(2 * NT) * (2 * sizeof(float4) + sizeof(int) + sizeof(uint32_t)) +
           (NT / 32) * kWarpList * (2 * sizeof(float4) + sizeof(uint16_t)) + 32 * sizeof(float4) +
           (MODE == 2 ? (size_t)(2 * NT) * feat_stride<CH>() * sizeof(float) : 0);
}

template <int MODE, int CH, bool BINARY, bool O16, bool CL>
__device__ __forceinline__ void cells_body(const VoxParams& P) {
    constexpr int NT = kThreads, LPR = 4;   // 4 lanes (one float4 each) along z per row: cells of 2 x 4 x 16 voxels
    constexpr int NW = NT / 32;             // warps per CTA
    constexpr int SC = 2 * NT;    // atoms staged per round
    constexpr int ROWS = 32 / LPR;
    constexpr int RY = (LPR >= 8) ? 2 : 4;
    constexpr int RX = ROWS / RY;
    constexpr int CZ = 4 * LPR;
    constexpr int NCX = kTile / RX, NCY = kTile / RY;
    constexpr int FS = feat_stride<CH>();
    static_assert(NCX * NCY * (64 / CZ) <= 32, "cell mask is 32 bits");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sA = reinterpret_cast<float4*>(smem_raw);        // rel x, rel y, rel z, r^2 + tau
    float4* sB = sA + SC;                             // r^2 - tau, gaussian coef, forbidden planes, type | radius
    int* sN = reinterpret_cast<int*>(sB + SC);        // global atom id (exact recheck)
    uint32_t* sM = reinterpret_cast<uint32_t*>(sN + SC);   // cells reached by the cutoff sphere
    float4* wA_all = reinterpret_cast<float4*>(sM + SC);
    float4* wB_all = wA_all + NW * kWarpList;
    uint16_t* wI_all = reinterpret_cast<uint16_t*>(wB_all + NW * kWarpList);
    float4* sBox = reinterpret_cast<float4*>(wI_all + NW * kWarpList);   // per cell: box centre x, y, z, half z
    float* sF = reinterpret_cast<float*>(sBox + 32);               // features [SC][CH + 4]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t = blockIdx.x;
    const int zc = t % P.nzc; t /= P.nzc;
    const int col = t % P.ncol;
    const int mol = t / P.ncol;
    const int x0 = (col / P.ncx) * kTile, y0 = (col % P.ncx) * kTile, z0 = zc * P.tz;
    const int D = P.dim;
    const int z1 = min(D, z0 + P.tz);
    const size_t plane = (size_t)D * D * D;
    constexpr int es = O16 ? 2 : 4;
    char* out_mol = reinterpret_cast<char*>(P.out) + (size_t)mol * P.Cout * plane * es;

    const uint2 bin = P.bins[(size_t)mol * P.ncol + col];
    const int cnt = (int)bin.y;

    if (cnt == 0) {   // empty column: pure zero fill
        zero_tile<O16, kThreads, CL>(P, out_mol, plane, D, x0, y0, z0, z1, tid);
        return;
    }

    const ColEntry* ent = P.entries + (size_t)P.mol_offsets[mol] * (size_t)P.maxcols + bin.x;
    const float resf = (float)P.res;
    const int mask_shift = zc * ((P.tz + CZ - 1) / CZ) * (NCX * NCY);
    const bool single_round = cnt <= SC;
    int staged_c0 = -1;

    float4* wA = wA_all + warp * kWarpList;
    float4* wB = wB_all + warp * kWarpList;
    uint16_t* wI = wI_all + warp * kWarpList;

    const int row = lane / LPR, zq = lane % LPR;
    const int rx = row / RY, ry = row % RY;
    const int ncz = (z1 - z0 + CZ - 1) / CZ;
    const int ncells = NCX * NCY * ncz;
    // half extents of a cell's voxel-centre box (x, y constant; z shorter in the last layer)
    const float bhx = 0.5f * (RX - 1) * resf, bhy = 0.5f * (RY - 1) * resf;
    const float inv_res = 1.0f / resf;
    if (tid < 32) {
        const int bz = tid / (NCX * NCY), bxy = tid % (NCX * NCY);
        const int czn = max(1, min(CZ, z1 - z0 - bz * CZ));
        const float bhz = 0.5f * (czn - 1) * resf;
        sBox[tid] = make_float4(((bxy / NCY) * RX) * resf + bhx, ((bxy % NCY) * RY) * resf + bhy, (z0 + bz * CZ) * resf + bhz, bhz);
    }

    for (int c0 = P.c_begin; c0 < P.c_end; c0 += CH) {
        for (int cell0 = 0; cell0 < ncells; cell0 += NW) {
            const int cell = cell0 + warp;
            const bool cell_ok = cell < ncells;
            const int cz = cell / (NCX * NCY), cxy = cell % (NCX * NCY);
            const int cxl = cxy / NCY, cyl = cxy % NCY;
            const int lx = cxl * RX + rx, ly = cyl * RY + ry, lzv = cz * CZ + zq * 4;   // tile-relative voxel
            const int x = x0 + lx, y = y0 + ly, z = z0 + lzv;
            const bool valid = cell_ok && x < D && y < D && z < z1;
            const uint32_t lane_key = (uint32_t)lx | ((uint32_t)ly << 8);
            const float ox = (float)((double)lx * P.res), oy = (float)((double)ly * P.res);
            float oz[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) oz[k] = (float)((double)(z + k) * P.res);   // grid-absolute, like ColEntry::az

            float acc[CH][4];
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;

            for (int r0 = 0; r0 < cnt; r0 += SC) {
                const int nc = min(SC, cnt - r0);
                if (!(single_round && staged_c0 == c0)) {
                    __syncthreads();
                    for (int i = tid; i < nc; i += NT) {   // coalesced copy of the expanded entries
                        const float4* src = reinterpret_cast<const float4*>(ent + r0 + i);
                        float4 e0 = src[0], e1 = src[1];
                        const float4 e2 = src[2];
                        if (MODE == 2 && P.chan_radii != nullptr) {   // channel-wise features: this channel's radius
                            const float r = P.chan_radii[c0];
                            const float r2 = r * r;
                            const float tau = r * P.tau_lin + r2 * P.tau_quad;
                            const double rs = (double)r * P.sigma;
                            e0.w = r2 + tau; e1.x = r2 - tau; e1.y = (float)(-0.5 * 1.4426950408889634 / (rs * rs));
                            e1.w = r;
                        }
                        sA[i] = e0; sB[i] = e1;
                        sN[i] = (int)__float_as_uint(e2.x);
                        const unsigned long long m64 = (unsigned long long)__float_as_uint(e2.y) |
                                                       ((unsigned long long)__float_as_uint(e2.z) << 32);
                        sM[i] = (uint32_t)(m64 >> mask_shift);
                    }
                    if (MODE == 2) {
                        for (int i = tid; i < nc * CH; i += NT) {
                            const int j = i / CH, c = i - j * CH;
                            const int n = (int)ent[r0 + j].n;
                            sF[j * FS + c] = (c0 + c < P.C) ? P.features[(size_t)n * P.C + c0 + c] : 0.f;
                        }
                    }
                    __syncthreads();
                    // cell masks when the expand pass could not precompute them (> 64 cells per column or another
                    // cell shape): lane = cell, one staged atom per warp iteration, exact sphere / voxel-centre-box
                    // test; the ballot IS the atom's 32-bit mask of cells its cutoff sphere reaches
                    if (!P.masks) {
                        const float4 box = sBox[lane];
                        for (int i = warp; i < nc; i += NW) {
                            const float4 A = sA[i];
                            const float ex = fmaxf(fabsf(A.x - box.x) - bhx, 0.f);
                            const float ey = fmaxf(fabsf(A.y - box.y) - bhy, 0.f);
                            const float ez = fmaxf(fabsf(A.z - box.z) - box.w, 0.f);
                            const bool in = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) <= A.w + 1e-4f * (1.f + A.w) && lane < ncells;
                            const uint32_t m = __ballot_sync(0xffffffffu, in);
                            if (lane == 0) sM[i] = m;
                        }
                        __syncthreads();
                    }
                    staged_c0 = c0;
                }
                if (!cell_ok) continue;   // warp-uniform

                int base = 0;
                while (base < nc) {
                    // 1. warp filter: compact the staged atoms whose cell mask names this cell
                    int wn = 0;
                    while (base < nc && wn <= kWarpList - 32) {
                        const int i = base + lane;
                        const bool in = (i < nc) && ((sM[i] >> cell) & 1u);
                        const uint32_t m = __ballot_sync(0xffffffffu, in);
                        if (in) {
                            const int pos = wn + __popc(m & ((1u << lane) - 1u));
                            wA[pos] = sA[i]; wB[pos] = sB[i]; wI[pos] = (uint16_t)i;
                        }
                        wn += __popc(m);
                        base += 32;
                    }
                    __syncwarp();
                    // 2. every lane tests the warp list for its 4 voxels (nearest of the 4 along z); candidate
                    //    hits -> two 32-bit masks.  The decision proper (incl. the exact band) is taken in 3.
                    uint32_t mask_lo = 0u, mask_hi = 0u;
                    if (valid) {
                        auto near_hit = [&](const float4 A) -> bool {
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float tz_ = A.z - oz[0];
                            const float dzc = fmaf(-resf, fminf(fmaxf(rintf(tz_ * inv_res), 0.f), 3.f), tz_);   // nearest of the 4
                            return fmaf(dzc, dzc, fmaf(dx, dx, dy * dy)) <= A.w;
                        };
                        const int n_lo = min(wn, 32);
#pragma unroll 4
                        for (int j = 0; j < n_lo; ++j)
                            if (near_hit(wA[j])) mask_lo |= 1u << j;
#pragma unroll 4
                        for (int j = 32; j < wn; ++j)
                            if (near_hit(wA[j])) mask_hi |= 1u << (j - 32);
                    }
                    // 3. lane-private walk over the set bits, ascending (fixed fp32 summation order)
                    while (__any_sync(0xffffffffu, (mask_lo | mask_hi) != 0u)) {
                        if ((mask_lo | mask_hi) != 0u) {
                            int j;
                            if (mask_lo != 0u) { j = __ffs((int)mask_lo) - 1; mask_lo &= mask_lo - 1u; }
                            else { j = 31 + __ffs((int)mask_hi); mask_hi &= mask_hi - 1u; }
                            const float4 A = wA[j];
                            const float4 Bv = wB[j];
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float dxy = fmaf(dx, dx, dy * dy);
                            // block-cull emulation: voxels on this atom's forbidden planes take nothing from it
                            bool off[4] = {false, false, false, false};
                            if (P.cull) {   // uniform
                                const uint32_t forb = __float_as_uint(Bv.z);
                                const uint32_t tx = forb ^ lane_key;
                                const bool row_off = (tx & 0xFFu) == 0u || (tx & 0xFF00u) == 0u;
                                const int dzf = (int)(forb >> 16) - z;
#pragma unroll
                                for (int k = 0; k < 4; ++k) off[k] = row_off || dzf == k;
                            }
                            float sk[4], w[4];
                            bool band = false;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float dz = A.z - oz[k];
                                sk[k] = fmaf(dz, dz, dxy);
                                w[k] = (sk[k] < Bv.x && !off[k]) ? (BINARY ? 1.0f : fast_exp2(sk[k] * Bv.y)) : 0.f;
                                band = band || (sk[k] >= Bv.x && sk[k] <= A.w);
                            }
                            if (band) {   // rare: voxels inside the tolerance band replay the reference's fp64 arithmetic
                                const int n = sN[wI[j]];
                                const float r32 = (MODE == 1) ? P.recs[n].r : Bv.w;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (sk[k] >= Bv.x && sk[k] <= A.w && !off[k] &&
                                        exact_hit(P.recs + n, r32, x, y, z + k, P.res, P.half_width))
                                        w[k] = BINARY ? 1.0f : fast_exp2(sk[k] * Bv.y);
                                }
                            }
                            if (MODE == 0) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc[0][k] += w[k];
                            } else if (MODE == 1) {
                                const int ct = __float_as_int(Bv.w) - c0;
#pragma unroll
                                for (int c = 0; c < CH; ++c)
#pragma unroll
                                    for (int k = 0; k < 4; ++k) acc[c][k] += (ct == c) ? w[k] : 0.f;
                            } else {
                                const float* frow = sF + (int)wI[j] * FS;
#pragma unroll
                                for (int c4 = 0; c4 < CH; c4 += 4) {
                                    float f[4];
                                    if (CH >= 4) {
                                        const float4 fv = *reinterpret_cast<const float4*>(frow + c4);
                                        f[0] = fv.x; f[1] = fv.y; f[2] = fv.z; f[3] = fv.w;
                                    } else {
                                        f[0] = frow[0]; f[1] = f[2] = f[3] = 0.f;
                                    }
                                    if (CH >= 4) {   // channel pairs through the packed FMA
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            ffma2(acc[c4][k], acc[c4 + 1 < CH ? c4 + 1 : c4][k], f[0], f[1], w[k]);
                                            ffma2(acc[c4 + 2 < CH ? c4 + 2 : c4][k], acc[c4 + 3 < CH ? c4 + 3 : c4][k], f[2], f[3], w[k]);
                                        }
                                    } else {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) acc[0][k] = fmaf(f[0], w[k], acc[0][k]);
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            store_lane<CH, O16, CL>(P, out_mol, plane, D, x, y, z, c0, acc, valid);
        }
    }
}


template <int MODE, int CH, bool BINARY, bool O16, bool CL>
__global__ void __launch_bounds__(kThreads, (CH <= 8 ? 3 : 2)) mvx_voxelize_cells_kernel(const VoxParams P) {
    cells_body<MODE, CH, BINARY, O16, CL>(P);
}

// The same kernel capped at 112 registers: two CTAs leave 8 192 registers of an SM free, room for one 128-thread CTA of
// the per-atom prep / binning kernels of the NEXT batch (mvx_voxelize_split: they run on a second stream in the shadow of
// this HBM-bound kernel).  Costs a 24-byte spill, which a write-bound kernel does not notice.
template <int MODE, int CH, bool BINARY, bool O16, bool CL>
__global__ void __maxnreg__(112) mvx_voxelize_cells_lean_kernel(const VoxParams P) {
    cells_body<MODE, CH, BINARY, O16, CL>(P);
}

// ---------------------------------------------------------------------------------------------
// voxelize, "tile" form (the main path; needs D % 4 == 0).  One CTA per tile of 8 x 8 x tz voxels.
//   0. staging: ONE flat, coalesced copy of the tile's layered entries (records + feature rows, prepared by
//      mvx_expand_layers_kernel) and of their 8-bit cell masks into shared memory — no per-atom arithmetic,
//      a single global round trip;
//   1. a warp owns a cell of 2 x 4 x 16 voxels (lane = one float4 along z) and compacts, from ITS layer's
//      sub-range only, the atoms whose mask names the cell into a warp-private list (32 atoms per ballot);
//   2. every lane tests that short list against the nearest of its 4 voxels -> hit bitmask;
//   3. each lane walks its own set bits in ascending order (the reference's atom order) and accumulates
//      CH channels in registers: the accumulation runs max-hits-per-lane times per warp, not once per atom
//      that touches any lane;
//   4. one 128-bit streaming store per channel per lane.  Empty tiles are zero-filled division-free.
// ---------------------------------------------------------------------------------------------
constexpr int kStageMaxEntries = 512;
template <int MODE>
__host__ __device__ constexpr int stage_bytes() { return MODE == 2 ? 60 * 1024 : 24 * 1024; }
template <int MODE>
constexpr size_t tiles_smem_bytes() {
    return (size_t)stage_bytes<MODE>() + kStageMaxEntries * sizeof(uint32_t) + 4 * sizeof(uint2) +
           (kThreads / 32) * kWarpList * (2 * sizeof(float4) + sizeof(uint16_t));
}

template <int MODE, int CH, bool BINARY, bool O16, bool CL>
__device__ __forceinline__ void tiles_body(const VoxParams& P, const int tile_id) {
    constexpr int LPR = 4, RX = kCellX, RY = kCellY, CZ = kCellZ;
    constexpr int NCY = kTile / RY;
    constexpr int NW = kThreads / 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sE = reinterpret_cast<float4*>(smem_raw);                                   // staged layered entries
    uint32_t* sM = reinterpret_cast<uint32_t*>(smem_raw + stage_bytes<MODE>());         // their cell masks
    uint2* sL = reinterpret_cast<uint2*>(sM + kStageMaxEntries);                        // layer sub-ranges (start, end)
    float4* wA_all = reinterpret_cast<float4*>(sL + 4);
    float4* wB_all = wA_all + NW * kWarpList;
    uint16_t* wI_all = reinterpret_cast<uint16_t*>(wB_all + NW * kWarpList);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t = tile_id;
    const int zc = t % P.nzc; t /= P.nzc;
    const int col = t % P.ncol;
    const int mol = t / P.ncol;
    const int x0 = (col / P.ncx) * kTile, y0 = (col % P.ncx) * kTile, z0 = zc * P.tz;
    const int D = P.dim;
    const int z1 = min(D, z0 + P.tz);
    const size_t plane = (size_t)D * D * D;
    constexpr int es = O16 ? 2 : 4;
    char* out_mol = reinterpret_cast<char*>(P.out) + (size_t)mol * P.Cout * plane * es;

    const size_t gcol = (size_t)mol * P.ncol + col;
    const uint2 bin = P.bins[gcol];
    if (bin.y == 0) {   // empty column
        zero_tile<O16, kThreads, CL>(P, out_mol, plane, D, x0, y0, z0, z1, tid);
        return;
    }
    const int ncz_max = (P.tz + CZ - 1) / CZ;
    const int ncz = (z1 - z0 + CZ - 1) / CZ;
    const uint2* lb = P.lbins + gcol * P.nlayers + (size_t)zc * ncz_max;
    const uint2 lb_first = lb[0], lb_last = lb[ncz - 1];
    const int seg_start = (int)lb_first.x;
    const int total = (int)(lb_last.x + lb_last.y) - seg_start;   // layers of a chunk are consecutive
    if (total == 0) {   // the column has atoms, none reaches this z chunk
        zero_tile<O16, kThreads, CL>(P, out_mol, plane, D, x0, y0, z0, z1, tid);
        return;
    }
    if (tid < ncz) {
        const uint2 v = lb[tid];
        sL[tid] = make_uint2(v.x - seg_start, v.x - seg_start + v.y);
    }
    const int ES4 = P.es4;
    const int SC = min(kStageMaxEntries, stage_bytes<MODE>() / (ES4 * (int)sizeof(float4)));
    const size_t colseg = (size_t)P.mol_offsets[mol] * (size_t)P.maxcols * (size_t)P.zl + bin.x + seg_start;
    const float4* src = P.lent + colseg * (size_t)ES4;
    const bool single_round = total <= SC;
    bool staged = false;

    float4* wA = wA_all + warp * kWarpList;
    float4* wB = wB_all + warp * kWarpList;
    uint16_t* wI = wI_all + warp * kWarpList;

    const float resf = (float)P.res;
    const float inv_res = 1.0f / resf;
    const int row = lane / LPR, zq = lane % LPR;
    const int rx = row / RY, ry = row % RY;
    const int ncells = kCellsXY * ncz;

    for (int c0 = P.c_begin; c0 < P.c_end; c0 += CH) {
        for (int cell0 = 0; cell0 < ncells; cell0 += NW) {
            const int cell = cell0 + warp;
            const bool cell_ok = cell < ncells;
            const int cz = cell / kCellsXY, cxy = cell % kCellsXY;
            const int cxl = cxy / NCY, cyl = cxy % NCY;
            const int lx = cxl * RX + rx, ly = cyl * RY + ry, lzv = cz * CZ + zq * 4;
            const int x = x0 + lx, y = y0 + ly, z = z0 + lzv;
            const bool valid = cell_ok && x < D && y < D && z < z1;
            const uint32_t lane_key = (uint32_t)lx | ((uint32_t)ly << 8);
            const float ox = (float)lx * resf, oy = (float)ly * resf;
            float oz[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) oz[k] = (float)(z + k) * resf;   // grid-absolute, like the entries' z

            float acc[CH][4];
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;

            for (int r0 = 0; r0 < total; r0 += SC) {
                const int nc = min(SC, total - r0);
                if (!(single_round && staged)) {
                    __syncthreads();
                    const float4* g = src + (size_t)r0 * ES4;
                    const int nq = nc * ES4;
                    for (int q = tid; q < nq; q += kThreads) sE[q] = g[q];
                    for (int i = tid; i < nc; i += kThreads) sM[i] = __float_as_uint(g[(size_t)i * ES4 + 2].y);   // cell mask, word 2 of the entry
                    __syncthreads();
                    if (MODE == 2 && P.chan_radii != nullptr) {   // channel-wise features: this channel's radius
                        const float r = P.chan_radii[c0];
                        const float r2 = r * r;
                        const float tau = r * P.tau_lin + r2 * P.tau_quad;
                        const double rs = (double)r * P.sigma;
                        const float kc = (float)(-0.5 * 1.4426950408889634 / (rs * rs));
                        for (int i = tid; i < nc; i += kThreads) {
                            sE[i * ES4].w = r2 + tau;
                            float4 b = sE[i * ES4 + 1];
                            b.x = r2 - tau; b.y = kc; b.w = r;
                            sE[i * ES4 + 1] = b;
                        }
                        __syncthreads();
                    }
                    staged = true;
                }
                if (!cell_ok) continue;   // warp-uniform

                const uint2 lr = sL[cz];
                int base = max((int)lr.x, r0) - r0;
                const int end = min((int)lr.y, r0 + nc) - r0;
                while (base < end) {
                    // 1. warp filter over this cell's layer: atoms whose mask names the cell, order kept
                    int wn = 0;
                    while (base < end && wn <= kWarpList - 32) {
                        const int i = base + lane;
                        const bool in = (i < end) && ((sM[i] >> cxy) & 1u);
                        const uint32_t m = __ballot_sync(0xffffffffu, in);
                        if (in) {
                            const int pos = wn + __popc(m & ((1u << lane) - 1u));
                            wA[pos] = sE[i * ES4]; wB[pos] = sE[i * ES4 + 1]; wI[pos] = (uint16_t)i;
                        }
                        wn += __popc(m);
                        base += 32;
                    }
                    __syncwarp();
                    // 2. every lane tests the warp list against the nearest of its 4 voxels -> hit bitmasks
                    uint32_t mask_lo = 0u, mask_hi = 0u;
                    if (valid) {
                        auto near_hit = [&](const float4 A) -> bool {
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float tz_ = A.z - oz[0];
                            const float dzc = fmaf(-resf, fminf(fmaxf(rintf(tz_ * inv_res), 0.f), 3.f), tz_);
                            return fmaf(dzc, dzc, fmaf(dx, dx, dy * dy)) <= A.w;
                        };
                        const int n_lo = min(wn, 32);
#pragma unroll 4
                        for (int j = 0; j < n_lo; ++j)
                            if (near_hit(wA[j])) mask_lo |= 1u << j;
#pragma unroll 4
                        for (int j = 32; j < wn; ++j)
                            if (near_hit(wA[j])) mask_hi |= 1u << (j - 32);
                    }
                    // 3. lane-private walk over the set bits, ascending (fixed fp32 summation order)
                    while (__any_sync(0xffffffffu, (mask_lo | mask_hi) != 0u)) {
                        if ((mask_lo | mask_hi) != 0u) {
                            int j;
                            if (mask_lo != 0u) { j = __ffs((int)mask_lo) - 1; mask_lo &= mask_lo - 1u; }
                            else { j = 31 + __ffs((int)mask_hi); mask_hi &= mask_hi - 1u; }
                            const float4 A = wA[j];
                            const float4 Bv = wB[j];
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float dxy = fmaf(dx, dx, dy * dy);
                            bool off[4] = {false, false, false, false};
                            if (P.cull) {   // uniform: block-cull emulation, voxels on the atom's forbidden planes take nothing
                                const uint32_t forb = __float_as_uint(Bv.z);
                                const uint32_t tx = forb ^ lane_key;
                                const bool row_off = (tx & 0xFFu) == 0u || (tx & 0xFF00u) == 0u;
                                const int dzf = (int)(forb >> 16) - z;
#pragma unroll
                                for (int k = 0; k < 4; ++k) off[k] = row_off || dzf == k;
                            }
                            // trigger: |s - r^2| <= tau, WIDENED, for any of the 4 voxels -> those voxels replay in fp64.  The
                            // midpoint and half-width of r^2 +- tau are fp32 results with up to 4e-7 * r^2 of rounding error
                            // (0.2 % of tau next to 1.0): the margin below covers it, so no s in [r^2 - tau, r^2 + tau] slips
                            // through; the decision proper is the pair of direct compares inside
                            const float r2c = 0.5f * (A.w + Bv.x), tauh = fmaf(4e-7f, r2c, 0.505f * (A.w - Bv.x));
                            float sk[4], w[4], dmin = 3.0e38f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float dz = A.z - oz[k];
                                sk[k] = fmaf(dz, dz, dxy);
                                w[k] = (sk[k] < Bv.x && !off[k]) ? (BINARY ? 1.0f : fast_exp2(sk[k] * Bv.y)) : 0.f;
                                dmin = fminf(dmin, fabsf(sk[k] - r2c));
                            }
                            const bool band = dmin <= tauh;
                            const float4* ent = sE + (int)wI[j] * ES4;
                            if (band) {   // rare
                                const int n = (int)__float_as_uint(ent[2].x);
                                const float r32 = (MODE == 1) ? P.recs[n].r : Bv.w;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (sk[k] >= Bv.x && !off[k]) {
                                        const bool hit = sk[k] <= A.w && exact_hit(P.recs + n, r32, x, y, z + k, P.res, P.half_width);
                                        w[k] = hit ? (BINARY ? 1.0f : fast_exp2(sk[k] * Bv.y)) : 0.f;
                                    }
                                }
                            }
                            if (MODE == 0) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc[0][k] += w[k];
                            } else if (MODE == 1) {
                                const int ct = __float_as_int(Bv.w) - c0;
#pragma unroll
                                for (int c = 0; c < CH; ++c)
#pragma unroll
                                    for (int k = 0; k < 4; ++k) acc[c][k] += (ct == c) ? w[k] : 0.f;
                            } else if (CH >= 4) {
                                const float4* frow = ent + 3 + (c0 >> 2);
#pragma unroll
                                for (int c4 = 0; c4 < CH; c4 += 4) {
                                    const float4 fv = frow[c4 >> 2];
                                    const float f[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {   // channel pairs through the packed FMA
                                        ffma2(acc[c4][k], acc[c4 + 1][k], f[0], f[1], w[k]);
                                        ffma2(acc[c4 + 2][k], acc[c4 + 3][k], f[2], f[3], w[k]);
                                    }
                                }
                            } else {
                                const float f = reinterpret_cast<const float*>(ent + 3)[c0];
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc[0][k] = fmaf(f, w[k], acc[0][k]);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            store_lane<CH, O16, CL>(P, out_mol, plane, D, x, y, z, c0, acc, valid);
        }
    }
}

template <int MODE, int CH, bool BINARY, bool O16, bool CL>
__global__ void __launch_bounds__(kThreads, 2) mvx_voxelize_tiles_kernel(const VoxParams P) {
    tiles_body<MODE, CH, BINARY, O16, CL>(P, (int)blockIdx.x);
}

// Companion of the pipelined form: a small grid scans the tile descriptors and does, with the tile form's
// multi-round staging, the few tiles that have more entries than the pipelined form takes (usually none).
constexpr int kSweepList = 64;
template <int MODE, int CH, bool BINARY, bool O16, bool CL>
__global__ void __launch_bounds__(kThreads, 2) mvx_voxelize_sweep_kernel(const VoxParams P, const unsigned ntiles) {
    __shared__ int s_list[kSweepList];
    __shared__ int s_n;
    const unsigned chunk = (ntiles + gridDim.x - 1) / gridDim.x;
    const unsigned t0 = blockIdx.x * chunk, t1 = min(ntiles, t0 + chunk);
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (unsigned t = t0 + threadIdx.x; t < t1; t += kThreads) {
        if (P.tdesc[t].total > (uint32_t)P.pipe_sc) {
            const int i = atomicAdd(&s_n, 1);
            if (i < kSweepList) s_list[i] = (int)t;
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (n <= kSweepList) {
        for (int i = 0; i < n; ++i) {
            tiles_body<MODE, CH, BINARY, O16, CL>(P, s_list[i]);
            __syncthreads();
        }
    } else {   // more overflow tiles than the list holds: walk the chunk
        for (unsigned t = t0; t < t1; ++t) {
            if (P.tdesc[t].total > (uint32_t)P.pipe_sc) {
                tiles_body<MODE, CH, BINARY, O16, CL>(P, (int)t);
                __syncthreads();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// voxelize, "pipelined" form (dense batches; needs D % 4 == 0).  Same tile, cell and lane layout and the same
// arithmetic as the tile form, but with ONE PERSISTENT CTA PER SM (12 warps, up to 168 registers: no spills) that
// walks the tiles round-robin, and a tile's staging is an asynchronous bulk copy (cp.async.bulk, completion on
// an mbarrier) issued several tiles ahead by thread 0 into a shared-memory RING allocated by bytes:
//   descriptor of tile k+1 (32 B)  ->  layered entries of tile k (one contiguous block)  ->  compute.
// There is no CTA-wide barrier in the steady state.  The work items of a tile (cell x channel chunk) are handed
// out dynamically (a shared-memory counter per slot), and the warps drift apart by up to the ring depth: a warp waits on
// the "full" mbarrier of its tile's slot, draws cells until none is left, arrives on the slot's "empty" mbarrier
// and moves on.  The global-memory latency of staging is hidden behind earlier tiles and the warps no longer
// reach their store bursts together.  Thread 0 polls (never blocks on) the "empty" barriers between its cells.
// Tiles with more entries than half the ring are skipped here and done by the tile form (second launch,
// which exits at once everywhere else).
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {   // may suspend up to ~20 us
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// Bounded wait: a pipeline bug traps (the launch fails with an error) instead of hanging the GPU.  The bound is wall
// time (20 s on %globaltimer, looked at every 4096 failed attempts), so time-slicing with other contexts cannot trip it.
template <int TAG>
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spins & 4095u) == 0u) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 20000000000ull) __trap();
        }
    }
}
// TAG: one copy of the out-of-line slow path per register-budget region of a warp-specialised kernel (a function shared by
// callers under different setmaxnreg budgets makes ptxas' register allocation fail)
template <int TAG = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow<TAG>(bar, parity);
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned); completion counts on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Shared-memory accesses of the hot loops by explicit 32-bit shared addresses: nvcc otherwise rebuilds the shared
// window base (S2UR SR_CgaCtaId + ULEA) next to every access group inside the divergent loops.
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}

template <int MODE, int CH, bool BINARY, bool O16, bool MULTI, bool CL>
__global__ void __launch_bounds__(kPipeThreads, 1) mvx_voxelize_pipe_kernel(const VoxParams P, const unsigned ntiles) {
    constexpr int LPR = 4, RX = kCellX, RY = kCellY, CZ = kCellZ;
    constexpr int NCY = kTile / RY;
    constexpr int NW = kPipeWarps;
    constexpr int ND = kPipeSlots;
    const int Q = P.pipe_q;   // float4 words of the entry ring (smaller when the hit cache is present)

    extern __shared__ __align__(128) float4 smem_q[];
    float4* const ring = smem_q;
    TileDesc* const sDesc = reinterpret_cast<TileDesc*>(smem_q + Q);
    uint64_t* const full = reinterpret_cast<uint64_t*>(sDesc + ND);   // slot s: the tile's entries have landed
    uint64_t* const empty = full + ND;                                 // slot s: all warps are done with the tile
    uint64_t* const dfull = empty + ND;                                // slot s: the descriptor has landed
    int* const sOff = reinterpret_cast<int*>(dfull + ND);              // slot s: ring offset of the tile's entries
    int* const sNext = sOff + ND;                                      // slot s: next work item (cell x channel chunk) to hand out
    // warp lists: float4 words [WA0, WA0 + NW * kWarpList) of smem_q (record A of each listed atom), then their
    // entry indices (u16).  Hot code indexes smem_q directly so that every access is a plain shared-memory one.
    const int WA0 = Q + kPipeSlots * ((int)sizeof(TileDesc) + 3 * 8 + 2 * 4) / 16;
    uint32_t sq = smem_u32(smem_q);   // shared address of smem_q, pinned in a register
    asm volatile("" : "+r"(sq));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = P.dim;
    const size_t plane = (size_t)D * D * D;
    constexpr int es = O16 ? 2 : 4;
    const int ES4 = P.es4;
    const uint32_t SC = (uint32_t)P.pipe_sc;   // largest tile (entries) this form takes
    const unsigned G = gridDim.x;

    const float resf = (float)P.res;
    const float inv_res = 1.0f / resf;
    const int row = lane / LPR, zq = lane % LPR;
    const int rx = row / RY, ry = row % RY;

    // ---- producer (thread 0) -------------------------------------------------------------------
    unsigned next_k = 0;   // first iteration whose entries have not been requested yet
    unsigned rel_k = 0;    // oldest iteration not yet seen released (its slot and ring bytes are still in use)
    int head = 0;          // ring position (float4 words) of the next allocation
    auto issue_desc = [&](unsigned k) {   // descriptor of iteration k's tile -> slot k % ND
        const unsigned t = blockIdx.x + k * G;
        if (t < ntiles) {
            const int s = (int)(k & (ND - 1));
            mbar_arrive_expect_tx(&dfull[s], (uint32_t)sizeof(TileDesc));
            bulk_g2s(&sDesc[s], P.tdesc + t, (uint32_t)sizeof(TileDesc), &dfull[s]);
        }
    };
    // Requests the entries of iteration next_k (and the descriptor after it) if a slot and ring space are free.
    auto produce = [&](bool block) -> bool {
        const unsigned k = next_k;
        if (blockIdx.x + k * G >= ntiles) return false;
        const int s = (int)(k & (ND - 1));
        auto release_oldest = [&]() -> bool {   // wait for / poll the release of the oldest tile in flight
            uint64_t* const eb = &empty[rel_k & (ND - 1)];
            const uint32_t epar = (rel_k / ND) & 1u;
            if (block) mbar_wait(eb, epar);
            else if (!mbar_test(eb, epar)) return false;
            ++rel_k;
            return true;
        };
        // slots: iteration k + 1's descriptor goes to slot (k + 1) % ND, so at most ND - 2 tiles stay in flight
        while (next_k - rel_k > (unsigned)(ND - 2))
            if (!release_oldest()) return false;
        mbar_wait(&dfull[s], (k / ND) & 1u);   // requested when iteration k - 1 was produced
        const unsigned long long start = sDesc[s].start;
        const uint32_t total = sDesc[s].total;
        const int need = (total > 0u && total <= SC) ? (int)total * ES4 : 0;
        int at;
        for (;;) {
            if (next_k == rel_k) { at = 0; break; }            // nothing in flight: restart at the ring's base
            if (need == 0) { at = head; break; }
            const int tail = sOff[rel_k & (ND - 1)];
            if (head >= tail) {
                if (need <= Q - head) { at = head; break; }
                if (need < tail) { at = 0; break; }
            } else if (need < tail - head) { at = head; break; }
            if (!release_oldest()) return false;
        }
        sOff[s] = at;
        sNext[s] = 0;
        if (need > 0) {
            const uint32_t bytes = (uint32_t)need * (uint32_t)sizeof(float4);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(ring + at, P.lent + start * (unsigned long long)ES4, bytes, &full[s]);
        } else {
            mbar_arrive(&full[s]);   // nothing to stage: empty tile, or an overflow tile left to the tile form
        }
        head = at + need;
        issue_desc(k + 1);
        next_k = k + 1;
        return true;
    };

    if (tid == 0) {
        for (int i = 0; i < ND; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NW); mbar_init(&dfull[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) issue_desc(0);

    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
        if (tid == 0) {
            while (next_k <= it) produce(true);
            while (produce(false)) {}
        }
        __syncwarp();
        const int s = (int)(it & (ND - 1));
        mbar_wait(&dfull[s], (it / ND) & 1u);
        const uint32_t total = sDesc[s].total;

        // which voxels the tile covers comes with its descriptor (no integer divisions per warp and tile)
        const uint32_t origin = sDesc[s].origin;
        const int mol = (int)sDesc[s].mol;
        const int x0 = (int)(origin & 1023u), y0 = (int)((origin >> 10) & 1023u), z0 = (int)(origin >> 20);
        const int z1 = min(D, z0 + P.tz);
        if (total != 0u && total <= SC) {
            char* const out_mol = reinterpret_cast<char*>(P.out) + (size_t)mol * P.Cout * plane * es;
            const int ncells = kCellsXY * ((z1 - z0 + CZ - 1) / CZ);
            mbar_wait(&full[s], (it / ND) & 1u);
            const uint32_t sbq = sq + (uint32_t)sOff[s] * 16u;   // shared address of this tile's entries (ES4 float4 words each)
            const uint32_t ebytes = (uint32_t)ES4 * 16u;
            const uint32_t wAq = sq + (uint32_t)(WA0 + warp * kWarpList) * 16u;                       // warp list: record A
            const uint32_t wIq = sq + (uint32_t)(WA0 + NW * kWarpList) * 16u + (uint32_t)(warp * kWarpList) * 2u;   // entry index
            // hit cache (MULTI): lane-major float4 weights, then the entries' shared addresses
            const uint32_t cw0 = sq + (uint32_t)(WA0 + NW * kWarpList) * 16u + (uint32_t)(NW * kWarpList) * 2u;
            const uint32_t cwq = cw0 + (uint32_t)(warp * kPipeHitCache * 32) * 16u + (uint32_t)lane * 16u;
            const uint32_t ceq = cw0 + (uint32_t)(NW * kPipeHitCache * 32) * 16u + (uint32_t)(warp * kPipeHitCache * 32) * 4u + (uint32_t)lane * 4u;
            // cells are handed out dynamically: the producer warp and warps that drew heavy cells simply take fewer
            for (;;) {
                int cell = 0;
                if (lane == 0) cell = atomicAdd(&sNext[s], 1);
                cell = __shfl_sync(0xffffffffu, cell, 0);
                if (cell >= ncells) break;
                const int cz = cell / kCellsXY, cxy = cell % kCellsXY;
                const int cxl = cxy / NCY, cyl = cxy % NCY;
                const int lx = cxl * RX + rx, ly = cyl * RY + ry, lzv = cz * CZ + zq * 4;
                const int x = x0 + lx, y = y0 + ly, z = z0 + lzv;
                const bool valid = x < D && y < D && z < z1;
                const uint32_t lane_key = (uint32_t)lx | ((uint32_t)ly << 8);
                const float ox = (float)lx * resf, oy = (float)ly * resf;
                float oz[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) oz[k] = (float)(z + k) * resf;   // grid-absolute, like the entries' z
                const int base0 = cz > 0 ? (int)sDesc[s].lend[cz - 1] : 0;
                const int end = (int)sDesc[s].lend[cz];
                int base = base0, wn = 0;
                uint32_t mask_lo = 0u, mask_hi = 0u;
                float acc[CH][4];
                bool single = false;   // the whole layer fits one warp list

                // 1. warp filter over this cell's layer: atoms whose mask names the cell, order kept;
                // 2. every lane tests the warp list against the nearest of its 4 voxels -> hit bitmasks
                auto build_list = [&]() {
                    wn = 0;
                    while (base < end && wn <= kWarpList - 32) {
                        const int i = base + lane;
                        const bool in = (i < end) && ((lds32(sbq + (uint32_t)i * ebytes + 36u) >> cxy) & 1u);
                        const uint32_t m = __ballot_sync(0xffffffffu, in);
                        if (in) {
                            const int pos = wn + __popc(m & ((1u << lane) - 1u));
                            sts128(wAq + (uint32_t)pos * 16u, lds128(sbq + (uint32_t)i * ebytes));
                            sts16(wIq + (uint32_t)pos * 2u, (uint32_t)i);
                        }
                        wn += __popc(m);
                        base += 32;
                    }
                    __syncwarp();
                    mask_lo = 0u; mask_hi = 0u;
                    if (valid) {
                        auto near_hit = [&](const float4 A) -> bool {
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float tz_ = A.z - oz[0];
                            const float dzc = fmaf(-resf, fminf(fmaxf(rintf(tz_ * inv_res), 0.f), 3.f), tz_);
                            return fmaf(dzc, dzc, fmaf(dx, dx, dy * dy)) <= A.w;
                        };
                        const int n_lo = min(wn, 32);
#pragma unroll 4
                        for (int j = 0; j < n_lo; ++j)
                            if (near_hit(lds128(wAq + (uint32_t)j * 16u))) mask_lo |= 1u << j;
#pragma unroll 4
                        for (int j = 32; j < wn; ++j)
                            if (near_hit(lds128(wAq + (uint32_t)j * 16u))) mask_hi |= 1u << (j - 32);
                    }
                };
                // adds one hit (weights w of this lane's 4 voxels, entry at shared address eb) to the channel chunk at c0
                auto accumulate = [&](const uint32_t eb, const float (&w)[4], const int c0, const uint32_t typew) {
                    if (MODE == 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[0][k] += w[k];
                    } else if (MODE == 1) {
                        const int ct = (int)typew - c0;
#pragma unroll
                        for (int c = 0; c < CH; ++c)
#pragma unroll
                            for (int k = 0; k < 4; ++k) acc[c][k] += (ct == c) ? w[k] : 0.f;
                    } else if (CH >= 4) {
                        const uint32_t fb = eb + 48u + (uint32_t)c0 * 4u;
#pragma unroll
                        for (int c4 = 0; c4 < CH; c4 += 4) {
                            const float4 fv = lds128(fb + (uint32_t)c4 * 4u);
                            const float f[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {   // packed FMA over voxel pairs: the accumulators of a channel stay one
                                ffma2(acc[c4 + c][0], acc[c4 + c][1], w[0], w[1], f[c]);   // float4 in register order, so the
                                ffma2(acc[c4 + c][2], acc[c4 + c][3], w[2], w[3], f[c]);   // stores need no shuffles
                            }
                        }
                    } else {
                        const float f = __uint_as_float(lds32(eb + 48u + (uint32_t)c0 * 4u));
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[0][k] = fmaf(f, w[k], acc[0][k]);
                    }
                };
                // 3. lane-private walk over the set bits, ascending (fixed fp32 summation order); the trip count
                //    is the largest hit count of the warp.  MULTI: chunk 0 caches each hit's weights, later chunks replay them.
                auto walk = [&](uint32_t mlo, uint32_t mhi, const int c0) {
                    const int mycnt = __popc(mlo) + __popc(mhi);
                    const int nmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)mycnt);
                    const bool cached = MULTI && single && nmax <= kPipeHitCache;   // warp-uniform
                    if (MULTI && cached && c0 != P.c_begin) {
                        for (int h = 0; h < nmax; ++h) {
                            if (h < mycnt) {
                                const float4 wv = lds128(cwq + (uint32_t)h * 512u);
                                const uint32_t eb = lds32(ceq + (uint32_t)h * 128u);
                                const float w[4] = {wv.x, wv.y, wv.z, wv.w};
                                accumulate(eb, w, c0, MODE == 1 ? lds32(eb + 28u) : 0u);
                            }
                        }
                        __syncwarp();
                        return;
                    }
                    for (int h = 0; h < nmax; ++h) {
                        if ((mlo | mhi) != 0u) {
                            int j;
                            if (mlo != 0u) { j = __ffs((int)mlo) - 1; mlo &= mlo - 1u; }
                            else { j = 31 + __ffs((int)mhi); mhi &= mhi - 1u; }
                            const uint32_t eb = sbq + lds16(wIq + (uint32_t)j * 2u) * ebytes;   // this atom's staged entry
                            const float4 A = lds128(wAq + (uint32_t)j * 16u);
                            const float4 Bv = lds128(eb + 16u);
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float dxy = fmaf(dx, dx, dy * dy);
                            // trigger: |s - r^2| <= tau, widened by the rounding error of the fp32 midpoint / half-width
                            // (see the tile form), for any of the 4 voxels -> those voxels replay in fp64
                            const float r2c = 0.5f * (A.w + Bv.x), tauh = fmaf(4e-7f, r2c, 0.505f * (A.w - Bv.x));
                            float sk[4], w[4], dmin = 3.0e38f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float dz = A.z - oz[k];
                                sk[k] = fmaf(dz, dz, dxy);
                                w[k] = (sk[k] < Bv.x) ? (BINARY ? 1.0f : fast_exp2(sk[k] * Bv.y)) : 0.f;
                                dmin = fminf(dmin, fabsf(sk[k] - r2c));
                            }
                            const bool band = dmin <= tauh;
                            const uint32_t forb = __float_as_uint(Bv.z);
                            if (band || forb != kNoForb) {   // rare: tolerance band, or a block-cull plane in reach
                                bool off[4];
                                {   // block-cull emulation: voxels on the atom's forbidden planes take nothing from it
                                    const uint32_t tx = forb ^ lane_key;
                                    const bool row_off = (tx & 0xFFu) == 0u || (tx & 0xFF00u) == 0u;
                                    const int dzf = (int)(forb >> 16) - z;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) off[k] = row_off || dzf == k;
                                }
                                if (band) {
                                    const int n = (int)lds32(eb + 32u);
                                    const float r32 = (MODE == 1) ? P.recs[n].r : Bv.w;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        if (sk[k] >= Bv.x && !off[k]) {
                                            const bool hit = sk[k] <= A.w && exact_hit(P.recs + n, r32, x, y, z + k, P.res, P.half_width);
                                            w[k] = hit ? (BINARY ? 1.0f : fast_exp2(sk[k] * Bv.y)) : 0.f;
                                        }
                                    }
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) w[k] = off[k] ? 0.f : w[k];
                            }
                            if (MULTI && cached) {
                                sts128(cwq + (uint32_t)h * 512u, make_float4(w[0], w[1], w[2], w[3]));
                                sts32(ceq + (uint32_t)h * 128u, eb);
                            }
                            accumulate(eb, w, c0, __float_as_uint(Bv.w));
                        }
                    }
                    __syncwarp();
                };
                auto clear = [&]() {
#pragma unroll
                    for (int c = 0; c < CH; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;
                };
                auto store = [&](const int c0) {
                    store_lane<CH, O16, CL>(P, out_mol, plane, D, x, y, z, c0, acc, valid);
                };

                // When the whole layer fits one warp list (the common case) its hit masks serve every channel chunk;
                // otherwise each chunk walks the layer's list rounds again.
                build_list();
                single = base >= end;
                for (int c0 = P.c_begin; c0 < P.c_end; c0 += CH) {
                    clear();
                    if (!single && c0 != P.c_begin) { base = base0; build_list(); }
                    for (;;) {
                        walk(mask_lo, mask_hi, c0);
                        if (base >= end) break;
                        build_list();
                    }
                    store(c0);
                }
                if (tid == 0) produce(false);   // keep the ring topped up without ever blocking
            }
        } else if (total == 0u) {
            zero_tile<O16, kPipeThreads, CL>(P, reinterpret_cast<char*>(P.out) + (size_t)mol * P.Cout * plane * es, plane, D, x0, y0, z0, z1, tid);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}

}  // namespace mvx
