// mvx_vox_ws.cuh — EXPERIMENT (not part of the default build; compile with -DMVX_WITH_WS, select with MVX_WS=4|8|20):
// the warp-specialised variant of the pipelined voxelize form.  Correct (GPU parity suite green with MVX_WS set) but
// slower than mvx_voxelize_pipe_kernel on B200 — measurements and analysis in profiles/r2b/warp_specialisation.txt.
#pragma once
#include "mvx_vox_kernels.cuh"

namespace mvx {

// Job-slot generations of the warp-specialised form: a slot's counter is the number of jobs consumed from it so far.
__device__ __forceinline__ uint32_t lds_acquire(uint32_t a) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(uint32_t a, uint32_t v) {
    asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
template <int TAG>
static __device__ __noinline__ void gen_wait_slow(uint32_t a, uint32_t want) {   // bounded like mbar_wait_slow
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spins = 1;; ++spins) {
        if (lds_acquire(a) == want) return;
        __nanosleep(64);
        if ((spins & 4095u) == 0u) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 20000000000ull) __trap();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// voxelize, "pipelined" form, WARP-SPECIALISED (dense batches).  Same tiles, ring, bulk copies and cell arithmetic as
// mvx_voxelize_pipe_kernel, but the two halves of a cell's work run on different warps of the persistent CTA:
//   * NB list-builder warps (setmaxnreg.dec: few registers) draw cells, run the warp filter and the near test, and put
//     a JOB — the cell, the compacted entry indices, the per-lane hit masks — into a shared-memory job ring
//     (tickets from two shared counters, an mbarrier pair per slot).  Thread 0 (builder warp 0) is the bulk-copy producer.
//   * NWK accumulator warps (setmaxnreg.inc: the 64 accumulators + the hit walk) take jobs in ticket order, walk the hits
//     of their lanes in ascending atom order, and store.  They never build lists.
// The CTA's register file is split by role instead of evenly: more resident warps where the work is latency-bound.
// A cell whose layer does not fit one list (rare) travels as a "slow" job: the accumulator warp walks ALL of the layer's
// entries in rounds of 64 with every bit set — the weight computation itself rejects the misses, so the near test is only
// ever an optimisation.  Tile release: builders arrive on the tile's "empty" barrier themselves; the accumulator warp
// that completes the tile's last cell (a shared counter) arrives for all of them.
// ---------------------------------------------------------------------------------------------
template <int MODE, int CH, bool BINARY, bool O16, bool MULTI, int NB, int NWK, int RB, int RW>
__global__ void __launch_bounds__((NB + NWK) * 32, 1) mvx_voxelize_ws_kernel(const VoxParams P, const unsigned ntiles) {
    constexpr int LPR = 4, RX = kCellX, RY = kCellY, CZ = kCellZ;
    constexpr int NCY = kTile / RY;
    constexpr int NJ = kWsJobs;
    constexpr int ND = kPipeSlots;
    constexpr uint32_t JB = (uint32_t)kWsJobBytes;
    static_assert(NB % 4 == 0 && NWK % 4 == 0, "roles are whole warpgroups");
    static_assert((NJ & (NJ - 1)) == 0, "job ring size is a power of two");
    const int Q = P.pipe_q;

    extern __shared__ __align__(128) float4 smem_q[];
    float4* const ring = smem_q;
    TileDesc* const sDesc = reinterpret_cast<TileDesc*>(smem_q + Q);
    uint64_t* const full = reinterpret_cast<uint64_t*>(sDesc + ND);
    uint64_t* const empty = full + ND;
    uint64_t* const dfull = empty + ND;
    int* const sOff = reinterpret_cast<int*>(dfull + ND);
    int* const sNext = sOff + ND;     // slot s: next cell to hand out
    int* const sDone = sNext + ND;    // slot s: cells completed by the accumulator warps
    int* const ticks = sDone + ND;    // [0] next job ticket to fill, [1] next job ticket to take
    unsigned char* const jobs = reinterpret_cast<unsigned char*>(ticks + 8);
    uint64_t* const jbar = reinterpret_cast<uint64_t*>(jobs + NJ * kWsJobBytes);   // per slot: full barrier (8 B), generation (4 B), pad
    uint32_t sq = smem_u32(smem_q);
    asm volatile("" : "+r"(sq));
    const uint32_t jobs_q = sq + (uint32_t)Q * 16u + (uint32_t)(ND * ((int)sizeof(TileDesc) + 3 * 8 + 3 * 4) + 32);
    const uint32_t jgen_q = jobs_q + (uint32_t)(NJ * kWsJobBytes) + 8u;    // slot i's generation: jgen_q + 16 i
    const uint32_t bwa_q = jobs_q + (uint32_t)(NJ * (kWsJobBytes + 16));   // builders' record lists
    const uint32_t cache_q = bwa_q + (uint32_t)(NB * kWsList * 16);        // hit cache (MULTI)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = P.dim;
    const size_t plane = (size_t)D * D * D;
    constexpr int es = O16 ? 2 : 4;
    const int ES4 = P.es4;
    const uint32_t ebytes = (uint32_t)ES4 * 16u;
    const uint32_t SC = (uint32_t)P.pipe_sc;
    const unsigned G = gridDim.x;
    const float resf = (float)P.res;
    const int row = lane / LPR, zq = lane % LPR;
    const int rx = row / RY, ry = row % RY;

    if (tid == 0) {
        for (int i = 0; i < ND; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NB + 1); mbar_init(&dfull[i], 1); }
        for (int i = 0; i < NJ; ++i) { mbar_init(&jbar[i * 2], 1); jbar[i * 2 + 1] = 0ull; }
        ticks[0] = 0; ticks[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp < NB) {
        // =================================== list builders ===================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RB));
        const float inv_res = 1.0f / resf;
        // ---- producer (thread 0): as in mvx_voxelize_pipe_kernel ----
        unsigned next_k = 0, rel_k = 0;
        int head = 0;
        auto issue_desc = [&](unsigned k) {
            const unsigned t = blockIdx.x + k * G;
            if (t < ntiles) {
                const int s = (int)(k & (ND - 1));
                mbar_arrive_expect_tx(&dfull[s], (uint32_t)sizeof(TileDesc));
                bulk_g2s(&sDesc[s], P.tdesc + t, (uint32_t)sizeof(TileDesc), &dfull[s]);
            }
        };
        auto produce = [&](bool block) -> bool {
            const unsigned k = next_k;
            if (blockIdx.x + k * G >= ntiles) return false;
            const int s = (int)(k & (ND - 1));
            auto release_oldest = [&]() -> bool {
                uint64_t* const eb = &empty[rel_k & (ND - 1)];
                const uint32_t epar = (rel_k / ND) & 1u;
                if (block) mbar_wait(eb, epar);
                else if (!mbar_test(eb, epar)) return false;
                ++rel_k;
                return true;
            };
            while (next_k - rel_k > (unsigned)(ND - 2))
                if (!release_oldest()) return false;
            mbar_wait(&dfull[s], (k / ND) & 1u);
            const unsigned long long start = sDesc[s].start;
            const uint32_t total = sDesc[s].total;
            const int need = (total > 0u && total <= SC) ? (int)total * ES4 : 0;
            int at;
            for (;;) {
                if (next_k == rel_k) { at = 0; break; }
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
            sDone[s] = 0;
            if (need > 0) {
                const uint32_t bytes = (uint32_t)need * (uint32_t)sizeof(float4);
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(ring + at, P.lent + start * (unsigned long long)ES4, bytes, &full[s]);
            } else {
                mbar_arrive(&full[s]);
            }
            head = at + need;
            issue_desc(k + 1);
            next_k = k + 1;
            return true;
        };
        if (tid == 0) issue_desc(0);

        const uint32_t wAq = bwa_q + (uint32_t)(warp * kWsList) * 16u;   // this builder's record list (near test)
        // takes the next job ticket and waits until its slot of the ring is free, i.e. until every earlier job of the
        // slot has been consumed (a generation count, not a parity: tickets may run ahead of a slow job by more than one
        // lap of the ring); returns the slot
        auto claim = [&]() -> uint32_t {
            uint32_t t = 0u;
            if (lane == 0) {
                t = (uint32_t)atomicAdd(&ticks[0], 1);
                const uint32_t ga = jgen_q + (t & (uint32_t)(NJ - 1)) * 16u;
                if (lds_acquire(ga) != t / (uint32_t)NJ) gen_wait_slow<0>(ga, t / (uint32_t)NJ);
            }
            t = __shfl_sync(0xffffffffu, t, 0);
            __syncwarp();   // lane 0's acquire orders the whole warp's writes to the slot
            return t & (uint32_t)(NJ - 1);
        };
        auto publish = [&](const uint32_t slot) {   // after the job's words are written by the warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&jbar[slot * 2u]);
        };

        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
            if (tid == 0) {
                while (next_k <= it) produce(true);
                while (produce(false)) {}
            }
            __syncwarp();
            const int s = (int)(it & (ND - 1));
            const uint32_t par = (it / ND) & 1u;
            mbar_wait(&dfull[s], par);
            const uint32_t total = sDesc[s].total;
            const uint32_t origin = sDesc[s].origin;
            const int x0 = (int)(origin & 1023u), y0 = (int)((origin >> 10) & 1023u), z0 = (int)(origin >> 20);
            const int z1 = min(D, z0 + P.tz);
            const bool has_cells = total != 0u && total <= SC;
            if (has_cells) {
                const int ncells = kCellsXY * ((z1 - z0 + CZ - 1) / CZ);
                // tag: slot, tile iteration (8 bits: tells two tiles of one slot apart for 256 iterations), cells of the tile
                const uint32_t tag = (uint32_t)s | ((it & 0xFFu) << 8) | ((uint32_t)ncells << 24);
                mbar_wait(&full[s], par);
                const uint32_t sbq = sq + (uint32_t)sOff[s] * 16u;
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
                    const float ox = (float)lx * resf, oy = (float)ly * resf, oz0 = (float)z * resf;
                    int base = cz > 0 ? (int)sDesc[s].lend[cz - 1] : 0;
                    const int end = (int)sDesc[s].lend[cz];
                    const uint32_t slot = claim();
                    const uint32_t jq = jobs_q + slot * JB;
                    // 1. warp filter over this cell's layer: atoms whose mask names the cell, order kept
                    int wn = 0;
                    while (base < end && wn <= kWsList - 32) {
                        const int i = base + lane;
                        const bool in = (i < end) && ((lds32(sbq + (uint32_t)i * ebytes + 36u) >> cxy) & 1u);
                        const uint32_t m = __ballot_sync(0xffffffffu, in);
                        if (in) {
                            const int pos = wn + __popc(m & ((1u << lane) - 1u));
                            sts128(wAq + (uint32_t)pos * 16u, lds128(sbq + (uint32_t)i * ebytes));
                            sts16(jq + 16u + (uint32_t)pos * 2u, (uint32_t)i);
                        }
                        wn += __popc(m);
                        base += 32;
                    }
                    const bool slow = base < end;   // the layer does not fit one list: the accumulator warp walks all of it
                    __syncwarp();
                    // 2. every lane tests the list against the nearest of its 4 voxels -> hit bitmasks (32 atoms per word)
                    uint32_t hm[4] = {0u, 0u, 0u, 0u};
                    if (valid && !slow) {
                        auto near_hit = [&](const float4 A) -> bool {
                            const float dx = A.x - ox, dy = A.y - oy;
                            const float tz_ = A.z - oz0;
                            const float dzc = fmaf(-resf, fminf(fmaxf(rintf(tz_ * inv_res), 0.f), 3.f), tz_);
                            return fmaf(dzc, dzc, fmaf(dx, dx, dy * dy)) <= A.w;
                        };
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int nq = min(wn - 32 * q, 32);   // warp-uniform
                            uint32_t mq = 0u;
#pragma unroll 4
                            for (int j = 0; j < nq; ++j)
                                if (near_hit(lds128(wAq + (uint32_t)(32 * q + j) * 16u))) mq |= 1u << j;
                            hm[q] = mq;
                        }
                    }
                    sts128(jq + 16u + (uint32_t)kWsList * 2u + (uint32_t)lane * 16u,
                           make_float4(__uint_as_float(hm[0]), __uint_as_float(hm[1]), __uint_as_float(hm[2]), __uint_as_float(hm[3])));
                    if (lane == 0)
                        sts128(jq, make_float4(__uint_as_float(tag), __uint_as_float((uint32_t)cell), __uint_as_float(slow ? 1u : 0u), 0.f));
                    publish(slot);
                    if (tid == 0) produce(false);
                }
            } else if (total == 0u) {
                zero_tile<O16, NB * 32, false>(P, reinterpret_cast<char*>(P.out) + (size_t)sDesc[s].mol * P.Cout * plane * es, plane, D, x0, y0, z0, z1, tid);
            }
            // end of tile for this builder; a tile without cells has no accumulator-side arrival: thread 0 stands in
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (tid == 0 && !has_cells) mbar_arrive(&empty[s]);
        }
        // quit jobs: NWK in total, each builder's after its own last cell job
        for (int i = warp; i < NWK; i += NB) {
            const uint32_t slot = claim();
            if (lane == 0) sts128(jobs_q + slot * JB, make_float4(__uint_as_float(0xFFFFFFFFu), 0.f, 0.f, 0.f));
            publish(slot);
        }
        return;
    }

    // =================================== accumulator warps ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RW));
    const int wk = warp - NB;
    // hit cache (MULTI): lane-major float4 weights, then the entries' shared addresses
    const uint32_t cwq = cache_q + (uint32_t)(wk * kPipeHitCache * 32) * 16u + (uint32_t)lane * 16u;
    const uint32_t ceq = cache_q + (uint32_t)(NWK * kPipeHitCache * 32) * 16u + (uint32_t)(wk * kPipeHitCache * 32) * 4u + (uint32_t)lane * 4u;
    uint32_t cur_tag = 0xFFFFFFFFu;
    uint32_t sbq = 0u;
    int x0 = 0, y0 = 0, z0 = 0, z1 = 0;
    char* out_mol = nullptr;

    for (;;) {
        uint32_t t = 0u;
        if (lane == 0) t = (uint32_t)atomicAdd(&ticks[1], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        const uint32_t slot = t & (uint32_t)(NJ - 1);
        const uint32_t jq = jobs_q + slot * JB;
        mbar_wait<1>(&jbar[slot * 2u], (t / (uint32_t)NJ) & 1u);
        const float4 hdr = lds128(jq);
        const uint32_t tag = __float_as_uint(hdr.x);
        if (tag == 0xFFFFFFFFu) {   // quit (warp-uniform)
            __syncwarp();
            if (lane == 0) red_release_add(jgen_q + slot * 16u, 1u);
            break;
        }
        const int s = (int)(tag & 0xFFu);
        if (tag != cur_tag) {   // first job of a tile for this warp: its descriptor and entries have landed (the builder saw both)
            const uint32_t par = (((tag >> 8) & 0xFFu) / (uint32_t)ND) & 1u;
            mbar_wait<1>(&dfull[s], par);
            mbar_wait<1>(&full[s], par);
            const uint32_t origin = sDesc[s].origin;
            x0 = (int)(origin & 1023u); y0 = (int)((origin >> 10) & 1023u); z0 = (int)(origin >> 20);
            z1 = min(D, z0 + P.tz);
            out_mol = reinterpret_cast<char*>(P.out) + (size_t)sDesc[s].mol * P.Cout * plane * es;
            sbq = sq + (uint32_t)sOff[s] * 16u;
            cur_tag = tag;
        }
        const int cell = (int)__float_as_uint(hdr.y);
        const bool slow = __float_as_uint(hdr.z) != 0u;
        const int cz = cell / kCellsXY, cxy = cell % kCellsXY;
        const int cxl = cxy / NCY, cyl = cxy % NCY;
        const int lx = cxl * RX + rx, ly = cyl * RY + ry, lzv = cz * CZ + zq * 4;
        const int x = x0 + lx, y = y0 + ly, z = z0 + lzv;
        const bool valid = x < D && y < D && z < z1;
        const uint32_t lane_key = (uint32_t)lx | ((uint32_t)ly << 8);
        const float ox = (float)lx * resf, oy = (float)ly * resf;
        float oz[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) oz[kk] = (float)(z + kk) * resf;
        const uint32_t wIq = jq + 16u;
        uint32_t hm0, hm1, hm2, hm3;   // this lane's hits among the job's list, 32 atoms per word
        {
            const float4 mv = lds128(jq + 16u + (uint32_t)kWsList * 2u + (uint32_t)lane * 16u);
            hm0 = __float_as_uint(mv.x); hm1 = __float_as_uint(mv.y); hm2 = __float_as_uint(mv.z); hm3 = __float_as_uint(mv.w);
        }
        float acc[CH][4];

        auto accumulate = [&](const uint32_t eb, const float (&w)[4], const int c0, const uint32_t typew) {
            if (MODE == 0) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) acc[0][kk] += w[kk];
            } else if (MODE == 1) {
                const int ct = (int)typew - c0;
#pragma unroll
                for (int c = 0; c < CH; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) acc[c][kk] += (ct == c) ? w[kk] : 0.f;
            } else if (CH >= 4) {
                const uint32_t fb = eb + 48u + (uint32_t)c0 * 4u;
#pragma unroll
                for (int c4 = 0; c4 < CH; c4 += 4) {
                    const float4 fv = lds128(fb + (uint32_t)c4 * 4u);
                    const float f[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        ffma2(acc[c4 + c][0], acc[c4 + c][1], w[0], w[1], f[c]);
                        ffma2(acc[c4 + c][2], acc[c4 + c][3], w[2], w[3], f[c]);
                    }
                }
            } else {
                const float f = __uint_as_float(lds32(eb + 48u + (uint32_t)c0 * 4u));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) acc[0][kk] = fmaf(f, w[kk], acc[0][kk]);
            }
        };
        // lane-private walk over the set bits, ascending (fixed fp32 summation order); the trip count is the largest hit
        // count of the warp.  MULTI: chunk 0 caches each hit's weights, later chunks replay them.
        auto walk = [&](uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, const int c0) {
            const int mycnt = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
            const int nmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)mycnt);
            const bool cached = MULTI && !slow && nmax <= kPipeHitCache;   // warp-uniform
            if (MULTI && cached && c0 != P.c_begin) {
                for (int h = 0; h < nmax; ++h) {
                    if (h < mycnt) {
                        const float4 wv = lds128(cwq + (uint32_t)h * 512u);
                        const uint32_t eb = lds32(ceq + (uint32_t)h * 128u);
                        const float w[4] = {wv.x, wv.y, wv.z, wv.w};
                        accumulate(eb, w, c0, MODE == 1 ? lds32(eb + 28u) : 0u);
                    }
                }
                return;
            }
            for (int h = 0; h < nmax; ++h) {
                if ((m0 | m1 | m2 | m3) != 0u) {
                    int j;
                    if (m0 != 0u) { j = __ffs((int)m0) - 1; m0 &= m0 - 1u; }
                    else if (m1 != 0u) { j = 31 + __ffs((int)m1); m1 &= m1 - 1u; }
                    else if (m2 != 0u) { j = 63 + __ffs((int)m2); m2 &= m2 - 1u; }
                    else { j = 95 + __ffs((int)m3); m3 &= m3 - 1u; }
                    const uint32_t eb = sbq + lds16(wIq + (uint32_t)j * 2u) * ebytes;   // this atom's staged entry
                    const float4 A = lds128(eb);
                    const float4 Bv = lds128(eb + 16u);
                    const float dx = A.x - ox, dy = A.y - oy;
                    const float dxy = fmaf(dx, dx, dy * dy);
                    // trigger: |s - r^2| <= tau, widened by the rounding error of the fp32 midpoint / half-width
                    // (see the tile form), for any of the 4 voxels -> those voxels replay in fp64
                    const float r2c = 0.5f * (A.w + Bv.x), tauh = fmaf(4e-7f, r2c, 0.505f * (A.w - Bv.x));
                    float sk[4], w[4], dmin = 3.0e38f;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float dz = A.z - oz[kk];
                        sk[kk] = fmaf(dz, dz, dxy);
                        w[kk] = (sk[kk] < Bv.x) ? (BINARY ? 1.0f : fast_exp2(sk[kk] * Bv.y)) : 0.f;
                        dmin = fminf(dmin, fabsf(sk[kk] - r2c));
                    }
                    const bool band = dmin <= tauh;
                    const uint32_t forb = __float_as_uint(Bv.z);
                    if (band || forb != kNoForb) {   // rare: tolerance band, or a block-cull plane in reach
                        bool off[4];
                        {
                            const uint32_t tx = forb ^ lane_key;
                            const bool row_off = (tx & 0xFFu) == 0u || (tx & 0xFF00u) == 0u;
                            const int dzf = (int)(forb >> 16) - z;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) off[kk] = row_off || dzf == kk;
                        }
                        if (band) {
                            const int na = (int)lds32(eb + 32u);
                            const float r32 = (MODE == 1) ? P.recs[na].r : Bv.w;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (sk[kk] >= Bv.x && !off[kk]) {
                                    const bool hit = sk[kk] <= A.w && exact_hit(P.recs + na, r32, x, y, z + kk, P.res, P.half_width);
                                    w[kk] = hit ? (BINARY ? 1.0f : fast_exp2(sk[kk] * Bv.y)) : 0.f;
                                }
                            }
                        }
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) w[kk] = off[kk] ? 0.f : w[kk];
                    }
                    if (MULTI && cached) {
                        sts128(cwq + (uint32_t)h * 512u, make_float4(w[0], w[1], w[2], w[3]));
                        sts32(ceq + (uint32_t)h * 128u, eb);
                    }
                    accumulate(eb, w, c0, __float_as_uint(Bv.w));
                }
            }
        };

        // A normal job is one round with the builder's list and masks.  A slow job: rounds of 128 consecutive entries of
        // the layer, every bit set (this warp writes the identity indices into its job slot).
        const int lbase = slow ? (cz > 0 ? (int)sDesc[s].lend[cz - 1] : 0) : 0;
        const int lend = slow ? (int)sDesc[s].lend[cz] : 1;
        for (int c0 = P.c_begin; c0 < P.c_end; c0 += CH) {
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) acc[c][kk] = 0.f;
            for (int b = lbase; b < lend; b += kWsList) {
                if (slow) {   // warp-uniform
                    const int nr = min(kWsList, lend - b);
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts16(wIq + (uint32_t)(lane + 32 * q) * 2u, (uint32_t)(b + lane + 32 * q));
                    __syncwarp();
                    auto word = [&](const int q) -> uint32_t {
                        const int nq = nr - 32 * q;
                        return (!valid || nq <= 0) ? 0u : (nq >= 32 ? 0xFFFFFFFFu : (1u << nq) - 1u);
                    };
                    hm0 = word(0); hm1 = word(1); hm2 = word(2); hm3 = word(3);
                }
                walk(hm0, hm1, hm2, hm3, c0);
            }
            store_lane<CH, O16, false>(P, out_mol, plane, D, x, y, z, c0, acc, valid);
        }
        __syncwarp();
        if (lane == 0) {
            red_release_add(jgen_q + slot * 16u, 1u);   // the job slot (indices, masks) may be refilled
            const int ncells = (int)(tag >> 24);
            if (atomicAdd(&sDone[s], 1) == ncells - 1) mbar_arrive(&empty[s]);   // the tile's last cell: release it
        }
    }
}

}  // namespace mvx
