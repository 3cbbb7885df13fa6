// mvx_rigid.cuh — the random rigid transform of the reference's forward_* calls, on the device.
//
//   draw_rigid   per-molecule parameters from a counter-based generator (Philox4x32-10; Salmon et al., SC'11) keyed by
//                (seed, global molecule index): the reference's draw — u1, u2, u3 ~ U[0,1) -> unit quaternion
//                (numpy/_quaternion.py:13-21), translation ~ U(-t, t)^3 rounded to fp32 (numpy/transform.py:74-76) — with
//                the device's RNG in place of numpy's global Mersenne Twister.
//   apply_rigid  the reference's arithmetic, operation for operation and without FMA contraction: two quaternion
//                products q.(0,p).q^-1 (numpy/_quaternion.py:28-54) in the dtype numpy would use (fp32 when coordinates
//                and centre are fp32 — python-float quaternion components are weak scalars —, else fp64), then the
//                translation added twice when rotating (numpy/transform.py:56-59) or once (torch/transform.py:56-60).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvx {

enum : int { kTfRotate = 1, kTfTranslate = 2, kTfTranslateOnce = 4 };

struct Rigid {
    double q[4];   // unit quaternion (q0, q1, q2, q3)
    double t[3];   // translation, fp32-representable
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0, 1) from two 32-bit words, the way numpy's legacy generator forms a double
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

constexpr uint32_t kPhiloxDomain = 0x6D767874u;   // "mvxt": keeps this stream apart from any other use of the key

__device__ __forceinline__ void draw_rigid(const uint64_t seed, const uint64_t gmol, const int flags, const double rt, Rigid& R) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t m0 = (uint32_t)gmol, m1 = (uint32_t)(gmol >> 32);
    uint32_t a[4], b[4], c[4];
    philox4x32_10(m0, m1, 0u, kPhiloxDomain, k0, k1, a);
    philox4x32_10(m0, m1, 1u, kPhiloxDomain, k0, k1, b);
    philox4x32_10(m0, m1, 2u, kPhiloxDomain, k0, k1, c);
    R.q[0] = 1.0; R.q[1] = 0.0; R.q[2] = 0.0; R.q[3] = 0.0;   // identity (scalar part first)
    R.t[0] = R.t[1] = R.t[2] = 0.0;
    if (flags & kTfRotate) {   // numpy/_quaternion.py:13-21
        const double u1 = u53(a[0], a[1]), u2 = u53(a[2], a[3]), u3 = u53(b[0], b[1]);
        const double pi2 = 6.283185307179586;   // 2 * math.pi
        const double sq1 = sqrt(__dsub_rn(1.0, u1)), sqr = sqrt(u1);
        double s2, c2, s3, c3;
        sincos(__dmul_rn(pi2, u2), &s2, &c2);
        sincos(__dmul_rn(pi2, u3), &s3, &c3);
        R.q[0] = __dmul_rn(sq1, s2); R.q[1] = __dmul_rn(sq1, c2);
        R.q[2] = __dmul_rn(sqr, s3); R.q[3] = __dmul_rn(sqr, c3);
    }
    if (flags & kTfTranslate) {   // np.random.uniform(-t, t, size=(1, 3)).astype(np.float32), numpy/transform.py:74-76
        const double u[3] = {u53(b[2], b[3]), u53(c[0], c[1]), u53(c[2], c[3])};
#pragma unroll
        for (int k = 0; k < 3; ++k) R.t[k] = (double)(float)__dadd_rn(-rt, __dmul_rn(__dsub_rn(rt, -rt), u[k]));
    }
}

// round-to-nearest arithmetic that nvcc cannot contract into FMAs, in the dtype numpy computes in
template <typename F> struct Rn;
template <> struct Rn<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};
template <> struct Rn<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
};

// p <- the reference's do_transform(p, None, translation, quaternion), p already centred (numpy/voxelizer.py:263-265)
template <typename F>
__device__ __forceinline__ void apply_rigid(F (&p)[3], const Rigid& R, const int flags) {
    using A = Rn<F>;
    if (flags & kTfRotate) {
        const F q0 = (F)R.q[0], q1 = (F)R.q[1], q2 = (F)R.q[2], q3 = (F)R.q[3];
        const F x = p[0], y = p[1], z = p[2], zero = (F)0;
        // qp = quaternion * (0, x, y, z): multiply_quaternion, numpy/_quaternion.py:28-35 (left-to-right evaluation)
        const F a0 = A::sub(A::sub(A::sub(A::mul(q0, zero), A::mul(q1, x)), A::mul(q2, y)), A::mul(q3, z));
        const F a1 = A::sub(A::add(A::add(A::mul(q0, x), A::mul(q1, zero)), A::mul(q2, z)), A::mul(q3, y));
        const F a2 = A::add(A::add(A::sub(A::mul(q0, y), A::mul(q1, z)), A::mul(q2, zero)), A::mul(q3, x));
        const F a3 = A::add(A::sub(A::add(A::mul(q0, z), A::mul(q1, y)), A::mul(q2, x)), A::mul(q3, zero));
        // qp * inverse(quaternion), inverse = (q0, -q1, -q2, -q3)  (:24-25, :48-54); only x, y, z are kept
        const F b0 = q0, b1 = (F)(R.q[1] * -1), b2 = (F)(R.q[2] * -1), b3 = (F)(R.q[3] * -1);
        p[0] = A::sub(A::add(A::add(A::mul(a0, b1), A::mul(a1, b0)), A::mul(a2, b3)), A::mul(a3, b2));
        p[1] = A::add(A::add(A::sub(A::mul(a0, b2), A::mul(a1, b3)), A::mul(a2, b0)), A::mul(a3, b1));
        p[2] = A::add(A::sub(A::add(A::mul(a0, b3), A::mul(a1, b2)), A::mul(a2, b1)), A::mul(a3, b0));
        if ((flags & kTfTranslate) && !(flags & kTfTranslateOnce)) {   // `coords += translation` inside the rotation branch
#pragma unroll
            for (int k = 0; k < 3; ++k) p[k] = A::add(p[k], (F)R.t[k]);
        }
    }
    if (flags & kTfTranslate) {   // `coords = coords + translation`
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = A::add(p[k], (F)R.t[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic ligands of the virtual-screening sweep (SURVEY.md section 8d, cfg3 / cfg4 recipe), generated on the
// device and keyed by the GLOBAL molecule index, so any sharding or chunking of the sweep sees the same molecules:
// V ~ U{vmin..vmax} atoms, a 3-D random walk with `step` Angstrom steps recentred to the origin, coordinates rounded
// to fp32-representable values, types uniform over num_types.  Benchmark / test input only — not on the hot path.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kSynthDomain = 0x6D767873u;   // "mvxs"

__host__ __device__ __forceinline__ int synth_count(uint64_t seed, uint64_t gmol, int vmin, int vmax) {
    uint32_t w[4];
    philox4x32_10((uint32_t)gmol, (uint32_t)(gmol >> 32), 0u, kSynthDomain, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    return vmin + (int)(w[0] % (uint32_t)(vmax - vmin + 1));
}

struct SynthParams {
    unsigned long long seed, first_mol;
    int B, vmin, vmax, num_types, coords_f64;
    double step;
    const int32_t* mol_offsets;   // (B+1), relative to the first molecule; nullptr: only counts are written
    int32_t* counts;              // (B) or nullptr
    void* coords;                 // (N,3) f32 | f64
    int32_t* types;               // (N) or nullptr
};

struct DrawParams {
    unsigned long long seed, offset;
    int B, flags;
    double rt;
    double* out;   // (B,7)
};

}  // namespace mvx
