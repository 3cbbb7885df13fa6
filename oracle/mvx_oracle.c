/*
 * mvx_oracle.c — CPU restatement of molvoxel's numpy backend (the parity oracle).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under molvoxel_b200/ may import, link or call this
 * file; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker, never as the product path.
 *
 * Parity pinning: this restatement is checked bit-for-bit (binary density) and to <=2e-6
 * (Gaussian, libm expf vs numpy's SIMD exp) against outputs of the live reference
 * (library="numpy", precision=32) generated in the build container by
 * tests/golden/make_golden.py and committed under tests/golden/.  The reference's own
 * tests hold no golden vectors for this path.
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/molvoxel/voxelizer/.  The one third-party piece of arithmetic,
 * scipy.spatial.distance.cdist (scipy 1.18.1, un-vendored), is restated from its
 * published algorithm: euclidean distance sqrt((dx*dx + dy*dy) + dz*dz) in fp64 with no
 * FMA contraction (SURVEY.md hazard 4; compile with -ffp-contract=off).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MVXO_MODE_SINGLE 0
#define MVXO_MODE_TYPES 1
#define MVXO_MODE_FEATURES 2

#define MVXO_RADII_SCALAR 0
#define MVXO_RADII_CHANNEL 1
#define MVXO_RADII_ATOM 2

typedef struct {
    double resolution;   /* base/voxelizer.py:26 */
    int32_t dimension;   /* base/voxelizer.py:27 */
    int32_t binary;      /* density_type == "binary" */
    double sigma;        /* base/voxelizer.py:37-38 */
    int32_t radii_mode;  /* MVXO_RADII_* */
    int32_t blockdim;    /* numpy/voxelizer.py:38 (None -> 8) */
} mvxo_spec;

/* axis[i] = i*res - width/2 with width = res*(dim-1): numpy/voxelizer.py:41-43, base/voxelizer.py:28 */
static double axis_at(const mvxo_spec *s, int i, double half_width) {
    double t = (double)i * s->resolution;
    return t - half_width;
}

/* One atom-voxel term: numpy/voxelizer.py:544-560 (+ scipy cdist). */
static inline float pair_term(double px, double py, double pz, double gx, double gy, double gz,
                              float r32, float sigma32, int binary) {
    double dx = px - gx, dy = py - gy, dz = pz - gz;
    double xx = dx * dx, yy = dy * dy, zz = dz * dz;
    double s = (xx + yy) + zz;          /* cdist association, SURVEY.md App. A.3 */
    float d32 = (float)sqrt(s);         /* dist.astype(float32), :545 */
    float dr = d32 / r32;               /* np.divide(dist, radii), :548 */
    if (binary) return (dr <= 1.0f) ? 1.0f : 0.0f; /* :554-555 */
    float q = dr / sigma32;             /* (dr / sigma), :558 */
    float e = expf(-0.5f * (q * q));    /* np.exp(-0.5 * (..)**2), :558 */
    if (dr > 1.0f) e = 0.0f;            /* out_grid[dr > 1.0] = 0, :559 */
    return e;
}

/* The term when the scalar radius is an np.float64 (a strongly typed scalar under NEP 50): np.divide(dist32, r64)
 * promotes to fp64 (:548), so the cutoff compares in fp64 (:555, :559) and the Gaussian is evaluated in fp64 (:558,
 * sigma a python float); the caller narrows on accumulation (out32 += res64, :365/:476) or after the fp64 matmul (:226-235). */
static inline double pair_term_r64(double px, double py, double pz, double gx, double gy, double gz,
                                   double r64, double sigma, int binary) {
    double dx = px - gx, dy = py - gy, dz = pz - gz;
    double xx = dx * dx, yy = dy * dy, zz = dz * dz;
    double s = (xx + yy) + zz;
    float d32 = (float)sqrt(s);         /* dist.astype(float32), :545 */
    double dr = (double)d32 / r64;      /* fp32 / np.float64 -> fp64, :548 */
    if (binary) return (dr <= 1.0) ? 1.0 : 0.0;
    double q = dr / sigma;
    double e = exp(-0.5 * (q * q));
    if (dr > 1.0) e = 0.0;
    return e;
}

/* The same term with precision=64 (numpy/voxelizer.py:34): dist stays fp64 (:545 is a no-op), dr = dist / radii in fp64
 * (:548), np.exp(-0.5 * (dr / sigma) ** 2) in fp64 (:558), cutoff dr > 1.0 (:559) / dr <= 1.0 (:555). */
static inline double pair_term64(double px, double py, double pz, double gx, double gy, double gz,
                                 double r, double sigma, int binary) {
    double dx = px - gx, dy = py - gy, dz = pz - gz;
    double xx = dx * dx, yy = dy * dy, zz = dz * dz;
    double s = (xx + yy) + zz;
    double dr = sqrt(s) / r;
    if (binary) return (dr <= 1.0) ? 1.0 : 0.0;
    double q = dr / sigma;
    double e = exp(-0.5 * (q * q));
    if (dr > 1.0) e = 0.0;
    return e;
}

/*
 * One molecule.  coords: (V,3) f32|f64; center: (3,) f32|f64 or NULL; types: (V,) int32
 * (values already reduced to int16 range by the caller); features: (V,C) f32;
 * radius: python-float scalar; radii: (C,) or (V,) f32.  out: (out_channels, D, D, D) f32.
 * Returns 0, or a negative code on bad arguments.
 *
 * Follows forward_types :240-315, forward_features :97-169, forward_single :370-436.
 */
static int mvxo_forward_p(const mvxo_spec *s, int mode, int V, const void *coords, int coords_f64,
                          const void *center, int center_f64, const int32_t *types, const float *features,
                          int C, double radius, const float *radii, void *out_v, int out_channels, int prec64,
                          int radius_kind) {   /* scalar radius: 0 python float, 1 np.float64, 2 np.float32 */
    const int radius_f64 = radius_kind == 1;
    float *out = (float *)out_v;      /* precision=32 (the default oracle) */
    double *out64 = (double *)out_v;  /* precision=64: get_empty_grid dtype :60-70, features/radii astype(fp64) :127-130 */
    const int D = s->dimension;
    const int bd = s->blockdim > 0 ? s->blockdim : 8;
    const int nb = (D + bd - 1) / bd;                     /* :44 */
    const double res = s->resolution;
    const double width = res * (double)(D - 1);           /* base/voxelizer.py:28 */
    const double half_width = width / 2.0;                /* :42 */
    const double upper = width / 2.0;                     /* base/voxelizer.py:33 */
    const double lower = -1 * upper;                      /* base/voxelizer.py:34 */
    const float sigma32 = (float)s->sigma;                /* python float is weak: fp32 divide */
    const size_t plane = (size_t)D * D * D;

    if (mode == MVXO_MODE_SINGLE && s->radii_mode == MVXO_RADII_CHANNEL) return -1; /* :443 */
    if (mode != MVXO_MODE_SINGLE && C > out_channels) return -2;                    /* :337 */

    /* out init: types/single zero (:279-281, :402-404); features writes every voxel (:158-160,:232-235). */
    memset(out_v, 0, (prec64 ? sizeof(double) : sizeof(float)) * plane * (size_t)out_channels);
    if (V == 0) return 0;
    /* np.float64 scalar radius, precision=32: features go through an fp64 matmul that is narrowed once (:226-235) */
    const int r64 = radius_f64 && !prec64 && s->radii_mode == MVXO_RADII_SCALAR;
    double *acc64 = NULL;
    if (r64 && mode == MVXO_MODE_FEATURES) {
        acc64 = (double *)calloc(plane * (size_t)out_channels, sizeof(double));
        if (!acc64) return -3;
    }

    double *p = (double *)malloc(sizeof(double) * 3 * (size_t)V);
    float *rr = (float *)malloc(sizeof(float) * (size_t)V);   /* per-atom kernel radius */
    int32_t *keep = (int32_t *)malloc(sizeof(int32_t) * (size_t)V);
    int32_t *blist = (int32_t *)malloc(sizeof(int32_t) * (size_t)V);
    double *bounds = (double *)malloc(sizeof(double) * (size_t)(nb > 1 ? nb - 1 : 1));
    if (!p || !rr || !keep || !blist || !bounds) return -3;

    /* prologue: coords - center in the promoted dtype, then fp64 (:263-268) */
    for (int n = 0; n < V; ++n) {
        for (int k = 0; k < 3; ++k) {
            double v;
            if (center == NULL) {
                v = coords_f64 ? ((const double *)coords)[3 * n + k] : (double)((const float *)coords)[3 * n + k];
            } else if (!coords_f64 && !center_f64) {
                float a = ((const float *)coords)[3 * n + k], b = ((const float *)center)[k];
                float d = a - b;
                v = (double)d;
            } else {
                double a = coords_f64 ? ((const double *)coords)[3 * n + k] : (double)((const float *)coords)[3 * n + k];
                double b = center_f64 ? ((const double *)center)[k] : (double)((const float *)center)[k];
                v = a - b;
            }
            p[3 * n + k] = v;
        }
    }

    /* per-atom radius used by the clip and the cull ("atom_size") and by the density kernel */
    const int scalar_form =
        (s->radii_mode == MVXO_RADII_SCALAR) || (mode == MVXO_MODE_FEATURES && s->radii_mode == MVXO_RADII_CHANNEL);
    double size_scalar = radius;   /* scalar form: python float (:138, :286) */
    int thr_f32 = 0;               /* features + channel-wise: radii.max() is np.float32 (:138) */
    if (mode == MVXO_MODE_FEATURES && s->radii_mode == MVXO_RADII_CHANNEL) {
        float m = radii[0];
        for (int c = 1; c < C; ++c) if (radii[c] > m) m = radii[c];
        size_scalar = (double)m;
        thr_f32 = !prec64;   /* precision=64: radii.astype(float64).max() is np.float64, the bounds stay fp64 (:130, :138) */
    }
    /* an np.float32 SCALAR radius: python-float bound -/+ np.float32 is fp32 arithmetic as well (:487-488, NEP 50);
     * the block bounds are np.float64 array elements, so the cull stays fp64 (:504-511) */
    if (s->radii_mode == MVXO_RADII_SCALAR && radius_kind == 2) thr_f32 = 1;
    for (int n = 0; n < V; ++n) {
        if (s->radii_mode == MVXO_RADII_SCALAR) rr[n] = (float)radius;
        else if (s->radii_mode == MVXO_RADII_ATOM) rr[n] = radii[n];
        else if (mode == MVXO_MODE_TYPES) rr[n] = radii[types[n]];   /* :284-285 */
        else rr[n] = (float)size_scalar;                              /* replaced per channel below */
    }

    /* global clip, strict inequalities: _get_overlap :481-494 */
    int nk = 0;
    {
        double lo_thr, hi_thr;
        if (thr_f32) {  /* python float (weak) op np.float32 -> float32 arithmetic (NEP 50) */
            float lo32 = (float)lower - (float)size_scalar;
            float hi32 = (float)upper + (float)size_scalar;
            lo_thr = (double)lo32; hi_thr = (double)hi32;
        } else {
            lo_thr = lower - size_scalar; hi_thr = upper + size_scalar;
        }
        for (int n = 0; n < V; ++n) {
            int ok = 1;
            for (int k = 0; k < 3 && ok; ++k) {
                double v = p[3 * n + k];
                if (scalar_form) {
                    ok = (v > lo_thr) && (v < hi_thr);             /* :487-488 */
                } else {
                    double a = (double)rr[n];
                    double vp = v + a, vm = v - a;
                    ok = (vp > lower) && (vm < upper);             /* :491-492 */
                }
            }
            if (ok) keep[nk++] = n;
        }
    }

    /* bounds[i-1] = axis[i*bd] + res/2, i = 1..nb-1  (:55) */
    for (int i = 1; i < nb; ++i) bounds[i - 1] = axis_at(s, i * bd, half_width) + (res / 2.0);

    /* block loop: itertools.product(range(nb), repeat=3)  (:297-311, :150-165, :418-432) */
    for (int bx = 0; bx < nb; ++bx)
    for (int by = 0; by < nb; ++by)
    for (int bz = 0; bz < nb; ++bz) {
        int bidx[3] = {bx, by, bz};
        /* _get_overlap_blocks :496-527 — per-axis tests ANDed; np.where keeps ascending order */
        int nbk = 0;
        for (int i = 0; i < nk; ++i) {
            int n = keep[i];
            double a = scalar_form ? size_scalar : (double)rr[n];
            int ok = 1;
            if (nb > 1) {
                for (int k = 0; k < 3 && ok; ++k) {
                    double v = p[3 * n + k];
                    int b = bidx[k];
                    if (b > 0) { double t = bounds[b - 1] - a; ok = ok && (v > t); }        /* :507,:510 */
                    if (b < nb - 1) { double t = bounds[b] + a; ok = ok && (v < t); }       /* :504,:511 */
                }
            }
            if (ok) blist[nbk++] = n;
        }
        if (nbk == 0) continue;   /* features zero-fills (:158-160): already zero */

        int x0 = bx * bd, x1 = x0 + bd > D ? D : x0 + bd;
        int y0 = by * bd, y1 = y0 + bd > D ? D : y0 + bd;
        int z0 = bz * bd, z1 = z0 + bd > D ? D : z0 + bd;

        for (int i = 0; i < nbk; ++i) {   /* ascending atom order: :364-365 */
            int n = blist[i];
            double px = p[3 * n], py = p[3 * n + 1], pz = p[3 * n + 2];
            /* speed only: voxels further than rmax + 1 voxel contribute exactly 0 */
            double rmax = (double)rr[n];
            if (scalar_form && size_scalar > rmax) rmax = size_scalar;
            double reach = rmax * 1.0001 + res;
            int ax0 = (int)floor((px - reach + half_width) / res); if (ax0 < x0) ax0 = x0;
            int ax1 = (int)ceil((px + reach + half_width) / res) + 1; if (ax1 > x1) ax1 = x1;
            int ay0 = (int)floor((py - reach + half_width) / res); if (ay0 < y0) ay0 = y0;
            int ay1 = (int)ceil((py + reach + half_width) / res) + 1; if (ay1 > y1) ay1 = y1;
            int az0 = (int)floor((pz - reach + half_width) / res); if (az0 < z0) az0 = z0;
            int az1 = (int)ceil((pz + reach + half_width) / res) + 1; if (az1 > z1) az1 = z1;
            for (int x = ax0; x < ax1; ++x) {
                double gx = axis_at(s, x, half_width);
                for (int y = ay0; y < ay1; ++y) {
                    double gy = axis_at(s, y, half_width);
                    for (int z = az0; z < az1; ++z) {
                        double gz = axis_at(s, z, half_width);
                        size_t vox = ((size_t)x * D + y) * D + z;
                        if (prec64) {   /* the same accumulation in fp64 (array radii are fp32 at the C boundary, widened) */
                            if (mode == MVXO_MODE_FEATURES && s->radii_mode == MVXO_RADII_CHANNEL) {
                                for (int c = 0; c < C; ++c)
                                    out64[(size_t)c * plane + vox] += (double)features[(size_t)n * C + c] *
                                        pair_term64(px, py, pz, gx, gy, gz, (double)radii[c], s->sigma, s->binary);
                            } else {
                                double r64 = (s->radii_mode == MVXO_RADII_SCALAR) ? radius : (double)rr[n];
                                double t = pair_term64(px, py, pz, gx, gy, gz, r64, s->sigma, s->binary);
                                if (mode == MVXO_MODE_TYPES) out64[(size_t)types[n] * plane + vox] += t;
                                else if (mode == MVXO_MODE_SINGLE) out64[vox] += t;
                                else if (t != 0.0)
                                    for (int c = 0; c < C; ++c) out64[(size_t)c * plane + vox] += (double)features[(size_t)n * C + c] * t;
                            }
                            continue;
                        }
                        if (r64) {
                            double t = pair_term_r64(px, py, pz, gx, gy, gz, radius, s->sigma, s->binary);
                            if (mode == MVXO_MODE_TYPES) {
                                float *o = &out[(size_t)types[n] * plane + vox];
                                *o = (float)((double)*o + t);                    /* fp32 += fp64: fp64 add, narrowed (:365) */
                            } else if (mode == MVXO_MODE_SINGLE) {
                                out[vox] = (float)((double)out[vox] + t);        /* :476 */
                            } else if (t != 0.0) {
                                for (int c = 0; c < C; ++c) acc64[(size_t)c * plane + vox] += (double)features[(size_t)n * C + c] * t;
                            }
                            continue;
                        }
                        if (mode == MVXO_MODE_TYPES) {
                            float t = pair_term(px, py, pz, gx, gy, gz, rr[n], sigma32, s->binary);
                            out[(size_t)types[n] * plane + vox] += t;            /* :365 */
                        } else if (mode == MVXO_MODE_SINGLE) {
                            float t = pair_term(px, py, pz, gx, gy, gz, rr[n], sigma32, s->binary);
                            out[vox] += t;                                       /* :476 */
                        } else if (s->radii_mode == MVXO_RADII_CHANNEL) {
                            for (int c = 0; c < C; ++c) {                        /* :213-224 */
                                float t = pair_term(px, py, pz, gx, gy, gz, radii[c], sigma32, s->binary);
                                float f = features[(size_t)n * C + c];
                                float prod = f * t;
                                out[(size_t)c * plane + vox] += prod;
                            }
                        } else {
                            float t = pair_term(px, py, pz, gx, gy, gz, rr[n], sigma32, s->binary);
                            if (t != 0.0f) {
                                for (int c = 0; c < C; ++c) {                    /* :226-235 (sgemm) */
                                    float f = features[(size_t)n * C + c];
                                    float prod = f * t;
                                    out[(size_t)c * plane + vox] += prod;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    if (acc64) {
        for (size_t i = 0; i < plane * (size_t)out_channels; ++i) out[i] = (float)acc64[i];
        free(acc64);
    }
    free(p); free(rr); free(keep); free(blist); free(bounds);
    return 0;
}

int mvxo_forward(const mvxo_spec *s, int mode, int V, const void *coords, int coords_f64,
                 const void *center, int center_f64, const int32_t *types, const float *features,
                 int C, double radius, const float *radii, float *out, int out_channels) {
    return mvxo_forward_p(s, mode, V, coords, coords_f64, center, center_f64, types, features, C, radius, radii, out,
                          out_channels, 0, 0);
}

/* precision=64 variant: out is (out_channels, D, D, D) float64. */
int mvxo_forward64(const mvxo_spec *s, int mode, int V, const void *coords, int coords_f64,
                   const void *center, int center_f64, const int32_t *types, const float *features,
                   int C, double radius, const float *radii, double *out, int out_channels) {
    return mvxo_forward_p(s, mode, V, coords, coords_f64, center, center_f64, types, features, C, radius, radii, out,
                          out_channels, 1, 0);
}

/*
 * Batch driver = B independent reference calls (test/test_time_numpy.py:11-15), molecules in
 * CSR form.  centers: (B,3) or NULL.  radii: (C,) channel-wise, (N,) atom-wise.  out:
 * (B, out_channels, D, D, D).  num_threads > 1 spreads molecules over pthreads (the reference
 * itself is single-threaded; the threaded form is the "all host cores" baseline).
 */
typedef struct {
    const mvxo_spec *s; int mode; int B; const int32_t *mol_offsets; const void *coords; int coords_f64;
    const void *centers; int centers_f64; const int32_t *types; const float *features; int C;
    double radius; const float *radii; float *out; int out_channels;
    int next; int rc; pthread_mutex_t mu; int radius_kind;
} mvxo_job;

static void *mvxo_worker(void *arg) {
    mvxo_job *j = (mvxo_job *)arg;
    const mvxo_spec *s = j->s;
    const size_t per_mol = (size_t)j->out_channels * s->dimension * s->dimension * s->dimension;
    const size_t csz = j->coords_f64 ? sizeof(double) : sizeof(float);
    const size_t zsz = j->centers_f64 ? sizeof(double) : sizeof(float);
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int b = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->B) break;
        int a0 = j->mol_offsets[b], V = j->mol_offsets[b + 1] - a0;
        const void *cb = (const char *)j->coords + csz * 3 * (size_t)a0;
        const void *zb = j->centers ? (const void *)((const char *)j->centers + zsz * 3 * (size_t)b) : NULL;
        const float *rb = j->radii;
        if (j->radii && s->radii_mode == MVXO_RADII_ATOM) rb = j->radii + a0;
        int rc = mvxo_forward_p(s, j->mode, V, cb, j->coords_f64, zb, j->centers_f64,
                                j->types ? j->types + a0 : NULL,
                                j->features ? j->features + (size_t)a0 * j->C : NULL, j->C, j->radius, rb,
                                j->out + per_mol * (size_t)b, j->out_channels, 0, j->radius_kind);
        if (rc != 0) { pthread_mutex_lock(&j->mu); j->rc = rc; pthread_mutex_unlock(&j->mu); }
    }
    return NULL;
}

/* radius_kind: 0 the scalar radius is a python float; 1 an np.float64 (numpy then divides in fp64, :546-548);
 * 2 an np.float32 (fp32 clip bounds, :487-488). */
int mvxo_forward_batch(const mvxo_spec *s, int mode, int B, const int32_t *mol_offsets, const void *coords,
                       int coords_f64, const void *centers, int centers_f64, const int32_t *types,
                       const float *features, int C, double radius, const float *radii, float *out,
                       int out_channels, int num_threads, int radius_kind) {
    mvxo_job j = {s, mode, B, mol_offsets, coords, coords_f64, centers, centers_f64, types, features, C,
                  radius, radii, out, out_channels, 0, 0, PTHREAD_MUTEX_INITIALIZER, radius_kind};
    if (num_threads < 1) num_threads = 1;
    if (num_threads > 256) num_threads = 256;
    if (num_threads > B) num_threads = B > 0 ? B : 1;
    if (num_threads == 1) { mvxo_worker(&j); return j.rc; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < num_threads; ++t)
        if (pthread_create(&th[started], NULL, mvxo_worker, &j) == 0) ++started;
    if (started == 0) mvxo_worker(&j);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    return j.rc;
}

int mvxo_version(void) { return 1; }
