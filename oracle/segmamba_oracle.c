/*
 * segmamba_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, sequential recurrences) of the reference algorithms on the SegMamba
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (segmamba_b200/) never does.
 *
 * Parity pinning: the reference ships no golden vectors for this path, so
 * this restatement is pinned against outputs of the reference's OWN pure-PyTorch functions
 * (selective_scan_ref, causal_conv1d_ref, and autograd through them) generated in the build
 * container by oracle/gen_golden.py and committed under tests/golden/.
 *
 * Each function cites the reference lines it restates.  Paths are relative to /root/reference.
 *
 * Arithmetic: ORC_REAL (double unless -DORC_REAL=float).  The double build is the checker; the
 * float build (+OpenMP) is what bench.py times as the CPU "port" baseline.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORC_REAL
#define ORC_REAL double
#endif
typedef ORC_REAL real_t;

#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_real_bytes(void) { return (int)sizeof(real_t); }

ORC_API int orc_num_threads(void) {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* F.softplus with threshold 20: mamba/mamba_ssm/ops/selective_scan_interface.py:106-107,
 * mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:153-156. */
static inline real_t orc_softplus(real_t x) { return x <= (real_t)20 ? (real_t)log1p(exp((double)x)) : x; }
static inline real_t orc_sigmoid(real_t x) { return (real_t)(1.0 / (1.0 + exp(-(double)x))); }

/*
 * Selective scan forward.
 * Restates selective_scan_ref, mamba/mamba_ssm/ops/selective_scan_interface.py:86-152
 * (variable B and C, real A), and the chunk-state output `x` of the CUDA op,
 * mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:236-254 / selective_scan.cpp:307-313.
 *
 * Layouts (all contiguous fp32):
 *   u, delta, z, out_y, out_z : (batch, dim, L)
 *   A : (dim, N)   Bm, Cm : (batch, G, N, L)   D, delta_bias : (dim)  (NULL = absent)
 *   last_state : (batch, dim, N) or NULL
 *   xchunks : (batch, dim, ceil(L/chunk), 2N) or NULL -- entry [2n] = running product of
 *             exp(delta*A) from t=0 to the chunk end, entry [2n+1] = state h at the chunk end.
 * out_y is the pre-gate output y = sum_n C h + D u; out_z = y * silu(z) (only when z != NULL).
 */
ORC_API void orc_selective_scan_fwd(
    const float *u, const float *delta, const float *A, const float *Bm, const float *Cm,
    const float *D, const float *z, const float *delta_bias, int delta_softplus,
    int batch, int dim, int L, int N, int G, int chunk,
    float *out_y, float *out_z, float *last_state, float *xchunks)
{
    const int n_chunks = chunk > 0 ? (L + chunk - 1) / chunk : 0;
    const int rows = batch * dim;
#pragma omp parallel for schedule(dynamic, 1)
    for (int row = 0; row < rows; ++row) {
        const int b = row / dim, d = row % dim;
        const int g = d / (dim / G);
        const float *ur = u + (size_t)row * L, *dr = delta + (size_t)row * L;
        const float *zr = z ? z + (size_t)row * L : NULL;
        const float *Br = Bm + ((size_t)b * G + g) * N * L;
        const float *Cr = Cm + ((size_t)b * G + g) * N * L;
        real_t *h = (real_t *)calloc((size_t)N, sizeof(real_t));
        real_t *ap = (real_t *)malloc((size_t)N * sizeof(real_t));
        for (int n = 0; n < N; ++n) ap[n] = 1;
        const real_t bias = delta_bias ? (real_t)delta_bias[d] : 0;
        const real_t Dv = D ? (real_t)D[d] : 0;
        for (int t = 0; t < L; ++t) {
            real_t dt = (real_t)dr[t] + bias;                                  /* ssi.py:104-105 */
            if (delta_softplus) dt = orc_softplus(dt);                         /* ssi.py:106-107 */
            const real_t uv = (real_t)ur[t];
            real_t y = 0;
            for (int n = 0; n < N; ++n) {
                const real_t a = (real_t)exp((double)(dt * (real_t)A[d * N + n]));  /* ssi.py:121 */
                const real_t bu = dt * (real_t)Br[(size_t)n * L + t] * uv;          /* ssi.py:126 */
                h[n] = a * h[n] + bu;                                               /* ssi.py:134 */
                ap[n] *= a;
                y += h[n] * (real_t)Cr[(size_t)n * L + t];                          /* ssi.py:139 */
            }
            y += uv * Dv;                                                           /* ssi.py:148 */
            if (out_y) out_y[(size_t)row * L + t] = (float)y;
            if (zr && out_z) {
                const real_t zv = (real_t)zr[t];
                out_z[(size_t)row * L + t] = (float)(y * zv * orc_sigmoid(zv));     /* ssi.py:150 */
            }
            if (xchunks && ((t + 1) % chunk == 0 || t == L - 1)) {
                const int c = t / chunk;
                float *xc = xchunks + ((size_t)row * n_chunks + c) * 2 * N;
                for (int n = 0; n < N; ++n) { xc[2 * n] = (float)ap[n]; xc[2 * n + 1] = (float)h[n]; }
            }
        }
        if (last_state) for (int n = 0; n < N; ++n) last_state[(size_t)row * N + n] = (float)h[n];
        free(h); free(ap);
    }
}

/*
 * Selective scan backward (variable B, C; real A).  Restates the math of
 * mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:146-489, which
 * is what autograd through selective_scan_ref produces.
 *
 *   dout : gradient w.r.t. the op's returned tensor (out_z when z != NULL, else y).
 *   du, ddelta, dz : (batch, dim, L);  dA : (dim, N);  dB, dC : (batch, G, N, L) fp32, summed
 *   over the channels of each group;  dD, ddelta_bias : (dim).  Any output pointer may be NULL.
 */
ORC_API void orc_selective_scan_bwd(
    const float *u, const float *delta, const float *A, const float *Bm, const float *Cm,
    const float *D, const float *z, const float *delta_bias, int delta_softplus,
    const float *dout,
    int batch, int dim, int L, int N, int G,
    float *du, float *ddelta, float *dA, float *dB, float *dC, float *dD, float *ddelta_bias,
    float *dz)
{
    const size_t bc_elems = (size_t)batch * G * N * L;
    if (dA) memset(dA, 0, sizeof(float) * (size_t)dim * N);
    if (dB) memset(dB, 0, sizeof(float) * bc_elems);
    if (dC) memset(dC, 0, sizeof(float) * bc_elems);
    if (dD) memset(dD, 0, sizeof(float) * (size_t)dim);
    if (ddelta_bias) memset(ddelta_bias, 0, sizeof(float) * (size_t)dim);
    /* double accumulators for the cross-channel reductions; channels run in parallel under OpenMP and
     * add into dB/dC with atomics (summation order varies at the 1e-16 level only). */
    double *dBacc = dB ? (double *)calloc(bc_elems, sizeof(double)) : NULL;
    double *dCacc = dC ? (double *)calloc(bc_elems, sizeof(double)) : NULL;
    const int dpg = dim / G;
#pragma omp parallel for schedule(dynamic, 1)
    for (int d = 0; d < dim; ++d) {
        const int g = d / dpg;
        real_t *hs = (real_t *)malloc(sizeof(real_t) * (size_t)L * N);   /* h_t,n           */
        real_t *as = (real_t *)malloc(sizeof(real_t) * (size_t)L * N);   /* a_t,n           */
        real_t *dts = (real_t *)malloc(sizeof(real_t) * (size_t)L);      /* softplus'd delta */
        real_t *gs = (real_t *)malloc(sizeof(real_t) * (size_t)L);       /* grad wrt y      */
        real_t *lam = (real_t *)malloc(sizeof(real_t) * (size_t)N);
        {
            double dA_acc[256]; for (int n = 0; n < N; ++n) dA_acc[n] = 0;
            double dD_acc = 0, dbias_acc = 0;
            for (int b = 0; b < batch; ++b) {
                const size_t row = (size_t)b * dim + d;
                const float *ur = u + row * L, *dr = delta + row * L, *gor = dout + row * L;
                const float *zr = z ? z + row * L : NULL;
                const float *Br = Bm + ((size_t)b * G + g) * N * L;
                const float *Cr = Cm + ((size_t)b * G + g) * N * L;
                const real_t bias = delta_bias ? (real_t)delta_bias[d] : 0;
                const real_t Dv = D ? (real_t)D[d] : 0;
                /* forward recompute (bwd_kernel.cuh:264-268) */
                for (int n = 0; n < N; ++n) lam[n] = 0;   /* reuse as running h */
                for (int t = 0; t < L; ++t) {
                    real_t dt = (real_t)dr[t] + bias;
                    if (delta_softplus) dt = orc_softplus(dt);
                    dts[t] = dt;
                    const real_t uv = (real_t)ur[t];
                    real_t y = uv * Dv;
                    for (int n = 0; n < N; ++n) {
                        const real_t a = (real_t)exp((double)(dt * (real_t)A[d * N + n]));
                        lam[n] = a * lam[n] + dt * (real_t)Br[(size_t)n * L + t] * uv;
                        as[(size_t)t * N + n] = a;
                        hs[(size_t)t * N + n] = lam[n];
                        y += lam[n] * (real_t)Cr[(size_t)n * L + t];
                    }
                    /* gate (bwd_kernel.cuh:171-207) */
                    real_t gy = (real_t)gor[t];
                    if (zr) {
                        const real_t zv = (real_t)zr[t], sg = orc_sigmoid(zv);
                        if (dz) dz[row * L + t] = (float)(gy * y * sg * ((real_t)1 + zv * ((real_t)1 - sg)));
                        gy *= zv * sg;
                    }
                    gs[t] = gy;
                }
                /* reverse adjoint scan (bwd_kernel.cuh:244-294) */
                for (int n = 0; n < N; ++n) lam[n] = 0;
                for (int t = L - 1; t >= 0; --t) {
                    const real_t dt = dts[t], uv = (real_t)ur[t], gy = gs[t];
                    real_t du_t = Dv * gy, ddt = 0;
                    dD_acc += (double)(gy * uv);
                    for (int n = 0; n < N; ++n) {
                        const real_t anext = (t + 1 < L) ? as[(size_t)(t + 1) * N + n] : (real_t)0;
                        const real_t Bv = (real_t)Br[(size_t)n * L + t], Cv = (real_t)Cr[(size_t)n * L + t];
                        const real_t l = gy * Cv + anext * lam[n];
                        lam[n] = l;
                        const real_t h = hs[(size_t)t * N + n];
                        const real_t hm = h - dt * Bv * uv;             /* a_t h_{t-1}       */
                        du_t += l * Bv * dt;                             /* :280-281          */
                        ddt += l * Bv * uv + l * (real_t)A[d * N + n] * hm;   /* :282-283     */
                        dA_acc[n] += (double)(l * dt * hm);              /* :284              */
                        if (dBacc) {
                            const double v = (double)(l * dt * uv);                                         /* :292 */
#pragma omp atomic
                            dBacc[(((size_t)b * G + g) * N + n) * L + t] += v;
                        }
                        if (dCacc) {
                            const double v = (double)(gy * h);                                              /* :294 */
#pragma omp atomic
                            dCacc[(((size_t)b * G + g) * N + n) * L + t] += v;
                        }
                    }
                    if (du) du[row * L + t] = (float)du_t;
                    /* through softplus (bwd_kernel.cuh:439-453) */
                    real_t dd = ddt;
                    if (delta_softplus) {
                        const real_t raw = (real_t)dr[t] + bias;
                        if (raw <= (real_t)20) dd = ddt * orc_sigmoid(raw);
                    }
                    if (ddelta) ddelta[row * L + t] = (float)dd;
                    dbias_acc += (double)dd;
                }
            }
            if (dA) for (int n = 0; n < N; ++n) dA[d * N + n] = (float)dA_acc[n];
            if (dD) dD[d] = (float)dD_acc;
            if (ddelta_bias) ddelta_bias[d] = (float)dbias_acc;
        }
        free(hs); free(as); free(dts); free(gs); free(lam);
    }
    if (dB) { for (size_t i = 0; i < bc_elems; ++i) dB[i] = (float)dBacc[i]; free(dBacc); }
    if (dC) { for (size_t i = 0; i < bc_elems; ++i) dC[i] = (float)dCacc[i]; free(dCacc); }
}

/*
 * Depthwise causal conv1d (+ optional SiLU).  Restates causal_conv1d_ref,
 * causal-conv1d/causal_conv1d/causal_conv1d_interface.py:49-65 and the kernel
 * causal-conv1d/csrc/causal_conv1d_fwd.cu:103-118.
 *   x, out : (batch, dim, L) contiguous;  w : (dim, width);  bias : (dim) or NULL.
 */
ORC_API void orc_causal_conv1d_fwd(const float *x, const float *w, const float *bias, int silu,
                                   int batch, int dim, int L, int width, float *out)
{
    const int rows = batch * dim;
#pragma omp parallel for schedule(static)
    for (int row = 0; row < rows; ++row) {
        const int d = row % dim;
        const float *xr = x + (size_t)row * L;
        for (int t = 0; t < L; ++t) {
            real_t acc = bias ? (real_t)bias[d] : 0;
            for (int k = 0; k < width; ++k) {
                const int src = t - (width - 1 - k);
                if (src >= 0) acc += (real_t)w[d * width + k] * (real_t)xr[src];
            }
            if (silu) acc = acc * orc_sigmoid(acc);
            out[(size_t)row * L + t] = (float)acc;
        }
    }
}

/*
 * Causal conv1d backward.  Restates causal-conv1d/csrc/causal_conv1d_bwd.cu:153-239
 * (last paragraph).  dw : (dim, width), dbias : (dim), dx : (batch, dim, L).
 */
ORC_API void orc_causal_conv1d_bwd(const float *x, const float *w, const float *bias, const float *dout,
                                   int silu, int batch, int dim, int L, int width,
                                   float *dx, float *dw, float *dbias)
{
#pragma omp parallel for schedule(static)
    for (int d = 0; d < dim; ++d) {
        double dw_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double db_acc = 0;
        real_t *gh = (real_t *)malloc(sizeof(real_t) * (size_t)L);
        for (int b = 0; b < batch; ++b) {
            const size_t row = (size_t)b * dim + d;
            const float *xr = x + row * L, *gr = dout + row * L;
            for (int t = 0; t < L; ++t) {
                real_t g = (real_t)gr[t];
                if (silu) {
                    real_t o = bias ? (real_t)bias[d] : 0;
                    for (int k = 0; k < width; ++k) {
                        const int src = t - (width - 1 - k);
                        if (src >= 0) o += (real_t)w[d * width + k] * (real_t)xr[src];
                    }
                    const real_t sg = orc_sigmoid(o);
                    g *= sg * ((real_t)1 + o * ((real_t)1 - sg));          /* bwd.cu:161-163 */
                }
                gh[t] = g;
                db_acc += (double)g;
                for (int k = 0; k < width; ++k) {
                    const int src = t - (width - 1 - k);
                    if (src >= 0) dw_acc[k] += (double)((real_t)xr[src] * g);   /* bwd.cu:216-222 */
                }
            }
            if (dx) for (int t = 0; t < L; ++t) {
                real_t acc = 0;
                for (int k = 0; k < width; ++k) {
                    const int dst = t + (width - 1 - k);
                    if (dst < L) acc += (real_t)w[d * width + k] * gh[dst];     /* bwd.cu:197-204 */
                }
                dx[row * L + t] = (float)acc;
            }
        }
        if (dw) for (int k = 0; k < width; ++k) dw[d * width + k] = (float)dw_acc[k];
        if (dbias) dbias[d] = (float)db_acc;
        free(gh);
    }
}
