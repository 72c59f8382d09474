// Internal parameter block for the fused instance-norm kernels.
#pragma once

#include "common.cuh"

namespace smb {

struct NormP {
    int batch, channels;
    int64_t spatial;
    int act, mode2;              // act: 0 none, 1 relu, 2 leaky relu; mode2: 0 none, 1 raw add, 2 normalised add
    float slope, eps;
    int n_cta;
    int64_t rows_per_cta;
    const void *x, *x2, *dy;
    void *y, *dx, *dx2;
    float *stats, *stats2;       // (batch, channels, 2): mean, rstd
    float *partial, *sums;       // workspace
};

// rows per CTA / CTA count: >= 8 CTAs per SM in total, at least 8 row-iterations per thread
inline void norm_plan(int batch, int channels, int64_t spatial, int elem_bytes, int *n_cta, int64_t *rows_per_cta) {
    const int V = 16 / elem_bytes;
    const int CV = channels / V;
    const int RB = 256 / CV;
    const int64_t target = (148 * 8 + batch - 1) / batch;
    int64_t rows = (spatial + target - 1) / target;
    if (rows < (int64_t)RB * 8) rows = (int64_t)RB * 8;
    *rows_per_cta = rows;
    *n_cta = (int)((spatial + rows - 1) / rows);
}

cudaError_t instnorm_dispatch(const NormP &p, int dtype, bool bwd, cudaStream_t st);

}  // namespace smb
