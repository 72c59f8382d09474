// Internal parameter block for the conv1d kernels.
#pragma once

#include "common.cuh"

namespace smb {

struct ConvP {
    int batch, dim, L, width;
    bool silu, reverse;
    const void *x, *dout;
    const float *weight, *bias;
    void *out, *dx;
    float *dweight, *dbias;
    int64_t x_bs, x_ds, out_bs, out_ds, dout_bs, dout_ds, dx_bs, dx_ds, w_ds, w_ws;
};

cudaError_t conv1d_dispatch(const ConvP &p, int dtype, bool bwd, cudaStream_t st);
// 16 positions per thread for 16-bit activations, conv1d_v2.cu; chosen by conv1d_dispatch when SMB_CONV_V2=1
cudaError_t conv1d_v2_dispatch(const ConvP &p, int dtype, bool bwd, cudaStream_t st);
// 4-byte-access variant for 16-bit activations (conv1d_v2.cu); cudaErrorNotSupported = shape / alignment not eligible
cudaError_t seq_permute_v2_dispatch(const void *src, void *dst, int64_t src_rs, int64_t dst_rs, int rows, int L, int ns, int inverse,
                                    int accumulate, int dtype, cudaStream_t st);
// layout.cu: rows x row_bytes strided copy, 16-byte vectors
cudaError_t copy2d_launch(const void *src, int64_t src_pitch_bytes, void *dst, int64_t dst_pitch_bytes, int64_t rows, int64_t row_bytes,
                          cudaStream_t st);
cudaError_t seq_permute_dispatch(const void *src, void *dst, int64_t src_rs, int64_t dst_rs, int rows, int L, int ns,
                                 int inverse, int accumulate, int dtype, cudaStream_t st);

}  // namespace smb
