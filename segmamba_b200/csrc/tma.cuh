// TMA (cp.async.bulk.tensor) + mbarrier wrappers for the warp-private tile pipelines, and the host-side tensor-map encoder.
//
// One elected lane arms the warp's mbarrier with the byte count of the tile and issues the bulk tensor copies; every lane then
// waits on the barrier's phase parity.  The 16-bit activation tiles use CU_TENSOR_MAP_SWIZZLE_64B, whose pattern on 64-byte rows
// (16-byte chunk index ^= address bits 7..8) is exactly the (row >> 1) & 3 swizzle of raw_tiles.cuh, provided the tile base is
// 512-byte aligned.  Out-of-range coordinates (rows past the tensor, tokens past either end of the sequence) are zero-filled by
// the hardware, so ragged tiles need no separate path on the load side.
// libcuda is NOT linked: cuTensorMapEncodeTiled is resolved at run time through cudaGetDriverEntryPoint, so the library still
// loads on a machine without a driver (the CPU-only tests do that).  Under SMB_EMU the tensor map holds a plain description and
// tools/simt_emu performs the (deferred, swizzled, zero-filled) copy when a thread waits on the barrier.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace smb {

struct TmapDesc {                 // what a tensor map describes (also the emulator's in-map representation)
    const void *base;
    int rank, elem_bytes, swizzle64;
    uint64_t dims[5];             // elements, dims[0] innermost (unit stride)
    uint64_t strides[5];          // bytes, strides[i] = stride of dims[i + 1]
    uint32_t box[5];
};

#ifdef SMB_EMU
inline cudaError_t tmap_encode(CUtensorMap *m, const TmapDesc &d, int /*dtype*/) {
    static_assert(sizeof(TmapDesc) <= sizeof(CUtensorMap), "emulated tensor map does not fit");
    memset(m, 0, sizeof(*m));
    memcpy(m, &d, sizeof(d));
    return cudaSuccess;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) { emu::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) { emu::mbar_wait(bar, parity); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void tma_load_3d(void *smem, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
    const int c[3] = {c0, c1, c2};
    emu::tma_load(smem, m, c, 3, bar);
}
__device__ __forceinline__ void tma_load_4d(void *smem, const CUtensorMap *m, int c0, int c1, int c2, int c3, uint64_t *bar) {
    const int c[4] = {c0, c1, c2, c3};
    emu::tma_load(smem, m, c, 4, bar);
}
#else
// dtype: 1 = fp16, 2 = bf16 (the activation dtype codes of the C ABI)
inline cudaError_t tmap_encode(CUtensorMap *m, const TmapDesc &d, int dtype) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !p) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    cuuint64_t dims[5], strides[5];
    cuuint32_t box[5], estr[5];
    for (int i = 0; i < d.rank; ++i) { dims[i] = d.dims[i]; box[i] = d.box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < d.rank; ++i) strides[i] = d.strides[i];
    const CUresult r = fn(m, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)d.rank,
                          const_cast<void *>(d.base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          d.swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a tile lands within microseconds, so a phase that has not completed after ~2^22 polls (each try_wait itself
// suspends for a hardware-defined interval) can only mean a mis-armed barrier (expect_tx != bytes delivered) or a faulted copy.
// Trap instead of spinning forever: the launch then fails with an error the host sees, and the box is not held.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    for (unsigned spins = 0;; ++spins) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void *smem, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem)),
                 "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem, const CUtensorMap *m, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(smem)),
                 "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace smb
