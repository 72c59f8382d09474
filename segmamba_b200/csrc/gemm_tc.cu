// Tensor-core GEMM for the pointwise contractions of the SegMamba hot path (sm_100a: TMA -> shared memory -> tcgen05.mma with
// the accumulator in tensor memory -> tcgen05.ld epilogue).
//
//     D[M, N] (+)= epilogue( A[M, K] . B[N, K]^T )          16-bit operands, fp32 accumulation
//
// replaces the library calls behind   Mamba.in_proj / out_proj (mamba_simple.py:204-208,264),  MlpChannel.fc1 / fc2
// (segmamba.py:81-89),  GSC.proj3 / proj4 (segmamba.py:103-107,121-128),  UnetResBlock.conv3 (dynunet_block.py:66-69,105-109)
// -- 1x1x1 convolutions on channels-last activations are plain (tokens, C_in) x (C_out, C_in)^T products -- and their two
// backward products (data gradient: the same kernel on the transposed weight view; weight gradient: contraction over the token
// axis, both operands MN-major, split over CTAs along K with fp32 atomics).
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes of A (128 x 64) and B (BN x 64) per 64-wide K block into a
//               ring of shared-memory stages (CU_TENSOR_MAP_SWIZZLE_128B), completion on the stage's `full` mbarrier
//   warp 1      allocates tensor memory; one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN, K = 16)
//               four per K block, releases the stage with tcgen05.commit -> `empty`, and after the last K block commits to
//               the accumulator's `tmem_full` barrier.  Two accumulator buffers (2 x BN columns) overlap the epilogue of
//               tile i with the main loop of tile i + 1.
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 16 columns per instruction) -> bias / GELU -> convert -> padded staging tile
//               in shared memory -> coalesced 16-byte global stores (or fp32 atomics for split-K) by the same 128 threads.
// Either operand may be K-major (rows = M or N index, K contiguous) or MN-major (rows = K index, M or N contiguous); the
// shared-memory matrix descriptors and the instruction descriptor carry the difference, the TMA boxes stay (64 elements = 128
// bytes) x rows.  Out-of-range rows / columns of a box are zero-filled by the TMA unit, so ragged M, N and K need no separate
// path on the load side; the store side bounds every access.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "gemm_internal.h"

namespace smb {
void count_launch();

namespace {

constexpr int kBM = 128;            // UMMA M (one CTA, cta_group::1): accumulator row i lives in tensor-memory lane i
constexpr int kBK = 64;             // K block: 64 sixteen-bit elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 6;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a phase that does not complete within ~2^24 polls means a mis-armed barrier or a faulted copy; trap so the
// launch fails visibly instead of holding the GPU
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
        if (spins > (1u << 24)) __trap();
    }
}
// optional profiling (GemmP::prof != nullptr): cycles a role spends inside a barrier wait, summed over the grid
__device__ __forceinline__ void mbar_wait_prof(uint64_t *bar, unsigned parity, unsigned long long *acc) {
    if (!acc) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    atomicAdd(acc, (unsigned long long)(clock64() - t0));
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *m, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem)),
                 "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(unsigned d_tmem, uint64_t a_desc, uint64_t b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4f(float *a, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), leading byte offset >> 4
// at [16,30), stride byte offset >> 4 at [32,46), version 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64)).
//   K-major  tile = rows x 128 B : 8-row groups are 1024 B apart (SBO); the leading offset is unused inside one swizzle row.
//   MN-major tile = 64 K-rows x 128 B per 64 MN elements: 8-K-row groups 1024 B apart (SBO), next 64 MN elements one tile on (LBO).
__device__ __forceinline__ uint64_t make_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct alignas(16) F4 { float x, y, z, w; };

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

template <typename T> __device__ __forceinline__ unsigned pack2(float a, float b);
template <> __device__ __forceinline__ unsigned pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<unsigned *>(&h);
}
template <> __device__ __forceinline__ unsigned pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<unsigned *>(&h);
}

// packed 16-bit pair `old` + (a, b), rounded once
template <typename T> __device__ __forceinline__ unsigned add2(unsigned old, float a, float b);
template <> __device__ __forceinline__ unsigned add2<__nv_bfloat16>(unsigned old, float a, float b) {
    const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162 *>(&old);
    return pack2<__nv_bfloat16>(__low2float(o) + a, __high2float(o) + b);
}
template <> __device__ __forceinline__ unsigned add2<__half>(unsigned old, float a, float b) {
    const __half2 o = *reinterpret_cast<const __half2 *>(&old);
    return pack2<__half>(__low2float(o) + a, __high2float(o) + b);
}

template <typename T> __device__ __forceinline__ unsigned sum2(unsigned a, unsigned b);
template <> __device__ __forceinline__ unsigned sum2<__nv_bfloat16>(unsigned a, unsigned b) {
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162 *>(&a), y = *reinterpret_cast<const __nv_bfloat162 *>(&b);
    return pack2<__nv_bfloat16>(__low2float(x) + __low2float(y), __high2float(x) + __high2float(y));
}
template <> __device__ __forceinline__ unsigned sum2<__half>(unsigned a, unsigned b) {
    const __half2 x = *reinterpret_cast<const __half2 *>(&a), y = *reinterpret_cast<const __half2 *>(&b);
    return pack2<__half>(__low2float(x) + __low2float(y), __high2float(x) + __high2float(y));
}

// ---------------------------------------------------------------------------------------------------------------------------
// TOut: __nv_bfloat16 / __half (store) or float (store, or atomic accumulate when p.atomic)
// ---------------------------------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const GemmP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int BN = p.BN;
    const int a_tile = kBM * 128;                       // bytes per stage
    const int b_rows = p.b_mn ? ((BN + 63) / 64) * 64 : BN;
    const int b_tile = p.b_mn ? ((BN + 63) / 64) * 8192 : BN * 128;
    const int stage_bytes = a_tile + ((b_tile + 1023) & ~1023);
    uint8_t *stage0 = smem;
    uint8_t *staging = smem + (size_t)p.stages * stage_bytes;                 // 8 warps x 32 rows x 144 bytes
    uint64_t *bars = reinterpret_cast<uint64_t *>(staging + kEpiWarps * 32 * 144);
    uint64_t *full = bars, *empty = bars + kMaxStages, *tfull = bars + 2 * kMaxStages, *tempty = bars + 2 * kMaxStages + 2;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + 2 * kMaxStages + 4);
    (void)b_rows;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;
    const long long t_start = clock64();

    const int m_tiles = (p.M + kBM - 1) / kBM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int kb_total = (p.K + kBK - 1) / kBK;
    const int kb_per = (kb_total + p.split_k - 1) / p.split_k;
    const int n_work = m_tiles * n_tiles * p.split_k;

    if (warp == 0) {
        if (lane == 0) {
            unsigned it = 0;
            for (int t = blockIdx.x; t < n_work; t += gridDim.x) {
                const int sp = t % p.split_k, nb = (t / p.split_k) % n_tiles, mb = t / (p.split_k * n_tiles);
                const int kb0 = sp * kb_per, kb1 = min(kb_total, kb0 + kb_per);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % p.stages;
                    const unsigned ph = (it / p.stages) & 1;
                    mbar_wait_prof(&empty[s], ph ^ 1, p.prof ? p.prof + 0 : nullptr);
                    uint8_t *sa = stage0 + (size_t)s * stage_bytes, *sb = sa + a_tile;
                    mbar_expect_tx(&full[s], (unsigned)(a_tile + b_tile));
                    if (!p.a_mn) {
                        tma_load_2d(sa, &tmA, kb * kBK, mb * kBM, &full[s]);
                    } else {
                        tma_load_2d(sa, &tmA, mb * kBM, kb * kBK, &full[s]);
                        tma_load_2d(sa + 8192, &tmA, mb * kBM + 64, kb * kBK, &full[s]);
                    }
                    if (!p.b_mn) {
                        tma_load_2d(sb, &tmB, kb * kBK, nb * BN, &full[s]);
                    } else {
                        for (int c = 0; c * 64 < BN; ++c) tma_load_2d(sb + c * 8192, &tmB, nb * BN + c * 64, kb * kBK, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a / b format (F16 0, BF16 1) at
            // [7,10) / [10,13), a / b major (0 K, 1 MN) at bits 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29)
            const unsigned fmt = p.dtype == 2 ? 1u : 0u;
            const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)p.a_mn << 15) | ((unsigned)p.b_mn << 16) |
                                   ((unsigned)(BN >> 3) << 17) | ((unsigned)(kBM >> 4) << 24);
            unsigned it = 0, tc = 0;
            for (int t = blockIdx.x; t < n_work; t += gridDim.x, ++tc) {
                const int sp = t % p.split_k;
                const int kb0 = sp * kb_per, kb1 = min(kb_total, kb0 + kb_per);
                const int as = tc & 1;
                mbar_wait_prof(&tempty[as], ((tc >> 1) & 1) ^ 1, p.prof ? p.prof + 1 : nullptr);
                tc_fence_after();
                const unsigned d_tmem = tmem_base + (unsigned)(as * BN);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % p.stages;
                    mbar_wait_prof(&full[s], (it / p.stages) & 1, p.prof ? p.prof + 2 : nullptr);
                    tc_fence_after();
                    const unsigned sa = smem_u32(stage0 + (size_t)s * stage_bytes), sb = sa + a_tile;
                    const int k_left = p.K - kb * kBK;
                    const int nk = k_left >= kBK ? kBK / kUmmaK : (k_left + kUmmaK - 1) / kUmmaK;
                    for (int k = 0; k < nk; ++k) {
                        const uint64_t da = p.a_mn ? make_desc(sa + k * (kUmmaK * 128), 8192, 1024) : make_desc(sa + k * (kUmmaK * 2), 16, 1024);
                        const uint64_t db = p.b_mn ? make_desc(sb + k * (kUmmaK * 128), 8192, 1024) : make_desc(sb + k * (kUmmaK * 2), 16, 1024);
                        tc_mma(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[s]);                 // the stage may be refilled once these MMAs have read it
                }
                tc_commit(&tfull[as]);                    // accumulator complete
            }
        }
    } else {
        // ------------------------------- epilogue warps (2..5): TMEM lane quadrant = warp % 4 -------------------------------
        // A thread drains one accumulator row (its TMEM lane), 32 columns per tcgen05.ld with the next load in flight, and parks
        // the converted values in a warp-private staging tile of 32 rows x 128 bytes; the warp then writes the tile out with 8
        // lanes per row, i.e. one full 128-byte line per row and four lines per store instruction.  (Storing each thread's own row
        // straight from registers costs one LSU transaction per 16 bytes: measured 3.2 us per 128 x 256 tile, four times the
        // HBM time of the tile.)  No CTA-level barrier: the tile is private to the warp.
        const int q = warp & 3;                          // TMEM lanes 32 q .. 32 q + 31 (a warp may only touch its own quadrant)
        const int half = (warp - 2) >> 2;                // two warps share a quadrant: each drains half of the column groups
        unsigned tc = 0;
        TOut *const D = reinterpret_cast<TOut *>(p.D);
        constexpr int kPitch = 144;                      // staging row pitch: 128 + 16 bytes -> conflict-free 16-byte accesses
        constexpr int kGC = 128 / (int)sizeof(TOut);     // columns per staged group: 64 (16-bit) or 32 (fp32)
        constexpr int kVec = 16 / (int)sizeof(TOut);
        uint8_t *stg = staging + (size_t)(warp - 2) * (32 * kPitch);
        for (int t = blockIdx.x; t < n_work; t += gridDim.x, ++tc) {
            const int nb = (t / p.split_k) % n_tiles, mb = t / (p.split_k * n_tiles);
            const int sp = t % p.split_k;
            const int as = tc & 1;
            mbar_wait_prof(&tfull[as], (tc >> 1) & 1, (p.prof && lane == 0 && half == 0) ? p.prof + 3 + q : nullptr);
            tc_fence_after();
            const long long t_tile0 = p.prof ? clock64() : 0;
            const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(as * BN);
            const int n0 = nb * BN;
            const int m0 = mb * kBM + q * 32;             // first row of this warp
            const int n_valid = min(BN, p.N - n0);
            const float bias_m = (p.epilogue == GEMM_EPI_BIAS_M && p.bias && m0 + lane < p.M && sp == 0) ? p.bias[m0 + lane] : 0.f;
            // column groups of this warp: [g_begin, g_end), the first / second half of the tile's groups
            const int n_groups = (n_valid + kGC - 1) / kGC;
            const int g_begin = half ? ((n_groups + 1) / 2) * kGC : 0;
            const int g_end = half ? n_valid : min(n_valid, ((n_groups + 1) / 2) * kGC);
            if (m0 < p.M && g_begin < g_end) {            // whole-warp condition: tcgen05.ld is .sync.aligned
                unsigned cur[32];
                tc_ld32(taddr + (unsigned)g_begin, cur);
                for (int g0 = g_begin; g0 < g_end; g0 += kGC) {
#pragma unroll
                    for (int h = 0; h < kGC / 32; ++h) {
                        const int c0 = g0 + 32 * h;
                        if (c0 < g_end) {
                            const long long tq0 = p.prof ? clock64() : 0;
                            tc_wait_ld();
                            const long long tq1 = p.prof ? clock64() : 0;
                            if (p.prof && warp == 2 && lane == 0) atomicAdd(p.prof + 8, (unsigned long long)(tq1 - tq0));
                            float v[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(cur[i]);
                            // the next 32 columns load into `cur` while `v` is converted and staged (v is live across the load, so
                            // the two never share registers; `cur` is not read again before the next wait)
                            if (c0 + 32 < g_end) tc_ld32(taddr + (unsigned)(c0 + 32), cur);
                            if (p.epilogue == GEMM_EPI_BIAS_N || p.epilogue == GEMM_EPI_BIAS_N_GELU) {
                                if (p.bias && sp == 0) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) v[i] += (n0 + c0 + i < p.N) ? __ldg(p.bias + n0 + c0 + i) : 0.f;
                                }
                                if (p.epilogue == GEMM_EPI_BIAS_N_GELU) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                                }
                            } else if (p.epilogue == GEMM_EPI_BIAS_M) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] += bias_m;
                            }
                            uint8_t *srow = stg + lane * kPitch + h * 64;
                            if constexpr (sizeof(TOut) == 4) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) *reinterpret_cast<F4 *>(srow + 16 * i) = F4{v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    *reinterpret_cast<uint4 *>(srow + 16 * i) =
                                        make_uint4(pack2<TOut>(v[8 * i], v[8 * i + 1]), pack2<TOut>(v[8 * i + 2], v[8 * i + 3]),
                                                   pack2<TOut>(v[8 * i + 4], v[8 * i + 5]), pack2<TOut>(v[8 * i + 6], v[8 * i + 7]));
                            }
                        }
                    }
                    __syncwarp();
                    const long long tq2 = p.prof ? clock64() : 0;
                    // ---- write the group out: lane -> (row = 4 it + lane / 8, 16-byte piece = lane % 8) ----
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = 4 * it + (lane >> 3), pc = lane & 7;
                        const int col = g0 + pc * kVec, m = m0 + r;
                        if (m < p.M && col < n_valid) {
                            const uint8_t *src = stg + r * kPitch + pc * 16;
                            TOut *dst = D + (int64_t)m * p.ldd + n0 + col;
                            const bool full = (col + kVec <= n_valid) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
                            if constexpr (sizeof(TOut) == 4) {
                                const F4 sv = *reinterpret_cast<const F4 *>(src);
                                float *fd = reinterpret_cast<float *>(dst);
                                if (full) {
                                    if (p.atomic) red_add_v4f(fd, sv.x, sv.y, sv.z, sv.w);
                                    else *reinterpret_cast<F4 *>(fd) = sv;
                                } else {
                                    const float e4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        if (col + i < n_valid) {
                                            if (p.atomic) atomicAdd(fd + i, e4[i]);
                                            else fd[i] = e4[i];
                                        }
                                    }
                                }
                            } else {
                                uint4 sv = *reinterpret_cast<const uint4 *>(src);
                                if (full) {
                                    if (p.atomic) {                   // D += product (16-bit outputs: read-modify-write, no split-K)
                                        const uint4 o = *reinterpret_cast<const uint4 *>(dst);
                                        sv = make_uint4(sum2<TOut>(o.x, sv.x), sum2<TOut>(o.y, sv.y), sum2<TOut>(o.z, sv.z), sum2<TOut>(o.w, sv.w));
                                    }
                                    *reinterpret_cast<uint4 *>(dst) = sv;
                                } else {
                                    const TOut *se = reinterpret_cast<const TOut *>(&sv);
#pragma unroll
                                    for (int i = 0; i < kVec; ++i) {
                                        if (col + i < n_valid) {
                                            const float o = p.atomic ? static_cast<float>(dst[i]) + static_cast<float>(se[i]) : static_cast<float>(se[i]);
                                            dst[i] = static_cast<TOut>(o);
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (p.prof && warp == 2 && lane == 0) atomicAdd(p.prof + 10, (unsigned long long)(clock64() - tq2));
                }
            }
            if (p.prof && warp == 2 && lane == 0) atomicAdd(p.prof + 11, (unsigned long long)(clock64() - t_tile0));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);      // accumulator buffer drained: the MMA warp may reuse it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.prof && threadIdx.x == 0) atomicAdd(p.prof + 7, (unsigned long long)(clock64() - t_start));
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

cudaError_t encode_2d(CUtensorMap *m, const void *base, int dtype, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !ptr) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeFn>(ptr);
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {outer_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                          dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) return cudaErrorDeviceUninitialized;
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace

int gemm_pick_bn(int N) {
    if (N >= 256) return 256;
    return ((N + 15) / 16) * 16;
}

cudaError_t gemm_tc_launch(GemmP p, const void *A, int64_t lda, const void *B, int64_t ldb, cudaStream_t st, const char **where) {
    const char *dummy;
    if (!where) where = &dummy;
    *where = "setup";
    static int sms = 0, max_smem = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    if (p.BN <= 0) p.BN = gemm_pick_bn(p.N);
    const int BN = p.BN;
    const size_t b_tile = p.b_mn ? (size_t)((BN + 63) / 64) * 8192 : (size_t)BN * 128;
    const size_t stage_bytes = (size_t)kBM * 128 + ((b_tile + 1023) & ~(size_t)1023);
    const size_t fixed = 1024 + kEpiWarps * 32 * 144 + (2 * kMaxStages + 4) * 8 + 16;
    int stages = (int)(((size_t)max_smem - fixed) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    const int kb_total = (p.K + kBK - 1) / kBK;
    if (stages < 2) return cudaErrorInvalidConfiguration;
    p.stages = stages;
    int cols = 32;
    while (cols < 2 * BN + ((BN & 31) ? 16 : 0)) cols <<= 1;   // the epilogue reads 32 columns at a time: keep its last read inside the allocation
    p.tmem_cols = cols;
    if (p.split_k < 1) p.split_k = 1;
    if (p.split_k > kb_total) p.split_k = kb_total;
    {   // no K slice may be empty: an accumulator that no MMA wrote would be added as is
        const int kb_per = (kb_total + p.split_k - 1) / p.split_k;
        p.split_k = (kb_total + kb_per - 1) / kb_per;
    }
    const size_t smem = fixed + (size_t)stages * stage_bytes;

    CUtensorMap tmA, tmB;
    cudaError_t e;
    {   // cuTensorMapEncodeTiled is a DRIVER call: it needs a current context in the calling thread.  PyTorch's autograd worker
        // threads select their device lazily, so a backward whose first CUDA work is this GEMM arrives here with no context
        // bound (seen on hardware as CUDA_ERROR_INVALID_CONTEXT).  Bind the primary context of the device that owns A.
        *where = "cudaPointerGetAttributes(A)";
        cudaPointerAttributes attr;
        if ((e = cudaPointerGetAttributes(&attr, A)) != cudaSuccess) return e;
        if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) return cudaErrorInvalidDevicePointer;
        if ((e = cudaSetDevice(attr.device)) != cudaSuccess) return e;
    }
    *where = "tensor map of A (cuTensorMapEncodeTiled)";
    // A: K-major -> dims (K, M), box (64, 128); MN-major -> dims (M, K), box (64, 64) loaded twice per stage
    e = p.a_mn ? encode_2d(&tmA, A, p.dtype, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda * 2, 64, 64)
               : encode_2d(&tmA, A, p.dtype, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, 64, kBM);
    if (e != cudaSuccess) return e;
    *where = "tensor map of B (cuTensorMapEncodeTiled)";
    e = p.b_mn ? encode_2d(&tmB, B, p.dtype, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb * 2, 64, 64)
               : encode_2d(&tmB, B, p.dtype, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldb * 2, 64, (uint32_t)BN);
    if (e != cudaSuccess) return e;

    const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + BN - 1) / BN;
    const long n_work = (long)m_tiles * n_tiles * p.split_k;
    const int ctas_per_sm = smem * 2 + 2048 <= (size_t)228 * 1024 && cols <= 256 ? 2 : 1;
    const int grid = (int)(n_work < (long)sms * ctas_per_sm ? n_work : (long)sms * ctas_per_sm);
#define SMB_GEMM_LAUNCH(T)                                                                                                  \
    do {                                                                                                                    \
        static size_t set = 0;                                                                                              \
        *where = "cudaFuncSetAttribute";                                                                                    \
        if (smem > set) {                                                                                                   \
            e = cudaFuncSetAttribute(gemm_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);        \
            if (e != cudaSuccess) return e;                                                                                 \
            set = (size_t)max_smem;                                                                                         \
        }                                                                                                                   \
        *where = "kernel launch";                                                                                           \
        gemm_tc_kernel<T><<<grid, kThreads, smem, st>>>(tmA, tmB, p);                                                       \
    } while (0)
    if (p.out_dtype == 0) SMB_GEMM_LAUNCH(float);
    else if (p.out_dtype == 1) SMB_GEMM_LAUNCH(__half);
    else SMB_GEMM_LAUNCH(__nv_bfloat16);
#undef SMB_GEMM_LAUNCH
    count_launch();
    return cudaGetLastError();
}

}  // namespace smb
