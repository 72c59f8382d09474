// Forward selective scan, software-pipelined variant for 16-bit activations (opt-in: SMB_FWD_V2=1).
//
// Same three passes, mapping and arithmetic as scan_fwd.cu (lane == channel, all N states in registers, B/C broadcast from a
// warp-private tile).  What changes is how a tile reaches the SM.  In scan_fwd.cu every warp alternates between a fill phase
// (global loads -> convert -> shared) and a compute phase (16 ex2 per position); the warps of a CTA run in lockstep, so the
// MUFU pipe idles while they all fill (ncu: XU 64 % / 46 % busy in the two passes, top stall long_scoreboard).  Here the
// NEXT tile is copied global -> shared with cp.async while the current one is computed:
//   * u / delta / z tiles stay in their 16-bit storage type in two alternating stages (2 KB each: 32 rows x 64 bytes, the
//     16-byte units of a row XOR-swizzled with (row >> 1) & 3 so that the per-lane LDS.128 of the compute phase and the
//     8-byte cp.async / store accesses are both conflict-free); conversion, softplus and SiLU happen in registers when a lane
//     reads its 8 positions, and y / out_z are written back in place over the consumed u / z units for the coalesced store;
//   * B / C rows land in a raw staging buffer and are transposed into the position-major fp32 broadcast tile at the top of
//     the iteration, before the staging buffer is handed to the next cp.async.
// Shared memory per warp: 11 KB (pass 1) / 18 KB (pass 3), against 10 / 16 KB in scan_fwd.cu.  Tiles that are ragged or not
// 8-byte aligned take a synchronous fill into the same stage layout, so there is one compute path.
#include "raw_tiles.cuh"
#include "scan_internal.h"
#include "scan_steps.cuh"

namespace smb {

// ---------------------------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------------------------
template <int N> struct AggSmem { static constexpr int kWarpBytes = 4 * kRawTileBytes + kTile * N * 4 + N * kTile * 2; };

template <typename T, int N, bool kRev>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) scan_fwd_agg2_kernel(const ScanP p) {
    static_assert(sizeof(T) == 2, "the pipelined variant is for 16-bit activations");
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerCta + warp;
    if (w >= p.n_work) return;
    const WorkItem wi = decode_work(p, w);
    const int d = wi.d0 + lane;
    const bool active = lane < wi.nrows;

    unsigned char *wb = reinterpret_cast<unsigned char *>(smem) + (size_t)warp * AggSmem<N>::kWarpBytes;
    constexpr int kStageBytes = 2 * kRawTileBytes;           // stage s: u at wb + s * kStageBytes, delta one tile further
    float *s_B = reinterpret_cast<float *>(wb + 4 * kRawTileBytes);
    T *rawB = reinterpret_cast<T *>(wb + 4 * kRawTileBytes + kTile * N * 4);

    float2 A2[N / 2], h[N / 2];
#pragma unroll
    for (int m = 0; m < N / 2; ++m) {
        A2[m] = active ? f2(p.A[(int64_t)d * N + 2 * m] * kLog2e, p.A[(int64_t)d * N + 2 * m + 1] * kLog2e) : f2(0.f, 0.f);
        h[m] = f2(0.f, 0.f);
    }
    const float bias = (active && p.delta_bias) ? p.delta_bias[d] : 0.f;
    float sumdt = 0.f;

    const T *u = reinterpret_cast<const T *>(p.u) + wi.b * p.u_bs + (int64_t)wi.d0 * p.u_ds;
    const T *dl = reinterpret_cast<const T *>(p.delta) + wi.b * p.delta_bs + (int64_t)wi.d0 * p.delta_ds;
    const T *Bm = reinterpret_cast<const T *>(p.B) + wi.b * p.B_bs + (int64_t)wi.g * p.B_gs;

    const int j_begin = wi.seg * p.S;
    const int j_end = min(p.L, j_begin + p.S);
    const bool fast = stream_aligned(u, p.u_ds, p.L, kRev) && stream_aligned(dl, p.delta_ds, p.L, kRev) && bc_aligned(Bm, p.B_ns, p.L, kRev);
    const LanePtr<T> lpu = lane_ptr(u, p.u_ds, p.L, kRev, lane), lpd = lane_ptr(dl, p.delta_ds, p.L, kRev, lane);

    int stage = 0;
    bool pending = false;
    if (j_begin < j_end && fast && j_begin + kTile <= p.L) {
        issue_tile<T>(wb, lpu, wi.nrows, j_begin, kRev, lane);
        issue_tile<T>(wb + kRawTileBytes, lpd, wi.nrows, j_begin, kRev, lane);
        issue_bc<T, N>(rawB, Bm, p.B_ns, j_begin, p.L, kRev, lane);
        cp_async_commit();
        pending = true;
    }
    for (int j0 = j_begin; j0 < j_end; j0 += kTile) {
        unsigned char *t_u = wb + stage * kStageBytes, *t_d = t_u + kRawTileBytes;
        unsigned char *n_u = wb + (stage ^ 1) * kStageBytes, *n_d = n_u + kRawTileBytes;
        if (pending) {
            cp_async_wait_all();
            __syncwarp();
            convert_bc<T, N, kRev>(s_B, rawB, lane);
        } else {
            fill_raw_sync<T, kRev>(t_u, u, p.u_ds, wi.nrows, j0, p.L, lane);
            fill_raw_sync<T, kRev>(t_d, dl, p.delta_ds, wi.nrows, j0, p.L, lane);
            float *const bt[1] = {s_B};
            const T *const bb[1] = {Bm};
            const int64_t bns[1] = {p.B_ns}, bls[1] = {p.B_ls};
            fill_bc_tiles<T, N, 1>(bt, bb, bns, bls, j0, p.L, kRev, lane);
        }
        __syncwarp();
        const int jn = j0 + kTile;
        const bool next_async = jn < j_end && fast && jn + kTile <= p.L;
        if (next_async) {                                   // in flight during the compute below
            issue_tile<T>(n_u, lpu, wi.nrows, jn, kRev, lane);
            issue_tile<T>(n_d, lpd, wi.nrows, jn, kRev, lane);
            issue_bc<T, N>(rawB, Bm, p.B_ns, jn, p.L, kRev, lane);
            cp_async_commit();
        }
        const int nvalid = j_end - j0;
#pragma unroll 1
        for (int u8 = 0; u8 < kTile / 8; ++u8) {
            float uu[8], dd[8];
            read_unit<T, kRev>(t_u, lane, u8, uu);
            read_unit<T, kRev>(t_d, lane, u8, dd);
            unit_dt(dd, bias, p.softplus, 8 * u8, nvalid);
            agg_block<N, 0>(s_B + 8 * u8 * N, uu, dd, A2, h, sumdt);
        }
        __syncwarp();                                       // s_B and this stage are free for the next iteration's writers
        pending = next_async;
        stage ^= 1;
    }
    if (active) {
        const int64_t o = (((int64_t)wi.b * p.n_seg + wi.seg) * N) * p.dim + d;
#pragma unroll
        for (int m = 0; m < N / 2; ++m) {
            p.P[o + (int64_t)(2 * m) * p.dim] = ex2(A2[m].x * sumdt);
            p.P[o + (int64_t)(2 * m + 1) * p.dim] = ex2(A2[m].y * sumdt);
            p.H[o + (int64_t)(2 * m) * p.dim] = h[m].x;
            p.H[o + (int64_t)(2 * m + 1) * p.dim] = h[m].y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 3
// ---------------------------------------------------------------------------------------------
template <int N> struct MainSmem { static constexpr int kWarpBytes = 6 * kRawTileBytes + 2 * kTile * N * 4 + 2 * N * kTile * 2; };

template <typename T, int N, bool kHasZ, bool kRev>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3) scan_fwd_main2_kernel(const ScanP p) {
    static_assert(sizeof(T) == 2, "the pipelined variant is for 16-bit activations");
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerCta + warp;
    if (w >= p.n_work) return;
    const WorkItem wi = decode_work(p, w);
    const int d = wi.d0 + lane;
    const bool active = lane < wi.nrows;

    unsigned char *wb = reinterpret_cast<unsigned char *>(smem) + (size_t)warp * MainSmem<N>::kWarpBytes;
    constexpr int kStageBytes = 3 * kRawTileBytes;           // stage s: u, delta, z tiles at wb + s * kStageBytes
    float *s_B = reinterpret_cast<float *>(wb + 6 * kRawTileBytes);
    float *s_C = s_B + kTile * N;
    T *rawB = reinterpret_cast<T *>(wb + 6 * kRawTileBytes + 2 * kTile * N * 4);
    T *rawC = rawB + N * kTile;

    float2 A2[N / 2], h[N / 2];
    {
        const int64_t o = (((int64_t)wi.b * p.n_seg + wi.seg) * N) * p.dim + d;
#pragma unroll
        for (int m = 0; m < N / 2; ++m) {
            A2[m] = active ? f2(p.A[(int64_t)d * N + 2 * m] * kLog2e, p.A[(int64_t)d * N + 2 * m + 1] * kLog2e) : f2(0.f, 0.f);
            h[m] = active ? f2(p.hin[o + (int64_t)(2 * m) * p.dim], p.hin[o + (int64_t)(2 * m + 1) * p.dim]) : f2(0.f, 0.f);
        }
    }
    const float bias = (active && p.delta_bias) ? p.delta_bias[d] : 0.f;
    const float Dv = (active && p.D) ? p.D[d] : 0.f;

    const T *u = reinterpret_cast<const T *>(p.u) + wi.b * p.u_bs + (int64_t)wi.d0 * p.u_ds;
    const T *dl = reinterpret_cast<const T *>(p.delta) + wi.b * p.delta_bs + (int64_t)wi.d0 * p.delta_ds;
    const T *z = kHasZ ? reinterpret_cast<const T *>(p.z) + wi.b * p.z_bs + (int64_t)wi.d0 * p.z_ds : nullptr;
    const T *Bm = reinterpret_cast<const T *>(p.B) + wi.b * p.B_bs + (int64_t)wi.g * p.B_gs;
    const T *Cm = reinterpret_cast<const T *>(p.C) + wi.b * p.C_bs + (int64_t)wi.g * p.C_gs;
    T *out = p.out ? reinterpret_cast<T *>(p.out) + wi.b * p.out_bs + (int64_t)wi.d0 * p.out_ds : nullptr;
    T *out_z = kHasZ ? reinterpret_cast<T *>(p.out_z) + wi.b * p.out_z_bs + (int64_t)wi.d0 * p.out_z_ds : nullptr;

    const int j_begin = wi.seg * p.S;
    const int j_end = min(p.L, j_begin + p.S);
    const bool fast = stream_aligned(u, p.u_ds, p.L, kRev) && stream_aligned(dl, p.delta_ds, p.L, kRev) &&
                      (!kHasZ || stream_aligned(z, p.z_ds, p.L, kRev)) && bc_aligned(Bm, p.B_ns, p.L, kRev) &&
                      bc_aligned(Cm, p.C_ns, p.L, kRev);
    const LanePtr<T> lpu = lane_ptr(u, p.u_ds, p.L, kRev, lane), lpd = lane_ptr(dl, p.delta_ds, p.L, kRev, lane);
    const LanePtr<T> lpzi = lane_ptr(kHasZ ? z : u, kHasZ ? p.z_ds : p.u_ds, p.L, kRev, lane);
    const bool fast_out = out ? stream_aligned(out, p.out_ds, p.L, kRev) : true;
    const bool fast_oz = kHasZ ? stream_aligned(out_z, p.out_z_ds, p.L, kRev) : true;
    const LanePtr<T> lpo = lane_ptr(out ? (const T *)out : u, out ? p.out_ds : p.u_ds, p.L, kRev, lane);
    const LanePtr<T> lpz = lane_ptr(kHasZ ? (const T *)out_z : u, kHasZ ? p.out_z_ds : p.u_ds, p.L, kRev, lane);

    int stage = 0;
    bool pending = false;
    if (j_begin < j_end && fast && j_begin + kTile <= p.L) {
        issue_tile<T>(wb, lpu, wi.nrows, j_begin, kRev, lane);
        issue_tile<T>(wb + kRawTileBytes, lpd, wi.nrows, j_begin, kRev, lane);
        if (kHasZ) issue_tile<T>(wb + 2 * kRawTileBytes, lpzi, wi.nrows, j_begin, kRev, lane);
        issue_bc<T, N>(rawB, Bm, p.B_ns, j_begin, p.L, kRev, lane);
        issue_bc<T, N>(rawC, Cm, p.C_ns, j_begin, p.L, kRev, lane);
        cp_async_commit();
        pending = true;
    }
    for (int j0 = j_begin; j0 < j_end; j0 += kTile) {
        unsigned char *t_u = wb + stage * kStageBytes, *t_d = t_u + kRawTileBytes, *t_z = t_u + 2 * kRawTileBytes;
        unsigned char *n_u = wb + (stage ^ 1) * kStageBytes, *n_d = n_u + kRawTileBytes, *n_z = n_u + 2 * kRawTileBytes;
        if (p.hstates && (j0 % kCkpt) == 0 && active) {
            const int64_t o = (((int64_t)wi.b * (p.nck + 1) + j0 / kCkpt) * N) * p.dim + d;
#pragma unroll
            for (int m = 0; m < N / 2; ++m) {
                p.hstates[o + (int64_t)(2 * m) * p.dim] = h[m].x;
                p.hstates[o + (int64_t)(2 * m + 1) * p.dim] = h[m].y;
            }
        }
        if (pending) {
            cp_async_wait_all();
            __syncwarp();
            convert_bc<T, N, kRev>(s_B, rawB, lane);
            convert_bc<T, N, kRev>(s_C, rawC, lane);
        } else {
            fill_raw_sync<T, kRev>(t_u, u, p.u_ds, wi.nrows, j0, p.L, lane);
            fill_raw_sync<T, kRev>(t_d, dl, p.delta_ds, wi.nrows, j0, p.L, lane);
            if (kHasZ) fill_raw_sync<T, kRev>(t_z, z, p.z_ds, wi.nrows, j0, p.L, lane);
            float *const bt[2] = {s_B, s_C};
            const T *const bb[2] = {Bm, Cm};
            const int64_t bns[2] = {p.B_ns, p.C_ns}, bls[2] = {p.B_ls, p.C_ls};
            fill_bc_tiles<T, N, 2>(bt, bb, bns, bls, j0, p.L, kRev, lane);
        }
        __syncwarp();
        const int jn = j0 + kTile;
        const bool next_async = jn < j_end && fast && jn + kTile <= p.L;
        if (next_async) {
            issue_tile<T>(n_u, lpu, wi.nrows, jn, kRev, lane);
            issue_tile<T>(n_d, lpd, wi.nrows, jn, kRev, lane);
            if (kHasZ) issue_tile<T>(n_z, lpzi, wi.nrows, jn, kRev, lane);
            issue_bc<T, N>(rawB, Bm, p.B_ns, jn, p.L, kRev, lane);
            issue_bc<T, N>(rawC, Cm, p.C_ns, jn, p.L, kRev, lane);
            cp_async_commit();
        }
        const int nvalid = j_end - j0;
#pragma unroll 1
        for (int u8 = 0; u8 < kTile / 8; ++u8) {
            float uu[8], dd[8], yy[8];
            read_unit<T, kRev>(t_u, lane, u8, uu);
            read_unit<T, kRev>(t_d, lane, u8, dd);
            unit_dt(dd, bias, p.softplus, 8 * u8, nvalid);
            main_block<N, 0>(s_B + 8 * u8 * N, s_C + 8 * u8 * N, uu, dd, Dv, A2, h, yy);
            if (kHasZ) {
                float zz[8];
                read_unit<T, kRev>(t_z, lane, u8, zz);
#pragma unroll
                for (int e = 0; e < 8; ++e) zz[e] = yy[e] * zz[e] * sigmoidf(zz[e]);
                write_unit<T, kRev>(t_z, lane, u8, zz);          // in place: this lane consumed the unit
            }
            if (out) write_unit<T, kRev>(t_u, lane, u8, yy);
        }
        __syncwarp();
        const bool full = j0 + kTile <= p.L;
        if (out) {
            if (full && fast_out) store_raw_fast<T>(t_u, const_cast<T *>(lpo.lp), lpo.rowstep, wi.nrows, j0, kRev, lane);
            else store_raw_sync<T, kRev>(t_u, out, p.out_ds, wi.nrows, j0, p.L, lane);
        }
        if (kHasZ) {
            if (full && fast_oz) store_raw_fast<T>(t_z, const_cast<T *>(lpz.lp), lpz.rowstep, wi.nrows, j0, kRev, lane);
            else store_raw_sync<T, kRev>(t_z, out_z, p.out_z_ds, wi.nrows, j0, p.L, lane);
        }
        __syncwarp();
        pending = next_async;
        stage ^= 1;
    }
    if (p.hstates && j_end == p.L && active) {
        const int64_t o = (((int64_t)wi.b * (p.nck + 1) + p.nck) * N) * p.dim + d;
#pragma unroll
        for (int m = 0; m < N / 2; ++m) {
            p.hstates[o + (int64_t)(2 * m) * p.dim] = h[m].x;
            p.hstates[o + (int64_t)(2 * m + 1) * p.dim] = h[m].y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side (same orchestration as launch_fwd in scan_fwd.cu)
// ---------------------------------------------------------------------------------------------
template <typename T, int N, bool kRev>
static cudaError_t launch_fwd2(const ScanP &p, bool has_z, float *x, cudaStream_t st) {
    const int ctas = (p.n_work + kWarpsPerCta - 1) / kWarpsPerCta;
    const size_t sm1 = (size_t)kWarpsPerCta * AggSmem<N>::kWarpBytes;
    const size_t sm3 = (size_t)kWarpsPerCta * MainSmem<N>::kWarpBytes;
    cudaError_t e;
    SMB_SET_SMEM_ONCE((scan_fwd_agg2_kernel<T, N, kRev>), sm1);
    if (p.n_seg > 1) {
        scan_fwd_agg2_kernel<T, N, kRev><<<ctas, kWarpsPerCta * 32, sm1, st>>>(p); count_launch();
        if ((e = carry_launch(p.P, p.H, p.hin, x ? p.cumP : nullptr, p.batch, p.n_seg, N, p.dim, 0, st)) != cudaSuccess) return e;
    } else {
        if ((e = cudaMemsetAsync(p.hin, 0, sizeof(float) * (size_t)p.batch * N * p.dim, st)) != cudaSuccess) return e;
        if (x) {
            scan_fwd_agg2_kernel<T, N, kRev><<<ctas, kWarpsPerCta * 32, sm1, st>>>(p); count_launch();
            if ((e = cudaMemcpyAsync(p.cumP, p.P, sizeof(float) * (size_t)p.batch * N * p.dim, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        }
    }
    if (has_z) {
        SMB_SET_SMEM_ONCE((scan_fwd_main2_kernel<T, N, true, kRev>), sm3);
        scan_fwd_main2_kernel<T, N, true, kRev><<<ctas, kWarpsPerCta * 32, sm3, st>>>(p); count_launch();
    } else {
        SMB_SET_SMEM_ONCE((scan_fwd_main2_kernel<T, N, false, kRev>), sm3);
        scan_fwd_main2_kernel<T, N, false, kRev><<<ctas, kWarpsPerCta * 32, sm3, st>>>(p); count_launch();
    }
    if (x) return x_finalize_launch(p, N, x, st);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_fwd2_t(const ScanP &p, int N, bool has_z, float *x, cudaStream_t st) {
    if (N == 16) return p.reverse ? launch_fwd2<T, 16, true>(p, has_z, x, st) : launch_fwd2<T, 16, false>(p, has_z, x, st);
    return p.reverse ? launch_fwd2<T, 8, true>(p, has_z, x, st) : launch_fwd2<T, 8, false>(p, has_z, x, st);
}

cudaError_t scan_fwd_v2_dispatch(const ScanP &p, int dtype, int N, bool has_z, float *x, cudaStream_t st) {
    if (dtype == 1) return launch_fwd2_t<__half>(p, N, has_z, x, st);
    return launch_fwd2_t<__nv_bfloat16>(p, N, has_z, x, st);
}

}  // namespace smb
