// cp.async (LDGSTS) wrappers: global -> shared copies that need no registers while in flight, so a warp can compute on one
// tile while the next one is on its way.  Under SMB_EMU (tools/simt_emu) the copy is deferred until the issuing thread waits
// for its group, which is exactly the contract the hardware gives.
#pragma once

#include "common.cuh"

namespace smb {

#ifdef SMB_EMU
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) { emu::cp_async_issue(smem_dst, gmem_src, 8); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) { emu::cp_async_issue(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_commit() { emu::cp_async_commit(); }
__device__ __forceinline__ void cp_async_wait_all() { emu::cp_async_wait(-1); }
#else
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

}  // namespace smb
