// ptx.cuh -- inline-PTX helpers: mbarrier + 1-D bulk async copies (the TMA unit's
// non-tensor path, SASS: UBLKCP) used to stage D x D tiles into shared memory.
#pragma once
#include <cstdint>

namespace vgpa {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async proxy
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// pull a contiguous global range into L2 ahead of its use (no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes)
{
#ifndef VGPA_NO_L2_PREFETCH
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
#endif
}
// ---- cp.async (SASS LDGSTS): 16-byte global -> shared copies without register staging ----
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// a 40 x 40 row-major matrix into a padded shared tile (row pitch `pitch` doubles), split over `nth` threads
__device__ __forceinline__ void cp_async_matrix(double* dst, const double* src, int pitch, int tid, int nth)
{
    for (int c = tid; c < 40 * 20; c += nth) {
        const int i = c / 20, k = c - i * 20;
        cp_async16(dst + i * pitch + 2 * k, src + i * 40 + 2 * k);
    }
}
// a 40-vector (20 chunks)
__device__ __forceinline__ void cp_async_vector(double* dst, const double* src, int tid, int first)
{
    const int c = tid - first;
    if (c >= 0 && c < 20) cp_async16(dst + 2 * c, src + 2 * c);
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order generic-proxy shared-memory writes before async-proxy reads of them
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace vgpa
