// 1-D TMA bulk copies (cp.async.bulk) + mbarrier helpers for the streaming (HBM-bound) kernels of
// loss.cu and detect.cu: a contiguous tile of anchor rows is pulled into shared memory by ONE
// instruction (no register staging, full-line DRAM requests) and written back the same way.
// Requirements of the instruction: 16-byte aligned global and shared addresses, size % 16 == 0.
#pragma once
#include <cstdint>
#include <cstdio>

namespace ssdb {
namespace bulk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) { printf("ssdb bulk: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
    }
}
// global -> shared, completion signalled on the mbarrier (transaction bytes)
__device__ __forceinline__ void load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global (bulk async-group)
__device__ __forceinline__ void store_1d(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tile helpers: a CTA of NT threads moves a contiguous run of `nfl` floats ----
// pull one or two tiles into shared memory: one TMA bulk copy per tensor (issued by thread 0, completion on `bar`),
// or plain coalesced loads when the buffers do not meet the 16-byte rules
template <int NT>
__device__ __forceinline__ void tile_load(float* dst0, const float* src0, float* dst1, const float* src1, int nfl, int use_bulk,
                                          unsigned long long* bar) {
    const int tid = threadIdx.x;
    if (use_bulk) {
        const uint32_t bar_a = smem_u32(bar);
        if (tid == 0) mbar_init(bar_a, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)nfl * 4u;
            mbar_expect_tx(bar_a, src1 ? 2u * bytes : bytes);
            load_1d(smem_u32(dst0), src0, bytes, bar_a);
            if (src1) load_1d(smem_u32(dst1), src1, bytes, bar_a);
        }
    } else {
        for (int i = tid; i < nfl; i += NT) { dst0[i] = src0[i]; if (src1) dst1[i] = src1[i]; }
    }
}
__device__ __forceinline__ void tile_load_wait(int use_bulk, unsigned long long* bar) {
    if (use_bulk) mbar_wait(smem_u32(bar), 0);
    __syncthreads();
}
// write a finished shared-memory tile to global (every thread calls; contains a CTA barrier)
template <int NT>
__device__ __forceinline__ void tile_store(float* dst, const float* src_smem, int nfl, int use_bulk) {
    if (use_bulk) {
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            store_1d(dst, smem_u32(src_smem), (uint32_t)nfl * 4u);
            store_commit();
            store_wait_all();
        }
    } else {
        __syncthreads();
        for (int i = threadIdx.x; i < nfl; i += NT) dst[i] = src_smem[i];
    }
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace bulk
}  // namespace ssdb
