// Shared device helpers for the sm_100a kernels of the Lovasz / confusion-matrix path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtda_lovasz.h"

typedef unsigned int u32;
typedef unsigned long long u64;

#define RTDA_FULL_MASK 0xffffffffu

#define RTDA_RETURN_IF_CUDA_ERROR(expr)            \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define RTDA_CHECK_LAUNCH() RTDA_RETURN_IF_CUDA_ERROR(cudaGetLastError())

namespace rtda {

// Optional per-launch timing (profile.cu).  A ProfileScope brackets one launch with two
// events on the launch stream when profiling is enabled; otherwise it does nothing.
enum Stage {
    kStageMemset = 0, kStagePrep, kStageScan, kStagePass0, kStagePass1, kStagePass2, kStageRankLast,
    kStageFinalize, kStageBwd, kStageConfusion, kStageConfusionArgmax, kStageOther,
    kStageSplit, kStagePart, kStageLocal, kStageHistFb, kStagePlan
};
bool profile_enabled();
void profile_begin(int stage, cudaStream_t st);
void profile_end(int stage, cudaStream_t st);
struct ProfileScope {
    int stage; cudaStream_t st; bool on;
    ProfileScope(int s, cudaStream_t t) : stage(s), st(t), on(profile_enabled()) { if (on) profile_begin(stage, st); }
    ~ProfileScope() { if (on) profile_end(stage, st); }
};

__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Relaxed, gpu-scope accesses for the decoupled look-back descriptors: one word
// carries flag + value, so no ordering with other data is needed, only coherence
// at L2 (never a stale L1 line).
__device__ __forceinline__ u32 ld_relaxed(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u32 *p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Streaming (read-once) loads that do not allocate in L1.
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ u64 ld_stream_u64(const u64 *p) {
    u64 r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// Exclusive scan of one value per thread over a 256-thread block.
// warp_tmp: shared scratch of 8 elements.  Ends with a barrier so warp_tmp can be reused.
template <typename T>
__device__ __forceinline__ T block_excl_scan_256(T v, T *warp_tmp, T *total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(RTDA_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) warp_tmp[warp] = incl;
    __syncthreads();
    T woff = 0, tot = 0;
#pragma unroll
    for (u32 w = 0; w < 8; ++w) {
        T x = warp_tmp[w];
        if (w < warp) woff += x;
        tot += x;
    }
    __syncthreads();
    if (total) *total = tot;
    return woff + incl - v;
}

// mbarrier + bulk asynchronous copy (cp.async.bulk, the 1-D TMA path): one thread starts the copy,
// the data arrives in shared memory without passing through registers, every reader waits on the barrier.
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// Label element readers: value as int64 whatever the storage type.
template <typename LabelT>
__device__ __forceinline__ long long load_label(const LabelT *p, size_t i) {
    return (long long)p[i];
}

inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace rtda
