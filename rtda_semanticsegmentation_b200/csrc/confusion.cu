// Confusion-matrix histogram for sm_100a (B200).
//
// Replaces the per-batch slice of the reference's validation loop
// (validation.py:98-109: two D2H copies, ignore filter, utils.py:62-89 fast_hist,
// `hist +=`) with one pass over device-resident label / prediction maps, and
// optionally fuses validation.py:95 (argmax over the logits).
//
// HBM-bound integer work (16 B/pixel for int64 maps, 2 888 B out).  Real label
// maps are spatially coherent and mostly predicted right, so nearly all
// increments land on the n_class diagonal bins: each warp first groups equal
// bins with match.any and one lane adds the group size to a WARP-PRIVATE
// shared-memory histogram (no atomics, no same-address serialisation); blocks
// flush non-zero bins to the int64 matrix with one atomicAdd each.
#include "common.cuh"

namespace rtda {

constexpr int kCmThreads = 1024;  // one block per SM, 32 warps: the kernel is latency-bound at 2 Mpx per call (profiles/)
constexpr int kCmWarps = kCmThreads / 32;
constexpr int kCmUnroll = 8;

// mode 0: one private histogram per warp; mode 1: one per block (shared atomics);
// mode 2: no shared histogram (global atomics), for very large n_class.
template <int MODE>
__device__ __forceinline__ void cm_add(u32 *sh, int nbins, u32 bin, bool ok, unsigned long long *hist) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 m = __match_any_sync(RTDA_FULL_MASK, ok ? bin : 0xffffffffu);
    if (ok && lane == (u32)(__ffs(m) - 1)) {
        const u32 cnt = __popc(m);
        if (MODE == 0) sh[warp * nbins + bin] += cnt;
        else if (MODE == 1) atomicAdd(&sh[bin], cnt);
        else atomicAdd(&hist[bin], (unsigned long long)cnt);
    }
    if (MODE == 0) __syncwarp();
}

template <int MODE>
__device__ __forceinline__ void cm_flush(u32 *sh, int nbins, unsigned long long *hist) {
    if (MODE == 2) return;
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += kCmThreads) {
        u32 s = 0;
        if (MODE == 0) {
#pragma unroll
            for (int w = 0; w < kCmWarps; ++w) s += sh[w * nbins + b];
        } else {
            s = sh[b];
        }
        if (s) atomicAdd(&hist[b], (unsigned long long)s);
    }
}

template <typename LabelT, typename PredT, int MODE>
__global__ void __launch_bounds__(kCmThreads)
k_confusion(const LabelT *__restrict__ labels, const PredT *__restrict__ preds, long long n, int n_class,
            long long ignore, int has_ignore, unsigned long long *__restrict__ hist) {
    extern __shared__ u32 sh[];
    const int nbins = n_class * n_class;
    if (MODE != 2) {
        const int words = (MODE == 0 ? kCmWarps : 1) * nbins;
        for (int i = threadIdx.x; i < words; i += kCmThreads) sh[i] = 0;
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * kCmThreads * kCmUnroll;
    // every lane of a warp runs the same number of iterations (match.any needs the full warp)
    for (long long base = (long long)blockIdx.x * kCmThreads * kCmUnroll; base < n; base += stride) {
        long long t[kCmUnroll], p[kCmUnroll];
        bool in[kCmUnroll];
#pragma unroll
        for (int k = 0; k < kCmUnroll; ++k) {
            const long long i = base + (long long)k * kCmThreads + threadIdx.x;
            in[k] = i < n;
            t[k] = in[k] ? (long long)labels[i] : -1;
            p[k] = in[k] ? (long long)preds[i] : -1;
        }
#pragma unroll
        for (int k = 0; k < kCmUnroll; ++k) {
            const bool ok = in[k] && !(has_ignore && t[k] == ignore) &&      // validation.py:102
                            t[k] >= 0 && t[k] < n_class && p[k] >= 0 && p[k] < n_class;  // utils.py:78-83
            const u32 bin = ok ? (u32)(t[k] * n_class + p[k]) : 0u;          // utils.py:84-87
            cm_add<MODE>(sh, nbins, bin, ok, hist);
        }
    }
    cm_flush<MODE>(sh, nbins, hist);
}

// ---------------------------------------------------------------------------
// k_confusion_stream: the streaming form for maps whose bins fit a shared histogram and whose pointers
// allow vector loads - the validation loop's case.
//
//   * a thread owns kCmsRun * V CONSECUTIVE pixels (8 for int64 maps: four 16-byte loads per map), all
//     loads issued before any pixel is looked at: 1024 resident threads keep 128 KB per SM in flight;
//   * label maps are spatially coherent and predictions mostly right, so the pixels of a thread mostly share
//     one bin: the thread counts that bin in a register and only the others go to the shared histogram one
//     by one; one more atomic per thread delivers the register count.  About 0.3 shared atomics per pixel
//     on the 80 %-right synthetic maps, fewer on real ones, and no warp-wide operation at all;
//   * the lanes of a warp are kCmsRun * V pixels apart, i.e. in different label tiles: the atomics of one
//     warp instruction mostly hit different bins (few same-address replays);
//   * ONE histogram per block (n_class^2 words): zeroing and the flush touch 361 words, not 32 x 361;
//     blocks take equal contiguous chunks, two 512-thread blocks per SM.
#ifndef RTDA_CM_THREADS
#define RTDA_CM_THREADS 512
#endif
#ifndef RTDA_CM_BLOCKS_PER_SM
#define RTDA_CM_BLOCKS_PER_SM 2
#endif
#ifndef RTDA_CM_RUN
#define RTDA_CM_RUN 4
#endif
constexpr int kCmsThreads = RTDA_CM_THREADS;
constexpr int kCmsRun = RTDA_CM_RUN;

template <typename T, int V> struct VecLoad;      // V consecutive elements of type T in one load
template <typename T> struct VecLoad<T, 1> { static __device__ __forceinline__ void ld(const T *p, long long *o) { o[0] = (long long)*p; } };
template <> struct VecLoad<long long, 2> {
    static __device__ __forceinline__ void ld(const long long *p, long long *o) {
        const longlong2 v = *reinterpret_cast<const longlong2 *>(p); o[0] = v.x; o[1] = v.y; }
};
template <> struct VecLoad<int, 2> {
    static __device__ __forceinline__ void ld(const int *p, long long *o) {
        const int2 v = *reinterpret_cast<const int2 *>(p); o[0] = v.x; o[1] = v.y; }
};
template <> struct VecLoad<unsigned char, 2> {
    static __device__ __forceinline__ void ld(const unsigned char *p, long long *o) {
        const uchar2 v = *reinterpret_cast<const uchar2 *>(p); o[0] = v.x; o[1] = v.y; }
};
template <> struct VecLoad<int, 4> {
    static __device__ __forceinline__ void ld(const int *p, long long *o) {
        const int4 v = *reinterpret_cast<const int4 *>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct VecLoad<unsigned char, 4> {
    static __device__ __forceinline__ void ld(const unsigned char *p, long long *o) {
        const uchar4 v = *reinterpret_cast<const uchar4 *>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};

template <typename LabelT, typename PredT, int V>
__global__ void __launch_bounds__(kCmsThreads, RTDA_CM_BLOCKS_PER_SM)
k_confusion_stream(const LabelT *__restrict__ labels, const PredT *__restrict__ preds, long long n, int n_class,
                   long long ignore, int has_ignore, unsigned long long *__restrict__ hist) {
    extern __shared__ u32 sh[];
    constexpr int kPx = kCmsRun * V;                       // consecutive pixels of one thread item
    const int nbins = n_class * n_class;
    const u32 tid = threadIdx.x;
    const bool ign_live = has_ignore && ignore >= 0 && ignore < n_class;
    for (int i = tid; i < nbins; i += kCmsThreads) sh[i] = 0;
    // this block's contiguous chunk of thread items; the last item of the map may be partial
    const long long items = (n + kPx - 1) / kPx;
    const long long per_block = (items + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per_block;
    const long long hi = lo + per_block < items ? lo + per_block : items;
    __syncthreads();
    for (long long it = lo + tid; it < hi; it += kCmsThreads) {
        long long t[kPx], p[kPx];
        const long long px0 = it * kPx;
        if (px0 + kPx <= n) {
#pragma unroll
            for (int r = 0; r < kCmsRun; ++r) {
                VecLoad<LabelT, V>::ld(labels + px0 + r * V, t + r * V);
                VecLoad<PredT, V>::ld(preds + px0 + r * V, p + r * V);
            }
        } else {
#pragma unroll
            for (int q = 0; q < kPx; ++q) {
                t[q] = -1; p[q] = -1;
                if (px0 + q < n) { t[q] = (long long)labels[px0 + q]; p[q] = (long long)preds[px0 + q]; }
            }
        }
        u32 cur = 0xffffffffu, cnt = 0;
#pragma unroll
        for (int q = 0; q < kPx; ++q) {
            // validation.py:102 (ignored label) and utils.py:78-83 (0 <= label < n).  An ignore value outside
            // [0, n) - 255 with 19 classes - is already excluded by the range test: high words zero (one OR, one
            // test for both values) and two unsigned 32-bit comparisons instead of five 64-bit ones.
            const bool ok = (((u32)((unsigned long long)t[q] >> 32) | (u32)((unsigned long long)p[q] >> 32)) == 0u) &&
                            (u32)t[q] < (u32)n_class && (u32)p[q] < (u32)n_class && !(ign_live && t[q] == ignore);
            if (!ok) continue;
            const u32 bin = (u32)t[q] * (u32)n_class + (u32)p[q];                     // utils.py:84-87
            if (bin == cur) ++cnt;
            else if (cnt == 0) { cur = bin; cnt = 1; }
            else atomicAdd(&sh[bin], 1u);
        }
        if (cnt) atomicAdd(&sh[cur], cnt);
    }
    __syncthreads();
    for (int b = tid; b < nbins; b += kCmsThreads) {
        const u32 s = sh[b];
        if (s) atomicAdd(&hist[b], (unsigned long long)s);
    }
}

// ---------------------------------------------------------------------------
// k_confusion_tma: the streaming form with the maps staged in shared memory by bulk asynchronous copies (TMA 1-D).
//
// k_confusion_stream keeps 128 KB per SM in flight - what the registers of 1024 threads hold - so a 1024 x 2048 map
// (33.5 MB of int64 labels + predictions, 226 KB per SM) takes 1.8 dependent rounds of load -> count, and at that size
// the kernel is all latency: 10.5 us where a plain copy of the same bytes takes 6.1.  Here one block per SM owns a
// contiguous run of 2048-pixel chunks and a ring of kCtStages chunk buffers (6 x 32 KB for int64 maps): one thread
// issues two bulk copies per chunk (labels, predictions) on the chunk's mbarrier, the first kCtStages of them before
// anything else happens, so nearly the whole share of an SM is in flight at once and no register waits for memory.
// The threads read their pixels from shared memory (16-byte, conflict-free), count them as k_confusion_stream does
// (a register run for the thread's majority bin, shared atomics for the rest) and the buffer is refilled as soon
// as every thread has passed it.  Pixels past the last 16-pixel boundary of the map (the copies move multiples of 16
// bytes of either map) are read from global memory by the first threads.
#ifndef RTDA_CM_TMA            // 0 = k_confusion_stream only
#define RTDA_CM_TMA 1
#endif
// The ring is a template parameter set (THREADS, CHUNK pixels per stage, STAGES); <1024, 4096, 3> is what runs: few, large
// copies and one block-wide barrier per 4096 pixels (<512, 2048, 6> and <1024, 2048, 6>: 52.9 / 50.8 us for 8 maps against 42.5).
constexpr int kCtMaxClasses = 77;               // 192 KB of stages + 77^2 bins fit the 227 KB of an SM

template <typename T>
__device__ __forceinline__ void ct_pair(const T *base, u32 i, long long *o) { o[0] = (long long)base[i]; o[1] = (long long)base[i + 1]; }
template <>
__device__ __forceinline__ void ct_pair<long long>(const long long *base, u32 i, long long *o) {
    const longlong2 v = *reinterpret_cast<const longlong2 *>(base + i); o[0] = v.x; o[1] = v.y;
}
template <>
__device__ __forceinline__ void ct_pair<int>(const int *base, u32 i, long long *o) {
    const int2 v = *reinterpret_cast<const int2 *>(base + i); o[0] = v.x; o[1] = v.y;
}

template <typename LabelT, typename PredT, int kCtThreads, int kCtChunk, int kCtStages>
__global__ void __launch_bounds__(kCtThreads, 1)
k_confusion_tma(const LabelT *__restrict__ labels, const PredT *__restrict__ preds, long long n, int n_class,
                long long ignore, int has_ignore, unsigned long long *__restrict__ hist) {
    extern __shared__ __align__(128) unsigned char ct_smem[];
    constexpr int kCtPer = kCtChunk / kCtThreads;   // pixels per thread and chunk (4: two pairs kCtChunk / 2 apart)
    static_assert(kCtPer == 4, "a thread takes two pairs of pixels per chunk");
    constexpr u32 kStageBytes = kCtChunk * (sizeof(LabelT) + sizeof(PredT));
    u64 *full = reinterpret_cast<u64 *>(ct_smem + (size_t)kCtStages * kStageBytes);
    u32 *sh = reinterpret_cast<u32 *>(full + kCtStages);
    const int nbins = n_class * n_class;
    const u32 tid = threadIdx.x;
    const bool ign_live = has_ignore && ignore >= 0 && ignore < n_class;

    // this block's chunks: [c0, c1)
    const long long chunks = (n + kCtChunk - 1) / kCtChunk;
    const long long per = (chunks + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per;
    const long long c1 = c0 + per < chunks ? c0 + per : chunks;
    const int nloc = c1 > c0 ? (int)(c1 - c0) : 0;

    // pixels of chunk c that the bulk copies bring (a multiple of 16), and its pixels in all
    auto chunk_px = [&](long long c) -> u32 { const long long r = n - c * kCtChunk; return r < kCtChunk ? (u32)r : (u32)kCtChunk; };
    auto issue = [&](int k) {                     // chunk c0 + k into stage k % kCtStages (one thread)
        const int st = k % kCtStages;
        const long long c = c0 + k;
        const u32 al = chunk_px(c) & ~15u;
        if (al == 0u) return;
        unsigned char *buf = ct_smem + (size_t)st * kStageBytes;
        bulk_load(buf, labels + c * kCtChunk, al * (u32)sizeof(LabelT), &full[st]);
        bulk_load(buf + kCtChunk * sizeof(LabelT), preds + c * kCtChunk, al * (u32)sizeof(PredT), &full[st]);
    };
    if (tid == 0) {
        for (int s = 0; s < kCtStages; ++s) mbar_init(&full[s], 2);      // two copies arrive per chunk
        for (int k = 0; k < kCtStages && k < nloc; ++k) issue(k);
    }
    for (int i = tid; i < nbins; i += kCtThreads) sh[i] = 0;
    __syncthreads();

    u32 cur = 0xffffffffu, cnt = 0;               // the thread's run: cnt pixels counted for bin cur, not yet delivered
    constexpr u32 kNoBin = 0xffffffffu;
    auto bin_of = [&](long long t, long long p) -> u32 {
        // validation.py:102 (ignored label) and utils.py:78-83 (0 <= label < n); see k_confusion_stream
        const bool ok = (((u32)((unsigned long long)t >> 32) | (u32)((unsigned long long)p >> 32)) == 0u) &&
                        (u32)t < (u32)n_class && (u32)p < (u32)n_class && !(ign_live && t == ignore);
        return ok ? (u32)t * (u32)n_class + (u32)p : kNoBin;                 // utils.py:84-87
    };
    auto add_run = [&](u32 bin, u32 c) {
        if (bin == kNoBin) return;
        if (bin == cur) cnt += c;
        else if (cnt == 0) { cur = bin; cnt = c; }
        else atomicAdd(&sh[bin], c);
    };
    for (int k = 0; k < nloc; ++k) {
        const int st = k % kCtStages;
        const long long c = c0 + k;
        const u32 npx = chunk_px(c), al = npx & ~15u;
        if (al) {
            mbar_wait(&full[st], (u32)(k / kCtStages) & 1u);
            const LabelT *lab = reinterpret_cast<const LabelT *>(ct_smem + (size_t)st * kStageBytes);
            const PredT *prd = reinterpret_cast<const PredT *>(ct_smem + (size_t)st * kStageBytes + kCtChunk * sizeof(LabelT));
            long long t[kCtPer], p[kCtPer];
#pragma unroll
            for (int h = 0; h < kCtPer / 2; ++h) {
                const u32 i = (u32)h * (kCtChunk / (kCtPer / 2)) + 2u * tid;    // pairs of one warp instruction are contiguous
                ct_pair<LabelT>(lab, i, t + 2 * h);
                ct_pair<PredT>(prd, i, p + 2 * h);
            }
            // the bins of the thread's pixels, independently of each other; pixels past the copied part count nothing
            u32 b[kCtPer];
            bool same = true;
#pragma unroll
            for (int h = 0; h < kCtPer / 2; ++h) {
                const bool in = (u32)h * (kCtChunk / (kCtPer / 2)) + 2u * tid < al;         // (al is even)
                b[2 * h] = in ? bin_of(t[2 * h], p[2 * h]) : kNoBin;
                b[2 * h + 1] = in ? bin_of(t[2 * h + 1], p[2 * h + 1]) : kNoBin;
                same = same && b[2 * h] == b[0] && b[2 * h + 1] == b[0];
            }
            // coherent maps: the pixels of a thread (two neighbours, and two more a row or two below) mostly share a bin
            if (same) add_run(b[0], (u32)kCtPer);
            else {
#pragma unroll
                for (int q = 0; q < kCtPer; ++q) add_run(b[q], 1u);
            }
        }
        if (tid < npx - al) {                     // fewer than 16 pixels at the very end of the map
            const long long i = c * kCtChunk + al + tid;
            add_run(bin_of((long long)labels[i], (long long)preds[i]), 1u);
        }
        if (k + kCtStages < nloc) {               // block-uniform: the stage is refilled once everyone has read it
            __syncthreads();
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(k + kCtStages);
            }
        }
    }
    if (cnt) atomicAdd(&sh[cur], cnt);
    __syncthreads();
    for (int b = tid; b < nbins; b += kCtThreads) {
        const u32 s = sh[b];
        if (s) atomicAdd(&hist[b], (unsigned long long)s);
    }
}
template <typename LabelT, typename PredT>
static size_t confusion_tma_smem(int n_class, int chunk, int stages) {
    return (size_t)stages * chunk * (sizeof(LabelT) + sizeof(PredT)) + stages * sizeof(u64) + (size_t)n_class * n_class * sizeof(u32);
}
template <typename LabelT, typename PredT, int THREADS, int CHUNK, int STAGES>
static int launch_confusion_tma(const LabelT *l, const PredT *p, long long n, int n_class, long long ignore, int has_ignore,
                                unsigned long long *h, int max_classes, cudaStream_t st) {
    static bool attr_done[64] = {false};
    int dev = 0;
    RTDA_RETURN_IF_CUDA_ERROR(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        RTDA_RETURN_IF_CUDA_ERROR(cudaFuncSetAttribute(k_confusion_tma<LabelT, PredT, THREADS, CHUNK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)confusion_tma_smem<LabelT, PredT>(max_classes, CHUNK, STAGES)));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const long long chunks = (n + CHUNK - 1) / CHUNK;
    long long tgrid = sm_count();
    if (tgrid > chunks) tgrid = chunks;
    k_confusion_tma<LabelT, PredT, THREADS, CHUNK, STAGES><<<(unsigned)tgrid, THREADS, confusion_tma_smem<LabelT, PredT>(n_class, CHUNK, STAGES), st>>>(
        l, p, n, n_class, ignore, has_ignore, h);
    RTDA_CHECK_LAUNCH();
    return 0;
}

// Fused validation.py:95 (argmax over classes, first maximum wins) + the histogram.
template <typename LabelT, int MODE>
__global__ void __launch_bounds__(kCmThreads)
k_confusion_argmax(const float *__restrict__ logits, const LabelT *__restrict__ labels, long long B,
                   int n_class, long long HW, long long ignore, int has_ignore,
                   unsigned long long *__restrict__ hist) {
    extern __shared__ u32 sh[];
    const int nbins = n_class * n_class;
    if (MODE != 2) {
        const int words = (MODE == 0 ? kCmWarps : 1) * nbins;
        for (int i = threadIdx.x; i < words; i += kCmThreads) sh[i] = 0;
        __syncthreads();
    }
    const long long n = B * HW;
    const long long stride = (long long)gridDim.x * kCmThreads;
    for (long long base = (long long)blockIdx.x * kCmThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool in = i < n;
        long long t = -1;
        int arg = 0;
        if (in) {
            t = (long long)labels[i];
            const long long b = i / HW, hw = i - b * HW;
            const float *px = logits + (size_t)(b * n_class) * HW + hw;
            float best = ld_stream_f1(px);
            for (int c = 1; c < n_class; ++c) {
                const float v = ld_stream_f1(px + (size_t)c * HW);
                if (v > best || (v != v && best == best)) { best = v; arg = c; }  // NaN counts as maximal, like torch
            }
        }
        const bool ok = in && !(has_ignore && t == ignore) && t >= 0 && t < n_class;
        const u32 bin = ok ? (u32)(t * n_class + arg) : 0u;
        cm_add<MODE>(sh, nbins, bin, ok, hist);
    }
    cm_flush<MODE>(sh, nbins, hist);
}

static int cm_mode(int n_class, size_t *smem) {
    const size_t nb = (size_t)n_class * n_class * sizeof(u32);
    if (nb * kCmWarps <= 96 * 1024) { *smem = nb * kCmWarps; return 0; }
    if (nb <= 160 * 1024) { *smem = nb; return 1; }
    *smem = 0;
    return 2;
}

template <typename LabelT, typename PredT>
static int launch_confusion(const void *labels, const void *preds, long long n, int n_class, long long ignore,
                            int has_ignore, int64_t *hist, cudaStream_t st) {
    size_t smem = 0;
    const int mode = cm_mode(n_class, &smem);
    // one persistent 1024-thread block per SM: all loads of a block's share are in flight at once,
    // and every block ends with at most n_class^2 atomics into the int64 matrix
    long long grid = (n + kCmThreads * kCmUnroll - 1) / (kCmThreads * kCmUnroll);
    const long long cap = sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    auto *h = reinterpret_cast<unsigned long long *>(hist);
    const LabelT *l = (const LabelT *)labels;
    const PredT *p = (const PredT *)preds;
    ProfileScope ps(kStageConfusion, st);
    {
        // streaming form: V pixels per load (16 bytes of the wider map), when both pointers are aligned for it
        constexpr int kWide = sizeof(LabelT) > sizeof(PredT) ? (int)sizeof(LabelT) : (int)sizeof(PredT);
        constexpr int V = kWide == 8 ? 2 : 4;
        const size_t ssm = (size_t)n_class * n_class * sizeof(u32);
#if RTDA_CM_TMA
        // staged by bulk copies: 16-byte aligned maps (torch allocations are; a slice starting at an odd element is not)
        if (((uintptr_t)labels & 15) == 0 && ((uintptr_t)preds & 15) == 0) {
            // several 4096-pixel chunks per SM: the three-stage ring (8 maps per call: 42.5 us against 47.7 us of k_confusion_stream)
            if (n_class <= kCtMaxClasses && n >= 4096ll * 8 * sm_count())
                return launch_confusion_tma<LabelT, PredT, 1024, 4096, 3>(l, p, n, n_class, ignore, has_ignore, h, kCtMaxClasses, st);
            // (one 1024 x 2048 map through a ring of seven 2048-pixel stages - 225.5 KB, its whole share of an SM in flight at
            // once, nothing refilled - measured 10.3 us, as k_confusion_stream below: bytes in flight are not what one map waits for)
        }
#endif
        if (ssm <= 40 * 1024 && ((uintptr_t)labels % (V * sizeof(LabelT))) == 0 && ((uintptr_t)preds % (V * sizeof(PredT))) == 0) {
            const long long items = (n + kCmsRun * V - 1) / (kCmsRun * V);
            long long sgrid = (long long)RTDA_CM_BLOCKS_PER_SM * sm_count();
            const long long need = (items + kCmsThreads - 1) / kCmsThreads;          // at least one item per thread
            if (sgrid > need) sgrid = need;
            if (sgrid < 1) sgrid = 1;
            k_confusion_stream<LabelT, PredT, V><<<(unsigned)sgrid, kCmsThreads, ssm, st>>>(l, p, n, n_class, ignore, has_ignore, h);
            RTDA_CHECK_LAUNCH();
            return 0;
        }
    }
    if (mode == 0) {
        if (smem > 48 * 1024)
            RTDA_RETURN_IF_CUDA_ERROR(cudaFuncSetAttribute(k_confusion<LabelT, PredT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_confusion<LabelT, PredT, 0><<<(unsigned)grid, kCmThreads, smem, st>>>(l, p, n, n_class, ignore, has_ignore, h);
    } else if (mode == 1) {
        if (smem > 48 * 1024)
            RTDA_RETURN_IF_CUDA_ERROR(cudaFuncSetAttribute(k_confusion<LabelT, PredT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_confusion<LabelT, PredT, 1><<<(unsigned)grid, kCmThreads, smem, st>>>(l, p, n, n_class, ignore, has_ignore, h);
    } else {
        k_confusion<LabelT, PredT, 2><<<(unsigned)grid, kCmThreads, 0, st>>>(l, p, n, n_class, ignore, has_ignore, h);
    }
    RTDA_CHECK_LAUNCH();
    return 0;
}

template <typename LabelT>
static int launch_confusion_pred(const void *labels, const void *preds, int pred_kind, long long n, int n_class,
                                 long long ignore, int has_ignore, int64_t *hist, cudaStream_t st) {
    switch (pred_kind) {
        case RTDA_LABEL_I64: return launch_confusion<LabelT, long long>(labels, preds, n, n_class, ignore, has_ignore, hist, st);
        case RTDA_LABEL_U8: return launch_confusion<LabelT, unsigned char>(labels, preds, n, n_class, ignore, has_ignore, hist, st);
        case RTDA_LABEL_I32: return launch_confusion<LabelT, int>(labels, preds, n, n_class, ignore, has_ignore, hist, st);
    }
    return RTDA_E_ARG;
}

template <typename LabelT>
static int launch_confusion_argmax(const float *logits, const void *labels, long long B, int n_class, long long HW,
                                   long long ignore, int has_ignore, int64_t *hist, cudaStream_t st) {
    size_t smem = 0;
    const int mode = cm_mode(n_class, &smem);
    const long long n = B * HW;
    long long grid = (n + kCmThreads - 1) / kCmThreads;
    const long long cap = 2ll * sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    auto *h = reinterpret_cast<unsigned long long *>(hist);
    const LabelT *l = (const LabelT *)labels;
    ProfileScope ps(kStageConfusionArgmax, st);
    if (mode == 0) {
        if (smem > 48 * 1024)
            RTDA_RETURN_IF_CUDA_ERROR(cudaFuncSetAttribute(k_confusion_argmax<LabelT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_confusion_argmax<LabelT, 0><<<(unsigned)grid, kCmThreads, smem, st>>>(logits, l, B, n_class, HW, ignore, has_ignore, h);
    } else if (mode == 1) {
        if (smem > 48 * 1024)
            RTDA_RETURN_IF_CUDA_ERROR(cudaFuncSetAttribute(k_confusion_argmax<LabelT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_confusion_argmax<LabelT, 1><<<(unsigned)grid, kCmThreads, smem, st>>>(logits, l, B, n_class, HW, ignore, has_ignore, h);
    } else {
        k_confusion_argmax<LabelT, 2><<<(unsigned)grid, kCmThreads, 0, st>>>(logits, l, B, n_class, HW, ignore, has_ignore, h);
    }
    RTDA_CHECK_LAUNCH();
    return 0;
}

}  // namespace rtda

using namespace rtda;

extern "C" int rtda_confusion_accumulate(const void *labels, int label_kind, const void *preds, int pred_kind,
                                         int64_t n, int n_class, int64_t ignore, int has_ignore,
                                         int64_t *hist, void *stream) {
    if (!hist || n < 0 || n_class <= 0) return RTDA_E_ARG;
    if (n_class > 46340 || n >= (1ll << 40)) return RTDA_E_RANGE;  // n_class^2 must fit int32, block counters u32
    if (n == 0) return 0;
    if (!labels || !preds) return RTDA_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (label_kind) {
        case RTDA_LABEL_I64: return launch_confusion_pred<long long>(labels, preds, pred_kind, n, n_class, ignore, has_ignore, hist, st);
        case RTDA_LABEL_U8: return launch_confusion_pred<unsigned char>(labels, preds, pred_kind, n, n_class, ignore, has_ignore, hist, st);
        case RTDA_LABEL_I32: return launch_confusion_pred<int>(labels, preds, pred_kind, n, n_class, ignore, has_ignore, hist, st);
    }
    return RTDA_E_ARG;
}

extern "C" int rtda_confusion_argmax_accumulate(const float *logits, const void *labels, int label_kind,
                                                int64_t B, int n_class, int64_t HW, int64_t ignore, int has_ignore,
                                                int64_t *hist, void *stream) {
    if (!hist || B < 0 || HW < 0 || n_class <= 0) return RTDA_E_ARG;
    if (n_class > 46340 || B * HW >= (1ll << 40)) return RTDA_E_RANGE;
    if (B * HW == 0) return 0;
    if (!logits || !labels) return RTDA_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (label_kind) {
        case RTDA_LABEL_I64: return launch_confusion_argmax<long long>(logits, labels, B, n_class, HW, ignore, has_ignore, hist, st);
        case RTDA_LABEL_U8: return launch_confusion_argmax<unsigned char>(logits, labels, B, n_class, HW, ignore, has_ignore, hist, st);
        case RTDA_LABEL_I32: return launch_confusion_argmax<int>(logits, labels, B, n_class, HW, ignore, has_ignore, hist, st);
    }
    return RTDA_E_ARG;
}

extern "C" int rtda_version(void) { return RTDA_ABI_VERSION; }

extern "C" const char *rtda_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case RTDA_E_ARG: return "rtda: bad argument (null pointer, non-positive size or unknown label kind)";
        case RTDA_E_RANGE: return "rtda: size out of range (classes > RTDA_MAX_CLASSES or segment >= 2^30 pixels)";
        case RTDA_E_WORKSPACE: return "rtda: workspace smaller than rtda_lovasz_workspace_bytes()";
        case RTDA_E_ALIGN: return "rtda: pointer not 16-byte aligned";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "rtda: unknown error";
}
