// Bucket path of the Lovasz-Softmax forward: two data movements instead of four radix passes.
//
// The ranks lovasz_grad needs (lovasz_loss.py:47-70,173-176) are those of a full descending sort of
// every class's errors.  The LSD radix path (lovasz.cu) moves every (key, payload) pair through
// shared memory four times.  For the smooth key distributions softmax outputs have, a sample of the
// keys predicts the distribution well enough to finish in two movements:
//
//   k_split   a few blocks per segment (group, class): 12-bit-prefix histogram of a 1/16 sample of the
//             segment's keys -> a monotone piecewise-linear map key -> bucket with equal expected
//             occupancy (kBMean elements per bucket).
//   k_part    level 1: every element goes to its bucket's fixed-capacity region (kBTile slots) through
//             a shared-memory exchange, one global cursor claim per (tile, bucket).  Unordered inside
//             a bucket: level 2 orders by (key, pixel index) explicitly, so no stability is needed and
//             there is no look-back chain.  The segment's map comes in by one bulk copy (TMA 1-D).
//   k_plan    one thread per bucket: a 32-byte record {count, elements and foreground elements of the
//             earlier buckets, flags; in-bucket linear key map} so that a level-2 block starts after a
//             single round trip to L2.
//   k_local   level 2: one block per bucket, staged in shared memory by one bulk copy.  Interpolation
//             sort: fine slot = monotone function of the key (the straight line of k_plan, else the
//             segment's map), one atomic on packed counters per element, a scan, one placement;
//             elements sharing a slot (about 15 %) are ordered by direct comparison.  rank = elements
//             in earlier buckets + local rank, likewise the foreground rank; lovasz_grad, the dot
//             product and the gradient scatter follow as in k_rank_last.
//
// The order produced is exactly the LSD path's: ascending key (= descending error), ties by ascending
// pixel index, so both paths are bit-identical.  Anything the map cannot handle - a bucket that
// overflows its region (a massive exact tie that is not on an interval boundary), more buckets than
// kBMaxBuckets - raises the segment's fallback flag on the device and the LSD kernels, which always
// run behind (persistent grids that exit at once when no flag is set), redo that segment.  No host
// synchronisation is involved.
//
// Elements with error exactly 0 are not sorted at all: they sit behind every other element, contribute
// 0 * g to the loss and sgn(0) = 0 to the gradient, and dropping a suffix changes no other rank.
#pragma once
// (included inside namespace rtda by lovasz.cu, after error_key / div_rn_counts / LossAcc)

template <int NW, typename T>
__device__ __forceinline__ T block_excl_scan(T v, T *warp_tmp /*[NW]*/, T *total) {
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
    for (u32 w = 0; w < (u32)NW; ++w) {
        T x = warp_tmp[w];
        if (w < warp) woff += x;
        tot += x;
    }
    __syncthreads();
    if (total) *total = tot;
    return woff + incl - v;
}

// Monotone re-parametrisation of the sort key in which softmax errors are smoothly distributed.
// Ascending key = descending error.  Errors in [0.5, 1] come from |1 - p| with p small (a confident
// miss) or from p itself close to 1: there 1 - err is exact in fp32 and spreads over many binades
// while err is squeezed against 1, so tau follows the bits of 1 - err.  Errors below 0.5 follow the
// bits of err (descending).  In tau the density is that of a log-scale histogram, which a
// piecewise-linear map models well; in the raw key the last bin under err = 1 holds a spike.
__device__ __forceinline__ u32 tau_of(u32 key) {
    const u32 eb = ~key;                             // bits of the error
    // both forms are evaluated (the subtraction is harmless on any bit pattern) and selected: two
    // compares and two selects; written as a chain of `if`s the compiler emits divergence-safe branches
    const u32 hi = __float_as_uint(__fsub_rn(1.0f, __uint_as_float(eb)));
    const u32 lo = 0x7E000000u - eb;
    u32 t, r;
    asm("{\n.reg .pred p;\nsetp.ge.u32 p, %1, 0x3F000000;\nselp.u32 %0, %2, %3, p;\n}" : "=r"(t) : "r"(eb), "r"(hi), "r"(lo));
    asm("{\n.reg .pred p;\nsetp.gt.u32 p, %1, 0x3F800000;\nselp.u32 %0, 0, %2, p;\n}" : "=r"(r) : "r"(eb), "r"(t));
    return r;                                        // err > 1 or NaN (garbage probabilities): first
}

// Monotone map (key, pixel index) -> position in the sampled cumulative distribution, in units of
// buckets, 16.16 fixed point: the top 12 bits of tau select a table interval [tab[p], tab[p+1]), its
// low 19 bits interpolate inside it.  Integer part = bucket, fractional part = position inside the
// bucket (level 2 turns it into a fine slot).
//
// Boundary ties.  Saturated predictions put thousands of elements on a handful of keys: a foreground
// probability that rounded to 1 - k * 2^-24 has the error k * 2^-24 exactly, and for small k that
// value has an all-zero low mantissa, i.e. it is the FIRST key of its table interval (low 19 bits of
// tau = 0).  k_split counts the samples sitting exactly there; when they are many the interval is
// flagged (bit 31 of tab[p]) and ties[p] holds the width T, in the same units, that those elements
// deserve.  They are then spread over [x0, x0 + T) by pixel index - which is their sort order - and
// the other keys of the interval share [x0 + T, x0 + dx).  So a tie of any size is split across
// buckets and fine slots instead of overflowing one.
constexpr u32 kTabTie = 1u << 31;
__device__ __forceinline__ u32 cdf_pos_ties(u32 key, u32 idx, u32 n, const u32 *tab, const u32 *__restrict__ ties, u32 xmax) {
    const u32 tau = tau_of(key);
    const u32 p = tau >> 19;                         // <= 0x7E0 << 1 = 4032
    const u32 t0 = tab[p];
    const u32 x0 = t0 & ~kTabTie;
    u32 dx = (tab[p + 1] & ~kTabTie) - x0;
    u32 x = x0;
    if (t0 & kTabTie) {
        const u32 T = ties[p];
        if ((tau & 0x7FFFFu) == 0u) return x0 + (u32)(((u64)idx * T) / n);     // < x0 + T <= x0 + dx
        x += T;
        dx -= T;
    }
    x += __umulhi(tau << 13, dx);                    // (low19 * dx) >> 19
    return x < xmax ? x : xmax;
}
// The same map for a segment without flagged intervals (the table then carries no flag bits).
__device__ __forceinline__ u32 cdf_pos_plain(u32 key, const u32 *tab, u32 xmax) {
    const u32 tau = tau_of(key);
    const u32 p = tau >> 19;
    const u32 x0 = tab[p], dx = tab[p + 1] - x0;
    const u32 x = x0 + __umulhi(tau << 13, dx);
    return x < xmax ? x : xmax;
}
__device__ __forceinline__ u32 cdf_pos(bool has_ties, u32 key, u32 idx, u32 n, const u32 *tab, const u32 *__restrict__ ties, u32 xmax) {
    return has_ties ? cdf_pos_ties(key, idx, n, tab, ties, xmax) : cdf_pos_plain(key, tab, xmax);
}
constexpr u32 kNbTies = 1u << 31;                    // bit of nb[seg]: the segment's map has flagged intervals

// ---------------------------------------------------------------------------
// k_split: sample histogram -> cumulative map of one segment.  gridDim.y blocks share a segment:
// each adds its part of the sample to the segment's global histogram, the last one to finish
// builds the map.
__global__ void __launch_bounds__(kSplitThreads, 2)
k_split(const float *__restrict__ probas, const unsigned char *__restrict__ codes, Geom geo,
        const u32 *__restrict__ seg_fg, int force_lsd, u32 *__restrict__ ghist /*[S][kSplitWords], zeroed*/,
        u32 *__restrict__ table, u32 *__restrict__ ties, u32 *__restrict__ nb_out, u32 *__restrict__ fb_flag,
        u32 *__restrict__ fb_any, float *__restrict__ seg_scale, float *__restrict__ grad) {
    // hist: sample count per bin.  zc: how many of them sit exactly on the bin's first key (low 19
    // bits of tau = 0): the boundary ties cdf_pos spreads by pixel index.
    extern __shared__ u32 split_smem[];
    u32 *hist = split_smem, *zc = split_smem + kTabBins;
    __shared__ u32 wtmp[kSplitThreads / 32];
    __shared__ u32 s_mvalid, s_last;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 seg = blockIdx.x;
    const u32 g = seg / (u32)geo.C, c = seg - g * (u32)geo.C;
    if (seg_fg[seg] == 0) {                       // class absent from this group
        if (tid == 0 && blockIdx.y == 0) nb_out[seg] = 0;
        if (geo.final_grad && grad) {
            // the final gradient has no separate zero-filling pass: the planes of an absent class are zeroed here
            float *gb = grad + geo.grad_seg_off(seg);            // class c of the group's (first) image
            for (u32 bi = 0; bi < (u32)geo.Bg; ++bi, gb += (size_t)geo.C * (size_t)geo.HW)
                for (u32 i = blockIdx.y * kSplitThreads + tid; i < (u32)geo.HW; i += gridDim.y * kSplitThreads) gb[i] = 0.0f;
        }
        return;
    }
    if (tid == 0 && blockIdx.y == 0) {
        // what the forward multiplies this segment's gradient by: 1 for the unit gradient; in final-gradient
        // mode the autograd factors of the two means, (grad_out / G) / K with the divisions k_bwd does
        float sc = 1.0f;
        if (geo.final_grad) {
            sc = geo.grad_hint;                    // the grad_out the gradient is written for (1 unless the caller said otherwise)
            int K = 0;
            for (int k = 0; k < geo.C; ++k) K += seg_fg[g * geo.C + k] > 0 ? 1 : 0;
            if (geo.G > 1) sc = __fdiv_rn(sc, (float)geo.G);
            if (K > 1) sc = __fdiv_rn(sc, (float)K);
        }
        seg_scale[seg] = sc;
    }
    if (force_lsd) {
        if (tid == 0 && blockIdx.y == 0) { nb_out[seg] = 0; fb_flag[seg] = 1u; *fb_any = 1u; }
        return;
    }
    // (the sample is binned in shared memory first: direct reductions on the global histogram were
    // measured slower, 69 vs 34 us at c2 - the populated bins are few and hot)
    u32 *gh = ghist + (size_t)seg * kSplitWords;
    u32 *gzc = gh + kTabBins + 2;
    for (int i = tid; i < kTabBins; i += kSplitThreads) { hist[i] = 0; zc[i] = 0; }
    if (tid == 0) s_mvalid = 0;
    __syncthreads();
    const u32 n = (u32)geo.n, HW = (u32)geo.HW;
    u32 mvalid = 0;
    {
        // runs of 32 consecutive pixels (one 128-byte line of probabilities per warp load), one run of
        // every SG (all of them for small segments)
        u32 SG = n / 16384u;
        SG = SG < 1u ? 1u : (SG > (u32)kSampleStride ? (u32)kSampleStride : SG);
        const u32 runs = (n + 32u * SG - 1u) / (32u * SG);
        const unsigned char *cg = codes + (size_t)g * n;
        const u32 wpb = kSplitThreads / 32;
        // kSplitUnroll runs per warp and iteration: their loads are issued before any of them is
        // binned, so a warp keeps that many DRAM round trips in flight instead of one
        const u32 ustep = gridDim.y * wpb;
        // (the index arithmetic was 42 % of this kernel's instructions as two integer divisions per sample:
        // the run's place inside its stride by a mask when SG is a power of two - it is 16 for every segment of
        // 256 K elements or more -, the image of a pixel by a multiplication with 1 / HW and one correction step)
        const u32 sg_mask = (SG & (SG - 1u)) == 0u ? SG - 1u : 0xFFFFFFFFu;
        const float rcp_hw = 1.0f / (float)HW;
        const float *pc = probas + ((size_t)g * geo.Bg * geo.C + c) * (size_t)HW;   // class c of the group's first image
        const size_t img_step = (size_t)(geo.C - 1) * HW;                          // i + bi * img_step: same class, image bi
        for (u32 u0 = blockIdx.y * wpb + (tid >> 5); u0 < runs; u0 += kSplitUnroll * ustep) {
            u32 code[kSplitUnroll];
            float p[kSplitUnroll];
#pragma unroll
            for (int q = 0; q < kSplitUnroll; ++q) {
                const u32 u = u0 + (u32)q * ustep;
                const u32 u7 = u * 7u;
                const u32 i = 32u * SG * u + 32u * (sg_mask != 0xFFFFFFFFu ? (u7 & sg_mask) : (u7 % SG)) + lane;
                code[q] = 255u;
                p[q] = 0.0f;
                if (u < runs && i < n) {
                    u32 bi = 0;
                    if (geo.Bg > 1) {
                        bi = (u32)(__uint2float_rz(i) * rcp_hw);      // within 1 of i / HW (i < 2^30)
                        if (bi * HW > i) --bi;
                        if ((bi + 1u) * HW <= i) ++bi;
                    }
                    code[q] = cg[i];                 // two independent loads
                    p[q] = pc[(size_t)i + bi * img_step];
                }
            }
#pragma unroll
            for (int q = 0; q < kSplitUnroll; ++q) {
                if (code[q] == 255u) continue;
                ++mvalid;
                u32 neg;
                const u32 key = error_key(p[q], true, code[q] == c, &neg);
                if (key == 0xFFFFFFFFu) continue;      // error exactly 0: never sorted
                const u32 tau = tau_of(key);
                const u32 bin = tau >> 19;
                atomicAdd(&hist[bin], 1u);
                if ((tau & 0x7FFFFu) == 0u) atomicAdd(&zc[bin], 1u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mvalid += __shfl_xor_sync(RTDA_FULL_MASK, mvalid, o);
    if (lane == 0 && mvalid) atomicAdd(&s_mvalid, mvalid);
    __syncthreads();
    if (gridDim.y > 1) {
        for (int i = tid; i < kTabBins; i += kSplitThreads) {
            if (hist[i]) {
                atomicAdd(&gh[i], hist[i]);
                if (zc[i]) atomicAdd(&gzc[i], zc[i]);
            }
        }
        if (tid == 0 && s_mvalid) atomicAdd(&gh[kTabBins], s_mvalid);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&gh[kTabBins + 1], 1u) == gridDim.y - 1u) ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int i = tid; i < kTabBins; i += kSplitThreads) { hist[i] = ld_relaxed(&gh[i]); zc[i] = ld_relaxed(&gzc[i]); }
        if (tid == 0) s_mvalid = ld_relaxed(&gh[kTabBins]);
        __syncthreads();
    }
    constexpr int kPerT = kTabBins / kSplitThreads;
    u32 h[kPerT], hsum = 0;
#pragma unroll
    for (int k = 0; k < kPerT; ++k) { h[k] = hist[kPerT * tid + k]; hsum += h[k]; }
    u32 m_nz = 0;
    const u32 excl = block_excl_scan<kSplitThreads / 32, u32>(hsum, wtmp, &m_nz);
    u32 valid = 0;
    for (int k = 0; k < geo.C; ++k) valid += seg_fg[g * geo.C + k];
    const u32 mv = s_mvalid;
    const double est = mv ? (double)valid * (double)m_nz / (double)mv : (double)valid;
    const double q = ceil(est / (double)kBMean);
    const u32 NB = q < 1.0 ? 1u : (q > 1e9 ? 1000000000u : (u32)q);
    if (NB > (u32)geo.NBK) {                      // more buckets than the workspace was sized for
        if (tid == 0) { nb_out[seg] = 0; fb_flag[seg] = 1u; *fb_any = 1u; }
        return;
    }
    // cumulative map with a floor of 1/kPrior sample per bin: a bin the sample missed still gets a
    // little width, so the elements that do fall there spread over slots instead of piling into one
    u32 *tab = table + (size_t)seg * kTabStride;
    u32 *tie = ties + (size_t)seg * kTabBins;
    const u64 scale = (u64)NB << 16;
    const u64 denom = (u64)kPrior * m_nz + kTabBins;
    u64 cum = (u64)kPrior * excl + (u64)kPerT * tid;
    u32 SG = (u32)geo.n / 16384u;
    SG = SG < 1u ? 1u : (SG > (u32)kSampleStride ? (u32)kSampleStride : SG);
    bool any_tie = false;
#pragma unroll
    for (int k = 0; k < kPerT; ++k) {
        const u32 bin = kPerT * tid + k;
        const u32 z = zc[bin];
        // boundary ties worth spreading: an estimated kTieMin elements or more on one key (smaller
        // ties are a bucket's own business: 8-bit slot counters, else its bitonic sort).  Never bin 0:
        // tau = 0 is shared by the error 1.0 and by garbage (errors above 1).
        const bool spread = bin != 0u && z * SG >= (u32)kTieMin;
        const u32 T = spread ? (u32)(((u64)kPrior * z * scale) / denom) : 0u;
        tab[bin] = (u32)((cum * scale) / denom) | (T ? kTabTie : 0u);
        tie[bin] = T;
        any_tie |= T != 0u;
        cum += (u64)kPrior * h[k] + 1u;
    }
    const int seg_ties = __syncthreads_or(any_tie);
    if (tid == 0) { tab[kTabBins] = (u32)scale; nb_out[seg] = NB | (seg_ties ? kNbTies : 0u); }
}

// (mbarrier + bulk-copy helpers: common.cuh)

// ---------------------------------------------------------------------------
// k_part: level 1.  One block per (segment, tile of kPTile pixels), segment-major.
constexpr u32 kPZeroBytes = 2048;                    // source of the zero fill (RTDA_SKIP_ZERO_STORES)
struct PartShared {
    u64 exch[kPTile];
    u32 tab[kTabBins + 4];
#if RTDA_SKIP_ZERO_STORES
    __align__(16) u32 zeros[kPZeroBytes / 4];
#endif
    u64 mbar;
    u32 scan_tmp[kPWarps];
    u32 skip;
    u32 tile_add;
    float *zbase;        // where the segment's gradient starts (Geom.grad_seg_off)
    u32 arrays[1];       // cnt[nbk_pad] (count | fg << 16, then the global destination), pre[nbk_pad]
};
constexpr u32 kNoDst = 0xFFFFFFFFu;
static inline size_t part_smem_bytes(int nbk) {
    const size_t pad = ((size_t)nbk + kPThreads - 1) / kPThreads * kPThreads;
    return sizeof(PartShared) + 2 * pad * sizeof(u32);
}

template <int TPB>                                  // consecutive tiles of one segment per block (part_tiles_per_block)
__global__ void __launch_bounds__(kPThreads, kPBlocksPerSM)
k_part(const float *__restrict__ probas, const unsigned char *__restrict__ codes, Geom geo,
       const u32 *__restrict__ seg_fg, const u32 *__restrict__ table, const u32 *__restrict__ ties,
       const u32 *__restrict__ nb, u64 *__restrict__ cursor, u64 *__restrict__ out, u32 *__restrict__ fb_flag,
       u32 *__restrict__ fb_any, float *__restrict__ unit_grad) {
    constexpr u32 tpb = (u32)TPB;
    extern __shared__ __align__(16) unsigned char part_smem[];
    PartShared &sm = *reinterpret_cast<PartShared *>(part_smem);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a block takes tpb consecutive tiles of one segment (part_tiles_per_block): the map is loaded once for them
    const u32 bps = ((u32)geo.T1 + tpb - 1u) / tpb;
    const u32 seg = blockIdx.x / bps;
    const u32 tile0 = (blockIdx.x - seg * bps) * tpb;
    // other blocks of this kernel may raise the flag while this one starts: one thread decides for the block
    if (tid == 0) {
        sm.skip = (seg_fg[seg] == 0 || ld_relaxed(fb_flag + seg) != 0) ? 1u : 0u;
        sm.zbase = unit_grad ? unit_grad + geo.grad_seg_off(seg) : nullptr;
        mbar_init(&sm.mbar, 1);
        // the segment's map (16 KB) comes in by one bulk copy; it is first needed after the tile's own
        // loads have returned, so its latency is hidden behind them (a register-staged copy stalled
        // every block for a full round trip to L2: 12 % of this kernel's stall samples)
        if (!sm.skip) bulk_load(sm.tab, table + (size_t)seg * kTabStride, kTabStride * (u32)sizeof(u32), &sm.mbar);
    }
#if RTDA_SKIP_ZERO_STORES
    for (u32 i = tid; i < kPZeroBytes / 4; i += kPThreads) sm.zeros[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the bulk stores below read these words
#endif
    __syncthreads();
    if (sm.skip) return;
    const u32 nb_raw = nb[seg];
    const u32 NB = nb_raw & ~kNbTies;
    const bool has_ties = (nb_raw & kNbTies) != 0u;
    const u32 xmax = (NB << 16) - 1u;
    u32 *sm_cnt = sm.arrays;
    u32 *sm_pre = sm.arrays + ((u32)geo.NBK + kPThreads - 1) / kPThreads * kPThreads;
    const u32 n = (u32)geo.n;
    const u32 *tie = ties + (size_t)seg * kTabBins;         // read through L1: touched only for flagged intervals
    for (u32 tile = tile0; tile < tile0 + tpb; ++tile) {
    const u32 tile_base = tile * (u32)kPTile;
    if (tile_base >= n) break;
    const u32 count = (n - tile_base < (u32)kPTile) ? n - tile_base : (u32)kPTile;
    __syncthreads();          // the previous tile no longer uses the shared arrays

    // per element: the key and one word of state {fg:31, neg:30, live:29, bucket:P..P+10, rank in (tile, bucket):0..P-1}
    // with P = kPPosBits (a rank is below the tile size) - the layout of the exchanged payload, so that
    // one AND/OR builds it; the pixel index is recomputed from the element's position in the tile
    constexpr u32 kPB = (u32)kPPosBits, kPosMask = (1u << kPB) - 1u;
    u32 key[kPItems], st[kPItems];
    const u32 idx0 = warp * (32 * kPItems) + lane;
    const u32 i0 = tile_base + idx0;
    constexpr u32 kLive = 1u << 29;
    {
        const u32 g = seg / (u32)geo.C, c = seg - g * (u32)geo.C;
        const u32 HW = (u32)geo.HW;
        const u32 bi = i0 / HW;
        u32 hw = i0 - bi * HW;
        const unsigned char *cg = codes + (size_t)g * n + i0;
        const float *pp = probas + (((size_t)g * geo.Bg + bi) * geo.C + c) * (size_t)HW + hw;
        const size_t wrap = (size_t)(geo.C - 1) * HW;          // same class, next image
        float p[kPItems];
        u32 code[kPItems];
        // the probability is loaded whether or not the pixel turns out to be ignored: the two loads
        // are independent, so a tile waits for one global-memory latency, not two
        const u32 hw_tile = tile_base % HW;
        if (count == (u32)kPTile && hw_tile + (u32)kPTile <= HW) {
            // full tile inside one image (the common case): fixed offsets, no predicates
#pragma unroll
            for (int j = 0; j < kPItems; ++j) {
                code[j] = cg[32 * j];
                p[j] = ld_stream_f1(pp + 32 * j);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPItems; ++j) {
                const bool in = idx0 + 32 * j < count;
                p[j] = 0.0f;
                code[j] = 255u;
                if (in) {
                    code[j] = cg[32 * j];
                    p[j] = ld_stream_f1(pp);
                }
                pp += 32;
                hw += 32;
                while (hw >= HW) { hw -= HW; pp += wrap; }
            }
        }
        for (u32 b = tid; b < NB; b += kPThreads) sm_cnt[b] = 0;
#if RTDA_SKIP_ZERO_STORES
        // The gradient of the whole tile is first written as the zero most of it is (two thirds of lovasz_grad's
        // first differences round to exactly 0 in fp32 at c2), so that level 2 scatters only the non-zero third:
        // its 4-byte stores, one LSU cycle per lane, are what bounds its last phase.  The fill is a handful of bulk
        // copies (TMA, shared -> global) from a block of zeros, issued by one thread: no LSU work at all.  Segments
        // whose length is not a multiple of 4 elements (16-byte granularity of the copies) store element-wise below.
        // (several images per group in final-gradient mode, Geom.img_step: a tile inside one image is contiguous in NCHW too)
        // (Geom.img_step != 0: HW is a multiple of the tile, so a tile lies in one image and is contiguous in NCHW too)
        // (the copies need 16-byte aligned addresses: a segment length that is a multiple of 4 with HW % 4 != 0 - four
        // 33 x 33 maps - comes with a buffer that is only promised to be 4-byte aligned)
        const bool bulk_zero = unit_grad != nullptr && (n & 3u) == 0u && (reinterpret_cast<size_t>(unit_grad) & 15) == 0;
        // the pixel index of the tile's first element; read after the barriers below, in the read-out
        if (tid == 0) sm.tile_add = tile_base + (tile_base - hw_tile) / HW * (u32)geo.img_step;
        if (bulk_zero && tid == 0) {
            char *zdst = reinterpret_cast<char *>(sm.zbase + sm.tile_add);
            for (u32 off = 0; off < count * 4u; off += kPZeroBytes) {
                const u32 bytes = count * 4u - off < kPZeroBytes ? count * 4u - off : kPZeroBytes;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(zdst + off), "r"(smem_u32(sm.zeros)), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (unit_grad != nullptr && !bulk_zero) {    // (block-uniform, rare; ignored pixels included, as in the bulk fill)
            float *ugz = sm.zbase;
            for (u32 j = 0; j < (u32)kPItems; ++j)
                if (idx0 + 32u * j < count) ugz[i0 + 32u * j] = 0.0f;     // (img_step == 0 here)
        }
#endif
#pragma unroll
        for (int j = 0; j < kPItems; ++j) {
            const bool fg = code[j] == c;
            u32 neg;
            key[j] = error_key(p[j], true, fg, &neg);
            st[j] = (fg ? kPayloadFg : 0u) | neg;
            if (code[j] != 255u) {
                if (key[j] != 0xFFFFFFFFu) st[j] |= kLive;
#if RTDA_SKIP_ZERO_STORES
#else
                else if (unit_grad)       // error exactly 0: gradient 0, not sorted (rare: the address is worked out here)
                    unit_grad[(size_t)seg * n + i0 + 32 * j] = 0.0f;
#endif
            }
        }
    }
    __syncthreads();
    if (tile == tile0) mbar_wait(&sm.mbar, 0);       // the map has arrived (block-uniform; later tiles find it there)
#if RTDA_EXPERIMENT_PART_CUT == 3
    { u32 x = 0;
#pragma unroll
      for (int j = 0; j < kPItems; ++j) x ^= key[j] ^ st[j];
      if (x == 0x12345678u) fb_flag[seg] = 1u;
      continue; }
#endif

    if (!has_ties) {                                 // block-uniform: the common case keeps the short map
#pragma unroll
        for (int j = 0; j < kPItems; ++j) {
            if (st[j] & kLive) {
                const u32 b = cdf_pos_plain(key[j], sm.tab, xmax) >> 16;
                const u32 old = atomicAdd(&sm_cnt[b], 1u | ((st[j] >> 31) << 16));
                st[j] |= (b << kPB) | (old & 0xffffu);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < kPItems; ++j) {
            if (st[j] & kLive) {
                const u32 b = cdf_pos_ties(key[j], i0 + 32 * j, n, sm.tab, tie, xmax) >> 16;
                const u32 old = atomicAdd(&sm_cnt[b], 1u | ((st[j] >> 31) << 16));
                st[j] |= (b << kPB) | (old & 0xffffu);
            }
        }
    }
    __syncthreads();
#if RTDA_EXPERIMENT_PART_CUT == 2
    { u32 x = 0;
#pragma unroll
      for (int j = 0; j < kPItems; ++j) x ^= key[j] ^ st[j];
      if (x == 0x12345678u) fb_flag[seg] = 1u;
      continue; }
#endif

    // exclusive scan of the bucket counts (consecutive buckets per thread)
    const u32 per = ((u32)geo.NBK + kPThreads - 1) / kPThreads;   // <= kBMaxBuckets / kPThreads
    u32 sum = 0;
    for (u32 k = 0; k < per; ++k) {
        const u32 b = per * tid + k;
        if (b < NB) sum += sm_cnt[b] & 0xffffu;
    }
    u32 total;
    u32 pre = block_excl_scan<kPWarps, u32>(sum, sm.scan_tmp, &total);
    for (u32 k = 0; k < per; ++k) {
        const u32 b = per * tid + k;
        if (b < NB) { sm_pre[b] = pre; pre += sm_cnt[b] & 0xffffu; }
    }
    __syncthreads();

    // one cursor claim per non-empty (tile, bucket), buckets strided over the threads.  The first
    // four claims of a thread are issued here and consumed after the exchange writes below, which
    // do not depend on them: their round trip to L2 is hidden.
    u64 *cur = cursor + (size_t)seg * geo.NBK;
    u32 off4[4], cn4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u32 b = tid + k * kPThreads;
        cn4[k] = 0;
        off4[k] = 0;
        if (b < NB) {
            const u32 cq = sm_cnt[b];
            cn4[k] = cq & 0xffffu;
            if (cn4[k]) off4[k] = (u32)atomicAdd(&cur[b], (u64)cn4[k] | ((u64)(cq >> 16) << 32));
        }
    }

#pragma unroll
    for (int j = 0; j < kPItems; ++j)
        if (st[j] & kLive)
            // payload while in shared memory: {fg, neg, bucket:P..P+10, position in the tile:0..P-1}
            sm.exch[sm_pre[(st[j] >> kPB) & 0x7ffu] + (st[j] & kPosMask)] =
                ((u64)key[j] << 32) | (st[j] & (kPayloadFg | kPayloadNeg | (0x7ffu << kPB))) | (idx0 + 32 * j);

    bool overflow = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u32 b = tid + k * kPThreads;
        if (b < NB) {
            u32 d = kNoDst;
            if (cn4[k]) {
                if (off4[k] + cn4[k] <= (u32)kBTile) d = b * (u32)kBTile + off4[k] - sm_pre[b];
                else overflow = true;
            }
            sm_cnt[b] = d;
        }
    }
    for (u32 b = tid + 4 * kPThreads; b < NB; b += kPThreads) {     // more than 1024 buckets: plain claims
        const u32 cq = sm_cnt[b], cn = cq & 0xffffu;
        u32 d = kNoDst;
        if (cn) {
            const u32 off = (u32)atomicAdd(&cur[b], (u64)cn | ((u64)(cq >> 16) << 32));
            if (off + cn <= (u32)kBTile) d = b * (u32)kBTile + off - sm_pre[b];
            else overflow = true;
        }
        sm_cnt[b] = d;
    }
    if (overflow) { fb_flag[seg] = 1u; *fb_any = 1u; }   // bucket region full: the LSD path redoes the segment
    __syncthreads();

    u64 *dst = out + (size_t)seg * geo.NBK * (size_t)kBTile;
    const u32 t_add = sm.tile_add;                   // the pixel index of the tile's first element
#pragma unroll
    for (int k = 0; k < kPItems; ++k) {
        const u32 pos = k * kPThreads + tid;
        if (pos < total) {
            const u64 v = sm.exch[pos];
            const u32 pl = (u32)v;
            const u32 d = sm_cnt[(pl >> kPB) & 0x7ffu];
            // global payload: {fg, neg, pixel index in the segment}
#if RTDA_EXPERIMENT_PART_CUT == 1
            if (d != kNoDst && v == 0x123456789ull) {
#else
            if (d != kNoDst) {
#endif
                dst[d + pos] = (v & 0xFFFFFFFFC0000000ull) | (t_add + (pl & kPosMask));
            }
        }
    }
    }
#if RTDA_SKIP_ZERO_STORES
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // the zero fills are complete before the block ends
#endif
}

// ---------------------------------------------------------------------------
// k_local: level 2.  One block per (segment, bucket), segment-major so that the 4-byte gradient
// scatter of one segment stays inside one L2-resident plane.
//
// The bucket (at most 64 KB, contiguous) is brought into shared memory by ONE bulk asynchronous copy
// (cp.async.bulk, the 1-D TMA path) signalled on an mbarrier; every later access to an element is a
// shared-memory load.
//
// The fractional part of an element's cumulative position is its place inside the bucket.  Its top
// bits select one of 32768 "fine slots", many more than there are elements (load 0.1 - 0.2), so
// most elements are alone in their slot and their local rank is simply the number of elements
// in lower slots: a prefix sum over the slot counters.  Elements sharing a slot are ordered by
// direct comparison of (key, pixel index).  The counters are W-bit fields packed in 32-bit words
// (one shared-memory atomic per element returns the arrival ticket); W = 4 first, and if a field
// overflows (16 or more elements in one slot: exact ties) the block starts over with W = 8 (16384
// slots); with 256 or more in one slot the block sorts its whole bucket with a bitonic network instead.
constexpr int kCntWords = kBTile / 2;    // packed counters, 4 fine slots per slot of capacity: 16 KB = 32768 4-bit or 16384 8-bit fine slots for 8192
constexpr int kColMax = kBTile / 8;      // dense list of the elements that share a fine slot
constexpr int kSlotBits4 = kCntWords == 4096 ? 15 : 14;    // log2(kCntWords * 8)
static_assert(kCntWords == 4096 || kCntWords == 2048, "fine-slot bits follow the counter words; the counter scan takes 128 words per warp and round");
struct LocalShared {
    u64 stage[kBTile];                   // the bucket, in the order k_part wrote it
    u32 cntw[kCntWords];
    unsigned short wpre[kCntWords];      // elements in lower words of the same warp chunk
    unsigned short where[kBTile];        // [first rank of a shared slot + ticket] -> index into stage
    u32 col[kColMax];                    // elements sharing a slot: {first rank:0..12, ticket:13..20, count:21..28} -> rank
    uint2 fg[kBTile / 32];               // .x bit r: the element of local rank 32 * i + r is foreground; .y: foreground elements of lower words
    u32 scan_tmp[kBWarps];
    u32 woff[kBWarps];
    long long red[2][kBWarps];
    u64 mbar;
    u32 ncol;
};

template <int W>
__device__ __forceinline__ u32 field_sum(u32 w) {   // sum of the W-bit fields of w
    if (W == 8) return __dp4a(w, 0x01010101u, 0u);
    return __dp4a(w & 0x0f0f0f0fu, 0x01010101u, 0u) + __dp4a((w >> 4) & 0x0f0f0f0fu, 0x01010101u, 0u);
}

// One attempt with W-bit counters.  Returns false (block-uniform) when a counter overflowed.
// rk[j] ends as the local rank of element j * kBThreads + tid.
struct LinMap { u32 lo, R, sh, M; };             // in-bucket linear map (k_plan); R = 0xFFFFFFFF: use the segment's map
template <int W>
__device__ __forceinline__ bool local_ranks(LocalShared &sm, const u32 *__restrict__ tab, const u32 *__restrict__ tie,
                                            bool has_ties, const LinMap lm, const Geom &geo, u32 cnt, u32 xmax, u32 (&rk)[kBItems]) {
    const u32 n = (u32)geo.n;
    const u32 lin_lo = lm.lo, lin_R = lm.R, lin_sh = lm.sh, lin_M = lm.M;
    const bool lin = lin_R != 0xFFFFFFFFu;           // block-uniform
    constexpr u32 FW = 32 / W;                       // fields per word
    constexpr u32 FM = (1u << W) - 1u;               // largest field value
    constexpr u32 FSHIFT = (W == 4) ? (16 - kSlotBits4) : (17 - kSlotBits4);   // 16 fractional bits -> 15 / 14 bits of fine slot (at 4096 words)
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // phase A: one atomic per element; {fine slot:0..14, foreground:15, ticket:16..23}
    bool over = false;
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        const u32 idx = j * kBThreads + tid;
        rk[j] = 0;
        if (j * kBThreads >= cnt) continue;          // block-uniform: the bucket is smaller than its capacity
        if (idx < cnt) {
            // the keys of one bucket share a few table entries: the loads broadcast out of L1
            const u64 e = sm.stage[idx];
            u32 x16;
            if (lin) {
                // any monotone map of the key orders the bucket; where the segment's map is close to one
                // straight line over the bucket's key range, that line needs no table lookups
                const u32 tau = tau_of((u32)(e >> 32));
                u32 d = (tau > lin_lo ? tau : lin_lo) - lin_lo;
                d = d < lin_R ? d : lin_R;
                x16 = __umulhi(d << lin_sh, lin_M);
            } else {
                x16 = cdf_pos(has_ties, (u32)(e >> 32), geo.seg_index((u32)e & kPayloadIdx), n, tab, tie, xmax) & 0xffffu;
            }
            const u32 f = x16 >> FSHIFT;
            const u32 sh = (f % FW) * W;
            const u32 t = (atomicAdd(&sm.cntw[f / FW], 1u << sh) >> sh) & FM;
            over |= t == FM;                         // the field was already full: this add spilled into the next one
            rk[j] = f | (((u32)e >> 31) << 15) | (t << 16);
        }
    }
    if (__syncthreads_or(over)) return false;

    // exclusive scan of the word sums inside each warp's chunk of 256 words (4 consecutive per lane)
    {
        constexpr int kChunk = kCntWords / kBWarps;
        const uint4 *c4p = reinterpret_cast<const uint4 *>(sm.cntw + warp * kChunk);
        uint2 *o2p = reinterpret_cast<uint2 *>(sm.wpre + warp * kChunk);
        u32 running = 0;
#pragma unroll
        for (int r = 0; r < kChunk / 128; ++r) {
            const uint4 c4 = c4p[r * 32 + lane];
            const u32 sx = field_sum<W>(c4.x), sy = field_sum<W>(c4.y), sz = field_sum<W>(c4.z), sw = field_sum<W>(c4.w);
            const u32 tot = sx + sy + sz + sw;
            u32 incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(RTDA_FULL_MASK, incl, o);
                if (lane >= (u32)o) incl += t;
            }
            const u32 e0 = running + incl - tot, e1 = e0 + sx, e2 = e1 + sy, e3 = e2 + sz;
            o2p[r * 32 + lane] = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
            running += __shfl_sync(RTDA_FULL_MASK, incl, 31);
        }
        if (lane == 0) sm.woff[warp] = running;
    }
    __syncthreads();
    if (warp == 0) {                                 // exclusive scan of the 16 chunk totals
        const u32 v = lane < (u32)kBWarps ? sm.woff[lane] : 0u;
        u32 incl = v;
#pragma unroll
        for (int o = 1; o < kBWarps; o <<= 1) {
            const u32 t = __shfl_up_sync(RTDA_FULL_MASK, incl, o);
            if (lane >= (u32)o) incl += t;
        }
        if (lane < (u32)kBWarps) sm.woff[lane] = incl - v;
    }
    __syncthreads();

    // phase B: rank of the elements that are alone in their slot; the others register in `where`
    // and in the dense list.  rk[j] becomes {first rank of the slot:0..12, ticket:13..20, count:21..28}
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        const u32 idx = j * kBThreads + tid;
        if (j * kBThreads >= cnt) break;
        if (idx < cnt) {
            const u32 f = rk[j] & 0x7fffu, t = rk[j] >> 16, fg = rk[j] & 0x8000u;
            const u32 wi = f / FW, sh = (f % FW) * W;
            const u32 word = sm.cntw[wi];
            const u32 c = (word >> sh) & FM;
            const u32 first = (u32)sm.wpre[wi] + sm.woff[wi / (kCntWords / kBWarps)] + field_sum<W>(word & ((1u << sh) - 1u));
            rk[j] = first | (t << 13) | (c << 21);
            if (c == 1) {
                if (fg) atomicOr(&sm.fg[first >> 5].x, 1u << (first & 31u));
            } else {
                sm.where[first + t] = (unsigned short)idx;
                const u32 q = atomicAdd(&sm.ncol, 1u);          // dense list: phase C runs without idle lanes
                if (q < (u32)kColMax) { sm.col[q] = rk[j]; rk[j] = 0x80000000u | q; }
            }
        }
    }
    __syncthreads();

    // phase C: elements sharing a slot are ordered by (key, pixel index)
    const u64 cmp_mask = ~((u64)(kPayloadFg | kPayloadNeg));
    const u32 ncol_all = sm.ncol;
    const u32 ncol = ncol_all < (u32)kColMax ? ncol_all : (u32)kColMax;
    for (u32 i = tid; i < ncol; i += kBThreads) {
        const u32 d = sm.col[i];
        const u32 first = d & 0x1fffu, c = d >> 21;
        const u64 e = sm.stage[sm.where[first + ((d >> 13) & 0xffu)]];
        const u64 ce = e & cmp_mask;
        u32 pred = 0;
        for (u32 m = 0; m < c; ++m) pred += ((sm.stage[sm.where[first + m]] & cmp_mask) < ce) ? 1u : 0u;
        const u32 r = first + pred;
        if ((u32)e & kPayloadFg) atomicOr(&sm.fg[r >> 5].x, 1u << (r & 31u));
        sm.col[i] = r;
    }
    if (ncol_all > (u32)kColMax) {                    // list full (heavy ties): the rest in place, lanes idle
#pragma unroll
        for (int j = 0; j < kBItems; ++j) {
            if (j * kBThreads >= cnt) break;
            const u32 c = (rk[j] >> 21) & 0xffu;
            if (!(rk[j] & 0x80000000u) && c > 1) {
                const u32 first = rk[j] & 0x1fffu;
                const u64 e = sm.stage[j * kBThreads + tid];
                const u64 ce = e & cmp_mask;
                u32 pred = 0;
                for (u32 m = 0; m < c; ++m) pred += ((sm.stage[sm.where[first + m]] & cmp_mask) < ce) ? 1u : 0u;
                const u32 r = first + pred;
                if ((u32)e & kPayloadFg) atomicOr(&sm.fg[r >> 5].x, 1u << (r & 31u));
                rk[j] = r;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        if (j * kBThreads >= cnt) break;
        rk[j] = (rk[j] & 0x80000000u) ? sm.col[rk[j] & 0x7fffffffu] : (rk[j] & 0x1fffu);
    }
    return true;
}

__device__ __forceinline__ void local_reset(LocalShared &sm) {
    for (u32 i = threadIdx.x; i < (u32)kCntWords; i += kBThreads) sm.cntw[i] = 0;
    if (threadIdx.x < kBTile / 32) sm.fg[threadIdx.x].x = 0;
    if (threadIdx.x == 0) sm.ncol = 0;
}

// Full in-place sort of the staged bucket by (key, pixel index): bitonic network over the next power
// of two (padding sorts last).  ~90 barrier-separated stages for 8192 elements - only for buckets
// whose fine-slot counters overflowed twice.  Leaves the bucket sorted (rank = position) and the
// foreground bitmap filled.
__device__ __noinline__ void local_sort_all(LocalShared &sm, u32 cnt) {
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u64 cmp_mask = ~((u64)(kPayloadFg | kPayloadNeg));
    u32 N = 32;
    while (N < cnt) N <<= 1;
    __syncthreads();
    for (u32 i = cnt + tid; i < N; i += kBThreads) sm.stage[i] = ~0ull;
    __syncthreads();
    for (u32 k = 2; k <= N; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = tid; t < N / 2; t += kBThreads) {
                const u32 i = 2u * t - (t & (j - 1u));
                const u32 l = i + j;
                const u64 a = sm.stage[i], b = sm.stage[l];
                const bool up = (i & k) == 0u;
                if (((a & cmp_mask) > (b & cmp_mask)) == up) { sm.stage[i] = b; sm.stage[l] = a; }
            }
            __syncthreads();
        }
    }
    for (u32 q = tid; q < N; q += kBThreads) {       // N is a multiple of 32: whole warps
        const bool fg = q < cnt && ((u32)sm.stage[q] & kPayloadFg);
        const u32 m = __ballot_sync(RTDA_FULL_MASK, fg);
        if (lane == 0) sm.fg[q >> 5].x = m;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// k_plan: between the two levels, one block per segment.  Everything a level-2 block needs to know
// before it can start is worked out once here and handed over in ONE 32-byte record per bucket, so a
// k_local block begins with a single round trip to L2 instead of a chain of dependent loads (with two
// blocks per SM and ~17 us per block, every microsecond of start-up latency is 6 % of the kernel):
//   a = {elements, elements in earlier buckets, foreground elements in earlier buckets, flags}
//   b = {tau_lo, R, shift, multiplier} of the in-bucket linear map (flag kPlanLin)
// The linear map: any monotone function of the key orders a bucket.  Inverting the segment's map at
// the bucket's two ends gives its key range [tau_lo, tau_lo + R]; where the map's slope varies by less
// than RTDA_LOCAL_LINEAR x over the intervals the bucket spans, a straight line over that range
// spreads the elements over the fine slots about as evenly as the map itself and needs no table
// lookups in k_local (their L1 latency was the largest single stall there).
constexpr u32 kPlanLive = 1u, kPlanTies = 2u, kPlanLin = 4u, kPlanTail = 8u;
// grid (S, ceil(NBK / kPlanThreads)): one thread per bucket, so the searches of a segment run side by side
__global__ void __launch_bounds__(kPlanThreads)
k_plan(Geom geo, const u32 *__restrict__ seg_fg, const u32 *__restrict__ nb, const u32 *__restrict__ table,
       const u64 *__restrict__ cursor, const u32 *__restrict__ fb_flag, uint4 *__restrict__ plan) {
    __shared__ __align__(16) u32 tab[kTabStride];
    __shared__ u64 scan_tmp[kPlanThreads / 32];
    __shared__ u64 mbar;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 seg = blockIdx.x;
    const u32 NBK = (u32)geo.NBK;
    const u32 b0 = blockIdx.y * (u32)kPlanThreads;                  // this block's buckets: [b0, b0 + kPlanThreads)
    const u32 b = b0 + tid;
    uint4 *out = plan + (size_t)seg * NBK * 2;
    const u32 nb_raw = nb[seg];
    const u32 NB = nb_raw & ~kNbTies;
    const bool has_ties = (nb_raw & kNbTies) != 0u;
    if (seg_fg[seg] == 0 || fb_flag[seg] != 0 || NB == 0) {          // nothing for level 2 in this segment
        if (b < NBK) out[2 * b] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const bool want_map = !has_ties && RTDA_LOCAL_LINEAR != 0;
    if (tid == 0 && want_map) {                                      // the map arrives while the cursors are summed
        mbar_init(&mbar, 1);
        bulk_load(tab, table + (size_t)seg * kTabStride, kTabStride * (u32)sizeof(u32), &mbar);
    }
    // elements (low word) and foreground elements (high word) of the buckets before this one
    const u64 *cur = cursor + (size_t)seg * NBK;
    const u64 c = b < NB ? cur[b] : 0ull;
    u64 before = 0;
    for (u32 i = tid; i < b0 && i < NB; i += kPlanThreads) before += cur[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(RTDA_FULL_MASK, before, o);
    u64 total_before;
    {
        if (lane == 0) scan_tmp[tid >> 5] = before;
        __syncthreads();
        u64 t = 0;
#pragma unroll
        for (int w = 0; w < kPlanThreads / 32; ++w) t += scan_tmp[w];
        total_before = t;
        __syncthreads();
    }
    const u64 run = total_before + block_excl_scan<kPlanThreads / 32, u64>(c, scan_tmp, nullptr);   // (barriers: mbar is initialised)
    if (b >= NBK) return;                                            // (after the last barrier)
    uint4 a = make_uint4(0u, 0u, 0u, 0u), l = make_uint4(0u, 0u, 0u, 0u);
    const u32 cnt = (u32)c;
    // cnt > kBTile: the region overflowed in k_part, the segment's flag is already raised
    if (b < NB && cnt != 0u && cnt <= (u32)kBTile) a = make_uint4(cnt, (u32)run, (u32)(run >> 32), kPlanLive | (has_ties ? kPlanTies : 0u));
    // Exact tail pruning: a bucket that starts after the segment's LAST foreground element (every one of the G
    // foreground elements - zero-error ones included, which are never sorted and keep this from triggering - lies in
    // earlier buckets) holds only background elements whose lovasz_grad is exactly 0: nothing of the intersection is
    // left, J = 1 - 0/U = 1 at their rank and the rank before.  Level 2 writes their zeros and skips the ranking.
    // A few per cent of the buckets on trained-like predictions, more than 90 % on diffuse ones (early training).
    if (a.w && (u32)(run >> 32) == seg_fg[seg]) { a.w |= kPlanTail; out[2 * b] = a; return; }
    if (want_map) mbar_wait(&mbar, 0);
    if (a.w && want_map) {
        // interval p with tab[p] <= X < tab[p + 1] at both ends of the bucket (tab[0] = 0, tab[kTabBins] = NB << 16 > X)
        u32 pe[2], te[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const u32 X = e ? (((b + 1u) << 16) - 1u) : (b << 16);
            u32 p = 0;
#pragma unroll
            for (u32 step = kTabBins / 2; step >= 1u; step >>= 1)
                if (p + step <= (u32)kTabBins - 1u && tab[p + step] <= X) p += step;
            const u32 tp = tab[p], dx = tab[p + 1u] - tp;           // dx >= 1
            const u64 num = (u64)(X - tp) << 19;
            // lower end: floor, never above the bucket's smallest tau; upper end: ceil, never below its largest
            u32 off = e ? (u32)((num + dx - 1u) / dx) + 1u : (u32)(num / dx);
            off = off < 0x7FFFFu ? off : 0x7FFFFu;
            pe[e] = p;
            te[e] = (p << 19) + off;
        }
        const u32 span = pe[1] - pe[0] + 1u;
        if (span <= 32u && te[1] >= te[0]) {
            u32 dmin = 0xFFFFFFFFu, dmax = 0u;
            for (u32 q = pe[0]; q <= pe[1]; ++q) {
                const u32 dd = tab[q + 1u] - tab[q];
                dmin = dd < dmin ? dd : dmin;
                dmax = dd > dmax ? dd : dmax;
            }
            if (dmin != 0u && dmax <= (u64)RTDA_LOCAL_LINEAR * dmin) {
                const u32 R = te[1] - te[0], Rn = R ? R : 1u;
                const u32 sh = (u32)__clz(Rn);
                l = make_uint4(te[0], R, sh, (u32)((65535ull << 32) / ((u64)Rn << sh)));   // (Rn << sh) in [2^31, 2^32)
                a.w |= kPlanLin;
            }
        }
    }
    out[2 * b] = a;
    out[2 * b + 1] = l;
}

__global__ void __launch_bounds__(kBThreads, RTDA_LOCAL_BLOCKS)
k_local(const u64 *__restrict__ regions, Geom geo, const u32 *__restrict__ seg_fg, const u32 *__restrict__ nb,
        const u32 *__restrict__ table, const u32 *__restrict__ ties, const uint4 *__restrict__ plan,
        unsigned long long *__restrict__ seg_sum, float *__restrict__ unit_grad, u32 *__restrict__ seg_bad,
        const float *__restrict__ seg_scale) {
    extern __shared__ __align__(128) unsigned char local_smem[];
    LocalShared &sm = *reinterpret_cast<LocalShared *>(local_smem);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 seg = blockIdx.x / (u32)geo.NBK;
    const u32 b = blockIdx.x - seg * (u32)geo.NBK;
    // k_plan's record for this bucket: one round trip decides whether the block runs and tells it how
    // (all four loads are issued before the first of them is looked at)
    const uint4 pa = plan[2 * (size_t)blockIdx.x];
    const uint4 pl = plan[2 * (size_t)blockIdx.x + 1];
    const u32 G = seg_fg[seg];
    const u32 nb_raw = nb[seg];
    const float gscale = seg_scale[seg];             // 1, or the (1 / G) / K of the final gradient (k_split)
    if (!(pa.w & kPlanLive)) return;
    const u32 cnt = pa.x;
#if RTDA_SKIP_ZERO_STORES
    if (pa.w & kPlanTail) return;                    // all zeros, already written by k_part
#endif
    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        // one bulk copy brings the bucket in (16-byte granularity: an odd count reads one slot of padding)
        bulk_load(sm.stage, regions + ((size_t)seg * geo.NBK + b) * (size_t)kBTile, ((cnt + 1u) & ~1u) * 8u, &sm.mbar);
    }
#if RTDA_LOCAL_PREFETCH > 0
    // hint: blocks start in index order; the bucket of the block RTDA_LOCAL_PREFETCH places ahead is asked
    // into L2 now, so that block's bulk copy does not wait for DRAM
    if (tid == 32) {
        const size_t pb = (size_t)blockIdx.x + RTDA_LOCAL_PREFETCH;
        if (pb < (size_t)gridDim.x) {
            const uint4 fa = plan[2 * pb];
            if ((fa.w & kPlanLive) && !(RTDA_SKIP_ZERO_STORES && (fa.w & kPlanTail)))
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(regions + pb * (size_t)kBTile), "r"(((fa.x + 1u) & ~1u) * 8u) : "memory");
        }
    }
#endif
    const bool has_ties = (pa.w & kPlanTies) != 0u;
    const u32 xmax = ((nb_raw & ~kNbTies) << 16) - 1u;
    const u32 *tab = table + (size_t)seg * kTabStride;
    const u32 *tie = ties + (size_t)seg * kTabBins;
    const u32 base_cnt = pa.y, base_fg = pa.z;
    const LinMap lm = {pl.x, (pa.w & kPlanLin) ? pl.y : 0xFFFFFFFFu, pl.z, pl.w};
    if (pa.w & kPlanTail) {
        // only background elements behind the last foreground element (k_plan): gradient exactly 0, err * 0 to the loss
#if !RTDA_SKIP_ZERO_STORES                         // (else k_part has written these zeros, and the bucket is not even staged)
        __syncthreads();                           // the barrier is initialised
        mbar_wait(&sm.mbar, 0);
        if (unit_grad) {
            float *ug0 = unit_grad + geo.grad_seg_off(seg);
            for (u32 i = tid; i < cnt; i += kBThreads) ug0[(u32)sm.stage[i] & kPayloadIdx] = 0.0f;
        }
#endif
        return;
    }
    local_reset(sm);
    __syncthreads();                               // the barrier is initialised, the counters are zero
    mbar_wait(&sm.mbar, 0);
#if RTDA_EXPERIMENT_LOCAL_CUT == 4
    if (sm.stage[tid] == 0x123456789ull) seg_bad[seg] = 1;
    return;
#endif

    u32 rk[kBItems];
    if (!local_ranks<4>(sm, tab, tie, has_ties, lm, geo, cnt, xmax, rk)) {
        local_reset(sm);
        __syncthreads();
        if (!local_ranks<8>(sm, tab, tie, has_ties, lm, geo, cnt, xmax, rk)) {
            // 256 or more elements in one fine slot (exact ties of saturated probabilities): order the
            // whole bucket by (key, pixel index) with a bitonic network in place; rank = position
            local_sort_all(sm, cnt);
#pragma unroll
            for (int j = 0; j < kBItems; ++j) rk[j] = j * kBThreads + tid;
        }
    }
#if RTDA_EXPERIMENT_LOCAL_CUT == 3
    { u32 x = 0;
#pragma unroll
      for (int j = 0; j < kBItems; ++j) x ^= rk[j];
      if (x == 0x12345678u) seg_bad[seg] = 1;
      return; }
#endif
    {
        const u32 v = tid < kBTile / 32 ? (u32)__popc(sm.fg[tid].x) : 0u;
        const u32 e = block_excl_scan<kBWarps, u32>(v, sm.scan_tmp, nullptr);
        if (tid < kBTile / 32) sm.fg[tid].y = e;
    }
    __syncthreads();

    // phase D: ranks -> lovasz_grad, dot product, gradient (lovasz_loss.py:60-69,180), as in k_rank_last
    float *ug = unit_grad ? unit_grad + geo.grad_seg_off(seg) : nullptr;
    LossAcc acc = {0, 0};
    u32 eb_max = 0;
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        const u32 idx = j * kBThreads + tid;
        if (j * kBThreads >= cnt) break;
        if (idx < cnt) {
            const u64 e = sm.stage[idx];
            const u32 pay = (u32)e;
            const u32 lr = rk[j];
            const u32 r = base_cnt + lr;                                  // 0-based rank in the sorted segment
            const uint2 fw = sm.fg[lr >> 5];
            const u32 f_ex = base_fg + fw.y + (u32)__popc(fw.x & ((1u << (lr & 31u)) - 1u));
            const u32 f_in = f_ex + (pay >> 31);
#if RTDA_EXPERIMENT_LOCAL_CUT == 2
            if (ug) ug[pay & kPayloadIdx] = __uint_as_float(f_in + r);
            continue;
#endif
            const float jac = __fsub_rn(1.0f, div_rn_counts(__uint2float_rn(G - f_in), __uint2float_rn(G + r + 1u - f_in)));
            const float jprev = __fsub_rn(1.0f, div_rn_counts(__uint2float_rn(G - f_ex), __uint2float_rn(G + r - f_ex)));
            const float gr = __fsub_rn(jac, jprev);
            const float err = __uint_as_float(~(u32)(e >> 32));
            eb_max = ~(u32)(e >> 32) > eb_max ? ~(u32)(e >> 32) : eb_max;
            acc.add(err, gr);
            // -sgn(fg - p) * g, times the segment's factor (err != 0 here)
#if RTDA_EXPERIMENT_LOCAL_CUT == 1
            if (ug && gr == 123.0f) ug[pay & kPayloadIdx] = __fmul_rn((pay & kPayloadNeg) ? gr : -gr, gscale);
#elif RTDA_SKIP_ZERO_STORES
            if (ug && gr != 0.0f) ug[pay & kPayloadIdx] = __fmul_rn((pay & kPayloadNeg) ? gr : -gr, gscale);
#else
            if (ug) ug[pay & kPayloadIdx] = __fmul_rn((pay & kPayloadNeg) ? gr : -gr, gscale);
#endif
        }
    }
    flag_nonfinite(eb_max, seg_bad + seg);         // (a NaN's bit pattern compares above every finite one)
    acc.warp_reduce();
    if (lane == 0) { sm.red[0][warp] = acc.a; sm.red[1][warp] = acc.b; }
    __syncthreads();
    if (tid < 2) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < kBWarps; ++w) t += sm.red[tid][w];
        atomicAdd(&seg_sum[(size_t)tid * geo.S + seg], (unsigned long long)t);    // integer: the order of the adds does not matter
    }
}
