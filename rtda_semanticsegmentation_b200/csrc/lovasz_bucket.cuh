// Bucket path of the Lovasz-Softmax forward: two data movements instead of four radix passes.
//
// The ranks lovasz_grad needs (lovasz_loss.py:47-70,173-176) are those of a full descending sort of
// every class's errors.  The LSD radix path (lovasz.cu) moves every (key, payload) pair through
// shared memory four times.  For the smooth key distributions softmax outputs have, a sample of the
// keys predicts the distribution well enough to finish in two movements:
//
//   k_split   one block per segment (group, class): 11-bit-prefix histogram of a 1/16 sample of the
//             segment's keys -> a monotone piecewise-linear map key -> bucket with equal expected
//             occupancy (kBMean elements per bucket).
//   k_part    level 1: every element goes to its bucket's fixed-capacity region (kBTile slots) through
//             a shared-memory exchange, one global cursor claim per (tile, bucket).  Unordered inside
//             a bucket: level 2 orders by (key, pixel index) explicitly, so no stability is needed and
//             there is no look-back chain.
//   k_local   level 2: one block per bucket.  Interpolation sort in shared memory: slot =
//             (key - min) * count / (max - min + 1), one atomic per element, a scan, one placement;
//             elements sharing a slot (about one neighbour on average) are ordered by direct
//             comparison.  rank = elements in earlier buckets + local rank, likewise the foreground
//             rank; lovasz_grad, the dot product and the gradient scatter follow as in k_rank_last.
//
// The order produced is exactly the LSD path's: ascending key (= descending error), ties by ascending
// pixel index, so both paths are bit-identical.  Anything the map cannot handle - a bucket that
// overflows its region, a slot chain longer than kChainMax (massive exact ties), more buckets than
// kBMaxBuckets - raises the segment's fallback flag on the device and the LSD kernels, which always
// run behind (persistent grids that exit at once when no flag is set), redo that segment.  No host
// synchronisation is involved.
//
// Elements with error exactly 0 are not sorted at all: they sit behind every other element, contribute
// 0 * g to the loss and sgn(0) = 0 to the gradient, and dropping a suffix changes no other rank.
#pragma once
// (included inside namespace rtda by lovasz.cu, after error_key / div_rn_counts / loss_term_fix)

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

// Monotone map key -> position in the sampled cumulative distribution, in units of buckets, 16.16 fixed
// point: the 11-bit key prefix (sign / top exponent bits are constant for errors in [0,1]) selects a
// table interval [tab[p], tab[p+1]), the low 19 key bits interpolate inside it.  Integer part = bucket,
// fractional part = position inside the bucket (level 2 turns it into a slot).
__device__ __forceinline__ u32 cdf_pos(u32 key, const u32 *tab, u32 xmax) {
    if (key < 0xC0000000u) return 0u;     // errors > 2 / NaN (garbage probabilities): first bucket, first slot
    const u32 p = (key >> 19) - 0x1800u;
    const u32 x0 = tab[p], dx = tab[p + 1] - x0;
    const u32 x = x0 + __umulhi(key << 13, dx);                 // (low19 * dx) >> 19
    return x < xmax ? x : xmax;
}

__device__ __forceinline__ u32 table_bin(u32 key) {
    return key < 0xC0000000u ? 0u : (key >> 19) - 0x1800u;
}

// ---------------------------------------------------------------------------
// k_split: sample histogram -> cumulative map of one segment.  gridDim.y blocks share a segment:
// each adds its part of the sample to the segment's global histogram, the last one to finish
// builds the map.
__global__ void __launch_bounds__(kSplitThreads)
k_split(const float *__restrict__ probas, const unsigned char *__restrict__ codes, Geom geo,
        const u32 *__restrict__ seg_fg, int force_lsd, u32 *__restrict__ ghist /*[S][kTabBins + 2], zeroed*/,
        u32 *__restrict__ table, u32 *__restrict__ nb_out, u32 *__restrict__ fb_flag, u32 *__restrict__ fb_any) {
    __shared__ u32 hist[kTabBins];
    __shared__ u32 wtmp[kSplitThreads / 32];
    __shared__ u32 s_mvalid, s_last;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 seg = blockIdx.x;
    const u32 g = seg / (u32)geo.C, c = seg - g * (u32)geo.C;
    if (seg_fg[seg] == 0) {                       // class absent from this group
        if (tid == 0 && blockIdx.y == 0) nb_out[seg] = 0;
        return;
    }
    if (force_lsd || geo.n >= (1ll << 29)) {      // bit 29 of the payload marks the first element of a slot in k_local
        if (tid == 0 && blockIdx.y == 0) { nb_out[seg] = 0; fb_flag[seg] = 1u; *fb_any = 1u; }
        return;
    }
    for (int i = tid; i < kTabBins; i += kSplitThreads) hist[i] = 0;
    if (tid == 0) s_mvalid = 0;
    __syncthreads();
    const u32 n = (u32)geo.n, HW = (u32)geo.HW;
    u32 mvalid = 0;
    {
        // groups of 4 consecutive pixels, one group of every SG (all of them for small segments)
        u32 SG = n / 16384u;
        SG = SG < 1u ? 1u : (SG > (u32)kSampleStride ? (u32)kSampleStride : SG);
        const u32 groups = (n + 4u * SG - 1u) / (4u * SG);
        const unsigned char *cg = codes + (size_t)g * n;
        for (u32 u = blockIdx.y * kSplitThreads + tid; u < groups; u += gridDim.y * kSplitThreads) {
            const u32 i0 = 4u * SG * u + 4u * ((u * 7u) % SG);
#pragma unroll
            for (u32 k = 0; k < 4; ++k) {
                const u32 i = i0 + k;
                if (i >= n) break;
                const u32 code = cg[i];
                if (code == 255u) continue;
                ++mvalid;
                const u32 bi = i / HW, hw = i - bi * HW;
                const float p = probas[(((size_t)g * geo.Bg + bi) * geo.C + c) * (size_t)HW + hw];
                u32 neg;
                const u32 key = error_key(p, true, code == c, &neg);
                if (key == 0xFFFFFFFFu) continue;          // error exactly 0: never sorted
                atomicAdd(&hist[table_bin(key)], 1u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mvalid += __shfl_xor_sync(RTDA_FULL_MASK, mvalid, o);
    if (lane == 0 && mvalid) atomicAdd(&s_mvalid, mvalid);
    __syncthreads();
    u32 *gh = ghist + (size_t)seg * (kTabBins + 2);
    if (gridDim.y > 1) {
        for (int i = tid; i < kTabBins; i += kSplitThreads)
            if (hist[i]) atomicAdd(&gh[i], hist[i]);
        if (tid == 0 && s_mvalid) atomicAdd(&gh[kTabBins], s_mvalid);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&gh[kTabBins + 1], 1u) == gridDim.y - 1u) ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int i = tid; i < kTabBins; i += kSplitThreads) hist[i] = ld_relaxed(&gh[i]);
        if (tid == 0) s_mvalid = ld_relaxed(&gh[kTabBins]);
        __syncthreads();
    }
    const u32 h0 = hist[2 * tid], h1 = hist[2 * tid + 1];
    u32 m_nz = 0;
    const u32 excl = block_excl_scan<kSplitThreads / 32, u32>(h0 + h1, wtmp, &m_nz);
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
    u32 *tab = table + (size_t)seg * (kTabBins + 1);
    const u64 scale = (u64)NB << 16;
    const u64 denom = (u64)kPrior * m_nz + kTabBins;
    const u64 c0 = (u64)kPrior * excl + 2u * tid;
    const u64 c1 = c0 + (u64)kPrior * h0 + 1u;
    tab[2 * tid] = (u32)((c0 * scale) / denom);
    tab[2 * tid + 1] = (u32)((c1 * scale) / denom);
    if (tid == 0) { tab[kTabBins] = (u32)scale; nb_out[seg] = NB; }
}

// ---------------------------------------------------------------------------
// k_part: level 1.  One block per (segment, tile of kPTile pixels), segment-major.
struct PartShared {
    u64 exch[kPTile];
    u32 tab[kTabBins + 4];
    u32 scan_tmp[kPWarps];
    u32 skip;
    u32 arrays[1];       // cnt[nbk_pad] (count | fg << 16, then the global destination), pre[nbk_pad]
};
constexpr u32 kNoDst = 0xFFFFFFFFu;
static inline size_t part_smem_bytes(int nbk) {
    const size_t pad = ((size_t)nbk + kPThreads - 1) / kPThreads * kPThreads;
    return sizeof(PartShared) + 2 * pad * sizeof(u32);
}

__global__ void __launch_bounds__(kPThreads, 4)
k_part(const float *__restrict__ probas, const unsigned char *__restrict__ codes, Geom geo,
       const u32 *__restrict__ seg_fg, const u32 *__restrict__ table, const u32 *__restrict__ nb,
       u64 *__restrict__ cursor, u64 *__restrict__ out, u32 *__restrict__ fb_flag, u32 *__restrict__ fb_any,
       float *__restrict__ unit_grad) {
    extern __shared__ __align__(16) unsigned char part_smem[];
    PartShared &sm = *reinterpret_cast<PartShared *>(part_smem);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 seg = blockIdx.x / (u32)geo.T1;
    const u32 tile = blockIdx.x - seg * (u32)geo.T1;
    // other blocks of this kernel may raise the flag while this one starts: one thread decides for the block
    if (tid == 0) sm.skip = (seg_fg[seg] == 0 || ld_relaxed(fb_flag + seg) != 0) ? 1u : 0u;
    __syncthreads();
    if (sm.skip) return;
    const u32 NB = nb[seg];
    const u32 xmax = (NB << 16) - 1u;
    u32 *sm_cnt = sm.arrays;
    u32 *sm_pre = sm.arrays + ((u32)geo.NBK + kPThreads - 1) / kPThreads * kPThreads;
    const u32 n = (u32)geo.n;
    const u32 tile_base = tile * (u32)kPTile;
    const u32 count = (n - tile_base < (u32)kPTile) ? n - tile_base : (u32)kPTile;
    const size_t seg_off = (size_t)seg * n;

    // per element: the key and one word of state {fg:31, neg:30, live:29, bucket:16..26, rank in (tile, bucket):0..12};
    // the pixel index is recomputed from the element's position in the tile
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
        {
            const u32 *tsrc = table + (size_t)seg * (kTabBins + 1);
            for (u32 k = tid; k < (u32)kTabBins + 1u; k += kPThreads) sm.tab[k] = tsrc[k];
            for (u32 b = tid; b < NB; b += kPThreads) sm_cnt[b] = 0;
        }
        float *ug = unit_grad ? unit_grad + seg_off : nullptr;
#pragma unroll
        for (int j = 0; j < kPItems; ++j) {
            const bool fg = code[j] == c;
            u32 neg;
            key[j] = error_key(p[j], true, fg, &neg);
            st[j] = (fg ? kPayloadFg : 0u) | neg;
            if (code[j] != 255u) {
                if (key[j] != 0xFFFFFFFFu) st[j] |= kLive;
                else if (ug) ug[i0 + 32 * j] = 0.0f;      // error exactly 0: gradient 0, not sorted
            }
        }
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < kPItems; ++j) {
        if (st[j] & kLive) {
            const u32 b = cdf_pos(key[j], sm.tab, xmax) >> 16;
            const u32 old = atomicAdd(&sm_cnt[b], 1u | ((st[j] & kPayloadFg) ? 0x10000u : 0u));
            st[j] |= (b << 16) | (old & 0xffffu);
        }
    }
    __syncthreads();

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
            sm.exch[sm_pre[(st[j] >> 16) & 0x7ffu] + (st[j] & 0x1fffu)] =
                ((u64)key[j] << 32) | (st[j] & (kPayloadFg | kPayloadNeg)) | (i0 + 32 * j);

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
#pragma unroll
    for (int k = 0; k < kPItems; ++k) {
        const u32 pos = k * kPThreads + tid;
        if (pos < total) {
            const u64 v = sm.exch[pos];
            const u32 d = sm_cnt[cdf_pos((u32)(v >> 32), sm.tab, xmax) >> 16];
            if (d != kNoDst) dst[d + pos] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// k_local: level 2.  One block per (segment, bucket), segment-major so that the 4-byte gradient
// scatter of one segment stays inside one L2-resident plane.
struct LocalShared {
    u64 buf[kBTile];
    u32 slot[kBTile];
    u32 tab[kTabBins + 4];
    u32 fgbits[kBTile / 32];
    u32 fgpre[kBTile / 32];
    u32 woff[kBWarps];
    u32 scan_tmp[kBWarps];
    u64 red64[kBWarps];
    u32 redchain[kBWarps];
    long long red[kBWarps];
    u32 chain;
    u32 skip;
    u64 base;
};
constexpr u32 kSlotFirst = 1u << 29;    // payload bit: first element (in arrival order) of its slot

__global__ void __launch_bounds__(kBThreads, 2)
k_local(const u64 *__restrict__ regions, Geom geo, const u32 *__restrict__ seg_fg, const u32 *__restrict__ nb,
        const u32 *__restrict__ table, const u64 *__restrict__ cursor, u32 *__restrict__ fb_flag,
        u32 *__restrict__ fb_any, long long *__restrict__ partial_b, float *__restrict__ unit_grad) {
    extern __shared__ __align__(16) unsigned char local_smem[];
    LocalShared &sm = *reinterpret_cast<LocalShared *>(local_smem);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 seg = blockIdx.x / (u32)geo.NBK;
    const u32 b = blockIdx.x - seg * (u32)geo.NBK;
    const u32 G = seg_fg[seg];
    if (tid == 0) sm.skip = (G == 0 || ld_relaxed(fb_flag + seg) != 0 || b >= nb[seg]) ? 1u : 0u;
    __syncthreads();
    if (sm.skip) return;
    const u64 *cur = cursor + (size_t)seg * geo.NBK;
    const u32 cnt = (u32)cur[b];
    if (cnt > (u32)kBTile) return;                 // overflowed in k_part (flag already raised)
    if (cnt == 0) {
        if (tid == 0) partial_b[(size_t)seg * geo.NBK + b] = 0;
        return;
    }
    const u64 *src = regions + ((size_t)seg * geo.NBK + b) * (size_t)kBTile;
    const u32 xmax = (nb[seg] << 16) - 1u;
    // the keys of phase A are requested first: the setup below runs while they travel
    u32 keyA[kBItems];
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        const u32 idx = j * kBThreads + tid;
        keyA[j] = 0;
        if (idx < cnt) keyA[j] = reinterpret_cast<const u32 *>(src + idx)[1];
    }

    // elements (low word) and foreground elements (high word) of the earlier buckets
    u64 acc64 = 0;
    for (u32 i = tid; i < b; i += kBThreads) acc64 += cur[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc64 += __shfl_xor_sync(RTDA_FULL_MASK, acc64, o);
    if (lane == 0) sm.red64[warp] = acc64;
    {
        const u32 *tsrc = table + (size_t)seg * (kTabBins + 1);
        for (u32 k = tid; k < (u32)kTabBins + 1u; k += kBThreads) sm.tab[k] = tsrc[k];
    }
    for (u32 i = tid; i < ((cnt + 127u) & ~127u); i += kBThreads) sm.slot[i] = 0;   // whole uint4 groups of the scan
    if (tid < kBTile / 32) sm.fgbits[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        u64 t = 0;
#pragma unroll
        for (int w = 0; w < kBWarps; ++w) t += sm.red64[w];
        sm.base = t;
    }

    // phase A: slot = fractional position inside the bucket x count; the ticket is the element's
    // arrival order inside its slot.  {slot:0..12, ticket:16..28} per element stay in registers.
    u32 stt[kBItems];
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        const u32 idx = j * kBThreads + tid;
        stt[j] = 0;
        if (idx < cnt) {
            const u32 x = cdf_pos(keyA[j], sm.tab, xmax);
            const u32 s = ((x & 0xffffu) * cnt) >> 16;
            stt[j] = s | (atomicAdd(&sm.slot[s], 1u) << 16);
        }
    }
    __syncthreads();
    const u32 base_cnt = (u32)sm.base, base_fg = (u32)(sm.base >> 32);

    // exclusive scan of the slot counts: warp w owns slots [512 w, 512 w + 512), 4 consecutive per lane
    {
        constexpr int kChunk = kBTile / kBWarps;
        uint4 *sl4 = reinterpret_cast<uint4 *>(sm.slot + warp * kChunk);
        u32 running = 0, maxc = 0;
#pragma unroll
        for (int r = 0; r < kChunk / 128; ++r) {
            if (warp * kChunk + r * 128 >= cnt) break;           // slots >= cnt are never used (warp-uniform)
            uint4 c4 = sl4[r * 32 + lane];
            const u32 m01 = c4.x > c4.y ? c4.x : c4.y, m23 = c4.z > c4.w ? c4.z : c4.w;
            maxc = maxc > m01 ? maxc : m01;
            maxc = maxc > m23 ? maxc : m23;
            const u32 tot = c4.x + c4.y + c4.z + c4.w;
            u32 incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(RTDA_FULL_MASK, incl, o);
                if (lane >= (u32)o) incl += t;
            }
            u32 e = running + incl - tot;
            uint4 o4;
            o4.x = e; e += c4.x;
            o4.y = e; e += c4.y;
            o4.z = e; e += c4.z;
            o4.w = e;
            sl4[r * 32 + lane] = o4;
            running += __shfl_sync(RTDA_FULL_MASK, incl, 31);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u32 t = __shfl_xor_sync(RTDA_FULL_MASK, maxc, o);
            maxc = t > maxc ? t : maxc;
        }
        if (lane == 0) { sm.woff[warp] = running; sm.redchain[warp] = maxc; }
    }
    __syncthreads();
    if (tid == 0) {
        u32 run = 0, mc = 0;
#pragma unroll
        for (int w = 0; w < kBWarps; ++w) {
            const u32 t = sm.woff[w];
            sm.woff[w] = run;
            run += t;
            mc = sm.redchain[w] > mc ? sm.redchain[w] : mc;
        }
        sm.chain = mc;
    }
    __syncthreads();
    if (sm.chain > (u32)kChainMax) {               // massive ties: quadratic chain work, let the LSD path do it
        if (tid == 0) { fb_flag[seg] = 1u; *fb_any = 1u; }
        return;
    }

    // phase B: placement in slot order; foreground bitmap by position; bit 29 marks a slot's first element
#pragma unroll
    for (int j = 0; j < kBItems; ++j) {
        const u32 idx = j * kBThreads + tid;
        if (idx < cnt) {
            const u64 v = src[idx];
            const u32 s = stt[j] & 0xffffu, t = stt[j] >> 16;
            const u32 pos = sm.slot[s] + sm.woff[s / (kBTile / kBWarps)] + t;
            sm.buf[pos] = t ? v : (v | kSlotFirst);
            if ((u32)v & kPayloadFg) atomicOr(&sm.fgbits[pos >> 5], 1u << (pos & 31u));
        }
    }
    __syncthreads();
    {
        const u32 v = tid < kBTile / 32 ? (u32)__popc(sm.fgbits[tid]) : 0u;
        const u32 e = block_excl_scan<kBWarps, u32>(v, sm.scan_tmp, nullptr);
        if (tid < kBTile / 32) sm.fgpre[tid] = e;
    }
    __syncthreads();

    // phase C: order inside the slot by (key, pixel index), ranks, lovasz_grad, dot product, gradient
    float *ug = unit_grad ? unit_grad + (size_t)seg * (size_t)geo.n : nullptr;
    const u64 cmp_mask = ~((u64)(kPayloadFg | kPayloadNeg | kSlotFirst));
    long long acc = 0;
#pragma unroll 2
    for (int k = 0; k < kBItems; ++k) {
        const u32 q = k * kBThreads + tid;
        if (q >= cnt) break;
        const u64 e = sm.buf[q];
        const u32 key = (u32)(e >> 32), pay = (u32)e;
        const u64 ce = e & cmp_mask;
        u32 pred = 0, fpred = 0;
        u32 cs = q;
        if (!(pay & kSlotFirst)) {
            for (;;) {                              // elements of the same slot to the left, down to its first
                const u64 x = sm.buf[--cs];
                if ((x & cmp_mask) < ce) { ++pred; fpred += ((u32)x >> 31); }
                if ((u32)x & kSlotFirst) break;
            }
        }
        for (u32 j = q + 1; j < cnt; ++j) {        // and to the right, up to the next slot's first
            const u64 x = sm.buf[j];
            if ((u32)x & kSlotFirst) break;
            if ((x & cmp_mask) < ce) { ++pred; fpred += ((u32)x >> 31); }
        }
        const u32 fgq = pay >> 31;
        const u32 r = base_cnt + cs + pred;                                  // 0-based rank in the sorted segment
        const u32 f_ex = base_fg + sm.fgpre[cs >> 5] + (u32)__popc(sm.fgbits[cs >> 5] & ((1u << (cs & 31u)) - 1u)) + fpred;
        const u32 f_in = f_ex + fgq;
        // lovasz_loss.py:60-69, as in k_rank_last
        const float jac = __fsub_rn(1.0f, div_rn_counts(__uint2float_rn(G - f_in), __uint2float_rn(G + r + 1u - f_in)));
        const float jprev = __fsub_rn(1.0f, div_rn_counts(__uint2float_rn(G - f_ex), __uint2float_rn(G + r - f_ex)));
        const float gr = __fsub_rn(jac, jprev);
        const float err = __uint_as_float(~key);
        acc += loss_term_fix(err, gr);
        if (ug) ug[pay & (kSlotFirst - 1u)] = (pay & kPayloadNeg) ? gr : -gr;       // err != 0 here
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(RTDA_FULL_MASK, acc, o);
    if (lane == 0) sm.red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < kBWarps; ++w) t += sm.red[w];
        partial_b[(size_t)seg * geo.NBK + b] = t;
    }
}
