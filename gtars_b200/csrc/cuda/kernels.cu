// kernels.cu — sm_100a kernels for count / any / find / tokenize.
//
// Nothing here is a dense contraction, so tensor cores are deliberately unused: the work is HBM-bound integer
// search + stream compaction.  What matters is (1) query rows staged by TMA bulk copies, coalesced everywhere else,
// (2) O(1) resolution through L2-resident tables (window records for find / tokenize, rank LUTs for counting) instead
// of 17–26-level bisections, (3) a single pass over the queries for enumeration: count → scan → two-level decoupled
// look-back across tiles → emit, so no per-query count/offset array ever round-trips through HBM, and (4) a
// persistent grid sized to the SM count (a lean record-only kernel with the full kernel as its on-device fallback).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace gtgpu {

// ================================================================================================================
// device helpers
// ================================================================================================================
__device__ __forceinline__ uint32_t lut_lower_bound(const uint32_t* __restrict__ arr, const uint32_t* __restrict__ lut,
                                                    uint32_t nb, uint32_t n, uint32_t shift, uint32_t key) {
    uint32_t b = key >> shift;
    if (b >= nb) return n;  // key is beyond the last bin: every element is smaller
    uint32_t lo = __ldg(lut + b), hi = __ldg(lut + b + 1);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(arr + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// lower_bound over a chromosome-level sorted array through its rank LUT: one 8-byte load in the common case.
__device__ __forceinline__ uint32_t rank_lower_bound(const uint32_t* __restrict__ arr, const unsigned long long* __restrict__ rl,
                                                     uint32_t nb, uint32_t n, uint32_t shift, uint32_t ninl, uint32_t key) {
    const uint32_t b = key >> shift;
    if (b >= nb) return n;
    const unsigned long long w = __ldg(rl + b);
    const uint32_t base = (uint32_t)w, cnt = (uint32_t)(w >> 32) & 7u;
    const uint32_t r = key & ((1u << shift) - 1);
    if (cnt != 7u) {
        uint32_t below = 0;
        unsigned long long offs = w >> 35;
        for (uint32_t j = 0; j < cnt; ++j) {
            below += (uint32_t)(offs & ((1u << shift) - 1)) < r;
            offs >>= shift;
        }
        return base + below;
    }
    uint32_t lo = base, hi = (uint32_t)__ldg(rl + b + 1);  // crowded bin: bisect inside it
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(arr + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Candidate range of one segment for query [s,e): positions [lo, ub) (global), where ub = #starts < e and
// lo = first position whose running max end exceeds s.  Everything outside cannot overlap; inside, an interval
// overlaps iff its own end > s (always true when the segment's ends are monotone).
__device__ __forceinline__ void seg_range(const IndexView& ix, const SegMeta& m, uint32_t s, uint32_t e, uint32_t& lo,
                                          uint32_t& ub) {
    uint32_t u = lut_lower_bound(ix.starts + m.off, ix.lut + m.lut_s, m.nb_s, m.len, ix.shift, e);
    uint32_t l = u;
    if (u != 0 && s != 0xFFFFFFFFu) {
        l = lut_lower_bound(ix.pmax + m.off, ix.lut + m.lut_p, m.nb_p, m.len, ix.shift, s + 1);
        if (l > u) l = u;
    }
    lo = m.off + l;
    ub = m.off + u;
}

__device__ __forceinline__ bool is_hit(const IndexView& ix, uint32_t i, uint32_t s, uint32_t e, int32_t min_bp,
                                       bool mono) {
    // multi_chrom_overlapper.rs:489-494: the bp filter only applies when min_overlap > 1.
    if (min_bp <= 1) return mono || __ldg(ix.ends + i) > s;
    uint32_t ie = __ldg(ix.ends + i);
    if (ie <= s) return false;
    uint32_t is = __ldg(ix.starts + i);
    int64_t bp = (int64_t)min(e, ie) - (int64_t)max(s, is);
    return bp >= (int64_t)min_bp;
}

__device__ __forceinline__ uint32_t count_range(const IndexView& ix, uint32_t lo, uint32_t ub, uint32_t s, uint32_t e,
                                                int32_t min_bp, bool mono) {
    if (mono && min_bp <= 1) return ub - lo;
    uint32_t c = 0;
    for (uint32_t i = lo; i < ub; ++i) c += is_hit(ix, i, s, e, min_bp, mono);
    return c;
}

__device__ __forceinline__ SegMeta load_seg(const IndexView& ix, uint32_t sidx) {
    const uint4* p = reinterpret_cast<const uint4*>(ix.segs + sidx);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    SegMeta m;
    m.off = a.x; m.len = a.y; m.lut_s = a.z; m.nb_s = a.w;
    m.lut_p = b.x; m.nb_p = b.y; m.mono = b.z; m.pad = b.w;
    return m;
}

__device__ __forceinline__ ChromMeta load_chrom(const IndexView& ix, uint32_t c) {
    const uint4* p = reinterpret_cast<const uint4*>(ix.chroms + c);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    ChromMeta m;
    m.seg_begin = a.x; m.seg_end = a.y; m.off = a.z; m.len = a.w;
    m.lut_cs = b.x; m.nb_cs = b.y; m.lut_ce = b.z; m.nb_ce = b.w;
    return m;
}

// Total hits of a query over all segments of its chromosome (walk path).
__device__ __forceinline__ uint32_t count_query_walk(const IndexView& ix, uint32_t c, uint32_t s, uint32_t e,
                                                     int32_t min_bp) {
    if (c >= ix.n_chroms) return 0;
    const uint2 sr = __ldg(reinterpret_cast<const uint2*>(ix.chroms + c));
    uint32_t total = 0;
    for (uint32_t si = sr.x; si < sr.y; ++si) {
        SegMeta m = load_seg(ix, si);
        uint32_t lo, ub;
        seg_range(ix, m, s, e, lo, ub);
        total += count_range(ix, lo, ub, s, e, min_bp, m.mono != 0);
    }
    return total;
}

// ================================================================================================================
// count / any / Bits::count
// ================================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) count_kernel(IndexView ix, uint64_t n, const uint32_t* __restrict__ chr,
                                                    const uint32_t* __restrict__ start,
                                                    const uint32_t* __restrict__ end, int32_t min_bp,
                                                    void* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t c = __ldg(chr + i), s = __ldg(start + i), e = __ldg(end + i);
        if (MODE == COUNT_BITS_RAW_U64) {
            // bits.rs:337-344 with its wrapping arithmetic: (#starts < e) - (#ends < s+1), s+1 wrapping in u32.
            uint64_t r = 0;
            if (c < ix.n_chroms) {
                ChromMeta cm = load_chrom(ix, c);
                uint64_t last = rank_lower_bound(ix.cs_starts + cm.off, ix.rank_lut + cm.lut_cs, cm.nb_cs, cm.len, ix.rank_shift, ix.rank_inline, e);
                uint64_t first = rank_lower_bound(ix.cs_ends + cm.off, ix.rank_lut + cm.lut_ce, cm.nb_ce, cm.len, ix.rank_shift, ix.rank_inline, s + 1u);
                r = last - first;
            }
            reinterpret_cast<uint64_t*>(out)[i] = r;
        } else {
            uint32_t cnt = 0;
            if (c < ix.n_chroms) {
                if (min_bp <= 1 && ix.proper && s < e) {
                    // The BITS identity equals the enumerated count whenever every interval has start <= end and
                    // the query has start < end (no interval can both end <= s and start >= e).
                    ChromMeta cm = load_chrom(ix, c);
                    uint32_t last = rank_lower_bound(ix.cs_starts + cm.off, ix.rank_lut + cm.lut_cs, cm.nb_cs, cm.len, ix.rank_shift, ix.rank_inline, e);
                    uint32_t first = rank_lower_bound(ix.cs_ends + cm.off, ix.rank_lut + cm.lut_ce, cm.nb_ce, cm.len, ix.rank_shift, ix.rank_inline, s + 1u);
                    cnt = last - first;
                } else {
                    cnt = count_query_walk(ix, c, s, e, min_bp);
                }
            }
            if (MODE == COUNT_ANY_U8) reinterpret_cast<uint8_t*>(out)[i] = cnt != 0;
            else reinterpret_cast<uint32_t*>(out)[i] = cnt;
        }
    }
}

// Four consecutive queries per thread (128-bit query loads): the rank-LUT words of all four are requested before any
// of them is consumed, so a thread keeps eight table loads in flight instead of two one after the other.  Same
// results as count_kernel; needs 16-byte aligned query arrays (the launcher checks).
__device__ __forceinline__ unsigned long long rank_fetch(const unsigned long long* __restrict__ rl, uint32_t nb, uint32_t shift,
                                                         uint32_t key) {
    const uint32_t b = key >> shift;
    return __ldg(rl + min(b, nb));  // word nb is the sentinel (base = n, no inline entries)
}

__device__ __forceinline__ uint32_t rank_resolve(unsigned long long w, const uint32_t* __restrict__ arr,
                                                 const unsigned long long* __restrict__ rl, uint32_t nb, uint32_t n, uint32_t shift,
                                                 uint32_t key) {
    const uint32_t b = key >> shift;
    if (b >= nb) return n;
    const uint32_t base = (uint32_t)w, cnt = (uint32_t)(w >> 32) & 7u;
    const uint32_t r = key & ((1u << shift) - 1);
    if (cnt != 7u) {
        uint32_t below = 0;
        unsigned long long offs = w >> 35;
        for (uint32_t j = 0; j < cnt; ++j) {
            below += (uint32_t)(offs & ((1u << shift) - 1)) < r;
            offs >>= shift;
        }
        return base + below;
    }
    uint32_t lo = base, hi = (uint32_t)__ldg(rl + b + 1);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(arr + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <int MODE>
__device__ __forceinline__ void count_store(void* __restrict__ out, uint64_t i, uint64_t r) {
    if (MODE == COUNT_BITS_RAW_U64) reinterpret_cast<uint64_t*>(out)[i] = r;
    else if (MODE == COUNT_ANY_U8) reinterpret_cast<uint8_t*>(out)[i] = r != 0;
    else reinterpret_cast<uint32_t*>(out)[i] = (uint32_t)r;
}

template <int MODE>
__global__ void __launch_bounds__(256) count_kernel_x4(IndexView ix, uint64_t n, const uint32_t* __restrict__ chr,
                                                       const uint32_t* __restrict__ start,
                                                       const uint32_t* __restrict__ end, int32_t min_bp,
                                                       void* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n4 = n / 4;
    const bool identity_ok = MODE == COUNT_BITS_RAW_U64 || (min_bp <= 1 && ix.proper);
    // queries and results stream (evict-first): the L2 is for the LUT words.  The next step's queries are requested
    // before this step's LUT words are consumed, so the DRAM latency of the stream hides behind the table look-ups.
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 nc = make_uint4(0, 0, 0, 0), ns = nc, ne = nc;
    if (g < n4) {
        nc = __ldcs(reinterpret_cast<const uint4*>(chr) + g);
        ns = __ldcs(reinterpret_cast<const uint4*>(start) + g);
        ne = __ldcs(reinterpret_cast<const uint4*>(end) + g);
    }
    for (; g < n4; g += stride) {
        const uint4 c4 = nc, s4 = ns, e4 = ne;
        if (g + stride < n4) {
            nc = __ldcs(reinterpret_cast<const uint4*>(chr) + g + stride);
            ns = __ldcs(reinterpret_cast<const uint4*>(start) + g + stride);
            ne = __ldcs(reinterpret_cast<const uint4*>(end) + g + stride);
        }
        const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w}, e[4] = {e4.x, e4.y, e4.z, e4.w};
        bool fast[4];
        uint4 off_len[4];  // ChromMeta words 0-3: seg_begin, seg_end, off (.z), len (.w)
        uint4 luts[4];     // ChromMeta words 4-7: lut_cs, nb_cs, lut_ce, nb_ce
        unsigned long long wl[4], wf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            fast[k] = c[k] < ix.n_chroms && identity_ok && (MODE == COUNT_BITS_RAW_U64 || s[k] < e[k]);
            wl[k] = wf[k] = 0;
            if (fast[k]) {
                const uint4* p = reinterpret_cast<const uint4*>(ix.chroms + c[k]);
                off_len[k] = __ldg(p);
                luts[k] = __ldg(p + 1);
                wl[k] = rank_fetch(ix.rank_lut + luts[k].x, luts[k].y, ix.rank_shift, e[k]);
                wf[k] = rank_fetch(ix.rank_lut + luts[k].z, luts[k].w, ix.rank_shift, s[k] + 1u);
            }
        }
        uint64_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (fast[k]) {
                const uint64_t last = rank_resolve(wl[k], ix.cs_starts + off_len[k].z, ix.rank_lut + luts[k].x, luts[k].y,
                                                   off_len[k].w, ix.rank_shift, e[k]);
                const uint64_t first = rank_resolve(wf[k], ix.cs_ends + off_len[k].z, ix.rank_lut + luts[k].z, luts[k].w,
                                                    off_len[k].w, ix.rank_shift, s[k] + 1u);
                r[k] = MODE == COUNT_BITS_RAW_U64 ? last - first : (uint64_t)((uint32_t)last - (uint32_t)first);
            } else {
                r[k] = (MODE == COUNT_BITS_RAW_U64 || c[k] >= ix.n_chroms) ? 0 : count_query_walk(ix, c[k], s[k], e[k], min_bp);
            }
        }
        if (MODE == COUNT_U32) {
            __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4((uint32_t)r[0], (uint32_t)r[1], (uint32_t)r[2], (uint32_t)r[3]));
        } else if (MODE == COUNT_ANY_U8) {
            __stcs(reinterpret_cast<uchar4*>(out) + g, make_uchar4(r[0] != 0, r[1] != 0, r[2] != 0, r[3] != 0));
        } else {
            __stcs(reinterpret_cast<ulonglong2*>(out) + 2 * g, make_ulonglong2(r[0], r[1]));
            __stcs(reinterpret_cast<ulonglong2*>(out) + 2 * g + 1, make_ulonglong2(r[2], r[3]));
        }
    }
    // the last n % 4 queries: one thread each, same arithmetic as count_kernel
    if (blockIdx.x == 0 && threadIdx.x < (uint32_t)(n - n4 * 4)) {
        const uint64_t i = n4 * 4 + threadIdx.x;
        const uint32_t c = __ldg(chr + i), s = __ldg(start + i), e = __ldg(end + i);
        uint64_t r = 0;
        if (c < ix.n_chroms) {
            if (identity_ok && (MODE == COUNT_BITS_RAW_U64 || s < e)) {
                ChromMeta cm = load_chrom(ix, c);
                uint64_t last = rank_lower_bound(ix.cs_starts + cm.off, ix.rank_lut + cm.lut_cs, cm.nb_cs, cm.len, ix.rank_shift, ix.rank_inline, e);
                uint64_t first = rank_lower_bound(ix.cs_ends + cm.off, ix.rank_lut + cm.lut_ce, cm.nb_ce, cm.len, ix.rank_shift, ix.rank_inline, s + 1u);
                r = MODE == COUNT_BITS_RAW_U64 ? last - first : (uint64_t)((uint32_t)last - (uint32_t)first);
            } else if (MODE != COUNT_BITS_RAW_U64) {
                r = count_query_walk(ix, c, s, e, min_bp);
            }
        }
        count_store<MODE>(out, i, r);
    }
}

// ---- TMA (bulk async copy) staging of a tile's query rows into shared memory ------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy) : "memory");
}
// L2 residency: the bin table is re-read by every tile (evict_last), queries and ids are touched once (evict_first).
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---- bucketed counting: databases whose rank LUTs exceed the L2 ----------------------------------------------
// A search through the rank LUT is one 8-byte load, but with unsorted queries over a 50 M-interval database (C3:
// 0.8-1.5 GB of LUT words) every one of them is a random DRAM sector.  The bucketed pass makes the counting kernel sweep
// the LUT once, slice by slice, each slice small enough to stay in the L2 while its bucket is being resolved — WITHOUT
// a global partition of the queries (round 1: histogram + scan + partition + count + gather, 8.0 GB of traffic for
// 2.0 GB of algorithmic bytes):
//   1. count_stage_kernel: a tile of CP_TILE queries is reduced to the two linearised search keys of each query
//      (IndexView::rank_lin: 8 bytes instead of 12, no chromosome look-up later), grouped by bucket in shared memory
//      (bucket = position of the starts-LUT word >> bucket_shift) and written to the TILE'S OWN region of the staging
//      array, runs padded to whole 64-byte lines.  Per tile and bucket one word (run start | run length) goes into a
//      bucket-major table; per query a 16-bit staged position.  No histogram pass, no scan, no slots.
//   2. count_runs_kernel: one warp per (bucket, tile) run, runs taken bucket-major by a grid that is resident all at
//      once: the two LUT words of a query are its only gathers.  Results are written over the same staged positions.
//   3. count_unstage_kernel: a tile's results come back as ONE contiguous read and are put back into query order through
//      shared memory (position 0xFFFF = a query the identity does not cover: walked here, from the original arrays).
// Results do not depend on the staging: out[] is identical to the direct pass.
constexpr int CP_MAX_BUCKETS = 256;
constexpr int CP_THREADS = 1024;
constexpr int CP_ITEMS = 4;  // consecutive queries per thread: 128-bit query loads
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;
constexpr int CP_PAD = 8;               // runs start on multiples of 8 staged entries (64 B of keys, 32 B of u32 results)
constexpr uint32_t CP_WALK = 0xFFFFu;   // staged position of a query that bypasses the identity
constexpr int CP_CHROM_CACHE = 256;     // rank_lin entries kept in shared memory
static_assert(CP_TILE + CP_PAD * CP_MAX_BUCKETS < (int)CP_WALK, "staged positions are 16 bits");
static_assert(CP_TILE <= 0xFFFF, "run lengths are 16 bits");

// RAW: Bits::count's wrapping identity (every query with a known chromosome takes it); else count / any (start < end).
// The three query rows of the block's NEXT tile are fetched by bulk copies (TMA, evict-first) into shared memory while
// the current tile is grouped and written out (the phases of a tile are separated by block barriers, and the two
// 1024-thread blocks of an SM need no registers for rows that are still on their way).
template <bool RAW>
__global__ void __launch_bounds__(CP_THREADS, 2)
count_stage_kernel(IndexView ix, uint64_t n, uint32_t n_tiles, uint32_t bucket_shift, uint32_t nb, uint32_t cap,
                   const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                   uint2* __restrict__ keys, uint16_t* __restrict__ pos, uint32_t* __restrict__ runs, uint32_t* __restrict__ tile_used,
                   uint32_t* __restrict__ run_counter) {
    __shared__ uint32_t s_cnt[CP_MAX_BUCKETS];   // tile histogram = rank dispenser
    __shared__ uint32_t s_base[CP_MAX_BUCKETS];  // first staged position of each bucket's run
    __shared__ uint32_t s_used;
    __shared__ uint4 s_lin[CP_CHROM_CACHE];
    __shared__ __align__(8) uint64_t s_bar;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    uint32_t* s_in = reinterpret_cast<uint32_t*>(s_dyn);                  // 3 x CP_TILE: chr, start, end of a whole tile
    uint2* s_keys = reinterpret_cast<uint2*>(s_dyn + 3 * CP_TILE * 4);    // cap entries: the tile grouped by bucket
    const uint32_t tid = threadIdx.x, rs = ix.rank_shift;
    // a whole tile comes through the bulk copies (the launcher checked the 16-byte alignment of the three arrays)
    auto whole = [&](uint32_t tile) { return tile < n_tiles && ((uint64_t)tile + 1) * CP_TILE <= n; };
    auto fetch = [&](uint32_t tile) {  // one thread
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the buffer are done
        mbar_expect_tx(&s_bar, 3 * CP_TILE * 4);
        const uint64_t q0 = (uint64_t)tile * CP_TILE, pol = policy_evict_first();
        bulk_g2s(s_in, chr + q0, CP_TILE * 4, &s_bar, pol);
        bulk_g2s(s_in + CP_TILE, start + q0, CP_TILE * 4, &s_bar, pol);
        bulk_g2s(s_in + 2 * CP_TILE, end + q0, CP_TILE * 4, &s_bar, pol);
    };
    for (uint32_t i = tid; i < (uint32_t)CP_CHROM_CACHE && i < ix.n_chroms; i += CP_THREADS) s_lin[i] = __ldg(ix.rank_lin + i);
    if (tid == 0) {
        if (blockIdx.x == 0) *run_counter = 0;  // the chunk dispenser of count_runs_kernel
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (whole(blockIdx.x)) fetch(blockIdx.x);
    }
    __syncthreads();
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t t0 = (uint64_t)tile * CP_TILE;
        const uint32_t tn = (uint32_t)min((uint64_t)CP_TILE, n - t0), j0 = tid * CP_ITEMS;
        if (tid < CP_MAX_BUCKETS) s_cnt[tid] = 0;
        uint32_t qc[CP_ITEMS], qs[CP_ITEMS], qe[CP_ITEMS];
        if (tn == (uint32_t)CP_TILE) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
            const uint4 c4 = reinterpret_cast<const uint4*>(s_in)[tid];
            const uint4 s4 = reinterpret_cast<const uint4*>(s_in + CP_TILE)[tid];
            const uint4 e4 = reinterpret_cast<const uint4*>(s_in + 2 * CP_TILE)[tid];
            qc[0] = c4.x, qc[1] = c4.y, qc[2] = c4.z, qc[3] = c4.w;
            qs[0] = s4.x, qs[1] = s4.y, qs[2] = s4.z, qs[3] = s4.w;
            qe[0] = e4.x, qe[1] = e4.y, qe[2] = e4.z, qe[3] = e4.w;
        } else {
#pragma unroll
            for (int k = 0; k < CP_ITEMS; ++k) {
                const bool ok = j0 + k < tn;
                qc[k] = ok ? __ldcs(chr + t0 + j0 + k) : 0xFFFFFFFFu;  // past the end: an unknown chromosome (never stored)
                qs[k] = ok ? __ldcs(start + t0 + j0 + k) : 0;
                qe[k] = ok ? __ldcs(end + t0 + j0 + k) : 0;
            }
        }
        __syncthreads();  // s_cnt zeroed, the staged rows are in registers
        if (tid == 0 && whole(tile + gridDim.x)) fetch(tile + gridDim.x);
        uint32_t br[CP_ITEMS];  // bucket << 16 | rank inside the tile's bucket; all ones = not staged
#pragma unroll
        for (int k = 0; k < CP_ITEMS; ++k) {
            br[k] = 0xFFFFFFFFu;
            if (qc[k] < ix.n_chroms && (RAW || qs[k] < qe[k])) {
                const uint4 L = qc[k] < (uint32_t)CP_CHROM_CACHE ? s_lin[qc[k]] : __ldg(ix.rank_lin + qc[k]);
                qe[k] = L.x + min(qe[k], L.y);        // key of the search over the starts: #starts < end
                qs[k] = L.z + min(qs[k] + 1u, L.w);   // key of the search over the ends: #ends < start + 1 (wrapping, bits.rs:337-344)
                br[k] = min((qe[k] >> rs) >> bucket_shift, nb - 1);
            }
        }
        // ranks from one shared-memory atomic per query.  Letting the lanes that drew the same bucket share an atomic was
        // slower both ways it was tried (0.46 ms per 1e8 queries as is; __match_any_sync 1.43 ms, one ballot per bucket
        // bit 0.71 ms).
#pragma unroll
        for (int k = 0; k < CP_ITEMS; ++k)
            if (br[k] != 0xFFFFFFFFu) br[k] = br[k] << 16 | atomicAdd(&s_cnt[br[k]], 1u);
        __syncthreads();
        if (tid < 32) {  // exclusive scan of the padded bucket sizes (eight per lane)
            uint32_t v[8], s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = (s_cnt[8 * tid + k] + (CP_PAD - 1)) & ~(uint32_t)(CP_PAD - 1);
                s += v[k];
            }
            uint32_t incl = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (tid >= (uint32_t)d) incl += t;
            }
            uint32_t run = incl - s;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s_base[8 * tid + k] = run;
                run += v[k];
            }
            if (tid == 31) {
                s_used = incl;
                tile_used[tile] = incl;
            }
        }
        __syncthreads();
        if (tid < nb) runs[(uint64_t)tid * n_tiles + tile] = s_base[tid] << 16 | s_cnt[tid];
        uint32_t p[CP_ITEMS];
#pragma unroll
        for (int k = 0; k < CP_ITEMS; ++k) {
            p[k] = CP_WALK;
            if (br[k] != 0xFFFFFFFFu) {
                p[k] = s_base[br[k] >> 16] + (br[k] & 0xFFFFu);
                s_keys[p[k]] = make_uint2(qe[k], qs[k]);
            }
        }
        if (tn == (uint32_t)CP_TILE) {
            __stcs(reinterpret_cast<uint2*>(pos + t0) + tid, make_uint2(p[0] | p[1] << 16, p[2] | p[3] << 16));
        } else {
#pragma unroll
            for (int k = 0; k < CP_ITEMS; ++k)
                if (j0 + k < tn) pos[t0 + j0 + k] = (uint16_t)p[k];
        }
        __syncthreads();
        // the padding between the runs travels along (whole lines; nobody reads it)
        const uint32_t used2 = s_used / 2;
        uint4* dst = reinterpret_cast<uint4*>(keys + (uint64_t)tile * cap);
        for (uint32_t i = tid; i < used2; i += CP_THREADS) __stcs(dst + i, reinterpret_cast<const uint4*>(s_keys)[i]);
        __syncthreads();
    }
}

// The chromosome a LUT word belongs to, for the rare bin with more entries than a word holds inline (the bisection
// needs the chromosome's slice of the sorted array): the last chromosome whose LUT starts at or before the word.
__device__ __noinline__ uint32_t rank_lin_bisect(const ChromMeta* __restrict__ chroms, uint32_t n_chroms, const uint32_t* __restrict__ sorted,
                                                 uint32_t lut_word /* 2 = lut_cs, 3 = lut_ce */, uint32_t mask, uint32_t word, uint32_t lo,
                                                 uint32_t hi, uint32_t r) {
    uint32_t a = 0, b = n_chroms;  // invariant: lut(a) <= word
    while (b - a > 1) {
        const uint32_t m = (a + b) >> 1;
        if (__ldg(reinterpret_cast<const uint2*>(chroms + m) + lut_word).x <= word) a = m;
        else b = m;
    }
    const uint32_t* arr = sorted + __ldg(reinterpret_cast<const uint4*>(chroms + a)).z;  // ChromMeta::off
    while (lo < hi) {  // every entry in [lo, hi) lies in the key's bin: the in-bin offsets decide
        const uint32_t mid = (lo + hi) >> 1;
        if ((__ldg(arr + mid) & mask) < r) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t rank_lin_resolve(const IndexView& ix, bool ends, unsigned long long w, uint32_t lin) {
    const uint32_t shift = ix.rank_shift, mask = (1u << shift) - 1, base = (uint32_t)w, cnt = (uint32_t)(w >> 32) & 7u, r = lin & mask;
    if (cnt == 7u) {
        const uint32_t word = (lin >> shift) + (ends ? ix.rank_ends_off : 0u);
        return rank_lin_bisect(ix.chroms, ix.n_chroms, ends ? ix.cs_ends : ix.cs_starts, ends ? 3 : 2, mask, word, base,
                               (uint32_t)__ldg(ix.rank_lut + word + 1), r);
    }
    uint32_t below = 0;
    unsigned long long offs = w >> 35;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {  // at most four inline entries (index.cu); branch-free
        below += (j < cnt) & ((uint32_t)(offs & mask) < r);
        offs >>= shift;
    }
    return base + below;
}

// Warps draw chunks of CR_RUNS consecutive (bucket, tile) runs from one counter, in order: the whole grid works at one
// front that moves through the buckets, so the LUT slices in use stay in the L2 (a static round-robin let the warps drift
// several buckets apart: 5.5 GB of DRAM reads instead of 1.7).  A chunk's runs are walked as one flat sequence, two
// staged queries per lane and round; their four LUT words are in flight together, and the next round's keys, the next
// chunk's run words and the chunk after that's number are requested before the current ones are consumed.
constexpr uint32_t CR_RUNS = 4;
#ifndef GT_CR_MINBLOCKS
#define GT_CR_MINBLOCKS 4
#endif
template <typename R>
__global__ void __launch_bounds__(256, GT_CR_MINBLOCKS) count_runs_kernel(IndexView ix, uint32_t n_runs, uint32_t n_tiles, uint32_t cap,
                                                         const uint32_t* __restrict__ runs, const uint2* __restrict__ keys,
                                                         R* __restrict__ res, uint32_t* __restrict__ counter) {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31, rs = ix.rank_shift;
    const unsigned long long* lut_s = ix.rank_lut;
    const unsigned long long* lut_e = ix.rank_lut + ix.rank_ends_off;
    const uint32_t n_chunks = (n_runs + CR_RUNS - 1) / CR_RUNS;
    auto run_word = [&](uint32_t chunk) {
        const uint32_t r = chunk * CR_RUNS + lane;
        return lane < CR_RUNS && chunk < n_chunks && r < n_runs ? __ldg(runs + r) : 0u;
    };
    uint32_t cur = 0;
    if (lane == 0) cur = atomicAdd(counter, 2u);
    cur = __shfl_sync(FULL, cur, 0);
    uint32_t nxt = cur + 1, w_cur = run_word(cur);
    while (cur < n_chunks) {
        const uint32_t w_nxt = run_word(nxt);
        uint32_t after = 0;
        if (lane == 0) after = atomicAdd(counter, 1u);
        // lanes 0 .. CR_RUNS-1 hold one run each: length, and where staged entry f of the flat sequence lives (adj + f)
        const uint32_t cnt = w_cur & 0xFFFFu;
        uint32_t incl = cnt;
#pragma unroll
        for (uint32_t d = 1; d < CR_RUNS; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t adj = ((cur * CR_RUNS + lane) % n_tiles) * cap + (w_cur >> 16) - (incl - cnt);
        const uint32_t p0 = __shfl_sync(FULL, incl, 0), p1 = __shfl_sync(FULL, incl, 1), p2 = __shfl_sync(FULL, incl, 2);
        const uint32_t total = __shfl_sync(FULL, incl, 3);
        static_assert(CR_RUNS == 4, "the run of a flat position is found with three comparisons");
        auto where = [&](uint32_t f) { return __shfl_sync(FULL, adj, (f >= p0) + (f >= p1) + (f >= p2)) + f; };
        uint32_t f = lane, ia = where(f), ib = where(f + 32);
        uint2 ka = f < total ? __ldcs(keys + ia) : make_uint2(0, 0);  // (0, 0): the first word of either block
        uint2 kb = f + 32 < total ? __ldcs(keys + ib) : make_uint2(0, 0);
        for (uint32_t f0 = 0; f0 < total; f0 += 64) {
            const unsigned long long wa_s = __ldg(lut_s + (ka.x >> rs)), wa_e = __ldg(lut_e + (ka.y >> rs));
            const unsigned long long wb_s = __ldg(lut_s + (kb.x >> rs)), wb_e = __ldg(lut_e + (kb.y >> rs));
            const uint32_t f2 = f + 64, ia2 = where(f2), ib2 = where(f2 + 32);
            const uint2 ka2 = f2 < total ? __ldcs(keys + ia2) : make_uint2(0, 0);
            const uint2 kb2 = f2 + 32 < total ? __ldcs(keys + ib2) : make_uint2(0, 0);
            if (f < total) {
                const uint32_t last = rank_lin_resolve(ix, false, wa_s, ka.x), first = rank_lin_resolve(ix, true, wa_e, ka.y);
                __stcs(res + ia, sizeof(R) == 8 ? (R)((uint64_t)last - (uint64_t)first) : (R)(last - first));
            }
            if (f + 32 < total) {
                const uint32_t last = rank_lin_resolve(ix, false, wb_s, kb.x), first = rank_lin_resolve(ix, true, wb_e, kb.y);
                __stcs(res + ib, sizeof(R) == 8 ? (R)((uint64_t)last - (uint64_t)first) : (R)(last - first));
            }
            f = f2, ia = ia2, ib = ib2, ka = ka2, kb = kb2;
        }
        cur = nxt;
        w_cur = w_nxt;
        nxt = __shfl_sync(FULL, after, 0);
    }
}

// A tile's results (one contiguous block) and its staged positions arrive by bulk copies, double-buffered: the next
// tile's are in flight while this one is put back into query order.
template <int MODE, typename R>
__global__ void __launch_bounds__(CP_THREADS, 2)
count_unstage_kernel(IndexView ix, uint64_t n, uint32_t n_tiles, uint32_t cap, const uint32_t* __restrict__ chr,
                     const uint32_t* __restrict__ start, const uint32_t* __restrict__ end, int32_t min_bp,
                     const uint16_t* __restrict__ pos, const R* __restrict__ res, const uint32_t* __restrict__ tile_used,
                     void* __restrict__ out, bool out_aligned) {
    __shared__ __align__(8) uint64_t s_bar[2];
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const uint32_t buf_bytes = cap * (uint32_t)sizeof(R) + CP_TILE * 2;  // cap results, then CP_TILE positions
    const uint32_t tid = threadIdx.x;
    auto fetch = [&](uint32_t tile, uint32_t b) {  // one thread
        const uint32_t res_bytes = __ldg(tile_used + tile) * (uint32_t)sizeof(R);  // a multiple of 32
        const bool whole = ((uint64_t)tile + 1) * CP_TILE <= n;
        unsigned char* buf = s_dyn + b * buf_bytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the buffer are done
        if (res_bytes + (whole ? CP_TILE * 2 : 0) == 0) {
            mbar_arrive(&s_bar[b]);
            return;
        }
        mbar_expect_tx(&s_bar[b], res_bytes + (whole ? CP_TILE * 2 : 0));
        const uint64_t pol = policy_evict_first();
        if (res_bytes) bulk_g2s(buf, res + (uint64_t)tile * cap, res_bytes, &s_bar[b], pol);
        if (whole) bulk_g2s(buf + cap * sizeof(R), pos + (uint64_t)tile * CP_TILE, CP_TILE * 2, &s_bar[b], pol);
    };
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x < n_tiles) fetch(blockIdx.x, 0);
    }
    __syncthreads();
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t b = it & 1u;
        if (tid == 0 && tile + gridDim.x < n_tiles) fetch(tile + gridDim.x, b ^ 1u);  // that buffer was consumed a round ago
        const uint64_t t0 = (uint64_t)tile * CP_TILE;
        const uint32_t tn = (uint32_t)min((uint64_t)CP_TILE, n - t0), j0 = tid * CP_ITEMS;
        const R* s_res = reinterpret_cast<const R*>(s_dyn + b * buf_bytes);
        mbar_wait(&s_bar[b], (it >> 1) & 1u);
        uint32_t p[CP_ITEMS];
        const bool whole = tn == (uint32_t)CP_TILE;
        if (whole) {
            const uint2 pp = reinterpret_cast<const uint2*>(s_dyn + b * buf_bytes + cap * sizeof(R))[tid];
            p[0] = pp.x & 0xFFFFu, p[1] = pp.x >> 16, p[2] = pp.y & 0xFFFFu, p[3] = pp.y >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < CP_ITEMS; ++k) p[k] = j0 + k < tn ? pos[t0 + j0 + k] : CP_WALK;
        }
        uint64_t v[CP_ITEMS];
#pragma unroll
        for (int k = 0; k < CP_ITEMS; ++k) {
            if (p[k] != CP_WALK) {
                v[k] = s_res[p[k]];
            } else {
                v[k] = 0;
                if (MODE != COUNT_BITS_RAW_U64 && j0 + k < tn) {
                    const uint32_t c = __ldg(chr + t0 + j0 + k);
                    if (c < ix.n_chroms) v[k] = count_query_walk(ix, c, __ldg(start + t0 + j0 + k), __ldg(end + t0 + j0 + k), min_bp);
                }
            }
        }
        if (whole && out_aligned) {
            const uint64_t g = t0 / CP_ITEMS + tid;
            if (MODE == COUNT_U32) {
                __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4((uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]));
            } else if (MODE == COUNT_ANY_U8) {
                __stcs(reinterpret_cast<uchar4*>(out) + g, make_uchar4(v[0] != 0, v[1] != 0, v[2] != 0, v[3] != 0));
            } else {
                __stcs(reinterpret_cast<ulonglong2*>(out) + 2 * g, make_ulonglong2(v[0], v[1]));
                __stcs(reinterpret_cast<ulonglong2*>(out) + 2 * g + 1, make_ulonglong2(v[2], v[3]));
            }
        } else {
#pragma unroll
            for (int k = 0; k < CP_ITEMS; ++k)
                if (j0 + k < tn) count_store<MODE>(out, t0 + j0 + k, v[k]);
        }
        __syncthreads();  // buffer b may be refilled
    }
}

// A grid that is resident all at once (blocks per SM from the occupancy calculator): every block then walks the
// queries in step with the others, which is what keeps one bucket's LUT slice in the L2 while it is being used.
template <typename K>
static int resident_grid(const gtgpu_ctx* ctx, K kernel, int threads, uint64_t blocks_needed, size_t dyn_smem = 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, dyn_smem) != cudaSuccess || occ < 1) occ = 1;
    return (int)std::min<uint64_t>(std::max<uint64_t>(blocks_needed, 1), (uint64_t)ctx->sm_count * occ);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// One counting pass: four queries per thread when the arrays allow 128-bit accesses, else one.
template <int MODE>
static void launch_count_mode(const gtgpu_ctx* ctx, const IndexView& v, uint64_t n, const uint32_t* c, const uint32_t* s,
                              const uint32_t* e, int32_t min_bp, void* out, bool resident) {
    cudaStream_t st = ctx->stream;
    if (aligned16(c) && aligned16(s) && aligned16(e) && aligned16(out)) {
        const uint64_t blocks = (n / 4 + 255) / 256;
        const int grid = resident ? resident_grid(ctx, count_kernel_x4<MODE>, 256, blocks)
                                  : (int)std::min<uint64_t>(std::max<uint64_t>(blocks, 1), (uint64_t)ctx->sm_count * 32);
        count_kernel_x4<MODE><<<grid, 256, 0, st>>>(v, n, c, s, e, min_bp, out);
    } else {
        const uint64_t blocks = (n + 255) / 256;
        const int grid = resident ? resident_grid(ctx, count_kernel<MODE>, 256, blocks)
                                  : (int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 32);
        count_kernel<MODE><<<grid, 256, 0, st>>>(v, n, c, s, e, min_bp, out);
    }
}

// Whether this launch goes through the bucketed pass: the identity path only, an index whose linearised search keys fit
// 32 bits, LUTs well beyond what stays in the L2, enough queries to pay for the extra passes.  GTGPU_COUNT_PARTITION=0 / 1
// forces it off / on (tests, measurements).
static bool count_wants_partition(const gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start,
                                  const uint32_t* d_end, int32_t min_overlap) {
    if (min_overlap > 1 || !ix->view.proper || n >= (1ull << 31) || ix->rank_lut_len == 0 || !ix->view.rank_lin) return false;  // (staged positions are 32-bit)
    if (!aligned16(d_chr) || !aligned16(d_start) || !aligned16(d_end)) return false;
    if (const char* env = getenv("GTGPU_COUNT_PARTITION")) return env[0] == '1';
    return ix->rank_lut_len * 8 > (64ull << 20) && n >= (4ull << 20);
}

template <int MODE, typename R>
static int32_t launch_count_bucketed_mode(gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start,
                                          const uint32_t* d_end, int32_t min_overlap, void* d_out) {
    gtgpu_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    // Buckets: (a) a slice of at most 16 MB of LUT words (an eighth of the L2; half of it starts, half ends); (b) the
    // chunks the resident warps of count_runs_kernel hold at any time (two each) should span at most about 64 MB of LUT —
    // with few queries per call (a rank's block of a query-sharded batch) that span, not the slice, is what has to fit the
    // L2.  Runs shorter than 32 queries cost more than they save: at most 128 buckets (measured at 1.25e7, 5e7 and 1e8
    // queries against 0.78 GB of LUT: 128 / 64 / 64 buckets are the fastest).
    const uint64_t lut_bytes = ix->rank_lut_len * 8;
    const uint64_t held = (uint64_t)ctx->sm_count * GT_CR_MINBLOCKS * 8 * 2 * CR_RUNS * CP_TILE;  // queries held x buckets
    uint32_t nb = 32;
    while (nb < 128 && (lut_bytes / nb > (16ull << 20) || (held / nb) * (double)lut_bytes / (double)n > (double)(64ull << 20))) nb *= 2;
    if (const char* env = getenv("GTGPU_COUNT_BUCKETS")) nb = (uint32_t)std::min(CP_MAX_BUCKETS, std::max(2, atoi(env)));
    uint32_t bucket_shift = 0;
    while ((ix->view.rank_ends_off >> bucket_shift) >= nb) ++bucket_shift;
    const uint32_t n_tiles = (uint32_t)((n + CP_TILE - 1) / CP_TILE), n_runs = n_tiles * nb;
    const uint32_t cap = CP_TILE + CP_PAD * nb;  // staged entries per tile: every run padded to a multiple of CP_PAD
    uint2* keys;
    R* res;
    uint16_t* pos;
    uint32_t* runs;
    GT_TRY(ctx->scratch_get(SC_CNT_KEYS, (size_t)n_tiles * cap * 8, (void**)&keys));
    GT_TRY(ctx->scratch_get(SC_CNT_RES, (size_t)n_tiles * cap * sizeof(R), (void**)&res));
    GT_TRY(ctx->scratch_get(SC_CNT_POS, (size_t)n_tiles * CP_TILE * 2, (void**)&pos));
    GT_TRY(ctx->scratch_get(SC_CNT_RUNS, ((size_t)n_runs + n_tiles + 1) * 4, (void**)&runs));
    uint32_t* tile_used = runs + n_runs;
    uint32_t* run_counter = tile_used + n_tiles;  // zeroed by the staging kernel
    constexpr bool RAW = MODE == COUNT_BITS_RAW_U64;
    const size_t stage_smem = (size_t)cap * 8 + 3 * CP_TILE * 4, unstage_smem = 2 * ((size_t)cap * sizeof(R) + CP_TILE * 2);
    GT_CUDA(cudaFuncSetAttribute(count_stage_kernel<RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_smem));
    GT_CUDA(cudaFuncSetAttribute(count_unstage_kernel<MODE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)unstage_smem));
    const int stage_grid = resident_grid(ctx, count_stage_kernel<RAW>, CP_THREADS, n_tiles, stage_smem);
    const int runs_grid = resident_grid(ctx, count_runs_kernel<R>, 256, ((uint64_t)n_runs + 7) / 8);
    const int unstage_grid = resident_grid(ctx, count_unstage_kernel<MODE, R>, CP_THREADS, n_tiles, unstage_smem);
    ctx->time_begin();
    count_stage_kernel<RAW><<<stage_grid, CP_THREADS, stage_smem, st>>>(ix->view, n, n_tiles, bucket_shift, nb, cap, d_chr, d_start,
                                                                         d_end, keys, pos, runs, tile_used, run_counter);
    count_runs_kernel<R><<<runs_grid, 256, 0, st>>>(ix->view, n_runs, n_tiles, cap, runs, keys, res, run_counter);
    count_unstage_kernel<MODE, R><<<unstage_grid, CP_THREADS, unstage_smem, st>>>(ix->view, n, n_tiles, cap, d_chr, d_start, d_end,
                                                                                  min_overlap, pos, res, tile_used, d_out,
                                                                                  aligned16(d_out));
    ctx->time_end();
    ctx->launches += 3;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

static int32_t launch_count_partitioned(gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start,
                                        const uint32_t* d_end, int32_t min_overlap, int mode, void* d_out) {
    switch (mode) {
        case COUNT_U32: return launch_count_bucketed_mode<COUNT_U32, uint32_t>(ix, n, d_chr, d_start, d_end, min_overlap, d_out);
        case COUNT_ANY_U8: return launch_count_bucketed_mode<COUNT_ANY_U8, uint32_t>(ix, n, d_chr, d_start, d_end, min_overlap, d_out);
        default: return launch_count_bucketed_mode<COUNT_BITS_RAW_U64, unsigned long long>(ix, n, d_chr, d_start, d_end, min_overlap, d_out);
    }
}

// Gives back what a find on this ctx set aside in the L2 for its window table: the stream's access-policy window, the
// persisting lines and the set-aside itself (the next find re-establishes all three).
void release_l2_window(gtgpu_ctx* ctx) {
    if (!ctx->l2_window_owner) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    cudaGetLastError();
    ctx->l2_window_owner = nullptr;
}

int32_t launch_count(gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start,
                     const uint32_t* d_end, int32_t min_overlap, int mode, void* d_out) {
    if (n == 0) return GTGPU_OK;
    gtgpu_ctx* ctx = ix->ctx;
    if (const char* env = getenv("GTGPU_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(env));  // tuning knob
    if (count_wants_partition(ix, n, d_chr, d_start, d_end, min_overlap)) {
        // The bucketed pass lives on its LUT slices staying in the L2: give back what an earlier find on this ctx set aside
        // for its window table (persisting lines + the stream's access-policy window; the next find re-establishes both).
        release_l2_window(ctx);
        return launch_count_partitioned(ix, n, d_chr, d_start, d_end, min_overlap, mode, d_out);
    }
    ctx->time_begin();
    switch (mode) {
        case COUNT_U32: launch_count_mode<COUNT_U32>(ctx, ix->view, n, d_chr, d_start, d_end, min_overlap, d_out, false); break;
        case COUNT_ANY_U8: launch_count_mode<COUNT_ANY_U8>(ctx, ix->view, n, d_chr, d_start, d_end, min_overlap, d_out, false); break;
        default: launch_count_mode<COUNT_BITS_RAW_U64>(ctx, ix->view, n, d_chr, d_start, d_end, min_overlap, d_out, false); break;
    }
    ctx->time_end();
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// ================================================================================================================
// fused find: count → scan → decoupled look-back across tiles → emit, one pass over the queries
// ================================================================================================================
// A tile is FUSED_BLOCK x ROWS consecutive queries.  Inside a tile each WARP owns 32 x ROWS consecutive queries in
// a striped layout (lane l holds queries l, l+32, …), so every query load and every id store of a row is one fully
// coalesced 128-byte access.  Tiles are handed out in order by an atomic counter (persistent grid), which is what
// makes the look-back deadlock-free.  The look-back is two-level (supertiles of FUSED_SUPER tiles, see below): at
// >1e11 queries/s about 600 tiles are in flight, far more than any single window of predecessors can follow.
#define ST_FLAG_AGG (1ull << 62)
#define ST_FLAG_PREFIX (2ull << 62)
#define ST_MASK ((1ull << 62) - 1)

// Two-level look-back: FUSED_SUPER consecutive tiles form a supertile with ONE accumulator word
//   bit 63 = the exclusive prefix of the supertile's first tile has been added (by that tile, once)
//   bits 56-62 = number of tiles whose aggregate has been added; FUSED_SUPER (64) of them set exactly bit 62
//   bits 0-55 = sum of what has been added
// so (word >> 62) reads 1 = complete aggregate, 3 = complete inclusive prefix, 0 / 2 = not usable yet.  A tile then
// needs the statuses of the < 64 tiles before it inside its supertile (two warps, one word per lane) and the words of
// the preceding supertiles (a third warp) — all prefetched one resolve phase earlier, one L2 round trip, one block
// barrier — instead of walking ~600 in-flight tiles 256 at a time (1.8 dependent round trips, measured).
#define FUSED_SUPER 64u
#define SUPER_MASK ((1ull << 56) - 1)

struct FusedWorkspace {
    uint64_t* status;      // [n_tiles] flag<<62 | value
    uint64_t* super;       // [n_tiles / FUSED_SUPER + 1] supertile accumulators
    uint32_t* tile_file;   // [n_tiles] 0 = no file boundary in this tile, else 0xFFFFFFFF - first file index
    uint32_t* counter;     // dynamic tile counter
    uint32_t* lean_flag;   // raised by the lean kernel when the full kernel has to redo the launch
};

static inline uint64_t n_tiles_for(uint64_t n) { return (n + FUSED_TILE - 1) / FUSED_TILE; }

// Two sets of status / accumulator words and counters (the lean launch and its fallback each need zeroed ones), one
// tile_file array (read-only for both), one lean flag.
static size_t fused_chain_bytes(uint64_t t) { return (size_t)(t * 8 + (t / FUSED_SUPER + 1) * 8 + 16); }

size_t fused_workspace_bytes(uint64_t n) {
    uint64_t t = n_tiles_for(n);
    return 2 * fused_chain_bytes(t) + (size_t)(((t * 4 + 15) / 16) * 16 + 64);
}

static FusedWorkspace carve(void* ws, uint64_t n, int set) {
    uint64_t t = n_tiles_for(n);
    char* base = reinterpret_cast<char*>(ws);
    FusedWorkspace w;
    w.status = reinterpret_cast<uint64_t*>(base + (size_t)set * fused_chain_bytes(t));
    w.super = w.status + t;
    w.counter = reinterpret_cast<uint32_t*>(w.super + t / FUSED_SUPER + 1);
    w.tile_file = reinterpret_cast<uint32_t*>(base + 2 * fused_chain_bytes(t));
    w.lean_flag = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(w.tile_file) + ((t * 4 + 15) / 16) * 16);
    return w;
}

__global__ void mark_file_tiles_kernel(uint64_t n_files, const uint64_t* __restrict__ file_offsets, uint64_t n_tiles,
                                       uint32_t* __restrict__ tile_file) {
    uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f > n_files) return;
    uint64_t t = file_offsets[f] / FUSED_TILE;
    if (t >= n_tiles) t = n_tiles - 1;
    atomicMax(tile_file + t, 0xFFFFFFFFu - (uint32_t)f);
}

__global__ void fill_from_base_kernel(uint64_t count, uint64_t* __restrict__ out, const uint64_t* __restrict__ base) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = base ? *base : 0;
}

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }


__device__ __forceinline__ bool hit_bp(uint32_t cs, uint32_t ce, uint32_t s, uint32_t e, int32_t min_bp) {
    bool h = cs < e && ce > s;
    if (min_bp > 1) h = h && ((int64_t)min(e, ce) - (int64_t)max(s, cs) >= (int64_t)min_bp);
    return h;
}

// Resolves one query through the window table when the direct path of the fused kernel could not: pooled candidate
// lists, and queries that span several two-bin windows (up to BT_MAX_WINDOWS).  Windows are visited in emission order
// (ascending for Bits, descending for AIList); an interval that touches several windows is taken from the lowest one
// (it starts before every later window).  Returns false when the table cannot serve the query (no table on this
// chromosome, degenerate or very wide query, an overflow window, several AIList components across windows): the caller
// then falls back to the LUT + walk path.  With EMIT the hits' vals are written from `pos` on; `count` gets the hits.
template <bool EMIT>
__device__ __forceinline__ bool window_walk(const IndexView& ix, uint32_t c, uint32_t s, uint32_t e, int32_t min_bp,
                                            uint32_t* __restrict__ out_ids, uint64_t pos, uint64_t capacity, uint32_t& count) {
    count = 0;
    if (c >= ix.n_chroms || s >= e) return false;
    const uint2 cb = __ldg(reinterpret_cast<const uint2*>(ix.chrom_bt + c));
    if (cb.y == BT_GENERIC_CHROM) return false;
    const uint32_t nb = cb.y & BT_NBINS_MASK, sh = ix.bt_shift;
    const uint32_t b1 = s >> sh, b2 = (e - 1) >> sh;
    const uint32_t nwin = (b2 - b1) / 2 + 1;
    if (nwin > BT_MAX_WINDOWS || (nwin > 1 && (cb.y & BT_MULTI_COMP))) return false;
    const bool desc = ix.descending != 0;
    // pass 1: no window may be an overflow window
    for (uint32_t j = 0; j < nwin; ++j) {
        const uint32_t b = b1 + 2 * j;
        if (b < nb && __ldg(ix.bt_lut + cb.x + b) == BT_OVERFLOW) return false;
    }
    for (uint32_t jj = 0; jj < nwin; ++jj) {
        const uint32_t j = desc ? nwin - 1 - jj : jj;
        const uint32_t b = b1 + 2 * j;
        if (b >= nb) continue;
        const uint32_t w = __ldg(ix.bt_lut + cb.x + b);
        if (w == 0) continue;
        const uint32_t wstart = j == 0 ? 0u : b << sh;  // later windows skip what an earlier window already listed
        const bool pool = (w & BT_POOL_FLAG) != 0;
        const uint32_t n = pool ? (w & BT_POOL_MAX) : (w & 3u);
        const uint32_t first = pool ? (w & ~BT_POOL_FLAG) >> BT_POOL_SHIFT : w >> 2;
        for (uint32_t t = 0; t < n; ++t) {
            // pool lists are stored in emission order; direct runs ascend, so AIList reads them backwards
            const uint32_t idx = pool ? __ldg(ix.bt_pool + first + t) : (desc ? first + n - 1 - t : first + t);
            const uint4 E = ldg128(ix.bt_ent + idx);
            if (E.x < wstart || !hit_bp(E.x, E.y, s, e, min_bp)) continue;
            if (EMIT && pos + count < capacity) out_ids[pos + count] = E.z;
            ++count;
        }
    }
    return true;
}

// Emits every hit of one query through the generic LUT + walk path, in reference order; returns the count.
__device__ __noinline__ uint32_t emit_query_walk(const IndexView& ix, uint32_t c, uint32_t s, uint32_t e, int32_t min_bp,
                                                 uint32_t* __restrict__ out_ids, uint64_t pos, uint64_t capacity) {
    if (c >= ix.n_chroms) return 0;
    uint32_t written = 0;
    if (window_walk<true>(ix, c, s, e, min_bp, out_ids, pos, capacity, written)) return written;
    written = 0;
    const uint2 sr = __ldg(reinterpret_cast<const uint2*>(ix.chroms + c));
    for (uint32_t si = sr.x; si < sr.y; ++si) {
        SegMeta m = load_seg(ix, si);
        uint32_t l, u;
        seg_range(ix, m, s, e, l, u);
        const bool mono = m.mono != 0;
        if (!ix.descending) {
            for (uint32_t i = l; i < u; ++i)
                if (is_hit(ix, i, s, e, min_bp, mono)) {
                    if (pos + written < capacity) out_ids[pos + written] = __ldg(ix.vals + i);
                    ++written;
                }
        } else {
            for (uint32_t i = u; i > l; --i)
                if (is_hit(ix, i - 1, s, e, min_bp, mono)) {
                    if (pos + written < capacity) out_ids[pos + written] = __ldg(ix.vals + i - 1);
                    ++written;
                }
        }
    }
    return written;
}

// Queries the direct window path cannot serve: pool list when the window has one, else the LUT + walk path.
__device__ __noinline__ uint32_t count_query_walk_noinline(const IndexView& ix, uint32_t c, uint32_t s, uint32_t e,
                                                           int32_t min_bp) {
    uint32_t cnt;
    if (window_walk<false>(ix, c, s, e, min_bp, nullptr, 0, 0, cnt)) return cnt;
    return count_query_walk(ix, c, s, e, min_bp);
}

// Count for a query that left the record path, plus its first two hits' vals (emission order) in ab[] when the window
// table could serve it (multi-window queries, pool lists): bit 31 of the result says so.
__device__ __noinline__ uint32_t count_query_first2(const IndexView& ix, uint32_t c, uint32_t s, uint32_t e, int32_t min_bp,
                                                    uint32_t* ab) {
    uint32_t cnt;
    ab[0] = ab[1] = 0;
    if (window_walk<true>(ix, c, s, e, min_bp, ab, 0, 2, cnt)) return cnt | 0x80000000u;
    return count_query_walk(ix, c, s, e, min_bp);
}

__device__ __forceinline__ uint64_t ld_status(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_status(uint64_t* p, uint64_t v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t ldg32_keep(const void* p, uint64_t policy) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint4 ldg128_keep(const void* p, uint64_t policy) {
    uint4 v;
    // one sector per lane and no reuse inside an SM: do not hold an L1 line for it (measured 7.44 -> 7.33 ms)
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}

#ifdef GT_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16 * 8];  // [warp][phase]
#define PHASE_MARK(i)                                                            \
    do {                                                                         \
        long long _now = clock64();                                              \
        if (lane == 0) s_acc[warp][i] += (unsigned long long)(_now - _t);        \
        _t = _now;                                                               \
    } while (0)
#define PHASE_START() long long _t = clock64()
#else
#define PHASE_MARK(i)
#define PHASE_START()
#endif

// Warp sum of 64-bit values below 2^63: three hardware warp reductions (REDUX.SUM) over 21-bit limbs instead of ten
// dependent shuffles — this sits on the critical path between the two barriers of a step.
__device__ __forceinline__ uint64_t warp_sum_u63(uint64_t v) {
    const uint32_t a = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(v & 0x1FFFFFu));
    const uint32_t b = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)((v >> 21) & 0x1FFFFFu));
    const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(v >> 42));
    return (uint64_t)a + ((uint64_t)b << 21) + ((uint64_t)c << 42);
}

// One step of a warp inclusive scan: v + (the value d lanes below, when there is such a lane).  shfl.up's own predicate
// says whether the source lane exists, so a step is two instructions instead of shuffle + compare + select + add.
__device__ __forceinline__ uint32_t shfl_up_add(uint32_t v, int d) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "shfl.sync.up.b32 %0|p, %1, %2, 0, 0xffffffff;\n"  // a lane without a source gets its own value back
        "@p add.u32 %0, %0, %1;\n"
        "}\n" : "=&r"(r) : "r"(v), "r"(d));
    return r;
}

// Per-thread result of resolving ROWS queries of one tile, kept in registers while the NEXT tile is resolved
// (software pipeline, lag 1): by the time a tile's look-back runs, its predecessors published their aggregates a
// whole resolve phase ago, so the look-back is one L2 round trip and (almost) never spins.
template <int ROWS>
struct TileState {
    uint32_t tile;
    uint32_t v0[ROWS];   // first hit's val per query (second one, when there is one, waits in s_aux)
    uint32_t cntpack;    // 2 bits per row: min(count, 3)
    uint32_t offpack;    // 8 bits per row: offset inside the warp slice, valid when the warp had no generic query
    uint32_t slow;       // bit k: query k goes through the generic LUT + walk path (count and emit); bit 31: warp-wide
    uint32_t warp_excl, tile_agg;
};

template <bool FILTER>
__device__ __forceinline__ bool cand_hit(uint32_t cs, uint32_t ce, uint32_t s, uint32_t e, int32_t min_bp) {
    bool h = cs < e && ce > s;
    // multi_chrom_overlapper.rs:489-494: the bp filter only exists for min_overlap > 1
    if (FILTER) h = h && ((int64_t)min(e, ce) - (int64_t)max(s, cs) >= (int64_t)min_bp);
    return h;
}

// DESC: AIList emission order; FILTER: min_overlap > 1; OFFS: per-query offsets are written (gtgpu_find).
// LEAN: the record path only — no window walk, no generic walk, no call at all, which is worth 4 registers, every
// spill and 5 % of the step.  A warp that meets a query the records cannot resolve raises *lean_flag; the launcher
// queues the full kernel right behind the lean one with run_if = that flag, so the full kernel either exits at once
// (the usual case) or redoes the whole launch — still one asynchronous sequence on the stream, no host round trip.
// Round 1's lean kernel needed 48 registers: five CTAs (40 warps) per SM instead of four (6.93 -> 6.14 ms; six CTAs at 40
// registers spilled: 6.33 ms).  After round 2's instruction diet (one-word candidates, byte-packed counts, predicated scan
// steps; profiles/r02/sweep_instruction_diet_*.log) it fits 40 registers without a spill: six CTAs (48 warps) per SM,
// 5.99 -> 5.69 ms per 1e9 queries.
#ifndef GT_LEAN_MINBLOCKS
#define GT_LEAN_MINBLOCKS 6
#endif
// UNK1: the per-query [unk] rule of fragment tokenization (fragments.rs:42-47: every fragment is its own tokenize() call) —
// a query without a hit emits the single id unk_id, so the id stream IS the token stream and the per-query offsets index it.
// TAG (lean kernel only): every emitted id is accompanied by its query's tag (out_tags[pos] = tag_in[query]) — fragment
// tokenization gets its (barcode, token) pairs straight from the find, without per-query offsets going to HBM and back.
template <int ROWS, bool DESC, bool FILTER, bool OFFS, bool LEAN, bool UNK1 = false, bool TAG = false>
__global__ void __launch_bounds__(FUSED_BLOCK, LEAN ? ((OFFS || TAG) ? 5 : GT_LEAN_MINBLOCKS) : GT_FUSED_MINBLOCKS)  // with per-query offsets or tags 40 registers spill
fused_find_kernel(IndexView ix, uint64_t n, uint32_t n_tiles, uint64_t n_files, const uint64_t* __restrict__ file_offsets,
                  const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                  int32_t min_bp, int tma_ok, uint32_t* __restrict__ out_ids, uint64_t capacity,
                  uint64_t* __restrict__ out_offsets, uint64_t* __restrict__ out_file_tok, FusedWorkspace ws,
                  const uint64_t* __restrict__ d_base, uint64_t* __restrict__ d_total, uint32_t* __restrict__ d_err,
                  uint32_t* __restrict__ lean_flag, const uint32_t* __restrict__ run_if, uint32_t unk_id,
                  const uint32_t* __restrict__ tag_in, uint32_t* __restrict__ out_tags) {
    static_assert(!TAG || LEAN, "tags are written by the lean kernel only (its fallback writes offsets; fragments.cu tags from those)");
    static_assert(ROWS == 4, "state packing (2-bit counts, 8-bit offsets) is written for four rows per thread");
    if (!LEAN && run_if && *run_if == 0) return;  // fallback launch, and the lean kernel resolved everything
    constexpr int WARPS = FUSED_BLOCK / 32;
    constexpr int WTILE = 32 * ROWS;          // queries per warp
    constexpr int TILE = FUSED_BLOCK * ROWS;  // queries per block tile
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr uint32_t NO_TILE = 0xFFFFFFFFu;
    constexpr uint32_t WARP_SLOW = 1u << 31;  // the warp keeps full offsets in s_aux (some query has more than 7 hits)
    constexpr uint32_t WARP_SEMI = 1u << 30;  // packed offsets + row bases in s_rowbase (pool-list queries, <= 7 hits each)
    // TMA-staged query rows (chr, start, end).  ONE buffer is enough: a tile's rows move to registers at the top of its
    // resolve, and the next tile is only staged after barrier B2 of the same step.  Shared memory is deliberately kept
    // small (32 KB per CTA): the unified L1 holds the in-flight gather lines, and a larger carve-out costs throughput.
    __shared__ __align__(128) uint32_t s_q[3][TILE];
    __shared__ __align__(8) uint64_t s_bar;
    // Per-query side word, by iteration parity.  A warp on the packed path keeps the SECOND hit's val of its two-hit
    // queries here; a warp that had a generic query keeps every query's full offset instead (its two-hit queries are
    // then emitted by the generic walk too, so the two uses never meet in one warp).
    __shared__ uint32_t s_aux[2][TILE];
    __shared__ uint2 s_chrom[CHROM_CACHE];
    __shared__ uint32_t s_wtot[2][WARPS];      // per-warp hit totals, double-buffered by iteration parity
    __shared__ uint64_t s_lb_sum[2][WARPS];    // look-back partial sums per 32-tile window
    __shared__ uint32_t s_lb_p[2][WARPS];      // 1 when the window contains an inclusive prefix
    __shared__ uint32_t s_tile[2];             // tile index for this / the next iteration
    __shared__ uint32_t s_rowbase[2][WARPS];   // packed row bases of a WARP_SEMI warp, by iteration parity
    __shared__ uint32_t s_mark;                // file-boundary mark of the tile being emitted
    __shared__ uint32_t s_staged[2];           // 1 when the tile's queries were requested through TMA

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef GT_PHASE_TIMING
    __shared__ unsigned long long s_acc[WARPS][10];
    if (lane < 10) s_acc[warp][lane] = 0;
#endif
    const uint32_t nchr = ix.n_chroms;
    // the last cached entry is then an "unknown chromosome" sentinel; the lean kernel is only launched for such indexes
    const bool chrom_cached = LEAN || nchr < CHROM_CACHE;

    // One thread claims a tile and, when it is a full aligned tile, starts the bulk copies of its three query rows.
    // Hand the next tile to the CTA: its index and (when the rows are whole and aligned) the bulk copies of its query
    // rows.  Every hand-over completes one phase of s_bar — with the copies' bytes or with a plain arrive — so the
    // threads that wait on it at the top of the next step also see s_tile / s_staged (arrive releases, wait acquires).
    auto stage_tile = [&](uint32_t buf, const uint32_t t) {
        const bool staged = tma_ok && t < n_tiles && (uint64_t)(t + 1) * TILE <= n;
        s_tile[buf] = t;
        s_staged[buf] = staged;
        if (staged) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this buffer are done
            mbar_expect_tx(&s_bar, 3 * TILE * 4);
            const uint64_t q0 = (uint64_t)t * TILE;
            const uint64_t pol = policy_evict_first();
            bulk_g2s(&s_q[0][0], chr + q0, TILE * 4, &s_bar, pol);
            bulk_g2s(&s_q[1][0], start + q0, TILE * 4, &s_bar, pol);
            bulk_g2s(&s_q[2][0], end + q0, TILE * 4, &s_bar, pol);
        } else {
            mbar_arrive(&s_bar);
        }
    };

    for (uint32_t i = tid; i < CHROM_CACHE; i += FUSED_BLOCK)
    {
        uint2 cb = i < nchr ? __ldg(reinterpret_cast<const uint2*>(ix.chrom_bt + i)) : make_uint2(0, 0);
        if (LEAN) cb.y &= BT_NBINS_MASK;  // the lean kernel wants the bin count alone (no generic chromosome on its indexes)
        s_chrom[i] = cb;
    }
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stage_tile(0, atomicAdd(ws.counter, 1u));
    }
    __syncthreads();

    const uint64_t base = d_base ? *d_base : 0;
    uint64_t* status = ws.status;
    const uint32_t shift = ix.bt_shift;
    uint32_t bar_phase = 0;  // parity to wait for on s_bar
    const uint64_t keep = policy_evict_last();

    // One pipeline step: resolve `cur` (tile of this iteration), then look back + emit `prev` (tile of the last one).
    auto step = [&](TileState<ROWS>& cur, TileState<ROWS>& prev, const uint32_t par) -> bool {
        PHASE_START();
        mbar_wait(&s_bar, bar_phase);  // the hand-over of this step's tile (and its query rows, when staged)
        bar_phase ^= 1u;
        uint32_t tile = s_tile[par];
        if (tile >= n_tiles) tile = NO_TILE;
        cur.tile = tile;
        cur.slow = 0;

        // prev's look-back words are fetched in the middle of this step — after the gathers of `cur` have landed, so the
        // words are as fresh as possible (a stale "not ready" costs a reload between the barriers), early enough that the
        // round trip hides behind the warp scan and barrier B2.  Warps 0-1: the tiles before prev inside its supertile
        // (lanes past the supertile's first tile read as an empty aggregate); warp 2: the 32 supertiles before prev's
        // (past the beginning: a zero prefix); one thread of warp 3: the file-boundary mark of prev.
        constexpr uint32_t CLAIMER = FUSED_SUPER + 32;
        uint64_t lb_pre = 0;
        auto prefetch_lookback = [&]() {
            if (prev.tile == NO_TILE) return;
            if (tid < FUSED_SUPER) {
                const int64_t j = (int64_t)prev.tile - 1 - (int64_t)tid;
                lb_pre = j >= (int64_t)(prev.tile & ~(FUSED_SUPER - 1)) ? ld_status(status + j) : ST_FLAG_AGG;
            } else if (tid < FUSED_SUPER + 32) {
                const int64_t sj = (int64_t)(prev.tile / FUSED_SUPER) - 1 - (int64_t)lane;
                lb_pre = sj >= 0 ? ld_status(ws.super + sj) : (3ull << 62);
            } else if (tid == CLAIMER + 1) {
                if (out_file_tok) lb_pre = __ldg(ws.tile_file + prev.tile);
            }
        };

        if (tile != NO_TILE) {
            const uint64_t tile_start = (uint64_t)tile * TILE;
            const uint32_t wl = warp * WTILE + lane;
            // ---- queries: ROWS coalesced rows per array, from the TMA-staged copy when there is one -----------------
            uint32_t qc[ROWS], qs[ROWS], qe[ROWS];
            uint32_t vmask = (1u << ROWS) - 1;  // bit k: row k is a real query (only the last, partial tile has others)
            if (s_staged[par]) {
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    qc[k] = s_q[0][wl + 32 * k];
                    qs[k] = s_q[1][wl + 32 * k];
                    qe[k] = s_q[2][wl + 32 * k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint64_t q = tile_start + wl + 32 * k;
                    const bool ok = q < n;
                    if (!ok) vmask &= ~(1u << k);
                    qc[k] = ok ? __ldcs(chr + q) : 0xFFFFFFFFu;
                    qs[k] = ok ? __ldcs(start + q) : 0;
                    qe[k] = ok ? __ldcs(end + q) : 1;  // past the end: an unknown-chromosome query [0, 1) the records resolve to nothing
                }
            }
            PHASE_MARK(0);
            // ---- resolve through the window records: one gather per query, candidates inline (common.cuh) -----------
            uint32_t cnt[ROWS];
            uint32_t mine = 0;  // LEAN: the four counts, one byte per row — all the record path ever needs of them
            {
                const uint32_t bin_mask = (1u << shift) - 1;
                uint4 r[ROWS];
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t c = qc[k], s = qs[k], e = qe[k];
                    uint2 cb;
                    if (chrom_cached) cb = s_chrom[min(c, (uint32_t)CHROM_CACHE - 1)];
                    else cb = c < nchr ? __ldg(reinterpret_cast<const uint2*>(ix.chrom_bt + c)) : make_uint2(0, 0);
                    // Record path: the query starts in bin b1 and its last position e-1 lies inside window b1.  That also
                    // admits empty and REVERSED queries (e <= s, as gtars-scoring builds them) whose e is past the window
                    // base: a hit needs start < e <= s < end, so it contains position s, touches bin b1 and is listed in
                    // the window; the window-relative comparison below is then exact for them too.
                    const uint32_t b1 = s >> shift;
                    const uint32_t d = e - 1u - (b1 << shift);  // wraps to a huge value when e-1 is before the window
                    if constexpr (LEAN) {
                        // an index the lean kernel runs on has no generic chromosome, and its tables end far enough below
                        // 2^32 that e == 0 (never a hit) cannot pass the test on d (index.cu).
                        if (d >= (2u << shift)) cur.slow |= 1u << k;
                    } else {
                        if ((cb.y == BT_GENERIC_CHROM) | (e == 0u) | (d >= (2u << shift))) cur.slow |= 0x11u << k;  // bit k+4: not a window query
                    }
                    // bins past the chromosome's last one read the empty record the builder put behind it (index.cu); an
                    // absent or unknown chromosome has (off, n_bins) = (0, 0): record 0, the global empty sentinel
                    const uint32_t li = b1 < (LEAN ? cb.y : cb.y & BT_NBINS_MASK) ? cb.x + b1 : 0u;
                    r[k] = ldg128_keep(ix.bt_rec + (size_t)li * 4, keep);
                }
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t s = qs[k] & bin_mask;  // a slow query's values are never used
                    bool h0, h1, rec_slow;
                    // (d + 0x8000) - (s_rel << 16) with d = end - 1 - base = qe - qs + s_rel - 1 (common.cuh, window records)
                    const uint32_t qd = (qe[k] - qs[k] + 0x7FFFu) - s * 0xFFFFu;
                    if constexpr (FILTER) {
                        // multi_chrom_overlapper.rs:489-494, 560-563: min(e, ce) - max(s, cs) >= min_bp (> 1, so it implies the
                        // overlap itself).  That difference is the smallest of e - s, e - cs, ce - s and ce - cs, and the two
                        // halves of T are e - cs and ce - s (each + 0x7FFF): no coordinate has to be unpacked.
                        const int w = (int)(qe[k] - qs[k]);
                        auto bp = [&](uint32_t word) {
                            const uint32_t T = word + qd;
                            const int a = (int)(T & 0xFFFFu) - 0x7FFF, b = (int)(T >> 16) - 0x7FFF;
                            return min(min(w, a), min(b, a + b - w));
                        };
                        h0 = bp(r[k].x) >= min_bp;
                        h1 = bp(r[k].z) >= min_bp;
                    } else {
                        h0 = ((r[k].x + qd) & 0x80008000u) == 0x80008000u;
                        h1 = ((r[k].z + qd) & 0x80008000u) == 0x80008000u;
                    }
                    rec_slow = r[k].x == BT_REC_SLOW;
                    cur.v0[k] = h0 ? r[k].y : r[k].w;
                    if constexpr (LEAN) {
                        mine += h0 ? (1u << (8 * k)) : 0u;
                        mine += h1 ? (1u << (8 * k)) : 0u;
                        s_aux[par][wl + 32 * k] = r[k].w;  // only read back for a two-hit query: no predicate needed
                        if (rec_slow) cur.slow |= 1u << k;  // pool list or overflow: not for this kernel
                        if (UNK1 && !(h0 | h1) && ((vmask >> k) & 1)) {
                            mine += 1u << (8 * k);  // no hit: the fragment's token is [unk] (moot if some row is slow: the launch is redone)
                            cur.v0[k] = unk_id;
                        }
                    } else {
                        cnt[k] = (uint32_t)h0 + (uint32_t)h1;
                        if (h0 & h1) s_aux[par][wl + 32 * k] = r[k].w;
                        if (rec_slow) {  // pool list or overflow (whatever the tests above said is overridden by the walk)
                            cur.slow |= 1u << k;
                            cur.v0[k] = r[k].y;  // the full kernel walks the list from this word
                        }
                        if (UNK1 && cnt[k] == 0 && !((cur.slow >> k) & 1) && ((vmask >> k) & 1)) {
                            cnt[k] = 1;  // no hit: the fragment's token is [unk]
                            cur.v0[k] = unk_id;
                        }
                    }
                }
            }
            prefetch_lookback();
            PHASE_MARK(1);
            // ---- warp scan: exclusive offset of every query inside the warp's slice -----------------------------------
            uint32_t warp_total = 0;
            if (!__any_sync(FULL, cur.slow != 0)) {
                // every count is 0, 1 or 2: scan the four rows at once, one byte per row (row sums <= 64)
                if constexpr (!LEAN) mine = cnt[0] | (cnt[1] << 8) | (cnt[2] << 16) | (cnt[3] << 24);
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) incl = shfl_up_add(incl, d);
                const uint32_t tot = __shfl_sync(FULL, incl, 31);
                // row bases 0, t0, t0+t1, t0+t1+t2 (each <= 192: no carry between the bytes), added bytewise to the exclusive
                // offsets inside the rows
                const uint32_t t0 = tot & 0xFF, t1 = (tot >> 8) & 0xFF, t2 = (tot >> 16) & 0xFF, t3 = tot >> 24;
                cur.offpack = incl - mine + ((t0 << 8) | ((t0 + t1) << 16) | ((t0 + t1 + t2) << 24));
                warp_total = t0 + t1 + t2 + t3;
                if constexpr (LEAN) cur.cntpack = mine;  // the lean kernel keeps its counts one byte per row
                else cur.cntpack = cnt[0] | (cnt[1] << 2) | (cnt[2] << 4) | (cnt[3] << 6);
            } else if constexpr (LEAN) {
                // not resolvable by records alone: this launch's output will be replaced by the full kernel's
                if (lane == 0) *lean_flag = 1u;
                cur.cntpack = 0;
                cur.offpack = 0;
                cur.slow = 0;
            } else {
                // Some queries left the record path.  Pool-list windows (nested intervals, several AIList components) are
                // walked right here from the window word the record carried; everything else goes through the window walk /
                // the generic LUT + walk.  Either way a query with at most two hits rejoins the register path (its vals are in
                // emission order, so an AIList pair is stored swapped for the emit's swap); longer ones are walked again at emit.
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    if (!((cur.slow >> k) & 1)) continue;
                    uint32_t c, a = 0, b = 0;
                    bool ordered;
                    const uint32_t w = cur.v0[k];
                    if (((cur.slow >> k) & 0x10u) == 0 && w != BT_OVERFLOW && qs[k] < qe[k]) {
                        const uint32_t nl = w & BT_POOL_MAX, first = (w & ~BT_POOL_FLAG) >> BT_POOL_SHIFT;
                        c = 0;
#pragma unroll 1
                        for (uint32_t t = 0; t < nl; ++t) {
                            const uint4 E = ldg128(ix.bt_ent + __ldg(ix.bt_pool + first + t));
                            if (!cand_hit<FILTER>(E.x, E.y, qs[k], qe[k], min_bp)) continue;
                            if (c == 0) a = E.z;
                            else if (c == 1) b = E.z;
                            ++c;
                        }
                        ordered = true;
                    } else {
                        uint32_t ab[2];
                        const uint32_t r = count_query_first2(ix, qc[k], qs[k], qe[k], min_bp, ab);
                        c = r & 0x7FFFFFFFu;
                        ordered = (r >> 31) != 0;
                        a = ab[0];
                        b = ab[1];
                    }
                    if (UNK1 && c == 0 && ((vmask >> k) & 1)) {
                        c = 1;
                        a = unk_id;
                        ordered = true;
                    }
                    cnt[k] = c;
                    if (ordered && c <= 2) {
                        cur.slow &= ~(1u << k);
                        cur.v0[k] = (DESC && c == 2) ? b : a;
                        if (c == 2) s_aux[par][wl + 32 * k] = DESC ? a : b;
                    }
                }
                cur.slow &= 0xFu;
                cur.cntpack = min(cnt[0], 3u) | (min(cnt[1], 3u) << 2) | (min(cnt[2], 3u) << 4) | (min(cnt[3], 3u) << 6);
                if (!__any_sync(FULL, (cnt[0] | cnt[1] | cnt[2] | cnt[3]) > BT_SEMI_MAX)) {
                    // pool-list windows (nested intervals, several AIList components): every count is at most 7, so a row
                    // sums to at most 224 — two scans of 16-bit fields; offsets inside a row stay packed in a register,
                    // the three row bases (warp-uniform) go to shared memory, and s_aux keeps the two-hit second vals
                    cur.slow |= WARP_SEMI;
                    const uint32_t lo = cnt[0] | (cnt[1] << 16), hi = cnt[2] | (cnt[3] << 16);
                    uint32_t ilo = lo, ihi = hi;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t a = __shfl_up_sync(FULL, ilo, d), b = __shfl_up_sync(FULL, ihi, d);
                        if (lane >= d) { ilo += a; ihi += b; }
                    }
                    const uint32_t tlo = __shfl_sync(FULL, ilo, 31), thi = __shfl_sync(FULL, ihi, 31);
                    const uint32_t t0 = tlo & 0xFFFF, t1 = tlo >> 16, t2 = thi & 0xFFFF, t3 = thi >> 16;
                    const uint32_t elo = ilo - lo, ehi = ihi - hi;
                    cur.offpack = (elo & 0xFF) | ((elo >> 16) << 8) | ((ehi & 0xFF) << 16) | ((ehi >> 16) << 24);
                    if (lane == 0) s_rowbase[par][warp] = t0 | ((t0 + t1) << 10) | ((t0 + t1 + t2) << 20);
                    warp_total = t0 + t1 + t2 + t3;
                } else {
                    // unbounded counts: one 32-bit scan per row; full offsets go to shared memory
                    cur.slow |= WARP_SLOW;
                    uint64_t wide = 0;
#pragma unroll
                    for (int k = 0; k < ROWS; ++k) {
                        uint32_t incl = cnt[k];
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            uint32_t t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        s_aux[par][wl + 32 * k] = warp_total + incl - cnt[k];
                        if (cnt[k] == 2) cur.slow |= 1u << k;  // its second val was just overwritten: emit through the walk
                        warp_total += __shfl_sync(FULL, incl, 31);
                        wide += cnt[k];
                    }
                    cur.offpack = 0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) wide += __shfl_down_sync(FULL, wide, d);
                    if (lane == 0 && wide > 0x0FFFFFFFull) atomicExch(d_err, 1u);  // the sum over the 8 warps must fit 32 bits
                }
            }
            if (lane == 0) s_wtot[par][warp] = warp_total;
            PHASE_MARK(2);
        } else {
            prefetch_lookback();
        }
        __syncthreads();  // B2: warp totals of `cur` visible; s_tile[par] / s_q consumed by everyone
        PHASE_MARK(3);

        // Between the two barriers only the look-back warps (0-2) have real work; everything that talks to L2 with a
        // "strong" operation (each holds its warp for hundreds of cycles) is given to one thread of an otherwise idle
        // warp: the claim to warp 3, the aggregate's publication to warp 4, the prefix's (after B3) to warp 5.
        constexpr uint32_t PUBLISH_AGG = 128, PUBLISH_PREFIX = 160;
        static_assert(FUSED_BLOCK >= 192, "role threads");
        auto aggregate = [&]() {
            // lane w < WARPS holds warp w's total: two hardware warp sums instead of eight loads + adds in every lane
            static_assert(WARPS <= 32, "one lane per warp total");
            const uint32_t t = lane < (uint32_t)WARPS ? s_wtot[par][lane] : 0u;
            cur.warp_excl = __reduce_add_sync(FULL, lane < warp ? t : 0u);
            cur.tile_agg = __reduce_add_sync(FULL, t);
        };
        if (tile != NO_TILE && warp > FUSED_SUPER / 32) {
            aggregate();
            if (tid == PUBLISH_AGG) {
                st_status(status + tile, ST_FLAG_AGG | (uint64_t)cur.tile_agg);
                red_status(ws.super + tile / FUSED_SUPER, (1ull << 56) + (uint64_t)cur.tile_agg);
            } else if (tid == CLAIMER) {
                // next tile: nobody waits for the atomic's round trip before barrier B3 — the claimer picks the result
                // up after it and stages the query copies
                lb_pre = atomicAdd(ws.counter, 1u);
            }
        }
        PHASE_MARK(4);

        if (prev.tile != NO_TILE) {
            // ---- two-level decoupled look-back for the PREVIOUS tile (see FUSED_SUPER above) --------------------------
            const uint32_t ppar = par ^ 1;  // parity under which prev's shared-memory side state was written
            if (warp < FUSED_SUPER / 32) {
                const int64_t j = (int64_t)prev.tile - 1 - (int64_t)tid;
                uint64_t v = lb_pre;
                for (uint32_t spin = 0; __any_sync(FULL, (v >> 62) == 0); ++spin) {
#ifdef GT_PHASE_TIMING
                    if (lane == 0) s_acc[warp][8] += 1;
#endif
                    if ((v >> 62) == 0) {
                        __nanosleep(32);
                        v = ld_status(status + j);
                    }
                    if ((spin & 1023) == 1023 && *(volatile uint32_t*)d_err) break;  // a failed launch never hangs
                }
                const uint32_t pmask = __ballot_sync(FULL, (v >> 62) == 2);
                uint64_t val = v & ST_MASK;
                if (pmask && lane > (uint32_t)(__ffs(pmask) - 1)) val = 0;
                val = warp_sum_u63(val);
                if (lane == 0) {
                    s_lb_sum[par][warp] = val;
                    s_lb_p[par][warp] = pmask != 0;
                }
            } else if (warp == FUSED_SUPER / 32) {
                int64_t sj = (int64_t)(prev.tile / FUSED_SUPER) - 1 - (int64_t)lane;
                uint64_t v = lb_pre, acc = 0;
                for (;;) {
                    for (uint32_t spin = 0; __any_sync(FULL, ((v >> 62) & 1) == 0); ++spin) {
#ifdef GT_PHASE_TIMING
                        if (lane == 0) s_acc[warp][8] += 1;
#endif
                        if (((v >> 62) & 1) == 0) {
                            __nanosleep(32);
                            v = ld_status(ws.super + sj);
                        }
                        if ((spin & 1023) == 1023 && *(volatile uint32_t*)d_err) { v = 3ull << 62; }
                    }
                    const uint32_t pmask = __ballot_sync(FULL, (v >> 62) == 3);
                    uint64_t val = v & SUPER_MASK;
                    if (pmask && lane > (uint32_t)(__ffs(pmask) - 1)) val = 0;
                    acc += warp_sum_u63(val);
                    if (pmask) break;
                    sj -= 32;  // rare: no finished supertile among the last 32 (2 048 tiles)
                    v = sj >= 0 ? ld_status(ws.super + sj) : (3ull << 62);
                }
                if (lane == 0) s_lb_sum[par][warp] = acc;
            } else if (tid == CLAIMER + 1) {
                s_mark = (uint32_t)lb_pre;
            }
            __syncthreads();  // B3
            if (tid == CLAIMER) stage_tile(par ^ 1, tile != NO_TILE ? (uint32_t)lb_pre : NO_TILE);
            if (tile != NO_TILE && warp <= FUSED_SUPER / 32) aggregate();  // the look-back warps catch up on `cur`
            uint64_t excl = s_lb_sum[par][0];
            if (!s_lb_p[par][0]) {
                excl += s_lb_sum[par][1];
                if (!s_lb_p[par][1]) excl += s_lb_sum[par][2];
            }
            PHASE_MARK(5);
#ifdef GT_PHASE_TIMING
            if (lane == 0) s_acc[warp][9] += 1;
#endif
            const uint64_t tile_start = (uint64_t)prev.tile * TILE;
            const uint64_t tile_base = base + excl;
            if (tid == PUBLISH_PREFIX) {
                st_status(status + prev.tile, ST_FLAG_PREFIX | (excl + prev.tile_agg));
                // the supertile's first tile contributes the prefix in front of the supertile, exactly once
                if ((prev.tile & (FUSED_SUPER - 1)) == 0) red_status(ws.super + prev.tile / FUSED_SUPER, (1ull << 63) + excl);
                if (excl + prev.tile_agg >= (1ull << 55)) atomicExch(d_err, 1u);  // accumulator words hold 56 bits
                if (prev.tile == n_tiles - 1) {
                    *d_total = tile_base + prev.tile_agg;
                    if (OFFS) out_offsets[n] = tile_base + prev.tile_agg;
                }
            }

            // ---- emit: a row's ids land in consecutive words ------------------------------------------------------------
            const uint32_t wl = warp * WTILE + lane;
            const uint64_t warp_base = tile_base + prev.warp_excl;
            uint32_t* outp = out_ids + warp_base;
            const bool fits = tile_base + prev.tile_agg <= capacity;
            if (!LEAN && (prev.slow & WARP_SEMI)) {
                const uint32_t rb = s_rowbase[ppar][warp];  // bases of rows 1..3, 10 bits each (row 0 starts at zero)
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t c = (prev.cntpack >> (2 * k)) & 3;
                    const uint32_t o = ((prev.offpack >> (8 * k)) & 0xFF) + (k == 0 ? 0u : (rb >> (10 * (k - 1))) & 0x3FF);
                    const uint64_t pos = warp_base + o;
                    const uint64_t q = tile_start + wl + 32 * k;
                    if (OFFS && q < n) out_offsets[q] = pos;
                    if (c == 0) continue;
                    if ((prev.slow >> k) & 1) {
                        emit_query_walk(ix, __ldg(chr + q), __ldg(start + q), __ldg(end + q), min_bp, out_ids, pos, capacity);
                    } else {
                        uint32_t a = prev.v0[k], b = 0;
                        if (c == 2) {
                            b = s_aux[ppar][wl + 32 * k];
                            if (DESC) { const uint32_t t = a; a = b; b = t; }
                        }
                        if (pos < capacity) out_ids[pos] = a;
                        if (c == 2 && pos + 1 < capacity) out_ids[pos + 1] = b;
                    }
                }
            } else if (LEAN && fits) {
                // the usual case of the lean kernel, branch-free per row: counts one byte per row, plain streaming stores
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t c = (prev.cntpack >> (8 * k)) & 3, o = (prev.offpack >> (8 * k)) & 0xFF;
                    if (OFFS && tile_start + wl + 32 * k < n) out_offsets[tile_start + wl + 32 * k] = warp_base + o;
                    uint32_t a = prev.v0[k], b = s_aux[ppar][wl + 32 * k];
                    if (DESC && c == 2) { const uint32_t t = a; a = b; b = t; }
                    if (c != 0) __stcs(outp + o, a);
                    if (c == 2) __stcs(outp + o + 1, b);
                    if constexpr (TAG) {
                        if (c != 0) {  // (a query past the end has no id)
                            const uint32_t t = __ldcs(tag_in + tile_start + wl + 32 * k);
                            __stcs(out_tags + warp_base + o, t);
                            if (c == 2) __stcs(out_tags + warp_base + o + 1, t);
                        }
                    }
                }
            } else if (LEAN || !(prev.slow & WARP_SLOW)) {
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t c = (prev.cntpack >> ((LEAN ? 8 : 2) * k)) & 3, o = (prev.offpack >> (8 * k)) & 0xFF;
                    if (OFFS && tile_start + wl + 32 * k < n) out_offsets[tile_start + wl + 32 * k] = warp_base + o;
                    if (c == 0) continue;
                    uint32_t a = prev.v0[k], b = 0;
                    if (c == 2) {
                        b = s_aux[ppar][wl + 32 * k];
                        if (DESC) { const uint32_t t = a; a = b; b = t; }
                    }
                    if (fits) {
                        __stcs(outp + o, a);
                        if (c == 2) __stcs(outp + o + 1, b);
                    } else {
                        if (warp_base + o < capacity) out_ids[warp_base + o] = a;
                        if (c == 2 && warp_base + o + 1 < capacity) out_ids[warp_base + o + 1] = b;
                    }
                    if constexpr (TAG) {
                        const uint32_t t = __ldcs(tag_in + tile_start + wl + 32 * k);
                        if (warp_base + o < capacity) out_tags[warp_base + o] = t;
                        if (c == 2 && warp_base + o + 1 < capacity) out_tags[warp_base + o + 1] = t;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t c = (prev.cntpack >> (2 * k)) & 3, o = s_aux[ppar][wl + 32 * k];
                    const uint64_t pos = warp_base + o;
                    const uint64_t q = tile_start + wl + 32 * k;
                    if (OFFS && q < n) out_offsets[q] = pos;
                    if (c == 0) continue;
                    if ((prev.slow >> k) & 1) {
                        // q < n here: an out-of-range query has count 0
                        emit_query_walk(ix, __ldg(chr + q), __ldg(start + q), __ldg(end + q), min_bp, out_ids, pos, capacity);
                    } else if (pos < capacity) {
                        out_ids[pos] = prev.v0[k];  // exactly one hit: two-hit queries of this warp took the walk
                    }
                }
            }
            PHASE_MARK(6);

            // ---- file boundaries inside this tile: raw token offset of each file's first query ------------------------
            if (out_file_tok) {
                const uint32_t mark = s_mark;  // block-uniform, prefetched before the resolve
                if (mark != 0) {
                    // every warp publishes full offsets (tile-relative) for this rare tile, then boundaries are looked up
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < ROWS; ++k) {
                        uint32_t o = (!LEAN && (prev.slow & WARP_SLOW)) ? s_aux[ppar][wl + 32 * k] : (prev.offpack >> (8 * k)) & 0xFF;
                        if (!LEAN && (prev.slow & WARP_SEMI) && k > 0) o += (s_rowbase[ppar][warp] >> (10 * (k - 1))) & 0x3FF;
                        s_aux[ppar][wl + 32 * k] = prev.warp_excl + o;
                    }
                    __syncthreads();
                    const uint64_t limit = (prev.tile == n_tiles - 1) ? n + 1 : tile_start + TILE;
                    for (uint64_t f = (uint64_t)(0xFFFFFFFFu - mark) + tid; f <= n_files; f += FUSED_BLOCK) {
                        const uint64_t qi = file_offsets[f];
                        if (qi >= limit) break;
                        const uint32_t r = (uint32_t)(qi - tile_start);
                        out_file_tok[f] = tile_base + (r < (uint32_t)TILE ? s_aux[ppar][r] : prev.tile_agg);
                    }
                    __syncthreads();
                }
            }
            PHASE_MARK(7);
        } else {
            if (tid == CLAIMER) stage_tile(par ^ 1, tile != NO_TILE ? (uint32_t)lb_pre : NO_TILE);
            if (tile != NO_TILE && warp <= FUSED_SUPER / 32) aggregate();
        }
        return tile != NO_TILE;
    };

    // Ping-pong the two tile states so nothing is copied between iterations.
    TileState<ROWS> ta, tb;
    tb.tile = NO_TILE;
    for (;;) {
        if (!step(ta, tb, 0)) break;
        if (!step(tb, ta, 1)) break;
    }
#ifdef GT_PHASE_TIMING
    __syncwarp();
    if (lane < 10) atomicAdd(&g_phase_cycles[warp * 16 + lane], s_acc[warp][lane]);
#endif
}

template <bool DESC, bool FILTER, bool OFFS, bool LEAN, bool UNK1 = false, bool TAG = false>
static cudaError_t launch_variant(int grid, cudaStream_t st, const IndexView& view, uint64_t n, uint32_t n_tiles,
                                  uint64_t n_files, const uint64_t* d_file_offsets, const uint32_t* d_chr,
                                  const uint32_t* d_start, const uint32_t* d_end, int32_t min_overlap, uint32_t* d_out_ids,
                                  uint64_t cap, uint64_t* d_out_offsets, uint64_t* d_out_file_tok, FusedWorkspace ws,
                                  const uint64_t* d_base, uint64_t* d_total_out, uint32_t* d_errflag, const uint32_t* run_if,
                                  int* blocks_per_sm, uint32_t unk_id = 0, const uint32_t* d_tag_in = nullptr,
                                  uint32_t* d_out_tags = nullptr) {
    auto kern = fused_find_kernel<FUSED_ROWS, DESC, FILTER, OFFS, LEAN, UNK1, TAG>;
    if (blocks_per_sm) {
        // 4 CTAs x 24 KB fit the 100 KB shared-memory configuration (6 CTAs of the lean kernel: 164 KB; 60-70 % measured
        // alike, 75 % costs 6 %); the rest of the unified L1 serves the gathers
        int carve = LEAN ? (GT_LEAN_MINBLOCKS >= 6 && !OFFS && !TAG ? 65 : 50) : 40;
        if (const char* env = getenv("GTGPU_CARVEOUT")) carve = atoi(env);  // tuning knob, percent
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, FUSED_BLOCK, 0);
    }
    // bulk copies need 16-byte aligned sources; tile starts are multiples of 4 KiB, so only the bases matter
    static const bool no_tma = getenv("GTGPU_NO_TMA") != nullptr;  // tuning knob: plain coalesced loads instead of bulk copies
    const int tma_ok = !no_tma && ((reinterpret_cast<uintptr_t>(d_chr) | reinterpret_cast<uintptr_t>(d_start) |
                                    reinterpret_cast<uintptr_t>(d_end)) & 15) == 0;
    kern<<<grid, FUSED_BLOCK, 0, st>>>(view, n, n_tiles, n_files, d_file_offsets, d_chr, d_start, d_end, min_overlap, tma_ok,
                                       d_out_ids, cap, d_out_offsets, d_out_file_tok, ws, d_base, d_total_out, d_errflag,
                                       ws.lean_flag, run_if, unk_id, d_tag_in, d_out_tags);
    return cudaGetLastError();
}

#ifdef GT_PHASE_TIMING
extern "C" void gtgpu_debug_phase_cycles(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16 * 8);
    if (reset) {
        unsigned long long z[16 * 8] = {0};
        cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
    }
}
#endif

int32_t launch_fused_find(gtgpu_index* ix, uint64_t n, uint64_t n_files, const uint64_t* d_file_offsets,
                          const uint32_t* d_chr, const uint32_t* d_start, const uint32_t* d_end,
                          int32_t min_overlap, uint32_t* d_out_ids, uint64_t ids_capacity,
                          uint64_t* d_out_offsets, uint64_t* d_out_file_tok, void* d_workspace,
                          const uint64_t* d_base, uint64_t* d_total_out, uint32_t* d_errflag, int unk_per_query, uint32_t unk_id,
                          const uint32_t* d_tag_in, uint32_t* d_out_tags, const uint32_t** tags_pending_if) {
    gtgpu_ctx* ctx = ix->ctx;
    if (d_out_tags && (!unk_per_query || !d_tag_in || !tags_pending_if))
        return fail(GTGPU_ERR_INVALID, "fused_find: tags go with the per-query [unk] rule");
    if (tags_pending_if) *tags_pending_if = nullptr;
    if (unk_per_query && (!d_out_offsets || min_overlap > 1 || d_out_file_tok || n == 0))
        return fail(GTGPU_ERR_INVALID, "fused_find: the per-query [unk] rule needs per-query offsets, no bp filter, no files, n > 0");
    cudaStream_t st = ctx->stream;
    if (n == 0) {
        // No queries: every offset equals the base, the total is the base.
        uint64_t cnt = d_out_file_tok ? n_files + 1 : 0;
        if (cnt) {
            fill_from_base_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(cnt, d_out_file_tok, d_base);
            ctx->launches++;
        }
        if (d_out_offsets) {
            fill_from_base_kernel<<<1, 32, 0, st>>>(1, d_out_offsets, d_base);
            ctx->launches++;
        }
        fill_from_base_kernel<<<1, 32, 0, st>>>(1, d_total_out, d_base);
        ctx->launches++;
        GT_CUDA(cudaGetLastError());
        return GTGPU_OK;
    }
    uint64_t tiles64 = n_tiles_for(n);
    if (tiles64 > 0x7FFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "fused_find: too many queries for one launch");
    uint32_t n_tiles = (uint32_t)tiles64;
    // Lean kernel first when every window of the index is a plain record (no pool lists, no overflow windows, no
    // chromosome without a table) and it has not failed on this index before; the full kernel follows as its
    // on-device fallback (see fused_find_kernel).  GTGPU_LEAN=0 forces the full kernel.
    if (ix->h_lean_probe && *ix->h_lean_probe) ix->lean_off = true;  // a previous launch had to fall back: stop trying
    static const bool lean_env = !(getenv("GTGPU_LEAN") && getenv("GTGPU_LEAN")[0] == '0');
    const bool use_lean = lean_env && ix->bt_clean && !ix->lean_off;
    FusedWorkspace ws = carve(d_workspace, n, 0);
    const FusedWorkspace ws_fallback = carve(d_workspace, n, 1);
    GT_CUDA(cudaMemsetAsync(d_workspace, 0, fused_workspace_bytes(n), st));
    if (d_out_file_tok) {
        uint64_t cnt = n_files + 1;
        mark_file_tiles_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(n_files, d_file_offsets, n_tiles, ws.tile_file);
        ctx->launches++;
    }
    // Keep the bin table resident in L2 while 12 B/query of streaming input flows past it: an access-policy window on
    // the ctx stream, moved whenever another index launches and dropped when its index is freed.
    if (ctx->l2_window_owner != ix) {
        ctx->l2_window_owner = ix;
        const char* env = getenv("GTGPU_L2_PERSIST");
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof attr);
        if (ix->bt_bins && !(env && env[0] == '0')) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            size_t bytes = (size_t)ix->bt_bins * sizeof(uint32_t) * BT_REC_WORDS;
            size_t win = std::min<size_t>(bytes, (size_t)std::max(max_window, 0));
            if (max_persist > 0 && win > 0) {
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(win, (size_t)max_persist));
                attr.accessPolicyWindow.base_ptr = const_cast<uint32_t*>(ix->view.bt_rec);
                attr.accessPolicyWindow.num_bytes = win;
                attr.accessPolicyWindow.hitRatio = std::min(1.0f, (float)max_persist / (float)win);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            }
        }
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);  // num_bytes == 0 clears the window
        cudaGetLastError();
    }
    const bool desc = ix->view.descending != 0, filt = min_overlap > 1, offs = d_out_offsets != nullptr;
    const int variant = unk_per_query ? (desc ? 9 : 8) : (desc ? 4 : 0) | (filt ? 2 : 0) | (offs ? 1 : 0);
    int (&blocks_per_sm)[2][12] = ctx->fused_bps;  // per ctx = per device: carve-out and occupancy are set on this device
    int grid = 0;
    cudaError_t err = cudaSuccess;
#define GT_LAUNCH(D, F, O, LEAN, WS, RUN_IF, U)                                                                           \
    do {                                                                                                                \
        int& bps = blocks_per_sm[LEAN ? 1 : 0][variant];                                                                \
        if (!bps) {                                                                                                     \
            err = launch_variant<D, F, O, LEAN, U>(0, st, ix->view, n, n_tiles, n_files, d_file_offsets, d_chr, d_start, d_end, \
                                                min_overlap, d_out_ids, ids_capacity, d_out_offsets, d_out_file_tok, WS,  \
                                                d_base, d_total_out, d_errflag, RUN_IF, &bps, unk_id);                  \
            if (err != cudaSuccess) break;                                                                              \
            if (bps < 1) bps = 1;                                                                                       \
        }                                                                                                               \
        grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * bps);                                         \
        err = launch_variant<D, F, O, LEAN, U>(grid, st, ix->view, n, n_tiles, n_files, d_file_offsets, d_chr, d_start, d_end,  \
                                            min_overlap, d_out_ids, ids_capacity, d_out_offsets, d_out_file_tok, WS, d_base, \
                                            d_total_out, d_errflag, RUN_IF, nullptr, unk_id);                            \
        ctx->launches++;                                                                                                \
    } while (0)
#define GT_VARIANT_AT(V, D, F, O, U)                                                                                     \
    case V:                                                                                                             \
        ctx->time_begin();                                                                                              \
        if (use_lean) {                                                                                                 \
            GT_LAUNCH(D, F, O, true, ws, nullptr, U);                                                                   \
            if (err == cudaSuccess) GT_LAUNCH(D, F, O, false, ws_fallback, ws.lean_flag, U);                            \
        } else {                                                                                                        \
            GT_LAUNCH(D, F, O, false, ws, nullptr, U);                                                                  \
        }                                                                                                               \
        ctx->time_end();                                                                                                \
        break;
#define GT_VARIANT(D, F, O) GT_VARIANT_AT(((D ? 4 : 0) | (F ? 2 : 0) | (O ? 1 : 0)), D, F, O, false)
    // Fragment tokenization with tags: the lean kernel writes (id, tag) pairs and no offsets; its fallback is the full
    // kernel WITH offsets (the caller then tags from those: *tags_pending_if = the flag that says so, on the device).
#define GT_TAGGED(D)                                                                                                    \
    do {                                                                                                                \
        int& bps = blocks_per_sm[1][10 + (D ? 1 : 0)];                                                                  \
        ctx->time_begin();                                                                                              \
        if (!bps) {                                                                                                     \
            err = launch_variant<D, false, false, true, true, true>(0, st, ix->view, n, n_tiles, n_files, d_file_offsets, d_chr, d_start, \
                                                                    d_end, min_overlap, d_out_ids, ids_capacity, nullptr, nullptr, ws,    \
                                                                    d_base, d_total_out, d_errflag, nullptr, &bps, unk_id, d_tag_in,      \
                                                                    d_out_tags);                                        \
            if (bps < 1) bps = 1;                                                                                       \
        }                                                                                                               \
        grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * bps);                                         \
        if (err == cudaSuccess)                                                                                         \
            err = launch_variant<D, false, false, true, true, true>(grid, st, ix->view, n, n_tiles, n_files, d_file_offsets, d_chr,       \
                                                                    d_start, d_end, min_overlap, d_out_ids, ids_capacity, nullptr,        \
                                                                    nullptr, ws, d_base, d_total_out, d_errflag, nullptr, nullptr, unk_id, \
                                                                    d_tag_in, d_out_tags);                              \
        ctx->launches++;                                                                                                \
        if (err == cudaSuccess) GT_LAUNCH(D, false, true, false, ws_fallback, ws.lean_flag, true);                      \
        ctx->time_end();                                                                                                \
        *tags_pending_if = ws.lean_flag;                                                                                \
    } while (0)
    if (d_out_tags && use_lean) {
        if (desc) GT_TAGGED(true);
        else GT_TAGGED(false);
    } else
    switch (variant) {
        GT_VARIANT(false, false, false)
        GT_VARIANT(false, false, true)
        GT_VARIANT(false, true, false)
        GT_VARIANT(false, true, true)
        GT_VARIANT(true, false, false)
        GT_VARIANT(true, false, true)
        GT_VARIANT(true, true, false)
        GT_VARIANT(true, true, true)
        GT_VARIANT_AT(8, false, false, true, true)  // per-query [unk] rule (fragments): offsets on, no filter
        GT_VARIANT_AT(9, true, false, true, true)
    }
#undef GT_TAGGED
#undef GT_VARIANT
#undef GT_VARIANT_AT
#undef GT_LAUNCH
    if (err != cudaSuccess) return fail(GTGPU_ERR_CUDA, std::string("fused_find launch: ") + cudaGetErrorString(err));
    if (use_lean && ix->h_lean_probe) cudaMemcpyAsync(ix->h_lean_probe, ws.lean_flag, 4, cudaMemcpyDeviceToHost, st);
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// ================================================================================================================
// chromosome ids given as runs (sorted BED files): expand the runs that intersect [q0, q0 + cn) into out[0..cn)
// ================================================================================================================
__global__ void expand_runs_kernel(uint64_t n_runs, const uint64_t* __restrict__ run_offsets, const uint32_t* __restrict__ run_chr,
                                   uint64_t q0, uint64_t cn, uint32_t* __restrict__ out) {
    __shared__ uint64_t s_first;
    const uint64_t per_block = (cn + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = q0 + (uint64_t)blockIdx.x * per_block, hi = min(q0 + cn, lo + per_block);
    if (lo >= hi) return;
    if (threadIdx.x == 0) {  // last run whose offset is <= lo
        uint64_t a = 0, b = n_runs;
        while (a < b) {
            uint64_t m = (a + b + 1) >> 1;
            if (m < n_runs && run_offsets[m] <= lo) a = m;
            else b = m - 1;
        }
        s_first = a;
    }
    __syncthreads();
    for (uint64_t r = s_first; r < n_runs && run_offsets[r] < hi; ++r) {
        const uint64_t a = max(run_offsets[r], lo), b = min(run_offsets[r + 1], hi);
        const uint32_t c = run_chr[r];
        for (uint64_t i = a + threadIdx.x; i < b; i += blockDim.x) out[i - q0] = c;
    }
}

int32_t launch_expand_runs(gtgpu_ctx* ctx, uint64_t n_runs, const uint64_t* d_run_offsets, const uint32_t* d_run_chr, uint64_t q0,
                           uint64_t cn, uint32_t* d_out) {
    if (cn == 0) return GTGPU_OK;
    int grid = (int)std::min<uint64_t>((cn + 8191) / 8192, (uint64_t)ctx->sm_count * 16);
    expand_runs_kernel<<<grid, 256, 0, ctx->stream>>>(n_runs, d_run_offsets, d_run_chr, q0, cn, d_out);
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// end = start + 16-bit width for a chunk of queries, then the exceptions (queries whose width does not fit, or whose end
// precedes their start) are patched from the caller's list: 2 bytes per query cross PCIe instead of 4.
__global__ void expand_widths_kernel(uint64_t n, const uint32_t* __restrict__ start, const uint16_t* __restrict__ w16,
                                     uint32_t* __restrict__ end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) end[i] = start[i] + (uint32_t)w16[i];
}
__global__ void patch_wide_kernel(uint64_t n_wide, const uint64_t* __restrict__ wide_index, const uint32_t* __restrict__ wide_end,
                                  uint64_t q0, uint32_t* __restrict__ end) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_wide) end[wide_index[i] - q0] = wide_end[i];
}
int32_t launch_expand_widths(gtgpu_ctx* ctx, uint64_t n, const uint32_t* d_start, const uint16_t* d_w16, uint32_t* d_end,
                             uint64_t n_wide, const uint64_t* d_wide_index, const uint32_t* d_wide_end, uint64_t q0) {
    if (n) {
        const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
        expand_widths_kernel<<<grid, 256, 0, ctx->stream>>>(n, d_start, d_w16, d_end);
        ctx->launches++;
    }
    if (n_wide) {
        patch_wide_kernel<<<(unsigned)((n_wide + 255) / 256), 256, 0, ctx->stream>>>(n_wide, d_wide_index, d_wide_end, q0, d_end);
        ctx->launches++;
    }
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// Packed wire format (gtgpu_tokenize_files_packed): one 32-bit word per query = offset from the anchor of its 32-query block
// | width << off_bits.  `packed` may alias `start` (every thread reads its word before it writes).  Exceptions carry their
// absolute (start, end).  4 bytes per query (+ 1/8 for the anchors) cross PCIe instead of 6.
__global__ void expand_packed_kernel(uint64_t n, const uint32_t* packed, const uint32_t* __restrict__ anchors, uint32_t off_bits,
                                     uint32_t* start, uint32_t* __restrict__ end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t off_mask = (1u << off_bits) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t w = packed[i];
        const uint32_t s = anchors[i >> 5] + (w & off_mask);
        start[i] = s;
        end[i] = s + (w >> off_bits);
    }
}
__global__ void patch_exceptions_kernel(uint64_t n_exc, const uint64_t* __restrict__ exc_index, const uint32_t* __restrict__ exc_start,
                                        const uint32_t* __restrict__ exc_end, uint64_t q0, uint32_t* __restrict__ start,
                                        uint32_t* __restrict__ end) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_exc) {
        start[exc_index[i] - q0] = exc_start[i];
        end[exc_index[i] - q0] = exc_end[i];
    }
}
int32_t launch_expand_packed(gtgpu_ctx* ctx, uint64_t n, const uint32_t* d_packed, const uint32_t* d_anchors, uint32_t width_bits,
                             uint32_t* d_start, uint32_t* d_end, uint64_t n_exc, const uint64_t* d_exc_index,
                             const uint32_t* d_exc_start, const uint32_t* d_exc_end, uint64_t q0) {
    if (n) {
        const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
        expand_packed_kernel<<<grid, 256, 0, ctx->stream>>>(n, d_packed, d_anchors, 32u - width_bits, d_start, d_end);
        ctx->launches++;
    }
    if (n_exc) {
        patch_exceptions_kernel<<<(unsigned)((n_exc + 255) / 256), 256, 0, ctx->stream>>>(n_exc, d_exc_index, d_exc_start, d_exc_end, q0,
                                                                                           d_start, d_end);
        ctx->launches++;
    }
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// ================================================================================================================
// per-call [unk] rule (tokenizer.rs:158-160): a file whose raw id run is empty becomes the single id unk
// ================================================================================================================
// Single block: out_tok[f] = raw_tok[f] + #empty files before f  (n_files is small next to the query count).
__global__ void unk_offsets_kernel(uint64_t n_files, const uint64_t* __restrict__ raw_tok, uint64_t* __restrict__ out_tok,
                                   uint64_t* __restrict__ n_empty) {
    __shared__ uint64_t s_carry;
    __shared__ uint32_t s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base <= n_files; base += blockDim.x) {
        uint64_t f = base + tid;
        uint32_t empty = (f < n_files) ? (raw_tok[f + 1] == raw_tok[f]) : 0;
        uint32_t incl = empty;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t wex = 0, tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            if (w < warp) wex += s_w[w];
            tot += s_w[w];
        }
        uint64_t carry = s_carry;
        if (f <= n_files) out_tok[f] = raw_tok[f] + carry + wex + incl - empty;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (tid == 0) *n_empty = s_carry;
}

__global__ void unk_expand_kernel(uint64_t n_files, const uint64_t* __restrict__ raw_tok,
                                  const uint64_t* __restrict__ out_tok, const uint32_t* __restrict__ raw_ids,
                                  uint32_t unk_id, uint32_t* __restrict__ out_ids) {
    for (uint64_t f = blockIdx.x; f < n_files; f += gridDim.x) {
        uint64_t r0 = raw_tok[f], r1 = raw_tok[f + 1], o0 = out_tok[f];
        if (r1 == r0) {
            if (threadIdx.x == 0) out_ids[o0] = unk_id;
        } else {
            for (uint64_t j = threadIdx.x; j < r1 - r0; j += blockDim.x) out_ids[o0 + j] = raw_ids[r0 + j];
        }
    }
}

int32_t launch_unk_offsets(gtgpu_ctx* ctx, uint64_t n_files, const uint64_t* d_raw_file_tok, uint64_t* d_out_file_tok,
                           uint64_t* d_n_empty) {
    unk_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(n_files, d_raw_file_tok, d_out_file_tok, d_n_empty);
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

int32_t launch_unk_expand(gtgpu_ctx* ctx, uint64_t n_files, const uint64_t* d_raw_file_tok,
                          const uint64_t* d_out_file_tok, const uint32_t* d_raw_ids, uint32_t unk_id,
                          uint32_t* d_out_ids) {
    if (n_files == 0) return GTGPU_OK;
    int grid = (int)std::min<uint64_t>(n_files, (uint64_t)ctx->sm_count * 8);
    unk_expand_kernel<<<grid, 256, 0, ctx->stream>>>(n_files, d_raw_file_tok, d_out_file_tok, d_raw_ids, unk_id, d_out_ids);
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

}  // namespace gtgpu
