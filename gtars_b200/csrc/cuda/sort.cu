// sort.cu — hand-written device primitives: exclusive scan and a stable LSD radix sort of (key, value) pairs.
//
// Used by the barcode group-by of fragment tokenization (fragments.cu).  Both are plain multi-kernel designs (no
// single-pass tricks): they run once per call over arrays that are small next to the fused find kernel's traffic.
//
//   exclusive_scan<T>:  reduce per 2 048-element tile -> scan of the tile sums by one block -> per-tile rescan + offset.
//   radix_sort_pairs:   8 (or 9, when that saves a pass) bits per pass; per pass  (1) per-tile digit histograms, stored digit-major,
//                       (2) exclusive scan of that table = global start of every (digit, tile) bucket,
//                       (3) scatter: each warp owns 512 consecutive elements, ranks them round by round with
//                           __match_any_sync, so equal digits keep their input order (stable).
#include <algorithm>

#include "common.cuh"

namespace gtgpu {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp, T& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    T warp_excl = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const T t = s_warp[w];
        if (w < warp) warp_excl += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return warp_excl + incl - v;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T* __restrict__ in, T* __restrict__ tile_sums, uint64_t n) {
    __shared__ T s_warp[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    T sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) sum += in[i];
    }
    T total;
    block_exclusive_scan<T>(sum, s_warp, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(T* __restrict__ tile_sums, uint64_t n_tiles) {
    __shared__ T s_warp[SCAN_THREADS / 32];
    T carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        const uint64_t i = base + threadIdx.x;
        const T v = i < n_tiles ? tile_sums[i] : 0;
        T total;
        const T excl = block_exclusive_scan<T>(v, s_warp, total);
        if (i < n_tiles) tile_sums[i] = carry + excl;
        carry += total;
    }
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                  const T* __restrict__ tile_offsets, uint64_t n) {
    __shared__ T s_warp[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;  // blocked: 8 consecutive
    T v[SCAN_ITEMS];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        sum += v[k];
    }
    T total;
    T run = tile_offsets[blockIdx.x] + block_exclusive_scan<T>(sum, s_warp, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

size_t exclusive_scan_temp_bytes(uint64_t n, size_t elem) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE + 1) * elem; }

template <typename T>
int32_t exclusive_scan(gtgpu_ctx* ctx, const T* d_in, T* d_out, uint64_t n, void* d_temp) {
    if (n == 0) return GTGPU_OK;
    const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles > 0x7FFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "exclusive_scan: input too large");
    T* sums = reinterpret_cast<T*>(d_temp);
    scan_reduce_kernel<T><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, sums, n);
    scan_spine_kernel<T><<<1, SCAN_THREADS, 0, ctx->stream>>>(sums, tiles);
    scan_down_kernel<T><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, d_out, sums, n);
    ctx->launches += 3;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}
template int32_t exclusive_scan<uint32_t>(gtgpu_ctx*, const uint32_t*, uint32_t*, uint64_t, void*);
template int32_t exclusive_scan<unsigned long long>(gtgpu_ctx*, const unsigned long long*, unsigned long long*, uint64_t, void*);

// ---- inclusive running maximum (64-bit) ---------------------------------------------------------------------------------
// Same three-kernel shape as exclusive_scan with max as the operator.  The index builders use it for segmented running
// maxima: with the segment number in the high word and the value in the low word, a later segment always wins, so the
// low word of the scan is the running maximum inside the element's own segment.
__device__ __forceinline__ unsigned long long block_inclusive_max(unsigned long long v, unsigned long long* s_warp,
                                                                  unsigned long long& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl = max(incl, t);
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const unsigned long long t = s_warp[w];
        if (w < warp) before = max(before, t);
        tot = max(tot, t);
    }
    __syncthreads();
    total = tot;
    return max(before, incl);
}

__global__ void __launch_bounds__(SCAN_THREADS) max_reduce_kernel(const unsigned long long* __restrict__ in,
                                                                  unsigned long long* __restrict__ tile_max, uint64_t n) {
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    unsigned long long m = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) m = max(m, in[i]);
    }
    unsigned long long total;
    block_inclusive_max(m, s_warp, total);
    if (threadIdx.x == 0) tile_max[blockIdx.x] = total;
}

// tile_max[t] becomes the maximum over all tiles BEFORE t (0 for the first)
__global__ void __launch_bounds__(SCAN_THREADS) max_spine_kernel(unsigned long long* __restrict__ tile_max, uint64_t n_tiles) {
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prev[SCAN_THREADS];
    unsigned long long carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_max[i] : 0;
        unsigned long long total;
        const unsigned long long incl = block_inclusive_max(v, s_warp, total);
        s_prev[threadIdx.x] = incl;
        __syncthreads();
        const unsigned long long excl = threadIdx.x ? s_prev[threadIdx.x - 1] : 0;
        if (i < n_tiles) tile_max[i] = max(carry, excl);
        carry = max(carry, total);
        __syncthreads();
    }
}

// in and out may be the same array (every thread reads its own eight elements before it writes them): no __restrict__
__global__ void __launch_bounds__(SCAN_THREADS) max_down_kernel(const unsigned long long* in, unsigned long long* out,
                                                                const unsigned long long* __restrict__ tile_before, uint64_t n) {
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prev[SCAN_THREADS];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS], m = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        m = max(m, v[k]);
    }
    unsigned long long total;
    s_prev[threadIdx.x] = block_inclusive_max(m, s_warp, total);
    __syncthreads();
    unsigned long long run = max(tile_before[blockIdx.x], threadIdx.x ? s_prev[threadIdx.x - 1] : 0ull);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        run = max(run, v[k]);
        if (base + k < n) out[base + k] = run;
    }
}

// out[i] = max(in[0..i]); in and out may alias; d_temp needs exclusive_scan_temp_bytes(n, 8).
int32_t inclusive_max_scan_u64(gtgpu_ctx* ctx, const unsigned long long* d_in, unsigned long long* d_out, uint64_t n, void* d_temp) {
    if (n == 0) return GTGPU_OK;
    const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles > 0x7FFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "inclusive_max_scan: input too large");
    unsigned long long* tm = reinterpret_cast<unsigned long long*>(d_temp);
    max_reduce_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, tm, n);
    max_spine_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(tm, tiles);
    max_down_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, d_out, tm, n);
    ctx->launches += 3;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// ---- radix sort -----------------------------------------------------------------------------------------------------------
// lanes with equal digits: one ballot per digit bit (C5: 26.3 ms per 1e9 fragments) or __match_any_sync (27.6 ms)
#ifndef GT_RS_BALLOT
#define GT_RS_BALLOT 1
#endif
constexpr int RS_THREADS = 512;
constexpr int RS_ROUNDS = 16;                          // elements per lane
constexpr int RS_WARP_TILE = 32 * RS_ROUNDS;           // 512 consecutive elements per warp
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;        // 8 192 per block
constexpr int RS_WARPS = RS_THREADS / 32;

// DB = digit width of the pass: 8 bits, or 9 when that saves a whole pass (17-18 and 25-27 significant bits)
template <int DB>
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n_cap,
                                                                const uint64_t* __restrict__ d_n, int shift,
                                                                uint32_t n_tiles, uint32_t* __restrict__ hist) {
    constexpr uint32_t ND = 1u << DB;
    const uint64_t n = d_n ? min((uint64_t)*d_n, n_cap) : n_cap;  // element count known on the device only (no host sync)
    // one histogram per warp (no contention between warps; a warp's 32 lanes rarely share a digit), summed at the end
    __shared__ uint32_t s_hist[RS_WARPS / 2][ND];
    for (uint32_t i = threadIdx.x; i < (RS_WARPS / 2) * ND; i += RS_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
    uint32_t* mine = s_hist[(threadIdx.x >> 5) >> 1];
    // all sixteen loads of a thread in flight before the first atomic (four at a time left the DRAM at a quarter of its rate)
    uint32_t key[RS_ROUNDS];
#pragma unroll
    for (int k = 0; k < RS_ROUNDS; ++k) {
        const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
        key[k] = i < n ? __ldcs(keys + i) : 0;
    }
#pragma unroll
    for (int k = 0; k < RS_ROUNDS; ++k)
        if (base + (uint64_t)k * RS_THREADS + threadIdx.x < n) atomicAdd(&mine[(key[k] >> shift) & (ND - 1)], 1u);
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < ND; d += RS_THREADS) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS / 2; ++w) t += s_hist[w][d];
        hist[(uint64_t)d * n_tiles + blockIdx.x] = t;  // digit-major: one scan orders everything
    }
}

// Scatter of one pass.  A block sorts its 8 192-element tile by digit in SHARED memory first (stable: per-warp ranks from
// __match_any_sync round by round, warps in order), then writes the tile out digit by digit — a digit's elements of one
// tile go to consecutive addresses, so the global stores are coalesced runs (32-64 B per digit and tile on average)
// instead of 8 192 single words.  The first version stored every element straight from its ranking round: 16 scattered
// 4-byte stores per lane and array, 9.8 ms per 2.5e8 pairs; this one: see profiles/r02.
template <int DB>
__global__ void __launch_bounds__(RS_THREADS, 2) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                   const uint32_t* __restrict__ vals_in, uint64_t n_cap,
                                                                   const uint64_t* __restrict__ d_n, int shift,
                                                                   uint32_t n_tiles, const uint32_t* __restrict__ bucket_start,
                                                                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    constexpr uint32_t ND = 1u << DB;
    const uint64_t n = d_n ? min((uint64_t)*d_n, n_cap) : n_cap;
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t* s_key = s_mem;                        // [RS_TILE] the tile, sorted by digit
    uint32_t* s_val = s_mem + RS_TILE;              // [RS_TILE]
    uint32_t* s_cnt = s_mem + 2 * RS_TILE;          // [RS_WARPS][ND] per-warp digit counts, then each warp's first slot per digit
    uint32_t* s_dbase = s_cnt + RS_WARPS * ND;      // [ND] first slot of each digit in the sorted tile
    uint32_t* s_gbase = s_dbase + ND;               // [ND] global position of the digit's first element minus s_dbase
    __shared__ uint32_t s_wsum[RS_WARPS];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile0 = (uint64_t)blockIdx.x * RS_TILE;
    if (tile0 >= n) return;
    const uint32_t tile_n = (uint32_t)min((uint64_t)RS_TILE, n - tile0);
    for (uint32_t i = tid; i < RS_WARPS * ND; i += RS_THREADS) s_cnt[i] = 0;
    __syncthreads();
    // ---- 1. rank inside the warp, round by round (stable) ----------------------------------------------------------------
    const uint32_t wbase = warp * RS_WARP_TILE + lane;
    uint32_t key[RS_ROUNDS];
    uint16_t rank[RS_ROUNDS];
    uint32_t* my_cnt = s_cnt + warp * ND;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t j = wbase + 32 * r;
        key[r] = j < tile_n ? __ldcs(keys_in + tile0 + j) : 0;
    }
    // the sixteen matches do not depend on each other: issue them together (a match per round in front of that round's
    // counter update left every warp waiting on its result: a third of the kernel's stall samples) ...
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const bool ok = wbase + 32 * r < tile_n;
        const uint32_t d = ok ? (key[r] >> shift) & (ND - 1) : ND + lane;  // invalid lanes never match anyone
#if GT_RS_BALLOT
        uint32_t peers = __ballot_sync(0xFFFFFFFFu, ok);
#pragma unroll
        for (int b = 0; b < DB; ++b) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? bal : ~bal;
        }
        if (!ok) peers = 1u << lane;
#else
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
#endif
        rank[r] = (uint16_t)(__popc(peers & ((1u << lane) - 1)) | __popc(peers) << 8);  // lanes before me | group size
    }
    // ... then the per-warp digit counters advance round by round (stable)
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const bool ok = wbase + 32 * r < tile_n;
        const uint32_t d = (key[r] >> shift) & (ND - 1), before = rank[r] & 0xFFu, group = rank[r] >> 8;
        uint32_t prev = 0;
        if (ok) prev = my_cnt[d];
        __syncwarp();
        if (ok && before == 0) my_cnt[d] = prev + group;
        __syncwarp();
        rank[r] = (uint16_t)(prev + before);
    }
    __syncthreads();
    // ---- 2. digit totals -> first slot of every digit, and of every (warp, digit) ------------------------------------------
    uint32_t tot = 0;  // thread d: tile total of digit d (ND <= RS_THREADS)
    if (tid < ND) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = s_cnt[w * ND + tid];
            s_cnt[w * ND + tid] = run;  // elements of this digit in earlier warps
            run += c;
        }
        tot = run;
    }
    uint32_t incl = tot;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, k);
        if (lane >= (uint32_t)k) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (tid < ND) {
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += s_wsum[w];
        const uint32_t first = before + incl - tot;
        s_dbase[tid] = first;
        s_gbase[tid] = bucket_start[(uint64_t)tid * n_tiles + blockIdx.x] - first;  // wrapping on purpose
    }
    __syncthreads();
    // ---- 3. the tile, sorted by digit, in shared memory ------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        if (wbase + 32 * r < tile_n) {
            const uint32_t d = (key[r] >> shift) & (ND - 1);
            const uint32_t p = s_dbase[d] + my_cnt[d] + rank[r];
            s_key[p] = key[r];
            s_val[p] = __ldcs(vals_in + tile0 + wbase + 32 * r);  // values are only touched here: 16 fewer live registers while ranking
        }
    }
    __syncthreads();
    // ---- 4. write out: consecutive threads = consecutive slots = (inside a digit) consecutive global addresses ------------------
    for (uint32_t p = tid; p < tile_n; p += RS_THREADS) {
        const uint32_t k = s_key[p];
        const uint32_t dst = s_gbase[(k >> shift) & (ND - 1)] + p;
        keys_out[dst] = k;
        vals_out[dst] = s_val[p];
    }
}

template <int DB>
static constexpr size_t radix_scatter_smem() { return (size_t)(2 * RS_TILE + RS_WARPS * (1 << DB) + 2 * (1 << DB)) * 4; }

// passes and digit width for `bits` significant key bits: as few passes as 9-bit digits allow, 8-bit digits otherwise
void radix_plan(int bits, int* passes, int* width) {
    *passes = (bits + 8) / 9;
    *width = ((bits + *passes - 1) / *passes) <= 8 ? 8 : 9;
}

size_t radix_sort_temp_bytes(uint64_t n) {
    const uint64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    const uint64_t table = 512 * tiles;
    return (size_t)(table * 4 * 2 + exclusive_scan_temp_bytes(table, 4) + 256);
}

// Sorts by bits [0, bits) of the key.  keys/vals ping-pong between the a and b buffers; *result_in_b tells where the
// sorted data ended up.  n < 2^32.
// d_n (optional, device): the real element count when only the device knows it; n is then the capacity the grid is sized for.
int32_t radix_sort_pairs(gtgpu_ctx* ctx, uint64_t n, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                         int bits, void* d_temp, int* result_in_b, const uint64_t* d_n) {
    *result_in_b = 0;
    if (n == 0 || bits <= 0) return GTGPU_OK;
    if (n >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "radix_sort_pairs: too many elements");
    int passes, width;
    radix_plan(std::min(bits, 32), &passes, &width);
    // > 48 KB of dynamic shared memory is an opt-in per kernel and per device: set it on every call (cheap)
    GT_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)radix_scatter_smem<8>()));
    GT_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)radix_scatter_smem<9>()));
    const uint32_t tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    const uint64_t table = ((uint64_t)1 << width) * tiles;
    uint32_t* hist = reinterpret_cast<uint32_t*>(d_temp);
    uint32_t* starts = hist + 512ull * tiles;
    void* scan_tmp = starts + 512ull * tiles;
    uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
    for (int p = 0, shift = 0; p < passes; ++p, shift += width) {
        if (width == 8) radix_hist_kernel<8><<<tiles, RS_THREADS, 0, ctx->stream>>>(ki, n, d_n, shift, tiles, hist);
        else radix_hist_kernel<9><<<tiles, RS_THREADS, 0, ctx->stream>>>(ki, n, d_n, shift, tiles, hist);
        ctx->launches++;
        GT_TRY(exclusive_scan<uint32_t>(ctx, hist, starts, table, scan_tmp));
        if (width == 8)
            radix_scatter_kernel<8><<<tiles, RS_THREADS, radix_scatter_smem<8>(), ctx->stream>>>(ki, vi, n, d_n, shift, tiles, starts, ko, vo);
        else
            radix_scatter_kernel<9><<<tiles, RS_THREADS, radix_scatter_smem<9>(), ctx->stream>>>(ki, vi, n, d_n, shift, tiles, starts, ko, vo);
        ctx->launches++;
        std::swap(ki, ko);
        std::swap(vi, vo);
        *result_in_b ^= 1;
    }
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

}  // namespace gtgpu
