// kernels.cu — sm_100a kernels for count / any / find / tokenize.
//
// Nothing here is a dense contraction, so tensor cores are deliberately unused: the work is HBM-bound integer
// search + stream compaction.  What matters is (1) 128-bit coalesced query loads (blocked 4 queries / thread),
// (2) O(1) searches through L2-resident bin LUTs instead of 17–26-level bisections, (3) a single pass over the
// queries for enumeration: count → block scan → decoupled look-back across tiles → emit, so no per-query
// count/offset array ever round-trips through HBM, and (4) a persistent grid sized to the SM count.
#include <algorithm>

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
                uint64_t last = lut_lower_bound(ix.cs_starts + cm.off, ix.lut + cm.lut_cs, cm.nb_cs, cm.len, ix.shift, e);
                uint64_t first = lut_lower_bound(ix.cs_ends + cm.off, ix.lut + cm.lut_ce, cm.nb_ce, cm.len, ix.shift, s + 1u);
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
                    uint32_t last = lut_lower_bound(ix.cs_starts + cm.off, ix.lut + cm.lut_cs, cm.nb_cs, cm.len, ix.shift, e);
                    uint32_t first = lut_lower_bound(ix.cs_ends + cm.off, ix.lut + cm.lut_ce, cm.nb_ce, cm.len, ix.shift, s + 1u);
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

int32_t launch_count(gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start,
                     const uint32_t* d_end, int32_t min_overlap, int mode, void* d_out) {
    if (n == 0) return GTGPU_OK;
    gtgpu_ctx* ctx = ix->ctx;
    uint64_t blocks_needed = (n + 255) / 256;
    int grid = (int)std::min<uint64_t>(blocks_needed, (uint64_t)ctx->sm_count * 32);
    ctx->time_begin();
    switch (mode) {
        case COUNT_U32:
            count_kernel<COUNT_U32><<<grid, 256, 0, ctx->stream>>>(ix->view, n, d_chr, d_start, d_end, min_overlap, d_out);
            break;
        case COUNT_ANY_U8:
            count_kernel<COUNT_ANY_U8><<<grid, 256, 0, ctx->stream>>>(ix->view, n, d_chr, d_start, d_end, min_overlap, d_out);
            break;
        default:
            count_kernel<COUNT_BITS_RAW_U64><<<grid, 256, 0, ctx->stream>>>(ix->view, n, d_chr, d_start, d_end, min_overlap, d_out);
            break;
    }
    ctx->time_end();
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

// ================================================================================================================
// fused find: count → block scan → decoupled look-back → emit
// ================================================================================================================
#define ST_FLAG_AGG (1ull << 62)
#define ST_FLAG_PREFIX (2ull << 62)
#define ST_MASK ((1ull << 62) - 1)
#define MULTI_SEG 0xFFFFFFFFu

struct FusedWorkspace {
    uint64_t* status;      // [n_tiles] flag<<62 | value
    uint32_t* tile_file;   // [n_tiles] 0 = no file boundary in this tile, else 0xFFFFFFFF - first file index
    uint32_t* counter;     // dynamic tile counter
};

static inline uint64_t n_tiles_for(uint64_t n) { return (n + FUSED_TILE - 1) / FUSED_TILE; }

size_t fused_workspace_bytes(uint64_t n) {
    uint64_t t = n_tiles_for(n);
    return (size_t)(t * 8 + ((t * 4 + 7) / 8) * 8 + 64);
}

static FusedWorkspace carve(void* ws, uint64_t n) {
    uint64_t t = n_tiles_for(n);
    FusedWorkspace w;
    w.status = reinterpret_cast<uint64_t*>(ws);
    w.tile_file = reinterpret_cast<uint32_t*>(w.status + t);
    w.counter = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + t * 8 + ((t * 4 + 7) / 8) * 8);
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

template <int BLOCK, int ITEMS>
__global__ void __launch_bounds__(BLOCK)
fused_find_kernel(IndexView ix, uint64_t n, uint32_t n_tiles, uint64_t n_files, const uint64_t* __restrict__ file_offsets,
                  const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                  int32_t min_bp, int vec_ok, uint32_t* __restrict__ out_ids, uint64_t capacity,
                  uint64_t* __restrict__ out_offsets, uint64_t* __restrict__ out_file_tok, FusedWorkspace ws,
                  const uint64_t* __restrict__ d_base, uint64_t* __restrict__ d_total, uint32_t* __restrict__ d_err) {
    constexpr int TILE = BLOCK * ITEMS;
    constexpr int WARPS = BLOCK / 32;
    static_assert(ITEMS == 4, "query loads are written for one uint4 per array per thread");
    __shared__ uint32_t s_qoff[TILE + 1];
    __shared__ uint32_t s_warp[WARPS];
    __shared__ uint64_t s_tile_excl;
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = d_base ? *d_base : 0;
    volatile uint64_t* status = ws.status;

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(ws.counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint64_t tile_start = (uint64_t)tile * TILE;
        const uint32_t first = tid * ITEMS;

        // ---- load 4 queries per thread -----------------------------------------------------------------------
        uint32_t qc[ITEMS], qs[ITEMS], qe[ITEMS];
        if (vec_ok && tile_start + TILE <= n) {
            uint4 c4 = __ldcs(reinterpret_cast<const uint4*>(chr + tile_start) + tid);
            uint4 s4 = __ldcs(reinterpret_cast<const uint4*>(start + tile_start) + tid);
            uint4 e4 = __ldcs(reinterpret_cast<const uint4*>(end + tile_start) + tid);
            qc[0] = c4.x; qc[1] = c4.y; qc[2] = c4.z; qc[3] = c4.w;
            qs[0] = s4.x; qs[1] = s4.y; qs[2] = s4.z; qs[3] = s4.w;
            qe[0] = e4.x; qe[1] = e4.y; qe[2] = e4.z; qe[3] = e4.w;
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                uint64_t q = tile_start + first + k;
                bool ok = q < n;
                qc[k] = ok ? __ldg(chr + q) : 0xFFFFFFFFu;
                qs[k] = ok ? __ldg(start + q) : 0;
                qe[k] = ok ? __ldg(end + q) : 0;
            }
        }

        // ---- resolve: candidate range + hit count per query -----------------------------------------------------
        uint32_t lo[ITEMS], ub[ITEMS], cnt[ITEMS];
        uint32_t mono_bits = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            lo[k] = ub[k] = cnt[k] = 0;
            if (qc[k] < ix.n_chroms) {
                const uint2 sr = __ldg(reinterpret_cast<const uint2*>(ix.chroms + qc[k]));
                if (sr.y - sr.x == 1) {
                    SegMeta m = load_seg(ix, sr.x);
                    seg_range(ix, m, qs[k], qe[k], lo[k], ub[k]);
                    cnt[k] = count_range(ix, lo[k], ub[k], qs[k], qe[k], min_bp, m.mono != 0);
                    mono_bits |= (m.mono != 0) << k;
                } else if (sr.y > sr.x) {
                    cnt[k] = count_query_walk(ix, qc[k], qs[k], qe[k], min_bp);
                    ub[k] = MULTI_SEG;
                }
            }
        }

        // ---- block exclusive scan of per-thread totals ---------------------------------------------------------
        uint64_t wide = (uint64_t)cnt[0] + cnt[1] + cnt[2] + cnt[3];
        uint32_t thread_total = (uint32_t)wide;
        uint32_t incl = thread_total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        // 64-bit block total, only to detect tiles whose local offsets would not fit 32 bits.
        uint64_t wsum = wide;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wsum += __shfl_down_sync(0xFFFFFFFFu, wsum, d);
        if (lane == 0 && wsum > 0xFFFFFFFFull) atomicExch(d_err, 1u);
        __syncthreads();
        uint32_t warp_excl = 0, tile_agg = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t t = s_warp[w];
            if (w < warp) warp_excl += t;
            tile_agg += t;
        }
        uint32_t running = warp_excl + incl - thread_total;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            s_qoff[first + k] = running;
            running += cnt[k];
        }
        if (tid == BLOCK - 1) s_qoff[TILE] = tile_agg;

        // ---- decoupled look-back (warp 0): exclusive prefix of this tile ----------------------------------------
        if (warp == 0) {
            if (lane == 0) status[tile] = ST_FLAG_AGG | (uint64_t)tile_agg;
            uint64_t excl = 0;
            int64_t j = (int64_t)tile - 1 - lane;
            for (;;) {
                uint64_t v = j >= 0 ? status[j] : ST_FLAG_PREFIX;
                while (__any_sync(0xFFFFFFFFu, (v >> 62) == 0)) {
                    if ((v >> 62) == 0) v = status[j];
                }
                uint32_t pmask = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
                uint64_t val = v & ST_MASK;
                if (pmask && lane > (__ffs(pmask) - 1)) val = 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xFFFFFFFFu, val, d);
                val = __shfl_sync(0xFFFFFFFFu, val, 0);
                excl += val;
                if (pmask) break;
                j -= 32;
            }
            if (lane == 0) {
                status[tile] = ST_FLAG_PREFIX | (excl + tile_agg);
                s_tile_excl = excl;
                if (tile == n_tiles - 1) {
                    *d_total = base + excl + tile_agg;
                    if (out_offsets) out_offsets[n] = base + excl + tile_agg;
                }
            }
        }
        __syncthreads();
        const uint64_t tile_base = base + s_tile_excl;

        // ---- emit --------------------------------------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            uint64_t q = tile_start + first + k;
            uint64_t pos = tile_base + s_qoff[first + k];
            if (out_offsets && q < n) out_offsets[q] = pos;
            if (cnt[k] == 0) continue;
            const uint32_t s = qs[k], e = qe[k];
            if (ub[k] != MULTI_SEG) {
                const bool mono = (mono_bits >> k) & 1;
                if (!ix.descending) {
                    for (uint32_t i = lo[k]; i < ub[k]; ++i)
                        if (is_hit(ix, i, s, e, min_bp, mono)) {
                            if (pos < capacity) out_ids[pos] = __ldg(ix.vals + i);
                            ++pos;
                        }
                } else {
                    for (uint32_t i = ub[k]; i > lo[k]; --i)
                        if (is_hit(ix, i - 1, s, e, min_bp, mono)) {
                            if (pos < capacity) out_ids[pos] = __ldg(ix.vals + i - 1);
                            ++pos;
                        }
                }
            } else {
                const uint2 sr = __ldg(reinterpret_cast<const uint2*>(ix.chroms + qc[k]));
                for (uint32_t si = sr.x; si < sr.y; ++si) {
                    SegMeta m = load_seg(ix, si);
                    uint32_t l, u;
                    seg_range(ix, m, s, e, l, u);
                    const bool mono = m.mono != 0;
                    if (!ix.descending) {
                        for (uint32_t i = l; i < u; ++i)
                            if (is_hit(ix, i, s, e, min_bp, mono)) {
                                if (pos < capacity) out_ids[pos] = __ldg(ix.vals + i);
                                ++pos;
                            }
                    } else {
                        for (uint32_t i = u; i > l; --i)
                            if (is_hit(ix, i - 1, s, e, min_bp, mono)) {
                                if (pos < capacity) out_ids[pos] = __ldg(ix.vals + i - 1);
                                ++pos;
                            }
                    }
                }
            }
        }

        // ---- file boundaries that fall into this tile: raw token offset of each file's first query -------------
        if (out_file_tok) {
            uint32_t mark = ws.tile_file[tile];
            if (mark != 0) {
                const uint64_t limit = (tile == n_tiles - 1) ? n + 1 : tile_start + TILE;
                for (uint64_t f = (uint64_t)(0xFFFFFFFFu - mark) + tid; f <= n_files; f += BLOCK) {
                    uint64_t qi = file_offsets[f];
                    if (qi >= limit) break;
                    out_file_tok[f] = tile_base + s_qoff[qi - tile_start];
                }
            }
        }
        __syncthreads();  // s_qoff / s_tile are reused by the next tile
    }
}

int32_t launch_fused_find(gtgpu_index* ix, uint64_t n, uint64_t n_files, const uint64_t* d_file_offsets,
                          const uint32_t* d_chr, const uint32_t* d_start, const uint32_t* d_end,
                          int32_t min_overlap, uint32_t* d_out_ids, uint64_t ids_capacity,
                          uint64_t* d_out_offsets, uint64_t* d_out_file_tok, void* d_workspace,
                          const uint64_t* d_base, uint64_t* d_total_out, uint32_t* d_errflag) {
    gtgpu_ctx* ctx = ix->ctx;
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
    FusedWorkspace ws = carve(d_workspace, n);
    GT_CUDA(cudaMemsetAsync(d_workspace, 0, fused_workspace_bytes(n), st));
    if (d_out_file_tok) {
        uint64_t cnt = n_files + 1;
        mark_file_tiles_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(n_files, d_file_offsets, n_tiles, ws.tile_file);
        ctx->launches++;
    }
    int vec_ok = ((reinterpret_cast<uintptr_t>(d_chr) | reinterpret_cast<uintptr_t>(d_start) |
                   reinterpret_cast<uintptr_t>(d_end)) & 15) == 0;
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        GT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fused_find_kernel<FUSED_BLOCK, FUSED_ITEMS>,
                                                              FUSED_BLOCK, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * blocks_per_sm);
    ctx->time_begin();
    fused_find_kernel<FUSED_BLOCK, FUSED_ITEMS><<<grid, FUSED_BLOCK, 0, st>>>(
        ix->view, n, n_tiles, n_files, d_file_offsets, d_chr, d_start, d_end, min_overlap, vec_ok, d_out_ids,
        ids_capacity, d_out_offsets, d_out_file_tok, ws, d_base, d_total_out, d_errflag);
    ctx->time_end();
    ctx->launches++;
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
