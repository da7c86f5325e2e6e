// fused_ws.cuh — warp-specialised variant of the fused find kernel (included by kernels.cu).
//
// Same algorithm and same status-word protocol as fused_find_kernel, but the two halves of a tile's life run in
// different warps of the CTA and only meet in shared memory:
//   * 8 RESOLVER warps: TMA-staged queries -> window table -> counts / vals -> warp scan -> tile aggregate; they write
//     (count, tile-local offset, vals) of every query into a result slot and move on to the next tile at once;
//   * 2 EMITTER warps: wait for a full slot, run the decoupled look-back for that tile (256 predecessors per round:
//     4 status words per emitter thread), publish the inclusive prefix, store the ids (and per-query / per-file
//     offsets), and hand the slot back.
// The latency chain of a tile (two dependent L2 gathers, then an L2 round trip for the look-back, then the stores) is
// what bounds the single-role kernel at 4 CTAs/SM; here the resolve chain and the look-back/emit chain overlap inside
// a CTA, three slots deep, so neither waits for the other.  Emitters handle tile i only after tile i+1 has been
// resolved (one tile of lag), by which time the predecessors' aggregates are published.
#pragma once

namespace gtgpu {

constexpr int WS_RESOLVERS = 256;                 // 8 warps
constexpr int WS_EMITTERS = 64;                   // 2 warps
constexpr int WS_THREADS = WS_RESOLVERS + WS_EMITTERS;
constexpr int WS_ROWS = 4;
constexpr int WS_TILE = WS_RESOLVERS * WS_ROWS;   // 1 024 queries, same tiles as the single-role kernel
constexpr int WS_SLOTS = 3;
static_assert(WS_TILE == FUSED_TILE, "both kernels share the workspace layout (status words per 1 024-query tile)");

struct WsSmem {
    uint32_t q[2][3][WS_TILE];          // TMA-staged chr / start / end rows, double-buffered
    uint32_t v0[WS_SLOTS][WS_TILE];     // first hit's val
    uint32_t v1[WS_SLOTS][WS_TILE];     // second hit's val
    uint32_t co[WS_SLOTS][WS_TILE];     // count (bits 0-1, 3 = "3 or more") | generic flag (bit 2) | tile-local offset << 3
    uint2 chrom[CHROM_CACHE];
    unsigned long long qbar[2];         // TMA arrival
    unsigned long long full[WS_SLOTS];  // resolvers -> emitters
    unsigned long long empty[WS_SLOTS]; // emitters -> resolvers
    uint32_t wtot[2][WS_RESOLVERS / 32];
    uint32_t tile[2];
    uint32_t staged[2];
    uint32_t slot_tile[WS_SLOTS];
    uint32_t slot_agg[WS_SLOTS];
    unsigned long long lb_sum[WS_EMITTERS / 32];
    uint32_t lb_p[WS_EMITTERS / 32];
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

#ifdef GT_PHASE_TIMING
#define WS_MARK(i)                                                                   \
    do {                                                                             \
        long long _now = clock64();                                                  \
        if (lane == 0) atomicAdd(&s_wsacc[i], (unsigned long long)(_now - _t));      \
        _t = _now;                                                                   \
    } while (0)
#define WS_START() long long _t = clock64()
#else
#define WS_MARK(i)
#define WS_START()
#endif

template <bool DESC, bool FILTER, bool OFFS>
__global__ void __launch_bounds__(WS_THREADS, 3)
fused_find_ws_kernel(IndexView ix, uint64_t n, uint32_t n_tiles, uint64_t n_files, const uint64_t* __restrict__ file_offsets,
                     const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                     int32_t min_bp, int tma_ok, uint32_t* __restrict__ out_ids, uint64_t capacity,
                     uint64_t* __restrict__ out_offsets, uint64_t* __restrict__ out_file_tok, FusedWorkspace ws,
                     const uint64_t* __restrict__ d_base, uint64_t* __restrict__ d_total, uint32_t* __restrict__ d_err) {
    constexpr int ROWS = WS_ROWS, TILE = WS_TILE, WTILE = 32 * WS_ROWS, RWARPS = WS_RESOLVERS / 32;
    constexpr uint32_t FULL = 0xFFFFFFFFu, NO_TILE = 0xFFFFFFFFu;
    extern __shared__ __align__(128) unsigned char ws_smem_raw[];
    WsSmem& sm = *reinterpret_cast<WsSmem*>(ws_smem_raw);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
#ifdef GT_PHASE_TIMING
    __shared__ unsigned long long s_wsacc[16];
    if (tid < 16) s_wsacc[tid] = 0;
#endif
    const uint32_t nchr = ix.n_chroms;
    const bool chrom_cached = nchr < CHROM_CACHE;
    const uint64_t base = d_base ? *d_base : 0;
    unsigned long long* status = reinterpret_cast<unsigned long long*>(ws.status);
    const uint32_t shift = ix.bt_shift;

    auto claim_and_stage = [&](uint32_t buf) {
        const uint32_t t = atomicAdd(ws.counter, 1u);
        sm.tile[buf] = t;
        uint32_t staged = 0;
        if (tma_ok && t < n_tiles && (uint64_t)(t + 1) * TILE <= n) {
            staged = 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(reinterpret_cast<uint64_t*>(&sm.qbar[buf]), 3 * TILE * 4);
            const uint64_t q0 = (uint64_t)t * TILE;
            const uint64_t pol = policy_evict_first();
            bulk_g2s(&sm.q[buf][0][0], chr + q0, TILE * 4, reinterpret_cast<uint64_t*>(&sm.qbar[buf]), pol);
            bulk_g2s(&sm.q[buf][1][0], start + q0, TILE * 4, reinterpret_cast<uint64_t*>(&sm.qbar[buf]), pol);
            bulk_g2s(&sm.q[buf][2][0], end + q0, TILE * 4, reinterpret_cast<uint64_t*>(&sm.qbar[buf]), pol);
        }
        sm.staged[buf] = staged;
    };

    for (uint32_t i = tid; i < CHROM_CACHE; i += WS_THREADS)
        sm.chrom[i] = i < nchr ? __ldg(reinterpret_cast<const uint2*>(ix.chrom_bt + i)) : make_uint2(0, 0);
    if (tid == 0) {
        mbar_init(reinterpret_cast<uint64_t*>(&sm.qbar[0]), 1);
        mbar_init(reinterpret_cast<uint64_t*>(&sm.qbar[1]), 1);
        for (int s = 0; s < WS_SLOTS; ++s) {
            mbar_init(reinterpret_cast<uint64_t*>(&sm.full[s]), WS_RESOLVERS);
            mbar_init(reinterpret_cast<uint64_t*>(&sm.empty[s]), WS_EMITTERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        claim_and_stage(0);
    }
    __syncthreads();

    if (tid < WS_RESOLVERS) {
        // ======================================== RESOLVER WARPS ========================================
        const uint32_t warp = tid >> 5;
        const uint32_t wl = warp * WTILE + lane;
        const uint64_t keep = policy_evict_last();
        uint32_t qphase = 0;
        for (uint32_t it = 0;; ++it) {
            WS_START();
            const uint32_t qb = it & 1, slot = it % WS_SLOTS;
            uint32_t tile = sm.tile[qb];
            if (tile >= n_tiles) tile = NO_TILE;
            // the slot must have been drained by the emitters (its previous use was iteration it - WS_SLOTS)
            if (it >= WS_SLOTS) mbar_wait(reinterpret_cast<uint64_t*>(&sm.empty[slot]), ((it / WS_SLOTS) - 1) & 1);
            if (tile == NO_TILE) {
                if (tid == 0) sm.slot_tile[slot] = NO_TILE;
                mbar_arrive(&sm.full[slot]);
                break;
            }
            WS_MARK(0);
            const uint64_t tile_start = (uint64_t)tile * TILE;
            uint32_t qc[ROWS], qs[ROWS], qe[ROWS];
            if (sm.staged[qb]) {
                mbar_wait(reinterpret_cast<uint64_t*>(&sm.qbar[qb]), (qphase >> qb) & 1);
                qphase ^= 1u << qb;
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    qc[k] = sm.q[qb][0][wl + 32 * k];
                    qs[k] = sm.q[qb][1][wl + 32 * k];
                    qe[k] = sm.q[qb][2][wl + 32 * k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint64_t q = tile_start + wl + 32 * k;
                    const bool ok = q < n;
                    qc[k] = ok ? __ldcs(chr + q) : 0xFFFFFFFFu;
                    qs[k] = ok ? __ldcs(start + q) : 0;
                    qe[k] = ok ? __ldcs(end + q) : 0;
                }
            }
            if (qc[0] == 0x12345678u && qs[0] == 0x9abcdefu) atomicExch(d_err, 2u);
            WS_MARK(1);
            uint32_t slow = 0;
            uint32_t w[ROWS];
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                const uint32_t c = qc[k], s = qs[k], e = qe[k];
                uint2 cb;
                if (chrom_cached) cb = sm.chrom[min(c, (uint32_t)CHROM_CACHE - 1)];
                else cb = c < nchr ? __ldg(reinterpret_cast<const uint2*>(ix.chrom_bt + c)) : make_uint2(0, 0);
                const uint32_t b1 = s >> shift, b2 = (e - 1) >> shift;
                if ((cb.y == BT_GENERIC_CHROM) | (s >= e) | (b2 - b1 > 1)) slow |= 1u << k;
                const uint32_t li = b1 < (cb.y & BT_NBINS_MASK) ? cb.x + b1 : 0u;
                w[k] = ldg32_keep(ix.bt_lut + li, keep);
            }
            uint32_t cnt[ROWS], v0[ROWS], v1[ROWS];
            {
                uint4 E0[ROWS], E1[ROWS];
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    if (w[k] & BT_POOL_FLAG) {
                        slow |= 1u << k;
                        w[k] = 0;
                    }
                    E0[k] = make_uint4(0xFFFFFFFFu, 0, 0, 0);
                    E1[k] = make_uint4(0xFFFFFFFFu, 0, 0, 0);
                    const uint4* ep = ix.bt_ent + (w[k] >> 2);
                    if (w[k] & 3) E0[k] = ldg128_keep(ep, keep);
                    if ((w[k] & 3) == 2) E1[k] = ldg128_keep(ep + 1, keep);
                }
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    const uint32_t s = qs[k], e = qe[k];
                    const bool h0 = cand_hit<FILTER>(E0[k].x, E0[k].y, s, e, min_bp);
                    const bool h1 = cand_hit<FILTER>(E1[k].x, E1[k].y, s, e, min_bp);
                    cnt[k] = (uint32_t)h0 + (uint32_t)h1;
                    v0[k] = h0 ? E0[k].z : E1[k].z;
                    v1[k] = E1[k].z;
                }
            }
            WS_MARK(2);
            // ---- warp scan -----------------------------------------------------------------------------------------
            uint32_t off[ROWS];
            uint32_t warp_total = 0;
            if (!__any_sync(FULL, slow != 0)) {
                const uint32_t mine = cnt[0] | (cnt[1] << 8) | (cnt[2] << 16) | (cnt[3] << 24);
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t t = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += t;
                }
                const uint32_t tot = __shfl_sync(FULL, incl, 31);
                const uint32_t excl = incl - mine;
                const uint32_t t0 = tot & 0xFF, t1 = (tot >> 8) & 0xFF, t2 = (tot >> 16) & 0xFF, t3 = tot >> 24;
                off[0] = excl & 0xFF;
                off[1] = t0 + ((excl >> 8) & 0xFF);
                off[2] = t0 + t1 + ((excl >> 16) & 0xFF);
                off[3] = t0 + t1 + t2 + (excl >> 24);
                warp_total = t0 + t1 + t2 + t3;
            } else {
#pragma unroll
                for (int k = 0; k < ROWS; ++k)
                    if ((slow >> k) & 1) cnt[k] = count_query_walk_noinline(ix, qc[k], qs[k], qe[k], min_bp);
                uint64_t wide = 0;
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    uint32_t incl = cnt[k];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        uint32_t t = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += t;
                    }
                    off[k] = warp_total + incl - cnt[k];
                    warp_total += __shfl_sync(FULL, incl, 31);
                    wide += cnt[k];
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) wide += __shfl_down_sync(FULL, wide, d);
                if (lane == 0 && wide > 0x0FFFFFFFull) atomicExch(d_err, 1u);  // tile-local offsets must fit 29 bits
            }
            if (lane == 0) sm.wtot[qb][warp] = warp_total;
            WS_MARK(3);
            named_bar_sync(1, WS_RESOLVERS);  // warp totals visible; everyone is done with sm.tile[qb] / sm.q[qb ^ 1]
            WS_MARK(4);
            uint32_t warp_excl = 0, tile_agg = 0;
#pragma unroll
            for (int ww = 0; ww < RWARPS; ++ww) {
                const uint32_t t = sm.wtot[qb][ww];
                if (ww < (int)warp) warp_excl += t;
                tile_agg += t;
            }
            if (tid == 0) {
                st_status(reinterpret_cast<uint64_t*>(status) + tile, ST_FLAG_AGG | (uint64_t)tile_agg);
                sm.slot_tile[slot] = tile;
                sm.slot_agg[slot] = tile_agg;
                claim_and_stage(qb ^ 1);
            }
            // ---- results into the slot ----------------------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                const uint32_t i = wl + 32 * k;
                const bool g = (slow >> k) & 1;
                sm.co[slot][i] = min(cnt[k], 3u) | (g ? 4u : 0u) | ((warp_excl + off[k]) << 3);
                if (cnt[k] != 0 && !g) {
                    sm.v0[slot][i] = v0[k];
                    if (cnt[k] == 2) sm.v1[slot][i] = v1[k];
                }
            }
            mbar_arrive(&sm.full[slot]);
            WS_MARK(5);
            named_bar_sync(1, WS_RESOLVERS);  // sm.tile[qb ^ 1] (next tile) visible to all resolvers
            WS_MARK(6);
        }
    } else {
        // ========================================= EMITTER WARPS =========================================
        const uint32_t et = tid - WS_RESOLVERS;  // 0..63
        const uint32_t ewarp = et >> 5;
        for (uint32_t it = 0;; ++it) {
            WS_START();
            const uint32_t slot = it % WS_SLOTS;
            mbar_wait(reinterpret_cast<uint64_t*>(&sm.full[slot]), (it / WS_SLOTS) & 1);
            const uint32_t tile = sm.slot_tile[slot];
            if (tile == NO_TILE) break;
            WS_MARK(8);
            const uint32_t tile_agg = sm.slot_agg[slot];
            // one tile of lag: wait until the next tile has been resolved too (or the resolvers are done)
            mbar_wait(reinterpret_cast<uint64_t*>(&sm.full[(it + 1) % WS_SLOTS]), ((it + 1) / WS_SLOTS) & 1);

            WS_MARK(9);
            // ---- look-back: 4 status words per emitter thread, nearest first -> 256 predecessors per round -------
            uint64_t excl = 0;
            for (int64_t win = (int64_t)tile - 1;; win -= 4 * WS_EMITTERS) {
                const int64_t j0 = win - 4 * (int64_t)et;
                uint64_t v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = j0 - k >= 0 ? ld_status(reinterpret_cast<uint64_t*>(status) + (j0 - k)) : ST_FLAG_PREFIX;
                for (;;) {
                    bool missing = false;
#pragma unroll
                    for (int k = 0; k < 4; ++k) missing |= (v[k] >> 62) == 0;
                    if (!__any_sync(FULL, missing)) break;
                    if (missing) {
                        __nanosleep(64);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if ((v[k] >> 62) == 0) v[k] = ld_status(reinterpret_cast<uint64_t*>(status) + (j0 - k));
                    }
                }
                uint64_t val = 0;
                bool hasp = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!hasp) {
                        val += v[k] & ST_MASK;
                        hasp = (v[k] >> 62) == 2;
                    }
                }
                const uint32_t pmask = __ballot_sync(FULL, hasp);
                if (pmask && lane > (uint32_t)(__ffs(pmask) - 1)) val = 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(FULL, val, d);
                if (lane == 0) {
                    sm.lb_sum[ewarp] = val;
                    sm.lb_p[ewarp] = pmask != 0;
                }
                named_bar_sync(2, WS_EMITTERS);
                bool done = false;
#pragma unroll
                for (int ww = 0; ww < WS_EMITTERS / 32; ++ww) {
                    if (!done) {
                        excl += sm.lb_sum[ww];
                        done = sm.lb_p[ww] != 0;
                    }
                }
                named_bar_sync(2, WS_EMITTERS);  // partials may be rewritten
                if (done) break;
            }
            WS_MARK(10);
            const uint64_t tile_start = (uint64_t)tile * TILE;
            const uint64_t tile_base = base + excl;
            if (et == 0) {
                st_status(reinterpret_cast<uint64_t*>(status) + tile, ST_FLAG_PREFIX | (excl + tile_agg));
                if (tile == n_tiles - 1) {
                    *d_total = tile_base + tile_agg;
                    if (OFFS) out_offsets[n] = tile_base + tile_agg;
                }
            }
            // ---- emit: consecutive emitter threads take consecutive queries ----------------------------------------------
            const bool fits = tile_base + tile_agg <= capacity;
            uint32_t* const outp = out_ids + tile_base;
            for (uint32_t i = et; i < (uint32_t)TILE; i += WS_EMITTERS) {
                const uint32_t co = sm.co[slot][i];
                const uint32_t c = co & 3, o = co >> 3;
                const uint64_t q = tile_start + i;
                if (OFFS && q < n) out_offsets[q] = tile_base + o;
                if (c == 0) continue;
                if (co & 4) {
                    emit_query_walk(ix, __ldg(chr + q), __ldg(start + q), __ldg(end + q), min_bp, out_ids, tile_base + o, capacity);
                    continue;
                }
                uint32_t a = sm.v0[slot][i], b = 0;
                if (c == 2) {
                    b = sm.v1[slot][i];
                    if (DESC) { const uint32_t t = a; a = b; b = t; }
                }
                if (fits) {
                    __stcs(outp + o, a);
                    if (c == 2) __stcs(outp + o + 1, b);
                } else {
                    if (tile_base + o < capacity) out_ids[tile_base + o] = a;
                    if (c == 2 && tile_base + o + 1 < capacity) out_ids[tile_base + o + 1] = b;
                }
            }
            WS_MARK(11);
            // ---- file boundaries inside this tile ------------------------------------------------------------------------
            if (out_file_tok) {
                const uint32_t mark = __ldg(ws.tile_file + tile);
                if (mark != 0) {
                    const uint64_t limit = (tile == n_tiles - 1) ? n + 1 : tile_start + TILE;
                    for (uint64_t f = (uint64_t)(0xFFFFFFFFu - mark) + et; f <= n_files; f += WS_EMITTERS) {
                        const uint64_t qi = file_offsets[f];
                        if (qi >= limit) break;
                        const uint32_t r = (uint32_t)(qi - tile_start);
                        out_file_tok[f] = tile_base + (r < (uint32_t)TILE ? sm.co[slot][r] >> 3 : tile_agg);
                    }
                }
            }
            mbar_arrive(&sm.empty[slot]);
            WS_MARK(12);
        }
    }
#ifdef GT_PHASE_TIMING
    __syncthreads();
    if (tid < 16) atomicAdd(&g_phase_cycles[tid], s_wsacc[tid]);
#endif
}

}  // namespace gtgpu
