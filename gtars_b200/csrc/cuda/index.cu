// index.cu — host-side build of the device index (gtgpu_index_build).
//
// Order semantics of the reference are fixed here, once:
//   Bits   (gtars-overlaprs/src/bits.rs:101-128):   stable sort by (start,end); one segment per chromosome.
//   AIList (gtars-overlaprs/src/ailist.rs:105-151, 198-236): stable sort by start, then repeated peeling of
//          "long" intervals (>= 10 of the next 19 end earlier) into further components; no cap on components.
// Both get the same device layout: start-sorted SoA segments + running max of ends + bin LUTs, so the kernels
// locate candidates identically and only the emission direction differs.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "common.cuh"

namespace gtgpu {

namespace {

struct HostSeg {
    uint32_t chrom;
    std::vector<uint32_t> order;  // indices into the caller's arrays
};

// lut[b] = lower_bound(arr, b << shift) for b in [0, nb], nb = (max >> shift) + 1, lut[nb] = n.
void build_lut(const uint32_t* arr, uint32_t n, uint32_t shift, std::vector<uint32_t>& lut, uint32_t& lut_off,
               uint32_t& nb) {
    lut_off = (uint32_t)lut.size();
    if (n == 0) {
        nb = 0;
        lut.push_back(0);
        return;
    }
    nb = (arr[n - 1] >> shift) + 1;
    lut.resize(lut.size() + (size_t)nb + 1);
    uint32_t* L = lut.data() + lut_off;
    uint32_t i = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        uint64_t key = (uint64_t)b << shift;
        while (i < n && arr[i] < key) ++i;
        L[b] = i;
    }
    L[nb] = n;
}

uint64_t lut_entries(const std::vector<uint32_t>& maxima, uint32_t shift) {
    uint64_t t = 0;
    for (uint32_t m : maxima) t += ((uint64_t)m >> shift) + 2;
    return t;
}

template <class T>
int32_t upload(gtgpu_index* ix, const std::vector<T>& v, const T** out) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return fail(GTGPU_ERR_NOMEM, std::string("cudaMalloc(index): ") + cudaGetErrorString(e));
    ix->allocs.push_back(d);
    ix->device_bytes += bytes;
    if (!v.empty()) GT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return GTGPU_OK;
}

}  // namespace

}  // namespace gtgpu

using namespace gtgpu;

extern "C" int32_t gtgpu_index_build(gtgpu_ctx* ctx, int32_t kind, uint32_t n_chroms, const uint64_t* chrom_offsets,
                                     const uint32_t* starts, const uint32_t* ends, const uint32_t* vals,
                                     gtgpu_index** out_index) {
    if (!ctx || !out_index || (n_chroms && !chrom_offsets)) return fail(GTGPU_ERR_INVALID, "index_build: null argument");
    if (kind != GTGPU_KIND_BITS && kind != GTGPU_KIND_AILIST) return fail(GTGPU_ERR_INVALID, "index_build: bad kind");
    uint64_t total = n_chroms ? chrom_offsets[n_chroms] : 0;
    if (total >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: more than 2^32-2 intervals");
    if (total && (!starts || !ends)) return fail(GTGPU_ERR_INVALID, "index_build: null coordinate arrays");
    for (uint32_t c = 0; c < n_chroms; ++c)
        if (chrom_offsets[c] > chrom_offsets[c + 1]) return fail(GTGPU_ERR_INVALID, "index_build: chrom_offsets not monotone");
    GT_CUDA(cudaSetDevice(ctx->device));

    // ---- 1. per-chromosome ordering + AIList decomposition ------------------------------------------------
    std::vector<HostSeg> segs;
    std::vector<ChromMeta> chroms(n_chroms);
    uint64_t max_components = 0;
    bool proper = true;
    for (uint32_t c = 0; c < n_chroms; ++c) {
        uint64_t lo = chrom_offsets[c], hi = chrom_offsets[c + 1];
        chroms[c].seg_begin = (uint32_t)segs.size();
        if (hi > lo) {
            std::vector<uint32_t> order(hi - lo);
            std::iota(order.begin(), order.end(), (uint32_t)lo);
            if (kind == GTGPU_KIND_BITS) {
                std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                    return starts[a] != starts[b] ? starts[a] < starts[b] : ends[a] < ends[b];
                });
                segs.push_back(HostSeg{c, std::move(order)});
            } else {
                std::stable_sort(order.begin(), order.end(),
                                 [&](uint32_t a, uint32_t b) { return starts[a] < starts[b]; });
                const size_t min_cov = 10;
                std::vector<uint32_t> next;
                while (!order.empty()) {
                    HostSeg seg{c, {}};
                    next.clear();
                    for (size_t i = 0; i < order.size(); ++i) {
                        size_t covered = 0;
                        uint32_t e_i = ends[order[i]];
                        for (size_t j = 1; j < 2 * min_cov && i + j < order.size(); ++j)
                            covered += e_i > ends[order[i + j]];
                        if (covered >= min_cov) next.push_back(order[i]);
                        else seg.order.push_back(order[i]);
                    }
                    // A pass that keeps nothing cannot happen: the last interval of a list always has covered == 0.
                    segs.push_back(std::move(seg));
                    order.swap(next);
                }
            }
        }
        chroms[c].seg_end = (uint32_t)segs.size();
        max_components = std::max<uint64_t>(max_components, chroms[c].seg_end - chroms[c].seg_begin);
    }

    // ---- 2. flatten to SoA ---------------------------------------------------------------------------------
    std::vector<uint32_t> h_starts(total), h_ends(total), h_pmax(total), h_vals(total), h_cs(total), h_ce(total);
    std::vector<SegMeta> seg_meta(segs.size());
    {
        uint64_t pos = 0;
        for (size_t s = 0; s < segs.size(); ++s) {
            SegMeta& m = seg_meta[s];
            m.off = (uint32_t)pos;
            m.len = (uint32_t)segs[s].order.size();
            m.mono = 1;
            m.pad = 0;
            uint32_t mx = 0;
            for (uint32_t id : segs[s].order) {
                h_starts[pos] = starts[id];
                h_ends[pos] = ends[id];
                h_vals[pos] = vals ? vals[id] : id;
                if (ends[id] < mx) m.mono = 0;
                mx = std::max(mx, ends[id]);
                h_pmax[pos] = mx;
                if (starts[id] > ends[id]) proper = false;
                ++pos;
            }
        }
        for (uint32_t c = 0; c < n_chroms; ++c) {
            uint64_t lo = chrom_offsets[c], hi = chrom_offsets[c + 1];
            chroms[c].off = (uint32_t)lo;
            chroms[c].len = (uint32_t)(hi - lo);
            std::copy(starts + lo, starts + hi, h_cs.begin() + lo);
            std::copy(ends + lo, ends + hi, h_ce.begin() + lo);
            std::sort(h_cs.begin() + lo, h_cs.begin() + hi);
            std::sort(h_ce.begin() + lo, h_ce.begin() + hi);
        }
    }

    // ---- 3. LUT shift: smallest shift whose LUT families stay within the bin budget -------------------------
    std::vector<uint32_t> max_s, max_p, max_cs, max_ce;
    for (const auto& m : seg_meta) {
        max_s.push_back(m.len ? h_starts[m.off + m.len - 1] : 0);
        max_p.push_back(m.len ? h_pmax[m.off + m.len - 1] : 0);
    }
    for (const auto& c : chroms) {
        max_cs.push_back(c.len ? h_cs[c.off + c.len - 1] : 0);
        max_ce.push_back(c.len ? h_ce[c.off + c.len - 1] : 0);
    }
    uint64_t budget = std::max<uint64_t>(2 * total, 4096);
    uint64_t cap = 8ull << 20;
    if (const char* env = getenv("GTGPU_LUT_MAX_BINS")) cap = std::max<uint64_t>(strtoull(env, nullptr, 10), 1024);
    budget = std::min(budget, cap);
    uint32_t shift = 0;
    while (shift < 31 && std::max(lut_entries(max_s, shift), lut_entries(max_p, shift)) > budget) ++shift;
    if (const char* env = getenv("GTGPU_LUT_SHIFT")) shift = (uint32_t)std::min(31, std::max(0, atoi(env)));

    std::vector<uint32_t> lut;
    for (auto& m : seg_meta) {
        build_lut(h_starts.data() + m.off, m.len, shift, lut, m.lut_s, m.nb_s);
        build_lut(h_pmax.data() + m.off, m.len, shift, lut, m.lut_p, m.nb_p);
    }
    // rank LUTs (see IndexView::rank_lut) over the chromosome-level sorted starts / ends
    uint32_t rank_shift = 0;
    {
        uint64_t rbudget = std::max<uint64_t>(2 * total, 4096);
        if (const char* env = getenv("GTGPU_RANK_BINS_PER_INTERVAL")) rbudget = std::max<uint64_t>(strtoull(env, nullptr, 10) * total, 4096);
        while (rank_shift < 29 && std::max(lut_entries(max_cs, rank_shift), lut_entries(max_ce, rank_shift)) > rbudget) ++rank_shift;
    }
    const uint32_t rank_inline = rank_shift == 0 ? 4 : std::min<uint32_t>(4, 29 / rank_shift);
    std::vector<unsigned long long> rank_lut;
    auto build_rank = [&](const uint32_t* arr, uint32_t n, uint32_t& off, uint32_t& nb) {
        off = (uint32_t)rank_lut.size();
        nb = n ? (arr[n - 1] >> rank_shift) + 1 : 0;
        rank_lut.resize(rank_lut.size() + (size_t)nb + 1);
        unsigned long long* L = rank_lut.data() + off;
        uint32_t i = 0;
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t base = i;
            unsigned long long word = base;
            uint32_t cnt = 0;
            while (i < n && (arr[i] >> rank_shift) == b) {
                if (cnt < rank_inline)
                    word |= (unsigned long long)(arr[i] & ((1u << rank_shift) - 1)) << (35 + cnt * rank_shift);
                ++cnt;
                ++i;
            }
            word |= (unsigned long long)(cnt <= rank_inline ? cnt : 7u) << 32;
            L[b] = word;
        }
        L[nb] = n;  // sentinel: base = n, count 0
    };
    for (auto& c : chroms) {
        build_rank(h_cs.data() + c.off, c.len, c.lut_cs, c.nb_cs);
        build_rank(h_ce.data() + c.off, c.len, c.lut_ce, c.nb_ce);
    }
    if (rank_lut.size() >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: rank LUT too large");
    if (lut.size() >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: LUT too large");

    // ---- 3b. bin table (fast path of find/tokenize; see common.cuh) ------------------------------------------
    std::vector<ChromBT> chrom_bt(n_chroms);
    std::vector<uint32_t> bt_lut;   // per window: direct run, pool list or overflow (encodings in common.cuh)
    std::vector<uint32_t> bt_pool;  // candidate lists (entry indices, in emission order) of the pool windows
    std::vector<uint4> bt_ent;      // {start, end, val, 0} per interval, segment order (padded by two)
    std::vector<uint32_t> bt_rec;   // per window: the candidates inline, window-relative (fast path; common.cuh)
    uint32_t bt_shift = 0;
    uint64_t bt_overflow = 0, bt_pool_windows = 0;
    {
        std::vector<uint64_t> cover_end(n_chroms, 0);  // exclusive end of the positions the chromosome's intervals touch
        for (uint32_t c = 0; c < n_chroms; ++c) {
            bool ok = chroms[c].seg_end > chroms[c].seg_begin;
            uint64_t ce = 0;
            for (uint32_t si = chroms[c].seg_begin; ok && si < chroms[c].seg_end; ++si) {
                const SegMeta& m = seg_meta[si];
                for (uint32_t i = m.off; i < m.off + m.len; ++i) {
                    if (h_starts[i] > h_ends[i]) { ok = false; break; }  // start > end intervals: generic path only
                    ce = std::max<uint64_t>(ce, std::max<uint64_t>(h_ends[i], (uint64_t)h_starts[i] + 1));
                }
            }
            cover_end[c] = ok ? ce : 0;
        }
        auto total_bins = [&](uint32_t sh) {
            uint64_t t = 0;
            for (uint32_t c = 0; c < n_chroms; ++c)
                if (cover_end[c]) t += ((cover_end[c] - 1) >> sh) + 1;
            return t;
        };
        uint64_t per_iv = 4;  // bins per interval: narrow enough that a two-bin window rarely holds > 2 candidates
        if (const char* env = getenv("GTGPU_BT_BINS_PER_INTERVAL")) per_iv = std::max<uint64_t>(strtoull(env, nullptr, 10), 1);
        uint64_t bt_budget = std::max<uint64_t>(per_iv * total, 4096);
        uint64_t bt_cap = 6ull << 20;
        if (const char* env = getenv("GTGPU_BT_MAX_BINS")) bt_cap = strtoull(env, nullptr, 10);
        bt_budget = std::min(bt_budget, bt_cap);
        while (bt_shift < 31 && total_bins(bt_shift) > bt_budget) ++bt_shift;
        // window records hold window-relative coordinates in 2 (bt_shift + 2) bits: sparse universes get narrower bins
        if (bt_shift > BT_REC_MAX_SHIFT && total_bins(BT_REC_MAX_SHIFT) <= bt_cap) bt_shift = BT_REC_MAX_SHIFT;
        if (const char* env = getenv("GTGPU_BT_SHIFT")) bt_shift = (uint32_t)std::min(31, std::max(0, atoi(env)));
        // A table only pays off while most windows hold a few candidates: skip it for dense databases.
        bool enabled = bt_cap > 0 && total_bins(bt_shift) <= (64ull << 20) && total <= total_bins(bt_shift) &&
                       total < (1ull << 29) && bt_shift <= BT_REC_MAX_SHIFT;
        uint64_t pos = 1;  // word 0 is an always-empty sentinel: unknown chromosomes and out-of-range bins read it
        for (uint32_t c = 0; c < n_chroms; ++c) {
            chrom_bt[c].off = (uint32_t)pos;
            if (chroms[c].seg_end == chroms[c].seg_begin) {
                chrom_bt[c].n_bins = 0;  // absent chromosome: nothing can hit
            } else if (!cover_end[c] || !enabled) {
                chrom_bt[c].n_bins = BT_GENERIC_CHROM;
                chrom_bt[c].off = 0;
            } else {
                uint64_t nb = ((cover_end[c] - 1) >> bt_shift) + 1;
                if (nb >= BT_NBINS_MASK) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: too many bins on one chromosome");
                chrom_bt[c].n_bins = (uint32_t)nb;
                pos += nb;
            }
        }
        // Pass 1: per window, how many intervals touch it and the index range they span.
        std::vector<uint32_t> w_min(pos, 0xFFFFFFFFu), w_max(pos, 0);
        std::vector<uint8_t> w_cnt(pos, 0);
        auto for_each_window = [&](uint32_t c, uint32_t i, auto&& fn) {
            const uint64_t last_pos = std::max<uint64_t>(h_ends[i], (uint64_t)h_starts[i] + 1) - 1;
            // window b serves queries that START in bin b and end before bin b+2: it lists what touches [b, b+2)
            uint64_t b0 = h_starts[i] >> bt_shift, b1 = last_pos >> bt_shift;
            if (b0 > 0) --b0;
            for (uint64_t b = b0; b <= b1; ++b) fn(chrom_bt[c].off + b);
        };
        for (uint32_t c = 0; c < n_chroms; ++c) {
            if (chrom_bt[c].n_bins == 0 || chrom_bt[c].n_bins == BT_GENERIC_CHROM) continue;
            for (uint32_t si = chroms[c].seg_begin; si < chroms[c].seg_end; ++si)
                for (uint32_t i = seg_meta[si].off; i < seg_meta[si].off + seg_meta[si].len; ++i)
                    for_each_window(c, i, [&](uint64_t k) {
                        w_min[k] = std::min(w_min[k], i);
                        w_max[k] = std::max(w_max[k], i);
                        if (w_cnt[k] < 255) w_cnt[k]++;
                    });
        }
        // Classify: a contiguous run of <= 2 intervals on a single-segment chromosome is addressed directly; up to
        // BT_POOL_MAX candidates of any shape (nested intervals, several AIList components) go to a pool list.
        bt_lut.assign(pos, 0);
        std::vector<uint8_t> fill(pos, 0);
        uint64_t pool_size = 0;
        for (uint32_t c = 0; c < n_chroms; ++c) {
            if (chrom_bt[c].n_bins == 0 || chrom_bt[c].n_bins == BT_GENERIC_CHROM) continue;
            const bool single = chroms[c].seg_end - chroms[c].seg_begin == 1;
            for (uint64_t k = chrom_bt[c].off; k < (uint64_t)chrom_bt[c].off + chrom_bt[c].n_bins; ++k) {
                const uint32_t n = w_cnt[k];
                if (n == 0) continue;
                if (single && n <= 2 && w_max[k] - w_min[k] + 1 == n) {
                    bt_lut[k] = (w_min[k] << 2) | n;
                } else if (n <= BT_POOL_MAX && pool_size + n < (1ull << (31 - BT_POOL_SHIFT)) - 16) {
                    bt_lut[k] = BT_POOL_FLAG | ((uint32_t)pool_size << BT_POOL_SHIFT) | n;
                    pool_size += n;
                    ++bt_pool_windows;
                } else {
                    bt_lut[k] = BT_OVERFLOW;
                    ++bt_overflow;
                }
            }
        }
        // Pass 2: fill the pool lists in emission order — Bits: ascending sorted position; AIList: component-major,
        // descending position inside a component (ailist.rs:153-178, 238-263).
        bt_pool.assign(pool_size + 1, 0);
        for (uint32_t c = 0; c < n_chroms; ++c) {
            if (chrom_bt[c].n_bins == 0 || chrom_bt[c].n_bins == BT_GENERIC_CHROM) continue;
            for (uint32_t si = chroms[c].seg_begin; si < chroms[c].seg_end; ++si) {
                const SegMeta& m = seg_meta[si];
                for (uint32_t t = 0; t < m.len; ++t) {
                    const uint32_t i = kind == GTGPU_KIND_AILIST ? m.off + m.len - 1 - t : m.off + t;
                    for_each_window(c, i, [&](uint64_t k) {
                        const uint32_t w = bt_lut[k];
                        if (w == BT_OVERFLOW || !(w & BT_POOL_FLAG)) return;
                        bt_pool[((w & ~BT_POOL_FLAG) >> BT_POOL_SHIFT) + fill[k]++] = i;
                    });
                }
            }
        }
        for (uint32_t c = 0; c < n_chroms; ++c)
            if (chrom_bt[c].n_bins != 0 && chrom_bt[c].n_bins != BT_GENERIC_CHROM && chroms[c].seg_end - chroms[c].seg_begin > 1)
                chrom_bt[c].n_bins |= BT_MULTI_COMP;
        bt_ent.resize(total + 2);
        for (uint64_t i = 0; i < total; ++i) bt_ent[i] = make_uint4(h_starts[i], h_ends[i], h_vals[i], 0);
        bt_ent[total] = bt_ent[total + 1] = make_uint4(0xFFFFFFFFu, 0, 0, 0);
        // Window records: the direct runs again, inline and window-relative (record 0 stays the empty sentinel).
        bt_rec.assign(pos * BT_REC_WORDS, 0);
        const uint32_t rel_bits = bt_shift + 2;
        const uint64_t rel_max = 2ull << bt_shift;
        auto rel = [&](uint32_t i, uint64_t base) {
            const uint64_t s = h_starts[i] <= base ? 0 : std::min<uint64_t>(h_starts[i] - base, rel_max);
            const uint64_t e = h_ends[i] <= base ? 0 : std::min<uint64_t>(h_ends[i] - base, rel_max);
            return (uint32_t)(s | (e << rel_bits));
        };
        for (uint32_t c = 0; c < n_chroms; ++c) {
            if (chrom_bt[c].n_bins == 0 || (chrom_bt[c].n_bins & BT_GENERIC_CHROM)) continue;
            const uint64_t nb = chrom_bt[c].n_bins & BT_NBINS_MASK;
            for (uint64_t b = 0; b < nb; ++b) {
                const uint64_t k = chrom_bt[c].off + b, base = b << bt_shift;
                const uint32_t w = bt_lut[k];
                uint32_t* r = &bt_rec[k * BT_REC_WORDS];
                if (w == 0) continue;
                if (w & BT_POOL_FLAG) {  // pool list or overflow (BT_OVERFLOW has the flag bit too)
                    r[0] = 3;
                    r[1] = w;  // the full kernel walks the pool list straight from the record
                    continue;
                }
                const uint32_t first = w >> 2, cnt = w & 3;
                r[0] = cnt | (rel(first, base) << 2);
                r[1] = h_vals[first];
                if (cnt == 2) {
                    r[2] = rel(first + 1, base) << 2;
                    r[3] = h_vals[first + 1];
                }
            }
        }
    }

    // ---- 4. upload ---------------------------------------------------------------------------------------------
    gtgpu_index* ix = new gtgpu_index();
    ix->ctx = ctx;
    ix->kind = kind;
    ix->n_intervals = total;
    ix->n_segments = segs.size();
    ix->max_components = max_components;
    IndexView& v = ix->view;
    int32_t st = GTGPU_OK;
    auto up = [&](auto& vec, auto** dst) { if (st == GTGPU_OK) st = upload(ix, vec, dst); };
    up(chrom_bt, &v.chrom_bt);
    up(bt_lut, &v.bt_lut);
    up(bt_rec, &v.bt_rec);
    up(bt_pool, &v.bt_pool);
    ix->bt_pool_windows = bt_pool_windows;
    up(bt_ent, &v.bt_ent);
    v.bt_shift = bt_shift;
    ix->bt_bins = bt_lut.size();
    ix->bt_overflow_bins = bt_overflow;
    ix->bt_clean = bt_overflow == 0 && bt_pool_windows == 0 && total > 0;
    for (uint32_t c = 0; c < n_chroms; ++c)
        if (chrom_bt[c].n_bins & BT_GENERIC_CHROM) ix->bt_clean = false;
    if (ix->bt_clean && cudaHostAlloc((void**)&ix->h_lean_probe, 4, cudaHostAllocDefault) == cudaSuccess) *ix->h_lean_probe = 0;
    else { ix->h_lean_probe = nullptr; cudaGetLastError(); }
    up(chroms, &v.chroms);
    up(seg_meta, &v.segs);
    up(h_starts, &v.starts);
    up(h_ends, &v.ends);
    up(h_pmax, &v.pmax);
    up(h_vals, &v.vals);
    for (uint32_t x : h_vals) ix->max_val = std::max(ix->max_val, x);
    up(h_cs, &v.cs_starts);
    up(h_ce, &v.cs_ends);
    up(lut, &v.lut);
    up(rank_lut, &v.rank_lut);
    v.rank_shift = rank_shift;
    ix->rank_lut_len = rank_lut.size();
    v.rank_inline = rank_inline;
    if (st != GTGPU_OK) {
        gtgpu_index_free(ix);
        return st;
    }
    v.n_chroms = n_chroms;
    v.shift = shift;
    v.descending = kind == GTGPU_KIND_AILIST;
    v.proper = proper;
    *out_index = ix;
    return GTGPU_OK;
}

extern "C" int32_t gtgpu_index_free(gtgpu_index* ix) {
    if (!ix) return GTGPU_OK;
    cudaSetDevice(ix->ctx->device);
    for (void* p : ix->allocs) cudaFree(p);
    if (ix->h_lean_probe) {
        cudaStreamSynchronize(ix->ctx->stream);  // a probe copy may still be in flight
        cudaFreeHost(ix->h_lean_probe);
    }
    delete ix;
    return GTGPU_OK;
}

extern "C" int32_t gtgpu_index_info(const gtgpu_index* ix, uint64_t info[12]) {
    if (!ix || !info) return fail(GTGPU_ERR_INVALID, "index_info: null argument");
    info[0] = ix->n_intervals;
    info[1] = ix->n_segments;
    info[2] = ix->device_bytes;
    info[3] = ix->view.shift;
    info[4] = ix->max_components;
    info[5] = ix->view.proper;
    info[6] = ix->bt_bins;
    info[7] = ix->bt_overflow_bins;
    info[8] = ix->view.bt_shift;
    info[9] = ix->bt_pool_windows;
    info[10] = ix->bt_clean;
    info[11] = ix->lean_off || (ix->h_lean_probe && *ix->h_lean_probe);
    return GTGPU_OK;
}
