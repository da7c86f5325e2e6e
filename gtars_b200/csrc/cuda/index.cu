// index.cu — build of the device index (gtgpu_index_build).
//
// Order semantics of the reference are fixed here, once:
//   Bits   (gtars-overlaprs/src/bits.rs:101-128):   stable sort by (start,end); one segment per chromosome.
//   AIList (gtars-overlaprs/src/ailist.rs:105-151, 198-236): stable sort by start, then repeated peeling of
//          "long" intervals (>= 10 of the next 19 end earlier) into further components; no cap on components.
// Both get the same device layout: start-sorted SoA segments + running max of ends + bin LUTs, so the kernels
// locate candidates identically and only the emission direction differs.
//
// The O(n log n) part — the stable sorts and the independently sorted start / end arrays of the counting identity —
// runs on the device for anything but small inputs (PermSorter, build.cu: stable LSD radix passes over a permutation;
// a 50 M-interval database is ordered in tens of milliseconds instead of ~20 s of std::stable_sort).  What follows is
// linear: AIList peeling, flattening, LUTs and the window table are filled per chromosome by a pool of host threads
// from the sorted order, once, into a HostIndex that can be uploaded to any number of devices (multi-device contexts
// replicate the index).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <thread>

#include "common.cuh"

namespace gtgpu {

int32_t launch_fill_set_ids(gtgpu_ctx* ctx, uint64_t n_sets, const uint64_t* d_set_offsets, uint32_t* d_set_of);

namespace {

// fn(i) for i in [0, n) on up to hardware_concurrency threads (dynamic hand-out: chromosomes differ 10 000x in size)
template <class F>
void parallel_for(uint64_t n, F&& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* env = getenv("GTGPU_BUILD_THREADS")) hw = (unsigned)std::max(1, atoi(env));
    const unsigned nt = (unsigned)std::min<uint64_t>(std::max(1u, std::min(hw, 32u)), n);
    if (nt <= 1) {
        for (uint64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&]() {
            for (uint64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto& t : th) t.join();
}

// lut[b] = lower_bound(arr, b << shift) for b in [0, nb], nb = (max >> shift) + 1, lut[nb] = n (n == 0: one entry).
uint32_t lut_bins(const uint32_t* arr, uint32_t n, uint32_t shift) { return n ? (arr[n - 1] >> shift) + 1 : 0; }
void fill_lut(const uint32_t* arr, uint32_t n, uint32_t shift, uint32_t nb, uint32_t* L) {
    uint32_t i = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        const uint64_t key = (uint64_t)b << shift;
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

struct DevTemp {  // device temporaries of the sort stage
    std::vector<void*> ptrs;
    template <class T>
    int32_t get(uint64_t count, T** out) {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, std::max<size_t>(count * sizeof(T), 16));
        if (e != cudaSuccess) return fail(GTGPU_ERR_NOMEM, std::string("cudaMalloc(index build): ") + cudaGetErrorString(e));
        ptrs.push_back(d);
        *out = (T*)d;
        return GTGPU_OK;
    }
    ~DevTemp() {
        for (void* p : ptrs) cudaFree(p);
    }
};

// Device stage: order[] = every chromosome's intervals in the backend's sorted order (global indices, chromosome c in
// [offs[c], offs[c+1])), cs[] / ce[] = each chromosome's starts / ends sorted independently.
int32_t device_order(gtgpu_ctx* ctx, int32_t kind, uint32_t n_chroms, const uint64_t* chrom_offsets, uint64_t total,
                     const uint32_t* starts, const uint32_t* ends, uint32_t* order, uint32_t* cs, uint32_t* ce) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    DevTemp tp;
    uint32_t *d_s, *d_e, *d_chr, *d_tmp;
    uint64_t* d_co;
    GT_TRY(tp.get(total, &d_s));
    GT_TRY(tp.get(total, &d_e));
    GT_TRY(tp.get(total, &d_chr));
    GT_TRY(tp.get(total, &d_tmp));
    GT_TRY(tp.get(n_chroms + 1, &d_co));
    GT_CUDA(cudaMemcpyAsync(d_s, starts, total * 4, cudaMemcpyHostToDevice, st));
    GT_CUDA(cudaMemcpyAsync(d_e, ends, total * 4, cudaMemcpyHostToDevice, st));
    GT_CUDA(cudaMemcpyAsync(d_co, chrom_offsets, (n_chroms + 1) * 8, cudaMemcpyHostToDevice, st));
    GT_CUDA(cudaStreamSynchronize(st));
    GT_TRY(launch_fill_set_ids(ctx, n_chroms, d_co, d_chr));
    const int chr_bits = bits_for_value(n_chroms);
    {
        PermSorter by_start;
        GT_TRY(by_start.init(ctx, total));
        if (kind == GTGPU_KIND_BITS) GT_TRY(by_start.pass(d_e, 0));  // Bits: ties on start are ordered by end (interval.rs:18-31)
        GT_TRY(by_start.pass(d_s, 0));
        GT_TRY(by_start.pass(d_chr, chr_bits));
        GT_CUDA(cudaMemcpyAsync(order, by_start.perm, total * 4, cudaMemcpyDeviceToHost, st));
        GT_TRY(launch_gather_u32(ctx, total, d_s, by_start.perm, d_tmp));  // starts in that order = sorted per chromosome
        GT_CUDA(cudaMemcpyAsync(cs, d_tmp, total * 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
    }
    {
        PermSorter by_end;
        GT_TRY(by_end.init(ctx, total));
        GT_TRY(by_end.pass(d_e, 0));
        GT_TRY(by_end.pass(d_chr, chr_bits));
        GT_TRY(launch_gather_u32(ctx, total, d_e, by_end.perm, d_tmp));
        GT_CUDA(cudaMemcpyAsync(ce, d_tmp, total * 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
    }
    return GTGPU_OK;
}

struct HostSeg {
    std::vector<uint32_t> order;  // indices into the caller's arrays
};

}  // namespace

// Everything gtgpu_index needs, in host memory: built once, uploaded to one device or to every device of a group.
struct HostIndex {
    int32_t kind = 0;
    uint32_t n_chroms = 0;
    uint64_t total = 0, n_segments = 0, max_components = 0;
    bool proper = true;
    std::vector<ChromMeta> chroms;
    std::vector<SegMeta> seg_meta;
    std::vector<uint32_t> h_starts, h_ends, h_pmax, h_vals, h_cs, h_ce, lut;
    std::vector<unsigned long long> rank_lut;
    std::vector<uint4> rank_lin;  // per chromosome, empty when the linearised coordinates do not fit 32 bits (IndexView::rank_lin)
    uint32_t rank_ends_off = 0;
    uint32_t shift = 0, rank_shift = 0, rank_inline = 0, bt_shift = 0, max_val = 0;
    std::vector<ChromBT> chrom_bt;
    std::vector<uint32_t> bt_lut, bt_pool, bt_rec;
    std::vector<uint4> bt_ent;
    uint64_t bt_overflow = 0, bt_pool_windows = 0;
};

void host_index_free(HostIndex* h) { delete h; }

int32_t host_index_build(gtgpu_ctx* ctx, int32_t kind, uint32_t n_chroms, const uint64_t* chrom_offsets, const uint32_t* starts,
                         const uint32_t* ends, const uint32_t* vals, HostIndex** out) {
    if (!ctx || !out || (n_chroms && !chrom_offsets)) return fail(GTGPU_ERR_INVALID, "index_build: null argument");
    if (kind != GTGPU_KIND_BITS && kind != GTGPU_KIND_AILIST) return fail(GTGPU_ERR_INVALID, "index_build: bad kind");
    const uint64_t total = n_chroms ? chrom_offsets[n_chroms] : 0;
    if (total >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: more than 2^32-2 intervals");
    if (total && (!starts || !ends)) return fail(GTGPU_ERR_INVALID, "index_build: null coordinate arrays");
    for (uint32_t c = 0; c < n_chroms; ++c)
        if (chrom_offsets[c] > chrom_offsets[c + 1]) return fail(GTGPU_ERR_INVALID, "index_build: chrom_offsets not monotone");
    if (n_chroms && chrom_offsets[0] != 0) return fail(GTGPU_ERR_INVALID, "index_build: chrom_offsets[0] must be 0");
    GT_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<HostIndex> hp(new HostIndex());
    HostIndex& H = *hp;
    H.kind = kind;
    H.n_chroms = n_chroms;
    H.total = total;

    // ---- 1. per-chromosome ordering (device radix sort, or the host's stable sort for small inputs) ---------------------
    std::vector<uint32_t> order_all(total);
    H.h_cs.resize(total);
    H.h_ce.resize(total);
    bool on_device = total >= (1u << 16);
    if (const char* env = getenv("GTGPU_BUILD_SORT")) on_device = total > 0 && env[0] == 'd';  // "device" / "host"
    if (on_device) {
        GT_TRY(device_order(ctx, kind, n_chroms, chrom_offsets, total, starts, ends, order_all.data(), H.h_cs.data(), H.h_ce.data()));
    } else {
        parallel_for(n_chroms, [&](uint64_t c) {
            const uint64_t lo = chrom_offsets[c], hi = chrom_offsets[c + 1];
            std::iota(order_all.begin() + lo, order_all.begin() + hi, (uint32_t)lo);
            if (kind == GTGPU_KIND_BITS)
                std::stable_sort(order_all.begin() + lo, order_all.begin() + hi, [&](uint32_t a, uint32_t b) {
                    return starts[a] != starts[b] ? starts[a] < starts[b] : ends[a] < ends[b];
                });
            else
                std::stable_sort(order_all.begin() + lo, order_all.begin() + hi,
                                 [&](uint32_t a, uint32_t b) { return starts[a] < starts[b]; });
            std::copy(starts + lo, starts + hi, H.h_cs.begin() + lo);
            std::copy(ends + lo, ends + hi, H.h_ce.begin() + lo);
            std::sort(H.h_cs.begin() + lo, H.h_cs.begin() + hi);
            std::sort(H.h_ce.begin() + lo, H.h_ce.begin() + hi);
        });
    }

    // ---- 1b. segments: Bits = the chromosome; AIList = repeated peeling of the sorted list (ailist.rs:128-142, 198-236) --
    std::vector<std::vector<HostSeg>> chrom_segs(n_chroms);
    parallel_for(n_chroms, [&](uint64_t c) {
        const uint64_t lo = chrom_offsets[c], hi = chrom_offsets[c + 1];
        if (hi == lo) return;
        std::vector<uint32_t> order(order_all.begin() + lo, order_all.begin() + hi);
        if (kind == GTGPU_KIND_BITS) {
            chrom_segs[c].push_back(HostSeg{std::move(order)});
            return;
        }
        const size_t min_cov = 10;
        std::vector<uint32_t> next;
        while (!order.empty()) {
            HostSeg seg;
            next.clear();
            for (size_t i = 0; i < order.size(); ++i) {
                size_t covered = 0;
                const uint32_t e_i = ends[order[i]];
                for (size_t j = 1; j < 2 * min_cov && i + j < order.size(); ++j) covered += e_i > ends[order[i + j]];
                if (covered >= min_cov) next.push_back(order[i]);
                else seg.order.push_back(order[i]);
            }
            // A pass that keeps nothing cannot happen: the last interval of a list always has covered == 0.
            chrom_segs[c].push_back(std::move(seg));
            order.swap(next);
        }
    });
    std::vector<uint32_t>().swap(order_all);
    std::vector<ChromMeta>& chroms = H.chroms;
    chroms.assign(n_chroms, ChromMeta{});
    std::vector<HostSeg*> segs;
    for (uint32_t c = 0; c < n_chroms; ++c) {
        chroms[c].seg_begin = (uint32_t)segs.size();
        for (auto& s : chrom_segs[c]) segs.push_back(&s);
        chroms[c].seg_end = (uint32_t)segs.size();
        chroms[c].off = (uint32_t)chrom_offsets[c];
        chroms[c].len = (uint32_t)(chrom_offsets[c + 1] - chrom_offsets[c]);
        H.max_components = std::max<uint64_t>(H.max_components, chroms[c].seg_end - chroms[c].seg_begin);
    }
    H.n_segments = segs.size();

    // ---- 2. flatten to SoA ---------------------------------------------------------------------------------
    std::vector<uint32_t>&h_starts = H.h_starts, &h_ends = H.h_ends, &h_pmax = H.h_pmax, &h_vals = H.h_vals, &h_cs = H.h_cs, &h_ce = H.h_ce;
    h_starts.resize(total);
    h_ends.resize(total);
    h_pmax.resize(total);
    h_vals.resize(total);
    std::vector<SegMeta>& seg_meta = H.seg_meta;
    seg_meta.assign(segs.size(), SegMeta{});
    {
        uint64_t pos = 0;
        for (size_t s = 0; s < segs.size(); ++s) {
            seg_meta[s].off = (uint32_t)pos;
            seg_meta[s].len = (uint32_t)segs[s]->order.size();
            pos += segs[s]->order.size();
        }
        std::vector<uint8_t> improper(segs.size(), 0);
        std::vector<uint32_t> seg_max_val(segs.size(), 0);
        parallel_for(segs.size(), [&](uint64_t s) {
            SegMeta& m = seg_meta[s];
            m.mono = 1;
            m.pad = 0;
            uint32_t mx = 0, mv = 0;
            uint64_t p = m.off;
            for (uint32_t id : segs[s]->order) {
                h_starts[p] = starts[id];
                h_ends[p] = ends[id];
                h_vals[p] = vals ? vals[id] : id;
                mv = std::max(mv, h_vals[p]);
                if (ends[id] < mx) m.mono = 0;
                mx = std::max(mx, ends[id]);
                h_pmax[p] = mx;
                if (starts[id] > ends[id]) improper[s] = 1;
                ++p;
            }
            seg_max_val[s] = mv;
            std::vector<uint32_t>().swap(segs[s]->order);
        });
        for (size_t s = 0; s < segs.size(); ++s) {
            if (improper[s]) H.proper = false;
            H.max_val = std::max(H.max_val, seg_max_val[s]);
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
    H.shift = shift;

    std::vector<uint32_t>& lut = H.lut;
    {
        uint64_t len = 0;
        for (auto& m : seg_meta) {
            m.nb_s = lut_bins(h_starts.data() + m.off, m.len, shift);
            m.nb_p = lut_bins(h_pmax.data() + m.off, m.len, shift);
            m.lut_s = (uint32_t)len;
            len += (uint64_t)m.nb_s + 1;
            m.lut_p = (uint32_t)len;
            len += (uint64_t)m.nb_p + 1;
        }
        if (len >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: LUT too large");
        lut.resize(len);
        parallel_for(seg_meta.size(), [&](uint64_t s) {
            const SegMeta& m = seg_meta[s];
            fill_lut(h_starts.data() + m.off, m.len, shift, m.nb_s, lut.data() + m.lut_s);
            fill_lut(h_pmax.data() + m.off, m.len, shift, m.nb_p, lut.data() + m.lut_p);
        });
    }
    // rank LUTs (see IndexView::rank_lut) over the chromosome-level sorted starts / ends
    uint32_t rank_shift = 0;
    {
        // about one bin per interval: four inline offsets per word cover all but 0.4 % of the bins of a uniformly spread
        // database, and the LUT is half the size of a two-bins-per-interval one (C3: 0.78 instead of 1.55 GB; the bucketed
        // pass 1.72 instead of 1.94 ms, the direct pass 4.08 instead of 4.32 ms per 1e8 queries)
        uint64_t rbudget = std::max<uint64_t>(total, 4096);
        if (const char* env = getenv("GTGPU_RANK_BINS_PER_INTERVAL")) rbudget = std::max<uint64_t>(strtoull(env, nullptr, 10) * total, 4096);
        while (rank_shift < 29 && std::max(lut_entries(max_cs, rank_shift), lut_entries(max_ce, rank_shift)) > rbudget) ++rank_shift;
    }
    if (const char* env = getenv("GTGPU_RANK_SHIFT_EXTRA")) rank_shift = (uint32_t)std::min(29, (int)rank_shift + std::max(0, atoi(env)));
    const uint32_t rank_inline = rank_shift == 0 ? 4 : std::min<uint32_t>(4, 29 / rank_shift);
    H.rank_shift = rank_shift;
    H.rank_inline = rank_inline;
    std::vector<unsigned long long>& rank_lut = H.rank_lut;
    auto fill_rank = [&](const uint32_t* arr, uint32_t n, uint32_t nb, unsigned long long* L) {
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
    {
        // Layout: the starts LUTs of all chromosomes back to back, then their ends LUTs.  A chromosome's position in its
        // block, shifted up by rank_shift, is what linearises its coordinates for the bucketed counting pass
        // (IndexView::rank_lin): word = lin >> rank_shift, in-bin offset = the low bits, no chromosome look-up.
        uint64_t len = 0;
        for (auto& c : chroms) {
            c.nb_cs = lut_bins(h_cs.data() + c.off, c.len, rank_shift);
            c.nb_ce = lut_bins(h_ce.data() + c.off, c.len, rank_shift);
            c.lut_cs = (uint32_t)len;
            len += (uint64_t)c.nb_cs + 1;
        }
        const uint64_t words_s = len;
        for (auto& c : chroms) {
            c.lut_ce = (uint32_t)len;
            len += (uint64_t)c.nb_ce + 1;
        }
        if (len >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: rank LUT too large");
        H.rank_ends_off = (uint32_t)words_s;
        if ((words_s << rank_shift) <= 0xFFFFFFFFull && ((len - words_s) << rank_shift) <= 0xFFFFFFFFull) {
            H.rank_lin.reserve(chroms.size());
            for (auto& c : chroms)
                H.rank_lin.push_back(make_uint4(c.lut_cs << rank_shift, c.nb_cs << rank_shift,
                                                (uint32_t)(c.lut_ce - words_s) << rank_shift, c.nb_ce << rank_shift));
        }
        rank_lut.resize(len);
        parallel_for(2 * (uint64_t)chroms.size(), [&](uint64_t k) {
            const ChromMeta& c = chroms[k >> 1];
            if (k & 1) fill_rank(h_ce.data() + c.off, c.len, c.nb_ce, rank_lut.data() + c.lut_ce);
            else fill_rank(h_cs.data() + c.off, c.len, c.nb_cs, rank_lut.data() + c.lut_cs);
        });
    }

    // ---- 3b. bin table (fast path of find/tokenize; see common.cuh) ------------------------------------------
    std::vector<ChromBT>& chrom_bt = H.chrom_bt;
    chrom_bt.assign(n_chroms, ChromBT{});
    std::vector<uint32_t>& bt_lut = H.bt_lut;    // per window: direct run, pool list or overflow (encodings in common.cuh)
    std::vector<uint32_t>& bt_pool = H.bt_pool;  // candidate lists (entry indices, in emission order) of the pool windows
    std::vector<uint4>& bt_ent = H.bt_ent;       // {start, end, val, 0} per interval, segment order (padded by two)
    std::vector<uint32_t>& bt_rec = H.bt_rec;    // per window: the candidates inline, window-relative (fast path; common.cuh)
    uint32_t bt_shift = 0;
    {
        std::vector<uint64_t> cover_end(n_chroms, 0);  // exclusive end of the positions the chromosome's intervals touch
        parallel_for(n_chroms, [&](uint64_t c) {
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
        });
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
        H.bt_shift = bt_shift;
        // A table only pays off while most windows hold a few candidates: skip it for dense databases.
        bool enabled = bt_cap > 0 && total_bins(bt_shift) <= (64ull << 20) && total <= total_bins(bt_shift) &&
                       total < (1ull << 29) && bt_shift <= BT_REC_MAX_SHIFT;
        uint64_t pos = 1;  // word 0 is an always-empty sentinel: unknown chromosomes and out-of-range bins read it
        for (uint32_t c = 0; c < n_chroms; ++c) {
            chrom_bt[c].off = (uint32_t)pos;
            if (chroms[c].seg_end == chroms[c].seg_begin) {
                chrom_bt[c].n_bins = 0;  // absent chromosome: nothing can hit
                chrom_bt[c].off = 0;     // (its one readable record is the global empty sentinel)
            } else if (!cover_end[c] || !enabled || (uint64_t)cover_end[c] > 0xFFFFFFFFull - (8ull << bt_shift)) {
                // (a table never reaches the top of the u32 range: the lean kernel's window test relies on it for e == 0)
                chrom_bt[c].n_bins = BT_GENERIC_CHROM;
                chrom_bt[c].off = 0;
            } else {
                uint64_t nb = ((cover_end[c] - 1) >> bt_shift) + 1;
                if (nb >= BT_NBINS_MASK) return fail(GTGPU_ERR_UNSUPPORTED, "index_build: too many bins on one chromosome");
                chrom_bt[c].n_bins = (uint32_t)nb;
                pos += nb + 1;  // + one always-empty record: where the fused kernel sends bins past the chromosome's last one
            }
        }
        auto has_table = [&](uint32_t c) { return chrom_bt[c].n_bins != 0 && !(chrom_bt[c].n_bins & BT_GENERIC_CHROM); };
        // Pass 1: per window, how many intervals touch it and the index range they span (chromosomes own disjoint windows).
        std::vector<uint32_t> w_min(pos, 0xFFFFFFFFu), w_max(pos, 0);
        std::vector<uint8_t> w_cnt(pos, 0);
        auto for_each_window = [&](uint32_t c, uint32_t i, auto&& fn) {
            const uint64_t last_pos = std::max<uint64_t>(h_ends[i], (uint64_t)h_starts[i] + 1) - 1;
            // window b serves queries that START in bin b and end before bin b+2: it lists what touches [b, b+2)
            uint64_t b0 = h_starts[i] >> bt_shift, b1 = last_pos >> bt_shift;
            if (b0 > 0) --b0;
            for (uint64_t b = b0; b <= b1; ++b) fn(chrom_bt[c].off + b);
        };
        parallel_for(n_chroms, [&](uint64_t cc) {
            const uint32_t c = (uint32_t)cc;
            if (!has_table(c)) return;
            for (uint32_t si = chroms[c].seg_begin; si < chroms[c].seg_end; ++si)
                for (uint32_t i = seg_meta[si].off; i < seg_meta[si].off + seg_meta[si].len; ++i)
                    for_each_window(c, i, [&](uint64_t k) {
                        w_min[k] = std::min(w_min[k], i);
                        w_max[k] = std::max(w_max[k], i);
                        if (w_cnt[k] < 255) w_cnt[k]++;
                    });
        });
        // Classify: a contiguous run of <= 2 intervals on a single-segment chromosome is addressed directly; up to
        // BT_POOL_MAX candidates of any shape (nested intervals, several AIList components) go to a pool list.  Pool
        // offsets follow window order across chromosomes; once the pool is full the remaining windows overflow.
        bt_lut.assign(pos, 0);
        std::vector<uint8_t> fill(pos, 0);
        const uint64_t pool_limit = (1ull << (31 - BT_POOL_SHIFT)) - 16;
        auto is_direct = [&](uint32_t c, uint64_t k) {
            const uint32_t n = w_cnt[k];
            return chroms[c].seg_end - chroms[c].seg_begin == 1 && n <= 2 && w_max[k] - w_min[k] + 1 == n;
        };
        std::vector<uint64_t> pool_need(n_chroms + 1, 0);  // pool entries per chromosome when nothing overflows the pool
        parallel_for(n_chroms, [&](uint64_t cc) {
            const uint32_t c = (uint32_t)cc;
            if (!has_table(c)) return;
            uint64_t need = 0;
            for (uint64_t k = chrom_bt[c].off; k < (uint64_t)chrom_bt[c].off + chrom_bt[c].n_bins; ++k) {
                const uint32_t n = w_cnt[k];
                if (n && !is_direct(c, k) && n <= BT_POOL_MAX) need += n;
            }
            pool_need[c + 1] = need;
        });
        for (uint32_t c = 0; c < n_chroms; ++c) pool_need[c + 1] += pool_need[c];
        const bool pool_fits = pool_need[n_chroms] < pool_limit;
        std::vector<uint64_t> c_overflow(n_chroms, 0), c_poolw(n_chroms, 0);
        auto classify = [&](uint32_t c, uint64_t& pool_size) {
            for (uint64_t k = chrom_bt[c].off; k < (uint64_t)chrom_bt[c].off + chrom_bt[c].n_bins; ++k) {
                const uint32_t n = w_cnt[k];
                if (n == 0) continue;
                if (is_direct(c, k)) {
                    bt_lut[k] = (w_min[k] << 2) | n;
                } else if (n <= BT_POOL_MAX && pool_size + n < pool_limit) {
                    bt_lut[k] = BT_POOL_FLAG | ((uint32_t)pool_size << BT_POOL_SHIFT) | n;
                    pool_size += n;
                    ++c_poolw[c];
                } else {
                    bt_lut[k] = BT_OVERFLOW;
                    ++c_overflow[c];
                }
            }
        };
        uint64_t pool_size = 0;
        if (pool_fits) {
            parallel_for(n_chroms, [&](uint64_t cc) {
                if (!has_table((uint32_t)cc)) return;
                uint64_t ps = pool_need[cc];
                classify((uint32_t)cc, ps);
            });
            pool_size = pool_need[n_chroms];
        } else {
            for (uint32_t c = 0; c < n_chroms; ++c)
                if (has_table(c)) classify(c, pool_size);
        }
        for (uint32_t c = 0; c < n_chroms; ++c) {
            H.bt_overflow += c_overflow[c];
            H.bt_pool_windows += c_poolw[c];
        }
        // Pass 2: fill the pool lists in emission order — Bits: ascending sorted position; AIList: component-major,
        // descending position inside a component (ailist.rs:153-178, 238-263).
        bt_pool.assign(pool_size + 1, 0);
        parallel_for(n_chroms, [&](uint64_t cc) {
            const uint32_t c = (uint32_t)cc;
            if (!has_table(c)) return;
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
        });
        bt_ent.resize(total + 2);
        parallel_for((total + (1 << 20) - 1) >> 20, [&](uint64_t blk) {
            const uint64_t lo = blk << 20, hi = std::min<uint64_t>(total, lo + (1 << 20));
            for (uint64_t i = lo; i < hi; ++i) bt_ent[i] = make_uint4(h_starts[i], h_ends[i], h_vals[i], 0);
        });
        bt_ent[total] = bt_ent[total + 1] = make_uint4(0xFFFFFFFFu, 0, 0, 0);
        // Window records: the direct runs again, inline and window-relative (record 0 stays the empty sentinel).
        bt_rec.assign(pos * BT_REC_WORDS, 0);
        const uint64_t rel_max = 2ull << bt_shift;
        for (uint64_t k = 0; k < pos; ++k) bt_rec[k * BT_REC_WORDS] = bt_rec[k * BT_REC_WORDS + 2] = BT_REC_EMPTY;
        auto rel = [&](uint32_t i, uint64_t base) {
            const uint64_t s = h_starts[i] <= base ? 0 : std::min<uint64_t>(h_starts[i] - base, rel_max);
            const uint64_t e = h_ends[i] <= base ? 0 : std::min<uint64_t>(h_ends[i] - base, rel_max);
            return (uint32_t)(((e + 0x7FFFu) << 16) - s);
        };
        parallel_for(n_chroms, [&](uint64_t cc) {
            const uint32_t c = (uint32_t)cc;
            if (!has_table(c)) return;
            const uint64_t nb = chrom_bt[c].n_bins & BT_NBINS_MASK;
            for (uint64_t b = 0; b < nb; ++b) {
                const uint64_t k = chrom_bt[c].off + b, base = b << bt_shift;
                const uint32_t w = bt_lut[k];
                uint32_t* r = &bt_rec[k * BT_REC_WORDS];
                if (w == 0) continue;
                if (w & BT_POOL_FLAG) {  // pool list or overflow (BT_OVERFLOW has the flag bit too)
                    r[0] = BT_REC_SLOW;
                    r[1] = w;  // the full kernel walks the pool list straight from the record
                    continue;
                }
                const uint32_t first = w >> 2, cnt = w & 3;
                r[0] = rel(first, base);
                r[1] = h_vals[first];
                if (cnt == 2) {
                    r[2] = rel(first + 1, base);
                    r[3] = h_vals[first + 1];
                }
            }
        });
        for (uint32_t c = 0; c < n_chroms; ++c)
            if (has_table(c) && chroms[c].seg_end - chroms[c].seg_begin > 1) chrom_bt[c].n_bins |= BT_MULTI_COMP;
    }
    *out = hp.release();
    return GTGPU_OK;
}

namespace {

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

int32_t index_free_impl(gtgpu_index* ix);

int32_t host_index_upload(gtgpu_ctx* ctx, const HostIndex& H, gtgpu_index** out_index) {
    GT_CUDA(cudaSetDevice(ctx->device));
    gtgpu_index* ix = new gtgpu_index();
    ix->ctx = ctx;
    ix->kind = H.kind;
    ix->n_intervals = H.total;
    ix->n_segments = H.n_segments;
    ix->max_components = H.max_components;
    IndexView& v = ix->view;
    int32_t st = GTGPU_OK;
    auto up = [&](auto& vec, auto** dst) { if (st == GTGPU_OK) st = upload(ix, vec, dst); };
    up(H.chrom_bt, &v.chrom_bt);
    up(H.bt_lut, &v.bt_lut);
    up(H.bt_rec, &v.bt_rec);
    up(H.bt_pool, &v.bt_pool);
    ix->bt_pool_windows = H.bt_pool_windows;
    up(H.bt_ent, &v.bt_ent);
    v.bt_shift = H.bt_shift;
    ix->bt_bins = H.bt_lut.size();
    ix->bt_overflow_bins = H.bt_overflow;
    ix->bt_clean = H.bt_overflow == 0 && H.bt_pool_windows == 0 && H.total > 0;
    for (uint32_t c = 0; c < H.n_chroms; ++c)
        if (H.chrom_bt[c].n_bins & BT_GENERIC_CHROM) ix->bt_clean = false;
    if (H.n_chroms >= (uint32_t)CHROM_CACHE) ix->bt_clean = false;  // the lean kernel reads chromosome entries from its shared-memory cache only
    if (ix->bt_clean && cudaHostAlloc((void**)&ix->h_lean_probe, 4, cudaHostAllocDefault) == cudaSuccess) *ix->h_lean_probe = 0;
    else { ix->h_lean_probe = nullptr; cudaGetLastError(); }
    up(H.chroms, &v.chroms);
    up(H.seg_meta, &v.segs);
    up(H.h_starts, &v.starts);
    up(H.h_ends, &v.ends);
    up(H.h_pmax, &v.pmax);
    up(H.h_vals, &v.vals);
    ix->max_val = H.max_val;
    up(H.h_cs, &v.cs_starts);
    up(H.h_ce, &v.cs_ends);
    up(H.lut, &v.lut);
    up(H.rank_lut, &v.rank_lut);
    v.rank_shift = H.rank_shift;
    ix->rank_lut_len = H.rank_lut.size();
    v.rank_inline = H.rank_inline;
    v.rank_lin = nullptr;
    if (!H.rank_lin.empty()) up(H.rank_lin, &v.rank_lin);
    v.rank_ends_off = H.rank_ends_off;
    if (st != GTGPU_OK) {
        index_free_impl(ix);
        return st;
    }
    v.n_chroms = H.n_chroms;
    v.shift = H.shift;
    v.descending = H.kind == GTGPU_KIND_AILIST;
    v.proper = H.proper;
    *out_index = ix;
    return GTGPU_OK;
}

int32_t index_free_impl(gtgpu_index* ix) {
    if (!ix) return GTGPU_OK;
    cudaSetDevice(ix->ctx->device);
    if (ix->ctx->l2_window_owner == ix) release_l2_window(ix->ctx);  // the stream's access-policy window points into this index
    for (void* p : ix->allocs) cudaFree(p);
    if (ix->h_lean_probe) {
        cudaStreamSynchronize(ix->ctx->stream);  // a probe copy may still be in flight
        cudaFreeHost(ix->h_lean_probe);
    }
    delete ix;
    return GTGPU_OK;
}

}  // namespace gtgpu

using namespace gtgpu;

extern "C" int32_t gtgpu_index_build(gtgpu_ctx* ctx, int32_t kind, uint32_t n_chroms, const uint64_t* chrom_offsets,
                                     const uint32_t* starts, const uint32_t* ends, const uint32_t* vals,
                                     gtgpu_index** out_index) try {
    if (!ctx || !out_index) return fail(GTGPU_ERR_INVALID, "index_build: null argument");
    HostIndex* H = nullptr;
    GT_TRY(host_index_build(ctx, kind, n_chroms, chrom_offsets, starts, ends, vals, &H));
    std::unique_ptr<HostIndex> hold(H);
    const size_t D = ctx->peers.size();
    if (D <= 1) return host_index_upload(ctx, *H, out_index);
    // multi-device group: the index is replicated — built once on the host, uploaded to every device in parallel
    std::lock_guard<std::mutex> glk(ctx->group_mu);
    std::vector<gtgpu_index*> reps(D, nullptr);
    const int32_t s = for_each_device(D, [&](size_t r) -> int32_t { return host_index_upload(ctx->peers[r], *H, &reps[r]); });
    if (s != GTGPU_OK) {
        const std::string msg = gtgpu_last_error();
        for (gtgpu_index* r : reps) index_free_impl(r);
        return fail(s, msg);
    }
    reps[0]->replicas = reps;
    *out_index = reps[0];
    return GTGPU_OK;
} GT_CATCH

extern "C" int32_t gtgpu_index_free(gtgpu_index* ix) try {
    if (!ix) return GTGPU_OK;
    const std::vector<gtgpu_index*> reps = ix->replicas;
    for (size_t r = 1; r < reps.size(); ++r) index_free_impl(reps[r]);
    return index_free_impl(ix);
} GT_CATCH

extern "C" int32_t gtgpu_index_info(const gtgpu_index* ix, uint64_t info[12]) try {
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
} GT_CATCH
