// marshal.cu — host-side marshalling helpers of the C ABI (no device work): what a caller does between its own region
// arrays and the wire formats of gtgpu_tokenize_files_compact / gtgpu_tokenize_files_packed.  Multi-threaded, one pass over the queries.
#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace gtgpu;

namespace gtgpu {

int32_t marshal_compact_impl(uint64_t n, const uint32_t* chr, const uint32_t* start, const uint32_t* end, uint64_t n_files,
                             const uint64_t* file_offsets, int32_t threads, uint16_t* out_width16, uint64_t run_capacity,
                             uint64_t* out_run_offsets, uint32_t* out_run_chr, uint64_t* out_n_runs, uint64_t wide_capacity,
                             uint64_t* out_wide_index, uint32_t* out_wide_end, uint64_t* out_n_wide) {
    if (!out_n_runs || !out_n_wide || (n && (!chr || !start || !end || !out_width16)) || (n_files && !file_offsets) ||
        (run_capacity && (!out_run_offsets || !out_run_chr)) || (wide_capacity && (!out_wide_index || !out_wide_end)))
        return fail(GTGPU_ERR_INVALID, "marshal_compact: null argument");
    for (uint64_t f = 0; f < n_files; ++f)
        if (file_offsets[f] > file_offsets[f + 1]) return fail(GTGPU_ERR_INVALID, "marshal_compact: file_offsets not monotone");
    if (n_files && (file_offsets[0] != 0 || file_offsets[n_files] != n))
        return fail(GTGPU_ERR_INVALID, "marshal_compact: file_offsets must span [0, n]");
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(nt, (n + (1 << 20) - 1) >> 20));
    struct Part {
        std::vector<uint64_t> run_off, wide_idx;
        std::vector<uint32_t> run_chr, wide_end;
    };
    std::vector<Part> parts(nt);
    auto work = [&](unsigned t) {
        const uint64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        Part& p = parts[t];
        // next file boundary at or after lo (a run never crosses a file boundary: files are separate encode() calls)
        const uint64_t* fb = n_files ? std::lower_bound(file_offsets, file_offsets + n_files + 1, lo) : nullptr;
        const uint64_t* fb_end = n_files ? file_offsets + n_files + 1 : nullptr;
        uint32_t prev = lo ? chr[lo - 1] : 0;
        for (uint64_t i = lo; i < hi; ++i) {
            const uint32_t c = chr[i], s = start[i], e = end[i];
            bool cut = i == 0 || c != prev;
            while (fb != fb_end && *fb < i) ++fb;
            if (fb != fb_end && *fb == i) cut = true;
            if (cut) {
                p.run_off.push_back(i);
                p.run_chr.push_back(c);
            }
            prev = c;
            const uint32_t w = e - s;
            if (e < s || w > 0xFFFEu) {  // does not fit 16 bits (or end < start): exception list, width field = 0xFFFF
                out_width16[i] = 0xFFFFu;
                p.wide_idx.push_back(i);
                p.wide_end.push_back(e);
            } else {
                out_width16[i] = (uint16_t)w;
            }
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    uint64_t n_runs = 0, n_wide = 0;
    for (auto& p : parts) {
        n_runs += p.run_off.size();
        n_wide += p.wide_idx.size();
    }
    *out_n_runs = n_runs;
    *out_n_wide = n_wide;
    if (n_runs > run_capacity || n_wide > wide_capacity)
        return fail(GTGPU_ERR_CAPACITY, "marshal_compact: run / exception capacity too small (needed counts returned)");
    uint64_t r = 0, w = 0;
    for (auto& p : parts) {
        std::copy(p.run_off.begin(), p.run_off.end(), out_run_offsets + r);
        std::copy(p.run_chr.begin(), p.run_chr.end(), out_run_chr + r);
        r += p.run_off.size();
        std::copy(p.wide_idx.begin(), p.wide_idx.end(), out_wide_index + w);
        std::copy(p.wide_end.begin(), p.wide_end.end(), out_wide_end + w);
        w += p.wide_idx.size();
    }
    if (run_capacity || n_runs) out_run_offsets[n_runs] = n;
    return GTGPU_OK;
}


// ---- packed wire format (gtgpu_tokenize_files_packed): ONE 32-bit word per query -----------------------------------------
// Blocks of 32 consecutive queries share a 32-bit anchor; word = (start - anchor) | width << (32 - width_bits).  Queries of a
// sorted BED file (RegionSet::sort: by chromosome, then start) are ascending inside a chromosome run, so the offsets inside a
// block are small; whatever does not fit — the far side of a run boundary inside a block, wide or reversed regions — goes to
// an exception list with its absolute (start, end).
namespace {

constexpr uint64_t PK_BLOCK = 32;

// exceptions of block [b0, b1) for a given anchor and field split
inline uint32_t pk_misfits(const uint32_t* start, const uint32_t* end, uint64_t b0, uint64_t b1, uint32_t anchor, uint32_t off_bits,
                           uint32_t width_bits) {
    uint32_t bad = 0;
    for (uint64_t i = b0; i < b1; ++i) {
        const uint32_t off = start[i] - anchor, w = end[i] - start[i];
        bad += (start[i] < anchor) | (end[i] < start[i]) | ((uint64_t)off >> off_bits != 0) | ((uint64_t)w >> width_bits != 0);
    }
    return bad;
}

// anchor of a block: the start of its first query, or of the first query after a descent (a run boundary inside the block),
// whichever leaves fewer exceptions — at most four candidates are tried
inline uint32_t pk_anchor(const uint32_t* start, const uint32_t* end, uint64_t b0, uint64_t b1, uint32_t off_bits, uint32_t width_bits) {
    uint32_t best = start[b0], best_bad = pk_misfits(start, end, b0, b1, best, off_bits, width_bits);
    int tried = 1;
    for (uint64_t i = b0 + 1; i < b1 && best_bad && tried < 4; ++i) {
        if (start[i] >= start[i - 1]) continue;
        ++tried;
        const uint32_t bad = pk_misfits(start, end, b0, b1, start[i], off_bits, width_bits);
        if (bad < best_bad) {
            best_bad = bad;
            best = start[i];
        }
    }
    return best;
}

}  // namespace

int32_t marshal_packed_impl(uint64_t n, const uint32_t* chr, const uint32_t* start, const uint32_t* end, uint64_t n_files,
                            const uint64_t* file_offsets, int32_t threads, uint32_t width_bits, uint32_t* out_packed,
                            uint32_t* out_anchors, uint64_t run_capacity, uint64_t* out_run_offsets, uint32_t* out_run_chr,
                            uint64_t* out_n_runs, uint64_t exc_capacity, uint64_t* out_exc_index, uint32_t* out_exc_start,
                            uint32_t* out_exc_end, uint64_t* out_n_exc, uint32_t* out_width_bits) {
    if (!out_n_runs || !out_n_exc || !out_width_bits || (n && (!chr || !start || !end || !out_packed || !out_anchors)) ||
        (n_files && !file_offsets) || (run_capacity && (!out_run_offsets || !out_run_chr)) ||
        (exc_capacity && (!out_exc_index || !out_exc_start || !out_exc_end)))
        return fail(GTGPU_ERR_INVALID, "marshal_packed: null argument");
    if (width_bits > 24) return fail(GTGPU_ERR_INVALID, "marshal_packed: width_bits must be 0 (choose) or 1..24");
    for (uint64_t f = 0; f < n_files; ++f)
        if (file_offsets[f] > file_offsets[f + 1]) return fail(GTGPU_ERR_INVALID, "marshal_packed: file_offsets not monotone");
    if (n_files && (file_offsets[0] != 0 || file_offsets[n_files] != n))
        return fail(GTGPU_ERR_INVALID, "marshal_packed: file_offsets must span [0, n]");
    const uint64_t n_blocks = (n + PK_BLOCK - 1) / PK_BLOCK;
    if (width_bits == 0) {
        // choose the split on a sample of blocks spread over the batch: fewest exceptions wins, ties go to the narrower width
        const uint64_t step = std::max<uint64_t>(1, n_blocks / 2048);
        uint64_t best_bad = ~0ull;
        width_bits = 10;
        for (uint32_t wb = 6; wb <= 16; ++wb) {
            uint64_t bad = 0;
            for (uint64_t b = 0; b < n_blocks; b += step) {
                const uint64_t b0 = b * PK_BLOCK, b1 = std::min(n, b0 + PK_BLOCK);
                bad += pk_misfits(start, end, b0, b1, pk_anchor(start, end, b0, b1, 32 - wb, wb), 32 - wb, wb);
            }
            if (bad < best_bad) {
                best_bad = bad;
                width_bits = wb;
            }
        }
    }
    *out_width_bits = width_bits;
    const uint32_t off_bits = 32 - width_bits;
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(nt, (n + (1 << 20) - 1) >> 20));
    struct Part {
        std::vector<uint64_t> run_off, exc_idx;
        std::vector<uint32_t> run_chr, exc_start, exc_end;
    };
    std::vector<Part> parts(nt);
    auto work = [&](unsigned t) {
        // thread ranges are whole blocks
        const uint64_t lo = std::min(n, n_blocks * t / nt * PK_BLOCK), hi = std::min(n, n_blocks * (t + 1) / nt * PK_BLOCK);
        Part& p = parts[t];
        const uint64_t* fb = n_files ? std::lower_bound(file_offsets, file_offsets + n_files + 1, lo) : nullptr;
        const uint64_t* fb_end = n_files ? file_offsets + n_files + 1 : nullptr;
        uint32_t prev = lo ? chr[lo - 1] : 0;
        uint64_t next_file = fb != fb_end ? *fb : ~0ull;  // next file boundary at or after the current query
        const uint32_t off_lim = off_bits == 32 ? 0xFFFFFFFFu : (1u << off_bits) - 1u, w_lim = (1u << width_bits) - 1u;
        for (uint64_t b0 = lo; b0 < hi; b0 += PK_BLOCK) {
            const uint64_t b1 = std::min(hi, b0 + PK_BLOCK);
            // fast path (almost every block of a sorted file): one chromosome, no file boundary, everything fits behind the
            // first query's start — one branch-free pass
            uint32_t anchor = start[b0];
            bool plain = b0 != 0 && next_file >= b1;
            if (plain) {
                uint32_t bad = 0;
                for (uint64_t i = b0; i < b1; ++i) {
                    const uint32_t s = start[i], e = end[i], off = s - anchor, w = e - s;
                    bad |= (uint32_t)(chr[i] != prev) | (uint32_t)(s < anchor) | (uint32_t)(e < s) | (uint32_t)(off > off_lim) |
                           (uint32_t)(w > w_lim);
                    out_packed[i] = off | (w << off_bits);
                }
                plain = bad == 0;
            }
            if (plain) {
                out_anchors[b0 / PK_BLOCK] = anchor;
                continue;
            }
            anchor = pk_anchor(start, end, b0, b1, off_bits, width_bits);
            out_anchors[b0 / PK_BLOCK] = anchor;
            for (uint64_t i = b0; i < b1; ++i) {
                const uint32_t c = chr[i], s = start[i], e = end[i];
                bool cut = i == 0 || c != prev;
                while (fb != fb_end && *fb < i) ++fb;
                if (fb != fb_end && *fb == i) cut = true;
                if (cut) {
                    p.run_off.push_back(i);
                    p.run_chr.push_back(c);
                }
                prev = c;
                const uint32_t off = s - anchor, w = e - s;
                if (s < anchor || e < s || off > off_lim || w > w_lim) {
                    out_packed[i] = 0;
                    p.exc_idx.push_back(i);
                    p.exc_start.push_back(s);
                    p.exc_end.push_back(e);
                } else {
                    out_packed[i] = off | (w << off_bits);
                }
            }
            while (fb != fb_end && *fb < b1) ++fb;
            next_file = fb != fb_end ? *fb : ~0ull;
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    uint64_t n_runs = 0, n_exc = 0;
    for (auto& p : parts) {
        n_runs += p.run_off.size();
        n_exc += p.exc_idx.size();
    }
    *out_n_runs = n_runs;
    *out_n_exc = n_exc;
    if (n_runs > run_capacity || n_exc > exc_capacity)
        return fail(GTGPU_ERR_CAPACITY, "marshal_packed: run / exception capacity too small (needed counts returned)");
    uint64_t r = 0, x = 0;
    for (auto& p : parts) {
        std::copy(p.run_off.begin(), p.run_off.end(), out_run_offsets + r);
        std::copy(p.run_chr.begin(), p.run_chr.end(), out_run_chr + r);
        r += p.run_off.size();
        std::copy(p.exc_idx.begin(), p.exc_idx.end(), out_exc_index + x);
        std::copy(p.exc_start.begin(), p.exc_start.end(), out_exc_start + x);
        std::copy(p.exc_end.begin(), p.exc_end.end(), out_exc_end + x);
        x += p.exc_idx.size();
    }
    if (run_capacity || n_runs) out_run_offsets[n_runs] = n;
    return GTGPU_OK;
}

}  // namespace gtgpu

extern "C" int32_t gtgpu_marshal_compact(uint64_t n, const uint32_t* chr, const uint32_t* start, const uint32_t* end,
                                         uint64_t n_files, const uint64_t* file_offsets, int32_t threads, uint16_t* out_width16,
                                         uint64_t run_capacity, uint64_t* out_run_offsets, uint32_t* out_run_chr,
                                         uint64_t* out_n_runs, uint64_t wide_capacity, uint64_t* out_wide_index,
                                         uint32_t* out_wide_end, uint64_t* out_n_wide) try {
    try {
        return marshal_compact_impl(n, chr, start, end, n_files, file_offsets, threads, out_width16, run_capacity, out_run_offsets,
                                    out_run_chr, out_n_runs, wide_capacity, out_wide_index, out_wide_end, out_n_wide);
    } catch (const std::bad_alloc&) {
        return fail(GTGPU_ERR_NOMEM, "marshal_compact: out of host memory");
    } catch (...) {
        return fail(GTGPU_ERR_INVALID, "marshal_compact: unexpected exception");
    }
} GT_CATCH

extern "C" int32_t gtgpu_marshal_packed(uint64_t n, const uint32_t* chr, const uint32_t* start, const uint32_t* end, uint64_t n_files,
                                        const uint64_t* file_offsets, int32_t threads, uint32_t width_bits, uint32_t* out_packed,
                                        uint32_t* out_anchors, uint64_t run_capacity, uint64_t* out_run_offsets,
                                        uint32_t* out_run_chr, uint64_t* out_n_runs, uint64_t exc_capacity, uint64_t* out_exc_index,
                                        uint32_t* out_exc_start, uint32_t* out_exc_end, uint64_t* out_n_exc,
                                        uint32_t* out_width_bits) try {
    try {
        return marshal_packed_impl(n, chr, start, end, n_files, file_offsets, threads, width_bits, out_packed, out_anchors, run_capacity,
                                   out_run_offsets, out_run_chr, out_n_runs, exc_capacity, out_exc_index, out_exc_start, out_exc_end,
                                   out_n_exc, out_width_bits);
    } catch (const std::bad_alloc&) {
        return fail(GTGPU_ERR_NOMEM, "marshal_packed: out of host memory");
    } catch (...) {
        return fail(GTGPU_ERR_INVALID, "marshal_packed: unexpected exception");
    }
} GT_CATCH
