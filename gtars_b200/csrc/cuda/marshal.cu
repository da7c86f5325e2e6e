// marshal.cu — host-side marshalling helpers of the C ABI (no device work): what a caller does between its own region
// arrays and the compact wire format of gtgpu_tokenize_files_compact.  Multi-threaded, one pass over the queries.
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
