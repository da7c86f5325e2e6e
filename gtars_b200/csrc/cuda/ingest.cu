// ingest.cu — BED text -> sorted SoA queries on the device.
//
// Once the find kernel runs at > 1e11 queries/s, parsing "chr\tstart\tend\n" on a host core (~1e8 lines/s) is the
// end-to-end bottleneck by three orders of magnitude.  This file restates RegionSet::try_from
// (gtars-core/src/models/region_set.rs:60-185: line rules, u32 parsing, header / comment rows) and RegionSet::sort
// (:502-505: stable, by chromosome string then start) as kernels over the raw bytes of one (already decompressed) file:
//   1. newline positions: per-64-byte-chunk counts -> hand-written exclusive scan -> positions
//   2. one thread per line: comment / header rules, chromosome name -> dense id (hash + byte compare against the
//      caller's name table), str::parse::<u32>() for start and end; the first malformed line is reported
//   3. compaction of the kept lines (scan + scatter)
//   4. order check; when the file is not already sorted: two stable passes of the hand-written radix sort
//      (start, then chromosome rank) and a gather
// gtgpu_tokenize_bed chains the fused find kernel behind it, so a BED file goes from text to token ids without its
// regions ever existing on the host.  Chromosome names that are not in the table map to GTGPU_UNKNOWN_CHROM and sort
// after every known name (the reference sorts them by their own strings; they produce no output either way).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "common.cuh"

namespace gtgpu {

constexpr int INGEST_CHUNK = 64;  // bytes per thread in the newline passes

__global__ void ingest_count_newlines_kernel(uint64_t n_bytes, const char* __restrict__ text, uint32_t* __restrict__ counts) {
    const uint64_t n_chunks = (n_bytes + INGEST_CHUNK - 1) / INGEST_CHUNK;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
        const uint64_t a = c * INGEST_CHUNK, b = min(a + INGEST_CHUNK, n_bytes);
        uint32_t k = 0;
        if (b - a == INGEST_CHUNK) {
            const uint4* p = reinterpret_cast<const uint4*>(text + a);  // text is 256-byte aligned device scratch
#pragma unroll
            for (int i = 0; i < INGEST_CHUNK / 16; ++i) {
                const uint4 v = __ldg(p + i);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t x = w[j] ^ 0x0A0A0A0Au;  // a zero byte where the text has '\n'
                    k += __popc(~((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) | 0x7F7F7F7Fu));
                }
            }
        } else {
            for (uint64_t i = a; i < b; ++i) k += text[i] == '\n';
        }
        counts[c] = k;
    }
}

__global__ void ingest_newline_positions_kernel(uint64_t n_bytes, const char* __restrict__ text, const uint32_t* __restrict__ rank,
                                                uint32_t* __restrict__ nl_pos) {
    const uint64_t n_chunks = (n_bytes + INGEST_CHUNK - 1) / INGEST_CHUNK;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
        const uint64_t a = c * INGEST_CHUNK, b = min(a + INGEST_CHUNK, n_bytes);
        uint32_t r = rank[c];
        for (uint64_t i = a; i < b; ++i)
            if (text[i] == '\n') nl_pos[r++] = (uint32_t)i;
    }
}

struct NameTable {
    const char* blob;             // names back to back
    const uint32_t* offsets;      // n_names + 1
    const unsigned long long* hash;  // FNV-1a of every name, sorted ascending
    const uint32_t* hash_id;      // name id of the k-th sorted hash
    uint32_t n_names;
};

__device__ __forceinline__ bool is_digit(char c) { return c >= '0' && c <= '9'; }

// str::parse::<u32>(): optional '+', at least one digit, digits only, no overflow
__device__ bool parse_u32_field(const char* __restrict__ t, uint32_t a, uint32_t b, uint32_t& out) {
    if (a < b && t[a] == '+') ++a;
    if (a >= b) return false;
    unsigned long long v = 0;
    for (uint32_t i = a; i < b; ++i) {
        const char c = t[i];
        if (!is_digit(c)) return false;
        v = v * 10 + (unsigned long long)(c - '0');
        if (v > 0xFFFFFFFFull) return false;
    }
    out = (uint32_t)v;
    return true;
}

__device__ bool has_prefix(const char* __restrict__ t, uint32_t a, uint32_t b, const char* p, uint32_t n) {
    if (b - a < n) return false;
    for (uint32_t i = 0; i < n; ++i)
        if (t[a + i] != p[i]) return false;
    return true;
}

// status per line: 0 = skipped (comment / header), 1 = region, 2 = malformed
__global__ void ingest_parse_lines_kernel(uint32_t n_lines, uint32_t n_newlines, uint64_t n_bytes, const char* __restrict__ text,
                                          const uint32_t* __restrict__ nl_pos, NameTable names, uint32_t* __restrict__ keep,
                                          uint32_t* __restrict__ chr, uint32_t* __restrict__ start, uint32_t* __restrict__ end,
                                          uint32_t* __restrict__ first_bad_line) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_lines; li += stride) {
        uint32_t a = li ? nl_pos[li - 1] + 1 : 0;
        uint32_t b = li < n_newlines ? nl_pos[li] : (uint32_t)n_bytes;
        if (b > a && text[b - 1] == '\r') --b;  // BufRead::lines strips "\r\n"
        uint32_t k = 0, c = GTGPU_UNKNOWN_CHROM, s = 0, e = 0;
        if (has_prefix(text, a, b, "browser", 7) || has_prefix(text, a, b, "track", 5) || has_prefix(text, a, b, "#", 1)) {
            k = 0;
        } else {
            // tab-separated fields 0, 1, 2
            uint32_t t1 = a;
            while (t1 < b && text[t1] != '\t') ++t1;
            uint32_t t2 = t1 < b ? t1 + 1 : b;
            while (t2 < b && text[t2] != '\t') ++t2;
            uint32_t t3 = t2 < b ? t2 + 1 : b;
            while (t3 < b && text[t3] != '\t') ++t3;
            const bool three = t1 < b && t2 < b;  // at least two tabs = three parts
            const bool s_ok = three && parse_u32_field(text, t1 + 1, t2, s);
            if (li == 0 && three && !s_ok) {
                k = 0;  // a column-header first row without '#' (region_set.rs:118-131)
            } else if (!s_ok || !parse_u32_field(text, t2 + 1, t3, e)) {
                k = 2;
            } else {
                k = 1;
                unsigned long long h = 1469598103934665603ull;  // FNV-1a 64
                for (uint32_t i = a; i < t1; ++i) h = (h ^ (unsigned char)text[i]) * 1099511628211ull;
                uint32_t lo = 0, hi = names.n_names;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (names.hash[mid] < h) lo = mid + 1;
                    else hi = mid;
                }
                for (; lo < names.n_names && names.hash[lo] == h; ++lo) {  // equal hashes: compare the bytes
                    const uint32_t id = names.hash_id[lo], na = names.offsets[id], nb = names.offsets[id + 1];
                    bool same = nb - na == t1 - a;
                    for (uint32_t i = 0; same && i < nb - na; ++i) same = names.blob[na + i] == text[a + i];
                    if (same) { c = id; break; }
                }
            }
        }
        if (k == 2) atomicMin(first_bad_line, li);
        keep[li] = k == 1;
        chr[li] = c;
        start[li] = s;
        end[li] = e;
    }
}

__global__ void ingest_compact_kernel(uint32_t n_lines, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos,
                                      const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start,
                                      const uint32_t* __restrict__ end, uint32_t* __restrict__ o_chr, uint32_t* __restrict__ o_start,
                                      uint32_t* __restrict__ o_end) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += stride)
        if (keep[i]) {
            const uint32_t p = pos[i];
            o_chr[p] = chr[i];
            o_start[p] = start[i];
            o_end[p] = end[i];
        }
}

__device__ __forceinline__ uint32_t rank_of(const uint32_t* __restrict__ name_rank, uint32_t n_names, uint32_t c) {
    return c < n_names ? name_rank[c] : n_names;  // unknown names after every known one
}

__global__ void ingest_check_sorted_kernel(uint32_t n, const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start,
                                           const uint32_t* __restrict__ name_rank, uint32_t n_names, uint32_t* __restrict__ unsorted) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += stride) {
        const uint32_t ra = rank_of(name_rank, n_names, chr[i]), rb = rank_of(name_rank, n_names, chr[i + 1]);
        if (ra > rb || (ra == rb && start[i] > start[i + 1])) *unsorted = 1;
    }
}

__global__ void ingest_iota_kernel(uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = i;
}

__global__ void ingest_rank_keys_kernel(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ chr,
                                        const uint32_t* __restrict__ name_rank, uint32_t n_names, uint32_t* __restrict__ keys) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) keys[i] = rank_of(name_rank, n_names, chr[order[i]]);
}

__global__ void ingest_gather_kernel(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ chr,
                                     const uint32_t* __restrict__ start, const uint32_t* __restrict__ end, uint32_t* __restrict__ o_chr,
                                     uint32_t* __restrict__ o_start, uint32_t* __restrict__ o_end) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t j = order[i];
        o_chr[i] = chr[j];
        o_start[i] = start[j];
        o_end[i] = end[j];
    }
}

static int igrid(gtgpu_ctx* ctx, uint64_t n) {
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16));
}

// Parses and sorts on the device.  On success *d_chr / *d_start / *d_end point at n_out regions in device scratch
// (valid until the next call on this ctx).  The caller holds ctx->mu.
static int32_t parse_bed_locked(gtgpu_ctx* ctx, const char* text, uint64_t n_bytes, uint32_t n_names, const char* names,
                                const uint32_t* name_offsets, uint64_t* n_out, uint32_t** d_chr, uint32_t** d_start,
                                uint32_t** d_end) {
    if (n_bytes >= 0xFFFFFFF0ull) return fail(GTGPU_ERR_UNSUPPORTED, "parse_bed: at most 4 GiB of text per call (split at a line boundary)");
    if (n_names >= 0x7FFFFFFFu) return fail(GTGPU_ERR_INVALID, "parse_bed: too many chromosome names");
    cudaStream_t st = ctx->stream;
    *n_out = 0;
    if (n_bytes == 0) return fail(GTGPU_ERR_INVALID, "parse_bed: EmptyRegionSet (no regions in the text)");

    // ---- name table: hashes sorted for the device lookup, lexicographic ranks for the sort ----------------------------
    std::vector<unsigned long long> hash(n_names);
    std::vector<uint32_t> hash_id(n_names), rank(n_names), by_name(n_names);
    for (uint32_t i = 0; i < n_names; ++i) {
        unsigned long long h = 1469598103934665603ull;
        for (uint32_t k = name_offsets[i]; k < name_offsets[i + 1]; ++k) h = (h ^ (unsigned char)names[k]) * 1099511628211ull;
        hash[i] = h;
    }
    std::iota(hash_id.begin(), hash_id.end(), 0u);
    std::sort(hash_id.begin(), hash_id.end(), [&](uint32_t a, uint32_t b) { return hash[a] != hash[b] ? hash[a] < hash[b] : a < b; });
    std::vector<unsigned long long> hash_sorted(n_names);
    for (uint32_t i = 0; i < n_names; ++i) hash_sorted[i] = hash[hash_id[i]];
    std::iota(by_name.begin(), by_name.end(), 0u);
    auto name_of = [&](uint32_t i) { return std::string(names + name_offsets[i], names + name_offsets[i + 1]); };
    std::stable_sort(by_name.begin(), by_name.end(), [&](uint32_t a, uint32_t b) { return name_of(a) < name_of(b); });
    for (uint32_t r = 0; r < n_names; ++r) rank[by_name[r]] = r;  // equal strings cannot occur in a name table
    const uint32_t blob_bytes = n_names ? name_offsets[n_names] : 0;

    char* d_text;
    uint32_t *d_counts, *d_crank, *d_nl;
    void* d_tmp;
    char* d_names;
    const uint64_t n_chunks = (n_bytes + INGEST_CHUNK - 1) / INGEST_CHUNK;
    GT_TRY(ctx->scratch_get(SC_OUT_IDS2, n_bytes + 64, (void**)&d_text));
    GT_TRY(ctx->scratch_get(SC_COUNTS, n_chunks * 4 + 4, (void**)&d_counts));
    GT_TRY(ctx->scratch_get(SC_IN3_CHR, n_chunks * 4 + 4, (void**)&d_crank));
    const size_t names_bytes = ((size_t)blob_bytes + 15) / 16 * 16 + ((size_t)n_names + 1) * 4 + (size_t)n_names * 8 + (size_t)n_names * 4 * 2 + 64;
    GT_TRY(ctx->scratch_get(SC_SET_ID, names_bytes, (void**)&d_names));
    GT_CUDA(cudaMemcpyAsync(d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
    // layout of the name scratch: hashes (8-byte aligned first), offsets, hash ids, ranks, blob
    unsigned long long* d_hash = reinterpret_cast<unsigned long long*>(d_names);
    uint32_t* d_noff = reinterpret_cast<uint32_t*>(d_hash + n_names);
    uint32_t* d_hid = d_noff + n_names + 1;
    uint32_t* d_rank = d_hid + n_names;
    char* d_blob = reinterpret_cast<char*>(d_rank + n_names);
    if (n_names) {
        GT_CUDA(cudaMemcpyAsync(d_hash, hash_sorted.data(), (size_t)n_names * 8, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_hid, hash_id.data(), (size_t)n_names * 4, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_rank, rank.data(), (size_t)n_names * 4, cudaMemcpyHostToDevice, st));
        if (blob_bytes) GT_CUDA(cudaMemcpyAsync(d_blob, names, blob_bytes, cudaMemcpyHostToDevice, st));
    }
    GT_CUDA(cudaMemcpyAsync(d_noff, name_offsets ? name_offsets : (const uint32_t*)&blob_bytes, ((size_t)n_names + 1) * 4,
                            cudaMemcpyHostToDevice, st));

    // ---- 1. newline positions ---------------------------------------------------------------------------------------------
    GT_TRY(ctx->scratch_get(SC_IN3_START, exclusive_scan_temp_bytes(n_chunks, 4), &d_tmp));
    ingest_count_newlines_kernel<<<igrid(ctx, n_chunks), 256, 0, st>>>(n_bytes, d_text, d_counts);
    ctx->launches++;
    GT_TRY(exclusive_scan<uint32_t>(ctx, d_counts, d_crank, n_chunks, d_tmp));
    uint32_t last[2];
    GT_CUDA(cudaMemcpyAsync(&last[0], d_crank + n_chunks - 1, 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaMemcpyAsync(&last[1], d_counts + n_chunks - 1, 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));  // the host's text buffer may be reused after this point
    const uint32_t n_newlines = last[0] + last[1];
    const uint32_t n_lines = n_newlines + (text[n_bytes - 1] != '\n' ? 1u : 0u);
    GT_TRY(ctx->scratch_get(SC_IN3_END, (size_t)n_newlines * 4 + 4, (void**)&d_nl));
    ingest_newline_positions_kernel<<<igrid(ctx, n_chunks), 256, 0, st>>>(n_bytes, d_text, d_crank, d_nl);
    ctx->launches++;

    // ---- 2. parse, 3. compact ---------------------------------------------------------------------------------------------
    uint32_t *d_keep, *d_pos, *p_chr, *p_start, *p_end, *o_chr, *o_start, *o_end;
    uint64_t* d_misc;
    const size_t lb = (size_t)n_lines * 4 + 4;
    GT_TRY(ctx->scratch_get(SC_IN3_START, std::max(exclusive_scan_temp_bytes(n_lines, 4), radix_sort_temp_bytes(n_lines)), &d_tmp));
    GT_TRY(ctx->scratch_get(SC_BARCODE, lb, (void**)&d_keep));
    GT_TRY(ctx->scratch_get(SC_FILE_TOK2, lb, (void**)&d_pos));
    GT_TRY(ctx->scratch_get(SC_IN2_CHR, lb, (void**)&p_chr));
    GT_TRY(ctx->scratch_get(SC_IN2_START, lb, (void**)&p_start));
    GT_TRY(ctx->scratch_get(SC_IN2_END, lb, (void**)&p_end));
    GT_TRY(ctx->scratch_get(SC_CHR, lb, (void**)&o_chr));
    GT_TRY(ctx->scratch_get(SC_START, lb, (void**)&o_start));
    GT_TRY(ctx->scratch_get(SC_END, lb, (void**)&o_end));
    GT_TRY(ctx->scratch_get(SC_MISC, 64, (void**)&d_misc));
    uint32_t* d_flags = reinterpret_cast<uint32_t*>(d_misc);  // [0] first malformed line, [1] unsorted
    const uint32_t init[2] = {0xFFFFFFFFu, 0u};
    GT_CUDA(cudaMemcpyAsync(d_flags, init, 8, cudaMemcpyHostToDevice, st));
    NameTable nt{d_blob, d_noff, d_hash, d_hid, n_names};
    ingest_parse_lines_kernel<<<igrid(ctx, n_lines), 256, 0, st>>>(n_lines, n_newlines, n_bytes, d_text, d_nl, nt, d_keep, p_chr, p_start,
                                                                   p_end, d_flags);
    ctx->launches++;
    GT_TRY(exclusive_scan<uint32_t>(ctx, d_keep, d_pos, n_lines, d_tmp));
    ingest_compact_kernel<<<igrid(ctx, n_lines), 256, 0, st>>>(n_lines, d_keep, d_pos, p_chr, p_start, p_end, o_chr, o_start, o_end);
    ctx->launches++;
    uint32_t tail[2], flags[2];
    GT_CUDA(cudaMemcpyAsync(&tail[0], d_pos + n_lines - 1, 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaMemcpyAsync(&tail[1], d_keep + n_lines - 1, 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaMemcpyAsync(flags, d_flags, 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));
    if (flags[0] != 0xFFFFFFFFu)
        return fail(GTGPU_ERR_INVALID, "parse_bed: RegionParseError: cannot parse start / end position on line " + std::to_string((uint64_t)flags[0] + 1));
    const uint32_t n = tail[0] + tail[1];
    if (n == 0) return fail(GTGPU_ERR_INVALID, "parse_bed: EmptyRegionSet (no regions in the text)");

    // ---- 4. RegionSet::sort: stable by (chromosome string, start) -----------------------------------------------------------
    ingest_check_sorted_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, o_chr, o_start, d_rank, n_names, d_flags + 1);
    ctx->launches++;
    GT_CUDA(cudaMemcpyAsync(flags, d_flags, 8, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));
    if (flags[1]) {
        uint32_t *k_a = d_keep, *v_a = d_pos, *k_b, *v_b;  // the line-sized arrays are free again
        GT_TRY(ctx->scratch_get(SC_OUT_IDS, (size_t)n * 8 + 8, (void**)&k_b));
        v_b = k_b + n;
        GT_CUDA(cudaMemcpyAsync(k_a, o_start, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        ingest_iota_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, v_a);
        ctx->launches++;
        int in_b = 0;
        GT_TRY(radix_sort_pairs(ctx, n, k_a, v_a, k_b, v_b, 32, d_tmp, &in_b));
        if (in_b) { std::swap(k_a, k_b); std::swap(v_a, v_b); }
        ingest_rank_keys_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, v_a, o_chr, d_rank, n_names, k_a);
        ctx->launches++;
        int bits = 1;
        while (bits < 32 && (1ull << bits) < (uint64_t)n_names + 1) ++bits;
        GT_TRY(radix_sort_pairs(ctx, n, k_a, v_a, k_b, v_b, bits, d_tmp, &in_b));
        const uint32_t* order = in_b ? v_b : v_a;
        ingest_gather_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, order, o_chr, o_start, o_end, p_chr, p_start, p_end);
        ctx->launches++;
        o_chr = p_chr; o_start = p_start; o_end = p_end;
    }
    GT_CUDA(cudaGetLastError());
    *n_out = n;
    *d_chr = o_chr;
    *d_start = o_start;
    *d_end = o_end;
    return GTGPU_OK;
}

static int32_t to_host_buf(gtgpu_ctx* ctx, const uint32_t* d_src, uint64_t n, gtgpu_buf** out) {
    gtgpu_buf* buf = new gtgpu_buf();
    buf->ctx = ctx;
    buf->len = n;
    int32_t s = ctx->pinned_get(n * 4, &buf->block);
    if (s != GTGPU_OK) {
        delete buf;
        return s;
    }
    cudaError_t e = n ? cudaMemcpyAsync(buf->block.ptr, d_src, n * 4, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        ctx->pinned_put(buf->block);
        delete buf;
        return fail(GTGPU_ERR_CUDA, std::string("ingest: D2H: ") + cudaGetErrorString(e));
    }
    *out = buf;
    return GTGPU_OK;
}

}  // namespace gtgpu

using namespace gtgpu;

extern "C" int32_t gtgpu_parse_bed(gtgpu_ctx* ctx, const char* text, uint64_t n_bytes, uint32_t n_names, const char* names,
                                   const uint32_t* name_offsets, uint64_t* out_n, gtgpu_buf** out_chr, gtgpu_buf** out_start,
                                   gtgpu_buf** out_end) {
    if (!ctx || !out_n || !out_chr || !out_start || !out_end || (n_bytes && !text) || (n_names && (!names || !name_offsets)))
        return fail(GTGPU_ERR_INVALID, "parse_bed: null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    uint32_t *d_chr, *d_start, *d_end;
    uint64_t n = 0;
    GT_TRY(parse_bed_locked(ctx, text, n_bytes, n_names, names, name_offsets, &n, &d_chr, &d_start, &d_end));
    gtgpu_buf* bufs[3] = {nullptr, nullptr, nullptr};
    const uint32_t* src[3] = {d_chr, d_start, d_end};
    for (int k = 0; k < 3; ++k) {
        int32_t s = to_host_buf(ctx, src[k], n, &bufs[k]);
        if (s != GTGPU_OK) {
            for (int j = 0; j < k; ++j) { ctx->pinned_put(bufs[j]->block); delete bufs[j]; }
            return s;
        }
    }
    *out_n = n;
    *out_chr = bufs[0];
    *out_start = bufs[1];
    *out_end = bufs[2];
    return GTGPU_OK;
}

extern "C" int32_t gtgpu_tokenize_bed(gtgpu_index* ix, const char* text, uint64_t n_bytes, uint32_t n_names, const char* names,
                                      const uint32_t* name_offsets, uint32_t unk_id, gtgpu_buf** out_ids) {
    if (!ix || !out_ids || (n_bytes && !text) || (n_names && (!names || !name_offsets)))
        return fail(GTGPU_ERR_INVALID, "tokenize_bed: null argument");
    gtgpu_ctx* ctx = ix->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    uint32_t *d_chr, *d_start, *d_end, *d_ids = nullptr;
    uint64_t n = 0, total = 0;
    GT_TRY(parse_bed_locked(ctx, text, n_bytes, n_names, names, name_offsets, &n, &d_chr, &d_start, &d_end));
    GT_TRY(fused_find_all(ix, n, 0, nullptr, d_chr, d_start, d_end, nullptr, nullptr, &d_ids, &total));
    if (total == 0) {  // Tokenizer::tokenize: a call without a single token yields [unk] (tokenizer.rs:156-160)
        gtgpu_buf* buf = new gtgpu_buf();
        buf->ctx = ctx;
        buf->len = 1;
        int32_t s = ctx->pinned_get(4, &buf->block);
        if (s != GTGPU_OK) {
            delete buf;
            return s;
        }
        *reinterpret_cast<uint32_t*>(buf->block.ptr) = unk_id;
        *out_ids = buf;
        return GTGPU_OK;
    }
    return to_host_buf(ctx, d_ids, total, out_ids);
}
