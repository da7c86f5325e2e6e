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
        // BufRead::lines strips "\n" and a "\r" right before it; a final line WITHOUT a newline keeps its '\r' (the reference then
        // fails to parse the field, and so do we)
        if (li < n_newlines && b > a && text[b - 1] == '\r') --b;
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

// ---- fragment files (gtars-tokenizers/src/utils/fragments.rs:12-82) ---------------------------------------------------------
__device__ __forceinline__ bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }  // ASCII White_Space

// One thread per line: '#' lines are skipped; split_whitespace must give at least 5 fields (chr start end barcode ...),
// start and end parse as u32.  The barcode is carried as (FNV-1a hash, byte span).
__global__ void ingest_parse_fragments_kernel(uint32_t n_lines, uint32_t n_newlines, uint64_t n_bytes, const char* __restrict__ text,
                                              const uint32_t* __restrict__ nl_pos, NameTable names, uint32_t* __restrict__ keep,
                                              uint32_t* __restrict__ chr, uint32_t* __restrict__ start, uint32_t* __restrict__ end,
                                              unsigned long long* __restrict__ bc_hash, uint32_t* __restrict__ bc_off,
                                              uint32_t* __restrict__ bc_len, uint32_t* __restrict__ first_bad_line) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_lines; li += stride) {
        uint32_t a = li ? nl_pos[li - 1] + 1 : 0;
        uint32_t b = li < n_newlines ? nl_pos[li] : (uint32_t)n_bytes;
        if (li < n_newlines && b > a && text[b - 1] == '\r') --b;  // only a '\r' in front of a '\n' is part of the line ending
        uint32_t k = 0, c = GTGPU_UNKNOWN_CHROM, s = 0, e = 0, bo = 0, bl = 0;
        unsigned long long h = 1469598103934665603ull;
        if (!(b > a && text[a] == '#')) {
            uint32_t fa[5], fb[5], nf = 0, p = a;
            while (nf < 5) {
                while (p < b && is_space(text[p])) ++p;
                if (p >= b) break;
                fa[nf] = p;
                while (p < b && !is_space(text[p])) ++p;
                fb[nf++] = p;
            }
            if (nf < 5 || !parse_u32_field(text, fa[1], fb[1], s) || !parse_u32_field(text, fa[2], fb[2], e)) {
                k = 2;
            } else {
                k = 1;
                unsigned long long hc = 1469598103934665603ull;
                for (uint32_t i = fa[0]; i < fb[0]; ++i) hc = (hc ^ (unsigned char)text[i]) * 1099511628211ull;
                uint32_t lo = 0, hi = names.n_names;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (names.hash[mid] < hc) lo = mid + 1;
                    else hi = mid;
                }
                for (; lo < names.n_names && names.hash[lo] == hc; ++lo) {
                    const uint32_t id = names.hash_id[lo], na = names.offsets[id], nb = names.offsets[id + 1];
                    bool same = nb - na == fb[0] - fa[0];
                    for (uint32_t i = 0; same && i < nb - na; ++i) same = names.blob[na + i] == text[fa[0] + i];
                    if (same) { c = id; break; }
                }
                bo = fa[3];
                bl = fb[3] - fa[3];
                for (uint32_t i = fa[3]; i < fb[3]; ++i) h = (h ^ (unsigned char)text[i]) * 1099511628211ull;
                if (h == 0) h = 1;  // 0 marks an empty slot of the barcode table
            }
        }
        if (k == 2) atomicMin(first_bad_line, li);
        keep[li] = k == 1;
        chr[li] = c;
        start[li] = s;
        end[li] = e;
        bc_hash[li] = h;
        bc_off[li] = bo;
        bc_len[li] = bl;
    }
}

__global__ void ingest_compact_fragments_kernel(uint32_t n_lines, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos,
                                                const uint32_t* __restrict__ chr, const uint32_t* __restrict__ start,
                                                const uint32_t* __restrict__ end, const unsigned long long* __restrict__ h,
                                                const uint32_t* __restrict__ bo, const uint32_t* __restrict__ bl,
                                                uint32_t* __restrict__ o_chr, uint32_t* __restrict__ o_start, uint32_t* __restrict__ o_end,
                                                unsigned long long* __restrict__ o_h, uint32_t* __restrict__ o_bo, uint32_t* __restrict__ o_bl) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += stride)
        if (keep[i]) {
            const uint32_t p = pos[i];
            o_chr[p] = chr[i];
            o_start[p] = start[i];
            o_end[p] = end[i];
            o_h[p] = h[i];
            o_bo[p] = bo[i];
            o_bl[p] = bl[i];
        }
}

// Barcode table (open addressing on the 64-bit hash): every fragment finds or claims its barcode's slot and lowers the
// slot's first-appearance index to its own position.
__global__ void ingest_barcode_insert_kernel(uint32_t n, const unsigned long long* __restrict__ h, uint32_t mask,
                                             unsigned long long* __restrict__ keys, uint32_t* __restrict__ first,
                                             uint32_t* __restrict__ slot_of) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const unsigned long long key = h[k];
        uint32_t slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(keys + slot, 0ull, key);
            if (prev == 0ull || prev == key) break;
            slot = (slot + 1) & mask;
        }
        atomicMin(first + slot, k);
        slot_of[k] = slot;
    }
}

// 1 for the fragment that is the first appearance of its barcode; also checks that a fragment's barcode bytes equal
// those of the first appearance (two different barcodes with one 64-bit hash would otherwise be merged silently).
__global__ void ingest_barcode_heads_kernel(uint32_t n, const char* __restrict__ text, const uint32_t* __restrict__ slot_of,
                                            const uint32_t* __restrict__ first, const uint32_t* __restrict__ bo,
                                            const uint32_t* __restrict__ bl, uint32_t* __restrict__ head, uint32_t* __restrict__ collision) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t f = first[slot_of[k]];
        head[k] = f == k;
        if (f != k) {
            bool same = bl[f] == bl[k];
            for (uint32_t i = 0; same && i < bl[k]; ++i) same = text[bo[f] + i] == text[bo[k] + i];
            if (!same) *collision = 1;
        }
    }
}

__global__ void ingest_barcode_ids_kernel(uint32_t n, const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ first,
                                          const uint32_t* __restrict__ rank, const uint32_t* __restrict__ bo, const uint32_t* __restrict__ bl,
                                          uint32_t* __restrict__ bc_id, uint32_t* __restrict__ spans) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t f = first[slot_of[k]];
        const uint32_t id = rank[f];  // first appearances are numbered in file order by the scan over `head`
        bc_id[k] = id;
        if (f == k) {
            spans[2 * id] = bo[k];
            spans[2 * id + 1] = bl[k];
        }
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
    // the text is either the caller's host buffer or already on the device (gtgpu_*_gz: inflated there, ctx->ingest_d_text)
    if (ctx->ingest_d_text) d_text = (char*)ctx->ingest_d_text;
    else GT_TRY(ctx->scratch_get(SC_OUT_IDS2, n_bytes + 64, (void**)&d_text));
    GT_TRY(ctx->scratch_get(SC_COUNTS, n_chunks * 4 + 4, (void**)&d_counts));
    GT_TRY(ctx->scratch_get(SC_IN3_CHR, n_chunks * 4 + 4, (void**)&d_crank));
    const size_t names_bytes = ((size_t)blob_bytes + 15) / 16 * 16 + ((size_t)n_names + 1) * 4 + (size_t)n_names * 8 + (size_t)n_names * 4 * 2 + 64;
    GT_TRY(ctx->scratch_get(SC_SET_ID, names_bytes, (void**)&d_names));
    if (!ctx->ingest_d_text) GT_CUDA(cudaMemcpyAsync(d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
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
    char last_byte = 0;  // read back from the device: with device-resident text there is no host copy
    GT_CUDA(cudaMemcpyAsync(&last_byte, d_text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));  // the host's text buffer may be reused after this point
    const uint32_t n_newlines = last[0] + last[1];
    const uint32_t n_lines = n_newlines + (last_byte != '\n' ? 1u : 0u);
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
                                   gtgpu_buf** out_end) try {
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
} GT_CATCH

extern "C" int32_t gtgpu_tokenize_bed(gtgpu_index* ix, const char* text, uint64_t n_bytes, uint32_t n_names, const char* names,
                                      const uint32_t* name_offsets, uint32_t unk_id, gtgpu_buf** out_ids) try {
    if (!ix || !out_ids || (n_bytes && !text) || (n_names && (!names || !name_offsets)))
        return fail(GTGPU_ERR_INVALID, "tokenize_bed: null argument");
    gtgpu_ctx* ctx = ix->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    return tokenize_bed_locked(ix, text, n_bytes, n_names, names, name_offsets, unk_id, out_ids);
} GT_CATCH

// the caller holds ctx->mu; `text` is ignored when ctx->ingest_d_text points at the text on the device
int32_t gtgpu::tokenize_bed_locked(gtgpu_index* ix, const char* text, uint64_t n_bytes, uint32_t n_names, const char* names,
                                   const uint32_t* name_offsets, uint32_t unk_id, gtgpu_buf** out_ids) {
    gtgpu_ctx* ctx = ix->ctx;
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

// tokenize_fragment_file (fragments.rs:61-82) from the file's text: parse on the device, barcodes -> dense ids in
// first-appearance order through a device hash table, then the fragment tokenizer core.
extern "C" int32_t gtgpu_tokenize_fragments_text(gtgpu_index* ix, const char* text, uint64_t n_bytes, uint32_t n_names,
                                                 const char* names, const uint32_t* name_offsets, uint32_t unk_id,
                                                 uint32_t* out_n_barcodes, gtgpu_buf** out_barcode_spans,
                                                 gtgpu_buf** out_barcode_offsets, gtgpu_buf** out_ids) try {
    if (!ix || !out_n_barcodes || !out_barcode_spans || !out_barcode_offsets || !out_ids || (n_bytes && !text) ||
        (n_names && (!names || !name_offsets)))
        return fail(GTGPU_ERR_INVALID, "tokenize_fragments_text: null argument");
    gtgpu_ctx* ctx = ix->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    return tokenize_fragments_text_locked(ix, text, n_bytes, n_names, names, name_offsets, unk_id, out_n_barcodes, out_barcode_spans,
                                          out_barcode_offsets, out_ids);
} GT_CATCH

// the caller holds ctx->mu; `text` is ignored when ctx->ingest_d_text points at the text on the device
int32_t gtgpu::tokenize_fragments_text_locked(gtgpu_index* ix, const char* text, uint64_t n_bytes, uint32_t n_names, const char* names,
                                              const uint32_t* name_offsets, uint32_t unk_id, uint32_t* out_n_barcodes,
                                              gtgpu_buf** out_barcode_spans, gtgpu_buf** out_barcode_offsets, gtgpu_buf** out_ids) {
    if (n_bytes >= 0xFFFFFFF0ull) return fail(GTGPU_ERR_UNSUPPORTED, "tokenize_fragments_text: at most 4 GiB of text per call");
    gtgpu_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;

    uint32_t n = 0, n_barcodes = 0;
    uint32_t *o_chr = nullptr, *o_start = nullptr, *o_end = nullptr, *d_bcid = nullptr, *d_spans = nullptr;
    if (n_bytes) {
        // ---- name table (as in parse_bed_locked) -------------------------------------------------------------------------
        std::vector<unsigned long long> hash(n_names), hash_sorted(n_names);
        std::vector<uint32_t> hash_id(n_names);
        for (uint32_t i = 0; i < n_names; ++i) {
            unsigned long long h = 1469598103934665603ull;
            for (uint32_t k = name_offsets[i]; k < name_offsets[i + 1]; ++k) h = (h ^ (unsigned char)names[k]) * 1099511628211ull;
            hash[i] = h;
        }
        std::iota(hash_id.begin(), hash_id.end(), 0u);
        std::sort(hash_id.begin(), hash_id.end(), [&](uint32_t a, uint32_t b) { return hash[a] != hash[b] ? hash[a] < hash[b] : a < b; });
        for (uint32_t i = 0; i < n_names; ++i) hash_sorted[i] = hash[hash_id[i]];
        const uint32_t blob_bytes = n_names ? name_offsets[n_names] : 0;
        char *d_text, *d_names;
        uint32_t *d_counts, *d_crank, *d_nl;
        void* d_tmp;
        const uint64_t n_chunks = (n_bytes + INGEST_CHUNK - 1) / INGEST_CHUNK;
        if (ctx->ingest_d_text) d_text = (char*)ctx->ingest_d_text;
        else GT_TRY(ctx->scratch_get(SC_OUT_IDS2, n_bytes + 64, (void**)&d_text));
        GT_TRY(ctx->scratch_get(SC_COUNTS, n_chunks * 4 + 4, (void**)&d_counts));
        GT_TRY(ctx->scratch_get(SC_IN3_CHR, n_chunks * 4 + 4, (void**)&d_crank));
        GT_TRY(ctx->scratch_get(SC_SET_ID, (size_t)n_names * 16 + ((size_t)n_names + 1) * 4 + blob_bytes + 64, (void**)&d_names));
        if (!ctx->ingest_d_text) GT_CUDA(cudaMemcpyAsync(d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
        unsigned long long* d_hash = reinterpret_cast<unsigned long long*>(d_names);
        uint32_t* d_noff = reinterpret_cast<uint32_t*>(d_hash + n_names);
        uint32_t* d_hid = d_noff + n_names + 1;
        char* d_blob = reinterpret_cast<char*>(d_hid + n_names);
        if (n_names) {
            GT_CUDA(cudaMemcpyAsync(d_hash, hash_sorted.data(), (size_t)n_names * 8, cudaMemcpyHostToDevice, st));
            GT_CUDA(cudaMemcpyAsync(d_hid, hash_id.data(), (size_t)n_names * 4, cudaMemcpyHostToDevice, st));
            if (blob_bytes) GT_CUDA(cudaMemcpyAsync(d_blob, names, blob_bytes, cudaMemcpyHostToDevice, st));
        }
        GT_CUDA(cudaMemcpyAsync(d_noff, name_offsets ? name_offsets : (const uint32_t*)&blob_bytes, ((size_t)n_names + 1) * 4,
                                cudaMemcpyHostToDevice, st));
        // ---- lines -------------------------------------------------------------------------------------------------------------
        GT_TRY(ctx->scratch_get(SC_IN3_START, exclusive_scan_temp_bytes(n_chunks, 4), &d_tmp));
        ingest_count_newlines_kernel<<<igrid(ctx, n_chunks), 256, 0, st>>>(n_bytes, d_text, d_counts);
        ctx->launches++;
        GT_TRY(exclusive_scan<uint32_t>(ctx, d_counts, d_crank, n_chunks, d_tmp));
        uint32_t last[2];
        GT_CUDA(cudaMemcpyAsync(&last[0], d_crank + n_chunks - 1, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaMemcpyAsync(&last[1], d_counts + n_chunks - 1, 4, cudaMemcpyDeviceToHost, st));
        char last_byte = 0;  // read back from the device: with device-resident text there is no host copy
        GT_CUDA(cudaMemcpyAsync(&last_byte, d_text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        const uint32_t n_newlines = last[0] + last[1];
        const uint32_t n_lines = n_newlines + (last_byte != '\n' ? 1u : 0u);
        GT_TRY(ctx->scratch_get(SC_IN3_END, (size_t)n_newlines * 4 + 4, (void**)&d_nl));
        ingest_newline_positions_kernel<<<igrid(ctx, n_chunks), 256, 0, st>>>(n_bytes, d_text, d_crank, d_nl);
        ctx->launches++;
        // ---- parse + compact ---------------------------------------------------------------------------------------------------
        const size_t lb = (size_t)n_lines * 4 + 8;
        uint32_t *d_keep, *d_pos, *p_chr, *p_start, *p_end, *p_bo, *p_bl, *o_bo, *o_bl;
        unsigned long long *p_h, *o_h;
        uint64_t* d_misc;
        GT_TRY(ctx->scratch_get(SC_IN3_START, exclusive_scan_temp_bytes(n_lines, 4), &d_tmp));
        GT_TRY(ctx->scratch_get(SC_ING_0, lb, (void**)&d_keep));
        GT_TRY(ctx->scratch_get(SC_ING_1, lb, (void**)&d_pos));
        GT_TRY(ctx->scratch_get(SC_IN2_CHR, lb, (void**)&p_chr));
        GT_TRY(ctx->scratch_get(SC_IN2_START, lb, (void**)&p_start));
        GT_TRY(ctx->scratch_get(SC_IN2_END, lb, (void**)&p_end));
        GT_TRY(ctx->scratch_get(SC_ING_2, lb * 2, (void**)&p_h));
        GT_TRY(ctx->scratch_get(SC_ING_3, lb, (void**)&p_bo));
        GT_TRY(ctx->scratch_get(SC_ING_4, lb, (void**)&p_bl));
        GT_TRY(ctx->scratch_get(SC_CHR, lb, (void**)&o_chr));
        GT_TRY(ctx->scratch_get(SC_START, lb, (void**)&o_start));
        GT_TRY(ctx->scratch_get(SC_END, lb, (void**)&o_end));
        GT_TRY(ctx->scratch_get(SC_BARCODE, lb, (void**)&d_bcid));
        GT_TRY(ctx->scratch_get(SC_ING_5, lb * 2, (void**)&o_h));
        GT_TRY(ctx->scratch_get(SC_ING_6, lb, (void**)&o_bo));
        GT_TRY(ctx->scratch_get(SC_ING_7, lb, (void**)&o_bl));
        GT_TRY(ctx->scratch_get(SC_MISC, 64, (void**)&d_misc));
        uint32_t* d_flags = reinterpret_cast<uint32_t*>(d_misc);  // [0] first malformed line, [1] hash collision
        const uint32_t init[2] = {0xFFFFFFFFu, 0u};
        GT_CUDA(cudaMemcpyAsync(d_flags, init, 8, cudaMemcpyHostToDevice, st));
        NameTable nt{d_blob, d_noff, d_hash, d_hid, n_names};
        ingest_parse_fragments_kernel<<<igrid(ctx, n_lines), 256, 0, st>>>(n_lines, n_newlines, n_bytes, d_text, d_nl, nt, d_keep, p_chr,
                                                                           p_start, p_end, p_h, p_bo, p_bl, d_flags);
        ctx->launches++;
        GT_TRY(exclusive_scan<uint32_t>(ctx, d_keep, d_pos, n_lines, d_tmp));
        ingest_compact_fragments_kernel<<<igrid(ctx, n_lines), 256, 0, st>>>(n_lines, d_keep, d_pos, p_chr, p_start, p_end, p_h, p_bo, p_bl,
                                                                             o_chr, o_start, o_end, o_h, o_bo, o_bl);
        ctx->launches++;
        uint32_t tail[2], flags[2];
        GT_CUDA(cudaMemcpyAsync(&tail[0], d_pos + n_lines - 1, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaMemcpyAsync(&tail[1], d_keep + n_lines - 1, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaMemcpyAsync(flags, d_flags, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        if (flags[0] != 0xFFFFFFFFu)
            return fail(GTGPU_ERR_INVALID, "tokenize_fragments_text: Invalid fragment file detected at line: " + std::to_string(flags[0]));
        n = tail[0] + tail[1];
        if (n) {
            // ---- barcodes -> dense ids in first-appearance order -----------------------------------------------------------------
            uint64_t cap = 1024;
            while (cap < 2ull * n) cap <<= 1;
            unsigned long long* d_keys;
            uint32_t *d_first, *d_slot_of = d_keep, *d_head = d_pos, *d_rank = p_chr;  // line-sized arrays are free again
            GT_TRY(ctx->scratch_get(SC_MATRIX, cap * 12, (void**)&d_keys));
            d_first = reinterpret_cast<uint32_t*>(d_keys + cap);
            GT_CUDA(cudaMemsetAsync(d_keys, 0, cap * 8, st));
            GT_CUDA(cudaMemsetAsync(d_first, 0xFF, cap * 4, st));
            ingest_barcode_insert_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, o_h, (uint32_t)(cap - 1), d_keys, d_first, d_slot_of);
            ingest_barcode_heads_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, d_text, d_slot_of, d_first, o_bo, o_bl, d_head, d_flags + 1);
            ctx->launches += 2;
            GT_TRY(exclusive_scan<uint32_t>(ctx, d_head, d_rank, n, d_tmp));
            GT_CUDA(cudaMemcpyAsync(&tail[0], d_rank + n - 1, 4, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaMemcpyAsync(&tail[1], d_head + n - 1, 4, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaMemcpyAsync(flags, d_flags, 8, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            if (flags[1]) return fail(GTGPU_ERR_UNSUPPORTED, "tokenize_fragments_text: two barcodes share a 64-bit hash; use the array entry point");
            n_barcodes = tail[0] + tail[1];
            GT_TRY(ctx->scratch_get(SC_FILE_OFFS, (size_t)n_barcodes * 8 + 8, (void**)&d_spans));
            ingest_barcode_ids_kernel<<<igrid(ctx, n), 256, 0, st>>>(n, d_slot_of, d_first, d_rank, o_bo, o_bl, d_bcid, d_spans);
            ctx->launches++;
            GT_CUDA(cudaGetLastError());
        }
    }
    // ---- results -----------------------------------------------------------------------------------------------------------------
    *out_n_barcodes = n_barcodes;
    gtgpu_buf* spans = nullptr;
    GT_TRY(to_host_buf(ctx, d_spans, (uint64_t)n_barcodes * 2, &spans));
    gtgpu_buf* offs = new gtgpu_buf();
    offs->ctx = ctx;
    offs->len = (uint64_t)n_barcodes + 1;
    offs->elem_size = 8;
    int32_t status = ctx->pinned_get(offs->len * 8, &offs->block);
    gtgpu_buf* ids = nullptr;
    if (status == GTGPU_OK) {
        if (n) {
            status = tokenize_fragments_core(ix, n, o_chr, o_start, o_end, d_bcid, n_barcodes, unk_id, (uint64_t*)offs->block.ptr, &ids);
        } else {
            *(uint64_t*)offs->block.ptr = 0;
            status = to_host_buf(ctx, nullptr, 0, &ids);
        }
    }
    if (status != GTGPU_OK) {
        ctx->pinned_put(spans->block);
        delete spans;
        if (offs->block.ptr) ctx->pinned_put(offs->block);
        delete offs;
        return status;
    }
    *out_barcode_spans = spans;
    *out_barcode_offsets = offs;
    *out_ids = ids;
    return GTGPU_OK;
}
