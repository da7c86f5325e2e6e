// fragments.cu — tokenize_fragment_file on the device (gtars-tokenizers/src/utils/fragments.rs:12-82).
//
// Every fragment is its own Tokenizer::tokenize call, so a fragment without a hit (or on an unknown chromosome)
// contributes one unk id; ids are appended to the fragment's barcode list in input order.  Device plan:
//   1. fused find over all fragments with per-fragment offsets            (kernels.cu, our kernel)
//   2. stable LSD radix sort of (barcode id, fragment index) by barcode   (sort.cu, hand-written)
//   3. tokens per fragment in sorted order = max(hits, 1); exclusive scan (sort.cu, hand-written)
//   4. barcode offsets by binary search over the sorted keys; scatter-copy of every fragment's ids
// Output is barcode-major: out_barcode_offsets[n_barcodes + 1] + ids.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace gtgpu {

__global__ void iota_kernel(uint64_t n, uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}

// tokens of the k-th fragment in barcode order
__global__ void frag_token_counts_kernel(uint64_t n, const uint32_t* __restrict__ order, const uint64_t* __restrict__ offsets,
                                         unsigned long long* __restrict__ counts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t i = order[k];
        const uint64_t c = offsets[i + 1] - offsets[i];
        counts[k] = c ? c : 1;
    }
}

__global__ void frag_barcode_offsets_kernel(uint32_t n_barcodes, uint64_t n, const uint32_t* __restrict__ sorted_bc,
                                            const unsigned long long* __restrict__ dst,
                                            const unsigned long long* __restrict__ last_count, uint64_t* __restrict__ out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_barcodes) return;
    uint64_t lo = 0, hi = n;  // first k with sorted_bc[k] >= b
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (sorted_bc[mid] < b) lo = mid + 1;
        else hi = mid;
    }
    out[b] = lo < n ? dst[lo] : (n ? dst[n - 1] + last_count[n - 1] : 0);
}

__global__ void frag_scatter_kernel(uint64_t n, const uint32_t* __restrict__ order, const uint64_t* __restrict__ offsets,
                                    const unsigned long long* __restrict__ dst, const uint32_t* __restrict__ raw_ids, uint32_t unk_id,
                                    uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t i = order[k];
        const uint64_t a = offsets[i], b = offsets[i + 1];
        uint64_t d = dst[k];
        if (a == b) {
            out[d] = unk_id;
        } else {
            for (uint64_t j = a; j < b; ++j) out[d++] = raw_ids[j];
        }
    }
}

// frag_scatter_kernel with a caller-provided output: ids past `capacity` are dropped; the thread of the last fragment (in
// barcode order) reports the total, or ~0 when the raw hits did not fit their staging buffer (raw_total > raw_cap).
__global__ void frag_scatter_cap_kernel(uint64_t n, const uint32_t* __restrict__ order, const uint64_t* __restrict__ offsets,
                                        const unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ counts,
                                        const uint32_t* __restrict__ raw_ids, uint32_t unk_id, uint32_t* __restrict__ out,
                                        uint64_t capacity, const uint64_t* __restrict__ raw_total, uint64_t raw_cap,
                                        uint64_t* __restrict__ out_total) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t i = order[k];
        const uint64_t a = offsets[i], b = offsets[i + 1];
        uint64_t d = dst[k];
        if (a == b) {
            if (d < capacity) out[d] = unk_id;
        } else {
            for (uint64_t j = a; j < b; ++j, ++d)
                if (d < capacity) out[d] = raw_ids[j];
        }
        if (k == n - 1) *out_total = *raw_total > raw_cap ? ~0ull : dst[k] + counts[k];
    }
}

}  // namespace gtgpu

using namespace gtgpu;

namespace gtgpu {

// Device-resident core (the caller holds ctx->mu): fragments in d_chr / d_start / d_end, dense barcode ids in d_bc.
int32_t tokenize_fragments_core(gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start, const uint32_t* d_end,
                                uint32_t* d_bc, uint32_t n_barcodes, uint32_t unk_id, uint64_t* out_barcode_offsets,
                                gtgpu_buf** out_ids) {
    gtgpu_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    uint32_t *d_bc_sorted, *d_idx, *d_order, *d_raw = nullptr, *d_out = nullptr;
    uint64_t *d_off, *d_bco, *d_misc;
    unsigned long long *d_cnt, *d_dst;
    void* d_ws;
    GT_TRY(ctx->scratch_get(SC_IN2_CHR, n * 4, (void**)&d_bc_sorted));
    GT_TRY(ctx->scratch_get(SC_IN2_START, n * 4, (void**)&d_idx));
    GT_TRY(ctx->scratch_get(SC_IN2_END, n * 4, (void**)&d_order));
    GT_TRY(ctx->scratch_get(SC_OUT_OFFS, (n + 1) * 8, (void**)&d_off));
    GT_TRY(ctx->scratch_get(SC_COUNTS, (n + 1) * 8, (void**)&d_cnt));
    GT_TRY(ctx->scratch_get(SC_IN3_CHR, (n + 1) * 8, (void**)&d_dst));
    GT_TRY(ctx->scratch_get(SC_FILE_TOK, ((uint64_t)n_barcodes + 1) * 8, (void**)&d_bco));
    GT_TRY(ctx->scratch_get(SC_TILE_STATUS, fused_workspace_bytes(n), &d_ws));
    GT_TRY(ctx->scratch_get(SC_MISC, 64, (void**)&d_misc));

    // 1. hits of every fragment, raw (no unk yet), with per-fragment offsets
    uint64_t cap = n + n / 4 + 1024, total = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        GT_TRY(ctx->scratch_get(SC_OUT_IDS, cap * 4, (void**)&d_raw));
        GT_CUDA(cudaMemsetAsync(d_misc, 0, 64, st));
        GT_TRY(launch_fused_find(ix, n, 0, nullptr, d_chr, d_start, d_end, 0, d_raw, cap, d_off, nullptr, d_ws, nullptr, d_misc,
                                 (uint32_t*)(d_misc + 2)));
        GT_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_misc, 24, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        total = ctx->h_scalars[0];
        if ((uint32_t)ctx->h_scalars[2] != 0) return fail(GTGPU_ERR_UNSUPPORTED, "tokenize_fragments: tile overflow");
        if (total <= cap) break;
        if (attempt == 1) return fail(GTGPU_ERR_CAPACITY, "tokenize_fragments: output capacity exceeded twice");
        cap = total;
    }

    uint64_t final_total = 0;
    if (n) {
        // 2. stable sort of fragment indices by barcode
        const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
        iota_kernel<<<grid, 256, 0, st>>>(n, d_idx);
        ctx->launches++;
        int bits = 1;
        while (bits < 32 && (1ull << bits) < n_barcodes) ++bits;
        void* d_tmp = nullptr;
        GT_TRY(ctx->scratch_get(SC_IN3_START, std::max(radix_sort_temp_bytes(n), exclusive_scan_temp_bytes(n, 8)), &d_tmp));
        int in_b = 0;
        GT_TRY(radix_sort_pairs(ctx, n, d_bc, d_idx, d_bc_sorted, d_order, bits, d_tmp, &in_b));
        if (!in_b) {  // even number of passes: the sorted data sits in the input buffers
            std::swap(d_bc, d_bc_sorted);
            std::swap(d_idx, d_order);
        }
        // 3. tokens per fragment (per-fragment unk rule) and their destinations
        frag_token_counts_kernel<<<grid, 256, 0, st>>>(n, d_order, d_off, d_cnt);
        ctx->launches++;
        GT_TRY(exclusive_scan<unsigned long long>(ctx, d_cnt, d_dst, n, d_tmp));
        // 4. barcode offsets + scatter
        frag_barcode_offsets_kernel<<<(n_barcodes + 1 + 255) / 256, 256, 0, st>>>(n_barcodes, n, d_bc_sorted, d_dst, d_cnt, d_bco);
        ctx->launches++;
        GT_CUDA(cudaMemcpyAsync(out_barcode_offsets, d_bco, ((uint64_t)n_barcodes + 1) * 8, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        final_total = out_barcode_offsets[n_barcodes];
        GT_TRY(ctx->scratch_get(SC_OUT_IDS2, final_total * 4, (void**)&d_out));
        frag_scatter_kernel<<<grid, 256, 0, st>>>(n, d_order, d_off, d_dst, d_raw, unk_id, d_out);
        ctx->launches++;
        GT_CUDA(cudaGetLastError());
    } else {
        for (uint32_t b = 0; b <= n_barcodes; ++b) out_barcode_offsets[b] = 0;
    }

    gtgpu_buf* buf = new gtgpu_buf();
    buf->ctx = ctx;
    buf->len = final_total;
    int32_t s = ctx->pinned_get(final_total * 4, &buf->block);
    if (s != GTGPU_OK) {
        delete buf;
        return s;
    }
    cudaError_t e = cudaSuccess;
    if (final_total) e = cudaMemcpyAsync(buf->block.ptr, d_out, final_total * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        ctx->pinned_put(buf->block);
        delete buf;
        return fail(GTGPU_ERR_CUDA, std::string("tokenize_fragments: D2H: ") + cudaGetErrorString(e));
    }
    *out_ids = buf;
    return GTGPU_OK;
}

}  // namespace gtgpu

static int32_t tokenize_fragments_one(gtgpu_index* ix, uint64_t n, const uint32_t* chr, const uint32_t* start, const uint32_t* end,
                                      const uint32_t* barcode_id, uint32_t n_barcodes, uint32_t unk_id,
                                      uint64_t* out_barcode_offsets, gtgpu_buf** out_ids) {
    gtgpu_ctx* ctx = ix->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint32_t *d_chr, *d_start, *d_end, *d_bc;
    GT_TRY(ctx->scratch_get(SC_CHR, n * 4, (void**)&d_chr));
    GT_TRY(ctx->scratch_get(SC_START, n * 4, (void**)&d_start));
    GT_TRY(ctx->scratch_get(SC_END, n * 4, (void**)&d_end));
    GT_TRY(ctx->scratch_get(SC_BARCODE, n * 4, (void**)&d_bc));
    if (n) {
        GT_CUDA(cudaMemcpyAsync(d_chr, chr, n * 4, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_start, start, n * 4, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_end, end, n * 4, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_bc, barcode_id, n * 4, cudaMemcpyHostToDevice, st));
    }
    return tokenize_fragments_core(ix, n, d_chr, d_start, d_end, d_bc, n_barcodes, unk_id, out_barcode_offsets, out_ids);
}

// Multi-device group: the fragments are dealt to the devices in contiguous blocks, every device groups its block by
// barcode, and each barcode's final list is the concatenation of its per-device lists in device order — file order inside
// a barcode is preserved because the blocks are contiguous ranges of the file (no collective).
static int32_t tokenize_fragments_group(gtgpu_index* ix, uint64_t n, const uint32_t* chr, const uint32_t* start, const uint32_t* end,
                                        const uint32_t* barcode_id, uint32_t n_barcodes, uint32_t unk_id,
                                        uint64_t* out_barcode_offsets, gtgpu_buf** out_ids) {
    std::lock_guard<std::mutex> glk(ix->ctx->group_mu);
    const size_t D = ix->replicas.size();
    std::vector<gtgpu_buf*> parts(D, nullptr);
    std::vector<std::vector<uint64_t>> offs(D, std::vector<uint64_t>((size_t)n_barcodes + 1, 0));
    int32_t s = for_each_device(D, [&](size_t r) -> int32_t {
        uint64_t lo, hi;
        block_range(n, D, r, &lo, &hi);
        return tokenize_fragments_one(ix->replicas[r], hi - lo, chr + lo, start + lo, end + lo, barcode_id + lo, n_barcodes, unk_id,
                                      offs[r].data(), &parts[r]);
    });
    gtgpu_buf* buf = nullptr;
    if (s == GTGPU_OK) {
        out_barcode_offsets[0] = 0;
        for (uint32_t b = 0; b < n_barcodes; ++b) {
            uint64_t len = 0;
            for (size_t r = 0; r < D; ++r) len += offs[r][b + 1] - offs[r][b];
            out_barcode_offsets[b + 1] = out_barcode_offsets[b] + len;
        }
        buf = new gtgpu_buf();
        buf->ctx = ix->ctx;
        buf->len = out_barcode_offsets[n_barcodes];
        {
            std::lock_guard<std::mutex> lk(ix->ctx->mu);
            cudaSetDevice(ix->ctx->device);
            s = ix->ctx->pinned_get(buf->len * 4, &buf->block);
        }
        if (s != GTGPU_OK) {
            delete buf;
            buf = nullptr;
        }
    }
    if (s == GTGPU_OK) {
        uint32_t* dst = (uint32_t*)buf->block.ptr;
        for_each_device(D, [&](size_t t) -> int32_t {  // D host threads, each a contiguous range of barcodes
            uint64_t b0, b1;
            block_range(n_barcodes, D, t, &b0, &b1);
            for (uint64_t b = b0; b < b1; ++b) {
                uint64_t pos = out_barcode_offsets[b];
                for (size_t r = 0; r < D; ++r) {
                    const uint64_t a = offs[r][b], len = offs[r][b + 1] - a;
                    if (len) memcpy(dst + pos, (const uint32_t*)parts[r]->block.ptr + a, len * 4);
                    pos += len;
                }
            }
            return GTGPU_OK;
        });
        *out_ids = buf;
    }
    for (auto* p : parts)
        if (p) gtgpu_buf_free(p);
    return s;
}

extern "C" int32_t gtgpu_tokenize_fragments(gtgpu_index* ix, uint64_t n, const uint32_t* chr, const uint32_t* start,
                                            const uint32_t* end, const uint32_t* barcode_id, uint32_t n_barcodes,
                                            uint32_t unk_id, uint64_t* out_barcode_offsets, gtgpu_buf** out_ids) try {
    if (!ix || !out_barcode_offsets || !out_ids || (n && (!chr || !start || !end || !barcode_id)))
        return fail(GTGPU_ERR_INVALID, "tokenize_fragments: null argument");
    if (n >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "tokenize_fragments: more than 2^32-2 fragments per call");
    for (uint64_t i = 0; i < n; ++i)
        if (barcode_id[i] >= n_barcodes) return fail(GTGPU_ERR_INVALID, "tokenize_fragments: barcode id out of range");
    if (ix->replicas.size() > 1 && n >= (1u << 20))
        return tokenize_fragments_group(ix, n, chr, start, end, barcode_id, n_barcodes, unk_id, out_barcode_offsets, out_ids);
    return tokenize_fragments_one(ix, n, chr, start, end, barcode_id, n_barcodes, unk_id, out_barcode_offsets, out_ids);
} GT_CATCH

// Device-resident form of gtgpu_tokenize_fragments: nothing crosses PCIe and nothing synchronises — find, stable radix
// sort by barcode, per-fragment [unk] rule, scan and scatter are queued on the ctx stream.
extern "C" int32_t gtgpu_tokenize_fragments_dev(gtgpu_index* ix, uint64_t n, const uint32_t* d_chr, const uint32_t* d_start,
                                                const uint32_t* d_end, const uint32_t* d_barcode_id, uint32_t n_barcodes,
                                                uint32_t unk_id, uint64_t* d_out_barcode_offsets, uint32_t* d_out_ids,
                                                uint64_t ids_capacity, uint64_t* d_out_total) try {
    if (!ix || !d_out_barcode_offsets || !d_out_total || (ids_capacity && !d_out_ids) ||
        (n && (!d_chr || !d_start || !d_end || !d_barcode_id)))
        return fail(GTGPU_ERR_INVALID, "tokenize_fragments_dev: null argument");
    if (n >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "tokenize_fragments_dev: more than 2^32-2 fragments per call");
    gtgpu_ctx* ctx = ix->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (n == 0) {
        GT_CUDA(cudaMemsetAsync(d_out_barcode_offsets, 0, ((uint64_t)n_barcodes + 1) * 8, st));
        GT_CUDA(cudaMemsetAsync(d_out_total, 0, 8, st));
        return GTGPU_OK;
    }
    uint32_t *d_bc, *d_bc_sorted, *d_idx, *d_order, *d_raw;
    uint64_t *d_off, *d_misc;
    unsigned long long *d_cnt, *d_dst;
    void *d_ws, *d_tmp;
    const uint64_t raw_cap = n + n / 4 + 1024;
    GT_TRY(ctx->scratch_get(SC_BARCODE, n * 4, (void**)&d_bc));
    GT_TRY(ctx->scratch_get(SC_IN2_CHR, n * 4, (void**)&d_bc_sorted));
    GT_TRY(ctx->scratch_get(SC_IN2_START, n * 4, (void**)&d_idx));
    GT_TRY(ctx->scratch_get(SC_IN2_END, n * 4, (void**)&d_order));
    GT_TRY(ctx->scratch_get(SC_OUT_OFFS, (n + 1) * 8, (void**)&d_off));
    GT_TRY(ctx->scratch_get(SC_COUNTS, (n + 1) * 8, (void**)&d_cnt));
    GT_TRY(ctx->scratch_get(SC_IN3_CHR, (n + 1) * 8, (void**)&d_dst));
    GT_TRY(ctx->scratch_get(SC_OUT_IDS, raw_cap * 4, (void**)&d_raw));
    GT_TRY(ctx->scratch_get(SC_TILE_STATUS, fused_workspace_bytes(n), &d_ws));
    GT_TRY(ctx->scratch_get(SC_MISC, 64, (void**)&d_misc));
    GT_TRY(ctx->scratch_get(SC_IN3_START, std::max(radix_sort_temp_bytes(n), exclusive_scan_temp_bytes(n, 8)), &d_tmp));
    GT_CUDA(cudaMemsetAsync(d_misc, 0, 64, st));
    GT_TRY(launch_fused_find(ix, n, 0, nullptr, d_chr, d_start, d_end, 0, d_raw, raw_cap, d_off, nullptr, d_ws, nullptr, d_misc,
                             (uint32_t*)(d_misc + 2)));
    GT_CUDA(cudaMemcpyAsync(d_bc, d_barcode_id, n * 4, cudaMemcpyDeviceToDevice, st));  // the sort ping-pongs its inputs
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    iota_kernel<<<grid, 256, 0, st>>>(n, d_idx);
    ctx->launches++;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < n_barcodes) ++bits;
    int in_b = 0;
    GT_TRY(radix_sort_pairs(ctx, n, d_bc, d_idx, d_bc_sorted, d_order, bits, d_tmp, &in_b));
    if (!in_b) {
        std::swap(d_bc, d_bc_sorted);
        std::swap(d_idx, d_order);
    }
    frag_token_counts_kernel<<<grid, 256, 0, st>>>(n, d_order, d_off, d_cnt);
    GT_TRY(exclusive_scan<unsigned long long>(ctx, d_cnt, d_dst, n, d_tmp));
    frag_barcode_offsets_kernel<<<(n_barcodes + 1 + 255) / 256, 256, 0, st>>>(n_barcodes, n, d_bc_sorted, d_dst, d_cnt, d_out_barcode_offsets);
    frag_scatter_cap_kernel<<<grid, 256, 0, st>>>(n, d_order, d_off, d_dst, d_cnt, d_raw, unk_id, d_out_ids, ids_capacity, d_misc, raw_cap,
                                                  d_out_total);
    ctx->launches += 3;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
} GT_CATCH
