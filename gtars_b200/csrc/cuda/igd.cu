// igd.cu — gtars-igd / gtars-lola overlap-count matrices on the device.
//
// The reference (gtars-igd/src/igd.rs) bins every database interval into 16 384-bp tiles and walks, per query, every
// record of the touched tiles (igd.rs:753-847).  For min_overlap >= 1 the result is tile-independent:
//     hit(q, r, m)  <=>  same contig  &&  r.start < q.end  &&  min(r.end, q.end) - max(r.start, max(q.start, 0)) >= m
// so the device keeps, per contig, ONE pooled start-sorted SoA of all database records (start, end, file) plus
//   pmax[i]  = max end over records [0, i]            -> candidate range [first i with pmax > q.start, #starts < q.end)
//   psame[i] = max end over EARLIER records of the same file with length >= m (m = 1 precomputed)
// One warp resolves one query: the lanes stride over the candidate range, test the closed form and add into the
// [n_sets x n_files] matrix with 64-bit atomics (the 80 MB matrix of the LOLA configuration is L2-resident).
//   count_set_overlaps (igd.rs:544-556): every hit counts.
//   count_region_hits  (igd.rs:563-590): a (query, file) pair counts once.  A hit r is the FIRST hit of its file for
//       this query iff no earlier same-file record hits, i.e. iff psame[r] - q.start < m (derivation in DESIGN.md);
//       that replaces the reference's O(n_files) scratch zero + scan per region.
// min_overlap <= 0 is tile-dependent in the reference (touching records count only when co-tiled) and is rejected.
#include <algorithm>
#include <map>
#include <numeric>

#include "common.cuh"

struct gtgpu_igd {
    gtgpu_ctx* ctx = nullptr;
    uint64_t n_files = 0, n_records = 0;
    uint32_t n_chroms = 0, shift = 0;
    // per chromosome (n_chroms + 1 offsets into the record arrays; LUT offsets / bin counts)
    std::vector<uint32_t> h_off;
    uint32_t *d_off = nullptr, *d_lut_s_off = nullptr, *d_nb_s = nullptr, *d_lut_p_off = nullptr, *d_nb_p = nullptr;
    int32_t *d_start = nullptr, *d_end = nullptr, *d_pmax = nullptr, *d_psame1 = nullptr;
    uint32_t *d_file = nullptr, *d_lut = nullptr;
    // host copies kept to derive psame for other min_overlap values on demand
    std::vector<int32_t> h_start, h_end;
    std::vector<uint32_t> h_file;
    std::map<int32_t, int32_t*> psame_by_m;
    std::vector<void*> allocs;
    uint64_t device_bytes = 0;
};

namespace gtgpu {

namespace {

template <class T>
int32_t up(gtgpu_igd* g, const std::vector<T>& v, T** out) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return fail(GTGPU_ERR_NOMEM, std::string("cudaMalloc(igd): ") + cudaGetErrorString(e));
    g->allocs.push_back(d);
    g->device_bytes += bytes;
    if (!v.empty()) GT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (T*)d;
    return GTGPU_OK;
}

void build_lut_i32(const int32_t* arr, uint32_t n, uint32_t shift, std::vector<uint32_t>& lut, uint32_t& off, uint32_t& nb) {
    off = (uint32_t)lut.size();
    if (n == 0) {
        nb = 0;
        lut.push_back(0);
        return;
    }
    nb = ((uint32_t)arr[n - 1] >> shift) + 1;
    lut.resize(lut.size() + (size_t)nb + 1);
    uint32_t* L = lut.data() + off;
    uint32_t i = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        int64_t key = (int64_t)b << shift;
        while (i < n && arr[i] < key) ++i;
        L[b] = i;
    }
    L[nb] = n;
}

// psame[i] = max end over earlier same-file records (pooled order) with end - start >= m; -1 when there is none.
std::vector<int32_t> compute_psame(const gtgpu_igd* g, int32_t m) {
    std::vector<int32_t> ps(g->n_records);
    std::vector<int32_t> last(g->n_files);
    for (uint32_t c = 0; c < g->n_chroms; ++c) {
        std::fill(last.begin(), last.end(), -1);
        for (uint32_t i = g->h_off[c]; i < g->h_off[c + 1]; ++i) {
            uint32_t f = g->h_file[i];
            ps[i] = last[f];
            if ((int64_t)g->h_end[i] - g->h_start[i] >= m) last[f] = std::max(last[f], g->h_end[i]);
        }
    }
    return ps;
}

}  // namespace

__device__ __forceinline__ uint32_t lb_i32(const int32_t* __restrict__ arr, const uint32_t* __restrict__ lut, uint32_t nb,
                                           uint32_t n, uint32_t shift, int32_t key) {
    if (key <= 0) return 0;
    uint32_t b = (uint32_t)key >> shift;
    if (b >= nb) return n;
    uint32_t lo = __ldg(lut + b), hi = __ldg(lut + b + 1);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(arr + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

struct IgdView {
    const uint32_t *off, *lut_s_off, *nb_s, *lut_p_off, *nb_p, *file, *lut;
    const int32_t *start, *end, *pmax, *psame;
    uint32_t n_chroms, shift;
    uint64_t n_files;
};

// One warp per query.  BINARY selects count_region_hits semantics.
template <bool BINARY>
__global__ void __launch_bounds__(256) igd_count_kernel(IgdView v, uint64_t n, const uint32_t* __restrict__ set_of,
                                                         const uint32_t* __restrict__ chr, const uint32_t* __restrict__ qstart,
                                                         const uint32_t* __restrict__ qend, int32_t m,
                                                         unsigned long long* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n; q += warps) {
        const uint32_t c = __ldg(chr + q);
        int32_t s = (int32_t)__ldg(qstart + q), e = (int32_t)__ldg(qend + q);
        if (c >= v.n_chroms || s >= e || e <= 0) continue;  // igd.rs:514-522
        s = max(s, 0);
        const uint32_t o = __ldg(v.off + c), len = __ldg(v.off + c + 1) - o;
        if (len == 0) continue;
        // records that can reach the query: start < e, and some record at or before them ends after s
        const uint32_t ub = lb_i32(v.start + o, v.lut + __ldg(v.lut_s_off + c), __ldg(v.nb_s + c), len, v.shift, e);
        const uint32_t lo = lb_i32(v.pmax + o, v.lut + __ldg(v.lut_p_off + c), __ldg(v.nb_p + c), len, v.shift, s + 1);
        unsigned long long* row = out + (uint64_t)__ldg(set_of + q) * v.n_files;
        for (uint32_t i = o + lo + lane; i < o + ub; i += 32) {
            const int32_t rs = __ldg(v.start + i), re = __ldg(v.end + i);
            if (min(re, e) - max(rs, s) < m) continue;
            if (BINARY && (int64_t)__ldg(v.psame + i) - s >= m) continue;  // an earlier record of this file already hit
            atomicAdd(row + __ldg(v.file + i), 1ull);
        }
    }
}

// set index of every query from the set offsets (one thread per set writes its range)
__global__ void fill_set_ids_kernel(uint64_t n_sets, const uint64_t* __restrict__ set_offsets, uint32_t* __restrict__ set_of) {
    for (uint64_t s = blockIdx.x; s < n_sets; s += gridDim.x)
        for (uint64_t i = set_offsets[s] + threadIdx.x; i < set_offsets[s + 1]; i += blockDim.x) set_of[i] = (uint32_t)s;
}

int32_t launch_fill_set_ids(gtgpu_ctx* ctx, uint64_t n_sets, const uint64_t* d_set_offsets, uint32_t* d_set_of) {
    if (!n_sets) return GTGPU_OK;
    fill_set_ids_kernel<<<(unsigned)std::min<uint64_t>(n_sets, 4096), 256, 0, ctx->stream>>>(n_sets, d_set_offsets, d_set_of);
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

}  // namespace gtgpu

using namespace gtgpu;

extern "C" int32_t gtgpu_igd_build(gtgpu_ctx* ctx, uint64_t n_files, const uint64_t* file_offsets, uint32_t n_chroms,
                                   const uint32_t* chr, const uint32_t* start, const uint32_t* end, gtgpu_igd** out_igd) {
    if (!ctx || !out_igd || !file_offsets) return fail(GTGPU_ERR_INVALID, "igd_build: null argument");
    uint64_t total = file_offsets[n_files];
    if (total && (!chr || !start || !end)) return fail(GTGPU_ERR_INVALID, "igd_build: null record arrays");
    if (total >= 0xFFFFFFFFull || n_files >= 0xFFFFFFFFull) return fail(GTGPU_ERR_UNSUPPORTED, "igd_build: too many records");
    GT_CUDA(cudaSetDevice(ctx->device));
    gtgpu_igd* g = new gtgpu_igd();
    g->ctx = ctx;
    g->n_files = n_files;
    g->n_chroms = n_chroms;

    // keep what Igd::add keeps (igd.rs:109-116, 285-301): start < end as u32, then 0 <= start < end as i32
    std::vector<uint32_t> keep;
    keep.reserve(total);
    std::vector<uint32_t> file_of(total);
    for (uint64_t f = 0; f < n_files; ++f)
        for (uint64_t i = file_offsets[f]; i < file_offsets[f + 1]; ++i) file_of[i] = (uint32_t)f;
    for (uint64_t i = 0; i < total; ++i) {
        if (chr[i] >= n_chroms || !(start[i] < end[i])) continue;
        int32_t s = (int32_t)start[i], e = (int32_t)end[i];
        if (s < 0 || e < 0 || s >= e) continue;
        keep.push_back((uint32_t)i);
    }
    // pooled order: by chromosome, then start; ties keep (file, insertion) order = the order Igd::add saw them
    std::stable_sort(keep.begin(), keep.end(), [&](uint32_t a, uint32_t b) {
        return chr[a] != chr[b] ? chr[a] < chr[b] : start[a] < start[b];
    });
    g->n_records = keep.size();
    g->h_start.resize(keep.size());
    g->h_end.resize(keep.size());
    g->h_file.resize(keep.size());
    g->h_off.assign(n_chroms + 1, 0);
    std::vector<int32_t> pmax(keep.size());
    for (size_t k = 0; k < keep.size(); ++k) {
        uint32_t i = keep[k];
        g->h_start[k] = (int32_t)start[i];
        g->h_end[k] = (int32_t)end[i];
        g->h_file[k] = file_of[i];
        g->h_off[chr[i] + 1]++;
    }
    for (uint32_t c = 0; c < n_chroms; ++c) g->h_off[c + 1] += g->h_off[c];
    int32_t max_coord = 0;
    for (uint32_t c = 0; c < n_chroms; ++c) {
        int32_t mx = 0;
        for (uint32_t k = g->h_off[c]; k < g->h_off[c + 1]; ++k) {
            mx = std::max(mx, g->h_end[k]);
            pmax[k] = mx;
            max_coord = std::max(max_coord, mx);
        }
    }
    // LUT shift: at most ~8 M bins per family
    uint32_t shift = 0;
    auto bins = [&](uint32_t sh) {
        uint64_t t = 0;
        for (uint32_t c = 0; c < n_chroms; ++c)
            if (g->h_off[c + 1] > g->h_off[c]) t += ((uint32_t)pmax[g->h_off[c + 1] - 1] >> sh) + 2;
        return t;
    };
    uint64_t budget = std::min<uint64_t>(std::max<uint64_t>(keep.size(), 4096), 8ull << 20);
    while (shift < 31 && bins(shift) > budget) ++shift;
    g->shift = shift;
    std::vector<uint32_t> lut, lso(n_chroms), nbs(n_chroms), lpo(n_chroms), nbp(n_chroms);
    for (uint32_t c = 0; c < n_chroms; ++c) {
        uint32_t o = g->h_off[c], len = g->h_off[c + 1] - o;
        build_lut_i32(g->h_start.data() + o, len, shift, lut, lso[c], nbs[c]);
        build_lut_i32(pmax.data() + o, len, shift, lut, lpo[c], nbp[c]);
    }
    std::vector<int32_t> psame1 = compute_psame(g, 1);
    int32_t st = GTGPU_OK;
    auto U = [&](auto& vec, auto** dst) { if (st == GTGPU_OK) st = up(g, vec, dst); };
    U(g->h_off, &g->d_off);
    U(lso, &g->d_lut_s_off);
    U(nbs, &g->d_nb_s);
    U(lpo, &g->d_lut_p_off);
    U(nbp, &g->d_nb_p);
    U(g->h_start, &g->d_start);
    U(g->h_end, &g->d_end);
    U(pmax, &g->d_pmax);
    U(psame1, &g->d_psame1);
    U(g->h_file, &g->d_file);
    U(lut, &g->d_lut);
    if (st != GTGPU_OK) {
        gtgpu_igd_free(g);
        return st;
    }
    g->psame_by_m[1] = g->d_psame1;
    *out_igd = g;
    return GTGPU_OK;
}

extern "C" int32_t gtgpu_igd_free(gtgpu_igd* g) {
    if (!g) return GTGPU_OK;
    cudaSetDevice(g->ctx->device);
    for (void* p : g->allocs) cudaFree(p);
    delete g;
    return GTGPU_OK;
}

extern "C" int32_t gtgpu_igd_info(const gtgpu_igd* g, uint64_t info[4]) {
    if (!g || !info) return fail(GTGPU_ERR_INVALID, "igd_info: null argument");
    info[0] = g->n_files;
    info[1] = g->n_records;
    info[2] = g->device_bytes;
    info[3] = g->shift;
    return GTGPU_OK;
}

namespace {

int32_t igd_count_dev_impl(gtgpu_igd* g, bool binary, uint64_t n, const uint32_t* d_set_of, const uint32_t* d_chr,
                           const uint32_t* d_start, const uint32_t* d_end, int32_t m, uint64_t* d_out) {
    gtgpu_ctx* ctx = g->ctx;
    if (m < 1)
        return fail(GTGPU_ERR_UNSUPPORTED,
                    "igd count: min_overlap < 1 depends on the reference's tile layout (igd.rs:786-795) and is not supported");
    const int32_t* d_psame = nullptr;
    if (binary) {
        auto it = g->psame_by_m.find(m);
        if (it == g->psame_by_m.end()) {
            std::vector<int32_t> ps = compute_psame(g, m);
            int32_t* d = nullptr;
            GT_TRY(up(g, ps, &d));
            it = g->psame_by_m.emplace(m, d).first;
        }
        d_psame = it->second;
    }
    if (n == 0) return GTGPU_OK;
    IgdView v{g->d_off, g->d_lut_s_off, g->d_nb_s, g->d_lut_p_off, g->d_nb_p, g->d_file, g->d_lut,
              g->d_start, g->d_end, g->d_pmax, d_psame, g->n_chroms, g->shift, g->n_files};
    uint64_t warps_needed = n;
    int grid = (int)std::min<uint64_t>((warps_needed + 7) / 8, (uint64_t)ctx->sm_count * 8);
    ctx->time_begin();
    if (binary)
        igd_count_kernel<true><<<grid, 256, 0, ctx->stream>>>(v, n, d_set_of, d_chr, d_start, d_end, m, (unsigned long long*)d_out);
    else
        igd_count_kernel<false><<<grid, 256, 0, ctx->stream>>>(v, n, d_set_of, d_chr, d_start, d_end, m, (unsigned long long*)d_out);
    ctx->time_end();
    ctx->launches++;
    GT_CUDA(cudaGetLastError());
    return GTGPU_OK;
}

int32_t igd_count_host(gtgpu_igd* g, bool binary, uint64_t n_sets, const uint64_t* set_offsets, const uint32_t* chr,
                       const uint32_t* start, const uint32_t* end, int32_t m, uint64_t* out) {
    if (!g || !set_offsets || !out) return fail(GTGPU_ERR_INVALID, "igd count: null argument");
    uint64_t n = set_offsets[n_sets];
    if (n && (!chr || !start || !end)) return fail(GTGPU_ERR_INVALID, "igd count: null query arrays");
    for (uint64_t s = 0; s < n_sets; ++s)
        if (set_offsets[s] > set_offsets[s + 1]) return fail(GTGPU_ERR_INVALID, "igd count: set_offsets not monotone");
    gtgpu_ctx* ctx = g->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint32_t *d_chr, *d_start, *d_end, *d_set;
    uint64_t *d_so, *d_out;
    GT_TRY(ctx->scratch_get(SC_CHR, n * 4, (void**)&d_chr));
    GT_TRY(ctx->scratch_get(SC_START, n * 4, (void**)&d_start));
    GT_TRY(ctx->scratch_get(SC_END, n * 4, (void**)&d_end));
    GT_TRY(ctx->scratch_get(SC_SET_ID, n * 4, (void**)&d_set));
    GT_TRY(ctx->scratch_get(SC_FILE_OFFS, (n_sets + 1) * 8, (void**)&d_so));
    const uint64_t cells = n_sets * g->n_files;
    GT_TRY(ctx->scratch_get(SC_MATRIX, cells * 8, (void**)&d_out));
    if (n) {
        GT_CUDA(cudaMemcpyAsync(d_chr, chr, n * 4, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_start, start, n * 4, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(d_end, end, n * 4, cudaMemcpyHostToDevice, st));
    }
    GT_CUDA(cudaMemcpyAsync(d_so, set_offsets, (n_sets + 1) * 8, cudaMemcpyHostToDevice, st));
    GT_CUDA(cudaMemsetAsync(d_out, 0, std::max<uint64_t>(cells * 8, 8), st));
    if (n_sets && n) GT_TRY(launch_fill_set_ids(ctx, n_sets, d_so, d_set));
    GT_TRY(igd_count_dev_impl(g, binary, n, d_set, d_chr, d_start, d_end, m, d_out));
    if (cells) GT_CUDA(cudaMemcpyAsync(out, d_out, cells * 8, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));
    return GTGPU_OK;
}

}  // namespace

extern "C" int32_t gtgpu_igd_count_set_overlaps(gtgpu_igd* igd, uint64_t n_sets, const uint64_t* set_offsets,
                                                const uint32_t* chr, const uint32_t* start, const uint32_t* end,
                                                int32_t min_overlap, uint64_t* out) {
    return igd_count_host(igd, false, n_sets, set_offsets, chr, start, end, min_overlap, out);
}

extern "C" int32_t gtgpu_igd_count_region_hits(gtgpu_igd* igd, uint64_t n_sets, const uint64_t* set_offsets,
                                               const uint32_t* chr, const uint32_t* start, const uint32_t* end,
                                               int32_t min_overlap, uint64_t* out) {
    return igd_count_host(igd, true, n_sets, set_offsets, chr, start, end, min_overlap, out);
}

extern "C" int32_t gtgpu_igd_count_dev(gtgpu_igd* igd, int32_t binary, uint64_t n, const uint32_t* d_set_of,
                                       const uint32_t* d_chr, const uint32_t* d_start, const uint32_t* d_end,
                                       int32_t min_overlap, uint64_t* d_out) {
    if (!igd || (n && (!d_set_of || !d_chr || !d_start || !d_end || !d_out)))
        return fail(GTGPU_ERR_INVALID, "igd_count_dev: null argument");
    std::lock_guard<std::mutex> lk(igd->ctx->mu);
    GT_CUDA(cudaSetDevice(igd->ctx->device));
    return igd_count_dev_impl(igd, binary != 0, n, d_set_of, d_chr, d_start, d_end, min_overlap, d_out);
}
