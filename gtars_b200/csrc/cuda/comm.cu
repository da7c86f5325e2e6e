// comm.cu — the one collective on this path: LOLA overlap-count matrices with the database sharded by region set.
//
// Every rank (one process per GPU) holds a gtgpu_igd over its own slice of the database's region sets, computes its
// column block of the [n_sets x n_files] matrix, and the blocks are combined with ONE ncclAllGather over
// NVLink / NVSwitch, enqueued on the compute stream right behind the count kernel (no host synchronisation in
// between); a small kernel then transposes the gathered blocks into the row-major matrix the reference returns.
// Tokenize / count / find shard by query and need no collective at all.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that single-GPU users never load it and so that a process
// which already carries torch's bundled NCCL reuses that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"

struct gtgpu_igd;
extern "C" int32_t gtgpu_igd_info(const gtgpu_igd* g, uint64_t info[4]);
extern "C" int32_t gtgpu_igd_count_dev(gtgpu_igd* igd, int32_t binary, uint64_t n, const uint32_t* d_set_of,
                                       const uint32_t* d_chr, const uint32_t* d_start, const uint32_t* d_end,
                                       int32_t min_overlap, uint64_t* d_out);

namespace gtgpu {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) api.lib = nullptr;
        }
    }
    return api.lib ? &api : nullptr;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
};

// gathered[r][s][j] (j < cols_per_rank) -> out[s][r * cols_per_rank + j] for columns that exist
__global__ void untranspose_blocks_kernel(uint32_t world, uint64_t n_sets, uint64_t cols_per_rank, uint64_t n_files,
                                          const uint64_t* __restrict__ gathered, uint64_t* __restrict__ out) {
    const uint64_t total = n_sets * n_files;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        const uint64_t s = k / n_files, f = k % n_files;
        const uint64_t r = f / cols_per_rank, j = f % cols_per_rank;
        out[k] = gathered[(r * n_sets + s) * cols_per_rank + j];
    }
}

}  // namespace gtgpu

using namespace gtgpu;

#define GT_NCCL(api, expr)                                                                                           \
    do {                                                                                                             \
        ncclResult_t _r = (expr);                                                                                    \
        if (_r != ncclSuccess)                                                                                       \
            return fail(GTGPU_ERR_NCCL, std::string(#expr) + ": " + ((api)->GetErrorString ? (api)->GetErrorString(_r) : "?")); \
    } while (0)

extern "C" int32_t gtgpu_comm_unique_id(uint8_t out_id[128]) try {
    if (!out_id) return fail(GTGPU_ERR_INVALID, "comm_unique_id: null argument");
    NcclApi* api = nccl_api();
    if (!api) return fail(GTGPU_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    GT_NCCL(api, api->GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out_id, &id, 128);
    return GTGPU_OK;
} GT_CATCH

extern "C" int32_t gtgpu_comm_init(gtgpu_ctx* ctx, int32_t world, int32_t rank, const uint8_t id_bytes[128]) try {
    if (!ctx || !id_bytes || world < 1 || rank < 0 || rank >= world) return fail(GTGPU_ERR_INVALID, "comm_init: bad argument");
    NcclApi* api = nccl_api();
    if (!api) return fail(GTGPU_ERR_NCCL, "libnccl.so.2 could not be loaded");
    std::lock_guard<std::mutex> lk(ctx->mu);
    GT_CUDA(cudaSetDevice(ctx->device));
    if (ctx->comm) return fail(GTGPU_ERR_INVALID, "comm_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    Comm* c = new Comm();
    c->world = world;
    c->rank = rank;
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(GTGPU_ERR_NCCL, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "?"));
    }
    ctx->comm = c;
    return GTGPU_OK;
} GT_CATCH

extern "C" int32_t gtgpu_comm_free(gtgpu_ctx* ctx) try {
    if (!ctx || !ctx->comm) return GTGPU_OK;
    NcclApi* api = nccl_api();
    Comm* c = (Comm*)ctx->comm;
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
    ctx->comm = nullptr;
    return GTGPU_OK;
} GT_CATCH

// Database sharded by region set: rank r owns global files [r * cols, min((r+1) * cols, n_files_global)),
// cols = ceil(n_files_global / world); `igd` was built over exactly those files (possibly none).
extern "C" int32_t gtgpu_igd_count_sharded(gtgpu_ctx* ctx, gtgpu_igd* igd, int32_t binary, uint64_t n_files_global,
                                           uint64_t n_sets, const uint64_t* set_offsets, const uint32_t* chr,
                                           const uint32_t* start, const uint32_t* end, int32_t min_overlap, uint64_t* out) try {
    if (!ctx || !igd || !set_offsets || !out) return fail(GTGPU_ERR_INVALID, "igd_count_sharded: null argument");
    Comm* c = (Comm*)ctx->comm;
    const int world = c ? c->world : 1, rank = c ? c->rank : 0;
    const uint64_t cols = (n_files_global + world - 1) / world;
    uint64_t info[4];
    GT_TRY(gtgpu_igd_info(igd, info));
    const uint64_t lo = std::min<uint64_t>((uint64_t)rank * cols, n_files_global);
    const uint64_t hi = std::min<uint64_t>(lo + cols, n_files_global);
    if (info[0] != hi - lo) return fail(GTGPU_ERR_INVALID, "igd_count_sharded: this rank's igd does not hold its slice of the sets");
    const uint64_t n = set_offsets[n_sets];
    NcclApi* api = world > 1 ? nccl_api() : nullptr;
    if (world > 1 && !api) return fail(GTGPU_ERR_NCCL, "libnccl.so.2 could not be loaded");

    uint32_t *d_chr, *d_start, *d_end, *d_set;
    uint64_t *d_so, *d_local, *d_gather, *d_full;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        GT_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        GT_TRY(ctx->scratch_get(SC_CHR, n * 4, (void**)&d_chr));
        GT_TRY(ctx->scratch_get(SC_START, n * 4, (void**)&d_start));
        GT_TRY(ctx->scratch_get(SC_END, n * 4, (void**)&d_end));
        GT_TRY(ctx->scratch_get(SC_SET_ID, n * 4, (void**)&d_set));
        GT_TRY(ctx->scratch_get(SC_FILE_OFFS, (n_sets + 1) * 8, (void**)&d_so));
        GT_TRY(ctx->scratch_get(SC_MATRIX, n_sets * cols * 8 + 8, (void**)&d_local));
        GT_TRY(ctx->scratch_get(SC_IN3_CHR, (uint64_t)world * n_sets * cols * 8 + 8, (void**)&d_gather));
        GT_TRY(ctx->scratch_get(SC_IN3_START, n_sets * n_files_global * 8 + 8, (void**)&d_full));
        if (n) {
            GT_CUDA(cudaMemcpyAsync(d_chr, chr, n * 4, cudaMemcpyHostToDevice, st));
            GT_CUDA(cudaMemcpyAsync(d_start, start, n * 4, cudaMemcpyHostToDevice, st));
            GT_CUDA(cudaMemcpyAsync(d_end, end, n * 4, cudaMemcpyHostToDevice, st));
        }
        GT_CUDA(cudaMemcpyAsync(d_so, set_offsets, (n_sets + 1) * 8, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemsetAsync(d_local, 0, n_sets * cols * 8 + 8, st));
    }
    // The local block is computed with row stride = the igd's own file count; make that equal to `cols` by padding:
    // ranks whose slice is shorter than `cols` (the last one) count into a [n_sets x local] matrix first.
    const uint64_t local = hi - lo;
    uint64_t* d_cnt = d_local;
    uint64_t* d_tmp = nullptr;
    if (local != cols) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        GT_TRY(ctx->scratch_get(SC_IN3_END, n_sets * local * 8 + 8, (void**)&d_tmp));
        GT_CUDA(cudaMemsetAsync(d_tmp, 0, n_sets * local * 8 + 8, ctx->stream));
        d_cnt = d_tmp;
    }
    if (n && n_sets) {
        GT_TRY(launch_fill_set_ids(ctx, n_sets, d_so, d_set));
        if (local) GT_TRY(gtgpu_igd_count_dev(igd, binary, n, d_set, d_chr, d_start, d_end, min_overlap, d_cnt));
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    if (local != cols && local)
        GT_CUDA(cudaMemcpy2DAsync(d_local, cols * 8, d_tmp, local * 8, local * 8, n_sets, cudaMemcpyDeviceToDevice, st));
    const uint64_t* d_src = d_local;
    if (world > 1) {
        GT_NCCL(api, api->AllGather(d_local, d_gather, n_sets * cols, ncclUint64, c->comm, st));
        d_src = d_gather;
    }
    const uint64_t cells = n_sets * n_files_global;
    if (cells) {
        int grid = (int)std::min<uint64_t>((cells + 255) / 256, (uint64_t)ctx->sm_count * 16);
        untranspose_blocks_kernel<<<grid, 256, 0, st>>>((uint32_t)world, n_sets, cols, n_files_global, d_src, d_full);
        ctx->launches++;
        GT_CUDA(cudaMemcpyAsync(out, d_full, cells * 8, cudaMemcpyDeviceToHost, st));
    }
    GT_CUDA(cudaStreamSynchronize(st));
    return GTGPU_OK;
} GT_CATCH
