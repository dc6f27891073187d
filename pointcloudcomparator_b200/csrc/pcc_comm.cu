// pcc_comm.cu -- multi-GPU entry points of the C ABI: one process per GPU, the caller owns an NCCL
// communicator over the ranks; the reference grid is REPLICATED (broadcast from the rank that built it) and the queries are
// sharded, so the only collectives are: the grid broadcast, the optional gather of per-shard result rows, and the 17-double
// all-reduce of ICP's correspondence sums (inside pcc_icp_step when the index has a communicator).
//
// NCCL is not a link-time dependency: its few entry points are resolved with dlopen("libnccl.so.2") at pcc_comm_init, so the
// library the host program already loaded (its own NCCL, or the one a framework brought in) is the one used, the communicator
// handed in comes from that same library, and a single-GPU consumer never needs NCCL installed.
#include <dlfcn.h>
#include <algorithm>
#include <vector>

#include <nccl.h>          // types and enums only; no symbol of libnccl is referenced at link time

#include "pcc_internal.h"

namespace pcc {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.handle) return PCC_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);        // the copy the process already uses, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(PCC_ERR_STATE, "pcc_comm_init: libnccl.so.2 cannot be loaded (%s)", dlerror());
#define PCC_SYM(field, name)                                                                                           \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                                                        \
    if (!g_nccl.field) return fail(PCC_ERR_STATE, "pcc_comm_init: libnccl has no symbol %s", name)
    PCC_SYM(Broadcast, "ncclBroadcast"); PCC_SYM(AllReduce, "ncclAllReduce"); PCC_SYM(AllGather, "ncclAllGather");
    PCC_SYM(GroupStart, "ncclGroupStart"); PCC_SYM(GroupEnd, "ncclGroupEnd"); PCC_SYM(CommCount, "ncclCommCount");
    PCC_SYM(CommUserRank, "ncclCommUserRank"); PCC_SYM(GetErrorString, "ncclGetErrorString");
#undef PCC_SYM
    g_nccl.handle = h;
    return PCC_OK;
}
#define PCC_NCCL(expr)                                                                                                 \
    do {                                                                                                               \
        ncclResult_t r__ = (expr);                                                                                     \
        if (r__ != ncclSuccess) return pcc::fail(PCC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
    } while (0)

// sum-all-reduce of n doubles in place on device memory (ICP's 16 sums + count); no-op without a communicator
int comm_allreduce_f64(pcc_index *idx, double *d_buf, int n, cudaStream_t s) {
    if (!idx->comm || idx->comm_world <= 1) return PCC_OK;
    PCC_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)idx->comm, s));
    return PCC_OK;
}

__global__ void scatter_rows_kernel(const uint32_t *__restrict__ src, const int32_t *__restrict__ rows, int64_t n_rows, int words_per_row, uint32_t *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * words_per_row) return;
    const int64_t r = i / words_per_row; const int w = (int)(i - r * words_per_row);
    dst[(int64_t)rows[r] * words_per_row + w] = src[i];
}

}  // namespace pcc

using namespace pcc;

extern "C" {

int pcc_comm_init(pcc_index *idx, void *nccl_comm, int rank, int world) {
    if (!idx) return fail(PCC_ERR_INVALID, "idx is NULL");
    if (!nccl_comm) { idx->comm = nullptr; idx->comm_rank = 0; idx->comm_world = 1; return PCC_OK; }       // detach
    if (world < 1 || rank < 0 || rank >= world) return fail(PCC_ERR_INVALID, "pcc_comm_init: rank %d / world %d", rank, world);
    PCC_TRY(load_nccl());
    int n = 0, r = 0;
    PCC_NCCL(g_nccl.CommCount((ncclComm_t)nccl_comm, &n));
    PCC_NCCL(g_nccl.CommUserRank((ncclComm_t)nccl_comm, &r));
    if (n != world || r != rank) return fail(PCC_ERR_INVALID, "pcc_comm_init: communicator says rank %d of %d, caller says %d of %d", r, n, rank, world);
    idx->comm = nccl_comm; idx->comm_rank = rank; idx->comm_world = world;
    return PCC_OK;
}

int pcc_comm_info(const pcc_index *idx, int *rank, int *world) {
    if (!idx) return fail(PCC_ERR_INVALID, "idx is NULL");
    if (rank) *rank = idx->comm ? idx->comm_rank : 0;
    if (world) *world = idx->comm ? idx->comm_world : 1;
    return PCC_OK;
}

int pcc_broadcast_index(pcc_index *idx, int root, void *stream) {
    if (!idx) return fail(PCC_ERR_INVALID, "idx is NULL");
    if (!idx->comm) return fail(PCC_ERR_STATE, "pcc_broadcast_index: no communicator (call pcc_comm_init first)");
    if (root < 0 || root >= idx->comm_world) return fail(PCC_ERR_INVALID, "root %d out of range", root);
    PCC_CUDA(cudaSetDevice(idx->device));
    cudaStream_t s = (cudaStream_t)stream;
    ncclComm_t comm = (ncclComm_t)idx->comm;
    // 1. the grid description (16 doubles) through a device staging buffer
    double meta[16]; void *ptrs[2] = {nullptr, nullptr};
    for (int i = 0; i < 16; ++i) meta[i] = 0;
    meta[15] = -1;                                           // marks "root had no built index"
    if (idx->comm_rank == root) { if (!idx->built) return fail(PCC_ERR_STATE, "pcc_broadcast_index: root index not built"); PCC_TRY(pcc_export(idx, meta, ptrs)); meta[15] = 1; }
    PCC_TRY(idx->sel_params.reserve(16 * sizeof(double)));
    double *d_meta = idx->sel_params.as<double>();
    PCC_CUDA(cudaMemcpyAsync(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    PCC_NCCL(g_nccl.Broadcast(d_meta, d_meta, 16, ncclFloat64, root, comm, s));
    PCC_CUDA(cudaMemcpyAsync(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost, s));
    PCC_CUDA(cudaStreamSynchronize(s));
    if (meta[15] != 1) return fail(PCC_ERR_STATE, "pcc_broadcast_index: the root rank had no built index");
    // 2. adopt (allocates the two arrays), then the arrays themselves straight into place
    if (idx->comm_rank != root) { PCC_TRY(pcc_adopt(idx, meta, stream)); PCC_TRY(pcc_export(idx, meta, ptrs)); }
    const size_t n_pts = (size_t)meta[0], n_cells = (size_t)meta[11];
    if (n_pts) PCC_NCCL(g_nccl.Broadcast(ptrs[0], ptrs[0], n_pts * sizeof(float4), ncclUint8, root, comm, s));
    PCC_NCCL(g_nccl.Broadcast(ptrs[1], ptrs[1], (n_cells + 1) * sizeof(uint32_t), ncclUint8, root, comm, s));
    PCC_CUDA(cudaStreamSynchronize(s));
    return PCC_OK;
}

int pcc_gather(pcc_index *idx, const void *local, int64_t n_local, int row_bytes, const int32_t *local_rows, void *out, int64_t n_total, void *stream) {
    if (!idx) return fail(PCC_ERR_INVALID, "idx is NULL");
    if (!idx->comm) return fail(PCC_ERR_STATE, "pcc_gather: no communicator (call pcc_comm_init first)");
    if (n_local < 0 || row_bytes <= 0 || (row_bytes & 3) || n_total < 0 || (n_total > 0 && !out) || (n_local > 0 && !local))
        return fail(PCC_ERR_INVALID, "pcc_gather: bad arguments (n_local=%lld row_bytes=%d n_total=%lld)", (long long)n_local, row_bytes, (long long)n_total);
    PCC_CUDA(cudaSetDevice(idx->device));
    cudaStream_t s = (cudaStream_t)stream;
    ncclComm_t comm = (ncclComm_t)idx->comm;
    const int world = idx->comm_world, rank = idx->comm_rank;
    // 1. every rank's row count
    PCC_TRY(idx->sel_params.reserve((size_t)(world + 1) * sizeof(int64_t) + 128));
    int64_t *d_counts = idx->sel_params.as<int64_t>();
    PCC_CUDA(cudaMemcpyAsync(d_counts + world, &n_local, sizeof(int64_t), cudaMemcpyHostToDevice, s));
    PCC_NCCL(g_nccl.AllGather(d_counts + world, d_counts, 1, ncclInt64, comm, s));
    std::vector<int64_t> counts((size_t)world), offs((size_t)world + 1, 0);
    PCC_CUDA(cudaMemcpyAsync(counts.data(), d_counts, (size_t)world * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    PCC_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < world; ++r) offs[(size_t)r + 1] = offs[(size_t)r] + counts[(size_t)r];
    if (offs[(size_t)world] != n_total) return fail(PCC_ERR_INVALID, "pcc_gather: the ranks hold %lld rows in total, n_total says %lld", (long long)offs[(size_t)world], (long long)n_total);
    // 2. one broadcast per rank inside a group: rank r's block lands at offset offs[r] (rank-order concatenation)
    uint8_t *stage = (uint8_t *)out; int32_t *stage_rows = nullptr;
    if (local_rows) {
        PCC_TRY(idx->keys64.reserve((size_t)n_total * row_bytes + 64));
        PCC_TRY(idx->keys64b.reserve((size_t)n_total * sizeof(int32_t) + 64));
        stage = idx->keys64.as<uint8_t>(); stage_rows = idx->keys64b.as<int32_t>();
    }
    PCC_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < world; ++r) {
        const size_t bytes = (size_t)counts[(size_t)r] * row_bytes;
        if (!bytes) continue;
        PCC_NCCL(g_nccl.Broadcast(r == rank ? local : nullptr, stage + (size_t)offs[(size_t)r] * row_bytes, bytes, ncclUint8, r, comm, s));
        if (local_rows) PCC_NCCL(g_nccl.Broadcast(r == rank ? (const void *)local_rows : nullptr, stage_rows + offs[(size_t)r], (size_t)counts[(size_t)r] * sizeof(int32_t), ncclUint8, r, comm, s));
    }
    PCC_NCCL(g_nccl.GroupEnd());
    // 3. back to original row order
    if (local_rows && n_total > 0) {
        const int wpr = row_bytes / 4;
        const int64_t words = n_total * wpr;
        scatter_rows_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>((const uint32_t *)stage, stage_rows, n_total, wpr, (uint32_t *)out);
        PCC_LAUNCHED();
        PCC_CUDA(cudaGetLastError());
    }
    return PCC_OK;
}

int pcc_allreduce_f64(pcc_index *idx, double *device_buf, int n, void *stream) {
    if (!idx) return fail(PCC_ERR_INVALID, "idx is NULL");
    if (!idx->comm) return fail(PCC_ERR_STATE, "pcc_allreduce_f64: no communicator (call pcc_comm_init first)");
    PCC_CUDA(cudaSetDevice(idx->device));
    return comm_allreduce_f64(idx, device_buf, n, (cudaStream_t)stream);
}

}  // extern "C"
