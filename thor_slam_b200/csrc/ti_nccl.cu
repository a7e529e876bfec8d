// Multi-GPU exchange step of the ingest path: ONE gather of per-GPU body-frame clouds to the
// fusing rank over NVLink 5 / NVSwitch (one process per GPU).
//
//  * ti_gather_clouds: grouped ncclSend / ncclRecv on the ctx stream (NCCL is dlopen()ed so the
//    library has no link-time dependency and shares torch's libnccl when torch is loaded).
//  * ti_peer_*: CUDA-IPC peer buffers, so a back-projection kernel on rank r can store its xyz
//    output straight into the root GPU's memory through NVLink - the gather is then fused into
//    the producing kernel (its `dst` simply is a peer pointer) and only a barrier remains.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "ti_common.cuh"

namespace {

struct NcclId { char internal[128]; };
typedef void* ncclComm_t;
typedef int ncclResult_t;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(NcclId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.why = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
        return api;
    }
#define LOAD(field, sym)                                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
    if (!api.field) { api.why = std::string("missing NCCL symbol ") + sym; api.handle = nullptr; return api; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return api;
}

constexpr int NCCL_UINT8 = 1;   // ncclUint8
constexpr int NCCL_INT32 = 2;   // ncclInt32
constexpr int NCCL_UINT32 = 3;  // ncclUint32
constexpr int NCCL_SUM = 0;     // ncclSum

#define TI_NCCL(ctx, expr)                                                                     \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != 0)                                                                           \
            return ti::fail((ctx), TI_ENCCL, "%s failed: %s", #expr, nccl().GetErrorString(_r)); \
    } while (0)

}  // namespace

// The exchange step runs on its own stream so that it overlaps the next batch's kernels: the comm
// stream first waits for what the ingest stream has enqueued so far, and whoever needs the result waits for ev_gather.
int ti_comm_ready(ti_ctx* ctx) {
    if (ctx->s_comm) return TI_OK;
    int prio_low = 0, prio_high = 0;  // the exchange stream's (small) kernels go first whenever an SM has room
    TI_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    const char* prio_env = getenv("TI_EXCHANGE_PRIORITY");  // bring-up: "0" = default priority
    TI_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->s_comm, cudaStreamNonBlocking, (prio_env && prio_env[0] == '0') ? prio_low : prio_high));
    TI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
    TI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_gather, cudaEventDisableTiming));
    TI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_counts, cudaEventDisableTiming));
    TI_CUDA(ctx, cudaMalloc(&ctx->d_comm_words, 256 * sizeof(uint32_t)));
    TI_CUDA(ctx, cudaMemset(ctx->d_comm_words, 0, 256 * sizeof(uint32_t)));
    TI_CUDA(ctx, cudaMallocHost(&ctx->h_comm_words, 256 * sizeof(uint32_t)));
    return TI_OK;
}

int ti_comm_follow_compute(ti_ctx* ctx) {
    TI_CUDA(ctx, cudaEventRecord(ctx->ev_compute, ctx->stream));
    TI_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comm, ctx->ev_compute, 0));
    return TI_OK;
}

void ti_nccl_teardown(ti_ctx* ctx) {
    if (ctx->nccl_comm && nccl().handle) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    if (ctx->s_comm) {
        cudaStreamSynchronize(ctx->s_comm);
        cudaEventDestroy(ctx->ev_compute); cudaEventDestroy(ctx->ev_gather); cudaEventDestroy(ctx->ev_counts);
        for (auto& ev : ctx->ev_fence) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
        ctx->counts_pending = false; ctx->gather_pending = false;
        cudaFree(ctx->d_comm_words); cudaFreeHost(ctx->h_comm_words);
        cudaStreamDestroy(ctx->s_comm);
        ctx->s_comm = nullptr; ctx->d_comm_words = nullptr; ctx->h_comm_words = nullptr;
    }
}

extern "C" {

int ti_nccl_unique_id(void* id128) {
    if (!id128) return ti::fail(nullptr, TI_EINVAL, "ti_nccl_unique_id: null buffer");
    NcclApi& a = nccl();
    if (!a.handle) return ti::fail(nullptr, TI_ENCCL, "%s", a.why.c_str());
    NcclId id;
    ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) return ti::fail(nullptr, TI_ENCCL, "ncclGetUniqueId: %s", a.GetErrorString(r));
    memcpy(id128, &id, sizeof id);
    return TI_OK;
}

int ti_nccl_init(ti_ctx* ctx, const void* id128, int rank, int world) {
    if (!ctx) return TI_EINVAL;
    if (!id128 || world <= 0 || rank < 0 || rank >= world) return ti::fail(ctx, TI_EINVAL, "ti_nccl_init: bad rank/world");
    NcclApi& a = nccl();
    if (!a.handle) return ti::fail(ctx, TI_ENCCL, "%s", a.why.c_str());
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    ti_nccl_teardown(ctx);
    NcclId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    TI_NCCL(ctx, a.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return TI_OK;
}

static int gather_on_comm_stream(ti_ctx* ctx, const void* local, void* gathered, const uint64_t* bytes_per_rank, int root, bool follow) {
    if (!ctx) return TI_EINVAL;
    if (!ctx->nccl_comm) return ti::fail(ctx, TI_ESTATE, "ti_gather_clouds: call ti_nccl_init first");
    if (!bytes_per_rank || root < 0 || root >= ctx->world) return ti::fail(ctx, TI_EINVAL, "ti_gather_clouds: bad arguments");
    if (ctx->rank == root && !gathered) return ti::fail(ctx, TI_EINVAL, "ti_gather_clouds: root needs a gathered buffer");
    NcclApi& a = nccl();
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t mine = bytes_per_rank[ctx->rank];
    if (mine && !local) return ti::fail(ctx, TI_EINVAL, "ti_gather_clouds: null local buffer");
    int rc = ti_comm_ready(ctx);
    if (rc != TI_OK) return rc;
    if (follow && (rc = ti_comm_follow_compute(ctx)) != TI_OK) return rc;
    if (ctx->rank == root) {
        uint64_t off = 0;
        for (int r = 0; r < root; ++r) off += bytes_per_rank[r];
        if (mine)  // own slice: device-to-device copy, no NCCL self send
            TI_CUDA(ctx, cudaMemcpyAsync((uint8_t*)gathered + off, local, mine, cudaMemcpyDeviceToDevice, ctx->s_comm));
        ncclResult_t first = 0;  // the group is always closed, whatever a Recv answers
        ncclResult_t r0 = a.GroupStart();
        if (r0 != 0) return ti::fail(ctx, TI_ENCCL, "ncclGroupStart failed: %s", a.GetErrorString(r0));
        off = 0;
        for (int r = 0; r < ctx->world; ++r) {
            if (r != root && bytes_per_rank[r] && first == 0)
                first = a.Recv((uint8_t*)gathered + off, bytes_per_rank[r], NCCL_UINT8, r, comm, ctx->s_comm);
            off += bytes_per_rank[r];
        }
        const ncclResult_t r1 = a.GroupEnd();
        if (first != 0) return ti::fail(ctx, TI_ENCCL, "ncclRecv failed: %s", a.GetErrorString(first));
        if (r1 != 0) return ti::fail(ctx, TI_ENCCL, "ncclGroupEnd failed: %s", a.GetErrorString(r1));
    } else if (mine) {
        TI_NCCL(ctx, a.Send(local, mine, NCCL_UINT8, root, comm, ctx->s_comm));
    }
    TI_CUDA(ctx, cudaEventRecord(ctx->ev_gather, ctx->s_comm));
    ctx->gather_pending = true;
    return TI_OK;
}

int ti_gather_clouds(ti_ctx* ctx, const void* local, void* gathered, const uint64_t* bytes_per_rank, int root) {
    return gather_on_comm_stream(ctx, local, gathered, bytes_per_rank, root, /*follow=*/true);
}

int ti_gather_records(ti_ctx* ctx, const uint64_t* records, uint64_t* gathered, const uint32_t* counts, int root) {
    if (!ctx) return TI_EINVAL;
    if (!counts) return ti::fail(ctx, TI_EINVAL, "ti_gather_records: null counts");
    if (ctx->world > 256) return ti::fail(ctx, TI_EINVAL, "ti_gather_records: at most 256 ranks");
    uint64_t bytes[256];
    for (int r = 0; r < ctx->world; ++r) bytes[r] = (uint64_t)counts[r] * sizeof(uint64_t);
    // ordered on the exchange stream behind the ti_gather_counts_begin that sized it - NOT behind what the ingest stream
    // has been given since (the next batch's kernels)
    return gather_on_comm_stream(ctx, records, gathered, bytes, root, /*follow=*/false);
}

int ti_exchange_fence(ti_ctx* ctx, uint64_t* fence) {
    if (!ctx) return TI_EINVAL;
    if (!fence) return ti::fail(ctx, TI_EINVAL, "ti_exchange_fence: null fence");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = ti_comm_ready(ctx);
    if (rc != TI_OK) return rc;
    constexpr uint64_t kRing = sizeof(ctx->ev_fence) / sizeof(ctx->ev_fence[0]);
    const uint64_t f = ctx->fences + 1;
    if (!ctx->ev_fence[f % kRing]) TI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fence[f % kRing], cudaEventDisableTiming));
    TI_CUDA(ctx, cudaEventRecord(ctx->ev_fence[f % kRing], ctx->s_comm));
    ctx->fences = f;
    *fence = f;
    return TI_OK;
}

int ti_exchange_wait(ti_ctx* ctx, uint64_t fence, int on_stream) {
    if (!ctx) return TI_EINVAL;
    if (fence == 0) return TI_OK;
    constexpr uint64_t kRing = sizeof(ctx->ev_fence) / sizeof(ctx->ev_fence[0]);
    if (fence > ctx->fences) return ti::fail(ctx, TI_EINVAL, "ti_exchange_wait: fence %llu was never issued", (unsigned long long)fence);
    if (ctx->fences - fence >= kRing)
        return ti::fail(ctx, TI_EINVAL, "ti_exchange_wait: fence %llu is more than %d fences old", (unsigned long long)fence, (int)kRing);
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (on_stream) TI_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_fence[fence % kRing], 0));
    else TI_CUDA(ctx, cudaEventSynchronize(ctx->ev_fence[fence % kRing]));
    return TI_OK;
}

int ti_gather_wait(ti_ctx* ctx, int on_stream) {
    if (!ctx) return TI_EINVAL;
    if (!ctx->gather_pending) return TI_OK;
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (on_stream) {
        TI_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_gather, 0));
    } else {
        TI_CUDA(ctx, cudaEventSynchronize(ctx->ev_gather));
        ctx->gather_pending = false;
    }
    return TI_OK;
}

int ti_gather_counts_begin(ti_ctx* ctx, const uint32_t* n_local) {
    if (!ctx) return TI_EINVAL;
    if (!ctx->nccl_comm) return ti::fail(ctx, TI_ESTATE, "ti_gather_counts_begin: call ti_nccl_init first");
    if (!n_local) return ti::fail(ctx, TI_EINVAL, "ti_gather_counts_begin: null argument");
    if (ctx->world > 128) return ti::fail(ctx, TI_EINVAL, "ti_gather_counts_begin: at most 128 ranks");
    if (ctx->counts_pending) return ti::fail(ctx, TI_ESTATE, "ti_gather_counts_begin: the previous counts were not collected (ti_gather_counts_finish)");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ti_comm_ready(ctx);
    if (rc != TI_OK) return rc;
    if ((rc = ti_comm_follow_compute(ctx)) != TI_OK) return rc;
    TI_NCCL(ctx, nccl().AllGather(n_local, ctx->d_comm_words, 1, NCCL_UINT32, (ncclComm_t)ctx->nccl_comm, ctx->s_comm));
    TI_CUDA(ctx, cudaMemcpyAsync(ctx->h_comm_words, ctx->d_comm_words, sizeof(uint32_t) * ctx->world, cudaMemcpyDeviceToHost, ctx->s_comm));
    TI_CUDA(ctx, cudaEventRecord(ctx->ev_counts, ctx->s_comm));
    ctx->counts_pending = true;
    return TI_OK;
}

int ti_gather_counts_finish(ti_ctx* ctx, uint32_t* counts) {
    if (!ctx) return TI_EINVAL;
    if (!counts) return ti::fail(ctx, TI_EINVAL, "ti_gather_counts_finish: null argument");
    if (!ctx->counts_pending) return ti::fail(ctx, TI_ESTATE, "ti_gather_counts_finish: no ti_gather_counts_begin is outstanding");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaEventSynchronize(ctx->ev_counts));  // the ingest stream keeps running what it was given meanwhile
    for (int r = 0; r < ctx->world; ++r) counts[r] = ctx->h_comm_words[r];
    ctx->counts_pending = false;
    return TI_OK;
}

int ti_gather_counts(ti_ctx* ctx, const uint32_t* n_local, uint32_t* counts) {
    const int rc = ti_gather_counts_begin(ctx, n_local);
    return rc != TI_OK ? rc : ti_gather_counts_finish(ctx, counts);
}

int ti_nccl_barrier(ti_ctx* ctx) {
    if (!ctx) return TI_EINVAL;
    if (!ctx->nccl_comm) return ti::fail(ctx, TI_ESTATE, "ti_nccl_barrier: call ti_nccl_init first");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = ti_comm_ready(ctx);
    if (rc != TI_OK) return rc;
    int* d_flag = reinterpret_cast<int*>(ctx->d_comm_words + 255);  // stays 0: a sum of zeros
    TI_NCCL(ctx, nccl().AllReduce(d_flag, d_flag, 1, NCCL_INT32, NCCL_SUM, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    TI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TI_OK;
}

int ti_peer_alloc(ti_ctx* ctx, uint64_t bytes, void** dev_ptr, void* handle64) {
    if (!ctx) return TI_EINVAL;
    if (!dev_ptr || !handle64 || bytes == 0) return ti::fail(ctx, TI_EINVAL, "ti_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    TI_CUDA(ctx, cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return ti::fail(ctx, TI_ECUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, sizeof h);
    *dev_ptr = p;
    return TI_OK;
}

int ti_peer_open(ti_ctx* ctx, const void* handle64, void** dev_ptr) {
    if (!ctx) return TI_EINVAL;
    if (!handle64 || !dev_ptr) return ti::fail(ctx, TI_EINVAL, "ti_peer_open: bad arguments");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    TI_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TI_OK;
}

int ti_peer_close(ti_ctx* ctx, void* dev_ptr) {
    if (!ctx) return TI_EINVAL;
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return TI_OK;
}

int ti_peer_free(ti_ctx* ctx, void* dev_ptr) {
    if (!ctx) return TI_EINVAL;
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaFree(dev_ptr));
    return TI_OK;
}

}  // extern "C"
