// The cloud exchange as our own kernels over NVLink 5 / NVSwitch peer memory - no collective library, no host round trip.
//
// The fusing rank owns an INBOX (ti_peer_alloc: cudaMalloc + CUDA-IPC handle) that every rank maps (ti_peer_open):
//
//     [ 128-byte header: n_records | done | gen | error ][ u64 records ... ]
//
// A producing rank appends its variable-length record list (ti_voxel_cloud output) with
//   1. push_reserve_kernel : one thread waits until the inbox is at generation `gen` (the root has emptied it for this round),
//                            reads the local count - a DEVICE value, the host never sees it - and reserves that many slots with
//                            one system-scope atomicAdd on the root's header;
//   2. push_copy_kernel    : coalesced 16-byte stores straight into the root's HBM through the peer mapping; the last block to
//                            finish fences and bumps `done`.
// The root takes a round with inbox_wait_kernel (spins until done == world), a copy into the caller's buffer, and
// inbox_release_kernel (zero the header, gen + 1).  Everything is enqueued on the context's exchange stream (the one
// ti_gather_clouds uses), behind an event of the ingest stream, so batch k's exchange overlaps batch k + 1's kernels.
// Two inboxes used alternately keep a producer from ever waiting for the root in steady state.
//
// Every spin has a deadline (TI_PUSH_TIMEOUT_NS): a missing peer turns into header.error / a TI_ECUDA from ti_inbox_take's
// caller-visible status word, never into a hung GPU.
#include <stdlib.h>

#include "ti_common.cuh"
#include "ti_tma.cuh"

namespace ti {

constexpr uint64_t TI_PUSH_TIMEOUT_NS = 4000000000ull;  // 4 s
constexpr int PUSH_THREADS = 64;   // small CTAs of <= 32 registers per thread: they fit beside the resident CTAs of the persistent
                                    // ingest kernels (which leave ~3 K registers and no shared memory to spare per SM), so the
                                    // copies really run UNDER the next batch's kernels instead of queueing behind them

struct InboxHdr {
    uint32_t n_records;  // slots reserved so far in this generation (may exceed the capacity: the list was truncated)
    uint32_t done;       // ranks whose records have all landed
    uint32_t gen;        // generation the inbox currently accepts
    uint32_t error;      // != 0: somebody's deadline passed
    uint32_t pad[28];
};
static_assert(sizeof(InboxHdr) == TI_INBOX_HEADER_BYTES, "inbox header is 128 bytes");

#ifndef TI_EMULATE
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t now_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// words: [0] base, [1] n to copy, [2] blocks finished (self-resetting), [3] sticky error
__global__ void push_reserve_kernel(InboxHdr* hdr, uint64_t* records, const uint32_t* n_local, uint64_t records_capacity, uint32_t gen,
                                    uint64_t capacity, uint32_t* words, int debug) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t t0 = now_ns();
    while (ld_acquire_sys(&hdr->gen) != gen) {
        if (now_ns() - t0 > TI_PUSH_TIMEOUT_NS) {
            words[0] = 0; words[1] = 0; words[3] = 1;
            atomicExch_system(&hdr->error, 1u);
            return;
        }
        __nanosleep(500);
    }
    // Every rank reserves an EVEN number of slots, so every rank's run starts 16-byte aligned in the inbox (bulk copies need
    // that); an odd list is padded with one zero record - no voxel encodes as 0 (key fields are biased, |k| < 16383).
    uint32_t n = (debug & 16) ? 0u : *n_local;  // bring-up switch 16: the whole protocol, no payload
    if ((uint64_t)n > records_capacity) n = (uint32_t)records_capacity;  // a truncated list reports more than it holds (ti_voxel_cloud)
    if (n & 1u) {
        if ((uint64_t)n < records_capacity) { records[n] = 0ull; ++n; }  // the spare slot takes the pad
        else --n;                                                         // a full, odd list: its last record stays behind
    }
    const uint32_t base = atomicAdd_system(&hdr->n_records, n);
    words[0] = base;
    words[1] = (uint64_t)base >= capacity ? 0u : (uint32_t)min((unsigned long long)n, (unsigned long long)((capacity & ~1ull) - base));
}

__global__ void __launch_bounds__(PUSH_THREADS, 24) push_copy_kernel(InboxHdr* hdr, const uint64_t* __restrict__ local, uint32_t* words) {
    uint64_t* remote = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(hdr) + TI_INBOX_HEADER_BYTES);
    const uint32_t base = words[0], n = words[1];
    // 16-byte stores where source and destination line up (base even), 8-byte otherwise.  Four independent 16-byte loads are
    // in flight per thread before the first store: the few small CTAs that fit beside the ingest kernels have to cover the
    // NVLink round trip with bytes in flight, not with thread count.
    const uint64_t tid = (uint64_t)blockIdx.x * PUSH_THREADS + threadIdx.x, nth = (uint64_t)gridDim.x * PUSH_THREADS;
    if ((base & 1u) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(local);
        uint4* d4 = reinterpret_cast<uint4*>(remote + base);
        const uint64_t n4 = n / 2;
        uint64_t i = tid;
        for (; i + 3 * nth < n4; i += 4 * nth) {
            const uint4 a = s4[i], b = s4[i + nth], c = s4[i + 2 * nth], d = s4[i + 3 * nth];
            d4[i] = a; d4[i + nth] = b; d4[i + 2 * nth] = c; d4[i + 3 * nth] = d;
        }
        for (; i < n4; i += nth) d4[i] = s4[i];
        if (tid == 0 && (n & 1u)) remote[base + n - 1] = local[n - 1];
    } else {
        for (uint64_t i = tid; i < n; i += nth) remote[base + i] = local[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t finished = atomicAdd(&words[2], 1u) + 1;
        if (finished == gridDim.x) {
            words[2] = 0;
            __threadfence_system();
            if (words[3] == 0) atomicAdd_system(&hdr->done, 1u);  // a rank whose reservation timed out never reports done
        }
    }
}

// The same copy driven by the TMA unit: ONE warp per CTA, one lane issuing bulk copies local HBM -> shared memory -> the root's
// HBM (cp.async.bulk both ways), three 4 KB stages in flight.  It costs the SM no load/store instructions and a handful of issue
// slots, so the next batch's kernels - which share the SM - keep their pace; its 12 KB of shared memory fit beside the voxel
// kernel's CTAs (not beside the rectify kernel's, which fill the SM: the copy then waits for that kernel's last CTAs).
constexpr int TPC_STAGE_BYTES = 4096;
constexpr int TPC_STAGES = 3;

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(32) push_copy_tma_kernel(InboxHdr* hdr, const uint64_t* __restrict__ local, uint32_t* words) {
    __shared__ __align__(128) uint8_t buf[TPC_STAGES][TPC_STAGE_BYTES];
    __shared__ __align__(8) uint64_t full[TPC_STAGES];
    if (threadIdx.x == 0) {
        uint8_t* remote = reinterpret_cast<uint8_t*>(hdr) + TI_INBOX_HEADER_BYTES + (uint64_t)words[0] * 8u;  // base is even: 16-byte aligned
        const uint8_t* src = reinterpret_cast<const uint8_t*>(local);
        const uint64_t bytes = (uint64_t)words[1] * 8u;                                                    // a multiple of 16
        const uint64_t n_chunks = (bytes + TPC_STAGE_BYTES - 1) / TPC_STAGE_BYTES;
        for (int s = 0; s < TPC_STAGES; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
        auto chunk_bytes = [&](uint64_t c) { return (uint32_t)min((uint64_t)TPC_STAGE_BYTES, bytes - c * TPC_STAGE_BYTES); };
        // chunks blockIdx.x, blockIdx.x + gridDim.x, ... ; load i + STAGES - 1 is requested before store i is issued
        uint64_t c_load = blockIdx.x, c_store = blockIdx.x;
        uint32_t i_load = 0, i_store = 0;
        for (; i_load < TPC_STAGES - 1 && c_load < n_chunks; ++i_load, c_load += gridDim.x) {
            const uint32_t nb = chunk_bytes(c_load);
            mbar_arrive_expect_tx(full + i_load % TPC_STAGES, nb);
            bulk_load_1d(buf[i_load % TPC_STAGES], src + c_load * TPC_STAGE_BYTES, nb, full + i_load % TPC_STAGES);
        }
        for (; c_store < n_chunks; ++i_store, c_store += gridDim.x) {
            if (c_load < n_chunks) {
                const int s = i_load % TPC_STAGES;  // last read by store i_load - STAGES: at most STAGES - 2 younger stores may still be reading
                bulk_wait_read<TPC_STAGES - 2>();
                const uint32_t nb = chunk_bytes(c_load);
                mbar_arrive_expect_tx(full + s, nb);
                bulk_load_1d(buf[s], src + c_load * TPC_STAGE_BYTES, nb, full + s);
                ++i_load; c_load += gridDim.x;
            }
            const int s = i_store % TPC_STAGES;
            mbar_wait(full + s, (i_store / TPC_STAGES) & 1u);
            bulk_store_1d(remote + c_store * TPC_STAGE_BYTES, buf[s], chunk_bytes(c_store));
            bulk_commit();
        }
        bulk_wait_all();  // every store of this CTA has been performed
        __threadfence_system();
        const uint32_t finished = atomicAdd(&words[2], 1u) + 1;
        if (finished == gridDim.x) {
            words[2] = 0;
            __threadfence_system();
            if (words[3] == 0) atomicAdd_system(&hdr->done, 1u);
        }
    }
}

// status: [0] records in the inbox (reserved; may exceed capacity), [1] error
__global__ void inbox_wait_kernel(InboxHdr* hdr, uint32_t world, uint32_t* status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t t0 = now_ns();
    while (ld_acquire_sys(&hdr->done) < world) {
        if (now_ns() - t0 > TI_PUSH_TIMEOUT_NS || ld_acquire_sys(&hdr->error)) {
            status[0] = 0; status[1] = 1;
            return;
        }
        __nanosleep(500);
    }
    status[0] = ld_acquire_sys(&hdr->n_records);
    status[1] = ld_acquire_sys(&hdr->error);
}

__global__ void __launch_bounds__(PUSH_THREADS, 32) inbox_copy_kernel(const InboxHdr* hdr, uint64_t capacity, const uint32_t* status, uint64_t* dst,
                                                                  uint64_t dst_capacity) {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(reinterpret_cast<const uint8_t*>(hdr) + TI_INBOX_HEADER_BYTES);
    const uint64_t n = min(min((uint64_t)status[0], capacity), dst_capacity);
    for (uint64_t i = (uint64_t)blockIdx.x * PUSH_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * PUSH_THREADS) dst[i] = src[i];
}

__global__ void inbox_release_kernel(InboxHdr* hdr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    hdr->n_records = 0;
    hdr->done = 0;
    __threadfence_system();
    st_release_sys(&hdr->gen, hdr->gen + 1);
}
#endif  // !TI_EMULATE

}  // namespace ti

using namespace ti;

#ifndef TI_EMULATE
int ti_comm_ready(ti_ctx* ctx);           // ti_nccl.cu
int ti_comm_follow_compute(ti_ctx* ctx);  // ti_nccl.cu

static int push_debug() {  // bring-up switches, TI_PUSH_DEBUG: 16 = the whole protocol without payload, 32 = no kernels at all
    static const int v = getenv("TI_PUSH_DEBUG") ? atoi(getenv("TI_PUSH_DEBUG")) : 0;
    return v;
}

extern "C" {

int ti_inbox_init(ti_ctx* ctx, void* inbox) {
    if (!ctx) return TI_EINVAL;
    if (!inbox) return fail(ctx, TI_EINVAL, "ti_inbox_init: null inbox");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaMemset(inbox, 0, TI_INBOX_HEADER_BYTES));
    return TI_OK;
}

int ti_cloud_push(ti_ctx* ctx, uint64_t* records, const uint32_t* n_records, uint64_t records_capacity, void* inbox, uint64_t inbox_capacity,
                  uint32_t gen) {
    if (!ctx) return TI_EINVAL;
    if (!records || !n_records || !inbox) return fail(ctx, TI_EINVAL, "ti_cloud_push: null argument");
    if ((uintptr_t)records % 16 || (uintptr_t)inbox % 16) return fail(ctx, TI_EINVAL, "ti_cloud_push: buffers must be 16-byte aligned");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ti_comm_ready(ctx);
    if (rc != TI_OK) return rc;
    if ((rc = ti_comm_follow_compute(ctx)) != TI_OK) return rc;
    InboxHdr* hdr = reinterpret_cast<InboxHdr*>(inbox);
    uint32_t* words = ctx->d_comm_words + 128;
    if (push_debug() & 32) {  // bring-up: no kernels at all on the exchange stream
        TI_CUDA(ctx, cudaEventRecord(ctx->ev_gather, ctx->s_comm));
        ctx->gather_pending = true;
        return TI_OK;
    }
    push_reserve_kernel<<<1, 32, 0, ctx->s_comm>>>(hdr, records, n_records, records_capacity, gen, inbox_capacity, words, push_debug());
    TI_CHECK_LAUNCH(ctx);
    if (ctx->push_tma) push_copy_tma_kernel<<<ctx->push_blocks > 0 ? ctx->push_blocks : ctx->sm_count, 32, 0, ctx->s_comm>>>(hdr, records, words);
    else push_copy_kernel<<<ctx->push_blocks > 0 ? ctx->push_blocks : 2 * ctx->sm_count, PUSH_THREADS, 0, ctx->s_comm>>>(hdr, records, words);
    TI_CHECK_LAUNCH(ctx);
    TI_CUDA(ctx, cudaEventRecord(ctx->ev_gather, ctx->s_comm));
    ctx->gather_pending = true;
    return TI_OK;
}

int ti_inbox_take(ti_ctx* ctx, void* inbox, uint64_t inbox_capacity, uint32_t world, uint64_t* dst, uint64_t dst_capacity, uint32_t* status) {
    if (!ctx) return TI_EINVAL;
    if (!inbox || !status || (dst_capacity && !dst)) return fail(ctx, TI_EINVAL, "ti_inbox_take: null argument");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ti_comm_ready(ctx);
    if (rc != TI_OK) return rc;
    InboxHdr* hdr = reinterpret_cast<InboxHdr*>(inbox);
    if (push_debug() & 32) {
        TI_CUDA(ctx, cudaEventRecord(ctx->ev_gather, ctx->s_comm));
        ctx->gather_pending = true;
        return TI_OK;
    }
    inbox_wait_kernel<<<1, 32, 0, ctx->s_comm>>>(hdr, world, status);
    TI_CHECK_LAUNCH(ctx);
    if (dst_capacity) {
        inbox_copy_kernel<<<ctx->push_blocks > 0 ? ctx->push_blocks : 2 * ctx->sm_count, PUSH_THREADS, 0, ctx->s_comm>>>(hdr, inbox_capacity, status, dst, dst_capacity);
        TI_CHECK_LAUNCH(ctx);
    }
    inbox_release_kernel<<<1, 32, 0, ctx->s_comm>>>(hdr);
    TI_CHECK_LAUNCH(ctx);
    TI_CUDA(ctx, cudaEventRecord(ctx->ev_gather, ctx->s_comm));
    ctx->gather_pending = true;
    return TI_OK;
}

}  // extern "C"
#endif
