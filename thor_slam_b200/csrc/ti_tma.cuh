// TMA (cp.async.bulk[.tensor]) + mbarrier wrappers for sm_100a, with the CPU-emulation doubles used
// by tests/emu.  Only what the ingest kernels need: 3-D u8 tiled tensor maps, bulk 1-D copies,
// single-CTA mbarriers.
#pragma once
#include "ti_common.cuh"

#ifndef TI_EMULATE
#include <cuda.h>  // CUtensorMap + enums (types only; the encoder is fetched through the runtime)
#endif

namespace ti {

#ifdef TI_EMULATE
struct alignas(64) TiTensorMap {
    const uint8_t* base;
    int32_t dim[3];     // elements (bytes) per dimension, x fastest
    int64_t stride[3];  // bytes; stride[0] == 1
    int32_t box[3];
    int32_t elem;       // bytes per element (1 or 4); dim / box / coordinates count elements
    char pad[128 - 8 - 12 - 24 - 12 - 4 - 4];
};
static_assert(sizeof(TiTensorMap) == 128, "same size as CUtensorMap");
#else
typedef CUtensorMap TiTensorMap;
#endif

// host: describe a [n][h][w] u8 tensor (row pitch `pitch_y`, frame pitch `pitch_z` bytes) read in
// boxes of box_x x box_y x 1; out-of-bounds elements read as zero.
int tma_encode_u8_3d(ti_ctx* ctx, TiTensorMap* out, const void* base, int w, int h, int n, uint64_t pitch_y,
                     uint64_t pitch_z, int box_x, int box_y);
// same with elements of `elem_bytes` (1 or 4): w, box_x and the x coordinate of a load count elements, pitches stay bytes
int tma_encode_3d(ti_ctx* ctx, TiTensorMap* out, const void* base, int elem_bytes, int w, int h, int n, uint64_t pitch_y,
                  uint64_t pitch_z, int box_x, int box_y);

#if defined(TI_EMULATE)
// ---- emulation -----------------------------------------------------------------------------------
struct EmuMbar {  // lives in the 8 bytes of "shared memory" reserved for an mbarrier
    uint32_t completed;  // phases completed so far
    int16_t pending;     // arrivals still expected in the current phase
    int16_t init;
};
static_assert(sizeof(EmuMbar) == 8, "mbarrier is 8 bytes");
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    m->completed = 0; m->pending = (int16_t)count; m->init = (int16_t)count;
}
__device__ __forceinline__ void mbar_fence_init() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
__device__ __forceinline__ void emu_arrive(uint64_t* bar) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    if (__atomic_sub_fetch(&m->pending, 1, __ATOMIC_ACQ_REL) == 0) {
        __atomic_store_n(&m->pending, m->init, __ATOMIC_RELAXED);
        __atomic_add_fetch(&m->completed, 1, __ATOMIC_RELEASE);
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { emu_arrive(bar); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t) { emu_arrive(bar); }  // copies already done
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    uint64_t spins = 0;
    while ((__atomic_load_n(&m->completed, __ATOMIC_ACQUIRE) & 1u) == parity) {
        std::this_thread::yield();
        if (++spins == (1ull << 24)) {  // a kernel that deadlocks must not hang the test run: say where and stop
            fprintf(stderr, "EMU: mbarrier wait stuck: block %u thread %u barrier at smem+%ld parity %u completed %u pending %d\n",
                    ti_emu::tls.bid.x, ti_emu::tls.tid.x, (long)(reinterpret_cast<uint8_t*>(bar) - ti_emu::tls.dyn_smem), parity,
                    m->completed, (int)m->pending);
            fflush(stderr);
            std::this_thread::sleep_for(std::chrono::seconds(2));
            abort();
        }
    }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, unsigned) { mbar_wait(bar, parity); }
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    return (__atomic_load_n(&m->completed, __ATOMIC_ACQUIRE) & 1u) != parity;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const TiTensorMap* map, int x, int y, int z, uint64_t*) {
    uint8_t* d = static_cast<uint8_t*>(smem_dst);
    ti_emu::check_align(smem_dst, 128);
    const int es = map->elem;
    for (int by = 0; by < map->box[1]; ++by)
        for (int bx = 0; bx < map->box[0]; ++bx) {
            const int gx = x + bx, gy = y + by;
            const bool in = gx >= 0 && gx < map->dim[0] && gy >= 0 && gy < map->dim[1] && z >= 0 && z < map->dim[2];
            for (int e = 0; e < es; ++e)
                d[((size_t)by * map->box[0] + bx) * es + e] =
                    in ? map->base[(int64_t)z * map->stride[2] + (int64_t)gy * map->stride[1] + (int64_t)gx * es + e] : 0;
        }
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t*) {
    ti_emu::check_align(smem_dst, 16); ti_emu::check_align(gsrc, 16);
    std::memcpy(smem_dst, gsrc, bytes);
}
__device__ __forceinline__ void tma_prefetch_desc(const TiTensorMap*) {}
#elif defined(__CUDACC__)
// ---- sm_100a ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Wait used by the producer side (not latency critical): back off between probes so the spinning warp
// does not eat issue slots the consumer warps need.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)   // suspend-time hint (ns)
            : "memory");
        if (!done) __nanosleep(200);
    } while (!done);
}
// The hardware suspends the thread until the phase completes or `ns` have passed: no probe instructions in between.
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, unsigned ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 20000;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity);
// Poll with a fixed sleep between probes: for a lone issuer thread whose wait is about one item long.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
    while (!mbar_test(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // one non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const TiTensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const TiTensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace ti
