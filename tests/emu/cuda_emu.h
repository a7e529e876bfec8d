// Minimal CPU stand-in for the CUDA execution model  -  TEST INFRASTRUCTURE ONLY.
//
// tests/emu builds the library's kernel sources (thor_slam_b200/csrc/*.cu) with g++ and
// -DTI_EMULATE so that the *same source text* that nvcc compiles for sm_100a can be run here,
// one std::thread per CUDA thread, one CTA after another, against the oracle.  It exists to
// catch indexing / packing mistakes in a container without a GPU; it is never built into
// libthoringest.so and nothing under thor_slam_b200/ can reach it.
#pragma once

#include <cuda_runtime.h>  // vector types + host API declarations only

#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

namespace ti_emu {

struct ThreadState {
    uint3 tid{0, 0, 0};
    uint3 bid{0, 0, 0};
    dim3 bdim{1, 1, 1};
    dim3 gdim{1, 1, 1};
    std::barrier<>* cta_barrier = nullptr;
    std::barrier<>* warp_barrier = nullptr;
    uint32_t* warp_scratch = nullptr;  // 32 words shared by the warp
    uint8_t* dyn_smem = nullptr;
};
inline thread_local ThreadState tls;

inline const void* check_align(const void* p, size_t a) {
    if (reinterpret_cast<uintptr_t>(p) % a) {
        fprintf(stderr, "EMU: misaligned %zu-byte access at %p\n", a, p);
        abort();
    }
    return p;
}
inline void* check_align(void* p, size_t a) { return const_cast<void*>(check_align(static_cast<const void*>(p), a)); }

inline unsigned lane() { return tls.tid.x & 31u; }

inline void syncthreads() { tls.cta_barrier->arrive_and_wait(); }
inline void syncwarp() { tls.warp_barrier->arrive_and_wait(); }

// all-lanes-active warp collectives (the kernels only use full masks)
inline uint32_t shfl_idx(uint32_t v, int src) {
    tls.warp_scratch[lane()] = v;
    syncwarp();
    const uint32_t r = tls.warp_scratch[src & 31];
    syncwarp();
    return r;
}
inline uint32_t ballot(int pred) {
    tls.warp_scratch[lane()] = pred ? 1u : 0u;
    syncwarp();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= tls.warp_scratch[i] << i;
    syncwarp();
    return r;
}
inline uint32_t reduce_or(uint32_t v) {
    tls.warp_scratch[lane()] = v;
    syncwarp();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= tls.warp_scratch[i];
    syncwarp();
    return r;
}
inline uint32_t reduce_add(uint32_t v) {
    tls.warp_scratch[lane()] = v;
    syncwarp();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r += tls.warp_scratch[i];
    syncwarp();
    return r;
}

constexpr size_t MAX_DYN_SMEM = 232448;

template <typename F>
void run_grid(dim3 grid, dim3 block, size_t dyn_smem_bytes, F&& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads % 32) {
        fprintf(stderr, "EMU: block size must be a multiple of 32\n");
        abort();
    }
    if (dyn_smem_bytes > MAX_DYN_SMEM) {
        fprintf(stderr, "EMU: %zu B of dynamic shared memory exceeds the 227 KB limit\n", dyn_smem_bytes);
        abort();
    }
    const unsigned nwarps = nthreads / 32;
    std::vector<uint8_t> smem(dyn_smem_bytes + 128);
    uint8_t* smem_aligned = smem.data() + ((128 - reinterpret_cast<uintptr_t>(smem.data()) % 128) % 128);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::barrier<> cta_barrier(nthreads);
                std::vector<std::unique_ptr<std::barrier<>>> warp_barriers;
                std::vector<std::vector<uint32_t>> scratch(nwarps, std::vector<uint32_t>(32));
                for (unsigned w = 0; w < nwarps; ++w) warp_barriers.emplace_back(new std::barrier<>(32));
                std::memset(smem_aligned, 0xCD, dyn_smem_bytes);  // poison: uninitialised reads show up
                std::vector<std::thread> threads;
                threads.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t) {
                    threads.emplace_back([&, t]() {
                        tls.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        tls.bid = uint3{bx, by, bz};
                        tls.bdim = block;
                        tls.gdim = grid;
                        tls.cta_barrier = &cta_barrier;
                        tls.warp_barrier = warp_barriers[t / 32].get();
                        tls.warp_scratch = scratch[t / 32].data();
                        tls.dyn_smem = smem_aligned;
                        body();
                    });
                }
                for (auto& th : threads) th.join();
            }
}

}  // namespace ti_emu

// ---- CUDA spellings used by the kernel sources ------------------------------------------------
#undef __shared__
#undef __global__
#undef __device__
#undef __host__
#undef __launch_bounds__
#undef __grid_constant__
#undef __forceinline__
#define __global__
#define __device__
#define __host__
#define __launch_bounds__(...)
#define __grid_constant__
#define __forceinline__ inline __attribute__((always_inline))
#define threadIdx (ti_emu::tls.tid)
#define blockIdx (ti_emu::tls.bid)
#define blockDim (ti_emu::tls.bdim)
#define gridDim (ti_emu::tls.gdim)
#define __shared__ static
#define __syncthreads() ti_emu::syncthreads()
#define __syncwarp() ti_emu::syncwarp()

#define TI_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ti_emu::run_grid(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
#define TI_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(ti_emu::tls.dyn_smem)

inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t pool = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t b = (uint32_t)(pool >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;  // sign replication mode
        r |= b << (8 * i);
    }
    return r;
}
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline uint32_t __ballot_sync(uint32_t, int pred) { return ti_emu::ballot(pred); }
inline int __any_sync(uint32_t, int pred) { return ti_emu::ballot(pred) != 0u; }
inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) { return ti_emu::reduce_add(v); }
inline uint32_t __reduce_or_sync(uint32_t, uint32_t v) { return ti_emu::reduce_or(v); }
inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r = std::max(r, ti_emu::shfl_idx(v, i));
    return r;
}
template <typename T>
inline T __shfl_sync(uint32_t, T v, int src) {
    static_assert(sizeof(T) == 4, "emu shuffles 32-bit values");
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u = ti_emu::shfl_idx(u, src);
    std::memcpy(&v, &u, 4);
    return v;
}
template <typename T>
inline T __shfl_down_sync(uint32_t m, T v, int delta) {
    const int src = (int)ti_emu::lane() + delta;
    T r = __shfl_sync(m, v, src & 31);
    return src < 32 ? r : v;
}
template <typename T>
inline T __shfl_up_sync(uint32_t m, T v, int delta) {
    const int src = (int)ti_emu::lane() - delta;
    T r = __shfl_sync(m, v, src & 31);
    return src >= 0 ? r : v;
}
template <typename T>
inline T __shfl_xor_sync(uint32_t m, T v, int x) { return __shfl_sync(m, v, (int)ti_emu::lane() ^ x); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}
// dp2a: a = two u16 halves, b = four u8; lo uses b bytes 0,1 - hi uses bytes 2,3
inline uint32_t __dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return c + (a & 0xFFFF) * (b & 0xFF) + (a >> 16) * ((b >> 8) & 0xFF); }
inline uint32_t __dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { return c + (a & 0xFFFF) * ((b >> 16) & 0xFF) + (a >> 16) * ((b >> 24) & 0xFF); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline int __float2int_rn(float a) { return (int)std::nearbyintf(a); }
inline float __uint2float_rn(uint32_t v) { return (float)v; }
inline float __int2float_rn(int v) { return (float)v; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
using std::max;
using std::min;

// ---- CUDA runtime calls made by ti_api.cu: "device memory" is host memory here -----------------
namespace ti_emu {
inline cudaError_t ok() { return cudaSuccess; }
inline cudaError_t emu_malloc(void** p, size_t n) {
    *p = std::aligned_alloc(256, (n + 255) / 256 * 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
inline cudaError_t emu_malloc(T** p, size_t n) { return emu_malloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t emu_memcpy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h) {
    for (size_t r = 0; r < h; ++r) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t emu_props(cudaDeviceProp* p) {
    std::memset(p, 0, sizeof *p);
    p->major = 10; p->minor = 0; p->multiProcessorCount = 2;  // tiny "GPU": persistent loops still iterate
    return cudaSuccess;
}
}  // namespace ti_emu
#define cudaGetDeviceCount(n) (*(n) = 1, cudaSuccess)
#define cudaSetDevice(d) ti_emu::ok()
#undef cudaGetDeviceProperties
#define cudaGetDeviceProperties(p, d) ti_emu::emu_props(p)
#define cudaMalloc(p, n) ti_emu::emu_malloc((p), (n))
#define cudaFree(p) (std::free(p), cudaSuccess)
#define cudaMemcpy(d, s, n, k) (std::memcpy((d), (s), (n)), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, k, st) (std::memcpy((d), (s), (n)), cudaSuccess)
#define cudaMemcpy2DAsync(d, dp, s, sp, w, h, k, st) ti_emu::emu_memcpy2d((d), (dp), (s), (sp), (w), (h))
#define cudaMemsetAsync(p, v, n, st) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaMemset(p, v, n) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaStreamCreateWithFlags(s, f) (*(s) = nullptr, cudaSuccess)
#define cudaStreamDestroy(s) ti_emu::ok()
#define cudaStreamSynchronize(s) ti_emu::ok()
#define cudaStreamWaitEvent(s, e, f) ti_emu::ok()
#define cudaEventCreateWithFlags(e, f) (*(e) = nullptr, cudaSuccess)
#define cudaEventRecord(e, s) ti_emu::ok()
#define cudaEventSynchronize(e) ti_emu::ok()
#define cudaEventDestroy(e) ti_emu::ok()
#define cudaDeviceSynchronize() ti_emu::ok()
#define cudaGetLastError() ti_emu::ok()
#define cudaGetErrorString(e) "emulated"
