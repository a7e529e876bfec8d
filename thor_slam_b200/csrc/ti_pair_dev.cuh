// Device helpers shared by the pair-window remap kernels (ti_rectify_pair.cu: mono, ti_rectify_c3.cu: 3-channel):
// shared-window addressing, mbarrier wait / arrive by address, raw PRMT, LUT pixel-word expansion.
#pragma once
#include "ti_common.cuh"
#include "ti_tma.cuh"

namespace ti {

#ifdef TI_EMULATE
typedef const uint8_t* p4_addr_t;
__device__ __forceinline__ p4_addr_t p4_addr(const uint8_t* p) { return p; }
template <int OFF>
__device__ __forceinline__ uint32_t p4_lds(p4_addr_t a) { return *reinterpret_cast<const uint32_t*>(ti_emu::check_align(a + OFF, 4)); }
__device__ __forceinline__ uint2 ld_keep_u2(const void* p) { return *reinterpret_cast<const uint2*>(ti_emu::check_align(p, 8)); }
__device__ __forceinline__ uint4 p4_lds128(p4_addr_t a) { return *reinterpret_cast<const uint4*>(ti_emu::check_align(a, 16)); }
#else
typedef uint32_t p4_addr_t;
__device__ __forceinline__ p4_addr_t p4_addr(const uint8_t* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ uint32_t p4_lds(p4_addr_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint4 p4_lds128(p4_addr_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 ld_keep_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p), "l"(policy_evict_last()));
    return r;
}
#endif

// mbarrier wait / arrive on a barrier named by its shared-window address (no generic->shared conversion per use)
#ifdef TI_EMULATE
__device__ __forceinline__ void p4_wait(p4_addr_t bar, uint32_t parity) {
    mbar_wait(reinterpret_cast<uint64_t*>(const_cast<uint8_t*>(bar)), parity);
}
__device__ __forceinline__ void p4_arrive(p4_addr_t bar) { mbar_arrive(reinterpret_cast<uint64_t*>(const_cast<uint8_t*>(bar))); }
#else
__device__ __forceinline__ void p4_wait(p4_addr_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(20000u)  // suspend-time hint (ns): the warp sleeps until the phase completes instead of re-probing
        : "memory");
}
__device__ __forceinline__ void p4_arrive(p4_addr_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#endif

// One arrival per warp on `bar`, after every lane's work: the arriving lane is picked by elect.sync (no thread-index
// arithmetic in the loop); the emulation lets lane 0 do it.
__device__ __forceinline__ void p4_warp_arrive(p4_addr_t bar) {
#ifdef TI_EMULATE
    __syncwarp();
    if ((threadIdx.x & 31) == 0) p4_arrive(bar);
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "bar.warp.sync 0xffffffff;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(bar)
        : "memory");
#endif
}

// prmt.b32 with the selector taken as it is (CUDA's __byte_perm masks it with 0x7777 first - one more
// instruction per window row; bit 3 of a nibble only matters in flagged window words, see below)
__device__ __forceinline__ uint32_t p4_prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
#ifdef TI_EMULATE
    return __byte_perm(lo, hi, sel);
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
#endif
}

// LUT pixel word -> the two weight words of the pixel.  e = (32-fx) | fy << 6 | z << 11 | fx << 16 for a
// pixel with a tap inside the image (z = 1 iff fx = fy = 0), 0 otherwise.
//   Wtop = 64*(32-fy) * {32-fx, fx} - z      (16-bit halves; 65536 -> 65535 in the one case it occurs)
//   Wbot = 64*fy      * {32-fx, fx}
__device__ __forceinline__ void p4_expand(uint32_t e, uint32_t& wtop, uint32_t& wbot) {
    const uint32_t pw = e & 0x001F003Fu, fy64 = e & (31u << 6);
    wtop = (2048u - fy64) * pw - ((e >> 11) & 1u);
    wbot = fy64 * pw;
}

}  // namespace ti
