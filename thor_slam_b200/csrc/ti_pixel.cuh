// Per-pixel integer arithmetic shared by the convert and rectify kernels (bit-exact with OpenCV).
#pragma once
#include "ti_common.cuh"

namespace ti {

// cv2.COLOR_BGR2GRAY: 15-bit fixed point
__host__ __device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

__host__ __device__ __forceinline__ int sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// cv2.COLOR_YUV2RGB_NV12: BT.601 limited range, OpenCV's ITUR_BT_601_* constants, 20-bit shift
__host__ __device__ __forceinline__ void yuv_to_rgb(int y, int u, int v, int& r, int& g, int& b) {
    const int yy = (y > 16 ? y - 16 : 0) * 1220542;
    u -= 128;
    v -= 128;
    r = sat_u8((yy + 1673527 * v + (1 << 19)) >> 20);
    g = sat_u8((yy - 852492 * v - 409993 * u + (1 << 19)) >> 20);
    b = sat_u8((yy + 2116026 * u + (1 << 19)) >> 20);
}

// cv2.remap INTER_LINEAR on u8: weights (32-fx)(32-fy)/1024 are exact multiples of 2^-10, OpenCV
// scales them to 2^15 and rounds with +2^14 >> 15, which equals (S + 512) >> 10 on the exact sum S.
__host__ __device__ __forceinline__ uint32_t bilinear_u8(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11,
                                                         uint32_t fx, uint32_t fy) {
    const uint32_t top = t00 * (32u - fx) + t01 * fx;
    const uint32_t bot = t10 * (32u - fx) + t11 * fx;
    return (top * (32u - fy) + bot * fy + 512u) >> 10;
}


// Saturate two s32 to u8 and pack: result = { c[15:0], sat_u8(hi), sat_u8(lo) }  (one instruction on the GPU)
__device__ __forceinline__ uint32_t pack_sat_u8(int hi, int lo, uint32_t c) {
#if defined(TI_EMULATE) || !defined(__CUDACC__)
    return (c << 16) | ((uint32_t)sat_u8(hi) << 8) | (uint32_t)sat_u8(lo);
#else
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
    return d;
#endif
}

// NV12 chroma terms shared by the 2x2 pixels of one (U, V) pair (rounding constant folded in)
struct ChromaTerms {
    int r, g, b;
};
__device__ __forceinline__ ChromaTerms chroma_terms(int u, int v) {
    u -= 128;
    v -= 128;
    return ChromaTerms{1673527 * v + (1 << 19), -852492 * v - 409993 * u + (1 << 19), 2116026 * u + (1 << 19)};
}
__device__ __forceinline__ int luma_term(int y) { return (y > 16 ? y - 16 : 0) * 1220542; }

}  // namespace ti
