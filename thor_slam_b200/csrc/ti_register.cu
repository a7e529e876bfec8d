// Depth -> RGB registration: one colour per depth pixel.
//
// The reference hands nvblox a 16UC1 depth image and an rgb8 image that come from different sensors unless
// depth_align_to_rgb is set (thor_slam/camera/drivers/luxonis.py:1018-1051); get_rgbd_extrinsics() gives
// depth (CAM_B) -> RGB (CAM_A) (luxonis.py:1068-1091).  This kernel does the lookup nvblox would do per voxel,
// once per depth pixel:
//
//     ray = A * [u - cx_d, v - cy_d, 1]        A = R_rgb<-depth * diag(1/fx_d, 1/fy_d, 1)   (float64 on the host, rounded once)
//     p   = z * ray + t                         z = d * 0.001
//     (ur, vr) = (fx_r * p.x / p.z + cx_r, fy_r * p.y / p.z + cy_r),  pinhole on the RGB image's K (no distortion, as nvblox)
//     colour = rgb[rint(vr)][rint(ur)]          nearest pixel, round-half-even; (0,0,0) when d == 0, p.z <= 0 or outside
//
// Every float operation is a separately rounded IEEE fp32 multiply / add / divide (no FMA contraction), in a fixed order,
// so a float32 numpy restatement (oracle/backproject.py:register_colour) selects bit-identical pixels.
// Streaming kernel: 2 B/px depth in, 3 B/px colour out, RGB gathers are spatially coherent (L1/L2 hits).
#include "ti_common.cuh"
#include "ti_register.cuh"

namespace ti {

constexpr int RG_THREADS = 256;

struct RegJobDev {
    const uint16_t* depth;
    const uint8_t* rgb;
    uint8_t* colour;
    uint64_t depth_stride, rgb_stride, colour_stride;
    RegConst reg;
    int dw, dh;
};

// reg_pixel_index (ti_register.cuh) is the arithmetic above with the two divisions replaced, off the rounding boundaries, by one
// reciprocal: same selected pixel, a third fewer instructions
__device__ __forceinline__ uint32_t register_pixel(const RegJobDev& J, const uint8_t* rgb, int u, int v, uint32_t d) {
    const int idx = reg_pixel_index(J.reg, u, v, d);
    if (idx < 0) return 0u;
    const uint8_t* s = rgb + (size_t)idx * 3;
    return (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16);
}

__global__ void __launch_bounds__(RG_THREADS) register_colour_kernel(const RegJobDev J, int n_batch) {
    const uint64_t npx = (uint64_t)J.dw * J.dh;
    const uint64_t total = npx * n_batch;
    for (uint64_t i = (uint64_t)blockIdx.x * RG_THREADS + threadIdx.x; i < total; i += (uint64_t)gridDim.x * RG_THREADS) {
        const uint64_t b = i / npx;
        const uint32_t px = (uint32_t)(i - b * npx);
        const int v = (int)(px / (uint32_t)J.dw), u = (int)(px - (uint32_t)v * J.dw);
        const uint16_t* depth = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(J.depth) + b * J.depth_stride);
        const uint32_t c = register_pixel(J, J.rgb + b * J.rgb_stride, u, v, depth[px]);
        uint8_t* o = J.colour + b * J.colour_stride + (size_t)px * 3;
        o[0] = (uint8_t)c; o[1] = (uint8_t)(c >> 8); o[2] = (uint8_t)(c >> 16);
    }
}

// 4 consecutive depth pixels per thread: one 64-bit depth load, twelve colour bytes as three 32-bit streaming stores.
__global__ void __launch_bounds__(RG_THREADS) register_colour_vec_kernel(const RegJobDev J, int n_batch) {
    const uint32_t quads_per_frame = (uint32_t)J.dw * J.dh / 4;
    const uint64_t total = (uint64_t)quads_per_frame * n_batch;
    for (uint64_t i = (uint64_t)blockIdx.x * RG_THREADS + threadIdx.x; i < total; i += (uint64_t)gridDim.x * RG_THREADS) {
        const uint64_t b = i / quads_per_frame;
        const uint32_t px = (uint32_t)(i - b * quads_per_frame) * 4;
        const int v = (int)(px / (uint32_t)J.dw), u = (int)(px - (uint32_t)v * J.dw);  // dw % 4 == 0: the quad stays in one row
        const uint2 d = ld_stream_u2(reinterpret_cast<const uint8_t*>(J.depth) + b * J.depth_stride + (size_t)px * 2);
        const uint8_t* rgb = J.rgb + b * J.rgb_stride;
        const uint32_t c0 = register_pixel(J, rgb, u, v, d.x & 0xFFFFu), c1 = register_pixel(J, rgb, u + 1, v, d.x >> 16);
        const uint32_t c2 = register_pixel(J, rgb, u + 2, v, d.y & 0xFFFFu), c3 = register_pixel(J, rgb, u + 3, v, d.y >> 16);
        uint8_t* o = J.colour + b * J.colour_stride + (size_t)px * 3;
        st_stream_u1(o, c0 | (c1 << 24));
        st_stream_u1(o + 4, (c1 >> 8) | (c2 << 16));
        st_stream_u1(o + 8, (c2 >> 16) | (c3 << 8));
    }
}

int launch_register_colour(ti_ctx* ctx, int camera, const uint16_t* depth, const uint8_t* rgb, uint8_t* colour, int n_batch,
                           uint64_t depth_stride, uint64_t rgb_stride, uint64_t colour_stride) {
    if (n_batch <= 0) return TI_OK;
    const CameraSlot& C = ctx->cams[camera];
    RegJobDev J{};
    J.depth = depth; J.rgb = rgb; J.colour = colour;
    J.depth_stride = depth_stride ? depth_stride : (uint64_t)C.reg_dw * C.reg_dh * 2;
    J.rgb_stride = rgb_stride ? rgb_stride : (uint64_t)C.reg_rw * C.reg_rh * 3;
    J.colour_stride = colour_stride ? colour_stride : (uint64_t)C.reg_dw * C.reg_dh * 3;
    if (J.depth_stride % 2) return fail(ctx, TI_EINVAL, "register_colour: depth frame stride must be even");
    J.reg = reg_constants(C);
    J.dw = C.reg_dw; J.dh = C.reg_dh;
    const uint64_t total = (uint64_t)J.dw * J.dh * n_batch;
    const bool vec = J.dw % 4 == 0 && ((uintptr_t)depth % 8 == 0) && (J.depth_stride % 8 == 0) && ((uintptr_t)colour % 4 == 0) &&
                     (J.colour_stride % 4 == 0);
    if (vec) {
        const int grid = (int)std::min<uint64_t>((total / 4 + RG_THREADS - 1) / RG_THREADS, (uint64_t)ctx->sm_count * 16);
        TI_LAUNCH(register_colour_vec_kernel, grid, RG_THREADS, 0, ctx->stream, J, n_batch);
        TI_CHECK_LAUNCH(ctx);
        return TI_OK;
    }
    const int grid = (int)std::min<uint64_t>((total + RG_THREADS - 1) / RG_THREADS, (uint64_t)ctx->sm_count * 16);
    TI_LAUNCH(register_colour_kernel, grid, RG_THREADS, 0, ctx->stream, J, n_batch);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

}  // namespace ti
