// depth (u16 mm) -> body-frame xyz (f32) + valid mask (u8) + per-frame valid count (u32).
//
// Restates, fused in one pass, what the reference leaves to nvblox (pinhole back-projection with
// the depth image's K - scripts/run_pipeline.py:247-256), the rig pose composition
// world_T_camera = rig_T_source @ source_T_camera (thor_slam/camera/rig.py:35-70), the RDF->FLU
// rotation (thor_slam/slam/adapters/isaac_ros.py:42-49) and the viewer's `depth > 0` mask / count
// (examples/rgbd_stream.py:121-123,270-276):
//
//     z = d * 1e-3 ;  ray = A * [u - cx, v - cy, 1] ,  A = R * diag(1/fx, 1/fy, 1)
//     p = z * ray + t      (d == 0 -> p = 0, mask = 0)        evaluated in double, stored as float
//
// HBM-bound streaming kernel: 2 B/px in, 12 + 1 B/px out.  Each thread owns 8 consecutive pixels
// (one 16-byte depth load, one 8-byte mask store); the 96 B of xyz per thread are transposed
// through a warp-private shared-memory buffer so every global store instruction writes 512
// contiguous bytes; counts are reduced with redux.sync and one atomic per warp.
#include "ti_common.cuh"
#include "ti_register.cuh"

namespace ti {

constexpr int BP_THREADS = 256;
constexpr int BP_PX_PER_THREAD = 8;
constexpr int BP_TILE = BP_THREADS * BP_PX_PER_THREAD;  // 2048 px
constexpr int MAX_BP_JOBS = 16;
constexpr int BP_LANE_STRIDE = 7;  // float4 slots per lane in the transpose buffer (6 used + 1 pad: conflict-free)

// Constants in double: p = d * (au * u + av * v + ac) + t with d the raw depth in millimetres - the 1e-3, the
// intrinsics and the principal point are folded in on the host (ti_upload_projection).
struct BpCam {
    double au[3], av[3], ac[3];
    double t[3];
    int width, height;
};

struct BpJobDev {
    const uint16_t* depth;
    float* xyz;
    uint8_t* mask;
    uint32_t* count;
    uint64_t depth_stride, xyz_stride, mask_stride;
    BpCam cam;
    uint32_t tile_begin;  // prefix of tiles per frame set
    // fused colour lookup (backproject_vec_kernel<true> only)
    const uint8_t* rgb;
    uint8_t* colour;
    uint64_t rgb_stride, colour_stride;
    RegConst reg;
};

struct BpParams {
    BpJobDev job[MAX_BP_JOBS];
    uint32_t tiles_per_set;
    int n_jobs;
    int n_batch;
};

// The arithmetic is done in double and rounded to float once per coordinate.  In float the error of a point is about
// 6e-8 of the LARGEST term, z * ray or t, whatever the size of their sum - a point that lands within a few millimetres of
// the body origin (z * ray = -t) then misses the 1e-5 relative bar (found by tools/fuzz_convert_backproject.py: 1.08e-5).
// The kernel moves 15 B/px and stays HBM-bound: ten double-precision operations per pixel are a fraction of the FP64 pipe.
__device__ __forceinline__ void project(const BpCam& c, double bx, double by, double bz, double ud, uint32_t d, float& x,
                                        float& y, float& z) {
    const double dd = (double)d;
    const bool ok = d != 0;
    x = ok ? (float)fma(dd, fma(c.au[0], ud, bx), c.t[0]) : 0.f;
    y = ok ? (float)fma(dd, fma(c.au[1], ud, by), c.t[1]) : 0.f;
    z = ok ? (float)fma(dd, fma(c.au[2], ud, bz), c.t[2]) : 0.f;
}

template <bool COLOUR>
__global__ void __launch_bounds__(BP_THREADS, 4) backproject_vec_kernel(const __grid_constant__ BpParams P) {
    __shared__ float4 xbuf[BP_THREADS / 32][32 * BP_LANE_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Every CTA owns a CONTIGUOUS run of tiles (decoded incrementally: no divisions per 8 pixels), so a
    // warp sees at most a couple of (frame, stream) changes and can keep the valid count in a register:
    // one atomic per warp per frame instead of one per warp per tile (same-address atomics serialise in
    // L2: per-tile atomics alone cost as much as the whole kernel).  The depth vector of the NEXT tile is
    // requested before the current one is processed.
    struct Cur { int j; uint32_t b, p0, warp_p0, npx; bool live; uint4 dv; };
    const uint64_t total = (uint64_t)P.tiles_per_set * P.n_batch;
    const uint64_t chunk = (total + gridDim.x - 1) / gridDim.x;
    const uint64_t t_begin = (uint64_t)blockIdx.x * chunk;
    uint64_t left = t_begin < total ? (total - t_begin < chunk ? total - t_begin : chunk) : 0;  // tiles still to fetch
    uint32_t b = (uint32_t)(t_begin / P.tiles_per_set);
    uint32_t r = (uint32_t)(t_begin - (uint64_t)b * P.tiles_per_set);
    int j = 0;
    auto fetch = [&](Cur& c) {
        c.live = left > 0;
        c.dv = make_uint4(0u, 0u, 0u, 0u);
        if (!c.live) return;
        --left;
        if (r >= P.tiles_per_set) { r = 0; ++b; j = 0; }
        while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
        const BpJobDev& J = P.job[j];
        c.j = j; c.b = b;
        c.npx = (uint32_t)J.cam.width * J.cam.height;
        c.warp_p0 = (r - J.tile_begin) * BP_TILE + warp * (32 * BP_PX_PER_THREAD);
        c.p0 = c.warp_p0 + lane * BP_PX_PER_THREAD;
        if (c.p0 < c.npx) c.dv = ld_stream_u4(J.depth + (uint64_t)b * (J.depth_stride / 2) + c.p0);
        ++r;
    };
    Cur cur, nxt;
    fetch(cur);
    uint32_t acc = 0, acc_b = cur.b;  // this lane's running valid count for (acc_j, acc_b)
    int acc_j = cur.j;
    auto flush = [&]() {
        const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, acc);
        if (lane == 0 && wsum && P.job[acc_j].count) atomicAdd(P.job[acc_j].count + acc_b, wsum);
        acc = 0;
    };
    while (cur.live) {
        fetch(nxt);
        const BpJobDev& J = P.job[cur.j];
        const uint32_t p0 = cur.p0, npx = cur.npx, warp_p0 = cur.warp_p0;
        uint32_t nvalid = 0;
        if (p0 < npx) {  // width % 8 == 0 -> a thread's 8 pixels never straddle a row or the frame end
            const uint32_t dw[4] = {cur.dv.x, cur.dv.y, cur.dv.z, cur.dv.w};
            const int v = (int)(p0 / (uint32_t)J.cam.width), u0 = (int)(p0 - (uint32_t)v * J.cam.width);
            const double vd = (double)v;
            const double bx = fma(J.cam.av[0], vd, J.cam.ac[0]);
            const double by = fma(J.cam.av[1], vd, J.cam.ac[1]);
            const double bz = fma(J.cam.av[2], vd, J.cam.ac[2]);
            const double u0d = (double)u0;
            float f[24];
            uint32_t m0 = 0, m1 = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t d = (dw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                project(J.cam, bx, by, bz, u0d + (double)k, d, f[3 * k], f[3 * k + 1], f[3 * k + 2]);
                const uint32_t ok = d != 0 ? 1u : 0u;
                nvalid += ok;
                if (k < 4) m0 |= ok << (8 * k); else m1 |= ok << (8 * (k - 4));
            }
            if (J.mask) st_stream_u2(J.mask + (uint64_t)cur.b * J.mask_stride + p0, make_uint2(m0, m1));
            if (COLOUR) {  // 8 pixels x RGB = 24 bytes = three 64-bit stores (p0 is a multiple of 8)
                const uint8_t* rgb = J.rgb + (uint64_t)cur.b * J.rgb_stride;
                int idx[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) idx[k] = reg_pixel_index(J.reg, u0 + k, v, (dw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
                uint32_t c[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {  // the gathers of all 8 pixels are in flight together (no control flow between them)
                    const uint8_t* s = rgb + (size_t)max(idx[k], 0) * 3;
                    const uint32_t px = (uint32_t)__ldg(s) | ((uint32_t)__ldg(s + 1) << 8) | ((uint32_t)__ldg(s + 2) << 16);
                    c[k] = idx[k] >= 0 ? px : 0u;
                }
                uint8_t* o = J.colour + (uint64_t)cur.b * J.colour_stride + (size_t)p0 * 3;
                st_stream_u2(o, make_uint2(c[0] | (c[1] << 24), (c[1] >> 8) | (c[2] << 16)));
                st_stream_u2(o + 8, make_uint2((c[2] >> 16) | (c[3] << 8), c[4] | (c[5] << 24)));
                st_stream_u2(o + 16, make_uint2((c[5] >> 8) | (c[6] << 16), (c[6] >> 16) | (c[7] << 8)));
            }
#pragma unroll
            for (int q = 0; q < 6; ++q)
                xbuf[warp][lane * BP_LANE_STRIDE + q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        }
        __syncwarp();
        // coalesced write-out: float4 slot q of the warp's 256-pixel span sits at lane (q / 6), sub-slot (q % 6)
        if (warp_p0 < npx) {
            const uint32_t live_px = min((uint32_t)(32 * BP_PX_PER_THREAD), npx - warp_p0);
            const uint32_t live_q = live_px * 3 / 4;  // live_px is a multiple of 8
            float4* out = reinterpret_cast<float4*>(J.xyz + (uint64_t)cur.b * (J.xyz_stride / 4) + (uint64_t)warp_p0 * 3);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const uint32_t q = i * 32 + lane;
                if (q < live_q) {
                    const float4 val = xbuf[warp][(q / 6) * BP_LANE_STRIDE + (q % 6)];
                    st_stream_u4(out + q, make_uint4(__float_as_uint(val.x), __float_as_uint(val.y),
                                                     __float_as_uint(val.z), __float_as_uint(val.w)));
                }
            }
        }
        __syncwarp();
        if (cur.j != acc_j || cur.b != acc_b) {  // warp-uniform
            flush();
            acc_j = cur.j; acc_b = cur.b;
        }
        acc += nvalid;
        cur = nxt;
    }
    flush();
}

// any width / alignment: one pixel per thread
__global__ void __launch_bounds__(BP_THREADS) backproject_scalar_kernel(BpJobDev J, int n_batch) {
    const uint32_t npx = (uint32_t)J.cam.width * J.cam.height;
    const uint64_t total = (uint64_t)npx * n_batch;
    // every thread takes part in every warp reduction, so the loop bound is uniform per warp
    const uint64_t rounds = (total + (uint64_t)gridDim.x * BP_THREADS - 1) / ((uint64_t)gridDim.x * BP_THREADS);
    for (uint64_t it = 0; it < rounds; ++it) {
        const uint64_t t = it * gridDim.x * BP_THREADS + (uint64_t)blockIdx.x * BP_THREADS + threadIdx.x;
        uint32_t ok = 0, b = 0;
        if (t < total) {
            b = (uint32_t)(t / npx);
            const uint32_t p = (uint32_t)(t - (uint64_t)b * npx);
            const int v = (int)(p / (uint32_t)J.cam.width), u = (int)(p - (uint32_t)v * J.cam.width);
            const uint32_t d = J.depth[(uint64_t)b * (J.depth_stride / 2) + p];
            const double vd = (double)v;
            float x, y, z;
            project(J.cam, fma(J.cam.av[0], vd, J.cam.ac[0]), fma(J.cam.av[1], vd, J.cam.ac[1]), fma(J.cam.av[2], vd, J.cam.ac[2]),
                    (double)u, d, x, y, z);
            float* o = J.xyz + (uint64_t)b * (J.xyz_stride / 4) + (uint64_t)p * 3;
            o[0] = x; o[1] = y; o[2] = z;
            ok = d != 0;
            if (J.mask) J.mask[(uint64_t)b * J.mask_stride + p] = (uint8_t)ok;
        }
        if (J.count && ok) atomicAdd(J.count + b, 1u);
    }
}

// valid count / min / max / sum of a depth frame (examples/rgbd_stream.py:270-276 prints count, mean, min, max of depth > 0):
// out[b] = {count, min, max, 0, sum lo, sum hi} as six u32; 128-bit loads, warp reductions, one set of atomics per warp.
__global__ void __launch_bounds__(256) depth_stats_kernel(const uint16_t* __restrict__ depth, uint64_t stride_bytes, uint32_t npx, int n_batch,
                                                          uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int b = blockIdx.y; b < n_batch; b += gridDim.y) {
        const uint16_t* d = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(depth) + (uint64_t)b * stride_bytes);
        uint32_t cnt = 0, mn = 0xFFFFFFFFu, mx = 0;
        uint64_t sum = 0;
        for (uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 8; i < npx; i += gridDim.x * 256 * 8) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if (i + 8 <= npx && (((uintptr_t)(d + i)) & 15) == 0) {
                const uint4 q = ld_stream_u4(d + i);
                w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            } else {
                for (uint32_t k = 0; k < 8 && i + k < npx; ++k) w[k >> 1] |= (uint32_t)d[i + k] << ((k & 1) * 16);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t v = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                if (v) { ++cnt; sum += v; mn = min(mn, v); mx = max(mx, v); }
            }
        }
#ifdef TI_EMULATE
        uint32_t wc = __reduce_add_sync(0xFFFFFFFFu, cnt), wlo = 0, whi = 0, wmn = mn, wmx = mx;
        for (int o = 16; o; o >>= 1) { wmn = min(wmn, __shfl_xor_sync(0xFFFFFFFFu, wmn, o)); wmx = max(wmx, __shfl_xor_sync(0xFFFFFFFFu, wmx, o)); }
        uint64_t ws = 0;
        for (int l = 0; l < 32; ++l) ws += ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)(sum >> 32), l) << 32) | __shfl_sync(0xFFFFFFFFu, (uint32_t)sum, l);
        wlo = (uint32_t)ws; whi = (uint32_t)(ws >> 32);
#else
        const uint32_t wc = __reduce_add_sync(0xFFFFFFFFu, cnt), wmn = __reduce_min_sync(0xFFFFFFFFu, mn), wmx = __reduce_max_sync(0xFFFFFFFFu, mx);
        uint64_t ws = sum;
        for (int o = 16; o; o >>= 1) ws += __shfl_xor_sync(0xFFFFFFFFu, ws, o);
        const uint32_t wlo = (uint32_t)ws, whi = (uint32_t)(ws >> 32);
#endif
        if (lane == 0 && wc) {
            uint32_t* o = out + (size_t)b * 6;
            atomicAdd(o, wc);
            atomicMin(o + 1, wmn);
            atomicMax(o + 2, wmx);
#ifdef TI_EMULATE
            uint64_t* s64 = reinterpret_cast<uint64_t*>(o + 4);
            __atomic_fetch_add(s64, ((uint64_t)whi << 32) | wlo, __ATOMIC_RELAXED);
#else
            atomicAdd(reinterpret_cast<unsigned long long*>(o + 4), ((unsigned long long)whi << 32) | wlo);
#endif
        }
    }
}

int launch_depth_stats(ti_ctx* ctx, const uint16_t* depth, int width, int height, int n_batch, uint64_t stride, uint32_t* out) {
    if (n_batch <= 0) return TI_OK;
    const uint32_t npx = (uint32_t)width * (uint32_t)height;
    // {0, 0xFFFFFFFF, 0, 0, 0, 0} per frame: min starts at all ones (an empty frame is reported as count 0, min 0 by the caller's view below)
    std::vector<uint32_t> init((size_t)n_batch * 6, 0u);
    for (int b = 0; b < n_batch; ++b) init[(size_t)b * 6 + 1] = 0xFFFFFFFFu;
    TI_CUDA(ctx, cudaMemcpyAsync(out, init.data(), init.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
#ifndef TI_EMULATE
    TI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `init` is pageable host memory: the copy must have left it
#endif
    const int gx = (int)std::min<uint32_t>((npx / 8 + 255) / 256 + 1, (uint32_t)ctx->sm_count * 4);
    dim3 grid((unsigned)gx, (unsigned)std::min(n_batch, 64));
    TI_LAUNCH(depth_stats_kernel, grid, 256, 0, ctx->stream, depth, stride ? stride : (uint64_t)npx * 2, npx, n_batch, out);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

int launch_backproject(ti_ctx* ctx, const BackprojectJob* jobs, int n_jobs, int n_batch) {
    if (n_jobs <= 0 || n_batch <= 0) return TI_OK;
    // two passes over the job list: plain jobs share launches of <false>, jobs with a colour output share launches of <true>
    for (int want_colour = 0; want_colour < 2; ++want_colour) {
        int done = 0;
        while (done < n_jobs) {
            BpParams P{};
            int nv = 0;
            uint32_t tiles = 0;
            for (; done < n_jobs && nv < MAX_BP_JOBS; ++done) {
                const BackprojectJob& J = jobs[done];
                if ((J.colour != nullptr) != (want_colour != 0)) continue;
                if (J.camera < 0 || J.camera >= TI_MAX_CAMERAS || !ctx->cams[J.camera].has_proj)
                    return fail(ctx, TI_ESTATE, "backproject: camera slot %d has no projection (call ti_upload_projection)", J.camera);
                if (!J.depth || !J.xyz) return fail(ctx, TI_EINVAL, "backproject: null depth/xyz pointer");
                const CameraSlot& C = ctx->cams[J.camera];
                BpJobDev D{};
                D.depth = J.depth; D.xyz = J.xyz; D.mask = J.mask; D.count = J.count;
                D.depth_stride = J.depth_stride; D.xyz_stride = J.xyz_stride; D.mask_stride = J.mask_stride;
                for (int i = 0; i < 3; ++i) { D.cam.au[i] = C.proj_au[i]; D.cam.av[i] = C.proj_av[i]; D.cam.ac[i] = C.proj_ac[i]; D.cam.t[i] = C.proj_t[i]; }
                D.cam.width = C.proj_w; D.cam.height = C.proj_h;
                if (J.depth_stride % 2 || J.xyz_stride % 4)
                    return fail(ctx, TI_EINVAL, "backproject: frame strides must keep element alignment");
                if (J.colour) {
                    if (!C.has_reg) return fail(ctx, TI_ESTATE, "backproject: camera slot %d has no registration (call ti_upload_registration)", J.camera);
                    if (!J.rgb) return fail(ctx, TI_EINVAL, "backproject: a colour output needs the RGB image");
                    if (C.reg_dw != C.proj_w || C.reg_dh != C.proj_h)
                        return fail(ctx, TI_EINVAL, "backproject: registration (%dx%d) and projection (%dx%d) of slot %d describe different depth images",
                                    C.reg_dw, C.reg_dh, C.proj_w, C.proj_h, J.camera);
                    D.rgb = J.rgb; D.colour = J.colour;
                    D.rgb_stride = J.rgb_stride ? J.rgb_stride : (uint64_t)C.reg_rw * C.reg_rh * 3;
                    D.colour_stride = J.colour_stride ? J.colour_stride : (uint64_t)C.proj_w * C.proj_h * 3;
                    D.reg = reg_constants(C);
                }
                if (J.count) {
#ifndef TI_EMULATE
                    TI_CUDA(ctx, cudaMemsetAsync(J.count, 0, sizeof(uint32_t) * (size_t)n_batch, ctx->stream));
#else
                    for (int b = 0; b < n_batch; ++b) J.count[b] = 0;
#endif
                }
                const bool aligned = (C.proj_w % 8 == 0) && ((uintptr_t)J.depth % 16 == 0) && ((uintptr_t)J.xyz % 16 == 0) &&
                                     (J.depth_stride % 16 == 0) && (J.xyz_stride % 16 == 0) &&
                                     (!J.mask || (((uintptr_t)J.mask | J.mask_stride) % 8 == 0)) &&
                                     (!J.colour || (((uintptr_t)J.colour | D.colour_stride) % 8 == 0));
                if (!aligned) {
                    const uint64_t total = (uint64_t)C.proj_w * C.proj_h * n_batch;
                    const int grid = (int)std::min<uint64_t>((total + BP_THREADS - 1) / BP_THREADS, (uint64_t)ctx->sm_count * 8);
                    TI_LAUNCH(backproject_scalar_kernel, grid, BP_THREADS, 0, ctx->stream, D, n_batch);
                    TI_CHECK_LAUNCH(ctx);
                    if (J.colour) {  // any size / alignment: the stand-alone registration kernel after it
                        const int rc = launch_register_colour(ctx, J.camera, J.depth, J.rgb, J.colour, n_batch, J.depth_stride, J.rgb_stride, J.colour_stride);
                        if (rc != TI_OK) return rc;
                    }
                    continue;
                }
                D.tile_begin = tiles;
                tiles += (uint32_t)(((uint64_t)C.proj_w * C.proj_h + BP_TILE - 1) / BP_TILE);
                P.job[nv++] = D;
            }
            if (nv == 0) continue;
            P.tiles_per_set = tiles;
            P.n_jobs = nv;
            P.n_batch = n_batch;
            const uint64_t total = (uint64_t)tiles * n_batch;
            // two resident CTAs per SM measured best for the plain kernel (fewer concurrent DRAM streams; the register prefetch hides
            // latency); the colour kernel waits on gathers and takes what fits
            const int fit = want_colour ? resident_ctas(backproject_vec_kernel<true>, BP_THREADS, 0, 3)
                                        : std::min(2, resident_ctas(backproject_vec_kernel<false>, BP_THREADS, 0, 4));
            const int grid = (int)std::min<uint64_t>(total, (uint64_t)ctx->sm_count * (ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : fit));
            if (want_colour) TI_LAUNCH(backproject_vec_kernel<true>, grid, BP_THREADS, 0, ctx->stream, P);
            else TI_LAUNCH(backproject_vec_kernel<false>, grid, BP_THREADS, 0, ctx->stream, P);
            TI_CHECK_LAUNCH(ctx);
        }
    }
    return TI_OK;
}

}  // namespace ti
