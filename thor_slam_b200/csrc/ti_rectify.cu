// Format conversion fused with the undistort / rectify remap (u8, bit-exact with
// cv2.remap(cvtColor(src), mapx, mapy, INTER_LINEAR, BORDER_CONSTANT 0)).
//
// The reference publishes raw images and lets cuVSLAM undistort them
// (thor_slam/slam/adapters/isaac_ros.py:364-411, Makefile:77-80); here the remap runs on the GPU.
//
// Tiled kernel (rectify_tile_kernel): one CTA = RT_W x RT_H (128 x 16) output pixels.
//   1. every thread fetches the packed LUT entries of its 8 consecutive output pixels (2 x 128-bit),
//   2. the CTA copies the tile's source bounding box (precomputed per tile at upload time) into
//      shared memory with coalesced 128-bit loads, zero-filling what lies outside the image
//      (that is BORDER_CONSTANT 0),
//   3. four bilinear taps per pixel are read from shared memory, blended in exact integer
//      arithmetic, packed and written with one 64-bit store per channel-row.
// Direct kernel (rectify_direct_kernel): one thread per output pixel, taps straight from global
// memory, conversion applied per tap.  Used for sizes/alignments the tiled kernel does not take
// and for the two-stage conversions (BGR->gray, NV12->rgb) before remap.
#include "ti_common.cuh"
#include "ti_pixel.cuh"
#include "ti_tma.cuh"
#include "ti_rectify_pair.cuh"
#include <deque>

namespace ti {

constexpr int MAX_RECT_JOBS = 16;
constexpr int RT_PX_PER_THREAD = 8;
constexpr int RT_MAX_SMEM = 96 * 1024;  // per-CTA cap for the staged source box

struct RectJobDev {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t src_stride, dst_stride;
    const lut_t* lut;  // padded: lut_rows x lut_pitch
    const TileBox* boxes;
    int lut_pitch;
    int tiles_x, tiles_y;
    int dst_w, dst_h, src_w, src_h;
    uint32_t tile_begin;
};

struct RectParams {
    RectJobDev job[MAX_RECT_JOBS];
    uint32_t tiles_per_set;
    int n_jobs;
    int n_batch;
};

__device__ __forceinline__ uint4 ld_src_u4(const void* p) {
#ifdef TI_EMULATE
    return *reinterpret_cast<const uint4*>(ti_emu::check_align(p, 16));
#else
    uint4 r;  // source rows are re-read by neighbouring tiles (halo): default L2 policy, no L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
#endif
}

// C = channels of the staged source (1: MONO8 / NV12 luma, 3: BGR8 written out as RGB8)
template <int C>
__global__ void __launch_bounds__(RT_THREADS) rectify_tile_kernel(const __grid_constant__ RectParams P) {
    TI_DYNAMIC_SMEM(uint8_t, smem);
    const int tid = threadIdx.x;
    const int lx = (tid & 15) * RT_PX_PER_THREAD, ly = tid >> 4;
    const uint64_t total = (uint64_t)P.tiles_per_set * P.n_batch;
    for (uint64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const uint32_t b = (uint32_t)(t / P.tiles_per_set);
        const uint32_t r = (uint32_t)(t - (uint64_t)b * P.tiles_per_set);
        int j = 0;
        while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
        const RectJobDev& J = P.job[j];
        const uint32_t tile = r - J.tile_begin;
        const int ty = (int)(tile / (uint32_t)J.tiles_x), tx = (int)(tile - (uint32_t)ty * J.tiles_x);
        const TileBox box = J.boxes[tile];
        const int u = tx * RT_W + lx, v = ty * RT_H + ly;

        // (1) LUT entries of this thread's 8 pixels (LUT is padded to whole tiles: no guards)
        const lut_t* lp = J.lut + (size_t)v * J.lut_pitch + u;
        lut_t e[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // 64-bit entries: two per 128-bit load
            const uint4 l = ld_keep_u4(lp + 2 * q);
            e[2 * q] = (lut_t)l.x | ((lut_t)l.y << 32);
            e[2 * q + 1] = (lut_t)l.z | ((lut_t)l.w << 32);
        }

        // (2) stage the source box
        const int c0 = (box.x0 * C) & ~15;               // first staged byte column (may be negative)
        const int c1 = (box.x1 * C + 15) & ~15;          // one past the last staged byte column
        const int pitch = c1 - c0;
        const int rows = box.y1 - box.y0;
        const bool any = box.x1 > box.x0;
        if (any) {
            const uint8_t* src = J.src + (uint64_t)b * J.src_stride;
            const int row_bytes = J.src_w * C;
            const int nvec = pitch >> 4;
            const int total_vec = nvec * rows;
            for (int i = tid; i < total_vec; i += RT_THREADS) {
                const int row = i / nvec, vc = i - row * nvec;
                const int gy = box.y0 + row, gx = c0 + (vc << 4);
                uint4 val = make_uint4(0u, 0u, 0u, 0u);
                if (gy >= 0 && gy < J.src_h && gx >= 0 && gx + 16 <= row_bytes)
                    val = ld_src_u4(src + (size_t)gy * row_bytes + gx);
                *reinterpret_cast<uint4*>(smem + row * pitch + (vc << 4)) = val;
            }
        }
        __syncthreads();

        // (3) taps + blend
        if (v < J.dst_h && u < J.dst_w) {
            uint32_t o[C][2];
#pragma unroll
            for (int c = 0; c < C; ++c) o[c][0] = o[c][1] = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const lut_t ek = e[k];
                if (ek == LUT_OUTSIDE) continue;
                const int sx = lut_x0(ek), sy = lut_y0(ek);
                const uint32_t fx = lut_fx(ek), fy = lut_fy(ek);
                const uint8_t* s = smem + (sy - box.y0) * pitch + (sx * C - c0);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int sc = C == 3 ? 2 - c : c;  // BGR source -> RGB output
                    const uint32_t val = bilinear_u8(s[sc], s[sc + C], s[sc + pitch], s[sc + pitch + C], fx, fy);
                    o[c][k >> 2] |= val << ((k & 3) * 8);
                }
            }
            uint8_t* dst = J.dst + (uint64_t)b * J.dst_stride + ((size_t)v * J.dst_w + u) * C;
            if (u + 8 <= J.dst_w) {
                if (C == 1) {
                    st_stream_u2(dst, make_uint2(o[0][0], o[0][1]));
                } else {
                    // interleave 8 px x 3 channels = 24 bytes = 3 x 8-byte stores
                    uint32_t w[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) w[i] = 0u;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const uint32_t val = (o[c][k >> 2] >> ((k & 3) * 8)) & 0xFFu;
                            const int bi = k * 3 + c;
                            w[bi >> 2] |= val << ((bi & 3) * 8);
                        }
                    st_stream_u2(dst, make_uint2(w[0], w[1]));
                    st_stream_u2(dst + 8, make_uint2(w[2], w[3]));
                    st_stream_u2(dst + 16, make_uint2(w[4], w[5]));
                }
            } else {  // ragged right edge (dst_w % 8 != 0 is rejected by the launcher; kept for safety)
                for (int k = 0; k < 8 && u + k < J.dst_w; ++k)
                    for (int c = 0; c < C; ++c) dst[k * C + c] = (uint8_t)((o[c][k >> 2] >> ((k & 3) * 8)) & 0xFFu);
            }
        }
        __syncthreads();  // the next tile's staging overwrites smem
    }
}

// ---- fast mono kernel ("v2") --------------------------------------------------------------------
// One CTA = 128 x 32 output pixels; a warp owns 4 rows; in a row, lane L owns pixels L, L+32, L+64, L+96.
//  * LUT entries are tile-relative: bits 31..16 = shared-memory byte address of the top tap pair,
//    bits 10..6 = fy, bits 4..0 = fx (built at upload, see ti_upload_rectify_map).
//  * The source box is staged twice per row: copy A as is, copy B shifted by one byte, so that the
//    pair (p[x0], p[x0+1]) is always one 2-byte-aligned LDS.U16 whatever the parity of x0; the pair
//    one source row below is +M2_ROW_BYTES (an immediate).  The 32 lanes of a warp read ~32
//    consecutive source bytes (<= 9 words of copy A, <= 9 words of copy B, the copies 16 banks
//    apart): no bank conflicts; stores are one byte per lane, 32 contiguous bytes per warp.
//  * blend: top/bot = dp2a((32-fx, fx), pair); out = ((32-fy)*top + fy*bot + 512) >> 10, evaluated
//    as byte 2 of 64*(fy*(bot-top) + 32*top + 512) so that no shift is needed before packing.
struct Rect2JobDev {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t src_stride, dst_stride;
    const uint32_t* lut2;
    const TileBox2* boxes2;
    int tiles_x, n_tiles;
    int dst_w, dst_h, src_w, src_h;
    uint32_t tile_begin;
};

struct Rect2Params {
    Rect2JobDev job[MAX_RECT_JOBS];
    uint32_t tiles_per_set;
    int n_jobs;
    int n_batch;
};

__device__ __forceinline__ uint32_t lds_u16(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }

// Taps addressed in the 32-bit shared window: base + (entry >> 16) is ONE LEA.HI, the row below an immediate.
#ifdef TI_EMULATE
typedef const uint8_t* smem_base_t;
__device__ __forceinline__ smem_base_t smem_base(const uint8_t* p) { return p; }
template <int OFF>
__device__ __forceinline__ uint32_t lds_u16_at(smem_base_t base, uint32_t e) { return lds_u16(base + (e >> 16) + OFF); }
#else
typedef uint32_t smem_base_t;
__device__ __forceinline__ smem_base_t smem_base(const uint8_t* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ uint32_t lds_u16_at(smem_base_t base, uint32_t e) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(base + (e >> 16)), "n"(OFF));
    return v;
}
#endif

__device__ __forceinline__ uint32_t ld_src_u32(const void* p) {
#ifdef TI_EMULATE
    return *reinterpret_cast<const uint32_t*>(ti_emu::check_align(p, 4));
#else
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
#endif
}

template <int ROW_PITCH = M3_PITCH>
__device__ __forceinline__ uint32_t blend64(uint32_t e, smem_base_t smem) {
    const uint32_t pt = lds_u16_at<0>(smem, e), pb = lds_u16_at<ROW_PITCH>(smem, e);
    const uint32_t fx = e & 31u;
    const uint32_t aw = fx * 65535u + 32u;  // (32 - fx) | fx << 16
    const uint32_t top = __dp2a_lo(aw, pt, 0u), bot = __dp2a_lo(aw, pb, 0u);
    const uint32_t g = e & (31u << 6);      // 64 * fy
    return g * (bot - top) + (top * 2048u + 32768u);  // 64 * S; result pixel = byte 2
}

__global__ void __launch_bounds__(M2_THREADS, 6) rectify_mono_kernel(const __grid_constant__ Rect2Params P) {
    TI_DYNAMIC_SMEM(uint8_t, smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < M2_ZERO_BYTES / 16) reinterpret_cast<uint4*>(smem)[tid] = make_uint4(0u, 0u, 0u, 0u);
    // work item = (frame b, job j, tile): walked with a stride of gridDim.x, decoded incrementally
    // (no divisions: every per-tile scalar instruction costs 1/16 instruction per pixel)
    uint32_t r = blockIdx.x, b = 0;
    int j = 0;
    const int vc = tid & 15, row0 = tid >> 4;  // staging role: 16-byte column vc of rows row0, row0+16, ...
    while (true) {
        while (r >= P.tiles_per_set) { r -= P.tiles_per_set; ++b; j = 0; }
        if (b >= (uint32_t)P.n_batch) break;
        while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
        const Rect2JobDev& J = P.job[j];
        const uint32_t tile = r - J.tile_begin;
        TileBox2 box;
        {
            const uint4 raw = *reinterpret_cast<const uint4*>(J.boxes2 + tile);
            box.c0 = (int16_t)(raw.x & 0xFFFF); box.y0 = (int16_t)(raw.x >> 16);
            box.nvec = (int16_t)(raw.y & 0xFFFF); box.rows = (int16_t)(raw.y >> 16);
            box.u0 = (int16_t)(raw.z & 0xFFFF); box.v0 = (int16_t)(raw.z >> 16);
        }

        // (1) LUT: 4 rows x 4 px per thread, issued before the staging loads
        const uint32_t* lp = J.lut2 + (size_t)tile * (M2_TW * M2_TH) + (warp * 4) * M2_TW + lane * 4;
        uint4 l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) l[q] = ld_keep_u4(lp + q * M2_TW);

        // (2) stage rows [y0, y0 + rows) x columns [c0, c0 + 16 * nvec): copy A and copy B
        if (vc < box.nvec) {
            const int gx = box.c0 + (vc << 4);
            const bool col_in = gx >= 0 && gx < J.src_w;   // gx is a multiple of 16 and src_w % 16 == 0
            const bool col_left = gx == -16;                // copy B of the left-border vector needs p[0]
            const bool has_next = gx + 16 < J.src_w;
            const uint8_t* g = J.src + (uint64_t)b * J.src_stride + (int64_t)(box.y0 + row0) * J.src_w + gx;
            uint8_t* d = smem + M2_ZERO_BYTES + row0 * M2_ROW_BYTES + (vc << 4);
            for (int row = row0; row < box.rows; row += 16, g += (size_t)16 * J.src_w, d += 16 * M2_ROW_BYTES) {
                const int gy = box.y0 + row;
                const bool row_in = gy >= 0 && gy < J.src_h;
                uint4 a = make_uint4(0u, 0u, 0u, 0u);
                uint32_t nx = 0u;
                if (row_in && col_in) {
                    a = ld_src_u4(g);
                    if (has_next) nx = ld_src_u32(g + 16);
                } else if (row_in && col_left) {
                    nx = ld_src_u32(g + 16);
                }
                *reinterpret_cast<uint4*>(d) = a;
                *reinterpret_cast<uint4*>(d + M2_COPY_BYTES) =
                    make_uint4(__funnelshift_r(a.x, a.y, 8), __funnelshift_r(a.y, a.z, 8), __funnelshift_r(a.z, a.w, 8),
                               __funnelshift_r(a.w, nx, 8));
            }
        }
        __syncthreads();

        // (3) taps + blend + packed store
        const smem_base_t sm2 = smem_base(smem);
        const int u = box.u0 + lane, v = box.v0 + warp * 4;
        uint8_t* dp = J.dst + (uint64_t)b * J.dst_stride + (size_t)v * J.dst_w + u;
        const int live_rows = J.dst_h - v;            // rows of this warp that exist in the image
        const int live_cols = (J.dst_w - u + 31) >> 5;  // of this lane's 4 pixels, how many exist
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t s0 = blend64<M2_ROW_BYTES>(l[q].x, sm2), s1 = blend64<M2_ROW_BYTES>(l[q].y, sm2);
            const uint32_t s2 = blend64<M2_ROW_BYTES>(l[q].z, sm2), s3 = blend64<M2_ROW_BYTES>(l[q].w, sm2);
            if (q < live_rows) {
                uint8_t* o = dp + (size_t)q * J.dst_w;
                if (live_cols > 0) st_stream_b8(o, s0 >> 16);
                if (live_cols > 1) st_stream_b8(o + 32, s1 >> 16);
                if (live_cols > 2) st_stream_b8(o + 64, s2 >> 16);
                if (live_cols > 3) st_stream_b8(o + 96, s3 >> 16);
            }
        }
        __syncthreads();
        r += gridDim.x;
    }
}

// ---- TMA-pipelined mono kernel ("v3") --------------------------------------------------------------
// Persistent, warp-specialised, M3_STAGES-deep shared-memory ring, no __syncthreads in steady state.
//
// Work unit = (stream, output tile, chunk of up to M3_FRAMES frames of the batch).  The tile's LUT
// (4 B per output pixel - twice the pixel payload) is fetched ONCE per unit into registers and reused
// for every frame of the chunk, so per frame only the source box streams in.  Work item = one frame
// of one unit = one ring stage.
//  * Warps 8..9 are producers.  Thread 0 of warp 8 issues, one item ahead, the TMA loads of copy A
//    (256 x 8 byte boxes; coordinates outside the image read as zero = BORDER_CONSTANT 0) onto the
//    stage's `raw` mbarrier and writes the stage header (destination pointer, live rows / columns,
//    LUT pointers).  TMA boxes must start on 16-byte boundaries, so the one-byte-shifted copy B cannot
//    come from TMA: the 64 producer threads build it from copy A in shared memory (LDS.128, funnel
//    shifts, STS.128) and arrive on the stage's `full` mbarrier.
//  * Warps 0..7 are consumers: wait `full`, blend exactly as the v2 kernel (aligned pair taps +
//    dp2a), store, arrive on `empty`.  They spend no instruction on staging, never wait on source
//    pixels from global memory, and prefetch the next unit's LUT during the last frame of a unit.
struct Rect3JobDev {
    const uint32_t* lut3;
    const TileBox2* boxes3;
    uint8_t* dst;
    uint64_t dst_stride;
    int dst_w, dst_h;
    int rows_alloc;
    uint32_t tile_begin;
};

struct Rect3Params {
    TiTensorMap map[MAX_RECT_JOBS];
    Rect3JobDev job[MAX_RECT_JOBS];
    uint32_t tiles_per_set;
    int n_jobs;
    int n_batch;
    int frames_per_unit;
    int rows_alloc_max;  // launch-wide stage geometry
    int stages;          // ring depth (3 .. M3_MAX_STAGES)
    int debug;           // bring-up switches (TI_OPT_DEBUG); 0 in production
};

// Stage header, 64 bytes at the end of a stage: written by the issuing thread, read by everybody.
//  h0 = {dst_lo, dst_hi, dst_w, live_rows}   h1 = {live_cols, rows, nvec, rows_alloc}
//  h2 = {lut_lo, lut_hi, next_lut_lo, next_lut_hi}   h3 = {flags, 0, 0, 0}
constexpr uint32_t H3_FIRST = 1u;  // first frame of a unit: (re)load the LUT registers
constexpr uint32_t H3_LAST_OF_UNIT = 2u;  // last frame of a unit: prefetch the next unit's LUT
constexpr uint32_t H3_LAST_ITEM = 4u;     // last item of this CTA
constexpr uint32_t H3_ALIGNED2 = 8u;      // destination tile rows start on even addresses (16-bit stores allowed)

template <int TH, bool PREFETCH>
__global__ void __launch_bounds__(M3_THREADS, PREFETCH ? 3 : 4) rectify_mono_tma_kernel(const __grid_constant__ Rect3Params P) {
    TI_DYNAMIC_SMEM(uint8_t, smem);
    constexpr int ROWS_PER_WARP = TH / M3_CONSUMER_WARPS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ab_bytes = 2u * (uint32_t)P.rows_alloc_max * M3_PITCH;
    const uint32_t stage_bytes = 128u + ab_bytes + 128u;
    const uint32_t hdr_off = 128u + ab_bytes;
    const int S = P.stages;
    uint64_t* raw = reinterpret_cast<uint64_t*>(smem);  // [S] TMA landed
    uint64_t* full = raw + M3_MAX_STAGES;                // [S] copy B ready
    uint64_t* empty = full + M3_MAX_STAGES;              // [S] consumers done
    uint8_t* stage0 = smem + 256;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(raw + s, 1);
            mbar_init(full + s, M3_PRODUCER_WARPS * 32);
            mbar_init(empty + s, M3_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    for (int s = 0; s < S; ++s)  // the always-zero block of every stage
        if (tid < 8) reinterpret_cast<uint4*>(stage0 + (size_t)s * stage_bytes)[tid] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    const uint32_t n_chunks = (uint32_t)((P.n_batch + P.frames_per_unit - 1) / P.frames_per_unit);
    const uint64_t total_units = (uint64_t)P.tiles_per_set * n_chunks;
    const uint32_t units_mine = total_units > blockIdx.x ? (uint32_t)((total_units - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    if (units_mine == 0) return;

    if (warp == M3_CONSUMER_WARPS + M3_PRODUCER_WARPS) {
        // ------------------------------------------------ issuer (one thread) ------------------------
        if (lane != 0) return;

        // issuer state (ptid == 0): the unit being issued, the unit after it (for LUT prefetch), item counter
        struct Unit { int j; uint32_t tile, b0, nb; uint4 box; uint64_t lut; };
        Unit cur{}, nxt{};
        uint32_t k_next = 0;   // index (among this CTA's units) of the unit held in `nxt`
        uint32_t f = 0;        // next frame of `cur` to issue
        uint32_t issued = 0;   // items issued so far
        int is = 0;            // stage of the next item to issue
        uint32_t iphase = 1;   // parity to wait for on empty[is]: 1 on a stage's first use (passes at once)
        bool more = true;      // `cur` is valid
        auto load_unit = [&](uint32_t k, Unit& U) {
            const uint64_t ug = (uint64_t)blockIdx.x + (uint64_t)k * gridDim.x;
            const uint32_t c = (uint32_t)(ug / P.tiles_per_set), r = (uint32_t)(ug - (uint64_t)c * P.tiles_per_set);
            int j = 0;
            while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
            U.j = j; U.tile = r - P.job[j].tile_begin;
            U.b0 = c * (uint32_t)P.frames_per_unit;
            U.nb = min((uint32_t)P.frames_per_unit, (uint32_t)P.n_batch - U.b0);
            U.box = *reinterpret_cast<const uint4*>(P.job[j].boxes3 + U.tile);
            U.lut = (uint64_t)(uintptr_t)(P.job[j].lut3 + (size_t)U.tile * (TH * M3_TW));
        };
        auto issue_one = [&]() {  // TMA copy A of frame f of `cur` into stage issued % S, then advance
            const Rect3JobDev& J = P.job[cur.j];
            const int c0 = (int16_t)(cur.box.x & 0xFFFF), y0 = (int16_t)(cur.box.x >> 16);
            const int nvec = (int16_t)(cur.box.y & 0xFFFF), rows = (int16_t)(cur.box.y >> 16);
            const int u0 = (int16_t)(cur.box.z & 0xFFFF), v0 = (int16_t)(cur.box.z >> 16);
            const uint32_t b = cur.b0 + f;
            const int s = is;
            uint8_t* sb = stage0 + (size_t)s * stage_bytes;
            mbar_wait_relaxed(empty + s, iphase);  // consumers have released the stage's previous item
            if (rows > 0) tma_load_3d(sb + 128, &P.map[cur.j], c0, y0, (int)b, raw + s);  // one box: 256 x rows_alloc bytes
            const bool last_of_unit = f + 1 == cur.nb;
            const bool have_next = k_next < units_mine;
            const uint64_t dst = (uint64_t)(uintptr_t)(J.dst + (uint64_t)b * J.dst_stride + (size_t)v0 * J.dst_w + u0);
            uint32_t flags = (f == 0 ? H3_FIRST : 0u) | (last_of_unit ? H3_LAST_OF_UNIT : 0u) |
                             (last_of_unit && !have_next ? H3_LAST_ITEM : 0u) |
                             (((dst | (uint64_t)J.dst_w) & 1ull) == 0 ? H3_ALIGNED2 : 0u);
            const uint64_t nlut = have_next ? nxt.lut : 0ull;
            uint4* h = reinterpret_cast<uint4*>(sb + hdr_off);
            h[0] = make_uint4((uint32_t)(dst & 0xFFFFFFFFu), (uint32_t)(dst >> 32), (uint32_t)J.dst_w, (uint32_t)(J.dst_h - v0));
            h[1] = make_uint4((uint32_t)(J.dst_w - u0), (uint32_t)rows, (uint32_t)nvec, (uint32_t)J.rows_alloc);
            h[2] = make_uint4((uint32_t)(cur.lut & 0xFFFFFFFFu), (uint32_t)(cur.lut >> 32), (uint32_t)(nlut & 0xFFFFFFFFu), (uint32_t)(nlut >> 32));
            h[3] = make_uint4(flags, 0u, 0u, 0u);
            mbar_arrive_expect_tx(raw + s, rows > 0 ? (uint32_t)J.rows_alloc * M3_PITCH : 0u);  // releases the header too
            ++issued;
            if (++is == S) { is = 0; iphase ^= 1u; }
            if (++f == cur.nb) {  // unit finished: move on
                f = 0;
                if (have_next) {
                    cur = nxt;
                    ++k_next;
                    if (k_next < units_mine) load_unit(k_next, nxt);
                } else {
                    more = false;
                }
            }
        };

        load_unit(0, cur);
        k_next = 1;
        if (units_mine > 1) load_unit(1, nxt);
        while (more) issue_one();  // runs ahead of the builders / consumers by up to S stages
        return;
    }
    if (warp >= M3_CONSUMER_WARPS) {
        // ------------------------------------------------ builder warps ------------------------------
        const int ptid = tid - M3_CONSUMER_WARPS * 32;  // 0 .. 32 * M3_PRODUCER_WARPS - 1
        const int vc = ptid & 15, r0 = ptid >> 4;       // column vector vc of rows r0, r0 + RSTEP, ...
        constexpr int RSTEP = M3_PRODUCER_WARPS * 2;
        int s = 0;
        uint32_t phase = 0;
        for (uint32_t i = 0;; ++i) {
            uint8_t* sb = stage0 + (size_t)s * stage_bytes;
            mbar_wait_relaxed(raw + s, phase);
            const uint4 h1 = *reinterpret_cast<const uint4*>(sb + hdr_off + 16);
            const uint32_t flags = *reinterpret_cast<const uint32_t*>(sb + hdr_off + 48);
            const int rows = (int)h1.y, nvec = (int)h1.z, rows_alloc = (int)h1.w;
            // copy B[64 + i] = A[i + 1]: odd-x0 pairs become 2-byte aligned, 16 banks away from copy A.
            // Every lane loads its 16-byte vector; the first word of the NEXT vector comes from the next
            // lane by shuffle (lanes 0..15 / 16..31 of a warp hold the 16 vectors of one row each).
            {
                const int nv = min(nvec, 12);
                const uint8_t* ap = sb + 128 + r0 * M3_PITCH + (vc << 4);
                uint8_t* bp = sb + 128 + (size_t)rows_alloc * M3_PITCH + 64 + r0 * M3_PITCH + (vc << 4);
                const int rows_pad = (rows + RSTEP - 1) / RSTEP * RSTEP;  // warp-uniform trip count (rows_alloc >= rows_pad)
                for (int row = r0; row < rows_pad; row += RSTEP, ap += RSTEP * M3_PITCH, bp += RSTEP * M3_PITCH) {
                    const uint4 x = *reinterpret_cast<const uint4*>(ap);
                    const uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, x.x, 1);
                    if (vc < nv && row < rows)
                        *reinterpret_cast<uint4*>(bp) = make_uint4(__funnelshift_r(x.x, x.y, 8), __funnelshift_r(x.y, x.z, 8),
                                                                   __funnelshift_r(x.z, x.w, 8), __funnelshift_r(x.w, nx, 8));
                }
            }
            mbar_arrive(full + s);  // every producer thread: each releases its own shared-memory writes
            if (flags & H3_LAST_ITEM) break;
            if (++s == S) { s = 0; phase ^= 1u; }
        }
        return;
    }
    // ---------------------------------------------------- consumers ---------------------------------
    uint4 l[ROWS_PER_WARP], ln[PREFETCH ? ROWS_PER_WARP : 1];  // LUT of the current unit / prefetched LUT of the next unit
#pragma unroll
    for (int q = 0; q < (PREFETCH ? ROWS_PER_WARP : 1); ++q) ln[q] = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t lut_lane_off = (uint32_t)((warp * ROWS_PER_WARP) * M3_TW + lane * 4);  // in u32 entries (lane's uint4)
    int s = 0;
    uint32_t phase = 0;
    for (uint32_t i = 0;; ++i) {
        const uint8_t* sb = stage0 + (size_t)s * stage_bytes;
        mbar_wait(full + s, phase);
        const uint4 h0 = *reinterpret_cast<const uint4*>(sb + hdr_off);
        const uint4 h2 = *reinterpret_cast<const uint4*>(sb + hdr_off + 32);
        const uint32_t live_cols = *reinterpret_cast<const uint32_t*>(sb + hdr_off + 16);
        const uint32_t flags = *reinterpret_cast<const uint32_t*>(sb + hdr_off + 48);
        if (flags & H3_FIRST) {
            if (i == 0 || !PREFETCH) {  // very first unit (or no prefetching): fetch now
                const uint32_t* lp = reinterpret_cast<const uint32_t*>((uintptr_t)(((uint64_t)h2.y << 32) | h2.x)) + lut_lane_off;
#pragma unroll
                for (int q = 0; q < ROWS_PER_WARP; ++q) l[q] = ld_keep_u4(lp + q * M3_TW);
            } else {
#pragma unroll
                for (int q = 0; q < ROWS_PER_WARP; ++q) l[q] = ln[PREFETCH ? q : 0];
            }
        }
        if (PREFETCH && (flags & H3_LAST_OF_UNIT) && (h2.z | h2.w)) {  // prefetch the next unit's LUT behind this frame's work
            const uint32_t* lp = reinterpret_cast<const uint32_t*>((uintptr_t)(((uint64_t)h2.w << 32) | h2.z)) + lut_lane_off;
#pragma unroll
            for (int q = 0; q < ROWS_PER_WARP; ++q) ln[q] = ld_keep_u4(lp + q * M3_TW);
        }
        uint8_t* dp = reinterpret_cast<uint8_t*>((uintptr_t)(((uint64_t)h0.y << 32) | h0.x));
        const smem_base_t sbase = smem_base(sb);
        const int dst_w = (int)h0.z, live_rows = (int)h0.w - warp * ROWS_PER_WARP;
        dp += (size_t)(warp * ROWS_PER_WARP) * dst_w + 2 * lane;
        // lane L owns pixels 2L, 2L+1, 64+2L, 64+2L+1 of the row: two 16-bit stores per row (64 B per warp store)
        if (live_rows >= ROWS_PER_WARP && live_cols >= (uint32_t)M3_TW && (flags & H3_ALIGNED2)) {  // warp-uniform
#pragma unroll
            for (int q = 0; q < ROWS_PER_WARP; ++q) {
                const uint32_t s0 = blend64(l[q].x, sbase), s1 = blend64(l[q].y, sbase), s2 = blend64(l[q].z, sbase), s3 = blend64(l[q].w, sbase);
                st_stream_b16(dp, __byte_perm(s0, s1, 0x0062));
                st_stream_b16(dp + 64, __byte_perm(s2, s3, 0x0062));
                dp += dst_w;
            }
        } else {
            const int c0 = 2 * lane, c2 = 64 + 2 * lane;
#pragma unroll
            for (int q = 0; q < ROWS_PER_WARP; ++q) {
                const uint32_t s0 = blend64(l[q].x, sbase), s1 = blend64(l[q].y, sbase), s2 = blend64(l[q].z, sbase), s3 = blend64(l[q].w, sbase);
                if (q < live_rows) {
                    if (c0 < (int)live_cols) st_stream_b8(dp, s0 >> 16);
                    if (c0 + 1 < (int)live_cols) st_stream_b8(dp + 1, s1 >> 16);
                    if (c2 < (int)live_cols) st_stream_b8(dp + 64, s2 >> 16);
                    if (c2 + 1 < (int)live_cols) st_stream_b8(dp + 65, s3 >> 16);
                }
                dp += dst_w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (flags & H3_LAST_ITEM) break;
        if (++s == S) { s = 0; phase ^= 1u; }
    }
}

// ---- direct kernel -----------------------------------------------------------------------------
enum DirectMode { DM_MONO = 0, DM_BGR_TO_RGB = 1, DM_BGR_TO_GRAY = 2, DM_NV12_TO_RGB = 3 };

struct DirectJob {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t src_stride, dst_stride;
    const lut_t* lut;
    int lut_pitch;
    int dst_w, dst_h, src_w, src_h;
    int mode;
};

template <int MODE>
__device__ __forceinline__ void fetch(const DirectJob& J, const uint8_t* src, int x, int y, uint32_t out[3]) {
    out[0] = out[1] = out[2] = 0u;
    if (x < 0 || x >= J.src_w || y < 0 || y >= J.src_h) return;  // BORDER_CONSTANT 0 (after conversion)
    const size_t p = (size_t)y * J.src_w + x;
    if (MODE == DM_MONO) {
        out[0] = src[p];
    } else if (MODE == DM_BGR_TO_RGB) {
        out[0] = src[3 * p + 2]; out[1] = src[3 * p + 1]; out[2] = src[3 * p];
    } else if (MODE == DM_BGR_TO_GRAY) {
        out[0] = gray_of(src[3 * p], src[3 * p + 1], src[3 * p + 2]);
    } else {
        const uint8_t* uv = src + (size_t)J.src_h * J.src_w + (size_t)(y >> 1) * J.src_w + (x & ~1);
        int r, g, b;
        yuv_to_rgb(src[p], uv[0], uv[1], r, g, b);
        out[0] = (uint32_t)r; out[1] = (uint32_t)g; out[2] = (uint32_t)b;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) rectify_direct_kernel(DirectJob J, int n_batch) {
    constexpr int C = (MODE == DM_BGR_TO_RGB || MODE == DM_NV12_TO_RGB) ? 3 : 1;
    const uint64_t npx = (uint64_t)J.dst_w * J.dst_h;
    const uint64_t total = npx * n_batch;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const uint64_t b = t / npx;
        const uint32_t p = (uint32_t)(t - b * npx);
        const int v = (int)(p / (uint32_t)J.dst_w), u = (int)(p - (uint32_t)v * J.dst_w);
        const lut_t ek = J.lut[(size_t)v * J.lut_pitch + u];
        uint8_t* dst = J.dst + b * J.dst_stride + (size_t)p * C;
        if (ek == LUT_OUTSIDE) {
            for (int c = 0; c < C; ++c) dst[c] = 0;
            continue;
        }
        const int sx = lut_x0(ek), sy = lut_y0(ek);
        const uint32_t fx = lut_fx(ek), fy = lut_fy(ek);
        const uint8_t* src = J.src + b * J.src_stride;
        uint32_t t00[3], t01[3], t10[3], t11[3];
        fetch<MODE>(J, src, sx, sy, t00);
        fetch<MODE>(J, src, sx + 1, sy, t01);
        fetch<MODE>(J, src, sx, sy + 1, t10);
        fetch<MODE>(J, src, sx + 1, sy + 1, t11);
        for (int c = 0; c < C; ++c) dst[c] = (uint8_t)bilinear_u8(t00[c], t01[c], t10[c], t11[c], fx, fy);
    }
}

// The overflow pixels of pair-window slots (ti_rectify_pair.cu): one thread per listed pixel and frame, taps from global
// memory through the generic LUT - the arithmetic of rectify_direct_kernel<DM_MONO>.  ONE launch for all jobs of a call
// (eight launches of a few thousand threads each cost a fisheye rig 40 us per call).
constexpr int MAX_POINT_JOBS = 16;
struct PointsParams {
    DirectJob job[MAX_POINT_JOBS];
    const uint32_t* pts[MAX_POINT_JOBS];
    uint32_t begin[MAX_POINT_JOBS + 1];  // prefix sum of points per frame over the jobs
    int n_jobs;
    int n_batch;
};

__global__ void __launch_bounds__(256) rectify_points_kernel(const __grid_constant__ PointsParams P) {
    const uint32_t per_frame = P.begin[P.n_jobs];
    const uint64_t total = (uint64_t)per_frame * P.n_batch;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const uint64_t b = t / per_frame;
        const uint32_t r = (uint32_t)(t - b * per_frame);
        int j = 0;
        while (j + 1 < P.n_jobs && r >= P.begin[j + 1]) ++j;
        const DirectJob& J = P.job[j];
        const uint32_t p = P.pts[j][r - P.begin[j]];
        const int v = (int)(p / (uint32_t)J.dst_w), u = (int)(p - (uint32_t)v * J.dst_w);
        const lut_t ek = J.lut[(size_t)v * J.lut_pitch + u];
        uint8_t* dst = J.dst + b * J.dst_stride + p;
        if (ek == LUT_OUTSIDE) { *dst = 0; continue; }
        const int sx = lut_x0(ek), sy = lut_y0(ek);
        const uint32_t fx = lut_fx(ek), fy = lut_fy(ek);
        const uint8_t* src = J.src + b * J.src_stride;
        uint32_t t00[3], t01[3], t10[3], t11[3];
        fetch<DM_MONO>(J, src, sx, sy, t00);
        fetch<DM_MONO>(J, src, sx + 1, sy, t01);
        fetch<DM_MONO>(J, src, sx, sy + 1, t10);
        fetch<DM_MONO>(J, src, sx + 1, sy + 1, t11);
        *dst = (uint8_t)bilinear_u8(t00[0], t01[0], t10[0], t11[0], fx, fy);
    }
}

static int direct_mode(int s, int d) {
    if ((s == TI_FMT_MONO8 || s == TI_FMT_NV12) && d == TI_FMT_MONO8) return DM_MONO;
    if (s == TI_FMT_BGR8 && d == TI_FMT_RGB8) return DM_BGR_TO_RGB;
    if (s == TI_FMT_BGR8 && d == TI_FMT_MONO8) return DM_BGR_TO_GRAY;
    if (s == TI_FMT_NV12 && d == TI_FMT_RGB8) return DM_NV12_TO_RGB;
    return -1;
}

static int launch_direct(ti_ctx* ctx, const DirectJob& D, int n_batch) {
    const uint64_t total = (uint64_t)D.dst_w * D.dst_h * n_batch;
    const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 16);
    switch (D.mode) {
        case DM_MONO: TI_LAUNCH(rectify_direct_kernel<DM_MONO>, grid, 256, 0, ctx->stream, D, n_batch); break;
        case DM_BGR_TO_RGB: TI_LAUNCH(rectify_direct_kernel<DM_BGR_TO_RGB>, grid, 256, 0, ctx->stream, D, n_batch); break;
        case DM_BGR_TO_GRAY: TI_LAUNCH(rectify_direct_kernel<DM_BGR_TO_GRAY>, grid, 256, 0, ctx->stream, D, n_batch); break;
        default: TI_LAUNCH(rectify_direct_kernel<DM_NV12_TO_RGB>, grid, 256, 0, ctx->stream, D, n_batch); break;
    }
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

template <int C>
static int launch_tiled(ti_ctx* ctx, RectParams& P, size_t smem_bytes) {
    const uint64_t total = (uint64_t)P.tiles_per_set * P.n_batch;
    if (total == 0) return TI_OK;
    if (smem_bytes > 48 * 1024) TI_CUDA(ctx, ensure_dynamic_smem(rectify_tile_kernel<C>, RT_MAX_SMEM, ctx->device));  // per device, not per process
    const int per_sm = resident_ctas(rectify_tile_kernel<C>, RT_THREADS, smem_bytes, 4);
    const int grid = (int)std::min<uint64_t>(total, (uint64_t)ctx->sm_count * per_sm);
    TI_LAUNCH(rectify_tile_kernel<C>, grid, RT_THREADS, smem_bytes, ctx->stream, P);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

// Two-pass paths.  BGR8 -> MONO8 rectified on a slot that qualifies for a fast mono kernel: the exact OpenCV gray
// conversion (convert_vec_kernel) into library scratch, then the mono remap on the gray frames - identical to
// cv2.remap(cvtColor(BGR2GRAY)) by construction.  NV12 -> RGB8 rectified: the exact NV12 -> BGR conversion into
// scratch, then the 3-channel window remap (which swaps to RGB) - remap is per channel, so this equals
// cv2.remap(cvtColor(YUV2RGB_NV12)).  Both are several times faster than converting every tap inside the generic kernel.
//
// TI_OPT_L2_SCRATCH_KB > 0 cuts the batch into chunks whose scratch fits that budget, every chunk converted and remapped
// before the next one overwrites the same scratch lines, so that the second pass reads what the first just wrote from the L2.
// Measured on the B200 (profiles/r02_summary.md): 40 MB chunks 0.39 of peak against 0.49 for the whole batch in one chunk
// (BGR8 -> MONO8, 4 x 1920 x 1200 x 16) - the smaller launches lose more than the saved re-read gains - so the default is one
// chunk.  mid_fmt_of() says which jobs take the two-pass route; `two_pass` holds them with their scratch offsets.

struct TwoPassJob {
    RectifyJob job;     // as given by the caller
    int mid_fmt;
    size_t mid_frame;   // bytes of one intermediate frame (multiple of 16)
};

static int mid_fmt_of(ti_ctx* ctx, const RectifyJob& J, int th4, int thk, int* mid) {
    *mid = -1;
    if (ctx->force_generic_rectify || J.camera < 0 || J.camera >= TI_MAX_CAMERAS) return TI_OK;
    CameraSlot& C = ctx->cams[J.camera];
    if (!C.has_map || C.src_w % 16 != 0 || !J.src || ((uintptr_t)J.src % 16) || (J.src_stride % 16)) return TI_OK;
    if (J.src_fmt == TI_FMT_BGR8 && J.dst_fmt == TI_FMT_MONO8 && ctx->mono_variant >= 3 && (C.has_pair[th4] || C.has_tma_mono[thk])) {
        *mid = TI_FMT_MONO8;
    } else if (J.src_fmt == TI_FMT_NV12 && J.dst_fmt == TI_FMT_RGB8 && ctx->mono_variant == 4 && !((C.src_w | C.src_h) & 1)) {
        if (!C.c3_tried) {
            const int rc = build_c3_tables(ctx, C);
            if (rc != TI_OK) return rc;
        }
        if (C.has_c3) *mid = TI_FMT_BGR8;
    }
    return TI_OK;
}

static int launch_rectify_direct(ti_ctx* ctx, const RectifyJob* jobs_in, int n_jobs, int n_batch);

int launch_rectify(ti_ctx* ctx, const RectifyJob* jobs_in, int n_jobs, int n_batch) {
    if (n_jobs <= 0 || n_batch <= 0) return TI_OK;
    for (int i = 0; i < n_jobs; ++i)
        if (!jobs_in[i].src || !jobs_in[i].dst) return fail(ctx, TI_EINVAL, "rectify: null src/dst pointer");
    const int th4 = p4_th_index(ctx->tma_tile_h), thk = m3_th_index(ctx->tma_tile_h);
    std::vector<RectifyJob> single;
    std::vector<TwoPassJob> two_pass;
    size_t mid_per_frame_set = 0;
    for (int i = 0; i < n_jobs; ++i) {
        int mid = -1;
        const int rc = mid_fmt_of(ctx, jobs_in[i], th4, thk, &mid);
        if (rc != TI_OK) return rc;
        if (mid < 0) { single.push_back(jobs_in[i]); continue; }
        const CameraSlot& C = ctx->cams[jobs_in[i].camera];
        const size_t frame = (size_t)frame_bytes(mid, C.src_w, C.src_h);  // multiple of 16 (src_w is)
        two_pass.push_back(TwoPassJob{jobs_in[i], mid, frame});
        mid_per_frame_set += frame;
    }
    if (!single.empty()) {
        const int rc = launch_rectify_direct(ctx, single.data(), (int)single.size(), n_batch);
        if (rc != TI_OK) return rc;
    }
    if (two_pass.empty()) return TI_OK;
    const int chunk = ctx->l2_scratch_kb <= 0 ? n_batch
                                              : (int)std::max<size_t>(1, std::min<size_t>((size_t)n_batch, ((size_t)ctx->l2_scratch_kb << 10) / mid_per_frame_set));
    const size_t need = mid_per_frame_set * (size_t)chunk;
    if (need > ctx->scratch_cap) {
        if (ctx->scratch) cudaFree(ctx->scratch);  // synchronises: earlier launches that read it have finished
        ctx->scratch = nullptr; ctx->scratch_cap = 0;
        TI_CUDA(ctx, cudaMalloc(&ctx->scratch, need));
        ctx->scratch_cap = need;
    }
    std::vector<ConvertJob> conv(two_pass.size());
    std::vector<RectifyJob> second(two_pass.size());
    for (int b0 = 0; b0 < n_batch; b0 += chunk) {
        const int nb = std::min(chunk, n_batch - b0);
        size_t off = 0;
        for (size_t i = 0; i < two_pass.size(); ++i) {
            const TwoPassJob& T = two_pass[i];
            const CameraSlot& C = ctx->cams[T.job.camera];
            uint8_t* mid = static_cast<uint8_t*>(ctx->scratch) + off;
            off += T.mid_frame * (size_t)chunk;
            conv[i] = ConvertJob{T.job.src + (uint64_t)b0 * T.job.src_stride, mid, T.job.src_stride, (uint64_t)T.mid_frame, C.src_w, C.src_h,
                                 T.job.src_fmt, T.mid_fmt};
            second[i] = RectifyJob{mid, T.job.dst + (uint64_t)b0 * T.job.dst_stride, (uint64_t)T.mid_frame, T.job.dst_stride, T.job.camera,
                                   T.mid_fmt, T.job.dst_fmt};
        }
        for (size_t i = 0; i < conv.size(); i += TI_MAX_STREAMS) {  // launch_convert takes TI_MAX_STREAMS jobs at a time
            const int rc = launch_convert(ctx, conv.data() + i, (int)std::min<size_t>(TI_MAX_STREAMS, conv.size() - i), nb);
            if (rc != TI_OK) return rc;
        }
        const int rc = launch_rectify_direct(ctx, second.data(), (int)second.size(), nb);
        if (rc != TI_OK) return rc;
    }
    return TI_OK;
}

// Everything that is one pass over the source: routes every job to the fastest kernel its slot qualifies for.
static int launch_rectify_direct(ti_ctx* ctx, const RectifyJob* jobs_in, int n_jobs, int n_batch) {
    if (n_jobs <= 0 || n_batch <= 0) return TI_OK;
    std::vector<RectifyJob> jobs(jobs_in, jobs_in + n_jobs);
    RectParams P1{}, P3{};  // tiled launches for 1-channel and 3-channel sources
    Rect2Params P2{};       // fast mono launch (v2: thread-staged)
    Rect3Params PT{};       // fast mono launch (v3: TMA-pipelined)
    std::deque<Rect4Params> pair_groups;  // fast mono launches (v4: pair / quad windows), one per (staged row pitch, layout)
    Rect5Params PC{}, PCW{};  // fast 3-channel launches (BGR8 -> RGB8 windows): standard and wide source boxes
    std::vector<DirectJob> pair_overflow;  // pair-window jobs whose slot has an overflow list (mode = camera slot)
    const int thk = m3_th_index(ctx->tma_tile_h);
    const int th4 = p4_th_index(ctx->tma_tile_h);
    size_t smem1 = 0, smem3 = 0, smem2 = 0;
    for (int i = 0; i < n_jobs; ++i) {
        const RectifyJob& J = jobs[i];
        if (J.camera < 0 || J.camera >= TI_MAX_CAMERAS || !ctx->cams[J.camera].has_map)
            return fail(ctx, TI_ESTATE, "rectify: camera slot %d has no remap LUT (call ti_upload_rectify_map)", J.camera);
        const int mode = direct_mode(J.src_fmt, J.dst_fmt);
        if (mode < 0) return fail(ctx, TI_EINVAL, "rectify: unsupported conversion %d -> %d", J.src_fmt, J.dst_fmt);
        if (!J.src || !J.dst) return fail(ctx, TI_EINVAL, "rectify: null src/dst pointer");
        const CameraSlot& C = ctx->cams[J.camera];
        if (J.src_fmt == TI_FMT_NV12 && ((C.src_w | C.src_h) & 1))
            return fail(ctx, TI_EINVAL, "NV12 needs even width and height (got %dx%d)", C.src_w, C.src_h);
        const int lut_pitch = C.tiles_x * RT_W;
        const int ch = (mode == DM_BGR_TO_RGB) ? 3 : 1;
        const size_t need = C.tile_smem[ch == 1 ? 0 : 1];
        const bool tiled_ok = (mode == DM_MONO || mode == DM_BGR_TO_RGB) && (C.src_w * ch) % 16 == 0 && C.dst_w % 8 == 0 &&
                              ((uintptr_t)J.src % 16 == 0) && (J.src_stride % 16 == 0) && ((uintptr_t)J.dst % 8 == 0) &&
                              (J.dst_stride % 8 == 0) && need <= (size_t)RT_MAX_SMEM;
        const bool fast_ok = mode == DM_MONO && C.has_fast_mono && ((uintptr_t)J.src % 16 == 0) &&
                             (J.src_stride % 16 == 0) && P2.n_jobs < MAX_RECT_JOBS && !ctx->force_generic_rectify && ctx->mono_variant >= 2;
        const bool tma_ok = mode == DM_MONO && C.has_tma_mono[thk] && ((uintptr_t)J.src % 16 == 0) && (J.src_stride % 16 == 0) &&
                            PT.n_jobs < MAX_RECT_JOBS && !ctx->force_generic_rectify && ctx->mono_variant >= 3;
        const bool pair_ok = mode == DM_MONO && C.has_pair[th4] && ((uintptr_t)J.src % 16 == 0) && (J.src_stride % 16 == 0) &&
                             !ctx->force_generic_rectify && ctx->mono_variant == 4;
        if (mode == DM_BGR_TO_RGB && !ctx->force_generic_rectify && ctx->mono_variant == 4 && C.src_w % 16 == 0 &&
            ((uintptr_t)J.src % 16 == 0) && (J.src_stride % 16 == 0)) {
            CameraSlot& CM = ctx->cams[J.camera];
            if (!CM.c3_tried) {
                const int rc = build_c3_tables(ctx, CM);
                if (rc != TI_OK) return rc;
            }
            Rect5Params& QC = CM.pitch5 == C3_PITCH_WIDE ? PCW : PC;  // jobs of a launch stage rows of one pitch
            if (CM.has_c3 && QC.n_jobs < MAX_PAIR_JOBS) {
                const int rc = tma_encode_3d(ctx, &QC.map[QC.n_jobs], J.src, 4, 3 * C.src_w / 4, C.src_h, n_batch, (uint64_t)3 * C.src_w,
                                             J.src_stride, CM.pitch5 / 4, CM.rows5_alloc);
                if (rc != TI_OK) return rc;
                Rect5JobDev D{};
                D.lut5 = CM.d_lut5; D.boxes5 = CM.d_boxes5; D.dst = J.dst; D.dst_stride = J.dst_stride;
                D.dst_w = C.dst_w; D.dst_h = C.dst_h; D.rows_alloc = CM.rows5_alloc;
                D.tile_begin = QC.tiles_per_set;
                QC.tiles_per_set += (uint32_t)(CM.tiles5_x * CM.tiles5_y);
                QC.rows_alloc_max = std::max(QC.rows_alloc_max, D.rows_alloc);
                QC.pitch = CM.pitch5;
                QC.job[QC.n_jobs++] = D;
                continue;
            }
        }
        if (pair_ok) {
            // jobs of a launch stage rows of one pitch, in one layout
            const bool wide_slot = C.pitch4[th4] == P4_PITCH_WIDE;
            Rect4Params* Qp = nullptr;
            for (Rect4Params& G : pair_groups)
                if (G.pitch == C.pitch4[th4] && G.quad == (C.quad4[th4] ? 1 : 0) && G.n_jobs < MAX_PAIR_JOBS) Qp = &G;
            if (!Qp) {
                pair_groups.emplace_back();
                Qp = &pair_groups.back();
                *Qp = Rect4Params{};
                Qp->pitch = C.pitch4[th4];
                Qp->quad = C.quad4[th4] ? 1 : 0;
            }
            Rect4Params& Q = *Qp;
            {
                const int rc = wide_slot ? tma_encode_3d(ctx, &Q.map[Q.n_jobs], J.src, 4, C.src_w / 4, C.src_h, n_batch, (uint64_t)C.src_w, J.src_stride,
                                                         P4_PITCH_WIDE / 4, C.rows4_alloc[th4])
                                         : tma_encode_u8_3d(ctx, &Q.map[Q.n_jobs], J.src, C.src_w, C.src_h, n_batch, (uint64_t)C.src_w, J.src_stride,
                                                            C.pitch4[th4], C.rows4_alloc[th4]);
                if (rc != TI_OK) return rc;
                Rect4JobDev D{};
                D.lut4 = C.d_lut4[th4]; D.boxes4 = C.d_boxes4[th4]; D.exc4 = C.d_exc4[th4]; D.dst = J.dst; D.dst_stride = J.dst_stride;
                D.dst_w = C.dst_w; D.dst_h = C.dst_h; D.rows_alloc = C.rows4_alloc[th4]; D.exc_per_warp = C.exc4_per_warp[th4];
                D.tile_begin = Q.tiles_per_set;
                Q.tiles_per_set += (uint32_t)(C.tiles4_x[th4] * C.tiles4_y[th4]);
                Q.rows_alloc_max = std::max(Q.rows_alloc_max, D.rows_alloc);
                Q.exc_max = std::max(Q.exc_max, D.exc_per_warp);
                Q.job[Q.n_jobs++] = D;
                if (C.n_over4[th4] > 0)
                    pair_overflow.push_back(DirectJob{J.src, J.dst, J.src_stride, J.dst_stride, C.d_lut, lut_pitch, C.dst_w, C.dst_h, C.src_w, C.src_h, J.camera});
                continue;
            }
        }
        if (tma_ok) {
            const int rc = tma_encode_u8_3d(ctx, &PT.map[PT.n_jobs], J.src, C.src_w, C.src_h, n_batch, (uint64_t)C.src_w,
                                            J.src_stride, M3_PITCH, C.rows3_alloc[thk]);
            if (rc != TI_OK) return rc;
            Rect3JobDev D{};
            D.lut3 = C.d_lut3[thk]; D.boxes3 = C.d_boxes3[thk]; D.dst = J.dst; D.dst_stride = J.dst_stride;
            D.dst_w = C.dst_w; D.dst_h = C.dst_h; D.rows_alloc = C.rows3_alloc[thk];
            D.tile_begin = PT.tiles_per_set;
            PT.tiles_per_set += (uint32_t)(C.tiles3_x[thk] * C.tiles3_y[thk]);
            PT.rows_alloc_max = std::max(PT.rows_alloc_max, C.rows3_alloc[thk]);
            PT.job[PT.n_jobs++] = D;
            continue;
        }
        if (fast_ok) {
            Rect2JobDev D{};
            D.src = J.src; D.dst = J.dst; D.src_stride = J.src_stride; D.dst_stride = J.dst_stride;
            D.lut2 = C.d_lut2; D.boxes2 = C.d_boxes2; D.tiles_x = C.tiles2_x; D.n_tiles = C.tiles2_x * C.tiles2_y;
            D.dst_w = C.dst_w; D.dst_h = C.dst_h; D.src_w = C.src_w; D.src_h = C.src_h;
            D.tile_begin = P2.tiles_per_set;
            P2.tiles_per_set += (uint32_t)D.n_tiles;
            P2.job[P2.n_jobs++] = D;
            smem2 = std::max(smem2, (size_t)M2_ZERO_BYTES + (size_t)std::max(C.rows2_max, 2) * M2_ROW_BYTES);
            continue;
        }
        RectParams& P = ch == 1 ? P1 : P3;
        if (!tiled_ok || P.n_jobs == MAX_RECT_JOBS) {
            DirectJob D{J.src, J.dst, J.src_stride, J.dst_stride, C.d_lut, lut_pitch, C.dst_w, C.dst_h, C.src_w, C.src_h, mode};
            const int rc = launch_direct(ctx, D, n_batch);
            if (rc != TI_OK) return rc;
            continue;
        }
        RectJobDev D{};
        D.src = J.src; D.dst = J.dst; D.src_stride = J.src_stride; D.dst_stride = J.dst_stride;
        D.lut = C.d_lut; D.boxes = C.d_boxes; D.lut_pitch = lut_pitch;
        D.tiles_x = C.tiles_x; D.tiles_y = C.tiles_y;
        D.dst_w = C.dst_w; D.dst_h = C.dst_h; D.src_w = C.src_w; D.src_h = C.src_h;
        D.tile_begin = P.tiles_per_set;
        P.tiles_per_set += (uint32_t)(C.tiles_x * C.tiles_y);
        P.job[P.n_jobs++] = D;
        size_t& smem = ch == 1 ? smem1 : smem3;
        smem = std::max(smem, need);
    }
    P1.n_batch = P3.n_batch = P2.n_batch = PT.n_batch = n_batch;
    PT.debug = ctx->debug;
    PC.n_batch = PCW.n_batch = n_batch;
    if (PC.n_jobs) {
        const int rc = launch_rectify_c3(ctx, PC);
        if (rc != TI_OK) return rc;
    }
    if (PCW.n_jobs) {
        const int rc = launch_rectify_c3(ctx, PCW);
        if (rc != TI_OK) return rc;
    }
    if (!pair_groups.empty()) {
        for (Rect4Params& G : pair_groups) {
            G.n_batch = n_batch;
            const int rc = launch_rectify_pair(ctx, G, th4);
            if (rc != TI_OK) return rc;
        }
        for (size_t i0 = 0; i0 < pair_overflow.size(); i0 += MAX_POINT_JOBS) {  // slots with more exceptions in some (tile, warp) than its list holds
            PointsParams Q{};
            for (size_t i = i0; i < std::min(pair_overflow.size(), i0 + MAX_POINT_JOBS); ++i) {
                const CameraSlot& C = ctx->cams[pair_overflow[i].mode];  // mode carries the slot here
                Q.job[Q.n_jobs] = pair_overflow[i];
                Q.job[Q.n_jobs].mode = DM_MONO;
                Q.pts[Q.n_jobs] = C.d_over4[th4];
                Q.begin[Q.n_jobs + 1] = Q.begin[Q.n_jobs] + (uint32_t)C.n_over4[th4];
                ++Q.n_jobs;
            }
            Q.n_batch = n_batch;
            const uint64_t total = (uint64_t)Q.begin[Q.n_jobs] * n_batch;
            const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 8);
            TI_LAUNCH(rectify_points_kernel, grid, 256, 0, ctx->stream, Q);
            TI_CHECK_LAUNCH(ctx);
        }
    }
    if (PT.n_jobs) {
        const int TH = M3_TILE_HEIGHTS[thk];
        const size_t stage = 128 + 2 * (size_t)PT.rows_alloc_max * M3_PITCH + 128;
        PT.frames_per_unit = std::max(1, std::min(n_batch, ctx->frames_per_unit));
        PT.stages = std::max(2, std::min(ctx->stages, M3_MAX_STAGES));
        while (PT.stages > 2 && 256 + (size_t)PT.stages * stage > 220 * 1024) --PT.stages;  // tall source boxes: fewer stages
        const size_t smem = 256 + (size_t)PT.stages * stage;
        const uint64_t total = (uint64_t)PT.tiles_per_set * ((n_batch + PT.frames_per_unit - 1) / PT.frames_per_unit);
        typedef void (*Kern)(const Rect3Params);
        static const Kern kernels[3][2] = {{rectify_mono_tma_kernel<16, false>, rectify_mono_tma_kernel<16, true>},
                                           {rectify_mono_tma_kernel<32, false>, rectify_mono_tma_kernel<32, true>},
                                           {rectify_mono_tma_kernel<24, false>, rectify_mono_tma_kernel<24, true>}};
        const Kern kern = kernels[thk][ctx->lut_prefetch ? 1 : 0];
        (void)TH;
        TI_CUDA(ctx, ensure_dynamic_smem(kern, smem, ctx->device));
        int per_sm = resident_ctas(kern, M3_THREADS, smem, 3);
        if (ctx->ctas_per_sm > 0) per_sm = ctx->ctas_per_sm;
        const int grid = (int)std::min<uint64_t>(total, (uint64_t)ctx->sm_count * per_sm);
        TI_LAUNCH(kern, grid, M3_THREADS, smem, ctx->stream, PT);
        TI_CHECK_LAUNCH(ctx);
    }
    if (P2.n_jobs) {
        const uint64_t total = (uint64_t)P2.tiles_per_set * n_batch;
        TI_CUDA(ctx, ensure_dynamic_smem(rectify_mono_kernel, M2_ZERO_BYTES + M2_MAX_ROWS * M2_ROW_BYTES, ctx->device));
        int per_sm = resident_ctas(rectify_mono_kernel, M2_THREADS, smem2, 4);
        if (ctx->ctas_per_sm > 0) per_sm = ctx->ctas_per_sm;
        const int grid = (int)std::min<uint64_t>(total, (uint64_t)ctx->sm_count * per_sm);
        TI_LAUNCH(rectify_mono_kernel, grid, M2_THREADS, smem2, ctx->stream, P2);
        TI_CHECK_LAUNCH(ctx);
    }
    int rc = launch_tiled<1>(ctx, P1, smem1);
    if (rc != TI_OK) return rc;
    return launch_tiled<3>(ctx, P3, smem3);
}

}  // namespace ti
