// 3-channel window remap ("c3"): BGR8 -> RGB8 fused with the undistort / rectify remap, bit-exact with
// cv2.remap(cv2.cvtColor(src, BGR2RGB), INTER_LINEAR, BORDER_CONSTANT 0).
//
// Takes over, for the colour stereo streams of the long-range cameras (BASELINE config 4), the BGR2RGB swap of
// thor_slam/slam/adapters/isaac_ros.py:351-358 and the undistortion the reference leaves to cuVSLAM
// (isaac_ros.py:364-411).
//
// Same machinery as the mono pair-window kernel (ti_rectify_pair.cu): persistent CTAs, one TMA-issuer warp and
// eight consumer warps, a ring of source boxes on mbarriers, the tile's LUT slice prefetched into shared memory
// by a bulk copy one unit (tile x frames of the batch) ahead and expanded once per unit into 2-D weights held in
// registers.  What differs:
//  * the source row is 3 bytes per pixel; the TMA box is described in 32-bit elements (128 of them = 512 B per
//    row), so one box still covers a 128-pixel tile;
//  * every output pixel has its own WINDOW: the three aligned words that contain the six bytes
//    B0 G0 R0 B1 G1 R1 of its two horizontal taps, and the three words below.  Two funnel shifts (shift from the
//    LUT) line the six bytes up, two PRMT with fixed selectors regroup them as (B0 B1 G0 G1) and (R0 R1 . .), and
//    each channel is two IDP.2A against the pixel's 16-bit 2-D weights - result 64*(S+512), the channel value is
//    byte 2.  No pixel depends on a neighbour, so there are no exceptions to fix up;
//  * for the window loads lane L of a warp takes pixels L, L+32, L+64, L+96 of a tile row: the 32 windows of one load
//    instruction then cover ~100 consecutive source bytes (<= 32 distinct words: no bank conflicts - with 4 consecutive
//    pixels per lane the windows are 3 words apart and 57 % of the shared-memory wavefronts were conflict replays);
//    the results cross a 512-byte per-warp buffer so that for the stores a lane owns 4 consecutive pixels = 12 output
//    bytes (R G B order), written as three 32-bit streaming stores.
#include "ti_rectify_pair.cuh"
#include "ti_pair_dev.cuh"

#include <cstring>

namespace ti {

constexpr int C3_RPW = C3_TH / C3_CONSUMER_WARPS;  // tile rows per consumer warp (2)
constexpr uint32_t C3_LUT_BYTES = (uint32_t)C3_TH * C3_LUT_ROW_WORDS * 4u;

struct C3Unit { int j; uint32_t tile, b0, nb; };
__device__ __forceinline__ C3Unit c3_unit(const Rect5Params& P, uint32_t k) {
    const uint64_t ug = (uint64_t)blockIdx.x + (uint64_t)k * gridDim.x;
    const uint32_t c = (uint32_t)(ug / P.tiles_per_set), r = (uint32_t)(ug - (uint64_t)c * P.tiles_per_set);
    int j = 0;
    while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
    C3Unit U;
    U.j = j; U.tile = r - P.job[j].tile_begin;
    U.b0 = c * (uint32_t)P.frames_per_unit;
    U.nb = min((uint32_t)P.frames_per_unit, (uint32_t)P.n_batch - U.b0);
    return U;
}

__device__ __forceinline__ void c3_st32(void* p, uint32_t v) {
#ifdef TI_EMULATE
    *reinterpret_cast<uint32_t*>(ti_emu::check_align(p, 4)) = v;
#else
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// one output pixel: window word m (offset << 16 | 8 * byte alignment), weights wt / wb -> (R, G, B, .) in bytes 0..2
template <int PITCH>
__device__ __forceinline__ uint32_t c3_pixel(p4_addr_t base, uint32_t m, uint32_t wt, uint32_t wb) {
    const p4_addr_t a = base + (m >> 16);
    const uint32_t t0 = p4_lds<0>(a), t1 = p4_lds<4>(a), t2 = p4_lds<8>(a);
    const uint32_t b0 = p4_lds<PITCH>(a), b1 = p4_lds<PITCH + 4>(a), b2 = p4_lds<PITCH + 8>(a);
    const uint32_t ta = __funnelshift_r(t0, t1, m), tb = __funnelshift_r(t1, t2, m);  // bytes B0 G0 R0 B1 | G1 R1 . .
    const uint32_t ba = __funnelshift_r(b0, b1, m), bb = __funnelshift_r(b1, b2, m);
    const uint32_t tx = __byte_perm(ta, tb, 0x4130), ty = __byte_perm(ta, tb, 0x5252);  // (B0 B1 G0 G1), (R0 R1 R0 R1)
    const uint32_t bx = __byte_perm(ba, bb, 0x4130), by = __byte_perm(ba, bb, 0x5252);
    const uint32_t cb = __dp2a_lo(wb, bx, __dp2a_lo(wt, tx, 32768u));
    const uint32_t cg = __dp2a_hi(wb, bx, __dp2a_hi(wt, tx, 32768u));
    const uint32_t cr = __dp2a_lo(wb, by, __dp2a_lo(wt, ty, 32768u));
    return __byte_perm(__byte_perm(cr, cg, 0x0062), cb, 0x0610);
}

// DSTW > 0: the destination row pitch in pixels is this compile-time constant (row stores become immediate offsets).
template <int DSTW, int PITCH = C3_PITCH>
__global__ void __launch_bounds__(C3_THREADS, 3) rectify_c3_kernel(const __grid_constant__ Rect5Params P) {
    TI_DYNAMIC_SMEM(uint8_t, smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t stage_bytes = (uint32_t)P.rows_alloc_max * PITCH;
    const int S = P.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [S] box landed
    uint64_t* empty = full + C3_MAX_STAGES;              // [S] consumers done
    uint64_t* lut_full = empty + C3_MAX_STAGES;          // LUT slice of the unit landed
    uint64_t* lut_empty = lut_full + 1;                  // every consumer warp has expanded its part of the slice
    uint8_t* stage0 = smem + 256;
    uint8_t* lutbuf = stage0 + (size_t)S * stage_bytes;
    // after the LUT slice: one 512-byte transpose buffer per consumer warp

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, C3_CONSUMER_WARPS);
        }
        mbar_init(lut_full, 1);
        mbar_init(lut_empty, C3_CONSUMER_WARPS);
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t n_chunks = (uint32_t)((P.n_batch + P.frames_per_unit - 1) / P.frames_per_unit);
    const uint64_t total_units = (uint64_t)P.tiles_per_set * n_chunks;
    const uint32_t units_mine = total_units > blockIdx.x ? (uint32_t)((total_units - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    if (units_mine == 0) return;

    if (warp == C3_CONSUMER_WARPS) {
        // ------------------------------------------------ issuer (one thread) ------------------------
        if (lane != 0) return;
        struct Unit { int j; uint32_t tile, b0, nb; uint4 box; };
        auto load_unit = [&](uint32_t k, Unit& U) {
            const C3Unit u = c3_unit(P, k);
            U.j = u.j; U.tile = u.tile; U.b0 = u.b0; U.nb = u.nb;
            U.box = *reinterpret_cast<const uint4*>(P.job[u.j].boxes5 + u.tile);
        };
        auto issue_lut = [&](const Unit& U) {
            bulk_load_1d(lutbuf, P.job[U.j].lut5 + (size_t)U.tile * (C3_TH * C3_LUT_ROW_WORDS), C3_LUT_BYTES, lut_full);
            mbar_arrive_expect_tx(lut_full, C3_LUT_BYTES);
        };
        Unit cur{}, nxt{};
        load_unit(0, cur);
        issue_lut(cur);
        if (units_mine > 1) load_unit(1, nxt);
        int s = 0;
        uint32_t phase = 1;  // parity to wait for on empty[s]: 1 on a stage's first use (passes at once)
        for (uint32_t k = 0; k < units_mine; ++k) {
            const Rect5JobDev& J = P.job[cur.j];
            const int c0 = (int16_t)(cur.box.x & 0xFFFF), y0 = (int16_t)(cur.box.x >> 16);
            const int rows = (int16_t)(cur.box.y >> 16);
            const uint32_t tx = rows > 0 ? (uint32_t)J.rows_alloc * PITCH : 0u;
            bool lut_pending = k + 1 < units_mine;  // the next unit's LUT slice still has to be requested
            for (uint32_t f = 0; f < cur.nb; ++f) {
                uint8_t* sb = stage0 + (size_t)s * stage_bytes;
                mbar_wait_hint(empty + s, phase, 20000);  // consumers have released the stage's previous item (suspended in hardware, no polling)
                if (rows > 0) tma_load_3d(sb, &P.map[cur.j], c0 / 4, y0, (int)(cur.b0 + f), full + s);  // 128 u32 x rows_alloc
                mbar_arrive_expect_tx(full + s, tx);
                if (++s == S) { s = 0; phase ^= 1u; }
                if (lut_pending && mbar_test(lut_empty, k & 1u)) { issue_lut(nxt); lut_pending = false; }
            }
            if (lut_pending) { mbar_wait_relaxed(lut_empty, k & 1u); issue_lut(nxt); }
            cur = nxt;
            if (k + 2 < units_mine) load_unit(k + 2, nxt);
        }
        return;
    }
    // ---------------------------------------------------- consumers ---------------------------------
    // Lane L blends pixels L, L+32, L+64, L+96 of each of its warp's C3_RPW tile rows and stores pixels 4L .. 4L+3.
    uint32_t mw[C3_RPW][4], wt[C3_RPW][4], wb[C3_RPW][4];
    const p4_addr_t sm0 = p4_addr(smem);
    const p4_addr_t stage_first = sm0 + 256;
    p4_addr_t base = stage_first, bar = sm0;  // current stage, its `full` barrier (`empty` is 64 bytes further)
    int s_left = S;                            // stages until the ring wraps
    uint32_t* const tbuf = reinterpret_cast<uint32_t*>(lutbuf + C3_LUT_BYTES) + warp * 128;  // this warp's transpose buffer
    uint32_t phase = 0;
    for (uint32_t k = 0; k < units_mine; ++k) {
        const C3Unit U = c3_unit(P, k);
        const Rect5JobDev& J = P.job[U.j];
        const uint4 box = *reinterpret_cast<const uint4*>(J.boxes5 + U.tile);
        const int u0 = (int16_t)(box.z & 0xFFFF), v0 = (int16_t)(box.z >> 16) + warp * C3_RPW;
        const int dst_w = J.dst_w, live_rows = J.dst_h - v0, live_cols = J.dst_w - u0 - 4 * lane;  // of this lane's 4 pixels
        const uint64_t dst_stride = J.dst_stride;
        uint8_t* dp = J.dst + (uint64_t)U.b0 * dst_stride + ((size_t)v0 * dst_w + u0 + 4 * lane) * 3;
        const bool whole = live_rows >= C3_RPW && J.dst_w - u0 >= C3_TW &&
                           ((((uint64_t)(uintptr_t)J.dst | dst_stride | (uint64_t)(3 * dst_w)) & 3ull) == 0);  // warp-uniform

        mbar_wait(lut_full, k & 1u);
        {
            const uint4* lp = reinterpret_cast<const uint4*>(lutbuf) + ((size_t)(warp * C3_RPW) * (C3_LUT_ROW_WORDS / 4) + lane * 2);
#pragma unroll
            for (int q = 0; q < C3_RPW; ++q) {
                const uint4 e0 = lp[q * (C3_LUT_ROW_WORDS / 4)], e1 = lp[q * (C3_LUT_ROW_WORDS / 4) + 1];  // {m0,p0,m1,p1},{m2,p2,m3,p3}
                mw[q][0] = e0.x; mw[q][1] = e0.z; mw[q][2] = e1.x; mw[q][3] = e1.z;
                p4_expand(e0.y, wt[q][0], wb[q][0]);
                p4_expand(e0.w, wt[q][1], wb[q][1]);
                p4_expand(e1.y, wt[q][2], wb[q][2]);
                p4_expand(e1.w, wt[q][3], wb[q][3]);
            }
        }
        __syncwarp();  // every lane's reads of the LUT slice are ordered before the release below
        if (lane == 0) mbar_arrive(lut_empty);

        for (uint32_t f = 0; f < U.nb; ++f) {
            p4_wait(bar, phase);
            uint8_t* rp = dp;
#pragma unroll
            for (int q = 0; q < C3_RPW; ++q) {
                uint32_t px[4];
                __syncwarp();  // the previous row's reads of the transpose buffer are done
#pragma unroll
                for (int j = 0; j < 4; ++j) tbuf[32 * j + lane] = c3_pixel<PITCH>(base, mw[q][j], wt[q][j], wb[q][j]);  // pixel 32j + lane
                __syncwarp();
                {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(tbuf + 4 * lane);  // pixels 4 lane .. 4 lane + 3
                    px[0] = t4.x; px[1] = t4.y; px[2] = t4.z; px[3] = t4.w;
                }
                if (whole && DSTW > 0) {
                    c3_st32(dp + q * (DSTW * 3), __byte_perm(px[0], px[1], 0x4210));
                    c3_st32(dp + q * (DSTW * 3) + 4, __byte_perm(px[1], px[2], 0x5421));
                    c3_st32(dp + q * (DSTW * 3) + 8, __byte_perm(px[2], px[3], 0x6542));
                } else if (whole) {
                    c3_st32(rp, __byte_perm(px[0], px[1], 0x4210));
                    c3_st32(rp + 4, __byte_perm(px[1], px[2], 0x5421));
                    c3_st32(rp + 8, __byte_perm(px[2], px[3], 0x6542));
                } else if (q < live_rows) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < live_cols) {
                            st_stream_b8(rp + 3 * j, px[j]);
                            st_stream_b8(rp + 3 * j + 1, px[j] >> 8);
                            st_stream_b8(rp + 3 * j + 2, px[j] >> 16);
                        }
                }
                rp += (size_t)dst_w * 3;
            }
            p4_warp_arrive(bar + 64);
            dp += dst_stride;
            bar += 8; base += stage_bytes;
            if (--s_left == 0) { s_left = S; bar -= 8u * (uint32_t)S; base -= (uint32_t)S * stage_bytes; phase ^= 1u; }
        }
    }
}

// ---- launcher ------------------------------------------------------------------------------------
int launch_rectify_c3(ti_ctx* ctx, Rect5Params& P) {
    if (P.n_jobs == 0 || P.n_batch <= 0) return TI_OK;
    const size_t stage = (size_t)P.rows_alloc_max * P.pitch;
    int stages = std::max(2, std::min(ctx->stages4, C3_MAX_STAGES));
    const size_t tail = C3_LUT_BYTES + (size_t)C3_CONSUMER_WARPS * 512;  // LUT slice + per-warp transpose buffers
    // ... and never so deep that the SM has no shared memory left for anybody else: the exchange kernels (ti_push.cu: a 12 KB TMA
    // copy CTA, one-warp flag kernels, each with its 1 KB of system shared memory) run BESIDE this kernel's resident CTAs.  When
    // they did not fit, whichever came first displaced a CTA of this persistent grid, which then started late and stretched the
    // kernel by a third (measured: 14 us per step on the fusing rank).  Ring depth beyond four stages buys nothing (round 1).
    while (stages > 2 && (256 + (size_t)stages * stage + tail + 1024) * 3 > (228 - ctx->smem_headroom_kb) * 1024) --stages;
    P.stages = stages;
    const size_t smem = 256 + (size_t)stages * stage + tail;
    if (smem > 220 * 1024) return fail(ctx, TI_EINVAL, "rectify (3-channel): source boxes of %d rows do not fit shared memory", P.rows_alloc_max);
    typedef void (*Kern)(const Rect5Params);
    const bool wide = P.pitch == C3_PITCH_WIDE;
    Kern kern = wide ? (Kern)rectify_c3_kernel<0, C3_PITCH_WIDE> : (Kern)rectify_c3_kernel<0>;
    if (!wide) {  // every job of the launch writes rows of the same common pitch: immediate row offsets
        int dw = P.job[0].dst_w;
        for (int j = 1; j < P.n_jobs; ++j)
            if (P.job[j].dst_w != dw) dw = 0;
        if (dw == 1920) kern = rectify_c3_kernel<1920>;
        else if (dw == 1280) kern = rectify_c3_kernel<1280>;
        else if (dw == 640) kern = rectify_c3_kernel<640>;
    }
    TI_CUDA(ctx, ensure_dynamic_smem(kern, smem, ctx->device));
    int per_sm = resident_ctas(kern, C3_THREADS, smem, 3);
    if (ctx->ctas_per_sm > 0) per_sm = ctx->ctas_per_sm;
    const uint64_t grid_max = (uint64_t)ctx->sm_count * per_sm;
    // frames per unit: every unit costs one LUT fetch + expansion (measured: about one frame of work) and the CTAs of
    // an SM share its throughput, so what counts is few units - as long as every CTA still gets about four of them
    int fpu = ctx->frames_per_unit4;
    if (fpu <= 0) {
        const uint64_t want_units = 4 * grid_max;
        const int chunks = (int)std::max<uint64_t>(1, std::min<uint64_t>((want_units + P.tiles_per_set - 1) / P.tiles_per_set,
                                                                        (uint64_t)std::max(1, P.n_batch / 8)));
        fpu = (P.n_batch + chunks - 1) / chunks;
    }
    P.frames_per_unit = std::max(1, std::min(P.n_batch, fpu));
    const uint64_t total = (uint64_t)P.tiles_per_set * ((P.n_batch + P.frames_per_unit - 1) / P.frames_per_unit);
    const int grid = (int)std::min<uint64_t>(total, grid_max);
    TI_LAUNCH(kern, grid, C3_THREADS, smem, ctx->stream, P);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

// ---- tables (host, on the first 3-channel rectify of a slot) -----------------------------------------
void free_c3_tables(CameraSlot& C) {
    if (C.d_lut5) cudaFree(C.d_lut5);
    if (C.d_boxes5) cudaFree(C.d_boxes5);
    C.d_lut5 = nullptr; C.d_boxes5 = nullptr;
    C.has_c3 = false; C.c3_tried = false; C.rows5_alloc = 0;
}

int build_c3_tables(ti_ctx* ctx, CameraSlot& C) {
    C.c3_tried = true;
    C.has_c3 = false;
    if (C.src_w % 16 != 0 || C.h_lut.empty()) return TI_OK;  // TMA row pitch 3 * src_w must be a multiple of 16 bytes
    const int dst_w = C.dst_w, dst_h = C.dst_h, lut_pitch = C.h_lut_pitch;
    const int tx_n = (dst_w + C3_TW - 1) / C3_TW, ty_n = (dst_h + C3_TH - 1) / C3_TH;
    const size_t n_tiles = (size_t)tx_n * ty_n;
    auto entry = [&](int u, int v) -> lut_t { return (u < dst_w && v < dst_h) ? C.h_lut[(size_t)v * lut_pitch + u] : LUT_OUTSIDE; };
    std::vector<TileBox2> boxes(n_tiles);
    std::vector<uint32_t> lut5(n_tiles * C3_TH * C3_LUT_ROW_WORDS, 0u);
    int rows_max = 0;
    // bytes per staged source row: the widest tile decides (a 2 x downscale map needs the wide boxes)
    int span_max = 0;
    for (int ty = 0; ty < ty_n; ++ty)
        for (int tx = 0; tx < tx_n; ++tx) {
            int bx0 = 1 << 20, bx1 = -(1 << 20);
            for (int v = ty * C3_TH; v < std::min(dst_h, (ty + 1) * C3_TH); ++v)
                for (int u = tx * C3_TW; u < std::min(dst_w, (tx + 1) * C3_TW); ++u) {
                    const lut_t e = entry(u, v);
                    if (e == LUT_OUTSIDE) continue;
                    bx0 = std::min(bx0, lut_x0(e)); bx1 = std::max(bx1, lut_x0(e) + 2);
                }
            if (bx1 > bx0) span_max = std::max(span_max, 3 * bx1 - ((3 * bx0) & ~15));
        }
    if (span_max > C3_PITCH_WIDE) return TI_OK;  // not eligible: generic kernels
    const int pitch = span_max > C3_PITCH ? C3_PITCH_WIDE : C3_PITCH;
    for (int ty = 0; ty < ty_n; ++ty)
        for (int tx = 0; tx < tx_n; ++tx) {
            int bx0 = 1 << 20, by0 = 1 << 20, bx1 = -(1 << 20), by1 = -(1 << 20);
            for (int v = ty * C3_TH; v < std::min(dst_h, (ty + 1) * C3_TH); ++v)
                for (int u = tx * C3_TW; u < std::min(dst_w, (tx + 1) * C3_TW); ++u) {
                    const lut_t e = entry(u, v);
                    if (e == LUT_OUTSIDE) continue;
                    const int x0 = lut_x0(e), y0 = lut_y0(e);
                    bx0 = std::min(bx0, x0); by0 = std::min(by0, y0); bx1 = std::max(bx1, x0 + 2); by1 = std::max(by1, y0 + 2);
                }
            const size_t tile = (size_t)ty * tx_n + tx;
            TileBox2& B = boxes[tile];
            B = TileBox2{0, 0, 0, 0, (int16_t)(tx * C3_TW), (int16_t)(ty * C3_TH), 0, 0};
            if (bx1 <= bx0) continue;
            const int c0 = (3 * bx0) & ~15;  // byte column of the box start (floor to 16: -3 -> -16)
            if (3 * bx1 - c0 > pitch || by1 - by0 > C3_MAX_ROWS) return TI_OK;  // not eligible: generic kernels
            B.c0 = (int16_t)c0; B.y0 = (int16_t)by0; B.nvec = (int16_t)((3 * bx1 - c0 + 15) / 16); B.rows = (int16_t)(by1 - by0);
            rows_max = std::max(rows_max, by1 - by0);
            for (int row = 0; row < C3_TH; ++row)
                for (int lu = 0; lu < C3_TW; ++lu) {
                    const lut_t e = entry(tx * C3_TW + lu, ty * C3_TH + row);
                    if (e == LUT_OUTSIDE) continue;  // {0, 0}: zero weights
                    const int x0 = lut_x0(e), y0 = lut_y0(e);
                    const uint32_t fx = lut_fx(e), fy = lut_fy(e);
                    const int bp = 3 * x0 - c0, wordx = bp & ~3, s = bp & 3;
                    const uint32_t off = (uint32_t)((y0 - by0) * pitch + wordx);
                    // lane (lu % 32) blends pixel lu as its (lu / 32)-th: its eight LUT words are contiguous
                    uint32_t* w = lut5.data() + (tile * C3_TH + row) * C3_LUT_ROW_WORDS + (size_t)((lu & 31) * 4 + (lu >> 5)) * 2;
                    w[0] = (off << 16) | (uint32_t)(8 * s);
                    w[1] = (32u - fx) | (fy << 6) | ((fx == 0 && fy == 0 ? 1u : 0u) << 11) | (fx << 16);  // see p4_expand
                }
        }
    const int rows_alloc = std::max(8, (rows_max + 7) / 8 * 8);
    TI_CUDA(ctx, cudaMalloc(&C.d_lut5, lut5.size() * sizeof(uint32_t)));
    TI_CUDA(ctx, cudaMalloc(&C.d_boxes5, boxes.size() * sizeof(TileBox2)));
    TI_CUDA(ctx, cudaMemcpy(C.d_lut5, lut5.data(), lut5.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    TI_CUDA(ctx, cudaMemcpy(C.d_boxes5, boxes.data(), boxes.size() * sizeof(TileBox2), cudaMemcpyHostToDevice));
    C.tiles5_x = tx_n; C.tiles5_y = ty_n; C.rows5_alloc = rows_alloc; C.pitch5 = pitch;
    C.has_c3 = true;
    return TI_OK;
}

}  // namespace ti
