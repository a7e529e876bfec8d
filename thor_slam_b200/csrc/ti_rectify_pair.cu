// Pair-window mono remap ("v4"), bit-exact with cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) on u8.
//
// Takes over the undistortion the reference leaves to cuVSLAM (thor_slam/slam/adapters/isaac_ros.py:364-411,
// rectified_images:=false in Makefile:77-80).
//
// What changed against the v3 kernel (ti_rectify.cu): the 1-byte-shifted second copy of the source box,
// the two builder warps that made it and the per-pixel fixed-point arithmetic are gone.
//  * A tile's source box lands by ONE TMA load (pitch P4_PITCH, out-of-image bytes = 0 = BORDER_CONSTANT)
//    and is used as it lies.
//  * Two horizontally adjacent output pixels (a, b) share one WINDOW: the two aligned 32-bit words
//    that contain source byte x0(a) of row y0, and the two words below them.  One PRMT per row picks
//    the four bytes (a.left, a.right, b.left, b.right) out of the 8-byte window - the selector is
//    precomputed per pair - so a pair costs 4 LDS.32 + 2 PRMT whatever the alignment of x0.
//  * The whole bilinear blend of a pixel is two IDP.2A: the LUT holds the four 2-D weights
//    64*(32-fx)(32-fy), 64*fx(32-fy), 64*(32-fx)fy, 64*fx*fy as two words of 16-bit halves, so
//    R = dp2a(Wbot, bottom bytes, dp2a(Wtop, top bytes, 32768)) = 64*(S + 512) and the output pixel
//    (S + 512) >> 10 is byte 2 of R.  The one weight that does not fit 16 bits (65536 when
//    fx = fy = 0) is stored as 65535: byte 2 of 65535*p + 32768 is still p for every p <= 255.
//  * Pairs whose pixel b does not lie in a's window (different source row, or more than 6 bytes to
//    the right of the window start: ~0.5 % of the pairs of a stereo rectification map) are
//    EXCEPTIONS.  The row loop does not know about them: it blends every pair the same way (the
//    window word of an exception still serves pixel a; what it produces for pixel b is garbage).
//    Afterwards each consumer warp runs ONE fix-up pass per frame: lane i takes entry i of the warp's
//    exception list for the tile (window word, the two weight words, destination column / row),
//    blends that single pixel from its own window and overwrites the byte (ordered behind the row
//    loop's stores by __syncwarp).  The cost per frame is flat - one short pass whatever the number
//    of exceptions - so the eight warps that share a stage stay in step.  The 33rd and further exceptions of a
//    (tile, warp) go to the slot's overflow list and are repaired after the kernel (rectify_points_kernel,
//    ti_rectify.cu); only cameras with source boxes wider than P4_PITCH_WIDE keep using the v3 / v2 kernels.
//  * Slots whose tiles span more than P4_PITCH source bytes (a 2 x downscale map: 128 output pixels sample 256 source pixels)
//    run the PITCH = P4_PITCH_WIDE instantiation: the box is loaded as 80 32-bit elements per row (a TMA box dimension holds at
//    most 256 elements), everything else is the same code.
//  * QUAD layout (round 2; the default where a map allows it): a lane owns FOUR consecutive output pixels whose taps all lie in
//    one 8-byte window per source row (at scale ~1 the four pixels span 4-5 source bytes; the window starts at the 4-byte word
//    of the leftmost) - half the window loads of the pair layout and one 32-bit store per row instead of two 16-bit ones.  The
//    LUT slot of a lane is the same six words {window word + selector of pixels 0-1, pixel 0, pixel 1, selector of pixels 2-3,
//    pixel 2, pixel 3}; pixels that do not fit (another source row, further right) are exceptions like pixel b above - three
//    times as many as with pairs, which the flat fix-up pass does not care about.  Slots with more than the lists hold fall
//    back to the pair layout.
//  * The LUT is 6 bytes per output pixel in memory (per pair: the window word and one word per pixel
//    holding 32-fx, fy, fx and the fx = fy = 0 flag); a tile's 24 KB slice is prefetched into shared
//    memory by a 1-D bulk copy one unit ahead (unit = tile x up to frames_per_unit frames of the
//    batch), expanded ONCE per unit into the four 2-D weight words per pair, which then stay in
//    registers for every frame of the unit.
//
// Shared memory: [full[8] | empty[8] | lut_full | lut_empty mbarriers, 256 B], `stages` boxes of
// rows_alloc_max x P4_PITCH bytes, the LUT slice of the current / next unit (TH x P4_LUT_ROW_WORDS
// words), two exception tables (8 warps x exc_max x 16 bytes; units alternate between them).
#include "ti_rectify_pair.cuh"
#include "ti_pair_dev.cuh"

#include <cstring>

namespace ti {

struct Taps { uint32_t t0, t1, b0, b1; };

// Unit k of this CTA: which job / tile, which frames of the batch.  Issuer and consumers walk the same list.
struct P4Unit { int j; uint32_t tile, b0, nb; };
__device__ __forceinline__ P4Unit p4_unit(const Rect4Params& P, uint32_t k) {
    const uint64_t ug = (uint64_t)blockIdx.x + (uint64_t)k * gridDim.x;
    const uint32_t c = (uint32_t)(ug / P.tiles_per_set), r = (uint32_t)(ug - (uint64_t)c * P.tiles_per_set);
    int j = 0;
    while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
    P4Unit U;
    U.j = j; U.tile = r - P.job[j].tile_begin;
    U.b0 = c * (uint32_t)P.frames_per_unit;
    U.nb = min((uint32_t)P.frames_per_unit, (uint32_t)P.n_batch - U.b0);
    return U;
}

template <int PITCH>
__device__ __forceinline__ Taps p4_fetch(p4_addr_t base, uint32_t window_word) {
    const p4_addr_t a = base + (window_word >> 16);
    Taps T;
    T.t0 = p4_lds<0>(a); T.t1 = p4_lds<4>(a); T.b0 = p4_lds<PITCH>(a); T.b1 = p4_lds<PITCH + 4>(a);
    return T;
}

// both pixels of a pair from the window `T`: 64 * (S + 512) of pixel a / pixel b (the pixel is byte 2).
__device__ __forceinline__ void p4_blend(const Taps& T, const uint4& w, uint32_t m, uint32_t& ra, uint32_t& rb) {
    const uint32_t wt = p4_prmt(T.t0, T.t1, m), wb = p4_prmt(T.b0, T.b1, m);
    ra = __dp2a_lo(w.y, wb, __dp2a_lo(w.x, wt, 32768u));
    rb = __dp2a_hi(w.w, wb, __dp2a_hi(w.z, wt, 32768u));
}

// One frame of one tile for one consumer warp: RPW rows x (2 pairs per lane), then the fix-up pass.
// WHOLE: every pixel of the warp's rows exists and rows start on even addresses (16-bit stores).
// PREFETCH: software pipeline, the windows of row q+1 are in flight while row q is blended and stored (8 more registers).
// fixes (warp-uniform): some lane has an exception entry; i_fix: this lane has one, at shared address my_exc =
// {window word, Wtop, Wbot, row * dst_w + column - 2 * lane}.  Its window is fetched ahead of the last row's blend so that the
// latency of the dependent loads hides behind that row; its store follows the row loop's stores (__syncwarp).
// DSTW > 0: the destination row pitch is this compile-time constant, so row q is an immediate offset from the lane's first-row pointer.
template <int RPW, bool WHOLE, bool PREFETCH, bool HALF_LOADS = false, int DSTW = 0, int PITCH = P4_PITCH, bool QUAD = false, int FIXQ = RPW - 1>
__device__ __forceinline__ void p4_rows(const uint4 (&w0)[RPW], const uint4 (&w1)[RPW], const uint2 (&mw)[RPW], p4_addr_t base,
                                        uint8_t* dp, int dst_w, int live_rows, int live_cols, int lane, bool fixes, bool i_fix,
                                        p4_addr_t my_exc) {
    // HALF_LOADS (bring-up only, wrong pixels): pair 1 reuses pair 0's window - how much of the time is the window loads?
    // QUAD: pixels 2-3 of the lane read pixel 0-1's window (mw[q].y is only their selector)
    Taps A = p4_fetch<PITCH>(base, mw[0].x), B = (HALF_LOADS || QUAD) ? A : p4_fetch<PITCH>(base, mw[0].y);
    uint8_t* const lane_dst = dp;  // this lane's first pixel of the warp's first row
    uint4 fe = make_uint4(0u, 0u, 0u, 0u);
    Taps X = A;
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        Taps An = A, Bn = B;
        if (PREFETCH && q + 1 < RPW) { An = p4_fetch<PITCH>(base, mw[q + 1].x); Bn = (HALF_LOADS || QUAD) ? An : p4_fetch<PITCH>(base, mw[q + 1].y); }
        if (q == FIXQ && fixes && i_fix) {
            fe = p4_lds128(my_exc);
            X = p4_fetch<PITCH>(base, fe.x);
        }
        uint32_t ra0, rb0, ra1, rb1;
        p4_blend(A, w0[q], mw[q].x, ra0, rb0);
        p4_blend(B, w1[q], mw[q].y, ra1, rb1);
        const uint32_t o0 = __byte_perm(ra0, rb0, 0x0062), o1 = __byte_perm(ra1, rb1, 0x0062);
        if (QUAD) {
            const uint32_t o = __byte_perm(o0, o1, 0x5410);  // the lane's four pixels
            if (WHOLE && DSTW > 0) {
                st_stream_b32(dp + q * DSTW, o);
            } else if (WHOLE) {
                st_stream_b32(dp, o);
            } else if (q < live_rows) {
                const int c = 4 * lane;
                if (c < live_cols) st_stream_b8(dp, o);
                if (c + 1 < live_cols) st_stream_b8(dp + 1, o >> 8);
                if (c + 2 < live_cols) st_stream_b8(dp + 2, o >> 16);
                if (c + 3 < live_cols) st_stream_b8(dp + 3, o >> 24);
            }
        } else if (WHOLE && DSTW > 0) {
            st_stream_b16(dp + q * DSTW, o0);
            st_stream_b16(dp + q * DSTW + 64, o1);
        } else if (WHOLE) {
            st_stream_b16(dp, o0);
            st_stream_b16(dp + 64, o1);
        } else if (q < live_rows) {
            const int c0 = 2 * lane, c1 = 64 + 2 * lane;
            if (c0 < live_cols) st_stream_b8(dp, o0);
            if (c0 + 1 < live_cols) st_stream_b8(dp + 1, o0 >> 8);
            if (c1 < live_cols) st_stream_b8(dp + 64, o1);
            if (c1 + 1 < live_cols) st_stream_b8(dp + 65, o1 >> 8);
        }
        if (!(WHOLE && DSTW > 0)) dp += dst_w;
        if (PREFETCH) { A = An; B = Bn; }
        else if (q + 1 < RPW) { A = p4_fetch<PITCH>(base, mw[q + 1].x); B = QUAD ? A : p4_fetch<PITCH>(base, mw[q + 1].y); }
    }
    if (fixes) {
        __syncwarp();  // the fix-up stores land behind the row loop's stores to the same bytes
        if (i_fix) {
            const uint32_t xt = p4_prmt(X.t0, X.t1, fe.x), xb = p4_prmt(X.b0, X.b1, fe.x);
            const uint32_t r = __dp2a_lo(fe.z, xb, __dp2a_lo(fe.y, xt, 32768u));
            st_stream_b8(lane_dst + (int32_t)fe.w, r >> 16);  // offset relative to the repairing lane's own first pixel
        }
    }
}

template <int TH, bool DEBUG, int DSTW = 0, int PITCH = P4_PITCH, bool QUAD = false>
__global__ void __launch_bounds__(P4_THREADS, TH == 32 ? 3 : 4) rectify_mono_pair_kernel(const __grid_constant__ Rect4Params P) {
    TI_DYNAMIC_SMEM(uint8_t, smem);
    constexpr int RPW = TH / P4_CONSUMER_WARPS;  // tile rows per consumer warp
    constexpr uint32_t LUT_BYTES = (uint32_t)TH * P4_LUT_ROW_WORDS * 4u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t stage_bytes = (uint32_t)P.rows_alloc_max * PITCH;  // rows_alloc is even: 128-byte granular for both pitches
    const uint32_t exc_buf_bytes = (uint32_t)P.exc_max * (P4_CONSUMER_WARPS * 16);
    const int S = P.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [S] box landed
    uint64_t* empty = full + P4_MAX_STAGES;              // [S] consumers done
    uint64_t* lut_full = empty + P4_MAX_STAGES;          // LUT slice of the unit landed
    uint64_t* lut_empty = lut_full + 1;                  // every consumer warp has expanded its part of the slice
    uint8_t* stage0 = smem + 256;
    uint8_t* lutbuf = stage0 + (size_t)S * stage_bytes;
    uint8_t* excbuf = lutbuf + LUT_BYTES;  // two tables of exc_buf_bytes

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, P4_CONSUMER_WARPS);
        }
        mbar_init(lut_full, 1);
        mbar_init(lut_empty, P4_CONSUMER_WARPS);
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t n_chunks = (uint32_t)((P.n_batch + P.frames_per_unit - 1) / P.frames_per_unit);
    const uint64_t total_units = (uint64_t)P.tiles_per_set * n_chunks;
    const uint32_t units_mine = total_units > blockIdx.x ? (uint32_t)((total_units - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    if (units_mine == 0) return;

    if (warp == P4_CONSUMER_WARPS) {
        // ------------------------------------------------ issuer (one thread) ------------------------
        if (lane != 0) return;
        struct Unit { int j; uint32_t tile, b0, nb; uint4 box; };
        auto load_unit = [&](uint32_t k, Unit& U) {
            const P4Unit u = p4_unit(P, k);
            U.j = u.j; U.tile = u.tile; U.b0 = u.b0; U.nb = u.nb;
            U.box = *reinterpret_cast<const uint4*>(P.job[u.j].boxes4 + u.tile);
        };
        auto issue_lut = [&](const Unit& U, uint32_t k) {  // LUT slice + exception table of unit k
            const Rect4JobDev& J = P.job[U.j];
            const uint32_t eb = (uint32_t)J.exc_per_warp * (P4_CONSUMER_WARPS * 16);
            bulk_load_1d(lutbuf, J.lut4 + (size_t)U.tile * (TH * P4_LUT_ROW_WORDS), LUT_BYTES, lut_full);
            if (eb) bulk_load_1d(excbuf + (k & 1u) * exc_buf_bytes, J.exc4 + (size_t)U.tile * (eb / 4), eb, lut_full);
            mbar_arrive_expect_tx(lut_full, LUT_BYTES + eb);
        };
        Unit cur{}, nxt{};
        load_unit(0, cur);
        issue_lut(cur, 0);
        if (units_mine > 1) load_unit(1, nxt);
        int s = 0;
        uint32_t phase = 1;  // parity to wait for on empty[s]: 1 on a stage's first use (passes at once)
        for (uint32_t k = 0; k < units_mine; ++k) {
            const Rect4JobDev& J = P.job[cur.j];
            const int c0 = (int16_t)(cur.box.x & 0xFFFF), y0 = (int16_t)(cur.box.x >> 16);
            const int rows = (int16_t)(cur.box.y >> 16);
            const uint32_t tx = rows > 0 ? (uint32_t)J.rows_alloc * PITCH : 0u;
            bool lut_pending = k + 1 < units_mine;  // the next unit's LUT slice still has to be requested
            for (uint32_t f = 0; f < cur.nb; ++f) {
                const uint32_t b = cur.b0 + f;
                uint8_t* sb = stage0 + (size_t)s * stage_bytes;
                // consumers have released the stage's previous item.  Suspended in hardware: polling every 300 ns cost 8 % of the
                // kernel's issued instructions (ncu, profiles/r02_ncu_rect_quad.txt) - 1 % of its time
                mbar_wait_hint(empty + s, phase, 20000);
                if (rows > 0 && !(DEBUG && (P.debug & 2))) tma_load_3d(sb, &P.map[cur.j], PITCH == P4_PITCH ? c0 : c0 / 4, y0, (int)b, full + s);  // one box: PITCH x rows_alloc bytes (wide: 32-bit elements)
                mbar_arrive_expect_tx(full + s, (DEBUG && (P.debug & 2)) ? 0u : tx);
                if (++s == S) { s = 0; phase ^= 1u; }
                // the LUT buffer is free again once all consumer warps hold unit k in registers
                if (lut_pending && mbar_test(lut_empty, k & 1u)) { issue_lut(nxt, k + 1); lut_pending = false; }
            }
            if (lut_pending) { mbar_wait_relaxed(lut_empty, k & 1u); issue_lut(nxt, k + 1); }
            cur = nxt;
            if (k + 2 < units_mine) load_unit(k + 2, nxt);
        }
        return;
    }
    // ---------------------------------------------------- consumers ---------------------------------
    // Lane L owns, in each of its warp's RPW tile rows, pair 0 = pixels (2L, 2L+1) and pair 1 =
    // pixels (64+2L, 65+2L): the 32 windows of one load instruction cover ~64 consecutive source
    // bytes (17 words: conflict-free), stores are 64 contiguous bytes per warp instruction.
    uint4 w0[RPW], w1[RPW];  // weights of pair 0 / pair 1: {Wtop(a), Wbot(a), Wtop(b), Wbot(b)}
    uint2 mw[RPW];           // window words of pair 0 / pair 1
    const p4_addr_t sm0 = p4_addr(smem);
    const p4_addr_t stage_first = sm0 + 256;
    const bool skip_blend = DEBUG && (P.debug & 1) != 0;
    p4_addr_t base = stage_first, bar = sm0;  // current stage, its `full` barrier (`empty` is 64 bytes further)
    int s_left = S;                            // stages until the ring wraps
    uint32_t phase = 0;
    for (uint32_t k = 0; k < units_mine; ++k) {
        const P4Unit U = p4_unit(P, k);
        const Rect4JobDev& J = P.job[U.j];
        const uint4 box = *reinterpret_cast<const uint4*>(J.boxes4 + U.tile);
        const int u0 = (int16_t)(box.z & 0xFFFF), v0 = (int16_t)(box.z >> 16) + warp * RPW;
        const int dst_w = J.dst_w, live_rows = J.dst_h - v0, live_cols = J.dst_w - u0;
        const uint64_t dst_stride = J.dst_stride;
        uint8_t* dp = J.dst + (uint64_t)U.b0 * dst_stride + (size_t)v0 * dst_w + u0 + (QUAD ? 4 : 2) * lane;
        const bool whole = live_rows >= RPW && live_cols >= P4_TW &&
                           ((((uint64_t)(uintptr_t)J.dst | dst_stride | (uint64_t)dst_w) & (QUAD ? 3ull : 1ull)) == 0);  // warp-uniform: aligned row stores
        // expand this lane's part of the unit's LUT slice into weight registers
        mbar_wait(lut_full, k & 1u);
        {
            const uint2* lp = reinterpret_cast<const uint2*>(lutbuf) + ((size_t)(warp * RPW) * 32 + lane) * 3;
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
                const uint2 e0 = lp[q * 96], e1 = lp[q * 96 + 1], e2 = lp[q * 96 + 2];  // {m0, a0}, {b0, m1}, {a1, b1}
                mw[q] = make_uint2(e0.x, e1.y);
                p4_expand(e0.y, w0[q].x, w0[q].y);
                p4_expand(e1.x, w0[q].z, w0[q].w);
                p4_expand(e2.x, w1[q].x, w1[q].y);
                p4_expand(e2.y, w1[q].z, w1[q].w);
            }
        }
        // this lane's exception entry (if any): entries are dense from 0, an unused one has P4_EXC_UNUSED in its last word (NOT all
        // ones: -1 is a legitimate destination offset - the pixel left of the repairing lane's first one)
        const p4_addr_t my_exc = sm0 + 256u + (uint32_t)S * stage_bytes + LUT_BYTES + (k & 1u) * exc_buf_bytes +
                                 (uint32_t)(warp * J.exc_per_warp + lane) * 16u;
        const bool i_fix = lane < J.exc_per_warp && p4_lds<12>(my_exc) != P4_EXC_UNUSED;
        const bool warp_fixes = __ballot_sync(0xFFFFFFFFu, i_fix) != 0u;  // warp-uniform
        __syncwarp();  // every lane's reads of the LUT slice are ordered before the release below
        if (lane == 0) mbar_arrive(lut_empty);

        for (uint32_t f = 0; f < U.nb; ++f) {
            p4_wait(bar, phase);
            if (!skip_blend) {
                if (DEBUG && (P.debug & 8)) p4_rows<RPW, true, TH == 32, true, 0, PITCH, QUAD>(w0, w1, mw, base, dp, dst_w, live_rows, live_cols, lane, warp_fixes, i_fix, my_exc);
                else if (whole) p4_rows<RPW, true, TH == 32, false, DSTW, PITCH, QUAD, QUAD ? RPW - 3 : RPW - 1>(w0, w1, mw, base, dp, dst_w, live_rows, live_cols, lane, warp_fixes, i_fix, my_exc);
                else p4_rows<RPW, false, false, false, 0, PITCH, QUAD>(w0, w1, mw, base, dp, dst_w, live_rows, live_cols, lane, warp_fixes, i_fix, my_exc);
            }
            p4_warp_arrive(bar + 64);
            dp += dst_stride;
            bar += 8; base += stage_bytes;
            if (--s_left == 0) { s_left = S; bar -= 8u * (uint32_t)S; base -= (uint32_t)S * stage_bytes; phase ^= 1u; }
        }
    }
}

// ---- launcher ------------------------------------------------------------------------------------
int launch_rectify_pair(ti_ctx* ctx, Rect4Params& P, int th_index) {
    if (P.n_jobs == 0 || P.n_batch <= 0) return TI_OK;
    const int TH = P4_TILE_HEIGHTS[th_index];
    const size_t stage = (size_t)P.rows_alloc_max * P.pitch;
    const size_t lut_bytes = (size_t)TH * P4_LUT_ROW_WORDS * 4 + 2 * (size_t)P.exc_max * (P4_CONSUMER_WARPS * 16);
    typedef void (*Kern)(const Rect4Params);
    static const Kern kernels[2][P4_N_TH] = {
        {rectify_mono_pair_kernel<16, false>, rectify_mono_pair_kernel<32, false>, rectify_mono_pair_kernel<24, false>},
        {rectify_mono_pair_kernel<16, true>, rectify_mono_pair_kernel<32, true>, rectify_mono_pair_kernel<24, true>}};
    static const Kern kernels_wide[P4_N_TH] = {rectify_mono_pair_kernel<16, false, 0, P4_PITCH_WIDE>, rectify_mono_pair_kernel<32, false, 0, P4_PITCH_WIDE>,
                                               rectify_mono_pair_kernel<24, false, 0, P4_PITCH_WIDE>};
    const bool wide = P.pitch == P4_PITCH_WIDE;
    Kern kern = wide ? kernels_wide[th_index] : kernels[ctx->debug ? 1 : 0][th_index];
    // every job of the launch writes rows of the same common pitch: use the kernel whose row stores are immediate offsets
    int dw = P.job[0].dst_w;
    for (int j = 1; j < P.n_jobs; ++j)
        if (P.job[j].dst_w != dw) dw = 0;
    if (P.quad) {
        if (TH != 32 || wide) return fail(ctx, TI_ESTATE, "rectify: the quad layout exists for 32-row tiles of the standard pitch only");
        kern = ctx->debug  ? (Kern)rectify_mono_pair_kernel<32, true, 0, P4_PITCH, true>
             : dw == 1280 ? (Kern)rectify_mono_pair_kernel<32, false, 1280, P4_PITCH, true>
             : dw == 640  ? (Kern)rectify_mono_pair_kernel<32, false, 640, P4_PITCH, true>
             : dw == 1920 ? (Kern)rectify_mono_pair_kernel<32, false, 1920, P4_PITCH, true>
                          : (Kern)rectify_mono_pair_kernel<32, false, 0, P4_PITCH, true>;
    } else if (!ctx->debug && TH == 32 && !wide) {
        if (dw == 1280) kern = rectify_mono_pair_kernel<32, false, 1280>;
        else if (dw == 640) kern = rectify_mono_pair_kernel<32, false, 640>;
        else if (dw == 1920) kern = rectify_mono_pair_kernel<32, false, 1920>;
    }
    // ring depth: as asked, but never so deep that fewer CTAs fit an SM than the register budget allows
    const int want_ctas = TH == 32 ? 3 : 4;  // = the kernel's __launch_bounds__
    int stages = std::max(2, std::min(ctx->stages4, P4_MAX_STAGES));
    // ... and never so deep that the SM has no shared memory left for anybody else: the exchange kernels (ti_push.cu: a 12 KB TMA
    // copy CTA, one-warp flag kernels, each with its 1 KB of system shared memory) run BESIDE this kernel's resident CTAs.  When
    // they did not fit, whichever came first displaced a CTA of this persistent grid, which then started late and stretched the
    // kernel by a third (measured: 14 us per step on the fusing rank).  Ring depth beyond four stages buys nothing (round 1).
    while (stages > 2 && (256 + (size_t)stages * stage + lut_bytes + 1024) * want_ctas > (228 - ctx->smem_headroom_kb) * 1024) --stages;
    P.stages = stages;
    P.debug = ctx->debug;
    const size_t smem = 256 + (size_t)stages * stage + lut_bytes;
    TI_CUDA(ctx, ensure_dynamic_smem(kern, smem, ctx->device));
    int per_sm = resident_ctas(kern, P4_THREADS, smem, 3);
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
    TI_LAUNCH(kern, grid, P4_THREADS, smem, ctx->stream, P);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

// ---- tables (host, at calibration upload) -----------------------------------------------------------
void free_pair_tables(CameraSlot& C) {
    for (int k = 0; k < P4_N_TH; ++k) {
        if (C.d_lut4[k]) cudaFree(C.d_lut4[k]);
        if (C.d_boxes4[k]) cudaFree(C.d_boxes4[k]);
        if (C.d_exc4[k]) cudaFree(C.d_exc4[k]);
        if (C.d_over4[k]) cudaFree(C.d_over4[k]);
        C.d_lut4[k] = nullptr; C.d_boxes4[k] = nullptr; C.d_exc4[k] = nullptr; C.d_over4[k] = nullptr; C.n_over4[k] = 0;
        C.has_pair[k] = false; C.quad4[k] = false; C.exc4_per_warp[k] = 0; C.rows4_alloc[k] = 0;
    }
}

namespace {
struct Px {  // one output pixel decoded from the generic LUT
    bool in;  // at least one tap inside the source image
    int x0, y0;
    uint32_t fx, fy;
};
inline Px decode(lut_t e) {
    Px p{};
    p.in = e != LUT_OUTSIDE;
    if (p.in) {
        p.x0 = lut_x0(e); p.y0 = lut_y0(e);
        p.fx = lut_fx(e); p.fy = lut_fy(e);
    }
    return p;
}
inline void host_expand(uint32_t e, uint32_t& wtop, uint32_t& wbot) {  // same arithmetic as p4_expand
    const uint32_t pw = e & 0x001F003Fu, fy64 = e & (31u << 6);
    wtop = (2048u - fy64) * pw - ((e >> 11) & 1u);
    wbot = fy64 * pw;
}
inline uint32_t pixel_word(const Px& p) {  // see p4_expand
    if (!p.in) return 0u;
    return (32u - p.fx) | (p.fy << 6) | ((p.fx == 0 && p.fy == 0 ? 1u : 0u) << 11) | (p.fx << 16);
}
}  // namespace

int build_pair_tables(ti_ctx* ctx, CameraSlot& C, const std::vector<lut_t>& lut, int lut_pitch) {
    free_pair_tables(C);
    if (C.src_w % 16 != 0) return TI_OK;  // TMA row pitch must be a multiple of 16 bytes
    const int dst_w = C.dst_w, dst_h = C.dst_h;
    for (int k = 0; k < P4_N_TH; ++k) {
        const int TH = P4_TILE_HEIGHTS[k], RPW = TH / P4_CONSUMER_WARPS;
        const int tx_n = (dst_w + P4_TW - 1) / P4_TW, ty_n = (dst_h + TH - 1) / TH;
        const size_t n_tiles = (size_t)tx_n * ty_n;
        std::vector<TileBox2> boxes(n_tiles);
        auto entry = [&](int u, int v) -> lut_t { return (u < dst_w && v < dst_h) ? lut[(size_t)v * lut_pitch + u] : LUT_OUTSIDE; };
        bool ok = true;
        int rows_max = 0, span_max = 0;
        for (int ty = 0; ty < ty_n && ok; ++ty)
            for (int tx = 0; tx < tx_n && ok; ++tx) {
                int bx0 = 1 << 20, by0 = 1 << 20, bx1 = -(1 << 20), by1 = -(1 << 20);
                for (int v = ty * TH; v < std::min(dst_h, (ty + 1) * TH); ++v)
                    for (int u = tx * P4_TW; u < std::min(dst_w, (tx + 1) * P4_TW); ++u) {
                        const Px p = decode(entry(u, v));
                        if (!p.in) continue;
                        bx0 = std::min(bx0, p.x0); by0 = std::min(by0, p.y0); bx1 = std::max(bx1, p.x0 + 2); by1 = std::max(by1, p.y0 + 2);
                    }
                TileBox2& B = boxes[(size_t)ty * tx_n + tx];
                B = TileBox2{0, 0, 0, 0, (int16_t)(tx * P4_TW), (int16_t)(ty * TH), 0, 0};
                if (bx1 <= bx0) continue;
                // floor to 16 (-1 -> -16): a TMA box must start on a 16-byte boundary of global memory - probed on the B200:
                // starts shifted by 1, 4 or 7 bytes end in "an illegal instruction was encountered"
                const int c0 = bx0 & ~15;
                if (bx1 - c0 > P4_PITCH_WIDE || by1 - by0 > P4_MAX_ROWS) { ok = false; break; }
                span_max = std::max(span_max, bx1 - c0);
                B.c0 = (int16_t)c0; B.y0 = (int16_t)by0; B.nvec = (int16_t)((bx1 - c0 + 15) / 16); B.rows = (int16_t)(by1 - by0);
                rows_max = std::max(rows_max, by1 - by0);
            }
        if (!ok) continue;
        const int rows_alloc = std::max(8, (rows_max + 1) / 2 * 2);  // a stage (rows x 192 or 320 bytes) must be a multiple of 128 bytes
        const int pitch = span_max > P4_PITCH ? P4_PITCH_WIDE : P4_PITCH;  // staged bytes per source row

        std::vector<uint32_t> lut4;
        std::vector<std::vector<uint32_t>> exc;  // per (tile, warp): 4 words per entry
        std::vector<uint32_t> over;              // output pixels (v * dst_w + u) whose warp list was full
        // One exception: pixel `q` at output column `col` of tile row `row` gets an entry of its warp's list - its own window (bytes
        // left, right selected into bytes 0, 1), its weight words, where it goes.  A full list (a strongly bent map: a fisheye4 camera
        // at 1280 x 800 has up to 42 such pixels in one (tile, warp) with pairs) sends the pixel to the slot's OVERFLOW list, repaired
        // after the kernel by one thread per pixel and frame from the generic LUT (rectify_points_kernel) - a handful of pixels per
        // frame, so every reference camera model stays on this kernel.  False: beyond ~3 % the per-pixel pass costs more than v3.
        auto exception = [&](std::vector<uint32_t>& ex, const TileBox2& B, const Px& q, int row, int col, int v, int u, int lane_px) -> bool {
            if (ex.size() / 4 >= (size_t)(lane_px == 4 ? ctx->quad_exc_cap : P4_MAX_EXC)) {
                over.push_back((uint32_t)v * (uint32_t)dst_w + (uint32_t)u);
                return over.size() <= (size_t)dst_w * dst_h / 32;
            }
            const int wordq = (q.x0 - B.c0) & ~3;
            const uint32_t offq = (uint32_t)((q.y0 - B.y0) * pitch + wordq);
            const uint32_t sq = (uint32_t)(q.x0 - B.c0 - wordq);
            uint32_t wt, wb;
            host_expand(pixel_word(q), wt, wb);
            ex.push_back((offq << 16) | sq | ((sq + 1) << 4) | (sq << 8) | ((sq + 1) << 12));
            ex.push_back(wt);
            ex.push_back(wb);
            // destination of the pixel relative to the first pixel of the lane that will repair it (entry i -> lane i)
            const int fix_lane = (int)(ex.size() / 4);
            ex.push_back((uint32_t)((row % RPW) * dst_w + col - lane_px * fix_lane));
            return true;
        };
        // pair layout: lane L owns pixels (2L, 2L+1) and (64+2L, 65+2L) of a tile row, a window per pair
        auto fill_pairs = [&]() -> bool {
            lut4.assign(n_tiles * TH * P4_LUT_ROW_WORDS, 0u);
            exc.assign(n_tiles * P4_CONSUMER_WARPS, std::vector<uint32_t>());
            over.clear();
            for (int ty = 0; ty < ty_n; ++ty)
                for (int tx = 0; tx < tx_n; ++tx) {
                    const size_t tile = (size_t)ty * tx_n + tx;
                    const TileBox2& B = boxes[tile];
                    for (int row = 0; row < TH; ++row) {
                        uint32_t* rw = lut4.data() + (tile * TH + row) * P4_LUT_ROW_WORDS;
                        std::vector<uint32_t>& ex = exc[tile * P4_CONSUMER_WARPS + row / RPW];
                        for (int lane = 0; lane < 32; ++lane)
                            for (int p = 0; p < 2; ++p) {
                                const int ua = tx * P4_TW + p * 64 + 2 * lane, v = ty * TH + row;
                                const Px a = decode(entry(ua, v)), b = decode(entry(ua + 1, v));
                                uint32_t* w = rw + (lane * 2 + p) * 3;  // {window word, pixel a, pixel b}
                                w[1] = pixel_word(a);
                                w[2] = pixel_word(b);
                                uint32_t& m = w[0];
                                m = 0x3210u;
                                if (!a.in && !b.in) continue;
                                const Px& anchor = a.in ? a : b;  // the window is placed for pixel a (for b when a has no tap inside)
                                const int wordx = (anchor.x0 - B.c0) & ~3;
                                const uint32_t off = (uint32_t)((anchor.y0 - B.y0) * pitch + wordx);
                                const uint32_t sa = a.in ? (uint32_t)(a.x0 - B.c0 - wordx) : 0u;
                                const int sb_rel = b.in ? b.x0 - B.c0 - wordx : (int)sa;
                                const bool b_fits = !b.in || (b.y0 == anchor.y0 && sb_rel >= 0 && sb_rel <= 6);
                                if (b_fits) {
                                    m = (off << 16) | sa | ((sa + 1) << 4) | ((uint32_t)sb_rel << 8) | ((uint32_t)(sb_rel + 1) << 12);
                                } else {
                                    m = (off << 16) | sa | ((sa + 1) << 4) | (sa << 8) | ((sa + 1) << 12);  // b reads a's bytes: harmless
                                    if (!exception(ex, B, b, row, p * 64 + 2 * lane + 1, v, ua + 1, 2)) return false;
                                }
                            }
                    }
                }
            return true;
        };
        // quad layout: lane L owns pixels 4L .. 4L+3 of a tile row, ONE window for the four (the slot's second window word only carries
        // the byte selector of pixels 2-3).  The window is placed for the source row most of the four read (ties: the leftmost pixel's)
        // at the 4-byte word of that row's leftmost tap; whoever reads another row or lies beyond byte 6 is an exception.
        auto fill_quads = [&]() -> bool {
            lut4.assign(n_tiles * TH * P4_LUT_ROW_WORDS, 0u);
            exc.assign(n_tiles * P4_CONSUMER_WARPS, std::vector<uint32_t>());
            over.clear();
            for (int ty = 0; ty < ty_n; ++ty)
                for (int tx = 0; tx < tx_n; ++tx) {
                    const size_t tile = (size_t)ty * tx_n + tx;
                    const TileBox2& B = boxes[tile];
                    for (int row = 0; row < TH; ++row) {
                        uint32_t* rw = lut4.data() + (tile * TH + row) * P4_LUT_ROW_WORDS;
                        std::vector<uint32_t>& ex = exc[tile * P4_CONSUMER_WARPS + row / RPW];
                        for (int lane = 0; lane < 32; ++lane) {
                            const int u0 = tx * P4_TW + 4 * lane, v = ty * TH + row;
                            Px q[4];
                            for (int i = 0; i < 4; ++i) q[i] = decode(entry(u0 + i, v));
                            uint32_t* w = rw + lane * 6;  // {window word + selector 0-1, pixel 0, pixel 1, selector 2-3, pixel 2, pixel 3}
                            w[0] = w[3] = 0x3210u;
                            w[1] = pixel_word(q[0]); w[2] = pixel_word(q[1]); w[4] = pixel_word(q[2]); w[5] = pixel_word(q[3]);
                            int best = -1, best_n = 0;
                            for (int i = 0; i < 4; ++i) {
                                if (!q[i].in) continue;
                                int n = 0;
                                for (int j = 0; j < 4; ++j) n += q[j].in && q[j].y0 == q[i].y0;
                                if (n > best_n) { best_n = n; best = i; }
                            }
                            if (best < 0) continue;
                            const int wy = q[best].y0;
                            int xmin = 1 << 20;
                            for (int i = 0; i < 4; ++i)
                                if (q[i].in && q[i].y0 == wy) xmin = std::min(xmin, q[i].x0);
                            const int wordx = (xmin - B.c0) & ~3;
                            const uint32_t off = (uint32_t)((wy - B.y0) * pitch + wordx);
                            uint32_t sel[4];
                            for (int i = 0; i < 4; ++i) {
                                const int rel = q[i].in ? q[i].x0 - B.c0 - wordx : 0;
                                const bool fits = !q[i].in || (q[i].y0 == wy && rel >= 0 && rel <= 6);
                                sel[i] = fits ? (uint32_t)rel : 0u;  // an exception reads bytes 0, 1 of the window: harmless
                                if (!fits && !exception(ex, B, q[i], row, 4 * lane + i, v, u0 + i, 4)) return false;
                            }
                            w[0] = (off << 16) | sel[0] | ((sel[0] + 1) << 4) | (sel[1] << 8) | ((sel[1] + 1) << 12);
                            w[3] = sel[2] | ((sel[2] + 1) << 4) | (sel[3] << 8) | ((sel[3] + 1) << 12);
                        }
                    }
                }
            return true;
        };
        // Quads where nearly all their exceptions fit the lists: a strongly bent map (fisheye4 at 1280 x 800: 22 000 overflow pixels with
        // quads, a few hundred with pairs) is faster with pairs (measured 0.46 against 0.57 of the copy peak).  The per-pixel pass costs
        // about 9 ps per overflow pixel and frame, quads save about 3.6 ns per 1 MPix frame: break-even near 0.6 % of the image.
        bool quad = false;
        if (TH == 32 && pitch == P4_PITCH && ctx->rectify_quad) quad = fill_quads() && over.size() <= (size_t)dst_w * dst_h / 256;
        if (!quad) ok = fill_pairs();
        if (!ok) continue;
        size_t e_max = 0;
        for (const auto& e : exc) e_max = std::max(e_max, e.size() / 4);
        const int epw = (int)e_max;
        std::vector<uint32_t> exc_flat(std::max<size_t>(4, n_tiles * P4_CONSUMER_WARPS * epw * 4), P4_EXC_UNUSED);
        for (size_t i = 0; i < exc.size(); ++i)
            if (!exc[i].empty()) std::memcpy(exc_flat.data() + i * epw * 4, exc[i].data(), exc[i].size() * sizeof(uint32_t));

        TI_CUDA(ctx, cudaMalloc(&C.d_lut4[k], lut4.size() * sizeof(uint32_t)));
        TI_CUDA(ctx, cudaMalloc(&C.d_boxes4[k], boxes.size() * sizeof(TileBox2)));
        TI_CUDA(ctx, cudaMalloc(&C.d_exc4[k], exc_flat.size() * sizeof(uint32_t)));
        TI_CUDA(ctx, cudaMemcpy(C.d_lut4[k], lut4.data(), lut4.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        TI_CUDA(ctx, cudaMemcpy(C.d_boxes4[k], boxes.data(), boxes.size() * sizeof(TileBox2), cudaMemcpyHostToDevice));
        TI_CUDA(ctx, cudaMemcpy(C.d_exc4[k], exc_flat.data(), exc_flat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (!over.empty()) {
            TI_CUDA(ctx, cudaMalloc(&C.d_over4[k], over.size() * sizeof(uint32_t)));
            TI_CUDA(ctx, cudaMemcpy(C.d_over4[k], over.data(), over.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        C.n_over4[k] = (int)over.size();
        C.tiles4_x[k] = tx_n; C.tiles4_y[k] = ty_n; C.rows4_alloc[k] = rows_alloc; C.pitch4[k] = pitch; C.quad4[k] = quad; C.exc4_per_warp[k] = epw;
        C.has_pair[k] = true;
    }
    return TI_OK;
}

}  // namespace ti
