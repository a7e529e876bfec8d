// C ABI of libthoringest.so: context, calibration upload, per-frame entry points, host pipeline,
// NCCL gather and peer buffers.  See include/thoringest.h for the contract.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "ti_common.cuh"

namespace ti {

static std::mutex g_err_mu;
static std::string g_err;

void set_global_error(const char* msg) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = msg;
}

int fail(ti_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else set_global_error(buf);
    return code;
}

static void free_camera(CameraSlot& c) {
    if (c.d_lut) cudaFree(c.d_lut);
    if (c.d_boxes) cudaFree(c.d_boxes);
    if (c.d_valid) cudaFree(c.d_valid);
    if (c.d_lut2) cudaFree(c.d_lut2);
    if (c.d_boxes2) cudaFree(c.d_boxes2);
    for (int k = 0; k < 3; ++k) {
        if (c.d_lut3[k]) cudaFree(c.d_lut3[k]);
        if (c.d_boxes3[k]) cudaFree(c.d_boxes3[k]);
    }
    free_pair_tables(c);
    free_c3_tables(c);
    c = CameraSlot{};
}

// OpenCV's cvRound on float*32: round-half-to-even (SSE cvtss2si under the default MXCSR mode)
static inline int round_half_even(float v) { return (int)lrintf(v); }

}  // namespace ti

using namespace ti;

extern "C" {

int ti_abi_version(void) { return TI_ABI_VERSION; }

int ti_create(int device, ti_ctx** out) {
    if (!out) return fail(nullptr, TI_EINVAL, "ti_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, TI_ECUDA, "ti_create: no CUDA device (%s) - this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(nullptr, TI_EINVAL, "ti_create: device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, TI_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, TI_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, TI_ECUDA, "ti_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device, prop.major, prop.minor);
    ti_ctx* ctx = new (std::nothrow) ti_ctx();
    if (!ctx) return fail(nullptr, TI_ENOMEM, "ti_create: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return TI_OK;
}

int ti_destroy(ti_ctx* ctx) {
    if (!ctx) return TI_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& c : ctx->cams) free_camera(c);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->voxel_table) cudaFree(ctx->voxel_table);
    for (auto& h : ctx->hslot) {
        for (void* p : h.d_src) if (p) cudaFree(p);
        for (void* p : h.d_dst) if (p) cudaFree(p);
        for (void* p : h.d_mask) if (p) cudaFree(p);
        for (uint32_t* p : h.d_count) if (p) cudaFree(p);
        if (h.h2d_done) cudaEventDestroy(h.h2d_done);
        if (h.exec_done) cudaEventDestroy(h.exec_done);
        if (h.d2h_done) cudaEventDestroy(h.d2h_done);
    }
    if (ctx->host_ready) {
        for (auto& ev : ctx->ticket_done) cudaEventDestroy(ev);
        cudaStreamDestroy(ctx->s_h2d);
        cudaStreamDestroy(ctx->s_d2h);
        cudaStreamDestroy(ctx->s_exec);
    }
    ti_nccl_teardown(ctx);
    delete ctx;
    return TI_OK;
}

const char* ti_last_error(const ti_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mu);
    static thread_local std::string copy;
    copy = g_err;
    return copy.c_str();
}

int ti_set_stream(ti_ctx* ctx, void* cuda_stream) {
    if (!ctx) return TI_EINVAL;
    ctx->stream = (cudaStream_t)cuda_stream;
    return TI_OK;
}

int ti_sync(ti_ctx* ctx) {
    if (!ctx) return TI_EINVAL;
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TI_OK;
}

int ti_set_option(ti_ctx* ctx, int option, int value) {
    if (!ctx) return TI_EINVAL;
    switch (option) {
        case TI_OPT_FORCE_GENERIC_RECTIFY: ctx->force_generic_rectify = value != 0; return TI_OK;
        case TI_OPT_CTAS_PER_SM: ctx->ctas_per_sm = value; return TI_OK;
        case TI_OPT_DEBUG: ctx->debug = value; return TI_OK;
        case TI_OPT_LUT_PREFETCH: ctx->lut_prefetch = value != 0; return TI_OK;
        case TI_OPT_PUSH_BLOCKS: ctx->push_blocks = value > 0 ? value : 0; return TI_OK;
        case TI_OPT_PUSH_TMA: ctx->push_tma = value != 0; return TI_OK;
        case TI_OPT_SMEM_HEADROOM_KB: ctx->smem_headroom_kb = std::max(0, std::min(value, 128)); return TI_OK;
        case TI_OPT_RECTIFY_QUAD:
            ctx->rectify_quad = value != 0;
            ctx->quad_exc_cap = value >= 2 ? std::min(value, (int)ti::P4_MAX_EXC) : (int)ti::P4_MAX_EXC_QUAD;
            return TI_OK;
        case TI_OPT_L2_SCRATCH_KB: ctx->l2_scratch_kb = value > 0 ? value : 0; return TI_OK;
        case TI_OPT_STAGES:
            if (value < 2 || value > M3_MAX_STAGES) return fail(ctx, TI_EINVAL, "stages must be in [2,%d]", M3_MAX_STAGES);
            ctx->stages = value; ctx->stages4 = value; return TI_OK;
        case TI_OPT_FRAMES_PER_UNIT:
            if (value < 0) return fail(ctx, TI_EINVAL, "frames per unit must be >= 0 (0 = automatic, pair-window kernel only)");
            if (value == 0) { ctx->frames_per_unit4 = 0; return TI_OK; }
            ctx->frames_per_unit = value; ctx->frames_per_unit4 = value; return TI_OK;
        case TI_OPT_MONO_VARIANT:
            if (value < 1 || value > 4) return fail(ctx, TI_EINVAL, "mono variant must be 1, 2, 3 or 4");
            ctx->mono_variant = value; return TI_OK;
        case TI_OPT_TMA_TILE_H:
            if (value != 16 && value != 24 && value != 32) return fail(ctx, TI_EINVAL, "tile height must be 16, 24 or 32");
            ctx->tma_tile_h = value; return TI_OK;
        default: return fail(ctx, TI_EINVAL, "ti_set_option: unknown option %d", option);
    }
}

uint64_t ti_launch_count(const ti_ctx* ctx) { return ctx ? ctx->launches : 0; }
int ti_device_sm_count(const ti_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// ---- calibration ------------------------------------------------------------------------------

int ti_upload_rectify_map(ti_ctx* ctx, int camera, int dst_w, int dst_h, int src_w, int src_h, const float* mapx,
                          const float* mapy) {
    if (!ctx) return TI_EINVAL;
    if (camera < 0 || camera >= TI_MAX_CAMERAS) return fail(ctx, TI_EINVAL, "camera slot %d out of range", camera);
    if (!mapx || !mapy) return fail(ctx, TI_EINVAL, "ti_upload_rectify_map: null map");
    if (dst_w <= 0 || dst_h <= 0 || src_w <= 0 || src_h <= 0 || src_w > TI_MAX_DIM || src_h > TI_MAX_DIM)
        return fail(ctx, TI_EINVAL, "ti_upload_rectify_map: sizes must be in [1,%d] (dst %dx%d, src %dx%d)", TI_MAX_DIM,
                    dst_w, dst_h, src_w, src_h);
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    CameraSlot& C = ctx->cams[camera];
    if (C.d_lut) cudaFree(C.d_lut);
    if (C.d_boxes) cudaFree(C.d_boxes);
    if (C.d_valid) cudaFree(C.d_valid);
    if (C.d_lut2) cudaFree(C.d_lut2);
    if (C.d_boxes2) cudaFree(C.d_boxes2);
    C.d_lut = nullptr; C.d_boxes = nullptr; C.d_valid = nullptr; C.d_lut2 = nullptr; C.d_boxes2 = nullptr;
    C.has_map = false; C.has_fast_mono = false;
    free_pair_tables(C);
    free_c3_tables(C);
    for (int k = 0; k < 3; ++k) {
        if (C.d_lut3[k]) cudaFree(C.d_lut3[k]);
        if (C.d_boxes3[k]) cudaFree(C.d_boxes3[k]);
        C.d_lut3[k] = nullptr; C.d_boxes3[k] = nullptr; C.has_tma_mono[k] = false;
    }

    const int tiles_x = (dst_w + RT_W - 1) / RT_W, tiles_y = (dst_h + RT_H - 1) / RT_H;
    const int lut_pitch = tiles_x * RT_W, lut_rows = tiles_y * RT_H;
    std::vector<lut_t> lut((size_t)lut_pitch * lut_rows, LUT_OUTSIDE);
    std::vector<uint8_t> valid((size_t)dst_w * dst_h, 0);
    std::vector<TileBox> boxes((size_t)tiles_x * tiles_y);
    for (auto& b : boxes) { b.x0 = 32767; b.y0 = 32767; b.x1 = -32768; b.y1 = -32768; }

    for (int v = 0; v < dst_h; ++v) {
        for (int u = 0; u < dst_w; ++u) {
            const float mx = mapx[(size_t)v * dst_w + u], my = mapy[(size_t)v * dst_w + u];
            if (!(fabsf(mx) < 1e6f) || !(fabsf(my) < 1e6f)) continue;  // NaN / far away: no tap inside
            const int ix = round_half_even(mx * 32.0f), iy = round_half_even(my * 32.0f);
            const int x0 = ix >> 5, y0 = iy >> 5;  // arithmetic shift = floor
            if (x0 < -1 || x0 > src_w - 1 || y0 < -1 || y0 > src_h - 1) continue;  // all four taps outside
            lut[(size_t)v * lut_pitch + u] = lut_pack(x0, y0, (uint32_t)(ix & 31), (uint32_t)(iy & 31));
            valid[(size_t)v * dst_w + u] = (x0 >= 0 && x0 + 1 <= src_w - 1 && y0 >= 0 && y0 + 1 <= src_h - 1) ? 1 : 0;
            TileBox& b = boxes[(size_t)(v / RT_H) * tiles_x + (u / RT_W)];
            b.x0 = std::min<int16_t>(b.x0, (int16_t)x0); b.y0 = std::min<int16_t>(b.y0, (int16_t)y0);
            b.x1 = std::max<int16_t>(b.x1, (int16_t)(x0 + 2)); b.y1 = std::max<int16_t>(b.y1, (int16_t)(y0 + 2));
        }
    }
    size_t smem1 = 0, smem3 = 0;
    for (auto& b : boxes) {
        if (b.x1 <= b.x0) { b.x0 = b.y0 = b.x1 = b.y1 = 0; continue; }
        const int rows = b.y1 - b.y0;
        const int p1 = ((b.x1 + 15) & ~15) - (b.x0 & ~15);
        const int p3 = ((b.x1 * 3 + 15) & ~15) - ((b.x0 * 3) & ~15);
        smem1 = std::max(smem1, (size_t)rows * p1);
        smem3 = std::max(smem3, (size_t)rows * p3);
    }
    TI_CUDA(ctx, cudaMalloc(&C.d_lut, lut.size() * sizeof(lut_t)));
    TI_CUDA(ctx, cudaMalloc(&C.d_boxes, boxes.size() * sizeof(TileBox)));
    TI_CUDA(ctx, cudaMalloc(&C.d_valid, valid.size()));
    TI_CUDA(ctx, cudaMemcpy(C.d_lut, lut.data(), lut.size() * sizeof(lut_t), cudaMemcpyHostToDevice));
    TI_CUDA(ctx, cudaMemcpy(C.d_boxes, boxes.data(), boxes.size() * sizeof(TileBox), cudaMemcpyHostToDevice));
    TI_CUDA(ctx, cudaMemcpy(C.d_valid, valid.data(), valid.size(), cudaMemcpyHostToDevice));
    C.dst_w = dst_w; C.dst_h = dst_h; C.src_w = src_w; C.src_h = src_h;
    C.tiles_x = tiles_x; C.tiles_y = tiles_y;
    C.tile_smem[0] = smem1; C.tile_smem[1] = smem3;
    C.has_map = true;

    // ---- fast mono tables: tile-relative LUT ---------------------------------------------------
    // entry = off << 16 | fy << 6 | fx ; off = byte address in shared memory of the (2-byte aligned)
    // pair (src[y0][x0], src[y0][x0+1]): copy A holds the row as is (pairs with even x0 - c0), copy B
    // holds it shifted by one byte (odd x0 - c0); the pair one row below sits M2_ROW_BYTES further.
    // Within a tile row the entries are permuted so that lane L of a warp finds the entries of
    // pixels L, L+32, L+64, L+96 in one 128-bit word (lanes then read consecutive source bytes).
    // Pixels with no tap inside the image point at the zero block with fx = fy = 0.
    if (src_w % 16 == 0) {
        const int t2x = (dst_w + M2_TW - 1) / M2_TW, t2y = (dst_h + M2_TH - 1) / M2_TH;
        std::vector<TileBox2> boxes2((size_t)t2x * t2y);
        std::vector<uint32_t> lut2((size_t)t2x * t2y * M2_TW * M2_TH, 0u);
        bool ok = true;
        int rows_max = 0;
        for (int ty = 0; ty < t2y && ok; ++ty)
            for (int tx = 0; tx < t2x && ok; ++tx) {
                int bx0 = 1 << 20, by0 = 1 << 20, bx1 = -(1 << 20), by1 = -(1 << 20);
                for (int v = ty * M2_TH; v < std::min(dst_h, (ty + 1) * M2_TH); ++v)
                    for (int u = tx * M2_TW; u < std::min(dst_w, (tx + 1) * M2_TW); ++u) {
                        const lut_t e = lut[(size_t)v * lut_pitch + u];
                        if (e == LUT_OUTSIDE) continue;
                        const int x0 = lut_x0(e), y0 = lut_y0(e);
                        bx0 = std::min(bx0, x0); by0 = std::min(by0, y0);
                        bx1 = std::max(bx1, x0 + 2); by1 = std::max(by1, y0 + 2);
                    }
                TileBox2& B = boxes2[(size_t)ty * t2x + tx];
                B = TileBox2{0, 0, 0, 0, (int16_t)(tx * M2_TW), (int16_t)(ty * M2_TH), 0, 0};
                if (bx1 <= bx0) continue;
                const int c0 = bx0 & ~15;  // floor to 16 (two's complement: -1 -> -16)
                const int span = bx1 - c0;
                const int rows = by1 - by0;
                if (span > M2_SPAN_BYTES || rows > M2_MAX_ROWS) { ok = false; break; }
                B.c0 = (int16_t)c0; B.y0 = (int16_t)by0; B.nvec = (int16_t)((span + 15) / 16); B.rows = (int16_t)rows;
                rows_max = std::max(rows_max, rows);
                uint32_t* tl = lut2.data() + ((size_t)ty * t2x + tx) * M2_TW * M2_TH;
                for (int v = ty * M2_TH; v < std::min(dst_h, (ty + 1) * M2_TH); ++v)
                    for (int u = tx * M2_TW; u < std::min(dst_w, (tx + 1) * M2_TW); ++u) {
                        const lut_t e = lut[(size_t)v * lut_pitch + u];
                        if (e == LUT_OUTSIDE) continue;
                        const int x0 = lut_x0(e), y0 = lut_y0(e);
                        const uint32_t fx = lut_fx(e), fy = lut_fy(e);
                        const int rel = x0 - c0;
                        const int off = M2_ZERO_BYTES + (y0 - by0) * M2_ROW_BYTES + ((rel & 1) ? M2_COPY_BYTES + rel - 1 : rel);
                        const int lu = u - tx * M2_TW;  // lane (lu % 32) owns pixels lu, lu+32, lu+64, lu+96: stored as its uint4
                        tl[(size_t)(v - ty * M2_TH) * M2_TW + (lu & 31) * 4 + (lu >> 5)] = ((uint32_t)off << 16) | (fy << 6) | fx;
                    }
            }
        if (ok) {
            TI_CUDA(ctx, cudaMalloc(&C.d_lut2, lut2.size() * sizeof(uint32_t)));
            TI_CUDA(ctx, cudaMalloc(&C.d_boxes2, boxes2.size() * sizeof(TileBox2)));
            TI_CUDA(ctx, cudaMemcpy(C.d_lut2, lut2.data(), lut2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            TI_CUDA(ctx, cudaMemcpy(C.d_boxes2, boxes2.data(), boxes2.size() * sizeof(TileBox2), cudaMemcpyHostToDevice));
            C.tiles2_x = t2x; C.tiles2_y = t2y; C.rows2_max = rows_max;
            C.has_fast_mono = true;
        }
    }

    // ---- TMA-pipelined mono tables (TH = 16 and TH = 32) ----------------------------------------
    // Two passes per tile height: the tap offsets embed the camera's rows_alloc (start of copy B),
    // which is only known once every tile's box has been measured.
    for (int k = 0; k < 3 && src_w % 16 == 0; ++k) {
        const int TH = M3_TILE_HEIGHTS[k];
        const int t3x = (dst_w + M3_TW - 1) / M3_TW, t3y = (dst_h + TH - 1) / TH;
        std::vector<TileBox2> boxes3((size_t)t3x * t3y);
        bool ok = true;
        int rows_max = 0;
        auto for_tile = [&](int tx, int ty, auto&& fn) {
            for (int v = ty * TH; v < std::min(dst_h, (ty + 1) * TH); ++v)
                for (int u = tx * M3_TW; u < std::min(dst_w, (tx + 1) * M3_TW); ++u) {
                    const lut_t e = lut[(size_t)v * lut_pitch + u];
                    if (e == LUT_OUTSIDE) continue;
                    fn(u, v, lut_x0(e), lut_y0(e), lut_fx(e), lut_fy(e));
                }
        };
        for (int ty = 0; ty < t3y && ok; ++ty)
            for (int tx = 0; tx < t3x && ok; ++tx) {
                int bx0 = 1 << 20, by0 = 1 << 20, bx1 = -(1 << 20), by1 = -(1 << 20);
                for_tile(tx, ty, [&](int, int, int x0, int y0, uint32_t, uint32_t) {
                    bx0 = std::min(bx0, x0); by0 = std::min(by0, y0); bx1 = std::max(bx1, x0 + 2); by1 = std::max(by1, y0 + 2);
                });
                TileBox2& B = boxes3[(size_t)ty * t3x + tx];
                B = TileBox2{0, 0, 0, 0, (int16_t)(tx * M3_TW), (int16_t)(ty * TH), 0, 0};
                if (bx1 <= bx0) continue;
                const int c0 = bx0 & ~15;
                if (bx1 - c0 > M3_MAX_SPAN || by1 - by0 > M3_MAX_ROWS) { ok = false; break; }
                B.c0 = (int16_t)c0; B.y0 = (int16_t)by0; B.nvec = (int16_t)((bx1 - c0 + 15) / 16); B.rows = (int16_t)(by1 - by0);
                rows_max = std::max(rows_max, by1 - by0);
            }
        if (!ok) continue;
        const int rows_alloc = std::max(M3_BOX_ROWS, (rows_max + M3_BOX_ROWS - 1) / M3_BOX_ROWS * M3_BOX_ROWS);
        std::vector<uint32_t> lut3((size_t)t3x * t3y * M3_TW * TH, 0u);
        for (int ty = 0; ty < t3y; ++ty)
            for (int tx = 0; tx < t3x; ++tx) {
                const TileBox2& B = boxes3[(size_t)ty * t3x + tx];
                uint32_t* tl = lut3.data() + ((size_t)ty * t3x + tx) * M3_TW * TH;
                for_tile(tx, ty, [&](int u, int v, int x0, int y0, uint32_t fx, uint32_t fy) {
                    const int rel = x0 - B.c0, row = y0 - B.y0;
                    const int off = (rel & 1) ? 128 + (rows_alloc + row) * M3_PITCH + (rel - 1 + M3_B_SHIFT + 1)
                                              : 128 + row * M3_PITCH + rel;
                    // lane L of a warp owns pixels 2L, 2L+1, 64+2L, 64+2L+1 of the 128-pixel tile row (its uint4)
                    const int lu = u - tx * M3_TW;
                    const int lane = (lu & 63) >> 1, slot = ((lu >> 6) << 1) | (lu & 1);
                    tl[(size_t)(v - ty * TH) * M3_TW + lane * 4 + slot] = ((uint32_t)off << 16) | (fy << 6) | fx;
                });
            }
        TI_CUDA(ctx, cudaMalloc(&C.d_lut3[k], lut3.size() * sizeof(uint32_t)));
        TI_CUDA(ctx, cudaMalloc(&C.d_boxes3[k], boxes3.size() * sizeof(TileBox2)));
        TI_CUDA(ctx, cudaMemcpy(C.d_lut3[k], lut3.data(), lut3.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        TI_CUDA(ctx, cudaMemcpy(C.d_boxes3[k], boxes3.data(), boxes3.size() * sizeof(TileBox2), cudaMemcpyHostToDevice));
        C.tiles3_x[k] = t3x; C.tiles3_y[k] = t3y; C.rows3_alloc[k] = rows_alloc;
        C.has_tma_mono[k] = true;
    }
    const int rc_pair = build_pair_tables(ctx, C, lut, lut_pitch);
    C.h_lut = std::move(lut);
    C.h_lut_pitch = lut_pitch;
    return rc_pair;
}

int ti_upload_projection(ti_ctx* ctx, int camera, int width, int height, const double k[4], const double body_T_cam[12]) {
    if (!ctx) return TI_EINVAL;
    if (camera < 0 || camera >= TI_MAX_CAMERAS) return fail(ctx, TI_EINVAL, "camera slot %d out of range", camera);
    if (!k || !body_T_cam) return fail(ctx, TI_EINVAL, "ti_upload_projection: null argument");
    if (width <= 0 || height <= 0 || width > 16384 || height > 16384)
        return fail(ctx, TI_EINVAL, "ti_upload_projection: bad size %dx%d", width, height);
    if (!(k[0] != 0.0) || !(k[1] != 0.0)) return fail(ctx, TI_EINVAL, "ti_upload_projection: fx and fy must be non-zero");
    CameraSlot& C = ctx->cams[camera];
    for (int r = 0; r < 3; ++r) {
        const double ax = body_T_cam[4 * r + 0] / k[0], ay = body_T_cam[4 * r + 1] / k[1], az = body_T_cam[4 * r + 2];
        C.proj_au[r] = 1e-3 * ax;
        C.proj_av[r] = 1e-3 * ay;
        C.proj_ac[r] = 1e-3 * (az - ax * k[2] - ay * k[3]);
        C.proj_t[r] = body_T_cam[4 * r + 3];
    }
    C.proj_w = width; C.proj_h = height;
    C.has_proj = true;
    return TI_OK;
}

int ti_rectify_plan(ti_ctx* ctx, int camera, int32_t out[8]) {
    if (!ctx) return TI_EINVAL;
    if (camera < 0 || camera >= TI_MAX_CAMERAS || !out) return fail(ctx, TI_EINVAL, "ti_rectify_plan: bad argument");
    CameraSlot& C = ctx->cams[camera];
    if (!C.has_map) return fail(ctx, TI_ESTATE, "ti_rectify_plan: camera slot %d has no remap LUT", camera);
    const int th4 = p4_th_index(ctx->tma_tile_h), thk = m3_th_index(ctx->tma_tile_h);
    out[0] = 1; out[1] = RT_H; out[2] = 0; out[3] = 0; out[4] = 1; out[5] = 0; out[6] = 0; out[7] = 0;
    if (ctx->force_generic_rectify) return TI_OK;
    if (ctx->mono_variant == 4 && C.src_w % 16 == 0) {
        if (!C.c3_tried) {
            TI_CUDA(ctx, cudaSetDevice(ctx->device));
            const int rc = build_c3_tables(ctx, C);
            if (rc != TI_OK) return rc;
        }
        if (C.has_c3) { out[4] = 5; out[5] = C.rows5_alloc; }
    }
    if (ctx->mono_variant == 4 && C.has_pair[th4]) {
        out[0] = 4; out[1] = P4_TILE_HEIGHTS[th4]; out[2] = C.rows4_alloc[th4]; out[3] = C.exc4_per_warp[th4]; out[6] = C.n_over4[th4];
        out[7] = C.pitch4[th4] | (C.quad4[th4] ? 0x10000 : 0);
    } else if (ctx->mono_variant >= 3 && C.has_tma_mono[thk]) {
        out[0] = 3; out[1] = M3_TILE_HEIGHTS[thk]; out[2] = C.rows3_alloc[thk];
    } else if (ctx->mono_variant >= 2 && C.has_fast_mono) {
        out[0] = 2; out[1] = M2_TH; out[2] = C.rows2_max;
    }
    return TI_OK;
}

int ti_upload_registration(ti_ctx* ctx, int camera, int depth_w, int depth_h, const double k_depth[4], int rgb_w, int rgb_h,
                           const double k_rgb[4], const double rgb_T_depth[12]) {
    if (!ctx) return TI_EINVAL;
    if (camera < 0 || camera >= TI_MAX_CAMERAS) return fail(ctx, TI_EINVAL, "camera slot %d out of range", camera);
    if (!k_depth || !k_rgb || !rgb_T_depth || depth_w <= 0 || depth_h <= 0 || rgb_w <= 0 || rgb_h <= 0)
        return fail(ctx, TI_EINVAL, "ti_upload_registration: bad argument");
    if (k_depth[0] == 0.0 || k_depth[1] == 0.0) return fail(ctx, TI_EINVAL, "ti_upload_registration: zero focal length");
    CameraSlot& C = ctx->cams[camera];
    for (int i = 0; i < 3; ++i) {
        C.reg_a[3 * i + 0] = (float)(rgb_T_depth[4 * i + 0] / k_depth[0]);
        C.reg_a[3 * i + 1] = (float)(rgb_T_depth[4 * i + 1] / k_depth[1]);
        C.reg_a[3 * i + 2] = (float)rgb_T_depth[4 * i + 2];
        C.reg_t[i] = (float)rgb_T_depth[4 * i + 3];
    }
    C.reg_k[0] = (float)k_depth[2]; C.reg_k[1] = (float)k_depth[3];
    C.reg_k[2] = (float)k_rgb[0]; C.reg_k[3] = (float)k_rgb[1]; C.reg_k[4] = (float)k_rgb[2]; C.reg_k[5] = (float)k_rgb[3];
    C.reg_dw = depth_w; C.reg_dh = depth_h; C.reg_rw = rgb_w; C.reg_rh = rgb_h;
    C.has_reg = true;
    return TI_OK;
}

int ti_register_colour(ti_ctx* ctx, int camera, const uint16_t* depth, const uint8_t* rgb, uint8_t* colour, int n_batch,
                       uint64_t depth_frame_stride, uint64_t rgb_frame_stride, uint64_t colour_frame_stride) {
    if (!ctx) return TI_EINVAL;
    if (camera < 0 || camera >= TI_MAX_CAMERAS || !ctx->cams[camera].has_reg)
        return fail(ctx, TI_ESTATE, "ti_register_colour: camera slot %d has no registration (call ti_upload_registration)", camera);
    if (n_batch < 0 || (n_batch > 0 && (!depth || !rgb || !colour))) return fail(ctx, TI_EINVAL, "ti_register_colour: bad argument");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_register_colour(ctx, camera, depth, rgb, colour, n_batch, depth_frame_stride, rgb_frame_stride, colour_frame_stride);
}

int ti_get_valid_mask(ti_ctx* ctx, int camera, uint8_t* dst) {
    if (!ctx) return TI_EINVAL;
    if (camera < 0 || camera >= TI_MAX_CAMERAS || !ctx->cams[camera].has_map)
        return fail(ctx, TI_ESTATE, "camera slot %d has no remap LUT", camera);
    if (!dst) return fail(ctx, TI_EINVAL, "ti_get_valid_mask: null dst");
    const CameraSlot& C = ctx->cams[camera];
    TI_CUDA(ctx, cudaMemcpyAsync(dst, C.d_valid, (size_t)C.dst_w * C.dst_h, cudaMemcpyDeviceToDevice, ctx->stream));
    return TI_OK;
}

// ---- per-frame entry points --------------------------------------------------------------------

int ti_convert(ti_ctx* ctx, int src_format, int dst_format, const void* src, void* dst, int width, int height, int n_batch,
               uint64_t src_frame_stride, uint64_t dst_frame_stride) {
    if (!ctx) return TI_EINVAL;
    if (n_batch < 0) return fail(ctx, TI_EINVAL, "n_batch must be >= 0");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    ConvertJob J{(const uint8_t*)src, (uint8_t*)dst, src_frame_stride, dst_frame_stride, width, height, src_format, dst_format};
    return launch_convert(ctx, &J, 1, n_batch);
}

int ti_rectify(ti_ctx* ctx, int camera, int src_format, int dst_format, const void* src, void* dst, int n_batch,
               uint64_t src_frame_stride, uint64_t dst_frame_stride) {
    if (!ctx) return TI_EINVAL;
    if (n_batch < 0) return fail(ctx, TI_EINVAL, "n_batch must be >= 0");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    RectifyJob J{(const uint8_t*)src, (uint8_t*)dst, src_frame_stride, dst_frame_stride, camera, src_format, dst_format};
    return launch_rectify(ctx, &J, 1, n_batch);
}

int ti_backproject(ti_ctx* ctx, int camera, const uint16_t* depth, float* xyz, uint8_t* mask, uint32_t* count, int n_batch,
                   uint64_t depth_frame_stride, uint64_t xyz_frame_stride, uint64_t mask_frame_stride) {
    if (!ctx) return TI_EINVAL;
    if (n_batch < 0) return fail(ctx, TI_EINVAL, "n_batch must be >= 0");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    BackprojectJob J{depth, xyz, mask, count, depth_frame_stride, xyz_frame_stride, mask_frame_stride, camera};
    return launch_backproject(ctx, &J, 1, n_batch);
}

int ti_depth_stats(ti_ctx* ctx, const uint16_t* depth, int width, int height, int n_batch, uint64_t depth_frame_stride, uint32_t* stats) {
    if (!ctx) return TI_EINVAL;
    if (n_batch < 0 || width <= 0 || height <= 0 || (n_batch > 0 && (!depth || !stats))) return fail(ctx, TI_EINVAL, "ti_depth_stats: bad argument");
    if (depth_frame_stride % 2 || (uintptr_t)depth % 2) return fail(ctx, TI_EINVAL, "ti_depth_stats: depth must be 2-byte aligned");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_depth_stats(ctx, depth, width, height, n_batch, depth_frame_stride, stats);
}

int ti_backproject_colour(ti_ctx* ctx, int camera, const uint16_t* depth, const uint8_t* rgb, float* xyz, uint8_t* mask, uint32_t* count,
                          uint8_t* colour, int n_batch, uint64_t depth_frame_stride, uint64_t rgb_frame_stride, uint64_t xyz_frame_stride,
                          uint64_t mask_frame_stride, uint64_t colour_frame_stride) {
    if (!ctx) return TI_EINVAL;
    if (n_batch < 0) return fail(ctx, TI_EINVAL, "n_batch must be >= 0");
    if (n_batch > 0 && (!rgb || !colour)) return fail(ctx, TI_EINVAL, "ti_backproject_colour: null rgb / colour pointer");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    BackprojectJob J{depth, xyz, mask, count, depth_frame_stride, xyz_frame_stride, mask_frame_stride, camera};
    J.rgb = rgb; J.colour = colour; J.rgb_stride = rgb_frame_stride; J.colour_stride = colour_frame_stride;
    return launch_backproject(ctx, &J, 1, n_batch);
}

static int split_streams(ti_ctx* ctx, const ti_stream* streams, int n_streams, std::vector<ConvertJob>& cv,
                         std::vector<RectifyJob>& rc, std::vector<BackprojectJob>& bp) {
    if (n_streams < 0 || (n_streams > 0 && !streams)) return fail(ctx, TI_EINVAL, "ti_ingest: bad stream array");
    for (int i = 0; i < n_streams; ++i) {
        const ti_stream& S = streams[i];
        switch (S.kind) {
            case TI_KIND_CONVERT:
                cv.push_back({(const uint8_t*)S.src, (uint8_t*)S.dst, S.src_frame_stride, S.dst_frame_stride, S.width, S.height,
                              S.src_format, S.dst_format});
                break;
            case TI_KIND_RECTIFY:
                rc.push_back({(const uint8_t*)S.src, (uint8_t*)S.dst, S.src_frame_stride, S.dst_frame_stride, S.camera,
                              S.src_format, S.dst_format});
                break;
            case TI_KIND_BACKPROJECT:
                if (S.src_format != TI_FMT_DEPTH16 || S.dst_format != TI_FMT_XYZ32F)
                    return fail(ctx, TI_EINVAL, "stream %d: BACKPROJECT needs DEPTH16 -> XYZ32F", i);
                bp.push_back({(const uint16_t*)S.src, (float*)S.dst, (uint8_t*)S.mask, S.count, S.src_frame_stride,
                              S.dst_frame_stride, S.mask_frame_stride, S.camera});
                break;
            default: return fail(ctx, TI_EINVAL, "stream %d: unknown kind %d", i, S.kind);
        }
    }
    return TI_OK;
}

int ti_ingest(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch) {
    if (!ctx) return TI_EINVAL;
    if (n_batch < 0) return fail(ctx, TI_EINVAL, "n_batch must be >= 0");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<ConvertJob> cv; std::vector<RectifyJob> rc; std::vector<BackprojectJob> bp;
    int r = split_streams(ctx, streams, n_streams, cv, rc, bp);
    if (r != TI_OK) return r;
    if ((r = launch_rectify(ctx, rc.data(), (int)rc.size(), n_batch)) != TI_OK) return r;
    if ((r = launch_backproject(ctx, bp.data(), (int)bp.size(), n_batch)) != TI_OK) return r;
    for (size_t i = 0; i < cv.size(); i += TI_MAX_STREAMS) {
        const int n = (int)std::min<size_t>(TI_MAX_STREAMS, cv.size() - i);
        if ((r = launch_convert(ctx, cv.data() + i, n, n_batch)) != TI_OK) return r;
    }
    return TI_OK;
}

// ---- host-buffer pipeline ----------------------------------------------------------------------
// Frames arrive in (pinned) host memory; per chunk of `chunk` frame sets: H2D on one stream,
// kernels on a second, D2H on a third, three chunk slots in flight so the three overlap.

static uint64_t stream_src_bytes(const ti_ctx* ctx, const ti_stream& S) {
    if (S.kind == TI_KIND_RECTIFY) { const CameraSlot& C = ctx->cams[S.camera]; return frame_bytes(S.src_format, C.src_w, C.src_h); }
    if (S.kind == TI_KIND_BACKPROJECT) { const CameraSlot& C = ctx->cams[S.camera]; return frame_bytes(TI_FMT_DEPTH16, C.proj_w, C.proj_h); }
    return frame_bytes(S.src_format, S.width, S.height);
}
static uint64_t stream_dst_bytes(const ti_ctx* ctx, const ti_stream& S) {
    if (S.kind == TI_KIND_RECTIFY) { const CameraSlot& C = ctx->cams[S.camera]; return frame_bytes(S.dst_format, C.dst_w, C.dst_h); }
    if (S.kind == TI_KIND_BACKPROJECT) { const CameraSlot& C = ctx->cams[S.camera]; return frame_bytes(TI_FMT_XYZ32F, C.proj_w, C.proj_h); }
    return frame_bytes(S.dst_format, S.width, S.height);
}
static uint64_t stream_mask_bytes(const ti_ctx* ctx, const ti_stream& S) {
    if (S.kind != TI_KIND_BACKPROJECT || !S.mask) return 0;
    const CameraSlot& C = ctx->cams[S.camera];
    return (uint64_t)C.proj_w * C.proj_h;
}

static int ensure_cap(ti_ctx* ctx, void** p, size_t* cap, size_t need) {
    if (*cap >= need) return TI_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    if (need == 0) return TI_OK;
    TI_CUDA(ctx, cudaMalloc(p, need));
    *cap = need;
    return TI_OK;
}

// Enqueues one host-buffer call on the pipeline streams.  Chunks are numbered over the life of the context, so the
// three buffer slots and their events carry over from one call to the next and consecutive submissions overlap.
static int host_enqueue(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch, int chunk, bool ramp, const char* who) {
    if (n_batch < 0 || n_streams < 0 || (n_streams && !streams)) return fail(ctx, TI_EINVAL, "%s: bad arguments", who);
    if (n_batch == 0 || n_streams == 0) return TI_OK;
    if (chunk <= 0) chunk = 1;
    chunk = std::min(chunk, n_batch);
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n_streams; ++i) {
        const ti_stream& S = streams[i];
        if ((S.kind == TI_KIND_RECTIFY && (S.camera < 0 || S.camera >= TI_MAX_CAMERAS || !ctx->cams[S.camera].has_map)) ||
            (S.kind == TI_KIND_BACKPROJECT && (S.camera < 0 || S.camera >= TI_MAX_CAMERAS || !ctx->cams[S.camera].has_proj)))
            return fail(ctx, TI_ESTATE, "stream %d: camera slot %d not uploaded", i, S.camera);
        if (!S.src || !S.dst) return fail(ctx, TI_EINVAL, "stream %d: null src/dst", i);
    }
    if (!ctx->host_ready) {
        TI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        TI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        TI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_exec, cudaStreamNonBlocking));
        for (auto& h : ctx->hslot) {
            TI_CUDA(ctx, cudaEventCreateWithFlags(&h.h2d_done, cudaEventDisableTiming));
            TI_CUDA(ctx, cudaEventCreateWithFlags(&h.exec_done, cudaEventDisableTiming));
            TI_CUDA(ctx, cudaEventCreateWithFlags(&h.d2h_done, cudaEventDisableTiming));
        }
        for (auto& ev : ctx->ticket_done) TI_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->host_ready = true;
    }
    // device staging: tightly packed frames, `chunk` per stream per slot
    std::vector<uint64_t> sb(n_streams), db(n_streams), mb(n_streams);
    for (int i = 0; i < n_streams; ++i) {
        sb[i] = (stream_src_bytes(ctx, streams[i]) + 15) & ~15ull;
        db[i] = (stream_dst_bytes(ctx, streams[i]) + 15) & ~15ull;
        mb[i] = (stream_mask_bytes(ctx, streams[i]) + 15) & ~15ull;
    }
    bool grow = false;
    for (auto& h : ctx->hslot) {
        if ((int)h.d_src.size() < n_streams) { grow = true; break; }
        for (int i = 0; i < n_streams && !grow; ++i) {
            const size_t cb = (streams[i].kind == TI_KIND_BACKPROJECT && streams[i].count) ? sizeof(uint32_t) * chunk : 0;
            grow = h.cap_src[i] < sb[i] * chunk || h.cap_dst[i] < db[i] * chunk || h.cap_mask[i] < mb[i] * chunk || h.cap_count[i] < cb;
        }
    }
    if (grow) {
        // buffers of earlier submissions may still be in flight
        TI_CUDA(ctx, cudaStreamSynchronize(ctx->s_d2h));
        TI_CUDA(ctx, cudaStreamSynchronize(ctx->s_exec));
        TI_CUDA(ctx, cudaStreamSynchronize(ctx->s_h2d));
        for (auto& h : ctx->hslot) {
            if ((int)h.d_src.size() < n_streams) {
                h.d_src.resize(n_streams, nullptr); h.d_dst.resize(n_streams, nullptr); h.d_mask.resize(n_streams, nullptr);
                h.d_count.resize(n_streams, nullptr);
                h.cap_src.resize(n_streams, 0); h.cap_dst.resize(n_streams, 0); h.cap_mask.resize(n_streams, 0); h.cap_count.resize(n_streams, 0);
            }
            for (int i = 0; i < n_streams; ++i) {
                int r;
                if ((r = ensure_cap(ctx, &h.d_src[i], &h.cap_src[i], sb[i] * chunk)) != TI_OK) return r;
                if ((r = ensure_cap(ctx, &h.d_dst[i], &h.cap_dst[i], db[i] * chunk)) != TI_OK) return r;
                if ((r = ensure_cap(ctx, &h.d_mask[i], &h.cap_mask[i], mb[i] * chunk)) != TI_OK) return r;
                const size_t cb = (streams[i].kind == TI_KIND_BACKPROJECT && streams[i].count) ? sizeof(uint32_t) * chunk : 0;
                if ((r = ensure_cap(ctx, (void**)&h.d_count[i], &h.cap_count[i], cb)) != TI_OK) return r;
            }
        }
    }
    cudaStream_t user_stream = ctx->stream;
    int rc = TI_OK;
    // Chunk schedule.  A call that returns only when everything has landed has its first upload and its last download
    // overlapped with nothing - one half-size chunk at either end keeps them short.  (Smaller pieces cost more than
    // they save: a host<->device copy carries ~15 us of fixed cost, tools/pcie_probe.py.)  Submissions that overlap
    // their neighbours need no ramp.
    std::vector<int> sizes;
    {
        int left = n_batch;
        const int half = chunk / 2;
        ramp = ramp && half >= 1 && n_batch >= 3 * chunk;
        if (ramp) { sizes.push_back(half); left -= 2 * half; }
        while (left > 0) { const int nb = std::min(chunk, left); sizes.push_back(nb); left -= nb; }
        if (ramp) sizes.push_back(half);
    }
    const int n_chunks = (int)sizes.size();
    int b_next = 0;
    for (int c = 0; c < n_chunks && rc == TI_OK; ++c) {
        const uint64_t g = ctx->host_chunks++;
        auto& h = ctx->hslot[g % 3];
        const int b0 = b_next, nb = sizes[c];
        b_next += nb;
        // the slot's previous D2H must have drained before its buffers are overwritten
        if (g >= 3) {
            cudaStreamWaitEvent(ctx->s_h2d, h.d2h_done, 0);
            cudaStreamWaitEvent(ctx->s_exec, h.d2h_done, 0);
        }
        std::vector<ti_stream> dev(streams, streams + n_streams);
        for (int i = 0; i < n_streams; ++i) {
            const ti_stream& S = streams[i];
            const uint64_t raw = stream_src_bytes(ctx, S);
            const uint64_t hs = S.src_frame_stride ? S.src_frame_stride : raw;
            if (hs == raw && sb[i] == raw) {
                cudaMemcpyAsync(h.d_src[i], (const uint8_t*)S.src + (uint64_t)b0 * hs, raw * nb, cudaMemcpyHostToDevice, ctx->s_h2d);
            } else {
                cudaMemcpy2DAsync(h.d_src[i], sb[i], (const uint8_t*)S.src + (uint64_t)b0 * hs, hs, raw, nb, cudaMemcpyHostToDevice, ctx->s_h2d);
            }
            dev[i].src = h.d_src[i]; dev[i].dst = h.d_dst[i];
            dev[i].src_frame_stride = sb[i]; dev[i].dst_frame_stride = db[i];
            dev[i].mask = mb[i] ? h.d_mask[i] : nullptr; dev[i].mask_frame_stride = mb[i];
            dev[i].count = (S.kind == TI_KIND_BACKPROJECT && S.count) ? h.d_count[i] : nullptr;
        }
        cudaEventRecord(h.h2d_done, ctx->s_h2d);
        cudaStreamWaitEvent(ctx->s_exec, h.h2d_done, 0);
        ctx->stream = ctx->s_exec;
        rc = ti_ingest(ctx, dev.data(), n_streams, nb);
        ctx->stream = user_stream;
        if (rc != TI_OK) break;
        cudaEventRecord(h.exec_done, ctx->s_exec);
        cudaStreamWaitEvent(ctx->s_d2h, h.exec_done, 0);
        for (int i = 0; i < n_streams; ++i) {
            const ti_stream& S = streams[i];
            const uint64_t raw = stream_dst_bytes(ctx, S);
            const uint64_t hs = S.dst_frame_stride ? S.dst_frame_stride : raw;
            if (hs == raw && db[i] == raw) {
                cudaMemcpyAsync((uint8_t*)S.dst + (uint64_t)b0 * hs, h.d_dst[i], raw * nb, cudaMemcpyDeviceToHost, ctx->s_d2h);
            } else {
                cudaMemcpy2DAsync((uint8_t*)S.dst + (uint64_t)b0 * hs, hs, h.d_dst[i], db[i], raw, nb, cudaMemcpyDeviceToHost, ctx->s_d2h);
            }
            if (mb[i]) {
                const uint64_t mraw = stream_mask_bytes(ctx, S);
                const uint64_t ms = S.mask_frame_stride ? S.mask_frame_stride : mraw;
                cudaMemcpy2DAsync((uint8_t*)S.mask + (uint64_t)b0 * ms, ms, h.d_mask[i], mb[i], mraw, nb, cudaMemcpyDeviceToHost, ctx->s_d2h);
            }
            if (S.kind == TI_KIND_BACKPROJECT && S.count)
                cudaMemcpyAsync(S.count + b0, h.d_count[i], sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, ctx->s_d2h);
        }
        cudaEventRecord(h.d2h_done, ctx->s_d2h);
    }
    return rc;
}

static int host_drain(ti_ctx* ctx, const char* who) {
    if (!ctx->host_ready) return TI_OK;
    cudaError_t e1 = cudaStreamSynchronize(ctx->s_d2h);
    cudaError_t e2 = cudaStreamSynchronize(ctx->s_exec);
    cudaError_t e3 = cudaStreamSynchronize(ctx->s_h2d);
    cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, TI_ECUDA, "%s: %s", who, cudaGetErrorString(e));
    return TI_OK;
}

int ti_ingest_host(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch, int chunk) {
    if (!ctx) return TI_EINVAL;
    const int rc = host_enqueue(ctx, streams, n_streams, n_batch, chunk, /*ramp=*/true, "ti_ingest_host");
    const int rd = host_drain(ctx, "ti_ingest_host");
    return rc != TI_OK ? rc : rd;
}

int ti_ingest_host_submit(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch, int chunk, uint64_t* ticket) {
    if (!ctx) return TI_EINVAL;
    if (!ticket) return fail(ctx, TI_EINVAL, "ti_ingest_host_submit: null ticket");
    constexpr uint64_t kRing = sizeof(ctx->ticket_done) / sizeof(ctx->ticket_done[0]);
    const uint64_t t = ctx->host_tickets + 1;
    // the event of ticket t - kRing is about to be reused
    if (ctx->host_ready && t > kRing) TI_CUDA(ctx, cudaEventSynchronize(ctx->ticket_done[t % kRing]));
    const int rc = host_enqueue(ctx, streams, n_streams, n_batch, chunk, /*ramp=*/false, "ti_ingest_host_submit");
    if (rc != TI_OK) { host_drain(ctx, "ti_ingest_host_submit"); return rc; }
    if (!ctx->host_ready) {  // empty submission before any stream exists: nothing to wait for
        *ticket = 0;
        return TI_OK;
    }
    TI_CUDA(ctx, cudaEventRecord(ctx->ticket_done[t % kRing], ctx->s_d2h));
    ctx->host_tickets = t;
    *ticket = t;
    return TI_OK;
}

int ti_ingest_host_wait(ti_ctx* ctx, uint64_t ticket) {
    if (!ctx) return TI_EINVAL;
    if (ticket == 0) return TI_OK;
    if (ticket > ctx->host_tickets) return fail(ctx, TI_EINVAL, "ti_ingest_host_wait: ticket %llu was never issued", (unsigned long long)ticket);
    constexpr uint64_t kRing = sizeof(ctx->ticket_done) / sizeof(ctx->ticket_done[0]);
    if (ctx->host_tickets - ticket >= kRing) return TI_OK;  // its event was waited for before being reused
    cudaError_t e = cudaEventSynchronize(ctx->ticket_done[ticket % kRing]);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, TI_ECUDA, "ti_ingest_host_wait: %s", cudaGetErrorString(e));
    return TI_OK;
}

int ti_copy_async(ti_ctx* ctx, void* dst, const void* src, uint64_t bytes, void* cuda_stream) {
    if (!ctx) return TI_EINVAL;
    if (bytes == 0) return TI_OK;
    if (!dst || !src) return fail(ctx, TI_EINVAL, "ti_copy_async: null buffer");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream));
    return TI_OK;
}

}  // extern "C"
