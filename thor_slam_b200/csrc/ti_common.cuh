// Shared definitions of libthoringest.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/thoringest.h"

// TI_EMULATE is defined ONLY by tests/emu (a g++ build that runs these very kernel sources
// thread-by-thread on the CPU so indexing bugs are found before GPU time is spent).  The
// shipped library is always built by nvcc for sm_100a with TI_EMULATE undefined.
#ifdef TI_EMULATE
#include "cuda_emu.h"
#define TI_DEVICE_CODE 1
#else
#ifdef __CUDACC__
#define TI_DEVICE_CODE 1
#endif
#define TI_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TI_DYNAMIC_SMEM(type, name) extern __shared__ __align__(128) type name[]
#endif

namespace ti {

// ---------------------------------------------------------------------------------------------
// Remap LUT encoding (the GENERIC table: what the fast tile-relative tables are built from on the host, and what the
// fall-back kernels read).  One u64 per OUTPUT pixel, row-major:
//   bits  0..15  x0 + 1   (x0 = floor(mapx) after 1/32-px quantisation; -1 <= x0 <= src_w-1)
//   bits 16..31  y0 + 1
//   bits 32..36  fx       (5-bit fractional part, units of 1/32 px)
//   bits 37..41  fy
// A pixel none of whose four taps touches the source image is LUT_OUTSIDE.  (Round 1 packed this into 32 bits with 11-bit
// coordinates, which capped images at 2046 px; the 4000 x 3000 and 4224 x 3136 sensor modes of the reference's driver -
// thor_slam/camera/drivers/luxonis.py:36-44 - need 13.)
// ---------------------------------------------------------------------------------------------
typedef uint64_t lut_t;
constexpr lut_t LUT_OUTSIDE = ~0ull;
__host__ __device__ inline lut_t lut_pack(int x0, int y0, uint32_t fx, uint32_t fy) {
    return (lut_t)(uint32_t)(x0 + 1) | ((lut_t)(uint32_t)(y0 + 1) << 16) | ((lut_t)fx << 32) | ((lut_t)fy << 37);
}
__host__ __device__ inline int lut_x0(lut_t e) { return (int)(e & 0xFFFFu) - 1; }
__host__ __device__ inline int lut_y0(lut_t e) { return (int)((e >> 16) & 0xFFFFu) - 1; }
__host__ __device__ inline uint32_t lut_fx(lut_t e) { return (uint32_t)(e >> 32) & 31u; }
__host__ __device__ inline uint32_t lut_fy(lut_t e) { return (uint32_t)(e >> 37) & 31u; }

// Rectify tile: one CTA produces RT_W x RT_H output pixels from a source box staged in smem.
constexpr int RT_W = 128;
constexpr int RT_H = 16;
constexpr int RT_THREADS = 256;

struct TileBox {  // source bounding box of one output tile, in source pixels
    int16_t x0, y0;  // first column / row any tap touches (may be -1: zero border)
    int16_t x1, y1;  // one past the last column / row any tap touches; x1 <= x0 means "all outside"
};

// Fast mono path ("v2"): tiles of M2_TW x M2_TH output pixels, LUT entries hold tile-relative
// shared-memory byte offsets (see ti_rectify.cu).
constexpr int M2_TW = 128;
constexpr int M2_TH = 32;
constexpr int M2_THREADS = 256;
constexpr int M2_ROW_BYTES = 624;   // per staged source row: copy A at +0, copy B (shifted by one byte) at +320; 624 = 112 (mod 128) spreads consecutive rows over the banks
constexpr int M2_COPY_BYTES = 320;  // = 64 (mod 128): copy B sits 16 banks away from copy A
constexpr int M2_SPAN_BYTES = 256;  // staged bytes per row and copy
constexpr int M2_ZERO_BYTES = 128;  // always-zero block at the start of shared memory
constexpr int M2_MAX_ROWS = 96;

struct TileBox2 {
    int16_t c0;    // first staged source column (multiple of 16, may be -16)
    int16_t y0;    // first staged source row (may be -1)
    int16_t nvec;  // 16-byte vectors staged per row (<= 16); 0 = nothing to stage
    int16_t rows;  // staged rows
    int16_t u0, v0;  // output pixel of the tile's top-left corner
    int16_t pad0, pad1;
};
static_assert(sizeof(TileBox2) == 16, "TileBox2 is loaded as one 128-bit word");

// TMA-pipelined mono path ("v3"): tiles of M3_TW x TH (TH = 16 or 32); per stage of shared memory
//   [128 B zeros | copy A: rows_alloc x 256 B | copy B: rows_alloc x 256 B | LUT tile | 128 B header]
// copy A row r = source bytes [c0, c0+256) of source row y0+r; copy B = [c0-63, c0+193) (one byte
// of pixel shift for odd x0, 64 bytes = 16 banks of bank shift).  Rows are TMA boxes of 256 x 8.
constexpr int M3_TW = 128;
constexpr int M3_TILE_HEIGHTS[3] = {16, 32, 24};
inline int m3_th_index(int th) { return th == 16 ? 0 : (th == 24 ? 2 : 1); }
constexpr int M3_BOX_ROWS = 8;
constexpr int M3_PITCH = 256;
constexpr int M3_B_SHIFT = 63;     // copy B[i] = source column c0 - 63 + i (built in shared memory from copy A)
constexpr int M3_MAX_SPAN = 192;
constexpr int M3_MAX_ROWS = 96;
constexpr int M3_CONSUMER_WARPS = 8;
constexpr int M3_PRODUCER_WARPS = 2;
constexpr int M3_THREADS = (M3_CONSUMER_WARPS + M3_PRODUCER_WARPS + 1) * 32;  // consumers, copy-B builders, TMA issuer
constexpr int M3_MAX_STAGES = 8;

// Pair-window mono path ("v4"): tiles of P4_TW x TH (TH = 16 or 32).  The TMA box of a tile lands
// as is (pitch P4_PITCH, no shifted copy, no builder warps); two horizontally adjacent output pixels
// share one 8-byte window per source row (see ti_rectify_pair.cu).
constexpr int P4_TW = 128;
constexpr int P4_N_TH = 3;
constexpr int P4_TILE_HEIGHTS[P4_N_TH] = {16, 32, 24};
inline int p4_th_index(int th) { return th == 16 ? 0 : (th == 24 ? 2 : 1); }
constexpr int P4_PITCH = 192;        // = 16 banks (mod 32): a lane group of the pair layout (32 windows = 16 words) crossing into the next source row stays
                                     // conflict-free.  The quad layout's 32 windows span 32 words and do collide (ncu: 1.6 wavefronts per window load); a
                                     // 256-byte pitch removes that but costs a ring stage - measured equal (profiles/r02_summary.md)
constexpr int P4_PITCH_WIDE = 320;   // same residue; for maps whose 128-pixel tiles span up to 320 source bytes (a 2 x downscale:
                                     // config/slam_config.yaml's output_resolution done on the host instead of on the camera)
constexpr int P4_MAX_ROWS = 96;
constexpr int P4_CONSUMER_WARPS = 8;
constexpr int P4_THREADS = (P4_CONSUMER_WARPS + 1) * 32;  // consumers + TMA issuer
constexpr int P4_MAX_STAGES = 8;
constexpr int P4_SMEM_HEADROOM_KB = 20;  // shared memory per SM the window kernels leave to co-resident exchange kernels
constexpr int P4_SMEM_HEADROOM_NCCL_KB = 36;  // what a host may ask for instead (TI_OPT_SMEM_HEADROOM_KB) when other libraries' CTAs run beside the remap grids.
                                              // Measured for NCCL's send/recv at N = 4 and 8: no help (0.34-0.40 of the job without exchange either way), so nothing sets it
constexpr int P4_LUT_ROW_WORDS = 192;  // per tile row: 32 lanes x 2 pairs x {window word, pixel a word, pixel b word}
constexpr uint32_t P4_EXC_UNUSED = 0x80000000u;  // last word of an unused exception entry (no destination offset is -2^31)
constexpr int P4_MAX_EXC = 32;         // exception entries per (tile, warp): one lane each in the per-frame fix-up pass
constexpr int P4_MAX_EXC_QUAD = 24;    // the same in the quad layout: 2 x 24 x 8 x 16 bytes of tables leave room for a 5-stage ring of 40-row boxes
                                       // beside three CTAs per SM (measured on the bench rig: 4 stages 0.765 of the copy peak, 5 stages 0.81)

// 3-channel window path ("c3": BGR8 -> RGB8 rectified): tiles of C3_TW x C3_TH output pixels, one TMA box of
// 32-bit elements per tile-frame, one 12-byte window per pixel and source row (see ti_rectify_c3.cu).
constexpr int C3_TW = 128;
constexpr int C3_TH = 16;
constexpr int C3_PITCH = 512;          // bytes per staged source row (TMA box of 128 u32 elements)
constexpr int C3_PITCH_WIDE = 1024;    // 256 u32 elements, the most a TMA box dimension holds: maps that sample up to ~2.5 x as many source pixels
constexpr int C3_MAX_ROWS = 64;
constexpr int C3_CONSUMER_WARPS = 8;
constexpr int C3_THREADS = (C3_CONSUMER_WARPS + 1) * 32;
constexpr int C3_MAX_STAGES = 8;
constexpr int C3_LUT_ROW_WORDS = 256;  // per tile row: 128 pixels x {window word, pixel word}

struct CameraSlot {
    // rectification
    std::vector<lut_t> h_lut;                       // host copy of the generic LUT (lazy builds of further tables)
    int h_lut_pitch = 0;
    bool c3_tried = false, has_c3 = false;          // built on the first BGR8 -> RGB8 rectify of the slot
    uint32_t* d_lut5 = nullptr;                     // tiles * C3_TH * C3_LUT_ROW_WORDS
    TileBox2* d_boxes5 = nullptr;                   // c0 in BYTES of the BGR row (multiple of 16)
    int tiles5_x = 0, tiles5_y = 0, rows5_alloc = 0, pitch5 = 0;
    bool has_pair[P4_N_TH] = {false, false, false};              // tile height 16, 32, 24
    uint32_t* d_lut4[P4_N_TH] = {nullptr, nullptr, nullptr};     // tiles * TH * P4_LUT_ROW_WORDS
    TileBox2* d_boxes4[P4_N_TH] = {nullptr, nullptr, nullptr};
    uint32_t* d_exc4[P4_N_TH] = {nullptr, nullptr, nullptr};     // tiles * P4_CONSUMER_WARPS * exc4_per_warp entries of 4 words
    int tiles4_x[P4_N_TH] = {0, 0, 0}, tiles4_y[P4_N_TH] = {0, 0, 0};
    int rows4_alloc[P4_N_TH] = {0, 0, 0};
    int pitch4[P4_N_TH] = {0, 0, 0};
    bool quad4[P4_N_TH] = {false, false, false};                 // quad layout: a lane owns 4 consecutive pixels that share ONE 8-byte window per source row                             // P4_PITCH or P4_PITCH_WIDE: bytes per staged source row
    int exc4_per_warp[P4_N_TH] = {0, 0, 0};
    uint32_t* d_over4[P4_N_TH] = {nullptr, nullptr, nullptr};    // output pixels whose (tile, warp) exception list was full
    int n_over4[P4_N_TH] = {0, 0, 0};
    bool has_tma_mono[3] = {false, false, false};   // tile height 16, 32, 24 (index = M3_TH_INDEX)
    uint32_t* d_lut3[3] = {nullptr, nullptr, nullptr};
    TileBox2* d_boxes3[3] = {nullptr, nullptr, nullptr};
    int tiles3_x[3] = {0, 0, 0}, tiles3_y[3] = {0, 0, 0};
    int rows3_alloc[3] = {0, 0, 0};
    bool has_map = false;
    bool has_fast_mono = false;
    uint32_t* d_lut2 = nullptr;     // tiles * M2_TH * M2_TW, tile-major
    TileBox2* d_boxes2 = nullptr;   // tiles
    int tiles2_x = 0, tiles2_y = 0;
    int rows2_max = 0;
    int dst_w = 0, dst_h = 0, src_w = 0, src_h = 0;
    int tiles_x = 0, tiles_y = 0;
    lut_t* d_lut = nullptr;        // dst_h * dst_w
    TileBox* d_boxes = nullptr;    // tiles_y * tiles_x
    uint8_t* d_valid = nullptr;    // dst_h * dst_w
    size_t tile_smem[2] = {0, 0};  // largest staged box in bytes for 1- and 3-channel sources
    // depth -> RGB registration (ti_register.cu)
    bool has_reg = false;
    int reg_dw = 0, reg_dh = 0, reg_rw = 0, reg_rh = 0;
    float reg_a[9];  // R_rgb<-depth * diag(1/fx_d, 1/fy_d, 1), row-major
    float reg_t[3];
    float reg_k[6];  // cx_d, cy_d, fx_r, fy_r, cx_r, cy_r
    // projection
    bool has_proj = false;
    int proj_w = 0, proj_h = 0;
    // p = d_mm * (au * u + av * v + ac) + t : rows of 1e-3 * R * diag(1/fx, 1/fy, 1) with the principal point folded into ac
    double proj_au[3], proj_av[3], proj_ac[3], proj_t[3];
};

// remembered TMA descriptors (ti_tma.cu): key = everything cuTensorMapEncodeTiled is given
struct TmaKey {
    const void* base;
    uint64_t pitch_y, pitch_z;
    int elem_bytes, w, h, n, box_x, box_y;
};
struct TmaBlob { unsigned char bytes[128]; };

}  // namespace ti

struct ti_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int ctas_per_sm = 0;
    int debug = 0;
    int mono_variant = 4;   // 4: pair-window TMA kernel, 3: TMA-pipelined kernel with shifted copy, 2: thread-staged kernel, 1: generic only
    int tma_tile_h = 32;    // 16, 24 or 32
    int lut_prefetch = 0;   // 1: consumers prefetch the next unit's LUT into a second register set (costs 16 registers)
    int stages = 2;           // shared-memory ring depth of the TMA-pipelined kernel (2 stages -> 4 CTAs per SM)
    int stages4 = 6;          // ring depth of the pair-window kernels (reduced per launch until three CTAs fit an SM)
    int frames_per_unit = 16;  // frames of the batch that share one LUT fetch in the TMA-pipelined kernel
    int frames_per_unit4 = 0;  // same for the pair-window kernel; 0 = chosen per launch (least tail over the persistent grid)
    bool force_generic_rectify = false;  // tests: exercise the generic tiled / direct kernels
    ti::CameraSlot cams[TI_MAX_CAMERAS];
    // scratch for two-pass paths (BGR -> gray ahead of the mono remap); grown on demand, never visible to the caller
    void* scratch = nullptr;
    size_t scratch_cap = 0;
    int smem_headroom_kb = ti::P4_SMEM_HEADROOM_KB;  // TI_OPT_SMEM_HEADROOM_KB: shared memory per SM the persistent window kernels leave free
    bool rectify_quad = true;  // calibration upload tries the quad layout of the pair-window kernel first
    int quad_exc_cap = ti::P4_MAX_EXC_QUAD;  // TI_OPT_RECTIFY_QUAD values 2..32 set it (bring-up: ring depth against overflow pixels)
    int l2_scratch_kb = 0;  // two-pass rectify: scratch per chunk of the batch; 0 = the whole batch in one chunk (chunks that fit the
                            // L2 were measured SLOWER: 0.39 vs 0.49 of peak for BGR8 -> MONO8 - small launches cost more than the re-read)
    // host pipeline (ti_ingest_host)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_exec = nullptr;
    struct HostSlot {
        std::vector<void*> d_src, d_dst, d_mask;
        std::vector<uint32_t*> d_count;
        std::vector<size_t> cap_src, cap_dst, cap_mask, cap_count;
        cudaEvent_t h2d_done = nullptr, exec_done = nullptr, d2h_done = nullptr;
    } hslot[3];  // three chunk slots in flight: upload / kernels / download of consecutive chunks never share buffers
    bool host_ready = false;     // streams and events below exist
    uint64_t host_chunks = 0;    // chunks enqueued so far; chunk g uses slot g % 3
    uint64_t host_tickets = 0;   // submissions so far (ti_ingest_host_submit); ticket t completes at ticket_done[t % 8]
    cudaEvent_t ticket_done[8] = {};
    std::vector<std::pair<ti::TmaKey, ti::TmaBlob>> tma_cache;
    // voxel down-sampling (ti_voxel.cu)
    double voxel_size = 0.0;
    uint32_t voxel_max_depth = 65535;
    uint64_t* voxel_table = nullptr;  // hash set of the launch in flight; entries carry the epoch of the launch that wrote them
    uint64_t voxel_slots = 0;
    uint32_t voxel_epoch = 0;
    // NCCL (dlopen)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    // exchange step on its own stream (ti_nccl.cu)
    cudaStream_t s_comm = nullptr;
    cudaEvent_t ev_compute = nullptr, ev_gather = nullptr, ev_counts = nullptr;
    uint32_t* d_comm_words = nullptr;  // 256 words: gathered counts, push reservations, the barrier's zero
    uint32_t* h_comm_words = nullptr;  // pinned mirror
    bool gather_pending = false, counts_pending = false;
    cudaEvent_t ev_fence[16] = {};  // ti_exchange_fence ring
    uint64_t fences = 0;
    int push_blocks = 0;  // CTAs of the peer-store copy kernels (0 = one per SM)
    int push_tma = 1;     // 1: the peer copy is issued as TMA bulk copies by one lane per CTA; 0: 16-byte stores by small CTAs
};

void ti_nccl_teardown(ti_ctx* ctx);  // ti_nccl.cu

namespace ti {

int fail(ti_ctx* ctx, int code, const char* fmt, ...);
void set_global_error(const char* msg);

#define TI_CUDA(ctx, expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ti::fail((ctx), TI_ECUDA, "%s failed: %s (%s:%d)", #expr,                   \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                       \
    } while (0)

#ifdef TI_EMULATE
#define TI_CHECK_LAUNCH(ctx) ((ctx)->launches++)
#else
#define TI_CHECK_LAUNCH(ctx)                                                                   \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return ti::fail((ctx), TI_ECUDA, "kernel launch failed: %s (%s:%d)",               \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                       \
        (ctx)->launches++;                                                                     \
    } while (0)
#endif

// Resident CTAs per SM of a persistent kernel: grids are sized to exactly one wave.  The last answer per
// (kernel, threads, shared memory) is remembered - the query costs microseconds that matter for one-frame-set calls.
template <typename K>
inline int resident_ctas(K kernel, int threads, size_t dyn_smem, int fallback) {
#ifdef TI_EMULATE
    (void)kernel; (void)threads; (void)dyn_smem;
    return fallback > 2 ? 2 : fallback;
#else
    struct Memo { const void* k; int threads; size_t smem; int n; };
    static thread_local Memo memo[8] = {};
    static thread_local int next = 0;
    const void* key = reinterpret_cast<const void*>(kernel);
    for (const Memo& m : memo)
        if (m.k == key && m.threads == threads && m.smem == dyn_smem) return m.n;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fallback;
    }
    memo[next] = Memo{key, threads, dyn_smem, n};
    next = (next + 1) % 8;
    return n;
#endif
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the kernel needs more than it was last given
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kernel, size_t dyn_smem, int device) {
#ifdef TI_EMULATE
    (void)kernel; (void)dyn_smem; (void)device;
    return cudaSuccess;
#else
    struct Memo { const void* k; size_t smem; int device; };
    static thread_local Memo memo[16] = {};
    const void* key = reinterpret_cast<const void*>(kernel);
    Memo* slot = nullptr;
    for (Memo& m : memo) {
        if (m.k == key && m.device == device) { slot = &m; break; }
        if (!m.k && !slot) slot = &m;
    }
    if (slot && slot->k == key && slot->smem >= dyn_smem) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    if (e == cudaSuccess && slot) { slot->k = key; slot->smem = dyn_smem; slot->device = device; }
    return e;
#endif
}

inline int channels_of(int fmt) {
    switch (fmt) {
        case TI_FMT_MONO8: return 1;
        case TI_FMT_BGR8:
        case TI_FMT_RGB8: return 3;
        default: return 0;
    }
}

// bytes of one frame in `fmt`
inline uint64_t frame_bytes(int fmt, int w, int h) {
    switch (fmt) {
        case TI_FMT_MONO8: return (uint64_t)w * h;
        case TI_FMT_BGR8:
        case TI_FMT_RGB8: return (uint64_t)w * h * 3;
        case TI_FMT_NV12: return (uint64_t)w * h * 3 / 2;
        case TI_FMT_DEPTH16: return (uint64_t)w * h * 2;
        case TI_FMT_XYZ32F: return (uint64_t)w * h * 12;
        default: return 0;
    }
}

// launchers implemented in the kernel translation units -------------------------------------
struct ConvertJob {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t src_stride, dst_stride;
    int width, height, src_fmt, dst_fmt;
};
int launch_convert(ti_ctx* ctx, const ConvertJob* jobs, int n_jobs, int n_batch);

struct RectifyJob {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t src_stride, dst_stride;
    int camera, src_fmt, dst_fmt;
};
int launch_rectify(ti_ctx* ctx, const RectifyJob* jobs, int n_jobs, int n_batch);
// pair-window tables of one camera from its generic LUT (ti_rectify_pair.cu); frees / replaces the old ones
int build_pair_tables(ti_ctx* ctx, CameraSlot& C, const std::vector<lut_t>& lut, int lut_pitch);
void free_pair_tables(CameraSlot& C);
// 3-channel window tables (ti_rectify_c3.cu), built lazily from CameraSlot::h_lut
int build_c3_tables(ti_ctx* ctx, CameraSlot& C);
void free_c3_tables(CameraSlot& C);

struct BackprojectJob {
    const uint16_t* depth;
    float* xyz;
    uint8_t* mask;
    uint32_t* count;
    uint64_t depth_stride, xyz_stride, mask_stride;
    int camera;
    // fused depth -> RGB registration (both NULL: none): one RGB8 colour per depth pixel from `rgb` (the slot's registration)
    const uint8_t* rgb = nullptr;
    uint8_t* colour = nullptr;
    uint64_t rgb_stride = 0, colour_stride = 0;
};
int launch_backproject(ti_ctx* ctx, const BackprojectJob* jobs, int n_jobs, int n_batch);
int launch_depth_stats(ti_ctx* ctx, const uint16_t* depth, int width, int height, int n_batch, uint64_t stride, uint32_t* out);
int launch_register_colour(ti_ctx* ctx, int camera, const uint16_t* depth, const uint8_t* rgb, uint8_t* colour, int n_batch,
                           uint64_t depth_stride, uint64_t rgb_stride, uint64_t colour_stride);

#if defined(TI_EMULATE)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) { return *reinterpret_cast<const uint4*>(ti_emu::check_align(p, 16)); }
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) { return *reinterpret_cast<const uint2*>(ti_emu::check_align(p, 8)); }
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) { *reinterpret_cast<uint4*>(ti_emu::check_align(p, 16)) = v; }
__device__ __forceinline__ void st_stream_u2(void* p, uint2 v) { *reinterpret_cast<uint2*>(ti_emu::check_align(p, 8)) = v; }
__device__ __forceinline__ void st_stream_u1(void* p, uint32_t v) { *reinterpret_cast<uint32_t*>(ti_emu::check_align(p, 4)) = v; }
__device__ __forceinline__ uint4 ld_keep_u4(const void* p) { return *reinterpret_cast<const uint4*>(ti_emu::check_align(p, 16)); }
__device__ __forceinline__ void st_stream_b8(void* p, uint32_t v) { *reinterpret_cast<uint8_t*>(p) = (uint8_t)v; }
__device__ __forceinline__ void st_stream_b16(void* p, uint32_t v) { *reinterpret_cast<uint16_t*>(ti_emu::check_align(p, 2)) = (uint16_t)v; }
__device__ __forceinline__ void st_stream_b32(void* p, uint32_t v) { *reinterpret_cast<uint32_t*>(ti_emu::check_align(p, 4)) = v; }
#elif defined(__CUDACC__)
// ---- device helpers ------------------------------------------------------------------------
// L2 eviction policies (createpolicy is not volatile: the compiler hoists / CSEs it).
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Streaming accesses: inputs are read once, outputs written once -> keep them out of L1 and mark
// them evict-first in L2 so the (re-used) remap LUT stays resident.
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy_evict_first()));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p), "l"(policy_evict_first()));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p),
                 "r"(v.x), "r"(v.y), "l"(policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_stream_u1(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v),
                 "l"(policy_evict_first())
                 : "memory");
}
__device__ __forceinline__ void st_stream_b8(void* p, uint32_t v) {  // .cs = streaming (evict-first), no policy descriptor
    asm volatile("st.global.cs.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_b16(void* p, uint32_t v) {
    asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void st_stream_b32(void* p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// LUT reads: re-used by every frame of the batch -> prefer to keep in L2.
__device__ __forceinline__ uint4 ld_keep_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy_evict_last()));
    return r;
}
#endif

}  // namespace ti
