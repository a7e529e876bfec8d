// Pair-window mono remap ("v4"): launch parameters shared by ti_rectify.cu (job routing) and
// ti_rectify_pair.cu (kernel, tables).
#pragma once
#include "ti_common.cuh"
#include "ti_tma.cuh"

namespace ti {

constexpr int MAX_PAIR_JOBS = 16;

struct Rect4JobDev {
    const uint32_t* lut4;    // tiles x TH x P4_LUT_ROW_WORDS
    const TileBox2* boxes4;  // tiles
    const uint32_t* exc4;    // tiles x P4_CONSUMER_WARPS x exc_per_warp entries of 16 bytes
    uint8_t* dst;
    uint64_t dst_stride;
    int dst_w, dst_h;
    int rows_alloc;          // rows of this job's TMA box
    int exc_per_warp;        // 0 .. P4_MAX_EXC
    uint32_t tile_begin;
};

struct Rect4Params {
    TiTensorMap map[MAX_PAIR_JOBS];
    Rect4JobDev job[MAX_PAIR_JOBS];
    uint32_t tiles_per_set;
    int n_jobs;
    int n_batch;
    int frames_per_unit;
    int rows_alloc_max;  // launch-wide stage geometry
    int pitch;           // P4_PITCH or P4_PITCH_WIDE: every job of a launch stages rows of this pitch
    int quad;            // 1: every job of the launch uses the quad layout (4 consecutive pixels per lane, one window per row)
    int exc_max;         // largest exc_per_warp of the launch
    int stages;
    int debug;           // bring-up switches (TI_OPT_DEBUG): 1 = consumers skip the blend, 2 = issuer skips the loads; 0 in production
};

// th_index: index into P4_TILE_HEIGHTS
int launch_rectify_pair(ti_ctx* ctx, Rect4Params& P, int th_index);

// ---- 3-channel window remap (ti_rectify_c3.cu) --------------------------------------------------------
struct Rect5JobDev {
    const uint32_t* lut5;    // tiles x C3_TH x C3_LUT_ROW_WORDS
    const TileBox2* boxes5;  // tiles
    uint8_t* dst;
    uint64_t dst_stride;
    int dst_w, dst_h;
    int rows_alloc;
    uint32_t tile_begin;
};

struct Rect5Params {
    TiTensorMap map[MAX_PAIR_JOBS];
    Rect5JobDev job[MAX_PAIR_JOBS];
    uint32_t tiles_per_set;
    int n_jobs;
    int n_batch;
    int frames_per_unit;
    int rows_alloc_max;
    int pitch;  // C3_PITCH or C3_PITCH_WIDE
    int stages;
};

int launch_rectify_c3(ti_ctx* ctx, Rect5Params& P);

}  // namespace ti
