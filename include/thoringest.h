/*
 * thoringest.h - C ABI of the B200 ingest library (libthoringest.so).
 *
 * This is the drop-in boundary for ONE path of WT-MM/thor-slam: the per-frame-set
 * ingest stage between CameraRig.get_synchronized_frames() and the SLAM / mapping
 * consumers.  The reference is pure Python and has no FFI of its own; each entry
 * point below names the reference code whose per-pixel work it takes over
 * (paths relative to the reference repo root).  Host code binds these with ctypes
 * (thor_slam_b200/ingest/_lib.py; the stub a reference maintainer would add is in
 * INTEGRATION.md).
 *
 * Rules of the boundary
 *  - plain C types only; every pointer argument is either a HOST pointer or a
 *    DEVICE pointer as stated per argument - no torch / numpy types cross it;
 *  - the library never allocates a buffer the caller sees, except the explicitly
 *    named peer buffers (ti_peer_alloc); callers (PyTorch tensors in our host code)
 *    own all image / cloud memory and keep it alive until ti_sync() returns;
 *  - every function returns a ti_status; ti_last_error() gives the message;
 *  - one ti_ctx per (process, GPU); a ctx is not thread-safe (the rig lock that
 *    serialises CameraRig - thor_slam/camera/rig.py:114 - serialises it too);
 *  - all kernels are enqueued on the stream given to ti_set_stream() (default: the
 *    legacy NULL stream) and are asynchronous unless stated otherwise;
 *  - there is NO CPU fallback: without a CUDA device ti_create() fails.
 */
#ifndef THORINGEST_H_
#define THORINGEST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TI_ABI_VERSION 2 /* round 2: voxel clouds, the exchange stream, peer inboxes, fused colour, depth statistics */
#define TI_MAX_CAMERAS 64 /* calibration slots per context                     */
#define TI_MAX_STREAMS 32 /* streams of one kind per ti_ingest() call          */
#define TI_MAX_DIM 8190   /* max source width/height of a remap slot (the driver's largest sensor mode is 4224 x 3136,
                             thor_slam/camera/drivers/luxonis.py:36-44) */

typedef struct ti_ctx ti_ctx;

typedef enum ti_status {
    TI_OK = 0,
    TI_EINVAL = 1, /* bad argument (maps to ValueError on the Python side)      */
    TI_ECUDA = 2,  /* a CUDA call failed (RuntimeError)                         */
    TI_ENCCL = 3,  /* an NCCL call failed / NCCL not loadable (RuntimeError)    */
    TI_ESTATE = 4, /* call order violated, e.g. camera slot not uploaded        */
    TI_ENOMEM = 5
} ti_status;

/* Wire formats.  MONO8: HxW u8.  BGR8 / RGB8: HxWx3 u8 interleaved.  NV12: (H*3/2)xW u8,
 * luma plane then interleaved U,V at half resolution.  DEPTH16: HxW u16 millimetres,
 * 0 = invalid.  XYZ32F: HxWx3 f32 metres.  Rows are tightly packed. */
typedef enum ti_format {
    TI_FMT_MONO8 = 0,
    TI_FMT_BGR8 = 1,
    TI_FMT_RGB8 = 2,
    TI_FMT_NV12 = 3,
    TI_FMT_DEPTH16 = 4,
    TI_FMT_XYZ32F = 5
} ti_format;

typedef enum ti_kind {
    TI_KIND_CONVERT = 0,     /* format conversion only                                   */
    TI_KIND_RECTIFY = 1,     /* format conversion fused with the undistort/rectify remap */
    TI_KIND_BACKPROJECT = 2  /* depth -> body-frame xyz + valid mask + valid count       */
} ti_kind;

/* One stream of a frame-set batch.  Frame b of the batch lives at
 * src + b*src_frame_stride / dst + b*dst_frame_stride (bytes). */
typedef struct ti_stream {
    int32_t kind;       /* ti_kind                                                         */
    int32_t camera;     /* calibration slot (RECTIFY: remap LUT, BACKPROJECT: projection)   */
    int32_t src_format; /* ti_format of src                                                */
    int32_t dst_format; /* ti_format of dst                                                */
    int32_t width;      /* source width in pixels                                          */
    int32_t height;     /* source height in pixels (NV12: luma rows)                       */
    const void* src;    /* DEVICE (ti_ingest) or HOST (ti_ingest_host) pointer, 16-B aligned */
    void* dst;          /* likewise                                                        */
    uint64_t src_frame_stride;
    uint64_t dst_frame_stride;
    void* mask;               /* BACKPROJECT: u8 HxW valid mask per frame, or NULL          */
    uint64_t mask_frame_stride;
    uint32_t* count;          /* BACKPROJECT: one u32 valid-pixel count per frame, or NULL  */
} ti_stream;

/* ---- context ------------------------------------------------------------------------- */
int ti_abi_version(void);
int ti_create(int device, ti_ctx** out);
int ti_destroy(ti_ctx* ctx);
/* Message of the last failing call on ctx (ctx == NULL: last failing ti_create). */
const char* ti_last_error(const ti_ctx* ctx);
/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream). */
int ti_set_stream(ti_ctx* ctx, void* cuda_stream);
int ti_sync(ti_ctx* ctx);
/* Kernel launches issued by this ctx since creation (for bench.py "gpu_launches"). */
uint64_t ti_launch_count(const ti_ctx* ctx);
/* Tuning / test switches (never change results). */
typedef enum ti_option {
    TI_OPT_FORCE_GENERIC_RECTIFY = 1, /* 1: skip the fast mono remap kernel, use the generic ones */
    TI_OPT_CTAS_PER_SM = 2,           /* >0: resident CTAs per SM the persistent grids are sized for */
    TI_OPT_MONO_VARIANT = 3,          /* mono remap kernel: 4 pair-window TMA (default), 3 TMA + shifted copy, 2 thread-staged, 1 generic */
    TI_OPT_TMA_TILE_H = 4,            /* output tile height of the TMA-pipelined kernel: 16, 24 or 32 (default) */
    TI_OPT_DEBUG = 5,                 /* bring-up switches; 0 in production (non-zero MAY change results) */
    TI_OPT_FRAMES_PER_UNIT = 6,       /* frames of a batch sharing one LUT fetch in the TMA kernels (default 16; pair-window: 0 = automatic) */
    TI_OPT_STAGES = 7,                /* shared-memory ring depth of the TMA kernels, 2..8 (default 6 pair-window, reduced to fit; 2 shifted-copy) */
    TI_OPT_LUT_PREFETCH = 8,          /* 1: consumers prefetch the next unit's LUT into a second register set (default 0) */
    TI_OPT_L2_SCRATCH_KB = 10,        /* two-pass rectify (BGR8 -> MONO8, NV12 -> RGB8): KB of intermediate frames per chunk of the batch (default 0: one chunk) */
    TI_OPT_PUSH_TMA = 11,             /* peer copy of ti_cloud_push: 1 = TMA bulk copies issued by one lane per CTA (default), 0 = 16-byte stores */
    TI_OPT_RECTIFY_QUAD = 12,         /* pair-window kernel: 1 = maps uploaded from now on try the quad layout (4 pixels per lane and window) first (default), 0 = pairs; 2..32 = quad with that many exception entries per (tile, warp) (default 24) */
    TI_OPT_SMEM_HEADROOM_KB = 13,     /* KB of shared memory per SM the persistent remap kernels leave to kernels running beside them (default 20: room for the library's own exchange kernels) */
    TI_OPT_PUSH_BLOCKS = 9            /* CTAs of the peer-store copy kernels of ti_cloud_push / ti_inbox_take (default: one per SM) */
} ti_option;
int ti_set_option(ti_ctx* ctx, int option, int value);
int ti_device_sm_count(const ti_ctx* ctx);

/* ---- calibration upload (one-off; replaces nothing per-frame in the reference:
 *      Intrinsics/Extrinsics of thor_slam/camera/types.py:31-69 become device constants) */

/* Remap LUT of calibration slot `camera` from OpenCV-style float maps (HOST pointers,
 * dst_w*dst_h floats each, as produced by cv2.initUndistortRectifyMap(..., CV_32FC1)):
 * output pixel (u,v) samples the source image at (mapx[v][u], mapy[v][u]).
 * Quantised exactly like cv2.remap: ix = rint(mapx*32) (half-to-even), 1/32-px taps. */
int ti_upload_rectify_map(ti_ctx* ctx, int camera, int dst_w, int dst_h, int src_w, int src_h,
                          const float* mapx, const float* mapy);

/* Pinhole projection + pose of calibration slot `camera` for back-projection.
 * k = {fx, fy, cx, cy} of the DEPTH image (thor_slam/camera/drivers/luxonis.py:974-1066);
 * body_T_cam = row-major 3x4 [R|t], metres, = M * world_T_camera with world_T_camera from
 * RigCalibration.get_world_extrinsics (thor_slam/camera/rig.py:35-70) and M = RDF_TO_FLU
 * (thor_slam/slam/adapters/isaac_ros.py:42-49) or identity.  float64 in, rounded once. */
int ti_upload_projection(ti_ctx* ctx, int camera, int width, int height, const double k[4],
                         const double body_T_cam[12]);

/* Which remap kernels ti_rectify()/ti_ingest() will run on slot `camera` under the current options
 * (diagnostic; tests use it to prove there is no silent fall-back to a slower kernel).
 * MONO8/NV12 -> MONO8 (and BGR8 -> MONO8 after its gray pre-pass): out[0] = kernel variant (4 pair-window,
 * 3 TMA + shifted copy, 2 thread-staged, 1 generic), out[1] = output tile height, out[2] = source rows staged
 * per tile, out[3] = exception entries per (tile, warp) of the pair-window kernel.
 * BGR8 -> RGB8: out[4] = 5 (3-channel window kernel) or 1 (generic), out[5] = source rows staged per tile.
 * out[6] = output pixels of the pair-window kernel repaired by the per-pixel pass after it (a (tile, warp) holds 32 exceptions;
 * strongly bent maps - fisheye - have a few more), out[7] = bytes per staged source row of the pair-window kernel (192, or
 * 320 for maps whose tiles span more source pixels, e.g. a 2 x downscale) in its low 16 bits, | 0x10000 when the slot runs the
 * kernel's quad layout (four output pixels per lane and window instead of two). */
int ti_rectify_plan(ti_ctx* ctx, int camera, int32_t out[8]);

/* u8 dst_h x dst_w mask of slot `camera`: 1 where all four bilinear taps are inside the
 * source image.  Static per calibration.  dst: DEVICE pointer. */
int ti_get_valid_mask(ti_ctx* ctx, int camera, uint8_t* dst);

/* ---- per-frame work, DEVICE pointers --------------------------------------------------- */

/* Format conversion of n_batch frames.  Takes over cv2.cvtColor(BGR2RGB)
 * (thor_slam/slam/adapters/isaac_ros.py:357, scripts/run_pipeline.py:234) and the NV12 ->
 * BGR/GRAY conversion inside dai.ImgFrame.getCvFrame() (thor_slam/camera/drivers/luxonis.py:773).
 * Supported: BGR8->RGB8, BGR8->MONO8, NV12->MONO8, NV12->RGB8, NV12->BGR8, MONO8->MONO8. */
int ti_convert(ti_ctx* ctx, int src_format, int dst_format, const void* src, void* dst, int width,
               int height, int n_batch, uint64_t src_frame_stride, uint64_t dst_frame_stride);

/* Conversion fused with the bilinear remap of slot `camera` (the undistortion the reference
 * delegates to cuVSLAM with rectified_images:=false - isaac_ros.py:364-411, Makefile:77-80).
 * src has the slot's src_w x src_h, dst its dst_w x dst_h.  Bit-exact with
 * cv2.remap(cvtColor(src), INTER_LINEAR, BORDER_CONSTANT 0).  Supported: MONO8->MONO8,
 * NV12->MONO8, BGR8->MONO8, BGR8->RGB8, NV12->RGB8. */
int ti_rectify(ti_ctx* ctx, int camera, int src_format, int dst_format, const void* src, void* dst,
               int n_batch, uint64_t src_frame_stride, uint64_t dst_frame_stride);

/* depth (u16 mm) -> xyz (f32 x3, body frame, invalid = 0,0,0) + mask (u8) + per-frame count
 * (u32, overwritten).  Takes over the back-projection the reference leaves to nvblox
 * (scripts/run_pipeline.py:247-256) and examples/rgbd_stream.py:121-123,270-276 (mask/count).
 * mask and count may be NULL. */
int ti_backproject(ti_ctx* ctx, int camera, const uint16_t* depth, float* xyz, uint8_t* mask,
                   uint32_t* count, int n_batch, uint64_t depth_frame_stride,
                   uint64_t xyz_frame_stride, uint64_t mask_frame_stride);

/* Depth -> RGB registration constants of slot `camera` (thor_slam/camera/drivers/luxonis.py:1018-1091: depth and
 * RGB intrinsics of get_rgbd_intrinsics(), rgb_T_depth = inv(rgb extrinsics) * depth extrinsics from
 * get_rgbd_extrinsics()).  k_* = {fx, fy, cx, cy}; rgb_T_depth: row-major 3x4, metres.  float64 in, rounded once. */
int ti_upload_registration(ti_ctx* ctx, int camera, int depth_w, int depth_h, const double k_depth[4], int rgb_w,
                           int rgb_h, const double k_rgb[4], const double rgb_T_depth[12]);

/* One RGB8 colour per depth pixel: back-project with the depth K, move into the RGB camera, project with the RGB K,
 * take the nearest RGB pixel (round-half-even); (0,0,0) where depth is 0, the point is behind the RGB camera or falls
 * outside the image.  depth: u16 HxW; rgb: RGB8 of the registered size; colour: u8 HxWx3 (depth size).  Strides in
 * bytes, 0 = tightly packed.  The lookup nvblox does per voxel (scripts/run_pipeline.py:218-256) done once per pixel. */
int ti_register_colour(ti_ctx* ctx, int camera, const uint16_t* depth, const uint8_t* rgb, uint8_t* colour, int n_batch,
                       uint64_t depth_frame_stride, uint64_t rgb_frame_stride, uint64_t colour_frame_stride);

/* ---- voxel down-sampling of the body-frame cloud ---------------------------------- */

/* Voxel grid of the rig-wide cloud.  voxel_size_m: edge length in metres (nvblox `voxel_size`, launch/thor_nvblox.launch.py:26:
 * 0.05).  max_depth_mm: depth pixels beyond it are not integrated (`tsdf_integrator_max_integration_distance_m`,
 * launch/thor_nvblox.launch.py:27-31: 10 m); 0 = no cap. */
int ti_set_voxel_grid(ti_ctx* ctx, double voxel_size_m, uint32_t max_depth_mm);

typedef struct ti_depth_stream {
    int32_t camera;              /* calibration slot with a projection (ti_upload_projection)     */
    int32_t reserved;
    const uint16_t* depth;       /* DEVICE, u16 millimetres, frame b at depth + b*depth_frame_stride bytes */
    uint64_t depth_frame_stride; /* 0 = tightly packed                                              */
} ti_depth_stream;

/* depth frames of n_streams cameras x n_batch frame sets -> ONE list with one u64 record per voxel that at least one valid
 * pixel (0 < d <= max_depth_mm) of ANY camera of frame set b falls into - the valid-compacted, down-sampled form of the clouds
 * ti_backproject writes densely (same double-precision back-projection, then k = floor(p_body / voxel_size) per axis):
 *     record = tag << 56 | (set_base + b) << 45 | (kx + 16384) << 30 | (ky + 16384) << 15 | (kz + 16384)
 * records: DEVICE u64[capacity], appended in no defined order; *n_records (DEVICE u32, overwritten) = number of distinct
 * records found - when it exceeds capacity the list was truncated (and the count is then only a lower bound: the kernel stops
 * extending a list that has grown past twice its capacity).  set_counts: DEVICE u32[n_batch] (overwritten) or NULL.
 * tag < 256 (the producing rank in the multi-GPU gather), set_base + n_batch <= 2048.  One call at a time per ctx.
 * This is the cloud the reference's only cloud type carries (thor_slam/slam/interface.py:134-138, N x 3) before ti_voxel_points. */
int ti_voxel_cloud(ti_ctx* ctx, const ti_depth_stream* streams, int n_streams, int n_batch, uint32_t set_base, uint32_t tag,
                   uint64_t* records, uint64_t capacity, uint32_t* n_records, uint32_t* set_counts);

/* Records -> voxel centres, N x 3 f32 metres in the body frame (SlamMap.to_point_cloud, thor_slam/slam/interface.py:134-138).
 * Converts min(*n_records, max_records) records; xyz: DEVICE f32[max_records * 3]. */
int ti_voxel_points(ti_ctx* ctx, const uint64_t* records, const uint32_t* n_records, uint64_t max_records, float* xyz);

/* The viewer's depth statistics (examples/rgbd_stream.py:270-276: valid count, mean, min, max of depth > 0), one pass over
 * n_batch depth frames: stats[b] = six u32 {count, min, max, 0, sum low word, sum high word} (DEVICE; min is 0xFFFFFFFF for a
 * frame without a valid pixel; mean = sum / count is the caller's division).  depth_frame_stride 0 = tightly packed. */
int ti_depth_stats(ti_ctx* ctx, const uint16_t* depth, int width, int height, int n_batch, uint64_t depth_frame_stride,
                   uint32_t* stats);

/* ti_backproject and ti_register_colour in ONE pass over the depth image ("fused into the
 * back-projection kernel"): xyz + mask + count as ti_backproject, plus colour (u8 HxWx3, depth size) as ti_register_colour
 * would give it - bit-identical selection - without reading the depth a second time.  rgb: RGB8 of the registered size.
 * Strides in bytes; rgb / colour strides 0 = tightly packed.  Needs ti_upload_projection and ti_upload_registration of the
 * same depth size on `camera`. */
int ti_backproject_colour(ti_ctx* ctx, int camera, const uint16_t* depth, const uint8_t* rgb, float* xyz, uint8_t* mask,
                          uint32_t* count, uint8_t* colour, int n_batch, uint64_t depth_frame_stride,
                          uint64_t rgb_frame_stride, uint64_t xyz_frame_stride, uint64_t mask_frame_stride,
                          uint64_t colour_frame_stride);

/* Whole frame-set batch in at most one launch per kind: every stream x every frame.
 * This is the call behind CameraRig.get_synchronized_frames() (thor_slam/camera/rig.py:358-415)
 * in the drop-in rig.  streams: HOST array, DEVICE image pointers inside. */
int ti_ingest(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch);

/* Same, but src/dst/mask/count inside `streams` are HOST pointers (pinned for overlap):
 * host->device copy, kernels and device->host copy are pipelined over `chunk` frame sets at a
 * time on internal streams; returns after everything has landed in the dst buffers. */
int ti_ingest_host(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch, int chunk);

/* The same call split in two, for callers that stream batch after batch (a capture loop double-buffering its
 * host frames): submit enqueues the batch behind whatever was submitted before and returns at once with a
 * ticket; wait blocks until that batch has landed in its dst buffers.  Uploads of batch k+1 overlap downloads
 * of batch k.  The host buffers of a batch belong to the library from submit until its wait returns; at most 8
 * tickets may be outstanding (a ninth submit waits for the oldest).  Same role as above: get_synchronized_frames()
 * (thor_slam/camera/rig.py:358-415) called in a loop by SlamEngine.process_frame (thor_slam/slam/interface.py). */
int ti_ingest_host_submit(ti_ctx* ctx, const ti_stream* streams, int n_streams, int n_batch, int chunk, uint64_t* ticket);
int ti_ingest_host_wait(ti_ctx* ctx, uint64_t ticket);

/* Plumbing for hosts that stage frames themselves (the drop-in rig: one pinned slot per stream and queue entry): one asynchronous
 * copy of `bytes` between any two of {pinned host, device} buffers on `cuda_stream` (a cudaStream_t; NULL = the context's stream).
 * Replaces nothing of the reference - its frames never leave the host (thor_slam/camera/rig.py:297-356) - and exists because a
 * framework's stream context switch costs 40 us per copy where this call costs 2. */
int ti_copy_async(ti_ctx* ctx, void* dst, const void* src, uint64_t bytes, void* cuda_stream);

/* ---- multi-GPU: one process per GPU ------------------------------------------------------ */

/* NCCL is dlopen()ed on first use.  id: 128 bytes, created on rank 0, shipped by the caller
 * (torch.distributed object broadcast in our host code). */
int ti_nccl_unique_id(void* id128);
int ti_nccl_init(ti_ctx* ctx, const void* id128, int rank, int world);
/* Gather of per-GPU body-frame clouds to `root` over NVLink: rank r contributes
 * bytes_per_rank[r] bytes from `local` (DEVICE); on root they land back to back in `gathered`
 * (DEVICE, sum of bytes_per_rank) in rank order.  bytes_per_rank: HOST array of `world`.
 * The exchange runs on the context's own EXCHANGE STREAM: it starts after everything enqueued on the ingest stream so far
 * (an event), and work enqueued on the ingest stream afterwards - the next batch's kernels - overlaps it.  Nobody may read
 * `gathered` or overwrite `local` before ti_gather_wait(). */
int ti_gather_clouds(ti_ctx* ctx, const void* local, void* gathered, const uint64_t* bytes_per_rank,
                     int root);
/* Wait for the last exchange enqueued by ti_gather_clouds / ti_cloud_push / ti_inbox_take.  on_stream != 0: the ingest
 * stream waits (device side, returns at once); on_stream == 0: the calling thread blocks. */
int ti_gather_wait(ti_ctx* ctx, int on_stream);
/* Sizes of a variable-length gather: every rank contributes *n_local (DEVICE u32, e.g. the n_records of ti_voxel_cloud);
 * counts (HOST, `world` entries) receives all of them.  All-gather + read-back on the exchange stream; blocks the calling
 * thread until the counts are there (the ingest stream keeps running what was enqueued meanwhile). */
int ti_gather_counts(ti_ctx* ctx, const uint32_t* n_local, uint32_t* counts);
/* The same in two halves, so that the caller can hand the ingest stream its next batch before it blocks: begin enqueues the
 * all-gather behind what the ingest stream holds NOW; finish blocks until the counts are on the host.  One outstanding. */
int ti_gather_counts_begin(ti_ctx* ctx, const uint32_t* n_local);
int ti_gather_counts_finish(ti_ctx* ctx, uint32_t* counts);
/* Variable-length gather of ti_voxel_cloud lists sized by ti_gather_counts: rank r contributes records[0 .. counts[r]); on root
 * they land back to back in `gathered` in rank order.  counts: HOST, `world` entries.  Ordered on the exchange stream behind the
 * ti_gather_counts_begin that sized it, not behind whatever the ingest stream was given since - this is what lets batch k's
 * exchange run under batch k + 1's kernels.  Completion: ti_gather_wait / ti_exchange_fence. */
int ti_gather_records(ti_ctx* ctx, const uint64_t* records, uint64_t* gathered, const uint32_t* counts, int root);
/* A fence names everything enqueued on the exchange stream so far (ring of 16).  ti_exchange_wait: the ingest stream
 * (on_stream != 0) or the calling thread waits for it - e.g. before a record buffer handed to an exchange is overwritten. */
int ti_exchange_fence(ti_ctx* ctx, uint64_t* fence);
int ti_exchange_wait(ti_ctx* ctx, uint64_t fence, int on_stream);
int ti_nccl_barrier(ti_ctx* ctx);

/* Peer-visible cloud buffer: allocated on this GPU, exported as a 64-byte IPC handle, opened
 * on the other ranks so their back-projection kernels store straight into it over NVLink
 * (gather fused into the producing kernel). */
int ti_peer_alloc(ti_ctx* ctx, uint64_t bytes, void** dev_ptr, void* handle64);
int ti_peer_open(ti_ctx* ctx, const void* handle64, void** dev_ptr);
int ti_peer_close(ti_ctx* ctx, void* dev_ptr);
int ti_peer_free(ti_ctx* ctx, void* dev_ptr);


/* ---- the exchange as our own kernels over peer memory (no collective library, no host round trip) --------------------
 * An INBOX is a peer buffer (ti_peer_alloc on the fusing rank, ti_peer_open elsewhere) of TI_INBOX_HEADER_BYTES + 8 *
 * capacity bytes: a header {u32 n_records, done, gen, error} followed by u64 records.  Producers append their ti_voxel_cloud
 * lists with stores that cross NVLink; the root takes one generation at a time.  All calls enqueue on the exchange stream. */
#define TI_INBOX_HEADER_BYTES 128
/* Root, once, before the handle is shared: zero the header (generation 0). */
int ti_inbox_init(ti_ctx* ctx, void* inbox);
/* Append records[0 .. min(*n_records, records_capacity)) (DEVICE; *n_records is read on the device when the ingest stream reaches
 * this point - a truncated ti_voxel_cloud list reports more records than records_capacity holds) to
 * `inbox` (own or peer-mapped) for generation `gen`: waits on the device until the inbox is at `gen`, reserves the slots with
 * one system-scope atomic, copies with peer stores, then reports this rank done.  Records past inbox_capacity are dropped
 * (the header's count still includes them).  Every rank's run starts 16-byte aligned: an odd list is padded with ONE zero
 * record (written into the list's spare slot; a list that fills its buffer to an odd count leaves its last record behind; 0 is
 * not a valid record) - consumers skip zero records.  `records` may be reused after ti_gather_wait() / the fence taken behind this call. */
int ti_cloud_push(ti_ctx* ctx, uint64_t* records, const uint32_t* n_records, uint64_t records_capacity, void* inbox,
                  uint64_t inbox_capacity, uint32_t gen);
/* Root: wait on the device until `world` ranks have reported done for the inbox's current generation, copy
 * min(count, inbox_capacity, dst_capacity) records to dst (DEVICE; dst_capacity 0: no copy - the caller consumed the inbox in
 * place while it was the other slot's turn), write status[0] = count, status[1] = error flag (DEVICE u32[2]; error != 0: a
 * peer missed its 4 s deadline), empty the inbox and advance its generation. */
int ti_inbox_take(ti_ctx* ctx, void* inbox, uint64_t inbox_capacity, uint32_t world, uint64_t* dst, uint64_t dst_capacity,
                  uint32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* THORINGEST_H_ */
