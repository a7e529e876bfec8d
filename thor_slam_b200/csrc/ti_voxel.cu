// depth (u16 mm) of every camera of a frame set -> ONE list of occupied body-frame voxels (SURVEY section 8 (f) row 4,
// "rig-side voxel down-sampling of the cloud ahead of nvblox", and the N x 3 cloud contract of
// thor_slam/slam/interface.py:134-138).
//
// nvblox integrates at voxel_size = 0.05 m and ignores depth beyond tsdf_integrator_max_integration_distance_m = 10 m
// (launch/thor_nvblox.launch.py:26-31).  Shipping 12 B for every depth pixel to the fusing rank carries each voxel of
// a surface ~170 times (a 5 cm voxel at 3 m covers 13 x 13 pixels of a 1280 x 800 / f = 800 px camera).  This kernel
// emits each occupied voxel of a frame set once:
//
//     p   = d * (au * u + av * v + ac) + t          the back-projection of ti_backproject.cu (double, fused multiply-adds,
//                                                   same constants, RDF->FLU and rig pose folded in)
//     k   = floor(p / voxel_size)  per axis         computed as floor(d * (au' * u + av' * v + ac') + (t' + 16384)) - 16384 with
//                                                   the constants pre-multiplied by 1 / voxel_size on the host
//     valid = d > 0  and  d <= max_depth_mm         examples/rgbd_stream.py:121-123 and the nvblox distance cap
//     record = tag << 56 | set << 45 | (kx + 16384) << 30 | (ky + 16384) << 15 | (kz + 16384)
//
// One record per distinct (set, k) over ALL cameras of the call: the per-frame-set rig-wide fusion.  The order of the list
// is not defined (records are appended tile by tile); the SET of records and the per-set counts are exact.
//
// How duplicates are removed (every stage only ever drops a key that is provably already on its way out):
//   1. a lane owns 8 consecutive pixels of a row and forwards a key only when it differs from its predecessor's;
//   2. a direct-mapped cache of 2048 keys in shared memory (one 64-bit atomic exchange): whoever finds its own key there drops
//      it - the thread that put it there forwards it.  Keys carry the frame-set number, so the cache is never cleared;
//   3. an open-addressing hash set in global memory (one load + one 64-bit compare-and-swap per surviving key), whose
//      entries carry an 8-bit launch epoch so it is never cleared between launches either.  The lane whose CAS installs a key
//      emits it.  (Slots are fully hashed: keeping the voxels of an 8 x 8 x 8 block together - a few cache lines per
//      surface patch - was measured 2.2 x slower, the lanes of a warp then hammer the same L2 sectors.)
// Warps work on their own: a warp takes a 32 x 32 pixel tile, collects the survivors of stage 2 in its slice of shared memory,
// inserts them 32 at a time (all lanes in flight together: two dependent L2 round trips per tile, not per key), and appends
// the new ones to the list with one atomic per ~100 records, in coalesced runs.  No block-wide barrier after start-up.
// Traffic: 2 B/px of depth in, 8 B per occupied voxel out.
#include "ti_common.cuh"

namespace ti {

constexpr int VX_THREADS = 256;
constexpr int VX_WARPS = VX_THREADS / 32;
constexpr int VX_TW = 32, VX_TH = 32;  // pixels per warp tile: 4 lanes x 8 px wide; 8 lane rows x 4 sub-blocks high
constexpr int VX_CACHE = 2048;         // direct-mapped key cache per CTA (16 KB)
static_assert(VX_CACHE == 2048, "the cache index is the top 11 bits of a 32-bit hash");
constexpr int VX_BUF = 384;            // records per warp in shared memory: confirmed new ones, then pending survivors
constexpr int VX_FLUSH = 96;           // confirmed records that trigger an append to the global list
constexpr int MAX_VX_JOBS = 16;

struct VxCam {
    double au[3], av[3], ac[3], t[3];  // already divided by the voxel size; t also carries the +16384 bias of the key fields
    int width, height;
};

struct VxJobDev {
    const uint16_t* depth;
    uint64_t depth_stride;  // bytes between frames
    VxCam cam;
    uint32_t tile_begin, tiles_x;
    uint32_t vec;  // rows are 16-byte aligned: 128-bit loads
};

struct VxParams {
    VxJobDev job[MAX_VX_JOBS];
    uint64_t* table;       // hash set, table_mask + 1 slots
    uint64_t table_mask;
    uint32_t region_mask, region_shift;  // frame set s starts probing in region (s - set_base) & region_mask of 1 << region_shift slots
    uint64_t epoch;        // << 56
    uint64_t tag;          // << 56
    uint64_t* records;
    uint64_t capacity;
    uint32_t* n_records;
    uint32_t* set_counts;
    uint32_t tiles_per_set, set_base, max_depth;
    int n_jobs, n_batch;
    int debug;  // bring-up switches (TI_OPT_DEBUG): 1 = no hash-set stage, 2 = no cache stage, 4 = block-local slots (measured slower: same-sector contention)
};

__device__ __forceinline__ uint64_t vx_mix(uint64_t k) {  // murmur3 finaliser
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

// floor(x) as an int for |x| < 2^31: add 1.5 * 2^52 rounding towards minus infinity, take the low mantissa word
// (a plain FP64-pipe add; the double -> int conversion instruction runs at a quarter of that rate)
__device__ __forceinline__ int vx_floor(double x) {
#ifdef TI_EMULATE
    return (int)std::floor(x);
#else
    return __double2loint(__dadd_rd(x, 6755399441055744.0));
#endif
}

__device__ __forceinline__ double vx_u2d(uint32_t v) {  // exact u32 -> double without the conversion pipe
#ifdef TI_EMULATE
    return (double)v;
#else
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
#endif
}

#ifdef TI_EMULATE
__device__ __forceinline__ uint64_t vx_exch(uint64_t* p, uint64_t v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
__device__ __forceinline__ uint64_t vx_cas(uint64_t* p, uint64_t expect, uint64_t v) {
    __atomic_compare_exchange_n(p, &expect, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expect;
}
__device__ __forceinline__ uint64_t vx_load(const uint64_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ uint64_t vx_exch(uint64_t* p, uint64_t v) {
    return atomicExch(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
__device__ __forceinline__ uint64_t vx_cas(uint64_t* p, uint64_t expect, uint64_t v) {
    return atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)expect, (unsigned long long)v);
}
__device__ __forceinline__ uint64_t vx_load(const uint64_t* p) { return *reinterpret_cast<const volatile uint64_t*>(p); }
#endif

// true when this thread installed `key` (set number and voxel, 56 bits) in the global hash set
__device__ __forceinline__ bool vx_insert(const VxParams& P, uint64_t key) {
    const uint64_t entry = P.epoch | key;
    // Home slot: the frame set picks a REGION of the table, the hash a slot inside it.  Tiles are handed out in frame-set order,
    // so at any moment the whole grid works on one or two frame sets and their regions (a few MB) stay in the L2; with slots
    // hashed over the whole table every probe was a DRAM sector (ncu: 390 MB read for 131 MB of depth).  Probing runs on past
    // the end of a region, so a frame set with more voxels than its share just borrows from its neighbour.
    const uint64_t set = (key >> 45) & 0x7FFu;
    uint64_t slot;
    if (P.debug & 4) {
        const uint64_t h = vx_mix(key & ~0x00000001C0038007ull);  // bring-up: 8 x 8 x 8 blocks kept together (measured slower)
        slot = (((h >> 20) << 9) | ((key >> 24) & 0x1C0u) | ((key >> 12) & 0x38u) | (key & 7u)) & P.table_mask;
    } else if (P.debug & 8) {
        slot = (vx_mix(key) >> 11) & P.table_mask;  // bring-up: hashed over the whole table
    } else {
        slot = (((set - P.set_base) & P.region_mask) << P.region_shift) | ((vx_mix(key) >> 11) & ((1ull << P.region_shift) - 1));
    }
    uint64_t cur = vx_load(P.table + slot);
    for (uint32_t probes = 0; probes < 2048u; ++probes) {
        if (cur == entry) return false;
        if ((cur >> 56) != (P.epoch >> 56)) {  // left over from an earlier launch: free
            const uint64_t old = vx_cas(P.table + slot, cur, entry);
            if (old == cur) return true;
            cur = old;  // somebody else took the slot meanwhile: look at what they put there
            continue;
        }
        slot = (slot + 1) & P.table_mask;
        cur = vx_load(P.table + slot);
    }
    return false;  // 2048 probes: the table (2 x capacity slots) is as good as full - the list overflowed long ago, *n_records says so
}

__global__ void __launch_bounds__(VX_THREADS, 3) voxel_cloud_kernel(const __grid_constant__ VxParams P) {
    __shared__ uint64_t cache[VX_CACHE];
    __shared__ uint64_t wbuf[VX_WARPS][VX_BUF];
    for (int i = threadIdx.x; i < VX_CACHE; i += VX_THREADS) cache[i] = ~0ull;  // no key has its top byte set
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int lx = lane & 3, ly = lane >> 2;
    uint64_t* buf = wbuf[warp];
    uint32_t n_conf = 0, n_pend = 0, buf_b = 0;  // warp-uniform

    // pending survivors -> hash set, 32 at a time; the ones this warp installed move up behind the confirmed records
    auto drain = [&]() {
        __syncwarp();
        const uint32_t base = n_conf, n = n_pend;
        uint32_t out = n_conf;
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t idx = i0 + lane;
            const bool have = idx < n;
            const uint64_t key = have ? buf[base + idx] : 0ull;
            const bool is_new = have && ((P.debug & 1) || vx_insert(P, key));
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_new);
            __syncwarp();  // every read of this round precedes its writes (which never reach a later round's reads)
            if (is_new) buf[out + __popc(m & lt_mask)] = key;
            out += __popc(m);
        }
        n_conf = out;
        n_pend = 0;
        __syncwarp();
    };
    auto flush = [&](uint32_t b) {
        if (n_conf == 0) return;
        uint32_t base = 0;
        if (lane == 0) {
            base = atomicAdd(P.n_records, n_conf);
            if (P.set_counts) atomicAdd(P.set_counts + b, n_conf);
        }
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (uint32_t i = lane; i < n_conf; i += 32)
            if ((uint64_t)base + i < P.capacity) P.records[(uint64_t)base + i] = P.tag | buf[i];
        n_conf = 0;
        __syncwarp();
    };

    const uint64_t total = (uint64_t)P.tiles_per_set * P.n_batch;
    const uint64_t n_warps = (uint64_t)gridDim.x * VX_WARPS;
    for (uint64_t t = (uint64_t)blockIdx.x * VX_WARPS + warp; t < total; t += n_warps) {
        const uint32_t b = (uint32_t)(t / P.tiles_per_set);
        const uint32_t r = (uint32_t)(t - (uint64_t)b * P.tiles_per_set);
        int j = 0;
        while (j + 1 < P.n_jobs && r >= P.job[j + 1].tile_begin) ++j;
        const VxJobDev& J = P.job[j];
        const uint32_t lt = r - J.tile_begin;
        const int tile_y = (int)(lt / J.tiles_x), tile_x = (int)(lt - (uint32_t)tile_y * J.tiles_x);
        // a list far beyond its capacity (the caller sized it wrong; *n_records will say so) stops being extended: the hash set
        // holds 2 x capacity slots and must not be driven to full
        if (*reinterpret_cast<const volatile uint32_t*>(P.n_records) > 2 * P.capacity && P.capacity) continue;
        if (b != buf_b) {  // the confirmed records are counted per frame set
            flush(buf_b);
            buf_b = b;
        }
        const int u0 = tile_x * VX_TW + lx * 8;
        const uint16_t* frame = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(J.depth) + (uint64_t)b * J.depth_stride);
        auto load_sub = [&](int sub) -> uint4 {
            const int v = tile_y * VX_TH + sub * 8 + ly;
            if (sub >= 4 || v >= J.cam.height || u0 >= J.cam.width) return make_uint4(0u, 0u, 0u, 0u);
            const uint16_t* row = frame + (size_t)v * J.cam.width;
            if (J.vec) return ld_stream_u4(row + u0);  // width % 8 == 0: the 8 pixels are inside the row
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int k = 0; k < 8; ++k)
                if (u0 + k < J.cam.width) w[k >> 1] |= (uint32_t)row[u0 + k] << ((k & 1) * 16);
            return make_uint4(w[0], w[1], w[2], w[3]);
        };
        const uint32_t set_hi = (P.set_base + b) << 13;  // bits 45.. of the record
        const double ud0 = vx_u2d((uint32_t)u0);
        uint4 cur = load_sub(0);
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {
            const uint4 nxt = load_sub(sub + 1);  // next sub-block's depth is in flight while this one is processed
            if (n_conf + n_pend > VX_BUF - 256) {  // room for a sub-block in which every pixel is a new voxel
                drain();
                if (n_conf > VX_BUF - 256) flush(b);
            }
            const uint32_t dw[4] = {cur.x, cur.y, cur.z, cur.w};
            const double vd = vx_u2d((uint32_t)(tile_y * VX_TH + sub * 8 + ly));
            const double bx = fma(J.cam.av[0], vd, J.cam.ac[0]);
            const double by = fma(J.cam.av[1], vd, J.cam.ac[1]);
            const double bz = fma(J.cam.av[2], vd, J.cam.ac[2]);
            // stage 1: the warp's run-distinct keys of this sub-block, appended densely behind the pending ones (one ballot per
            // pixel column: stage 2 then walks ceil(n / 32) full rows instead of max-over-lanes sparse ones)
            uint64_t* pend = buf + n_conf + n_pend;
            uint32_t prev_lo = 0xFFFFFFFFu, prev_hi = 0xFFFFFFFFu, n_sub = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t d = (k & 1) ? (dw[k >> 1] >> 16) : (dw[k >> 1] & 0xFFFFu);
                const double ud = ud0 + (double)k, dd = vx_u2d(d);
                // t carries the +16384 bias of the key fields: the low word of the rounded-down sum IS the field
                const uint32_t kx = (uint32_t)vx_floor(fma(dd, fma(J.cam.au[0], ud, bx), J.cam.t[0]));
                const uint32_t ky = (uint32_t)vx_floor(fma(dd, fma(J.cam.au[1], ud, by), J.cam.t[1]));
                const uint32_t kz = (uint32_t)vx_floor(fma(dd, fma(J.cam.au[2], ud, bz), J.cam.t[2]));
                const uint32_t lo = kz | (ky << 15) | (kx << 30), hi = (kx >> 2) | set_hi;
                const bool keep = (d - 1u) < P.max_depth && (lo != prev_lo || hi != prev_hi);  // 0 < d <= max_depth, new run
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    pend[n_sub + __popc(m & lt_mask)] = ((uint64_t)hi << 32) | lo;
                    prev_lo = lo; prev_hi = hi;
                }
                n_sub += __popc(m);
            }
            __syncwarp();
            // stage 2, 32 keys at a time: the shared cache; survivors are packed to the front of the same region
            uint32_t out = 0;
            for (uint32_t i0 = 0; i0 < n_sub; i0 += 32) {
                bool fresh = i0 + lane < n_sub;
                uint64_t key = 0;
                if (fresh) {
                    key = pend[i0 + lane];
                    // index from the TOP bits of a product: every key bit reaches them (low product bits only see low key bits)
                    const uint32_t c = ((uint32_t)key ^ ((uint32_t)(key >> 32) * 0x9E3779B1u)) * 0x85EBCA6Bu;
                    fresh = (P.debug & 2) || vx_exch(&cache[c >> 21], key) != key;
                }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, fresh);  // also orders this row's reads before the writes below
                if (fresh) pend[out + __popc(m & lt_mask)] = key;      // out <= i0: never reaches an unread row
                out += __popc(m);
            }
            n_pend += out;
            __syncwarp();
            cur = nxt;
        }
        drain();
        if (n_conf >= VX_FLUSH) flush(b);
    }
    flush(buf_b);
}

// records -> centres of the voxels as N x 3 f32 (the cloud type of SlamMap.to_point_cloud, interface.py:134-138);
// rows past *n_records are left untouched
__global__ void __launch_bounds__(256) voxel_points_kernel(const uint64_t* records, const uint32_t* n_records, uint64_t max_records,
                                                           double voxel, float* xyz) {
    const uint64_t n = min((uint64_t)*n_records, max_records);
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const uint64_t r = records[i];
        const int kx = (int)((r >> 30) & 0x7FFF) - 16384, ky = (int)((r >> 15) & 0x7FFF) - 16384, kz = (int)(r & 0x7FFF) - 16384;
        xyz[3 * i + 0] = (float)(((double)kx + 0.5) * voxel);
        xyz[3 * i + 1] = (float)(((double)ky + 0.5) * voxel);
        xyz[3 * i + 2] = (float)(((double)kz + 0.5) * voxel);
    }
}

}  // namespace ti

using namespace ti;

extern "C" {

int ti_set_voxel_grid(ti_ctx* ctx, double voxel_size_m, uint32_t max_depth_mm) {
    if (!ctx) return TI_EINVAL;
    if (!(voxel_size_m > 0.0) || !(voxel_size_m < 1e6)) return fail(ctx, TI_EINVAL, "ti_set_voxel_grid: voxel size must be positive");
    ctx->voxel_size = voxel_size_m;
    ctx->voxel_max_depth = max_depth_mm ? std::min<uint32_t>(max_depth_mm, 65535u) : 65535u;
    return TI_OK;
}

int ti_voxel_cloud(ti_ctx* ctx, const ti_depth_stream* streams, int n_streams, int n_batch, uint32_t set_base, uint32_t tag,
                   uint64_t* records, uint64_t capacity, uint32_t* n_records, uint32_t* set_counts) {
    if (!ctx) return TI_EINVAL;
    if (n_streams < 0 || n_batch < 0 || (n_streams > 0 && !streams)) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: bad stream array");
    if (!n_records || (capacity && !records)) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: null records / n_records");
    if (!(ctx->voxel_size > 0.0)) return fail(ctx, TI_ESTATE, "ti_voxel_cloud: call ti_set_voxel_grid first");
    if (n_streams > MAX_VX_JOBS) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: at most %d depth streams per call", MAX_VX_JOBS);
    if ((uint64_t)set_base + (uint64_t)n_batch > 2048u) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: set_base + n_batch must be <= 2048 (11-bit set field)");
    if (tag > 255u) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: tag must fit 8 bits");
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    TI_CUDA(ctx, cudaMemsetAsync(n_records, 0, sizeof(uint32_t), ctx->stream));
    if (set_counts && n_batch) TI_CUDA(ctx, cudaMemsetAsync(set_counts, 0, sizeof(uint32_t) * (size_t)n_batch, ctx->stream));
    if (n_streams == 0 || n_batch == 0) return TI_OK;

    VxParams P{};
    const double inv = 1.0 / ctx->voxel_size;
    uint32_t tiles = 0;
    uint64_t px = 0;
    for (int i = 0; i < n_streams; ++i) {
        const ti_depth_stream& S = streams[i];
        if (S.camera < 0 || S.camera >= TI_MAX_CAMERAS || !ctx->cams[S.camera].has_proj)
            return fail(ctx, TI_ESTATE, "ti_voxel_cloud: camera slot %d has no projection (call ti_upload_projection)", S.camera);
        if (!S.depth) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: stream %d has a null depth pointer", i);
        const CameraSlot& C = ctx->cams[S.camera];
        VxJobDev& D = P.job[i];
        D.depth = S.depth;
        D.depth_stride = S.depth_frame_stride ? S.depth_frame_stride : (uint64_t)C.proj_w * C.proj_h * 2;
        if (D.depth_stride % 2 || (uintptr_t)S.depth % 2) return fail(ctx, TI_EINVAL, "ti_voxel_cloud: depth must be 2-byte aligned");
        double reach = 0.0;  // max |p| / voxel over the image and every valid depth: the ray is linear in (u, v), so a corner has it
        for (int r = 0; r < 3; ++r) {
            D.cam.au[r] = C.proj_au[r] * inv; D.cam.av[r] = C.proj_av[r] * inv; D.cam.ac[r] = C.proj_ac[r] * inv; D.cam.t[r] = C.proj_t[r] * inv + 16384.0;
            double ray = 0.0;
            for (int corner = 0; corner < 4; ++corner)
                ray = std::max(ray, fabs(D.cam.au[r] * ((corner & 1) ? C.proj_w - 1 : 0) + D.cam.av[r] * ((corner & 2) ? C.proj_h - 1 : 0) + D.cam.ac[r]));
            reach = std::max(reach, ray * ctx->voxel_max_depth + fabs(C.proj_t[r] * inv));
        }
        if (!(reach < 16383.0))
            return fail(ctx, TI_EINVAL, "ti_voxel_cloud: camera slot %d reaches %.0f voxels from the body origin; the 15-bit key fields hold "
                        "16383 (use a larger voxel or a depth cap)", S.camera, reach);
        D.cam.width = C.proj_w; D.cam.height = C.proj_h;
        D.tiles_x = (uint32_t)((C.proj_w + VX_TW - 1) / VX_TW);
        D.tile_begin = tiles;
        tiles += D.tiles_x * (uint32_t)((C.proj_h + VX_TH - 1) / VX_TH);
        D.vec = (C.proj_w % 8 == 0) && ((uintptr_t)S.depth % 16 == 0) && (D.depth_stride % 16 == 0);
        px += (uint64_t)C.proj_w * C.proj_h;
    }
    // hash set: twice the records the caller's list can take (a list that overflows is reported through *n_records; the
    // table then fills up harmlessly), at most twice the pixels of the launch (every pixel its own voxel), a power of two;
    // entries of earlier launches are recognised by their epoch byte, so it is cleared only when the epoch wraps or the
    // table grows
    uint64_t slots = 1u << 16;
    const uint64_t want = 2 * std::min<uint64_t>(std::max<uint64_t>(capacity, 1), px * (uint64_t)n_batch);
    while (slots < want) slots <<= 1;
    if (slots > ctx->voxel_slots) {
        TI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->voxel_table) cudaFree(ctx->voxel_table);
        ctx->voxel_table = nullptr; ctx->voxel_slots = 0;
        TI_CUDA(ctx, cudaMalloc(&ctx->voxel_table, slots * sizeof(uint64_t)));
        ctx->voxel_slots = slots;
        ctx->voxel_epoch = 0;
    }
    if (ctx->voxel_epoch == 0 || ctx->voxel_epoch >= 255) {
        TI_CUDA(ctx, cudaMemsetAsync(ctx->voxel_table, 0, ctx->voxel_slots * sizeof(uint64_t), ctx->stream));
        ctx->voxel_epoch = 0;
    }
    ctx->voxel_epoch++;
    P.table = ctx->voxel_table;
    P.table_mask = ctx->voxel_slots - 1;
    {
        uint32_t regions = 1;
        while ((int)regions < n_batch && (ctx->voxel_slots / (regions * 2)) >= 4096) regions <<= 1;
        uint32_t shift = 0;
        while ((ctx->voxel_slots >> (shift + 1)) >= regions) ++shift;  // 1 << shift = slots / regions
        P.region_mask = regions - 1;
        P.region_shift = shift;
    }
    P.epoch = (uint64_t)ctx->voxel_epoch << 56;
    P.tag = (uint64_t)tag << 56;
    P.records = records; P.capacity = capacity; P.n_records = n_records; P.set_counts = set_counts;
    P.tiles_per_set = tiles; P.set_base = set_base; P.max_depth = ctx->voxel_max_depth;
    P.n_jobs = n_streams; P.n_batch = n_batch;
    P.debug = ctx->debug;
    const uint64_t total = (uint64_t)tiles * n_batch;
    const int per_sm = ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : resident_ctas(voxel_cloud_kernel, VX_THREADS, 0, 4);
    const int grid = (int)std::min<uint64_t>((total + VX_WARPS - 1) / VX_WARPS, (uint64_t)ctx->sm_count * per_sm);
    TI_LAUNCH(voxel_cloud_kernel, grid, VX_THREADS, 0, ctx->stream, P);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

int ti_voxel_points(ti_ctx* ctx, const uint64_t* records, const uint32_t* n_records, uint64_t max_records, float* xyz) {
    if (!ctx) return TI_EINVAL;
    if (!n_records || (max_records && (!records || !xyz))) return fail(ctx, TI_EINVAL, "ti_voxel_points: null argument");
    if (!(ctx->voxel_size > 0.0)) return fail(ctx, TI_ESTATE, "ti_voxel_points: call ti_set_voxel_grid first");
    if (max_records == 0) return TI_OK;
    TI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int grid = (int)std::min<uint64_t>((max_records + 255) / 256, (uint64_t)ctx->sm_count * 8);
    TI_LAUNCH(voxel_points_kernel, grid, 256, 0, ctx->stream, records, n_records, max_records, ctx->voxel_size, xyz);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

}  // extern "C"
