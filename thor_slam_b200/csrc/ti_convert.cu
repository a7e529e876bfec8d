// Format conversion kernels (u8, bit-exact with OpenCV 4.x cvtColor).
//
//   BGR8 -> RGB8   byte swap                                (isaac_ros.py:357, run_pipeline.py:234)
//   BGR8 -> MONO8  Y = (B*3735 + G*19235 + R*9798 + 2^14) >> 15         (cv2.COLOR_BGR2GRAY)
//   NV12 -> MONO8  copy of the luma plane                               (cv2.COLOR_YUV2GRAY_NV12)
//   NV12 -> RGB8 / BGR8  BT.601 limited range, 20-bit fixed point       (cv2.COLOR_YUV2RGB_NV12)
//   MONO8 -> MONO8 copy                                                 (isaac_ros.py:352-353)
//
// All of them are pure streaming: each thread moves whole 16-byte vectors, a warp covers a
// contiguous span, nothing is re-read -> HBM-bound.  One launch covers every (stream, frame)
// of the batch; the grid is a multiple of the SM count and CTAs stride over "units" of work.
#include "ti_common.cuh"
#include "ti_pixel.cuh"

namespace ti {

constexpr int CV_THREADS = 256;
constexpr int MAX_CONVERT_JOBS = TI_MAX_STREAMS;

struct ConvertParams {
    ConvertJob job[MAX_CONVERT_JOBS];
    uint32_t unit_begin[MAX_CONVERT_JOBS + 1];  // prefix sum of units per frame over jobs
    int n_jobs;
    int n_batch;
};

__device__ __forceinline__ uint32_t byte_of(const uint32_t* w, int i) {  // i-th byte of a small register array
    return (w[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
}

// 48 bytes per lane (16 three-byte pixels) written so that every store instruction of the warp covers 512
// contiguous bytes: strided 16-byte stores (lane stride 48 B) touch 12 lines and half-fill every sector.
// Fast path when the 32 lanes' destinations are back to back (checked with a ballot); otherwise plain stores.
__device__ __forceinline__ void store48(const uint32_t o[12], uint8_t* dst, bool live, uint4* wbuf) {
    const int lane = threadIdx.x & 31;
    uint8_t* base = reinterpret_cast<uint8_t*>(
        ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)((uint64_t)(uintptr_t)dst >> 32), 0) << 32) |
        __shfl_sync(0xFFFFFFFFu, (uint32_t)(uintptr_t)dst, 0));
    const bool contiguous = __ballot_sync(0xFFFFFFFFu, live && dst == base + lane * 48) == 0xFFFFFFFFu;
    if (contiguous) {
        wbuf[lane * 3] = make_uint4(o[0], o[1], o[2], o[3]);
        wbuf[lane * 3 + 1] = make_uint4(o[4], o[5], o[6], o[7]);
        wbuf[lane * 3 + 2] = make_uint4(o[8], o[9], o[10], o[11]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 3; ++k) st_stream_u4(base + (k * 32 + lane) * 16, wbuf[k * 32 + lane]);
        __syncwarp();
    } else if (live) {
        st_stream_u4(dst, make_uint4(o[0], o[1], o[2], o[3]));
        st_stream_u4(dst + 16, make_uint4(o[4], o[5], o[6], o[7]));
        st_stream_u4(dst + 32, make_uint4(o[8], o[9], o[10], o[11]));
    }
}

// ---- unit processors: one "unit" = 16 pixels (or 2 rows x 16 pixels for NV12 colour) --------
// Each unit type is split into load() and emit() so that a thread can have the loads of several units
// in flight before it converts and stores the first one (memory-level parallelism: one 16-byte load per
// thread in flight cannot saturate HBM).

struct UnitCopy {  // MONO8 -> MONO8, NV12 luma -> MONO8: 16 bytes
    uint4 v;
    __device__ __forceinline__ void load(const uint8_t* src) { v = ld_stream_u4(src); }
    __device__ __forceinline__ void emit(uint8_t* dst) const { st_stream_u4(dst, v); }
};

struct UnitBgr {  // 16 BGR pixels = 48 bytes
    uint32_t w[12];
    __device__ __forceinline__ void load(const uint8_t* src) {
        const uint4 a = ld_stream_u4(src), b = ld_stream_u4(src + 16), c = ld_stream_u4(src + 32);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    }
    __device__ __forceinline__ void emit_rgb(uint8_t* dst, bool live, uint4* wbuf) const {
        uint32_t o[12];
        // every 12-byte group holds 4 pixels: B0G0R0B1 G1R1B2G2 R2B3G3R3 -> R0G0B0R1 G1B1R2G2 B2R3G3B3
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
            o[3 * g] = __byte_perm(w0, w1, 0x5012);                               // R0 G0 B0 R1
            o[3 * g + 1] = __byte_perm(w1, __byte_perm(w0, w2, 0x0043), 0x3540);  // G1 B1 R2 G2
            o[3 * g + 2] = __byte_perm(w2, w1, 0x1236);                           // B2 R3 G3 B3
        }
        store48(o, dst, live, wbuf);
    }
    __device__ __forceinline__ void emit_gray(uint8_t* dst) const {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t out = 0;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int px = q * 4 + p;
                out |= gray_of(byte_of(w, 3 * px), byte_of(w, 3 * px + 1), byte_of(w, 3 * px + 2)) << (8 * p);
            }
            o[q] = out;
        }
        st_stream_u4(dst, make_uint4(o[0], o[1], o[2], o[3]));
    }
};

struct UnitNv12 {  // 2 rows x 16 luma + 16 bytes of interleaved U,V
    uint32_t y0[4], y1[4], uv[4];
    __device__ __forceinline__ void load(const uint8_t* y0p, const uint8_t* y1p, const uint8_t* uvp) {
        const uint4 a = ld_stream_u4(y0p), b = ld_stream_u4(y1p), c = ld_stream_u4(uvp);
        y0[0] = a.x; y0[1] = a.y; y0[2] = a.z; y0[3] = a.w;
        y1[0] = b.x; y1[1] = b.y; y1[2] = b.z; y1[3] = b.w;
        uv[0] = c.x; uv[1] = c.y; uv[2] = c.z; uv[3] = c.w;
    }
    template <bool BGR_OUT>
    __device__ __forceinline__ void emit(uint8_t* d0, uint8_t* d1, bool live, uint4* wbuf) const {
        // chroma terms once per (U,V) pair (shared by 2 columns x 2 rows); per pixel 3 x (add, shift) and
        // saturate+pack two channels per instruction (cvt.pack.sat)
        ChromaTerms ct[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ct[i] = chroma_terms((int)byte_of(uv, 2 * i), (int)byte_of(uv, 2 * i + 1));
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            const uint32_t* yy = row ? y1 : y0;
            uint32_t o[12];
#pragma unroll
            for (int g = 0; g < 4; ++g) {  // 4 pixels -> 12 bytes -> 3 words
                int v[12];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int px = 4 * g + p;
                    const int l = luma_term((int)byte_of(yy, px));
                    const ChromaTerms& c = ct[px >> 1];
                    const int r = (l + c.r) >> 20, gg = (l + c.g) >> 20, b = (l + c.b) >> 20;
                    v[3 * p] = BGR_OUT ? b : r;
                    v[3 * p + 1] = gg;
                    v[3 * p + 2] = BGR_OUT ? r : b;
                }
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    o[3 * g + q] = pack_sat_u8(v[4 * q + 1], v[4 * q], pack_sat_u8(v[4 * q + 3], v[4 * q + 2], 0u));
            }
            store48(o, row ? d1 : d0, live, wbuf);
        }
    }
};

enum ConvMode { CM_COPY = 0, CM_BGR_RGB = 1, CM_BGR_GRAY = 2, CM_NV12_RGB = 3, CM_NV12_BGR = 4 };

// where unit t of the launch lives
struct UnitAddr {
    const uint8_t* src;
    uint8_t* dst;
    int w, h;
    uint32_t unit;
};

__device__ __forceinline__ UnitAddr locate(const ConvertParams& P, uint32_t units_per_set, uint64_t t) {
    const uint32_t b = (uint32_t)(t / units_per_set);
    const uint32_t r = (uint32_t)(t - (uint64_t)b * units_per_set);
    int j = 0;
    while (j + 1 < P.n_jobs && r >= P.unit_begin[j + 1]) ++j;
    const ConvertJob& J = P.job[j];
    return UnitAddr{J.src + (uint64_t)b * J.src_stride, J.dst + (uint64_t)b * J.dst_stride, J.width, J.height, r - P.unit_begin[j]};
}

// ---- vector kernels (width % 16 == 0, 16-byte aligned frames): one per conversion, UNROLL units in flight ----
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(CV_THREADS) convert_vec_kernel(const __grid_constant__ ConvertParams P) {
    __shared__ uint4 wbuf_all[(MODE == CM_COPY || MODE == CM_BGR_GRAY) ? 1 : CV_THREADS / 32][96];
    uint4* wbuf = wbuf_all[(MODE == CM_COPY || MODE == CM_BGR_GRAY) ? 0 : threadIdx.x >> 5];
    const uint32_t units_per_set = P.unit_begin[P.n_jobs];
    const uint64_t total = (uint64_t)units_per_set * P.n_batch;
    const uint64_t stride = (uint64_t)gridDim.x * CV_THREADS;
    // warp-uniform trip count: the coalesced store path uses warp collectives
    const uint64_t warp_t0 = (uint64_t)blockIdx.x * CV_THREADS + (threadIdx.x & ~31);
    for (uint64_t t0 = warp_t0 + (threadIdx.x & 31); t0 - (threadIdx.x & 31) < total; t0 += stride * UNROLL) {
        UnitAddr A[UNROLL];
        bool live[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            live[k] = t0 + k * stride < total;
            if (live[k]) A[k] = locate(P, units_per_set, t0 + k * stride);
        }
        if (MODE == CM_COPY) {
            UnitCopy U[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) if (live[k]) U[k].load(A[k].src + (uint64_t)A[k].unit * 16);
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) if (live[k]) U[k].emit(A[k].dst + (uint64_t)A[k].unit * 16);
        } else if (MODE == CM_BGR_RGB || MODE == CM_BGR_GRAY) {
            UnitBgr U[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) if (live[k]) U[k].load(A[k].src + (uint64_t)A[k].unit * 48);
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) {
                if (MODE == CM_BGR_RGB) U[k].emit_rgb(live[k] ? A[k].dst + (uint64_t)A[k].unit * 48 : nullptr, live[k], wbuf);
                else if (live[k]) U[k].emit_gray(A[k].dst + (uint64_t)A[k].unit * 16);
            }
        } else {
            UnitNv12 U[UNROLL];
            uint8_t* d0[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) d0[k] = nullptr;
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) if (live[k]) {
                const int w = A[k].w, h = A[k].h;
                const uint32_t upr = (uint32_t)w / 16;  // unit = 16 columns of a row PAIR
                const uint32_t rp = A[k].unit / upr, cx = (A[k].unit - rp * upr) * 16;
                const uint8_t* y0 = A[k].src + (uint64_t)(2 * rp) * w + cx;
                U[k].load(y0, y0 + w, A[k].src + (uint64_t)h * w + (uint64_t)rp * w + cx);
                d0[k] = A[k].dst + ((uint64_t)(2 * rp) * w + cx) * 3;
            }
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) {
                uint8_t* d1 = live[k] ? d0[k] + (uint64_t)A[k].w * 3 : nullptr;
                if (MODE == CM_NV12_BGR) U[k].template emit<true>(d0[k], d1, live[k], wbuf);
                else U[k].template emit<false>(d0[k], d1, live[k], wbuf);
            }
        }
    }
}

// ---- scalar kernel (any width / alignment): one pixel per thread ----------------------------
__global__ void __launch_bounds__(CV_THREADS) convert_scalar_kernel(ConvertJob J, int n_batch) {
    const uint64_t npx = (uint64_t)J.width * J.height;
    const uint64_t total = npx * n_batch;
    for (uint64_t t = (uint64_t)blockIdx.x * CV_THREADS + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * CV_THREADS) {
        const uint64_t b = t / npx, p = t - b * npx;
        const uint8_t* src = J.src + b * J.src_stride;
        uint8_t* dst = J.dst + b * J.dst_stride;
        if (J.src_fmt == TI_FMT_BGR8) {
            const uint32_t bb = src[3 * p], gg = src[3 * p + 1], rr = src[3 * p + 2];
            if (J.dst_fmt == TI_FMT_RGB8) {
                dst[3 * p] = (uint8_t)rr; dst[3 * p + 1] = (uint8_t)gg; dst[3 * p + 2] = (uint8_t)bb;
            } else {
                dst[p] = (uint8_t)gray_of(bb, gg, rr);
            }
        } else if (J.dst_fmt == TI_FMT_MONO8) {
            dst[p] = src[p];
        } else {
            const uint32_t y = (uint32_t)(p / J.width), x = (uint32_t)(p - (uint64_t)y * J.width);
            const uint8_t* uv = src + npx + (uint64_t)(y >> 1) * J.width + (x & ~1u);
            int r, g, bl;
            yuv_to_rgb(src[p], uv[0], uv[1], r, g, bl);
            const bool bgr = J.dst_fmt == TI_FMT_BGR8;
            dst[3 * p] = (uint8_t)(bgr ? bl : r); dst[3 * p + 1] = (uint8_t)g; dst[3 * p + 2] = (uint8_t)(bgr ? r : bl);
        }
    }
}

static bool convert_supported(int s, int d) {
    if (s == TI_FMT_BGR8) return d == TI_FMT_RGB8 || d == TI_FMT_MONO8;
    if (s == TI_FMT_NV12) return d == TI_FMT_MONO8 || d == TI_FMT_RGB8 || d == TI_FMT_BGR8;
    if (s == TI_FMT_MONO8) return d == TI_FMT_MONO8;
    return false;
}

static int conv_mode(int s, int d) {
    if (d == TI_FMT_MONO8 && (s == TI_FMT_MONO8 || s == TI_FMT_NV12)) return CM_COPY;
    if (s == TI_FMT_BGR8) return d == TI_FMT_RGB8 ? CM_BGR_RGB : CM_BGR_GRAY;
    return d == TI_FMT_BGR8 ? CM_NV12_BGR : CM_NV12_RGB;
}

template <int MODE, int UNROLL>
static int launch_mode(ti_ctx* ctx, const ConvertParams& P, uint32_t units) {
    const uint64_t total = (uint64_t)units * P.n_batch;
    const uint64_t want = (total + (uint64_t)CV_THREADS * UNROLL - 1) / ((uint64_t)CV_THREADS * UNROLL);
    const int per_sm = resident_ctas(convert_vec_kernel<MODE, UNROLL>, CV_THREADS, 0, 4);  // memoised per thread and kernel
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)ctx->sm_count * per_sm));
    TI_LAUNCH((convert_vec_kernel<MODE, UNROLL>), grid, CV_THREADS, 0, ctx->stream, P);
    TI_CHECK_LAUNCH(ctx);
    return TI_OK;
}

int launch_convert(ti_ctx* ctx, const ConvertJob* jobs, int n_jobs, int n_batch) {
    if (n_jobs <= 0 || n_batch <= 0) return TI_OK;
    if (n_jobs > MAX_CONVERT_JOBS) return fail(ctx, TI_EINVAL, "too many convert streams (%d > %d)", n_jobs, MAX_CONVERT_JOBS);
    ConvertParams P[5] = {};
    uint32_t units[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n_jobs; ++i) {
        const ConvertJob& J = jobs[i];
        if (!convert_supported(J.src_fmt, J.dst_fmt))
            return fail(ctx, TI_EINVAL, "unsupported conversion %d -> %d", J.src_fmt, J.dst_fmt);
        if (J.width <= 0 || J.height <= 0 || !J.src || !J.dst)
            return fail(ctx, TI_EINVAL, "convert: bad size or null pointer");
        if (J.src_fmt == TI_FMT_NV12 && ((J.width | J.height) & 1))
            return fail(ctx, TI_EINVAL, "NV12 needs even width and height (got %dx%d)", J.width, J.height);
        const bool aligned = (J.width % 16 == 0) && (((uintptr_t)J.src | (uintptr_t)J.dst | J.src_stride | J.dst_stride) % 16 == 0);
        if (!aligned) {
            const uint64_t total = (uint64_t)J.width * J.height * n_batch;
            const int grid = (int)std::min<uint64_t>((total + CV_THREADS - 1) / CV_THREADS, (uint64_t)ctx->sm_count * 8);
            TI_LAUNCH(convert_scalar_kernel, grid, CV_THREADS, 0, ctx->stream, J, n_batch);
            TI_CHECK_LAUNCH(ctx);
            continue;
        }
        const int m = conv_mode(J.src_fmt, J.dst_fmt);
        ConvertParams& Q = P[m];
        Q.job[Q.n_jobs] = J;
        Q.unit_begin[Q.n_jobs] = units[m];
        const bool pair = m == CM_NV12_RGB || m == CM_NV12_BGR;
        units[m] += (uint32_t)((uint64_t)J.width * J.height / (pair ? 32 : 16));
        ++Q.n_jobs;
    }
    for (int m = 0; m < 5; ++m) {
        if (!P[m].n_jobs) continue;
        P[m].unit_begin[P[m].n_jobs] = units[m];
        P[m].n_batch = n_batch;
        int rc = TI_OK;
        switch (m) {
            case CM_COPY: rc = launch_mode<CM_COPY, 4>(ctx, P[m], units[m]); break;
            case CM_BGR_RGB: rc = launch_mode<CM_BGR_RGB, 2>(ctx, P[m], units[m]); break;
            case CM_BGR_GRAY: rc = launch_mode<CM_BGR_GRAY, 2>(ctx, P[m], units[m]); break;
            case CM_NV12_RGB: rc = launch_mode<CM_NV12_RGB, 1>(ctx, P[m], units[m]); break;
            default: rc = launch_mode<CM_NV12_BGR, 1>(ctx, P[m], units[m]); break;
        }
        if (rc != TI_OK) return rc;
    }
    return TI_OK;
}

}  // namespace ti
