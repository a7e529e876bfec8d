// Host side of the TMA wrappers: tensor-map encoding through the driver entry point
// (cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint, so libcuda is not linked).
#include "ti_tma.cuh"

#include <string.h>

#include <utility>

namespace ti {

#ifdef TI_EMULATE
int tma_encode_3d(ti_ctx*, TiTensorMap* out, const void* base, int elem_bytes, int w, int h, int n, uint64_t pitch_y,
                  uint64_t pitch_z, int box_x, int box_y) {
    *out = TiTensorMap{};
    out->elem = elem_bytes;
    out->base = static_cast<const uint8_t*>(base);
    out->dim[0] = w; out->dim[1] = h; out->dim[2] = n;
    out->stride[0] = 1; out->stride[1] = (int64_t)pitch_y; out->stride[2] = (int64_t)pitch_z;
    out->box[0] = box_x; out->box[1] = box_y; out->box[2] = 1;
    return TI_OK;
}
#else
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int tma_encode_3d(ti_ctx* ctx, TiTensorMap* out, const void* base, int elem_bytes, int w, int h, int n, uint64_t pitch_y,
                  uint64_t pitch_z, int box_x, int box_y) {
    // A live rig replays the same few (ring slot, batch) combinations for ever: descriptors are remembered per context
    // (an encode costs ~1.5 us of driver time, eight of them were a third of a one-frame-set call).
    static_assert(sizeof(TiTensorMap) == sizeof(TmaBlob), "a tensor map is 128 bytes");
    TmaKey key;
    memset(&key, 0, sizeof key);  // padding bytes take part in the comparison
    key.base = base; key.pitch_y = pitch_y; key.pitch_z = pitch_z; key.elem_bytes = elem_bytes; key.w = w; key.h = h; key.n = n; key.box_x = box_x; key.box_y = box_y;
    if (ctx) {
        auto& cache = ctx->tma_cache;
        for (auto it = cache.begin(); it != cache.end(); ++it)
            if (memcmp(&it->first, &key, sizeof key) == 0) {
                memcpy(out, &it->second, sizeof(TiTensorMap));
                if (it != cache.begin()) std::iter_swap(it, it - 1);  // drift towards the front: hot entries are found first
                return TI_OK;
            }
    }
    EncodeTiledFn fn = encoder();
    if (!fn) return fail(ctx, TI_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    if ((elem_bytes != 1 && elem_bytes != 4) || ((uintptr_t)base % 16) || (pitch_y % 16) || (pitch_z % 16) ||
        ((box_x * elem_bytes) % 16) || box_x > 256 || box_y > 256)
        return fail(ctx, TI_EINVAL, "tensor map: base/pitches/box must be 16-byte multiples, box <= 256 elements");
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)std::max(n, 1)};
    const cuuint64_t strides[2] = {pitch_y, pitch_z ? pitch_z : pitch_y * (uint64_t)h};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, TI_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (ctx) {
        if (ctx->tma_cache.size() >= 512) ctx->tma_cache.erase(ctx->tma_cache.begin() + 256, ctx->tma_cache.end());  // cold half
        TmaBlob blob;
        memcpy(&blob, out, sizeof blob);
        ctx->tma_cache.emplace_back(key, blob);
    }
    return TI_OK;
}
#endif

int tma_encode_u8_3d(ti_ctx* ctx, TiTensorMap* out, const void* base, int w, int h, int n, uint64_t pitch_y, uint64_t pitch_z,
                     int box_x, int box_y) {
    return tma_encode_3d(ctx, out, base, 1, w, h, n, pitch_y, pitch_z, box_x, box_y);
}

}  // namespace ti
