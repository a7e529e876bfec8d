// Depth -> RGB registration shared by the stand-alone kernel (ti_register.cu) and the fused back-projection (ti_backproject.cu).
#pragma once
#include "ti_common.cuh"

namespace ti {

// depth -> RGB registration constants of one camera slot (ti_register.cu has the arithmetic's rationale)
struct RegConst {
    float a[9], t[3];
    float cx, cy, rfx, rfy, rcx, rcy;
    float guard;  // half-width of the band around x.5 inside which the reciprocal fast path hands over to the IEEE division
    int rw, rh;
};

// Which RGB pixel colours one depth pixel: the selection of register_pixel (ti_register.cu) - same float32 operations, each rounded
// on its own, so the float32 oracle picks the same RGB pixel - with the two IEEE divisions taken off the common path:
// q' = p.x * rcp(p.z) is within 2.4e-7 relative of fl(p.x / p.z), so u' = q' * fx + cx is within `guard` (a few 1e-3 px,
// bound derived in DESIGN.md) of the reference value; rint() of the two can only differ when u' lies within `guard` of a
// half-integer, and exactly those pixels (well under 1 %) are redone with the division.
__device__ __forceinline__ int reg_pixel_index(const RegConst& R, int u, int v, uint32_t d) {  // pixel index in the RGB image, -1: none
    const float fu = __fsub_rn((float)u, R.cx), fv = __fsub_rn((float)v, R.cy);
    const float z = __fmul_rn((float)d, 0.001f);
    float p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float r = __fadd_rn(__fadd_rn(__fmul_rn(R.a[3 * i], fu), __fmul_rn(R.a[3 * i + 1], fv)), R.a[3 * i + 2]);
        p[i] = __fadd_rn(__fmul_rn(r, z), R.t[i]);
    }
    if (d == 0 || !(p[2] > 0.f)) return -1;
#ifdef TI_EMULATE
    const float inv = 1.0f / p[2];
#else
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(p[2]));
#endif
    float ur = __fadd_rn(__fmul_rn(__fmul_rn(p[0], inv), R.rfx), R.rcx);
    float vr = __fadd_rn(__fmul_rn(__fmul_rn(p[1], inv), R.rfy), R.rcy);
    // far outside: rejected on either path (the relative error of the fast path cannot bring such a value back inside)
    if (!(ur > -2.f && ur < (float)R.rw + 1.f && vr > -2.f && vr < (float)R.rh + 1.f)) return -1;
    const float du = fabsf(__fsub_rn(__fsub_rn(ur, floorf(ur)), 0.5f)), dv = fabsf(__fsub_rn(__fsub_rn(vr, floorf(vr)), 0.5f));
    if (du < R.guard || dv < R.guard) {  // too close to a rounding boundary to trust the reciprocal: the reference's own operations
        ur = __fadd_rn(__fmul_rn(__fdiv_rn(p[0], p[2]), R.rfx), R.rcx);
        vr = __fadd_rn(__fmul_rn(__fdiv_rn(p[1], p[2]), R.rfy), R.rcy);
    }
    if (!(ur > -1.f && ur < (float)R.rw && vr > -1.f && vr < (float)R.rh)) return -1;
    const int iu = __float2int_rn(ur), iv = __float2int_rn(vr);
    if (iu < 0 || iu >= R.rw || iv < 0 || iv >= R.rh) return -1;
    return iv * R.rw + iu;
}

// host: constants of slot C
inline RegConst reg_constants(const CameraSlot& C) {
    RegConst R{};
    for (int i = 0; i < 9; ++i) R.a[i] = C.reg_a[i];
    for (int i = 0; i < 3; ++i) R.t[i] = C.reg_t[i];
    R.cx = C.reg_k[0]; R.cy = C.reg_k[1]; R.rfx = C.reg_k[2]; R.rfy = C.reg_k[3]; R.rcx = C.reg_k[4]; R.rcy = C.reg_k[5];
    R.rw = C.reg_rw; R.rh = C.reg_rh;
    // |u' - u| <= 3.6e-7 * |u - cx| + 2 * ulp(u) / 2: 1.7e-3 px for images up to 4096 px; twice that as the band
    R.guard = std::max(4e-3f, 1e-6f * (float)std::max(C.reg_rw, C.reg_rh));
    return R;
}

}  // namespace ti
