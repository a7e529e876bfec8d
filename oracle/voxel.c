/* Oracle of the voxel down-sampling stage  -  TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/__init__.py).
 *
 * The reference has no voxel code of its own: it hands depth images to nvblox, which integrates them at
 * `voxel_size` 0.05 m up to `tsdf_integrator_max_integration_distance_m` 10 m
 * (launch/thor_nvblox.launch.py:26-31).  What is restated here is the contract of ti_voxel_cloud
 * (include/thoringest.h): the pinhole back-projection of oracle/backproject.py (K of the depth image, pose
 * body_T_cam = M * world_T_camera, thor_slam/camera/rig.py:35-70, thor_slam/slam/adapters/isaac_ros.py:42-49),
 * evaluated in IEEE double with fused multiply-adds in a FIXED order, then floor() per axis:
 *
 *     a_u = 1e-3 * R[:,0] / fx,  a_v = 1e-3 * R[:,1] / fy,  a_c = 1e-3 * (R[:,2] - R[:,0]/fx * cx - R[:,1]/fy * cy)
 *     (each then multiplied by 1 / voxel_size, as is t)
 *     k = floor( fma(d, fma(a_u, u, fma(a_v, v, a_c)), t + 16384) ) - 16384      d = depth in millimetres; 16384 = bias of
 *                                                                                the 15-bit record fields, added before the one rounding
 *     valid = 0 < d <= max_depth_mm                                   examples/rgbd_stream.py:121-123 + the nvblox cap
 *
 * IEEE fma is correctly rounded everywhere, so these keys are bit-identical to the CUDA kernel's; tests also compare them
 * with floor(p / voxel_size) of the plain float64 numpy back-projection away from voxel boundaries.
 * "parity unpinned" by the reference (no voxel vectors exist there); pinned against the numpy float64 oracle instead.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle/voxel.c -o oracle/_build/libvoxel_oracle.so -lm   (oracle/voxel.py does it)
 */
#include <math.h>
#include <stdint.h>

void oracle_voxel_constants(const double k[4], const double m[12], double voxel, double out[12]) {
    const double inv = 1.0 / voxel;
    for (int r = 0; r < 3; ++r) {
        const double ax = m[4 * r + 0] / k[0], ay = m[4 * r + 1] / k[1], az = m[4 * r + 2];
        out[r] = (1e-3 * ax) * inv;
        out[3 + r] = (1e-3 * ay) * inv;
        out[6 + r] = (1e-3 * (az - ax * k[2] - ay * k[3])) * inv;
        out[9 + r] = m[4 * r + 3] * inv + 16384.0;
    }
}

/* keys: h*w*3 int32 (kx, ky, kz); valid: h*w u8.  Returns the number of valid pixels. */
int64_t oracle_voxel_keys(const uint16_t* depth, int w, int h, const double k[4], const double m[12], double voxel,
                          uint32_t max_depth_mm, int32_t* keys, uint8_t* valid) {
    double c[12];
    oracle_voxel_constants(k, m, voxel, c);
    int64_t n = 0;
    for (int v = 0; v < h; ++v) {
        const double bx = fma(c[3], (double)v, c[6]), by = fma(c[4], (double)v, c[7]), bz = fma(c[5], (double)v, c[8]);
        for (int u = 0; u < w; ++u) {
            const uint32_t d = depth[(int64_t)v * w + u];
            const double dd = (double)d, ud = (double)u;
            int32_t* o = keys + ((int64_t)v * w + u) * 3;
            o[0] = (int32_t)floor(fma(dd, fma(c[0], ud, bx), c[9])) - 16384;
            o[1] = (int32_t)floor(fma(dd, fma(c[1], ud, by), c[10])) - 16384;
            o[2] = (int32_t)floor(fma(dd, fma(c[2], ud, bz), c[11])) - 16384;
            const int ok = d != 0 && d <= max_depth_mm;
            valid[(int64_t)v * w + u] = (uint8_t)ok;
            n += ok;
        }
    }
    return n;
}
