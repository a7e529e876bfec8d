// Probe: is the texture unit's bilinear filter exact enough to reproduce cv2.remap's fixed-point result,
// and how fast is it on B200?  (development probe; not part of the library)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void probe(cudaTextureObject_t tex, const uint32_t* lut, uint8_t* out, float* raw, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t e = lut[i];  // x0 | y0<<11 | fx<<22 | fy<<27   (x0,y0 >= 0 here)
        const float x = (float)((e & 2047u) * 32u + ((e >> 22) & 31u)) * (1.f / 32.f) + 0.5f;
        const float y = (float)(((e >> 11) & 2047u) * 32u + (e >> 27)) * (1.f / 32.f) + 0.5f;
        const float v = tex2D<float>(tex, x, y);  // normalized float read: texel / 255
        if (raw) raw[i] = v;
        const int S = __float2int_rn(v * (255.f * 1024.f));
        out[i] = (uint8_t)((S + 512) >> 10);
    }
}

int main() {
    const int W = 1280, H = 800, N = W * H;
    std::vector<uint8_t> img(N);
    srand(1);
    for (auto& p : img) p = rand() & 255;
    std::vector<uint32_t> lut(N);
    std::vector<uint8_t> want(N);
    std::vector<int> wantS(N);
    for (int i = 0; i < N; ++i) {
        int x0 = rand() % (W - 1), y0 = rand() % (H - 1), fx = rand() & 31, fy = rand() & 31;
        if (i < 1024) { fx = i & 31; fy = (i >> 5) & 31; }
        lut[i] = x0 | (y0 << 11) | (fx << 22) | ((uint32_t)fy << 27);
        int t00 = img[y0 * W + x0], t01 = img[y0 * W + x0 + 1], t10 = img[(y0 + 1) * W + x0], t11 = img[(y0 + 1) * W + x0 + 1];
        int S = t00 * (32 - fx) * (32 - fy) + t01 * fx * (32 - fy) + t10 * (32 - fx) * fy + t11 * fx * fy;
        wantS[i] = S;
        want[i] = (uint8_t)((S + 512) >> 10);
    }
    uint8_t* d_img; size_t pitch;
    CK(cudaMallocPitch(&d_img, &pitch, W, H));
    CK(cudaMemcpy2D(d_img, pitch, img.data(), W, W, H, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d_img; rd.res.pitch2D.desc = cudaCreateChannelDesc<unsigned char>();
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = H; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    uint32_t* d_lut; uint8_t* d_out; float* d_raw;
    CK(cudaMalloc(&d_lut, N * 4)); CK(cudaMalloc(&d_out, N)); CK(cudaMalloc(&d_raw, N * 4));
    CK(cudaMemcpy(d_lut, lut.data(), N * 4, cudaMemcpyHostToDevice));
    probe<<<148 * 8, 256>>>(tex, d_lut, d_out, d_raw, N);
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> got(N); std::vector<float> raw(N);
    CK(cudaMemcpy(got.data(), d_out, N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(raw.data(), d_raw, N * 4, cudaMemcpyDeviceToHost));
    long bad = 0; double maxerr = 0;
    for (int i = 0; i < N; ++i) {
        if (got[i] != want[i]) ++bad;
        double err = fabs((double)raw[i] * 255.0 * 1024.0 - wantS[i]);
        if (err > maxerr) maxerr = err;
    }
    printf("pixels %d mismatches %ld (%.4f%%)  max |S_tex - S_exact| = %.3f (in units of 1/1024 grey level)\n", N, bad, 100.0 * bad / N, maxerr);
    for (int i = 0; i < 6; ++i) printf("  e.g. S_exact %d  tex*261120 = %.4f\n", wantS[i * 37 + 5], raw[i * 37 + 5] * 261120.0);
    // throughput (sequential-ish coordinates like a real map)
    for (int i = 0; i < N; ++i) { int u = i % W, v = i / W; int x0 = (int)(u * 0.96f) + 3, y0 = (int)(v * 0.97f) + 2;
        lut[i] = x0 | (y0 << 11) | ((u * 7 & 31) << 22) | ((uint32_t)(v * 5 & 31) << 27); }
    CK(cudaMemcpy(d_lut, lut.data(), N * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) probe<<<148 * 8, 256>>>(tex, d_lut, d_out, nullptr, N);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 50; ++rep) probe<<<148 * 8, 256>>>(tex, d_lut, d_out, nullptr, N);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("tex remap (L2-resident 1 MPix frame): %.2f GPix/s\n", 50.0 * N / (ms * 1e-3) / 1e9);
    return 0;
}
