// Compute-only rate of the per-pixel HSV + SAD arithmetic (no memory traffic): how many pixels
// per second can the two formulations the library ships (csrc/hsv_math.cuh, csrc/hsv_half2.cuh) retire?  The fused kernel needs
// 2.18 Tpx/s for 100 % of the measured HBM roofline (6541.8 GB/s / 3 B per pixel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hsv_rate hsv_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#include "../../pyscenedetect_b200/csrc/hsv_math.cuh"
#include "../../pyscenedetect_b200/csrc/hsv_half2.cuh"

using namespace psd;
constexpr int ITERS = 2048;

template <int VARIANT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) rate_kernel(uint32_t* out, uint32_t seed, long long* cyc) {
    extern __shared__ __align__(128) float lut[];
    const LutView7 lv7 = make_lut7((uint32_t)__cvta_generic_to_shared(lut), threadIdx.x & 31);
    if (VARIANT == 7) lut_fill7(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    uint32_t w[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) w[j] = seed * (j + 1) + threadIdx.x * 2654435761u;
    Px16 prev;
#pragma unroll
    for (int j = 0; j < 4; ++j) prev.h[j] = prev.s[j] = prev.v[j] = 0;
    uint32_t sh = 0, ss = 0, sv = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        Px16 cur;
        if (VARIANT == 7)
            hsv16_v7(w, cur, lv7);
        else
            hsv16_f32x2(w, cur);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sh = __vsadu4(cur.h[j], prev.h[j]) + sh;
            ss = __vsadu4(cur.s[j], prev.s[j]) + ss;
            sv = __vsadu4(cur.v[j], prev.v[j]) + sv;
        }
        prev = cur;
#pragma unroll
        for (int j = 0; j < 12; ++j) w[j] += 0x9E3779B9u + j;  // 12 IADD per 16 px of input churn
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sh ^ ss ^ sv;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int VARIANT, int THREADS, int MINB>
void run(uint32_t* out, long long* cyc) {
    const int grid = 148 * MINB;
    const int smem = (VARIANT == 7) ? 65536 : 0;
    cudaFuncSetAttribute(rate_kernel<VARIANT, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<VARIANT, THREADS, MINB><<<grid, THREADS, smem>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<VARIANT, THREADS, MINB><<<grid, THREADS, smem>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double px = (double)grid * THREADS * 16 * ITERS;
    printf("variant %d (%4d thr x %d CTA/SM): %8.3f Gpx/s  (%.3f ms)  = %.2f %% of the 2180 Gpx/s roofline pixel rate %s\n",
           VARIANT, THREADS, MINB, px / ms / 1e6, ms, 100.0 * px / ms / 1e6 / 2180.0,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 3 * 1024 * 4); cudaMalloc(&cyc, 148 * 3 * 8);
    run<2, 256, 3>(out, cyc);   // generic kernel arithmetic (hsv_math.cuh)
    run<7, 768, 1>(out, cyc);   // warp-specialised kernel arithmetic (hsv_half2.cuh)
    run<7, 1024, 1>(out, cyc);
    return 0;
}
