// Compute-only rate of the per-pixel HSV + SAD arithmetic (no memory traffic): how many pixels
// per second can each formulation in csrc/hsv_math.cuh retire?  The fused kernel needs
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

// ---- experiment: table lookups through the texture unit instead of the replicated smem LUT ----
__device__ __forceinline__ void hsv_px_tex(const uint32_t (&w)[12], int KB, cudaTextureObject_t ts,
                                           cudaTextureObject_t th, uint32_t& oh, uint32_t& os, uint32_t& ov);
template <int KB>
__device__ __forceinline__ void hsv_px_tex1(const uint32_t (&w)[12], cudaTextureObject_t ts, cudaTextureObject_t th,
                                          uint32_t& oh, uint32_t& os, uint32_t& ov) {
    const float B = magic_byte_dp4a<(KB + 0) & 3>(w[(KB + 0) >> 2]);
    const float G = magic_byte<(KB + 1) & 3>(w[(KB + 1) >> 2]);
    const float R = magic_byte_dp4a<(KB + 2) & 3>(w[(KB + 2) >> 2]);
    const float V = fmax3(B, G, R);
    const float mn = fmin3(B, G, R);
    const float d = V - mn;
    const float sdivp = tex1D<float>(ts, V - 8388608.0f);
    const float hdivp = tex1D<float>(th, d);
    const float yS = fma_rz(d, sdivp, 32768.5f);
    const float hR = G - B;
    const float hG = fmaf(d, 2.0f, B - R);
    const float hB = fmaf(d, 4.0f, R - G);
    const float h = (V == R) ? hR : ((V == G) ? hG : hB);
    float yH = fma_rm(h, hdivp, 49152.5f);
    yH = fmaf(fma_sat(yH, -256.0f, 12582912.0f), 180.0f, yH);
    oh = __float_as_uint(yH);
    os = __float_as_uint(yS);
    ov = __float_as_uint(V);
}
__device__ __forceinline__ void hsv16_tex(const uint32_t (&w)[12], Px16& o, cudaTextureObject_t ts, cudaTextureObject_t th) {
    uint32_t h[16], s[16], v[16];
#define PSD_PX(i) hsv_px_tex1<3 * (i)>(w, ts, th, h[i], s[i], v[i]);
    PSD_PX(0) PSD_PX(1) PSD_PX(2) PSD_PX(3) PSD_PX(4) PSD_PX(5) PSD_PX(6) PSD_PX(7)
    PSD_PX(8) PSD_PX(9) PSD_PX(10) PSD_PX(11) PSD_PX(12) PSD_PX(13) PSD_PX(14) PSD_PX(15)
#undef PSD_PX
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.h[j] = __byte_perm(__byte_perm(h[4 * j], h[4 * j + 1], 0x0051), __byte_perm(h[4 * j + 2], h[4 * j + 3], 0x0051), 0x5410);
        o.s[j] = __byte_perm(__byte_perm(s[4 * j], s[4 * j + 1], 0x0051), __byte_perm(s[4 * j + 2], s[4 * j + 3], 0x0051), 0x5410);
        o.v[j] = __byte_perm(__byte_perm(v[4 * j], v[4 * j + 1], 0x0040), __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x0040), 0x5410);
    }
}
__device__ cudaTextureObject_t g_ts, g_th;

template <int VARIANT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) rate_kernel(uint32_t* out, uint32_t seed, long long* cyc) {
    extern __shared__ __align__(128) float lut[];
    __shared__ int32_t sdiv[256], hdiv[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sdiv[i] = i ? __double2int_rn(1044480.0 / (double)i) : 0;
        hdiv[i] = i ? __double2int_rn(737280.0 / (6.0 * (double)i)) : 0;
    }
    LutView lv{0, 0};
    const LutView7 lv7 = make_lut7((uint32_t)__cvta_generic_to_shared(lut), threadIdx.x & 31);
    if (VARIANT == 7 || VARIANT == 8) lut_fill7(lut, threadIdx.x, blockDim.x);
    if (VARIANT == 4 || VARIANT == 6) {
        lut_fill(lut, threadIdx.x, blockDim.x);
        lv.s_addr = (uint32_t)__cvta_generic_to_shared(lut) + (threadIdx.x & 31) * 4;
        lv.h_addr = lv.s_addr + 128;
    }
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
        if (VARIANT == 17)
            hsv16_tex(w, cur, g_ts, g_th);
        else if (VARIANT == 7 || VARIANT == 8)
            hsv16_v7<VARIANT == 8>(w, cur, lv7, seed << 11);  // 12345 << 11 is not 2^24, irrelevant for the rate
        else if (VARIANT == 6)
            hsv16_v4pair(w, cur, lv);
        else if (VARIANT == 4)
            hsv16_v4(w, cur, lv);
        else
            hsv16<VARIANT>(w, cur, sdiv, hdiv);
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
    const int smem = (VARIANT == 4 || VARIANT == 6 || VARIANT == 7 || VARIANT == 8) ? 65536 : 0;
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

static cudaTextureObject_t make_tex(const float* host) {
    cudaArray_t arr;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaMallocArray(&arr, &cd, 256);
    cudaMemcpy2DToArray(arr, 0, 0, host, 256 * sizeof(float), 256 * sizeof(float), 1, cudaMemcpyHostToDevice);
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td{}; td.addressMode[0] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t t = 0; cudaCreateTextureObject(&t, &rd, &td, nullptr);
    return t;
}

int main() {
    {
        float hs[256], hh[256];
        for (int i = 0; i < 256; ++i) {
            hs[i] = i ? (float)nearbyint(1044480.0 / i) / 4096.0f : 0.0f;
            hh[i] = i ? (float)nearbyint(737280.0 / (6.0 * i)) / 4096.0f : 0.0f;
        }
        cudaTextureObject_t ts = make_tex(hs), th = make_tex(hh);
        cudaMemcpyToSymbol(g_ts, &ts, sizeof(ts)); cudaMemcpyToSymbol(g_th, &th, sizeof(th));
    }
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 3 * 1024 * 4); cudaMalloc(&cyc, 148 * 3 * 8);
    if (getenv("HSV_RATE_ALL")) {
        run<0, 256, 3>(out, cyc);
        run<1, 256, 3>(out, cyc);
        run<2, 256, 3>(out, cyc);
        run<3, 256, 3>(out, cyc);
        run<2, 256, 2>(out, cyc);
        run<17, 768, 1>(out, cyc);
    }
    printf("PSD_V4_PRMT_CHANNELS=%d\n", PSD_V4_PRMT_CHANNELS);
    run<4, 768, 1>(out, cyc);
    run<6, 768, 1>(out, cyc);
    printf("PSD_V7_ADDR=%d PSD_V7_HMNMX=%d\n", PSD_V7_ADDR, PSD_V7_HMNMX);
    run<7, 768, 1>(out, cyc);
    run<8, 768, 1>(out, cyc);
    run<7, 512, 1>(out, cyc);
    run<8, 1024, 1>(out, cyc);
    return 0;
}
