// Device twin of pyscenedetect_b200/synth.py:render_frames (bit-exact, 32-bit unsigned math).
#include "psd_common.cuh"

namespace psd {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) psd_synth_kernel(uint8_t* __restrict__ out,
                                                        const int32_t* __restrict__ params, int width,
                                                        int height, int64_t frame_stride) {
    const int64_t f = blockIdx.y;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(params + f * 24);
    const uint32_t gain = row[12], seed_t = row[13];
    const int nshift = (int)row[15];
    const int npix = width * height;
    uint8_t* o = out + f * frame_stride;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)idx / (uint32_t)width, x = (uint32_t)idx - y * (uint32_t)width;
        const uint32_t xy = (x * y) >> 8;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t p = (row[c] * x + row[3 + c] * y + row[6 + c] * xy) >> 4;
            uint32_t v = (p + row[9 + c]) & 255u;
            v = row[19 + c] + ((v * row[16 + c]) >> 8);
            v = (v * gain) >> 8;
            int vi = (int)v;
            if (nshift < 32) {
                const uint32_t h = mix32(seed_t + (uint32_t)idx * 3u + (uint32_t)c);
                vi += (int)(h >> nshift) - (1 << (31 - nshift));
            }
            o[(int64_t)idx * 3 + c] = (uint8_t)min(max(vi, 0), 255);
        }
    }
}

int launch_synth(uint8_t* out, const int32_t* d_params, int64_t n, int width, int height,
                 int64_t frame_stride, cudaStream_t stream) {
    PSD_REQUIRE(n > 0 && n <= 65535, "synth batch out of range");
    const int npix = width * height;
    dim3 grid((unsigned)min((npix + 255) / 256, 2048), (unsigned)n);
    psd_synth_kernel<<<grid, 256, 0, stream>>>(out, d_params, width, height, frame_stride);
    PSD_CHECK_LAUNCH();
    count_launch();
    return PSD_OK;
}

}  // namespace psd
