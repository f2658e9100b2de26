// Issue-rate microbenchmark, part 3 (round 2): open questions about the sub-partition's two 16-lane halves
// and the 32-lane "wide" class, asked while trimming the pixel-pair HSV loop (csrc/hsv_half2.cuh):
//   * which half does F2FP (cvt.rz.f16x2.f32) / I2IP (cvt.pack.sat.u8.s32) / HADD2 / 3-input IADD3 / LEA use?
//   * what does a wide instruction cost when it sits between half-rate ones (A F W A F W ...) compared with
//     the same instructions grouped (A F A F ... W W ...)?
//   * LDS.64 from a per-lane replicated table, FFMA with an immediate addend.
// Not product code.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes3 pipes3.cu ; run: ./pipes3
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

#define DEF_KERNEL(NAME, BODY)                                                                       \
    __global__ void __launch_bounds__(1024, 1) NAME(uint32_t* out, uint32_t seed, long long* cyc) {  \
        __shared__ float lut[256 * 32 * 2 / 8];                                                      \
        uint32_t x[CHAINS], y = ((seed + threadIdx.x * 7) & 0x03FF03FFu) | 0x64006400u, z = 0x3C003C00u + (threadIdx.x & 3); \
        float f[CHAINS], a = 1.0001f, b = 0.5f;                                                      \
        for (int c = 0; c < CHAINS; ++c) {                                                           \
            x[c] = ((seed + c * 977 + threadIdx.x) & 0x03FF03FFu) | 0x64006400u;                     \
            f[c] = 1.0f + c;                                                                         \
        }                                                                                            \
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) lut[i] = (float)i;                      \
        __syncthreads();                                                                             \
        const uint32_t lbase = (uint32_t)__cvta_generic_to_shared(lut) + (threadIdx.x & 31) * 8;     \
        (void)lbase; (void)a; (void)b; (void)z;                                                     \
        long long t0 = clock64();                                                                    \
        for (int it = 0; it < ITERS; ++it) {                                                         \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY; }                             \
        }                                                                                            \
        long long t1 = clock64();                                                                    \
        uint32_t acc = y ^ z;                                                                        \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ __float_as_uint(f[c]);      \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                            \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                             \
    }

#define HFMA2(x) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(z), "r"(y))
#define HADD2(x) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y))
#define PRMT(x) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x) : "r"(y))
#define LOP3(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(x) : "r"(y), "r"(z))
#define IMAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define IADD(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define IADD3(x) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z))
#define VABS4(x) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(x) : "r"(y), "r"(z))
#define SHF(x) asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y))
#define LEA(x) x = (x << 3) + y
#define FFMA(v) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v) : "f"(a), "f"(b))
#define FFMAI(v) asm volatile("fma.rz.f32 %0, %0, %1, 0f47000040;" : "+f"(v) : "f"(a))
#define F2FP(x, v) asm volatile("{ .reg .f32 t; mov.b32 t, %0; cvt.rz.f16x2.f32 %0, t, %1; }" : "+r"(x) : "f"(v))
#define I2IP(x) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define H2F(x, v) { float t; asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi; }" : "=f"(t) : "r"(x)); v = t; }
#define IDP2A(x) x = __dp2a_lo(x, 0x00008000u, y)
#define LDS64(xx, v) { float t0_, t1_; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(t0_), "=f"(t1_) : "r"(lbase + ((xx) & 0x700))); v += t0_ + t1_; }
#define LDS32(x, v) { float t; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(lbase + ((x) & 0x700))); v += t; }

DEF_KERNEL(k_hadd2, HADD2(x[c]))
DEF_KERNEL(k_hadd2_prmt, HADD2(x[c]); PRMT(y))
DEF_KERNEL(k_hadd2_imad, HADD2(x[c]); IMAD(y))
DEF_KERNEL(k_hadd2_hfma2, HADD2(x[c]); HFMA2(y))
DEF_KERNEL(k_f2fp_only, F2FP(x[c], f[c]))
DEF_KERNEL(k_f2fp_prmt, F2FP(x[c], f[c]); PRMT(y))
DEF_KERNEL(k_f2fp_imad, F2FP(x[c], f[c]); IMAD(y))
DEF_KERNEL(k_f2fp_hfma2, F2FP(x[c], f[c]); HFMA2(y))
DEF_KERNEL(k_i2ip, I2IP(x[c]))
DEF_KERNEL(k_i2ip_prmt, I2IP(x[c]); PRMT(y))
DEF_KERNEL(k_i2ip_imad, I2IP(x[c]); IMAD(y))
DEF_KERNEL(k_iadd3, IADD3(x[c]))
DEF_KERNEL(k_iadd3_prmt, IADD3(x[c]); PRMT(y))
DEF_KERNEL(k_iadd3_imad, IADD3(x[c]); IMAD(y))
DEF_KERNEL(k_iadd3_prmt_imad, IADD3(x[c]); PRMT(y); IMAD(z))
DEF_KERNEL(k_lea, LEA(x[c]))
DEF_KERNEL(k_lea_prmt, LEA(x[c]); PRMT(y))
DEF_KERNEL(k_lea_imad, LEA(x[c]); IMAD(y))
DEF_KERNEL(k_vabs4_imad, VABS4(x[c]); IMAD(y))
DEF_KERNEL(k_vabs4_prmt, VABS4(x[c]); PRMT(y))
DEF_KERNEL(k_ffmai, FFMAI(f[c]))
DEF_KERNEL(k_ffmai_prmt, FFMAI(f[c]); PRMT(y))
DEF_KERNEL(k_ffmai_imad, FFMAI(f[c]); IMAD(y))
DEF_KERNEL(k_ffmai_prmt_imad, FFMAI(f[c]); PRMT(y); IMAD(z))
// A F W interleaved (per chain) vs grouped (all A/F first, then all W)
DEF_KERNEL(k_afw_interleaved, PRMT(x[c]); IMAD(y); FFMAI(f[c]))
DEF_KERNEL(k_af2w, PRMT(x[c]); IMAD(y); PRMT(z); IMAD(y); FFMAI(f[c]))
__global__ void __launch_bounds__(1024, 1) k_afw_grouped(uint32_t* out, uint32_t seed, long long* cyc) {
    uint32_t x[CHAINS], y = seed + threadIdx.x, z = seed * 3 + threadIdx.x;
    float f[CHAINS], a = 1.0001f;
    for (int c = 0; c < CHAINS; ++c) { x[c] = seed + c + threadIdx.x; f[c] = 1.0f + c; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { PRMT(x[c]); IMAD(y); }
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { FFMAI(f[c]); }
    }
    long long t1 = clock64();
    uint32_t acc = y ^ z;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ __float_as_uint(f[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
DEF_KERNEL(k_h2f_ffmai_prmt, H2F(x[c], f[c]); FFMAI(f[c]); PRMT(y))
DEF_KERNEL(k_pixelmix, PRMT(x[c]); HFMA2(y); IDP2A(z); FFMAI(f[c]); LOP3(x[c]); H2F(y, f[c]))
DEF_KERNEL(k_lds32, LDS32(x[c], f[c]); x[c] += 0x100)
DEF_KERNEL(k_lds64, LDS64(x[c], f[c]); x[c] += 0x100)

template <typename K>
static void run(const char* name, K kernel, int ops_per_body, uint32_t* out, long long* cyc) {
    const int grid = 148, threads = 1024;
    kernel<<<grid, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    kernel<<<grid, threads>>>(out, 12345u, cyc);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double clk_per_body = avg / ((double)ITERS * CHAINS * (threads / 32) / 4.0);
    printf("%-30s ops/body=%d  clk/body/SMSP=%6.3f  clk/op=%6.3f %s\n", name, ops_per_body, clk_per_body,
           clk_per_body / ops_per_body, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
#define RUN(k, n) run(#k, k, n, out, cyc)
    RUN(k_hadd2, 1); RUN(k_hadd2_prmt, 2); RUN(k_hadd2_imad, 2); RUN(k_hadd2_hfma2, 2);
    RUN(k_f2fp_only, 1); RUN(k_f2fp_prmt, 2); RUN(k_f2fp_imad, 2); RUN(k_f2fp_hfma2, 2);
    RUN(k_i2ip, 1); RUN(k_i2ip_prmt, 2); RUN(k_i2ip_imad, 2);
    RUN(k_iadd3, 1); RUN(k_iadd3_prmt, 2); RUN(k_iadd3_imad, 2); RUN(k_iadd3_prmt_imad, 3);
    RUN(k_lea, 1); RUN(k_lea_prmt, 2); RUN(k_lea_imad, 2);
    RUN(k_vabs4_imad, 2); RUN(k_vabs4_prmt, 2);
    RUN(k_ffmai, 1); RUN(k_ffmai_prmt, 2); RUN(k_ffmai_imad, 2); RUN(k_ffmai_prmt_imad, 3);
    RUN(k_afw_interleaved, 3); RUN(k_af2w, 5); RUN(k_afw_grouped, 3);
    RUN(k_h2f_ffmai_prmt, 3); RUN(k_pixelmix, 6);
    RUN(k_lds32, 2); RUN(k_lds64, 2);
    return 0;
}
