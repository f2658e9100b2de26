// Issue-rate microbenchmark, part 2: the packed 16-bit / half2 instruction classes the pixel-pair
// HSV formulation (csrc/hsv_half2.cuh) is built from, and their mixes with the alu-half (PRMT/LOP3)
// and fma-half (IMAD/IDP) classes measured by pipes.cu.  Not product code.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu ; run: ./pipes2
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

#define DEF_KERNEL(NAME, BODY)                                                                       \
    __global__ void __launch_bounds__(1024, 1) NAME(uint32_t* out, uint32_t seed, long long* cyc) {  \
        uint32_t x[CHAINS], y = (seed & 0x03FF03FFu) | 0x64006400u, z = 0x3C003C00u;                 \
        float f[CHAINS], a = 1.0001f, b = 0.5f;                                                      \
        for (int c = 0; c < CHAINS; ++c) {                                                           \
            x[c] = ((seed + c * 977 + threadIdx.x) & 0x03FF03FFu) | 0x64006400u;                     \
            f[c] = 1.0f + c;                                                                         \
        }                                                                                            \
        long long t0 = clock64();                                                                    \
        for (int it = 0; it < ITERS; ++it) {                                                         \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY; }                             \
        }                                                                                            \
        long long t1 = clock64();                                                                    \
        uint32_t acc = 0;                                                                            \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ __float_as_uint(f[c]);      \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                            \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                             \
    }

#define HFMA2(x) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(z), "r"(y))
#define HADD2(x) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y))
#define HSET2(x) asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y))
#define HMNMX2(x) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y))
__device__ __forceinline__ uint32_t hmax3(uint32_t a, uint32_t b, uint32_t c) {
    __half2 r = __hmax2(__hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b)), *reinterpret_cast<__half2*>(&c));
    return *reinterpret_cast<uint32_t*>(&r);
}
#define VHMNMX(x) x = hmax3(x, y, z)
#define VIMNMX3(x) x = __vimax3_u16x2(x, y, z)
#define PRMT(x) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x) : "r"(y))
#define LOP3(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(x) : "r"(y), "r"(z))
#define IMAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define IDP2A(x) x = __dp2a_lo(x, 0x00008000u, y)
#define IADD(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define FFMA(v) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v) : "f"(a), "f"(b))
#define FFMA2RM(v0, v1) { unsigned long long d, aa, bb; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(v0), "f"(v1)); \
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a)); asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b)); \
    asm volatile("fma.rm.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(aa), "l"(bb)); asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(d)); }
// cvt of one half lane to f32 (HADD2.F32); the result is folded back so the chain stays alive
#define H2F(x, v) { float t; asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi; }" : "=f"(t) : "r"(x)); v = t; }

DEF_KERNEL(k_hfma2, HFMA2(x[c]))
DEF_KERNEL(k_hadd2, HADD2(x[c]))
DEF_KERNEL(k_hset2, HSET2(x[c]))
DEF_KERNEL(k_hmnmx2, HMNMX2(x[c]))
DEF_KERNEL(k_vhmnmx, VHMNMX(x[c]))
DEF_KERNEL(k_idp2a, IDP2A(x[c]))
DEF_KERNEL(k_h2f, H2F(x[c], f[c]); x[c] += __float_as_uint(f[c]))
DEF_KERNEL(k_mix_hfma2_prmt, HFMA2(x[c]); PRMT(y))
DEF_KERNEL(k_mix_hfma2_imad, HFMA2(x[c]); IMAD(y))
DEF_KERNEL(k_mix_hfma2_ffma, HFMA2(x[c]); FFMA(f[c]))
DEF_KERNEL(k_mix_hfma2_iadd, HFMA2(x[c]); IADD(y))
DEF_KERNEL(k_mix_hset2_lop3, HSET2(x[c]); LOP3(y))
DEF_KERNEL(k_mix_hset2_imad, HSET2(x[c]); IMAD(y))
DEF_KERNEL(k_mix_vhmnmx_prmt, VHMNMX(x[c]); PRMT(y))
DEF_KERNEL(k_mix_vhmnmx_imad, VHMNMX(x[c]); IMAD(y))
DEF_KERNEL(k_mix_vimnmx3_hfma2, VIMNMX3(x[c]); HFMA2(y))
DEF_KERNEL(k_mix_idp2a_prmt, IDP2A(x[c]); PRMT(y))
DEF_KERNEL(k_mix_idp2a_hfma2, IDP2A(x[c]); HFMA2(y))
DEF_KERNEL(k_mix_h2f_prmt, H2F(x[c], f[c]); x[c] = __byte_perm(x[c], __float_as_uint(f[c]), 0x4321))
DEF_KERNEL(k_mix_h2f_imad, H2F(x[c], f[c]); x[c] = x[c] * y + __float_as_uint(f[c]))
DEF_KERNEL(k_mix_h2f_hfma2, H2F(x[c], f[c]); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(__float_as_uint(f[c]))))
DEF_KERNEL(k_mix_h2f_ffma, H2F(x[c], f[c]); FFMA(f[c]); x[c] += __float_as_uint(f[c]))
DEF_KERNEL(k_mix3_hfma2_prmt_idp2a, HFMA2(x[c]); PRMT(y); z = __dp2a_lo(z, 0x00008000u, y))
DEF_KERNEL(k_mix4_pixel, HFMA2(x[c]); PRMT(y); z = __dp2a_lo(z, 0x00008000u, y); FFMA(f[c]))
DEF_KERNEL(k_ffma2rm, if ((c & 1) == 0) FFMA2RM(f[c], f[c + 1]))

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
    printf("%-26s ops/body=%d  clk/body/SMSP=%6.3f  clk/op=%6.3f %s\n", name, ops_per_body, clk_per_body,
           clk_per_body / ops_per_body, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run("HFMA2", k_hfma2, 1, out, cyc);
    run("HADD2", k_hadd2, 1, out, cyc);
    run("HSET2.BM", k_hset2, 1, out, cyc);
    run("HMNMX2", k_hmnmx2, 1, out, cyc);
    run("VHMNMX (3-in half2)", k_vhmnmx, 1, out, cyc);
    run("IDP2A", k_idp2a, 1, out, cyc);
    run("HADD2.F32 + IADD", k_h2f, 2, out, cyc);
    run("mix HFMA2+PRMT", k_mix_hfma2_prmt, 2, out, cyc);
    run("mix HFMA2+IMAD", k_mix_hfma2_imad, 2, out, cyc);
    run("mix HFMA2+FFMA", k_mix_hfma2_ffma, 2, out, cyc);
    run("mix HFMA2+IADD", k_mix_hfma2_iadd, 2, out, cyc);
    run("mix HSET2+LOP3", k_mix_hset2_lop3, 2, out, cyc);
    run("mix HSET2+IMAD", k_mix_hset2_imad, 2, out, cyc);
    run("mix VHMNMX+PRMT", k_mix_vhmnmx_prmt, 2, out, cyc);
    run("mix VHMNMX+IMAD", k_mix_vhmnmx_imad, 2, out, cyc);
    run("mix VIMNMX3.U16x2+HFMA2", k_mix_vimnmx3_hfma2, 2, out, cyc);
    run("mix IDP2A+PRMT", k_mix_idp2a_prmt, 2, out, cyc);
    run("mix IDP2A+HFMA2", k_mix_idp2a_hfma2, 2, out, cyc);
    run("mix HADD2.F32+PRMT", k_mix_h2f_prmt, 2, out, cyc);
    run("mix HADD2.F32+IMAD", k_mix_h2f_imad, 2, out, cyc);
    run("mix HADD2.F32+HFMA2", k_mix_h2f_hfma2, 2, out, cyc);
    run("mix HADD2.F32+FFMA+IADD", k_mix_h2f_ffma, 3, out, cyc);
    run("mix HFMA2+PRMT+IDP2A", k_mix3_hfma2_prmt_idp2a, 3, out, cyc);
    run("mix HFMA2+PRMT+IDP2A+FFMA", k_mix4_pixel, 4, out, cyc);
    run("FFMA2.RM (per 2 chains)", k_ffma2rm, 1, out, cyc);
    return 0;
}
