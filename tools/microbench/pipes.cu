// Issue-rate microbenchmark for the instruction classes the fused score kernel is built from.
// Not product code: it exists to choose the per-pixel arithmetic formulation from measured
// B200 pipe rates instead of guesses (results are summarised in profiles/).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;  // independent dependency chains per thread

#define DEF_KERNEL(NAME, DECL, BODY)                                                         \
    __global__ void __launch_bounds__(1024, 1) NAME(uint32_t* out, uint32_t seed, long long* cyc) { \
        DECL;                                                                                \
        long long t0 = clock64();                                                            \
        for (int it = 0; it < ITERS; ++it) {                                                 \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY; }                     \
        }                                                                                    \
        long long t1 = clock64();                                                            \
        uint32_t acc = 0;                                                                    \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc ^= SINK(c);                   \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                    \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                     \
    }

#define SINK(c) x[c]
#define U32DECL uint32_t x[CHAINS], y = seed | 1, z = seed * 3 + 7; for (int c = 0; c < CHAINS; ++c) x[c] = seed + c * 977 + threadIdx.x
DEF_KERNEL(k_iadd3, U32DECL, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)))
DEF_KERNEL(k_lop3, U32DECL, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)))
DEF_KERNEL(k_prmt, U32DECL, asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x[c]) : "r"(y)))
DEF_KERNEL(k_shf, U32DECL, asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(x[c]) : "r"(y)))
DEF_KERNEL(k_imad, U32DECL, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)))
DEF_KERNEL(k_vimnmx3, U32DECL, x[c] = __vimax3_u32(x[c], y + c, z))
DEF_KERNEL(k_vimnmx3_16x2, U32DECL, x[c] = __vimax3_u16x2(x[c], y + c, z))
DEF_KERNEL(k_vabsdiff4, U32DECL, x[c] = __vsadu4(x[c], y) + z)
DEF_KERNEL(k_dp4a, U32DECL, x[c] = __dp4a(x[c], y, z))
DEF_KERNEL(k_isetp_sel, U32DECL, x[c] = (x[c] > y + c) ? x[c] - z : x[c] + y)
DEF_KERNEL(k_redux, U32DECL, x[c] = __reduce_add_sync(0xffffffffu, x[c]) + y)
#undef SINK
#define SINK(c) __float_as_uint(f[c])
#define F32DECL float f[CHAINS], a = __uint_as_float((seed & 0xFFFF) | 0x3F800000), b = 0.5f; for (int c = 0; c < CHAINS; ++c) f[c] = 1.0f + c + threadIdx.x
DEF_KERNEL(k_ffma, F32DECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)))
DEF_KERNEL(k_ffma_rz, F32DECL, asm volatile("fma.rz.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)))
DEF_KERNEL(k_fadd, F32DECL, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(a)))
DEF_KERNEL(k_fmnmx, F32DECL, asm volatile("max.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(a)))
DEF_KERNEL(k_fmnmx3, F32DECL, asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)))
DEF_KERNEL(k_fsetp_fsel, F32DECL, { float t; asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; selp.f32 %0, %1, %3, p; }" : "=f"(t) : "f"(f[c]), "f"(a), "f"(b)); f[c] = t + a; })
DEF_KERNEL(k_ffma_sat, F32DECL, asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)))
DEF_KERNEL(k_mufu_rcp, F32DECL, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[c])))
DEF_KERNEL(k_i2fp, F32DECL, f[c] = __int2float_rn(__float_as_int(f[c]) & 0xFF) + a)
DEF_KERNEL(k_fsetp_sel, F32DECL, f[c] = (f[c] < a) ? f[c] + b : f[c])
#undef SINK
#define SINK(c) (uint32_t)(d[c] ^ (d[c] >> 32))
#define F2DECL unsigned long long d[CHAINS], a2, b2; { float2 t = make_float2(1.0001f, 0.9999f); a2 = *(unsigned long long*)&t; t = make_float2(0.5f, 0.25f); b2 = *(unsigned long long*)&t; } for (int c = 0; c < CHAINS; ++c) { float2 t = make_float2(1.0f + c, 2.0f + threadIdx.x); d[c] = *(unsigned long long*)&t; }
DEF_KERNEL(k_ffma2, F2DECL, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(a2), "l"(b2)))
DEF_KERNEL(k_ffma2_rz, F2DECL, asm volatile("fma.rz.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(a2), "l"(b2)))
DEF_KERNEL(k_fadd2, F2DECL, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(a2)))
#undef SINK
// mixed: alternate FMA-pipe and ALU-pipe ops (are the pipes co-issued?)
#define SINK(c) (x[c] ^ __float_as_uint(f[c]))
#define MIXDECL uint32_t x[CHAINS], y = seed | 1, z = seed * 3 + 7; float f[CHAINS], a = 1.0001f, b = 0.5f; for (int c = 0; c < CHAINS; ++c) { x[c] = seed + c + threadIdx.x; f[c] = 1.0f + c; }
DEF_KERNEL(k_mix_ffma_iadd, MIXDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)))
DEF_KERNEL(k_mix_ffma_prmt, MIXDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x[c]) : "r"(y)))
DEF_KERNEL(k_mix_imad_lop3, MIXDECL, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); f[c] = __uint_as_float(__float_as_uint(f[c]) ^ y))
DEF_KERNEL(k_mix_idp_prmt, MIXDECL, x[c] = __dp4a(x[c], y, z); f[c] = __uint_as_float(__byte_perm(__float_as_uint(f[c]), y, 0x4321)))
DEF_KERNEL(k_mix_fmnmx3_prmt, MIXDECL, asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x[c]) : "r"(y)))
DEF_KERNEL(k_mix_fmnmx3_imad, MIXDECL, asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)))
DEF_KERNEL(k_mix_ffma_imad, MIXDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)))
DEF_KERNEL(k_mix_ffma2_prmt_imad, MIXDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); z = __byte_perm(z, y, 0x4321))
DEF_KERNEL(k_mix_ffma_mufu, MIXDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b)); if (c < 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[c])); x[c] += y)
#undef SINK
// shared-memory lookups: conflict-free per-lane replicated table vs plain 256-entry table
#define SINK(c) x[c]
#define LDSDECL __shared__ uint32_t tab[8192]; for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (i * 2654435761u) >> 19; __syncthreads(); uint32_t x[CHAINS]; const uint32_t lane = threadIdx.x & 31; for (int c = 0; c < CHAINS; ++c) x[c] = (seed + c * 97 + threadIdx.x * 31) & 255
DEF_KERNEL(k_lds_replicated, LDSDECL, x[c] = tab[((x[c] & 255) << 5) | lane] & 255)
DEF_KERNEL(k_lds_plain, LDSDECL, x[c] = tab[(x[c] * 37 + c) & 255] & 255)
DEF_KERNEL(k_atoms_spread, LDSDECL, x[c] = (x[c] * 37 + 11) & 255; atomicAdd(&tab[x[c]], 1u))
DEF_KERNEL(k_atoms_perwarp, LDSDECL, x[c] = (x[c] * 37 + 11) & 255; atomicAdd(&tab[((threadIdx.x >> 5) << 8) | x[c]], 1u))

template <typename K>
static void run(const char* name, K kernel, int ops_per_body, uint32_t* out, long long* cyc, int threads) {
    const int grid = 148;
    kernel<<<grid, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kernel<<<grid, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double warp_instr = (double)ITERS * CHAINS * ops_per_body * (threads / 32);
    // clk/body/SMSP: SMSP cycles consumed per loop body (all SASS ops of one BODY for one warp)
    const double clk_per_body = avg / ((double)ITERS * CHAINS * (threads / 32) / 4.0);
    printf("%-22s thr=%4d  %8.3f warp-instr/clk/SM  (%.2f per SMSP)  clk/body/SMSP=%6.3f  %7.3f ms  clk~%.0f MHz %s\n", name, threads,
           warp_instr / avg, warp_instr / avg / 4.0, clk_per_body, ms, avg / (ms * 1e3), err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {1024}) {
        run("IADD", k_iadd3, 1, out, cyc, threads);
        run("LOP3", k_lop3, 1, out, cyc, threads);
        run("PRMT", k_prmt, 1, out, cyc, threads);
        run("SHF", k_shf, 1, out, cyc, threads);
        run("IMAD", k_imad, 1, out, cyc, threads);
        run("VIMNMX3.U32", k_vimnmx3, 1, out, cyc, threads);
        run("VIMNMX3.U16x2", k_vimnmx3_16x2, 1, out, cyc, threads);
        run("VABSDIFF4.ACC(+add)", k_vabsdiff4, 1, out, cyc, threads);
        run("IDP4A", k_dp4a, 1, out, cyc, threads);
        run("ISETP+SEL(+2 add)", k_isetp_sel, 4, out, cyc, threads);
        run("REDUX(+add)", k_redux, 2, out, cyc, threads);
        run("FFMA", k_ffma, 1, out, cyc, threads);
        run("FFMA.RZ", k_ffma_rz, 1, out, cyc, threads);
        run("FADD", k_fadd, 1, out, cyc, threads);
        run("FMNMX", k_fmnmx, 1, out, cyc, threads);
        run("FMNMX3", k_fmnmx3, 1, out, cyc, threads);
        run("FSETP+FSEL+FADD", k_fsetp_fsel, 3, out, cyc, threads);
        run("FFMA.SAT", k_ffma_sat, 1, out, cyc, threads);
        run("MUFU.RCP", k_mufu_rcp, 1, out, cyc, threads);
        run("LOP+I2FP+FADD", k_i2fp, 3, out, cyc, threads);
        run("FSETP+FADD+SEL", k_fsetp_sel, 3, out, cyc, threads);
        run("FFMA2", k_ffma2, 1, out, cyc, threads);
        run("FFMA2.RZ", k_ffma2_rz, 1, out, cyc, threads);
        run("FADD2", k_fadd2, 1, out, cyc, threads);
        run("mix FFMA+IADD", k_mix_ffma_iadd, 2, out, cyc, threads);
        run("mix FFMA+PRMT", k_mix_ffma_prmt, 2, out, cyc, threads);
        run("mix IMAD+LOP3", k_mix_imad_lop3, 2, out, cyc, threads);
        run("mix IDP4A+PRMT", k_mix_idp_prmt, 2, out, cyc, threads);
        run("mix FMNMX3+PRMT", k_mix_fmnmx3_prmt, 2, out, cyc, threads);
        run("mix FMNMX3+IMAD", k_mix_fmnmx3_imad, 2, out, cyc, threads);
        run("mix FFMA+IMAD", k_mix_ffma_imad, 2, out, cyc, threads);
        run("mix FFMA+IMAD+PRMT", k_mix_ffma2_prmt_imad, 3, out, cyc, threads);
        run("mix 8FFMA+2MUFU+8IADD", k_mix_ffma_mufu, 2, out, cyc, threads);
        run("LDS replicated(+2)", k_lds_replicated, 4, out, cyc, threads);
        run("LDS plain256(+3)", k_lds_plain, 4, out, cyc, threads);
        run("ATOMS spread(+3)", k_atoms_spread, 4, out, cyc, threads);
        run("ATOMS per-warp(+3)", k_atoms_perwarp, 4, out, cyc, threads);
    }
    return 0;
}
