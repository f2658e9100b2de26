// Canny classify on pixel PAIRS (two 16-bit lanes per register): Sobel, L1 magnitude, non-maximum
// suppression and the double threshold of cv2.Canny(V, low, high) (aperture 3, L2gradient=False) straight
// into two bit planes.  Restated from OpenCV's algorithm; oracle/intmath.py:canny is the CPU twin, and
// tests/test_edge_bits_model.py replays THIS arithmetic (float16 lanes and all) in numpy against it.
//
// Number representation.  Every quantity of the gradient stage is a small integer n (|n| <= 2040) and is
// kept as the binary16 value n * 2^-19 (= n in units of the sub-normal 0x0020).  binary16 adds, subtracts
// and fused multiply-adds of such values are exact as long as |n| < 2048, which holds for every
// intermediate below, so the lanes carry the same integers as OpenCV's short/int buffers:
//   * a pixel byte v next to the constant byte 0x19 is the 16-bit pattern 0x1900 + v = (1280 + v) * 2^-19.
//     One PRMT per pair builds that from the row window - no conversion instruction.
//   * c(i) = V(i+1) - V(i-1)  : one HADD2 on two such pairs (the 1280 cancels).
//   * h(i) = V(i-1) + 2 V(i) + V(i+1) : INTEGER add of the three patterns (IMAD + IADD; lanes cannot
//     carry, 4 * 0x1900 + 1020 < 2^16) gives 0x6400 + h = the binary16 number 1024 + h; times 2^-19 (one
//     HMUL2, exact: 11 significant bits) it is (1024 + h) * 2^-19, and the 1024 cancels in gy.
//   * gx = c(y-1) + 2 c(y) + c(y+1),  gy = h(y+1) - h(y-1),  m = |gx| + |gy|,  p = m + 1.
// Non-negative binary16 patterns order like unsigned 16-bit integers, so maxima are VIMNMX3.U16x2 and the
// comparisons HSET2 with a bit-mask result.  OpenCV's tests
//     keep(x) = m > low  and  m >  m(left)  and m >= m(right)         (and the same for the other sectors)
// become ONE comparison per pixel:  p > max(p(left), m(right), low + 1).
// The sector (OpenCV's fixed-point tangent test, TG22 = 13573 / 2^15) needs 25 bits: two FP32 FMAs per
// pixel, fma(|gx|, 13573/32768, -|gy|) and fma(|gx|, 79109/32768, -|gy|), each rounded once from the
// exact value, so their SIGNS are exact; PRMT's sign-replicate mode turns the sign bits into lane masks.
//
// One thread owns 8 consecutive columns (one byte of each bit plane per row) and marches down a band of
// kBandRows rows.  It evaluates the gradient in BOTH pair alignments - O_k = columns (2k, 2k+1) and
// L_k = columns (2k-1, 2k) relative to its first column - so every neighbour the suppression needs is
// already a register pair (left = L_k, right = L_k+1): the arithmetic sits on the FMA-side pipe, which has
// room, instead of more PRMTs on the ALU-side pipe, which has none.
#pragma once
#include <cuda_fp16.h>

#include "psd_common.cuh"

namespace psd {
namespace cp {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t hadd(uint32_t a, uint32_t b) { return u32(__hadd2(h2(a), h2(b))); }
__device__ __forceinline__ uint32_t hsub(uint32_t a, uint32_t b) { return u32(__hsub2(h2(a), h2(b))); }
__device__ __forceinline__ uint32_t hmul(uint32_t a, uint32_t b) { return u32(__hmul2(h2(a), h2(b))); }
__device__ __forceinline__ uint32_t hfma(uint32_t a, uint32_t b, uint32_t c) {
    return u32(__hfma2(h2(a), h2(b), h2(c)));
}
__device__ __forceinline__ uint32_t habs_sum(uint32_t a, uint32_t b) {
    return u32(__hadd2(__habs2(h2(a)), __habs2(h2(b))));
}
__device__ __forceinline__ uint32_t hgt_mask(uint32_t a, uint32_t b) { return __hgt2_mask(h2(a), h2(b)); }
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t bitsel(uint32_t mask, uint32_t a, uint32_t b) { return (a & mask) | (b & ~mask); }
// n * 2^-19 in both lanes
__device__ __forceinline__ uint32_t scaled2(int n) { return u32(__float2half2_rn(ldexpf((float)n, -19))); }

constexpr uint32_t kOne = 0x00200020u;     // 1 * 2^-19 in both lanes (sub-normal 32 * 2^-24)
constexpr uint32_t kTwo = 0x40004000u;     // 2.0 in both lanes
constexpr uint32_t kBias = 0x19191919u;    // the constant byte PRMT puts above every pixel byte
constexpr float kTg22 = 13573.0f / 32768.0f;              // OpenCV's TG22 in 2^15 fixed point, exact in FP32
constexpr float kTg67 = (13573.0f + 65536.0f) / 32768.0f; // tan(67.5) as OpenCV forms it: tg22 + 2

// horizontal sums of one row: c (scaled) and h (scaled, +1024) in both alignments
struct Sums {
    uint32_t cO[4], cL[5], hO[4], hL[5];
};
// one row of magnitudes as the suppression reads them
struct Mags {
    uint32_t mO[4], mL[5], pO[4], pL[5];
    uint32_t dlo[4], dhi[4];   // sector of the O pixels as lane masks: (hi, lo) = 00 left/right, 01 up/down,
                               // 10 the (y-1,x-1)/(y+1,x+1) diagonal, 11 the (y-1,x+1)/(y+1,x-1) diagonal
    bool any;                  // some O pixel is above the low threshold
};

// pixel pairs of one row window (16 bytes x0-4 .. x0+11): VO[k+1] = columns (2k, 2k+1), k = -1..4;
// VL[k] = columns (2k-1, 2k), k = 0..4 - relative to x0, each lane 0x1900 + byte
__device__ __forceinline__ void expand(const uint32_t (&w)[4], uint32_t (&VO)[6], uint32_t (&VL)[5]) {
    VO[0] = prmt(w[0], kBias, 0x4342);
    VO[1] = prmt(w[1], kBias, 0x4140);
    VO[2] = prmt(w[1], kBias, 0x4342);
    VO[3] = prmt(w[2], kBias, 0x4140);
    VO[4] = prmt(w[2], kBias, 0x4342);
    VO[5] = prmt(w[3], kBias, 0x4140);
    const uint32_t s0 = __funnelshift_r(w[0], w[1], 8);   // window bytes 1..4
    const uint32_t s1 = __funnelshift_r(w[1], w[2], 8);   // 5..8
    const uint32_t s2 = __funnelshift_r(w[2], w[3], 8);   // 9..12
    VL[0] = prmt(s0, kBias, 0x4342);
    VL[1] = prmt(s1, kBias, 0x4140);
    VL[2] = prmt(s1, kBias, 0x4342);
    VL[3] = prmt(s2, kBias, 0x4140);
    VL[4] = prmt(s2, kBias, 0x4342);
}

__device__ __forceinline__ void row_sums(const uint32_t (&w)[4], Sums& s) {
    uint32_t VO[6], VL[5];
    expand(w, VO, VL);
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // O_k: left = VL[k], centre = VO[k+1], right = VL[k+1]
        s.cO[k] = hsub(VL[k + 1], VL[k]);
        s.hO[k] = hmul(VO[k + 1] * 2u + VL[k] + VL[k + 1], kOne);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {   // L_k: left = VO[k], centre = VL[k], right = VO[k+1]
        s.cL[k] = hsub(VO[k + 1], VO[k]);
        s.hL[k] = hmul(VL[k] * 2u + VO[k] + VO[k + 1], kOne);
    }
}

// sector masks of one O pair from its gradient pairs
__device__ __forceinline__ void sector(uint32_t gx, uint32_t gy, uint32_t& dlo, uint32_t& dhi) {
    const float ax0 = fabsf(__low2float(h2(gx))), ax1 = fabsf(__high2float(h2(gx)));
    const float ay0 = fabsf(__low2float(h2(gy))), ay1 = fabsf(__high2float(h2(gy)));
    const float a0 = fmaf(ax0, kTg22, -ay0), a1 = fmaf(ax1, kTg22, -ay1);   // < 0: not horizontal
    const float b0 = fmaf(ax0, kTg67, -ay0), b1 = fmaf(ax1, kTg67, -ay1);   // < 0: vertical
    const uint32_t s22 = prmt(__float_as_uint(a0), __float_as_uint(a1), 0xFFBB);   // sign -> whole lane
    const uint32_t s67 = prmt(__float_as_uint(b0), __float_as_uint(b1), 0xFFBB);
    const uint32_t x = gx ^ gy;
    const uint32_t sxy = prmt(x, x, 0xBB99);                                       // signs differ
    dhi = s22 & ~s67;                 // diagonal
    dlo = s67 | (s22 & sxy);          // vertical, or the anti-diagonal
}

}  // namespace cp
}  // namespace psd
