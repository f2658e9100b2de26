// Per-pixel colour arithmetic shared by the fused score kernel and the test hook.
//
// Exact restatement of OpenCV's 8-bit BGR->HSV (H in [0,180)) and BGR->YUV-Y, the arithmetic
// behind content_detector.py:155 and histogram_detector.py:156 (see oracle/intmath.py, which is
// pinned against cv2 over all 2^24 colours):
//   V = max(B,G,R), d = V - min(B,G,R)
//   S = (d * sdiv[V] + 2048) >> 12,           sdiv[i] = rint((255<<12) / i),     sdiv[0] = 0
//   h = (V==R) ? G-B : (V==G) ? B-R+2d : R-G+4d
//   H = (h * hdiv[d] + 2048) >> 12 (arithmetic), hdiv[i] = rint((180<<12) / (6 i)), hdiv[0] = 0
//   H += 180 if H < 0
//   Y = (R*4899 + G*9617 + B*1868 + 8192) >> 14
#pragma once

#include <stdint.h>

namespace psd {

struct Px16 {  // 16 pixels, planar, 4 pixels per 32-bit word (pixel 4j+i in byte i of word j)
    uint32_t h[4], s[4], v[4];
};

// ---- variant 0: scalar integer arithmetic with the two 256-entry tables in shared memory ----
__device__ __forceinline__ void hsv_px_lut(uint32_t b, uint32_t g, uint32_t r, const int32_t* sdiv,
                                           const int32_t* hdiv, uint32_t& H, uint32_t& S,
                                           uint32_t& V) {
    const int32_t v = (int32_t)max(max(b, g), r);
    const int32_t mn = (int32_t)min(min(b, g), r);
    const int32_t d = v - mn;
    const int32_t s = (d * sdiv[v] + 2048) >> 12;
    int32_t hn;
    if (v == (int32_t)r)
        hn = (int32_t)g - (int32_t)b;
    else if (v == (int32_t)g)
        hn = (int32_t)b - (int32_t)r + 2 * d;
    else
        hn = (int32_t)r - (int32_t)g + 4 * d;
    int32_t h = (hn * hdiv[d] + 2048) >> 12;
    if (h < 0) h += 180;
    H = (uint32_t)h;
    S = (uint32_t)s;
    V = (uint32_t)v;
}

// ---- variant 1: table-free.  The table entries are recomputed with one MUFU.RCP each:
//   sdiv[V] = rint(1044480 / V):  q = 1044480 * rcp(V) has relative error <= 2^-23 * (1 + eps)
//   (rcp.approx.f32 is specified to 1 ulp), i.e. absolute error < 0.125 / V * 1.01, while the exact
//   quotient j/V is never closer than 1/(2V) to a rounding boundary k + 1/2 (1044480 = 2^12*255 has
//   no factor 2^13, so 2*1044480/V is never an odd integer).  Hence adding the 1.5*2^23 magic
//   constant inside the FMA rounds to exactly rint(1044480/V).  The same argument holds for
//   hdiv[d] = rint(122880 / d) (122880 = 2^13 * 15).  Verified exhaustively on the device by
//   tests/test_gpu_hsv.py (all 2^24 colours).
//   The products d*sdiv (< 2^28) and h*hdiv (|.| < 2^25) are formed in integer IMADs. ----
__device__ __forceinline__ int32_t rint_div_u8(float numer, int32_t x) {
    // rint(numer / x) for x in [1,255]; returns 0 for x == 0 via the max() (callers have d == 0)
    const float xf = __int2float_rn(max(x, 1));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xf));
    const float t = fmaf(numer, r, 12582912.0f);  // 1.5 * 2^23: low mantissa bits = rint(product)
    return __float_as_int(t) - 0x4B400000;
}

__device__ __forceinline__ void hsv_px_rcp(uint32_t b, uint32_t g, uint32_t r, uint32_t& H,
                                           uint32_t& S, uint32_t& V) {
    const int32_t v = (int32_t)max(max(b, g), r);
    const int32_t mn = (int32_t)min(min(b, g), r);
    const int32_t d = v - mn;
    const int32_t sd = rint_div_u8(1044480.0f, v);
    const int32_t hd = rint_div_u8(122880.0f, d);
    const int32_t s = (d * sd + 2048) >> 12;
    int32_t hn;
    if (v == (int32_t)r)
        hn = (int32_t)g - (int32_t)b;
    else if (v == (int32_t)g)
        hn = (int32_t)b - (int32_t)r + 2 * d;
    else
        hn = (int32_t)r - (int32_t)g + 4 * d;
    int32_t h = (hn * hd + 2048) >> 12;
    if (h < 0) h += 180;
    H = (uint32_t)h;
    S = (uint32_t)s;
    V = (uint32_t)v;
}

// ---- variant 2: float-domain pipeline on pixel PAIRS with packed FFMA2/FADD2 (sm_100 f32x2) ----
// Bytes are lifted to "magic" floats 2^23 + b by one PRMT each (mantissa ulp = 1, so integer
// add/sub/compare on them is exact and order-preserving).  The table values are regenerated as in
// variant 1; the two fixed-point products use directed rounding so that only bits below the
// 4096 quantum are dropped before the >> 12:
//   S:  x = fma.rz(d, sdiv, 2048)  (x >= 0: truncation never crosses a multiple of 4096)
//       y = fma.rz(x, 2^-12, 2^23)            -> mantissa = floor(x / 4096) = S
//   H:  x = fma.rm(h, hdiv, 2048)  (h may be negative: round toward -inf == arithmetic shift)
//       y = fma.rm(x, 2^-12, 1.5 * 2^23)      -> mantissa = 2^22 + floor(x / 4096)
// |d * sdiv| < 2^28 and |h * hdiv| < 2^25, so the dropped bits are at most 2^4 resp. 2^1 wide and
// 4096 k is always representable: floor(x/4096) is unchanged.  Pinned by the exhaustive test.
typedef unsigned long long f32x2_t;

__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2_rn(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2_t fma2_rz(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2_t fma2_rm(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// byte j of word w -> float 2^23 + byte  (PRMT with the constant 0x4B000000 as second source)
template <int J>
__device__ __forceinline__ float magic_byte(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | J));
}

struct PairOut {  // H, S, V of two pixels as magic-float bit patterns (value in the low byte)
    uint32_t h0, h1, s0, s1, v0, v1;
};

// bytes of pixels (P, P+1) inside the 12 packed words: pixel p occupies bytes 3p, 3p+1, 3p+2
template <int P>
__device__ __forceinline__ void hsv_pair_f32x2(const uint32_t (&w)[12], PairOut& o) {
    constexpr int kb0 = 3 * P, kb1 = 3 * P + 3;
    const float B0 = magic_byte<(kb0 + 0) & 3>(w[(kb0 + 0) >> 2]);
    const float G0 = magic_byte<(kb0 + 1) & 3>(w[(kb0 + 1) >> 2]);
    const float R0 = magic_byte<(kb0 + 2) & 3>(w[(kb0 + 2) >> 2]);
    const float B1 = magic_byte<(kb1 + 0) & 3>(w[(kb1 + 0) >> 2]);
    const float G1 = magic_byte<(kb1 + 1) & 3>(w[(kb1 + 1) >> 2]);
    const float R1 = magic_byte<(kb1 + 2) & 3>(w[(kb1 + 2) >> 2]);
    // max / min on the bit patterns (order-preserving for these positive floats)
    const uint32_t V0 = __vimax3_u32(__float_as_uint(B0), __float_as_uint(G0), __float_as_uint(R0));
    const uint32_t V1 = __vimax3_u32(__float_as_uint(B1), __float_as_uint(G1), __float_as_uint(R1));
    const uint32_t m0 = __vimin3_u32(__float_as_uint(B0), __float_as_uint(G0), __float_as_uint(R0));
    const uint32_t m1 = __vimin3_u32(__float_as_uint(B1), __float_as_uint(G1), __float_as_uint(R1));
    const f32x2_t Vm = pack2(__uint_as_float(V0), __uint_as_float(V1));
    const f32x2_t mn = pack2(__uint_as_float(m0), __uint_as_float(m1));
    const f32x2_t d2 = sub2(Vm, mn);  // plain floats 0..255 (exact)
    float d0, d1;
    unpack2(d2, d0, d1);
    // reciprocals of max(V,1), max(d,1): d == 0 whenever V == 0, and h == 0 whenever d == 0, so the
    // clamped entries are multiplied by zero exactly as the tables' zero entries would be
    const f32x2_t M23 = pack2(8388608.0f, 8388608.0f);
    const f32x2_t M15 = pack2(12582912.0f, 12582912.0f);
    const f32x2_t Vg = pack2(__uint_as_float(max(V0, 0x4B000001u)), __uint_as_float(max(V1, 0x4B000001u)));
    float Vp0, Vp1;
    unpack2(sub2(Vg, M23), Vp0, Vp1);
    const f32x2_t rV = pack2(rcp_approx(Vp0), rcp_approx(Vp1));
    const f32x2_t rd = pack2(rcp_approx(fmaxf(d0, 1.0f)), rcp_approx(fmaxf(d1, 1.0f)));
    const f32x2_t sdiv = sub2(fma2_rn(pack2(1044480.0f, 1044480.0f), rV, M15), M15);
    const f32x2_t hdiv = sub2(fma2_rn(pack2(122880.0f, 122880.0f), rd, M15), M15);
    const f32x2_t c2048 = pack2(2048.0f, 2048.0f);
    const f32x2_t cshift = pack2(0.000244140625f, 0.000244140625f);  // 2^-12
    const f32x2_t ys = fma2_rz(fma2_rz(d2, sdiv, c2048), cshift, M23);
    // hue numerators (differences of magic floats are exact small integers)
    const f32x2_t B2 = pack2(B0, B1), G2 = pack2(G0, G1), R2 = pack2(R0, R1);
    const f32x2_t hR = sub2(G2, B2);
    const f32x2_t hG = fma2_rn(d2, pack2(2.0f, 2.0f), sub2(B2, R2));
    const f32x2_t hB = fma2_rn(d2, pack2(4.0f, 4.0f), sub2(R2, G2));
    float hR0, hR1, hG0, hG1, hB0, hB1;
    unpack2(hR, hR0, hR1);
    unpack2(hG, hG0, hG1);
    unpack2(hB, hB0, hB1);
    const float h0 = (V0 == __float_as_uint(R0)) ? hR0 : ((V0 == __float_as_uint(G0)) ? hG0 : hB0);
    const float h1 = (V1 == __float_as_uint(R1)) ? hR1 : ((V1 == __float_as_uint(G1)) ? hG1 : hB1);
    float yh0, yh1;
    unpack2(fma2_rm(fma2_rm(pack2(h0, h1), hdiv, c2048), cshift, M15), yh0, yh1);
    if (yh0 < 12582912.0f) yh0 += 180.0f;
    if (yh1 < 12582912.0f) yh1 += 180.0f;
    float ys0, ys1;
    unpack2(ys, ys0, ys1);
    o.h0 = __float_as_uint(yh0); o.h1 = __float_as_uint(yh1);
    o.s0 = __float_as_uint(ys0); o.s1 = __float_as_uint(ys1);
    o.v0 = V0; o.v1 = V1;
}

// four magic words (value in the low byte) -> one planar word; the two IMADs run on the FMA pipe
__device__ __forceinline__ uint32_t pack4_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t lo = b * 256u + a;  // low 16 bits = [a.b0, b.b0] (a's upper bytes only reach bytes 2,3)
    const uint32_t hi = d * 256u + c;
    return __byte_perm(lo, hi, 0x5410);
}

__device__ __forceinline__ void hsv16_f32x2(const uint32_t (&w)[12], Px16& o) {
    PairOut p[8];
    hsv_pair_f32x2<0>(w, p[0]);
    hsv_pair_f32x2<2>(w, p[1]);
    hsv_pair_f32x2<4>(w, p[2]);
    hsv_pair_f32x2<6>(w, p[3]);
    hsv_pair_f32x2<8>(w, p[4]);
    hsv_pair_f32x2<10>(w, p[5]);
    hsv_pair_f32x2<12>(w, p[6]);
    hsv_pair_f32x2<14>(w, p[7]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.h[j] = pack4_low_bytes(p[2 * j].h0, p[2 * j].h1, p[2 * j + 1].h0, p[2 * j + 1].h1);
        o.s[j] = pack4_low_bytes(p[2 * j].s0, p[2 * j].s1, p[2 * j + 1].s0, p[2 * j + 1].s1);
        o.v[j] = pack4_low_bytes(p[2 * j].v0, p[2 * j].v1, p[2 * j + 1].v0, p[2 * j + 1].v1);
    }
}

// ---- variant 3: variant 2 re-balanced for the measured B200 pipes (profiles/r01_pipes.md):
// the ALU pipe (PRMT/VIMNMX/ISETP/SEL/FMNMX) issues at half the rate of the FMA pipe, so every
// step that can run on the FMA pipe is moved there:
//   * divide-by-zero guards: rcp(x + 2^-24) instead of rcp(max(x,1)); for x >= 1 the sum rounds
//     back to x, for x == 0 the huge-but-finite reciprocal is multiplied by d == 0 / h == 0;
//   * hue sector: instead of two compares and two selects, every candidate gets the penalty
//     4096 * (V - channel) (zero only for a channel that attains the max) and one 3-input
//     minimum picks the winner; ties between channels give the same H as OpenCV's R > G > B
//     priority for all 2^24 colours (checked exhaustively on the CPU and on the device);
//   * H < 0 -> H + 180 with a saturating add: m = sat(1.5*2^23 - y) is 1 exactly when the
//     integer H is negative, then y += 180 m. ----
__device__ __forceinline__ float fadd_sat(float a, float b) {
    float r;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int P>
__device__ __forceinline__ void hsv_pair_v3(const uint32_t (&w)[12], PairOut& o) {
    constexpr int kb0 = 3 * P, kb1 = 3 * P + 3;
    const float B0 = magic_byte<(kb0 + 0) & 3>(w[(kb0 + 0) >> 2]);
    const float G0 = magic_byte<(kb0 + 1) & 3>(w[(kb0 + 1) >> 2]);
    const float R0 = magic_byte<(kb0 + 2) & 3>(w[(kb0 + 2) >> 2]);
    const float B1 = magic_byte<(kb1 + 0) & 3>(w[(kb1 + 0) >> 2]);
    const float G1 = magic_byte<(kb1 + 1) & 3>(w[(kb1 + 1) >> 2]);
    const float R1 = magic_byte<(kb1 + 2) & 3>(w[(kb1 + 2) >> 2]);
    const uint32_t V0 = __vimax3_u32(__float_as_uint(B0), __float_as_uint(G0), __float_as_uint(R0));
    const uint32_t V1 = __vimax3_u32(__float_as_uint(B1), __float_as_uint(G1), __float_as_uint(R1));
    const uint32_t m0 = __vimin3_u32(__float_as_uint(B0), __float_as_uint(G0), __float_as_uint(R0));
    const uint32_t m1 = __vimin3_u32(__float_as_uint(B1), __float_as_uint(G1), __float_as_uint(R1));
    const f32x2_t Vm = pack2(__uint_as_float(V0), __uint_as_float(V1));
    const f32x2_t mn = pack2(__uint_as_float(m0), __uint_as_float(m1));
    const f32x2_t M23 = pack2(8388608.0f, 8388608.0f);
    const f32x2_t M15 = pack2(12582912.0f, 12582912.0f);
    const f32x2_t eps = pack2(5.9604644775390625e-8f, 5.9604644775390625e-8f);  // 2^-24
    const f32x2_t d2 = sub2(Vm, mn);
    float dg0, dg1, Vg0, Vg1;
    unpack2(add2(d2, eps), dg0, dg1);
    unpack2(add2(sub2(Vm, M23), eps), Vg0, Vg1);
    const f32x2_t rV = pack2(rcp_approx(Vg0), rcp_approx(Vg1));
    const f32x2_t rd = pack2(rcp_approx(dg0), rcp_approx(dg1));
    const f32x2_t sdiv = sub2(fma2_rn(pack2(1044480.0f, 1044480.0f), rV, M15), M15);
    const f32x2_t hdiv = sub2(fma2_rn(pack2(122880.0f, 122880.0f), rd, M15), M15);
    const f32x2_t c2048 = pack2(2048.0f, 2048.0f);
    const f32x2_t cshift = pack2(0.000244140625f, 0.000244140625f);  // 2^-12
    const f32x2_t ys = fma2_rz(fma2_rz(d2, sdiv, c2048), cshift, M23);
    const f32x2_t B2 = pack2(B0, B1), G2 = pack2(G0, G1), R2 = pack2(R0, R1);
    const f32x2_t K = pack2(4096.0f, 4096.0f);
    // candidate + 4096 * (V - channel)
    const f32x2_t hR = fma2_rn(sub2(Vm, R2), K, sub2(G2, B2));
    const f32x2_t hG = fma2_rn(sub2(Vm, G2), K, fma2_rn(d2, pack2(2.0f, 2.0f), sub2(B2, R2)));
    const f32x2_t hB = fma2_rn(sub2(Vm, B2), K, fma2_rn(d2, pack2(4.0f, 4.0f), sub2(R2, G2)));
    float hR0, hR1, hG0, hG1, hB0, hB1;
    unpack2(hR, hR0, hR1);
    unpack2(hG, hG0, hG1);
    unpack2(hB, hB0, hB1);
    const f32x2_t h2 = pack2(fmin3(hR0, hG0, hB0), fmin3(hR1, hG1, hB1));
    float yh0, yh1;
    unpack2(fma2_rm(fma2_rm(h2, hdiv, c2048), cshift, M15), yh0, yh1);
    yh0 = fmaf(fadd_sat(12582912.0f, -yh0), 180.0f, yh0);
    yh1 = fmaf(fadd_sat(12582912.0f, -yh1), 180.0f, yh1);
    float ys0, ys1;
    unpack2(ys, ys0, ys1);
    o.h0 = __float_as_uint(yh0); o.h1 = __float_as_uint(yh1);
    o.s0 = __float_as_uint(ys0); o.s1 = __float_as_uint(ys1);
    o.v0 = V0; o.v1 = V1;
}

__device__ __forceinline__ void hsv16_v3(const uint32_t (&w)[12], Px16& o) {
    PairOut p[8];
    hsv_pair_v3<0>(w, p[0]);
    hsv_pair_v3<2>(w, p[1]);
    hsv_pair_v3<4>(w, p[2]);
    hsv_pair_v3<6>(w, p[3]);
    hsv_pair_v3<8>(w, p[4]);
    hsv_pair_v3<10>(w, p[5]);
    hsv_pair_v3<12>(w, p[6]);
    hsv_pair_v3<14>(w, p[7]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.h[j] = pack4_low_bytes(p[2 * j].h0, p[2 * j].h1, p[2 * j + 1].h0, p[2 * j + 1].h1);
        o.s[j] = pack4_low_bytes(p[2 * j].s0, p[2 * j].s1, p[2 * j + 1].s0, p[2 * j + 1].s1);
        o.v[j] = pack4_low_bytes(p[2 * j].v0, p[2 * j].v1, p[2 * j + 1].v0, p[2 * j + 1].v1);
    }
}

// ---- variant 4: fewest lane-operations (the measured B200 model is "one 32-lane operation per
// clock per SM sub-partition", with IMAD/IDP4A confined to one 16-lane half and
// PRMT/LOP3/SHF/VIMNMX to the other - profiles/r01_pipes_microbench.txt).
//   * the two tables come back as a conflict-free shared-memory LUT, but pre-divided by 4096 and
//     stored as float, replicated per lane (row i = 32 x sdiv[i]/4096 | 32 x hdiv[i]/4096, 256 B),
//     so a lookup is one IMAD (address = value * 256 + lane offset; the 2^23 magic exponent
//     overflows out of the 32-bit product) and one LDS;
//   * S and H are then ONE directed-rounding FMA each: with the magic constant 2^15 (ulp 2^-8) the
//     +0.5 of the fixed-point rounding is representable, and the integer part of
//     d*sdiv/4096 + 0.5 lands byte-aligned in bits 8..15 of the result:
//       yS = fma.rz(d, sdiv/4096, 32768.5)          -> byte 1 = S
//       yH = fma.rm(h, hdiv/4096, 49152.5)          -> byte 1 = H mod 256, H < 0 <=> yH < 49152
//   * bytes are lifted to magic floats with IDP4A (A half) instead of PRMT (B half). ----
struct LutView {
    uint32_t s_addr;  // shared-window byte address of this lane's sdiv column
    uint32_t h_addr;  // ... of this lane's hdiv column
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fma_rz(float a, float b, float c) {
    float r;
    asm("fma.rz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fma_rm(float a, float b, float c) {
    float r;
    asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
template <int J>
__device__ __forceinline__ float magic_byte_dp4a(uint32_t w) {
    return __uint_as_float(__dp4a(w, 1u << (8 * J), 0x4B000000u));
}

// one pixel: bytes kb, kb+1, kb+2 of the packed words.  Outputs: yS / yH bit patterns (value in
// byte 1) and the magic-float bits of V (value in byte 0).
// The two 16-lane halves of a sub-partition must stay balanced: IDP4A/IMAD only run on one half,
// PRMT/VABSDIFF4 only on the other (everything else uses both).  With the two address IMADs on the
// first half and 3 packing PRMT + SAD per pixel on the second, lifting TWO of the three bytes with
// IDP4A and one with PRMT balances them (profiles/r01b_hsv_rate_all_variants.txt).
#ifndef PSD_V4_PRMT_CHANNELS
#define PSD_V4_PRMT_CHANNELS 1  // how many of B,G,R are extracted with PRMT instead of IDP4A
#endif
template <int KB>
__device__ __forceinline__ void hsv_px_v4(const uint32_t (&w)[12], const LutView& lut, uint32_t& oh,
                                          uint32_t& os, uint32_t& ov) {
    const float B = (PSD_V4_PRMT_CHANNELS >= 3) ? magic_byte<(KB + 0) & 3>(w[(KB + 0) >> 2])
                                                : magic_byte_dp4a<(KB + 0) & 3>(w[(KB + 0) >> 2]);
    const float G = (PSD_V4_PRMT_CHANNELS >= 1) ? magic_byte<(KB + 1) & 3>(w[(KB + 1) >> 2])
                                                : magic_byte_dp4a<(KB + 1) & 3>(w[(KB + 1) >> 2]);
    const float R = (PSD_V4_PRMT_CHANNELS >= 2) ? magic_byte<(KB + 2) & 3>(w[(KB + 2) >> 2])
                                                : magic_byte_dp4a<(KB + 2) & 3>(w[(KB + 2) >> 2]);
    const float V = fmax3(B, G, R);
    const float mn = fmin3(B, G, R);
    const float d = V - mn;  // exact, plain float 0..255
    const uint32_t vbits = __float_as_uint(V);
    // row address = value * 256: the 0x4B exponent byte of the magic float overflows out of the
    // 32-bit product, so (2^23 + V) * 256 == V * 256 and (V - mn) * 256 == V*256 - mn*256 (mod 2^32)
    const uint32_t a_s = vbits * 256u + lut.s_addr;
    const float sdivp = lds_f32(a_s);
    const float yS = fma_rz(d, sdivp, 32768.5f);
    const uint32_t a_h = a_s - __float_as_uint(mn) * 256u;  // one IMAD: (V - mn) * 256 + s_addr
    const float hdivp = lds_f32(a_h + 128u);                  // hdiv column = sdiv column + 128 B
    const float hR = G - B;
    const float hG = fmaf(d, 2.0f, B - R);
    const float hB = fmaf(d, 4.0f, R - G);
    const float h = (V == R) ? hR : ((V == G) ? hG : hB);
    float yH = fma_rm(h, hdivp, 49152.5f);
    yH = fmaf(fma_sat(yH, -256.0f, 12582912.0f), 180.0f, yH);  // += 180 when the integer part is < 0
    oh = __float_as_uint(yH);
    os = __float_as_uint(yS);
    ov = vbits;
}

__device__ __forceinline__ void hsv16_v4(const uint32_t (&w)[12], Px16& o, const LutView& lut) {
    uint32_t h[16], s[16], v[16];
#define PSD_PX(i) hsv_px_v4<3 * (i)>(w, lut, h[i], s[i], v[i]);
    PSD_PX(0) PSD_PX(1) PSD_PX(2) PSD_PX(3) PSD_PX(4) PSD_PX(5) PSD_PX(6) PSD_PX(7)
    PSD_PX(8) PSD_PX(9) PSD_PX(10) PSD_PX(11) PSD_PX(12) PSD_PX(13) PSD_PX(14) PSD_PX(15)
#undef PSD_PX
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // H, S: byte 1 of each result; V: byte 0
        o.h[j] = __byte_perm(__byte_perm(h[4 * j], h[4 * j + 1], 0x0051),
                             __byte_perm(h[4 * j + 2], h[4 * j + 3], 0x0051), 0x5410);
        o.s[j] = __byte_perm(__byte_perm(s[4 * j], s[4 * j + 1], 0x0051),
                             __byte_perm(s[4 * j + 2], s[4 * j + 3], 0x0051), 0x5410);
        o.v[j] = __byte_perm(__byte_perm(v[4 * j], v[4 * j + 1], 0x0040),
                             __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x0040), 0x5410);
    }
}

// ---- variant 4 on pixel PAIRS: same arithmetic, but the float adds/FMAs of two pixels share one
// FADD2/FFMA2 issue slot (the kernel is issue-bound, the FMA pipe has slack). ----
template <int KB>
__device__ __forceinline__ void hsv_pair_v4(const uint32_t (&w)[12], const LutView& lut, uint32_t (&oh)[2],
                                            uint32_t (&os)[2], uint32_t (&ov)[2]) {
    constexpr int K0 = KB, K1 = KB + 3;
    const float B0 = magic_byte_dp4a<(K0 + 0) & 3>(w[(K0 + 0) >> 2]);
    const float G0 = magic_byte<(K0 + 1) & 3>(w[(K0 + 1) >> 2]);
    const float R0 = magic_byte_dp4a<(K0 + 2) & 3>(w[(K0 + 2) >> 2]);
    const float B1 = magic_byte_dp4a<(K1 + 0) & 3>(w[(K1 + 0) >> 2]);
    const float G1 = magic_byte<(K1 + 1) & 3>(w[(K1 + 1) >> 2]);
    const float R1 = magic_byte_dp4a<(K1 + 2) & 3>(w[(K1 + 2) >> 2]);
    const float V0 = fmax3(B0, G0, R0), V1 = fmax3(B1, G1, R1);
    const float m0 = fmin3(B0, G0, R0), m1 = fmin3(B1, G1, R1);
    const f32x2_t B2 = pack2(B0, B1), G2 = pack2(G0, G1), R2 = pack2(R0, R1);
    const f32x2_t d2 = sub2(pack2(V0, V1), pack2(m0, m1));
    const uint32_t as0 = __float_as_uint(V0) * 256u + lut.s_addr;
    const uint32_t as1 = __float_as_uint(V1) * 256u + lut.s_addr;
    const f32x2_t sdivp = pack2(lds_f32(as0), lds_f32(as1));
    const f32x2_t hdivp = pack2(lds_f32(as0 - __float_as_uint(m0) * 256u + 128u),
                                lds_f32(as1 - __float_as_uint(m1) * 256u + 128u));
    const f32x2_t yS = fma2_rz(d2, sdivp, pack2(32768.5f, 32768.5f));
    const f32x2_t hR = sub2(G2, B2);
    const f32x2_t hG = fma2_rn(d2, pack2(2.0f, 2.0f), sub2(B2, R2));
    const f32x2_t hB = fma2_rn(d2, pack2(4.0f, 4.0f), sub2(R2, G2));
    float hR0, hR1, hG0, hG1, hB0, hB1;
    unpack2(hR, hR0, hR1);
    unpack2(hG, hG0, hG1);
    unpack2(hB, hB0, hB1);
    const float h0 = (V0 == R0) ? hR0 : ((V0 == G0) ? hG0 : hB0);
    const float h1 = (V1 == R1) ? hR1 : ((V1 == G1) ? hG1 : hB1);
    float yH0, yH1;
    unpack2(fma2_rm(pack2(h0, h1), hdivp, pack2(49152.5f, 49152.5f)), yH0, yH1);
    const f32x2_t mfix = pack2(fma_sat(yH0, -256.0f, 12582912.0f), fma_sat(yH1, -256.0f, 12582912.0f));
    unpack2(fma2_rn(mfix, pack2(180.0f, 180.0f), pack2(yH0, yH1)), yH0, yH1);
    float yS0, yS1;
    unpack2(yS, yS0, yS1);
    oh[0] = __float_as_uint(yH0); oh[1] = __float_as_uint(yH1);
    os[0] = __float_as_uint(yS0); os[1] = __float_as_uint(yS1);
    ov[0] = __float_as_uint(V0); ov[1] = __float_as_uint(V1);
}

__device__ __forceinline__ void hsv16_v4pair(const uint32_t (&w)[12], Px16& o, const LutView& lut) {
    uint32_t h[16], s[16], v[16];
#define PSD_PAIR(i) { uint32_t a[2], b[2], c[2]; hsv_pair_v4<6 * (i)>(w, lut, a, b, c); \
        h[2 * (i)] = a[0]; h[2 * (i) + 1] = a[1]; s[2 * (i)] = b[0]; s[2 * (i) + 1] = b[1]; \
        v[2 * (i)] = c[0]; v[2 * (i) + 1] = c[1]; }
    PSD_PAIR(0) PSD_PAIR(1) PSD_PAIR(2) PSD_PAIR(3) PSD_PAIR(4) PSD_PAIR(5) PSD_PAIR(6) PSD_PAIR(7)
#undef PSD_PAIR
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.h[j] = __byte_perm(__byte_perm(h[4 * j], h[4 * j + 1], 0x0051),
                             __byte_perm(h[4 * j + 2], h[4 * j + 3], 0x0051), 0x5410);
        o.s[j] = __byte_perm(__byte_perm(s[4 * j], s[4 * j + 1], 0x0051),
                             __byte_perm(s[4 * j + 2], s[4 * j + 3], 0x0051), 0x5410);
        o.v[j] = __byte_perm(__byte_perm(v[4 * j], v[4 * j + 1], 0x0040),
                             __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x0040), 0x5410);
    }
}

// fills the replicated LUT (64 KB) - called once per CTA by all threads.  Thread t < 512 computes
// ONE table value (row t>>1, sdiv or hdiv) and stores its 32 per-lane copies.
__device__ __forceinline__ void lut_fill(float* lut, int tid, int nthreads) {
    for (int t = tid; t < 512; t += nthreads) {
        const int row = t >> 1, which = t & 1;
        float v = 0.0f;
        if (row) {
            // exact table integers / 4096 (both are exact in float: < 2^21 and a power-of-two divisor)
            const int q = which ? __double2int_rn(737280.0 / (6.0 * (double)row))
                                : __double2int_rn(1044480.0 / (double)row);
            v = (float)q * 0.000244140625f;
        }
        float4* dst = reinterpret_cast<float4*>(lut + row * 64 + which * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v, v, v, v);
    }
}

__device__ __forceinline__ uint32_t y_px(uint32_t b, uint32_t g, uint32_t r) {
    return (r * 4899u + g * 9617u + b * 1868u + 8192u) >> 14;
}

// Y of pixel P straight from the packed words with IDP4A: the 14-bit coefficients are split into
// (lo, hi) bytes, 1868 = 7*256+76, 9617 = 37*256+145, 4899 = 19*256+35, so
// Y = (dp4a(w, lo) + 256 * dp4a(w, hi) + 8192) >> 14 with per-byte-position coefficient words.
__host__ __device__ constexpr uint32_t y_coef_word(int word, int k0, int which) {
    // which: 0 = lo bytes (76,145,35 for B,G,R), 1 = hi bytes (7,37,19)
    uint32_t r = 0;
    for (int b = 0; b < 4; ++b) {
        const int g = 4 * word + b - k0;  // 0,1,2 -> B,G,R of this pixel
        uint32_t c = 0;
        if (g == 0) c = which ? 7u : 76u;
        if (g == 1) c = which ? 37u : 145u;
        if (g == 2) c = which ? 19u : 35u;
        r |= c << (8 * b);
    }
    return r;
}

template <int P>
__device__ __forceinline__ uint32_t y_of_pixel(const uint32_t (&w)[12]) {
    constexpr int k0 = 3 * P, j0 = k0 >> 2, j1 = (k0 + 2) >> 2;
    uint32_t lo = __dp4a(w[j0], y_coef_word(j0, k0, 0), 8192u);
    uint32_t hi = __dp4a(w[j0], y_coef_word(j0, k0, 1), 0u);
    if (j1 != j0) {
        lo = __dp4a(w[j1], y_coef_word(j1, k0, 0), lo);
        hi = __dp4a(w[j1], y_coef_word(j1, k0, 1), hi);
    }
    return (hi * 256u + lo) >> 14;
}

// byte k (0..47) of 12 packed words
__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[12], int k) {
    return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
}

__device__ __forceinline__ void hsv16_f32x2(const uint32_t (&w)[12], Px16& o);
__device__ __forceinline__ void hsv16_v3(const uint32_t (&w)[12], Px16& o);

template <int VARIANT>
__device__ __forceinline__ void hsv16(const uint32_t (&w)[12], Px16& o, const int32_t* sdiv,
                                      const int32_t* hdiv) {
    if (VARIANT == 2) {
        hsv16_f32x2(w, o);
        return;
    }
    if (VARIANT == 3) {
        hsv16_v3(w, o);
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t hw = 0, sw = 0, vw = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = 4 * j + i;
            const uint32_t b = byte_of(w, 3 * p), g = byte_of(w, 3 * p + 1), r = byte_of(w, 3 * p + 2);
            uint32_t H, S, V;
            if (VARIANT == 0)
                hsv_px_lut(b, g, r, sdiv, hdiv, H, S, V);
            else
                hsv_px_rcp(b, g, r, H, S, V);
            hw |= H << (8 * i);
            sw |= S << (8 * i);
            vw |= V << (8 * i);
        }
        o.h[j] = hw;
        o.s[j] = sw;
        o.v[j] = vw;
    }
}

}  // namespace psd
