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

__device__ __forceinline__ uint32_t y_px(uint32_t b, uint32_t g, uint32_t r) {
    return (r * 4899u + g * 9617u + b * 1868u + 8192u) >> 14;
}

// byte k (0..47) of 12 packed words
__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[12], int k) {
    return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
}

template <int VARIANT>
__device__ __forceinline__ void hsv16(const uint32_t (&w)[12], Px16& o, const int32_t* sdiv,
                                      const int32_t* hdiv) {
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
