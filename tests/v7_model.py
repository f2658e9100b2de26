"""numpy restatement, instruction by instruction, of the pixel-pair HSV arithmetic in
pyscenedetect_b200/csrc/hsv_half2.cuh (engine variants 7 and 8).

It exists so the bit manipulations of that kernel (PRMT selectors, half2 magic lanes, lane masks,
IDP2A row offsets, the packed hue fix) can be pinned against the oracle on the CPU, over all 2^24
colours, before any GPU time is spent; the GPU test `test_hsv_and_y_exhaustive_2_24[7|8]` pins the
real instructions.  Test infrastructure only.
"""

from __future__ import annotations

import numpy as np

U32 = np.uint32


def prmt(a, b, sel):
    """PRMT in default mode: result byte i = byte ((sel >> 4 i) & 7) of the pool {a: 0..3, b: 4..7}."""
    a = np.asarray(a, dtype=U32)
    b = np.broadcast_to(np.asarray(b, dtype=U32), a.shape)
    pool = [(a >> U32(8 * i)) & U32(0xFF) for i in range(4)] + [(b >> U32(8 * i)) & U32(0xFF) for i in range(4)]
    out = np.zeros_like(a)
    for i in range(4):
        k = (sel >> (4 * i)) & 0xF
        assert k < 8, "sign-replicate mode is not used"
        out |= pool[k] << U32(8 * i)
    return out


def lanes(x):
    x = np.asarray(x, dtype=U32)
    return (x & U32(0xFFFF)).astype(np.int64), (x >> U32(16)).astype(np.int64)


def from_lanes(lo, hi):
    return (np.asarray(lo, dtype=np.int64) & 0xFFFF).astype(U32) | ((np.asarray(hi, dtype=np.int64) & 0xFFFF).astype(U32) << U32(16))


def half_bits_to_int(bits):
    """value of an fp16 bit pattern that is known to hold an integer (asserted)"""
    v = bits.astype(np.uint16).view(np.float16).astype(np.float64)
    assert np.all(v == np.round(v))
    return v.astype(np.int64)


def int_to_half_bits(v):
    assert np.all(np.abs(v) <= 2048), "not exactly representable with ulp 1 in fp16"
    return v.astype(np.float16).view(np.uint16).astype(np.int64)


def half2_op(fn, *ops):
    """lane-wise exact integer op on half2 registers holding integers"""
    lo = fn(*[half_bits_to_int(lanes(o)[0]) for o in ops])
    hi = fn(*[half_bits_to_int(lanes(o)[1]) for o in ops])
    return from_lanes(int_to_half_bits(lo), int_to_half_bits(hi))


def heq2_mask(a, b):
    al, ah = lanes(a)
    bl, bh = lanes(b)
    return from_lanes(np.where(al == bl, 0xFFFF, 0), np.where(ah == bh, 0xFFFF, 0))


def bitsel(m, x, y):
    return (m & x) | (~m & y)


def vimnmx3_u16x2(a, b, c, is_max):
    f = np.maximum if is_max else np.minimum
    (al, ah), (bl, bh), (cl, ch) = lanes(a), lanes(b), lanes(c)
    return from_lanes(f(f(al, bl), cl), f(f(ah, bh), ch))


def tables():
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, dtype=np.int64)
    hdiv = np.zeros(256, dtype=np.int64)
    sdiv[1:] = np.rint(1044480.0 / i).astype(np.int64)
    hdiv[1:] = np.rint(737280.0 / (6.0 * i)).astype(np.int64)
    return sdiv, hdiv


def pair(Bh, Gh, Rh, lane, sdiv, hdiv):
    Vh = vimnmx3_u16x2(Bh, Gh, Rh, True)
    mh = vimnmx3_u16x2(Bh, Gh, Rh, False)
    dh = half2_op(lambda v, m: v - m, Vh, mh)
    di = (Vh - mh).astype(U32)  # plain 32-bit subtraction
    hR = half2_op(lambda g, b: g - b, Gh, Bh)
    hG = half2_op(lambda d, b, r: 2 * d + (b - r), dh, Bh, Rh)
    hB = half2_op(lambda d, r, g: 4 * d + (r - g), dh, Rh, Gh)
    eR = heq2_mask(Vh, Rh)
    eG = heq2_mask(Vh, Gh)
    hh = bitsel(eR, hR, bitsel(eG, hG, hB))
    d0, d1 = [half_bits_to_int(x) for x in lanes(dh)]
    h0, h1 = [half_bits_to_int(x) for x in lanes(hh)]
    # PSD_V7_ADDR 0: IDP2A.LO(x, coef, c) = x.h0 * coef.b0 + x.h1 * coef.b1 + c (mod 2^32); the LUT base
    # is taken as 0: offsets must equal row * 128 + lane * 4 (sdiv) and 32768 + row * 128 + lane * 4
    cs = (lane * 4 - 0x6400 * 128) & 0xFFFFFFFF
    ch = 32768 + lane * 4
    vl, vhh = lanes(Vh)
    dl, dhh = lanes(di)
    aS0 = (vl * 128 + cs) & 0xFFFFFFFF
    aS1 = (vhh * 128 + cs) & 0xFFFFFFFF
    aH0 = (dl * 128 + ch) & 0xFFFFFFFF
    aH1 = (dhh * 128 + ch) & 0xFFFFFFFF
    outs = []
    for aS, aH, d, h in ((aS0, aH0, d0, h0), (aS1, aH1, d1, h1)):
        rowS = (aS - lane * 4) // 128
        rowH = (aH - 32768 - lane * 4) // 128
        assert np.all((aS - lane * 4) % 128 == 0) and np.all((rowS >= 0) & (rowS < 256))
        assert np.all((aH - 32768 - lane * 4) % 128 == 0) and np.all((rowH >= 0) & (rowH < 256))
        # fma.rz(d, sdiv/4096, 32768.5) and fma.rm(h, hdiv/4096, 49152.5): exact sum, then one rounding
        # toward -inf (both sums are positive) to the ulp 2^-8 of [2^15, 2^16)
        xs = d * sdiv[rowS] + 134219776  # * 4096
        xh = h * hdiv[rowH] + 201328640
        assert np.all((xs >> 12 >= 32768) & (xs >> 12 < 65536)) and np.all((xh >> 12 >= 32768) & (xh >> 12 < 65536))
        ys = ((xs >> 4) - (1 << 23)).astype(np.int64) | 0x47000000  # mantissa | exponent of 2^15
        yh = ((xh >> 4) - (1 << 23)).astype(np.int64) | 0x47000000
        outs.append((yh.astype(U32), ys.astype(U32)))
    return outs[0], outs[1], Vh


def fix_hue4(hw):
    m = hw & (hw + hw).astype(U32) & U32(0x80808080)
    return (hw - (m >> U32(7)) * U32(76)).astype(U32)


def hsv16_v7(w, lane=0):
    """w: uint32 array [n_groups, 12] (16 BGR pixels per group).  Returns H, S, V planes as packed words
    [n_groups, 4] exactly as hsv16_v7 in the kernel produces them."""
    w = np.asarray(w, dtype=U32)
    sdiv, hdiv = tables()
    K = U32(0x64646464)
    oh = np.zeros((w.shape[0], 4), dtype=U32)
    os_ = np.zeros_like(oh)
    ov = np.zeros_like(oh)
    for g in range(4):
        wa, wb, wc = w[:, 3 * g], w[:, 3 * g + 1], w[:, 3 * g + 2]
        B01 = prmt(wa, K, 0x4340)
        t01 = prmt(wa, wb, 0x5421)
        G01 = (t01 & U32(0x00FF00FF)) | U32(0x64006400)
        R01 = prmt(t01, K, 0x4341)
        R23 = prmt(wc, K, 0x4340)
        t23 = prmt(wb, wc, 0x6532)
        B23 = (t23 & U32(0x00FF00FF)) | U32(0x64006400)
        G23 = prmt(t23, K, 0x4341)
        (p_h0, p_s0), (p_h1, p_s1), p_v = pair(B01, G01, R01, lane, sdiv, hdiv)
        (q_h0, q_s0), (q_h1, q_s1), q_v = pair(B23, G23, R23, lane, sdiv, hdiv)
        hw = prmt(prmt(p_h0, p_h1, 0x0051), prmt(q_h0, q_h1, 0x0051), 0x5410)
        oh[:, g] = fix_hue4(hw)
        os_[:, g] = prmt(prmt(p_s0, p_s1, 0x0051), prmt(q_s0, q_s1, 0x0051), 0x5410)
        ov[:, g] = prmt(p_v, q_v, 0x6420)
    return oh, os_, ov


def planes_from_bgr(bgr):
    """bgr: uint8 [n, 3] with n a multiple of 16 -> (H, S, V) uint8 [n] through the v7 model"""
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    assert bgr.shape[0] % 16 == 0
    w = bgr.reshape(-1).view(U32).reshape(-1, 12)
    oh, os_, ov = hsv16_v7(w, lane=int(bgr.shape[0] // 16) % 32)
    unpack = lambda x: np.ascontiguousarray(x).view(np.uint8).reshape(-1)
    return unpack(oh), unpack(os_), unpack(ov)
