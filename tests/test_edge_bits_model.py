"""CPU restatement of the bit-plane edge path (csrc/edge_kernels.cu, csrc/canny_pairs.cuh), instruction-level where
the bit manipulation is the risk:

  * classify, integer form (`classify_bits`): per-strip horizontal sums h = V(i-1)+2V(i)+V(i+1), c = V(i+1)-V(i-1),
    vertical Sobel from them, zeroed outside the image, 2-bit direction sectors, byte-per-8-columns bit planes;
  * classify, the kernel's own arithmetic (`classify_pairs`): binary16 lanes holding n * 2^-19, the PRMT selectors
    that build them, the sign-mask sectors from two FP32 FMAs, `p > max(p(left), m(right), low + 1)` suppression,
    IDP.2A byte packing - must equal the integer form, cv2.Sobel and cv2.Canny;
  * hysteresis: 64x32 tiles, rows as 64-bit words, `run_fill` with the two carry walks, the one-pixel ring
    taken from the neighbouring tiles, rounds until no tile changes.

The result must equal cv2.Canny (through oracle.intmath.canny_thresholds) on images whose sizes exercise
partial strips, partial tiles, a single word per row and long weak chains that cross many tiles."""

import cv2
import numpy as np
import pytest

from oracle import intmath as M

TG22 = 13573
MASK64 = (1 << 64) - 1


def brev64(x):
    return int(f"{x:064b}"[::-1], 2)


def run_fill(t, c):
    up = ((((c + t) & MASK64) ^ c) & c) | t
    cr, tr = brev64(c), brev64(t)
    dn = brev64(((((cr + tr) & MASK64) ^ cr) & cr) | tr)
    return up | dn


def classify_bits(lum, low, high):
    """-> (E, C) bit planes as uint32 arrays [H][Wq], built strip by strip like the kernel."""
    H, W = lum.shape
    Wq = (W + 31) // 32
    E = np.zeros((H, Wq * 4), np.uint8)
    C = np.zeros((H, Wq * 4), np.uint8)
    V = lum.astype(np.int64)

    def sums(y, x0):
        yc = min(max(y, 0), H - 1)
        win = [int(V[yc, min(max(x0 - 4 + k, 0), W - 1)]) for k in range(16)]
        h = [win[2 + i] + 2 * win[3 + i] + win[4 + i] for i in range(10)]
        c = [win[4 + i] - win[2 + i] for i in range(10)]
        return h, c

    def grad(y, x0, sa, sb, sc):
        m, sec = [0] * 10, [0] * 10
        for i in range(10):
            x = x0 - 1 + i
            inside = 0 <= y < H and 0 <= x < W
            gx = sa[1][i] + 2 * sb[1][i] + sc[1][i] if inside else 0
            gy = sc[0][i] - sa[0][i] if inside else 0
            m[i] = abs(gx) + abs(gy)
            ax, ay = abs(gx), abs(gy) << 15
            t22 = ax * TG22
            t67 = t22 + (ax << 16)
            sec[i] = 0 if ay < t22 else 1 if ay > t67 else (3 if (gx ^ gy) < 0 else 2)
        return m, sec

    for x0 in range(0, W, 8):
        rows = {y: sums(y, x0) for y in range(-2, H + 2)}
        g = {y: grad(y, x0, rows[y - 1], rows[y], rows[y + 1]) for y in range(-1, H + 1)}
        for y in range(H):
            (mU, _), (mC, sC), (mD, _) = g[y - 1], g[y], g[y + 1]
            eb = cb = 0
            for i in range(1, 9):
                m = mC[i]
                if m > low and x0 + i - 1 < W:
                    d = sC[i]
                    if d == 0:
                        keep = m > mC[i - 1] and m >= mC[i + 1]
                    elif d == 1:
                        keep = m > mU[i] and m >= mD[i]
                    elif d == 2:
                        keep = m > mU[i - 1] and m > mD[i + 1]
                    else:
                        keep = m > mU[i + 1] and m > mD[i - 1]
                    if keep:
                        cb |= 1 << (i - 1)
                        if m > high:
                            eb |= 1 << (i - 1)
            E[y, x0 // 8] = eb
            C[y, x0 // 8] = cb
    return E.view("<u4").copy(), C.view("<u4").copy()


# ---- the pixel-pair classify kernel (csrc/canny_pairs.cuh): binary16 lanes holding n * 2^-19 ----
SCALE = np.float16(2.0 ** -19)          # the pattern 0x0020


def prmt(a, b, sel):
    """PTX prmt.b32, default mode: nibble bit 3 = replicate the sign of the selected byte."""
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for i in range(4):
        nib = (sel >> (4 * i)) & 0xF
        byte = src[nib & 7]
        if nib & 8:
            byte = 0xFF if byte & 0x80 else 0x00
        out |= byte << (8 * i)
    return out


def funnel_r(lo, hi, sh):
    return (((hi << 32) | lo) >> sh) & 0xFFFFFFFF


def expand_window(w):
    """canny_pairs.cuh:expand on four window words -> (VO[6], VL[5]) pair patterns."""
    K = 0x19191919
    VO = [prmt(w[0], K, 0x4342), prmt(w[1], K, 0x4140), prmt(w[1], K, 0x4342), prmt(w[2], K, 0x4140),
          prmt(w[2], K, 0x4342), prmt(w[3], K, 0x4140)]
    s0, s1, s2 = funnel_r(w[0], w[1], 8), funnel_r(w[1], w[2], 8), funnel_r(w[2], w[3], 8)
    VL = [prmt(s0, K, 0x4342), prmt(s1, K, 0x4140), prmt(s1, K, 0x4342), prmt(s2, K, 0x4140), prmt(s2, K, 0x4342)]
    return VO, VL


def test_pair_expansion_selectors():
    rng = np.random.default_rng(5)
    for _ in range(50):
        by = [int(v) for v in rng.integers(0, 256, 16)]          # window bytes, column u = index - 4
        w = [by[4 * j] | by[4 * j + 1] << 8 | by[4 * j + 2] << 16 | by[4 * j + 3] << 24 for j in range(4)]
        VO, VL = expand_window(w)
        col = lambda u: 0x1900 + by[u + 4]
        for j in range(6):       # VO[j] = columns (2k, 2k+1), k = j - 1
            k = j - 1
            assert VO[j] == col(2 * k) | col(2 * k + 1) << 16
        for k in range(5):       # VL[k] = columns (2k-1, 2k)
            assert VL[k] == col(2 * k - 1) | col(2 * k) << 16
    # sign-replicate selectors of the sector masks
    assert prmt(0x80000000, 0x00000001, 0xFFBB) == 0x0000FFFF
    assert prmt(0x7FFFFFFF, 0xFFFFFFFF, 0xFFBB) == 0xFFFF0000
    assert prmt(0x00008000, 0x00008000, 0xBB99) == 0x0000FFFF
    assert prmt(0x80007FFF, 0x80007FFF, 0xBB99) == 0xFFFF0000
    # the bit planes' byte from the lane masks: sum of 65535 * weight, negated, low byte
    for byte in range(256):
        acc = sum(65535 * (1 << i) for i in range(8) if byte >> i & 1)
        assert (-acc) & 0xFF == byte


def classify_pairs(lum, low, high):
    """Canny classify with the kernel's number representation: every lane a binary16 value n * 2^-19."""
    H, W = lum.shape
    Wq = (W + 31) // 32
    pad = np.pad(lum, 2, mode="edge").astype(np.uint16)                 # BORDER_REPLICATE, rows/cols -2 .. +1
    pat = (np.uint16(0x1900) + pad)                                      # 0x1900 + byte
    val = pat.view(np.float16)                                           # = (1280 + v) * 2^-19
    assert np.array_equal(val.astype(np.float64) * 2.0 ** 19, 1280.0 + pad)
    # horizontal sums for columns -1 .. W (index i -> column i - 1), rows -2 .. H+1
    c = val[:, 2:] - val[:, :-2]                                         # HADD2: exact, the 1280 cancels
    hbits = (pat[:, :-2] + np.uint16(2) * pat[:, 1:-1] + pat[:, 2:]).astype(np.uint16)   # integer lanes
    assert hbits.max() < 0x6800
    hs = hbits.view(np.float16) * SCALE                                  # HMUL2: (1024 + h) * 2^-19
    assert c.dtype == np.float16 and hs.dtype == np.float16
    two = np.float16(2.0)
    gx = (c[1:-1] * two + c[:-2]).astype(np.float16) + c[2:]             # HFMA2 (single rounding of an exact value), HADD2
    gy = hs[2:] - hs[:-2]
    m = np.abs(gx) + np.abs(gy)                                          # rows -1 .. H, columns -1 .. W
    assert m.dtype == np.float16
    rows = np.arange(-1, H + 1)[:, None]
    cols = np.arange(-1, W + 1)[None, :]
    m = np.where((rows >= 0) & (rows < H) & (cols >= 0) & (cols < W), m, np.float16(0))
    p = m + SCALE
    # the integers the lanes stand for
    gi = lambda a: np.rint(a.astype(np.float64) * 2.0 ** 19).astype(np.int64)
    assert np.array_equal(gi(m) * 2.0 ** -19, m.astype(np.float64))
    # sectors from the signs of two FP32 FMAs (exact here in float64: 11 x 17 significant bits)
    ax, ay = np.abs(gx).astype(np.float64), np.abs(gy).astype(np.float64)
    k22 = float(np.float32(13573.0 / 32768.0))
    k67 = float(np.float32((13573.0 + 65536.0) / 32768.0))
    assert k22 * 32768 == 13573 and k67 * 32768 == 79109
    s22 = (ax * k22 - ay) < 0
    s67 = (ax * k67 - ay) < 0
    sxy = np.signbit(gx) != np.signbit(gy)
    dhi = s22 & ~s67
    dlo = s67 | (s22 & sxy)
    # suppression: one comparison per pixel on the unsigned order of the patterns
    bits = lambda a: a.view(np.uint16).astype(np.int64)
    mb, pb = bits(m), bits(p)
    low1 = int(np.float16(np.float32(low + 1) * np.float32(2.0 ** -19)).view(np.uint16))
    high1 = np.float16(np.float32(high + 1) * np.float32(2.0 ** -19))
    C_, U_, D_ = slice(1, -1), slice(0, -2), slice(2, None)            # rows y, y-1, y+1
    X_, Lx, Rx = slice(1, -1), slice(0, -2), slice(2, None)            # columns x, x-1, x+1
    mx3 = lambda a, b: np.maximum(np.maximum(a, b), low1)
    n_h = mx3(pb[C_, Lx], mb[C_, Rx])
    n_v = mx3(pb[U_, X_], mb[D_, X_])
    n_d1 = mx3(pb[U_, Lx], pb[D_, Rx])
    n_d2 = mx3(pb[U_, Rx], pb[D_, Lx])
    hi_, lo_ = dhi[C_, X_], dlo[C_, X_]
    n = np.where(hi_, np.where(lo_, n_d2, n_d1), np.where(lo_, n_v, n_h))
    # HSET2.GT compares the VALUES; for non-negative binary16 that is the order of the patterns
    keep = pb[C_, X_] > n
    strong = keep & (p[C_, X_] > high1)
    pack = lambda k: np.packbits(np.pad(k, ((0, 0), (0, Wq * 32 - W))), axis=1, bitorder="little").view("<u4").copy()
    return pack(strong), pack(keep), gi(gx[C_, X_]), gi(gy[C_, X_])


@pytest.mark.parametrize("shape", [(40, 64), (97, 131), (33, 70), (64, 200)])
def test_pair_lane_classify_equals_integer_classify(shape):
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    hard = rng.integers(0, 2, (h, w), dtype=np.uint8) * 255          # gradients up to +-1020, magnitudes to 2040
    imgs = [hard, rng.integers(0, 256, (h, w), dtype=np.uint8),
            cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (5, 5), 0), _snake(h, w)]
    for lum in imgs:
        for low, high in (M.canny_thresholds(float(np.median(lum))), (0, 255), (20, 120), (255, 255)):
            E, C, gx, gy = classify_pairs(lum, low, high)
            E0, C0 = classify_bits(lum, low, high)
            assert np.array_equal(C, C0) and np.array_equal(E, E0)
        sx = cv2.Sobel(lum, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE)
        sy = cv2.Sobel(lum, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE)
        assert np.array_equal(gx, sx) and np.array_equal(gy, sy)


def hysteresis_bits(E, C, H, W):
    """Rounds over 64x32 tiles exactly as psd_hyst_bits_kernel schedules them; returns (E, rounds)."""
    Wq = E.shape[1]
    E = E.copy()
    tx_n, ty_n = (Wq + 1) // 2, (H + 31) // 32

    def row64(P, y, wq0):
        lo = int(P[y, wq0])
        hi = int(P[y, wq0 + 1]) if wq0 + 1 < Wq else 0
        return lo | (hi << 32)

    def side(y, wq0):
        left = (int(E[y, wq0 - 1]) >> 31) if wq0 > 0 else 0
        right = (int(E[y, wq0 + 2]) & 1) if wq0 + 2 < Wq else 0
        return left, right

    dirty = {(ty, tx) for ty in range(ty_n) for tx in range(tx_n)}
    rounds = 0
    while dirty:
        rounds += 1
        nxt = set()
        for ty, tx in sorted(dirty):
            y0, wq0 = ty * 32, tx * 2
            c = [row64(C, y0 + r, wq0) if y0 + r < H else 0 for r in range(32)]
            e = [row64(E, y0 + r, wq0) if y0 + r < H else 0 for r in range(32)]
            if not any(cc & ~ee for cc, ee in zip(c, e)):
                continue
            lr = [side(y0 + r, wq0) if y0 + r < H else (0, 0) for r in range(32)]
            above = (row64(E, y0 - 1, wq0), *side(y0 - 1, wq0)) if y0 > 0 else (0, 0, 0)
            below = (row64(E, y0 + 32, wq0), *side(y0 + 32, wq0)) if y0 + 32 < H else (0, 0, 0)
            seeds = []
            for r in range(32):
                lu = above[1] if r == 0 else lr[r - 1][0]
                ld = below[1] if r == 31 else lr[r + 1][0]
                ru = above[2] if r == 0 else lr[r - 1][1]
                rd = below[2] if r == 31 else lr[r + 1][1]
                seeds.append((1 if (lu | lr[r][0] | ld) else 0) | ((1 << 63) if (ru | lr[r][1] | rd) else 0))
            e_in = list(e)
            while True:
                new = []
                for r in range(32):
                    u = above[0] if r == 0 else e[r - 1]
                    d = below[0] if r == 31 else e[r + 1]
                    v = u | d
                    nb = (v | (v << 1) | (v >> 1) | seeds[r]) & MASK64
                    new.append(run_fill((nb & c[r]) | e[r], c[r]))
                changed = new != e
                e = new
                if not changed:
                    break
            if e != e_in:
                for r in range(32):
                    if y0 + r < H:
                        E[y0 + r, wq0] = e[r] & 0xFFFFFFFF
                        if wq0 + 1 < Wq:
                            E[y0 + r, wq0 + 1] = e[r] >> 32
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        if (dy or dx) and 0 <= ty + dy < ty_n and 0 <= tx + dx < tx_n:
                            nxt.add((ty + dy, tx + dx))
        dirty = nxt
    return E, rounds


def unpack(bits, H, W):
    b = np.unpackbits(bits.view(np.uint8).reshape(H, -1), axis=1, bitorder="little")
    return (b[:, :W] * 255).astype(np.uint8)


def test_run_fill_fills_whole_runs():
    rng = np.random.default_rng(0)
    for _ in range(300):
        c = int(rng.integers(0, 1 << 63)) | (int(rng.integers(0, 2)) << 63)
        t = c & int(rng.integers(0, 1 << 63)) & int(rng.integers(0, 1 << 63))
        want, x = 0, 0
        while x < 64:
            if (c >> x) & 1:
                x2 = x
                while x2 < 64 and (c >> x2) & 1:
                    x2 += 1
                run = ((1 << (x2 - x)) - 1) << x
                if run & t:
                    want |= run
                x = x2
            else:
                x += 1
        assert run_fill(t, c) == want


def _snake(h, w):
    """a one-pixel weak path that winds through the whole image and touches one strong blob: the chain
    crosses tile borders dozens of times"""
    img = np.full((h, w), 40, np.uint8)
    for k, y in enumerate(range(6, h - 6, 8)):
        img[y:y + 2, 6:w - 6] = 90
        xs = w - 10 if k % 2 == 0 else 6
        img[y:y + 10, xs:xs + 2] = 90
    img[4:12, 4:12] = 255
    return cv2.GaussianBlur(img, (3, 3), 0)


@pytest.mark.parametrize("shape", [(40, 64), (97, 131), (33, 70), (300, 29), (70, 200), (150, 260)])
def test_bit_plane_canny_equals_cv2(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    imgs = [cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (5, 5), 0),
            rng.integers(0, 256, (h, w), dtype=np.uint8),
            _snake(h, w),
            np.clip(np.add.outer(np.arange(h) * 2, np.arange(w)) % 256 + rng.integers(-6, 7, (h, w)), 0, 255).astype(np.uint8)]
    max_rounds = 0
    for lum in imgs:
        low, high = M.canny_thresholds(float(np.median(lum)))
        if lum is imgs[2]:
            low, high = 20, 120       # the path is weak, the blob strong
        E0, C = classify_bits(lum, low, high)
        E, rounds = hysteresis_bits(E0, C, h, w)
        max_rounds = max(max_rounds, rounds)
        assert np.array_equal(unpack(E, h, w), cv2.Canny(lum, low, high))
        assert not (E0 & ~C).any()
    assert max_rounds >= 2
