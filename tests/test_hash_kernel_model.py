"""CPU checks of the bit manipulation in csrc/hash_kernels.cu:psd_hash_rows_kernel (no GPU needed): the IDP4A form
of OpenCV's 15-bit gray, the PRMT selectors that cut four BGR pixels out of three words, the conversion-free
byte -> float32, and the first / whole-pixel run / last split of an INTER_AREA tap list."""

import numpy as np

from oracle import intmath as M
from tests.test_edge_bits_model import prmt


def dp4a(a, b, c):
    return (sum(((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF) for i in range(4)) + c) & 0xFFFFFFFF


def gray_word(px):
    hi = dp4a(px, 0x00264B0E, 64)
    return dp4a(px, 0x00462397, (hi << 8) & 0xFFFFFFFF) >> 15


def test_gray_by_two_dot_products_equals_opencv_fixed_point():
    rng = np.random.default_rng(3)
    cols = np.concatenate([rng.integers(0, 256, (5000, 3)), [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255]]])
    want = M.bgr_to_gray(cols.astype(np.uint8).reshape(-1, 1, 3)).reshape(-1)
    for (b, g, r), w in zip(cols, want):
        px = int(b) | int(g) << 8 | int(r) << 16 | int(rng.integers(0, 256)) << 24   # byte 3 carries weight 0
        assert gray_word(px) == int(w) == (int(b) * 3735 + int(g) * 19235 + int(r) * 9798 + 16384) >> 15


def test_four_pixels_from_three_words():
    rng = np.random.default_rng(4)
    for _ in range(200):
        by = [int(v) for v in rng.integers(0, 256, 12)]
        w0, w1, w2 = (by[4 * j] | by[4 * j + 1] << 8 | by[4 * j + 2] << 16 | by[4 * j + 3] << 24 for j in range(3))
        px = [w0, prmt(w0, w1, 0x0543), prmt(w1, w2, 0x0432), w2 >> 8]
        for k in range(4):
            b, g, r = by[3 * k:3 * k + 3]
            assert gray_word(px[k]) == (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15


def test_byte_to_float_without_conversion():
    g = np.arange(256, dtype=np.uint32)
    f = (np.uint32(0x4B000000) | g).view(np.float32) - np.float32(8388608.0)
    assert np.array_equal(f, g.astype(np.float32))


def test_area_taps_split_into_first_run_last():
    """hash_plan_create's xmid table: the whole-pixel taps of a destination column are consecutive source pixels
    with one common weight, so the kernel's three loops walk the same (index, weight) sequence as the table."""
    for ssize, dsize in [(1920, 16), (1080, 16), (100, 7), (97, 36), (64, 64), (33, 8)]:
        tab = M.area_tab(ssize, dsize)                      # (dst index, src index, weight) in source order
        si, alpha = [t[1] for t in tab], [float(t[2]) for t in tab]
        start = [next((k for k, t in enumerate(tab) if t[0] >= dx), len(tab)) for dx in range(dsize + 1)]
        scale = ssize / dsize
        for dx in range(dsize):
            fsx1, fsx2 = dx * scale, dx * scale + scale
            sx1, sx2 = int(np.ceil(fsx1)), int(np.floor(fsx2))
            sx2 = min(sx2, ssize - 1)
            sx1 = min(sx1, sx2)
            km = start[dx] + (1 if sx1 - fsx1 > 1e-3 else 0)
            nm = max(0, sx2 - sx1)
            seq = [(si[k], alpha[k]) for k in range(start[dx], km)]
            if nm:
                seq += [(si[km] + i, alpha[km]) for i in range(nm)]
            seq += [(si[k], alpha[k]) for k in range(km + nm, start[dx + 1])]
            assert seq == [(si[k], alpha[k]) for k in range(start[dx], start[dx + 1])], (ssize, dsize, dx)
