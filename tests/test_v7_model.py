"""CPU pin of the pixel-pair HSV arithmetic (csrc/hsv_half2.cuh, engine variants 7/8): its numpy
restatement (tests/v7_model.py) against the integer oracle over all 2^24 colours, with the colours
shuffled so that every pair / quad position of the kernel's byte layout sees every kind of colour."""

import numpy as np

from oracle import intmath as M
from tests import v7_model as V7


def test_prmt_model():
    a, b = np.uint32(0x33221100), np.uint32(0x77665544)
    assert int(V7.prmt(a, b, 0x3210)) == 0x33221100
    assert int(V7.prmt(a, b, 0x7654)) == 0x77665544
    assert int(V7.prmt(a, b, 0x4340)) == 0x44334400
    assert int(V7.prmt(a, b, 0x0051)) & 0xFFFF == 0x5511


def test_fix_hue4():
    # H mod 256 of -30..-1 -> 150..179, 0..179 unchanged, in every byte position
    vals = np.array(list(range(180)) + list(range(226, 256)), dtype=np.uint32)
    want = np.where(vals >= 226, vals - 76, vals)
    for sh in (0, 8, 16, 24):
        others = np.uint32(0xE2B300FF) & ~np.uint32(0xFF << sh)   # neighbours: negative, 179, 0, -1
        got = V7.fix_hue4((vals << np.uint32(sh)) | others)
        assert np.array_equal((got >> np.uint32(sh)) & 0xFF, want)


def test_variant7_model_exhaustive_2_24():
    step = 1 << 21
    for s0 in range(0, 1 << 24, step):
        c = np.arange(s0, s0 + step, dtype=np.uint32)
        bgr = np.stack([c & 255, (c >> 8) & 255, (c >> 16) & 255], axis=1).astype(np.uint8)
        bgr = bgr[np.random.default_rng(s0).permutation(step)]
        h, s, v = V7.planes_from_bgr(bgr)
        H, S, Vv = M.bgr_to_hsv(bgr)
        assert np.array_equal(h, H) and np.array_equal(s, S) and np.array_equal(v, Vv)
