"""Pin oracle.intmath (the integer restatement the CUDA kernels follow) against cv2/numpy."""

import cv2
import numpy as np
import pytest

from oracle import intmath as M
from oracle import ref_detectors as R
from pyscenedetect_b200.synth import ScenePlan, render_frames


def all_colours():
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v & 255), (v >> 8) & 255, (v >> 16) & 255], axis=-1).astype(np.uint8)
    return img.reshape(4096, 4096, 3)


def test_hsv_exhaustive_2_24():
    img = all_colours()
    want = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    h, s, v = M.bgr_to_hsv(img)
    assert np.array_equal(h, want[..., 0])
    assert np.array_equal(s, want[..., 1])
    assert np.array_equal(v, want[..., 2])


def test_y_exhaustive_2_24():
    img = all_colours()
    want = cv2.cvtColor(img, cv2.COLOR_BGR2YUV)[..., 0]
    assert np.array_equal(M.bgr_to_y(img), want)


@pytest.mark.parametrize("shape,dst", [((360, 640), (256, 144)), ((720, 1280), (256, 144)),
                                       ((1080, 1920), (256, 144)), ((288, 512), (256, 144)),
                                       ((270, 480), (160, 90)), ((97, 131), (50, 37)),
                                       ((2160, 3840), (256, 144))])
def test_resize_linear_bit_exact(shape, dst):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, size=(*shape, 3), dtype=np.uint8)
    want = cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)
    got = M.resize_linear(img, dst[0], dst[1])
    assert np.array_equal(got, want)


def test_mean_and_sad_bit_exact():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=(90, 160), dtype=np.uint8)
    b = rng.integers(0, 256, size=(90, 160), dtype=np.uint8)
    assert M.mean_from_sum(M.sad(a, b), a.size) == R.mean_pixel_distance(a, b)
    f = rng.integers(0, 256, size=(90, 160, 3), dtype=np.uint8)
    assert M.mean_from_sum(int(f.astype(np.int64).sum()), f.size) == np.mean(f)


@pytest.mark.parametrize("bins", [256, 128, 100, 7])
def test_hist_normalize_correl(bins):
    frames = render_frames(ScenePlan(40, seed=3, min_len=10, max_len=20).params, 160, 90)
    prev = None
    for f in frames:
        y = M.bgr_to_y(f)
        counts = M.hist_counts(y, bins)
        raw = cv2.calcHist([cv2.split(cv2.cvtColor(f, cv2.COLOR_BGR2YUV))[0]], [0], None, [bins], [0, 256])
        assert np.array_equal(counts, raw.ravel().astype(np.int64))
        norm = M.hist_normalize_l2(counts)
        want = R.calculate_histogram(f, bins)
        assert np.array_equal(norm, want)
        if prev is not None:
            got = M.hist_correl(prev, norm)
            ref = cv2.compareHist(prev, want, cv2.HISTCMP_CORREL)
            assert abs(got - ref) < 1e-9
        prev = want


def test_correl_edge_cases():
    one = np.zeros(128, np.float32); one[0] = 1.0
    two = np.zeros(128, np.float32); two[127] = 1.0
    flat = np.full(128, 1.0 / np.sqrt(128), np.float32)
    for a, b in [(one, one), (one, two), (flat, one), (flat, flat)]:
        assert abs(M.hist_correl(a, b) - cv2.compareHist(a, b, cv2.HISTCMP_CORREL)) < 1e-12


def _edge_images():
    rng = np.random.default_rng(7)
    out = [rng.integers(0, 256, size=(90, 160), dtype=np.uint8)]
    out.append(cv2.GaussianBlur(rng.integers(0, 256, size=(120, 200), dtype=np.uint8), (9, 9), 0))
    out.append(np.zeros((64, 96), np.uint8))
    out.append(np.full((64, 96), 255, np.uint8))
    box = np.zeros((80, 120), np.uint8); box[20:60, 30:90] = 200; out.append(box)
    ramp = (np.arange(200 * 150).reshape(150, 200) % 256).astype(np.uint8); out.append(ramp)
    frames = render_frames(ScenePlan(3, seed=5).params, 192, 108)
    for f in frames:
        out.append(cv2.split(cv2.cvtColor(f, cv2.COLOR_BGR2HSV))[2])
    return out


@pytest.mark.parametrize("i", range(9))
def test_median_canny_dilate(i):
    lum = _edge_images()[i]
    med = M.median_u8(lum)
    assert med == float(np.median(lum))
    low, high = M.canny_thresholds(med)
    want = cv2.Canny(lum, low, high)
    got = M.canny(lum, low, high)
    assert np.array_equal(got, want)
    for k in (3, 5, 13):
        assert np.array_equal(M.dilate_square(want, k), cv2.dilate(want, np.ones((k, k), np.uint8)))
