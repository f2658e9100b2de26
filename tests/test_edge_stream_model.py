"""CPU restatement of psd_canny_classify_stream_kernel (csrc/edge_kernels.cu): the register/shuffle
pipeline of one warp, emulated lane by lane, must give the same 0/1/2 class map as the whole-image
formulation of the oracle (Sobel with BORDER_REPLICATE, L1 magnitude zero-padded, fixed-point NMS,
strict double threshold)."""

import numpy as np

TG22 = 13573


def classify_reference(lum, low, high):
    h, w = lum.shape
    p = np.pad(lum.astype(np.int32), 1, mode="edge")
    gx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    gy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    mag = np.abs(gx) + np.abs(gy)
    m = np.pad(mag, 1, mode="constant")
    c = m[1:-1, 1:-1]
    ax = np.abs(gx).astype(np.int64)
    ay = np.abs(gy).astype(np.int64) << 15
    tg22x = ax * TG22
    tg67x = tg22x + (ax << 16)
    s = np.where((gx ^ gy) < 0, -1, 1)
    yy, xx = np.mgrid[0:h, 0:w]
    d1 = m[yy, xx + 1 - s]
    d2 = m[yy + 2, xx + 1 + s]
    keep = np.where(ay < tg22x, (c > m[1:-1, :-2]) & (c >= m[1:-1, 2:]),
                    np.where(ay > tg67x, (c > m[:-2, 1:-1]) & (c >= m[2:, 1:-1]), (c > d1) & (c > d2)))
    cand = keep & (c > low)
    return np.where(cand, np.where(c > high, 2, 1), 0).astype(np.uint8)


def shfl_up(v):  # lane l receives lane l-1 (lane 0 keeps its own value)
    return np.concatenate([v[:1], v[:-1]])


def shfl_down(v):
    return np.concatenate([v[1:], v[-1:]])


def classify_stream(lum, low, high, band_cols=28, band_rows=136):
    H, W = lum.shape
    out = np.full((H, W), 255, dtype=np.uint8)
    lane = np.arange(32)
    for yb in range(0, H, band_rows):
        ye = min(yb + band_rows, H)
        for xb in range(0, W, band_cols):
            x = xb - 2 + lane
            xc = np.clip(x, 0, W - 1)
            x_in = (x >= 0) & (x < W)
            writer = (lane >= 2) & (lane < 2 + band_cols) & (x < W)
            z = np.zeros(32, dtype=np.int64)
            l0, l1, rs0, rs1 = z, z, z, z
            mU, mUl, mUr, mC, mCl, mCr, gxC, gyC = z, z, z, z, z, z, z, z
            for r in range(yb - 2, ye + 2):
                l2 = lum[min(max(r, 0), H - 1), xc].astype(np.int64)
                rs2 = shfl_up(l2) + 2 * l2 + shfl_down(l2)
                col = l0 + 2 * l1 + l2
                gx = shfl_down(col) - shfl_up(col)
                gy = rs2 - rs0
                ok = x_in & (0 <= r - 1 < H)
                gx = np.where(ok, gx, 0)
                gy = np.where(ok, gy, 0)
                mD = np.abs(gx) + np.abs(gy)
                mDl, mDr = shfl_up(mD), shfl_down(mD)
                y = r - 2
                if yb <= y < ye:
                    m = mC
                    ax = np.abs(gxC)
                    ay = np.abs(gyC) << 15
                    tg22x = ax * TG22
                    tg67x = tg22x + (ax << 16)
                    keep = np.where(ay < tg22x, (m > mCl) & (m >= mCr),
                                    np.where(ay > tg67x, (m > mU) & (m >= mD),
                                             np.where((gxC ^ gyC) < 0, (m > mUr) & (m > mDl), (m > mUl) & (m > mDr))))
                    o = np.where((m > low) & keep, np.where(m > high, 2, 1), 0)
                    out[y, x[writer]] = o[writer]
                l0, l1, rs0, rs1 = l1, l2, rs1, rs2
                mU, mUl, mUr = mC, mCl, mCr
                mC, mCl, mCr = mD, mDl, mDr
                gxC, gyC = gx, gy
    return out


def test_stream_classify_equals_whole_image():
    rng = np.random.default_rng(5)
    for (h, w, bc, br) in [(37, 61, 28, 136), (50, 28, 28, 16), (9, 1, 28, 136), (1, 90, 28, 4), (64, 57, 28, 20)]:
        base = rng.integers(0, 256, (h, w), dtype=np.uint8)
        smooth = (np.cumsum(rng.integers(-6, 7, (h, w)), axis=1) % 256).astype(np.uint8)
        for lum in (base, smooth):
            for low, high in ((40, 120), (0, 255), (200, 300)):
                got = classify_stream(lum, low, high, bc, br)
                assert np.array_equal(got, classify_reference(lum, low, high)), (h, w, low, high)
