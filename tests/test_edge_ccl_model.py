"""CPU restatement of the union rule of csrc/edge_kernels.cu (psd_hyst_runs_kernel +
psd_hyst_union_kernel + mark/resolve): hysteresis as connected components over horizontal runs must
equal "weak pixels 8-connected to a strong pixel" for any class map."""

import numpy as np
import pytest

ndimage = pytest.importorskip("scipy.ndimage")


def hysteresis_reference(m):
    lab, _ = ndimage.label(m > 0, structure=np.ones((3, 3)))
    strong = np.unique(lab[m == 2])
    return np.isin(lab, strong[strong > 0]) & (m > 0)


def hysteresis_runs(m):
    H, W = m.shape
    f = m.reshape(-1)
    L = np.arange(H * W)
    for y in range(H):  # psd_hyst_runs_kernel: label = first pixel of the horizontal run
        start = -1
        for x in range(W):
            p = y * W + x
            if f[p]:
                start = p if start < 0 else start
                L[p] = start
            else:
                start = -1

    def find(x):
        while L[x] != x:
            x = L[x]
        return x

    def unite(a, b):
        ra, rb = find(a), find(b)
        if ra != rb:
            L[max(ra, rb)] = min(ra, rb)

    for p in range(W, H * W):  # psd_hyst_union_kernel
        if not f[p]:
            continue
        x = p % W
        w_edge = x > 0 and f[p - 1]
        n_edge = f[p - W]
        ne_edge = x + 1 < W and f[p - W + 1]
        if not w_edge:
            if n_edge:
                unite(p, p - W)
            else:
                if x > 0 and f[p - W - 1]:
                    unite(p, p - W - 1)
                if ne_edge:
                    unite(p, p - W + 1)
        elif not n_edge and ne_edge:
            unite(p, p - W + 1)
    out = f.copy()
    for p in range(H * W):  # mark
        if f[p] == 2 and find(p) != p:
            out[find(p)] = 2
    res = out.copy()
    for p in range(H * W):  # resolve
        if out[p] == 1 and find(p) != p and out[find(p)] == 2:
            res[p] = 2
    return (res == 2).reshape(H, W)


def test_hysteresis_union_rule():
    rng = np.random.default_rng(2)
    for _ in range(80):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        dens = float(rng.choice([0.2, 0.45, 0.7, 0.95]))
        m = rng.choice([0, 1, 2], size=(H, W), p=[1 - dens, dens * 0.9, dens * 0.1]).astype(np.uint8)
        assert np.array_equal(hysteresis_reference(m), hysteresis_runs(m))
