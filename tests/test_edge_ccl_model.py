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


def hysteresis_tiled(m, TW=8, TH=4):
    """psd_hyst_tile_kernel + psd_hyst_border_kernel (modes 0, 1, 2) + mark + resolve."""
    H, W = m.shape
    cls = m.copy()
    L = -np.ones(H * W, dtype=np.int64)  # labels of non-edge pixels are never read
    for y0 in range(0, H, TH):
        for x0 in range(0, W, TW):
            t = m[y0:y0 + TH, x0:x0 + TW]
            th, tw = t.shape
            lab = np.arange(th * tw)
            for r in range(th):  # run starts inside the tile row
                start = -1
                for c in range(tw):
                    if t[r, c]:
                        start = r * tw + c if start < 0 else start
                        lab[r * tw + c] = start
                    else:
                        start = -1

            def find(x):
                while lab[x] != x:
                    x = lab[x]
                return x

            def unite(a, b):
                ra, rb = find(a), find(b)
                if ra != rb:
                    lab[max(ra, rb)] = min(ra, rb)

            for r in range(1, th):
                for c in range(tw):
                    if not t[r, c]:
                        continue
                    p = r * tw + c
                    w_edge = c > 0 and t[r, c - 1]
                    n_edge = t[r - 1, c]
                    ne_edge = c + 1 < tw and t[r - 1, c + 1]
                    if not w_edge:
                        if n_edge:
                            unite(p, p - tw)
                        else:
                            if c > 0 and t[r - 1, c - 1]:
                                unite(p, p - tw - 1)
                            if ne_edge:
                                unite(p, p - tw + 1)
                    elif not n_edge and ne_edge:
                        unite(p, p - tw + 1)
            strong_root = set(find(r * tw + c) for r in range(th) for c in range(tw) if t[r, c] == 2)
            for r in range(th):
                for c in range(tw):
                    if t[r, c]:
                        root = find(r * tw + c)
                        L[(y0 + r) * W + x0 + c] = (y0 + root // tw) * W + x0 + root % tw
                        if t[r, c] == 1 and root in strong_root:
                            cls[y0 + r, x0 + c] = 2
    f = cls.reshape(-1)

    def gfind(x):
        while L[x] != x:
            x = L[x]
        return x

    def gunite(a, b):
        ra, rb = gfind(a), gfind(b)
        if ra != rb:
            L[max(ra, rb)] = min(ra, rb)

    for y in range(TH, H, TH):  # mode 0
        for x in range(W):
            p = y * W + x
            if f[p]:
                if f[p - W]: gunite(p, p - W)
                if x > 0 and f[p - W - 1]: gunite(p, p - W - 1)
                if x + 1 < W and f[p - W + 1]: gunite(p, p - W + 1)
    for k in range(1, (W - 1) // TW + 1):
        for y in range(H):
            p = y * W + k * TW  # mode 1
            if f[p]:
                if f[p - 1]: gunite(p, p - 1)
                if y > 0 and f[p - W - 1]: gunite(p, p - W - 1)
            p = y * W + k * TW - 1  # mode 2
            if f[p] and y > 0 and f[p - W + 1]:
                gunite(p, p - W + 1)
    out = f.copy()
    for p in range(H * W):
        if f[p] == 2 and gfind(p) != p:
            out[gfind(p)] = 2
    res = out.copy()
    for p in range(H * W):
        if out[p] == 1 and gfind(p) != p and out[gfind(p)] == 2:
            res[p] = 2
    return (res == 2).reshape(H, W)


def test_hysteresis_tile_decomposition():
    rng = np.random.default_rng(3)
    for _ in range(80):
        H, W = int(rng.integers(1, 30)), int(rng.integers(1, 50))
        dens = float(rng.choice([0.2, 0.45, 0.7, 0.95]))
        m = rng.choice([0, 1, 2], size=(H, W), p=[1 - dens, dens * 0.93, dens * 0.07]).astype(np.uint8)
        assert np.array_equal(hysteresis_reference(m), hysteresis_tiled(m))
