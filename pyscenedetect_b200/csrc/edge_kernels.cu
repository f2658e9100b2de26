// Edge component of ContentDetector (content_detector.py:213-239):
//   median = numpy.median(V);  low/high = int(max(0,(1-s)*median)), int(min(255,(1+s)*median)), s = 1/3
//   edges  = cv2.dilate(cv2.Canny(V, low, high), ones(k,k))
//   delta_edges = mean |edges_t - edges_{t-1}|                       (content_detector.py:171-175)
// Canny is restated from OpenCV's algorithm (aperture 3, L1 gradient): Sobel 3x3 with
// BORDER_REPLICATE, fixed-point non-maximum suppression (TG22 = 13573), double threshold with
// strict '>' and 8-connected hysteresis.  oracle/intmath.py:canny is the CPU twin pinned
// against cv2.Canny.  Stages: thresholds (from the V histogram the score pass produced) ->
// gradient/NMS/classify -> hysteresis -> separable k x k max -> SAD against the previous frame's
// dilated map.  Hysteresis ("weak pixels 8-connected to a strong pixel become edges") is solved as
// connected-component labelling with a lock-free union-find over the weak+strong pixels: three
// launches per batch whatever the length of the weak chains, no host round trip.  (The earlier
// tile-local fix-point iteration needed ~100 chained launches per batch on noisy frames; it is kept
// behind PSD_EDGE_HYSTERESIS=tiles as a cross-check.)
#include "psd_common.cuh"

namespace psd {

// ---- 1. per-frame Canny thresholds from the V histogram ----
__global__ void psd_edge_thresholds_kernel(const uint32_t* __restrict__ vhist, int n, int64_t n_pixels,
                                           int32_t* __restrict__ thr) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const uint32_t* h = vhist + (int64_t)f * 256;
    // numpy.median: mean of the order statistics (n-1)//2 and n//2 (0-based)
    const int64_t r_lo = (n_pixels - 1) / 2 + 1, r_hi = n_pixels / 2 + 1;
    int lo = -1, hi = -1;
    int64_t cum = 0;
    for (int i = 0; i < 256; ++i) {
        cum += h[i];
        if (lo < 0 && cum >= r_lo) lo = i;
        if (hi < 0 && cum >= r_hi) hi = i;
    }
    const double median = __ddiv_rn((double)(lo + hi), 2.0);
    const double sigma = __ddiv_rn(1.0, 3.0);
    const double lo_d = __dmul_rn(__dsub_rn(1.0, sigma), median);
    const double hi_d = __dmul_rn(__dadd_rn(1.0, sigma), median);
    int low = (int)(lo_d > 0.0 ? lo_d : 0.0);       // int(max(0, x)) truncates
    int high = (int)(hi_d < 255.0 ? hi_d : 255.0);  // int(min(255, x))
    if (low > high) { const int t = low; low = high; high = t; }  // cv2.Canny swaps
    thr[2 * f] = low;
    thr[2 * f + 1] = high;
}

// ---- 2. Sobel + L1 magnitude + NMS + double threshold ----
// 64x16-pixel tiles (4 pixels per thread): the 2-pixel lum halo and 1-pixel gradient halo cost
// 1.2x / 1.16x redundant work instead of 1.7x / 1.33x with 32x8 tiles, and 4x fewer CTAs.
constexpr int TX = 64, TY = 16, kClassifyThreads = 256;

__global__ void __launch_bounds__(kClassifyThreads) psd_canny_classify_kernel(
    const uint8_t* __restrict__ vplane, const int32_t* __restrict__ thr, uint8_t* __restrict__ map, int W,
    int H) {
    __shared__ uint8_t lum[TY + 4][TX + 4];
    __shared__ int16_t sgx[TY + 2][TX + 2];
    __shared__ int16_t sgy[TY + 2][TX + 2];
    const int f = blockIdx.z;
    const int64_t P = (int64_t)W * H;
    const uint8_t* src = vplane + f * P;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x;
    for (int i = tid; i < (TY + 4) * (TX + 4); i += kClassifyThreads) {
        const int ly = i / (TX + 4), lx = i - ly * (TX + 4);
        const int gy = min(max(y0 + ly - 2, 0), H - 1), gx = min(max(x0 + lx - 2, 0), W - 1);  // replicate
        lum[ly][lx] = src[(int64_t)gy * W + gx];
    }
    __syncthreads();
    for (int i = tid; i < (TY + 2) * (TX + 2); i += kClassifyThreads) {
        const int ly = i / (TX + 2), lx = i - ly * (TX + 2);
        const int gy = y0 + ly - 1, gx = x0 + lx - 1;
        int dx = 0, dy = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            // lum index of (gy,gx) is [ly+1][lx+1]
            const int a = lum[ly][lx], b = lum[ly][lx + 1], c = lum[ly][lx + 2];
            const int d = lum[ly + 1][lx], e = lum[ly + 1][lx + 2];
            const int g = lum[ly + 2][lx], h = lum[ly + 2][lx + 1], k = lum[ly + 2][lx + 2];
            dx = (c + 2 * e + k) - (a + 2 * d + g);
            dy = (g + 2 * h + k) - (a + 2 * b + c);
        }
        sgx[ly][lx] = (int16_t)dx;  // outside the image: 0 => magnitude 0
        sgy[ly][lx] = (int16_t)dy;
    }
    __syncthreads();
    const int low = thr[2 * f], high = thr[2 * f + 1];
    auto mag = [&](int yy, int xx) { return abs((int)sgx[yy][xx]) + abs((int)sgy[yy][xx]); };
    for (int i = tid; i < TX * TY; i += kClassifyThreads) {
        const int ty = i / TX, tx = i - ty * TX;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= W || y >= H) continue;
        const int ly = ty + 1, lx = tx + 1;
        const int gx = sgx[ly][lx], gy = sgy[ly][lx];
        const int m = abs(gx) + abs(gy);
        uint8_t out = 0;
        if (m > low) {
            const int ax = abs(gx);
            const int ay = abs(gy) << 15;
            const int tg22x = ax * 13573;
            const int tg67x = tg22x + (ax << 16);
            bool keep;
            if (ay < tg22x) {
                keep = (m > mag(ly, lx - 1)) && (m >= mag(ly, lx + 1));
            } else if (ay > tg67x) {
                keep = (m > mag(ly - 1, lx)) && (m >= mag(ly + 1, lx));
            } else {
                const int s = ((gx ^ gy) < 0) ? -1 : 1;
                keep = (m > mag(ly - 1, lx - s)) && (m > mag(ly + 1, lx + s));
            }
            if (keep) out = (m > high) ? 2 : 1;
        }
        map[f * P + (int64_t)y * W + x] = out;
    }
}

// ---- 2b. the same classification, streamed through registers ----
// One warp owns a band of kBandCols output columns x kBandRows rows and marches down it one image
// row per step.  Lane l holds column (band start - 2 + l): a 2-column apron on each side covers the
// Sobel and the non-maximum-suppression neighbourhoods, horizontal neighbours travel by warp shuffle,
// vertical ones stay in registers (3 rows of V, 3 rows of magnitudes).  Per row: one byte load,
// six shuffles, no shared memory, no barrier.  Same arithmetic as psd_canny_classify_kernel.
constexpr int kBandCols = 28, kBandRows = 136;

__global__ void __launch_bounds__(256) psd_canny_classify_stream_kernel(
    const uint8_t* __restrict__ vplane, const int32_t* __restrict__ thr, uint8_t* __restrict__ map, int W,
    int H, int bands_x, int bands_y, int64_t n_warps) {
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (wid >= n_warps) return;
    const int lane = threadIdx.x & 31;
    const int bx = (int)(wid % bands_x);
    const int by = (int)((wid / bands_x) % bands_y);
    const int64_t f = wid / ((int64_t)bands_x * bands_y);
    const int64_t P = (int64_t)W * H;
    const uint8_t* src = vplane + f * P;
    uint8_t* dst = map + f * P;
    const int low = thr[2 * f], high = thr[2 * f + 1];
    const int xb = bx * kBandCols, yb = by * kBandRows;
    const int ye = min(yb + kBandRows, H);
    const int x = xb - 2 + lane;                 // this lane's column (may lie outside the image)
    const int xc = min(max(x, 0), W - 1);        // BORDER_REPLICATE
    const bool x_in = x >= 0 && x < W;
    const bool writer = lane >= 2 && lane < 2 + kBandCols && x < W;
    const unsigned full = 0xFFFFFFFFu;

    int l0 = 0, l1 = 0;          // V of rows r-2, r-1
    int rs0 = 0, rs1 = 0;        // horizontal 1-2-1 sums of rows r-2, r-1
    int mU = 0, mUl = 0, mUr = 0;  // magnitudes of row r-3 (centre, left, right)
    int mC = 0, mCl = 0, mCr = 0;  // ... row r-2
    int gxC = 0, gyC = 0;          // gradient of row r-2
    auto load_row = [&](int r) { return (int)src[(int64_t)min(max(r, 0), H - 1) * W + xc]; };
    int l_next = load_row(yb - 2);
    for (int r = yb - 2; r <= ye + 1; ++r) {
        const int l2 = l_next;
        l_next = load_row(r + 1);  // one row ahead of the shuffle chain (clamped: always in bounds)
        const int rs2 = __shfl_up_sync(full, l2, 1) + 2 * l2 + __shfl_down_sync(full, l2, 1);
        // gradient and magnitude of row r-1 (zero outside the image, as cv2 pads the magnitude buffer)
        const int col = l0 + 2 * l1 + l2;
        int gx = __shfl_down_sync(full, col, 1) - __shfl_up_sync(full, col, 1);
        int gy = rs2 - rs0;
        if (!(x_in && r - 1 >= 0 && r - 1 < H)) { gx = 0; gy = 0; }
        const int mD = abs(gx) + abs(gy);
        const int mDl = __shfl_up_sync(full, mD, 1), mDr = __shfl_down_sync(full, mD, 1);
        // classify row r-2 from the magnitudes of rows r-3, r-2, r-1
        const int y = r - 2;
        if (y >= yb && y < ye && writer) {
            const int m = mC;
            uint8_t out = 0;
            if (m > low) {
                const int ax = abs(gxC);
                const int ay = abs(gyC) << 15;
                const int tg22x = ax * 13573;
                const int tg67x = tg22x + (ax << 16);
                bool keep;
                if (ay < tg22x) {
                    keep = (m > mCl) && (m >= mCr);
                } else if (ay > tg67x) {
                    keep = (m > mU) && (m >= mD);
                } else if ((gxC ^ gyC) < 0) {  // s = -1: compare (y-1, x+1) and (y+1, x-1)
                    keep = (m > mUr) && (m > mDl);
                } else {                       // s = +1: compare (y-1, x-1) and (y+1, x+1)
                    keep = (m > mUl) && (m > mDr);
                }
                if (keep) out = (m > high) ? 2 : 1;
            }
            dst[(int64_t)y * W + x] = out;
        }
        l0 = l1; l1 = l2;
        rs0 = rs1; rs1 = rs2;
        mU = mC; mUl = mCl; mUr = mCr;
        mC = mD; mCl = mDl; mCr = mDr;
        gxC = gx; gyC = gy;
    }
}

// ---- 3a. hysteresis as connected components (union-find with atomicCAS, roots = smallest index) ----
// labels[p] <= p always; a pixel is a root iff labels[p] == p.  Reads may see an older (larger)
// ancestor, which is still an ancestor, so every race is benign; see ECL-CC (Jaiganesh & Burtscher).
__device__ __forceinline__ int32_t ccl_find(int32_t* L, int32_t x) {
    volatile int32_t* V = L;
    int32_t y = V[x];
    if (y != x) {
        int32_t prev = x, next;
        while (y > (next = V[y])) {  // intermediate pointer jumping
            V[prev] = next;
            prev = y;
            y = next;
        }
    }
    return y;
}
__device__ __forceinline__ void ccl_unite(int32_t* L, int32_t a, int32_t b) {
    int32_t ra = ccl_find(L, a), rb = ccl_find(L, b);
    while (ra != rb) {
        if (ra < rb) { const int32_t t = ra; ra = rb; rb = t; }  // hang the larger root under the smaller
        const int32_t old = atomicCAS(&L[ra], ra, rb);
        if (old == ra) break;
        ra = old;  // ra had stopped being a root: continue from its parent
    }
}

// Initial labels: every edge pixel (class 1 or 2) points at the first pixel of its horizontal run, so
// horizontal chains have depth 1 before any union.  One warp per image row walks it 32 pixels at a
// time: the ballot of the edge flags gives the run start of every lane with two bit operations; a run
// that touches the end of a 32-pixel chunk hands its start to the next chunk.
__global__ void __launch_bounds__(256) psd_hyst_runs_kernel(const uint8_t* __restrict__ map,
                                                            int32_t* __restrict__ labels, int W, int H,
                                                            int64_t n_rows) {
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;  // frame * H + y
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int y = (int)(row % H);
    const uint8_t* m = map + row * W;
    int32_t* L = labels + row * W;
    int carry = -1;  // x of the start of a run that reached the end of the previous chunk
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const bool e = x < W && m[x] != 0;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, e);
        const uint32_t zeros_below = ~mask & ((2u << lane) - 1u);  // non-edge positions at or below this lane
        int start;
        if (zeros_below == 0u) start = carry >= 0 ? carry : x0;
        else start = x0 + 32 - __clz(zeros_below);
        if (e) L[x] = y * W + start;
        const int s31 = __shfl_sync(0xFFFFFFFFu, start, 31);
        carry = (mask >> 31) ? s31 : -1;
    }
}

// Edge pixels are a few percent of a frame and come in lines.  The three passes below give every
// thread kScanPerThread pixels spaced one CTA width apart: a byte-per-thread grid spends its time
// launching CTAs, and 16 CONSECUTIVE pixels per thread serialise the pointer chasing of a whole edge
// segment in one thread (measured 3x slower); strided pixels keep the loads coalesced and spread an
// edge segment over the threads of a warp.
constexpr int kScanPerThread = 16;
template <typename Fn>
__device__ __forceinline__ void for_each_class_byte16(const uint8_t* map, int64_t total, Fn fn) {
    const int64_t base = (int64_t)blockIdx.x * (kScanPerThread * 256) + threadIdx.x;
#pragma unroll 4
    for (int k = 0; k < kScanPerThread; ++k) {
        const int64_t g = base + k * 256;
        if (g >= total) return;
        const uint32_t c = map[g];
        if (c) fn(g, c);
    }
}

// Link the runs of adjacent rows (8-connectivity).  A run start looks at N, or at NW and NE when N is
// not an edge pixel (N's run already contains NW and NE otherwise); a pixel inside a run only has to
// add NE when N is not an edge pixel - every other contact was made by its W neighbour.
// (tests/test_edge_ccl_model.py restates this rule on the CPU.)
__global__ void __launch_bounds__(256) psd_hyst_union_kernel(const uint8_t* __restrict__ map,
                                                             int32_t* __restrict__ labels, int W, int H,
                                                             int64_t total) {
    const int64_t P = (int64_t)W * H;
    for_each_class_byte16(map, total, [&](int64_t g, uint32_t) {
        const int64_t f = g / P;
        const int32_t p = (int32_t)(g - f * P);
        const int y = p / W, x = p - y * W;
        if (y == 0) return;
        const uint8_t* m = map + f * P;
        int32_t* L = labels + f * P;
        const bool w_edge = x > 0 && m[p - 1];
        const bool n_edge = m[p - W] != 0;
        const bool ne_edge = x + 1 < W && m[p - W + 1];
        if (!w_edge) {
            if (n_edge) {
                ccl_unite(L, p, p - W);
            } else {
                if (x > 0 && m[p - W - 1]) ccl_unite(L, p, p - W - 1);
                if (ne_edge) ccl_unite(L, p, p - W + 1);
            }
        } else if (!n_edge && ne_edge) {
            ccl_unite(L, p, p - W + 1);
        }
    });
}

// Strong pixels mark the root of their component as strong (the root is itself an edge pixel of that
// component, so promoting it is part of the answer).  `sc` is the class that carries the mark: 2 when
// the labels come straight from the global union-find, 3 ("strong tile-local root") after
// psd_hyst_tile_kernel - then only one pixel per tile-local component has to chase its global root.
__global__ void __launch_bounds__(256) psd_hyst_mark_kernel(uint8_t* map, int32_t* __restrict__ labels,
                                                            int64_t P, int64_t total, uint32_t sc) {
    for_each_class_byte16(map, total, [&](int64_t g, uint32_t c) {
        if (c != sc) return;
        const int64_t f = g / P;
        const int32_t p = (int32_t)(g - f * P);
        const int32_t r = ccl_find(labels + f * P, p);
        if (r != p) map[f * P + r] = (uint8_t)sc;
    });
}

// ---- 3c. hysteresis, production path: tile-local components in shared memory, then border links ----
// The global union-find above is latency-bound: components are a few hundred pixels, but a vertical
// edge is a chain of runs, and every hop of find() through HBM/L2 costs ~1 us.  Here one CTA labels a
// 64x32 tile entirely in shared memory (run starts from row bit masks, unions and finds at ~30 cycles
// per hop), resolves "weak next to strong" inside the tile at once, and writes for every edge pixel the
// global index of its tile-local root.  Only contacts ACROSS tile borders go through the global
// union-find (three thin launches over border pixels), so global trees are as deep as a component
// is wide in tiles.  psd_hyst_mark_kernel / psd_hyst_resolve_kernel then finish components that span
// tiles.  (tests/test_edge_ccl_model.py restates the decomposition on the CPU.)
constexpr int CTW = 64, CTH = 32;

__device__ __forceinline__ int sm_find(int32_t* L, int x) {
    volatile int32_t* V = L;
    int y = V[x];
    if (y != x) {
        int prev = x, next;
        while (y > (next = V[y])) {
            V[prev] = next;
            prev = y;
            y = next;
        }
    }
    return y;
}
__device__ __forceinline__ void sm_unite(int32_t* L, int a, int b) {
    int ra = sm_find(L, a), rb = sm_find(L, b);
    while (ra != rb) {
        if (ra < rb) { const int t = ra; ra = rb; rb = t; }
        const int old = atomicCAS(&L[ra], ra, rb);
        if (old == ra) break;
        ra = old;
    }
}

// rec: 16 words per tile - [0] = number of strong tile-local roots, [1..15] = their in-frame pixel index
// (more than 15: the mark pass scans the tile instead)
constexpr int kTileRec = 16;
__global__ void __launch_bounds__(256) psd_hyst_tile_kernel(uint8_t* __restrict__ map,
                                                            int32_t* __restrict__ labels,
                                                            int32_t* __restrict__ tile_rec, int W, int H) {
    __shared__ __align__(8) uint8_t cls[CTH][CTW];
    __shared__ int32_t lab[CTH * CTW];
    __shared__ uint32_t rowmask[CTH][2];
    __shared__ __align__(8) uint8_t strong_root[CTH * CTW];
    __shared__ int any_edge;
    __shared__ int n_sroots;
    __shared__ int32_t sroots[kTileRec - 1];
    const int tid = threadIdx.x;
    const int row = tid >> 3, c0 = (tid & 7) * 8;  // 8 consecutive pixels of one tile row per thread
    const int x0 = blockIdx.x * CTW, y0 = blockIdx.y * CTH;
    const int64_t P = (int64_t)W * H;
    uint8_t* m = map + (int64_t)blockIdx.z * P;
    int32_t* Lg = labels + (int64_t)blockIdx.z * P;
    int32_t* rec = tile_rec + (((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kTileRec;
    if (tid < 2 * CTH) (&rowmask[0][0])[tid] = 0;
    if (tid == 0) { any_edge = 0; n_sroots = 0; }
    __syncthreads();
    const int gy = y0 + row;
    // the thread's 8 class bytes travel as one 64-bit word; all loops below walk only its non-zero
    // bytes (edge pixels are ~3 % of a frame), which also keeps the kernel small enough for the
    // instruction cache - the fully unrolled first version spent most of its time in fetch stalls
    unsigned long long cpack = 0;
    if (gy < H) {
        const int64_t g0 = (int64_t)gy * W + x0 + c0;
        if ((W & 7) == 0 && x0 + c0 + 8 <= W) {
            cpack = *reinterpret_cast<const unsigned long long*>(m + g0);
        } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i)
                if (x0 + c0 + i < W) cpack |= (unsigned long long)m[g0 + i] << (8 * i);
        }
    }
    *reinterpret_cast<unsigned long long*>(&cls[row][c0]) = cpack;
    *reinterpret_cast<unsigned long long*>(&strong_root[row * CTW + c0]) = 0ull;
    uint32_t mask8 = 0;  // bit i set <=> pixel i of this thread is an edge pixel (class 1 or 2)
    {
        // a byte is non-zero <=> (b | b>>1) & 1 here, classes being 0, 1, 2
        const unsigned long long nz = (cpack | (cpack >> 1)) & 0x0101010101010101ull;
        mask8 = (uint32_t)((nz * 0x0102040810204080ull) >> 56);  // gather the eight flags into one byte
    }
    if (mask8) {
        atomicOr(&rowmask[row][c0 >> 5], mask8 << (c0 & 31));
        any_edge = 1;
    }
    __syncthreads();
    if (!any_edge) {
        if (tid == 0) rec[0] = 0;
        return;
    }
    auto cls_of = [&](int i) { return (uint32_t)(cpack >> (8 * i)) & 0xFFu; };
    // run starts: label = first pixel of the horizontal run inside this tile row
    const unsigned long long m64 = (unsigned long long)rowmask[row][0] | ((unsigned long long)rowmask[row][1] << 32);
#pragma unroll 1
    for (uint32_t mm = mask8; mm; mm &= mm - 1) {
        const int col = c0 + __ffs(mm) - 1;
        const unsigned long long zb = ~m64 & ((2ull << col) - 1ull);  // non-edge columns at or left of col
        const int start = zb ? 64 - __clzll((long long)zb) : 0;
        lab[row * CTW + col] = row * CTW + start;
    }
    __syncthreads();
    // link the runs of adjacent tile rows (same rule as psd_hyst_union_kernel; outside the tile = no edge)
    if (row > 0) {
#pragma unroll 1
        for (uint32_t mm = mask8; mm; mm &= mm - 1) {
            const int col = c0 + __ffs(mm) - 1, p = row * CTW + col;
            const bool w_edge = col > 0 && cls[row][col - 1];
            const bool n_edge = cls[row - 1][col] != 0;
            const bool ne_edge = col + 1 < CTW && cls[row - 1][col + 1];
            if (!w_edge) {
                if (n_edge) {
                    sm_unite(lab, p, p - CTW);
                } else {
                    if (col > 0 && cls[row - 1][col - 1]) sm_unite(lab, p, p - CTW - 1);
                    if (ne_edge) sm_unite(lab, p, p - CTW + 1);
                }
            } else if (!n_edge && ne_edge) {
                sm_unite(lab, p, p - CTW + 1);
            }
        }
    }
    __syncthreads();
#pragma unroll 1
    for (uint32_t mm = mask8; mm; mm &= mm - 1) {
        const int i = __ffs(mm) - 1;
        if (cls_of(i) == 2u) strong_root[sm_find(lab, row * CTW + c0 + i)] = 1;
    }
    __syncthreads();
#pragma unroll 1
    for (uint32_t mm = mask8; mm; mm &= mm - 1) {
        const int i = __ffs(mm) - 1;
        const int r = sm_find(lab, row * CTW + c0 + i);
        const int64_t g = (int64_t)gy * W + x0 + c0 + i;
        Lg[g] = (y0 + (r >> 6)) * W + x0 + (r & 63);      // global index of the tile-local root
        // resolved inside the tile: members of a component with a strong pixel become 2, its root 3
        if (strong_root[r]) {
            const bool is_root = (r == row * CTW + c0 + i);
            const uint32_t nc = is_root ? 3u : 2u;
            if (cls_of(i) != nc) m[g] = (uint8_t)nc;
            if (is_root) {
                const int slot = atomicAdd(&n_sroots, 1);
                if (slot < kTileRec - 1) sroots[slot] = (int32_t)g;
            }
        }
    }
    __syncthreads();
    if (tid == 0) rec[0] = n_sroots;
    else if (tid < kTileRec && tid - 1 < n_sroots) rec[tid] = sroots[tid - 1];
}

// every strong tile-local root (class 3) marks its global root: one thread per tile record
__global__ void __launch_bounds__(256) psd_hyst_mark_tiles_kernel(uint8_t* map, int32_t* __restrict__ labels,
                                                                  const int32_t* __restrict__ tile_rec, int W,
                                                                  int H, int tiles_x, int tiles_y, int64_t n_tiles) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int32_t* rec = tile_rec + t * kTileRec;
    const int cnt = rec[0];
    if (cnt == 0) return;
    const int64_t P = (int64_t)W * H;
    const int64_t f = t / ((int64_t)tiles_x * tiles_y);
    uint8_t* m = map + f * P;
    int32_t* L = labels + f * P;
    if (cnt <= kTileRec - 1) {
        for (int k = 0; k < cnt; ++k) {
            const int32_t p = rec[1 + k];
            const int32_t r = ccl_find(L, p);
            if (r != p) m[r] = 3;
        }
    } else {  // more strong components in this tile than the record holds: look at every pixel of the tile
        const int tt = (int)(t - f * tiles_x * tiles_y);
        const int x0 = (tt % tiles_x) * CTW, y0 = (tt / tiles_x) * CTH;
        for (int y = y0; y < min(y0 + CTH, H); ++y)
            for (int x = x0; x < min(x0 + CTW, W); ++x) {
                const int32_t p = y * W + x;
                if (*(volatile uint8_t*)(m + p) != 3) continue;
                const int32_t r = ccl_find(L, p);
                if (r != p) m[r] = 3;
            }
    }
}

// contacts across tile borders.  mode 0: pixels of the first row of a tile row (N, NW, NE lie in other
// tiles); mode 1: first column of a tile column (W, NW); mode 2: last column of a tile column (NE).
__global__ void __launch_bounds__(256) psd_hyst_border_kernel(const uint8_t* __restrict__ map,
                                                              int32_t* __restrict__ labels, int W, int H,
                                                              int n, int mode) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t P = (int64_t)W * H;
    int x, y;
    int64_t f;
    if (mode == 0) {
        const int lines = (H - 1) / CTH;  // tile rows 1..lines start at y = CTH * k
        if (lines <= 0 || t >= (int64_t)n * lines * W) return;
        x = (int)(t % W);
        const int64_t q = t / W;
        y = ((int)(q % lines) + 1) * CTH;
        f = q / lines;
    } else {
        const int lines = (mode == 1) ? (W - 1) / CTW : (W - 1) / CTW;  // borders between tile columns
        if (lines <= 0 || t >= (int64_t)n * lines * H) return;
        y = (int)(t % H);
        const int64_t q = t / H;
        const int k = (int)(q % lines) + 1;
        x = (mode == 1) ? k * CTW : k * CTW - 1;
        f = q / lines;
    }
    const uint8_t* m = map + f * P;
    int32_t* L = labels + f * P;
    const int32_t p = y * W + x;
    if (m[p] == 0) return;
    if (mode == 0) {
        if (m[p - W]) ccl_unite(L, p, p - W);
        if (x > 0 && m[p - W - 1]) ccl_unite(L, p, p - W - 1);
        if (x + 1 < W && m[p - W + 1]) ccl_unite(L, p, p - W + 1);
    } else if (mode == 1) {
        if (m[p - 1]) ccl_unite(L, p, p - 1);
        if (y > 0 && m[p - W - 1]) ccl_unite(L, p, p - W - 1);
    } else {
        if (y > 0 && x + 1 < W && m[p - W + 1]) ccl_unite(L, p, p - W + 1);
    }
}

// ---- 3b. hysteresis, cross-check implementation: tile-local fix-point, repeated until no tile changes ----
constexpr int HTX = 64, HTY = 32;  // tile size (pixels)

// Launch i of a round reads flag[i-1] and returns at once when the previous launch changed
// nothing (the map is at its fix-point), so a round can be enqueued blind without host syncs.
// Per-tile dirty bytes (double-buffered by launch parity) restrict every launch after the first
// to the frontier: a tile is revisited only if it or one of its 8 neighbours changed last time.
__global__ void __launch_bounds__(256) psd_hysteresis_kernel(uint8_t* __restrict__ map, int W, int H,
                                                             const int32_t* __restrict__ prev_changed,
                                                             int32_t* __restrict__ changed,
                                                             const uint8_t* __restrict__ dirty_prev,
                                                             uint8_t* __restrict__ dirty_cur) {
    __shared__ uint8_t t[HTY + 2][HTX + 2 + 2];
    __shared__ int any_weak;
    if (prev_changed != nullptr && *prev_changed == 0) return;
    const int f = blockIdx.z;
    const int tid = threadIdx.x;
    const int tiles_x = gridDim.x, tiles_y = gridDim.y;
    const int64_t tile_id = ((int64_t)f * tiles_y + blockIdx.y) * tiles_x + blockIdx.x;
    if (dirty_prev != nullptr) {
        bool need = false;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int ty = (int)blockIdx.y + dy, tx = (int)blockIdx.x + dx;
                if (ty >= 0 && ty < tiles_y && tx >= 0 && tx < tiles_x)
                    need |= dirty_prev[((int64_t)f * tiles_y + ty) * tiles_x + tx] != 0;
            }
        if (!need) {
            if (tid == 0) dirty_cur[tile_id] = 0;
            return;
        }
    }
    const int64_t P = (int64_t)W * H;
    uint8_t* m = map + f * P;
    const int x0 = blockIdx.x * HTX, y0 = blockIdx.y * HTY;
    if (tid == 0) any_weak = 0;
    __syncthreads();
    int weak = 0;
    for (int i = tid; i < (HTY + 2) * (HTX + 2); i += 256) {
        const int ly = i / (HTX + 2), lx = i - ly * (HTX + 2);
        const int gy = y0 + ly - 1, gx = x0 + lx - 1;
        uint8_t v = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = m[(int64_t)gy * W + gx];
        t[ly][lx] = v;
        if (v == 1 && ly >= 1 && ly <= HTY && lx >= 1 && lx <= HTX) weak = 1;
    }
    if (weak) any_weak = 1;
    __syncthreads();
    if (!any_weak) {
        if (tid == 0) dirty_cur[tile_id] = 0;
        return;
    }
    // compact the weak pixels of the tile once; the fix-point loop then only visits those
    __shared__ uint16_t weak_list[HTX * HTY];
    __shared__ int n_weak;
    if (tid == 0) n_weak = 0;
    __syncthreads();
    for (int i = tid; i < HTX * HTY; i += 256) {
        const int ly = 1 + i / HTX, lx = 1 + (i % HTX);
        if (t[ly][lx] == 1) weak_list[atomicAdd(&n_weak, 1)] = (uint16_t)(ly * (HTX + 4) + lx);
    }
    __syncthreads();
    const int nw = n_weak;
    volatile uint8_t* vt = &t[0][0];
    constexpr int S = HTX + 4;  // row stride of the tile
    int tile_changed = 0;
    while (true) {
        int ch = 0;
        for (int i = tid; i < nw; i += 256) {
            const int p = weak_list[i];
            if (vt[p] == 1) {
                const bool s = vt[p - S - 1] == 2 || vt[p - S] == 2 || vt[p - S + 1] == 2 || vt[p - 1] == 2 ||
                               vt[p + 1] == 2 || vt[p + S - 1] == 2 || vt[p + S] == 2 || vt[p + S + 1] == 2;
                if (s) {
                    vt[p] = 2;
                    ch = 1;
                }
            }
        }
        if (!__syncthreads_or(ch)) break;
        tile_changed = 1;
    }
    if (tile_changed) {
        for (int i = tid; i < HTX * HTY; i += 256) {
            const int ly = 1 + i / HTX, lx = 1 + (i % HTX);
            const int gy = y0 + ly - 1, gx = x0 + lx - 1;
            if (gy < H && gx < W && t[ly][lx] == 2) m[(int64_t)gy * W + gx] = 2;
        }
        if (tid == 0) atomicExch(changed, 1);
    }
    if (tid == 0) dirty_cur[tile_id] = tile_changed ? 1 : 0;
}

// ---- 4. dilate on bit-packed edge maps (32 pixels per word) ----
// pack: bit i of word (y, wq) = pixel (y, 32*wq + i) is an edge; pixels beyond W are 0.  With sc != 0
// the last step of the hysteresis is folded in: a weak pixel (class 1) whose component root carries the
// mark `sc` is an edge too (and is written back as class 2 for the debug taps); classes >= 2 are edges.
__global__ void __launch_bounds__(256) psd_edge_pack_kernel(uint8_t* map, int32_t* __restrict__ labels,
                                                            uint32_t sc, uint32_t* __restrict__ bits, int W,
                                                            int H, int Wq) {
    const int64_t P = (int64_t)W * H;
    const int64_t f = blockIdx.z;
    const int y = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp -> 32 words of this row
    const int wq0 = warp * 32;
    if (wq0 >= Wq) return;
    uint8_t* mf = map + f * P;
    int32_t* L = labels + f * P;
    uint32_t mine = 0;
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
        const int x = (wq0 + j) * 32 + lane;
        bool edge = false;
        if (x < W) {
            const int32_t p = y * W + x;
            const uint32_t c = mf[p];
            edge = c >= 2u;
            if (c == 1u && sc != 0u) {
                const int32_t r = ccl_find(L, p);
                if (r != p && *(volatile uint8_t*)(mf + r) == sc) {
                    edge = true;
                    mf[p] = 2;
                }
            }
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, edge);
        if (lane == j) mine = b;
    }
    if (wq0 + lane < Wq) bits[(f * H + y) * Wq + wq0 + lane] = mine;
}

// rows: out = OR over |dx| <= r of the row shifted by dx (funnel shifts across word boundaries)
__global__ void __launch_bounds__(256) psd_edge_dilate_rows_bits_kernel(const uint32_t* __restrict__ in,
                                                                        uint32_t* __restrict__ out,
                                                                        int64_t n_words, int Wq, int r,
                                                                        uint32_t last_word_mask) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const int wq = (int)(i % Wq);
    const uint32_t cur = in[i];
    const uint32_t prv = (wq > 0) ? in[i - 1] : 0u;
    const uint32_t nxt = (wq + 1 < Wq) ? in[i + 1] : 0u;
    uint32_t o = cur;
    for (int s = 1; s <= r; ++s) {
        o |= __funnelshift_r(cur, nxt, s);  // pixel x+s -> bit position of x
        o |= __funnelshift_l(prv, cur, s);  // pixel x-s
    }
    // columns >= W of the last word must stay 0: the SAD counts whole words (found by the 131x97 case
    // of test_edge_intermediates_match_cv2: dilation spilled into the padding bits)
    if (wq == Wq - 1) o &= last_word_mask;
    out[i] = o;
}

// columns + SAD: dil[y] = OR over |dy| <= r of rows[y+dy]; count differing pixels vs previous frame
__global__ void __launch_bounds__(256) psd_edge_dilate_cols_bits_kernel(const uint32_t* __restrict__ rows,
                                                                        uint32_t* __restrict__ dil, int H,
                                                                        int Wq, int r) {
    const int64_t per_frame = (int64_t)H * Wq;
    const int64_t f = blockIdx.y;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= per_frame) return;
    const int y = (int)(i / Wq), wq = (int)(i - (int64_t)y * Wq);
    const uint32_t* base = rows + f * per_frame + wq;
    uint32_t o = 0;
    const int ya = max(y - r, 0), yb = min(y + r, H - 1);
    for (int yy = ya; yy <= yb; ++yy) o |= base[(int64_t)yy * Wq];
    dil[f * per_frame + i] = o;
}

__global__ void __launch_bounds__(256) psd_edge_sad_bits_kernel(const uint32_t* __restrict__ dil,
                                                                const uint32_t* __restrict__ carry,
                                                                int64_t per_frame, int have_prev,
                                                                psd_frame_sums* __restrict__ sums) {
    const int64_t f = blockIdx.y;
    if (f == 0 && !have_prev) return;
    const uint32_t* cur = dil + f * per_frame;
    const uint32_t* prv = (f == 0) ? carry : dil + (f - 1) * per_frame;
    uint32_t cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_frame;
         i += (int64_t)gridDim.x * blockDim.x)
        cnt += __popc(cur[i] ^ prv[i]);
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    __shared__ uint32_t part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        if (t) atomicAdd(reinterpret_cast<unsigned long long*>(&sums[f].sad_edges), 255ull * t);
    }
}

// debug/test tap: bit-packed map -> 0/255 bytes
__global__ void psd_edge_unpack_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ out, int W,
                                       int H, int Wq) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)W * H) return;
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    out[i] = ((bits[(int64_t)y * Wq + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
}

int edge_unpack(const uint32_t* bits, uint8_t* out, int W, int H, cudaStream_t stream) {
    const int Wq = (W + 31) / 32;
    psd_edge_unpack_kernel<<<(unsigned)(((int64_t)W * H + 255) / 256), 256, 0, stream>>>(bits, out, W, H, Wq);
    PSD_CHECK_LAUNCH();
    return PSD_OK;
}

int launch_edges(const EdgeBuffers& b, int n, int W, int H, int ksize, bool have_prev,
                 psd_frame_sums* sums, cudaStream_t stream) {
    PSD_REQUIRE(n > 0 && n <= 65535, "edge batch out of range");
    const int64_t P = (int64_t)W * H;
    psd_edge_thresholds_kernel<<<(n + 63) / 64, 64, 0, stream>>>(b.vhist, n, P, b.thresholds);
    PSD_CHECK_LAUNCH();
    dim3 cg((W + TX - 1) / TX, (H + TY - 1) / TY, (unsigned)n);
    static const bool use_tiles = [] {
        const char* v = getenv("PSD_EDGE_HYSTERESIS");
        return v && v[0] == 't';
    }();
    static const bool classify_tiles = [] {
        const char* v = getenv("PSD_EDGE_CLASSIFY");
        return v && v[0] == 't';
    }();
    if (classify_tiles) {
        psd_canny_classify_kernel<<<cg, kClassifyThreads, 0, stream>>>(b.vplane, b.thresholds, b.map, W, H);
    } else {
        const int bands_x = (W + kBandCols - 1) / kBandCols, bands_y = (H + kBandRows - 1) / kBandRows;
        const int64_t n_warps = (int64_t)bands_x * bands_y * n;
        psd_canny_classify_stream_kernel<<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, stream>>>(
            b.vplane, b.thresholds, b.map, W, H, bands_x, bands_y, n_warps);
    }
    PSD_CHECK_LAUNCH();
    count_launch(2);
    uint32_t resolve_class = 0;  // class that marks a strong component root; 0 = the map is already final
    if (!use_tiles) {
        const int64_t total = P * n;
        const unsigned blocks = (unsigned)((total + kScanPerThread * 256 - 1) / (kScanPerThread * 256));
        static const bool global_only = [] {
            const char* v = getenv("PSD_EDGE_HYSTERESIS");
            return v && v[0] == 'g';
        }();
        if (global_only) {  // cross-check: run labels + unions straight in global memory
            const int64_t n_rows = (int64_t)H * n;
            psd_hyst_runs_kernel<<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, stream>>>(b.map, b.labels, W, H, n_rows);
            PSD_CHECK_LAUNCH();
            psd_hyst_union_kernel<<<blocks, 256, 0, stream>>>(b.map, b.labels, W, H, total);
            PSD_CHECK_LAUNCH();
            psd_hyst_mark_kernel<<<blocks, 256, 0, stream>>>(b.map, b.labels, P, total, 2u);
            PSD_CHECK_LAUNCH();
            count_launch(3);
        } else {
            dim3 tg((W + CTW - 1) / CTW, (H + CTH - 1) / CTH, (unsigned)n);
            psd_hyst_tile_kernel<<<tg, 256, 0, stream>>>(b.map, b.labels, b.tile_rec, W, H);
            PSD_CHECK_LAUNCH();
            const int64_t hb = (int64_t)n * ((H - 1) / CTH) * W, vb = (int64_t)n * ((W - 1) / CTW) * H;
            if (hb > 0) psd_hyst_border_kernel<<<(unsigned)((hb + 255) / 256), 256, 0, stream>>>(b.map, b.labels, W, H, n, 0);
            if (vb > 0) {
                psd_hyst_border_kernel<<<(unsigned)((vb + 255) / 256), 256, 0, stream>>>(b.map, b.labels, W, H, n, 1);
                psd_hyst_border_kernel<<<(unsigned)((vb + 255) / 256), 256, 0, stream>>>(b.map, b.labels, W, H, n, 2);
            }
            PSD_CHECK_LAUNCH();
            const int64_t n_tiles = (int64_t)tg.x * tg.y * n;
            psd_hyst_mark_tiles_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(
                b.map, b.labels, b.tile_rec, W, H, (int)tg.x, (int)tg.y, n_tiles);
            PSD_CHECK_LAUNCH();
            count_launch(5);
        }
        resolve_class = global_only ? 2u : 3u;  // the resolve step itself is folded into psd_edge_pack_kernel
    } else {
        dim3 hg((W + HTX - 1) / HTX, (H + HTY - 1) / HTY, (unsigned)n);
        // Each launch reaches a fix-point inside every tile; edges crossing tiles need another launch.
        // A round enqueues kRound launches chained through device flags (a launch is a no-op once its
        // predecessor changed nothing) and only then asks the host whether another round is needed.
        constexpr int kRound = 8;
        const size_t tiles = (size_t)hg.x * hg.y * n;
        for (int round = 0; round < 100000; ++round) {
            PSD_CUDA(cudaMemsetAsync(b.changed, 0, kRound * sizeof(int32_t), stream));
            for (int rep = 0; rep < kRound; ++rep) {
                const int launch = round * kRound + rep;
                psd_hysteresis_kernel<<<hg, 256, 0, stream>>>(
                    b.map, W, H, rep ? b.changed + rep - 1 : nullptr, b.changed + rep,
                    launch ? b.dirty + (size_t)((launch - 1) & 1) * tiles : nullptr,
                    b.dirty + (size_t)(launch & 1) * tiles);
                PSD_CHECK_LAUNCH();
            }
            count_launch(kRound);
            PSD_CUDA(cudaMemcpyAsync(b.changed_host, b.changed + kRound - 1, sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, stream));
            PSD_CUDA(cudaStreamSynchronize(stream));
            if (*b.changed_host == 0) break;
        }
    }
    const int r = ksize / 2;
    const int Wq = (W + 31) / 32;
    const int64_t per_frame = (int64_t)H * Wq;
    dim3 kg((unsigned)((Wq + 31) / 32 * 32 + 255) / 256, (unsigned)H, (unsigned)n);  // 8 warps per block
    kg.x = (unsigned)(((Wq + 31) / 32 + 7) / 8);
    psd_edge_pack_kernel<<<kg, 256, 0, stream>>>(b.map, b.labels, resolve_class, b.bits_in, W, H, Wq);
    PSD_CHECK_LAUNCH();
    psd_edge_dilate_rows_bits_kernel<<<(unsigned)((per_frame * n + 255) / 256), 256, 0, stream>>>(
        b.bits_in, b.bits_row, per_frame * n, Wq, r, (W & 31) ? ((1u << (W & 31)) - 1u) : 0xFFFFFFFFu);
    PSD_CHECK_LAUNCH();
    dim3 cgd((unsigned)((per_frame + 255) / 256), (unsigned)n);
    psd_edge_dilate_cols_bits_kernel<<<cgd, 256, 0, stream>>>(b.bits_row, b.bits_dil, H, Wq, r);
    PSD_CHECK_LAUNCH();
    dim3 sg((unsigned)min((int64_t)64, (per_frame + 255) / 256), (unsigned)n);
    psd_edge_sad_bits_kernel<<<sg, 256, 0, stream>>>(b.bits_dil, b.carry_bits, per_frame, have_prev ? 1 : 0, sums);
    PSD_CHECK_LAUNCH();
    count_launch(4);
    PSD_CUDA(cudaMemcpyAsync(b.carry_bits, b.bits_dil + (int64_t)(n - 1) * per_frame,
                             (size_t)per_frame * 4, cudaMemcpyDeviceToDevice, stream));
    return PSD_OK;
}

}  // namespace psd
