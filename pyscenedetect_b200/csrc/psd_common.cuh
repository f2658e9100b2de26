// Shared helpers for the psd_b200 CUDA translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/psd_b200.h"

namespace psd {

// ---- error plumbing (thread-local message, no exceptions across the ABI) ----
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define PSD_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            psd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                           __LINE__);                                                            \
            return (_e == cudaErrorMemoryAllocation) ? PSD_ERR_OOM : PSD_ERR_CUDA;               \
        }                                                                                        \
    } while (0)

#define PSD_CHECK_LAUNCH()                                                                       \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            psd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                            \
            return PSD_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define PSD_REQUIRE(cond, ...)                                                                   \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            psd::set_error(__VA_ARGS__);                                                         \
            return PSD_ERR_INVALID;                                                              \
        }                                                                                        \
    } while (0)

// ---- fused score pass (score_kernel.cu) ----
struct ScoreArgs {
    const uint8_t* frames;   // n frames, frame_stride apart, tightly packed rows (3*W bytes)
    const uint8_t* prev;     // predecessor of frames[0] or nullptr
    int64_t frame_stride;
    int32_t n_frames;
    int32_t n_pixels;        // W*H
    int32_t chunk_frames;    // frames per time chunk (generic kernel)
    int32_t n_chunks;        // time chunks (the persistent kernel splits the frames into n_chunks near-equal runs)
    uint32_t features;       // PSD_F_* mask of the launch (device-side copy of the template argument)
    int32_t n_strips;
    int32_t tma_ok;          // base pointers and stride 16-byte aligned
    int32_t px_base;         // first pixel this launch covers (strips are relative to it)
    int32_t write_has_prev;  // this launch owns the has_prev flags
    psd_frame_sums* sums;    // [n] (pre-zeroed)
    uint32_t* yhist;         // [n][256] (pre-zeroed) or nullptr
    uint32_t* vhist;         // [n][256] (pre-zeroed) or nullptr
    uint8_t* vplane;         // [n][n_pixels] or nullptr
    uint32_t shift24;        // 0x01000000, passed at run time: the kernel derives a zero the compiler cannot fold from it
};
int launch_score(const ScoreArgs& a, uint32_t features, bool generic_only, cudaStream_t stream);
int score_kernel_smem_bytes();

// ---- resize (resize_kernel.cu) ----
struct ResizeTaps {  // device arrays built on the host exactly as OpenCV builds them
    const int32_t* xofs;  // [dw] source column of tap 0
    const int16_t* xa;    // [dw][2] 11-bit coefficients
    const int32_t* yofs;  // [dh]
    const int16_t* ya;    // [dh][2]
};
int launch_resize(const uint8_t* src, int64_t src_frame_stride, int64_t src_row_pitch, int sw, int sh,
                  uint8_t* dst, int dw, int dh, int64_t n, const ResizeTaps& taps, cudaStream_t stream);

// ---- edge path (edge_kernels.cu) ----
struct EdgeBuffers {
    uint8_t* vplane;    // [n][P] V of HSV (written by the score pass)
    uint32_t* vhist;    // [n][256]
    int32_t* thresholds;// [n][2] low, high
    uint32_t* cand;     // [n][edge_tile_words] Canny candidates (weak or strong pixels), 32 per word, TILE-MAJOR:
                        // tile (ty, tx) = 32 rows x 64 columns = 64 consecutive words, row r at words 2r, 2r+1
    uint32_t* bits_in;  // same layout: strong pixels after classify, the Canny map after hysteresis
    uint32_t* bits_dil; // [n][H][Wq] dilated edges, row-major
    uint32_t* carry_bits; // [H][Wq] dilated edges of the predecessor frame
    uint8_t* tmp;       // [P] scratch for debug taps
    uint8_t* dirty;     // [2][n][tiles] hysteresis: tiles to revisit (double-buffered by round parity)
    int32_t* hyst_flags;// [3] hysteresis: "some tile changed" per round (rotating)
};
int launch_edges(const EdgeBuffers& b, int n, int width, int height, int ksize, bool have_prev,
                 psd_frame_sums* sums, cudaStream_t stream);
int edge_unpack(const uint32_t* bits, uint8_t* out, int W, int H, bool tile_major, cudaStream_t stream);
int64_t edge_tile_words(int W, int H);   // words per frame of a tile-major bit plane

// ---- perceptual hash (hash_kernels.cu) ----
struct HashPlan {        // per-engine tables for one (frame size, hash size, lowpass)
    int n = 0, size = 0; // hash image edge (size * lowpass), low band edge
    int fast = 0;        // both area scale factors are integers
    int area_w = 0, area_h = 0;
    int32_t *xstart = nullptr, *xsi = nullptr, *ystart = nullptr, *ysi = nullptr;
    int32_t *xmid = nullptr;  // [n][2]: table index of the first whole-pixel tap of a destination column, their count
    float *xalpha = nullptr, *ybeta = nullptr;
    double* cosn = nullptr;   // [4n] cos(pi k / 2n)
    int levels = 1, len[8] = {0}, off[8] = {0};  // folded levels of a length-n vector (hash_kernels.cu:FoldPlan)
    float* rowbuf = nullptr;  // [max_batch][H][n] horizontal pass
};
int hash_plan_create(HashPlan* p, int W, int H, int size, int lowpass, int max_batch);
void hash_plan_destroy(HashPlan* p);
int launch_hash(const HashPlan& p, const uint8_t* frames, int64_t frame_stride, int n_frames, int W, int H,
                uint64_t* hashes, cudaStream_t stream);
int launch_hash_dist(const uint64_t* hashes, int64_t n, int size, const uint64_t* prev_hash, double* out,
                     cudaStream_t stream);

// ---- synthetic generator (synth_kernel.cu) ----
int launch_synth(uint8_t* out, const int32_t* d_params, int64_t n, int width, int height,
                 int64_t frame_stride, cudaStream_t stream);

}  // namespace psd
