/*
 * psd_b200.h - C ABI of the B200-native per-frame content-score engine for PySceneDetect.
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json: the
 * process_frame() arithmetic of ContentDetector / AdaptiveDetector / ThresholdDetector /
 * HistogramDetector / HashDetector plus the cv2.resize pre-step SceneManager applies.  The reference is
 * pure Python over cv2/numpy and has no FFI of its own, so every entry
 * point below replaces a cv2/numpy call sequence at the reference line cited; the Python
 * host (pyscenedetect_b200/_capi.py) binds them with ctypes.
 *
 * Conventions: every call returns an int status (PSD_OK == 0, negative = error class);
 * no C++ exception crosses this boundary; output buffers are caller-allocated;
 * psd_last_error() returns a thread-local human-readable message for the last failure.
 * There is NO CPU fallback: without an sm_100 device psd_engine_create fails.
 */
#ifndef PSD_B200_H
#define PSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSD_ABI_VERSION 1

/* status codes */
#define PSD_OK 0
#define PSD_ERR_INVALID (-1) /* bad argument */
#define PSD_ERR_CUDA (-2)    /* CUDA runtime/driver failure (message has the cudaError) */
#define PSD_ERR_OOM (-3)     /* host or device allocation failed */
#define PSD_ERR_STATE (-4)   /* call not valid in the engine's current state */
#define PSD_ERR_NODEVICE (-5)/* no usable sm_100 device */

/* feature mask: which per-frame integer results the fused pass produces */
#define PSD_F_HSV 1u    /* SAD of H,S,V planes vs previous frame: content_detector.py:29-36,155,166-175 */
#define PSD_F_BGRSUM 2u /* sum of all B,G,R bytes: numpy.mean(frame_img), threshold_detector.py:127 */
#define PSD_F_YHIST 4u  /* 256-bin histogram of YUV-Y: histogram_detector.py:156-159 */
#define PSD_F_EDGES 8u  /* Canny+dilate edge-map SAD: content_detector.py:213-239 (implies HSV) */
#define PSD_F_HASH 16u  /* perceptual hash of every frame: hash_detector.py:124-158 */
#define PSD_HASH_WORDS 4 /* a hash occupies 4 x uint64 (size * size <= 256 bits), bit u*size+v = D[u][v] > median */

/* engine flags (psd_config.flags) */
#define PSD_CFG_GENERIC_KERNEL 1u /* score every strip with the generic kernel instead of the persistent
                                     warp-specialised one (same results; exists so tests can cross-check the two) */

/* submit flags */
#define PSD_SUBMIT_PINNED 1u /* host buffer is page-locked (psd_host_alloc): DMA straight from it,
                                caller keeps it unchanged until psd_engine_sync() */

typedef struct psd_engine psd_engine;

/* Engine configuration.  (src_width,src_height) is the size of the frames submitted;
 * (width,height) the size the detectors score at.  If they differ the engine applies the
 * exact cv2.resize(..., INTER_LINEAR) fixed-point bilinear of scene_manager.py:670-678. */
typedef struct psd_config {
    int32_t struct_size; /* sizeof(psd_config) */
    int32_t device;      /* CUDA ordinal */
    int32_t src_width, src_height;
    int32_t width, height;
    uint32_t features;        /* PSD_F_* */
    int32_t edge_kernel_size; /* dilate kernel k (odd >= 3); 0 = content_detector.py:39-46 estimate */
    int32_t max_batch;        /* max frames per submit call (staging is sized for it) */
    uint32_t flags;           /* PSD_CFG_* */
    int32_t hash_size;        /* PSD_F_HASH: HashDetector(size=...), 0 = 8 */
    int32_t hash_lowpass;     /* PSD_F_HASH: HashDetector(lowpass=...), 0 = 2 */
    int32_t reserved[4];
} psd_config;

/* Per-frame integer results (device- and host-side layout, 64 bytes). */
typedef struct psd_frame_sums {
    uint64_t sad_hue;   /* sum |H_t - H_{t-1}|            */
    uint64_t sad_sat;   /* sum |S_t - S_{t-1}|            */
    uint64_t sad_lum;   /* sum |V_t - V_{t-1}|            */
    uint64_t sad_edges; /* sum |E_t - E_{t-1}|, E in {0,255} */
    uint64_t bgr_sum;   /* sum of all 3*W*H bytes          */
    uint64_t has_prev;  /* 0 for the first frame of a stream (content_detector.py:161-164) */
    uint64_t reserved[2];
} psd_frame_sums;

/* ---- library ---- */
int psd_abi_version(void);
const char* psd_version(void);
const char* psd_last_error(void);
int psd_device_count(void);
/* name_out may be NULL; fills compute capability, SM count and total memory. */
int psd_device_info(int device, char* name_out, size_t name_cap, int* cc_major, int* cc_minor,
                    int* sm_count, uint64_t* total_mem);
/* PCI bus id "0000:3b:00.0" of a device (lets a host place page-locked buffers on the GPU's NUMA node) */
int psd_device_pci_bus_id(int device, char* out, size_t cap);
/* total kernel launches issued by this library since load (for bench.py's gpu_launches). */
uint64_t psd_launch_count(void);

/* page-locked host memory for zero-staging submits */
int psd_host_alloc(size_t bytes, void** out);
int psd_host_free(void* p);
/* plain device memory + copies, so a torch-free host can keep batches resident in HBM */
int psd_device_alloc(int device, size_t bytes, void** out);
int psd_device_free(int device, void* p);
int psd_memcpy_h2d(int device, void* dst, const void* src, size_t bytes);
int psd_memcpy_d2h(int device, void* dst, const void* src, size_t bytes);

/* ---- engine: replaces the per-frame cv2/numpy work of detector.process_frame() ---- */
int psd_engine_create(const psd_config* cfg, psd_engine** out);
void psd_engine_destroy(psd_engine* e);
/* forget all frames and the carried previous frame (detector re-use on a new video) */
int psd_engine_reset(psd_engine* e);
/* Set the predecessor of the NEXT submitted frame (the one-frame halo of a time shard).
 * Source-size BGR24.  host variant copies before returning. */
int psd_engine_set_halo_host(psd_engine* e, const uint8_t* bgr, int64_t row_pitch);
int psd_engine_set_halo_device(psd_engine* e, const void* dptr);
/* Score n frames of host memory.  frame_stride / row_pitch in bytes (numpy strides:
 * the SceneManager crop view of scene_manager.py:666-668 is non-contiguous).  Without
 * PSD_SUBMIT_PINNED the bytes are copied to internal page-locked staging before return. */
int psd_engine_submit_host(psd_engine* e, const uint8_t* bgr, int64_t n_frames,
                           int64_t frame_stride, int64_t row_pitch, uint32_t flags);
/* Score n tightly packed frames already resident in HBM (row pitch = 3*src_width).
 * The memory must stay valid until psd_engine_sync(). */
int psd_engine_submit_device(psd_engine* e, const void* dptr, int64_t n_frames,
                             int64_t frame_stride);
int psd_engine_sync(psd_engine* e);
/* the engine's compute stream (cudaStream_t): launch the psd_scan_* kernels (or record events)
 * on it to stay ordered after the engine's own kernels */
void* psd_engine_compute_stream(psd_engine* e);
int64_t psd_engine_frame_count(const psd_engine* e);
/* copy results for frames [first, first+n) to host (implies sync) */
int psd_engine_read_sums(psd_engine* e, int64_t first, int64_t n, psd_frame_sums* out);
int psd_engine_read_yhist(psd_engine* e, int64_t first, int64_t n, uint32_t* out /*[n][256]*/);
int psd_engine_read_hash(psd_engine* e, int64_t first, int64_t n, uint64_t* out /*[n][PSD_HASH_WORDS]*/);
/* device pointers of the engine-owned result arrays (valid until destroy/reset) */
int psd_engine_device_results(psd_engine* e, const psd_frame_sums** sums, const uint32_t** yhist);
int psd_engine_device_hash(psd_engine* e, const uint64_t** hashes /* stream frame i at hashes + i*PSD_HASH_WORDS;
                                                                      the halo frame's hash sits one entry before */);
/* CUDA-event time (ms) spent in the engine's kernels between the first launch after the last
 * psd_engine_timing_reset() and the last launch (on the engine's compute stream). */
int psd_engine_timing_reset(psd_engine* e);
int psd_engine_timing_ms(psd_engine* e, float* total_ms, float* score_kernel_ms,
                         uint64_t* score_kernel_launches);
/* effective dilate kernel size used for the edge component */
int psd_engine_edge_kernel_size(const psd_engine* e);
/* debug/test taps: copy intermediate planes of frame `index` of the LAST submitted batch.
 * which: 0 = scored-size BGR (after resize), 1 = V plane, 2 = Canny map (0/255), 3 = dilated edges */
int psd_engine_debug_plane(psd_engine* e, int which, int64_t index, uint8_t* out, size_t cap);

/* ---- trailing device scans over result arrays (all pointers are DEVICE pointers unless the
 *      name says host; `stream` is a cudaStream_t or NULL) ---- */
/* content_detector.py:166-180: components = sad / float(W*H); content_val = sum(c*w)/sum(|w|);
 * out_components[n][4], out_content_val[n]; frames with has_prev == 0 get 0.0 */
int psd_scan_content(const psd_frame_sums* sums, int64_t n, int64_t n_pixels, const double weights[4],
                     double weight_abs_sum /* sum(abs(w)) as the host computed it */,
                     double* out_components, double* out_content_val, void* stream);
/* adaptive_detector.py:100-143: ratio for target i uses scores[i-w .. i+w]; out_ratio[i] is NaN
 * where the window is incomplete.  scores[] is the content_val array incl. frame 0's 0.0. */
int psd_scan_adaptive(const double* scores, int64_t n, int32_t window_width, double min_content_val,
                      double* out_ratio, void* stream);
/* threshold_detector.py:127: average_rgb = bgr_sum / (3*W*H) */
int psd_scan_average(const psd_frame_sums* sums, int64_t n, int64_t n_values, double* out_avg,
                     void* stream);
/* histogram_detector.py:98,159-163: rebin 256 -> bins, L2-normalise to float32 as cv2.normalize,
 * HISTCMP_CORREL in fp64 against the previous frame; out_correl[0] (no predecessor in the
 * array) uses prev_hist if non-NULL else is NaN. */
int psd_scan_hist_correl(const uint32_t* yhist, int64_t n, int32_t bins, const uint32_t* prev_hist,
                         double* out_correl, void* stream);
/* hash_detector.py:95-99: hash_dist = popcount(hash_t xor hash_{t-1}) / (size*size); out[0] uses prev_hash if
 * non-NULL else is NaN */
int psd_scan_hash_dist(const uint64_t* hashes, int64_t n, int32_t hash_size, const uint64_t* prev_hash,
                       double* out_dist, void* stream);
/* >= / <= compare producing u8 flags (content_detector.py:210, histogram_detector.py:108) */
int psd_scan_compare(const double* values, int64_t n, double threshold, int32_t op /*0: >=, 1: <=, 2: <*/,
                     uint8_t* out_flags, void* stream);
/* ---- cut state machines on the device.  Frame-number domain, constant frame
 *      rate: `min_frames` is the host's conversion of min_scene_len (common.py:480-486,627-638).
 *      cuts[cap] receives absolute frame numbers (first_frame + index), *count the number found
 *      (may exceed cap: only cap are stored).  All pointers are DEVICE pointers. ---- */
/* detector.py:160-224 FlashFilter over the `score >= threshold` flags; mode 0 = MERGE, 1 = SUPPRESS */
int psd_cuts_flash_filter(const uint8_t* above, int64_t n, int64_t first_frame, int64_t min_frames,
                          int32_t mode, int64_t* cuts, int32_t* count, int32_t cap, void* stream);
/* adaptive_detector.py:134-143 (ratio from psd_scan_adaptive, NaN where the window is incomplete) */
int psd_cuts_adaptive(const double* ratio, const double* score, int64_t n, int64_t first_frame,
                      int32_t window_width, double adaptive_threshold, double min_content_val,
                      int64_t min_frames, int64_t* cuts, int32_t* count, int32_t cap, void* stream);
/* histogram_detector.py:87-112: cut where correl <= threshold and min_frames since the last cut */
int psd_cuts_histogram(const double* correl, int64_t n, int64_t first_frame, double threshold,
                       int64_t min_frames, int64_t* cuts, int32_t* count, int32_t cap, void* stream);
/* hash_detector.py:104-109: cut where dist >= threshold and min_frames since the last cut (NaN = no predecessor) */
int psd_cuts_hash(const double* dist, int64_t n, int64_t first_frame, double threshold, int64_t min_frames,
                  int64_t* cuts, int32_t* count, int32_t cap, void* stream);
/* threshold_detector.py:113-191: fade in/out automaton incl. the post_process final cut */
int psd_cuts_threshold(const double* average, int64_t n, int64_t first_frame, double threshold,
                       int32_t method_ceiling, double fade_bias, int64_t min_frames, int32_t add_final_scene,
                       int64_t* cuts, int32_t* count, int32_t cap, void* stream);

/* host-convenience wrappers: engine-owned sums -> host arrays (numpy), implies sync */
int psd_engine_scan_content_host(psd_engine* e, int64_t first, int64_t n, const double weights[4],
                                 double weight_abs_sum, double* out_components,
                                 double* out_content_val);
int psd_engine_scan_adaptive_host(psd_engine* e, const double* scores_host, int64_t n,
                                  int32_t window_width, double min_content_val, double* out_ratio);
int psd_engine_scan_average_host(psd_engine* e, int64_t first, int64_t n, double* out_avg);
int psd_engine_scan_hist_correl_host(psd_engine* e, int64_t first, int64_t n, int32_t bins,
                                     double* out_correl);
int psd_engine_scan_hash_dist_host(psd_engine* e, int64_t first, int64_t n, double* out_dist);

/* ---- synthetic input generator (bench.py / tests; pyscenedetect_b200/synth.py bit-exact twin) ---- */
/* params_host: [n][24] int32 rows of ScenePlan.params for frames first..first+n-1 */
int psd_synth_frames(int device, void* d_out, const int32_t* params_host, int64_t n, int32_t width,
                     int32_t height, int64_t frame_stride, void* stream);

/* ---- test hooks ---- */
/* device BGR (n pixels) -> H,S,V planes with the device functions the fused pass uses:
 * variant 7 = the warp-specialised kernel's arithmetic, 2 = the generic kernel's */
int psd_test_hsv(int device, const uint8_t* bgr_host, int64_t n_pixels, uint8_t* h_out, uint8_t* s_out,
                 uint8_t* v_out, uint8_t* y_out, int variant);

#ifdef __cplusplus
}
#endif
#endif /* PSD_B200_H */
