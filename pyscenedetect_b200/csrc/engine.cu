// psd_engine: the stateful C-ABI object behind the detectors' process_frame().
// Owns page-locked staging, device staging, the carried previous frame (the one-frame halo),
// per-frame result arrays in HBM, two streams (copy / compute) and the event plumbing that
// overlaps the H2D of batch k+1 with the kernels of batch k.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "psd_common.cuh"

namespace psd {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// OpenCV resize.cpp tap generation for INTER_LINEAR (float32 coefficient math, 11-bit fixed point).
static void build_taps(int src, int dst, std::vector<int32_t>& ofs, std::vector<int16_t>& coef) {
    ofs.resize(dst);
    coef.resize(2 * (size_t)dst);
    const double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        ofs[d] = s;
        const float c0 = (1.f - f) * 2048.f, c1 = f * 2048.f;
        coef[2 * d] = (int16_t)lrintf(c0);  // cvRound: round half to even (default FE_TONEAREST)
        coef[2 * d + 1] = (int16_t)lrintf(c1);
    }
}

}  // namespace psd

using namespace psd;

struct psd_engine {
    psd_config cfg{};
    int device = 0;
    int sw = 0, sh = 0, W = 0, H = 0;
    int64_t src_frame_bytes = 0, frame_bytes = 0, P = 0;
    bool resize = false;
    uint32_t features = 0;
    int ksize = 0;
    int max_batch = 0;
    bool generic_only = false;  // PSD_CFG_GENERIC_KERNEL: score with the generic kernel only (cross-check)
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    // staging (double buffered)
    uint8_t* pinned[2] = {nullptr, nullptr};
    uint8_t* dev_stage[2] = {nullptr, nullptr};
    cudaEvent_t slot_free[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};
    int next_slot = 0;
    // scored-size frames when resizing
    uint8_t* small = nullptr;
    int32_t* d_xofs = nullptr; int16_t* d_xa = nullptr; int32_t* d_yofs = nullptr; int16_t* d_ya = nullptr;
    // carry (predecessor of the next frame, scored size)
    uint8_t* carry = nullptr;
    bool have_carry = false;
    // results: slot 0 = halo frame, stream frame i at slot i+1
    psd_frame_sums* d_sums = nullptr;
    uint32_t* d_yhist = nullptr;
    uint64_t* d_hash = nullptr;   // [capacity][PSD_HASH_WORDS]
    HashPlan hash{};
    int64_t capacity = 0;
    int64_t n_frames = 0;
    bool halo_scored = false;
    // edge path
    EdgeBuffers eb{};
    // last batch bookkeeping for debug taps
    const uint8_t* last_scored = nullptr;
    int64_t last_scored_stride = 0;
    int64_t last_n = 0;
    // timing
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_score, ev_total;
    size_t ev_next = 0;
};

static int ensure_capacity(psd_engine* e, int64_t need_slots) {
    if (need_slots <= e->capacity) return PSD_OK;
    int64_t cap = e->capacity ? e->capacity : 4096;
    while (cap < need_slots) cap *= 2;
    psd_frame_sums* ns = nullptr;
    PSD_CUDA(cudaMalloc(&ns, (size_t)cap * sizeof(psd_frame_sums)));
    PSD_CUDA(cudaMemsetAsync(ns, 0, (size_t)cap * sizeof(psd_frame_sums), e->compute_stream));
    if (e->d_sums) {
        PSD_CUDA(cudaMemcpyAsync(ns, e->d_sums, (size_t)(e->n_frames + 1) * sizeof(psd_frame_sums),
                                 cudaMemcpyDeviceToDevice, e->compute_stream));
        PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
        cudaFree(e->d_sums);
    }
    e->d_sums = ns;
    if (e->features & PSD_F_YHIST) {
        uint32_t* nh = nullptr;
        PSD_CUDA(cudaMalloc(&nh, (size_t)cap * 256 * sizeof(uint32_t)));
        PSD_CUDA(cudaMemsetAsync(nh, 0, (size_t)cap * 256 * sizeof(uint32_t), e->compute_stream));
        if (e->d_yhist) {
            PSD_CUDA(cudaMemcpyAsync(nh, e->d_yhist, (size_t)(e->n_frames + 1) * 256 * sizeof(uint32_t),
                                     cudaMemcpyDeviceToDevice, e->compute_stream));
            PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
            cudaFree(e->d_yhist);
        }
        e->d_yhist = nh;
    }
    if (e->features & PSD_F_HASH) {
        uint64_t* nh = nullptr;
        PSD_CUDA(cudaMalloc(&nh, (size_t)cap * PSD_HASH_WORDS * sizeof(uint64_t)));
        PSD_CUDA(cudaMemsetAsync(nh, 0, (size_t)cap * PSD_HASH_WORDS * sizeof(uint64_t), e->compute_stream));
        if (e->d_hash) {
            PSD_CUDA(cudaMemcpyAsync(nh, e->d_hash, (size_t)(e->n_frames + 1) * PSD_HASH_WORDS * sizeof(uint64_t),
                                     cudaMemcpyDeviceToDevice, e->compute_stream));
            PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
            cudaFree(e->d_hash);
        }
        e->d_hash = nh;
    }
    e->capacity = cap;
    return PSD_OK;
}

static cudaEvent_t next_event(psd_engine* e) {
    if (e->ev_next == e->ev_pool.size()) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreate(&ev) != cudaSuccess) return nullptr;
        e->ev_pool.push_back(ev);
    }
    return e->ev_pool[e->ev_next++];
}

// Score `n` tightly-packed-row frames at `src` (source size) that are visible to compute_stream.
// slot0: result slot of the first frame (0 = halo slot).
static int run_batch(psd_engine* e, const uint8_t* src, int64_t src_frame_stride, int64_t n,
                     int64_t slot0, bool is_halo) {
    cudaStream_t st = e->compute_stream;
    cudaEvent_t t0 = next_event(e), t1 = next_event(e), k0 = next_event(e), k1 = next_event(e);
    if (!t0 || !t1 || !k0 || !k1) { set_error("cudaEventCreate failed"); return PSD_ERR_CUDA; }
    PSD_CUDA(cudaEventRecord(t0, st));
    const uint8_t* scored = src;
    int64_t scored_stride = src_frame_stride;
    if (e->resize) {
        ResizeTaps taps{e->d_xofs, e->d_xa, e->d_yofs, e->d_ya};
        int rc = launch_resize(src, src_frame_stride, (int64_t)e->sw * 3, e->sw, e->sh, e->small, e->W,
                               e->H, n, taps, st);
        if (rc) return rc;
        scored = e->small;
        scored_stride = e->frame_bytes;
    }
    PSD_CUDA(cudaMemsetAsync(e->d_sums + slot0, 0, (size_t)n * sizeof(psd_frame_sums), st));
    if (e->features & PSD_F_YHIST)
        PSD_CUDA(cudaMemsetAsync(e->d_yhist + slot0 * 256, 0, (size_t)n * 256 * sizeof(uint32_t), st));
    if (e->features & PSD_F_EDGES)
        PSD_CUDA(cudaMemsetAsync(e->eb.vhist, 0, (size_t)n * 256 * sizeof(uint32_t), st));
    ScoreArgs a{};
    a.frames = scored;
    a.prev = (e->have_carry && !is_halo) ? e->carry : nullptr;
    a.frame_stride = scored_stride;
    a.n_frames = (int32_t)n;
    a.n_pixels = (int32_t)e->P;
    a.chunk_frames = 0;  // launch_score picks the time-chunk length
    a.sums = e->d_sums + slot0;
    a.yhist = (e->features & PSD_F_YHIST) ? e->d_yhist + slot0 * 256 : nullptr;
    a.vhist = (e->features & PSD_F_EDGES) ? e->eb.vhist : nullptr;
    a.vplane = (e->features & PSD_F_EDGES) ? e->eb.vplane : nullptr;
    PSD_CUDA(cudaEventRecord(k0, st));
    int rc = PSD_OK;
    if (e->features & 15u) {  // the fused pass (HSV / byte sum / Y histogram / edges)
        rc = launch_score(a, e->features & 15u, e->generic_only, st);
        if (rc) return rc;
    }
    if (e->features & PSD_F_HASH) {
        rc = launch_hash(e->hash, scored, scored_stride, (int)n, e->W, e->H, e->d_hash + slot0 * PSD_HASH_WORDS, st);
        if (rc) return rc;
    }
    PSD_CUDA(cudaEventRecord(k1, st));
    e->ev_score.push_back({k0, k1});
    if (e->features & PSD_F_EDGES) {
        rc = launch_edges(e->eb, (int)n, e->W, e->H, e->ksize, a.prev != nullptr, e->d_sums + slot0, st);
        if (rc) return rc;
    }
    // carry the last frame (scored size) for the next batch
    PSD_CUDA(cudaMemcpyAsync(e->carry, scored + (n - 1) * scored_stride, (size_t)e->frame_bytes,
                             cudaMemcpyDeviceToDevice, st));
    e->have_carry = true;
    e->last_scored = scored;
    e->last_scored_stride = scored_stride;
    e->last_n = n;
    PSD_CUDA(cudaEventRecord(t1, st));
    e->ev_total.push_back({t0, t1});
    return PSD_OK;
}

extern "C" {

int psd_abi_version(void) { return PSD_ABI_VERSION; }
const char* psd_version(void) { return "psd_b200 0.1.0 (sm_100a)"; }
const char* psd_last_error(void) { return g_err; }
uint64_t psd_launch_count(void) { return g_launches.load(); }

int psd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int psd_device_info(int device, char* name_out, size_t name_cap, int* cc_major, int* cc_minor,
                    int* sm_count, uint64_t* total_mem) {
    cudaDeviceProp p{};
    PSD_CUDA(cudaGetDeviceProperties(&p, device));
    if (name_out && name_cap) { strncpy(name_out, p.name, name_cap - 1); name_out[name_cap - 1] = 0; }
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = (uint64_t)p.totalGlobalMem;
    return PSD_OK;
}

int psd_device_pci_bus_id(int device, char* out, size_t cap) {
    PSD_REQUIRE(out && cap >= 16, "psd_device_pci_bus_id: buffer too small");
    PSD_CUDA(cudaDeviceGetPCIBusId(out, (int)cap, device));
    return PSD_OK;
}

int psd_host_alloc(size_t bytes, void** out) {
    PSD_REQUIRE(out && bytes > 0, "psd_host_alloc: bad args");
    PSD_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return PSD_OK;
}
int psd_host_free(void* p) {
    if (p) PSD_CUDA(cudaFreeHost(p));
    return PSD_OK;
}
int psd_device_alloc(int device, size_t bytes, void** out) {
    PSD_REQUIRE(out && bytes > 0, "psd_device_alloc: bad args");
    PSD_CUDA(cudaSetDevice(device));
    PSD_CUDA(cudaMalloc(out, bytes));
    return PSD_OK;
}
int psd_device_free(int device, void* p) {
    PSD_CUDA(cudaSetDevice(device));
    if (p) PSD_CUDA(cudaFree(p));
    return PSD_OK;
}
int psd_memcpy_h2d(int device, void* dst, const void* src, size_t bytes) {
    PSD_CUDA(cudaSetDevice(device));
    PSD_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return PSD_OK;
}
int psd_memcpy_d2h(int device, void* dst, const void* src, size_t bytes) {
    PSD_CUDA(cudaSetDevice(device));
    PSD_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return PSD_OK;
}

void psd_engine_destroy(psd_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->compute_stream) cudaStreamSynchronize(e->compute_stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    for (int s = 0; s < 2; ++s) {
        if (e->pinned[s]) cudaFreeHost(e->pinned[s]);
        if (e->dev_stage[s]) cudaFree(e->dev_stage[s]);
        if (e->slot_free[s]) cudaEventDestroy(e->slot_free[s]);
        if (e->h2d_done[s]) cudaEventDestroy(e->h2d_done[s]);
    }
    cudaFree(e->small); cudaFree(e->d_xofs); cudaFree(e->d_xa); cudaFree(e->d_yofs); cudaFree(e->d_ya);
    cudaFree(e->carry); cudaFree(e->d_sums); cudaFree(e->d_yhist); cudaFree(e->d_hash);
    hash_plan_destroy(&e->hash);
    cudaFree(e->eb.vplane); cudaFree(e->eb.vhist); cudaFree(e->eb.thresholds); cudaFree(e->eb.cand);
    cudaFree(e->eb.tmp); cudaFree(e->eb.bits_in); cudaFree(e->eb.bits_dil);
    cudaFree(e->eb.carry_bits); cudaFree(e->eb.dirty); cudaFree(e->eb.hyst_flags);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->compute_stream) cudaStreamDestroy(e->compute_stream);
    delete e;
}

#define ENG_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            psd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                           __LINE__);                                                           \
            psd_engine_destroy(e);                                                              \
            return (_e == cudaErrorMemoryAllocation) ? PSD_ERR_OOM : PSD_ERR_CUDA;              \
        }                                                                                       \
    } while (0)

int psd_engine_create(const psd_config* cfg, psd_engine** out) {
    PSD_REQUIRE(cfg && out, "psd_engine_create: null argument");
    PSD_REQUIRE(cfg->struct_size == (int32_t)sizeof(psd_config), "psd_config.struct_size mismatch");
    PSD_REQUIRE(cfg->src_width > 0 && cfg->src_height > 0 && cfg->width > 0 && cfg->height > 0,
                "frame sizes must be positive");
    PSD_REQUIRE((int64_t)cfg->src_width * cfg->src_height < (1LL << 30), "frame too large");
    PSD_REQUIRE(cfg->features != 0 && (cfg->features & ~31u) == 0, "bad feature mask 0x%x", cfg->features);
    PSD_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch <= 4096, "max_batch must be in [1,4096]");
    PSD_REQUIRE(cfg->edge_kernel_size == 0 || (cfg->edge_kernel_size >= 3 && (cfg->edge_kernel_size & 1)),
                "kernel_size must be odd integer >= 3");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible (this engine has no CPU fallback)");
        return PSD_ERR_NODEVICE;
    }
    PSD_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "device %d out of range (%d visible)", cfg->device, ndev);
    cudaDeviceProp prop{};
    PSD_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major,
                  prop.minor);
        return PSD_ERR_NODEVICE;
    }
    PSD_CUDA(cudaSetDevice(cfg->device));
    psd_engine* e = new (std::nothrow) psd_engine();
    if (!e) { set_error("out of host memory"); return PSD_ERR_OOM; }
    e->cfg = *cfg;
    e->device = cfg->device;
    e->sw = cfg->src_width; e->sh = cfg->src_height; e->W = cfg->width; e->H = cfg->height;
    e->resize = (e->sw != e->W) || (e->sh != e->H);
    e->P = (int64_t)e->W * e->H;
    e->frame_bytes = e->P * 3;
    e->src_frame_bytes = (int64_t)e->sw * e->sh * 3;
    e->features = cfg->features | ((cfg->features & PSD_F_EDGES) ? PSD_F_HSV : 0);
    e->max_batch = cfg->max_batch;
    e->generic_only = (cfg->flags & PSD_CFG_GENERIC_KERNEL) != 0;
    if (e->features & PSD_F_EDGES) {
        int k = cfg->edge_kernel_size;
        if (k == 0) {  // content_detector.py:39-46; Python round() is half-to-even like nearbyint
            k = 4 + (int)nearbyint(sqrt((double)e->W * (double)e->H) / 192.0);
            if ((k & 1) == 0) k += 1;
        }
        if (k > 63) {  // the bit-plane dilation shifts words by at most 31 bits
            psd_engine_destroy(e);
            PSD_REQUIRE(false, "edge kernel size %d is not supported (odd sizes 3 .. 63)", k);
        }
        e->ksize = k;
    }
    if (e->features & PSD_F_HASH) {
        int rc = hash_plan_create(&e->hash, e->W, e->H, cfg->hash_size ? cfg->hash_size : 8,
                                  cfg->hash_lowpass ? cfg->hash_lowpass : 2, e->max_batch);
        if (rc) { psd_engine_destroy(e); return rc; }
    }
    ENG_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    ENG_CUDA(cudaStreamCreateWithFlags(&e->compute_stream, cudaStreamNonBlocking));
    for (int s = 0; s < 2; ++s) {
        ENG_CUDA(cudaEventCreateWithFlags(&e->slot_free[s], cudaEventDisableTiming));
        ENG_CUDA(cudaEventCreateWithFlags(&e->h2d_done[s], cudaEventDisableTiming));
    }
    ENG_CUDA(cudaMalloc(&e->carry, (size_t)e->frame_bytes));
    if (e->resize) {
        ENG_CUDA(cudaMalloc(&e->small, (size_t)e->frame_bytes * e->max_batch));
        std::vector<int32_t> xo, yo; std::vector<int16_t> xa, ya;
        build_taps(e->sw, e->W, xo, xa);
        build_taps(e->sh, e->H, yo, ya);
        ENG_CUDA(cudaMalloc(&e->d_xofs, xo.size() * 4)); ENG_CUDA(cudaMalloc(&e->d_xa, xa.size() * 2));
        ENG_CUDA(cudaMalloc(&e->d_yofs, yo.size() * 4)); ENG_CUDA(cudaMalloc(&e->d_ya, ya.size() * 2));
        ENG_CUDA(cudaMemcpy(e->d_xofs, xo.data(), xo.size() * 4, cudaMemcpyHostToDevice));
        ENG_CUDA(cudaMemcpy(e->d_xa, xa.data(), xa.size() * 2, cudaMemcpyHostToDevice));
        ENG_CUDA(cudaMemcpy(e->d_yofs, yo.data(), yo.size() * 4, cudaMemcpyHostToDevice));
        ENG_CUDA(cudaMemcpy(e->d_ya, ya.data(), ya.size() * 2, cudaMemcpyHostToDevice));
    }
    if (e->features & PSD_F_EDGES) {
        const size_t plane = (size_t)e->P * e->max_batch;
        ENG_CUDA(cudaMalloc(&e->eb.vplane, plane));
        const size_t words = (size_t)e->H * ((e->W + 31) / 32);
        // the two planes of the hysteresis are tile-major and padded to whole 64 x 32 tiles; the padding is never
        // written, so it is zeroed once here
        const size_t tiled = (size_t)edge_tile_words(e->W, e->H) * 4 * e->max_batch;
        ENG_CUDA(cudaMalloc(&e->eb.cand, tiled));
        ENG_CUDA(cudaMalloc(&e->eb.bits_in, tiled));
        ENG_CUDA(cudaMemset(e->eb.cand, 0, tiled));
        ENG_CUDA(cudaMemset(e->eb.bits_in, 0, tiled));
        ENG_CUDA(cudaMalloc(&e->eb.tmp, (size_t)e->P));
        ENG_CUDA(cudaMalloc(&e->eb.bits_dil, words * 4 * e->max_batch));
        ENG_CUDA(cudaMalloc(&e->eb.carry_bits, words * 4));
        ENG_CUDA(cudaMalloc(&e->eb.vhist, (size_t)e->max_batch * 256 * 4));
        ENG_CUDA(cudaMalloc(&e->eb.thresholds, (size_t)e->max_batch * 2 * 4));
        ENG_CUDA(cudaMalloc(&e->eb.hyst_flags, 64));
        const size_t n_tiles = (size_t)e->max_batch * ((e->W + 63) / 64) * ((e->H + 31) / 32);
        ENG_CUDA(cudaMalloc(&e->eb.dirty, 2 * n_tiles));
    }
    {
        int rc = ensure_capacity(e, 4096);
        if (rc) { psd_engine_destroy(e); return rc; }
    }
    *out = e;
    return PSD_OK;
}

int psd_engine_reset(psd_engine* e) {
    PSD_REQUIRE(e, "null engine");
    PSD_CUDA(cudaSetDevice(e->device));
    PSD_CUDA(cudaStreamSynchronize(e->copy_stream));
    PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
    e->n_frames = 0;
    e->have_carry = false;
    e->halo_scored = false;
    e->last_n = 0;
    e->ev_next = 0; e->ev_score.clear(); e->ev_total.clear();
    return PSD_OK;
}

static int ensure_staging(psd_engine* e) {
    for (int s = 0; s < 2; ++s) {
        if (!e->dev_stage[s]) PSD_CUDA(cudaMalloc(&e->dev_stage[s], (size_t)e->src_frame_bytes * e->max_batch));
    }
    return PSD_OK;
}
static int ensure_pinned(psd_engine* e) {
    for (int s = 0; s < 2; ++s) {
        if (!e->pinned[s])
            PSD_CUDA(cudaHostAlloc((void**)&e->pinned[s], (size_t)e->src_frame_bytes * e->max_batch,
                                   cudaHostAllocDefault));
    }
    return PSD_OK;
}

// copy host frames (arbitrary strides) -> device staging slot, tightly packed
static int stage_host(psd_engine* e, const uint8_t* bgr, int64_t n, int64_t frame_stride,
                      int64_t row_pitch, uint32_t flags, int slot) {
    const int64_t tight_row = (int64_t)e->sw * 3;
    PSD_CUDA(cudaEventSynchronize(e->slot_free[slot]));
    const uint8_t* src = bgr;
    int64_t fs = frame_stride, rp = row_pitch;
    if (!(flags & PSD_SUBMIT_PINNED)) {
        int rc = ensure_pinned(e);
        if (rc) return rc;
        uint8_t* dst = e->pinned[slot];
        if (rp == tight_row && fs == e->src_frame_bytes) {
            memcpy(dst, bgr, (size_t)(n * e->src_frame_bytes));
        } else {
            for (int64_t f = 0; f < n; ++f)
                for (int y = 0; y < e->sh; ++y)
                    memcpy(dst + f * e->src_frame_bytes + (int64_t)y * tight_row,
                           bgr + f * frame_stride + (int64_t)y * row_pitch, (size_t)tight_row);
        }
        src = dst; fs = e->src_frame_bytes; rp = tight_row;
    }
    if (rp == tight_row && fs == e->src_frame_bytes) {
        PSD_CUDA(cudaMemcpyAsync(e->dev_stage[slot], src, (size_t)(n * e->src_frame_bytes),
                                 cudaMemcpyHostToDevice, e->copy_stream));
    } else {
        for (int64_t f = 0; f < n; ++f)
            PSD_CUDA(cudaMemcpy2DAsync(e->dev_stage[slot] + f * e->src_frame_bytes, (size_t)tight_row,
                                       src + f * fs, (size_t)rp, (size_t)tight_row, (size_t)e->sh,
                                       cudaMemcpyHostToDevice, e->copy_stream));
    }
    PSD_CUDA(cudaEventRecord(e->h2d_done[slot], e->copy_stream));
    PSD_CUDA(cudaStreamWaitEvent(e->compute_stream, e->h2d_done[slot], 0));
    return PSD_OK;
}

int psd_engine_set_halo_device(psd_engine* e, const void* dptr) {
    PSD_REQUIRE(e && dptr, "psd_engine_set_halo_device: null argument");
    PSD_REQUIRE(e->n_frames == 0, "halo must be set before the first frame is submitted");
    PSD_CUDA(cudaSetDevice(e->device));
    e->have_carry = false;
    int rc = run_batch(e, (const uint8_t*)dptr, e->src_frame_bytes, 1, 0, true);
    if (rc) return rc;
    e->halo_scored = true;
    return PSD_OK;
}

int psd_engine_set_halo_host(psd_engine* e, const uint8_t* bgr, int64_t row_pitch) {
    PSD_REQUIRE(e && bgr, "psd_engine_set_halo_host: null argument");
    PSD_REQUIRE(e->n_frames == 0, "halo must be set before the first frame is submitted");
    PSD_REQUIRE(row_pitch >= (int64_t)e->sw * 3, "row_pitch smaller than a row");
    PSD_CUDA(cudaSetDevice(e->device));
    int rc = ensure_staging(e);
    if (rc) return rc;
    const int slot = e->next_slot;
    e->next_slot ^= 1;
    rc = stage_host(e, bgr, 1, row_pitch * e->sh, row_pitch, 0, slot);
    if (rc) return rc;
    rc = psd_engine_set_halo_device(e, e->dev_stage[slot]);
    if (rc) return rc;
    PSD_CUDA(cudaEventRecord(e->slot_free[slot], e->compute_stream));
    return PSD_OK;
}

int psd_engine_submit_device(psd_engine* e, const void* dptr, int64_t n, int64_t frame_stride) {
    PSD_REQUIRE(e && dptr, "psd_engine_submit_device: null argument");
    PSD_REQUIRE(n >= 0, "negative frame count");
    PSD_REQUIRE(frame_stride >= e->src_frame_bytes, "frame_stride smaller than a frame");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    int rc = ensure_capacity(e, e->n_frames + n + 1);
    if (rc) return rc;
    const uint8_t* p = (const uint8_t*)dptr;
    int64_t done = 0;
    while (done < n) {
        const int64_t b = (n - done < e->max_batch) ? (n - done) : e->max_batch;
        rc = run_batch(e, p + done * frame_stride, frame_stride, b, e->n_frames + 1, false);
        if (rc) return rc;
        e->n_frames += b;
        done += b;
    }
    return PSD_OK;
}

int psd_engine_submit_host(psd_engine* e, const uint8_t* bgr, int64_t n, int64_t frame_stride,
                           int64_t row_pitch, uint32_t flags) {
    PSD_REQUIRE(e && bgr, "psd_engine_submit_host: null argument");
    PSD_REQUIRE(n >= 0, "negative frame count");
    PSD_REQUIRE(row_pitch >= (int64_t)e->sw * 3, "row_pitch smaller than a row");
    PSD_REQUIRE(n <= 1 || frame_stride >= (int64_t)e->sw * 3, "bad frame_stride");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    int rc = ensure_staging(e);
    if (rc) return rc;
    rc = ensure_capacity(e, e->n_frames + n + 1);
    if (rc) return rc;
    int64_t done = 0;
    while (done < n) {
        const int64_t b = (n - done < e->max_batch) ? (n - done) : e->max_batch;
        const int slot = e->next_slot;
        e->next_slot ^= 1;
        rc = stage_host(e, bgr + done * frame_stride, b, frame_stride, row_pitch, flags, slot);
        if (rc) return rc;
        rc = run_batch(e, e->dev_stage[slot], e->src_frame_bytes, b, e->n_frames + 1, false);
        if (rc) return rc;
        PSD_CUDA(cudaEventRecord(e->slot_free[slot], e->compute_stream));
        e->n_frames += b;
        done += b;
    }
    return PSD_OK;
}

int psd_engine_sync(psd_engine* e) {
    PSD_REQUIRE(e, "null engine");
    PSD_CUDA(cudaSetDevice(e->device));
    PSD_CUDA(cudaStreamSynchronize(e->copy_stream));
    PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
    return PSD_OK;
}

void* psd_engine_compute_stream(psd_engine* e) { return e ? (void*)e->compute_stream : nullptr; }
int64_t psd_engine_frame_count(const psd_engine* e) { return e ? e->n_frames : -1; }
int psd_engine_edge_kernel_size(const psd_engine* e) { return e ? e->ksize : -1; }

int psd_engine_read_sums(psd_engine* e, int64_t first, int64_t n, psd_frame_sums* out) {
    PSD_REQUIRE(e && out, "psd_engine_read_sums: null argument");
    PSD_REQUIRE(first >= -1 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    int rc = psd_engine_sync(e);
    if (rc) return rc;
    if (n) PSD_CUDA(cudaMemcpy(out, e->d_sums + first + 1, (size_t)n * sizeof(psd_frame_sums), cudaMemcpyDeviceToHost));
    return PSD_OK;
}

int psd_engine_read_yhist(psd_engine* e, int64_t first, int64_t n, uint32_t* out) {
    PSD_REQUIRE(e && out, "psd_engine_read_yhist: null argument");
    PSD_REQUIRE(e->features & PSD_F_YHIST, "engine was created without PSD_F_YHIST");
    PSD_REQUIRE(first >= -1 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    int rc = psd_engine_sync(e);
    if (rc) return rc;
    if (n) PSD_CUDA(cudaMemcpy(out, e->d_yhist + (first + 1) * 256, (size_t)n * 256 * 4, cudaMemcpyDeviceToHost));
    return PSD_OK;
}

int psd_engine_read_hash(psd_engine* e, int64_t first, int64_t n, uint64_t* out) {
    PSD_REQUIRE(e && out, "psd_engine_read_hash: null argument");
    PSD_REQUIRE(e->features & PSD_F_HASH, "engine was created without PSD_F_HASH");
    PSD_REQUIRE(first >= -1 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    int rc = psd_engine_sync(e);
    if (rc) return rc;
    if (n) PSD_CUDA(cudaMemcpy(out, e->d_hash + (first + 1) * PSD_HASH_WORDS, (size_t)n * PSD_HASH_WORDS * 8, cudaMemcpyDeviceToHost));
    return PSD_OK;
}

int psd_engine_device_hash(psd_engine* e, const uint64_t** hashes) {
    PSD_REQUIRE(e && hashes, "psd_engine_device_hash: null argument");
    *hashes = e->d_hash ? e->d_hash + PSD_HASH_WORDS : nullptr;
    return PSD_OK;
}

int psd_engine_device_results(psd_engine* e, const psd_frame_sums** sums, const uint32_t** yhist) {
    PSD_REQUIRE(e, "null engine");
    if (sums) *sums = e->d_sums + 1;
    if (yhist) *yhist = e->d_yhist ? e->d_yhist + 256 : nullptr;
    return PSD_OK;
}

int psd_engine_timing_reset(psd_engine* e) {
    PSD_REQUIRE(e, "null engine");
    int rc = psd_engine_sync(e);
    if (rc) return rc;
    e->ev_next = 0; e->ev_score.clear(); e->ev_total.clear();
    return PSD_OK;
}

int psd_engine_timing_ms(psd_engine* e, float* total_ms, float* score_ms, uint64_t* score_launches) {
    PSD_REQUIRE(e, "null engine");
    int rc = psd_engine_sync(e);
    if (rc) return rc;
    float tot = 0.f, sc = 0.f;
    for (auto& p : e->ev_total) { float ms = 0; PSD_CUDA(cudaEventElapsedTime(&ms, p.first, p.second)); tot += ms; }
    for (auto& p : e->ev_score) { float ms = 0; PSD_CUDA(cudaEventElapsedTime(&ms, p.first, p.second)); sc += ms; }
    if (total_ms) *total_ms = tot;
    if (score_ms) *score_ms = sc;
    if (score_launches) *score_launches = e->ev_score.size();
    return PSD_OK;
}

int psd_engine_debug_plane(psd_engine* e, int which, int64_t index, uint8_t* out, size_t cap) {
    PSD_REQUIRE(e && out, "psd_engine_debug_plane: null argument");
    PSD_REQUIRE(index >= 0 && index < e->last_n, "index outside the last batch");
    int rc = psd_engine_sync(e);
    if (rc) return rc;
    if (which == 0) {
        PSD_REQUIRE(cap >= (size_t)e->frame_bytes, "buffer too small");
        PSD_CUDA(cudaMemcpy(out, e->last_scored + index * e->last_scored_stride, (size_t)e->frame_bytes, cudaMemcpyDeviceToHost));
        return PSD_OK;
    }
    PSD_REQUIRE(e->features & PSD_F_EDGES, "engine was created without PSD_F_EDGES");
    PSD_REQUIRE(cap >= (size_t)e->P, "buffer too small");
    PSD_REQUIRE(which >= 1 && which <= 3, "unknown plane %d", which);
    if (which == 2 || which == 3) {  // bit-packed maps -> 0/255 bytes
        const size_t words = (size_t)e->H * ((e->W + 31) / 32);
        const uint32_t* bits = which == 3 ? e->eb.bits_dil + index * words
                                          : e->eb.bits_in + index * (size_t)edge_tile_words(e->W, e->H);
        rc = edge_unpack(bits, e->eb.tmp, e->W, e->H, which == 2, e->compute_stream);
        if (rc) return rc;
        PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
        PSD_CUDA(cudaMemcpy(out, e->eb.tmp, (size_t)e->P, cudaMemcpyDeviceToHost));
        return PSD_OK;
    }
    PSD_CUDA(cudaMemcpy(out, e->eb.vplane + index * e->P, (size_t)e->P, cudaMemcpyDeviceToHost));
    return PSD_OK;
}

// ---- host-convenience scans over the engine-owned arrays ----
static int scan_to_host(psd_engine* e, double* d_tmp, double* out, size_t count) {
    PSD_CUDA(cudaStreamSynchronize(e->compute_stream));
    PSD_CUDA(cudaMemcpy(out, d_tmp, count * sizeof(double), cudaMemcpyDeviceToHost));
    return PSD_OK;
}

int psd_engine_scan_content_host(psd_engine* e, int64_t first, int64_t n, const double weights[4],
                                 double weight_abs_sum, double* out_components, double* out_content_val) {
    PSD_REQUIRE(e && weights && out_content_val, "psd_engine_scan_content_host: null argument");
    PSD_REQUIRE(first >= 0 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    double* tmp = nullptr;
    PSD_CUDA(cudaMalloc(&tmp, (size_t)n * 5 * sizeof(double)));
    int rc = psd_scan_content(e->d_sums + 1 + first, n, e->P, weights, weight_abs_sum, tmp + n, tmp, e->compute_stream);
    if (!rc) rc = scan_to_host(e, tmp, out_content_val, (size_t)n);
    if (!rc && out_components) rc = scan_to_host(e, tmp + n, out_components, (size_t)n * 4);
    cudaFree(tmp);
    return rc;
}

int psd_engine_scan_adaptive_host(psd_engine* e, const double* scores_host, int64_t n, int32_t window_width,
                                  double min_content_val, double* out_ratio) {
    PSD_REQUIRE(e && scores_host && out_ratio && n >= 0, "psd_engine_scan_adaptive_host: bad argument");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    double* tmp = nullptr;
    PSD_CUDA(cudaMalloc(&tmp, (size_t)n * 2 * sizeof(double)));
    int rc = PSD_OK;
    if (cudaMemcpyAsync(tmp, scores_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->compute_stream) != cudaSuccess) {
        set_error("scores upload failed"); rc = PSD_ERR_CUDA;
    }
    if (!rc) rc = psd_scan_adaptive(tmp, n, window_width, min_content_val, tmp + n, e->compute_stream);
    if (!rc) rc = scan_to_host(e, tmp + n, out_ratio, (size_t)n);
    cudaFree(tmp);
    return rc;
}

int psd_engine_scan_average_host(psd_engine* e, int64_t first, int64_t n, double* out_avg) {
    PSD_REQUIRE(e && out_avg, "psd_engine_scan_average_host: null argument");
    PSD_REQUIRE(first >= 0 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    double* tmp = nullptr;
    PSD_CUDA(cudaMalloc(&tmp, (size_t)n * sizeof(double)));
    int rc = psd_scan_average(e->d_sums + 1 + first, n, e->P * 3, tmp, e->compute_stream);
    if (!rc) rc = scan_to_host(e, tmp, out_avg, (size_t)n);
    cudaFree(tmp);
    return rc;
}

int psd_engine_scan_hist_correl_host(psd_engine* e, int64_t first, int64_t n, int32_t bins, double* out) {
    PSD_REQUIRE(e && out, "psd_engine_scan_hist_correl_host: null argument");
    PSD_REQUIRE(e->features & PSD_F_YHIST, "engine was created without PSD_F_YHIST");
    PSD_REQUIRE(first >= 0 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    double* tmp = nullptr;
    PSD_CUDA(cudaMalloc(&tmp, (size_t)n * sizeof(double)));
    // slot `first` (= stream frame first-1, or the halo slot) precedes slot first+1
    const uint32_t* prev = (first > 0 || e->halo_scored) ? e->d_yhist + first * 256 : nullptr;
    int rc = psd_scan_hist_correl(e->d_yhist + (first + 1) * 256, n, bins, prev, tmp, e->compute_stream);
    if (!rc) rc = scan_to_host(e, tmp, out, (size_t)n);
    cudaFree(tmp);
    return rc;
}

int psd_scan_hash_dist(const uint64_t* hashes, int64_t n, int32_t hash_size, const uint64_t* prev_hash, double* out,
                       void* stream) {
    PSD_REQUIRE(hashes && out && n >= 0 && hash_size >= 1 && hash_size <= 16, "psd_scan_hash_dist: bad arguments");
    return launch_hash_dist(hashes, n, hash_size, prev_hash, out, (cudaStream_t)stream);
}

int psd_engine_scan_hash_dist_host(psd_engine* e, int64_t first, int64_t n, double* out) {
    PSD_REQUIRE(e && out, "psd_engine_scan_hash_dist_host: null argument");
    PSD_REQUIRE(e->features & PSD_F_HASH, "engine was created without PSD_F_HASH");
    PSD_REQUIRE(first >= 0 && n >= 0 && first + n <= e->n_frames, "frame range out of bounds");
    if (n == 0) return PSD_OK;
    PSD_CUDA(cudaSetDevice(e->device));
    double* tmp = nullptr;
    PSD_CUDA(cudaMalloc(&tmp, (size_t)n * sizeof(double)));
    // slot `first` (= stream frame first-1, or the halo slot) precedes slot first+1
    const uint64_t* prev = (first > 0 || e->halo_scored) ? e->d_hash + first * PSD_HASH_WORDS : nullptr;
    int rc = psd_scan_hash_dist(e->d_hash + (first + 1) * PSD_HASH_WORDS, n, e->hash.size, prev, tmp, e->compute_stream);
    if (!rc) rc = scan_to_host(e, tmp, out, (size_t)n);
    cudaFree(tmp);
    return rc;
}

int psd_synth_frames(int device, void* d_out, const int32_t* params_host, int64_t n, int32_t width,
                     int32_t height, int64_t frame_stride, void* stream) {
    PSD_REQUIRE(d_out && params_host && n > 0 && width > 0 && height > 0, "psd_synth_frames: bad args");
    PSD_REQUIRE(frame_stride >= (int64_t)width * height * 3, "frame_stride smaller than a frame");
    PSD_CUDA(cudaSetDevice(device));
    int32_t* d_params = nullptr;
    PSD_CUDA(cudaMalloc(&d_params, (size_t)n * 24 * 4));
    PSD_CUDA(cudaMemcpy(d_params, params_host, (size_t)n * 24 * 4, cudaMemcpyHostToDevice));
    int rc = PSD_OK;
    int64_t done = 0;
    while (!rc && done < n) {
        const int64_t b = (n - done < 4096) ? (n - done) : 4096;
        rc = launch_synth((uint8_t*)d_out + done * frame_stride, d_params + done * 24, b, width, height,
                          frame_stride, (cudaStream_t)stream);
        done += b;
    }
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) { set_error("synth sync failed: %s", cudaGetErrorString(cudaGetLastError())); rc = PSD_ERR_CUDA; }
    cudaFree(d_params);
    return rc;
}

}  // extern "C"
