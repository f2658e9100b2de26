"""Python handle on a `psd_engine` (include/psd_b200.h): the batched, device-resident
replacement for the per-frame cv2/numpy work inside the reference detectors."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import F_BGRSUM, F_EDGES, F_HASH, F_HSV, F_YHIST, HASH_WORDS, SUMS_DTYPE, check


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (for zero-staging submits)."""

    def __init__(self, nbytes: int):
        lib = _capi.load()
        p = C.c_void_p()
        check(lib.psd_host_alloc(int(nbytes), C.byref(p)), "psd_host_alloc")
        self._p = p
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.nbytes,))

    def close(self):
        if self._p is not None and self._p.value:
            _capi.load().psd_host_free(self._p)
            self._p = None
            self.array = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBuffer:
    """Plain device allocation owned through the C-ABI (torch-free HBM residency)."""

    def __init__(self, nbytes: int, device: int = 0):
        lib = _capi.load()
        p = C.c_void_p()
        check(lib.psd_device_alloc(device, int(nbytes), C.byref(p)), "psd_device_alloc")
        self.ptr = p.value
        self.nbytes = int(nbytes)
        self.device = device

    def upload(self, arr: np.ndarray, offset: int = 0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        check(_capi.load().psd_memcpy_h2d(self.device, self.ptr + offset, arr.ctypes.data, arr.nbytes))

    def download(self, nbytes: int, offset: int = 0) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        check(_capi.load().psd_memcpy_d2h(self.device, out.ctypes.data, self.ptr + offset, nbytes))
        return out

    def close(self):
        if self.ptr:
            _capi.load().psd_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One streaming scorer.  Frames go in (host ndarray batches or device pointers); per-frame
    integer sums / histograms stay in HBM; `scan_*` run the trailing device scans that turn
    them into the detectors' float64 metrics."""

    def __init__(self, src_width: int, src_height: int, features: int, width: int | None = None,
                 height: int | None = None, device: int = 0, max_batch: int = 64,
                 edge_kernel_size: int = 0, generic_kernel: bool = False, hash_size: int = 8,
                 hash_lowpass: int = 2):
        self._lib = _capi.load()
        cfg = _capi.PsdConfig()
        cfg.struct_size = C.sizeof(_capi.PsdConfig)
        cfg.device = device
        cfg.src_width, cfg.src_height = int(src_width), int(src_height)
        cfg.width = int(width if width is not None else src_width)
        cfg.height = int(height if height is not None else src_height)
        cfg.features = int(features)
        cfg.edge_kernel_size = int(edge_kernel_size)
        cfg.max_batch = int(max_batch)
        cfg.flags = _capi.CFG_GENERIC_KERNEL if generic_kernel else 0  # cross-check switch (tests)
        cfg.hash_size, cfg.hash_lowpass = int(hash_size), int(hash_lowpass)
        self.hash_size = int(hash_size)
        h = C.c_void_p()
        check(self._lib.psd_engine_create(C.byref(cfg), C.byref(h)), "psd_engine_create")
        self._h = h
        self.device = device
        self.src_width, self.src_height = cfg.src_width, cfg.src_height
        self.width, self.height = cfg.width, cfg.height
        self.features = int(features) | (F_HSV if features & F_EDGES else 0)
        self.max_batch = int(max_batch)
        self.n_pixels = self.width * self.height
        self.src_frame_bytes = self.src_width * self.src_height * 3

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.psd_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(self._lib.psd_engine_reset(self._h), "psd_engine_reset")

    # -- input --
    def _check_frames(self, frames: np.ndarray) -> np.ndarray:
        if frames.dtype != np.uint8:
            raise ValueError("frames must be uint8 BGR24")
        if frames.ndim == 3:
            frames = frames[None]
        if frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("frames must have shape (N, H, W, 3)")
        if frames.shape[1] != self.src_height or frames.shape[2] != self.src_width:
            raise ValueError(
                f"frame size {frames.shape[2]}x{frames.shape[1]} does not match engine "
                f"{self.src_width}x{self.src_height}")
        if frames.strides[3] != 1 or frames.strides[2] != 3:
            frames = np.ascontiguousarray(frames)
        return frames

    def set_halo(self, frame: np.ndarray):
        f = self._check_frames(frame)
        check(self._lib.psd_engine_set_halo_host(self._h, f.ctypes.data, f.strides[1]),
              "psd_engine_set_halo_host")

    def set_halo_device(self, dptr: int):
        check(self._lib.psd_engine_set_halo_device(self._h, dptr), "psd_engine_set_halo_device")

    def submit(self, frames: np.ndarray, pinned: bool = False):
        """Score a batch of host frames (N,H,W,3) uint8; strided views (crop) are honoured."""
        f = self._check_frames(frames)
        n = f.shape[0]
        fs = f.strides[0] if n > 1 else f.strides[1] * self.src_height
        check(self._lib.psd_engine_submit_host(self._h, f.ctypes.data, n, fs, f.strides[1],
                                               _capi.SUBMIT_PINNED if pinned else 0),
              "psd_engine_submit_host")

    def submit_device(self, dptr: int, n_frames: int, frame_stride: int | None = None):
        check(self._lib.psd_engine_submit_device(self._h, dptr, int(n_frames),
                                                 int(frame_stride or self.src_frame_bytes)),
              "psd_engine_submit_device")

    def sync(self):
        check(self._lib.psd_engine_sync(self._h), "psd_engine_sync")

    @property
    def compute_stream(self) -> int:
        """cudaStream_t handle of the engine's compute stream."""
        return int(self._lib.psd_engine_compute_stream(self._h) or 0)

    @property
    def frame_count(self) -> int:
        return int(self._lib.psd_engine_frame_count(self._h))

    @property
    def edge_kernel_size(self) -> int:
        return int(self._lib.psd_engine_edge_kernel_size(self._h))

    # -- raw integer results --
    def read_sums(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.frame_count - first if n is None else n
        out = np.zeros(n, dtype=SUMS_DTYPE)
        check(self._lib.psd_engine_read_sums(self._h, first, n, out.ctypes.data), "psd_engine_read_sums")
        return out

    def read_yhist(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.frame_count - first if n is None else n
        out = np.zeros((n, 256), dtype=np.uint32)
        check(self._lib.psd_engine_read_yhist(self._h, first, n, out.ctypes.data), "psd_engine_read_yhist")
        return out

    def read_hash(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """-> (n, 4) uint64: bit u*size+v of the 256-bit word = DCT[u][v] > median (hash_detector.py:156)."""
        n = self.frame_count - first if n is None else n
        out = np.zeros((n, HASH_WORDS), dtype=np.uint64)
        check(self._lib.psd_engine_read_hash(self._h, first, n, out.ctypes.data), "psd_engine_read_hash")
        return out

    def device_hash(self) -> int | None:
        p = C.c_void_p()
        check(self._lib.psd_engine_device_hash(self._h, C.byref(p)))
        return p.value

    def device_results(self) -> tuple[int, int | None]:
        s, h = C.c_void_p(), C.c_void_p()
        check(self._lib.psd_engine_device_results(self._h, C.byref(s), C.byref(h)))
        return s.value, h.value

    # -- trailing device scans (host-array convenience forms) --
    def scan_content(self, weights, first: int = 0, n: int | None = None):
        """-> (content_val[n], components[n,4]) as float64; bit-identical to
        content_detector.py:166-180."""
        n = self.frame_count - first if n is None else n
        w = (C.c_double * 4)(*[float(x) for x in weights])
        wsum = float(sum(abs(x) for x in weights))  # same expression as content_detector.py:180
        val = np.zeros(n, dtype=np.float64)
        comps = np.zeros((n, 4), dtype=np.float64)
        check(self._lib.psd_engine_scan_content_host(self._h, first, n, w, wsum, comps.ctypes.data,
                                                     val.ctypes.data), "psd_engine_scan_content_host")
        return val, comps

    def scan_adaptive(self, scores: np.ndarray, window_width: int, min_content_val: float) -> np.ndarray:
        s = np.ascontiguousarray(scores, dtype=np.float64)
        out = np.zeros(s.shape[0], dtype=np.float64)
        check(self._lib.psd_engine_scan_adaptive_host(self._h, s.ctypes.data, s.shape[0],
                                                      int(window_width), float(min_content_val),
                                                      out.ctypes.data), "psd_engine_scan_adaptive_host")
        return out

    def scan_average(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.frame_count - first if n is None else n
        out = np.zeros(n, dtype=np.float64)
        check(self._lib.psd_engine_scan_average_host(self._h, first, n, out.ctypes.data),
              "psd_engine_scan_average_host")
        return out

    def scan_hist_correl(self, bins: int, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.frame_count - first if n is None else n
        out = np.zeros(n, dtype=np.float64)
        check(self._lib.psd_engine_scan_hist_correl_host(self._h, first, n, int(bins), out.ctypes.data),
              "psd_engine_scan_hist_correl_host")
        return out

    def scan_hash_dist(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """hash_dist of frames [first, first+n) against their predecessors; NaN where there is none."""
        n = self.frame_count - first if n is None else n
        out = np.zeros(n, dtype=np.float64)
        check(self._lib.psd_engine_scan_hash_dist_host(self._h, first, n, out.ctypes.data),
              "psd_engine_scan_hash_dist_host")
        return out

    # -- instrumentation --
    def timing_reset(self):
        check(self._lib.psd_engine_timing_reset(self._h))

    def timing_ms(self) -> tuple[float, float, int]:
        t, s, k = C.c_float(), C.c_float(), C.c_uint64()
        check(self._lib.psd_engine_timing_ms(self._h, C.byref(t), C.byref(s), C.byref(k)))
        return t.value, s.value, k.value

    def debug_plane(self, which: int, index: int) -> np.ndarray:
        if which == 0:
            out = np.zeros((self.height, self.width, 3), dtype=np.uint8)
        else:
            out = np.zeros((self.height, self.width), dtype=np.uint8)
        check(self._lib.psd_engine_debug_plane(self._h, which, index, out.ctypes.data, out.nbytes),
              "psd_engine_debug_plane")
        return out


def bind_host_to_gpu_numa_node(device: int = 0) -> dict:
    """Pin the calling process to the CPUs of the GPU's NUMA node so that page-locked staging
    buffers allocated afterwards are local to the GPU's PCIe root (first-touch placement).  Returns
    what was found; a no-op when the topology is not exposed."""
    import os
    lib = _capi.load()
    buf = C.create_string_buffer(32)
    check(lib.psd_device_pci_bus_id(device, buf, 32), "psd_device_pci_bus_id")
    bdf = buf.value.decode().lower()
    info = {"pci": bdf, "numa_node": None, "cpus": None}
    base = f"/sys/bus/pci/devices/{bdf}"
    try:
        node = int(open(f"{base}/numa_node").read().strip())
        info["numa_node"] = node
        cpulist = open(f"{base}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if node >= 0 and allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except (OSError, ValueError):
        pass
    return info


def synth_frames_device(dptr: int, params: np.ndarray, width: int, height: int,
                        frame_stride: int | None = None, device: int = 0):
    """Render ScenePlan rows straight into HBM (bit-exact twin of synth.render_frames)."""
    p = np.ascontiguousarray(params, dtype=np.int32)
    check(_capi.load().psd_synth_frames(device, dptr, p.ctypes.data, p.shape[0], width, height,
                                        int(frame_stride or width * height * 3), None),
          "psd_synth_frames")


__all__ = ["Engine", "PinnedBuffer", "DeviceBuffer", "synth_frames_device", "bind_host_to_gpu_numa_node", "F_HSV", "F_BGRSUM",
           "F_YHIST", "F_EDGES", "F_HASH"]
