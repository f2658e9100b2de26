"""Restatement of the reference detectors' per-frame logic with the reference's own
cv2/numpy calls.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Time is carried as plain integer frame numbers plus one `fractions.Fraction` frame rate
(the reference's frame-number-backed `FrameTimecode`, scenedetect/common.py:191-811,
reduces to this for constant-frame-rate input: `a - b` is `max(0, fa - fb)`
(common.py:755), `tc >= int` compares frame numbers and `tc >= float` compares
`frame_num >= round(seconds * rate)` (common.py:627-638, 480-486)).
"""

from __future__ import annotations

import math
from fractions import Fraction

import cv2
import numpy


def _as_rate(fps) -> Fraction:
    if isinstance(fps, Fraction):
        return fps
    # scenedetect/common.py:126-145 maps NTSC-like float rates onto x/1001 fractions.
    for num in (24000, 30000, 60000, 120000):
        if abs(float(fps) - num / 1001.0) < 1e-3:
            return Fraction(num, 1001)
    return Fraction(fps).limit_denominator(1000000)


def timecode_to_seconds(text: str, rate: Fraction) -> float:
    """common.py:488-533 (`_timecode_to_seconds`)."""
    text = text.strip()
    if text.isdigit():
        return int(text) / float(rate)
    if text.find(":") >= 0:
        values = text.split(":")
        if len(values) not in (2, 3):
            raise ValueError("Invalid timecode (too many separators).")
        if len(values) == 3:
            hrs, mins = int(values[0]), int(values[1])
            secs = float(values[2]) if "." in values[2] else int(values[2])
        else:
            hrs, mins = 0, int(values[0])
            secs = float(values[1]) if "." in values[1] else int(values[1])
        if not (hrs >= 0 and mins >= 0 and secs >= 0 and mins < 60 and secs < 60):
            raise ValueError("Invalid timecode range (values outside allowed range).")
        return secs + (hrs * 60 * 60) + (mins * 60)
    if text.endswith("s"):
        text = text[:-1]
    if not text.replace(".", "").isdigit():
        raise ValueError("All characters in timecode seconds string must be digits.")
    return float(text)


def min_len_to_frames(length, rate: Fraction) -> int:
    """`(tc_a - tc_b) >= length` threshold in frames for frame-number-backed timecodes
    (common.py:627-638, 535-541): int -> frames; float seconds -> round(seconds * rate);
    str -> round(_timecode_to_seconds(str) * rate)."""
    if isinstance(length, int):
        return length
    if isinstance(length, float):
        return round(length * rate)
    if isinstance(length, str):
        return round(timecode_to_seconds(length, rate) * rate)
    raise TypeError("unsupported min_scene_len")


class RefFlashFilter:
    """scenedetect/detector.py:106-224."""

    MERGE = 0
    SUPPRESS = 1

    def __init__(self, mode: int, length, rate: Fraction):
        self._mode = mode
        self._rate = rate
        self._filter_length = 0
        self._filter_secs = None
        # detector.py:130-137
        if isinstance(length, float):
            self._filter_secs = length
        elif isinstance(length, str) and not length.strip().isdigit():
            self._filter_secs = timecode_to_seconds(length, Fraction(100))
        else:
            self._filter_length = int(length)
        self._last_above = None
        self._merge_enabled = False
        self._merge_triggered = False
        self._merge_start = None

    @property
    def max_behind(self) -> int:  # detector.py:143-150
        if self._mode == RefFlashFilter.SUPPRESS:
            return 0
        if self._filter_secs is not None:
            return math.ceil(self._filter_secs * 240.0)
        return self._filter_length

    def _disabled(self) -> bool:  # detector.py:153-157
        if self._filter_secs is not None:
            return self._filter_secs <= 0.0
        return self._filter_length <= 0

    def _ge_secs(self, frames: int) -> bool:
        return frames >= round(self._filter_secs * self._rate)

    def filter(self, t: int, above: bool) -> list[int]:  # detector.py:159-224
        if self._disabled():
            return [t] if above else []
        if self._last_above is None:
            self._last_above = t
        if self._filter_secs is None:
            self._filter_secs = self._filter_length / float(self._rate)
        met = self._ge_secs(max(0, t - self._last_above))
        if self._mode == RefFlashFilter.SUPPRESS:
            if not (above and met):
                return []
            self._last_above = t
            return [t]
        if above:
            self._last_above = t
        if self._merge_triggered:
            if met and not above and self._ge_secs(max(0, self._last_above - self._merge_start)):
                self._merge_triggered = False
                return [self._last_above]
            return []
        if not above:
            return []
        if met:
            self._merge_enabled = True
            return [t]
        if self._merge_enabled:
            self._merge_triggered = True
            self._merge_start = t
        return []


def mean_pixel_distance(left: numpy.ndarray, right: numpy.ndarray):
    """content_detector.py:29-36."""
    num_pixels = float(left.shape[0] * left.shape[1])
    return numpy.sum(numpy.abs(left.astype(numpy.int32) - right.astype(numpy.int32))) / num_pixels


def estimated_kernel_size(frame_width: int, frame_height: int) -> int:
    """content_detector.py:39-46."""
    size = 4 + round(math.sqrt(frame_width * frame_height) / 192)
    if size % 2 == 0:
        size += 1
    return size


def detect_edges(lum: numpy.ndarray, kernel: numpy.ndarray) -> numpy.ndarray:
    """content_detector.py:213-239 (kernel passed in)."""
    sigma = 1.0 / 3.0
    median = numpy.median(lum)
    low = int(max(0, (1.0 - sigma) * median))
    high = int(min(255, (1.0 + sigma) * median))
    edges = cv2.Canny(lum, low, high)
    return cv2.dilate(edges, kernel)


class RefContentDetector:
    """content_detector.py:49-243.  `metrics[t]` mirrors StatsManager rows."""

    METRIC_KEYS = ["content_val", "delta_hue", "delta_sat", "delta_lum", "delta_edges"]

    def __init__(self, threshold=27.0, min_scene_len=15, weights=(1.0, 1.0, 1.0, 0.0),
                 luma_only=False, kernel_size=None, filter_mode=RefFlashFilter.MERGE,
                 fps=30.0, with_stats=False):
        self._threshold = threshold
        self._weights = tuple(weights)
        if luma_only:
            self._weights = (0.0, 0.0, 1.0, 0.0)
        self._kernel = None
        if kernel_size is not None:
            if kernel_size < 3 or kernel_size % 2 == 0:
                raise ValueError("kernel_size must be odd integer >= 3")
            self._kernel = numpy.ones((kernel_size, kernel_size), numpy.uint8)
        self._last = None
        self._frame_score = None
        self._rate = _as_rate(fps)
        self._flash_filter = RefFlashFilter(filter_mode, min_scene_len, self._rate)
        self.with_stats = with_stats
        self.metrics: dict[int, dict] = {}

    def _calculate_frame_score(self, t: int, frame_img: numpy.ndarray):
        hue, sat, lum = cv2.split(cv2.cvtColor(frame_img, cv2.COLOR_BGR2HSV))
        calculate_edges = (self._weights[3] > 0.0) or self.with_stats
        edges = None
        if calculate_edges:
            if self._kernel is None:
                k = estimated_kernel_size(lum.shape[1], lum.shape[0])
                self._kernel = numpy.ones((k, k), numpy.uint8)
            edges = detect_edges(lum, self._kernel)
        if self._last is None:
            self._last = (hue, sat, lum, edges)
            return 0.0
        comps = (
            mean_pixel_distance(hue, self._last[0]),
            mean_pixel_distance(sat, self._last[1]),
            mean_pixel_distance(lum, self._last[2]),
            0.0 if edges is None or self._last[3] is None
            else mean_pixel_distance(edges, self._last[3]),
        )
        score = sum(c * w for c, w in zip(comps, self._weights, strict=True)) / sum(
            abs(w) for w in self._weights)
        if self.with_stats:
            m = {"content_val": score}
            m.update(dict(zip(self.METRIC_KEYS[1:], comps)))
            self.metrics.setdefault(t, {}).update(m)
        self._last = (hue, sat, lum, edges)
        return score

    def process_frame(self, t: int, frame_img: numpy.ndarray) -> list[int]:
        self._frame_score = self._calculate_frame_score(t, frame_img)
        above = self._frame_score >= self._threshold
        return self._flash_filter.filter(t, above)

    def post_process(self, t: int) -> list[int]:
        return []


class RefAdaptiveDetector(RefContentDetector):
    """adaptive_detector.py:29-143."""

    def __init__(self, adaptive_threshold=3.0, min_scene_len=15, window_width=2,
                 min_content_val=15.0, weights=(1.0, 1.0, 1.0, 0.0), luma_only=False,
                 kernel_size=None, fps=30.0, with_stats=False):
        if window_width < 1:
            raise ValueError("window_width must be at least 1.")
        super().__init__(threshold=255.0, min_scene_len=0, weights=weights, luma_only=luma_only,
                         kernel_size=kernel_size, fps=fps, with_stats=with_stats)
        self.min_scene_len = min_len_to_frames(min_scene_len, self._rate)
        self.adaptive_threshold = adaptive_threshold
        self.min_content_val = min_content_val
        self.window_width = window_width
        self.ratio_key = "adaptive_ratio{} (w={})".format("_lum" if luma_only else "", window_width)
        self._buffer = []
        self._last_cut = None

    def process_frame(self, t: int, frame_img: numpy.ndarray) -> list[int]:
        super().process_frame(t, frame_img)
        if self._last_cut is None:
            self._last_cut = t
        required = 1 + 2 * self.window_width
        self._buffer.append((t, self._frame_score))
        if not len(self._buffer) >= required:
            return []
        self._buffer = self._buffer[-required:]
        target_t, target_score = self._buffer[self.window_width]
        avg = sum(s for i, (_t, s) in enumerate(self._buffer) if i != self.window_width) / (
            2.0 * self.window_width)
        zero = abs(avg) < 0.00001
        ratio = 0.0
        if not zero:
            ratio = min(target_score / avg, 255.0)
        elif zero and target_score >= self.min_content_val:
            ratio = 255.0
        if self.with_stats:
            self.metrics.setdefault(target_t, {})[self.ratio_key] = ratio
        met = ratio >= self.adaptive_threshold and target_score >= self.min_content_val
        if met and max(0, t - self._last_cut) >= self.min_scene_len:
            self._last_cut = target_t
            return [target_t]
        return []


class RefThresholdDetector:
    """threshold_detector.py:31-191."""

    FLOOR = 0
    CEILING = 1

    def __init__(self, threshold=12, min_scene_len=15, fade_bias=0.0, add_final_scene=False,
                 method=0, fps=30.0, with_stats=False):
        self.threshold = int(threshold)
        self.method = method
        self.fade_bias = fade_bias
        self._rate = _as_rate(fps)
        self.min_scene_len = min_len_to_frames(min_scene_len, self._rate)
        self.add_final_scene = add_final_scene
        self.processed_frame = False
        self.last_scene_cut = None
        self.last_fade_frame = None
        self.last_fade_type = None
        self.with_stats = with_stats
        self.metrics: dict[int, dict] = {}

    def process_frame(self, t: int, frame_img: numpy.ndarray) -> list[int]:
        if self.last_scene_cut is None:
            self.last_scene_cut = t
        cuts = []
        if self.with_stats and t in self.metrics and "average_rgb" in self.metrics[t]:
            frame_avg = self.metrics[t]["average_rgb"]
        else:
            frame_avg = numpy.mean(frame_img)
            if self.with_stats:
                self.metrics.setdefault(t, {})["average_rgb"] = frame_avg
        floor = self.method == RefThresholdDetector.FLOOR
        if self.processed_frame:
            if self.last_fade_type == "in" and (
                (floor and frame_avg < self.threshold) or (not floor and frame_avg >= self.threshold)
            ):
                self.last_fade_type = "out"
                self.last_fade_frame = t
            elif self.last_fade_type == "out" and (
                (floor and frame_avg >= self.threshold) or (not floor and frame_avg < self.threshold)
            ):
                if max(0, t - self.last_scene_cut) >= self.min_scene_len:
                    f_out = self.last_fade_frame
                    cuts.append(f_out + round((t - f_out) * (1.0 + self.fade_bias) / 2.0))
                    self.last_scene_cut = t
                self.last_fade_type = "in"
                self.last_fade_frame = t
        else:
            self.last_fade_frame = t
            self.last_fade_type = "out" if frame_avg < self.threshold else "in"
        self.processed_frame = True
        return cuts

    def post_process(self, t: int) -> list[int]:
        elapsed = t if self.last_scene_cut is None else max(0, t - self.last_scene_cut)
        if (self.last_fade_type == "out" and self.add_final_scene
                and self.last_fade_frame is not None and elapsed >= self.min_scene_len):
            return [self.last_fade_frame]
        return []


def calculate_histogram(frame_img: numpy.ndarray, bins: int = 256, normalize: bool = True):
    """histogram_detector.py:122-165."""
    y, _, _ = cv2.split(cv2.cvtColor(frame_img, cv2.COLOR_BGR2YUV))
    hist = cv2.calcHist([y], [0], None, [bins], [0, 256])
    if normalize:
        hist = cv2.normalize(hist, hist).flatten()
    return hist


class RefHistogramDetector:
    """histogram_detector.py:27-168."""

    def __init__(self, threshold=0.20, bins=128, min_scene_len=15, fps=30.0, with_stats=False):
        self._threshold = max(0.0, min(1.0, 1.0 - threshold))
        self._bins = bins
        self._rate = _as_rate(fps)
        self._min_scene_len = min_len_to_frames(min_scene_len, self._rate)
        self._last_hist = None
        self._last_cut = None
        self.metric_key = f"hist_diff [bins={bins}]"
        self.with_stats = with_stats
        self.metrics: dict[int, dict] = {}

    def process_frame(self, t: int, frame_img: numpy.ndarray) -> list[int]:
        cuts = []
        if frame_img.dtype != numpy.uint8:
            raise ValueError("Image must be 8-bit rgb for HistogramDetector")
        if frame_img.shape[2] != 3:
            raise ValueError("Image must have three color channels for HistogramDetector")
        # histogram_detector.py:87-88 tests `not self._last_cut`; a FrameTimecode is always truthy
        # (no __bool__/__len__), so this is an is-None check even for frame 0
        if self._last_cut is None:
            self._last_cut = t
        hist = calculate_histogram(frame_img, bins=self._bins)
        if self._last_hist is not None:
            hist_diff = cv2.compareHist(self._last_hist, hist, cv2.HISTCMP_CORREL)
            if hist_diff <= self._threshold and max(0, t - self._last_cut) >= self._min_scene_len:
                cuts.append(t)
                self._last_cut = t
            if self.with_stats:
                self.metrics.setdefault(t, {})[self.metric_key] = hist_diff
        self._last_hist = hist
        return cuts

    def post_process(self, t: int) -> list[int]:
        return []


def hash_frame(frame_img: numpy.ndarray, hash_size: int, factor: int) -> numpy.ndarray:
    """hash_detector.py:124-158 (pHash: gray -> INTER_AREA square -> /max -> DCT -> low band > median)."""
    gray_img = cv2.cvtColor(frame_img, cv2.COLOR_BGR2GRAY)
    imsize = hash_size * factor
    resized_img = cv2.resize(gray_img, (imsize, imsize), interpolation=cv2.INTER_AREA)
    max_value = numpy.max(numpy.max(resized_img))
    if max_value == 0:
        max_value = 1
    resized_img = numpy.asarray(numpy.float32(resized_img) / max_value)
    dct_complete = cv2.dct(resized_img)
    dct_low_freq = dct_complete[:hash_size, :hash_size]
    med = numpy.median(numpy.asarray(dct_low_freq, dtype=numpy.float32))
    return dct_low_freq > med


class RefHashDetector:
    """hash_detector.py:27-122."""

    def __init__(self, threshold=0.35, size=8, lowpass=2, min_scene_len=15, fps=30.0, with_stats=False):
        self._threshold = threshold
        self._rate = _as_rate(fps)
        self._min_scene_len = min_len_to_frames(min_scene_len, self._rate)
        self._size = size
        self._size_sq = float(size * size)
        self._factor = lowpass
        self._last_frame = None
        self._last_scene_cut = None
        self._last_hash = numpy.array([])
        self.metric_key = f"hash_dist [size={size} lowpass={lowpass}]"
        self.with_stats = with_stats
        self.metrics: dict[int, dict] = {}

    def process_frame(self, t: int, frame_img: numpy.ndarray) -> list[int]:
        cuts = []
        if self._last_scene_cut is None:
            self._last_scene_cut = t
        if self._last_frame is not None:
            curr_hash = hash_frame(frame_img, self._size, self._factor)
            last_hash = self._last_hash
            if last_hash.size == 0:
                last_hash = hash_frame(self._last_frame, self._size, self._factor)
            hash_dist = numpy.count_nonzero(curr_hash.flatten() != last_hash.flatten())
            hash_dist_norm = hash_dist / self._size_sq
            if self.with_stats:
                self.metrics.setdefault(t, {})[self.metric_key] = hash_dist_norm
            self._last_hash = curr_hash
            if hash_dist_norm >= self._threshold and max(0, t - self._last_scene_cut) >= self._min_scene_len:
                cuts.append(t)
                self._last_scene_cut = t
        self._last_frame = frame_img.copy()
        return cuts

    def post_process(self, t: int) -> list[int]:
        return []


# --- SceneManager pieces on the path (scene_manager.py) -------------------------------------


def compute_downscale_factor(frame_width: int, effective_width: int = 256) -> float:
    """scene_manager.py:123-140."""
    if frame_width < effective_width:
        return 1
    return frame_width / float(effective_width)


def downscaled_size(width: int, height: int, factor: float) -> tuple[int, int]:
    """scene_manager.py:673-676 (Python round = half-to-even)."""
    return max(1, round(width / factor)), max(1, round(height / factor))


def downscale_frame(frame: numpy.ndarray, factor: float, interpolation=cv2.INTER_LINEAR):
    """scene_manager.py:670-678."""
    if factor <= 1.0:
        return frame
    h, w = frame.shape[:2]
    dw, dh = downscaled_size(w, h, factor)
    return cv2.resize(frame, (dw, dh), interpolation=interpolation)


def run_detector(det, frames, downscale_factor: float = 1.0, first_frame: int = 0):
    """The SceneManager per-frame loop (scene_manager.py:410-440, 578-597, 621): returns the
    sorted unique cut list (scene_manager.py:403-408)."""
    cuts: list[int] = []
    t = first_frame
    for frame in frames:
        cuts += det.process_frame(t, downscale_frame(frame, downscale_factor))
        t += 1
    # scene_manager.py:621 passes video.position after the last read, i.e. the last frame.
    cuts += det.post_process(t - 1)
    return sorted(set(cuts))


def format_timecode(frame_num: int, rate: Fraction, precision: int = 3) -> str:
    """common.py:421-464 for frame-number-backed timecodes."""
    secs = frame_num / float(rate)
    hrs = int(secs / 3600.0)
    secs -= hrs * 3600.0
    mins = int(secs / 60.0)
    secs = max(0.0, secs - mins * 60.0)
    secs = round(secs, precision)
    secs = min(60.0, secs)
    if int(secs) == 60:
        secs = 0.0
        mins += 1
        if mins >= 60:
            mins = 0
            hrs += 1
    msec = format(secs, f".{precision + 1}f") if precision else ""
    msec_str = msec[-(2 + precision):-1]
    return f"{hrs:02d}:{mins:02d}:{int(secs):02d}{msec_str}"


def stats_csv(metrics: dict[int, dict], metric_keys, fps=30.0) -> str:
    """stats_manager.py:164-203: header + one row per frame that has any metric."""
    rate = _as_rate(fps)
    keys = sorted(set(metric_keys))
    lines = [",".join(["Frame Number", "Timecode", *keys])]
    for t in sorted(metrics.keys()):
        row = [str(t + 1), format_timecode(t, rate)]
        row += [str(metrics[t].get(k)) for k in keys]
        lines.append(",".join(row))
    return "\n".join(lines) + "\n"
