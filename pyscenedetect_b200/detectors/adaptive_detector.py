"""Drop-in for scenedetect.detectors.AdaptiveDetector (adaptive_detector.py:29-143)."""

from __future__ import annotations

import math

import numpy as np

from .content_detector import ContentDetector


class AdaptiveDetector(ContentDetector):
    """ContentDetector scores + a rolling window: ratio = target / mean(2w neighbours), computed
    by the trailing device scan (psd_scan_adaptive)."""

    ADAPTIVE_RATIO_KEY_TEMPLATE = "adaptive_ratio{luma_only} (w={window_width})"

    def __init__(
        self,
        adaptive_threshold: float = 3.0,
        min_scene_len=15,
        window_width: int = 2,
        min_content_val: float = 15.0,
        weights: ContentDetector.Components = ContentDetector.DEFAULT_COMPONENT_WEIGHTS,
        luma_only: bool = False,
        kernel_size: int | None = None,
    ):
        if window_width < 1:
            raise ValueError("window_width must be at least 1.")
        # adaptive_detector.py:71-77: the parent's own cut logic is disabled
        super().__init__(threshold=255.0, min_scene_len=0, weights=weights, luma_only=luma_only,
                         kernel_size=kernel_size)
        self.min_scene_len = min_scene_len
        self.adaptive_threshold = adaptive_threshold
        self.min_content_val = min_content_val
        self.window_width = window_width
        self._adaptive_ratio_key = AdaptiveDetector.ADAPTIVE_RATIO_KEY_TEMPLATE.format(
            window_width=window_width, luma_only="" if not luma_only else "_lum")
        self._timecodes: list = []
        self._last_cut = None
        self._next_target = window_width  # index (in this detector's frames) of the next target

    @property
    def event_buffer_length(self) -> int:
        return self.window_width

    def get_metrics(self) -> list[str]:
        return [*ContentDetector.METRIC_KEYS, self._adaptive_ratio_key]

    def _consume(self, timecodes: list, first: int) -> list:
        scores = self._score_batch(timecodes, first)
        if self._last_cut is None and timecodes:
            self._last_cut = timecodes[0]
        start = len(self._scores)
        self._scores.extend(scores)
        self._timecodes.extend(timecodes)
        if scores:
            self._frame_score = scores[-1]
        w = self.window_width
        total = len(self._scores)
        # targets whose 2w+1 window is now complete: [next_target, total - w)
        t0, t1 = self._next_target, total - w
        if t1 <= t0:
            return []
        lo = t0 - w
        window = np.asarray(self._scores[lo:total], dtype=np.float64)
        ratios = self._engine.scan_adaptive(window, w, self.min_content_val)
        cuts = []
        for t in range(t0, t1):
            ratio = ratios[t - lo]
            assert not math.isnan(ratio)
            target_tc, target_score = self._timecodes[t], self._scores[t]
            current_tc = self._timecodes[t + w]  # the frame whose arrival completed the window
            if self.stats_manager is not None:
                self.stats_manager.set_metrics(target_tc, {self._adaptive_ratio_key: ratio})
            threshold_met = ratio >= self.adaptive_threshold and target_score >= self.min_content_val
            min_length_met = (current_tc - self._last_cut) >= self.min_scene_len
            if threshold_met and min_length_met:
                self._last_cut = target_tc
                cuts.append(target_tc)
        self._next_target = t1
        del start
        return cuts
