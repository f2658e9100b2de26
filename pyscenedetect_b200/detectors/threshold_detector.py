"""Drop-in for scenedetect.detectors.ThresholdDetector (threshold_detector.py:31-191)."""

from __future__ import annotations

import warnings
from enum import Enum

from .._capi import F_BGRSUM
from ..compat import FrameTimecode
from ._base import EngineDetector


class ThresholdDetector(EngineDetector):
    """Detects fades in/out of a brightness threshold; `average_rgb` comes from the fused GPU
    pass (exact integer byte sum -> one fp64 divide, as numpy.mean does)."""

    class Method(Enum):
        FLOOR = 0
        CEILING = 1

    THRESHOLD_VALUE_KEY = "average_rgb"
    FEATURES = F_BGRSUM

    def __init__(
        self,
        threshold: float = 12,
        min_scene_len=15,
        fade_bias: float = 0.0,
        add_final_scene: bool = False,
        method: "ThresholdDetector.Method" = Method.FLOOR,
        block_size=None,
    ):
        if block_size is not None:
            warnings.warn("The `block_size` argument is deprecated and will be removed in v0.8.",
                          DeprecationWarning, stacklevel=2)
        super().__init__()
        self.threshold = int(threshold)
        self.method = ThresholdDetector.Method(method)
        self.fade_bias = fade_bias
        self.min_scene_len = min_scene_len
        self.processed_frame = False
        self.last_scene_cut = None
        self.add_final_scene = add_final_scene
        self.last_fade = {"frame": None, "type": None}
        self._metric_keys = [ThresholdDetector.THRESHOLD_VALUE_KEY]

    def get_metrics(self) -> list[str]:
        return self._metric_keys

    def _consume(self, timecodes: list, first: int) -> list:
        avgs = self._engine.scan_average(first=first, n=len(timecodes))
        cuts = []
        for timecode, device_avg in zip(timecodes, avgs):
            if self.last_scene_cut is None:
                self.last_scene_cut = timecode
            # cached-metric short circuit, threshold_detector.py:122-125
            if self.stats_manager is not None and self.stats_manager.metrics_exist(
                    timecode, self._metric_keys):
                frame_avg = self.stats_manager.get_metrics(timecode, self._metric_keys)[0]
            else:
                frame_avg = device_avg
                if self.stats_manager is not None:
                    self.stats_manager.set_metrics(timecode, {self._metric_keys[0]: frame_avg})
            floor = self.method == ThresholdDetector.Method.FLOOR
            if self.processed_frame:
                if self.last_fade["type"] == "in" and (
                    (floor and frame_avg < self.threshold)
                    or (not floor and frame_avg >= self.threshold)
                ):
                    self.last_fade["type"] = "out"
                    self.last_fade["frame"] = timecode
                elif self.last_fade["type"] == "out" and (
                    (floor and frame_avg >= self.threshold)
                    or (not floor and frame_avg < self.threshold)
                ):
                    if (timecode - self.last_scene_cut) >= self.min_scene_len:
                        f_out = self.last_fade["frame"]
                        duration_frames = timecode.frame_num - f_out.frame_num
                        split = f_out.frame_num + round(duration_frames * (1.0 + self.fade_bias) / 2.0)
                        cuts.append(FrameTimecode(split, fps=timecode))
                        self.last_scene_cut = timecode
                    self.last_fade["type"] = "in"
                    self.last_fade["frame"] = timecode
            else:
                self.last_fade["frame"] = timecode
                self.last_fade["type"] = "out" if frame_avg < self.threshold else "in"
            self.processed_frame = True
        return cuts

    def post_process(self, timecode) -> list:
        cuts = []
        elapsed = timecode if self.last_scene_cut is None else timecode - self.last_scene_cut
        if (self.last_fade["type"] == "out" and self.add_final_scene
                and self.last_fade["frame"] is not None and elapsed >= self.min_scene_len):
            cuts.append(self.last_fade["frame"])
        return cuts
