"""Batched counterpart of the reference `SceneManager` for the detectors in this package.

Mirrors the part of scenedetect/scene_manager.py that is on (or immediately around) the hot
path: `add_detector`, `detect_scenes`, `get_scene_list` / cut list, `auto_downscale` /
`downscale` / `crop`, StatsManager injection.  Differences by design:

* the reference hands one frame at a time to `detector.process_frame`
  (scene_manager.py:410-435); here frames are gathered into batches of `batch_size`, pushed
  through ONE fused GPU pass shared by all attached detectors, and each detector then runs
  its per-frame state machine over the device-computed metrics.  Cuts can therefore be
  emitted up to one batch late, which the reference's contract allows (cuts are sorted and
  de-duplicated at scene_manager.py:403-408; `post_process` is the final flush, :621);
* the `cv2.resize` downscale of the decode thread (scene_manager.py:670-678) runs on the
  device (exact fixed-point restatement) instead of on the host.

The resulting cut list, scene list and StatsManager CSV are identical to the reference's.
"""

from __future__ import annotations

import numpy as np

from .compat import FrameTimecode, StatsManager
from .detectors._base import EngineDetector
from .engine import Engine, PinnedBuffer

DEFAULT_MIN_WIDTH = 256


def compute_downscale_factor(frame_width: int, effective_width: int = DEFAULT_MIN_WIDTH) -> float:
    """scene_manager.py:123-140."""
    assert frame_width > 0 and effective_width > 0
    if frame_width < effective_width:
        return 1
    return frame_width / float(effective_width)


def get_scenes_from_cuts(cut_list, start_pos, end_pos):
    """scene_manager.py:171-210: contiguous (start, end) pairs from a sorted cut list."""
    scene_list = []
    if not cut_list:
        scene_list.append((start_pos, end_pos))
        return scene_list
    last_cut = start_pos
    for cut in cut_list:
        scene_list.append((last_cut, cut))
        last_cut = cut
    scene_list.append((last_cut, end_pos))
    return scene_list


class SceneManager:
    def __init__(self, stats_manager: StatsManager | None = None, device: int = 0,
                 batch_size: int = 64):
        self._detector_list: list[EngineDetector] = []
        self._cutting_list: list = []
        self._stats_manager = stats_manager
        self._device = device
        self._batch_size = int(batch_size)
        self._auto_downscale = True
        self._downscale = 1
        self._crop = None
        self._start_pos = None
        self._last_pos = None
        self._base_timecode = None
        self._engine: Engine | None = None
        self._frame_size = None

    # -- configuration (scene_manager.py:254-335) --
    @property
    def stats_manager(self):
        return self._stats_manager

    @property
    def auto_downscale(self) -> bool:
        return self._auto_downscale

    @auto_downscale.setter
    def auto_downscale(self, value: bool):
        self._auto_downscale = value

    @property
    def downscale(self) -> int:
        return self._downscale

    @downscale.setter
    def downscale(self, value: int):
        if value < 0:
            raise ValueError("Downscale factor must be a positive integer >= 1!")
        if self.auto_downscale:
            self._auto_downscale = False  # setting a factor disables auto (reference logs a warning)
        self._downscale = int(value) if value is not None else 1

    @property
    def crop(self):
        return self._crop

    @crop.setter
    def crop(self, value):
        if value is None:
            self._crop = None
            return
        if not (len(value) == 4 and all(isinstance(v, int) for v in value)):
            raise TypeError("crop region must be tuple of 4 ints.")
        if any(v < 0 for v in value):
            raise ValueError("crop coordinates must be >= 0")
        if value[2] <= value[0] or value[3] <= value[1]:
            raise ValueError("invalid crop region")
        self._crop = tuple(value)

    def add_detector(self, detector: EngineDetector) -> None:
        """scene_manager.py:337-352."""
        if not isinstance(detector, EngineDetector):
            raise TypeError("pyscenedetect_b200.SceneManager drives the GPU detectors of this "
                            "package; use the reference SceneManager for CPU detectors")
        detector.stats_manager = self._stats_manager
        if self._stats_manager is not None:
            self._stats_manager.register_metrics(detector.get_metrics())
        self._detector_list.append(detector)

    def clear(self) -> None:
        self._cutting_list.clear()
        self._last_pos = None
        self._start_pos = None

    # -- results (scene_manager.py:376-408) --
    def get_cut_list(self) -> list:
        if not self._cutting_list:
            return []
        return sorted(set(self._cutting_list))

    def get_scene_list(self, start_in_scene: bool = False) -> list:
        if self._base_timecode is None or self._last_pos is None:
            return []  # nothing processed yet / empty stream
        cut_list = self.get_cut_list()
        scene_list = get_scenes_from_cuts(cut_list, self._start_pos, self._last_pos + 1)
        if not cut_list and not start_in_scene:
            scene_list = []
        return sorted(scene_list)

    # -- the loop (scene_manager.py:446-623) --
    def _scored_size(self, w: int, h: int) -> tuple[int, int]:
        factor = compute_downscale_factor(max(w, h)) if self._auto_downscale else self._downscale
        if factor > 1.0:
            return max(1, round(w / factor)), max(1, round(h / factor))
        return w, h

    def detect_scenes(self, video, duration=None, end_time=None, frame_skip: int = 0,
                      show_progress: bool = False, callback=None) -> int:
        if not self._detector_list:
            raise ValueError("No detectors added")
        if frame_skip:
            raise NotImplementedError("frame_skip is not supported by the batched engine")
        self.clear()
        fw, fh = video.frame_size
        x0, y0, x1, y1 = (0, 0, fw, fh)
        if self._crop is not None:
            x0, y0, x1, y1 = self._crop
            if x0 >= fw or y0 >= fh:
                raise ValueError("crop starts outside boundaries of video frame")
            x1, y1 = min(x1, fw), min(y1, fh)
        w, h = x1 - x0, y1 - y0
        sw, sh = self._scored_size(w, h)
        features = 0
        ks = {d.edge_kernel_size_arg() for d in self._detector_list}
        for d in self._detector_list:
            features |= d.required_features()
        if len(ks) > 1:
            raise ValueError("detectors sharing a SceneManager must agree on kernel_size")
        self._engine = Engine(w, h, features, width=sw, height=sh, device=self._device,
                              max_batch=self._batch_size, edge_kernel_size=ks.pop())
        for d in self._detector_list:
            d.attach_engine(self._engine)
        self._base_timecode = video.position if hasattr(video, "position") else None
        fps = video.frame_rate
        total = 0
        if end_time is not None and duration is not None:
            raise ValueError("duration and end_time cannot be set at the same time!")
        limit = None
        if duration is not None:
            limit = int(duration) if not hasattr(duration, "frame_num") else duration.frame_num
        zero_copy = hasattr(video, "read_batch") and self._crop is None
        pinned = [None, None]
        pending = None  # (timecodes, frames_view, first engine index)
        which = 0
        done = False
        while True:
            # 1. gather the next batch while the GPU works on the previous one
            tcs, batch = [], None
            if not done:
                want = self._batch_size if limit is None else min(self._batch_size, limit - total)
                if want > 0:
                    if zero_copy:
                        pos0 = video.frame_number
                        view = video.read_batch(want)
                        if view is not None:
                            batch = view
                            tcs = [FrameTimecode(pos0 + i, fps) for i in range(view.shape[0])]
                    else:
                        if pinned[which] is None:
                            pinned[which] = PinnedBuffer(self._batch_size * w * h * 3)
                        buf = pinned[which].array.reshape(self._batch_size, h, w, 3)
                        k = 0
                        while k < want:
                            frame = video.read()
                            if frame is False:
                                break
                            np.copyto(buf[k], frame[y0:y1, x0:x1])
                            tcs.append(video.position)
                            k += 1
                        batch = buf[:k] if k else None
                if batch is None:
                    done = True
            # 2. launch the fused pass on the new batch (its H2D overlaps step 3's host work)
            nxt = None
            if batch is not None:
                use_pinned = (not zero_copy) or bool(getattr(video, "is_pinned", False))
                first = self._engine.frame_count
                self._engine.submit(batch, pinned=use_pinned)
                if self._start_pos is None:
                    self._start_pos = tcs[0]
                self._last_pos = tcs[-1]
                total += len(tcs)
                nxt = (tcs, batch, first)
                which ^= 1
            # 3. retire the previous batch (device scans + per-frame state machines)
            if pending is not None:
                self._consume(*pending, callback)
            pending = nxt
            if pending is None:
                break
        if self._last_pos is not None:
            for d in self._detector_list:
                self._cutting_list += d.post_process(self._last_pos)
        for p in pinned:
            if p is not None:
                p.close()
        return total

    def _consume(self, timecodes, frames, first, callback) -> None:
        for d in self._detector_list:
            cuts = d.process_batch(timecodes, frames, first=first)
            self._cutting_list += cuts
            if callback:
                for cut in cuts:
                    for tc, frame in zip(timecodes, frames):
                        if cut == tc:
                            callback(frame, tc)
