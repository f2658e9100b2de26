"""Minimal in-memory `VideoStream`-shaped source (the reference's interface at
scenedetect/video_stream.py:79-222) for decoded BGR24 frames.  Video decoding itself is out
of scope: the hot path starts at decoded frames."""

from __future__ import annotations

from fractions import Fraction

import numpy as np

from .compat import FrameTimecode, _to_fraction


class ArrayVideoStream:
    """Forward-only stream over an (N,H,W,3) uint8 array (optionally page-locked)."""

    BACKEND_NAME = "array"

    def __init__(self, frames: np.ndarray, fps=30.0, pinned: bool = False, repeat: int = 1):
        if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be (N,H,W,3) uint8")
        self._frames = frames
        self._fps: Fraction = _to_fraction(fps)
        self._pinned = pinned
        self._total = frames.shape[0] * int(repeat)
        self._n = 0

    path = property(lambda self: "array")
    name = property(lambda self: "array")
    is_seekable = property(lambda self: False)
    frame_rate = property(lambda self: self._fps)
    frame_size = property(lambda self: (self._frames.shape[2], self._frames.shape[1]))
    aspect_ratio = property(lambda self: 1.0)
    frame_number = property(lambda self: self._n)
    is_pinned = property(lambda self: self._pinned)

    @property
    def base_timecode(self):
        return FrameTimecode(0, self._fps)

    @property
    def duration(self):
        return FrameTimecode(self._total, self._fps)

    @property
    def position(self):
        return FrameTimecode(max(0, self._n - 1), self._fps)

    @property
    def position_ms(self) -> float:
        return 0.0 if self._n == 0 else 1000.0 * (self._n - 1) / float(self._fps)

    def read(self, decode: bool = True):
        if self._n >= self._total:
            return False
        frame = self._frames[self._n % self._frames.shape[0]]
        self._n += 1
        return frame if decode else True

    def read_batch(self, max_frames: int):
        """Zero-copy view of up to `max_frames` consecutive frames (None at EOF)."""
        if self._n >= self._total:
            return None
        base = self._frames.shape[0]
        i = self._n % base
        k = min(max_frames, self._total - self._n, base - i)
        self._n += k
        return self._frames[i:i + k]

    def reset(self):
        self._n = 0

    def seek(self, target):
        raise NotImplementedError("ArrayVideoStream is forward-only")
