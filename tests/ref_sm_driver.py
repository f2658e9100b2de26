"""Run in a child process with /root/reference on PYTHONPATH, so that pyscenedetect_b200.compat binds to the
REAL scenedetect classes (the drop-in situation):

  1. every golden case: this package's detector handed to the reference's own `SceneManager.add_detector` /
     `detect_scenes` (its decode thread, cv2.resize downscale, StatsManager) - cut list, scene list and CSV
     hash must equal what the reference's CPU detector produced (tests/golden/golden_v1.json);
  2. this package's batched SceneManager against the reference SceneManager driving the reference's CPU
     ContentDetector over end_time / duration (frames, seconds, timecode string) / frame_skip / crop settings.

`--engine fake` (CPU box) replaces the engine by the oracle-backed stand-in; `--engine gpu` uses the real one.
Prints one JSON line; exit code 0 iff everything matched."""

from __future__ import annotations

import argparse
import hashlib
import io
import json
import os
import sys
from fractions import Fraction

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from scenedetect.common import FrameTimecode  # noqa: E402
from scenedetect.detectors import ContentDetector as RefContentDetector  # noqa: E402
from scenedetect.scene_manager import SceneManager as RefSceneManager  # noqa: E402
from scenedetect.stats_manager import StatsManager as RefStatsManager  # noqa: E402
from scenedetect.video_stream import VideoStream  # noqa: E402

import pyscenedetect_b200.compat as compat  # noqa: E402

assert compat.USING_REFERENCE, "the reference must be importable before pyscenedetect_b200.compat"


class SyntheticStream(VideoStream):
    BACKEND_NAME = "synthetic"

    def __init__(self, frames, fps=30.0):
        self._frames, self._n = frames, 0
        self._fps = Fraction(fps).limit_denominator(1000000)

    path = property(lambda self: "synthetic")
    name = property(lambda self: "synthetic")
    is_seekable = property(lambda self: False)
    frame_rate = property(lambda self: self._fps)
    duration = property(lambda self: FrameTimecode(len(self._frames), self._fps))
    frame_size = property(lambda self: (self._frames.shape[2], self._frames.shape[1]))
    aspect_ratio = property(lambda self: 1.0)
    frame_number = property(lambda self: self._n)
    position = property(lambda self: FrameTimecode(max(0, self._n - 1), self._fps))
    position_ms = property(lambda self: 0.0 if self._n == 0 else 1000.0 * (self._n - 1) / float(self._fps))

    def read(self, decode=True):
        if self._n >= len(self._frames):
            return False
        self._n += 1
        return self._frames[self._n - 1] if decode else True

    def reset(self):
        self._n = 0

    def seek(self, target):
        raise NotImplementedError


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="fake", choices=["fake", "gpu"])
    ap.add_argument("--cases", default="")
    args = ap.parse_args()
    import pyscenedetect_b200.detectors._base as base_mod
    import pyscenedetect_b200.scene_manager as sm_mod
    if args.engine == "fake":
        from tests.fake_engine import OracleEngine
        base_mod.Engine = OracleEngine
        sm_mod.Engine = OracleEngine

        class FakePinned:
            def __init__(self, nbytes):
                self.array = np.zeros(nbytes, np.uint8)

            def close(self):
                pass
        sm_mod.PinnedBuffer = FakePinned
    from tests.golden_util import case_frames, case_names, get_case
    from tests.test_gpu_parity import _build
    failures, checked = [], 0
    wanted = [c for c in args.cases.split(",") if c] or case_names()
    # ---- 1. our detectors inside the reference's SceneManager ----
    for name in wanted:
        case = get_case(name)
        frames = case_frames(case)
        stats = RefStatsManager() if case["stats"] else None
        sm = RefSceneManager(stats)
        det = _build(case)
        sm.add_detector(det)
        if case["mode"] == "scene_manager" and case.get("auto_downscale"):
            sm.auto_downscale = True
        else:
            sm.auto_downscale = False
            sm.downscale = case.get("downscale", 1)
        n = sm.detect_scenes(SyntheticStream(frames, case["fps"]))
        cuts = [c.frame_num for c in sm.get_cut_list()]
        ok = n == frames.shape[0] and cuts == case["cuts"]
        if case["scene_list"] is not None:
            ok &= [[a.frame_num, b.frame_num] for a, b in sm.get_scene_list()] == case["scene_list"]
        if stats is not None and not any(k.startswith("hist_diff") for k in case["metric_keys"]):
            buf = io.StringIO()
            stats.save_to_csv(buf)
            ok &= hashlib.sha256(buf.getvalue().encode()).hexdigest() == case["csv_sha256"]
        det.close()
        checked += 1
        if not ok:
            failures.append(f"refsm:{name}")
    # ---- 2. our SceneManager vs the reference SceneManager + reference CPU detector ----
    case = get_case("content_default_nostats")
    frames = case_frames(case)
    from pyscenedetect_b200.detectors import ContentDetector
    from pyscenedetect_b200.scene_manager import SceneManager
    settings = [dict(), dict(end_time=100), dict(end_time=3.5), dict(end_time="00:00:05.100"), dict(duration=77),
                dict(duration=2.0), dict(duration="3s"), dict(end_time=0), dict(duration=0), dict(frame_skip=1),
                dict(frame_skip=3, end_time=120), dict(crop=(143, 10, 16, 81)), dict(crop=(0, 0, 40, 30), auto=True),
                dict(crop=(100, 50, 400, 300)), dict(start=40, duration=60), dict(start=40, end_time=90)]
    for st in settings:
        res = []
        for which in ("ref", "ours"):
            if which == "ref":
                sm, det, stream = RefSceneManager(), RefContentDetector(), SyntheticStream(frames, 30.0)
            else:
                sm, det = SceneManager(batch_size=16), ContentDetector()
                stream = SyntheticStream(frames, 30.0)  # no read_batch: the frame-by-frame path
            sm.add_detector(det)
            sm.auto_downscale = bool(st.get("auto", False))
            if "crop" in st:
                sm.crop = st["crop"]
            for _ in range(st.get("start", 0)):
                stream.read(decode=False)
            kw = {k: v for k, v in st.items() if k in ("end_time", "duration", "frame_skip")}
            n = sm.detect_scenes(stream, **kw)
            res.append((n, [c.frame_num for c in sm.get_cut_list()],
                        [[a.frame_num, b.frame_num] for a, b in sm.get_scene_list()]))
        checked += 1
        if res[0] != res[1]:
            failures.append(f"settings:{st}: ref={res[0]} ours={res[1]}")
        # and the zero-copy (read_batch) path of our SceneManager where it applies
        if "crop" not in st and not st.get("frame_skip") and not st.get("start"):
            from pyscenedetect_b200.video import ArrayVideoStream
            sm, det = SceneManager(batch_size=16), ContentDetector()
            sm.add_detector(det)
            sm.auto_downscale = False
            kw = {k: v for k, v in st.items() if k in ("end_time", "duration")}
            n = sm.detect_scenes(ArrayVideoStream(frames, 30.0), **kw)
            got = (n, [c.frame_num for c in sm.get_cut_list()], [[a.frame_num, b.frame_num] for a, b in sm.get_scene_list()])
            checked += 1
            if got != res[0]:
                failures.append(f"zero-copy settings:{st}: ref={res[0]} ours={got}")
    print(json.dumps({"checked": checked, "failures": failures, "engine": args.engine}))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
