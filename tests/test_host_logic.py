"""Host-side logic (detector state machines, batching, StatsManager rows, SceneManager) on a
CPU-only box: the engine is replaced by an oracle-backed fake, everything above it is the
product code.  Expected values are the golden fixtures recorded from the real reference."""

import hashlib
import io

import pytest

import pyscenedetect_b200.detectors._base as base_mod
import pyscenedetect_b200.scene_manager as sm_mod
from tests.fake_engine import OracleEngine
from tests.golden_util import case_frames, case_names, get_case, golden_metrics
from tests.test_gpu_parity import _build, _scored_size


@pytest.fixture(autouse=True)
def fake_engine(monkeypatch):
    monkeypatch.setattr(base_mod, "Engine", OracleEngine)
    monkeypatch.setattr(sm_mod, "Engine", OracleEngine)


def _check(case, stats, n):
    from pyscenedetect_b200 import FrameTimecode
    gold = golden_metrics(case)
    keys = case["metric_keys"]
    for t in range(n):
        vals = stats.get_metrics(FrameTimecode(t, case["fps"]), keys)
        if t not in gold:
            assert all(v is None for v in vals)
            continue
        for k, v in zip(keys, vals):
            want = gold[t][k]
            if want is None:
                assert v is None
            elif k.startswith("hist_diff"):
                assert abs(float(v) - want) < 1e-9
            else:
                assert float(v) == want, (t, k)
    if not any(k.startswith("hist_diff") for k in keys):
        buf = io.StringIO()
        stats.save_to_csv(buf)
        assert hashlib.sha256(buf.getvalue().encode()).hexdigest() == case["csv_sha256"]


@pytest.mark.parametrize("name", case_names())
def test_strict_mode_host_logic(name):
    from pyscenedetect_b200 import FrameTimecode, StatsManager
    case = get_case(name)
    frames = case_frames(case)
    n = frames.shape[0]
    det = _build(case)
    stats = StatsManager() if case["stats"] else None
    det.stats_manager = stats
    if stats is not None:
        stats.register_metrics(det.get_metrics())
    det.configure(scored_size=_scored_size(case))
    cuts = []
    for i in range(n):
        cuts += det.process_frame(FrameTimecode(i, case["fps"]), frames[i])
    cuts += det.post_process(FrameTimecode(n - 1, case["fps"]))
    assert sorted({c.frame_num for c in cuts}) == case["cuts"]
    if stats is not None:
        _check(case, stats, n)


@pytest.mark.parametrize("batch", [5, 64])
@pytest.mark.parametrize("name", case_names())
def test_batched_scene_manager_host_logic(name, batch):
    from pyscenedetect_b200 import StatsManager
    from pyscenedetect_b200.scene_manager import SceneManager
    from pyscenedetect_b200.video import ArrayVideoStream
    case = get_case(name)
    frames = case_frames(case)
    stats = StatsManager() if case["stats"] else None
    sm = SceneManager(stats, batch_size=batch)
    sm.add_detector(_build(case))
    if case["mode"] == "scene_manager" and case.get("auto_downscale"):
        sm.auto_downscale = True
    else:
        sm.auto_downscale = False
        sm.downscale = case.get("downscale", 1)
    assert sm.detect_scenes(ArrayVideoStream(frames, case["fps"])) == frames.shape[0]
    assert [c.frame_num for c in sm.get_cut_list()] == case["cuts"]
    if case["scene_list"] is not None:
        assert [[a.frame_num, b.frame_num] for a, b in sm.get_scene_list()] == case["scene_list"]
    if stats is not None:
        _check(case, stats, frames.shape[0])


def test_frame_by_frame_stream_and_crop_and_callback():
    """Streams without read_batch go through the pinned double buffer - patched out here -
    so exercise crop + callback + duration with the zero-copy path only."""
    import numpy as np

    from pyscenedetect_b200.detectors import ContentDetector
    from pyscenedetect_b200.scene_manager import SceneManager
    from pyscenedetect_b200.video import ArrayVideoStream
    case = get_case("content_default_nostats")
    frames = case_frames(case)
    seen = []
    sm = SceneManager(batch_size=16)
    sm.auto_downscale = False
    sm.add_detector(ContentDetector())
    n = sm.detect_scenes(ArrayVideoStream(frames, 30.0), duration=200,
                         callback=lambda f, tc: seen.append(tc.frame_num))
    assert n == 200
    want = [c for c in case["cuts"] if c < 200]
    assert [c.frame_num for c in sm.get_cut_list()] == want
    assert sorted(seen) == want  # callbacks fire for cuts inside the retained batch
    # crop setter: inclusive corners in any order, stored one past the end (scene_manager.py:293-306)
    sm.crop = (10, 10, 5, 20)
    assert sm.crop == (5, 10, 10, 20) and sm._crop == (5, 10, 11, 21)
    with pytest.raises(ValueError):
        sm.crop = (-1, 0, 5, 5)
    with pytest.raises(TypeError):
        sm.crop = (1.0, 2, 3, 4)
    # downscale setter keeps auto_downscale on and rejects 0 (scene_manager.py:313-325)
    sm2 = SceneManager()
    sm2.downscale = 3
    assert sm2.auto_downscale is True and sm2.downscale == 3
    with pytest.raises(ValueError):
        sm2.downscale = 0
    assert isinstance(np.zeros(1), np.ndarray)


@pytest.mark.parametrize("batch", [4, 7, 64])
def test_callbacks_cross_batch_boundaries(batch):
    """AdaptiveDetector reports a cut `window_width` frames behind the frame it is looking at and FlashFilter's
    MERGE mode reports `_last_above`: both can sit in the previous batch (reference `_frame_buffer`,
    scene_manager.py:422-434)."""
    import numpy as np

    from pyscenedetect_b200.detectors import AdaptiveDetector
    from pyscenedetect_b200.scene_manager import SceneManager
    from pyscenedetect_b200.video import ArrayVideoStream
    case = get_case("adaptive_w2")
    frames = case_frames(case)
    seen = []
    sm = SceneManager(batch_size=batch)
    sm.auto_downscale = False
    sm.add_detector(_build(case))
    sm.detect_scenes(ArrayVideoStream(frames, case["fps"]),
                     callback=lambda f, tc: seen.append((tc.frame_num, int(np.asarray(f, np.int64).sum()))))
    cuts = [c.frame_num for c in sm.get_cut_list()]
    assert cuts == case["cuts"] and len(cuts) >= 3
    assert sorted(t for t, _ in seen) == cuts          # every cut got its callback ...
    for t, total in seen:                              # ... with the frame of the cut, not a recycled buffer
        assert total == int(frames[t].astype(np.int64).sum())
    assert isinstance(AdaptiveDetector(), object)


def test_kernel_size_agreement_only_among_edge_detectors():
    from pyscenedetect_b200.detectors import ContentDetector, ThresholdDetector
    from pyscenedetect_b200.scene_manager import SceneManager
    from pyscenedetect_b200.video import ArrayVideoStream
    frames = case_frames(get_case("content_default_nostats"))[:20]
    sm = SceneManager(batch_size=8)
    sm.auto_downscale = False
    sm.add_detector(ContentDetector(weights=ContentDetector.Components(1.0, 1.0, 1.0, 0.5), kernel_size=5))
    sm.add_detector(ThresholdDetector())   # owns no dilation kernel: must not clash with kernel_size=5
    assert sm.detect_scenes(ArrayVideoStream(frames, 30.0)) == 20


def test_plain_stream_crop_empty_host_logic(monkeypatch):
    """The frame-by-frame (pinned double-buffer) path, crop and the empty stream, with the page-locked
    buffer replaced by plain numpy memory (no GPU here)."""
    import numpy as np

    import tests.test_gpu_parity as gpu_tests

    class FakePinned:
        def __init__(self, nbytes):
            self.array = np.zeros(nbytes, np.uint8)

        def close(self):
            pass

    monkeypatch.setattr(sm_mod, "PinnedBuffer", FakePinned)
    gpu_tests.test_scene_manager_plain_stream_crop_and_empty(None)


def test_threshold_detector_uses_cached_metrics():
    """threshold_detector.py:122-125: a metric already in the StatsManager wins over the computed one."""
    import numpy as np

    from pyscenedetect_b200 import FrameTimecode, StatsManager
    from pyscenedetect_b200.detectors import ThresholdDetector
    frames = np.full((40, 36, 64, 3), 100, np.uint8)   # constant brightness: no fades on its own
    stats = StatsManager()
    det = ThresholdDetector(min_scene_len=5)
    det.stats_manager = stats
    stats.register_metrics(det.get_metrics())
    for t in range(10, 20):                           # cached "dark" stretch => fade out / fade in
        stats.set_metrics(FrameTimecode(t, 30.0), {"average_rgb": 3.0})
    cuts = []
    for t in range(40):
        cuts += det.process_frame(FrameTimecode(t, 30.0), frames[t])
    assert [c.frame_num for c in cuts] == [15]        # f_out=10, fade in at 20 -> 10 + round(10*1/2)
    assert stats.get_metrics(FrameTimecode(12, 30.0), ["average_rgb"]) == [3.0]
    assert float(stats.get_metrics(FrameTimecode(30, 30.0), ["average_rgb"])[0]) == 100.0


def test_adaptive_event_buffer_and_metric_keys():
    from pyscenedetect_b200.detectors import AdaptiveDetector, ContentDetector, HistogramDetector, ThresholdDetector
    a = AdaptiveDetector(window_width=4, luma_only=True)
    assert a.event_buffer_length == 4
    assert a.get_metrics() == ["content_val", "delta_hue", "delta_sat", "delta_lum", "delta_edges",
                               "adaptive_ratio_lum (w=4)"]
    assert ContentDetector(min_scene_len=15).event_buffer_length == 15
    assert ContentDetector(min_scene_len=0.5).event_buffer_length == 120     # ceil(0.5 * 240)
    assert HistogramDetector(bins=64).get_metrics() == ["hist_diff [bins=64]"]
    assert ThresholdDetector().get_metrics() == ["average_rgb"]
