"""Pin oracle.ref_detectors against fixtures recorded from the real reference
(tests/golden/make_golden.py) - bit-exact metrics, identical cut lists, identical CSV."""

import hashlib

import pytest

from oracle import ref_detectors as R
from tests.golden_util import case_frames, case_names, get_case, golden_metrics


def build_ref(case, with_stats):
    kw = dict(case["kw"])
    det = case["det"]
    if "filter_mode" in kw:
        kw["filter_mode"] = {"MERGE": R.RefFlashFilter.MERGE, "SUPPRESS": R.RefFlashFilter.SUPPRESS}[kw["filter_mode"]]
    if "method" in kw:
        kw["method"] = {"FLOOR": 0, "CEILING": 1}[kw["method"]]
    cls = {"content": R.RefContentDetector, "adaptive": R.RefAdaptiveDetector,
           "threshold": R.RefThresholdDetector, "histogram": R.RefHistogramDetector,
           "hash": R.RefHashDetector}[det]
    return cls(fps=case["fps"], with_stats=with_stats, **kw)


def downscale_factor(case):
    if case["mode"] != "scene_manager":
        return 1.0
    w = case["gen"][1]
    h = case["gen"][2]
    if case.get("auto_downscale"):
        return R.compute_downscale_factor(max(w, h))
    return float(case.get("downscale", 1))


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference_golden(name):
    case = get_case(name)
    frames = case_frames(case)
    det = build_ref(case, with_stats=case["stats"])
    cuts = R.run_detector(det, frames, downscale_factor(case))
    assert cuts == case["cuts"]
    if case["stats"]:
        gold = golden_metrics(case)
        assert sorted(det.metrics.keys()) == sorted(gold.keys())
        for t, row in gold.items():
            for k, v in row.items():
                got = det.metrics[t].get(k)
                if v is None:
                    assert got is None, (t, k)
                else:
                    assert float(got) == v, (t, k, float(got), v)  # bit-exact
        csv = R.stats_csv(det.metrics, case["metric_keys"], case["fps"])
        assert csv.splitlines()[:4] == case["csv_head"]
        assert hashlib.sha256(csv.encode()).hexdigest() == case["csv_sha256"]
