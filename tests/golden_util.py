"""Helpers shared by the golden-fixture tests."""

from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from pyscenedetect_b200.synth import ScenePlan, render_frames

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_PATH = os.path.join(HERE, "golden", "golden_v1.json")
GOLDEN_V2_PATH = os.path.join(HERE, "golden", "golden_v2.json")  # HashDetector + a histogram SceneManager case with cuts

_cache: dict = {}


def load_golden() -> dict:
    if "g" not in _cache:
        with open(GOLDEN_PATH) as f:
            _cache["g"] = json.load(f)
    return _cache["g"]


def load_golden_v2() -> dict:
    if "g2" not in _cache:
        with open(GOLDEN_V2_PATH) as f:
            _cache["g2"] = json.load(f)
    return _cache["g2"]


def case_names(which: str = "all") -> list[str]:
    v1 = [c["name"] for c in load_golden()["cases"]]
    v2 = [c["name"] for c in load_golden_v2()["cases"]]
    return {"v1": v1, "v2": v2, "all": v1 + v2}[which]


def get_case(name: str) -> dict:
    for c in load_golden()["cases"] + load_golden_v2()["cases"]:
        if c["name"] == name:
            return c
    raise KeyError(name)


def case_frames(case: dict) -> np.ndarray:
    key = tuple(case["gen"])
    if key not in _cache:
        n, w, h, seed, mn, mx, ns = case["gen"]
        plan = ScenePlan(n, seed=seed, noise_shift=ns, min_len=mn, max_len=mx)
        frames = render_frames(plan.params, w, h)
        assert hashlib.sha256(frames.tobytes()).hexdigest() == case["frames_sha256"], \
            "synthetic generator drifted from the committed golden fixtures"
        _cache[key] = frames
    return _cache[key]


def golden_metrics(case: dict) -> dict[int, dict[str, float | None]]:
    keys = case["metric_keys"]
    out = {}
    for t, vals in case["metrics"].items():
        out[int(t)] = {k: (None if v is None else float.fromhex(v)) for k, v in zip(keys, vals)}
    return out
