"""Helpers shared by the golden-fixture tests."""

from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from pyscenedetect_b200.synth import ScenePlan, render_frames

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_PATH = os.path.join(HERE, "golden", "golden_v1.json")

_cache: dict = {}


def load_golden() -> dict:
    if "g" not in _cache:
        with open(GOLDEN_PATH) as f:
            _cache["g"] = json.load(f)
    return _cache["g"]


def case_names() -> list[str]:
    return [c["name"] for c in load_golden()["cases"]]


def get_case(name: str) -> dict:
    for c in load_golden()["cases"]:
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
