"""bench.py helpers that run without a GPU: the nvidia-smi clock sampler keeps only the samples that
arrived inside the timed region (it is started before the warm-up because nvidia-smi needs ~0.1 s to
come up) and falls back to the whole run when the region is shorter than one sample period."""

import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


class _Proc:
    def terminate(self): pass
    def wait(self, timeout=None): pass
    def kill(self): pass


LINE = "0, {sm}, 1965, {p}, 0x0, Not Active, Not Active, Not Active, {cap}"


def test_clock_sampler_windows_samples():
    b = load_bench()
    s = b.ClockSampler(0)
    s.proc = _Proc()
    t = time.perf_counter()
    s.lines = [(t - 1.0, LINE.format(sm=1200, p=150.0, cap="Not Active")),        # warm-up sample
               (t + 0.1, LINE.format(sm=1950, p=400.0, cap="Active")),
               (t + 0.2, LINE.format(sm=1965, p=410.0, cap="Not Active"))]
    s.window = (t, t + 0.5)
    r = s.stop()
    assert r["samples"] == 2 and r["scope"] == "timed region"
    assert r["sm_mhz"] == 1957.5 and r["sm_max_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"]
    s.window = (t + 5, t + 6)  # nothing inside: fall back to every sample and say so
    r = s.stop()
    assert r["samples"] == 3 and r["scope"] == "warm-up + timed region"


def test_clock_sampler_without_nvidia_smi():
    b = load_bench()
    s = b.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]
