"""Drop-in boundary, both directions, against the REAL reference (child process with /root/reference on
PYTHONPATH so pyscenedetect_b200.compat binds to the reference's own classes; see tests/ref_sm_driver.py):
this package's detectors inside the reference's `SceneManager`, and this package's `SceneManager` against
the reference's over end_time / duration / frame_skip / crop.  CPU box: oracle-backed fake engine; GPU box
(only if `scenedetect` is importable there): the real engine."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(engine):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(p for p in ("/root/reference", ROOT, env.get("PYTHONPATH", "")) if p)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_sm_driver.py"), "--engine", engine],
                         capture_output=True, text=True, env=env, timeout=1500)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["failures"] == [] and res["checked"] >= 40, res
    assert out.returncode == 0


@pytest.mark.refsrc
def test_reference_scene_manager_drives_our_detectors_fake_engine():
    _run("fake")


@pytest.mark.gpu
def test_reference_scene_manager_drives_our_detectors_real_engine():
    try:
        import scenedetect  # noqa: F401
    except ImportError:
        if not os.path.isdir("/root/reference/scenedetect"):
            pytest.skip("the reference is not importable on this box")
    _run("gpu")
