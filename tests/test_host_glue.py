"""Engine-side C++ glue (astre_b200/host): built here, exercised by tests/host/test_host_glue.cpp,
which compares GpuTransformSystem / GpuCameraSystem / GpuVisualSystem with the CPU oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "host", "_build", "test_host_glue")


def build():
    for d in ("oracle", os.path.join("astre_b200", "csrc"), os.path.join("astre_b200", "host"), os.path.join("tests", "host")):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, d)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr


def test_glue_builds_and_mirror_logic():
    build()
    r = subprocess.run([BIN, "--cpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_glue_systems_match_oracle():
    if not os.path.exists(BIN):
        build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr
