"""BASELINE.json configs[4] asks for a larger map than the reference's 3 x 30 x 100.  The dimensions are
compile-time constants of every library (as in the reference, gameplay.hpp:37), so the check runs in a process of
its own with SF_GEOMETRY=40x128: tests/large_map_check.py (Squad with the whole alphabet, Battle Royale with 16
players placed anywhere on the map; status + canonical-state hash every step, observations in both layouts on the
GPU).  Parity is pinned against the C oracle only: the reference cannot load such a map."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "large_map_check.py")
GEOMETRY = "40x128"


def run(mode, timeout):
    env = dict(os.environ, SF_GEOMETRY=GEOMETRY)
    env.pop("SF_LIB_PATH", None)
    out = subprocess.run([sys.executable, SCRIPT, mode], capture_output=True, text=True, timeout=timeout, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "LARGE MAP OK" in out.stdout
    return out.stdout


def test_host_build_of_the_tick_on_a_larger_map():
    out = run("host", 900)
    assert "host squad" in out and "host royale" in out


@pytest.mark.gpu
def test_cuda_library_on_a_larger_map():
    lib = os.path.join(ROOT, "strikeforce_b200", "libstrikeforce_b200_%s.so" % GEOMETRY)
    assert os.path.exists(lib), "%s is missing: __graft_entry__.build() builds it (SF_GEOMETRY=%s csrc/build.sh)" % (lib, GEOMETRY)
    out = run("gpu", 900)
    assert "gpu squad" in out and "gpu royale" in out
