"""The boundary used from plain C, no Python in the path: examples/ttmlblend_demo.c (built by
__graft_entry__.build() against include/fluc_ttmlblend.h only) and tools/cue_storm.c run to
completion on the GPU."""
import os
import subprocess

import pytest

import __graft_entry__ as graft

pytestmark = pytest.mark.gpu


def test_plain_c_demo_runs():
    exe = os.path.join(graft.ROOT, "build", "ttmlblend_demo")
    if not os.path.exists(exe):
        graft.build()
    r = subprocess.run([exe, "640", "360", "8", "20"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "device frames" in out and "host frames" in out and "one frame" in out and "stats:" in out
    frames = int(out.split("stats: ")[1].split(" frames")[0])
    assert frames == 8 * 25 + 8 * (3 + 20 // 4 + 1) + 2000
