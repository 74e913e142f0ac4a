"""tools/sanitize_small.py — every kernel family once on small inputs (ticks with the persistent solver and with one launch per
stage, rope edits, casts and queries, batched worlds, a 60-colour star, slabs with both exchange forms, island shards, pose sync)
— must run through without a CUDA error."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_small_tour_of_every_kernel_family():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_small.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "sanitize_small: done" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
