"""GPU, config 4: tracing over edit-dirtied pages.  The reference's CPU edit code dirties the DAG, the deltas
reach the product's replica through hashdag_b200/edits.py + hdt_apply_ranges, frames are compared with the
reference kernels and the oracle after every edit (tests/edit_scenario.py, run in its own process)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("recipe", ["d13", "d17"])
def test_frames_follow_the_reference_through_a_sequence_of_edits(recipe):
    if not ref.available(int(recipe[1:]), 256, 256):
        pytest.skip("oracle/_ref variant not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "edit_scenario.py"), recipe], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("EDIT_SCENARIO ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    report = json.loads(line[-1][len("EDIT_SCENARIO "):])
    assert len(report) == 6
    for e in report:
        assert e["mismatched_pixels"] == 0, e
        assert e["delta_bytes"] < e["full_upload_bytes"] / 20, e      # a delta, not a re-upload
    assert sum(e["changed"] for e in report) >= 5 and report[-1]["unique_leaves"] > 0
    assert r.returncode == 0
