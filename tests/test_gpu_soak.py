"""Short runs of the randomised soaks (scripts/parity_soak.py, color_leaf_soak.py, hash_table_soak.py): seeded random camera poses,
colour-leaf op lists and hash-table batches through the CUDA product against the oracles.  The long runs are under profiles/."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{") or l.startswith("PARITY_SOAK ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1].replace("PARITY_SOAK ", "", 1))


def test_random_poses_against_the_oracle():
    rep = run("parity_soak.py", "--poses", "20", "--seed", "3")
    assert len(rep["cases"]) == 9 and all(c["hit_pixels"] > 100000 for c in rep["cases"])
    assert rep["mismatched_pixels"] == 0 and rep["fog_max_channel_diff"] <= 1


def test_random_poses_against_the_reference_kernels():
    from oracle import ref
    if not ref.available(13, 256, 256):
        pytest.skip("oracle/_ref variant not built")
    rep = run("parity_soak.py", "--reference", "d13", "--poses", "25", "--seed", "5")
    assert len(rep["cases"]) == 2 and all(c["hit_pixels"] > 100000 for c in rep["cases"])
    assert rep["mismatched_pixels"] == 0 and rep["fog_max_channel_diff"] <= 1


def test_random_color_leaf_op_lists_against_the_oracle():
    rep = run("color_leaf_soak.py", "--cases", "12", "--seed", "4")
    assert rep["cases"] == 12 and rep["colours"] > 100000 and rep["different"] == 0, rep


def test_random_hash_table_batches_against_the_oracle():
    rep = run("hash_table_soak.py", "--batches", "20", "--seed", "6")
    assert rep["batches"] + rep["refused_by_both"] == 20 and rep["added"] > 1000
    assert rep["pointer_mismatches"] == 0 and rep["table_mismatches"] == 0, rep
