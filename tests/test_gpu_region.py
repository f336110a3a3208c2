"""GPU: hdt_get_values / hdt_is_empty (csrc/hdt_region.cuh) against the oracle, the fixture written
by the reference's DAGUtils functions, and the live reference on an edited HashDAG."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu
import region_cases as rc
from conftest import ROOT
from hashdag_b200 import tracer
from oracle import ref, region
from test_region_cpu import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["hash", "basic"])
def test_fixture_cases(kind):
    scene = gu.recipe_scene("d13")
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    dag = tracer.HashDAG.from_scene(scene) if kind == "hash" else tracer.BasicDAG.from_scene(scene)
    _, golden = load_golden()
    for st, sz, v, empties in golden:
        got, ms = t.get_values(dag, st, sz)
        assert np.array_equal(got.cpu().numpy(), v), (st, sz)
        assert [int(t.is_empty(dag, l, st, sz)) for l in rc.is_empty_levels(scene)] == empties, (st, sz)
    t.close()


def test_depth17_against_the_oracle():
    scene = gu.recipe_scene("d17")
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    dag, odag = tracer.HashDAG.from_scene(scene), region.HostDag.from_scene(scene, "hash")
    n_set = 0
    for st, sz in rc.cases(scene, n_random=6, seed=9):
        got, _ = t.get_values(dag, st, sz)
        want = region.get_values(odag, st, sz)
        n_set += int(want.sum())
        assert np.array_equal(got.cpu().numpy(), want), (st, sz)
        for l in rc.is_empty_levels(scene):
            assert t.is_empty(dag, l, st, sz) == region.is_empty(odag, l, st, sz), (l, st, sz)
    assert n_set > 10000
    t.close()


def test_large_region_properties():
    """512 x 256 x 512 voxels around the surface: every set voxel is a voxel of the DAG (spot-checked with the
    point query), the low planes are empty, and the count equals the count of two half regions put together."""
    scene = gu.recipe_scene("d17")
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    dag, odag = tracer.HashDAG.from_scene(scene), region.HostDag.from_scene(scene, "hash")
    c = 1 << (scene.levels - 1)
    h0 = int(scene.heights[(c, c)])
    st, sz = (c - 256, h0 - 128, c - 256), (512, 256, 512)
    v, ms = t.get_values(dag, st, sz)
    v = v.cpu().numpy()
    assert v.sum() > 200000 and v[0].sum() == 0 and v[:, 0].sum() == 0 and v[:, :, 0].sum() == 0
    zs, ys, xs = np.nonzero(v)
    rng = np.random.default_rng(2)
    for i in rng.integers(0, zs.size, 300):
        assert region.get_value(odag, (st[0] + int(xs[i]), st[1] + int(ys[i]), st[2] + int(zs[i])))
    for _ in range(300):
        x, y, z = (int(rng.integers(1, s)) for s in sz)
        assert bool(v[z, y, x]) == region.get_value(odag, (st[0] + x, st[1] + y, st[2] + z))
    # the right half as its own region: identical except for its own low-x plane
    h, _ = t.get_values(dag, (st[0] + 256, st[1], st[2]), (256, 256, 512))
    h = h.cpu().numpy()
    assert np.array_equal(h[:, :, 1:], v[:, :, 257:]) and h[:, :, 0].sum() == 0
    assert not t.is_empty(dag, scene.levels - 2, st, sz) and t.is_empty(dag, scene.levels - 2, (c, h0 + 4000, c), (64, 64, 64))
    t.close()


def test_live_reference_on_an_edited_dag():
    if not ref.available(13, 256, 256):
        pytest.skip("oracle/_ref variant not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "region_scenario.py")], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("REGION_SCENARIO ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.loads(line[-1][len("REGION_SCENARIO "):])
    assert rep["cases"] >= 20 and rep["set_voxels"] > 10000 and rep["changed_by_edits"] > 0
    assert rep["get_values_mismatches"] == 0 and rep["is_empty_mismatches"] == 0 and rep["oracle_mismatches"] == 0


def test_argument_errors():
    scene = gu.recipe_scene("d13")
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    dag = tracer.HashDAG.from_scene(scene)
    with pytest.raises(tracer.TracerError):
        t.is_empty(dag, scene.levels - 1, (0, 0, 0), (4, 4, 4))          # beyond the leaf level (checkAlways, dag_utils.h:264)
    with pytest.raises(tracer.TracerError):
        t.get_values(dag, ((1 << scene.levels) - 2, 0, 0), (4, 4, 4))    # outside the volume
    v, _ = t.get_values(dag, (5, 5, 5), (0, 3, 3))
    assert v.numel() == 0
    t.close()
