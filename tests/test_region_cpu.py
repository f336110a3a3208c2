"""CPU: the region-query oracle (oracle/region.py) against outputs of the reference's own DAGUtils::get_values /
is_empty (tests/golden/ref_regions_d13.npz), and against the point query DAGUtils::get_value."""
import json
import os

import numpy as np

import golden_util as gu
import region_cases as rc
from conftest import ROOT
from oracle import region

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_regions_d13.npz")


def load_golden():
    z = np.load(GOLDEN)
    meta = json.loads(str(z["meta"]))
    out = []
    for k, c in enumerate(meta["cases"]):
        n = int(np.prod(c["size"]))
        v = np.unpackbits(z[f"values_{k}"])[:n].reshape(c["size"][2], c["size"][1], c["size"][0])
        out.append((tuple(c["start"]), tuple(c["size"]), v, c["is_empty"]))
    return meta, out


def test_oracle_equals_the_reference_functions():
    scene = gu.recipe_scene("d13")
    meta, golden = load_golden()
    assert [(list(s), list(z)) for s, z in rc.cases(scene)] == [(c["start"], c["size"]) for c in meta["cases"]], "cases drifted from the fixture"
    assert sum(c["set"] for c in meta["cases"]) > 10000 and any(any(c["is_empty"]) for c in meta["cases"])
    for kind in ("hash", "basic"):
        dag = region.HostDag.from_scene(scene, kind)
        for st, sz, v, empties in golden:
            assert np.array_equal(region.get_values(dag, st, sz), v), (kind, st, sz)
            assert [int(region.is_empty(dag, l, st, sz)) for l in rc.is_empty_levels(scene)] == empties, (kind, st, sz)


def test_get_values_is_the_point_query_strictly_inside_the_region():
    scene = gu.recipe_scene("d13")
    dag = region.HostDag.from_scene(scene, "hash")
    st, sz = rc.cases(scene)[0]
    v = region.get_values(dag, st, sz)
    assert v.sum() > 100 and v[0].sum() == 0 and v[:, 0].sum() == 0 and v[:, :, 0].sum() == 0      # the planes through `start` stay 0
    rng = np.random.default_rng(1)
    for _ in range(600):
        x, y, z = (int(rng.integers(0, s)) for s in sz)
        want = region.get_value(dag, (st[0] + x, st[1] + y, st[2] + z)) and min(x, y, z) > 0
        assert bool(v[z, y, x]) == want


def test_every_traced_path_is_a_voxel_of_the_dag():
    """SURVEY.md §8c invariant (ii): DAGUtils::get_value(dag, path) holds for every non-null path trace_paths returns
    (oracle frame, BasicDAG and HashDAG), and fails one voxel above the terrain's highest hit."""
    from hashdag_b200 import camera
    from oracle import hdo
    scene = gu.recipe_scene("d13")
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    pose = gu.recipe_poses(scene)[0]
    prm = camera.trace_params(pose, info, scene.levels, 96, 96)
    for kind, dk in (("basic", hdo.DAG_BASIC), ("hash", hdo.DAG_HASH)):
        paths, _ = hdo.trace_paths(hdo.make_dag(scene, dk), 96, 96, prm)
        hit = paths[..., :3][paths[..., :3].any(-1)]
        assert hit.shape[0] > 2000
        dag = region.HostDag.from_scene(scene, kind)
        for p in hit[:: max(1, hit.shape[0] // 500)]:
            assert region.get_value(dag, tuple(int(v) for v in p))
        top = hit[hit[:, 1].argmax()]
        assert not region.get_value(dag, (int(top[0]), int(top[1]) + 300, int(top[2])))
