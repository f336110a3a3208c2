"""The reference's DAGUtils::get_values / is_empty (host functions, oracle/_ref) on its HashDAG before and after two
SphereEditor edits, beside the product (hdt_get_values / hdt_is_empty on the replica that followed the edits through
deltas) and the oracle on the edited host arrays.  Own process; called by tests/test_gpu_region.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import golden_util as gu                              # noqa: E402
import make_color_leaf_golden as mk                   # noqa: E402
import region_cases as rc                             # noqa: E402
from hashdag_b200 import edits, tracer                # noqa: E402
from oracle import ref, region                        # noqa: E402


def main(recipe="d13"):
    scene = gu.recipe_scene(recipe)
    rt = ref.RefTracer(scene.levels, gu.W, gu.H)
    rt.load_scene(scene)
    pool, table, first, top = rt.hash_dag()
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    rep = edits.HashDagReplica(t, pool, table, top, first, scene.levels, pool_capacity_pages=top + 4096)
    cases = rc.cases(scene)
    c = float(1 << (scene.levels - 1))
    h0 = float(scene.heights[(int(c), int(c))])
    cases += [((int(c) + 4, int(h0) - 10, int(c) - 4), (40, 36, 40)), ((int(c) - 36, int(h0) - 20, int(c) - 10), (30, 40, 30))]   # where the edits land
    before = [rt.get_values(st, sz) for st, sz in cases]
    gv_bad = ie_bad = or_bad = n_set = changed = n_cases = 0
    for phase in range(2):
        if phase == 1:
            for centre, radius, adding in mk.leaf_plan(scene)[:2]:
                rt.edit_sphere(centre, radius, adding)
            npool, ntable, nfirst, ntop = rt.hash_dag()
            rep.apply(edits.diff_hash_dag(pool, table, npool, ntable, nfirst, ntop))
            pool, table, first, top = npool, ntable, nfirst, ntop
        odag = region.HostDag(scene.levels, pool=pool, page_table=table, first_node_index=first)
        for k, (st, sz) in enumerate(cases):
            want = rt.get_values(st, sz)
            got, _ = t.get_values(rep.dag(), st, sz)
            gv_bad += int(not np.array_equal(got.cpu().numpy(), want))
            or_bad += int(not np.array_equal(region.get_values(odag, st, sz), want))
            n_set += int(want.sum())
            changed += int(phase == 1 and not np.array_equal(want, before[k]))
            for l in rc.is_empty_levels(scene):
                w = rt.is_empty(l, st, sz)
                ie_bad += int(t.is_empty(rep.dag(), l, st, sz) != w)
                or_bad += int(region.is_empty(odag, l, st, sz) != w)
            n_cases += 1
    t.close()
    rt.close()
    print("REGION_SCENARIO " + json.dumps({"cases": n_cases, "set_voxels": n_set, "changed_by_edits": changed, "get_values_mismatches": gv_bad,
                                           "is_empty_mismatches": ie_bad, "oracle_mismatches": or_bad}))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "d13")
