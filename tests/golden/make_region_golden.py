"""Generate tests/golden/ref_regions_d13.npz with the reference's own DAGUtils::get_values / is_empty
(dag_utils.h:175-411, host functions; oracle/_ref harness).  Run on a GPU box (the harness initialises CUDA):

    gpurun -- 'python tests/golden/make_region_golden.py gpurun_out/golden'
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util as gu  # noqa: E402
import region_cases as rc  # noqa: E402
from oracle import ref  # noqa: E402


def main(out_dir, recipe="d13"):
    os.makedirs(out_dir, exist_ok=True)
    scene = gu.recipe_scene(recipe)
    rt = ref.RefTracer(scene.levels, gu.W, gu.H)
    rt.load_scene(scene)
    arrays, meta_cases = {}, []
    for k, (st, sz) in enumerate(rc.cases(scene)):
        v = rt.get_values(st, sz)
        arrays[f"values_{k}"] = np.packbits(v.reshape(-1))
        empties = [int(rt.is_empty(l, st, sz)) for l in rc.is_empty_levels(scene)]
        meta_cases.append({"start": list(st), "size": list(sz), "set": int(v.sum()), "is_empty": empties})
        print(k, st, sz, int(v.sum()), empties, flush=True)
    rt.close()
    meta = dict(recipe=recipe, cases=meta_cases, generator="tests/golden/make_region_golden.py",
                reference=f"oracle/_ref libhashdag_ref_{recipe}_256x256: DAGUtils::get_values<5> / DAGUtils::is_empty on the reference's HashDAG")
    np.savez_compressed(os.path.join(out_dir, f"ref_regions_{recipe}.npz"), meta=json.dumps(meta), **arrays)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
