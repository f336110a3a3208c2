"""Generate tests/golden/ref_find_or_add_d13.npz with the reference's own HashTable::find_or_add_interior_node /
find_or_add_leaf_node (hash_table.h:470-560; oracle/_ref harness, ref_find_or_add).  Run on a GPU box (the harness
initialises CUDA):

    gpurun -- 'python tests/golden/make_find_or_add_golden.py gpurun_out/golden'
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util as gu  # noqa: E402
import hash_table_cases as hc  # noqa: E402
from oracle import ref  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    scene = gu.recipe_scene(hc.RECIPE)
    rt = ref.RefTracer(scene.levels, gu.W, gu.H)
    rt.load_scene(scene, extra_pool_pages=hc.SPARE_PAGES)
    pool, table, sizes = rt.hash_views()
    first, top = rt.hash_info()
    # the reference's factory and the scene builder lay the table out identically: the batches can be derived from either
    assert top == scene.hash_pool_top and first == scene.hash_first_node_index
    assert np.array_equal(pool[:top * 512], scene.hash_pool) and np.array_equal(table, scene.hash_page_table)
    assert np.array_equal(sizes[:scene.hash_bucket_sizes.size], scene.hash_bucket_sizes) and not sizes[scene.hash_bucket_sizes.size:].any()
    arrays, meta = {}, []
    for k, (name, level, leaves, nodes) in enumerate(hc.cases(scene.hash_pool, scene.hash_page_table, first, scene.levels)):
        ptrs = rt.find_or_add(level, nodes, leaves)
        _, top = rt.hash_info()
        arrays[f"ptrs_{k}"] = ptrs
        meta.append(dict(name=name, level=level, leaves=bool(leaves), n=len(nodes), pool_top=top, pool=digest(pool[:top * 512]), page_table=digest(table),
                         bucket_sizes=digest(sizes[:scene.hash_bucket_sizes.size])))
        print(k, name, len(nodes), top, flush=True)
    rt.close()
    info = dict(recipe=hc.RECIPE, cases=meta, generator="tests/golden/make_find_or_add_golden.py",
                reference="oracle/_ref libhashdag_ref_d13_256x256: HashTable::find_or_add_interior_node / find_or_add_leaf_node, one node after the other")
    np.savez_compressed(os.path.join(out_dir, f"ref_find_or_add_{hc.RECIPE}.npz"), meta=json.dumps(info), **arrays)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
