"""Every colour leaf the reference's ColorLeafBuilder writes during the six-edit scenario (oracle/_ref, GPU box),
decoded and re-encoded by the product (hdt_rebuild_color_leaf) and by the oracle: both must reproduce the
reference's arrays byte for byte.  Own process (the reference keeps its scene in globals); called by
tests/test_gpu_color_leaf.py."""
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
from hashdag_b200 import color_leaf as host, tracer   # noqa: E402
from oracle import color_leaf as cl                   # noqa: E402


def main(recipe):
    leaves = mk.reference_leaves(recipe)
    meta = json.loads(str(np.load(os.path.join(ROOT, "tests", "golden", f"ref_color_leaves_{recipe}.npz"))["meta"]))
    digests = [mk.leaf_digest(*l) for l in leaves]
    t = tracer.DAGTracer(True, gu.W, gu.H, int(recipe[1:]))
    gpu_bad = oracle_bad = 0
    n_colors = 0
    for w, b, m in leaves:
        n = cl.leaf_color_count(w, b, m)
        n_colors += n
        old = tracer.CompressedColorLeaf(tracer._to_device(w, "cuda:0"), tracer._to_device(b, "cuda:0"), tracer._to_device(m, "cuda:0"))
        bld = host.ColorLeafBuilder()
        bld.copy_colors(0, n)
        leaf, _ = bld.build(t, old)
        got = (np.zeros(0, np.uint32) if leaf.weights is None else leaf.weights.cpu().numpy().view(np.uint32),
               leaf.blocks.cpu().numpy().view(np.uint64), leaf.macro_blocks.cpu().numpy().view(np.uint64))
        gpu_bad += int(not all(np.array_equal(x, y) for x, y in zip(got, (w, b, m))))
        want = cl.encode(*cl.decode_range(w, b, m, 0, n))
        oracle_bad += int(not all(np.array_equal(x, y) for x, y in zip(want, (w, b, m))))
    t.close()
    print("COLOR_LEAF_SCENARIO " + json.dumps({"n_leaves": len(leaves), "n_colors": n_colors, "digests_match_fixture": digests == meta["digests"],
                                               "gpu_mismatches": gpu_bad, "oracle_mismatches": oracle_bad}))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "d13")
