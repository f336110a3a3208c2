"""Soak of hdt_rebuild_color_leaf against the oracle (oracle/color_leaf.py): seeded random op lists over the depth-13 recipe scene's
colour leaf -- copies of any length from any place (adjacent ranges that must merge again, ranges across old macro boundaries),
fills of every weight width with few distinct colours (runs that continue across ops), thousands of tiny ops in one macro block,
unique and shared (offset) old leaves.  Prints one JSON line.

    python scripts/color_leaf_soak.py [--cases 200] [--seed 1]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_ops(rng, n_old, host, budget):
    """op list with about `budget` colours in total"""
    rows, total = [], 0
    style = int(rng.integers(0, 4))
    palette = [int(x) for x in rng.integers(0, 1 << 32, 3)]
    if style == 1:
        budget = min(budget, 4500)          # (the oracle takes milliseconds per op)
    while total < budget:
        r = rng.random()
        if style == 0:      # mostly long copies
            count = int(rng.integers(1, 60000)) if r < 0.7 else int(rng.integers(1, 40))
        elif style == 1:    # tiny ops
            count = int(rng.integers(1, 6))
        elif style == 2:    # mixed
            count = int(rng.integers(1, 3000))
        else:               # whole macro blocks and near misses
            count = 16384 * int(rng.integers(1, 3)) + int(rng.integers(-2, 3))
        count = max(1, min(count, n_old - 1))
        if rng.random() < (0.75 if style != 1 else 0.5):
            if rows and rows[-1][2] == host.OP_COPY and rng.random() < 0.3:
                src = rows[-1][0] + rows[-1][1]                      # continues the previous copy: blocks must merge again
            elif rng.random() < 0.2:
                src = 16384 * int(rng.integers(0, n_old // 16384))   # starts on an old macro boundary
            else:
                src = int(rng.integers(0, n_old))
            src = min(src, n_old - count)
            rows.append((src, count, host.OP_COPY, 0, 0, 0))
        else:
            bpw = int(rng.integers(0, 5))
            rows.append((0, count, host.OP_FILL, bpw, palette[int(rng.integers(0, 3))], int(rng.integers(0, 1 << bpw))))
        total += count
    return np.array(rows, dtype=host.OP_DTYPE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import golden_util as gu
    from hashdag_b200 import color_leaf as host, tracer
    from oracle import color_leaf as cl
    scene = gu.recipe_scene("d13")
    old_host = (scene.weights, scene.blocks, scene.macro_blocks)
    n = int(scene.n_voxels)
    t = tracer.DAGTracer(True, 64, 64, 13)
    rep = {"cases": 0, "ops": 0, "colours": 0, "different": 0, "first_difference": None}
    for k in range(a.cases):
        rng = np.random.default_rng(a.seed * 100003 + k)
        offset = None if k % 3 else int(rng.integers(1, 200000))
        n_old = n - (offset or 0)
        ops = random_ops(rng, n_old, host, int(rng.integers(1000, 250000)))
        old = tracer.CompressedColorLeaf.from_scene(scene)
        if offset is not None:
            old.offset = offset
        leaf, _ = t.rebuild_color_leaf(ops, old)
        want = cl.rebuild(ops.astype(cl.OP_DTYPE), old_host + (offset,))
        got = [np.zeros(0, dt) if x is None else x.cpu().numpy().view(dt) for x, dt in ((leaf.weights, np.uint32), (leaf.blocks, np.uint64), (leaf.macro_blocks, np.uint64))]
        same = all(g.shape == w.shape and np.array_equal(g, w) for g, w in zip(got, want))
        rep["cases"] += 1
        rep["ops"] += int(ops.size)
        rep["colours"] += int(ops["count"].sum())
        if not same:
            rep["different"] += 1
            if rep["first_difference"] is None:
                rep["first_difference"] = {"case": k, "seed": a.seed, "ops": int(ops.size), "offset": offset}
        if k % 20 == 0:
            print(k, rep, file=sys.stderr, flush=True)
    t.close()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
