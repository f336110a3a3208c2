"""Soak of hdt_find_or_add against the oracle (oracle/hash_table.py): seeded random batches applied one after the other to the depth-13
recipe scene's hash table -- known nodes, changed ones, repeats inside a batch, all levels, batches aimed at a few buckets so that they
fill pages -- pointers compared after every batch, the whole table every tenth.  Prints one JSON line.

    python scripts/hash_table_soak.py [--batches 100] [--seed 1]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import torch
    import golden_util as gu
    import hash_table_cases as hc
    from hashdag_b200 import tracer
    from oracle import hash_table as ht
    from test_gpu_hash_table import DeviceTable
    from test_hash_table_cpu import run_oracle
    scene = gu.recipe_scene(hc.RECIPE)
    levels, leaf_level = scene.levels, scene.levels - 2
    by_level = hc.nodes_by_level(scene.hash_pool, scene.hash_page_table, scene.hash_first_node_index, levels)
    t = tracer.DAGTracer(True, 64, 64, levels)
    spare = 60000
    pool = np.zeros((scene.hash_pool_top + spare) * 512, dtype=np.uint32)
    pool[:scene.hash_pool.size] = scene.hash_pool
    table, sizes = scene.hash_page_table.copy(), scene.hash_bucket_sizes.copy()
    o = ht.HashTable(pool, table, sizes, scene.hash_pool_top, levels)
    dt = DeviceTable(torch, tracer, scene, (pool.copy(), table.copy(), sizes.copy(), scene.hash_pool_top))
    rng = np.random.default_rng(a.seed)
    rep = {"batches": 0, "nodes": 0, "added": 0, "pages_opened": 0, "pointer_mismatches": 0, "table_mismatches": 0, "refused_by_both": 0}
    for k in range(a.batches):
        level = int(rng.integers(1, leaf_level + 1))
        leaves = level == leaf_level
        n = int(rng.integers(1, 1500))
        nodes = []
        if leaves:
            known = by_level[leaf_level]
            for _ in range(n):
                r = rng.random()
                if r < 0.3:
                    nodes.append(hc.read_node(pool, table, rng.choice(known), True))
                elif r < 0.45 and nodes:
                    nodes.append(nodes[int(rng.integers(0, len(nodes)))].copy())
                else:
                    nodes.append(rng.integers(1, 1 << 32, 2).astype(np.uint32))
        else:
            known, below = by_level[level], by_level[level + 1]
            few = level >= 9 and rng.random() < 0.2     # aim at two buckets (4096-word ones: they take a few batches)
            while len(nodes) < n:
                r = rng.random()
                if r < 0.3:
                    nodes.append(hc.read_node(pool, table, rng.choice(known), False))
                elif r < 0.45 and nodes:
                    nodes.append(nodes[int(rng.integers(0, len(nodes)))].copy())
                else:
                    size = int(rng.integers(2, 10))
                    nb = ht.buckets_per_level(level)
                    got = hc._random_interior(rng, 1 if not few else 8, size, 2 if few else nb, nb, below)
                    nodes.extend(got)
            nodes = nodes[:n]
        before = o.pool_top
        snapshot = None
        try:
            want, want_added = run_oracle(o, level, leaves, nodes)
        except AssertionError:
            want = None                         # a bucket (or the pool) overflowed: the reference aborts here
        if want is None:
            # the oracle stopped half way; restore it from the device table (which must refuse the batch and stay unchanged)
            try:
                t.find_or_add(dt.pod, level, nodes, leaves)
                rep["pointer_mismatches"] += 1
            except tracer.TracerError as e:
                rep["refused_by_both"] += int(e.code == tracer.ERR_CAPACITY)
            gp, gt, gs, gtop = dt.host()
            pool[:] = gp; table[:] = gt; sizes[:] = gs; o.pool_top = gtop
            continue
        got, added, pages = t.find_or_add(dt.pod, level, nodes, leaves)
        rep["batches"] += 1
        rep["nodes"] += len(nodes)
        rep["added"] += added
        rep["pages_opened"] += pages
        rep["pointer_mismatches"] += int((got != want).sum()) + int(added != want_added) + int(pages != o.pool_top - before)
        if k % 10 == 9 or k == a.batches - 1:
            gp, gt, gs, gtop = dt.host()
            rep["table_mismatches"] += int(not (gtop == o.pool_top and np.array_equal(gp, pool) and np.array_equal(gt, table) and np.array_equal(gs, sizes)))
            print(k, rep, file=sys.stderr, flush=True)
    t.close()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
