"""hdt_find_or_add beside the reference's own HashTable::find_or_add_* (oracle/_ref, ref_find_or_add) on the table the
reference's factory built: result pointers, pool, page table, bucket fill counts and pool top after every batch.  Own process
(the reference keeps its table in globals and the insert changes it); called by tests/test_gpu_hash_table.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util as gu                              # noqa: E402
import hash_table_cases as hc                         # noqa: E402
from hashdag_b200 import tracer                       # noqa: E402
from oracle import ref                                # noqa: E402
from test_gpu_hash_table import DeviceTable           # noqa: E402


BIG_SPARE = 160000   # pool pages for the cases and the timed batch (up to one page per touched bucket)


def main():
    import torch
    scene = gu.recipe_scene(hc.RECIPE)
    rt = ref.RefTracer(scene.levels, gu.W, gu.H)
    rt.load_scene(scene, extra_pool_pages=BIG_SPARE)
    rpool, rtable, rsizes = rt.hash_views()
    first, top = rt.hash_info()
    n = scene.hash_bucket_sizes.size
    t = tracer.DAGTracer(True, 64, 64, scene.levels)
    dt = DeviceTable(torch, tracer, scene, (rpool.copy(), rtable.copy(), rsizes[:n].copy(), top))
    rep = dict(cases=0, nodes=0, added=0, pages_opened=0, pointer_mismatches=0, table_mismatches=0)
    for name, level, leaves, nodes in hc.cases(rpool, rtable, first, scene.levels):
        want = rt.find_or_add(level, nodes, leaves)
        got, added, pages = t.find_or_add(dt.pod, level, nodes, leaves)
        gpool, gtable, gsizes, gtop = dt.host()
        rtop = rt.hash_info()[1]
        rep["cases"] += 1
        rep["nodes"] += len(nodes)
        rep["added"] += added
        rep["pages_opened"] += pages
        rep["pointer_mismatches"] += int((got != want).sum())
        same = gtop == rtop and np.array_equal(gpool[:rtop * 512], rpool[:rtop * 512]) and np.array_equal(gtable, rtable) and np.array_equal(gsizes, rsizes[:n])
        rep["table_mismatches"] += int(not same)
    # one large batch, timed on both sides: new nodes of a low level (65536 buckets), then the same nodes again (mostly found)
    import time
    rng = np.random.default_rng(99)
    n_big = 200000
    rows = np.empty((n_big, 5), dtype=np.uint32)
    rows[:, 0] = 0x1E | (rng.integers(1, 1 << 20, n_big).astype(np.uint32) << np.uint32(8))
    rows[:, 1:] = rng.integers(1 << 20, 1 << 28, (n_big, 4))
    offsets = np.arange(n_big + 1, dtype=np.int64) * 5
    words_dev = torch.from_numpy(rows.reshape(-1).view(np.int32).copy()).cuda()
    offsets_dev = torch.from_numpy(offsets).cuda()
    timing = {}
    for label in ("new", "again"):
        t0 = time.perf_counter()
        want = np.empty(n_big, dtype=np.uint32)
        assert rt.lib.ref_find_or_add(10, 0, rows.ctypes.data, offsets.astype(np.uint64).ctypes.data, n_big, want.ctypes.data) == 0
        t1 = time.perf_counter()
        got, added, pages = t.find_or_add_device(dt.pod, 10, words_dev, offsets_dev, n_big, False)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rep["pointer_mismatches"] += int((got.cpu().numpy().view(np.uint32) != want).sum())
        timing[label] = dict(nodes=n_big, added=added, pages=pages, reference_ms=(t1 - t0) * 1e3, ours_ms=(t2 - t1) * 1e3)
    gpool, gtable, gsizes, gtop = dt.host()
    rtop = rt.hash_info()[1]
    same = gtop == rtop and np.array_equal(gpool[:rtop * 512], rpool[:rtop * 512]) and np.array_equal(gtable, rtable) and np.array_equal(gsizes, rsizes[:n])
    rep["table_mismatches"] += int(not same)
    rep["timing"] = timing
    t.close()
    rt.close()
    print("HASH_TABLE_SCENARIO " + json.dumps(rep), flush=True)


if __name__ == "__main__":
    main()
