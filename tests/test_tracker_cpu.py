"""hdt_dirty_tracker (csrc/hdt_multi.cuh, host code of the C ABI): an edit's dirty spans from the hash table's bucket fill
counts, against its numpy twin edits.delta_from_bucket_sizes and against a full comparison of the arrays."""
import numpy as np
import pytest

from hashdag_b200 import edits, tracer
from test_edits_cpu import _grow


def test_tracker_reproduces_the_numpy_twin_and_the_arrays():
    rng = np.random.default_rng(5)
    layout = edits.HashLayout(13)
    sizes = np.zeros(layout.n_buckets, np.uint32)
    table = np.zeros(layout.n_pages, np.uint32)
    pool = np.zeros(2400 * 512, np.uint32)
    top = _grow(rng, layout, sizes, table, pool, 1, 300)
    t = tracer.DirtyTracker(13)
    assert t.n_buckets == layout.n_buckets
    t.snapshot(sizes)
    dev_pool, dev_table, last = pool.copy(), table.copy(), sizes.copy()
    for step in range(4):
        top = _grow(rng, layout, sizes, table, pool, top, 60 + 25 * step)
        want = edits.delta_from_bucket_sizes(layout, last, sizes, pool, table, first_node_index=11 + step, pool_top=top)
        pod = t.delta_pod(sizes, pool, table, 11 + step, top)
        pr, pp, tr, tp = tracer.DirtyTracker.arrays(pod)
        assert (pod.first_node_index, pod.pool_top) == (11 + step, top)
        assert np.array_equal(pr, want.pool_ranges) and np.array_equal(pp, want.pool_payload)
        assert np.array_equal(tr, want.table_ranges) and np.array_equal(tp, want.table_payload)
        edits.apply_spans_host(dev_pool, pr, pp)
        edits.apply_spans_host(dev_table, tr, tp)
        assert np.array_equal(dev_pool, pool) and np.array_equal(dev_table, table)   # the delta alone reproduces the arrays
        last = sizes.copy()
    # nothing grew: an empty delta
    pod = t.delta_pod(sizes, pool, table, 3, top)
    assert pod.n_pool_ranges == 0 and pod.n_table_ranges == 0 and pod.n_pool_payload == 0
    # a bucket that shrank is not an append-only edit
    shrunk = sizes.copy()
    shrunk[np.flatnonzero(sizes)[0]] -= 1
    with pytest.raises(tracer.TracerError):
        t.delta_pod(shrunk, pool, table, 3, top)
    t.close()


def test_tracker_rejects_short_bucket_arrays():
    t = tracer.DirtyTracker(17)
    assert t.n_buckets == 9 * 1024 + 8 * 65536
    with pytest.raises(tracer.TracerError):
        t.snapshot(np.zeros(100, np.uint32))
    t.close()
