"""CPU: the dirty-span service (hashdag_b200/edits.py) -- span extraction, host mirror of the apply kernel,
and the world-2 broadcast (gloo) that replicas follow edits through."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import get_scene
from hashdag_b200 import edits


def test_spans_reproduce_the_new_array_and_stay_small():
    rng = np.random.default_rng(5)
    old = rng.integers(0, 2**32, 200_000, dtype=np.uint32)
    for trial in range(20):
        new = np.concatenate([old, rng.integers(0, 2**32, int(rng.integers(0, 3000)), dtype=np.uint32)]) if trial % 2 else old.copy()
        for _ in range(int(rng.integers(0, 12))):
            a = int(rng.integers(0, old.size - 600)); n = int(rng.integers(1, 600))
            new[a:a + n] = rng.integers(0, 2**32, n, dtype=np.uint32)
        n_new = new.size - (7 if trial % 4 == 3 else 0)
        ranges, payload = edits.dirty_spans(old, new, n_new)
        dst = np.zeros(max(old.size, n_new), np.uint32); dst[: old.size] = old
        edits.apply_spans_host(dst, ranges, payload)
        assert np.array_equal(dst[:n_new], new[:n_new])
        assert payload.size == int(ranges["n_words"].sum()) and payload.size <= (new[: old.size] != old).sum() + (new.size - old.size) + 32 * (len(ranges) + 12)
        assert np.all(ranges["dst_word"][1:] >= ranges["dst_word"][:-1] + ranges["n_words"][:-1])      # ordered, disjoint
    r, p = edits.dirty_spans(old, old)
    assert len(r) == 0 and p.size == 0


def _leaf(rng, n):
    return edits.ColorLeafArrays(rng.integers(0, 2**32, n, dtype=np.uint32), rng.integers(0, 2**63, n, dtype=np.uint64), rng.integers(0, 2**63, 2, dtype=np.uint64))


def test_delta_between_two_hash_dags():
    a, b = get_scene(13, 10), get_scene(13, 10, seed=99)
    rng = np.random.default_rng(1)
    la = [_leaf(rng, 5)]
    lb = [la[0], _leaf(rng, 9)]
    d = edits.diff_hash_dag(a.hash_pool, a.hash_page_table, b.hash_pool, b.hash_page_table, b.hash_first_node_index, b.hash_pool_top,
                            a.color_nodes, b.color_nodes, la, lb)
    pool = np.zeros(max(a.hash_pool.size, b.hash_pool.size), np.uint32); pool[: a.hash_pool.size] = a.hash_pool
    edits.apply_spans_host(pool, d.pool_ranges, d.pool_payload)
    table = a.hash_page_table.copy(); edits.apply_spans_host(table, d.table_ranges, d.table_payload)
    nodes = np.zeros(max(a.color_nodes.size, b.color_nodes.size), np.uint32); nodes[: a.color_nodes.size] = a.color_nodes
    edits.apply_spans_host(nodes, d.color_node_ranges, d.color_node_payload)
    assert np.array_equal(pool[: b.hash_pool.size], b.hash_pool) and np.array_equal(table, b.hash_page_table)
    assert np.array_equal(nodes[: b.color_nodes.size], b.color_nodes)
    assert sorted(d.color_leaves) == [1] and d.n_color_leaves == 2 and d.first_node_index == b.hash_first_node_index


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = get_scene(13, 10), get_scene(13, 10, seed=99)
        rng = np.random.default_rng(1)
        delta = None
        if rank == 0:   # only rank 0 runs the editor and knows the new version
            delta = edits.diff_hash_dag(a.hash_pool, a.hash_page_table, b.hash_pool, b.hash_page_table, b.hash_first_node_index, b.hash_pool_top,
                                        a.color_nodes, b.color_nodes, [], [_leaf(rng, 4), _leaf(rng, 11)])
        delta = edits.broadcast_delta(delta, src=0)
        pool = np.zeros(max(a.hash_pool.size, delta.pool_top * 512), np.uint32); pool[: a.hash_pool.size] = a.hash_pool
        edits.apply_spans_host(pool, delta.pool_ranges, delta.pool_payload)
        table = a.hash_page_table.copy(); edits.apply_spans_host(table, delta.table_ranges, delta.table_payload)
        ok = np.array_equal(pool[: b.hash_pool.size], b.hash_pool) and np.array_equal(table, b.hash_page_table)
        ok = ok and delta.first_node_index == b.hash_first_node_index and sorted(delta.color_leaves) == [0, 1] and delta.color_leaves[1].weights.size == 11
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok), delta.nbytes]))
    finally:
        dist.destroy_process_group()


def test_replicas_follow_an_edit_through_one_broadcast(tmp_path):
    get_scene(13, 10); get_scene(13, 10, seed=99)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "ok0.npy"), np.load(tmp_path / "ok1.npy")
    assert r0[0] == 1 and r1[0] == 1 and r0[1] == r1[1] > 0


def _grow(rng, layout, sizes, table, pool, top, n_buckets_to_grow):
    """Append random words to random buckets of a fake hash table, allocating physical pages on demand (hash_table.h:416-442)."""
    P = layout.PAGE
    cap = np.where(np.arange(layout.n_buckets) < (min(9, layout.levels) << 10), 1024, 4096)
    for b in rng.choice(layout.n_buckets, n_buckets_to_grow, replace=False):
        add = int(rng.integers(1, 700))
        add = min(add, int(cap[b] - sizes[b]))
        for k in range(add):
            v = int(layout.bucket_base[b] + sizes[b] + k)
            if table[v // P] == 0:
                table[v // P] = top
                top += 1
            pool[int(table[v // P]) * P + v % P] = rng.integers(1, 2 ** 32)
        sizes[b] += add
    return top


def test_bucket_size_tracker_reproduces_the_arrays_without_comparing_them():
    rng = np.random.default_rng(21)
    layout = edits.HashLayout(13)
    assert layout.n_buckets == 9 * 1024 + 4 * 65536 and layout.bucket_base[9 * 1024] == 9 * 1024 * 1024
    sizes = np.zeros(layout.n_buckets, np.uint32)
    table = np.zeros(layout.n_pages, np.uint32)
    pool = np.zeros(2000 * 512, np.uint32)
    top = _grow(rng, layout, sizes, table, pool, 1, 250)            # physical page 0 stays unused (hash_table.h:821)
    dev_pool, dev_table, last = pool.copy(), table.copy(), sizes.copy()
    for step in range(3):
        top = _grow(rng, layout, sizes, table, pool, top, 70)
        d = edits.delta_from_bucket_sizes(layout, last, sizes, pool, table, first_node_index=7, pool_top=top)
        full = edits.diff_hash_dag(dev_pool, dev_table, pool, table, 7, top)
        edits.apply_spans_host(dev_pool, d.pool_ranges, d.pool_payload)
        edits.apply_spans_host(dev_table, d.table_ranges, d.table_payload)
        assert np.array_equal(dev_pool, pool) and np.array_equal(dev_table, table)
        assert d.pool_payload.size < 1.5 * full.pool_payload.size + 64 * len(d.pool_ranges)    # a delta of the same order as the exact diff
        assert int(d.pool_ranges["n_words"].sum()) == d.pool_payload.size and int(d.table_ranges["n_words"].sum()) == d.table_payload.size
        last = sizes.copy()
    none = edits.delta_from_bucket_sizes(layout, sizes, sizes, pool, table, 7, top)
    assert len(none.pool_ranges) == 0 and len(none.table_ranges) == 0


def test_resolved_pool_host_mirror_against_a_walk_of_the_dag():
    """resolve_pool_host parses pages on their own (what resolve_pages_kernel does).  Ground truth: walk the HashDAG from
    its root through the page table; every child-pointer word met on the way must hold the child's physical index in the
    resolved pool, every other reachable word must be unchanged, and walking the resolved pool must meet the same nodes."""
    import golden_util as gu
    scene = gu.recipe_scene("d13")
    layout = edits.HashLayout(scene.levels)
    pool, table, top = scene.hash_pool, scene.hash_page_table, int(scene.hash_pool_top)
    res = edits.resolve_pool_host(layout, pool, table, top)
    assert res.shape == (top * 512,) and res.dtype == np.uint32
    phys = lambda v: int(table[v >> 9]) * 512 + (v & 511)
    leaf_level = scene.levels - 2
    frontier = {scene.hash_first_node_index}
    pointer_words = 0
    for level in range(leaf_level):
        nxt = set()
        for v in frontier:
            h = phys(v)
            hdr = int(pool[h])
            assert res[h] == hdr                                    # headers are copied
            n = bin(hdr & 0xFF).count("1")
            for k in range(1, n + 1):
                child = int(pool[h + k])
                assert int(res[h + k]) == phys(child)               # pointers are translated, once
                nxt.add(child)
                pointer_words += 1
        frontier = nxt
        if len(frontier) > 4000:                                   # bounded: a sample of every level is enough below the top
            frontier = set(sorted(frontier)[:: len(frontier) // 4000 + 1])
    for v in list(frontier)[:2000]:                                 # leaves: 64-bit masks, copied
        h = phys(v)
        assert res[h] == pool[h] and res[h + 1] == pool[h + 1]
    assert pointer_words > 20000
    # only pointer words differ, and resolving a subset of pages leaves the others as they are
    changed = np.flatnonzero(res != pool[: top * 512])
    assert 0 < changed.size
    some = sorted(set((changed[:: max(1, changed.size // 50)] // 512).tolist()))
    part = edits.resolve_pool_host(layout, pool, table, top, pages=some)
    mask = np.zeros(top * 512, bool)
    for p in some:
        mask[p * 512:(p + 1) * 512] = True
    assert np.array_equal(part[mask], res[mask]) and np.array_equal(part[~mask], pool[: top * 512][~mask])
