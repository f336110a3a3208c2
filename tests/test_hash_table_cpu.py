"""The oracle of the hash-table insert (oracle/hash_table.py) against the reference's own HashTable::find_or_add_*
(tests/golden/ref_find_or_add_d13.npz, written by tests/golden/make_find_or_add_golden.py on a GPU box)."""
import json
import os

import numpy as np
import pytest

import golden_util as gu
import hash_table_cases as hc
from oracle import hash_table as ht

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", f"ref_find_or_add_{hc.RECIPE}.npz")


def run_oracle(table, level, leaves, nodes):
    ptrs = np.empty(len(nodes), dtype=np.uint32)
    added = 0
    for i, w in enumerate(nodes):
        p, new = table.find_or_add_leaf(int(w[0]) | (int(w[1]) << 32)) if leaves else table.find_or_add_interior(level, w)
        ptrs[i] = p
        added += new
    return ptrs, added


def test_hashes_against_known_values():
    # Utils::murmurhash32xN / murmurhash64 (utils.h:77-110): the scalar restatement and the vectorised one of the case builder agree
    rng = np.random.default_rng(3)
    rows = rng.integers(0, 1 << 32, (64, 5), dtype=np.uint64).astype(np.uint32)
    assert [ht.murmur32xn(r) for r in rows] == list(hc._hash32xn(rows))
    v = rng.integers(0, 1 << 63, 64, dtype=np.uint64)
    assert [ht.murmur64(int(x)) & 0xFFFFFFFF for x in v] == list(hc._hash64(v))
    assert ht.murmur32xn([]) == 0 and ht.murmur64(0) == 0


def test_scene_nodes_sit_in_the_bucket_their_hash_names():
    # the scene builder's table obeys the reference's placement rule, so the batches below exercise real buckets
    scene = gu.recipe_scene(hc.RECIPE)
    by_level = hc.nodes_by_level(scene.hash_pool, scene.hash_page_table, scene.hash_first_node_index, scene.levels)
    for level in (3, 9, scene.levels - 3):
        for p in by_level[level][:50]:
            w = hc.read_node(scene.hash_pool, scene.hash_page_table, p, False)
            bucket = ht.murmur32xn(w) & (ht.buckets_per_level(level) - 1)
            base = ht.make_ptr(level, bucket, 0)
            assert base <= int(p) < base + ht.bucket_capacity(level)
            assert int(p) - base + len(w) <= scene.hash_bucket_sizes[ht.bucket_global_index(level, bucket)]


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="fixture not generated yet")
def test_oracle_equals_the_reference_insert():
    z = np.load(GOLDEN)
    meta = json.loads(str(z["meta"]))
    scene = gu.recipe_scene(meta["recipe"])
    pool, table, sizes, top = hc.fresh_table(scene)
    t = ht.HashTable(pool, table, sizes, top, scene.levels)
    cases = hc.cases(scene.hash_pool, scene.hash_page_table, scene.hash_first_node_index, scene.levels)
    assert len(cases) == len(meta["cases"])
    for k, ((name, level, leaves, nodes), want) in enumerate(zip(cases, meta["cases"])):
        assert (name, level, leaves, len(nodes)) == (want["name"], want["level"], want["leaves"], want["n"])
        ptrs, _ = run_oracle(t, level, leaves, nodes)
        assert np.array_equal(ptrs, z[f"ptrs_{k}"]), name
        assert t.pool_top == want["pool_top"], name
        assert hc.digest(pool[:t.pool_top * 512]) == want["pool"], name
        assert hc.digest(table) == want["page_table"] and hc.digest(sizes) == want["bucket_sizes"], name


def test_the_end_of_page_rule_and_overflow():
    # hash_table.h:327-333: the search of a page stops nodeSize words before its end, so the node appended last is not found again
    scene = gu.recipe_scene(hc.RECIPE)
    pool, table, sizes, top = hc.fresh_table(scene)
    t = ht.HashTable(pool, table, sizes, top, scene.levels)
    node = np.array([0x41 | (7 << 8), 1234, 5678], dtype=np.uint32)
    p0, new0 = t.find_or_add_interior(9, node)
    p1, new1 = t.find_or_add_interior(9, node)
    assert new0 and new1 and p1 != p0
    p2, new2 = t.find_or_add_interior(9, node)          # now two copies: the first is found
    assert not new2 and p2 == p0
    level, leaves, nodes = hc.overflow_case(scene.levels)
    with pytest.raises(AssertionError):
        run_oracle(t, level, leaves, nodes)
