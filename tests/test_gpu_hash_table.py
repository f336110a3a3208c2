"""hdt_find_or_add (GPU batch insert into the hash table) against the oracle (oracle/hash_table.py), the
reference's recorded results (tests/golden/ref_find_or_add_d13.npz) and, when oracle/_ref is built, the reference itself."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu
import hash_table_cases as hc
from test_hash_table_cpu import GOLDEN, ROOT, run_oracle

pytestmark = pytest.mark.gpu


class DeviceTable:
    def __init__(self, torch, tracer, scene, arrays=None):
        pool, table, sizes, top = arrays if arrays is not None else hc.fresh_table(scene)
        dev = "cuda:0"
        as_dev = lambda a: torch.from_numpy(a.view(np.int32).copy()).to(dev)
        self.pool, self.table, self.sizes = as_dev(pool), as_dev(table), as_dev(sizes)
        self.pod = tracer.HashTablePod(self.pool.data_ptr(), self.pool.numel(), self.table.data_ptr(), self.table.numel(),
                                       self.sizes.data_ptr(), self.sizes.numel(), top, scene.levels)

    def host(self):
        back = lambda t: t.cpu().numpy().view(np.uint32)
        return back(self.pool), back(self.table), back(self.sizes), int(self.pod.pool_top)


@pytest.fixture(scope="module")
def setup():
    import torch
    from hashdag_b200 import tracer
    from oracle import hash_table as ht
    scene = gu.recipe_scene(hc.RECIPE)
    t = tracer.DAGTracer(True, 64, 64, scene.levels)
    cases = hc.cases(scene.hash_pool, scene.hash_page_table, scene.hash_first_node_index, scene.levels)
    yield torch, tracer, ht, scene, t, cases
    t.close()


def test_batches_equal_the_sequential_insert(setup):
    torch, tracer, ht, scene, t, cases = setup
    dt = DeviceTable(torch, tracer, scene)
    pool, table, sizes, top = hc.fresh_table(scene)
    o = ht.HashTable(pool, table, sizes, top, scene.levels)
    golden = np.load(GOLDEN) if os.path.exists(GOLDEN) else None
    for k, (name, level, leaves, nodes) in enumerate(cases):
        before = o.pool_top
        want, want_added = run_oracle(o, level, leaves, nodes)
        got, added, pages = t.find_or_add(dt.pod, level, nodes, leaves)
        assert np.array_equal(got, want), name
        assert (added, pages) == (want_added, o.pool_top - before), name
        gpool, gtable, gsizes, gtop = dt.host()
        assert gtop == o.pool_top, name
        assert np.array_equal(gtable, table) and np.array_equal(gsizes, sizes), name
        assert np.array_equal(gpool, pool), name
        if golden is not None:
            assert np.array_equal(got, golden[f"ptrs_{k}"]), name
    if golden is not None:
        last = json.loads(str(golden["meta"]))["cases"][-1]
        gpool, gtable, gsizes, gtop = dt.host()
        assert (hc.digest(gpool[:gtop * 512]), hc.digest(gtable), hc.digest(gsizes)) == (last["pool"], last["page_table"], last["bucket_sizes"])


def test_one_at_a_time_equals_one_batch(setup):
    # the batch is defined as the sequential insert: feeding the nodes in pieces of any size gives the same table
    torch, tracer, ht, scene, t, cases = setup
    name, level, leaves, nodes = cases[2]
    a, b = DeviceTable(torch, tracer, scene), DeviceTable(torch, tracer, scene)
    whole, _, _ = t.find_or_add(a.pod, level, nodes, leaves)
    parts = []
    cuts = [0, 1, 2, 50, 51, 400, len(nodes)]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        parts.append(t.find_or_add(b.pod, level, nodes[lo:hi], leaves)[0])
    assert np.array_equal(np.concatenate(parts), whole)
    for x, y in zip(a.host(), b.host()):
        assert np.array_equal(x, y)


def test_a_voxel_edit_built_with_find_or_add_reads_back(setup):
    # one voxel toggled the way an edit does it (hash_dag_edits.h:544, :639): new leaf, then its ancestors bottom-up, each
    # through find_or_add; DAGUtils::get_values (hdt_get_values) through the new root differs in exactly that voxel
    torch, tracer, ht, scene, t, cases = setup
    dt = DeviceTable(torch, tracer, scene)
    pool, table, levels = scene.hash_pool, scene.hash_page_table, scene.levels
    node, pos, path = scene.hash_first_node_index, [0, 0, 0], []
    for level in range(levels - 2):
        w = hc.read_node(pool, table, node, False)
        child = int(w[0] & 0xFF).bit_length() - 1                       # last child: slot = popc(mask)
        slot = bin(int(w[0]) & 0xFF).count("1")
        sh = levels - (level + 1)
        for a in range(3):
            pos[a] |= ((child >> (2 - a)) & 1) << sh
        path.append((node, level, slot))
        node = int(w[slot])
    leaf = hc.read_node(pool, table, node, True)
    bits = int(leaf[0]) | (int(leaf[1]) << 32)
    bit = (bits & -bits).bit_length() - 1
    changed = bits ^ (1 << bit) if bits != (1 << bit) else bits | (1 << 63)
    bit = bit if bits != (1 << bit) else 63
    ptr, added, _ = t.find_or_add(dt.pod, levels - 2, [np.array([changed & 0xFFFFFFFF, changed >> 32], dtype=np.uint32)], True)
    below, inserted = int(ptr[0]), added
    for node, level, slot in reversed(path):
        w = hc.read_node(pool, table, node, False)
        w[slot] = below
        ptr, added, _ = t.find_or_add(dt.pod, level, [w], False)
        below, inserted = int(ptr[0]), inserted + added
    assert 2 <= inserted <= levels - 1 and below != scene.hash_first_node_index      # (the changed leaf may exist already)
    old = tracer.HashDAG(dt.pool, dt.table, int(dt.pod.pool_top), scene.hash_first_node_index, levels)
    new = tracer.HashDAG(dt.pool, dt.table, int(dt.pod.pool_top), below, levels)
    start = tuple(p - 4 for p in pos)      # get_values leaves the low planes of its region empty (dag_utils.h:190-207): start below the leaf
    v_old = t.get_values(old, start, (8, 8, 8))[0].cpu().numpy()
    v_new = t.get_values(new, start, (8, 8, 8))[0].cpu().numpy()
    x, y, z = ((bit >> 2) & 1) | ((bit >> 5) & 1) << 1, ((bit >> 1) & 1) | ((bit >> 4) & 1) << 1, (bit & 1) | ((bit >> 3) & 1) << 1
    diff = np.argwhere(v_old != v_new)
    assert diff.tolist() == [[z + 4, y + 4, x + 4]]
    far = (pos[0] ^ (1 << (levels - 1)), pos[1], pos[2])
    assert np.array_equal(t.get_values(old, far, (8, 8, 8))[0].cpu().numpy(), t.get_values(new, far, (8, 8, 8))[0].cpu().numpy())


def test_a_batch_that_does_not_fit_changes_nothing(setup):
    torch, tracer, ht, scene, t, cases = setup
    dt = DeviceTable(torch, tracer, scene)
    before = dt.host()
    level, leaves, nodes = hc.overflow_case(scene.levels)
    with pytest.raises(tracer.TracerError) as e:
        t.find_or_add(dt.pod, level, nodes, leaves)
    assert e.value.code == tracer.ERR_CAPACITY
    # a pool without a free page
    small = DeviceTable(torch, tracer, scene)
    small.pod.pool_capacity_words = scene.hash_pool_top * 512
    with pytest.raises(tracer.TracerError) as e:
        t.find_or_add(small.pod, cases[0][1], cases[0][3], cases[0][2])
    assert e.value.code == tracer.ERR_CAPACITY
    for d, tab in ((before, dt), (before, small)):
        for x, y in zip(d, tab.host()):
            assert np.array_equal(x, y)


def test_candidates_that_are_not_nodes_are_refused(setup):
    torch, tracer, ht, scene, t, cases = setup
    dt = DeviceTable(torch, tracer, scene)
    for level, leaves, nodes in ((5, False, [np.array([0xFF, 1, 2], dtype=np.uint32)]),            # mask says 8 children
                                 (scene.levels - 2, True, [np.array([1, 2, 3], dtype=np.uint32)]),  # a leaf is 64 bits
                                 (scene.levels - 2, False, [np.array([1, 2], dtype=np.uint32)]),   # interior insert at the leaf level
                                 (3, True, [np.array([1, 2], dtype=np.uint32)])):
        with pytest.raises(tracer.TracerError) as e:
            t.find_or_add(dt.pod, level, nodes, leaves)
        assert e.value.code == tracer.ERR_ARG
    assert t.find_or_add(dt.pod, 5, [], False)[1:] == (0, 0)


def test_against_the_reference_itself():
    # the reference's table lives in globals the insert changes: own process (tests/hash_table_scenario.py)
    from oracle import ref
    if not ref.available(13, gu.W, gu.H):
        pytest.skip("oracle/_ref variant not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hash_table_scenario.py")], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("HASH_TABLE_SCENARIO ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.loads(line[-1][len("HASH_TABLE_SCENARIO "):])
    assert rep["cases"] == 7 and rep["nodes"] > 10000 and rep["added"] > 5000 and rep["pages_opened"] > 100
    assert rep["pointer_mismatches"] == 0 and rep["table_mismatches"] == 0
