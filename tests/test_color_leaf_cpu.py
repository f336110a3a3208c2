"""CPU: the colour-leaf rebuild oracle (oracle/color_leaf.py) against leaves the reference's own
ColorLeafBuilder wrote (tests/golden/ref_color_leaves_d13.npz), its two restatements against each other, and the
host mirror of the builder interface (hashdag_b200/color_leaf.py)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from hashdag_b200 import color_leaf as host
from oracle import color_leaf as cl

GOLDEN = {r: os.path.join(ROOT, "tests", "golden", f"ref_color_leaves_{r}.npz") for r in ("d13", "d17")}


def random_stream(rng, n, run=6.0, bpw_choices=(0, 1, 2, 3, 4)):
    """Colours with runs of equal (colorBits, bpw), like a voxel walk over a surface."""
    n_runs = max(1, int(n / run) + 1)
    lens = rng.geometric(1.0 / run, n_runs)
    cb = np.repeat(rng.integers(0, 2 ** 32, n_runs, dtype=np.uint64).astype(np.uint32), lens)[:n]
    bpw = np.repeat(rng.choice(bpw_choices, n_runs).astype(np.uint32), lens)[:n]
    while cb.size < n:
        cb = np.concatenate((cb, cb)); bpw = np.concatenate((bpw, bpw))
    cb, bpw = cb[:n], bpw[:n]
    w = (rng.integers(0, 16, n).astype(np.uint32)) & ((1 << bpw) - 1).astype(np.uint32)
    return cb, w, bpw


def golden_leaves(recipe):
    z = np.load(GOLDEN[recipe])
    meta = json.loads(str(z["meta"]))
    return meta, [(z[f"weights_{k}"], z[f"blocks_{k}"], z[f"macro_{k}"]) for k in range(meta["stored"])]


def test_vectorised_encoder_equals_the_statement_by_statement_port():
    rng = np.random.default_rng(5)
    for n, run in ((1, 1.0), (17, 2.0), (16384, 3.0), (16385, 1.2), (40000, 9.0)):
        cb, w, bpw = random_stream(rng, n, run)
        p = cl.ColorLeafBuilderPort()
        for i in range(n):
            p.add(int(cb[i]), int(w[i]), int(bpw[i]))
        for a, b in zip(p.build(), cl.encode(cb, w, bpw)):
            assert a.dtype == b.dtype and np.array_equal(a, b)


def test_decode_loop_form_equals_vectorised_and_inverts_the_encoder():
    rng = np.random.default_rng(6)
    cb, w, bpw = random_stream(rng, 50000, 4.0)
    W, B, M = cl.encode(cb, w, bpw)
    dcb, dw, dbpw = cl.decode_range(W, B, M, 0, cb.size)
    assert np.array_equal(dcb, cb) and np.array_equal(dw, w) and np.array_equal(dbpw, bpw)
    for i in rng.integers(0, cb.size, 400).tolist() + [0, 16383, 16384, cb.size - 1]:
        assert cl.get_color_loop(W, B, M, i) == (cb[i], w[i], bpw[i])
    # a shared leaf: the offset is added to the colour index (vwsc.h:407-410)
    scb, sw, sb = cl.decode_range(W, B, M, 100, 3000, offset=20000)
    assert np.array_equal(scb, cb[20100:23100]) and np.array_equal(sw, w[20100:23100])
    assert cl.get_color_loop(W, B, M, 100, offset=20000) == (cb[20100], w[20100], bpw[20100])


def test_add_large_single_color_is_n_times_add():
    a, b = cl.ColorLeafBuilderPort(), cl.ColorLeafBuilderPort()
    for builder in (a, b):
        for i in range(16000):
            builder.add(0x1234 + i // 7, i % 4, 2)
    a.add_large_single_color(0xABCDE, 40000)
    for _ in range(40000):
        b.add(0xABCDE, 0, 0)
    a.add(7, 1, 1); b.add(7, 1, 1)
    assert a.color_index == b.color_index
    for x, y in zip(a.build(), b.build()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("recipe", ["d13", "d17"])
def test_oracle_reproduces_the_reference_builders_leaves(recipe):
    """The pin: every stored reference leaf is what the oracle encoder makes of its own colours."""
    meta, leaves = golden_leaves(recipe)
    assert meta["stored"] >= 7 and len(meta["digests"]) == meta["n_leaves_in_scenario"]
    if recipe == "d17":
        assert max(m.size for _, _, m in leaves) >= 4, "depth-17 leaves should span macro blocks"
    seen_bpw = set()
    for w, b, m in leaves:
        n = cl.leaf_color_count(w, b, m)
        cb, wt, bpw = cl.decode_range(w, b, m, 0, n)
        seen_bpw |= set(np.unique(bpw).tolist())
        W, B, M = cl.encode(cb, wt, bpw)
        assert np.array_equal(B, b) and np.array_equal(M, m) and np.array_equal(W, w)
    assert len(seen_bpw) >= 3, seen_bpw


def test_scalar_port_reproduces_a_reference_leaf():
    _, leaves = golden_leaves("d17")
    w, b, m = min(leaves, key=lambda l: abs(l[1].size - 3000))
    n = cl.leaf_color_count(w, b, m)
    p = cl.ColorLeafBuilderPort()
    for i in range(n):
        p.add(*[int(v) for v in cl.get_color_loop(w, b, m, i)])
    W, B, M = p.build()
    assert np.array_equal(B, b) and np.array_equal(M, m) and np.array_equal(W, w)


def test_rebuild_from_ops():
    rng = np.random.default_rng(8)
    cb, w, bpw = random_stream(rng, 60000, 5.0)
    old = cl.encode(cb, w, bpw)
    ops = np.array([(0, 20000, cl.OP_COPY, 0, 0, 0), (0, 1, cl.OP_FILL, 3, 0xDEADBEEF, 5), (0, 0, cl.OP_FILL, 0, 1, 0),
                    (0, 32768, cl.OP_FILL, 0, 0x3FF003FF, 0), (25000, 35000, cl.OP_COPY, 0, 0, 0)], dtype=cl.OP_DTYPE)
    ecb, ew, eb = cl.expand_ops(ops, old + (None,))
    assert ecb.size == 20000 + 1 + 32768 + 35000
    assert np.array_equal(ecb[:20000], cb[:20000]) and ecb[20000] == 0xDEADBEEF and ew[20000] == 5 and eb[20000] == 3
    assert np.array_equal(ew[-35000:], w[25000:]) and (ecb[20001:20001 + 32768] == 0x3FF003FF).all()
    W, B, M = cl.rebuild(ops, old + (None,))
    dcb, dw, db = cl.decode_range(W, B, M, 0, ecb.size)
    assert np.array_equal(dcb, ecb) and np.array_equal(dw, ew) and np.array_equal(db, eb)


def test_host_builder_records_ops_like_the_reference_calls():
    b = host.ColorLeafBuilder()
    b.copy_colors(10, 5)
    b.copy_colors(15, 7)             # contiguous: merged
    b.add(0x11, 1, 2)
    b.add(0x11, 1, 2)                # same colour: merged
    b.add(0x11, 2, 2)
    b.add_large_single_color((1.0, 0.5, 0.0), 70000)
    b.copy_colors(40, 0)             # empty: dropped
    ops = b.ops()
    assert ops.dtype.itemsize == 32 and b.get_color_index() == 12 + 3 + 70000
    assert [(int(o["kind"]), int(o["src_start"]), int(o["count"])) for o in ops] == [(0, 10, 12), (1, 0, 2), (1, 0, 1), (1, 0, 70000)]
    assert int(ops[3]["color_bits"]) == 1023 | (2047 << 10) and int(ops[3]["bits_per_weight"]) == 0
    with pytest.raises(ValueError):
        b.add(1, 4, 2)
    # the recorded list means the same colours as the reference calls
    cb, w, bpw = cl.expand_ops(ops.astype(cl.OP_DTYPE), (np.zeros(0, np.uint32), np.array([0x55 << 32 | 0xFFFF0000], np.uint64), np.array([0, 0], np.uint64), None))
    assert cb.size == b.get_color_index() and (cb[:12] == 0x55).all() and cb[-1] == ops[3]["color_bits"]
