"""GPU: colour-leaf rebuild (hdt_rebuild_color_leaf, csrc/hdt_color_leaf.cuh) against the oracle
(oracle/color_leaf.py), against leaves the reference's ColorLeafBuilder wrote (fixture + live oracle/_ref), and
through trace_colors."""
import json
import os

import numpy as np
import pytest

import golden_util as gu
from conftest import ROOT
from hashdag_b200 import camera, color_leaf as host, tracer
from oracle import color_leaf as cl, ref
from test_color_leaf_cpu import GOLDEN, golden_leaves, random_stream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def t():
    tr = tracer.DAGTracer(True, gu.W, gu.H, 13)
    yield tr
    tr.close()


def to_host(leaf):
    f = lambda x, dt: np.zeros(0, dt) if x is None else x.cpu().numpy().view(dt)
    return f(leaf.weights, np.uint32), f(leaf.blocks, np.uint64), f(leaf.macro_blocks, np.uint64)


def assert_same(got, want, what=""):
    for name, a, b in zip(("weights", "blocks", "macro_blocks"), got, want):
        assert a.shape == b.shape, f"{what}: {name} has {a.size} entries, the oracle {b.size}"
        assert np.array_equal(a, b), f"{what}: {name} differ at {np.flatnonzero(a != b)[:5]}"


def fill_ops(cb, w, bpw):
    ops = np.zeros(cb.size, dtype=host.OP_DTYPE)
    ops["count"], ops["kind"], ops["bits_per_weight"], ops["color_bits"], ops["weight"] = 1, host.OP_FILL, bpw, cb, w
    return ops


@pytest.mark.parametrize("n,run", [(1, 1.0), (15, 2.0), (16, 1.0), (1023, 3.0), (16384, 2.5), (16385, 1.1), (49152, 40.0), (250000, 6.0)])
def test_encoder_equals_oracle_on_random_streams(t, n, run):
    rng = np.random.default_rng(n)
    cb, w, bpw = random_stream(rng, n, run)
    leaf, ms = t.rebuild_color_leaf(fill_ops(cb, w, bpw))
    assert_same(to_host(leaf), cl.encode(cb, w, bpw), f"n={n}")


@pytest.mark.parametrize("bpw_choices", [(0,), (4,), (1, 3), (0, 2)])
def test_encoder_weight_widths(t, bpw_choices):
    rng = np.random.default_rng(11)
    cb, w, bpw = random_stream(rng, 70001, 3.0, bpw_choices)
    leaf, _ = t.rebuild_color_leaf(fill_ops(cb, w, bpw))
    assert_same(to_host(leaf), cl.encode(cb, w, bpw), str(bpw_choices))


def test_copy_and_fill_ops_from_a_unique_and_a_shared_leaf(t):
    scene = gu.recipe_scene("d13")
    old_host = (scene.weights, scene.blocks, scene.macro_blocks)
    n = int(scene.n_voxels)
    for offset in (None, 123457):
        old = tracer.CompressedColorLeaf.from_scene(scene)
        if offset is not None:
            old.offset = offset
        b = host.ColorLeafBuilder()
        b.copy_colors(0, 70000)
        b.add(0xDEADBEEF, 5, 3)
        b.add_large_single_color((0.2, 0.9, 0.4), 40000)
        b.copy_colors(90000, 300001)
        b.add(0x00C0FFEE, 0, 0)
        b.copy_colors(500000, n - 500000 - (offset or 0))
        leaf, ms = b.build(t, old)
        want = cl.rebuild(b.ops().astype(cl.OP_DTYPE), old_host + (offset,))
        assert_same(to_host(leaf), want, f"offset={offset}")
        assert ms > 0


def test_long_runs_with_weights_and_short_copies(t):
    # pieces (runs of one old block / one FILL) of every length: repeated weights of each width over several macro blocks
    # (periodic bit patterns, 3 bits does not divide a word), copies of a few colours each, copies over old macro boundaries
    scene = gu.recipe_scene("d13")
    old_host = (scene.weights, scene.blocks, scene.macro_blocks)
    n = int(scene.n_voxels)
    rng = np.random.default_rng(77)
    rows = []
    for bpw, count in ((3, 5000), (1, 40001), (4, 16384), (2, 7), (3, 33333), (0, 20000), (4, 1)):
        rows.append((0, count, host.OP_FILL, bpw, int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << bpw))))
        src = int(rng.integers(0, n - 40000))
        rows.append((src, int(rng.integers(1, 40)), host.OP_COPY, 0, 0, 0))
        rows.append((16384 * int(rng.integers(1, n // 16384 - 2)) - 11, 30000, host.OP_COPY, 0, 0, 0))
    for k in range(700):       # more segments in one macro block than a CTA has threads
        rows.append((int(rng.integers(0, n - 5000)), int(rng.integers(1, 30)), host.OP_COPY, 0, 0, 0))
        rows.append((0, int(rng.integers(1, 4)), host.OP_FILL, 2, 0xABC00000 + k % 3, k % 4))
    ops = np.array(rows, dtype=host.OP_DTYPE)
    for offset in (None, 4097):
        old = tracer.CompressedColorLeaf.from_scene(scene)
        if offset is not None:
            old.offset = offset
        leaf, _ = t.rebuild_color_leaf(ops, old)
        assert_same(to_host(leaf), cl.rebuild(ops.astype(cl.OP_DTYPE), old_host + (offset,)), f"offset={offset}")


@pytest.mark.parametrize("recipe", ["d13", "d17"])
def test_reference_leaves_survive_decode_and_reencode(t, recipe):
    """COPY(0, n) of a leaf the reference built must give that leaf back, byte for byte."""
    _, leaves = golden_leaves(recipe)
    for k, (w, b, m) in enumerate(leaves):
        old = tracer.CompressedColorLeaf(tracer._to_device(w, "cuda:0"), tracer._to_device(b, "cuda:0"), tracer._to_device(m, "cuda:0"))
        bld = host.ColorLeafBuilder()
        bld.copy_colors(0, cl.leaf_color_count(w, b, m))
        leaf, _ = bld.build(t, old)
        assert_same(to_host(leaf), (w, b, m), f"golden leaf {k}")


@pytest.mark.parametrize("recipe", ["d13", "d17"])
def test_every_leaf_of_the_live_reference_scenario(t, recipe):
    if not ref.available(int(recipe[1:]), 256, 256):
        pytest.skip("oracle/_ref variant not built")
    import subprocess, sys
    # own process: the reference keeps its scene in globals
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "color_leaf_scenario.py"), recipe], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("COLOR_LEAF_SCENARIO ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.loads(line[-1][len("COLOR_LEAF_SCENARIO "):])
    meta = json.loads(str(np.load(GOLDEN[recipe])["meta"]))
    assert rep["n_leaves"] == meta["n_leaves_in_scenario"] and rep["digests_match_fixture"]
    assert rep["gpu_mismatches"] == 0 and rep["oracle_mismatches"] == 0 and rep["n_leaves"] >= (50 if recipe == "d13" else 5)


def test_rebuilt_main_leaf_renders_the_same_frame(t):
    """End to end: the whole colour leaf of a scene re-encoded on the GPU, then traced."""
    scene = gu.recipe_scene("d13")
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    dag = tracer.BasicDAG.from_scene(scene)
    col = tracer.BasicDAGCompressedColors.from_scene(scene)
    b = host.ColorLeafBuilder()
    b.copy_colors(0, int(scene.n_voxels))
    leaf, ms = b.build(t, col.leaf)
    nw = 0 if leaf.weights is None else leaf.weights.numel()
    assert leaf.blocks.numel() == scene.blocks.size and nw <= scene.weights.size   # the scene builder word-aligns macro blocks, the reference does not
    col2 = tracer.BasicDAGCompressedColors(col.top_levels, col.enclosed_leaves, leaf)
    for pose in gu.recipe_poses(scene)[:3]:
        t.resolve_paths(pose, info, dag)
        t.resolve_colors(dag, col)
        a = t.read_colors()
        t.resolve_colors(dag, col2)
        assert np.array_equal(a, t.read_colors())


def test_capacity_and_argument_errors(t):
    import ctypes as C
    import torch
    lib = tracer.load_library()
    ops = fill_ops(*random_stream(np.random.default_rng(3), 5000, 1.0, (4,)))
    counts = (C.c_uint64 * 4)()
    ms = C.c_float()
    small = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    rc = lib.hdt_rebuild_color_leaf(t._ctx, None, 0, ops.ctypes.data, ops.size, small.data_ptr(), 16, small.data_ptr(), 16, small.data_ptr(), 16, counts, C.byref(ms))
    assert rc == tracer.ERR_CAPACITY and counts[0] == 5000 and counts[1] == (4 * 5000 + 31) // 32 and counts[3] == 2
    assert (small == 0).all()
    copy = np.array([(0, 10, host.OP_COPY, 0, 0, 0)], dtype=host.OP_DTYPE)
    with pytest.raises(tracer.TracerError):
        t.rebuild_color_leaf(copy, None)
    bad = np.array([(0, 10, host.OP_FILL, 2, 1, 7)], dtype=host.OP_DTYPE)
    with pytest.raises(tracer.TracerError):
        t.rebuild_color_leaf(bad, None)
    leaf, ms0 = t.rebuild_color_leaf(np.zeros(0, dtype=host.OP_DTYPE))
    assert leaf.blocks is None and ms0 == 0.0
