"""CPU checks of the oracle and of the synthetic scene formats (no GPU).

The reference ships no golden vectors for this path; what it does guarantee is
exercised here as invariants, and the fixtures under tests/golden/ (frames produced by the
reference's own CUDA kernels on a B200, see tests/golden/make_golden.py) pin the oracle itself.
"""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, get_scene, scene_cameras
from hashdag_b200 import camera
from oracle import hdo

W, H = 96, 64


def _params(scene, cam, w=W, h=H):
    return camera.trace_params(cam, camera.DAGInfo(scene.bounds_min, scene.bounds_max), scene.levels, w, h)


@pytest.mark.parametrize("levels,fp", [(12, 10), (13, 10), (17, 10)])
def test_basic_and_hash_dag_give_identical_paths(levels, fp):
    """Invariant (i) of SURVEY.md §8c: same tree, different addressing."""
    s = get_scene(levels, fp)
    db, dh = hdo.make_dag(s, hdo.DAG_BASIC), hdo.make_dag(s, hdo.DAG_HASH)
    any_hit = False
    for cam in scene_cameras(s, 3, fp):
        pb, sb = hdo.trace_paths(db, W, H, _params(s, cam))
        ph, sh = hdo.trace_paths(dh, W, H, _params(s, cam))
        assert np.array_equal(pb, ph)
        assert sb["n_word"] == sh["n_word"] and sh["n_page"] == sh["n_word"] + sh["n_leaf"] and sb["n_page"] == 0
        any_hit |= sb["n_hit"] > 0
    assert any_hit


def test_traced_voxels_are_set_and_in_bounds():
    """Invariant (ii): DAGUtils::get_value(dag, path) holds for every non-null path."""
    s = get_scene(13, 10)
    d = hdo.make_dag(s, hdo.DAG_BASIC)
    for cam in scene_cameras(s, 2, 10):
        p, st = hdo.trace_paths(d, W, H, _params(s, cam))
        ys, xs = np.nonzero(p[..., :3].any(-1))
        assert st["n_hit"] == len(ys)
        assert (p[..., :3] < (1 << s.levels)).all() and (p[..., 3] == 0).all()
        for y, x in list(zip(ys, xs))[::7]:
            assert hdo.get_value(d, *p[y, x, :3])


def test_row_flip_contract():
    """Paths row r holds camera row H-1-r (tracer.cu:251): flipping the up vector flips the frame."""
    s = get_scene(12, 10)
    d = hdo.make_dag(s, hdo.DAG_BASIC)
    cam = scene_cameras(s, 1, 10)[0]
    flipped = camera.CameraView(cam.position, (cam.rotation[0], tuple(-v for v in cam.rotation[1]), cam.rotation[2]))
    a, _ = hdo.trace_paths(d, W, H, _params(s, cam))
    b, _ = hdo.trace_paths(d, W, H, _params(s, flipped))
    # ray directions mirror exactly only up to rounding of rayMin/ddy: compare the hit masks coarsely
    assert (a[..., :3].any(-1) == b[::-1][..., :3].any(-1)).mean() > 0.98


def test_hash_colors_match_basic_colors():
    """Invariant (iv)-style: HashDAG + HashDAGColors reproduce BasicDAG + compressed colours."""
    s = get_scene(13, 10)
    db, dh = hdo.make_dag(s, hdo.DAG_BASIC), hdo.make_dag(s, hdo.DAG_HASH)
    cb, ch = hdo.make_colors(s, hdo.COLORS_COMPRESSED), hdo.make_colors(s, hdo.COLORS_HASH)
    for cam in scene_cameras(s, 2, 10):
        prm = _params(s, cam)
        p, _ = hdo.trace_paths(db, W, H, prm)
        a, _ = hdo.trace_colors(db, cb, p)
        b, _ = hdo.trace_colors(dh, ch, p)
        assert np.array_equal(a, b)
        sa, _ = hdo.trace_shadows(db, prm, p, a, 1.0, 3.0)
        sb, _ = hdo.trace_shadows(dh, prm, p, b, 1.0, 3.0)
        assert np.array_equal(sa, sb)
        assert (a >> 24 == 0xFF).all()
        sky = hdo.trace_colors(db, cb, np.zeros_like(p))[0]
        assert (sky == sky[0, 0]).all() and sky[0, 0] == (0xFF000000 | 186 | (241 << 8) | (249 << 16)) or sky[0, 0] == (0xFF000000 | 187 | (242 << 8) | (250 << 16))


def test_color_stream_has_every_weight_width_and_decodes():
    s = get_scene(13, 10)
    hdr = (s.blocks & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    none = (hdr >> 16) == 0xFFFF
    bpw = np.where(none, 0, ((hdr >> 14) & 3) + 1)
    assert set(np.unique(bpw).tolist()) == {0, 1, 2, 3, 4}
    assert s.macro_blocks.size == 2 * ((s.n_voxels + 16383) // 16384)
    # block colour indices restart at every macro block and increase inside it
    first = s.macro_blocks[0::2].astype(np.int64)
    assert (hdr[first] & 0x3FFF == 0).all()
    leaf = hdo.make_leaf(s)
    import ctypes as C
    for idx in (0, 1, 16383, 16384, s.n_voxels // 2, s.n_voxels - 1):
        c = hdo.lib().hdo_decode_color(C.byref(leaf), int(idx))
        assert c >> 24 == 0xFF


def test_uncompressed_and_error_views():
    s = get_scene(12, 10, uncompressed=True)
    d = hdo.make_dag(s, hdo.DAG_BASIC)
    p, _ = hdo.trace_paths(d, W, H, _params(s, scene_cameras(s, 1, 10)[0]))
    u, _ = hdo.trace_colors(d, hdo.make_colors(s, hdo.COLORS_UNCOMPRESSED), p)
    e, _ = hdo.trace_colors(d, hdo.make_colors(s, hdo.COLORS_ERRORS), p)
    hit = p[..., :3].any(-1)
    assert set(np.unique(e[hit]).tolist()) <= {0xFF000000, 0xFFFFFFFF} and len(np.unique(e[hit])) == 2
    assert len(np.unique(u[hit])) > 50


def test_sun_direction_matches_reference_sass_immediates():
    """The reference compiler folded normalize(0.3,1,0.5) and its reciprocals into the SASS of
    trace_shadows (FFMA ..., 0.2591605..., FMUL ..., 3.858612...).  The oracle computes the same."""
    import struct
    x, y, z = np.float32(0.3), np.float32(1.0), np.float32(0.5)
    ln = np.sqrt(np.float32(np.float32(x * x + y * y) + z * z), dtype=np.float32)
    r = np.float32(1.0) / ln
    sun = (np.float32(r * x), np.float32(r * y), np.float32(r * z))
    assert [float(v) for v in sun] == [0.25916054844856262207, 0.86386841535568237305, 0.43193420767784118652]
    inv = [float(np.float32(1.0) / v) for v in sun]
    assert inv == [3.858612060546875, 1.1575837135314941406, 2.3151674270629882812]


def test_hashdag_layout_rules():
    """Nodes never straddle a 512-word page and leaves are 8-byte aligned (hash_table.h:359-468)."""
    s = get_scene(13, 10)
    d = hdo.make_dag(s, hdo.DAG_HASH)
    pt, pool = s.hash_page_table, s.hash_pool
    assert pt.size == 9 * 1024 * 2 + (s.levels - 9) * 65536 * 8
    used = pt[pt != 0]
    assert used.size == s.hash_pool_top - 1 and np.array_equal(np.sort(used), np.arange(1, s.hash_pool_top))
    # walk a few root-to-leaf chains
    rng = np.random.default_rng(0)
    for _ in range(50):
        ptr, level = s.hash_first_node_index, 0
        while level < s.levels - 2:
            phys = int(pt[ptr // 512]) * 512 + ptr % 512
            mask = int(pool[phys]) & 0xFF
            n = bin(mask).count("1")
            assert ptr % 512 + n < 512
            ptr = int(pool[phys + 1 + rng.integers(n)])
            level += 1
        assert ptr % 2 == 0


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_d[0-9]*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_against_reference_golden_frames(path):
    """Frames written by the reference's own kernels on a B200 (tests/golden/make_golden.py)."""
    from golden_util import check_golden
    check_golden(path, impl="oracle")


def test_golden_fixtures_exist():
    if not GOLDEN:
        pytest.skip("no golden fixtures committed yet (they are produced on the GPU box)")
