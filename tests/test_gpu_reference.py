"""GPU: the UNMODIFIED reference kernels (oracle/_ref, built by oracle/build_ref.py) against the
oracle and the product on identical inputs.  This is what pins the oracle."""
import ctypes as C

import numpy as np
import pytest

import golden_util as gu
from hashdag_b200 import camera
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gu.RECIPES))
def test_reference_vs_oracle_vs_product(name):
    from hashdag_b200 import tracer
    with_unc = name in gu.UNCOMPRESSED_RECIPES
    scene = gu.recipe_scene(name, uncompressed=with_unc)
    if not ref.available(scene.levels, gu.W, gu.H):
        pytest.skip("oracle/_ref variant not built")
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.shared(scene, gu.W, gu.H, name, with_uncompressed=with_unc)
    # the reference's own BasicDAG -> HashDAG conversion reproduces our packer bit for bit
    pool, pt, first, top = rt.hash_dag()
    assert first == scene.hash_first_node_index and top == scene.hash_pool_top
    assert np.array_equal(pt, scene.hash_page_table)
    assert np.array_equal(pool, scene.hash_pool)
    if scene.has_hash_colors:
        nodes, offs = rt.hash_colors()
        assert np.array_equal(nodes, scene.color_nodes) and np.array_equal(offs, scene.color_offsets)
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    objs = {"basic": (t, tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene))}
    if scene.has_hash_colors:
        objs["hash"] = (t, tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene))
    for i, pose in enumerate(gu.recipe_poses(scene)):
        for kind, (dk, ck) in (("basic", (0, 1)), ("hash", (1, 3))):
            if kind not in objs:
                continue
            rt.resolve_paths(dk, pose, info)
            rp = rt.read_paths()
            rt.resolve_colors(dk, ck)
            rc = rt.read_colors()
            rt.resolve_shadows(dk, pose, info, 1.0, 0.0)
            rs = rt.read_colors()
            rt.resolve_colors(dk, ck)
            rt.resolve_shadows(dk, pose, info, 1.0, gu.FOG)
            rf = rt.read_colors()
            for impl in ("oracle", "cuda"):
                p, c, s, f = gu.render(impl, scene, kind, pose, gu.FOG, objs[kind])
                tag = f"{name} pose {i} {kind} {impl}"
                assert np.array_equal(p, rp), f"{tag}: {(p != rp).any(-1).sum()} path pixels differ from the reference"
                assert np.array_equal(c, rc), f"{tag}: {(c != rc).sum()} colour pixels differ"
                assert np.array_equal(s, rs), f"{tag}: {(s != rs).sum()} shaded pixels differ"
                assert gu.channel_diff(f, rf) <= 1, f"{tag}: fog differs by more than 1/255"
        if with_unc:
            # BasicDAGUncompressedColors and BasicDAGColorErrors (basic_dag.h:122-242, instantiated at tracer.cu:705-711):
            # the reference's kernels against oracle and product
            rt.resolve_paths(0, pose, info)
            rp = rt.read_paths()
            rt.resolve_colors(0, 0)
            ru = rt.read_colors()
            rt.resolve_colors(0, 2)
            re_ = rt.read_colors()
            assert (ru != rc).any() and (re_ == 0xFFFFFFFF).any() and (re_ == 0xFF000000).any(), "the two views show nothing"
            for impl in ("oracle", "cuda"):
                u, e = gu.render_uncompressed_views(impl, scene, pose, rp, objs["basic"])
                assert np.array_equal(u, ru), f"{name} pose {i} {impl}: {(u != ru).sum()} uncompressed-colour pixels differ from the reference"
                assert np.array_equal(e, re_), f"{name} pose {i} {impl}: {(e != re_).sum()} colour-error pixels differ from the reference"
    t.close()


def test_reference_debug_modes_vs_product():
    from hashdag_b200 import tracer
    scene = gu.recipe_scene("d13")
    if not ref.available(13, gu.W, gu.H):
        pytest.skip("oracle/_ref variant not built")
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.shared(scene, gu.W, gu.H, "d13")
    t = tracer.DAGTracer(True, gu.W, gu.H, 13)
    pose = gu.recipe_poses(scene)[0]
    for dk, ck, dag, col in ((0, 1, tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene)),
                             (1, 3, tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene))):
        rt.resolve_paths(dk, pose, info)
        t.resolve_paths(pose, info, dag)
        for dbg in range(1, 8):
            for lvl in (0, 5, 11):
                rt.resolve_colors(dk, ck, dbg, lvl)
                t.resolve_colors(dag, col, dbg, lvl)
                assert np.array_equal(t.read_colors(), rt.read_colors()), f"debug {dbg} level {lvl} dag {dk}"
    t.close()


def test_config5_depth16_1080p_basic_vs_hash_with_fog():
    """BASELINE.json config 5 at full size: depth-16 BasicDAG vs HashDAG of the same scene, compressed
    colours with every weight width, shadows + fog (density 5), 1920x1080, against the reference
    kernels.  Paths and colours bit-exact, fogged frame within 1/255, Basic == Hash."""
    from hashdag_b200 import tracer
    from conftest import get_scene, scene_cameras
    if not ref.available(16, 1920, 1080):
        pytest.skip("oracle/_ref variant not built")
    W, H = 1920, 1080
    scene = get_scene(16, 12, seed=5, n_spheres=12)
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.shared(scene, W, H, "config5")
    t = tracer.DAGTracer(True, W, H, 16)
    sets = ((0, 1, tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene)),
            (1, 3, tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene)))
    for pose in scene_cameras(scene, 2, 12)[:3]:
        frames = []
        for dk, ck, dag, col in sets:
            rt.resolve_paths(dk, pose, info)
            t.resolve_paths(pose, info, dag)
            rp, p = rt.read_paths(), t.read_paths()
            assert np.array_equal(p, rp), f"{(p != rp).any(-1).sum()} path pixels differ"
            rt.resolve_colors(dk, ck)
            t.resolve_colors(dag, col)
            assert np.array_equal(t.read_colors(), rt.read_colors())
            rt.resolve_shadows(dk, pose, info, 1.0, 5.0)
            t.resolve_shadows(pose, info, dag, 1.0, 5.0)
            rf, f = rt.read_colors(), t.read_colors()
            assert gu.channel_diff(f, rf) <= 1
            assert (f != rf).mean() < 1e-4
            frames.append((p, f))
        assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])
        assert frames[0][0][..., :3].any(-1).mean() > 0.3
    t.close()


def test_tool_overlay_matches_the_reference_built_with_TOOL_OVERLAY():
    """ToolInfo::strength + the lerp towards red (tracer.h:51-76, tracer.cu:276-286).  The BENCHMARK build of
    the reference compiles the overlay out (typedefs.h:70-72); oracle/_ref holds one variant built with
    TOOL_OVERLAY 1 for this test.  Sphere, cube and copy tools, BasicDAG and HashDAG, product and oracle."""
    from hashdag_b200 import tracer
    from oracle import hdo
    scene = gu.recipe_scene("d13")
    if not ref.available(13, gu.W, gu.H, overlay=True):
        pytest.skip("oracle/_ref overlay variant not built")
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.RefTracer(13, gu.W, gu.H, overlay=True)
    assert rt.tool_overlay_compiled()
    rt.load_scene(scene)
    t = tracer.DAGTracer(True, gu.W, gu.H, 13)
    pose = gu.recipe_poses(scene)[0]
    c = 1 << 12
    h0 = scene.heights[(c, c)]
    tools = [(0, (c + 10, h0 + 2, c + 5), 40.0, (0, 0, 0), (0, 0, 0)), (1, (c - 60, h0, c + 30), 7.5, (0, 0, 0), (0, 0, 0)),
             (3, (c + 35, h0 + 5, c - 20), 22.0, (0, 0, 0), (0, 0, 0)), (4, (c, h0, c), 30.0, (c + 50, h0 + 3, c + 50), (c - 45, h0 + 1, c - 40)),
             (5, (c + 5, h0 + 1, c + 5), 300.0, (c, h0, c), (c + 2, h0, c + 2))]
    hit_overlay = 0
    for dk, ck, dag, col, okind in ((0, 1, tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene), hdo.COLORS_COMPRESSED),
                                    (1, 3, tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene), hdo.COLORS_HASH)):
        odag, ocol = hdo.make_dag(scene, dk), hdo.make_colors(scene, okind)
        rt.resolve_paths(dk, pose, info)
        t.resolve_paths(pose, info, dag)
        paths = t.read_paths()
        assert np.array_equal(paths, rt.read_paths())
        rt.resolve_colors(dk, ck)
        plain = rt.read_colors()
        for kind, pos, radius, src, dst in tools:
            rt.resolve_colors_tool(dk, ck, kind, pos, radius, src, dst)
            want = rt.read_colors()
            ti = tracer.ToolInfo(kind, (C.c_uint32 * 3)(*pos), radius, (C.c_uint32 * 3)(*src), (C.c_uint32 * 3)(*dst))
            t.resolve_colors(dag, col, 0, 0, ti, True)
            got = t.read_colors()
            assert np.array_equal(got, want), f"tool {kind} dag {dk}: {(got != want).sum()} pixels differ from the reference"
            oti = hdo.ToolInfo(kind, (C.c_uint32 * 3)(*pos), radius, (C.c_uint32 * 3)(*src), (C.c_uint32 * 3)(*dst))
            oc, _ = hdo.trace_colors(odag, ocol, paths, 0, 0, oti, True)
            assert np.array_equal(oc, want), f"tool {kind} dag {dk}: oracle differs in {(oc != want).sum()} pixels"
            hit_overlay += int((want != plain).sum())
    assert hit_overlay > 1000, "the overlay never showed"
    t.close()
    rt.close()
