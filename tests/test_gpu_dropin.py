"""Drop-in proof: oracle/ref_harness/ref_dropin.cu includes the reference's own headers, aliases
DAGTracer to hashdag_b200::DAGTracer and makes the calls of Engine::tick (engine.cpp:575-648) with the reference's real
BasicDAG / HashDAG / colour / CameraView / DAGInfo / ToolInfo objects.  That it compiles (with the static_asserts on every
struct the library mirrors) is checked by oracle/build_ref.py; here the frames it renders are compared with the frames the
reference's own DAGTracer renders from the very same objects."""
import numpy as np
import pytest

import golden_util as gu
from hashdag_b200 import camera
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dags", [("d12", (0, 1, 2)), ("d13", (1, 3))])
def test_shim_with_reference_types_renders_the_reference_frames(name, dags):
    with_unc = name in gu.UNCOMPRESSED_RECIPES
    scene = gu.recipe_scene(name, uncompressed=with_unc)
    if not ref.available(scene.levels, gu.W, gu.H):
        pytest.skip("oracle/_ref variant not built")
    rt = ref.shared(scene, gu.W, gu.H, name, with_uncompressed=with_unc)
    if not hasattr(rt.lib, "ref_dropin_tick"):
        pytest.skip("oracle/_ref built without the drop-in proof")
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    for pose in gu.recipe_poses(scene)[:3]:
        for current_dag in dags:
            dk, ck = (1, 3) if current_dag == 3 else (0, current_dag)
            for dbg, lvl, fog in ((0, 0, 0.0), (7, 0, 0.0), (1, 3, 0.0), (0, 0, gu.FOG)):
                rt.resolve_paths(dk, pose, info)
                rp = rt.read_paths()
                hit = np.argwhere(rp[..., :3].any(-1))
                py, px = (int(v) for v in hit[len(hit) // 2])
                rt.resolve_colors(dk, ck, dbg, lvl)
                rt.resolve_shadows(dk, pose, info, 1.0, fog)
                rc = rt.read_colors()
                times, path = rt.dropin_tick(current_dag, pose, info, dbg, lvl, True, 1.0, fog, (px, py))
                assert all(t > 0 for t in times)
                p, c = rt.dropin_read_paths(), rt.dropin_read_colors()
                assert np.array_equal(p, rp), f"{name} dag {current_dag}: {(p != rp).any(-1).sum()} path pixels differ from the reference's DAGTracer"
                assert path == tuple(int(v) for v in rp[py, px, :3]), "get_path"
                if fog == 0.0:
                    assert np.array_equal(c, rc), f"{name} dag {current_dag} debug {dbg}: {(c != rc).sum()} pixels differ from the reference's DAGTracer"
                else:
                    assert gu.channel_diff(c, rc) <= 1
