"""Golden fixtures: frames produced by the UNMODIFIED reference kernels (oracle/_ref) on a B200.

A fixture stores the scene recipe (the builder is deterministic integer code), the camera poses
and, per pose, the reference's paths / colours / shaded frames for BasicDAG and HashDAG.
"""
import json

import numpy as np

from hashdag_b200 import camera
from hashdag_b200.scene import build_scene

# (levels, footprint_log2, seed, spheres): small enough to rebuild in well under a second
RECIPES = {
    "d12": dict(levels=12, footprint_log2=10, seed=11, n_spheres=5),
    "d13": dict(levels=13, footprint_log2=10, seed=12, n_spheres=5),
    "d17": dict(levels=17, footprint_log2=10, seed=13, n_spheres=5),
}
W = H = 256
FOG = 4.0
UNCOMPRESSED_RECIPES = ("d12",)   # fixtures that also hold the uncompressed-colours and colour-error views (basic_dag.h:122-242)


def recipe_scene(name, uncompressed=False):
    r = RECIPES[name]
    c = 1 << (r["levels"] - 1)
    return build_scene(r["levels"], r["footprint_log2"], seed=r["seed"], n_spheres=r["n_spheres"], build_uncompressed=uncompressed,
                       height_probes=[(c, c)])


def recipe_poses(scene):
    c = float(1 << (scene.levels - 1))
    h0 = float(scene.heights[(int(c), int(c))])
    poses = camera.orbit_poses((c, h0, c), 460.0, 230.0, 3, phase=0.7)
    poses.append(camera.look_at((c + 3.0, h0 + 40.0, c + 5.0), (c + 90.0, h0 - 10.0, c + 70.0)))
    poses.append(camera.CameraView((c, h0 + 500.0, c), ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))))
    return poses


def pose_to_list(p):
    return [list(p.position), [list(r) for r in p.rotation]]


def pose_from_list(l):
    return camera.CameraView(tuple(l[0]), tuple(tuple(r) for r in l[1]))


def render_uncompressed_views(impl, scene, pose, paths, tracer_objs=None):
    """-> (uncompressed colours, colour-error view) of a BasicDAG for the paths frame `paths`, oracle or CUDA product."""
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    if impl == "oracle":
        from oracle import hdo
        d = hdo.make_dag(scene, hdo.DAG_BASIC)
        return (hdo.trace_colors(d, hdo.make_colors(scene, hdo.COLORS_UNCOMPRESSED), paths)[0],
                hdo.trace_colors(d, hdo.make_colors(scene, hdo.COLORS_ERRORS), paths)[0])
    from hashdag_b200 import tracer
    t, dag, comp = tracer_objs
    unc = tracer.BasicDAGUncompressedColors.from_scene(scene)
    t.resolve_paths(pose, info, dag)
    t.resolve_colors(dag, unc)
    u = t.read_colors()
    t.resolve_colors(dag, tracer.BasicDAGColorErrors(comp, unc))
    return u, t.read_colors()


def render(impl, scene, kind, pose, fog, tracer_objs=None):
    """-> (paths[H,W,4], colors, shaded, shaded_fog) with the oracle or the CUDA product."""
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    hash_kind = kind == "hash"
    if impl == "oracle":
        from oracle import hdo
        prm = camera.trace_params(pose, info, scene.levels, W, H)
        d = hdo.make_dag(scene, hdo.DAG_HASH if hash_kind else hdo.DAG_BASIC)
        col = hdo.make_colors(scene, hdo.COLORS_HASH if hash_kind else hdo.COLORS_COMPRESSED)
        p, _ = hdo.trace_paths(d, W, H, prm)
        c, _ = hdo.trace_colors(d, col, p)
        s, _ = hdo.trace_shadows(d, prm, p, c, 1.0, 0.0)
        f, _ = hdo.trace_shadows(d, prm, p, c, 1.0, fog)
        return p, c, s, f
    t, dag, col = tracer_objs
    t.resolve_paths(pose, info, dag)
    p = t.read_paths()
    t.resolve_colors(dag, col)
    c = t.read_colors()
    t.resolve_shadows(pose, info, dag, 1.0, 0.0)
    s = t.read_colors()
    t.resolve_colors(dag, col)
    t.resolve_shadows(pose, info, dag, 1.0, fog)
    f = t.read_colors()
    return p, c, s, f


def channel_diff(a, b):
    a8 = a.view(np.uint8).reshape(a.shape + (4,)).astype(np.int16)
    b8 = b.view(np.uint8).reshape(b.shape + (4,)).astype(np.int16)
    return int(np.abs(a8 - b8).max())


def check_golden(path, impl="oracle"):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    with_unc = bool(meta.get("uncompressed_views"))
    scene = recipe_scene(meta["recipe"], uncompressed=with_unc)
    assert scene.n_voxels == meta["n_voxels"] and scene.basic.size == meta["basic_words"], "scene builder drifted from the fixture"
    kinds = ["basic"] + (["hash"] if meta["has_hash_colors"] else [])
    objs = {}
    if impl == "cuda":
        from hashdag_b200 import tracer
        t = tracer.DAGTracer(True, W, H, scene.levels)
        objs["basic"] = (t, tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene))
        if "hash" in kinds:
            objs["hash"] = (t, tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene))
    for i, pl in enumerate(meta["poses"]):
        pose = pose_from_list(pl)
        for kind in kinds:
            p, c, s, f = render(impl, scene, kind, pose, meta["fog"], objs.get(kind))
            gp = z[f"paths_{i}"]
            assert np.array_equal(p[..., :3], gp), f"pose {i} {kind}: paths differ from the reference in {(p[..., :3] != gp).any(-1).sum()} pixels"
            assert np.array_equal(c, z[f"colors_{i}"]), f"pose {i} {kind}: colours differ"
            assert np.array_equal(s, z[f"shadows_{i}"]), f"pose {i} {kind}: shaded frame differs"
            # fog goes through exp()/pow(): libm vs CUDA may differ in the last ulp -> <= 1/255
            assert channel_diff(f, z[f"fog_{i}"]) <= 1, f"pose {i} {kind}: fogged frame differs by more than 1/255"
            if kind == "basic" and with_unc:
                u, e = render_uncompressed_views(impl, scene, pose, p, objs.get("basic"))
                assert np.array_equal(u, z[f"uncompressed_{i}"]), f"pose {i}: uncompressed colours differ from the reference"
                assert np.array_equal(e, z[f"errors_{i}"]), f"pose {i}: colour-error view differs from the reference"
