"""Parity soak: the CUDA product against the CPU oracle on many seeded random camera poses (inside and outside the scene, grazing
and axis-parallel directions), HashDAG (plain and resolved + prefix pool) and BasicDAG.  Prints one JSON line.

    python scripts/parity_soak.py [--poses 150] [--seed 1]
    python scripts/parity_soak.py --reference d13|d17 [--poses 150]      # against the REFERENCE's kernels (oracle/_ref), own process per scene
"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def random_poses(rng, scene, footprint_log2, n):
    from hashdag_b200 import camera
    c = float(1 << (scene.levels - 1))
    h0 = float(scene.heights.get((int(c), int(c)), c))
    r = float(1 << footprint_log2)
    poses = []
    for k in range(n):
        kind = k % 5
        pos = (c + rng.uniform(-0.6, 0.6) * r, h0 + rng.uniform(-0.05, 0.6) * r, c + rng.uniform(-0.6, 0.6) * r)
        if kind == 0:      # look at a point near the ground
            tgt = (c + rng.uniform(-0.5, 0.5) * r, h0 + rng.uniform(-0.1, 0.1) * r, c + rng.uniform(-0.5, 0.5) * r)
            poses.append(camera.look_at(pos, tgt))
        elif kind == 1:    # any direction
            d = rng.normal(size=3)
            poses.append(camera.look_at(pos, tuple(np.array(pos) + d / np.linalg.norm(d) * 100.0)))
        elif kind == 2:    # axis-parallel view directions: zero components in the centre ray, inf / NaN slabs
            ax = int(rng.integers(0, 3)); sg = 1.0 if rng.integers(0, 2) else -1.0
            fwd = [0.0, 0.0, 0.0]; fwd[ax] = sg
            up = [0.0, 0.0, 0.0]; up[(ax + 1) % 3] = 1.0
            right = list(np.cross(fwd, up))
            poses.append(camera.CameraView(pos, (tuple(right), tuple(up), tuple(fwd))))
        elif kind == 3:    # grazing: almost horizontal, just above the ground
            a = rng.uniform(0, 2 * math.pi)
            p2 = (pos[0], h0 + rng.uniform(1.0, 30.0), pos[2])
            poses.append(camera.look_at(p2, (p2[0] + math.cos(a) * 500.0, p2[1] - rng.uniform(0.0, 3.0), p2[2] + math.sin(a) * 500.0)))
        else:              # far outside the volume, looking in
            far = (c + rng.choice([-1.0, 1.0]) * r * rng.uniform(1.5, 40.0), h0 + rng.uniform(0.0, 3.0) * r, c + rng.choice([-1.0, 1.0]) * r * rng.uniform(1.5, 40.0))
            poses.append(camera.look_at(far, (c, h0, c)))
    return poses


def reference_soak(a):
    """Product and oracle against the reference's own kernels (oracle/_ref, 256x256 variants) on the recipe scene `a.reference`."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    from hashdag_b200 import camera, tracer
    from oracle import hdo, ref
    scene = gu.recipe_scene(a.reference)
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.RefTracer(scene.levels, gu.W, gu.H)
    rt.load_scene(scene)
    t = tracer.DAGTracer(True, gu.W, gu.H, scene.levels)
    rng = np.random.default_rng(a.seed * 77 + scene.levels)
    poses = random_poses(rng, scene, gu.RECIPES[a.reference]["footprint_log2"], a.poses)
    hd = tracer.HashDAG.from_scene(scene)
    cases = [("basic", 0, 1, tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene), hdo.make_dag(scene, hdo.DAG_BASIC), hdo.make_colors(scene, hdo.COLORS_COMPRESSED))]
    if scene.has_hash_colors:
        cases.append(("hash resolved + prefix", 1, 3, t.resolve_hash_dag(hd), tracer.HashDAGColors.from_scene(scene), hdo.make_dag(scene, hdo.DAG_HASH), hdo.make_colors(scene, hdo.COLORS_HASH)))
    rep = {"reference_scene": a.reference, "poses_per_case": a.poses, "resolution": [gu.W, gu.H], "cases": [], "mismatched_pixels": 0, "fog_max_channel_diff": 0}
    for name, dk, ck, dag, col, odag, ocol in cases:
        bad_cuda = bad_oracle = hits = fog = 0
        for cam in poses:
            prm = camera.trace_params(cam, info, scene.levels, gu.W, gu.H)
            rt.resolve_paths(dk, cam, info); rp = rt.read_paths()
            rt.resolve_colors(dk, ck); rc = rt.read_colors()
            rt.resolve_shadows(dk, cam, info, 1.0, 0.0); rs = rt.read_colors()
            rt.resolve_colors(dk, ck)
            rt.resolve_shadows(dk, cam, info, 2.5, 5.0); rf = rt.read_colors()
            t.resolve_paths(cam, info, dag); p = t.read_paths()
            t.resolve_colors(dag, col); c = t.read_colors()
            t.resolve_shadows(cam, info, dag, 1.0, 0.0); sh = t.read_colors()
            t.resolve_colors(dag, col)
            t.resolve_shadows(cam, info, dag, 2.5, 5.0); fg = t.read_colors()
            op, st = hdo.trace_paths(odag, gu.W, gu.H, prm)
            oc, _ = hdo.trace_colors(odag, ocol, op)
            osh, _ = hdo.trace_shadows(odag, prm, op, oc, 1.0, 0.0)
            hits += int(st["n_hit"])
            bad_cuda += int((p != rp).any(-1).sum()) + int((c != rc).sum()) + int((sh != rs).sum())
            bad_oracle += int((op != rp).any(-1).sum()) + int((oc != rc).sum()) + int((osh != rs).sum())
            fog = max(fog, int(np.abs(fg.view(np.uint8).astype(np.int16) - rf.view(np.uint8).astype(np.int16)).max()))
        rep["cases"].append({"dag": name, "hit_pixels": hits, "cuda_vs_reference_kernels": bad_cuda, "oracle_vs_reference_kernels": bad_oracle, "fog_max_channel_diff": fog})
        rep["mismatched_pixels"] += bad_cuda + bad_oracle
        rep["fog_max_channel_diff"] = max(rep["fog_max_channel_diff"], fog)
        print(rep["cases"][-1], file=sys.stderr, flush=True)
    t.close()
    rt.close()
    print("PARITY_SOAK " + json.dumps(rep), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="", help="recipe scene (d13, d17): compare with the reference's kernels instead of the oracle only")
    ap.add_argument("--poses", type=int, default=150)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=200)
    a = ap.parse_args()
    if a.reference:
        return reference_soak(a)
    from hashdag_b200 import camera, tracer
    from hashdag_b200.scene import build_scene
    from oracle import hdo
    W, H = a.width, a.height
    rep = {"poses_per_case": a.poses, "resolution": [W, H], "cases": [], "mismatched_pixels": 0, "fog_max_channel_diff": 0}
    for levels, fp, seed in ((13, 10, 12), (17, 10, 13), (16, 11, 5)):
        scene = build_scene(levels, fp, seed=seed, n_spheres=5, height_probes=[(1 << (levels - 1), 1 << (levels - 1))])
        info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
        t = tracer.DAGTracer(True, W, H, levels)
        rng = np.random.default_rng(a.seed * 1000 + levels)
        poses = random_poses(rng, scene, fp, a.poses)
        variants = [("basic", tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene), hdo.make_dag(scene, hdo.DAG_BASIC), hdo.make_colors(scene, hdo.COLORS_COMPRESSED))]
        if scene.has_hash_colors:
            hd = tracer.HashDAG.from_scene(scene)
            hc = tracer.HashDAGColors.from_scene(scene)
            od, oc = hdo.make_dag(scene, hdo.DAG_HASH), hdo.make_colors(scene, hdo.COLORS_HASH)
            variants.append(("hash", hd, hc, od, oc))
            variants.append(("hash resolved + prefix", t.resolve_hash_dag(hd), hc, od, oc))
        for name, dag, col, odag, ocol in variants:
            bad = hits = 0
            fog = 0
            for cam in poses:
                prm = camera.trace_params(cam, info, levels, W, H)
                t.resolve_paths(cam, info, dag)
                p = t.read_paths()
                op, st = hdo.trace_paths(odag, W, H, prm)
                bad += int((p != op).any(-1).sum())
                hits += int(st["n_hit"])
                t.resolve_colors(dag, col)
                oc_, _ = hdo.trace_colors(odag, ocol, op)
                bad += int((t.read_colors() != oc_).sum())
                t.resolve_shadows(cam, info, dag, 1.0, 0.0)
                osh, _ = hdo.trace_shadows(odag, prm, op, oc_, 1.0, 0.0)
                bad += int((t.read_colors() != osh).sum())
                t.resolve_colors(dag, col)
                t.resolve_shadows(cam, info, dag, 2.5, 5.0)
                ofg, _ = hdo.trace_shadows(odag, prm, op, oc_, 2.5, 5.0)
                fog = max(fog, int(np.abs(t.read_colors().view(np.uint8).astype(np.int16) - ofg.view(np.uint8).astype(np.int16)).max()))
            rep["cases"].append({"levels": levels, "dag": name, "hit_pixels": hits, "mismatched_pixels": bad, "fog_max_channel_diff": fog})
            rep["mismatched_pixels"] += bad
            rep["fog_max_channel_diff"] = max(rep["fog_max_channel_diff"], fog)
            print(rep["cases"][-1], file=sys.stderr, flush=True)
        t.close()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
