"""A/B timing of library builds (same C ABI): per-pass kernel ms of synchronous frames, with an oracle check.

    python scripts/ab_bench.py [--footprint 14] [--poses 16] [--reps 3] [--check 1] [--basic 0] lib_a.so lib_b.so ...

One process, one scene: every library is loaded in turn (ctypes), renders the same poses through hdt_resolve_frame and through
the three separate calls, and is compared with the oracle on two frames.  One JSON line per library.
Environment knobs the libraries read at hdt_create (HDT_BEAMS, HDT_COLORS_RECORDED, ...) apply to all of them."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from hashdag_b200 import camera, tracer, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="*", default=[tracer.LIB_PATH])
ap.add_argument("--footprint", type=int, default=14)
ap.add_argument("--poses", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--basic", type=int, default=0)
ap.add_argument("--resolved", type=int, default=1)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
args = ap.parse_args()
W, H = args.width, args.height
scene, poses = workloads.build_workload(17, args.footprint, args.poses)
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
oracle_frames = {}


def oracle_frame(kind, pose_id):
    from oracle import hdo
    key = (kind, pose_id)
    if key not in oracle_frames:
        od = hdo.make_dag(scene, hdo.DAG_HASH if kind == "hash" else hdo.DAG_BASIC)
        oc = hdo.make_colors(scene, hdo.COLORS_HASH if kind == "hash" else hdo.COLORS_COMPRESSED)
        prm = camera.trace_params(poses[pose_id], info, 17, W, H)
        op, _ = hdo.trace_paths(od, W, H, prm)
        c, _ = hdo.trace_colors(od, oc, op)
        s, _ = hdo.trace_shadows(od, prm, op, c, 1.0, 0.0)
        oracle_frames[key] = (op, s)
    return oracle_frames[key]


for lib in args.libs:
    tracer._lib, tracer.LIB_PATH = None, os.path.abspath(lib)
    t = tracer.DAGTracer(True, W, H, 17)
    out = {"lib": os.path.basename(lib)}
    for kind in ("hash", "basic") if args.basic else ("hash",):
        if kind == "hash":
            dag, col = tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene)
            if args.resolved:
                dag = t.resolve_hash_dag(dag)
                t.sync()
        else:
            dag, col = tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene)
        for p in poses[:4]:
            t.resolve_frame(p, info, dag, col)
        frame = np.zeros(3)
        calls = np.zeros(3)
        for rep in range(args.reps):
            for p in poses:
                frame += np.array(t.resolve_frame(p, info, dag, col))
            for p in poses:
                calls += np.array([t.resolve_paths(p, info, dag), t.resolve_colors(dag, col), t.resolve_shadows(p, info, dag, 1.0, 0.0)])
        frame /= args.reps * len(poses)
        calls /= args.reps * len(poses)
        out[kind] = {"frame_call": {"paths_ms": round(frame[0], 4), "colors_ms": round(frame[1], 4), "shadows_ms": round(frame[2], 4), "frame_ms": round(frame.sum(), 4)},
                     "three_calls": {"paths_ms": round(calls[0], 4), "colors_ms": round(calls[1], 4), "shadows_ms": round(calls[2], 4), "frame_ms": round(calls.sum(), 4)},
                     "recorded_color_passes": t.recorded_color_passes()}
        if args.check:
            bad = 0
            for pid in (0, len(poses) // 2):
                t.resolve_frame(poses[pid], info, dag, col)
                gp, gc = t.read_paths(), t.read_colors()
                op, s = oracle_frame(kind, pid)
                bad += int((gp != op).any(-1).sum()) + int((gc != s).sum())
            out[kind]["mismatch_vs_oracle"] = bad
        del dag, col
    print(json.dumps(out), flush=True)
    t.close()
