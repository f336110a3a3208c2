"""A/B timing of library builds: HDT_LIB=<path> [AB_RESOLVED=0] [AB_CHECK=0] python scripts/ab_bench.py [F] [poses]"""
import os, sys, time, json
sys.path.insert(0, '.')
import numpy as np
from hashdag_b200 import camera, tracer, workloads
F = int(sys.argv[1]) if len(sys.argv) > 1 else 13
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 16
check = int(os.environ.get("AB_CHECK", "1"))
W, H = 1920, 1080
scene, poses = workloads.build_workload(17, F, NP)
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
t = tracer.DAGTracer(True, W, H, 17)
out = {"lib": os.path.basename(tracer.LIB_PATH), "persistent": os.environ.get("HDT_PERSISTENT", "0")}
for kind in ("hash", "basic"):
    if kind == "hash":
        dag, col = tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene)
        if os.environ.get("AB_RESOLVED", "1") != "0":      # the shipped configuration: resolved pool (hdt_hash_dag_resolve)
            dag = t.resolve_hash_dag(dag)
            t.sync()
    else:
        dag, col = tracer.BasicDAG.from_scene(scene), tracer.BasicDAGCompressedColors.from_scene(scene)
    for p in poses[:4]:
        t.resolve_frame(p, info, dag, col)
    acc = np.zeros(3)
    for rep in range(3):
        for p in poses:
            acc += np.array(t.resolve_frame(p, info, dag, col))
    acc /= 3 * len(poses)
    out[kind] = {"paths_ms": round(acc[0], 4), "colors_ms": round(acc[1], 4), "shadows_ms": round(acc[2], 4), "frame_ms": round(acc.sum(), 4)}
    if check:
        from oracle import hdo
        od = hdo.make_dag(scene, hdo.DAG_HASH if kind == "hash" else hdo.DAG_BASIC)
        oc = hdo.make_colors(scene, hdo.COLORS_HASH if kind == "hash" else hdo.COLORS_COMPRESSED)
        bad = 0
        for p in poses[:2]:
            prm = camera.trace_params(p, info, 17, W, H)
            t.resolve_frame(p, info, dag, col)
            gp, gc = t.read_paths(), t.read_colors()
            op, _ = hdo.trace_paths(od, W, H, prm)
            c, _ = hdo.trace_colors(od, oc, op)
            s, _ = hdo.trace_shadows(od, prm, op, c, 1.0, 0.0)
            bad += int((gp != op).any(-1).sum()) + int((gc != s).sum())
        out[kind]["mismatch_vs_oracle"] = bad
print(json.dumps(out), flush=True)
