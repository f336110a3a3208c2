"""Per-pass device timelines and beam usage: HDT_LIB=build/libhdt_dbg.so python scripts/timeline.py [F]"""
import os, sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np
from hashdag_b200 import camera, tracer, workloads
F = int(sys.argv[1]) if len(sys.argv) > 1 else 13
W, H = 1920, 1080
scene, poses = workloads.build_workload(17, F, 16)
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
t = tracer.DAGTracer(True, W, H, 17)
dag, col = tracer.HashDAG.from_scene(scene), tracer.HashDAGColors.from_scene(scene)
lib = tracer.load_library()
dbg = hasattr(lib, "hdt_debug_beam_counters")
def counters(reset=1):
    out = (C.c_uint32 * 8)()
    lib.hdt_debug_beam_counters(out, reset)
    return list(out)
for p in poses[:4]:
    t.resolve_frame(p, info, dag, col)
for mode in ("separate calls", "whole frame", "serial"):
    for cap in (32,):
        t.set_option(tracer.OPT_BEAM_MAX_VISITS, cap)
        if dbg: counters()
        acc = np.zeros((2, 3)); n = 0
        for p in poses[4:12]:
            t.set_option(tracer.OPT_BEAM_SERIAL, 1 if mode == "serial" else 0)
            if mode in ("separate calls", "serial"):
                t.resolve_paths(p, info, dag); t.resolve_colors(dag, col); t.resolve_shadows(p, info, dag, 1.0, 0.0)
            else:
                t.resolve_frame(p, info, dag, col)
            acc[0] += t.pass_timeline(0); acc[1] += t.pass_timeline(1); n += 1
        acc /= n
        print(mode, "cap", cap, "paths: setup_end %.3f beam_end %.3f trace_end %.3f | shadows: setup_end %.3f beam_end %.3f trace_end %.3f" % (*acc[0], *acc[1]),
              "warps by status [none,resume,hit,miss] paths/shadows:", counters() if dbg else None, flush=True)
t.set_option(tracer.OPT_BEAM_SERIAL, 0)
t.set_option(tracer.OPT_BEAMS, 0)
acc = np.zeros((2, 3)); n = 0
for p in poses[4:12]:
    t.resolve_paths(p, info, dag); t.resolve_colors(dag, col); t.resolve_shadows(p, info, dag, 1.0, 0.0)
    acc[0] += t.pass_timeline(0); acc[1] += t.pass_timeline(1); n += 1
print("beams off:", (acc / n).round(3).tolist())
