"""Would conservative content boxes per node pay?  Visits of the ordered DFS (primary rays) and of the any-hit DFS
(shadow rays) with and without pruning, counted on the CPU oracle.  python tests/analysis/aabb_prune.py [footprint_log2]"""
import os, subprocess, sys, ctypes as C, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                os.path.join(HERE, "aabb_prune.cpp"), "-o", os.path.join(ROOT, "build", "libaabb_prune.so")], check=True)
from hashdag_b200 import workloads, camera
from oracle import hdo
fp = int(sys.argv[1]) if len(sys.argv) > 1 else 14
scene, poses = workloads.build_workload(17, fp, 64)
lib = C.CDLL(os.path.join(ROOT, "build", "libaabb_prune.so"))
dag = hdo.make_dag(scene, hdo.DAG_HASH)
W, H, stride = 1920, 1080, 4
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
d3 = lambda v: (C.c_double * 3)(*v)
for pi in (0, 20, 40):
    prm = camera.trace_params(poses[pi], info, 17, W, H)
    paths, _ = hdo.trace_paths(dag, W, H, prm)
    paths = np.ascontiguousarray(paths)
    for shadow in (0, 1):
        base = None
        for mode, quant, margin in ((0, 0, 0.0), (1, 0, 1.0), (1, 1, 1.0), (2, 0, 1.0), (2, 1, 1.0), (2, 1, 0.25), (1, 1, 0.25)):
            out = np.zeros(4, np.uint64)
            res = np.zeros((H, W, 3), np.uint32)
            t = time.time()
            lib.ana_prune(C.byref(dag), W, H, stride, d3(prm[0]), d3(prm[1]), d3(prm[2]), d3(prm[3]), mode, quant, C.c_float(margin), shadow,
                          paths.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p))
            if base is None:
                base = (int(out[0]), res.copy())
            bad = int((res != base[1]).any(-1).sum())
            n = (W // stride) * (H // stride)
            print(f"pose {pi} {'shadow' if shadow else 'primary'} mode {mode} quant {quant} margin {margin}: visits/ray {out[0] / n:.2f} "
                  f"({out[0] / base[0]:.3f} of the reference), hits {int(out[1])}, mismatches {bad}  [{time.time() - t:.1f}s]", flush=True)
