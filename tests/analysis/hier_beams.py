"""ANALYSIS: visits saved by a second / third level of beams resuming from the tile beam (tests/analysis/hier_beams.cpp).
    python tests/analysis/hier_beams.py [footprint_log2=13]
Cost model: a beam (interval) visit costs ~2.7 ray visits (measured ratio of the kernels' instruction counts, profiles/r1_beams.md)."""
import ctypes as C, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
lib_path = os.path.join(ROOT, "build", "libhier_beams.so")
subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", os.path.join(HERE, "hier_beams.cpp"), "-o", lib_path], check=True)
from hashdag_b200 import camera, workloads
from oracle import hdo
fp = int(sys.argv[1]) if len(sys.argv) > 1 else 13
scene, poses = workloads.build_workload(17, fp, 64)
lib = C.CDLL(lib_path)
dag = hdo.make_dag(scene, hdo.DAG_HASH)
W, H = 1920, 1080
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
d3 = lambda v: (C.c_double * 3)(*v)
K = 2.7
for hier in ([(8, 4)], [(8, 4), (4, 2)], [(8, 4), (2, 2)], [(8, 4), (4, 2), (2, 1)], [(8, 4), (4, 4), (2, 2)], [(8, 4), (2, 1)]):
    tot = np.zeros(16, np.float64)
    for pi in (0, 20, 40):
        prm = camera.trace_params(poses[pi], info, 17, W, H)
        out = np.zeros(16, np.uint64)
        tiles = np.array([v for t in hier for v in t], dtype=np.uint32)
        lib.hier_paths(C.byref(dag), W, H, d3(prm[0]), d3(prm[1]), d3(prm[2]), d3(prm[3]), tiles.ctypes.data_as(C.c_void_p), len(hier), out.ctypes.data_as(C.c_void_p))
        assert out[10] == 0, "a pixel changed"
        tot += out
    n = 3 * W * H
    beams = [tot[2 + k] for k in range(len(hier))]
    cost = tot[1] + K * sum(beams)
    print(f"{str(hier):34s} ray visits/px {tot[1]/n:6.2f} (no beams {tot[0]/n:5.2f})  beam visits/px " + " ".join(f"{b/n:5.2f}" for b in beams) +
          f"  modelled cost/px {cost/n:6.2f}  warp cost (max over 8x4)/px {tot[11]*32/n:6.2f} (no beams {tot[12]*32/n:6.2f})", flush=True)
