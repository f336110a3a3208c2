"""Visits of the reference DFS per level (how much of a frame is top-of-tree work every ray of a tile shares)."""
import os, subprocess, sys, ctypes as C, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                os.path.join(HERE, "visit_histogram.cpp"), "-o", os.path.join(ROOT, "build", "libvisit_histogram.so")], check=True)
from hashdag_b200 import workloads, camera
from oracle import hdo
t=time.time()
fp = int(sys.argv[1]) if len(sys.argv)>1 else 13
scene, poses = workloads.build_workload(17, fp, 64)
print("build", time.time()-t, scene.n_voxels, scene.nodes_per_level)
lib = C.CDLL(os.path.join(ROOT, 'build', 'libvisit_histogram.so'))
dag = hdo.make_dag(scene, hdo.DAG_HASH)
W,H = 1920//2,1080//2
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
for pi in (0, 20, 40):
    prm = camera.trace_params(poses[pi], info, 17, W, H)
    vis = np.zeros(32, np.uint64); onp = np.zeros(32, np.uint64); rs = np.zeros(1024, np.uint64); cc = np.zeros(9, np.uint64); vc=np.zeros(9,np.uint64)
    d3 = lambda v: (C.c_double*3)(*v)
    lib.ana_paths(C.byref(dag), W, H, d3(prm[0]), d3(prm[1]), d3(prm[2]), d3(prm[3]), vis.ctypes.data_as(C.c_void_p), onp.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p), cc.ctypes.data_as(C.c_void_p), vc.ctypes.data_as(C.c_void_p))
    n = W*H
    print("pose", pi, "visits/ray", vis.sum()/n, "hits", onp[17]/n)
    print(" by level visits/ray:", [round(float(v)/n,2) for v in vis[1:18]])
    cs = np.cumsum(rs); print(" steps pct: p50", np.searchsorted(cs, n*0.5), "p90", np.searchsorted(cs, n*0.9), "p99", np.searchsorted(cs, n*0.99), "max", np.max(np.nonzero(rs)))
    print(" childCount hist", (cc/cc.sum()).round(3), "vm hist", (vc/vc.sum()).round(3))
