import sys, ctypes as C, numpy as np, time
sys.path.insert(0, '.')
from hashdag_b200 import workloads, camera
from oracle import hdo
fp = int(sys.argv[1]) if len(sys.argv)>1 else 13
scene, poses = workloads.build_workload(17, fp, 64)
lib = C.CDLL('scratch/libana.so')
dag = hdo.make_dag(scene, hdo.DAG_HASH)
W,H = 1920,1080
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
d3 = lambda v: (C.c_double*3)(*v)
for pi in (0, 20, 40):
    prm = camera.trace_params(poses[pi], info, 17, W, H)
    for (tw,th) in ((8,4),(4,4),(2,2)):
        out = np.zeros(4, np.uint64); hist = np.zeros(32, np.uint64)
        lib.ana_beam(C.byref(dag), W, H, d3(prm[0]), d3(prm[1]), d3(prm[2]), d3(prm[3]), tw, th, out.ctypes.data_as(C.c_void_p), hist.ctypes.data_as(C.c_void_p))
        print("pose", pi, "tile", tw, th, "visits", out[0], "in common prefix frac", float(out[1])/float(out[0]), "avg prefix", float(out[3])/float(out[2]), "divergence level hist", hist[:18].tolist())
