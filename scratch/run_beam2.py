import sys, ctypes as C, numpy as np, time
sys.path.insert(0, '.')
from hashdag_b200 import workloads, camera
from oracle import hdo
fp = int(sys.argv[1]) if len(sys.argv)>1 else 13
scene, poses = workloads.build_workload(17, fp, 64)
lib = C.CDLL('scratch/libbeam.so')
dag = hdo.make_dag(scene, hdo.DAG_HASH)
W,H = 1920,1080
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
d3 = lambda v: (C.c_double*3)(*v)
for pi in (0, 20, 40):
    prm = camera.trace_params(poses[pi], info, 17, W, H)
    for (tw,th) in ((8,4),):
        out = np.zeros(160, np.uint64); paths = np.zeros((H,W,4), np.uint32)
        lib.beam_paths(C.byref(dag), W, H, d3(prm[0]), d3(prm[1]), d3(prm[2]), d3(prm[3]), tw, th, paths.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        print("paths pose", pi, "tile", tw, th, "ray visits", out[0], "with beam", out[1], "ratio", round(float(out[1])/float(out[0]),3), "beam visits", out[2], "mismatch", out[3], "beams", out[4], "/", out[5], "resolved", out[6], "sum-of-max no beam", out[8], "with", out[7], "ratio", round(float(out[7])/float(out[8]),3), "simt eff before", round(float(out[0])/float(out[8])/tw/th,3), "after", round(float(out[1])/float(out[7])/tw/th,3))
        hist = out[16:144].astype(float); cs = np.cumsum(hist)/hist.sum(); print("   beam visits p50", np.searchsorted(cs,0.5), "p90", np.searchsorted(cs,0.9), "p99", np.searchsorted(cs,0.99), "max", np.max(np.nonzero(hist)))
        out = np.zeros(10, np.uint64)
        lib.beam_shadows(C.byref(dag), W, H, d3(prm[0]), d3(prm[1]), d3(prm[2]), d3(prm[3]), tw, th, paths.ctypes.data_as(C.c_void_p), C.c_float(1.0), out.ctypes.data_as(C.c_void_p))
        print("shadow pose", pi, "tile", tw, th, "ray visits", out[0], "with beam", out[1], "ratio", round(float(out[1])/float(out[0]),3), "beam visits", out[2], "mismatch", out[3], "beams", out[4], "/", out[5], "resolved", out[6])
