import sys, os, faulthandler
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
faulthandler.enable()
import numpy as np
import golden_util as gu
from hashdag_b200 import camera
from oracle import ref, hdo
name = sys.argv[1]
scene = gu.recipe_scene(name)
info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
rt = ref.RefTracer(scene.levels, gu.W, gu.H)
rt.load_scene(scene)
print('loaded', flush=True)
pose = gu.recipe_poses(scene)[0]
prm = camera.trace_params(pose, info, scene.levels, gu.W, gu.H)
for kind,(dk,ck) in (('basic',(0,1)),('hash',(1,3))):
    print(kind, 'paths', rt.resolve_paths(dk, pose, info), flush=True)
    rp = rt.read_paths()
    od = hdo.make_dag(scene, dk)
    op,_ = hdo.trace_paths(od, gu.W, gu.H, prm)
    print(' paths equal oracle', np.array_equal(rp, op), (rp!=op).any(-1).sum(), flush=True)
    if kind == 'hash' and not scene.has_hash_colors: continue
    print(kind, 'colors...', flush=True)
    print(rt.resolve_colors(dk, ck), flush=True)
    rc = rt.read_colors()
    oc,_ = hdo.trace_colors(od, hdo.make_colors(scene, ck), op)
    print(' colors equal oracle', np.array_equal(rc, oc), (rc!=oc).sum(), np.unique(rc)[:8], flush=True)
    print(kind, 'shadows', rt.resolve_shadows(dk, pose, info, 1.0, 0.0), flush=True)
    rs = rt.read_colors()
    os_,_ = hdo.trace_shadows(od, prm, op, oc, 1.0, 0.0)
    print(' shadows equal oracle', np.array_equal(rs, os_), (rs!=os_).sum(), flush=True)
rt.close()
print('closed ok', flush=True)
