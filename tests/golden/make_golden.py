"""Generate tests/golden/*.npz with the reference's own kernels.  Run on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  Needs oracle/_ref (oracle/build_ref.py).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util as gu  # noqa: E402
from hashdag_b200 import camera  # noqa: E402
from oracle import ref  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name in gu.RECIPES:
        with_unc = name in gu.UNCOMPRESSED_RECIPES     # BasicDAGUncompressedColors / BasicDAGColorErrors views too (basic_dag.h:122-242)
        scene = gu.recipe_scene(name, uncompressed=with_unc)
        info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
        rt = ref.shared(scene, gu.W, gu.H, name, with_uncompressed=with_unc)
        has_hash_colors = scene.levels - 2 > 10
        poses = gu.recipe_poses(scene)
        arrays = {}
        for i, pose in enumerate(poses):
            frames = {}
            for kind, (dk, ck) in (("basic", (0, 1)), ("hash", (1, 3))):
                if kind == "hash" and not has_hash_colors:
                    rt.resolve_paths(1, pose, info)
                    assert np.array_equal(rt.read_paths(), frames["basic"][0]), "reference: HashDAG paths != BasicDAG paths"
                    continue
                rt.resolve_paths(dk, pose, info)
                p = rt.read_paths()
                rt.resolve_colors(dk, ck)
                c = rt.read_colors()
                rt.resolve_shadows(dk, pose, info, 1.0, 0.0)
                s = rt.read_colors()
                rt.resolve_colors(dk, ck)
                rt.resolve_shadows(dk, pose, info, 1.0, gu.FOG)
                f = rt.read_colors()
                frames[kind] = (p, c, s, f)
            if "hash" in frames:
                for a, b in zip(frames["basic"], frames["hash"]):
                    assert np.array_equal(a, b), "reference: HashDAG frame != BasicDAG frame"
            p, c, s, f = frames["basic"]
            assert (p[..., 3] == 0).all()
            arrays[f"paths_{i}"] = p[..., :3].copy()
            arrays[f"colors_{i}"], arrays[f"shadows_{i}"], arrays[f"fog_{i}"] = c, s, f
            if with_unc:
                rt.resolve_paths(0, pose, info)
                rt.resolve_colors(0, 0)
                arrays[f"uncompressed_{i}"] = rt.read_colors()
                rt.resolve_colors(0, 2)
                arrays[f"errors_{i}"] = rt.read_colors()
                assert (arrays[f"uncompressed_{i}"] != c).any(), "uncompressed colours never differ from the compressed ones: nothing is pinned"
            print(name, "pose", i, "hits", int(p[..., :3].any(-1).sum()), flush=True)
        meta = dict(recipe=name, n_voxels=int(scene.n_voxels), basic_words=int(scene.basic.size), has_hash_colors=bool(has_hash_colors),
                    fog=gu.FOG, uncompressed_views=bool(with_unc), poses=[gu.pose_to_list(p) for p in poses], generator="tests/golden/make_golden.py",
                    reference="oracle/_ref libhashdag_ref (unmodified /root/reference/src, sm_100a, default -fmad)")
        np.savez_compressed(os.path.join(out_dir, f"ref_{name}.npz"), meta=json.dumps(meta), **arrays)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
