"""Generate tests/golden/ref_color_leaves_d{13,17}.npz: colour leaves written by the reference's own
ColorLeafBuilder::build (variable_weight_size_colors.h:643-685) while its SphereEditor edits a HashDAG
(hash_dag_edits.h:340-540).  Run on a GPU box (the reference's build() uploads every leaf to the GPU):

    gpurun -- 'python tests/golden/make_color_leaf_golden.py gpurun_out/golden'

then copy gpurun_out/golden/ref_color_leaves_d*.npz into tests/golden/.  Needs oracle/_ref.  At depth 13 a colour leaf
covers 8^3 voxels (colour tree depth 10, hash_dag_globals.h:7): many tiny leaves; at depth 17 it covers 128^3: leaves
of 10^4-10^5 colours that span several macro blocks and carry every weight width.
A size-balanced subset of the leaves is stored (smallest, largest, and evenly spaced in between) plus the
SHA-256 of every leaf of the scenario, which the GPU test recomputes from the live reference.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util as gu  # noqa: E402
from oracle import ref  # noqa: E402

BUDGET_BYTES = 1_500_000


def leaf_plan(scene):
    """The six-edit scenario of tests/edit_scenario.py (adds and carves, radii 3-25); at depth 17 the same shape with
    radii 10-60 (replays/replay_edits_add.csv uses radii 3-111)."""
    c = float(1 << (scene.levels - 1))
    h0 = float(scene.heights[(int(c), int(c))])
    if scene.levels >= 17:
        return [((c + 41.0, h0 + 12.0, c + 21.0), 40.0, True), ((c - 30.0, h0 - 4.0, c + 10.0), 30.0, False), ((c + 6.0, h0 + 28.0, c - 16.0), 10.0, True),
                ((c + 36.0, h0 + 18.0, c + 24.0), 24.0, False), ((c - 60.0, h0 + 8.0, c - 44.0), 60.0, True), ((c - 56.0, h0 + 20.0, c - 40.0), 35.0, False)]
    return [((c + 20.5, h0 + 6.0, c + 10.5), 12.0, True), ((c - 15.0, h0 - 2.0, c + 5.0), 9.0, False), ((c + 3.0, h0 + 14.0, c - 8.0), 3.0, True),
            ((c + 18.0, h0 + 9.0, c + 12.0), 7.0, False), ((c - 30.0, h0 + 4.0, c - 22.0), 25.0, True), ((c - 28.0, h0 + 10.0, c - 20.0), 14.0, False)]


def leaf_digest(w, b, m):
    h = hashlib.sha256()
    for a in (w, b, m):
        h.update(np.ascontiguousarray(a).tobytes())
        h.update(b"|")
    return h.hexdigest()


def reference_leaves(recipe="d13"):
    scene = gu.recipe_scene(recipe)
    rt = ref.RefTracer(scene.levels, gu.W, gu.H)
    rt.load_scene(scene)
    for centre, radius, adding in leaf_plan(scene):
        rt.edit_sphere(centre, radius, adding)
    leaves = [l for l in rt.color_leaves() if l[1].size]
    rt.close()
    return leaves


def main(out_dir, recipe):
    os.makedirs(out_dir, exist_ok=True)
    leaves = reference_leaves(recipe)
    order = sorted(range(len(leaves)), key=lambda i: sum(a.nbytes for a in leaves[i]))
    picked, used = [], 0
    cand = [order[0], order[-1]] + [order[i] for i in np.linspace(0, len(order) - 1, 40).astype(int)]
    for i in cand:
        nb = sum(a.nbytes for a in leaves[i])
        if i not in picked and used + nb <= BUDGET_BYTES:
            picked.append(i)
            used += nb
    arrays = {}
    for k, i in enumerate(picked):
        arrays[f"weights_{k}"], arrays[f"blocks_{k}"], arrays[f"macro_{k}"] = leaves[i]
    meta = dict(recipe=recipe, n_leaves_in_scenario=len(leaves), stored=len(picked), stored_indices=[int(i) for i in picked],
                digests=[leaf_digest(*l) for l in leaves], generator="tests/golden/make_color_leaf_golden.py",
                reference=f"oracle/_ref libhashdag_ref_{recipe}_256x256 (unmodified /root/reference/src): SphereEditor edits, ColorLeafBuilder::build")
    np.savez_compressed(os.path.join(out_dir, f"ref_color_leaves_{recipe}.npz"), meta=json.dumps(meta), **arrays)
    bpw_hist = np.zeros(5, np.int64)
    for w, b, m in leaves:
        h = (b & np.uint64(0xFFFFFFFF)).astype(np.int64)
        bpw = np.where((h >> 16) == 0xFFFF, 0, ((h >> 14) & 3) + 1)
        bpw_hist += np.bincount(bpw, minlength=5)
    print(f"{recipe}: {len(leaves)} leaves, {sum(l[1].size for l in leaves)} blocks, stored {len(picked)} ({used} bytes), blocks by bits-per-weight {bpw_hist.tolist()}, macro blocks max {max(l[2].size // 2 for l in leaves)}")


if __name__ == "__main__":
    # one process per recipe: the reference keeps its scene in globals
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"), sys.argv[2] if len(sys.argv) > 2 else "d13")
