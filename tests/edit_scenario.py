"""Config 4 in small: the reference's own CPU edits (SphereEditor add / remove, oracle/_ref) dirty pages of a
HashDAG; the spans go to the tracer's replica as one packed delta per edit (hashdag_b200/edits.py ->
hdt_apply_ranges); after every edit the product's frames must equal the reference kernels' frames on the edited
DAG, and the oracle's on the edited host arrays.  Run as a script (own process: the reference keeps its scene in
globals and the edits would leak into other tests); tests/test_gpu_edits.py calls it."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import golden_util as gu                      # noqa: E402
import make_color_leaf_golden as mk           # noqa: E402
from hashdag_b200 import camera, edits, tracer   # noqa: E402
from oracle import hdo, ref                   # noqa: E402


def oracle_colors_with_leaves(scene, nodes, leaves):
    col = hdo.make_colors(scene, hdo.COLORS_HASH)
    col.color_nodes, col.n_color_nodes = nodes.ctypes.data, nodes.size
    arr = (hdo.ColorLeaf * max(1, len(leaves)))()
    for i, (w, b, m) in enumerate(leaves):
        arr[i] = hdo.ColorLeaf(w.ctypes.data if w.size else None, w.size, b.ctypes.data if b.size else None, b.size, m.ctypes.data if m.size else None, m.size)
    col.unique_leaves, col.n_unique_leaves = C.cast(arr, C.c_void_p), len(leaves)
    col._keep2 = (nodes, leaves, arr)
    return col


def main(recipe="d13"):
    scene = gu.recipe_scene(recipe)
    W = H = gu.W
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.RefTracer(scene.levels, W, H)
    rt.load_scene(scene)
    pool, table, first, top = rt.hash_dag()
    nodes, offsets = rt.hash_colors()
    assert np.array_equal(pool, scene.hash_pool) and first == scene.hash_first_node_index

    t = tracer.DAGTracer(True, W, H, scene.levels)
    rep = edits.HashDagReplica(t, pool, table, top, first, scene.levels, pool_capacity_pages=top + 4096, color_nodes=nodes, color_offsets=offsets,
                               main_leaf=tracer.CompressedColorLeaf.from_scene(scene), color_node_capacity=nodes.size + 4096)
    c = float(1 << (scene.levels - 1))
    h0 = float(scene.heights[(int(c), int(c))])
    poses = gu.recipe_poses(scene)[:2] + [camera.look_at((c - 60.0, h0 + 45.0, c - 50.0), (c + 10.0, h0, c + 5.0))]
    # shaped like replays/replay_edits_add.csv / _remove.csv: spheres of mixed radii on the surface, added then carved
    # (radii 3-25 at depth 13, 10-60 at depth 17)
    plan = mk.leaf_plan(scene)
    leaves_host = []
    report = []
    layout = edits.HashLayout(scene.levels)
    vpool, vtable, vsizes = rt.hash_views()          # the reference's live host arrays + per-bucket fill counts
    assert vsizes.size == layout.n_buckets and vtable.size == layout.n_pages
    last_sizes = vsizes.copy()
    tracker = tracer.DirtyTracker(scene.levels)      # the product's tracker (C++, hdt_tracker_*); the numpy twin below checks it
    assert tracker.n_buckets == layout.n_buckets
    tracker.snapshot(vsizes)
    for k, (centre, radius, adding) in enumerate(plan):
        rt.edit_sphere(centre, radius, adding)
        npool, ntable, nfirst, ntop = rt.hash_dag()
        nnodes, _ = rt.hash_colors()
        nleaves = [edits.ColorLeafArrays(*l) for l in rt.color_leaves()]
        full = edits.diff_hash_dag(pool, table, npool, ntable, nfirst, ntop, nodes, nnodes, leaves_host, nleaves)
        # the product path: spans from the hash table's own bookkeeping, no array comparison (edits.delta_from_bucket_sizes)
        delta = edits.add_color_delta(edits.delta_from_bucket_sizes(layout, last_sizes, vsizes, vpool, vtable, nfirst, ntop), nodes, nnodes, leaves_host, nleaves)
        last_sizes = vsizes.copy()
        pod = tracker.delta_pod(vsizes, vpool, vtable, nfirst, ntop)
        for a, b in zip(tracer.DirtyTracker.arrays(pod), (delta.pool_ranges, delta.pool_payload, delta.table_ranges, delta.table_payload)):
            assert np.array_equal(a, b), "hdt_tracker_delta differs from edits.delta_from_bucket_sizes"
        assert (pod.first_node_index, pod.pool_top) == (nfirst, ntop)
        assert delta.pool_payload.size <= 2 * full.pool_payload.size + 64 * max(1, len(delta.pool_ranges))
        # host mirror of the device apply: the delta alone reproduces the new arrays
        hp = np.zeros(max(pool.size, ntop * 512), np.uint32); hp[: pool.size] = pool
        edits.apply_spans_host(hp, delta.pool_ranges, delta.pool_payload)
        ht = table.copy(); edits.apply_spans_host(ht, delta.table_ranges, delta.table_payload)
        assert np.array_equal(hp[: ntop * 512], npool) and np.array_equal(ht, ntable)
        rep.apply(delta, pod=pod)                    # pool / page table straight from the tracker's arrays (hdt_broadcast_dirty)
        full = npool.nbytes + ntable.nbytes + nnodes.nbytes
        changed = int((nfirst != first) or len(delta.pool_ranges) > 0)
        pool, table, first, top, nodes, leaves_host = npool, ntable, nfirst, ntop, nnodes, nleaves

        # the oracle on the edited host arrays
        sc = type("S", (), {})()
        odag = hdo.Dag(hdo.DAG_HASH, scene.levels, pool.ctypes.data, pool.size, table.ctypes.data, table.size, first)
        ocol = oracle_colors_with_leaves(scene, nodes, [(l.weights, l.blocks, l.macro_blocks) for l in leaves_host])
        bad = 0
        for pose in poses:
            prm = camera.trace_params(pose, info, scene.levels, W, H)
            rt.resolve_paths(1, pose, info); rp = rt.read_paths()
            rt.resolve_colors(1, 3); rc = rt.read_colors()
            rt.resolve_shadows(1, pose, info, 1.0, 0.0); rs = rt.read_colors()
            t.resolve_paths(pose, info, rep.dag()); p = t.read_paths()
            t.resolve_colors(rep.dag(), rep.colors()); col = t.read_colors()
            t.resolve_shadows(pose, info, rep.dag(), 1.0, 0.0); s = t.read_colors()
            op, _ = hdo.trace_paths(odag, W, H, prm)
            oc, _ = hdo.trace_colors(odag, ocol, op)
            osh, _ = hdo.trace_shadows(odag, prm, op, oc, 1.0, 0.0)
            bad += int((p != rp).any(-1).sum()) + int((col != rc).sum()) + int((s != rs).sum())
            bad += int((op != rp).any(-1).sum()) + int((oc != rc).sum()) + int((osh != rs).sum())
        report.append({"edit": k, "adding": adding, "radius": radius, "changed": changed, "delta_bytes": delta.nbytes, "full_upload_bytes": full,
                       "pool_spans": len(delta.pool_ranges), "table_spans": len(delta.table_ranges), "new_leaves": len(delta.color_leaves),
                       "unique_leaves": len(leaves_host), "mismatched_pixels": bad})
    t.close()
    rt.close()
    print("EDIT_SCENARIO " + json.dumps(report))
    return report


if __name__ == "__main__":
    rep = main(sys.argv[1] if len(sys.argv) > 1 else "d13")
    sys.exit(0 if all(r["mismatched_pixels"] == 0 for r in rep) and any(r["changed"] for r in rep) else 1)
