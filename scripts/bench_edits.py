"""BASELINE.json config 4 at full size: a replay_edits_add / replay_edits_remove-shaped sequence of sphere edits on the
depth-17 bench HashDAG, then frames over the edit-dirtied pages.  Prints one JSON line.

    python scripts/bench_edits.py [--edits 57] [--no-remove] [--footprint-log2 14]

The edits themselves are the REFERENCE's (SphereEditor through HashDAG::edit_threads on the host, oracle/_ref): this
measures what the tracer side owns -- getting each edit to the GPU and tracing over it -- beside the reference's own
upload (HashTable::upload_to_gpu) and kernels on the same edited DAG:
  * per edit: delta size vs the arrays the reference re-uploads, host time of diff + apply, first frame (3 passes),
    frames == reference kernels' frames (paths, colours, shaded) pixel for pixel;
  * afterwards: a fly-through over the fully edited DAG, product vs reference kernels.
Radii are those of replays/replay_edits_add.csv (57 EditSphere, radii 3-111 at depth 16) times two; centres are on
this scene's terrain ahead of the bench cameras (the replay's own centres belong to a scene that is not available).
Needs a GPU and the prebuilt oracle/_ref/libhashdag_ref_d17_1920x1080.so."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REPLAY_RADII = [10, 13, 17, 18, 18, 11, 9, 3, 3, 7, 11, 11, 11, 5, 5, 14, 14, 25, 33, 33, 50, 50, 50, 66, 66, 74, 74, 92, 111, 111, 103, 96, 100, 99, 101,
                96, 98, 103, 94, 99, 90, 90, 95, 104, 95, 84, 76, 68, 61, 52, 58, 37, 37, 16, 16, 30, 15]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edits", type=int, default=57)
    ap.add_argument("--no-remove", action="store_true")
    ap.add_argument("--levels", type=int, default=17)
    ap.add_argument("--footprint-log2", type=int, default=14)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--fly-frames", type=int, default=64)
    a = ap.parse_args()
    import torch
    from hashdag_b200 import camera, edits, tracer, workloads
    from oracle import ref

    W, H = a.width, a.height
    scene, poses = workloads.build_workload(a.levels, a.footprint_log2, 64)
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.RefTracer(a.levels, W, H)
    rt.load_scene(scene, extra_pool_pages=131072)
    pool, table, first, top = rt.hash_dag()
    nodes, offsets = rt.hash_colors()
    t = tracer.DAGTracer(True, W, H, a.levels)
    rep = edits.HashDagReplica(t, pool, table, top, first, a.levels, pool_capacity_pages=top + 131072, color_nodes=nodes, color_offsets=offsets,
                               main_leaf=tracer.CompressedColorLeaf.from_scene(scene), color_node_capacity=nodes.size + (1 << 20))
    # edit centres: on the ground ahead of the bench cameras, where those cameras look
    xz = workloads.flythrough_xz(a.levels, a.footprint_log2, 64)
    s = float(1 << a.footprint_log2)
    rng = np.random.default_rng(4)
    plan = []
    for k in range(a.edits):
        i = (k * 64) // max(a.edits, 1) % 64
        x0, z0 = xz[i]
        x1, z1 = xz[i + 1]
        d = float(np.hypot(x1 - x0, z1 - z0)) or 1.0
        ax, az = x0 + (x1 - x0) / d * 0.12 * s + float(rng.uniform(-60, 60)), z0 + (z1 - z0) / d * 0.12 * s + float(rng.uniform(-60, 60))
        h = scene.heights.get((int(x0 + (x1 - x0) / d * 0.12 * s), int(z0 + (z1 - z0) / d * 0.12 * s)), 1 << (a.levels - 1))
        r = 2.0 * REPLAY_RADII[k % len(REPLAY_RADII)]
        plan.append((i, (ax, float(h) + 0.4 * r, az), r))
    sequence = [(i, c, r, True) for i, c, r in plan] + ([] if a.no_remove else [(i, c, r, False) for i, c, r in plan])

    leaves_host, per_edit, bad_total = [], [], 0
    layout = edits.HashLayout(a.levels)
    vpool, vtable, vsizes = rt.hash_views()          # the reference's live host arrays + per-bucket fill counts (zero copy)
    last_sizes = vsizes.copy()
    del pool, table
    for k, (pi, centre, radius, adding) in enumerate(sequence):
        rt.edit_sphere(centre, radius, adding)
        ref_edit_ms, ref_upload_ms = rt.last_edit_ms()
        nfirst, ntop = rt.hash_info()
        nnodes, _ = rt.hash_colors()
        nleaves = [edits.ColorLeafArrays(*l) for l in rt.color_leaves()]
        h0 = time.perf_counter()
        delta = edits.delta_from_bucket_sizes(layout, last_sizes, vsizes, vpool, vtable, nfirst, ntop)
        last_sizes = vsizes.copy()
        hc = time.perf_counter()
        delta = edits.add_color_delta(delta, nodes, nnodes, leaves_host, nleaves)     # harness-side: the editor knows which leaves it rebuilt
        h1 = time.perf_counter()
        rep.apply(delta)
        ha = time.perf_counter()
        t.sync()
        h2 = time.perf_counter()
        first, top, nodes, leaves_host = nfirst, ntop, nnodes, nleaves
        pose = poses[pi]
        ms = t.resolve_frame(pose, info, rep.dag(), rep.colors(), 1.0, 0.0)
        p = t.read_paths()
        # per-pass frames for the comparison
        t.resolve_colors(rep.dag(), rep.colors()); col = t.read_colors()
        t.resolve_shadows(pose, info, rep.dag(), 1.0, 0.0); sh = t.read_colors()
        ra = rt.resolve_paths(1, pose, info); rp = rt.read_paths()
        rb = rt.resolve_colors(1, 3); rc = rt.read_colors()
        rcs = rt.resolve_shadows(1, pose, info, 1.0, 0.0); rs = rt.read_colors()
        bad = int((p != rp).any(-1).sum()) + int((col != rc).sum()) + int((sh != rs).sum())
        bad_total += bad
        per_edit.append({"edit": k, "adding": adding, "radius": radius, "delta_bytes": delta.nbytes,
                         "reference_upload_bytes": int(vtable.nbytes + sum(int(r["n_words"]) * 4 for r in delta.pool_ranges) + nnodes.nbytes),
                         "pool_spans": len(delta.pool_ranges), "table_spans": len(delta.table_ranges), "new_leaves": len(delta.color_leaves),
                         "tracker_host_ms": (hc - h0) * 1e3, "color_diff_harness_ms": (h1 - hc) * 1e3, "apply_enqueue_ms": (ha - h1) * 1e3, "apply_ms": (h2 - h1) * 1e3, "frame_ms": float(sum(ms)), "reference_frame_ms": float(ra + rb + rcs),
                         "reference_edit_ms": ref_edit_ms, "reference_upload_ms": ref_upload_ms, "mismatched_pixels": bad,
                         "pixels_changed_by_edit": None})
        sys.stderr.write(f"edit {k} r={radius} add={adding} delta={delta.nbytes} tracker={1e3 * (hc - h0):.2f}ms apply={1e3 * (h2 - h1):.2f}ms frame={sum(ms):.3f}ms ref_frame={ra + rb + rcs:.3f}ms bad={bad}\n")
    # fly-through over the edited DAG
    ours = refk = 0.0
    fly_bad = 0
    for i in range(a.fly_frames):
        pose = poses[i % len(poses)]
        ms = t.resolve_frame(pose, info, rep.dag(), rep.colors(), 1.0, 0.0)
        ours += float(sum(ms))
        img = t.read_colors()
        refk += rt.resolve_paths(1, pose, info) + rt.resolve_colors(1, 3) + rt.resolve_shadows(1, pose, info, 1.0, 0.0)
        fly_bad += int((img != rt.read_colors()).sum())
    t.close()
    rt.close()
    med = lambda key: float(np.median([e[key] for e in per_edit])) if per_edit else None
    out = {"config": "BASELINE.json configs[3]: sphere add/remove sequence (radii of replays/replay_edits_add.csv x2) on the depth-%d bench HashDAG, %dx%d" % (a.levels, W, H),
           "edits": len(sequence), "mismatched_pixels_total": bad_total, "flythrough_mismatched_pixels": fly_bad,
           "median": {k: med(k) for k in ("delta_bytes", "reference_upload_bytes", "tracker_host_ms", "apply_enqueue_ms", "apply_ms", "frame_ms", "reference_frame_ms", "reference_edit_ms", "reference_upload_ms")},
           "max": {k: float(max(e[k] for e in per_edit)) for k in ("delta_bytes", "apply_ms", "frame_ms", "new_leaves")} if per_edit else {},
           "flythrough": {"frames": a.fly_frames, "ms_per_frame": ours / max(a.fly_frames, 1), "reference_kernels_ms_per_frame": refk / max(a.fly_frames, 1)},
           "final": {"pool_pages": int(top), "unique_color_leaves": len(leaves_host), "color_nodes": int(nodes.size)},
           "per_edit": per_edit}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
