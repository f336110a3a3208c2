"""BASELINE.json config 4 on N GPUs (one process per GPU, `torchrun --nproc-per-node N`; N = 1 works too): sphere edits by the
REFERENCE's own editor on rank 0 (oracle/_ref), each edit's dirty spans found by the C ABI's tracker (hdt_tracker_*),
broadcast and applied on every rank's replica (hdt_broadcast_dirty / hdt_broadcast_ranges / hdt_replicate: NCCL inside
libhashdag_b200.so), then one 4K frame rendered by all ranks (64x64 tiles, tile t -> rank t % N) and assembled in pinned
host memory shared by the ranks (hdt_exchange_attach_host: every rank pushes its tiles over its own PCIe link).
Rank 0 compares every assembled frame with the reference kernels' frame of the edited DAG and prints one JSON line.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_edits_multi.py [--edits 16]

torch.distributed (gloo) carries only the rendezvous metadata (NCCL id, array sizes, the shared-memory name)."""
import argparse
import json
import os
import sys
import time
from multiprocessing import shared_memory

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.bench_edits import REPLAY_RADII  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edits", type=int, default=16)
    ap.add_argument("--levels", type=int, default=17)
    ap.add_argument("--footprint-log2", type=int, default=14)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from hashdag_b200 import camera, edits, tracer, workloads

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)

    def share(obj):
        if world == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    W, H = a.width, a.height
    rt = scene = None
    if rank == 0:
        from oracle import ref
        scene, poses = workloads.build_workload(a.levels, a.footprint_log2, 64)
        rt = ref.RefTracer(a.levels, W, H, device=local)
        rt.load_scene(scene, extra_pool_pages=131072)
        pool, table, first, top = rt.hash_dag()
        nodes, offsets = rt.hash_colors()
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
        sequence = [(i, c, r, True) for i, c, r in plan] + [(i, c, r, False) for i, c, r in plan]
        meta = {"sizes": {"pool": int(pool.size), "table": int(table.size), "nodes": int(nodes.size), "offsets": int(offsets.size),
                          "weights": int(scene.weights.size), "blocks": int(scene.blocks.size), "macro": int(scene.macro_blocks.size)},
                "top": int(top), "first": int(first), "n_edits": len(sequence),
                "poses": [[list(poses[i].position), [list(r) for r in poses[i].rotation]] for i, _, _, _ in sequence],
                "bounds": [list(scene.bounds_min), list(scene.bounds_max)],
                "nccl_id": tracer.comm_unique_id() if world > 1 else b""}
    else:
        meta = None
    meta = share(meta)
    info = camera.DAGInfo(tuple(meta["bounds"][0]), tuple(meta["bounds"][1]))
    edit_poses = [camera.CameraView(tuple(p[0]), tuple(tuple(r) for r in p[1])) for p in meta["poses"]]
    sz, top, first = meta["sizes"], meta["top"], meta["first"]
    if rank != 0:
        pool, table = np.zeros(sz["pool"], np.uint32), np.zeros(sz["table"], np.uint32)
        nodes, offsets = np.zeros(sz["nodes"], np.uint32), np.zeros(sz["offsets"], np.uint64)

    t = tracer.DAGTracer(True, W, H, a.levels, device=local)
    if world > 1:
        t.comm_init(meta["nccl_id"], rank, world)
        t.set_partition(rank, world, 6)
    if rank == 0:
        main_leaf = tracer.CompressedColorLeaf.from_scene(scene, dev)
    else:
        main_leaf = tracer.CompressedColorLeaf(torch.zeros(sz["weights"], dtype=torch.int32, device=dev), torch.zeros(sz["blocks"], dtype=torch.int64, device=dev),
                                               torch.zeros(sz["macro"], dtype=torch.int64, device=dev), tracer.UNIQUE_OFFSET)
    t0 = time.perf_counter()
    rep = edits.HashDagReplica(t, pool, table, top, first, a.levels, pool_capacity_pages=top + 131072, color_nodes=nodes, color_offsets=offsets,
                               main_leaf=main_leaf, color_node_capacity=sz["nodes"] + (1 << 20), device=dev, replicate_from=0)
    t.sync()
    replicate_s = time.perf_counter() - t0

    # the frame, assembled in host memory shared by all ranks
    shm = frame_host = None
    if world > 1:
        nbytes = t.exchange_block_bytes()
        name = None
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name = shm.name
        name = share(name)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name)
            try:   # only the creator unlinks
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        block = np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes)
        t.exchange_attach_host(block.ctypes.data, nbytes)
        from hashdag_b200 import partition
        mt, T = partition.max_tiles_per_rank(world, W, H, 6), 64
        frame_host = block[: world * mt * T * T * 4].view(np.uint32).reshape(world, mt, T, T)     # tile layout: rank-major compact tile buffers
        dist.barrier()

    per_edit, bad_total = [], 0
    if rank == 0:
        vpool, vtable, vsizes = rt.hash_views()
        trk = tracer.DirtyTracker(a.levels)
        trk.snapshot(vsizes)
        leaves_host = []
        del pool, table
    for k in range(meta["n_edits"]):
        pose = edit_poses[k]
        prm = camera.trace_params(pose, info, a.levels, W, H)
        delta = pod = None
        tracker_ms = 0.0
        if rank == 0:
            _, centre, radius, adding = sequence[k]
            rt.edit_sphere(centre, radius, adding)
            nfirst, ntop = rt.hash_info()
            nnodes, _ = rt.hash_colors()
            nleaves = [edits.ColorLeafArrays(*l) for l in rt.color_leaves()]
            h0 = time.perf_counter()
            pod = trk.delta_pod(vsizes, vpool, vtable, nfirst, ntop)
            tracker_ms = (time.perf_counter() - h0) * 1e3
            e = np.zeros(0, dtype=edits.RANGE_DTYPE)
            delta = edits.add_color_delta(edits.DagDelta(nfirst, ntop, e, np.zeros(0, np.uint32), e.copy(), np.zeros(0, np.uint32)), nodes, nnodes, leaves_host, nleaves)
            nodes, leaves_host = nnodes, nleaves
        if world > 1:
            dist.barrier()
        h1 = time.perf_counter()
        rep.apply(delta, root=0, pod=pod)
        t.sync()
        h2 = time.perf_counter()
        # first frame after the edit: every rank renders its tiles, the frame assembles in shared host memory
        dag, col = rep.dag(), rep.colors()
        t.enqueue_frame(prm, dag.pod(), dag.kind, col.pod(), col.kind, 1.0, 0.0, True, None)
        if world > 1:
            t.exchange_frame()
        t.sync()
        h3 = time.perf_counter()
        if rank == 0:
            img = partition.assemble([frame_host[r] for r in range(world)], W, H, 6) if world > 1 else t.read_colors()
            if world > 1:
                t.exchange_release()
            ra = rt.resolve_paths(1, pose, info)
            rb = rt.resolve_colors(1, 3)
            rc = rt.resolve_shadows(1, pose, info, 1.0, 0.0)
            bad = int((img != rt.read_colors()).sum())
            bad_total += bad
            per_edit.append({"edit": k, "adding": bool(sequence[k][3]), "radius": sequence[k][2], "pool_spans": int(pod.n_pool_ranges), "delta_words": int(pod.n_pool_payload + pod.n_table_payload),
                             "new_leaves": len(delta.color_leaves), "tracker_host_ms": tracker_ms, "broadcast_apply_ms": (h2 - h1) * 1e3, "first_frame_ms": (h3 - h2) * 1e3,
                             "reference_frame_ms": float(ra + rb + rc), "reference_upload_ms": rt.last_edit_ms()[1], "mismatched_pixels": bad})
            sys.stderr.write(f"edit {k}: tracker {tracker_ms:.2f} ms, broadcast+apply {1e3 * (h2 - h1):.2f} ms, first frame {1e3 * (h3 - h2):.2f} ms (reference kernels {ra + rb + rc:.2f} ms), bad {bad}\n")
    t.sync()
    if world > 1:
        dist.barrier()
    if rank == 0:
        med = lambda key: float(np.median([e[key] for e in per_edit])) if per_edit else None
        print(json.dumps({"config": f"BASELINE.json configs[3] on {world} GPU(s): reference SphereEditor edits on rank 0 -> hdt_tracker_delta -> hdt_broadcast_dirty (NCCL inside the "
                                    f"library) -> replicas; first frame at {W}x{H} by tiles, assembled in shared pinned host memory (hdt_exchange_attach_host)",
                          "n_gpus": world, "edits": len(per_edit), "mismatched_pixels_total": bad_total, "replicate_s": replicate_s,
                          "median": {k: med(k) for k in ("tracker_host_ms", "broadcast_apply_ms", "first_frame_ms", "reference_frame_ms", "reference_upload_ms", "delta_words")},
                          "max": {k: float(max(e[k] for e in per_edit)) for k in ("broadcast_apply_ms", "first_frame_ms", "new_leaves")} if per_edit else {},
                          "per_edit": per_edit}))
    t.close()
    if rt is not None:
        rt.close()
    if shm is not None:
        del frame_host, block
        try:
            shm.close()
        except BufferError:
            pass
        if rank == 0:
            shm.unlink()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
