#!/usr/bin/env python
"""Benchmark of the DAGTracer hot path (BASELINE.json metric: Mrays/s primary+shadow on a depth-17
HashDAG at 1080p / 4K, with the achieved fraction of the HBM gather roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference ...                            # the algorithm on the host cores (CPU arm)
    python bench.py --impl reference-cuda ...                       # the reference's own CUDA kernels, same frames (1 GPU)

A step is one frame of the seeded fly-through: resolve_paths + resolve_colors + resolve_shadows.
rays per step = W*H primary + one shadow ray per hit pixel.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mrays/s primary+shadow (depth-17 HashDAG)"
UNIT = "Mrays/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--levels", type=int, default=17)
    ap.add_argument("--footprint-log2", type=int, default=int(os.environ.get("HDT_BENCH_FOOTPRINT", "14")))
    ap.add_argument("--poses", type=int, default=64)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-sample-poses", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference's own CUDA kernels (oracle/_ref) run after the timed region at N = 1")
    ap.add_argument("--dag", default="hash", choices=["hash", "basic"])
    ap.add_argument("--no-beam-prefetch", action="store_true", help="keep the beam kernels of frame n+1 behind all of frame n")
    ap.add_argument("--no-resolved", action="store_true", help="trace the HashDAG through its page table (A/B); default: the resolved pool (hdt_hash_dag_resolve)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-fused", "nccl"],
                    help="N > 1: how the tiles reach rank 0's frame -- stored over peer memory by a scatter kernel (default), by the shadows kernel "
                         "itself (peer-fused, HDT_OPT_EXCHANGE_FUSED), or NCCL gather + assembly (A/B)")
    ap.add_argument("--e2e-exchange", default="host", choices=["host", "rank0"],
                    help="N > 1, end-to-end loop: how the frame reaches host memory -- every rank stores its tiles into pinned host memory shared by "
                         "the ranks over its own PCIe link (hdt_exchange_attach_host, default), or rank 0 copies the frame assembled in its device memory (A/B)")
    ap.add_argument("--frames-in-flight", type=int, default=3, choices=[1, 2, 3, 4],
                    help="tracer contexts (each with its own streams and frame buffers) the fly-through alternates between")
    return ap.parse_args()


def resolution(args, world):
    if args.width and args.height:
        return args.width, args.height
    return (1920, 1080) if world == 1 else (3840, 2160)   # configs[1] / configs[2]


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"], out["samples"] = sorted(reasons), len(sm)
        return out


def algorithmic_bytes(stats_paths, stats_colors, stats_shadows, hash_dag, w, h):
    """bytes the reference algorithm touches per frame, per pass."""
    def walk(st):
        return 4 * st["n_word"] + 8 * st["n_leaf"] + (4 * st["n_page"] if hash_dag else 0)
    px = w * h
    return {
        "paths": walk(stats_paths) + 16 * px,
        "colors": walk(stats_colors) + 8 * stats_colors["n_color_probe"] + 16 * px + 4 * px,
        "shadows": walk(stats_shadows) + (16 + 4 + 4) * px,
    }


def oracle_frame(hdo, odag, ocol, params, w, h, threads=0):
    t0 = time.perf_counter()
    p, sp = hdo.trace_paths(odag, w, h, params, threads)
    c, sc = hdo.trace_colors(odag, ocol, p, n_threads=threads)
    s, ss = hdo.trace_shadows(odag, params, p, c, 1.0, 0.0, threads)
    dt = time.perf_counter() - t0
    return dt, w * h + sp["n_hit"], (sp, sc, ss), s


def run_reference_cpu(args):
    """--impl reference: the algorithm restated for the host (oracle/, kind "port"; the reference has no
    host implementation of its __device__ traversal) on every host thread.  Each step is a
    quarter-resolution frame of the same fly-through so the run stays within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hashdag_b200 import camera, workloads
    from oracle import hdo
    W, H = resolution(args, args.gpus)
    w, h = W // 4, H // 4
    scene, poses = workloads.build_workload(args.levels, args.footprint_log2, args.poses)
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    hashed = args.dag == "hash"
    odag = hdo.make_dag(scene, hdo.DAG_HASH if hashed else hdo.DAG_BASIC)
    ocol = hdo.make_colors(scene, hdo.COLORS_HASH if hashed else hdo.COLORS_COMPRESSED)
    cores = os.cpu_count() or 1
    rays, secs = 0, 0.0
    for i in range(args.warmup + args.steps):
        prm = camera.trace_params(poses[i % len(poses)], info, scene.levels, w, h)
        dt, n, _, _ = oracle_frame(hdo, odag, ocol, prm, w, h)
        if i >= args.warmup:
            rays += n
            secs += dt
    value = rays / secs / 1e6
    sample = f"{args.steps} frames at {w}x{h} (1/16 of the {W}x{H} rays per frame), same DAG and camera path, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32+f64+u32", "data": "synthetic",
        "config": workload_config(args, scene, W, H, args.gpus),
        "implementation": {"what": "oracle/hdo_oracle.cpp: the reference's traversal restated for host cores (tracer.cu:7-697), rows of the image over threads",
                           "threads": cores, "frame": [w, h], "rays_per_frame_fraction": (w * h) / (W * H)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, scene, W, H, world):
    return {
        "workload": f"synthetic depth-{args.levels} {'HashDAG' if args.dag == 'hash' else 'BasicDAG'} terrain (footprint 2^{args.footprint_log2}, seed 1337), "
                    f"{W}x{H} primary+shadow+colour, {args.poses}-pose fly-through",
        "levels": args.levels, "resolution": [W, H], "voxels": int(scene.n_voxels), "dag_words": int(scene.basic.size),
        "hash_pool_mib": round(scene.hash_pool.nbytes / 2**20, 1) if scene.has_hash else 0,
        "color_mib": round((scene.weights.nbytes + scene.blocks.nbytes + scene.macro_blocks.nbytes) / 2**20, 1),
        "partition": "whole frame" if world == 1 else f"{1 << int(os.environ.get('HDT_BENCH_TILE_LOG2', '6'))}x{1 << int(os.environ.get('HDT_BENCH_TILE_LOG2', '6'))} screen tiles, tile t -> rank t % {world}, replicated DAG, " + ("NCCL gather to rank 0" if getattr(args, "exchange", "peer") == "nccl" else "tiles stored into rank 0's frame over NVLink peer memory"),
        "l2_policy": "inputs larger than L2: each step is a different camera pose over a DAG pool >> 126 MB",
        "shadow_bias": 1.0, "fog_density": 0.0,
        "scene_generator": "hashdag_b200/scene/scene_builder.cpp: seeded value-noise terrain + spheres (NOT the reference's FastNoise, src/FastNoise.cpp; "
                           "no scene file of the reference is available offline)",
    }


def product_implementation(args, world):
    """How OUR arm renders the workload (the reference arms describe themselves in their own `implementation`)."""
    hashed = getattr(args, "dag", "hash") == "hash"
    resolved = hashed and not getattr(args, "no_resolved", False)
    return {
        "beam_prefetch": not getattr(args, "no_beam_prefetch", False),
        "hash_dag_pointers": "resolved pool + prefix pool (hdt_hash_dag_resolve: child pointers pre-translated, sibling voxel counts pre-summed)" if resolved else "as stored",
        "colors": "from the ancestor records of the paths pass (HDT_OPT_COLORS_RECORDED)" if resolved and os.environ.get("HDT_COLORS_RECORDED", "1") != "0" else "full DAG walk",
        "frames_in_flight": max(2, int(os.environ.get("HDT_BENCH_LANES", "2"))) if world > 1 else getattr(args, "frames_in_flight", 1),
    }


def time_reference_kernels(rt, poses, info, step_ids, warmup_ids, dk, ck):
    """The reference's three synchronous calls per frame (dag_tracer.cu:116-219), timed by its own cudaEvents (:130-138)."""
    for i in warmup_ids:
        rt.resolve_paths(dk, poses[i], info); rt.resolve_colors(dk, ck); rt.resolve_shadows(dk, poses[i], info, 1.0, 0.0)
    tp = tc = ts = 0.0
    for i in step_ids:
        tp += rt.resolve_paths(dk, poses[i], info)
        tc += rt.resolve_colors(dk, ck)
        ts += rt.resolve_shadows(dk, poses[i], info, 1.0, 0.0)
    n = max(1, len(step_ids))
    return {"paths": tp / n, "colors": tc / n, "shadows": ts / n}


def run_reference_cuda(args):
    """Extra arm (not part of the driver contract): the UNMODIFIED reference kernels from oracle/_ref
    on the same GPU, DAG and camera path -- the number the north star's '>= 5x' refers to."""
    from hashdag_b200 import camera, workloads
    from oracle import ref
    W, H = resolution(args, 1)
    scene, poses = workloads.build_workload(args.levels, args.footprint_log2, args.poses)
    info = camera.DAGInfo(scene.bounds_min, scene.bounds_max)
    rt = ref.RefTracer(args.levels, W, H)
    rt.load_scene(scene)
    hashed = args.dag == "hash"
    dk, ck = (1, 3) if hashed else (0, 1)
    hits = []
    for pose in poses:   # untimed: shadow rays per pose
        rt.resolve_paths(dk, pose, info)
        hits.append(int(rt.read_paths()[..., :3].any(-1).sum()))
    tp = tc = ts = 0.0
    rays = 0
    for i in range(args.warmup + args.steps):
        pose = poses[i % len(poses)]
        a = rt.resolve_paths(dk, pose, info)
        b = rt.resolve_colors(dk, ck)
        c = rt.resolve_shadows(dk, pose, info, 1.0, 0.0)
        if i >= args.warmup:
            tp, tc, ts = tp + a, tc + b, ts + c
            rays += W * H + hits[i % len(poses)]
    total_ms = tp + tc + ts
    rt.close()
    print(json.dumps({
        "impl": "reference-cuda", "metric": METRIC, "value": rays / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "dtype": "f32+f64+u32", "data": "synthetic",
        "config": workload_config(args, scene, W, H, 1),
        "implementation": {"what": "oracle/_ref: the unmodified reference kernels (tracer.cu, dag_tracer.cu) compiled for sm_100a, three synchronous calls per frame"},
        "passes_ms": {"paths": tp / args.steps, "colors": tc / args.steps, "shadows": ts / args.steps},
        "mrays_s_paths_shadows_only": rays / ((tp + ts) * 1e-3) / 1e6,
        "note": "kernel times from the reference's own cudaEvents (dag_tracer.cu:130-138), summed; each call is synchronous",
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hashdag_b200 import camera, tracer, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the tracer has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    W, H = resolution(args, world)
    hashed = args.dag == "hash"

    # ---- scene: rank 0 builds, the replicas get it over NCCL -------------------------------
    t_build = time.perf_counter()
    scene = poses = None
    if rank == 0:
        scene, poses = workloads.build_workload(args.levels, args.footprint_log2, args.poses)
    if world > 1:
        box = [scene_meta(scene, poses) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        meta = box[0]
        tensors = {}
        for name, (n, itemsize) in meta["arrays"].items():
            dt = torch.int32 if itemsize == 4 else torch.int64
            if rank == 0:
                arr = getattr(scene, name)
                t = torch.from_numpy(arr.view(np.int32 if itemsize == 4 else np.int64)).to(dev)
            else:
                t = torch.empty(n, dtype=dt, device=dev)
            dist.broadcast(t, src=0)
            tensors[name] = t
        if rank != 0:
            poses = [camera.CameraView(tuple(p[0]), tuple(tuple(r) for r in p[1])) for p in meta["poses"]]
        dag, colors = replicas_from_tensors(tracer, tensors, meta, hashed)
        bounds = (tuple(meta["bounds_min"]), tuple(meta["bounds_max"]))
        replication_bytes = sum(n * isz for n, isz in meta["arrays"].values())
    else:
        if hashed:
            dag, colors = tracer.HashDAG.from_scene(scene, dev), tracer.HashDAGColors.from_scene(scene, dev)
        else:
            dag, colors = tracer.BasicDAG.from_scene(scene, dev), tracer.BasicDAGCompressedColors.from_scene(scene, dev)
        bounds = (scene.bounds_min, scene.bounds_max)
        replication_bytes = 0
    t_build = time.perf_counter() - t_build
    info = camera.DAGInfo(*bounds)

    tr = tracer.DAGTracer(True, W, H, args.levels, device=local_rank)
    tile_log2 = int(os.environ.get("HDT_BENCH_TILE_LOG2", "6"))     # partition tiles of 64x64 pixels
    if world > 1:
        tr.set_partition(rank, world, tile_log2)
    params = [camera.trace_params(p, info, args.levels, W, H) for p in poses]
    if hashed and not args.no_resolved:
        # the HashDAG's child pointers pushed through the page table once, into a second pool (csrc/hdt_resolve.cuh)
        torch.cuda.synchronize()
        dag = tr.resolve_hash_dag(dag)
        tr.sync()
    dag_pod, col_pod = dag.pod(), colors.pod()

    # ---- multi-GPU gather plumbing ----------------------------------------------------------
    # Two frames are in flight per rank: frame i's NCCL gather + assembly (queued behind its kernels
    # on stream i%2) overlaps frame i+1's kernels on the other stream.  No host synchronisation
    # inside a step; collectives are issued in the same order on every rank.
    gather = None
    lanes = [tr]
    if world > 1:
        class _Mem:
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        n_lanes = max(2, int(os.environ.get("HDT_BENCH_LANES", "2")))     # frames in flight per rank
        lanes = [tr]
        for _ in range(n_lanes - 1):
            t2 = tracer.DAGTracer(True, W, H, args.levels, device=local_rank)
            t2.set_partition(rank, world, tile_log2)
            lanes.append(t2)
        streams = [torch.cuda.Stream(device=dev) for _ in lanes]
        mine, gathered, frames = [], [], []
        for t_, s_ in zip(lanes, streams):
            _, cptr, n_owned, max_tiles = t_.partition_buffers()
            n = max_tiles << (2 * tile_log2)
            mine.append(torch.as_tensor(_Mem(cptr, n), device=dev))
            gathered.append(torch.empty(world * n, dtype=torch.int32, device=dev) if rank == 0 else None)
            frames.append(torch.empty(W * H, dtype=torch.int32, device=dev) if rank == 0 else None)
            t_.set_stream(s_.cuda_stream)
        frame = frames[0] if rank == 0 else None

        peer = args.exchange in ("peer", "peer-fused")
        if peer:
            # rank 0's frames are mapped into every rank (CUDA IPC); ranks store their tiles into them (csrc/hdt_exchange.cuh)
            frames = []
            for t_ in lanes:
                box = [None]
                if rank == 0:
                    handle, fptr = t_.exchange_create()
                    box = [handle]
                    frames.append(torch.as_tensor(_Mem(fptr, W * H), device=dev))
                else:
                    frames.append(None)
                dist.broadcast_object_list(box, src=0)
                if rank != 0:
                    t_.exchange_open(box[0])
            frame = frames[0] if rank == 0 else None
            dist.barrier()

        def gather(k, release=True):
            if os.environ.get("HDT_BENCH_NO_GATHER"):      # diagnostics only: how much of a step is the exchange
                return
            if peer:
                lanes[k].exchange_frame()
                if rank == 0 and release:
                    lanes[k].exchange_release()
                return
            with torch.cuda.stream(streams[k]):
                dist.gather(mine[k], list(gathered[k].chunk(world)) if rank == 0 else None, dst=0)
                if rank == 0:
                    lanes[k].assemble_colors(gathered[k], frames[k])
    elif args.frames_in_flight >= 2:
        # single GPU: frame i+1 (second context, own streams and buffers) is enqueued while frame i runs, so the
        # drain of one kernel overlaps the ramp-up of another instead of leaving SMs idle
        lanes = [tr] + [tracer.DAGTracer(True, W, H, args.levels, device=local_rank) for _ in range(args.frames_in_flight - 1)]
    host_frame = torch.empty(W * H, dtype=torch.int32).pin_memory() if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        k = i % len(lanes)
        lanes[k].enqueue_frame(params[i % len(params)], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
        if gather:
            gather(k)

    def sync_all():
        for t_ in lanes:
            t_.sync()

    # ---- hits per pose (untimed) -------------------------------------------------------------
    hits = []
    for i in range(len(params)):
        tr.enqueue_frame(params[i], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, False, None)
        tr.sync()
        hits.append(tr.count_hits())
    if world > 1:
        ht = torch.tensor(hits, dtype=torch.int64, device=dev)
        dist.all_reduce(ht)
        hits = ht.tolist()

    # The scene is static during the timed frames, so the ray setup + beam kernels of frame n+1 may run
    # beside the colours / shadows kernels of frame n (include/hashdag_b200.h, HDT_OPT_BEAM_PREFETCH).
    for t_ in lanes:
        t_.set_option(tracer.OPT_BEAM_PREFETCH, 0 if args.no_beam_prefetch else 1)

    def arm_fused(on):
        if world > 1 and args.exchange == "peer-fused" and not os.environ.get("HDT_BENCH_NO_GATHER"):
            for t_ in lanes:
                t_.set_option(tracer.OPT_EXCHANGE_FUSED, 1 if on else 0)
    arm_fused(True)     # from here on every frame with shadows is followed by its exchange

    # ---- warm-up ------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    sync_all()

    # ---- timed: K steps, inputs resident, device time, max over ranks ----------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = sum(t_.launch_count() for t_ in lanes)
    barrier()
    t0 = time.perf_counter()
    if gather:
        start = torch.cuda.Event(enable_timing=True)
        start.record()                       # default stream, idle after the barrier
        for s_ in streams:
            s_.wait_event(start)
        for i in range(args.steps):
            step_device(args.warmup + i)
        ends = []
        for s_ in streams:
            e_ = torch.cuda.Event(enable_timing=True)
            e_.record(s_)
            ends.append(e_)
        torch.cuda.synchronize()
        dev_ms = max(start.elapsed_time(e_) for e_ in ends)
    else:
        # one event before the first frame on every lane's stream, one after the last: device time of the whole batch
        for t_ in lanes:
            t_.timer_begin()
        for i in range(args.steps):
            step_device(args.warmup + i)
        dev_ms = max([t_.timer_end() for t_ in lanes])
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    # device events on the streams everything is queued on (kernels, NCCL gather, assembly); the
    # barrier-to-barrier wall clock is reported beside it
    elapsed_ms = dev_ms
    launches = sum(t_.launch_count() for t_ in lanes) - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    rays = sum(W * H + hits[(args.warmup + i) % len(hits)] for i in range(args.steps))
    value = rays / (elapsed_ms * 1e-3) / 1e6

    # ---- e2e: public API, host camera in, colour frame out to pinned host memory, every step --
    # Every step: camera pose -> get_trace_params on the host -> 96 bytes of kernel arguments -> three passes
    # -> the colour frame copied into pinned host memory.  Single GPU: the steps alternate between the
    # contexts (hdt_resolve_frame_async), so frame i's copy over PCIe overlaps frame i+1's kernels; a
    # context is synchronised (its host frame is complete and may be consumed) before it is given the next frame.
    for t_ in lanes:
        t_.set_option(tracer.OPT_BEAM_PREFETCH, 0)
    host_frames = [host_frame] + [torch.empty(W * H, dtype=torch.int32).pin_memory() for _ in lanes[1:]] if rank == 0 else None
    host_lanes, host_shm, host_views = [], [], []
    if world > 1 and args.e2e_exchange == "host" and not os.environ.get("HDT_BENCH_NO_GATHER"):
        # N > 1: three more contexts per rank whose exchange block is pinned HOST memory shared by all ranks (POSIX shared memory):
        # every rank pushes its own tiles over its own PCIe link, nothing funnels through rank 0's GPU
        from multiprocessing import shared_memory
        for _ in range(3):
            t_ = tracer.DAGTracer(True, W, H, args.levels, device=local_rank)
            t_.set_partition(rank, world, tile_log2)
            nbytes = t_.exchange_block_bytes()
            box = [None]
            if rank == 0:
                shm = shared_memory.SharedMemory(create=True, size=nbytes)      # zero-filled
                box = [shm.name]
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                shm = shared_memory.SharedMemory(name=box[0])
                try:   # only the creator unlinks: keep Python's resource tracker from trying (and warning) in the other ranks
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(shm._name, "shared_memory")
                except Exception:
                    pass
            view = np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes)
            t_.exchange_attach_host(view.ctypes.data, nbytes)
            host_lanes.append(t_); host_shm.append(shm); host_views.append(view)
        del view
        dist.barrier()
        for i in range(len(host_lanes)):          # warm-up: first touch of the mapped pages, one frame per lane
            host_lanes[i].enqueue_frame(params[i], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            host_lanes[i].exchange_frame()
        for t_ in host_lanes:
            t_.sync()
            if rank == 0:
                t_.exchange_release()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        p = poses[(args.warmup + i) % len(poses)]
        k = i % len(lanes)
        if host_lanes:
            # frame i: kernels -> tiles stored into the shared host frame of lane k (+ arrival flags); a lane is synchronised -- on
            # rank 0 that means the whole frame is in host memory -- and released before it is given the next frame
            k = i % len(host_lanes)
            if i >= len(host_lanes):
                host_lanes[k].sync()
                if rank == 0:
                    host_lanes[k].exchange_release()
            host_lanes[k].enqueue_frame(camera.trace_params(p, info, args.levels, W, H), dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            host_lanes[k].exchange_frame()
        elif world == 1:
            if i >= len(lanes):
                lanes[k].sync()
            lanes[k].enqueue_frame(camera.trace_params(p, info, args.levels, W, H), dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True,
                                   host_frames[k].data_ptr())
        else:
            # N > 1: the same two lanes as the device-timed loop.  Frame i (kernels -> NCCL gather -> assembly -> copy of the
            # assembled frame into pinned host memory, all on stream k) overlaps frame i+1 on the other stream; a lane is
            # synchronised -- its host frame complete -- before it is given the next frame.
            if i >= len(lanes):
                streams[k].synchronize()
            lanes[k].enqueue_frame(camera.trace_params(p, info, args.levels, W, H), dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            gather(k, release=False)
            if rank == 0:
                with torch.cuda.stream(streams[k]):
                    host_frames[k].copy_(frames[k], non_blocking=True)
                if peer:
                    lanes[k].exchange_release()          # after the copy on the same stream: the peers may overwrite frame k
    sync_all()
    for t_ in host_lanes:
        t_.sync()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = rays / (e2e_ms * 1e-3) / 1e6

    sync_all()
    arm_fused(False)    # the pass timings below render frames that are not exchanged

    # ---- N > 1: the SAME workload (same resolution, same poses) on ONE GPU, so that the scaling efficiency compares like
    # with like.  Rank 0 renders whole frames on three contexts, exactly as a 1-GPU run of this resolution would; the other
    # ranks wait at the barrier.
    single_ms = None
    if world > 1 and not os.environ.get("HDT_BENCH_NO_SINGLE"):
        if rank == 0:
            wl = [tracer.DAGTracer(True, W, H, args.levels, device=local_rank) for _ in range(3)]
            for t_ in wl:
                t_.set_option(tracer.OPT_BEAM_PREFETCH, 0 if args.no_beam_prefetch else 1)
            for i in range(args.warmup):
                wl[i % 3].enqueue_frame(params[i % len(params)], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            for t_ in wl:
                t_.sync()
            for t_ in wl:
                t_.timer_begin()
            for i in range(args.steps):
                wl[i % 3].enqueue_frame(params[(args.warmup + i) % len(params)], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            single_ms = max(t_.timer_end() for t_ in wl) / args.steps
            for t_ in wl:
                t_.close()
        barrier()
    # ---- per-pass kernel times on the sample poses (for the roofline) ------------------------
    sample_ids = [int(k * len(poses) / max(1, args.cpu_sample_poses)) for k in range(args.cpu_sample_poses)]
    pass_ms = {i: [0.0, 0.0, 0.0] for i in sample_ids}
    reps = 5
    if world == 1:
        for _ in range(reps):
            for i in sample_ids:
                ms = tr.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, None)
                for k in range(3):
                    pass_ms[i][k] += ms[k] / reps
    elif not args.no_cpu_baseline:
        # N > 1: each rank times its own (1/N of the tiles) passes on two sample poses; the slowest rank counts
        sample_ids = sample_ids[:2]
        pass_ms = {i: [0.0, 0.0, 0.0] for i in sample_ids}
        for _ in range(3):
            for i in sample_ids:
                ms = tr.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, None)
                for k in range(3):
                    pass_ms[i][k] += ms[k] / 3
        pt = torch.tensor([pass_ms[i] for i in sample_ids], dtype=torch.float64, device=dev)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        for row, i in zip(pt.tolist(), sample_ids):
            pass_ms[i] = row

    # ---- N > 1: the exchanged frame against the same frame rendered whole on rank 0 ----------
    exchange_mismatch = None
    if world > 1 and not os.environ.get("HDT_BENCH_NO_GATHER"):
        exchange_mismatch = 0
        arm_fused(True)
        whole = tracer.DAGTracer(True, W, H, args.levels, device=local_rank) if rank == 0 else None
        for i in sample_ids[:2]:
            lanes[0].enqueue_frame(params[i], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            gather(0, release=False)
            if rank == 0:
                with torch.cuda.stream(streams[0]):
                    host_frame.copy_(frames[0], non_blocking=True)
                if peer:
                    lanes[0].exchange_release()
            lanes[0].sync()
            streams[0].synchronize()
            if rank == 0:
                whole.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, None)
                exchange_mismatch += int((whole.read_colors() != host_frame.numpy().view(np.uint32).reshape(H, W)).sum())
        if whole is not None:
            whole.close()
        barrier()

    host_mismatch = None
    if host_lanes:
        # the frame assembled in shared host memory against the same frame rendered whole on rank 0
        host_mismatch = 0
        if rank == 0:
            for t_ in host_lanes:
                t_.exchange_release()
        whole = tracer.DAGTracer(True, W, H, args.levels, device=local_rank) if rank == 0 else None
        for i in sample_ids[:2]:
            host_lanes[0].enqueue_frame(params[i], dag_pod, dag.kind, col_pod, colors.kind, 1.0, 0.0, True, None)
            host_lanes[0].exchange_frame()
            host_lanes[0].sync()
            if rank == 0:
                from hashdag_b200 import partition
                T = 1 << tile_log2
                mt = partition.max_tiles_per_rank(world, W, H, tile_log2)
                tiles = host_views[0][: world * mt * T * T * 4].view(np.uint32).reshape(world, mt, T, T)
                got = partition.assemble([tiles[r] for r in range(world)], W, H, tile_log2)    # the host frame is in tile layout
                del tiles
                host_lanes[0].exchange_release()
                whole.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, None)
                host_mismatch += int((whole.read_colors() != got).sum())
        if whole is not None:
            whole.close()
        barrier()
        for t_ in host_lanes:
            t_.close()          # unregisters the pinned mapping
        del host_views[:]
        got = None
        for shm in host_shm:
            try:
                shm.close()
            except BufferError:  # a view is still alive somewhere: the mapping goes with the process
                pass
            if rank == 0:
                shm.unlink()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32+f64+u32", "data": "synthetic",
        "config": workload_config(args, scene, W, H, world),
        "implementation": product_implementation(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 96, "d2h_bytes_per_step": W * H * 4,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "setup_s": round(t_build, 1), "scene_build_s": round(scene.build_seconds, 1), "replication_bytes": int(replication_bytes),
        "timing": ("cudaEvents on the tracer streams around K enqueued frames; two frames in flight, " + ({"peer-fused": "tiles stored into rank 0's frame over peer memory by trace_shadows_kernel (hdt_exchange_*, fused)", "peer": "tiles stored into rank 0's frame over peer memory by a scatter kernel (hdt_exchange_*)", "nccl": "NCCL gather + assembly"}[args.exchange]) + " included, max over ranks" if gather
                   else f"cudaEvents on the tracer streams around K enqueued frames, {len(lanes)} frame(s) in flight"),
        "wall_ms_per_step": wall_ms / args.steps,
    }
    if exchange_mismatch is not None:
        out["parity_check_mismatched_pixels_vs_whole_frame_on_rank0"] = exchange_mismatch
    if host_mismatch is not None:
        out["parity_check_mismatched_pixels_host_frame_vs_whole_frame_on_rank0"] = host_mismatch
        out["e2e"]["path"] = ("every rank copies its compact tile buffer into its slice of pinned host memory shared by the ranks over its own PCIe link "
                              "(hdt_exchange_attach_host: one copy-engine transfer per rank and frame, tile layout), three frames in flight")
    if single_ms:
        single_value = rays / (single_ms * 1e-3 * args.steps) / 1e6
        out["single_gpu_same_workload"] = {"ms_per_step": single_ms, "value": single_value, "unit": UNIT, "frames_in_flight": 3,
                                           "what": f"the same {W}x{H} frames rendered whole on rank 0's GPU alone (three contexts in flight, as `bench.py --gpus 1 "
                                                   f"--width {W} --height {H}` does), measured in this run after the timed region"}
        out["strong_scaling_efficiency"] = value / (world * single_value)
        out["e2e_strong_scaling_efficiency"] = e2e_value / (world * single_value)

    # ---- CPU baseline + algorithmic bytes (bounded sample), roofline -------------------------
    if world == 1 and not args.no_cpu_baseline:
        from oracle import hdo
        odag = hdo.make_dag(scene, hdo.DAG_HASH if hashed else hdo.DAG_BASIC)
        ocol = hdo.make_colors(scene, hdo.COLORS_HASH if hashed else hdo.COLORS_COMPRESSED)
        cores = os.cpu_count() or 1
        cpu_rays, cpu_s = 0, 0.0
        bytes_pass = {"paths": 0, "colors": 0, "shadows": 0}
        gpu_ms_pass = {"paths": 0.0, "colors": 0.0, "shadows": 0.0}
        mismatch = 0
        for i in sample_ids:
            dt, n, (sp, sc, ss), img = oracle_frame(hdo, odag, ocol, params[i], W, H)
            cpu_rays += n
            cpu_s += dt
            b = algorithmic_bytes(sp, sc, ss, hashed, W, H)
            for k, name in enumerate(("paths", "colors", "shadows")):
                bytes_pass[name] += b[name]
                gpu_ms_pass[name] += pass_ms[i][k]
            tr.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, host_frame)
            mismatch += int((host_frame.numpy().view(np.uint32).reshape(H, W) != img).sum())
        out["cpu_baseline"] = {"value": cpu_rays / cpu_s / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{len(sample_ids)} full {W}x{H} frames (poses {sample_ids}) of the same fly-through, oracle/ on {cores} threads"}
        out["parity_check_mismatched_pixels_vs_oracle"] = mismatch
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        dominant = max(gpu_ms_pass, key=lambda k: gpu_ms_pass[k])
        traffic, limiter = None, None
        ncu_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(ncu_path):   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the pass' kernels, same config
            try:
                ncu = json.load(open(ncu_path))
                traffic = ncu.get("dram_bytes_per_pass", {}).get(dominant)
                k = next((v for n, v in ncu.get("kernels", {}).items() if n.startswith(f"trace_{dominant}")), None)
                if k:   # what the counters say bounds the kernel (the HBM figure above is the SURVEY §8d algorithmic-bytes roofline)
                    limiter = {"what": "instruction issue" if k.get("issue_active_pct", 0) > 60 and k.get("dram_pct_of_peak", 100) < 40 else "memory latency",
                               "issue_slots_busy_pct": k.get("issue_active_pct"), "threads_per_warp_instruction": k.get("threads_per_warp_inst"),
                               "dram_pct_of_peak": k.get("dram_pct_of_peak"), "l1_hit_pct": k.get("l1_hit_pct"), "l2_hit_pct": k.get("l2_hit_pct"),
                               "warp_instructions_per_launch": k.get("warp_inst_executed"), "source": "profiles/ncu_summary.json (ncu --set full of the shipped kernels, round 2)"}
            except Exception:
                traffic = None
        ach = {k: (bytes_pass[k] / (gpu_ms_pass[k] * 1e-3) / 1e9 if gpu_ms_pass[k] > 0 else 0.0) for k in bytes_pass}
        out["roofline"] = {"bound": "hbm", "kernel": f"{dominant} pass = setup_{dominant}_kernel + beam_{dominant}_kernel + trace_{dominant}_kernel" if dominant != "colors" else "trace_colors_kernel", "achieved": ach[dominant], "peak": peak, "unit": "GB/s",
                           "frac": ach[dominant] / peak, "traffic": traffic, "peak_source": peak_src,
                           "limited_by": limiter,
                           "algorithmic_bytes_per_launch": bytes_pass[dominant] / len(sample_ids),
                           "avg_launch_ms": gpu_ms_pass[dominant] / len(sample_ids),
                           "per_pass": {k: {"ms": gpu_ms_pass[k] / len(sample_ids), "algorithmic_GBps": ach[k],
                                            "bytes_per_launch": bytes_pass[k] / len(sample_ids)} for k in bytes_pass}}
    # ---- the north star's anchor: the reference's own CUDA kernels on this GPU, scene and camera path ----------------
    if world == 1 and not args.no_ref_cuda:
        from oracle import ref
        if ref.available(args.levels, W, H):
            step_ids = [(args.warmup + i) % len(poses) for i in range(args.steps)]
            warm_ids = [i % len(poses) for i in range(min(args.warmup, 3))]
            # ours, synchronous: one frame at a time, nothing else in flight -- (a) the whole-frame call (one synchronisation per
            # frame), (b) the three calls of the reference's interface, each synchronous like the reference's
            for t_ in lanes:
                t_.sync()
            for i in warm_ids:
                tr.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, None)
            fp = [0.0, 0.0, 0.0]
            t0 = time.perf_counter()
            for i in step_ids:
                ms = tr.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, None)
                fp = [a + b for a, b in zip(fp, ms)]
            frame_wall_ms = (time.perf_counter() - t0) * 1e3 / len(step_ids)
            sp = [0.0, 0.0, 0.0]
            for i in step_ids:
                sp[0] += tr.resolve_paths(poses[i], info, dag)
                sp[1] += tr.resolve_colors(dag, colors)
                sp[2] += tr.resolve_shadows(poses[i], info, dag, 1.0, 0.0)
            n = len(step_ids)
            ours_frame = {k: v / n for k, v in zip(("paths", "colors", "shadows"), fp)}
            ours_calls = {k: v / n for k, v in zip(("paths", "colors", "shadows"), sp)}
            ours_img = {}
            for i in sample_ids[:2]:
                tr.resolve_frame(poses[i], info, dag, colors, 1.0, 0.0, True, host_frame)
                ours_img[i] = host_frame.numpy().view(np.uint32).reshape(H, W).copy()
            rt = ref.RefTracer(args.levels, W, H)       # (the reference prints its progress: descriptor 1 is stderr, see main())
            rt.load_scene(scene)
            dk, ck = (1, 3) if hashed else (0, 1)
            ref_ms = time_reference_kernels(rt, poses, info, step_ids, warm_ids, dk, ck)
            ref_bad = 0
            for i in sample_ids[:2]:
                rt.resolve_paths(dk, poses[i], info); rt.resolve_colors(dk, ck); rt.resolve_shadows(dk, poses[i], info, 1.0, 0.0)
                ref_bad += int((rt.read_colors() != ours_img[i]).sum())
            rt.close()
            ref_total = sum(ref_ms.values())
            out["ref_cuda"] = {"what": "oracle/_ref: the UNMODIFIED reference kernels (tracer.cu:145-697 via dag_tracer.cu:116-219) compiled for sm_100a, same GPU, "
                                       "DAG and the same K poses, after the timed region; kernel times from the reference's own cudaEvents, three synchronous calls per frame",
                               "ms_per_step": ref_total, "passes_ms": ref_ms, "value": rays / (ref_total * 1e-3 * args.steps) / 1e6, "unit": UNIT}
            out["ours_synchronous"] = {"three_calls_ms": ours_calls, "three_calls_ms_per_step": sum(ours_calls.values()),
                                       "frame_call_ms": ours_frame, "frame_call_ms_per_step": sum(ours_frame.values()),
                                       "frame_call_wall_ms_per_step": frame_wall_ms,
                                       "note": "three_calls = hdt_resolve_paths / _colors / _shadows, each synchronous like DAGTracer::resolve_*; frame_call = hdt_resolve_frame "
                                               "(the same three passes, one synchronisation, the shadow pass' set-up and beams beside the colours kernel)"}
            out["vs_ref_cuda"] = {"pipelined": ref_total / (elapsed_ms / args.steps), "e2e": ref_total / (e2e_ms / args.steps),
                                  "synchronous": ref_total / sum(ours_calls.values()), "synchronous_frame_call": ref_total / sum(ours_frame.values()),
                                  "per_pass": {k: ref_ms[k] / ours_calls[k] for k in ref_ms},
                                  "note": "reference ms / ours ms on the same frames; pipelined = K frames enqueued on "
                                          f"{len(lanes)} contexts (the headline `value`), synchronous = one call at a time like the reference"}
            out["parity_check_mismatched_pixels_vs_reference_kernels"] = ref_bad
        else:
            out["ref_cuda"] = {"unavailable": f"oracle/_ref has no build for depth {args.levels} at {W}x{H} (oracle/build_ref.py)"}

    if world > 1 and not args.no_cpu_baseline:
        # roofline of one rank's launch: the frame's algorithmic bytes (oracle access counts on rank 0's host cores) / N,
        # over the slowest rank's pass time
        from oracle import hdo
        odag = hdo.make_dag(scene, hdo.DAG_HASH if hashed else hdo.DAG_BASIC)
        ocol = hdo.make_colors(scene, hdo.COLORS_HASH if hashed else hdo.COLORS_COMPRESSED)
        bytes_pass = {"paths": 0, "colors": 0, "shadows": 0}
        gpu_ms_pass = {"paths": 0.0, "colors": 0.0, "shadows": 0.0}
        for i in sample_ids:
            _, _, (sp, sc, ss), _ = oracle_frame(hdo, odag, ocol, params[i], W, H)
            b = algorithmic_bytes(sp, sc, ss, hashed, W, H)
            for k, name in enumerate(("paths", "colors", "shadows")):
                bytes_pass[name] += b[name] / world
                gpu_ms_pass[name] += pass_ms[i][k]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        dominant = max(gpu_ms_pass, key=lambda k: gpu_ms_pass[k])
        ach = {k: (bytes_pass[k] / (gpu_ms_pass[k] * 1e-3) / 1e9 if gpu_ms_pass[k] > 0 else 0.0) for k in bytes_pass}
        out["roofline"] = {"bound": "hbm", "kernel": f"{dominant} pass of ONE rank (1/{world} of the tiles) = setup_{dominant}_kernel + beam_{dominant}_kernel + trace_{dominant}_kernel" if dominant != "colors" else "trace_colors_kernel of one rank",
                           "achieved": ach[dominant], "peak": peak, "unit": "GB/s", "frac": ach[dominant] / peak, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": bytes_pass[dominant] / len(sample_ids), "avg_launch_ms": gpu_ms_pass[dominant] / len(sample_ids),
                           "note": f"per-GPU figure: the frame's algorithmic bytes / {world} over the slowest rank's synchronous pass time, {len(sample_ids)} sample poses"}
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def scene_meta(scene, poses):
    names = ["hash_pool", "hash_page_table", "weights", "blocks", "macro_blocks", "color_nodes", "color_offsets", "basic", "enclosed_leaves"]
    return {
        "arrays": {n: (int(getattr(scene, n).size), int(getattr(scene, n).dtype.itemsize)) for n in names},
        "levels": scene.levels, "top_levels": scene.top_levels, "pool_top": scene.hash_pool_top, "first": scene.hash_first_node_index,
        "bounds_min": list(scene.bounds_min), "bounds_max": list(scene.bounds_max),
        "poses": [[list(p.position), [list(r) for r in p.rotation]] for p in poses],
    }


def replicas_from_tensors(tracer, t, meta, hashed):
    leaf = tracer.CompressedColorLeaf(t["weights"], t["blocks"], t["macro_blocks"], tracer.UNIQUE_OFFSET)
    if hashed:
        dag = tracer.HashDAG(t["hash_pool"], t["hash_page_table"], meta["pool_top"], meta["first"], meta["levels"])
        return dag, tracer.HashDAGColors(t["color_nodes"], t["color_offsets"], leaf)
    dag = tracer.BasicDAG(t["basic"], meta["levels"])
    return dag, tracer.BasicDAGCompressedColors(meta["top_levels"], t["enclosed_leaves"], leaf)


def main():
    args = parse_args()
    # ONE JSON line on stdout.  Native libraries write to file descriptor 1 behind Python's back (NCCL's version banner, the
    # reference's progress messages and its "No leaks!" from a static destructor at exit), so descriptor 1 is pointed at
    # stderr for the whole run and Python's sys.stdout keeps the real stdout for the one line this script prints.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference_cpu(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
