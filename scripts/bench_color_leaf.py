"""Colour-leaf rebuild on the GPU against the CPU oracle.  Prints one JSON line.

    python scripts/bench_color_leaf.py [--footprint-log2 14] [--reps 10]

Workloads: (whole) the bench scene's main colour leaf re-encoded in one call -- COPY(0, n_voxels), the streaming
limit; (edit) a leaf of the size an edit touches at depth 17 (128^3 region): copy / fill / copy.
Algorithmic bytes = old leaf arrays read once + new leaf arrays written once.  cpu_baseline = oracle/color_leaf.py
(numpy restatement of ColorLeafBuilder + copy_colors, one thread) on a bounded sample of the same leaf."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--footprint-log2", type=int, default=14)
    ap.add_argument("--levels", type=int, default=17)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=16 * 1024 * 1024)
    a = ap.parse_args()
    import torch
    from hashdag_b200 import color_leaf as host, tracer
    from hashdag_b200.scene import build_scene
    from oracle import color_leaf as cl

    scene = build_scene(a.levels, a.footprint_log2, seed=1337, n_spheres=6)
    n = int(scene.n_voxels)
    t = tracer.DAGTracer(True, 256, 256, a.levels)
    old = tracer.CompressedColorLeaf.from_scene(scene)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6553.0))

    def run(builder, reps):
        ms, leaf = [], None
        for _ in range(reps):
            leaf, m = builder.build(t, old)
            ms.append(m)
        return leaf, float(np.median(ms)), float(np.min(ms))

    out = {"metric": "colour-leaf rebuild", "unit": "Mcolours/s", "levels": a.levels, "footprint_log2": a.footprint_log2}
    # ---- whole leaf ------------------------------------------------------------------------------------
    b = host.ColorLeafBuilder()
    b.copy_colors(0, n)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    leaf, med, best = run(b, a.reps)
    wall = (time.perf_counter() - w0) / a.reps * 1e3
    nw = 0 if leaf.weights is None else leaf.weights.numel()
    in_bytes = scene.weights.nbytes + scene.blocks.nbytes + scene.macro_blocks.nbytes
    out_bytes = 4 * nw + 8 * leaf.blocks.numel() + 8 * leaf.macro_blocks.numel()
    out["whole"] = {"colours": n, "kernel_ms": med, "kernel_ms_best": best, "call_wall_ms": wall, "mcolours_s": n / med / 1e3,
                    "algorithmic_bytes": in_bytes + out_bytes, "achieved_GBps": (in_bytes + out_bytes) / med / 1e6, "peak_GBps": peak,
                    "frac": (in_bytes + out_bytes) / med / 1e6 / peak, "blocks": int(leaf.blocks.numel()), "weight_words": int(nw)}
    # the same leaf from colour 5000 on: no macro block of the new leaf lines up with one of the old leaf (every segment is searched for,
    # spans two old macro blocks, and its weights move by a bit offset)
    u = host.ColorLeafBuilder()
    u.copy_colors(5000, n - 5000)
    _, umed, ubest = run(u, a.reps)
    out["whole_unaligned"] = {"colours": n - 5000, "kernel_ms": umed, "kernel_ms_best": ubest, "mcolours_s": (n - 5000) / umed / 1e3,
                              "frac": (in_bytes + out_bytes) / umed / 1e6 / peak}
    # parity on the sample the CPU baseline encodes
    ns = min(a.cpu_sample, n)
    c0 = time.perf_counter()
    want = cl.rebuild(np.array([(0, ns, cl.OP_COPY, 0, 0, 0)], dtype=cl.OP_DTYPE), (scene.weights, scene.blocks, scene.macro_blocks, None))
    cpu_s = time.perf_counter() - c0
    bs = host.ColorLeafBuilder(); bs.copy_colors(0, ns)
    got, _ = bs.build(t, old)
    same = (np.array_equal(got.blocks.cpu().numpy().view(np.uint64), want[1]) and np.array_equal(got.macro_blocks.cpu().numpy().view(np.uint64), want[2])
            and np.array_equal(got.weights.cpu().numpy().view(np.uint32), want[0]))
    out["cpu_baseline"] = {"value": ns / cpu_s / 1e6, "unit": "Mcolours/s", "cores": 1, "kind": "port",
                           "sample": f"first {ns} colours of the same leaf, oracle/color_leaf.py (numpy)"}
    out["parity_sample_identical"] = bool(same)
    # ---- edit-sized leaf ---------------------------------------------------------------------------------
    e = host.ColorLeafBuilder()
    start = n // 3
    e.copy_colors(start, 110_000)
    e.add_large_single_color((0.8, 0.3, 0.1), 65_536)
    for i in range(300):
        e.add(0x12340000 + i, i % 8, 3)
    e.copy_colors(start + 150_000, 115_000)
    w0 = time.perf_counter()
    leaf2, med2, best2 = run(e, 50)
    wall2 = (time.perf_counter() - w0) / 50 * 1e3
    out["edit"] = {"colours": e.get_color_index(), "ops": int(e.ops().size), "kernel_ms": med2, "kernel_ms_best": best2, "call_wall_ms": wall2}
    want2 = cl.rebuild(e.ops().astype(cl.OP_DTYPE), (scene.weights, scene.blocks, scene.macro_blocks, None))
    out["edit"]["identical_to_oracle"] = bool(np.array_equal(leaf2.blocks.cpu().numpy().view(np.uint64), want2[1])
                                              and np.array_equal(leaf2.weights.cpu().numpy().view(np.uint32), want2[0])
                                              and np.array_equal(leaf2.macro_blocks.cpu().numpy().view(np.uint64), want2[2]))
    t.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
