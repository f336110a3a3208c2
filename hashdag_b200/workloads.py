"""Synthetic workloads of BASELINE.json's configs: scene recipe + seeded camera fly-through.

The reference's own benchmark replays a recorded camera path over a scene file
(python/benchmark_rt.py:12-23, replays/*_move.csv); neither the scenes nor a depth-17 replay are
available offline, so the path here is a closed curve over the synthetic terrain."""
from __future__ import annotations

import math

from . import camera
from .scene import build_scene


def flythrough_xz(levels, footprint_log2, n_poses):
    c = float(1 << (levels - 1))
    s = float(1 << footprint_log2)
    pts = []
    for i in range(n_poses + 1):
        t = i / n_poses
        pts.append((c + 0.30 * s * math.cos(2 * math.pi * t), c + 0.22 * s * math.sin(4 * math.pi * t + 0.4)))
    return pts


def build_workload(levels=17, footprint_log2=14, n_poses=64, seed=1337, n_spheres=24, n_threads=0, uncompressed=False):
    """-> (scene, [CameraView]*n_poses).  Cameras fly 1.5-5% of the footprint above the ground,
    looking ahead and down, so most pixels hit geometry at full depth and the rest see sky."""
    xz = flythrough_xz(levels, footprint_log2, n_poses)
    s = float(1 << footprint_log2)
    probes = [(int(x), int(z)) for x, z in xz]
    ahead = []
    for i in range(n_poses):
        x0, z0 = xz[i]
        x1, z1 = xz[i + 1]
        d = math.hypot(x1 - x0, z1 - z0) or 1.0
        ahead.append((x0 + (x1 - x0) / d * 0.12 * s, z0 + (z1 - z0) / d * 0.12 * s))
    probes += [(int(x), int(z)) for x, z in ahead]
    scene = build_scene(levels, footprint_log2, seed=seed, n_spheres=n_spheres, n_threads=n_threads, build_uncompressed=uncompressed,
                        height_probes=probes)
    base = float(1 << (levels - 1))

    def ground(x, z):
        h = scene.heights.get((int(x), int(z)), -1)
        return float(h) if h >= 0 else base

    poses = []
    for i in range(n_poses):
        x, z = xz[i]
        ax, az = ahead[i]
        alt = s * (0.015 + 0.035 * (0.5 + 0.5 * math.sin(6 * math.pi * i / n_poses)))
        eye = (x, max(ground(x, z), ground(ax, az) - 0.02 * s) + alt, z)
        tgt = (ax, ground(ax, az), az)
        poses.append(camera.look_at(eye, tgt))
    return scene, poses
