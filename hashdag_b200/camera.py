"""Host-side camera maths of the tracer boundary.

Mirrors CameraView (/root/reference/src/camera_view.h:8-27), DAGInfo (dag_info.h:5-9) and
get_trace_params (/root/reference/src/dag_tracer.cu:71-113): same operations in the same order in
IEEE double, so the four double3 handed to the kernels are the ones the reference would compute.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

FOV_DEGREES = 60.0  # CameraView::fov, camera_view.h:10


@dataclass
class DAGInfo:
    bounds_min: tuple = (0.0, 0.0, 0.0)
    bounds_max: tuple = (1.0, 1.0, 1.0)


@dataclass
class CameraView:
    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = field(default_factory=lambda: ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)))

    def right(self):
        return tuple(-v for v in self.rotation[0])

    def up(self):
        return tuple(self.rotation[1])

    def forward(self):
        return tuple(self.rotation[2])


def _norm(v):
    l = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return (v[0] / l, v[1] / l, v[2] / l)


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def look_at(position, target, up_hint=(0.0, 1.0, 0.0)) -> CameraView:
    f = _norm(tuple(t - p for t, p in zip(target, position)))
    r = _norm(_cross(f, up_hint))
    u = _cross(r, f)
    return CameraView(tuple(float(p) for p in position), (tuple(-x for x in r), u, f))


def trace_params(camera: CameraView, info: DAGInfo, levels: int, width: int, height: int):
    """-> (cameraPosition, rayMin, rayDDx, rayDDy) in voxel space, each a 3-tuple of float."""
    fov = FOV_DEGREES / 2.0 * (math.pi / 180.0)
    aspect = float(width) / float(height)
    s, c = math.sin(fov), math.cos(fov)
    right, up, fwd = camera.right(), camera.up(), camera.forward()
    cam, rmin, ddx, ddy = [], [], [], []
    for k in range(3):
        X = right[k] * s * aspect
        Y = up[k] * s
        Z = fwd[k] * c
        p = camera.position[k]
        bl = p + Z - Y - X
        br = p + Z - Y + X
        tl = p + Z + Y - X
        translation = -info.bounds_min[k]
        scale = float(1 << levels) / (info.bounds_max[k] - info.bounds_min[k])
        fp = (p + translation) * scale
        fbl = (bl + translation) * scale
        ftl = (tl + translation) * scale
        fbr = (br + translation) * scale
        cam.append(fp)
        rmin.append(fbl)
        ddx.append((fbr - fbl) * (1.0 / width))
        ddy.append((ftl - fbl) * (1.0 / height))
    return tuple(cam), tuple(rmin), tuple(ddx), tuple(ddy)


def orbit_poses(center, radius, height, n, phase=0.0):
    """n cameras on a circle around `center`, looking at it (seed-free, deterministic)."""
    out = []
    for i in range(n):
        a = phase + 2.0 * math.pi * i / n
        pos = (center[0] + radius * math.cos(a), center[1] + height, center[2] + radius * math.sin(a))
        out.append(look_at(pos, center))
    return out
