"""CPU: the beam pre-pass (hashdag_b200/csrc/hdt_beam.cuh) as a host model on top of the oracle
(tests/cpp/beam_model.cpp): resuming every ray from its tile's interval-DFS state must give exactly
the pixels of the plain per-ray DFS -- primary and shadow rays, BasicDAG and HashDAG, ordinary and
nasty cameras -- and must actually save visits.  The CUDA implementation is checked against the
oracle on the GPU (test_gpu_parity.py); this model is what can run without one."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, get_scene, scene_cameras
from hashdag_b200 import camera
from oracle import hdo


@pytest.fixture(scope="module")
def model():
    out = os.path.join(ROOT, "build", "libbeam_model.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "beam_model.cpp")
    deps = [src, os.path.join(ROOT, "oracle", "hdo_oracle.cpp"), os.path.join(ROOT, "oracle", "hdo_oracle.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", src, "-o", out], check=True)
    return C.CDLL(out)


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _special(s, fp):
    c = float(1 << (s.levels - 1))
    size = float(1 << s.levels)
    h0 = float(s.heights.get((int(c), int(c)), c))
    r = float(1 << fp)
    return [
        camera.look_at((-0.3 * size, 1.2 * size, -0.2 * size), (c, h0, c)),
        camera.look_at((c - 2.0 * r, h0 + 0.1 * r, c), (c + r, h0 + 0.1 * r, c + 1e-3)),
        camera.CameraView((c + 0.25, h0 + 30.5, c + 0.75), ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))),
        camera.look_at((c + 7.3, h0 + 3.1, c - 5.2), (c + 40.0, h0 - 6.0, c + 33.0)),
    ]


@pytest.mark.parametrize("levels,fp,kind", [(13, 10, "hash"), (13, 10, "basic"), (17, 11, "hash")])
def test_resuming_from_the_tile_state_is_exact_and_saves_visits(model, levels, fp, kind):
    s = get_scene(levels, fp)
    dag = hdo.make_dag(s, hdo.DAG_HASH if kind == "hash" else hdo.DAG_BASIC)
    info = camera.DAGInfo(s.bounds_min, s.bounds_max)
    W, H = 240, 136
    tot = np.zeros(4, dtype=np.float64)
    for cam in scene_cameras(s, 2, fp) + _special(s, fp):
        prm = camera.trace_params(cam, info, levels, W, H)
        for tw, th in ((8, 4), (2, 2)):
            out = np.zeros(160, np.uint64)
            paths = np.zeros((H, W, 4), np.uint32)
            model.beam_paths(C.byref(dag), W, H, _d3(prm[0]), _d3(prm[1]), _d3(prm[2]), _d3(prm[3]), tw, th,
                             paths.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            assert out[3] == 0, f"{out[3]} primary rays change when resumed from a {tw}x{th} beam"
            want, _ = hdo.trace_paths(dag, W, H, prm)
            assert np.array_equal(paths, want)
            if (tw, th) == (8, 4):
                tot[0] += out[0]; tot[1] += out[1]
            out = np.zeros(160, np.uint64)
            model.beam_shadows(C.byref(dag), W, H, _d3(prm[0]), _d3(prm[1]), _d3(prm[2]), _d3(prm[3]), tw, th,
                               paths.ctypes.data_as(C.c_void_p), C.c_float(1.0), out.ctypes.data_as(C.c_void_p))
            assert out[3] == 0, f"{out[3]} shadow rays change when resumed from a {tw}x{th} beam"
            if (tw, th) == (8, 4):
                tot[2] += out[0]; tot[3] += out[1]
    assert tot[1] < 0.9 * tot[0] and tot[3] < 0.9 * tot[2], f"visits with/without beams: primary {tot[1]:.0f}/{tot[0]:.0f}, shadow {tot[3]:.0f}/{tot[2]:.0f}"
