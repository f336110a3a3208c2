"""GPU: the C++ DAGTracer shim (hashdag_b200/cpp/dag_tracer_b200.h), driven by a small headless C++
harness the way engine.cpp drives the reference's DAGTracer, against the CPU oracle."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, get_scene, scene_cameras
from hashdag_b200 import camera
from oracle import hdo

pytestmark = pytest.mark.gpu

W, H = 256, 160


def _blob(f, arr):
    raw = np.ascontiguousarray(arr).tobytes() if arr is not None else b""
    f.write(struct.pack("<Q", len(raw)))
    f.write(raw)


def test_cpp_shim_matches_oracle(tmp_path):
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    exe = os.path.join(ROOT, "build", "shim_harness")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    lib_dir = os.path.join(ROOT, "hashdag_b200")
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_harness.cpp"),
                    "-L" + lib_dir, "-lhashdag_b200", "-Xlinker", "-rpath=" + lib_dir], check=True)
    s = get_scene(13, 10)
    cam = scene_cameras(s, 1, 10)[0]
    scene_file, out_file = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(scene_file, "wb") as f:
        f.write(struct.pack("<8I", s.levels, W, H, s.hash_first_node_index, s.hash_pool_top, s.top_levels, 0, 0))
        f.write(struct.pack("<18d", *cam.position, *[v for r in cam.rotation for v in r], *s.bounds_min, *s.bounds_max))
        for a in (s.basic, s.enclosed_leaves, s.hash_pool, s.hash_page_table, s.weights, s.blocks, s.macro_blocks, s.color_nodes, s.color_offsets):
            _blob(f, a)
    r = subprocess.run([exe, scene_file, out_file], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_file, dtype=np.uint32)
    per = W * H * 5
    assert raw.size == 2 * per
    prm = camera.trace_params(cam, camera.DAGInfo(s.bounds_min, s.bounds_max), s.levels, W, H)
    for i, (dk, ck) in enumerate(((hdo.DAG_BASIC, hdo.COLORS_COMPRESSED), (hdo.DAG_HASH, hdo.COLORS_HASH))):
        paths = raw[i * per: i * per + W * H * 4].reshape(H, W, 4)
        colors = raw[i * per + W * H * 4: (i + 1) * per].reshape(H, W)
        d = hdo.make_dag(s, dk)
        op, _ = hdo.trace_paths(d, W, H, prm)
        oc, _ = hdo.trace_colors(d, hdo.make_colors(s, ck), op)
        osh, _ = hdo.trace_shadows(d, prm, op, oc, 1.0, 0.0)
        assert np.array_equal(paths, op)
        assert np.array_equal(colors, osh)
