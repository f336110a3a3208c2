import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Host libraries (scene builder, oracle) are compiled on demand; the CUDA library too where nvcc exists."""
    import shutil
    import __graft_entry__ as g
    if shutil.which("nvcc") and shutil.which("g++"):
        g.build()
    yield


_SCENES = {}


def get_scene(levels, footprint_log2, seed=1337, n_spheres=6, uncompressed=False):
    from hashdag_b200.scene import build_scene
    key = (levels, footprint_log2, seed, n_spheres, uncompressed)
    if key not in _SCENES:
        c = 1 << (levels - 1)
        _SCENES[key] = build_scene(levels, footprint_log2, seed=seed, n_spheres=n_spheres, build_uncompressed=uncompressed,
                                   height_probes=[(c, c)])
    return _SCENES[key]


def scene_cameras(scene, n=4, footprint_log2=10):
    """A few deterministic poses: orbit above the terrain, one close to the ground, one axis-parallel."""
    from hashdag_b200 import camera
    c = float(1 << (scene.levels - 1))
    h0 = float(scene.heights.get((int(c), int(c)), c))
    r = float(1 << footprint_log2) * 0.45
    poses = camera.orbit_poses((c, h0, c), r, r * 0.5, n, phase=0.3)
    poses.append(camera.look_at((c + 3.0, h0 + 40.0, c + 5.0), (c + 90.0, h0 - 10.0, c + 70.0)))
    # straight down an axis: exercises the inf / NaN slab cases of compute_intersection_mask
    poses.append(camera.CameraView((c, h0 + r, c), ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))))
    return poses
