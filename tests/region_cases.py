"""Region-query cases shared by the golden generator, the CPU suite and the GPU suite."""
import numpy as np


def cases(scene, n_random=10, seed=3):
    """[(start, size)]: around the terrain surface, in the air, underground, aligned and not, degenerate sizes."""
    c = 1 << (scene.levels - 1)
    h0 = int(scene.heights[(c, c)])
    out = [((c - 7, h0 - 13, c + 3), (29, 31, 18)), ((c, h0 - 8, c), (16, 16, 16)), ((c - 64, h0 - 32, c - 64), (128, 64, 128)),
           ((c + 5, h0 + 300, c + 5), (24, 24, 24)), ((c - 3, h0 - 2, c - 3), (1, 1, 1)), ((c - 3, h0 - 2, c - 3), (2, 5, 3)),
           ((c + 1, h0 - 40, c + 2), (4, 70, 4)), ((c - 100, h0 - 1, c + 40), (97, 3, 5)), ((0, 0, 0), (8, 8, 8)),
           (((1 << scene.levels) - 8, (1 << scene.levels) - 8, (1 << scene.levels) - 8), (8, 8, 8))]
    rng = np.random.default_rng(seed)
    for _ in range(n_random):
        st = (c + int(rng.integers(-200, 200)), h0 + int(rng.integers(-30, 30)), c + int(rng.integers(-200, 200)))
        sz = tuple(int(v) for v in rng.integers(1, 48, 3))
        out.append((st, sz))
    return out


def is_empty_levels(scene):
    return list(range(0, scene.levels - 1))      # 0 .. leaf level (dag_utils.h:264)
