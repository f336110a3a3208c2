"""ctypes front-end of the synthetic scene builder (scene_builder.cpp).

Produces, as numpy arrays, the reference's own data formats: BasicDAG words
(/root/reference/src/dags/basic_dag/basic_dag.h:11-45), HashDAG pool + page table
(hash_table.h:156-173), compressed colours (variable_weight_size_colors.h:157-191) and the
HashDAGColors tree (hash_dag_colors.h:65-73).  Inputs for tests and bench.py only.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhashdag_scene.so")


class _Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "levels", "footprint_log2", "seed", "n_spheres", "roughness", "n_threads",
        "build_hash", "build_colors", "build_uncompressed", "finest_cell_log2")]


class _Info(C.Structure):
    _fields_ = [
        ("levels", C.c_uint32), ("top_levels", C.c_uint32),
        ("n_voxels", C.c_uint64), ("basic_words", C.c_uint64), ("enclosed_leaves", C.c_uint64),
        ("hash_page_table_size", C.c_uint32), ("hash_pool_top", C.c_uint32),
        ("hash_first_node_index", C.c_uint32), ("has_hash_colors", C.c_uint32),
        ("n_weight_words", C.c_uint64), ("n_blocks", C.c_uint64), ("n_macro_words", C.c_uint64),
        ("n_color_nodes", C.c_uint64), ("n_color_offsets", C.c_uint64),
        ("n_uncompressed", C.c_uint64), ("build_seconds", C.c_double),
        ("nodes_per_level", C.c_uint64 * 32),
    ]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(_LIB_PATH)
        lib.hds_build.restype = C.c_void_p
        lib.hds_build.argtypes = [C.POINTER(_Params)]
        lib.hds_free.argtypes = [C.c_void_p]
        lib.hds_get_info.argtypes = [C.c_void_p, C.POINTER(_Info)]
        lib.hds_terrain_height.restype = C.c_int32
        lib.hds_terrain_height.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        for name, ty in (("hds_basic_data", C.c_uint32), ("hds_enclosed_leaves", C.c_uint64),
                         ("hds_hash_pool", C.c_uint32), ("hds_hash_page_table", C.c_uint32), ("hds_hash_bucket_sizes", C.c_uint32),
                         ("hds_color_weights", C.c_uint32), ("hds_color_blocks", C.c_uint64),
                         ("hds_color_macro_blocks", C.c_uint64), ("hds_color_uncompressed", C.c_uint32),
                         ("hds_hash_color_nodes", C.c_uint32), ("hds_hash_color_offsets", C.c_uint64)):
            fn = getattr(lib, name)
            fn.restype = C.POINTER(ty)
            fn.argtypes = [C.c_void_p]
        lib.hds_hash_bucket_count.restype = C.c_uint64
        lib.hds_hash_bucket_count.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


@dataclass
class Scene:
    """Host-side copy of one synthetic scene in the reference's formats."""
    levels: int
    top_levels: int
    n_voxels: int
    bounds_min: tuple = (0.0, 0.0, 0.0)
    bounds_max: tuple = (0.0, 0.0, 0.0)
    basic: np.ndarray = None             # uint32 BasicDAG words, root at 0
    enclosed_leaves: np.ndarray = None   # uint64
    hash_pool: np.ndarray = None         # uint32, pool_top*512 words
    hash_page_table: np.ndarray = None   # uint32
    hash_bucket_sizes: np.ndarray = None # uint32, words used per bucket (hash_table.h:18-35 order)
    hash_pool_top: int = 0
    hash_first_node_index: int = 0
    weights: np.ndarray = None           # uint32 (byte-swapped words)
    blocks: np.ndarray = None            # uint64
    macro_blocks: np.ndarray = None      # uint64 pairs
    uncompressed: np.ndarray = None      # uint32 RGB888 per voxel
    color_nodes: np.ndarray = None       # uint32 HashDAGColors tree
    color_offsets: np.ndarray = None     # uint64
    nodes_per_level: list = field(default_factory=list)
    build_seconds: float = 0.0
    heights: dict = field(default_factory=dict)
    params: dict = field(default_factory=dict)

    @property
    def has_hash(self):
        return self.hash_pool is not None and self.hash_pool.size > 0

    @property
    def has_hash_colors(self):
        return self.color_nodes is not None and self.color_nodes.size > 0


def build_scene(levels: int, footprint_log2: int | None = None, seed: int = 1337, n_spheres: int = 8,
                roughness: int = 1, n_threads: int = 0, build_hash: bool = True, build_colors: bool = True,
                build_uncompressed: bool = False, finest_cell_log2: int = 2,
                height_probes: list | None = None) -> Scene:
    lib = _load()
    if footprint_log2 is None:
        footprint_log2 = levels
    p = _Params(levels, footprint_log2, seed, n_spheres, roughness, n_threads,
                int(build_hash), int(build_colors), int(build_uncompressed), finest_cell_log2)
    h = lib.hds_build(C.byref(p))
    if not h:
        raise RuntimeError("hds_build failed")
    try:
        info = _Info()
        lib.hds_get_info(h, C.byref(info))
        size = float(1 << levels)
        sc = Scene(levels=info.levels, top_levels=info.top_levels, n_voxels=info.n_voxels,
                   bounds_min=(0.0, 0.0, 0.0), bounds_max=(size, size, size))
        sc.basic = _arr(lib.hds_basic_data(h), info.basic_words, np.uint32)
        sc.enclosed_leaves = _arr(lib.hds_enclosed_leaves(h), info.enclosed_leaves, np.uint64)
        if build_hash:
            sc.hash_pool = _arr(lib.hds_hash_pool(h), info.hash_pool_top * 512, np.uint32)
            sc.hash_page_table = _arr(lib.hds_hash_page_table(h), info.hash_page_table_size, np.uint32)
            sc.hash_bucket_sizes = _arr(lib.hds_hash_bucket_sizes(h), lib.hds_hash_bucket_count(h), np.uint32)
            sc.hash_pool_top = info.hash_pool_top
            sc.hash_first_node_index = info.hash_first_node_index
        if build_colors:
            sc.weights = _arr(lib.hds_color_weights(h), info.n_weight_words, np.uint32)
            sc.blocks = _arr(lib.hds_color_blocks(h), info.n_blocks, np.uint64)
            sc.macro_blocks = _arr(lib.hds_color_macro_blocks(h), info.n_macro_words, np.uint64)
            sc.color_nodes = _arr(lib.hds_hash_color_nodes(h), info.n_color_nodes, np.uint32)
            sc.color_offsets = _arr(lib.hds_hash_color_offsets(h), info.n_color_offsets, np.uint64)
            if build_uncompressed:
                sc.uncompressed = _arr(lib.hds_color_uncompressed(h), info.n_uncompressed, np.uint32)
        sc.nodes_per_level = [int(v) for v in info.nodes_per_level[: levels - 1]]
        sc.build_seconds = info.build_seconds
        for (x, z) in (height_probes or []):
            sc.heights[(x, z)] = int(lib.hds_terrain_height(h, int(x), int(z)))
        sc.params = dict(levels=levels, footprint_log2=footprint_log2, seed=seed, n_spheres=n_spheres,
                         roughness=roughness, finest_cell_log2=finest_cell_log2)
        return sc
    finally:
        lib.hds_free(h)
