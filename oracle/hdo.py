"""ctypes bindings of the CPU oracle (oracle/hdo_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from hashdag_b200/ (the product fails loudly without its CUDA library instead).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhdo_oracle.so")

DAG_BASIC, DAG_HASH = 0, 1
COLORS_UNCOMPRESSED, COLORS_COMPRESSED, COLORS_ERRORS, COLORS_HASH = 0, 1, 2, 3


class Dag(C.Structure):
    _fields_ = [("kind", C.c_int32), ("levels", C.c_uint32), ("data", C.c_void_p), ("n_words", C.c_uint64),
                ("page_table", C.c_void_p), ("page_table_size", C.c_uint32), ("first_node_index", C.c_uint32)]


class ColorLeaf(C.Structure):
    _fields_ = [("weights", C.c_void_p), ("n_weights", C.c_uint64), ("blocks", C.c_void_p), ("n_blocks", C.c_uint64),
                ("macro_blocks", C.c_void_p), ("n_macro_words", C.c_uint64)]


class Colors(C.Structure):
    _fields_ = [("kind", C.c_int32), ("top_levels", C.c_uint32),
                ("enclosed_leaves", C.c_void_p), ("n_enclosed", C.c_uint64),
                ("leaf", ColorLeaf),
                ("uncompressed", C.c_void_p), ("n_uncompressed", C.c_uint64),
                ("color_nodes", C.c_void_p), ("n_color_nodes", C.c_uint64),
                ("color_offsets", C.c_void_p), ("n_color_offsets", C.c_uint64),
                ("unique_leaves", C.c_void_p), ("n_unique_leaves", C.c_uint64)]


class ToolInfo(C.Structure):
    _fields_ = [("tool", C.c_int32), ("position", C.c_uint32 * 3), ("radius", C.c_float),
                ("copy_source", C.c_uint32 * 3), ("copy_dest", C.c_uint32 * 3)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_word", "n_leaf", "n_page", "n_hit", "n_steps", "n_color_probe")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: run __graft_entry__.build()")
        l = C.CDLL(_LIB_PATH)
        d3 = C.POINTER(C.c_double)
        l.hdo_camera_params.argtypes = [d3, d3, d3, d3, C.c_uint32, C.c_uint32, C.c_uint32, d3, d3, d3, d3]
        l.hdo_trace_paths.argtypes = [C.POINTER(Dag), C.c_uint32, C.c_uint32, d3, d3, d3, d3, C.c_void_p, C.c_uint32, C.POINTER(Stats)]
        l.hdo_trace_colors.argtypes = [C.POINTER(Dag), C.POINTER(Colors), C.c_uint32, C.c_uint32, C.c_void_p, C.c_int32, C.c_uint32,
                                       C.POINTER(ToolInfo), C.c_int32, C.c_void_p, C.c_uint32, C.POINTER(Stats)]
        l.hdo_trace_shadows.argtypes = [C.POINTER(Dag), C.c_uint32, C.c_uint32, d3, d3, d3, d3, C.c_float, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_uint32, C.POINTER(Stats)]
        l.hdo_get_value.argtypes = [C.POINTER(Dag), C.c_uint32, C.c_uint32, C.c_uint32]
        l.hdo_decode_color.restype = C.c_uint32
        l.hdo_decode_color.argtypes = [C.POINTER(ColorLeaf), C.c_uint64]
        _lib = l
    return _lib


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data


def _d3(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def make_dag(scene, kind):
    """hdo_dag over a hashdag_b200.scene.Scene (keeps references to the numpy arrays alive)."""
    if kind == DAG_BASIC:
        d = Dag(DAG_BASIC, scene.levels, _p(scene.basic), scene.basic.size, None, 0, 0)
        d._keep = (scene.basic,)
    else:
        d = Dag(DAG_HASH, scene.levels, _p(scene.hash_pool), scene.hash_pool.size, _p(scene.hash_page_table),
                scene.hash_page_table.size, scene.hash_first_node_index)
        d._keep = (scene.hash_pool, scene.hash_page_table)
    return d


def make_leaf(scene):
    return ColorLeaf(_p(scene.weights), scene.weights.size, _p(scene.blocks), scene.blocks.size,
                     _p(scene.macro_blocks), scene.macro_blocks.size)


def make_colors(scene, kind):
    c = Colors()
    c.kind = kind
    c.top_levels = scene.top_levels
    c.enclosed_leaves, c.n_enclosed = _p(scene.enclosed_leaves), scene.enclosed_leaves.size
    c.leaf = make_leaf(scene)
    if scene.uncompressed is not None:
        c.uncompressed, c.n_uncompressed = _p(scene.uncompressed), scene.uncompressed.size
    if kind == COLORS_HASH:
        c.color_nodes, c.n_color_nodes = _p(scene.color_nodes), scene.color_nodes.size
        c.color_offsets, c.n_color_offsets = _p(scene.color_offsets), scene.color_offsets.size
    c._keep = scene
    return c


def camera_params(pos, rot, bounds_min, bounds_max, levels, width, height):
    """dag_tracer.cu:71-113 -> (cam, ray_min, ddx, ddy) as 4 tuples of 3 doubles."""
    outs = [(C.c_double * 3)() for _ in range(4)]
    lib().hdo_camera_params(_d3(pos), _d3(np.asarray(rot, dtype=np.float64).reshape(9)), _d3(bounds_min), _d3(bounds_max),
                            levels, width, height, *outs)
    return tuple(tuple(o) for o in outs)


def trace_paths(dag, width, height, params, n_threads=0):
    cam, rmin, ddx, ddy = params
    paths = np.zeros((height, width, 4), dtype=np.uint32)
    st = Stats()
    rc = lib().hdo_trace_paths(C.byref(dag), width, height, _d3(cam), _d3(rmin), _d3(ddx), _d3(ddy), paths.ctypes.data, n_threads, C.byref(st))
    assert rc == 0
    return paths, st.as_dict()


def trace_colors(dag, colors, paths, debug_colors=0, debug_index_level=0, tool=None, tool_overlay=False, n_threads=0):
    h, w = paths.shape[:2]
    out = np.zeros((h, w), dtype=np.uint32)
    st = Stats()
    paths = np.ascontiguousarray(paths)
    rc = lib().hdo_trace_colors(C.byref(dag), C.byref(colors), w, h, paths.ctypes.data, debug_colors, debug_index_level,
                                C.byref(tool) if tool is not None else None, int(tool_overlay), out.ctypes.data, n_threads, C.byref(st))
    assert rc == 0
    return out, st.as_dict()


def trace_shadows(dag, params, paths, colors_img, shadow_bias=1.0, fog_density=0.0, n_threads=0):
    cam, rmin, ddx, ddy = params
    h, w = paths.shape[:2]
    out = np.ascontiguousarray(colors_img).copy()
    st = Stats()
    paths = np.ascontiguousarray(paths)
    rc = lib().hdo_trace_shadows(C.byref(dag), w, h, _d3(cam), _d3(rmin), _d3(ddx), _d3(ddy), shadow_bias, fog_density,
                                 paths.ctypes.data, out.ctypes.data, n_threads, C.byref(st))
    assert rc == 0
    return out, st.as_dict()


def get_value(dag, x, y, z):
    return bool(lib().hdo_get_value(C.byref(dag), int(x), int(y), int(z)))
