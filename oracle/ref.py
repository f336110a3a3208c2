"""ctypes front-end of oracle/_ref/libhashdag_ref_d*_*.so -- the UNMODIFIED reference tracer
(/root/reference/src/{tracer,dag_tracer}.cu + its HashDAG factory) compiled for sm_100a by
oracle/build_ref.py.  TEST INFRASTRUCTURE: needs a GPU; used by the `-m gpu` parity tests, by
tests/golden/make_golden.py and by `bench.py --impl reference-cuda`.

The reference keeps its scene in globals, so one process can hold one scene per library variant.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(depth, width, height, overlay=False):
    return os.path.join(_HERE, "_ref", f"libhashdag_ref_d{depth}_{width}x{height}{'_overlay' if overlay else ''}.so")


def available(depth, width, height, overlay=False):
    return os.path.exists(lib_path(depth, width, height, overlay))


def _d(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    return (C.c_double * v.size)(*v.tolist())


class RefTracer:
    def __init__(self, depth, width, height, device=0, overlay=False):
        path = lib_path(depth, width, height, overlay)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        self.lib = l = C.CDLL(path)
        assert (l.ref_levels(), l.ref_width(), l.ref_height()) == (depth, width, height)
        self.depth, self.width, self.height = depth, width, height
        d = C.POINTER(C.c_double)
        l.ref_set_basic_dag.argtypes = [C.c_void_p, C.c_uint64]
        l.ref_set_compressed_colors.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        l.ref_set_uncompressed_colors.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        l.ref_build_hash_dag.argtypes = [C.c_uint32, C.c_int]
        l.ref_hash_info.argtypes = [C.POINTER(C.c_uint32)] * 3
        l.ref_hash_copy.argtypes = [C.c_void_p, C.c_void_p]
        l.ref_hash_colors_info.argtypes = [C.POINTER(C.c_uint64)] * 2
        l.ref_hash_colors_copy.argtypes = [C.c_void_p, C.c_void_p]
        l.ref_resolve_paths.restype = C.c_float
        l.ref_resolve_paths.argtypes = [C.c_int, d, d, d, d]
        l.ref_resolve_colors.restype = C.c_float
        l.ref_resolve_colors.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]
        l.ref_resolve_shadows.restype = C.c_float
        l.ref_resolve_shadows.argtypes = [C.c_int, d, d, d, d, C.c_float, C.c_float]
        l.ref_read_paths.argtypes = [C.c_void_p]
        l.ref_read_colors.argtypes = [C.c_void_p]
        l.ref_write_colors.argtypes = [C.c_void_p]
        if hasattr(l, "ref_edit_sphere"):
            u3 = C.POINTER(C.c_uint32)
            l.ref_edit_sphere.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
            l.ref_color_leaf_count.restype = C.c_uint64
            l.ref_color_leaf_info.argtypes = [C.c_uint64] + [C.POINTER(C.c_uint64)] * 3
            l.ref_color_leaf_copy.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
            l.ref_resolve_colors_tool.restype = C.c_float
            l.ref_resolve_colors_tool.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, u3, C.c_float, u3, u3]
        if hasattr(l, "ref_hash_pointers"):
            l.ref_hash_pointers.argtypes = [C.POINTER(C.c_uint64)] * 3 + [C.POINTER(C.c_uint32)]
        if hasattr(l, "ref_last_edit_ms"):
            l.ref_last_edit_ms.argtypes = [C.POINTER(C.c_double)]
            l.ref_last_edit_ms.restype = None
        if hasattr(l, "ref_get_values"):
            u3 = C.POINTER(C.c_uint32)
            l.ref_get_values.argtypes = [u3, u3, C.c_void_p]
            l.ref_is_empty.argtypes = [C.c_uint32, u3, u3]
        if hasattr(l, "ref_dropin_tick"):   # oracle/ref_harness/ref_dropin.cu: the product's C++ shim driven with the reference's own types
            l.ref_dropin_tick.argtypes = [C.c_int, d, d, d, d, C.c_int, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_uint32, C.c_uint32,
                                          C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
            l.ref_dropin_read_paths.argtypes = [C.c_void_p]
            l.ref_dropin_read_colors.argtypes = [C.c_void_p]
        self._dropin = False
        if l.ref_init(device) != 0:
            raise RuntimeError("ref_init failed (no CUDA device?)")
        self.scene = None
        atexit.register(self.close)

    def close(self):
        """Release the reference's allocations (its leak check aborts the process otherwise)."""
        if self.lib is not None:
            if self._dropin:
                self.lib.ref_dropin_shutdown()
            self.lib.ref_shutdown()
            self.lib = None

    # -- the drop-in proof (ref_dropin.cu): Engine::tick's tracer calls made on hashdag_b200::DAGTracer with the reference's structs
    def dropin_tick(self, current_dag, camera, info, debug_colors=0, debug_level=0, shadows=True, shadow_bias=1.0, fog_density=0.0, pos=(0, 0)):
        """current_dag: engine.h's EDag (0 uncompressed, 1 compressed, 2 colour errors, 3 HashDAG).  -> ((paths, colours, shadows) ms, config.path)"""
        if not self._dropin:
            assert self.lib.ref_dropin_init(0) == 0
            self._dropin = True
        times, path = (C.c_double * 3)(), (C.c_uint32 * 3)()
        assert self.lib.ref_dropin_tick(int(current_dag), _d(camera.position), _d(camera.rotation), _d(info.bounds_min), _d(info.bounds_max), int(debug_colors),
                                        int(debug_level), int(bool(shadows)), shadow_bias, fog_density, int(pos[0]), int(pos[1]), times, path) == 0
        return tuple(times), tuple(path)

    def dropin_read_paths(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint32)
        assert self.lib.ref_dropin_read_paths(out.ctypes.data) == 0
        return out

    def dropin_read_colors(self):
        out = np.empty((self.height, self.width), dtype=np.uint32)
        assert self.lib.ref_dropin_read_colors(out.ctypes.data) == 0
        return out

    def load_scene(self, scene, with_hash=True, with_colors=True, with_uncompressed=False, extra_pool_pages=4096):
        assert scene.levels == self.depth
        l = self.lib
        self.scene = scene
        l.ref_set_basic_dag(scene.basic.ctypes.data, scene.basic.size)
        if with_colors:
            l.ref_set_compressed_colors(scene.top_levels, scene.enclosed_leaves.ctypes.data, scene.enclosed_leaves.size,
                                        scene.weights.ctypes.data, scene.weights.size, scene.blocks.ctypes.data, scene.blocks.size,
                                        scene.macro_blocks.ctypes.data, scene.macro_blocks.size)
        if with_uncompressed:
            l.ref_set_uncompressed_colors(scene.top_levels, scene.enclosed_leaves.ctypes.data, scene.enclosed_leaves.size,
                                          scene.uncompressed.ctypes.data, scene.uncompressed.size)
        if with_hash:
            hash_colors = with_colors and scene.levels - 2 > 10
            self._pool_pages = scene.hash_pool_top + int(extra_pool_pages)
            l.ref_build_hash_dag(self._pool_pages, int(hash_colors))

    def hash_dag(self):
        """(pool, page_table, first_node_index, pool_top) as built by the reference's own factory."""
        f, t, s = C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert self.lib.ref_hash_info(C.byref(f), C.byref(t), C.byref(s)) == 0
        pool = np.empty(t.value * 512, dtype=np.uint32)
        pt = np.empty(s.value, dtype=np.uint32)
        self.lib.ref_hash_copy(pool.ctypes.data, pt.ctypes.data)
        return pool, pt, f.value, t.value

    def hash_views(self):
        """Zero-copy numpy views of the reference's HOST arrays: (pool[:capacity], page_table, bucket_sizes).  They alias
        live memory: the next edit changes them in place."""
        p, t, b, n = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        assert self.lib.ref_hash_pointers(C.byref(p), C.byref(t), C.byref(b), C.byref(n)) == 0
        f, top, size = C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert self.lib.ref_hash_info(C.byref(f), C.byref(top), C.byref(size)) == 0
        view = lambda addr, count: np.ctypeslib.as_array(C.cast(addr.value, C.POINTER(C.c_uint32)), shape=(count,))
        return view(p, self._pool_pages * 512), view(t, size.value), view(b, n.value)

    def hash_info(self):
        """(first_node_index, pool_top)"""
        f, top, size = C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert self.lib.ref_hash_info(C.byref(f), C.byref(top), C.byref(size)) == 0
        return f.value, top.value

    def hash_colors(self):
        a, b = C.c_uint64(), C.c_uint64()
        assert self.lib.ref_hash_colors_info(C.byref(a), C.byref(b)) == 0
        nodes = np.empty(a.value, dtype=np.uint32)
        offs = np.empty(b.value, dtype=np.uint64)
        self.lib.ref_hash_colors_copy(nodes.ctypes.data, offs.ctypes.data)
        return nodes, offs

    def edit_sphere(self, center, radius, adding):
        """The reference's own CPU edit (SphereEditor<adding>, hash_dag_editors.h:273-313) + upload_to_gpu."""
        assert self.lib.ref_edit_sphere(float(center[0]), float(center[1]), float(center[2]), float(radius), int(bool(adding))) == 0

    def find_or_add(self, level, nodes, leaves=False):
        """The reference's HashTable::find_or_add_* (hash_table.h:470-560) for `nodes` (uint32 arrays), one after the other."""
        offsets = np.concatenate(([0], np.cumsum([len(w) for w in nodes]))).astype(np.uint64)
        words = np.ascontiguousarray(np.concatenate(nodes), dtype=np.uint32)
        ptrs = np.empty(len(nodes), dtype=np.uint32)
        assert self.lib.ref_find_or_add(int(level), int(bool(leaves)), words.ctypes.data, offsets.ctypes.data, len(nodes), ptrs.ctypes.data) == 0
        return ptrs

    def last_edit_ms(self):
        """(HashDAG::edit_threads, HashTable::upload_to_gpu) host milliseconds of the last edit_sphere."""
        out = (C.c_double * 2)()
        self.lib.ref_last_edit_ms(out)
        return float(out[0]), float(out[1])

    def color_leaves(self):
        """Unique colour leaves created by edits: list of (weights, blocks, macro_blocks)."""
        out = []
        for i in range(int(self.lib.ref_color_leaf_count())):
            a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
            assert self.lib.ref_color_leaf_info(i, C.byref(a), C.byref(b), C.byref(c)) == 0
            w, bl, m = np.zeros(a.value, np.uint32), np.zeros(b.value, np.uint64), np.zeros(c.value, np.uint64)
            assert self.lib.ref_color_leaf_copy(i, w.ctypes.data, bl.ctypes.data, m.ctypes.data) == 0
            out.append((w, bl, m))
        return out

    def get_values(self, start, size):
        """DAGUtils::get_values<5> on the reference's HashDAG (host walk) -> uint8[size.z, size.y, size.x]."""
        u3 = lambda v: (C.c_uint32 * 3)(*[int(x) for x in v])
        out = np.full((int(size[2]), int(size[1]), int(size[0])), 0xAA, dtype=np.uint8)
        assert self.lib.ref_get_values(u3(start), u3(size), out.ctypes.data) == 0
        return out

    def is_empty(self, max_level, start, size):
        u3 = lambda v: (C.c_uint32 * 3)(*[int(x) for x in v])
        r = self.lib.ref_is_empty(int(max_level), u3(start), u3(size))
        assert r in (0, 1)
        return bool(r)

    def tool_overlay_compiled(self):
        return bool(self.lib.ref_tool_overlay_compiled())

    def resolve_colors_tool(self, dag_kind, colors_kind, tool_kind, position, radius, copy_source=(0, 0, 0), copy_dest=(0, 0, 0), debug_colors=0, debug_level=0):
        u3 = lambda v: (C.c_uint32 * 3)(*[int(x) for x in v])
        return self.lib.ref_resolve_colors_tool(dag_kind, colors_kind, debug_colors, debug_level, int(tool_kind), u3(position), float(radius), u3(copy_source), u3(copy_dest))

    def resolve_paths(self, dag_kind, camera, info):
        return self.lib.ref_resolve_paths(dag_kind, _d(camera.position), _d(camera.rotation), _d(info.bounds_min), _d(info.bounds_max))

    def resolve_colors(self, dag_kind, colors_kind, debug_colors=0, debug_level=0):
        return self.lib.ref_resolve_colors(dag_kind, colors_kind, debug_colors, debug_level)

    def resolve_shadows(self, dag_kind, camera, info, shadow_bias=1.0, fog_density=0.0):
        return self.lib.ref_resolve_shadows(dag_kind, _d(camera.position), _d(camera.rotation), _d(info.bounds_min), _d(info.bounds_max),
                                            shadow_bias, fog_density)

    def read_paths(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint32)
        assert self.lib.ref_read_paths(out.ctypes.data) == 0
        return out

    def read_colors(self):
        out = np.empty((self.height, self.width), dtype=np.uint32)
        assert self.lib.ref_read_colors(out.ctypes.data) == 0
        return out

    def write_colors(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint32)
        assert self.lib.ref_write_colors(img.ctypes.data) == 0


_SHARED = {}


def shared(scene, width, height, key, **load_kwargs):
    """One RefTracer per library variant and process (the reference keeps its scene in globals and
    cannot load a second one); `key` names the scene so a different one is refused, not mixed up."""
    k = (scene.levels, width, height)
    if k in _SHARED:
        rt, have = _SHARED[k]
        if have != key:
            raise RuntimeError(f"reference variant d{k[0]}_{k[1]}x{k[2]} already holds scene {have!r}")
        return rt
    rt = RefTracer(scene.levels, width, height)
    rt.load_scene(scene, **load_kwargs)
    _SHARED[k] = (rt, key)
    return rt
