"""Host-side mirror of the reference's tracer interface over the C ABI (include/hashdag_b200.h).

`DAGTracer` keeps the reference class's methods and argument meaning
(/root/reference/src/dag_tracer.h:9-43): resolve_paths / resolve_colors / resolve_shadows return the
kernel time in milliseconds, get_path returns the voxel under a pixel.  The DAG and colour classes
are the device-resident counterparts of BasicDAG (basic_dag.h:11-45), HashDAG (hash_dag.h:214-253),
BasicDAG*Colors (basic_dag.h:47-242) and HashDAGColors (hash_dag_colors.h:9-73); each produces the
bytes the reference would pass to its kernels by value.

PyTorch is used for device memory only.  There is no CPU path: without libhashdag_b200.so or
without a CUDA device every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

from .camera import CameraView, DAGInfo, trace_params

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HDT_LIB", os.path.join(_HERE, "libhashdag_b200.so"))  # HDT_LIB: A/B builds of the same ABI

DAG_BASIC, DAG_HASH, DAG_HASH_RESOLVED = 0, 1, 2
COLORS_UNCOMPRESSED, COLORS_COMPRESSED, COLORS_ERRORS, COLORS_HASH = 0, 1, 2, 3
UNIQUE_OFFSET = 0xFFFFFFFFFFFFFFFF
OPT_BEAMS, OPT_BEAM_MAX_VISITS, OPT_BEAM_PREFETCH, OPT_BEAM_SERIAL, OPT_EXCHANGE_FUSED, OPT_COLORS_RECORDED, OPT_EXCHANGE_TIMEOUT_MS, OPT_L2_PERSIST = 1, 2, 3, 4, 5, 6, 7, 8

# EDebugColors, tracer.h:7-17
DEBUG_NONE, DEBUG_INDEX, DEBUG_POSITION, DEBUG_COLOR_TREE, DEBUG_COLOR_BITS, DEBUG_MIN_COLOR, DEBUG_MAX_COLOR, DEBUG_WEIGHT = range(8)

EXPORTS = (
    "hdt_create", "hdt_destroy", "hdt_last_error", "hdt_set_partition", "hdt_set_option", "hdt_beam_stats", "hdt_pass_timeline", "hdt_resolve_paths", "hdt_resolve_colors",
    "hdt_resolve_shadows", "hdt_resolve_frame", "hdt_resolve_frame_async", "hdt_sync", "hdt_timer_begin", "hdt_timer_end",
    "hdt_count_hits", "hdt_get_path", "hdt_read_paths", "hdt_read_colors",
    "hdt_partition_buffers", "hdt_assemble_colors", "hdt_exchange_create", "hdt_exchange_open", "hdt_exchange_block", "hdt_exchange_attach", "hdt_exchange_block_bytes", "hdt_exchange_attach_host",
    "hdt_exchange_frame", "hdt_exchange_release", "hdt_set_stream", "hdt_apply_ranges", "hdt_apply_ranges_host", "hdt_hash_dag_resolve", "hdt_rebuild_color_leaf", "hdt_get_values", "hdt_is_empty", "hdt_tracker_create", "hdt_tracker_destroy", "hdt_tracker_bucket_count", "hdt_tracker_snapshot", "hdt_tracker_delta",
    "hdt_comm_unique_id", "hdt_comm_init", "hdt_comm_destroy", "hdt_replicate", "hdt_broadcast_dirty", "hdt_broadcast_ranges",
    "hdt_find_or_add", "hdt_launch_count", "hdt_recorded_color_passes", "hdt_version",
)
ERR_ARG, ERR_CAPACITY = 1, 4   # HDT_ERR_* of include/hashdag_b200.h the tests tell apart


class TracerError(RuntimeError):
    code = 0                     # the C ABI's return code


class Range(C.Structure):     # hdt_range
    _fields_ = [("dst_word", C.c_uint64), ("src_word", C.c_uint64), ("n_words", C.c_uint64)]


class DagDeltaPod(C.Structure):   # hdt_dag_delta
    _fields_ = [("first_node_index", C.c_uint32), ("pool_top", C.c_uint32),
                ("pool_ranges", C.c_void_p), ("n_pool_ranges", C.c_uint32), ("pool_payload", C.c_void_p), ("n_pool_payload", C.c_uint64),
                ("table_ranges", C.c_void_p), ("n_table_ranges", C.c_uint32), ("table_payload", C.c_void_p), ("n_table_payload", C.c_uint64)]


class ReplicaPod(C.Structure):    # hdt_replica
    _fields_ = [("pool", C.c_void_p), ("pool_capacity_words", C.c_uint64), ("page_table", C.c_void_p), ("page_table_size", C.c_uint32),
                ("first_node_index", C.c_uint32), ("pool_top", C.c_uint32), ("resolved_pool", C.c_void_p), ("prefix_pool", C.c_void_p)]


class HashTablePod(C.Structure):   # hdt_hash_table
    _fields_ = [("pool", C.c_void_p), ("pool_capacity_words", C.c_uint64), ("page_table", C.c_void_p), ("page_table_size", C.c_uint32),
                ("bucket_sizes", C.c_void_p), ("n_buckets", C.c_uint32), ("pool_top", C.c_uint32), ("levels", C.c_uint32)]


class ToolInfo(C.Structure):  # tracer.h:33-39
    _fields_ = [("tool", C.c_int32), ("position", C.c_uint32 * 3), ("radius", C.c_float),
                ("copy_source", C.c_uint32 * 3), ("copy_dest", C.c_uint32 * 3)]


_lib = None


def load_library():
    """Load libhashdag_b200.so; raises TracerError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TracerError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    d3 = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    lib.hdt_last_error.restype = C.c_char_p
    lib.hdt_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    lib.hdt_destroy.argtypes = [C.c_void_p]
    lib.hdt_set_partition.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.hdt_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.hdt_beam_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.hdt_pass_timeline.argtypes = [C.c_void_p, C.c_int, fp]
    lib.hdt_resolve_paths.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, d3, d3, d3, d3, fp]
    lib.hdt_resolve_colors.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_uint32,
                                       C.POINTER(ToolInfo), C.c_int, fp]
    lib.hdt_resolve_shadows.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, d3, d3, d3, d3, C.c_float, C.c_float, fp]
    lib.hdt_resolve_frame.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, d3, d3, d3, d3,
                                      C.c_float, C.c_float, C.c_int, C.c_void_p, fp]
    lib.hdt_resolve_frame_async.argtypes = lib.hdt_resolve_frame.argtypes[:-1]
    lib.hdt_sync.argtypes = [C.c_void_p]
    lib.hdt_timer_begin.argtypes = [C.c_void_p]
    lib.hdt_timer_end.argtypes = [C.c_void_p, fp]
    lib.hdt_count_hits.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.hdt_get_path.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.hdt_read_paths.argtypes = [C.c_void_p, C.c_void_p]
    lib.hdt_read_colors.argtypes = [C.c_void_p, C.c_void_p]
    lib.hdt_partition_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.hdt_assemble_colors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hdt_exchange_create.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.hdt_exchange_open.argtypes = [C.c_void_p, C.c_char_p]
    lib.hdt_exchange_block.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.hdt_exchange_attach.argtypes = [C.c_void_p, C.c_void_p]
    lib.hdt_exchange_block_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.hdt_exchange_attach_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.hdt_exchange_frame.argtypes = [C.c_void_p]
    lib.hdt_exchange_release.argtypes = [C.c_void_p]
    lib.hdt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.hdt_apply_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.hdt_apply_ranges_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    lib.hdt_hash_dag_resolve.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    lib.hdt_rebuild_color_leaf.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                           C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), fp]
    u3 = C.POINTER(C.c_uint32)
    lib.hdt_get_values.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, u3, u3, C.c_void_p, fp]
    lib.hdt_is_empty.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_uint32, u3, u3, C.POINTER(C.c_int), fp]
    lib.hdt_tracker_create.argtypes = [C.c_uint32, C.POINTER(C.c_void_p)]
    lib.hdt_tracker_destroy.argtypes = [C.c_void_p]
    lib.hdt_tracker_bucket_count.argtypes = [C.c_void_p]
    lib.hdt_tracker_bucket_count.restype = C.c_uint32
    lib.hdt_tracker_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.hdt_tracker_delta.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(DagDeltaPod)]
    lib.hdt_comm_unique_id.argtypes = [C.c_char_p]
    lib.hdt_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32]
    lib.hdt_comm_destroy.argtypes = [C.c_void_p]
    lib.hdt_replicate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
    lib.hdt_broadcast_dirty.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(DagDeltaPod), C.POINTER(ReplicaPod)]
    lib.hdt_broadcast_ranges.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    lib.hdt_find_or_add.argtypes = [C.c_void_p, C.POINTER(HashTablePod), C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    lib.hdt_launch_count.restype = C.c_uint64
    lib.hdt_launch_count.argtypes = [C.c_void_p]
    lib.hdt_recorded_color_passes.restype = C.c_uint64
    lib.hdt_recorded_color_passes.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        err = TracerError(f"hashdag_b200 error {rc}: {load_library().hdt_last_error().decode()}")
        err.code = rc
        raise err


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise TracerError("hashdag_b200 needs a CUDA device (there is no CPU path)")
    return torch


def _to_device(arr: np.ndarray, device):
    """numpy (uint32/uint64) -> device tensor with the same bytes."""
    torch = _torch()
    if arr is None or arr.size == 0:
        return None
    view = {4: np.int32, 8: np.int64}[arr.dtype.itemsize]
    return torch.from_numpy(np.ascontiguousarray(arr).view(view)).to(device)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _n(t):
    return 0 if t is None else t.numel()


def _array(t):        # StaticArray<T>
    return struct.pack("<QQ", _ptr(t), _n(t))


def _dyn_array(t):    # DynamicArray<T>
    return struct.pack("<QQQ", _ptr(t), _n(t), _n(t))


_NULL_ARRAY = struct.pack("<QQ", 0, 0)
_NULL_DYN = struct.pack("<QQQ", 0, 0, 0)


class BasicDAG:
    """Device-resident BasicDAG (basic_dag.h:11-45): one uint32 array, root at word 0."""
    kind = DAG_BASIC

    def __init__(self, data_tensor, levels):
        self.data, self.levels = data_tensor, levels

    @classmethod
    def from_scene(cls, scene, device="cuda:0"):
        return cls(_to_device(scene.basic, device), scene.levels)

    def pod(self) -> bytes:
        return _array(self.data)


class HashDAG:
    """Device-resident HashDAG read side (hash_dag.h:214-253, hash_table.h:818-826)."""
    kind = DAG_HASH

    def __init__(self, pool, page_table, pool_top, first_node_index, levels):
        self.pool, self.page_table = pool, page_table
        self.pool_top, self.first_node_index, self.levels = pool_top, first_node_index, levels

    @classmethod
    def from_scene(cls, scene, device="cuda:0"):
        if not scene.has_hash:
            raise TracerError("scene was built without a HashDAG")
        return cls(_to_device(scene.hash_pool, device), _to_device(scene.hash_page_table, device),
                   scene.hash_pool_top, scene.hash_first_node_index, scene.levels)

    def pod(self) -> bytes:
        return struct.pack("<IIQQII", _n(self.page_table), self.pool_top, _ptr(self.page_table), _ptr(self.pool), self.first_node_index, 0)


class ResolvedHashDAG:
    """A HashDAG + its resolved pool (hdt_resolved_hash_dag, 48 B): the copy of the pool in which every child pointer
    already holds the child's physical word index, and optionally the prefix pool (voxels under a node's earlier children,
    which lets trace_colors skip the DAG walk); csrc/hdt_resolve.cuh.  Accepted wherever a HashDAG is; same frames."""
    kind = DAG_HASH_RESOLVED

    def __init__(self, dag: HashDAG, resolved_pool, prefix_pool=None):
        self.dag, self.resolved_pool, self.prefix_pool, self.levels = dag, resolved_pool, prefix_pool, dag.levels

    def pod(self) -> bytes:
        return self.dag.pod() + struct.pack("<QQ", _ptr(self.resolved_pool), _ptr(self.prefix_pool))


class CompressedColorLeaf:
    """CompressedColorLeaf (vwsc.h:157-191), GPU arrays only."""

    def __init__(self, weights, blocks, macro_blocks, offset=UNIQUE_OFFSET):
        self.weights, self.blocks, self.macro_blocks, self.offset = weights, blocks, macro_blocks, offset

    @classmethod
    def from_scene(cls, scene, device="cuda:0"):
        return cls(_to_device(scene.weights, device), _to_device(scene.blocks, device), _to_device(scene.macro_blocks, device), UNIQUE_OFFSET)

    def pod(self) -> bytes:
        return struct.pack("<Q", self.offset) + _array(self.weights) + _array(self.blocks) + _array(self.macro_blocks) + 3 * _NULL_ARRAY


class _BasicColorsBase:
    def __init__(self, top_levels, enclosed_leaves):
        self.top_levels, self.enclosed_leaves = top_levels, enclosed_leaves

    def _base_pod(self) -> bytes:
        return struct.pack("<II", self.top_levels, 0) + _array(self.enclosed_leaves)


class BasicDAGCompressedColors(_BasicColorsBase):
    kind = COLORS_COMPRESSED

    def __init__(self, top_levels, enclosed_leaves, leaf):
        super().__init__(top_levels, enclosed_leaves)
        self.leaf = leaf

    @classmethod
    def from_scene(cls, scene, device="cuda:0"):
        return cls(scene.top_levels, _to_device(scene.enclosed_leaves, device), CompressedColorLeaf.from_scene(scene, device))

    def pod(self) -> bytes:
        return self._base_pod() + self.leaf.pod()


class BasicDAGUncompressedColors(_BasicColorsBase):
    kind = COLORS_UNCOMPRESSED

    def __init__(self, top_levels, enclosed_leaves, colors):
        super().__init__(top_levels, enclosed_leaves)
        self.colors = colors

    @classmethod
    def from_scene(cls, scene, device="cuda:0"):
        if scene.uncompressed is None:
            raise TracerError("scene was built without uncompressed colours")
        return cls(scene.top_levels, _to_device(scene.enclosed_leaves, device), _to_device(scene.uncompressed, device))

    def pod(self) -> bytes:
        return self._base_pod() + _array(self.colors)


class BasicDAGColorErrors:
    kind = COLORS_ERRORS

    def __init__(self, compressed: BasicDAGCompressedColors, uncompressed: BasicDAGUncompressedColors):
        self.compressed, self.uncompressed = compressed, uncompressed

    def pod(self) -> bytes:
        unused_leaf = struct.pack("<Q", 0) + 6 * _NULL_ARRAY + _NULL_ARRAY
        return unused_leaf + self.compressed.pod() + self.uncompressed.pod()


class HashDAGColors:
    kind = COLORS_HASH

    def __init__(self, nodes, offsets, main_leaf, leaves=None):
        self.nodes, self.offsets, self.main_leaf, self.leaves = nodes, offsets, main_leaf, leaves

    @classmethod
    def from_scene(cls, scene, device="cuda:0"):
        if not scene.has_hash_colors:
            raise TracerError("HashDAGColors need levels-2 > 10 (hash_dag_factory.cpp:113)")
        return cls(_to_device(scene.color_nodes, device), _to_device(scene.color_offsets, device), CompressedColorLeaf.from_scene(scene, device))

    def pod(self) -> bytes:
        leaves = _NULL_DYN if self.leaves is None else struct.pack("<QQQ", _ptr(self.leaves), _n(self.leaves) // 13, _n(self.leaves) // 13)
        return _dyn_array(self.nodes) + leaves + _dyn_array(self.offsets) + self.main_leaf.pod() + 3 * _NULL_DYN


class DirtyTracker:
    """hdt_dirty_tracker: an edit's dirty spans from the hash table's bucket fill counts (host code, no GPU needed).
    C++ twin of edits.delta_from_bucket_sizes."""

    def __init__(self, levels: int):
        self._lib = load_library()
        self._t = C.c_void_p()
        _check(self._lib.hdt_tracker_create(levels, C.byref(self._t)))
        self.n_buckets = int(self._lib.hdt_tracker_bucket_count(self._t))

    def close(self):
        if getattr(self, "_t", None):
            self._lib.hdt_tracker_destroy(self._t)
            self._t = None

    __del__ = close

    def snapshot(self, bucket_sizes: np.ndarray):
        b = np.ascontiguousarray(bucket_sizes, dtype=np.uint32)
        _check(self._lib.hdt_tracker_snapshot(self._t, b.ctypes.data, b.size))

    def delta_pod(self, bucket_sizes, cpu_pool, cpu_page_table, first_node_index, pool_top) -> DagDeltaPod:
        """-> hdt_dag_delta whose arrays live inside the tracker until its next call (zero-copy hand-over to hdt_broadcast_dirty)."""
        b = np.ascontiguousarray(bucket_sizes, dtype=np.uint32)
        assert cpu_pool.dtype == np.uint32 and cpu_page_table.dtype == np.uint32 and cpu_pool.flags.c_contiguous and cpu_page_table.flags.c_contiguous
        pod = DagDeltaPod()
        _check(self._lib.hdt_tracker_delta(self._t, b.ctypes.data, b.size, cpu_pool.ctypes.data, cpu_page_table.ctypes.data, int(first_node_index), int(pool_top), C.byref(pod)))
        return pod

    @staticmethod
    def arrays(pod: DagDeltaPod):
        """numpy copies of a delta's four arrays: (pool_ranges, pool_payload, table_ranges, table_payload)."""
        from .edits import RANGE_DTYPE

        def view(ptr, n, dtype):
            if not n:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (int(n) * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).copy()
        return (view(pod.pool_ranges, pod.n_pool_ranges, RANGE_DTYPE), view(pod.pool_payload, pod.n_pool_payload, np.uint32),
                view(pod.table_ranges, pod.n_table_ranges, RANGE_DTYPE), view(pod.table_payload, pod.n_table_payload, np.uint32))


def comm_unique_id() -> bytes:
    """hdt_comm_unique_id (rank 0): 128 bytes for the host to hand to every rank."""
    buf = C.create_string_buffer(128)
    _check(load_library().hdt_comm_unique_id(buf))
    return buf.raw


def delta_pod_from_arrays(first_node_index, pool_top, pool_ranges, pool_payload, table_ranges, table_payload):
    """hdt_dag_delta over numpy arrays (kept alive by the returned tuple's second element)."""
    keep = [np.ascontiguousarray(pool_ranges), np.ascontiguousarray(pool_payload, dtype=np.uint32),
            np.ascontiguousarray(table_ranges), np.ascontiguousarray(table_payload, dtype=np.uint32)]
    pod = DagDeltaPod(int(first_node_index), int(pool_top), keep[0].ctypes.data, len(keep[0]), keep[1].ctypes.data, keep[1].size,
                      keep[2].ctypes.data, len(keep[2]), keep[3].ctypes.data, keep[3].size)
    return pod, keep


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class DAGTracer:
    """Mirror of the reference's DAGTracer over libhashdag_b200.so.

    Width, height and DAG depth are runtime values (compile-time constants in the reference,
    typedefs.h:517,683-684).  `head_less` is accepted for signature compatibility; there is no GL path.
    """

    def __init__(self, head_less: bool = True, width: int = 1920, height: int = 1080, levels: int = 17, device: int = 0):
        _torch()
        self._lib = load_library()
        self.head_less, self.width, self.height, self.levels, self.device = head_less, width, height, levels, device
        self._ctx = C.c_void_p()
        _check(self._lib.hdt_create(width, height, levels, device, C.byref(self._ctx)))
        self._ms = C.c_float()
        self.comm_rank, self.comm_world = 0, 1

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.hdt_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _params(self, camera: CameraView, dag_info: DAGInfo):
        return trace_params(camera, dag_info, self.levels, self.width, self.height)

    # -- the three passes -----------------------------------------------------------------
    def resolve_paths(self, camera: CameraView, dag_info: DAGInfo, dag) -> float:
        cam, rmin, ddx, ddy = self._params(camera, dag_info)
        pod = dag.pod()
        _check(self._lib.hdt_resolve_paths(self._ctx, dag.kind, pod, len(pod), _d3(cam), _d3(rmin), _d3(ddx), _d3(ddy), C.byref(self._ms)))
        return self._ms.value

    def resolve_colors(self, dag, colors, debug_colors: int = DEBUG_NONE, debug_colors_index_level: int = 0,
                       tool_info: ToolInfo | None = None, tool_overlay: bool = False) -> float:
        pod, cpod = dag.pod(), colors.pod()
        _check(self._lib.hdt_resolve_colors(self._ctx, dag.kind, pod, len(pod), colors.kind, cpod, len(cpod), int(debug_colors),
                                            int(debug_colors_index_level), C.byref(tool_info) if tool_info is not None else None,
                                            int(tool_overlay), C.byref(self._ms)))
        return self._ms.value

    def resolve_shadows(self, camera: CameraView, dag_info: DAGInfo, dag, shadow_bias: float = 1.0, fog_density: float = 0.0) -> float:
        cam, rmin, ddx, ddy = self._params(camera, dag_info)
        pod = dag.pod()
        _check(self._lib.hdt_resolve_shadows(self._ctx, dag.kind, pod, len(pod), _d3(cam), _d3(rmin), _d3(ddx), _d3(ddy),
                                             float(shadow_bias), float(fog_density), C.byref(self._ms)))
        return self._ms.value

    def resolve_frame(self, camera, dag_info, dag, colors, shadow_bias=1.0, fog_density=0.0, with_shadows=True, host_colors=None):
        """paths + colours + shadows with one synchronisation; returns the three kernel times (ms)."""
        cam, rmin, ddx, ddy = self._params(camera, dag_info)
        pod, cpod = dag.pod(), colors.pod()
        ms = (C.c_float * 3)()
        host_ptr = None
        if host_colors is not None:
            host_ptr = host_colors.data_ptr() if hasattr(host_colors, "data_ptr") else host_colors.ctypes.data
        _check(self._lib.hdt_resolve_frame(self._ctx, dag.kind, pod, len(pod), colors.kind, cpod, len(cpod), _d3(cam), _d3(rmin), _d3(ddx),
                                           _d3(ddy), float(shadow_bias), float(fog_density), int(with_shadows), host_ptr, ms))
        return tuple(ms)

    def enqueue_frame(self, params, dag_pod, dag_kind, colors_pod, colors_kind, shadow_bias=1.0, fog_density=0.0, with_shadows=True, host_ptr=None):
        """Asynchronous frame from precomputed trace params and POD bytes (bench inner loop)."""
        cam, rmin, ddx, ddy = params
        _check(self._lib.hdt_resolve_frame_async(self._ctx, dag_kind, dag_pod, len(dag_pod), colors_kind, colors_pod, len(colors_pod),
                                                 _d3(cam), _d3(rmin), _d3(ddx), _d3(ddy), float(shadow_bias), float(fog_density),
                                                 int(with_shadows), host_ptr))

    def sync(self):
        _check(self._lib.hdt_sync(self._ctx))

    def timer_begin(self):
        _check(self._lib.hdt_timer_begin(self._ctx))

    def timer_end(self) -> float:
        _check(self._lib.hdt_timer_end(self._ctx, C.byref(self._ms)))
        return self._ms.value

    def count_hits(self) -> int:
        n = C.c_uint64()
        _check(self._lib.hdt_count_hits(self._ctx, C.byref(n)))
        return int(n.value)

    def get_path(self, x: int, y: int):
        out = (C.c_uint32 * 3)()
        _check(self._lib.hdt_get_path(self._ctx, x, y, out))
        return tuple(out)

    # -- read-back ------------------------------------------------------------------------
    def read_paths(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.uint32)
        _check(self._lib.hdt_read_paths(self._ctx, out.ctypes.data))
        return out

    def read_colors(self) -> np.ndarray:
        out = np.empty((self.height, self.width), dtype=np.uint32)
        _check(self._lib.hdt_read_colors(self._ctx, out.ctypes.data))
        return out

    # -- multi-GPU ------------------------------------------------------------------------
    def set_partition(self, rank: int, world: int, tile_log2: int = 6):
        _check(self._lib.hdt_set_partition(self._ctx, rank, world, tile_log2))

    def partition_buffers(self):
        p, c, n, m = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(self._lib.hdt_partition_buffers(self._ctx, C.byref(p), C.byref(c), C.byref(n), C.byref(m)))
        return p.value, c.value, n.value, m.value

    def assemble_colors(self, gathered_tensor, frame_tensor=None):
        _check(self._lib.hdt_assemble_colors(self._ctx, gathered_tensor.data_ptr(), 0 if frame_tensor is None else frame_tensor.data_ptr()))

    # -- framebuffer exchange over peer memory (csrc/hdt_exchange.cuh) ---------------------------
    def exchange_create(self):
        """Rank 0: allocate the shared row-major frame -> (64-byte CUDA IPC handle for the other processes, frame device pointer)."""
        h = C.create_string_buffer(64)
        p = C.c_void_p()
        _check(self._lib.hdt_exchange_create(self._ctx, h, C.byref(p)))
        return h.raw, p.value

    def exchange_open(self, ipc_handle: bytes):
        """Other ranks (other processes): map rank 0's frame."""
        _check(self._lib.hdt_exchange_open(self._ctx, C.create_string_buffer(ipc_handle, 64)))

    def exchange_block(self) -> int:
        p = C.c_void_p()
        _check(self._lib.hdt_exchange_block(self._ctx, C.byref(p)))
        return p.value

    def exchange_attach(self, block_ptr: int):
        """Other ranks living in rank 0's process: attach by pointer (CUDA IPC cannot open a handle in its own process)."""
        _check(self._lib.hdt_exchange_attach(self._ctx, block_ptr))

    def exchange_block_bytes(self) -> int:
        n = C.c_uint64()
        _check(self._lib.hdt_exchange_block_bytes(self._ctx, C.byref(n)))
        return int(n.value)

    def exchange_attach_host(self, host_address: int, n_bytes: int):
        """Every rank (rank 0 too): use zero-initialised shared HOST memory as the exchange block; tiles then travel
        straight into host memory over each rank's own PCIe link."""
        _check(self._lib.hdt_exchange_attach_host(self._ctx, C.c_void_p(host_address), n_bytes))

    def exchange_frame(self):
        """Every rank, after the frame's passes: store the owned tiles into rank 0's frame (asynchronous)."""
        _check(self._lib.hdt_exchange_frame(self._ctx))

    def exchange_release(self):
        """Rank 0, once the work that reads the frame is queued: let the other ranks overwrite it."""
        _check(self._lib.hdt_exchange_release(self._ctx))

    def apply_ranges(self, dst_tensor, payload_tensor, ranges_tensor, n_ranges: int):
        """dst[r.dst_word + i] = payload[r.src_word + i] for every hdt_range r (edit-dirtied spans, see edits.py)."""
        _check(self._lib.hdt_apply_ranges(self._ctx, dst_tensor.data_ptr(), payload_tensor.data_ptr(), ranges_tensor.data_ptr(), n_ranges))

    def apply_ranges_host(self, dst_tensor, payload: np.ndarray, ranges: np.ndarray):
        """hdt_apply_ranges with host arrays (uint32 payload, hdt_range records): staged and applied in stream order, no sync."""
        if len(ranges) == 0:
            return
        payload = np.ascontiguousarray(payload)
        ranges = np.ascontiguousarray(ranges)
        assert payload.dtype.itemsize == 4 and ranges.dtype.itemsize == 24
        _check(self._lib.hdt_apply_ranges_host(self._ctx, dst_tensor.data_ptr(), payload.ctypes.data, payload.size, ranges.ctypes.data, len(ranges)))

    def resolve_hash_dag(self, dag: "HashDAG", resolved_pool=None, ranges: "np.ndarray | None" = None, prefix_pool=None, prefix: bool = True) -> "ResolvedHashDAG":
        """hdt_hash_dag_resolve: fill (ranges None) or refresh (the pages the edit's pool spans touch) the resolved pool and,
        unless prefix=False, the prefix pool.  Asynchronous on the tracer's stream."""
        torch = _torch()
        if resolved_pool is None:
            resolved_pool = torch.empty(dag.pool.numel(), dtype=torch.int32, device=dag.pool.device)
        if prefix_pool is None and prefix:
            prefix_pool = torch.empty(dag.pool.numel(), dtype=torch.int32, device=dag.pool.device)
        pod = dag.pod()
        rptr, n = (None, 0)
        if ranges is not None:
            ranges = np.ascontiguousarray(ranges)
            assert ranges.dtype.itemsize == 24
            rptr, n = ranges.ctypes.data, len(ranges)
        _check(self._lib.hdt_hash_dag_resolve(self._ctx, pod, len(pod), resolved_pool.data_ptr(), _ptr(prefix_pool), resolved_pool.numel(), rptr, n))
        return ResolvedHashDAG(dag, resolved_pool, prefix_pool)

    # -- multi-GPU host layer behind the C ABI (csrc/hdt_multi.cuh) -------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """hdt_comm_init: NCCL communicator of this context (collective); world == 1 needs no NCCL."""
        _check(self._lib.hdt_comm_init(self._ctx, C.create_string_buffer(unique_id, 128), rank, world))
        self.comm_rank, self.comm_world = rank, world

    def replicate(self, tensor, root: int = 0):
        """hdt_replicate: broadcast a device tensor's bytes from `root` (collective, tracer's stream)."""
        _check(self._lib.hdt_replicate(self._ctx, tensor.data_ptr(), tensor.numel() * tensor.element_size(), root))

    def broadcast_dirty(self, replica: ReplicaPod, delta: "DagDeltaPod | None", root: int = 0):
        """hdt_broadcast_dirty: one edit on every rank's replica (pool, page table, resolved / prefix pools); updates replica's root / pool top."""
        _check(self._lib.hdt_broadcast_dirty(self._ctx, root, C.byref(delta) if delta is not None else None, C.byref(replica)))

    def broadcast_ranges(self, dst_tensor, payload: "np.ndarray | None", ranges: "np.ndarray | None", root: int = 0):
        """hdt_broadcast_ranges: spans of any replicated uint32 array (the colour tree)."""
        if payload is None:
            _check(self._lib.hdt_broadcast_ranges(self._ctx, root, dst_tensor.data_ptr(), dst_tensor.numel(), None, 0, None, 0))
            return
        payload = np.ascontiguousarray(payload, dtype=np.uint32)
        ranges = np.ascontiguousarray(ranges)
        _check(self._lib.hdt_broadcast_ranges(self._ctx, root, dst_tensor.data_ptr(), dst_tensor.numel(), payload.ctypes.data, payload.size, ranges.ctypes.data, len(ranges)))

    # -- GPU batch insert into the hash table (HashTable::find_or_add_*, hash_table.h:470-560) -----------
    def find_or_add(self, table: HashTablePod, level: int, nodes, leaves: bool = False):
        """hdt_find_or_add: `nodes` = list of uint32 arrays (interior: [header, child pointers...]; leaf: 2 words) of ONE level in
        insertion order.  -> (virtual pointers uint32[n], nodes added, pages opened); `table` (device arrays) is updated in place."""
        torch = _torch()
        dev = f"cuda:{self.device}"
        sizes = np.array([len(w) for w in nodes], dtype=np.int64)
        offsets = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
        words = np.concatenate([np.asarray(w, dtype=np.uint32) for w in nodes]) if len(nodes) else np.zeros(1, np.uint32)
        ptrs, added, pages = self.find_or_add_device(table, level, _to_device(words, dev), torch.from_numpy(offsets).to(dev), len(nodes), leaves)
        return ptrs.cpu().numpy().view(np.uint32), added, pages

    def find_or_add_device(self, table: HashTablePod, level: int, words, offsets, n_nodes: int, leaves: bool = False):
        """hdt_find_or_add on device tensors: node i = words[offsets[i]:offsets[i+1]] (int32 words, int64 offsets[n+1]).
        -> (int32 tensor of virtual pointers, nodes added, pages opened)."""
        torch = _torch()
        ptrs = torch.empty(max(1, n_nodes), dtype=torch.int32, device=words.device)
        counts = (C.c_uint32 * 2)()
        torch.cuda.synchronize()
        _check(self._lib.hdt_find_or_add(self._ctx, C.byref(table), int(level), int(bool(leaves)), _ptr(words), offsets.data_ptr(), int(n_nodes), ptrs.data_ptr(), counts))
        return ptrs[:n_nodes], int(counts[0]), int(counts[1])

    def rebuild_color_leaf(self, ops: np.ndarray, old_leaf: "CompressedColorLeaf | None" = None, device=None):
        """Re-encode a colour leaf on the GPU from an op list (color_leaf.OP_DTYPE records = hdt_color_op), see
        hashdag_b200/color_leaf.py.  -> (CompressedColorLeaf with device tensors, kernel ms)."""
        torch = _torch()
        device = device or f"cuda:{self.device}"
        ops = np.ascontiguousarray(ops)
        assert ops.dtype.itemsize == 32, "ops must be hdt_color_op records"
        n = int(ops["count"].sum()) if ops.size else 0
        counts = (C.c_uint64 * 4)()
        pod = old_leaf.pod() if old_leaf is not None else None
        if n == 0:
            return CompressedColorLeaf(None, None, None, UNIQUE_OFFSET), 0.0
        # worst case: every colour its own block with 4-bit weights (12.5 B / colour); trimmed below
        blocks = torch.empty(n, dtype=torch.int64, device=device)
        macro = torch.empty(2 * ((n + 16383) // 16384), dtype=torch.int64, device=device)
        weights = torch.empty((4 * n + 31) // 32, dtype=torch.int32, device=device)
        _check(self._lib.hdt_rebuild_color_leaf(self._ctx, pod, len(pod) if pod else 0, ops.ctypes.data, ops.size, weights.data_ptr(), weights.numel(),
                                                blocks.data_ptr(), blocks.numel(), macro.data_ptr(), macro.numel(), counts, C.byref(self._ms)))
        assert counts[0] == n
        nw, nb, nm = int(counts[1]), int(counts[2]), int(counts[3])
        leaf = CompressedColorLeaf(weights[:nw].clone() if nw else None, blocks[:nb].clone(), macro[:nm].clone(), UNIQUE_OFFSET)
        return leaf, self._ms.value

    # -- region queries of the copy tool (DAGUtils, dag_utils.h:175-411) -------------------------
    def get_values(self, dag, start, size):
        """DAGUtils::get_values: -> (uint8 device tensor [size.z, size.y, size.x], kernel ms)."""
        torch = _torch()
        out = torch.empty((int(size[2]), int(size[1]), int(size[0])), dtype=torch.uint8, device=f"cuda:{self.device}")
        pod = dag.pod()
        u3 = lambda v: (C.c_uint32 * 3)(*[int(x) for x in v])
        _check(self._lib.hdt_get_values(self._ctx, dag.kind, pod, len(pod), u3(start), u3(size), out.data_ptr(), C.byref(self._ms)))
        return out, self._ms.value

    def is_empty(self, dag, max_level: int, start, size) -> bool:
        """DAGUtils::is_empty(dag, maxLevel, start, size)."""
        pod = dag.pod()
        u3 = lambda v: (C.c_uint32 * 3)(*[int(x) for x in v])
        e = C.c_int()
        _check(self._lib.hdt_is_empty(self._ctx, dag.kind, pod, len(pod), int(max_level), u3(start), u3(size), C.byref(e), C.byref(self._ms)))
        return bool(e.value)

    def set_stream(self, cuda_stream_handle):
        """Enqueue on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream); None = own stream."""
        _check(self._lib.hdt_set_stream(self._ctx, cuda_stream_handle))

    def set_option(self, option: int, value: int):
        """Tuning switches that never change results (OPT_BEAMS: per-tile beam pre-pass on/off)."""
        _check(self._lib.hdt_set_option(self._ctx, option, value))

    def beam_stats(self) -> dict:
        out = (C.c_uint64 * 5)()
        _check(self._lib.hdt_beam_stats(self._ctx, out))
        return {"root": int(out[0]), "resume": int(out[1]), "hit": int(out[2]), "miss": int(out[3]), "level_sum": int(out[4])}

    def pass_timeline(self, which: int):
        """(setup end, beam kernel end, per-ray kernel end) of the last paths (0) / shadows (1) pass, ms from its enqueue."""
        out = (C.c_float * 3)()
        _check(self._lib.hdt_pass_timeline(self._ctx, which, out))
        return tuple(out)

    def launch_count(self) -> int:
        return int(self._lib.hdt_launch_count(self._ctx))

    def recorded_color_passes(self) -> int:
        """Colour passes that read the paths pass' ancestor records instead of walking the DAG (OPT_COLORS_RECORDED)."""
        return int(self._lib.hdt_recorded_color_passes(self._ctx))
