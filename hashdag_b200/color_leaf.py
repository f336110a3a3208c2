"""Host mirror of the reference's ColorLeafBuilder for the GPU colour-leaf rebuild.

The reference re-encodes a colour leaf while the editor walks it (hash_dag_edits.h:340-540): untouched subtrees
call `leaf.copy_colors(builder, start, count)` (vwsc.h:416-542), entirely-full nodes
`builder.add_large_single_color(color, n)` (vwsc.h:614-641), single voxels `builder.add(color)` (vwsc.h:582-613),
and `builder.build(leaf)` (vwsc.h:643-685) packs blocks / weights / macro blocks on the host and uploads them.

Here the builder keeps the same four methods but only RECORDS what it is told (one 32-byte hdt_color_op per call,
consecutive equal calls merged); `build()` hands the list to `hdt_rebuild_color_leaf`, which decodes, run-length
encodes and bit-packs on the GPU (csrc/hdt_color_leaf.cuh) and leaves the three arrays in device memory, where
trace_colors reads them.  No CPU path: build() needs the tracer's CUDA context.
"""
from __future__ import annotations

import numpy as np

OP_COPY, OP_FILL = 0, 1
OP_DTYPE = np.dtype([("src_start", "<u8"), ("count", "<u8"), ("kind", "<u4"), ("bits_per_weight", "<u4"), ("color_bits", "<u4"), ("weight", "<u4")])
assert OP_DTYPE.itemsize == 32   # hdt_color_op


def float3_to_rgb101210(rgb) -> int:
    """ColorUtils::float3_to_rgb101210 (color_utils.h:52-60): float32 products, truncating casts."""
    f = np.float32
    r, g, b = (min(max(f(c), f(0.0)), f(1.0)) for c in rgb)
    return int(f(r) * f(1023.0)) | (int(f(g) * f(4095.0)) << 10) | (int(f(b) * f(1023.0)) << 22)


class ColorLeafBuilder:
    """Same calls as the reference's ColorLeafBuilder + CompressedColorLeaf::copy_colors; records ops."""

    def __init__(self):
        self._ops: list[tuple] = []
        self._n = 0

    def get_color_index(self) -> int:           # vwsc.h:687-690
        return self._n

    def _push(self, src, count, kind, bpw, bits, weight):
        if count <= 0:
            return
        if self._ops:
            s, c, k, b, cb, w = self._ops[-1]
            if k == kind and ((kind == OP_COPY and s + c == src) or (kind == OP_FILL and (b, cb, w) == (bpw, bits, weight))):
                self._ops[-1] = (s, c + count, k, b, cb, w)
                self._n += count
                return
        self._ops.append((src, count, kind, bpw, bits, weight))
        self._n += count

    def add(self, color_bits: int, weight: int = 0, bits_per_weight: int = 0):
        if not (0 <= bits_per_weight <= 4) or weight >> bits_per_weight:
            raise ValueError("weight does not fit bits_per_weight")
        self._push(0, 1, OP_FILL, int(bits_per_weight), int(color_bits) & 0xFFFFFFFF, int(weight))

    def add_large_single_color(self, single_color_rgb, num_voxels: int):
        self._push(0, int(num_voxels), OP_FILL, 0, float3_to_rgb101210(single_color_rgb), 0)

    def copy_colors(self, start: int, count: int):
        """`start` is the editor's oldLeavesCount: relative to the old leaf's view (its offset is added on the device)."""
        self._push(int(start), int(count), OP_COPY, 0, 0, 0)

    def ops(self) -> np.ndarray:
        return np.array(self._ops, dtype=OP_DTYPE) if self._ops else np.zeros(0, dtype=OP_DTYPE)

    def build(self, tracer, old_leaf=None):
        """-> (tracer.CompressedColorLeaf on the tracer's device, kernel milliseconds)."""
        return tracer.rebuild_color_leaf(self.ops(), old_leaf)
