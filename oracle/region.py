"""CPU oracle of the copy tool's region queries.  TEST INFRASTRUCTURE: only tests/ may import it.

Statement-by-statement restatement of /root/reference/src/dags/dag_utils.h for host arrays:
  * `should_edit`   the box test both functions prune with            :190-207 / :278-296
  * `is_empty`      DAGUtils::is_empty / is_empty_impl                :175-266
  * `get_values`    DAGUtils::get_values / get_values_impl            :268-411 (its task list / threads only
                    distribute the same recursion, :341-361, :381-410)
Plain Python recursion over the node arrays: small regions only.

Pinning: no golden vectors in the reference; tests/test_gpu_region.py runs the reference's own host functions
(oracle/_ref: ref_get_values / ref_is_empty, on the original and on an edited HashDAG) beside this oracle and the
CUDA product, and tests/golden/ref_regions_d13.npz keeps outputs of those reference calls for the CPU suite.
"""
from __future__ import annotations

import numpy as np

PAGE = 512


class HostDag:
    """get_sys_ptr-style access to a BasicDAG (`data`) or a HashDAG (`pool` + `page_table`) held in numpy arrays."""

    def __init__(self, levels, data=None, pool=None, page_table=None, first_node_index=0):
        self.levels, self.data, self.pool, self.table, self.root = levels, data, pool, page_table, (0 if data is not None else int(first_node_index))

    def word(self, ptr):
        if self.data is not None:
            return int(self.data[ptr])
        return int(self.pool[int(self.table[ptr // PAGE]) * PAGE + ptr % PAGE])     # hash_table.h:156-173

    @classmethod
    def from_scene(cls, scene, kind="hash"):
        if kind == "basic":
            return cls(scene.levels, data=scene.basic)
        return cls(scene.levels, pool=scene.hash_pool, page_table=scene.hash_page_table, first_node_index=scene.hash_first_node_index)


def should_edit(path, shift, start, size):
    """dag_utils.h:190-207: bounds of the node `path` at `shift` against [start, start+size)."""
    for a in range(3):
        bmin = path[a] << shift
        bmax = bmin + (1 << shift) - 1          # inclusive
        in_min, in_max = start[a], start[a] + size[a]
        if bmin >= in_max or bmax <= in_min or in_min >= bmax or in_max <= bmin:
            return False
    return True


def _descend(path, child):                      # Path::descend, path.h:30-36
    return ((path[0] << 1) | ((child >> 2) & 1), (path[1] << 1) | ((child >> 1) & 1), (path[2] << 1) | (child & 1))


def is_empty(dag: HostDag, max_level, start, size):
    leaf_level = dag.levels - 2
    assert max_level <= leaf_level              # checkAlways, :264

    def impl(node, path, level):
        if level == max_level:
            return False
        if level < leaf_level and not should_edit(path, dag.levels - level, start, size):
            return True
        child_mask = dag.word(node) & 0xFF
        off = 1
        for child in range(8):
            if child_mask & (1 << child):
                child_node = dag.word(node + off)
                off += 1
                if not impl(child_node, _descend(path, child), level + 1):
                    return False
        return True

    return impl(dag.root, (0, 0, 0), 0)


def get_values(dag: HostDag, start, size):
    """-> uint8[size.z, size.y, size.x] (values[x + size.x*y + size.x*size.y*z], :323)."""
    leaf_level = dag.levels - 2
    values = np.zeros((size[2], size[1], size[0]), dtype=np.uint8)      # the memset of :379

    def impl(node, path, level):
        if level < leaf_level and not should_edit(path, dag.levels - level, start, size):
            return
        if level == leaf_level:
            low, high = dag.word(node), dag.word(node + 1)
            for child1 in range(8):
                p1 = _descend(path, child1)
                if not should_edit(p1, 1, start, size):
                    continue
                for child2 in range(8):
                    p2 = _descend(p1, child2)
                    if should_edit(p2, 0, start, size):
                        mask = high if (child1 & 4) else low
                        bit = (child1 & 3) * 8 + child2
                        values[p2[2] - start[2], p2[1] - start[1], p2[0] - start[0]] = (mask >> bit) & 1
            return
        child_mask = dag.word(node) & 0xFF
        off = 1
        for child in range(8):
            if child_mask & (1 << child):
                child_node = dag.word(node + off)
                off += 1
                impl(child_node, _descend(path, child), level + 1)

    impl(dag.root, (0, 0, 0), 0)
    return values


def get_value(dag: HostDag, p):
    """DAGUtils::get_value (dag_utils.h:138-172): does voxel p exist?"""
    node = dag.root
    for level in range(dag.levels):
        if level < dag.levels - 2:
            cm = dag.word(node) & 0xFF
            sh = dag.levels - (level + 1)
            child = (((p[0] >> sh) & 1) << 2) | (((p[1] >> sh) & 1) << 1) | ((p[2] >> sh) & 1)
            if not cm & (1 << child):
                return False
            node = dag.word(node + bin(cm & ((1 << child) - 1)).count("1") + 1)
        else:
            leaf = dag.word(node) | (dag.word(node + 1) << 32)
            bit = ((p[0] & 1) << 2) | ((p[1] & 1) << 1) | (p[2] & 1) | ((p[0] & 2) << 4) | ((p[1] & 2) << 3) | ((p[2] & 2) << 2)
            return bool((leaf >> bit) & 1)
    return True
