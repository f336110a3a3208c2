// Region queries on the device: DAGUtils::get_values (dag_utils.h:268-411) and
// DAGUtils::is_empty (dag_utils.h:175-266), the two read-only DAG walks the copy tool runs on the host
// (hash_dag_editors.h:401, :476, :570).
//
// Both reference functions prune with the same box test (dag_utils.h:190-207, :278-296): a node with voxel bounds
// [bmin, bmax] (bmax inclusive) is visited iff, on every axis, bmin < start+size and bmax > start.  Because bmax is
// inclusive the test is strict on the low side too: a single voxel p passes iff start < p < start+size.  The test
// is monotone (a box that passes makes every box containing it pass), so the recursions collapse to point queries:
//   get_values : values[p - start] = voxel p exists AND start < p < start+size per axis (the p == start planes keep
//                the memset's zero, exactly like the reference);
//   is_empty   : false iff maxLevel == 0 or some EXISTING node of level maxLevel-1 has bmin < start+size and
//                bmax > start (the recursion returns false as soon as it reaches level maxLevel, without testing
//                that node's own box).
// One thread per 4x4x4 leaf cell (get_values) / per candidate node cell (is_empty), each walking down from the root.
#pragma once
#include "hdt_device.cuh"

namespace hdt {

struct RegionParams { u32 start[3]; u32 size[3]; u32 cell0[3]; u32 nCells[3]; };

// Descend `steps` levels from the root along the cell coordinates (cx,cy,cz), `steps` bits each, MSB first
// (Path::child_index, path.h:20-28).  Returns false if a child is missing; `h` = handle of the node reached.
template <class DAG>
__device__ __forceinline__ bool descend_to(const DAG& dag, u32 steps, u32 cx, u32 cy, u32 cz, u32& h)
{
    h = dag.root();
    for (u32 level = 0; level < steps; ++level) {
        const u32 cm = dag.header(h) & 0xFF;
        const u32 sh = steps - 1 - level;
        const u32 child = (((cx >> sh) & 1) << 2) | (((cy >> sh) & 1) << 1) | ((cz >> sh) & 1);
        if (!(cm & (1u << child))) return false;
        h = dag.child(h, __popc(cm & ((1u << child) - 1)) + 1);
    }
    return true;
}

template <class DAG>
__global__ void __launch_bounds__(128) get_values_kernel(const DAG dag, const u32 levels, const RegionParams rp, u8* __restrict__ values)
{
    const u64 nCells = u64(rp.nCells[0]) * rp.nCells[1] * rp.nCells[2];
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nCells) return;
    const u32 cx = rp.cell0[0] + u32(i % rp.nCells[0]), cy = rp.cell0[1] + u32((i / rp.nCells[0]) % rp.nCells[1]), cz = rp.cell0[2] + u32(i / (u64(rp.nCells[0]) * rp.nCells[1]));
    u32 h;
    uint2 leaf = make_uint2(0, 0);
    if (descend_to(dag, levels - 2, cx, cy, cz, h)) leaf = dag.leaf(h);
    const u32 endx = rp.start[0] + rp.size[0], endy = rp.start[1] + rp.size[1], endz = rp.start[2] + rp.size[2];
#pragma unroll
    for (u32 z = 0; z < 4; ++z) {
        const u32 pz = cz * 4 + z;
        if (pz < rp.start[2] || pz >= endz) continue;
#pragma unroll
        for (u32 y = 0; y < 4; ++y) {
            const u32 py = cy * 4 + y;
            if (py < rp.start[1] || py >= endy) continue;
            u8* row = values + (u64(pz - rp.start[2]) * rp.size[1] + (py - rp.start[1])) * rp.size[0];
#pragma unroll
            for (u32 x = 0; x < 4; ++x) {
                const u32 px = cx * 4 + x;
                if (px < rp.start[0] || px >= endx) continue;
                // leaf bit = child1 * 8 + child2 (dag_utils.h:159-165, :318-320)
                const u32 c1 = ((x >> 1) << 2) | ((y >> 1) << 1) | (z >> 1), c2 = ((x & 1) << 2) | ((y & 1) << 1) | (z & 1);
                const u32 word = (c1 & 4) ? leaf.y : leaf.x;
                const bool set = (word >> ((c1 & 3) * 8 + c2)) & 1;
                const bool inside = px > rp.start[0] && py > rp.start[1] && pz > rp.start[2];   // strict: see the file comment
                row[px - rp.start[0]] = u8(set && inside);
            }
        }
    }
}

// Candidate nodes of level `steps` = maxLevel-1 (cells of 2^shift voxels); found[0] != 0 once one exists.
template <class DAG>
__global__ void __launch_bounds__(128) is_empty_kernel(const DAG dag, const u32 steps, const RegionParams rp, u32* __restrict__ found)
{
    const u64 nCells = u64(rp.nCells[0]) * rp.nCells[1] * rp.nCells[2];
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nCells) return;
    if (*reinterpret_cast<volatile u32*>(found)) return;
    const u32 cx = rp.cell0[0] + u32(i % rp.nCells[0]), cy = rp.cell0[1] + u32((i / rp.nCells[0]) % rp.nCells[1]), cz = rp.cell0[2] + u32(i / (u64(rp.nCells[0]) * rp.nCells[1]));
    u32 h;
    if (descend_to(dag, steps, cx, cy, cz, h)) atomicOr(found, 1u);
}

}  // namespace hdt
