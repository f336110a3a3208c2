// Library-format companions of a HashDAG pool, built page by page on the GPU:
//
//   resolved pool  every child pointer pushed through the page table once (HashDagResolvedDev, hdt_device.cuh): a descent
//                  is two dependent loads (child word, child header) instead of three;
//   prefix pool    (optional) for every child-pointer word of a node of depth >= 10 -- the levels below the colour tree,
//                  hash_dag_globals.h:7 -- the number of voxels under the node's EARLIER children (low 24 bits; a subtree
//                  of those levels holds < 2^24 voxels, hash_table.h:772), i.e. the sum trace_colors forms by loading every
//                  preceding sibling (tracer.cu:391-420, get_leaves_count = header >> 8, hash_dag_colors.h:28-32; popcount
//                  of the 64-bit leaves at the last interior level).  With it the colour walk needs one load per level
//                  instead of 2 + 2 x (earlier siblings).  Pointer words to 64-bit leaves also carry, in the top byte,
//                  the leaf's first-level child mask (base_dag.h:16-40) for the traversal (HashDagPrefixDev).
//
// The reference's pool is an array of 512-word physical pages; a virtual page belongs to one bucket of one level
// (hash_table.h:45-63), interior nodes `[header][child pointer] x popc(header & 0xFF)` are packed from the start of a page
// and never straddle one (hash_table.h:416-442), leaf pages (level levels-2) hold 64-bit masks.  So a page is parsed on
// its own:
//   map_pages_kernel      page table -> physToVirt[physical page] = virtual page
//   resolve_pages_kernel  one warp per physical page: copy it; if it is an interior page, lane 0 walks the headers and
//                         records for every pointer word the position of its node's header, then all lanes translate the
//                         pointer words (pageTable[v >> 9] * 512 + (v & 511)) and, for the prefix pool, add up the
//                         counts of the node's earlier children.
// Pages are re-done wholesale after an edit (the pages its spans touch): the operation is idempotent because it always
// reads the caller's (virtual) pool, and nodes are immutable once written (the hash table only appends).
//
// What is NOT assumed about the caller's memory: the words behind the last node of a page (page tails, the gap
// add_interior_node leaves when a node would straddle a page, pages of buckets that are not full) may hold anything --
// the reference neither clears its pool nor pads with zeros.  The parser stops at the first header without children and
// may otherwise run through such words; what it makes of them lands in words no node points at.  Every index derived
// from pool contents is bounds-checked (page table size, pool size), so garbage can never be dereferenced.
#pragma once
#include "hdt_device.cuh"

namespace hdt {

struct HashLayoutDev {   // hash_dag_globals.h:7-38 with the defaults of typedefs.h:201-236
    u32 levels;
    u32 pageTableSize;
    u32 poolTop;         // pages
    __device__ __forceinline__ u32 level_of_page(u32 vpage) const
    {
        const u32 topPages = 9u * 1024u * (1024u / kPageWords);          // 9 top levels x 1024 buckets x 2 pages
        if (vpage < topPages) return vpage / (1024u * (1024u / kPageWords));
        return 9u + (vpage - topPages) / (65536u * (4096u / kPageWords));
    }
};

__global__ void map_pages_kernel(const u32* __restrict__ pageTable, u32 pageTableSize, u32* __restrict__ physToVirt, u32 poolTop)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= pageTableSize) return;
    const u32 p = pageTable[v];
    if (p && p < poolTop) physToVirt[p] = v;
}

// pages: list of physical pages to do, or nullptr = pages [0, nPages).  prefix may be nullptr.
__global__ void __launch_bounds__(128) resolve_pages_kernel(const u32* __restrict__ vpool, const u32* __restrict__ pageTable, const u32* __restrict__ physToVirt,
                                                             const u32* __restrict__ pages, u32 nPages, HashLayoutDev layout, u32* __restrict__ resolved,
                                                             u32* __restrict__ prefix)
{
    __shared__ u32 words[4][kPageWords];
    __shared__ u16 owner[4][kPageWords];   // pointer word -> position of its node's header; 0xFFFF = not a pointer word
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 i = blockIdx.x * 4 + warp;
    if (i >= nPages) return;
    const u32 page = pages ? pages[i] : i;
    if (page >= layout.poolTop) return;
    const u32 vpage = physToVirt[page];
    const u32* src = vpool + u64(page) * kPageWords;
    for (u32 k = lane; k < kPageWords; k += 32) { words[warp][k] = src[k]; owner[warp][k] = 0xFFFFu; }
    __syncwarp();
    const u32 level = vpage != 0xFFFFFFFFu ? layout.level_of_page(vpage) : 0xFFFFFFFFu;
    const bool interior = level < layout.levels - 2;
    if (interior && lane == 0) {
        u32 pos = 0;
        while (pos < kPageWords) {
            const u32 hdr = words[warp][pos];
            if ((hdr & 0xFF) == 0) break;                       // a node header has at least one child
            const u32 n = __popc(hdr & 0xFF);
            if (pos + n >= kPageWords) break;                   // nodes never straddle a page: not a node
            for (u32 k = pos + 1; k <= pos + n; ++k) owner[warp][k] = u16(pos);
            pos += 1 + n;
        }
    }
    __syncwarp();
    // translate: a pointer the page table cannot translate (garbage behind the last node) is left as it is
    const u64 poolWords = u64(layout.poolTop) * kPageWords;
    for (u32 k = lane; k < kPageWords; k += 32) {
        u32 w = words[warp][k];
        if (owner[warp][k] != 0xFFFFu && (w >> 9) < layout.pageTableSize) {
            const u32 phys = __ldg(pageTable + (w >> 9));
            w = phys * kPageWords + (w & (kPageWords - 1));
            if (phys == 0 || phys >= layout.poolTop) owner[warp][k] = 0xFFFEu;   // translated to nowhere: not counted below
        } else if (owner[warp][k] != 0xFFFFu) {
            owner[warp][k] = 0xFFFEu;
        }
        words[warp][k] = w;
        resolved[u64(page) * kPageWords + k] = w;
    }
    if (!prefix) return;
    __syncwarp();
    // prefix pool: voxels under the earlier children of the node, for the levels whose siblings trace_colors counts
    const bool counted = interior && level >= kColorTreeDepth;
    const bool leafChildren = level + 3 == layout.levels;      // children are the 64-bit leaves
    for (u32 k = lane; k < kPageWords; k += 32) {
        u32 sum = 0;
        const u32 own = owner[warp][k];
        if (leafChildren && own < 0xFFFEu) {   // the leaf this word points at: its first-level child mask
            const u32 c = words[warp][k];
            if (u64(c) + 1 < poolWords) sum = first_child_mask(__ldg(reinterpret_cast<const uint2*>(vpool + (c & ~1u)))) << 24;
        }
        if (counted && own < 0xFFFEu) {
            for (u32 j = own + 1; j < k; ++j) {
                if (owner[warp][j] != own) continue;            // an earlier pointer of this node that translated to nowhere
                const u32 c = words[warp][j];
                if (u64(c) + 1 >= poolWords) continue;
                if (leafChildren) { const uint2 l = __ldg(reinterpret_cast<const uint2*>(vpool + (c & ~1u))); sum += __popc(l.x) + __popc(l.y); }
                else sum += __ldg(vpool + c) >> 8;          // stays below 2^24: the sum is at most the node's own count24
            }
        }
        prefix[u64(page) * kPageWords + k] = sum;
    }
}

}  // namespace hdt
