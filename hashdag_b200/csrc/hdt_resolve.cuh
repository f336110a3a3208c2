// Resolving a HashDAG pool: every child pointer pushed through the page table once (HashDagResolvedDev, hdt_device.cuh).
//
// The reference's pool is an array of 512-word physical pages; a virtual page belongs to one bucket of one level
// (hash_table.h:45-63), interior nodes `[header][child pointer] x popc(header & 0xFF)` are packed from the start of a page
// and never straddle one (add_interior_node pads the tail with zeros, hash_table.h:416-430), leaf pages (level
// levels-2) hold 64-bit masks.  So a page can be parsed on its own:
//   map_pages_kernel      page table -> physToVirt[physical page] = virtual page
//   resolve_pages_kernel  one warp per physical page: copy it; if it is an interior page, lane 0 walks the headers and
//                         marks the pointer words in a 512-bit map, then all lanes translate the marked words
//                         (pageTable[v >> 9] * 512 + (v & 511)).
// Pages are re-resolved wholesale after an edit (the pages its spans touch): the operation is idempotent because it
// always reads the caller's (virtual) pool.
#pragma once
#include "hdt_device.cuh"

namespace hdt {

struct HashLayoutDev {   // hash_dag_globals.h:7-38 with the defaults of typedefs.h:201-236
    u32 levels;
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

// pages: list of physical pages to resolve, or nullptr = pages [0, nPages)
__global__ void __launch_bounds__(128) resolve_pages_kernel(const u32* __restrict__ vpool, const u32* __restrict__ pageTable, const u32* __restrict__ physToVirt,
                                                             const u32* __restrict__ pages, u32 nPages, HashLayoutDev layout, u32* __restrict__ resolved)
{
    __shared__ u32 words[4][kPageWords];
    __shared__ u32 marks[4][kPageWords / 32];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 i = blockIdx.x * 4 + warp;
    if (i >= nPages) return;
    const u32 page = pages ? pages[i] : i;
    const u32 vpage = physToVirt[page];
    const u32* src = vpool + u64(page) * kPageWords;
    u32* dst = resolved + u64(page) * kPageWords;
    for (u32 k = lane; k < kPageWords; k += 32) words[warp][k] = src[k];
    if (lane < kPageWords / 32) marks[warp][lane] = 0;
    __syncwarp();
    const bool interior = vpage != 0xFFFFFFFFu && layout.level_of_page(vpage) < layout.levels - 2;
    if (interior && lane == 0) {
        u32 pos = 0;
        while (pos < kPageWords) {
            const u32 hdr = words[warp][pos];
            if ((hdr & 0xFF) == 0) break;                       // padding / unused tail: a node header has at least one child
            const u32 n = __popc(hdr & 0xFF);
            for (u32 k = pos + 1; k <= pos + n && k < kPageWords; ++k) marks[warp][k >> 5] |= 1u << (k & 31);
            pos += 1 + n;
        }
    }
    __syncwarp();
    for (u32 k = lane; k < kPageWords; k += 32) {
        u32 w = words[warp][k];
        if ((marks[warp][k >> 5] >> (k & 31)) & 1) w = __ldg(pageTable + (w >> 9)) * kPageWords + (w & (kPageWords - 1));
        dst[k] = w;
    }
}

}  // namespace hdt
