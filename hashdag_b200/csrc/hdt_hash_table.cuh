// GPU batch insert into the reference's hash table: HashTable::find_or_add_interior_node /
// find_or_add_leaf_node (hash_table.h:470-560) for a batch of candidate nodes of ONE level, in insertion order.  The result is
// what the reference's host code produces when it is handed the same nodes one after the other: the same virtual pointers,
// and the same pool, page table, bucket fill counts and pool top afterwards -- including its quirks:
//   * a bucket is searched page by page, node by node (a match must start at a node boundary), and for interior nodes the walk
//     of a page stops `nodeSize` words before the page's end, so a node that sits exactly at the end of a (partly filled) page is
//     not found and is added again (hash_table.h:322-352);
//   * an interior node never straddles a page: if it does not fit, the bucket's fill count jumps to the next page first
//     (add_interior_node, hash_table.h:416-442); leaves are appended two words at a time (add_leaf_node, :355-400);
//   * a page is allocated when the first node lands in it, physical pages handed out in insertion order (allocate_page, :794-807).
// The reference's Bloom filter only skips pages a node cannot be in; it never changes a result and has no counterpart here.
// The reference never writes the padding in front of a page boundary but later searches walk over it: here padding inside a page
// that existed before the batch is read from the pool like every other old word, and pages the batch itself opens are taken to be
// zero-filled (the reference's pool pages come untouched from a fresh allocation, hash_table.cpp:60-76).
//
// How the batch is made parallel without changing the sequential result: candidates only interact inside their bucket, and
// through the ORDER in which new physical pages are handed out.
//   1. foa_hash_kernel      candidate i -> its bucket (murmur hashes of utils.h:68-110), sort key (bucket << 32 | i)
//      (keys sorted by cub::DeviceRadixSort: buckets become segments, insertion order kept inside each)
//   2. foa_bucket_kernel    one thread per bucket that has candidates walks them in insertion order: search the bucket's old
//                           words in the pool and the words appended earlier in this batch in a staging area; not found ->
//                           append (fill count, page rule) and note the virtual page a candidate is the first to need
//   3. foa_pages_kernel     exclusive scan of "candidate i opened a page" over i = the reference's allocation order
//   4. (host: does the batch fit the pool and the buckets?  else nothing has been written)
//      foa_apply_kernel     pageTable[page] = poolTop + rank, new bucket fill counts
//      foa_commit_kernel    added nodes are copied to their physical place
#pragma once
#include "hdt_device.cuh"

namespace hdt {

constexpr u32 kNoNode = 0xFFFFFFFFu;

struct HashTableDev {
    u32* pool; u32* pageTable; u32* bucketSizes;
    u32 pageTableSize, levels, poolTop;
    u64 poolCapacityWords;
    // hash_dag_globals.h:7-38 with the defaults of typedefs.h:201-236
    __host__ __device__ static u32 buckets_per_level(u32 level) { return level < 9 ? 1024u : 65536u; }
    __host__ __device__ static u32 bucket_capacity(u32 level) { return level < 9 ? 1024u : 4096u; }
    __host__ __device__ static u32 bucket_global_index(u32 level, u32 bucket) { return level < 9 ? level * 1024u + bucket : 9u * 1024u + (level - 9u) * 65536u + bucket; }
    __host__ __device__ static u32 make_ptr(u32 level, u32 bucket, u32 pos)   // hash_table.h:45-63
    {
        return level < 9 ? (level * 1024u + bucket) * 1024u + pos : 9u * 1024u * 1024u + ((level - 9u) * 65536u + bucket) * 4096u + pos;
    }
};

__host__ __device__ inline u32 murmur32xN(const u32* w, u32 n)   // utils.h:91-110 (USE_ALTERNATE_HASH)
{
    u32 h = 0;
    for (u32 i = 0; i < n; ++i) {
        u32 k = w[i];
        k *= 0xcc9e2d51u; k = (k << 15) | (k >> 17); k *= 0x1b873593u;
        h ^= k; h = (h << 13) | (h >> 19); h = h * 5 + 0xe6546b64u;
    }
    h ^= n;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
__host__ __device__ inline u32 murmur64_low(u64 h)             // uint32(Utils::murmurhash64(leaf)), utils.h:77-85
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return u32(h);
}

// candidates: node i = words[offsets[i] .. offsets[i+1])
// (errors |= 2 for a candidate that cannot be a node: a leaf is 2 words, an interior node 1 + popc(child mask) words, 2..9)
__global__ void foa_hash_kernel(const u32* __restrict__ words, const u64* __restrict__ offsets, u32 n, u32 level, bool leaves, u64* __restrict__ keys, u32* __restrict__ errors)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32* w = words + offsets[i];
    const u32 size = u32(offsets[i + 1] - offsets[i]);
    if (leaves ? size != 2 : (size < 2 || size > 9 || size != u32(__popc(w[0] & 0xFF)) + 1)) { atomicOr(errors, 2u); keys[i] = i; return; }
    const u32 hash = leaves ? murmur64_low(u64(w[0]) | (u64(w[1]) << 32)) : murmur32xN(w, size);
    keys[i] = (u64(hash & (HashTableDev::buckets_per_level(level) - 1)) << 32) | i;
}

// segment starts of the sorted keys: segStart[s] = first sorted index of the s-th distinct bucket; nSeg counted atomically
__global__ void foa_segments_kernel(const u64* __restrict__ keys, u32 n, u32* __restrict__ isStart)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    isStart[k] = (k == 0 || (keys[k] >> 32) != (keys[k - 1] >> 32)) ? 1u : 0u;
}

// per candidate: result pointer, 1 if appended, virtual page it opened (or kNoNode); per bucket segment: global bucket index and
// its new fill count.  Nothing of the table is written before the batch is known to fit (foa_apply_kernel).
struct FoaOut { u32* ptrs; u32* added; u32* openedPage; u32* errors; u32* segBucket; u32* segSize; };

__global__ void foa_bucket_kernel(const HashTableDev t, const u32 level, const bool leaves, const u32* __restrict__ words, const u64* __restrict__ offsets,
                                  const u64* __restrict__ keys, const u32 n, const u32* __restrict__ segFirst, const u32 nSeg, u32* __restrict__ staging, const FoaOut out)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nSeg) return;
    const u32 k0 = segFirst[s], k1 = (s + 1 < nSeg) ? segFirst[s + 1] : n;
    const u32 bucket = u32(keys[k0] >> 32);
    const u32 g = HashTableDev::bucket_global_index(level, bucket);
    const u32 base = HashTableDev::make_ptr(level, bucket, 0), cap = HashTableDev::bucket_capacity(level);
    const u32 orig = t.bucketSizes[g];
    u32 size = orig;
    u32* stage = staging + u64(k0) * 17;           // words appended in this batch: position p >= orig lives at stage[p - orig]
    auto word = [&](u32 pos) -> u32 {
        if (pos >= orig) return stage[pos - orig];
        const u32 phys = t.pageTable[(base + pos) / kPageWords];
        return t.pool[u64(phys) * kPageWords + (pos % kPageWords)];
    };
    for (u32 k = k0; k < k1; ++k) {
        const u32 i = u32(keys[k]);
        const u32* w = words + offsets[i];
        const u32 ns = u32(offsets[i + 1] - offsets[i]);
        u32 found = kNoNode;
        if (leaves) {   // find_leaf_node_in_bucket, hash_table.h:196-262
            for (u32 pos = 0; pos < size && found == kNoNode; pos += 2)
                if (word(pos) == w[0] && word(pos + 1) == w[1]) found = pos;
        } else {        // find_interior_node_in_bucket, hash_table.h:264-358
            for (u32 pindex = 0; pindex < size && found == kNoNode; pindex += kPageWords) {
                u32 pageEnd = min(size, pindex + kPageWords);
                if (pindex + ns >= pageEnd) break;           // `return 0xFFFFFFFF`: the search ends here
                pageEnd -= ns;
                for (u32 index = pindex; index < pageEnd;) {
                    const u32 first = word(index);
                    bool same = first == w[0];
                    for (u32 j = 1; same && j < ns; ++j) same = word(index + j) == w[j];
                    if (same) { found = index; break; }
                    index += __popc(first & 0xFF) + 1;
                }
            }
        }
        u32 added = 0, opened = kNoNode;
        if (found == kNoNode) {
            u32 pos = size;
            if (leaves) {   // add_leaf_node, hash_table.h:355-400
                if (pos % kPageWords == 0 && t.pageTable[(base + pos) / kPageWords] == 0) opened = (base + pos) / kPageWords;
            } else {        // add_interior_node, hash_table.h:401-468
                const u32 left = kPageWords - (pos % kPageWords);
                if (left == kPageWords || left < ns) {
                    if (left != kPageWords) {                // the bucket's fill count includes the padding up to the page boundary
                        // later searches walk over the padding: in a page that existed before the batch it is what the pool holds there
                        if (orig % kPageWords != 0 && pos / kPageWords == orig / kPageWords) {
                            const u32 phys = t.pageTable[(base + pos) / kPageWords];
                            for (u32 j = 0; j < left; ++j) stage[pos - orig + j] = t.pool[u64(phys) * kPageWords + (pos % kPageWords) + j];
                        }
                        pos += left;
                    }
                    if (pos < cap && t.pageTable[(base + pos) / kPageWords] == 0) opened = (base + pos) / kPageWords;
                }
            }
            if (pos + ns >= cap) { atomicOr(out.errors, 1u); pos = size; }   // "Bucket size on level %u too low" (hash_table.h:461): the batch is refused
            else {
                for (u32 j = 0; j < ns; ++j) stage[pos - orig + j] = w[j];
                size = pos + ns;
                added = 1;
            }
            found = pos;
        }
        out.ptrs[i] = base + found;
        out.added[i] = added;
        out.openedPage[i] = added ? opened : kNoNode;
    }
    out.segBucket[s] = g;
    out.segSize[s] = size;
}

// exclusive count of "opened a page" over the candidates in insertion order -> physical pages in the reference's allocation order
__global__ void __launch_bounds__(1024) foa_pages_kernel(const u32* __restrict__ openedPage, const u32 n, u32* __restrict__ pageRank, u32* __restrict__ totals)
{
    __shared__ u32 warpSums[32];
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += blockDim.x) {
        const u32 i = base + threadIdx.x;
        const u32 page = i < n ? openedPage[i] : kNoNode;
        const u32 flag = page != kNoNode ? 1u : 0u;
        const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        u32 inc = flag;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) warpSums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u32 v = warpSums[lane];
#pragma unroll
            for (u32 d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v += o; }
            warpSums[lane] = v;
        }
        __syncthreads();
        const u32 rank = carry + (warp ? warpSums[warp - 1] : 0) + inc - flag;
        if (i < n) pageRank[i] = rank;
        __syncthreads();
        if (threadIdx.x == 0) carry += warpSums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[0] = carry;
}

// the batch fits: page table entries of the opened pages, new fill counts
__global__ void foa_apply_kernel(const HashTableDev t, const u32 n, const u32 nSeg, const FoaOut out, const u32* __restrict__ pageRank)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && out.openedPage[i] != kNoNode) t.pageTable[out.openedPage[i]] = t.poolTop + pageRank[i];
    if (i < nSeg) t.bucketSizes[out.segBucket[i]] = out.segSize[i];
}

__global__ void foa_commit_kernel(const HashTableDev t, const u32* __restrict__ words, const u64* __restrict__ offsets, const u32 n, const FoaOut out, u32* __restrict__ totals)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !out.added[i]) return;
    const u32 ptr = out.ptrs[i];
    const u32 phys = t.pageTable[ptr / kPageWords];
    const u32* w = words + offsets[i];
    const u32 ns = u32(offsets[i + 1] - offsets[i]);
    u32* dst = t.pool + u64(phys) * kPageWords + (ptr % kPageWords);
    for (u32 j = 0; j < ns; ++j) dst[j] = w[j];
    atomicAdd(totals + 1, 1u);
}

}  // namespace hdt
