// Colour-leaf rebuild on the device: what the reference does on the host, one voxel at a
// time, whenever an edit touches a colour leaf --
//   CompressedColorLeaf::copy_colors   variable_weight_size_colors.h:416-542   (old colours of an untouched subtree)
//   ColorLeafBuilder::add / add_weight / add_large_single_color / build        :549-721
// driven from hash_dag_edits.h:381-396 (copy), :430-437 (entirely-full node), :488-516 (single voxels).
//
// The encoder is a run-length + bit-packing codec over the stream of CompressedColor{colorBits, weight,
// bitsPerWeight} in voxel order:
//   * a macro block starts every 16384 colours and stores {first block index, weight bit offset};
//   * a block starts at every macro boundary and wherever (colorBits, bitsPerWeight) differs from the previous
//     colour; its header holds the weight bit offset and colour index RELATIVE to the macro block;
//   * weights are appended MSB-first to a bit stream kept in 32-bit words, each byte-swapped by build().
// None of this depends on more than the previous colour and two prefix sums, so the device version is:
//   1. count_color_ops_kernel    one CTA per macro block of the NEW leaf, one thread per 16 consecutive colours:
//                                evaluate the op list (copy from the old leaf, or a constant colour) and reduce
//                                {blocks started, weight bits} per macro block;
//   2. scan_color_tiles_kernel   exclusive scan of those pairs (one CTA; a leaf has n/16384 macro blocks);
//   3. emit_color_leaf_kernel    one CTA per macro block: evaluate the ops again, CTA-wide scan, block headers,
//                                macro-block pairs, and the weight bits assembled in shared memory and stored as
//                                whole swapped words (the two words a macro block may share with its neighbours
//                                go through atomicOr).
// The colours are evaluated twice rather than staged in memory: a thread reads the old leaf like copy_colors does
// (one binary search, then a walk along the blocks), which costs a few cached loads per colour, whereas a staged
// stream would be 16 B of HBM traffic per colour against ~1.5 B of algorithmic bytes (old leaf in, new leaf out).
// Bit-exact with the reference builder by construction; pinned against leaves the reference built
// (tests/golden/ref_color_leaves_d13.npz, tests/test_gpu_color_leaf.py).
#pragma once
#include "hdt_colors.cuh"

namespace hdt {

constexpr u32 kRebuildThreads = 1024;
constexpr u32 kColorsPerThread = 16;
static_assert(u64(kRebuildThreads) * kColorsPerThread == kColorsPerMacroBlock, "one CTA per macro block");

// hdt_color_op with the exclusive prefix of the counts (where the op's first colour lands in the new leaf).
struct ColorOpDev { u64 dstStart; u64 srcStart; u32 kind; u32 bitsPerWeight; u32 colorBits; u32 weight; };
static_assert(sizeof(ColorOpDev) == 32, "uploaded as is");

// A colour of the stream: colorBits | bitsPerWeight << 32 | weight << 40.  Two colours continue the same block
// iff their low 40 bits agree (ColorLeafBuilder::add, vwsc.h:606).
__device__ __forceinline__ u64 pack_color(u32 colorBits, u32 bpw, u32 weight) { return u64(colorBits) | (u64(bpw) << 32) | (u64(weight) << 40); }
constexpr u64 kBlockKeyMask = (u64(1) << 40) - 1;

struct TilePair { u32 blocks; u32 bits; };   // per macro block: blocks started, weight bits appended

__device__ __forceinline__ u64 cta_exclusive_scan(u64 v, u64& total)
{
    __shared__ u64 warpSums[32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        const u64 o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const u32 nWarps = blockDim.x >> 5;
        u64 w = lane < nWarps ? warpSums[lane] : 0;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            const u64 o = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += o;
        }
        warpSums[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const u64 base = warp ? warpSums[warp - 1] : 0;
    total = warpSums[(blockDim.x >> 5) - 1];
    __syncthreads();          // warpSums may be reused by the caller's next scan
    return base + inc - v;
}

// Block starts and weight bits of one thread's 16 colours.  prevKey = key of the colour before the first one
// (ignored at the start of a macro block, which always starts a block).
__device__ __forceinline__ void thread_flags(const u64 (&c)[kColorsPerThread], u32 nValid, bool macroStart, u64 prevKey, u32& flags, u32& nBlocks, u32& nBits)
{
    flags = 0; nBits = 0;
#pragma unroll
    for (u32 j = 0; j < kColorsPerThread; ++j) {
        if (j < nValid) {
            const u64 key = c[j] & kBlockKeyMask;
            if ((j == 0 && macroStart) || key != prevKey) flags |= 1u << j;
            prevKey = key;
            nBits += u32(c[j] >> 32) & 0xFF;
        }
    }
    nBlocks = __popc(flags);
}

// Walks the old leaf like CompressedColorLeaf::copy_colors (vwsc.h:462-540): binary_search_blocks once, then block by
// block and macro block by macro block.
struct LeafCursor {
    u32 macro, block, lastBlock, nextStart;   // nextStart: local colour index where the next block of this macro block starts
    u32 hdr, colorBits, bpw;
    u64 macroWeightOffset;

    __device__ __forceinline__ void load_block(const ColorLeafDev& l)
    {
        const u64 b = __ldg(l.blocks + block);
        hdr = u32(b); colorBits = u32(b >> 32);
        bpw = ((hdr >> 16) == 0xFFFF) ? 0u : (((hdr >> 14) & 0x3) + 1);
        nextStart = block < lastBlock ? (u32(__ldg(l.blocks + block + 1)) & 0x3FFF) : u32(kColorsPerMacroBlock);
    }
    __device__ __forceinline__ void load_macro(const ColorLeafDev& l)
    {
        lastBlock = (2 * (u64(macro) + 1) < l.nMacroWords) ? u32(__ldg(l.macroBlocks + 2 * (macro + 1)) - 1) : u32(l.nBlocks - 1);
        macroWeightOffset = __ldg(l.macroBlocks + 2 * macro + 1);
    }
    // position on colour `colorIndex` (absolute: the shared leaf's offset already added), vwsc.h:343-372
    __device__ __forceinline__ void seek(const ColorLeafDev& l, u64 colorIndex)
    {
        const u32 local = u32(colorIndex % kColorsPerMacroBlock);
        macro = u32(colorIndex / kColorsPerMacroBlock);
        load_macro(l);
        u32 lo = u32(__ldg(l.macroBlocks + 2 * macro)), hi = lastBlock;
        u32 pos = (lo + hi) / 2;
        u32 idx = u32(__ldg(l.blocks + pos)) & 0x3FFF;
        while (idx != local && lo <= hi) {
            if (idx > local) hi = pos - 1; else lo = pos + 1;
            pos = (lo + hi) / 2;
            idx = u32(__ldg(l.blocks + pos)) & 0x3FFF;
        }
        block = pos;
        load_block(l);
    }
    // colour `local` of the current macro block (the cursor stands on its block), vwsc.h:374-403
    __device__ __forceinline__ u64 color(const ColorLeafDev& l, u32 local) const
    {
        u32 weight = 0;
        if (bpw) {
            const u64 bitPtr = macroWeightOffset + (hdr >> 16) + u64(local - (hdr & 0x3FFF)) * bpw;
            const u8* bytes = reinterpret_cast<const u8*>(l.weights) + (bitPtr >> 3);
            const u32 be16 = (u32(__ldg(bytes)) << 8) | u32(__ldg(bytes + 1));
            weight = (be16 >> (16 - bpw - u32(bitPtr & 7))) & ((1u << bpw) - 1);
        }
        return pack_color(colorBits, bpw, weight);
    }
    // step from colour `local` to `local + 1` (vwsc.h:497-520); returns the new local index
    __device__ __forceinline__ u32 advance(const ColorLeafDev& l, u32 local)
    {
        if (++local == kColorsPerMacroBlock) {
            ++macro; ++block; local = 0;
            load_macro(l);
            load_block(l);
        } else if (local >= nextStart) {
            ++block;
            load_block(l);
        }
        return local;
    }
};

// The 16 consecutive colours [first, first + nValid) of the new leaf, from the op list.
__device__ __forceinline__ void eval_ops(const ColorOpDev* __restrict__ ops, const u32 nOps, const ColorLeafDev& oldLeaf, const u64 first, const u32 nValid,
                                         u64 (&c)[kColorsPerThread])
{
#pragma unroll
    for (u32 j = 0; j < kColorsPerThread; ++j) c[j] = 0;
    if (!nValid) return;
    // op of the first colour: last op with dstStart <= first (ops[nOps].dstStart = nColors is a sentinel)
    u32 lo = 0, hi = nOps - 1;
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if (__ldg(&ops[mid].dstStart) <= first) lo = mid; else hi = mid - 1;
    }
    ColorOpDev op = ops[lo];
    u64 opEnd = __ldg(&ops[lo + 1].dstStart);
    LeafCursor cur;
    u32 local = 0;
    bool seeked = false;
#pragma unroll
    for (u32 j = 0; j < kColorsPerThread; ++j) {
        if (j < nValid) {
            const u64 i = first + j;
            while (i >= opEnd) { ++lo; op = ops[lo]; opEnd = __ldg(&ops[lo + 1].dstStart); seeked = false; }
            if (op.kind == HDT_COLOR_OP_COPY) {
                if (!seeked) {
                    const u64 src = op.srcStart + (i - op.dstStart) + (oldLeaf.is_shared() ? oldLeaf.offset : 0);
                    cur.seek(oldLeaf, src);
                    local = u32(src % kColorsPerMacroBlock);
                    seeked = true;
                } else {
                    local = cur.advance(oldLeaf, local);
                }
                c[j] = cur.color(oldLeaf, local);
            } else {
                c[j] = pack_color(op.colorBits, op.bitsPerWeight, op.weight);
            }
        }
    }
}

__global__ void __launch_bounds__(kRebuildThreads) count_color_ops_kernel(const ColorOpDev* __restrict__ ops, const u32 nOps, const ColorLeafDev oldLeaf,
                                                                            const u64 nColors, TilePair* __restrict__ tiles)
{
    __shared__ u64 lastKey[kRebuildThreads];
    const u64 first = u64(blockIdx.x) * kColorsPerMacroBlock + u64(threadIdx.x) * kColorsPerThread;
    const u32 nValid = first >= nColors ? 0u : u32(min(u64(kColorsPerThread), nColors - first));
    u64 c[kColorsPerThread];
    eval_ops(ops, nOps, oldLeaf, first, nValid, c);
    lastKey[threadIdx.x] = c[kColorsPerThread - 1] & kBlockKeyMask;
    __syncthreads();
    u32 flags, nBlocks, nBits;
    thread_flags(c, nValid, threadIdx.x == 0, threadIdx.x ? lastKey[threadIdx.x - 1] : 0, flags, nBlocks, nBits);
    u64 total;
    cta_exclusive_scan((u64(nBlocks) << 32) | nBits, total);
    if (threadIdx.x == 0) tiles[blockIdx.x] = TilePair{ u32(total >> 32), u32(total) };
}

// tiles[m] (counts) -> offsets[m] (exclusive prefix: first block index, weight bit offset); totals[0..1].
__global__ void __launch_bounds__(1024) scan_color_tiles_kernel(const TilePair* __restrict__ tiles, const u32 nTiles, ulonglong2* __restrict__ offsets,
                                                                 u64* __restrict__ totals)
{
    u64 carryBlocks = 0, carryBits = 0;
    for (u32 base = 0; base < nTiles; base += blockDim.x) {
        const u32 m = base + threadIdx.x;
        const TilePair t = m < nTiles ? tiles[m] : TilePair{ 0, 0 };
        u64 totB, totW;
        const u64 eb = cta_exclusive_scan(t.blocks, totB);
        const u64 ew = cta_exclusive_scan(t.bits, totW);
        if (m < nTiles) offsets[m] = make_ulonglong2(carryBlocks + eb, carryBits + ew);
        carryBlocks += totB; carryBits += totW;
    }
    if (threadIdx.x == 0) { totals[0] = carryBlocks; totals[1] = carryBits; }
}

// VariableColorsUtils::make_block_header, vwsc.h:32-52
__device__ __forceinline__ u32 make_block_header(u32 weightOffset, u32 bitsPerWeight, u32 index)
{
    if (bitsPerWeight == 0) weightOffset = 0xFFFF; else --bitsPerWeight;
    return (weightOffset << 16) | (bitsPerWeight << 14) | index;
}

__global__ void __launch_bounds__(kRebuildThreads) emit_color_leaf_kernel(const ColorOpDev* __restrict__ ops, const u32 nOps, const ColorLeafDev oldLeaf,
                                                                            const u64 nColors, const ulonglong2* __restrict__ offsets,
                                                                            u32* __restrict__ weights, u64* __restrict__ blocks, u64* __restrict__ macroBlocks)
{
    __shared__ u64 lastKey[kRebuildThreads];
    // a macro block holds at most 16384 * 4 weight bits = 2048 words, + 1 for the misaligned start, + 1 so that a
    // straddling store of the last weight stays in bounds
    __shared__ u32 words[kColorsPerMacroBlock * 4 / 32 + 2];
    const u64 first = u64(blockIdx.x) * kColorsPerMacroBlock + u64(threadIdx.x) * kColorsPerThread;
    const u32 nValid = first >= nColors ? 0u : u32(min(u64(kColorsPerThread), nColors - first));
    for (u32 k = threadIdx.x; k < sizeof(words) / 4; k += blockDim.x) words[k] = 0;
    u64 c[kColorsPerThread];
    eval_ops(ops, nOps, oldLeaf, first, nValid, c);
    lastKey[threadIdx.x] = c[kColorsPerThread - 1] & kBlockKeyMask;
    __syncthreads();
    u32 flags, nBlocks, nBits;
    thread_flags(c, nValid, threadIdx.x == 0, threadIdx.x ? lastKey[threadIdx.x - 1] : 0, flags, nBlocks, nBits);
    u64 total;
    const u64 excl = cta_exclusive_scan((u64(nBlocks) << 32) | nBits, total);   // also orders the zeroing of `words` before the atomics
    const ulonglong2 tile = offsets[blockIdx.x];
    if (threadIdx.x == 0) {   // MacroBlockStruct, vwsc.h:595-599 / build() :677-681
        macroBlocks[2 * u64(blockIdx.x)] = tile.x;
        macroBlocks[2 * u64(blockIdx.x) + 1] = tile.y;
    }
    u64 blockIndex = tile.x + (excl >> 32);
    u32 bit = u32(excl);                       // weight bit offset relative to the macro block
    const u32 skew = u32(tile.y & 31);         // the macro block's first bit within its first word
#pragma unroll
    for (u32 j = 0; j < kColorsPerThread; ++j) {
        if (j < nValid) {
            const u32 bpw = u32(c[j] >> 32) & 0xFF;
            if (flags & (1u << j)) blocks[blockIndex++] = (u64(u32(c[j])) << 32) | make_block_header(bit, bpw, threadIdx.x * kColorsPerThread + j);
            if (bpw) {   // ColorLeafBuilder::add_weight, vwsc.h:552-580: MSB-first bit stream
                const u32 w = u32(c[j] >> 40) & 0xFF, p = bit + skew, k = p >> 5, o = p & 31;
                if (o + bpw <= 32) {
                    atomicOr(&words[k], w << (32 - o - bpw));
                } else {
                    atomicOr(&words[k], w >> (o + bpw - 32));
                    atomicOr(&words[k + 1], w << (64 - o - bpw));
                }
                bit += bpw;
            }
        }
    }
    __syncthreads();
    const u32 bitsInTile = u32(total);
    if (bitsInTile == 0) return;
    const u32 nWords = (skew + bitsInTile + 31) >> 5;
    u32* out = weights + (tile.y >> 5);
    for (u32 k = threadIdx.x; k < nWords; k += blockDim.x) {
        const u32 v = __byte_perm(words[k], 0, 0x0123);   // ColorUtils::swap_byte_order, build() :671-674
        if (k == 0 || k == nWords - 1) { if (v) atomicOr(out + k, v); }   // words shared with the neighbouring macro blocks
        else out[k] = v;
    }
}

}  // namespace hdt
