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
//   * weights are appended MSB-first to a bit stream kept in 32-bit words, each byte-swapped by build() -- in
//     memory the stream is simply a sequence of bytes, first bit in the top bit of byte 0.
// The reference produces this one colour at a time.  Here the unit of work is a PIECE: a maximal run of colours of the
// new leaf that come from one block of the old leaf (COPY op) or from one FILL op, cut at the macro-block boundaries of
// the new leaf.  Inside a piece (colorBits, bitsPerWeight) is constant, so
//   * a block of the new leaf starts at a piece iff the piece starts a macro block or its key differs from the piece
//     before it (blocks the old leaf had split at ITS macro boundaries, or at an op boundary, merge again this way);
//   * the weights of a COPY piece are a contiguous bit range of the old stream, moved to their new bit position with
//     funnel shifts, 32 bits at a time; a FILL piece is a periodic pattern.
// Work per piece, not per colour: the bench leaf has 15.5 colours per block.  Kernels, one CTA per macro block of the new leaf:
//   1. count_color_pieces_kernel  enumerate the pieces, reduce {blocks started, weight bits};
//   2. scan_color_tiles_kernel    exclusive scan of those pairs (one CTA; a leaf has n/16384 macro blocks);
//   3. emit_color_pieces_kernel   enumerate the pieces again, CTA-wide scan, block headers, macro-block pairs, the weight bits
//                                 assembled in shared memory and stored as whole swapped words (the two words a macro block
//                                 may share with its neighbours go through atomicOr).
// Bit-exact with the reference builder by construction; pinned against leaves the reference built
// (tests/golden/ref_color_leaves_d13.npz, tests/test_gpu_color_leaf.py).
#pragma once
#include "hdt_colors.cuh"

namespace hdt {

constexpr u32 kPieceThreads = 256;
constexpr u32 kLongPieceWords = 12;   // pieces whose weights span more words than this are copied by the whole CTA

// hdt_color_op with the exclusive prefix of the counts (where the op's first colour lands in the new leaf).
struct ColorOpDev { u64 dstStart; u64 srcStart; u32 kind; u32 bitsPerWeight; u32 colorBits; u32 weight; };
static_assert(sizeof(ColorOpDev) == 32, "uploaded as is");

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

__device__ __forceinline__ u32 block_bits_per_weight(u32 hdr) { return ((hdr >> 16) == 0xFFFF) ? 0u : (((hdr >> 14) & 0x3) + 1); }   // vwsc.h:54-75

// Last block of macro block `macro` of the old leaf whose first colour is <= `local` (binary_search_blocks, vwsc.h:343-372).
__device__ __forceinline__ u32 find_block(const ColorLeafDev& l, u32 macro, u32 local, u32& lastBlock)
{
    u32 lo = u32(__ldg(l.macroBlocks + 2 * u64(macro)));
    lastBlock = (2 * (u64(macro) + 1) < l.nMacroWords) ? u32(__ldg(l.macroBlocks + 2 * (u64(macro) + 1)) - 1) : u32(l.nBlocks - 1);
    u32 hi = lastBlock;
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if ((u32(__ldg(l.blocks + mid)) & 0x3FFF) <= local) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Where a piece's weight bits come from: a bit position of the old stream, or (FILL) the first 64 bits of a periodic pattern.
struct BitSource {
    u64 at;        // COPY: absolute bit position in the old leaf's stream; FILL: the pattern, first bit in bit 63
    u32 period;    // 0: COPY; else bitsPerWeight of the FILL
};

// `take` (1..32) bits starting `skip` bits into the piece, left-aligned in the result.
__device__ __forceinline__ u32 read_bits(const ColorLeafDev& l, const BitSource& src, u32 skip, u32 take)
{
    u32 v;
    if (src.period) {
        v = u32((src.at << (skip % src.period)) >> 32);
    } else {
        const u64 p = src.at + skip;
        const u64 wi = p >> 5;
        const u32 sh = u32(p) & 31;
        const u32 hi = wi < l.nWeights ? __byte_perm(__ldg(l.weights + wi), 0, 0x0123) : 0u;
        const u32 lo = (sh + take > 32 && wi + 1 < l.nWeights) ? __byte_perm(__ldg(l.weights + wi + 1), 0, 0x0123) : 0u;
        v = __funnelshift_l(lo, hi, sh);
    }
    return v & (0xFFFFFFFFu << (32 - take));
}

// Bits [q, q + n) of the macro block's stream (`words`, shared, MSB-first, zeroed) <- the piece's bits; the destination words
// j = first, first + stride, ... of the range are handled by the caller (one thread: 0, 1; the CTA: threadIdx.x, blockDim.x).
__device__ __forceinline__ void write_bits(u32* words, const ColorLeafDev& l, const BitSource& src, u32 q, u32 n, u32 first, u32 stride)
{
    const u32 w0 = q >> 5, nWords = ((q + n - 1) >> 5) - w0 + 1;
    for (u32 j = first; j < nWords; j += stride) {
        const u32 ws = (w0 + j) << 5;
        const u32 a = max(q, ws), b = min(q + n, ws + 32);
        const u32 v = read_bits(l, src, a - q, b - a);
        if (b - a == 32) words[w0 + j] = v;          // the word belongs to this piece alone
        else atomicOr(&words[w0 + j], v >> (a & 31));
    }
}

// One CTA per macro block of the new leaf.  EMIT = false: tiles[blockIdx] = {blocks started, weight bits}.  EMIT = true: the
// macro block is written (offsets[blockIdx] = its first block index and weight bit offset).
template<bool EMIT>
__global__ void __launch_bounds__(kPieceThreads) color_pieces_kernel(const ColorOpDev* __restrict__ ops, const u32 nOps, const ColorLeafDev oldLeaf, const u64 nColors,
                                                                      TilePair* __restrict__ tiles, const ulonglong2* __restrict__ offsets,
                                                                      u32* __restrict__ weights, u64* __restrict__ blocks, u64* __restrict__ macroBlocks)
{
    constexpr u32 T = kPieceThreads;
    // the segments (op x this macro block) of the current chunk, one per thread
    __shared__ u32 segPieceStart[T + 1];    // exclusive prefix of the pieces per segment
    __shared__ u32 segDst[T], segLen[T], segBlock0[T], segLast0[T];
    __shared__ u64 segSrc[T];
    __shared__ u64 pieceKey[T];
    // a macro block holds at most 16384 * 4 weight bits = 2048 words, + 1 for the misaligned start
    __shared__ u32 words[EMIT ? kColorsPerMacroBlock * 4 / 32 + 2 : 1];
    __shared__ u32 longCount;
    __shared__ u32 longQ[EMIT ? T : 1], longN[EMIT ? T : 1], longPeriod[EMIT ? T : 1];
    __shared__ u64 longAt[EMIT ? T : 1];

    const u32 t = threadIdx.x;
    const u64 d0 = u64(blockIdx.x) * kColorsPerMacroBlock, d1 = min(d0 + kColorsPerMacroBlock, nColors);
    // ops that reach into [d0, d1): lo = last op starting at or before d0, hi = last op starting before d1
    u32 lo = 0, hi = nOps - 1;
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if (__ldg(&ops[mid].dstStart) <= d0) lo = mid; else hi = mid - 1;
    }
    u32 last = lo;
    hi = nOps - 1;
    while (last < hi) {
        const u32 mid = (last + hi + 1) >> 1;
        if (__ldg(&ops[mid].dstStart) < d1) last = mid; else hi = mid - 1;
    }
    const u32 nSeg = last - lo + 1;
    const u64 sharedOffset = oldLeaf.is_shared() ? oldLeaf.offset : 0;

    ulonglong2 tile = make_ulonglong2(0, 0);
    u32 skew = 0;
    if (EMIT) {
        tile = offsets[blockIdx.x];
        skew = u32(tile.y & 31);
        for (u32 k = t; k < sizeof(words) / 4; k += T) words[k] = 0;
        if (t == 0) {   // MacroBlockStruct, vwsc.h:595-599 / build() :677-681
            macroBlocks[2 * u64(blockIdx.x)] = tile.x;
            macroBlocks[2 * u64(blockIdx.x) + 1] = tile.y;
        }
    }
    u32 carryBlocks = 0, carryBits = 0;          // blocks started / weight bits of this macro block so far (CTA-uniform)
    u64 carryKey = 0;                            // key of the last piece so far

    for (u32 segBase = 0; segBase < nSeg; segBase += T) {
        const u32 nChunk = min(T, nSeg - segBase);
        // ---- segments of this chunk ----
        u32 myPieces = 0;
        if (t < nChunk) {
            const ColorOpDev op = ops[lo + segBase + t];
            const u64 a = max(op.dstStart, d0), b = min(__ldg(&ops[lo + segBase + t + 1].dstStart), d1);
            segDst[t] = u32(a - d0);
            segLen[t] = u32(b - a);
            myPieces = 1;
            if (op.kind == HDT_COLOR_OP_COPY) {
                const u64 s0 = op.srcStart + (a - op.dstStart) + sharedOffset, s1 = s0 + (b - a) - 1;
                u32 last0, last1;
                const u32 b0 = find_block(oldLeaf, u32(s0 / kColorsPerMacroBlock), u32(s0 % kColorsPerMacroBlock), last0);
                const u32 b1 = find_block(oldLeaf, u32(s1 / kColorsPerMacroBlock), u32(s1 % kColorsPerMacroBlock), last1);
                segSrc[t] = s0; segBlock0[t] = b0; segLast0[t] = last0;
                myPieces = b1 - b0 + 1;
            }
        }
        u64 totalPieces64;
        const u32 myStart = u32(cta_exclusive_scan(myPieces, totalPieces64));
        const u32 totalPieces = u32(totalPieces64);
        if (t < nChunk) segPieceStart[t] = myStart;
        if (t == 0) { segPieceStart[nChunk] = totalPieces; if (EMIT) longCount = 0; }
        __syncthreads();
        // ---- pieces of this chunk, T at a time ----
        for (u32 pBase = 0; pBase < totalPieces; pBase += T) {
            const u32 p = pBase + t;
            const bool valid = p < totalPieces;
            u64 key = 0;
            u32 dstLocal = 0, len = 0, bpw = 0;
            BitSource src{ 0, 0 };
            if (valid) {
                u32 s = 0, e = nChunk - 1;       // segment of piece p: last one starting at or before p
                while (s < e) {
                    const u32 mid = (s + e + 1) >> 1;
                    if (segPieceStart[mid] <= p) s = mid; else e = mid - 1;
                }
                const ColorOpDev op = ops[lo + segBase + s];
                if (op.kind == HDT_COLOR_OP_COPY) {
                    const u32 b = segBlock0[s] + (p - segPieceStart[s]);
                    const u64 s0 = segSrc[s], s1 = s0 + segLen[s];
                    const u64 macro = s0 / kColorsPerMacroBlock + (b > segLast0[s] ? 1 : 0);   // a segment spans at most two old macro blocks
                    const u64 blk = __ldg(oldLeaf.blocks + b);
                    const u32 hdr = u32(blk), startLocal = hdr & 0x3FFF;
                    const u32 nextLocal = (u64(b) + 1 < oldLeaf.nBlocks) ? (u32(__ldg(oldLeaf.blocks + b + 1)) & 0x3FFF) : 0u;
                    const u64 blockStart = macro * kColorsPerMacroBlock + startLocal;
                    // starts increase inside a macro block: a smaller or equal one belongs to the next macro block
                    const u64 blockEnd = macro * kColorsPerMacroBlock + (nextLocal > startLocal ? nextLocal : u32(kColorsPerMacroBlock));
                    const u64 ps = max(blockStart, s0), pe = min(blockEnd, s1);
                    bpw = block_bits_per_weight(hdr);
                    key = (blk >> 32) | (u64(bpw) << 32);
                    dstLocal = segDst[s] + u32(ps - s0);
                    len = u32(pe - ps);
                    if (EMIT && bpw) src.at = __ldg(oldLeaf.macroBlocks + 2 * macro + 1) + (hdr >> 16) + (ps - blockStart) * bpw;
                } else {
                    bpw = op.bitsPerWeight;
                    key = u64(op.colorBits) | (u64(bpw) << 32);
                    dstLocal = segDst[s];
                    len = segLen[s];
                    if (EMIT && bpw) {           // the stream of a repeated weight: its first 64 bits
                        u64 pat = 0;
                        for (u32 k = 0; k < 64; k += bpw) pat |= (u64(op.weight) << (64 - bpw)) >> k;
                        src.at = pat; src.period = bpw;
                    }
                }
            }
            pieceKey[t] = key;
            __syncthreads();
            const u64 prevKey = t ? pieceKey[t - 1] : carryKey;
            const bool starts = valid && (dstLocal == 0 || key != prevKey);   // ColorLeafBuilder::add, vwsc.h:606
            const u32 bits = len * bpw;
            u64 total;
            const u64 excl = cta_exclusive_scan((u64(starts ? 1u : 0u) << 32) | bits, total);
            const u32 nValid = min(T, totalPieces - pBase);
            const u64 lastKey = pieceKey[nValid - 1];
            if (EMIT && valid) {
                const u32 bit = carryBits + u32(excl);     // weight bit offset relative to the macro block
                if (starts) blocks[tile.x + carryBlocks + u32(excl >> 32)] = (u64(u32(key)) << 32) | make_block_header(bit, bpw, dstLocal);
                if (bits) {
                    const u32 q = bit + skew;
                    if (((q + bits - 1) >> 5) - (q >> 5) + 1 > kLongPieceWords) {
                        const u32 k = atomicAdd(&longCount, 1u);
                        longQ[k] = q; longN[k] = bits; longAt[k] = src.at; longPeriod[k] = src.period;
                    } else {
                        write_bits(words, oldLeaf, src, q, bits, 0, 1);
                    }
                }
            }
            carryBlocks += u32(total >> 32);
            carryBits += u32(total);
            carryKey = lastKey;
            __syncthreads();                                // pieceKey is rewritten by the next round; longCount is complete
            if (EMIT) {
                const u32 nLong = longCount;
                for (u32 k = 0; k < nLong; ++k) write_bits(words, oldLeaf, BitSource{ longAt[k], longPeriod[k] }, longQ[k], longN[k], t, T);
                __syncthreads();
                if (t == 0) longCount = 0;
            }
        }
        __syncthreads();                                    // the segment arrays are rewritten by the next chunk
    }
    if (!EMIT) {
        if (t == 0) tiles[blockIdx.x] = TilePair{ carryBlocks, carryBits };
        return;
    }
    __syncthreads();
    if (carryBits == 0) return;
    const u32 nWords = (skew + carryBits + 31) >> 5;
    u32* out = weights + (tile.y >> 5);
    for (u32 k = t; k < nWords; k += T) {
        const u32 v = __byte_perm(words[k], 0, 0x0123);   // ColorUtils::swap_byte_order, build() :671-674
        if (k == 0 || k == nWords - 1) { if (v) atomicOr(out + k, v); }   // words shared with the neighbouring macro blocks
        else out[k] = v;
    }
}

}  // namespace hdt
