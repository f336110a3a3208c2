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
//   * the weights need not be moved piece by piece: the pieces of a COPY segment (op x new macro block) are consecutive blocks of
//     the old leaf, so its weights are one contiguous bit range of the old stream per old macro block it touches (at most two),
//     moved with funnel shifts one destination word at a time; a FILL segment is a periodic pattern.
// Work per piece, not per colour: the bench leaf has 15.5 colours per block.  Kernels:
//   1. color_segments_kernel     one warp per macro block of the new leaf: the ops that reach into it, and for every COPY
//                                segment the old blocks of its first and last colour (the searches);
//   2. color_pieces_kernel       one CTA per macro block: the pieces, once -- CTA-wide scans of {blocks started, weight bits},
//                                finished block entries (they only hold positions relative to the macro block) staged in
//                                scratch memory, every segment's weight range, the macro block's {blocks, bits};
//   3. scan_color_tiles_kernel   first block index and weight bit offset of every macro block;
//   4. color_emit_kernel         one CTA per macro block, nothing per piece: staged entries to their place, macro-block pairs,
//                                weight words (the words two macro blocks share go through atomicOr).
// Bit-exact with the reference builder by construction; pinned against leaves the reference built
// (tests/golden/ref_color_leaves_d13.npz, tests/test_gpu_color_leaf.py).
#pragma once
#include "hdt_colors.cuh"

namespace hdt {

#ifndef HDT_PIECE_THREADS
#define HDT_PIECE_THREADS 128
#endif
#ifndef HDT_PIECES_PER_THREAD
#define HDT_PIECES_PER_THREAD 6
#endif
#ifndef HDT_PIECE_MIN_BLOCKS
#define HDT_PIECE_MIN_BLOCKS 12     // <= 42 registers: the pass is latency bound, residency pays (profiles/r2_ab_color_leaf*.jsonl)
#endif
constexpr u32 kPieceThreads = HDT_PIECE_THREADS;
constexpr u32 kPiecesPerThread = HDT_PIECES_PER_THREAD;

// hdt_color_op with the exclusive prefix of the counts (where the op's first colour lands in the new leaf).
struct ColorOpDev { u64 dstStart; u64 srcStart; u32 kind; u32 bitsPerWeight; u32 colorBits; u32 weight; };
static_assert(sizeof(ColorOpDev) == 32, "uploaded as is");

struct TilePieces { u32 stageBase, blocks, bits, pad; };   // per macro block: where its finished block entries are staged, how many, weight bits

// CTA-wide exclusive scan with ONE barrier: every warp leaves its sum in shared memory, every thread adds up the (few) warps in
// front of it.  Two buffers, used alternately (`phase`, a per-thread counter all threads advance together): a warp can only
// come back to a buffer after the barrier of the call in between, which every warp reaches after it has read this one.
template<typename V>
__device__ __forceinline__ V cta_exclusive_scan(V v, V& total, u32& phase)
{
    __shared__ V warpSums[2][32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    V inc = v;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        const V o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    V* sums = warpSums[phase & 1];
    ++phase;
    if (lane == 31) sums[warp] = inc;
    __syncthreads();
    V base = V(0), all = V(0);
    for (u32 w = 0; w < nWarps; ++w) {
        const V x = sums[w];
        if (w < warp) base += x;
        all += x;
    }
    total = all;
    return base + inc - v;
}

// tiles[m] (counts) -> offsets[m] (exclusive prefix: first block index, weight bit offset; offsets[nTiles] = the totals) and
// totals[0..1].  The count pass has added every tile's pair to groupSums[m / 256] (blocks << 32 | bits: a leaf has fewer than
// 2^30 blocks and 2^32 bits); one CTA per group of 256 tiles adds the groups in front of it and scans its own tiles.
// The weight words two macro blocks share (the word a macro block's first bit falls into) are put together with atomicOr by
// the emit pass: they are zeroed here, if the caller's buffer reaches that far.
constexpr u32 kTileGroup = 256;
__global__ void __launch_bounds__(kTileGroup) scan_color_tiles_kernel(const TilePieces* __restrict__ tiles, const u32 nTiles, const u64* __restrict__ groupSums,
                                                                       ulonglong2* __restrict__ offsets, u64* __restrict__ totals, u32* __restrict__ weights, const u64 weightsCapacity)
{
    const u32 t = threadIdx.x, g = blockIdx.x, m = g * kTileGroup + t;
    const TilePieces mine = m < nTiles ? tiles[m] : TilePieces{ 0, 0, 0, 0 };
    u64 front = t < g ? groupSums[t] : 0;            // at most 65536 / 256 = 256 groups
    u64 base;
    u32 phase = 0;
    cta_exclusive_scan(front, base, phase);          // (only the total is used)
    u64 total;
    const u64 at = base + cta_exclusive_scan((u64(mine.blocks) << 32) | mine.bits, total, phase);
    if (m < nTiles) {
        offsets[m] = make_ulonglong2(at >> 32, at & 0xFFFFFFFFull);
        const u64 word = (at & 0xFFFFFFFFull) >> 5;
        if (weights && word < weightsCapacity) weights[word] = 0;
    }
    if (m + 1 == nTiles) {
        const u64 end = at + ((u64(mine.blocks) << 32) | mine.bits);
        totals[0] = end >> 32; totals[1] = end & 0xFFFFFFFFull;
        offsets[nTiles] = make_ulonglong2(end >> 32, end & 0xFFFFFFFFull);
        if (weights && ((end & 0xFFFFFFFFull) >> 5) < weightsCapacity) weights[(end & 0xFFFFFFFFull) >> 5] = 0;
    }
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
    if (local == 0) return lo;                                   // a macro block's first colour is in its first block,
    if (local == kColorsPerMacroBlock - 1) return hi;            // its last colour in its last (copies of whole macro blocks search nothing)
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if ((u32(__ldg(l.blocks + mid)) & 0x3FFF) <= local) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// The same search by a whole warp, 32 probes at a time: two or three dependent loads instead of ten.  All lanes get the result.
__device__ __forceinline__ u32 find_block_warp(const ColorLeafDev& l, u32 macro, u32 local, u32& lastBlock)
{
    const u32 lane = threadIdx.x & 31;
    u32 lo = u32(__ldg(l.macroBlocks + 2 * u64(macro)));
    lastBlock = (2 * (u64(macro) + 1) < l.nMacroWords) ? u32(__ldg(l.macroBlocks + 2 * (u64(macro) + 1)) - 1) : u32(l.nBlocks - 1);
    u32 hi = lastBlock;                         // blocks[lo] starts at colour 0 <= local: the answer is in [lo, hi]
    if (local == 0) return lo;
    if (local == kColorsPerMacroBlock - 1) return hi;
    while (hi - lo >= 32) {
        const u32 step = (hi - lo) / 32;        // probes lo + step, lo + 2 step, ... lo + 32 step (<= hi)
        const u32 probe = lo + (lane + 1) * step;
        const u32 below = __popc(__ballot_sync(0xFFFFFFFFu, (u32(__ldg(l.blocks + probe)) & 0x3FFF) <= local));   // starts increase: a prefix of the lanes
        const u32 newLo = lo + below * step;
        hi = below == 32 ? hi : newLo + step - 1;
        lo = newLo;
    }
    const bool ok = lo + lane <= hi && (u32(__ldg(l.blocks + lo + lane)) & 0x3FFF) <= local;
    return lo + __popc(__ballot_sync(0xFFFFFFFFu, ok)) - 1;
}

// A segment: the part of one op that lands in one macro block of the new leaf.  Written once by color_segments_kernel (the
// searches in the old leaf are its dependent loads), read by the count and the emit pass.
struct SegmentDev {
    u32 dstLocal, len;          // first colour within the new macro block, colours
    u32 nPieces;
    u32 fill;                   // FILL: 1 | bitsPerWeight << 8 | weight << 16;  COPY: 0
    u32 colorBits;              // FILL
    u32 block0, last0;          // COPY: old block holding the first colour; last block of that block's macro block
    u32 src0;                   // COPY: the first colour's index within its old macro block
    u64 weightBase[2];          // COPY: weight bit offsets of that old macro block and of the one after it
};
static_assert(sizeof(SegmentDev) == 48, "three 16-byte loads");
struct TileSegments { u32 first, count; };   // macro block of the new leaf -> its segments

// The ops that reach into macro block m are ops[lo(m) .. last(m)]; their segments live in slots lo(m) + m + k (distinct for
// every (m, k): lo(m + 1) >= last(m), so the slot ranges of consecutive macro blocks do not overlap).  One warp per macro block.
__global__ void __launch_bounds__(256) color_segments_kernel(const ColorOpDev* __restrict__ ops, const u32 nOps, const ColorLeafDev oldLeaf, const u64 nColors,
                                                              const u32 nTiles, SegmentDev* __restrict__ segs, TileSegments* __restrict__ tileSegs)
{
    const u32 m = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (m >= nTiles) return;
    const u64 d0 = u64(m) * kColorsPerMacroBlock, d1 = min(d0 + kColorsPerMacroBlock, nColors);
    // lo = last op starting at or before d0, last = last op starting before d1
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
    if (lane == 0) tileSegs[m] = TileSegments{ lo + m, nSeg };
    const u64 sharedOffset = oldLeaf.is_shared() ? oldLeaf.offset : 0;
    const bool together = nSeg <= 4;            // a few segments: the warp searches each one with 32 probes at a time
    for (u32 k = together ? 0 : lane; k < nSeg; k += together ? 1 : 32) {
        const ColorOpDev op = ops[lo + k];
        const u64 a = max(op.dstStart, d0), b = min(__ldg(&ops[lo + k + 1].dstStart), d1);
        SegmentDev sg{};
        sg.dstLocal = u32(a - d0); sg.len = u32(b - a);
        sg.nPieces = 1;
        if (op.kind == HDT_COLOR_OP_COPY) {
            const u64 s0 = op.srcStart + (a - op.dstStart) + sharedOffset, s1 = s0 + (b - a) - 1;
            const u32 m0 = u32(s0 / kColorsPerMacroBlock), m1 = u32(s1 / kColorsPerMacroBlock);
            u32 last1, b1;
            if (together) {
                sg.block0 = find_block_warp(oldLeaf, m0, u32(s0 % kColorsPerMacroBlock), sg.last0);
                b1 = find_block_warp(oldLeaf, m1, u32(s1 % kColorsPerMacroBlock), last1);
            } else {
                sg.block0 = find_block(oldLeaf, m0, u32(s0 % kColorsPerMacroBlock), sg.last0);
                b1 = find_block(oldLeaf, m1, u32(s1 % kColorsPerMacroBlock), last1);
            }
            sg.nPieces = b1 - sg.block0 + 1;
            sg.src0 = u32(s0 % kColorsPerMacroBlock);
            sg.weightBase[0] = __ldg(oldLeaf.macroBlocks + 2 * u64(m0) + 1);
            sg.weightBase[1] = m1 != m0 ? __ldg(oldLeaf.macroBlocks + 2 * u64(m1) + 1) : 0;
        } else {
            sg.fill = 1u | (op.bitsPerWeight << 8) | (op.weight << 16);
            sg.colorBits = op.colorBits;
        }
        if (!together || lane == 0) segs[lo + m + k] = sg;
    }
}

// `take` (1..32) bits of the old leaf's weight stream from bit position `p`, left-aligned in the result.
__device__ __forceinline__ u32 read_stream_bits(const ColorLeafDev& l, u64 p, u32 take)
{
    const u64 wi = p >> 5;
    const u32 sh = u32(p) & 31;
    const u32 hi = wi < l.nWeights ? __byte_perm(__ldg(l.weights + wi), 0, 0x0123) : 0u;
    const u32 lo = (sh + take > 32 && wi + 1 < l.nWeights) ? __byte_perm(__ldg(l.weights + wi + 1), 0, 0x0123) : 0u;
    return __funnelshift_l(lo, hi, sh) & (0xFFFFFFFFu << (32 - take));
}

// A piece as a thread keeps it between the two halves of a round: colorBits, and
// first colour within the new macro block (14 bits) | colours << 14 (15 bits) | bitsPerWeight << 29.
__device__ __forceinline__ u32 pack_piece(u32 dstLocal, u32 bpw, u32 len) { return dstLocal | (len << 14) | (bpw << 29); }
__device__ __forceinline__ u32 piece_dst(u32 pk) { return pk & 0x3FFF; }
__device__ __forceinline__ u32 piece_bpw(u32 pk) { return pk >> 29; }
__device__ __forceinline__ u32 piece_bits(u32 pk) { return ((pk >> 14) & 0x7FFF) * piece_bpw(pk); }
__device__ __forceinline__ u64 piece_key(u32 colorBits, u32 pk) { return u64(colorBits) | (u64(piece_bpw(pk)) << 32); }   // equal keys continue a block

// Piece of a COPY segment that comes from block `b` of the old leaf (`blk`; `nextLocal` = first colour of block b + 1 within its
// macro block).  Colour indices are relative to the old macro block of the segment's first colour: the segment is
// [src0, src0 + len) there, at most 16384 long, so it ends before 2 * 16384.  -> packed piece; skipped = colours of the block in
// front of the piece (piece_src_at: where the piece's first weight bit sits in the old stream).
__device__ __forceinline__ u64 piece_src_at(const SegmentDev& sg, u32 b, u32 hdr, u32 skipped)
{
    return (b > sg.last0 ? sg.weightBase[1] : sg.weightBase[0]) + (hdr >> 16) + skipped * block_bits_per_weight(hdr);
}
__device__ __forceinline__ u32 copy_piece(const SegmentDev& sg, u32 b, u64 blk, u32 nextLocal, u32& skipped)
{
    const u32 second = b > sg.last0 ? 1u : 0u;
    const u32 hdr = u32(blk), startLocal = hdr & 0x3FFF;
    const u32 base = second * u32(kColorsPerMacroBlock);
    // starts increase inside a macro block: a smaller or equal one belongs to the next macro block
    const u32 blockStart = base + startLocal, blockEnd = base + (nextLocal > startLocal ? nextLocal : u32(kColorsPerMacroBlock));
    const u32 ps = max(blockStart, sg.src0), pe = min(blockEnd, sg.src0 + sg.len);
    const u32 bpw = block_bits_per_weight(hdr);
    skipped = ps - blockStart;
    return pack_piece(sg.dstLocal + (ps - sg.src0), bpw, pe - ps);
}

// Where a segment's weight bits go and come from, relative to the new macro block: bits [bit0, mid) are the old stream's
// bits from (bit + delta[0]) on, bits [mid, bit1) from (bit + delta[1]) on -- a segment touches at most two macro blocks of the
// old leaf, and the old stream need not be contiguous across them (the format lets a builder pad there).  FILL: `fill` as in
// SegmentDev, a periodic pattern.  Written by the pieces pass, read by the emit pass.
struct SegmentWeights { u32 bit0, mid, bit1, fill; u64 delta[2]; };
static_assert(sizeof(SegmentWeights) == 32, "two 16-byte loads");

// Pass 1, one CTA per macro block of the new leaf: everything that is per piece, once.  The block entries of a macro block do not
// depend on where the macro block ends up (weight offset and colour index are relative to it), so they are finished here and
// staged (`stage`, space taken from `stageTop`); so are its segments' weight ranges.  What is still unknown -- the macro block's
// first block index and weight bit offset -- comes from the scan of the {blocks, bits} pairs this pass leaves in `tiles`.
//
// The segments of the macro block are taken T at a time (a "chunk", one per thread), the pieces of a chunk T*K at a time (a
// "round": K consecutive pieces per thread, one CTA-wide scan of {blocks started, weight bits}).
__global__ void __launch_bounds__(kPieceThreads, HDT_PIECE_MIN_BLOCKS) color_pieces_kernel(const SegmentDev* __restrict__ segs, const TileSegments* __restrict__ tileSegs, const ColorLeafDev oldLeaf,
                                                                      TilePieces* __restrict__ tiles, u64* __restrict__ groupSums, u64* __restrict__ stage, u32* __restrict__ stageTop,
                                                                      SegmentWeights* __restrict__ segWeights)
{
    constexpr u32 T = kPieceThreads, K = kPiecesPerThread;
    __shared__ SegmentDev seg[T];
    __shared__ u32 segPieceStart[T + 1];    // exclusive prefix of the pieces per segment
    __shared__ u64 lastKeyOf[T];            // key of each thread's last piece of the round
    __shared__ u32 segBit0[T], segMid[T], segBit1[T];
    __shared__ u64 segDelta[2][T];
    __shared__ u32 stageBaseShared;

    const u32 t = threadIdx.x;
    const TileSegments mine = tileSegs[blockIdx.x];
    const u32 nSeg = mine.count;
    u32 carryBlocks = 0, carryBits = 0;          // blocks started / weight bits of this macro block so far (CTA-uniform)
    u32 scanPhase = 0;
    u64 carryKey = ~u64(0);                      // key of the last piece so far; no piece has this one: the first piece starts a block
    {   // room for the macro block's block entries: at most one per piece
        u32 pieces = 0;
        for (u32 k = t; k < nSeg; k += T) pieces += segs[mine.first + k].nPieces;
        u32 total;
        cta_exclusive_scan(pieces, total, scanPhase);
        if (t == 0) stageBaseShared = atomicAdd(stageTop, total);
        __syncthreads();
    }
    const u32 stageBase = stageBaseShared;

    for (u32 segBase = 0; segBase < nSeg; segBase += T) {
        const u32 nChunk = min(T, nSeg - segBase);
        // ---- segments of this chunk ----
        u32 myPieces = 0;
        if (t < nChunk) {
            const SegmentDev sg = segs[mine.first + segBase + t];
            seg[t] = sg;
            myPieces = sg.nPieces;
        }
        u32 totalPieces;
        const u32 myStart = cta_exclusive_scan(myPieces, totalPieces, scanPhase);
        if (t < nChunk) segPieceStart[t] = myStart;
        if (t == 0) segPieceStart[nChunk] = totalPieces;
        __syncthreads();
        const bool oneCopy = nChunk == 1 && !seg[0].fill;   // the usual case: the macro block is one COPY segment
        // ---- pieces of this chunk, T * K at a time ----
        for (u32 pBase = 0; pBase < totalPieces; pBase += T * K) {
            const u32 p0 = pBase + t * K;
            const u32 nMine = p0 >= totalPieces ? 0u : min(K, totalPieces - p0);
            u32 s0idx = 0;                       // segment of piece p0: last one starting at or before it
            if (nChunk > 1 && nMine) {
                u32 e = nChunk - 1;
                while (s0idx < e) {
                    const u32 mid = (s0idx + e + 1) >> 1;
                    if (segPieceStart[mid] <= p0) s0idx = mid; else e = mid - 1;
                }
            }
            u32 cb[K], pk[K];                    // the thread's pieces: colorBits, pack_piece()
            u32 myBits = 0;
            // oneCopy: where the old stream continues for the thread's first weighted piece of each half of the segment
            // (position of that piece's first bit minus the thread's bits in front of it), and which halves it has
            u64 at0 = 0, at1 = 0;
            u32 haveAt = 0;
#pragma unroll
            for (u32 j = 0; j < K; ++j) { cb[j] = 0; pk[j] = 0; }
            if (oneCopy) {
                // the thread's pieces are K consecutive blocks of the old leaf (+ the start of the one after them)
                const SegmentDev sg = seg[0];
                const u32 b = sg.block0 + p0;
                u64 blk[K + 1];
                const u32 nLoad = u32(min(u64(nMine) + 1, oldLeaf.nBlocks - b));   // (b < nBlocks: the thread has pieces)
#pragma unroll
                for (u32 j = 0; j <= K; ++j) blk[j] = j < nLoad ? __ldg(oldLeaf.blocks + b + j) : 0;
#pragma unroll
                for (u32 j = 0; j < K; ++j) {
                    if (j < nMine) {
                        u32 skipped;
                        pk[j] = copy_piece(sg, b + j, blk[j], u32(blk[j + 1]) & 0x3FFF, skipped);
                        cb[j] = u32(blk[j] >> 32);
                        const u32 bits = piece_bits(pk[j]), half = b + j > sg.last0 ? 1u : 0u;
                        if (bits && !(haveAt & (1u << half))) {
                            const u64 srcAt = piece_src_at(sg, b + j, u32(blk[j]), skipped);
                            if (half) at1 = srcAt - myBits; else at0 = srcAt - myBits;
                            haveAt |= 1u << half;
                        }
                        myBits += bits;
                    }
                }
            } else {
                u32 s = s0idx, segEnd = 0;       // pieces below segEnd belong to segment s, whose fields are in sg / pieceStart
                SegmentDev sg{};
                u32 pieceStart = 0;
#pragma unroll
                for (u32 j = 0; j < K; ++j) {
                    if (j < nMine) {
                        const u32 p = p0 + j;
                        if (p >= segEnd) {
                            if (segEnd) ++s;
                            while (p >= segPieceStart[s + 1]) ++s;
                            sg = seg[s]; pieceStart = segPieceStart[s]; segEnd = segPieceStart[s + 1];
                        }
                        if (!sg.fill) {
                            const u32 b = sg.block0 + (p - pieceStart);
                            const u64 blk = __ldg(oldLeaf.blocks + b);
                            const u32 nextLocal = (u64(b) + 1 < oldLeaf.nBlocks) ? (u32(__ldg(oldLeaf.blocks + b + 1)) & 0x3FFF) : 0u;
                            u32 skipped;
                            pk[j] = copy_piece(sg, b, blk, nextLocal, skipped);
                            cb[j] = u32(blk >> 32);
                        } else {
                            pk[j] = pack_piece(sg.dstLocal, (sg.fill >> 8) & 0xFF, sg.len);
                            cb[j] = sg.colorBits;
                        }
                        myBits += piece_bits(pk[j]);
                    }
                }
            }
            if (nMine) {
                u32 lastCb = cb[0], lastPk = pk[0];
#pragma unroll
                for (u32 j = 1; j < K; ++j) if (j < nMine) { lastCb = cb[j]; lastPk = pk[j]; }
                lastKeyOf[t] = piece_key(lastCb, lastPk);
            }
            __syncthreads();
            const u64 prevKey = t ? lastKeyOf[t - 1] : carryKey;      // (threads in front of a thread that has pieces have K each)
            u32 prevCb = u32(prevKey), prevBpw = u32(prevKey >> 32);
            u32 startsMask = 0;
#pragma unroll
            for (u32 j = 0; j < K; ++j) {
                const u32 bpw = piece_bpw(pk[j]);
                if (j < nMine && (cb[j] != prevCb || bpw != prevBpw)) startsMask |= 1u << j;   // ColorLeafBuilder::add, vwsc.h:606 (a macro block's first piece: carryKey)
                prevCb = cb[j]; prevBpw = bpw;
            }
            // {blocks started, weight bits} in one word: a macro block has at most 65536 bits, a round at most T * K pieces
            static_assert(T * K < 4096, "12 bits for the blocks a round starts");
            u32 total;
            const u32 excl = cta_exclusive_scan((u32(__popc(startsMask)) << 20) | myBits, total, scanPhase);
            const u32 nInRound = min(T * K, totalPieces - pBase);
            const u64 roundLastKey = lastKeyOf[(nInRound - 1) / K];
            if (nMine) {
                u64* out = stage + stageBase + carryBlocks + (excl >> 20);
                u32 bit = carryBits + (excl & 0xFFFFF);     // weight bit offset relative to the macro block
                if (oneCopy) {
                    const u32 b = seg[0].block0 + p0, last0 = seg[0].last0;
                    if (haveAt & 1) segDelta[0][0] = at0 - bit;                       // the same from every thread that has one
                    if (haveAt & 2) segDelta[1][0] = at1 - bit;
                    if (p0 == 0) segBit0[0] = bit;
#pragma unroll
                    for (u32 j = 0; j < K; ++j) {
                        if (j < nMine) {
                            if (startsMask & (1u << j)) *out++ = (u64(cb[j]) << 32) | make_block_header(bit, piece_bpw(pk[j]), piece_dst(pk[j]));
                            if (b + j == last0 + 1) segMid[0] = bit;                  // the first piece from the second old macro block
                            bit += piece_bits(pk[j]);
                        }
                    }
                    if (p0 + nMine == totalPieces) {
                        segBit1[0] = bit;
                        if (b + nMine - 1 <= last0) segMid[0] = bit;                  // no second half
                    }
                } else {
                    u32 s = s0idx, segEnd = 0, pieceStart = 0;
                    SegmentDev sg{};
#pragma unroll
                    for (u32 j = 0; j < K; ++j) {
                        if (j < nMine) {
                            const u32 p = p0 + j;
                            if (p >= segEnd) {
                                if (segEnd) ++s;
                                while (p >= segPieceStart[s + 1]) ++s;
                                sg = seg[s]; pieceStart = segPieceStart[s]; segEnd = segPieceStart[s + 1];
                            }
                            const u32 bits = piece_bits(pk[j]);
                            if (startsMask & (1u << j)) *out++ = (u64(cb[j]) << 32) | make_block_header(bit, piece_bpw(pk[j]), piece_dst(pk[j]));
                            u32 half = 0;
                            if (!sg.fill) {
                                const u32 b = sg.block0 + (p - pieceStart);
                                half = b > sg.last0 ? 1u : 0u;
                                if (b == sg.last0 + 1) segMid[s] = bit;
                                if (bits) {      // where the old stream continues for this half: the same from every piece of it
                                    const u64 blk = __ldg(oldLeaf.blocks + b);
                                    const u32 nextLocal = (u64(b) + 1 < oldLeaf.nBlocks) ? (u32(__ldg(oldLeaf.blocks + b + 1)) & 0x3FFF) : 0u;
                                    u32 skipped;
                                    copy_piece(sg, b, blk, nextLocal, skipped);
                                    segDelta[half][s] = piece_src_at(sg, b, u32(blk), skipped) - bit;
                                }
                            }
                            if (p == pieceStart) segBit0[s] = bit;
                            bit += bits;
                            if (p + 1 == segEnd) {
                                segBit1[s] = bit;
                                if (!half) segMid[s] = bit;                           // no second half
                            }
                        }
                    }
                }
            }
            carryBlocks += total >> 20;
            carryBits += total & 0xFFFFF;
            carryKey = roundLastKey;
            __syncthreads();                                // lastKeyOf is rewritten by the next round; the chunk's weight ranges are complete
        }
        if (t < nChunk) segWeights[mine.first + segBase + t] = SegmentWeights{ segBit0[t], segMid[t], segBit1[t], seg[t].fill, { segDelta[0][t], segDelta[1][t] } };
        __syncthreads();                                    // the segment arrays are rewritten by the next chunk
    }
    if (t == 0) {
        tiles[blockIdx.x] = TilePieces{ stageBase, carryBlocks, carryBits, 0 };
        atomicAdd(reinterpret_cast<unsigned long long*>(groupSums + blockIdx.x / kTileGroup), (u64(carryBlocks) << 32) | carryBits);
    }
}

// Pass 2, one CTA per macro block, nothing per piece: the staged block entries go to their place, the macro-block pair is written,
// and the weights are moved: every destination word of the macro block is put together by one thread from the segments that
// overlap it (a COPY segment's weights are a contiguous bit range of the old stream per old macro block: funnel shifts) and
// stored once; the words shared with a neighbouring macro block or chunk go through atomicOr (zeroed by the scan / below).
#ifndef HDT_EMIT_THREADS
#define HDT_EMIT_THREADS 128
#endif
constexpr u32 kEmitThreads = HDT_EMIT_THREADS;
__global__ void __launch_bounds__(kEmitThreads) color_emit_kernel(const SegmentWeights* __restrict__ segWeights, const TileSegments* __restrict__ tileSegs,
                                                                   const TilePieces* __restrict__ tiles, const u64* __restrict__ stage, const ulonglong2* __restrict__ offsets,
                                                                   const ColorLeafDev oldLeaf, u32* __restrict__ weights, u64* __restrict__ blocks, u64* __restrict__ macroBlocks)
{
    constexpr u32 T = kEmitThreads;
    __shared__ SegmentWeights sw[T];
    const u32 t = threadIdx.x;
    const ulonglong2 tile = offsets[blockIdx.x];
    const TilePieces mine = tiles[blockIdx.x];
    const TileSegments segsOf = tileSegs[blockIdx.x];
    if (t == 0) {   // MacroBlockStruct, vwsc.h:595-599 / build() :677-681
        macroBlocks[2 * u64(blockIdx.x)] = tile.x;
        macroBlocks[2 * u64(blockIdx.x) + 1] = tile.y;
    }
    const u32 nSeg = segsOf.count;
    if (t < min(T, nSeg)) sw[t] = segWeights[segsOf.first + t];   // the first chunk's weight ranges, in flight beside the copy below
    {   // four entries in flight per thread
        const u64* from = stage + mine.stageBase;
        u64* to = blocks + tile.x;
        u32 k = t;
        for (; k + 3 * T < mine.blocks; k += 4 * T) {
            const u64 e0 = __ldg(from + k), e1 = __ldg(from + k + T), e2 = __ldg(from + k + 2 * T), e3 = __ldg(from + k + 3 * T);
            to[k] = e0; to[k + T] = e1; to[k + 2 * T] = e2; to[k + 3 * T] = e3;
        }
        for (; k < mine.blocks; k += T) to[k] = __ldg(from + k);
    }
    if (mine.bits == 0) return;
    const u32 skew = u32(tile.y & 31);                      // the macro block's first bit within its first word
    u32* out = weights + (tile.y >> 5);
    if (nSeg > T) {
        // several chunks: the words two chunks share are put together with atomicOr as well.  Zero the macro block's words,
        // except the first and the last (shared with the neighbours, zeroed by the scan)
        const u64 endBit = offsets[blockIdx.x + 1].y;
        for (u64 w = (tile.y >> 5) + 1 + t; w + 1 <= (endBit >> 5); w += T) weights[w] = 0;
        __syncthreads();
    }
    for (u32 segBase = 0; segBase < nSeg; segBase += T) {
        const u32 nChunk = min(T, nSeg - segBase);
        if (segBase && t < nChunk) sw[t] = segWeights[segsOf.first + segBase + t];
        __syncthreads();
        const u32 chunkBit0 = sw[0].bit0, chunkBit1 = sw[nChunk - 1].bit1;
        if (chunkBit1 > chunkBit0) {
            const u32 q0 = skew + chunkBit0, q1 = skew + chunkBit1;
            const u32 w0 = q0 >> 5, w1 = (q1 - 1) >> 5;
            // one COPY segment, everything about it in registers: a whole word that comes from one old macro block is two
            // loads and a funnel shift; the others (first, last, across the old macro boundary) go the general way
            const bool oneCopy = nChunk == 1 && !sw[0].fill;
            const u32 mid0 = sw[0].mid;
            const u64 delta0 = sw[0].delta[0], delta1 = sw[0].delta[1];
#pragma unroll 4
            for (u32 w = w0 + t; w <= w1; w += T) {
                if (oneCopy && (w << 5) >= q0 && ((w + 1) << 5) <= q1) {
                    const u32 a = (w << 5) - skew;
                    if (a + 32 <= mid0 || a >= mid0) {
                        const u64 p = (a >= mid0 ? delta1 : delta0) + a;
                        if ((p >> 5) + 1 < oldLeaf.nWeights) {
                            const u32 shift = u32(p) & 31;
                            const u32* src = oldLeaf.weights + (p >> 5);
                            const u32 hi = __byte_perm(__ldg(src), 0, 0x0123), lo = shift ? __byte_perm(__ldg(src + 1), 0, 0x0123) : 0u;
                            out[w] = __byte_perm(__funnelshift_l(lo, hi, shift), 0, 0x0123);
                            continue;
                        }
                    }
                }
                const u32 a0 = max(q0, w << 5) - skew, b0 = min(q1, (w + 1) << 5) - skew;     // bits of the macro block in this word
                u32 sgi = 0;                                 // first segment whose bits end after a0
                if (nChunk > 1) {
                    u32 e = nChunk - 1;
                    while (sgi < e) {
                        const u32 mid = (sgi + e) >> 1;
                        if (sw[mid].bit1 > a0) e = mid; else sgi = mid + 1;
                    }
                }
                u32 v = 0;
                for (; sgi < nChunk && sw[sgi].bit0 < b0; ++sgi) {
                    const SegmentWeights g = sw[sgi];
                    if (max(a0, g.bit0) >= min(b0, g.bit1)) continue;
                    if (!g.fill) {
#pragma unroll
                        for (u32 half = 0; half < 2; ++half) {
                            const u32 a = max(a0, half ? g.mid : g.bit0), b = min(b0, half ? g.bit1 : g.mid);
                            if (a < b) v |= read_stream_bits(oldLeaf, (half ? g.delta[1] : g.delta[0]) + a, b - a) >> ((a + skew) & 31);
                        }
                    } else {                                 // the stream of a repeated weight, from the phase this word starts in
                        const u32 a = max(a0, g.bit0), b = min(b0, g.bit1);
                        const u32 bpw = (g.fill >> 8) & 0xFF;
                        u64 pat = 0;
                        for (u32 k = 0; k < 64; k += bpw) pat |= (u64(g.fill >> 16) << (64 - bpw)) >> k;
                        v |= (u32((pat << ((a - g.bit0) % bpw)) >> 32) & (0xFFFFFFFFu << (32 - (b - a)))) >> ((a + skew) & 31);
                    }
                }
                v = __byte_perm(v, 0, 0x0123);               // ColorUtils::swap_byte_order, build() :671-674
                if (b0 - a0 == 32) out[w] = v;               // the word belongs to this chunk alone
                else if (v) atomicOr(out + w, v);            // shared with the neighbouring chunk or macro block
            }
        }
        __syncthreads();                                    // `sw` is rewritten by the next chunk
    }
}

}  // namespace hdt
