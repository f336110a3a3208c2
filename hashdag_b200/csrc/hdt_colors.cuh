// Device-side colour path: variable-weight-size colour decode and the per-pixel colour walk.
// Reference: variable_weight_size_colors.h:86-155,343-413; color_utils.h:9-148; tracer.cu:254-451;
// basic_dag.h:47-242; hash_dag_colors.h:24-50.
#pragma once
#include "hdt_device.cuh"

// The full walk issues the next node's pointer load beside the colour-tree load instead of behind the sibling loop (two waits
// per level instead of three): 0.164 ms against 0.169 ms per 1080p pass on B200 (profiles/r2_ab.md).
#ifndef HDT_COLORS_HOIST
#define HDT_COLORS_HOIST 1
#endif

namespace hdt {

constexpr u64 kColorsPerMacroBlock = 16 * 1024;  // vwsc.h:8
constexpr u32 kColorTreeLevels = 10;             // hash_dag_globals.h:7
constexpr u64 kUniqueOffset = ~u64(0);           // vwsc.h:161

struct ArrayDev { const void* data; u64 size; };

struct ColorLeafDev {   // the GPU half of CompressedColorLeaf (vwsc.h:159-165)
    u64 offset;
    const u32* weights; u64 nWeights;
    const u64* blocks; u64 nBlocks;
    const u64* macroBlocks; u64 nMacroWords;
    __device__ __forceinline__ bool is_shared() const { return offset != kUniqueOffset; }
    __device__ __forceinline__ bool is_valid() const { return blocks != nullptr; }
    __device__ __forceinline__ bool is_valid_index(u64 index) const
    {
        if (is_shared()) index += offset;
        return (2 * index / kColorsPerMacroBlock) < nMacroWords;
    }
};

// Same 104-byte layout as hdt_color_leaf / CompressedColorLeaf, for reading HashDAGColors::leaves.
struct ColorLeafPod { u64 offset; ArrayDev wG, bG, mG, wC, bC, mC; };
static_assert(sizeof(ColorLeafPod) == 104, "CompressedColorLeaf layout");

__device__ __forceinline__ ColorLeafDev leaf_from_pod(const ColorLeafPod& p)
{
    ColorLeafDev l;
    l.offset = p.offset;
    l.weights = static_cast<const u32*>(p.wG.data); l.nWeights = p.wG.size;
    l.blocks = static_cast<const u64*>(p.bG.data); l.nBlocks = p.bG.size;
    l.macroBlocks = static_cast<const u64*>(p.mG.data); l.nMacroWords = p.mG.size;
    return l;
}

struct ColorsDev {
    int kind;                 // HDT_COLORS_*
    u32 topLevels;
    const u64* enclosedLeaves;
    ColorLeafDev leaf;        // default leaf (Basic) / main shared leaf (Hash)
    const u32* uncompressed; u64 nUncompressed;
    const u32* nodes;         // HashDAGColors::nodes_GPU
    const ColorLeafPod* leaves;
    const u64* offsets;
};

struct CompressedColorDev { u32 colorBits; u32 weight; u32 bitsPerWeight; };

__device__ __forceinline__ float clampf(float f, float a, float b) { return f < a ? a : f > b ? b : f; }

__device__ __forceinline__ u32 float3_to_rgb888(float r, float g, float b)   // color_utils.h:28-37
{
    r = clampf(r, 0.f, 1.f); g = clampf(g, 0.f, 1.f); b = clampf(b, 0.f, 1.f);
    return u32(__float2uint_rz(__fmul_rn(r, 255.0f))) | (u32(__float2uint_rz(__fmul_rn(g, 255.0f))) << 8) |
           (u32(__float2uint_rz(__fmul_rn(b, 255.0f))) << 16) | 0xff000000u;
}
__device__ __forceinline__ float3 rgb888_to_float3(u32 c)
{
    return make_float3(__fdiv_rn(__uint2float_rn(c & 0xFF), 255.0f), __fdiv_rn(__uint2float_rn((c >> 8) & 0xFF), 255.0f),
                       __fdiv_rn(__uint2float_rn((c >> 16) & 0xFF), 255.0f));
}
__device__ __forceinline__ float3 rgb565_to_float3(u32 c)
{
    return make_float3(__fdiv_rn(__uint2float_rn(c & 0x1F), 31.0f), __fdiv_rn(__uint2float_rn((c >> 5) & 0x3F), 63.0f),
                       __fdiv_rn(__uint2float_rn((c >> 11) & 0x1F), 31.0f));
}
__device__ __forceinline__ float3 rgb101210_to_float3(u32 c)
{
    return make_float3(__fdiv_rn(__uint2float_rn(c & 0x3FF), 1023.0f), __fdiv_rn(__uint2float_rn((c >> 10) & 0xFFF), 4095.0f),
                       __fdiv_rn(__uint2float_rn((c >> 22) & 0x3FF), 1023.0f));
}
__device__ __forceinline__ float color_weight(const CompressedColorDev& c)
{
    return c.bitsPerWeight == 0 ? 0.f : __fdiv_rn(__uint2float_rn(c.weight), __uint2float_rn((1u << c.bitsPerWeight) - 1));
}
__device__ __forceinline__ float3 color_value(const CompressedColorDev& c)     // vwsc.h:144-154
{
    if (c.bitsPerWeight == 0) return rgb101210_to_float3(c.colorBits);
    const float3 a = rgb565_to_float3(c.colorBits & 0xFFFF), b = rgb565_to_float3(c.colorBits >> 16);
    const float f = color_weight(c), g = __fsub_rn(1.0f, f);
    // lerp = a*(1-f) + b*f, which the reference build contracts to fma(b, f, a*(1-f))
    return make_float3(__fmaf_rn(b.x, f, __fmul_rn(a.x, g)), __fmaf_rn(b.y, f, __fmul_rn(a.y, g)), __fmaf_rn(b.z, f, __fmul_rn(a.z, g)));
}

// vwsc.h:343-413 (binary_search_blocks + get_color_for_block + get_color), extract_bits from
// color_utils.h:137-146 (byte-swapped weight stream).
__device__ __forceinline__ CompressedColorDev leaf_get_color(const ColorLeafDev& l, u64 colorIndex)
{
    if (l.is_shared()) colorIndex += l.offset;
    const u32 local = u32(colorIndex % kColorsPerMacroBlock);
    const u32 macro = u32(colorIndex / kColorsPerMacroBlock);
    u32 lo = u32(__ldg(l.macroBlocks + 2 * macro));
    u32 hi = (2 * (u64(macro) + 1) < l.nMacroWords) ? u32(__ldg(l.macroBlocks + 2 * (macro + 1)) - 1) : u32(l.nBlocks - 1);
    u32 pos = (lo + hi) / 2;
    u64 block = __ldg(l.blocks + pos);
    while ((u32(block) & 0x3FFF) != local && lo <= hi) {
        if ((u32(block) & 0x3FFF) > local) hi = pos - 1; else lo = pos + 1;
        pos = (lo + hi) / 2;
        block = __ldg(l.blocks + pos);
    }
    const u32 hdr = u32(block);
    CompressedColorDev out;
    out.colorBits = u32(block >> 32);
    out.weight = 0;
    out.bitsPerWeight = ((hdr >> 16) == 0xFFFF) ? 0u : (((hdr >> 14) & 0x3) + 1);
    if (out.bitsPerWeight) {
        const u64 bitPtr = __ldg(l.macroBlocks + 2 * macro + 1) + (hdr >> 16) + u64(local - (hdr & 0x3FFF)) * out.bitsPerWeight;
        const u8* bytes = reinterpret_cast<const u8*>(l.weights) + (bitPtr >> 3);
        const u32 be16 = (u32(__ldg(bytes)) << 8) | u32(__ldg(bytes + 1));
        out.weight = (be16 >> (16 - out.bitsPerWeight - u32(bitPtr & 7))) & ((1u << out.bitsPerWeight) - 1);
    }
    return out;
}

// The same block as leaf_get_color finds, with a shorter chain of dependent loads: the reference's binary search
// (vwsc.h:354-369) ends on the LAST block of the macro block whose first colour is <= local (blocks are sorted by their
// first colour, and a macro block's first block starts at colour 0) after ~log2(n) dependent 8-byte loads -- about ten for
// the ~1000 blocks per macro block of a noisy leaf.  Here every round probes K-1 evenly spaced headers at once
// (independent loads, one latency) and keeps the K-th of the range that contains the answer: log_K(n) rounds.
#ifndef HDT_COLOR_SEARCH_K
#define HDT_COLOR_SEARCH_K 4   // A/B on B200 (profiles/r2_ab.md): 4 -> 0.075 ms, 8 -> 0.079, 16 -> 0.092 per 1080p colour pass
#endif
__device__ __forceinline__ CompressedColorDev leaf_get_color_wide(const ColorLeafDev& l, u64 colorIndex)
{
    constexpr u32 K = HDT_COLOR_SEARCH_K;
    if (l.is_shared()) colorIndex += l.offset;
    const u32 local = u32(colorIndex % kColorsPerMacroBlock);
    const u32 macro = u32(colorIndex / kColorsPerMacroBlock);
    const u64 bitBase = __ldg(l.macroBlocks + 2 * macro + 1);      // issued with the two range loads; used after the search
    u32 lo = u32(__ldg(l.macroBlocks + 2 * macro));
    u32 hi = (2 * (u64(macro) + 1) < l.nMacroWords) ? u32(__ldg(l.macroBlocks + 2 * (macro + 1)) - 1) : u32(l.nBlocks - 1);
    const u32* hdrs = reinterpret_cast<const u32*>(l.blocks);      // low word of block i = hdrs[2 i]: weight offset | bpw-1 | first colour
    // invariant: the answer lies in [lo, hi]
    while (hi > lo) {
        const u32 n = hi - lo + 1;                                  // >= 2
        const u32 step = (n + K - 1) / K;                           // >= 1; probes at lo + i*step, i = 1 .. K-1, while <= hi
        u32 first[K - 1];
#pragma unroll
        for (u32 i = 1; i < K; ++i) {
            const u32 pos = lo + i * step;
            first[i - 1] = pos <= hi ? (__ldg(hdrs + 2 * u64(pos)) & 0x3FFFu) : 0xFFFFFFFFu;
        }
        u32 j = 0;                                                  // probes whose block starts at or before `local` (a prefix of them)
#pragma unroll
        for (u32 i = 1; i < K; ++i) j += first[i - 1] <= local ? 1u : 0u;
        lo += j * step;
        hi = min(hi, lo + step - 1);
    }
    const u64 block = __ldg(l.blocks + lo);
    const u32 hdr = u32(block);
    CompressedColorDev out;
    out.colorBits = u32(block >> 32);
    out.weight = 0;
    out.bitsPerWeight = ((hdr >> 16) == 0xFFFF) ? 0u : (((hdr >> 14) & 0x3) + 1);
    if (out.bitsPerWeight) {
        const u64 bitPtr = bitBase + (hdr >> 16) + u64(local - (hdr & 0x3FFF)) * out.bitsPerWeight;
        const u8* bytes = reinterpret_cast<const u8*>(l.weights) + (bitPtr >> 3);
        const u32 be16 = (u32(__ldg(bytes)) << 8) | u32(__ldg(bytes + 1));
        out.weight = (be16 >> (16 - out.bitsPerWeight - u32(bitPtr & 7))) & ((1u << out.bitsPerWeight) - 1);
    }
    return out;
}

__device__ __forceinline__ u32 murmurhash32(u32 h)   // utils.h:68-76
{
    h ^= h >> 16; h *= 0x85ebca6b; h ^= h >> 13; h *= 0xc2b2ae35; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ float tool_strength(const hdt_tool_info& t, u32 x, u32 y, u32 z)   // tracer.h:51-76
{
    auto sphere = [&](const u32* p, float radius) {
        const float dx = __uint2float_rn(p[0]) - __uint2float_rn(x), dy = __uint2float_rn(p[1]) - __uint2float_rn(y), dz = __uint2float_rn(p[2]) - __uint2float_rn(z);
        // length(): the reference's compiled dot is fma(z,z, fma(x,x, y*y)) (its PTX, TOOL_OVERLAY build)
        return 1.f - __fdiv_rn(__fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)))), radius);
    };
    auto cube = [&](const u32* p, float radius) {
        const float dx = fabsf(__uint2float_rn(p[0]) - __uint2float_rn(x)), dy = fabsf(__uint2float_rn(p[1]) - __uint2float_rn(y)), dz = fabsf(__uint2float_rn(p[2]) - __uint2float_rn(z));
        const float myz = dy > dz ? dy : dz, m = dx > myz ? dx : myz;
        return 1.f - __fdiv_rn(m, radius);
    };
    switch (t.tool) {
    case 0: case 1: case 2: return sphere(t.position, t.radius);
    case 3: return cube(t.position, t.radius);
    default: return fmaxf(fmaxf(sphere(t.copy_source, 3.f), sphere(t.copy_dest, 3.f)), cube(t.position, t.radius));
    }
}

struct ColorsParams {
    int debugColors; u32 debugIndexLevel; int overlay; hdt_tool_info tool;
};

// Where a pixel's colour goes: the tool overlay of tracer.cu:276-286 and the "invalid" checkerboard of :288-292.
struct ColorSink {
    const ColorsParams& prm;
    u32 px, py, pz;
    float strength;
    __device__ __forceinline__ ColorSink(const ColorsParams& p, u32 x, u32 y, u32 z) : prm(p), px(x), py(y), pz(z), strength(p.overlay ? tool_strength(p.tool, x, y, z) : 0.f) {}
    __device__ __forceinline__ u32 set(u32 color) const
    {
        if (prm.overlay && strength > 0.f) {
            const float3 c = rgb888_to_float3(color);
            const float f = clampf(100.f * strength, 0.f, .5f), g = 1.f - f;
            color = float3_to_rgb888(c.x * g + 1.f * f, c.y * g + 0.f * f, c.z * g + 0.f * f);
        }
        return color;
    }
    __device__ __forceinline__ u32 invalid() const
    {
        const u32 b = (px ^ py ^ pz) & 1;
        return set(float3_to_rgb888(1.f, __uint2float_rn(b), 1.f - __uint2float_rn(b)));
    }
};

__device__ __forceinline__ u32 sky_color() { return float3_to_rgb888(__fdiv_rn(187.f, 255.f), __fdiv_rn(242.f, 255.f), __fdiv_rn(250.f, 255.f)); }

// The end of tracer.cu:254-451 for compressed colours: colour `nofLeaves` of `leaf`, through the debug views.
template <bool WIDE>
__device__ __forceinline__ u32 compressed_color(const ColorLeafDev& leaf, u64 nofLeaves, int dbg, const ColorSink& sink)
{
    const CompressedColorDev cc = WIDE ? leaf_get_color_wide(leaf, nofLeaves) : leaf_get_color(leaf, nofLeaves);
    u32 color;
    if (dbg == HDT_DEBUG_COLOR_BITS) color = 0;   // debug hash is compiled out in the BENCHMARK configuration (vwsc.h:96-104)
    else if (dbg == HDT_DEBUG_MIN_COLOR) { const float3 c = rgb565_to_float3(cc.colorBits & 0xFFFF); color = float3_to_rgb888(c.x, c.y, c.z); }
    else if (dbg == HDT_DEBUG_MAX_COLOR) { const float3 c = rgb565_to_float3(cc.colorBits >> 16); color = float3_to_rgb888(c.x, c.y, c.z); }
    else if (dbg == HDT_DEBUG_WEIGHT) { const float w = color_weight(cc); color = float3_to_rgb888(w, w, w); }
    else { const float3 c = color_value(cc); color = float3_to_rgb888(c.x, c.y, c.z); }
    return sink.set(color);
}

// One pixel of tracer.cu:254-451.
template <class DAG>
__device__ u32 color_pixel(const DAG& dag, const ColorsDev& col, const u32 levels, const ColorsParams& prm, u32 px, u32 py, u32 pz)
{
    if ((px | py | pz) == 0) return sky_color();
    const ColorSink sink(prm, px, py, pz);
    auto set = [&](u32 color) { return sink.set(color); };
    auto invalid = [&]() { return sink.invalid(); };
    const u32 leafLevel = levels - 2;
    const bool hashColors = col.kind == HDT_COLORS_HASH;
    const u32 colorTreeLevels = hashColors ? kColorTreeLevels : 0;
    const int dbg = prm.debugColors;

    u64 nofLeaves = 0;
    u32 debugIndex = 0, colorNodeIndex = 0;
    ColorLeafDev leaf = col.leaf;
    if (hashColors) { leaf.blocks = nullptr; leaf.offset = 0; }   // get_default_leaf() == {} (hash_dag_colors.h:51-54)

    u32 level = 0;
    u32 handle = dag.root();
    u32 rawIndex = 0;   // the index the reference would hold (virtual pointer for HashDAG); only the Index/Position debug views show it
    if (dbg == HDT_DEBUG_INDEX || dbg == HDT_DEBUG_POSITION) rawIndex = dag.raw_root();
    while (level < leafLevel) {
        level++;
        const u32 childMask = dag.header(handle) & 0xFF;
        const u32 sh = levels - level;
        const u32 child = (((px >> sh) & 1) << 2) | (((py >> sh) & 1) << 1) | ((pz >> sh) & 1);
        if (!(childMask & (1u << child))) return set(0xFF00FF);
        const u32 childOff = __popc(childMask & ((1u << child) - 1u)) + 1;
#if HDT_COLORS_HOIST
        // the next node's pointer load is issued here, next to the colour-tree load below, instead of after the colour-tree
        // checks and the sibling loop
        const u32 nextHandle = dag.child(handle, childOff);
#endif

        if (level - 1 < colorTreeLevels) {
            colorNodeIndex = __ldg(col.nodes + colorNodeIndex + child);
            if (level == colorTreeLevels) {
                if (colorNodeIndex & 0x80000000u) leaf = leaf_from_pod(col.leaves[colorNodeIndex & 0x7FFFFFFFu]);
                else { leaf = col.leaf; leaf.offset = __ldg(col.offsets + colorNodeIndex); }
            } else if (!colorNodeIndex) return invalid();
        }

        if (dbg == HDT_DEBUG_INDEX || dbg == HDT_DEBUG_POSITION || dbg == HDT_DEBUG_COLOR_TREE) {
            if (dbg == HDT_DEBUG_INDEX && prm.debugIndexLevel == level - 1) debugIndex = rawIndex;
            if (level == leafLevel) {
                if (prm.debugIndexLevel == leafLevel) debugIndex = dag.raw_child(handle, childOff);
                if (dbg == HDT_DEBUG_INDEX) return set(murmurhash32(debugIndex));
                if (dbg == HDT_DEBUG_POSITION) {
                    float c = __fdiv_rn(__uint2float_rn((px ^ py ^ pz) & 0x7FF), 2047.f);
                    c = __double2float_rn(__ddiv_rn(__dadd_rn(double(c), 0.5), 2.0));
                    return set((rawIndex & 0x80000000u) ? float3_to_rgb888(c, 0.f, 0.f) : float3_to_rgb888(c, c, c));
                }
                const u32 off = levels - colorTreeLevels;
                const float c = __uint2float_rn(((px >> off) ^ (py >> off) ^ (pz >> off)) & 1);
                return set(float3_to_rgb888(c, c, c));
            }
            rawIndex = dag.raw_child(handle, childOff);
            handle = dag.to_handle(rawIndex);
            continue;
        }

        if (level == leafLevel) {
            // preceding leaves: issue all pointer loads, then all leaf loads (independent chains)
            for (u32 c = 0, k = 1; c < child; ++c)
                if (childMask & (1u << c)) {
                    const uint2 l = dag.leaf(dag.child(handle, k++));
                    nofLeaves += __popc(l.x) + __popc(l.y);
                }
#if HDT_COLORS_HOIST
            const uint2 l = dag.leaf(nextHandle);
#else
            const uint2 l = dag.leaf(dag.child(handle, childOff));
#endif
            const u32 bit = ((px & 1) ? 4 : 0) | ((py & 1) ? 2 : 0) | ((pz & 1) ? 1 : 0) | ((px & 2) ? 32 : 0) | ((py & 2) ? 16 : 0) | ((pz & 2) ? 8 : 0);
            const u64 l64 = (u64(l.y) << 32) | l.x;
            nofLeaves += __popcll(l64 & ((u64(1) << bit) - 1));
            break;
        }
        if (level > colorTreeLevels) {
            for (u32 c = 0, k = 1; c < child; ++c)
                if (childMask & (1u << c)) {
                    const u32 node = dag.header(dag.child(handle, k++));
                    const u32 upper = node >> 8;
                    // get_leaves_count: basic_dag.h:62-76 / hash_dag_colors.h:28-32
                    nofLeaves += (!hashColors && level < col.topLevels) ? __ldg(col.enclosedLeaves + upper) : u64(upper);
                }
        }
#if HDT_COLORS_HOIST
        handle = nextHandle;
#else
        handle = dag.child(handle, childOff);
#endif
    }

    if (col.kind == HDT_COLORS_UNCOMPRESSED) {
        if (!col.uncompressed || nofLeaves >= col.nUncompressed) return invalid();
        if (dbg == HDT_DEBUG_COLOR_BITS) return set(0);
        if (dbg == HDT_DEBUG_MIN_COLOR || dbg == HDT_DEBUG_MAX_COLOR || dbg == HDT_DEBUG_WEIGHT) return set(float3_to_rgb888(0.f, 0.f, 0.f));
        const float3 c = rgb888_to_float3(__ldg(col.uncompressed + u32(nofLeaves)));
        return set(float3_to_rgb888(c.x, c.y, c.z));
    }
    if (!leaf.is_valid() || !leaf.is_valid_index(nofLeaves)) return invalid();
    if (col.kind == HDT_COLORS_ERRORS) {
        if (!col.uncompressed || u32(nofLeaves) >= col.nUncompressed) return invalid();
        const float3 a = color_value(leaf_get_color(leaf, u32(nofLeaves))), b = rgb888_to_float3(__ldg(col.uncompressed + u32(nofLeaves)));
        const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
        const float err = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
        const float v = (double(err) > 0.04) ? 1.f : 0.f;
        if (dbg == HDT_DEBUG_COLOR_BITS) return set(0);
        if (dbg == HDT_DEBUG_MIN_COLOR || dbg == HDT_DEBUG_MAX_COLOR || dbg == HDT_DEBUG_WEIGHT) return set(float3_to_rgb888(0.f, 0.f, 0.f));
        const float3 c = rgb888_to_float3(float3_to_rgb888(v, v, v));
        return set(float3_to_rgb888(c.x, c.y, c.z));
    }
    return compressed_color<false>(leaf, nofLeaves, dbg, sink);
}

// The same pixel for a HashDAG with HashDAGColors when trace_paths left an AncestorRecord (hdt_device.cuh) and the
// resolved pool has a prefix pool (hdt_resolve.cuh).  What is left of the reference's walk:
//   * the ten levels of the colour tree (hash_dag_colors.h:33-50), a chain of dependent loads that needs only the path;
//   * below the tree the voxel's index inside its colour leaf = sum over the ancestors of depth 10 .. levels-3 of the
//     voxels under their earlier children -- ONE prefix-pool load per level at the recorded pointer word, all independent of
//     each other and of the colour-tree chain -- plus the set bits below the voxel's own bit in its 64-bit leaf (recorded);
//   * the colour decode, with the wide block search.
// No DAG node is visited: the existence checks of tracer.cu:316-320 cannot fail for a path trace_paths has just walked
// through this DAG (the launch code only takes this route for the DAG the paths frame was traced in).
// Debug views that show node indices or positions take color_pixel.
__device__ __forceinline__ u32 color_pixel_recorded(const u32* __restrict__ prefix, const ColorsDev& col, const u32 levels, const ColorsParams& prm,
                                                    const u32 px, const u32 py, const u32 pz, const uint4 a, const uint4 b)
{
    const ColorSink sink(prm, px, py, pz);
    // below the colour tree: independent loads, issued first
    const u32 w[kMaxAncestorWords] = { a.x, a.y, a.z, a.w, b.x, b.y };
    u32 below[kMaxAncestorWords];
#pragma unroll
    for (u32 i = 0; i < kMaxAncestorWords; ++i) below[i] = (kColorTreeDepth + i + 3 <= levels) ? (__ldg(prefix + w[i]) & 0xFFFFFFu) : 0u;   // top byte: leaf mask
    // the colour tree
    u32 colorNodeIndex = 0;
#pragma unroll 1
    for (u32 level = 1; level <= kColorTreeLevels; ++level) {
        const u32 sh = levels - level;
        const u32 child = (((px >> sh) & 1) << 2) | (((py >> sh) & 1) << 1) | ((pz >> sh) & 1);
        colorNodeIndex = __ldg(col.nodes + colorNodeIndex + child);
        if (level < kColorTreeLevels && !colorNodeIndex) return sink.invalid();
    }
    ColorLeafDev leaf;
    if (colorNodeIndex & 0x80000000u) leaf = leaf_from_pod(col.leaves[colorNodeIndex & 0x7FFFFFFFu]);
    else { leaf = col.leaf; leaf.offset = __ldg(col.offsets + colorNodeIndex); }
    u64 nofLeaves = 0;
#pragma unroll
    for (u32 i = 0; i < kMaxAncestorWords; ++i) nofLeaves += below[i];
    const u32 bit = ((px & 1) ? 4 : 0) | ((py & 1) ? 2 : 0) | ((pz & 1) ? 1 : 0) | ((px & 2) ? 32 : 0) | ((py & 2) ? 16 : 0) | ((pz & 2) ? 8 : 0);
    const u64 l64 = (u64(b.w) << 32) | b.z;
    nofLeaves += __popcll(l64 & ((u64(1) << bit) - 1));
    if (!leaf.is_valid() || !leaf.is_valid_index(nofLeaves)) return sink.invalid();
    return compressed_color<true>(leaf, nofLeaves, prm.debugColors, sink);
}

}  // namespace hdt
