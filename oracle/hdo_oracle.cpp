// CPU ORACLE -- TEST INFRASTRUCTURE, NOT THE PRODUCT (see hdo_oracle.h for scope and pinning).
//
// Build: g++ -O2 -std=c++17 -march=x86-64-v3 -ffp-contract=off  (no -ffast-math).  Every fused
// multiply-add the reference's compiled kernels contain is written as an explicit fma()/fmaf()
// below; everything else is a single correctly-rounded IEEE operation, so this file, the
// reference's sm_100a SASS and the product kernels compute the same bits.  The contraction sites
// were read off the reference's own PTX/SASS (nvcc 12.9, default -fmad=true) and are listed in
// DESIGN.md §4.
#include "hdo_oracle.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

struct Counters { u64 word = 0, leaf = 0, page = 0, hit = 0, steps = 0, probe = 0; };

// ---------------------------------------------------------------------------------------------
// DAG accessors: base_dag.h:74-80 concept over BasicDAG (basic_dag.h:20-35) and HashDAG
// (hash_dag.h:228-247 + hash_table.h:156-173).
// ---------------------------------------------------------------------------------------------
struct Dag {
    const hdo_dag& d;
    Counters& c;
    const bool hash;
    Dag(const hdo_dag& dd, Counters& cc) : d(dd), c(cc), hash(dd.kind == HDO_DAG_HASH) {}

    u32 levels() const { return d.levels; }
    u32 leaf_level() const { return d.levels - 2; }
    u32 first() const { return hash ? d.first_node_index : 0; }
    inline const u32* ptr(u32 index) const
    {
        if (!hash) return d.data + index;
        ++c.page;
        const u32 page = index / 512, off = index % 512;
        return d.data + u64(d.page_table[page]) * 512 + off;
    }
    u32 get_node(u32 index) const { ++c.word; return *ptr(index); }
    u32 get_child_index(u32 index, u8 childMask, u8 child) const
    {
        ++c.word;
        return *ptr(index + u32(__builtin_popcount(childMask & ((1u << child) - 1u))) + 1);
    }
    u64 get_leaf(u32 index) const
    {
        ++c.leaf;
        const u32* p = ptr(index);
        return u64(p[0]) | (u64(p[1]) << 32);
    }
};

// base_dag.h:16-58
inline u8 first_child_mask(u64 leaf)
{
    u8 m = 0;
    for (int i = 0; i < 8; ++i)
        if ((leaf >> (8 * i)) & 0xFF) m |= u8(1u << i);
    return m;
}
inline u8 second_child_mask(u64 leaf, u8 first) { return u8(leaf >> (8 * first)); }

struct F3 { float x, y, z; };
struct D3 { double x, y, z; };

// tracer.cu:19-136.  Node centre/radius are exact in float (coordinates < 2^24, radius a power of
// two), so radius*|invDir| is exact too and pmin/pmax come out the same fused or not.
template <bool isRoot>
inline u8 intersection_mask(u32 level, u32 levels, u32 px, u32 py, u32 pz, F3 o, F3 d, F3 inv)
{
    const u32 shift = levels - level;
    const float radius = float(1u << (shift - 1));
    const float cx = radius + float(px << shift), cy = radius + float(py << shift), cz = radius + float(pz << shift);
    const float rx = cx - o.x, ry = cy - o.y, rz = cz - o.z;
    const float tx = rx * inv.x, ty = ry * inv.y, tz = rz * inv.z;
    const float sx = radius * std::fabs(inv.x), sy = radius * std::fabs(inv.y), sz = radius * std::fabs(inv.z);
    const float ax = tx - sx, ay = ty - sy, az = tz - sz;
    // max(float3) is the ternary constexpr_max (cuda_math.h:42), the outer max is fmaxf (tracer.cu:46)
    const float ayz = (ay > az) ? ay : az;
    const float a3 = (ax > ayz) ? ax : ayz;
    const float tmin = fmaxf(a3, 0.0f);
    const float bx = tx + sx, by = ty + sy, bz = tz + sz;
    const float byz = (by < bz) ? by : bz;
    const float tmax = (bx < byz) ? bx : byz;
    if (isRoot && (tmin >= tmax)) return 0;

    u8 mask = 0;
    {
        const float h = 0.5f * (tmin + tmax);
        const float qx = h * d.x, qy = h * d.y, qz = h * d.z;
        const u8 first = u8(((qx >= rx) ? 4 : 0) + ((qy >= ry) ? 2 : 0) + ((qz >= rz) ? 1 : 0));
        mask |= u8(1u << first);
    }
    const float eps = 1e-4f;
    if (tmin <= tx && tx <= tmax) {
        const float qy = tx * d.y, qz = tx * d.z;
        u8 A = 0, B = 0;
        if (qy >= ry - eps) A |= 0xCC;
        if (qy <= ry + eps) A |= 0x33;
        if (qz >= rz - eps) B |= 0xAA;
        if (qz <= rz + eps) B |= 0x55;
        mask |= A & B;
    }
    if (tmin <= ty && ty <= tmax) {
        const float qx = ty * d.x, qz = ty * d.z;
        u8 C = 0, D = 0;
        if (qx >= rx - eps) C |= 0xF0;
        if (qx <= rx + eps) C |= 0x0F;
        if (qz >= rz - eps) D |= 0xAA;
        if (qz <= rz + eps) D |= 0x55;
        mask |= C & D;
    }
    if (tmin <= tz && tz <= tmax) {
        const float qx = tz * d.x, qy = tz * d.y;
        u8 E = 0, F = 0;
        if (qx >= rx - eps) E |= 0xF0;
        if (qx <= rx + eps) E |= 0x0F;
        if (qy >= ry - eps) F |= 0xCC;
        if (qy <= ry + eps) F |= 0x33;
        mask |= E & F;
    }
    return mask;
}

// tracer.cu:7-17
inline u8 next_child_ordered(u8 order, u8 mask)
{
    for (u8 child = 0; child < 8; ++child) {
        const u8 c = child ^ order;
        if (mask & (1u << c)) return c;
    }
    return 0;
}

struct StackEntry { u32 index; u8 childMask; u8 visitMask; };

// Shared DFS of tracer.cu:166-249 (ordered == true) and tracer.cu:458-542 (ordered == false:
// highest set bit first, any hit).  Returns true on a hit; path holds the voxel.
template <bool ordered>
inline bool traverse(const Dag& dag, F3 o, F3 d, F3 inv, u8 order, u32& outx, u32& outy, u32& outz)
{
    const u32 levels = dag.levels(), leafLevel = dag.leaf_level();
    u32 level = 0, px = 0, py = 0, pz = 0;
    StackEntry stack[32];
    StackEntry cache;
    u64 cachedLeaf = 0;

    cache.index = dag.first();
    cache.childMask = u8(dag.get_node(cache.index) & 0xFF);
    cache.visitMask = cache.childMask & intersection_mask<true>(0, levels, px, py, pz, o, d, inv);

    for (;;) {
        u32 newLevel = level;
        while (newLevel > 0 && !cache.visitMask) {
            newLevel--;
            cache = stack[newLevel];
        }
        if (newLevel == 0 && !cache.visitMask) { outx = outy = outz = 0; return false; }
        px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
        level = newLevel;

        const u8 nextChild = ordered ? next_child_ordered(order, cache.visitMask) : u8(31 - __builtin_clz(u32(cache.visitMask)));
        cache.visitMask &= u8(~(1u << nextChild));

        px = (px << 1) | ((nextChild & 4u) >> 2); py = (py << 1) | ((nextChild & 2u) >> 1); pz = (pz << 1) | (nextChild & 1u);
        stack[level] = cache;
        level++;
        ++dag.c.steps;

        if (level == levels) { outx = px; outy = py; outz = pz; return true; }

        if (level < leafLevel) {
            cache.index = dag.get_child_index(cache.index, cache.childMask, nextChild);
            cache.childMask = u8(dag.get_node(cache.index) & 0xFF);
        } else if (level == leafLevel) {
            const u32 addr = dag.get_child_index(cache.index, cache.childMask, nextChild);
            cachedLeaf = dag.get_leaf(addr);
            cache.childMask = first_child_mask(cachedLeaf);
        } else {
            cache.childMask = second_child_mask(cachedLeaf, nextChild);
        }
        cache.visitMask = cache.childMask & intersection_mask<false>(level, levels, px, py, pz, o, d, inv);
    }
}

// Primary-ray direction: tracer.cu:158 / :622.  Contraction as in the reference PTX:
// fma(px,ddx,rayMin) -> fma(py,ddy,.) -> -cam; dot = fma(z,z,fma(x,x,y*y)); * (1/sqrt).
inline D3 primary_direction(const double cam[3], const double rmin[3], const double ddx[3], const double ddy[3], u32 px, u32 py)
{
    const double fx = double(px), fy = double(py);
    const double vx = std::fma(fy, ddy[0], std::fma(fx, ddx[0], rmin[0])) - cam[0];
    const double vy = std::fma(fy, ddy[1], std::fma(fx, ddx[1], rmin[1])) - cam[1];
    const double vz = std::fma(fy, ddy[2], std::fma(fx, ddx[2], rmin[2])) - cam[2];
    const double len = std::sqrt(std::fma(vz, vz, std::fma(vx, vx, vy * vy)));
    const double r = 1.0 / len;
    return { vx * r, vy * r, vz * r };
}

template <class Fn>
void parallel_rows(u32 height, u32 nThreads, Fn fn)
{
    if (!nThreads) nThreads = std::max(1u, std::thread::hardware_concurrency());
    nThreads = std::min(nThreads, std::max(1u, height));
    std::vector<Counters> cs(nThreads);
    std::atomic<u32> next{ 0 };
    auto worker = [&](u32 t) {
        for (;;) {
            const u32 row = next.fetch_add(1);
            if (row >= height) break;
            fn(row, cs[t]);
        }
    };
    if (nThreads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (u32 t = 0; t < nThreads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    Counters tot;
    for (auto& c : cs) { tot.word += c.word; tot.leaf += c.leaf; tot.page += c.page; tot.hit += c.hit; tot.steps += c.steps; tot.probe += c.probe; }
    cs[0] = tot;
    fn(~0u, cs[0]);  // hand the totals back
}

inline void add_stats(hdo_stats* s, const Counters& c)
{
    if (!s) return;
    s->n_word += c.word; s->n_leaf += c.leaf; s->n_page += c.page; s->n_hit += c.hit; s->n_steps += c.steps; s->n_color_probe += c.probe;
}

// ---------------------------------------------------------------------------------------------
// Colours
// ---------------------------------------------------------------------------------------------
inline float clampf(float f, float a, float b) { return f < a ? a : f > b ? b : f; }
inline u32 float3_to_rgb888(F3 c)  // color_utils.h:28-37
{
    const float r = clampf(c.x, 0.f, 1.f), g = clampf(c.y, 0.f, 1.f), b = clampf(c.z, 0.f, 1.f);
    return (u32(r * 255.0f) << 0) | (u32(g * 255.0f) << 8) | (u32(b * 255.0f) << 16) | 0xff000000u;
}
inline F3 rgb888_to_float3(u32 rgb)
{
    return { float((rgb >> 0) & 0xFF) / 255.0f, float((rgb >> 8) & 0xFF) / 255.0f, float((rgb >> 16) & 0xFF) / 255.0f };
}
inline F3 rgb565_to_float3(u16 rgb)
{
    return { float((rgb >> 0) & 0x1F) / 31.0f, float((rgb >> 5) & 0x3F) / 63.0f, float((rgb >> 11) & 0x1F) / 31.0f };
}
inline F3 rgb101210_to_float3(u32 rgb)
{
    return { float((rgb >> 0) & 0x3FF) / 1023.0f, float((rgb >> 10) & 0xFFF) / 4095.0f, float((rgb >> 22) & 0x3FF) / 1023.0f };
}
inline u32 murmurhash32(u32 h)
{
    h ^= h >> 16; h *= 0x85ebca6b; h ^= h >> 13; h *= 0xc2b2ae35; h ^= h >> 16;
    return h;
}

struct CompressedColor {  // vwsc.h:86-155
    u32 colorBits = 0; u8 weight = 0; u8 bitsPerWeight = 0;
    F3 min_color() const { return rgb565_to_float3(u16(colorBits & 0xFFFF)); }
    F3 max_color() const { return rgb565_to_float3(u16((colorBits >> 16) & 0xFFFF)); }
    float get_weight() const { return bitsPerWeight == 0 ? 0.f : float(weight) / float((1 << bitsPerWeight) - 1); }
    F3 get_color() const
    {
        if (bitsPerWeight == 0) return rgb101210_to_float3(colorBits);
        // lerp(a,b,f) = a*(1-f) + b*f, contracted by nvcc to fma(b, f, a*(1-f))
        const F3 a = min_color(), b = max_color();
        const float f = get_weight(), g = 1 - f;
        return { fmaf(b.x, f, a.x * g), fmaf(b.y, f, a.y * g), fmaf(b.z, f, a.z * g) };
    }
};

struct ColorLeafView {   // CompressedColorLeaf (vwsc.h:157-413), with the shared-leaf offset
    const hdo_color_leaf* l = nullptr;
    u64 offset = ~u64(0);  // uniqueOffset
    bool is_shared() const { return offset != ~u64(0); }
    bool is_valid() const { return l && l->blocks != nullptr; }
    bool is_valid_index(u64 index) const
    {
        if (is_shared()) index += offset;
        return (2 * index / (16 * 1024)) < l->n_macro_words;
    }
    // color_utils.h:137-146 (byte-swapped weights)
    static u8 extract_bits(u32 bits, const u32* array, u64 bitPtr)
    {
        if (bits == 0) return 0;
        u16 dst;
        std::memcpy(&dst, reinterpret_cast<const u8*>(array) + bitPtr / 8, sizeof(u16));
        dst = u16((dst << 8) | (dst >> 8));
        const u64 ptrBit = bitPtr % 8;
        return u8((dst >> (16 - bits - ptrBit)) & ((1u << bits) - 1));
    }
    CompressedColor get_color(u64 colorIndex, Counters& c) const
    {
        if (is_shared()) colorIndex += offset;
        const u16 local = u16(colorIndex % (16 * 1024));
        const u32 macro = u32(colorIndex / (16 * 1024));
        // binary_search_blocks, vwsc.h:343-371
        u32 lo = u32(l->macro_blocks[2 * macro]);
        u32 hi = (2 * (u64(macro) + 1) < l->n_macro_words) ? u32(l->macro_blocks[2 * (macro + 1)] - 1) : u32(l->n_blocks - 1);
        u32 pos = (lo + hi) / 2;
        u32 hdr = u32(l->blocks[pos]); ++c.probe;
        while (u16(hdr & 0x3FFF) != local && lo <= hi) {
            if (u16(hdr & 0x3FFF) > local) hi = pos - 1; else lo = pos + 1;
            pos = (lo + hi) / 2;
            hdr = u32(l->blocks[pos]); ++c.probe;
        }
        // get_color_for_block, vwsc.h:373-403
        const u64 block = l->blocks[pos]; ++c.probe;
        const u32 bh = u32(block);
        CompressedColor out;
        out.colorBits = u32(block >> 32);
        out.bitsPerWeight = (u16(bh >> 16) == 0xFFFF) ? u8(0) : u8(((bh >> 14) & 0x3) + 1);
        if (out.bitsPerWeight) {
            const u64 wi = l->macro_blocks[2 * macro + 1] + (bh >> 16) + u32(local - u16(bh & 0x3FFF)) * out.bitsPerWeight;
            out.weight = extract_bits(out.bitsPerWeight, l->weights, wi);
        }
        return out;
    }
};

inline u64 leaves_count(const hdo_colors& col, u32 level, u32 node)
{
    if (col.kind == HDO_COLORS_HASH) return node >> 8;          // hash_dag_colors.h:28-32
    const u32 upper = node >> 8;                                  // basic_dag.h:62-76
    return level < col.top_levels ? col.enclosed_leaves[upper] : upper;
}

inline float tool_strength(const hdo_tool_info& t, u32 x, u32 y, u32 z)  // tracer.h:51-76
{
    auto sphere = [&](const u32* p, float radius) {
        const float dx = float(p[0]) - float(x), dy = float(p[1]) - float(y), dz = float(p[2]) - float(z);
        // length(): dot contracted by nvcc as fma(z,z, fma(x,x, y*y)) (read off the reference's PTX, TOOL_OVERLAY build)
        return 1 - std::sqrt(fmaf(dz, dz, fmaf(dx, dx, dy * dy))) / radius;
    };
    auto cube = [&](const u32* p, float radius) {
        const float dx = std::fabs(float(p[0]) - float(x)), dy = std::fabs(float(p[1]) - float(y)), dz = std::fabs(float(p[2]) - float(z));
        const float m = dx > (dy > dz ? dy : dz) ? dx : (dy > dz ? dy : dz);
        return 1 - m / radius;
    };
    switch (t.tool) {
    case 0: case 1: case 2: return sphere(t.position, t.radius);
    case 3: return cube(t.position, t.radius);
    default: return fmaxf(fmaxf(sphere(t.copy_source, 3), sphere(t.copy_dest, 3)), cube(t.position, t.radius));
    }
}

// One pixel of tracer.cu:254-451.
u32 color_pixel(const Dag& dag, const hdo_colors& col, u32 px, u32 py, u32 pz, int debugColors, u32 debugIndexLevel,
                const hdo_tool_info* tool, bool overlay)
{
    if (px == 0 && py == 0 && pz == 0) return float3_to_rgb888({ 187 / 255.f, 242 / 255.f, 250 / 255.f });
    const float strength = (overlay && tool) ? tool_strength(*tool, px, py, pz) : 0.f;
    auto set = [&](u32 color) {
        if (overlay && strength > 0) {
            const F3 c = rgb888_to_float3(color);
            const float f = clampf(100 * strength, 0.f, .5f), g = 1 - f;
            color = float3_to_rgb888({ c.x * g + 1.f * f, c.y * g + 0.f * f, c.z * g + 0.f * f });
        }
        return color;
    };
    auto invalid = [&]() {
        const u32 b = (px ^ py ^ pz) & 0x1;
        return set(float3_to_rgb888({ 1.f, float(b), 1.f - float(b) }));
    };
    const u32 levels = dag.levels(), leafLevel = dag.leaf_level();
    const bool hashColors = col.kind == HDO_COLORS_HASH;
    const u32 colorTreeLevels = hashColors ? 10 : 0;

    u64 nof_leaves = 0;
    u32 debugColorsIndex = 0, colorNodeIndex = 0;
    ColorLeafView leaf;
    if (!hashColors) { leaf.l = &col.leaf; leaf.offset = 0; /* shared with offset 0 == the default-constructed leaf (vwsc.h:160) */ }

    u32 level = 0, nodeIndex = dag.first();
    while (level < leafLevel) {
        level++;
        const u32 node = dag.get_node(nodeIndex);
        const u8 childMask = u8(node & 0xFF);
        const u32 sh = levels - level;
        const u8 child = u8((((px >> sh) & 1) ? 4 : 0) | (((py >> sh) & 1) ? 2 : 0) | (((pz >> sh) & 1) ? 1 : 0));
        if (!(childMask & (1 << child))) return set(0xFF00FF);

        if (level - 1 < colorTreeLevels) {
            colorNodeIndex = col.color_nodes[colorNodeIndex + child];
            if (level == colorTreeLevels) {
                // HashDAGColors::get_leaf, hash_dag_colors.h:37-50
                if (colorNodeIndex & 0x80000000u) { leaf.l = &col.unique_leaves[colorNodeIndex & 0x7FFFFFFFu]; leaf.offset = ~u64(0); }
                else { leaf.l = &col.leaf; leaf.offset = col.color_offsets[colorNodeIndex]; }
            } else if (!colorNodeIndex) return invalid();
        }

        if (debugColors == 1 || debugColors == 2 || debugColors == 3) {
            if (debugColors == 1 && debugIndexLevel == level - 1) debugColorsIndex = nodeIndex;
            if (level == leafLevel) {
                if (debugIndexLevel == leafLevel) debugColorsIndex = dag.get_child_index(nodeIndex, childMask, child);
                if (debugColors == 1) return set(murmurhash32(debugColorsIndex));
                if (debugColors == 2) {
                    float color = float((px ^ py ^ pz) & 0x7FF) / float(0x7FF);
                    color = float((double(color) + 0.5) / 2);
                    return set(float3_to_rgb888((nodeIndex & 0x80000000u) ? F3{ color, 0, 0 } : F3{ color, color, color }));
                }
                const u32 offset = levels - colorTreeLevels;
                const float color = float(((px >> offset) ^ (py >> offset) ^ (pz >> offset)) & 0x1);
                return set(float3_to_rgb888({ color, color, color }));
            }
            nodeIndex = dag.get_child_index(nodeIndex, childMask, child);
            continue;
        }

        if (level == leafLevel) {
            for (u8 c = 0; c < child; ++c)
                if (childMask & (1u << c)) nof_leaves += u64(__builtin_popcountll(dag.get_leaf(dag.get_child_index(nodeIndex, childMask, c))));
            const u64 l64 = dag.get_leaf(dag.get_child_index(nodeIndex, childMask, child));
            const u8 bit = u8(((px & 1) ? 4 : 0) | ((py & 1) ? 2 : 0) | ((pz & 1) ? 1 : 0) | ((px & 2) ? 32 : 0) | ((py & 2) ? 16 : 0) | ((pz & 2) ? 8 : 0));
            nof_leaves += u64(__builtin_popcountll(l64 & ((u64(1) << bit) - 1)));
            break;
        }
        if (level > colorTreeLevels)
            for (u8 c = 0; c < child; ++c)
                if (childMask & (1u << c)) nof_leaves += leaves_count(col, level, dag.get_node(dag.get_child_index(nodeIndex, childMask, c)));
        nodeIndex = dag.get_child_index(nodeIndex, childMask, child);
    }

    if (col.kind == HDO_COLORS_UNCOMPRESSED) {
        if (!col.uncompressed || nof_leaves >= col.n_uncompressed) return invalid();
        const u32 raw = col.uncompressed[u32(nof_leaves)];
        // UncompressedColor (vwsc.h:55-84): debug accessors are zero, colour is the raw RGB888
        if (debugColors == 4) return set(0);
        if (debugColors == 5 || debugColors == 6 || debugColors == 7) return set(float3_to_rgb888({ 0, 0, 0 }));
        return set(float3_to_rgb888(rgb888_to_float3(raw)));
    }
    if (!leaf.is_valid() || !leaf.is_valid_index(nof_leaves)) return invalid();
    if (col.kind == HDO_COLORS_ERRORS) {
        if (!col.uncompressed || u32(nof_leaves) >= col.n_uncompressed) return invalid();
        const F3 a = leaf.get_color(u32(nof_leaves), dag.c).get_color(), b = rgb888_to_float3(col.uncompressed[u32(nof_leaves)]);
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        const float err = std::sqrt(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
        const float v = (double(err) > 0.04) ? 1.f : 0.f;  // basic_dag.h:191
        if (debugColors == 4) return set(0);
        if (debugColors == 5 || debugColors == 6 || debugColors == 7) return set(float3_to_rgb888({ 0, 0, 0 }));
        return set(float3_to_rgb888(rgb888_to_float3(float3_to_rgb888({ v, v, v }))));
    }
    const CompressedColor cc = leaf.get_color(nof_leaves, dag.c);
    u32 color;
    if (debugColors == 4) color = 0;                      // COLOR_DEBUG off (BENCHMARK build): hash is 0
    else if (debugColors == 5) color = float3_to_rgb888(cc.min_color());
    else if (debugColors == 6) color = float3_to_rgb888(cc.max_color());
    else if (debugColors == 7) { const float w = cc.get_weight(); color = float3_to_rgb888({ w, w, w }); }
    else color = float3_to_rgb888(cc.get_color());
    return set(color);
}

// sun_direction(), tracer.cu:546-549: normalize(float3(0.3,1,0.5)) folded by the compiler with
// IEEE single ops; its components match the immediates in the reference SASS.
inline F3 sun_direction()
{
    const float x = 0.3f, y = 1.f, z = 0.5f;
    const float len = std::sqrt(x * x + y * y + z * z);
    const float r = 1 / len;
    return { r * x, r * y, r * z };
}

}  // namespace

extern "C" {

void hdo_camera_params(const double pos[3], const double rot[9], const double bmin[3], const double bmax[3], uint32_t levels,
                       uint32_t width, uint32_t height, double cam_out[3], double ray_min[3], double ray_ddx[3], double ray_ddy[3])
{
    const double fov = double(60.f) / 2.0 * (double(M_PI) / 180.);
    const double aspect = double(width) / double(height);
    const double s = std::sin(fov), c = std::cos(fov);
    for (int k = 0; k < 3; ++k) {
        const double right = -rot[0 + k], up = rot[3 + k], fwd = rot[6 + k];
        const double X = right * s * aspect, Y = up * s, Z = fwd * c;
        const double bl = pos[k] + Z - Y - X, br = pos[k] + Z - Y + X, tl = pos[k] + Z + Y - X;
        const double translation = -bmin[k];
        const double scale = double(1 << levels) / (bmax[k] - bmin[k]);
        const double fp = (pos[k] + translation) * scale, fbl = (bl + translation) * scale;
        const double ftl = (tl + translation) * scale, fbr = (br + translation) * scale;
        cam_out[k] = fp; ray_min[k] = fbl;
        ray_ddx[k] = (fbr - fbl) * (1.0 / width);
        ray_ddy[k] = (ftl - fbl) * (1.0 / height);
    }
}

int hdo_trace_paths(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                    const double ddy[3], uint32_t* paths, uint32_t nThreads, hdo_stats* stats)
{
    if (!dag || !paths || dag->levels < 3 || dag->levels > 31) return 1;
    const F3 o = { float(cam[0]), float(cam[1]), float(cam[2]) };
    parallel_rows(H, nThreads, [&](u32 y, Counters& c) {
        if (y == ~0u) { add_stats(stats, c); return; }
        Dag d(*dag, c);
        for (u32 x = 0; x < W; ++x) {
            const D3 dd = primary_direction(cam, rmin, ddx, ddy, x, y);
            const F3 dir = { float(dd.x), float(dd.y), float(dd.z) };
            const F3 inv = { 1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z };
            const u8 order = u8((dir.x < 0.f ? 4 : 0) + (dir.y < 0.f ? 2 : 0) + (dir.z < 0.f ? 1 : 0));
            u32 px, py, pz;
            if (traverse<true>(d, o, dir, inv, order, px, py, pz)) ++c.hit;
            u32* out = paths + (u64(H - 1 - y) * W + x) * 4;
            out[0] = px; out[1] = py; out[2] = pz; out[3] = 0;
        }
    });
    return 0;
}

int hdo_trace_colors(const hdo_dag* dag, const hdo_colors* colors, uint32_t W, uint32_t H, const uint32_t* paths, int32_t debugColors,
                     uint32_t debugIndexLevel, const hdo_tool_info* tool, int32_t overlay, uint32_t* out, uint32_t nThreads, hdo_stats* stats)
{
    if (!dag || !colors || !paths || !out) return 1;
    parallel_rows(H, nThreads, [&](u32 y, Counters& c) {
        if (y == ~0u) { add_stats(stats, c); return; }
        Dag d(*dag, c);
        for (u32 x = 0; x < W; ++x) {
            const u32* p = paths + (u64(y) * W + x) * 4;
            out[u64(y) * W + x] = color_pixel(d, *colors, p[0], p[1], p[2], debugColors, debugIndexLevel, tool, overlay != 0);
        }
    });
    return 0;
}

int hdo_trace_shadows(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                      const double ddy[3], float shadowBias, float fogDensity, const uint32_t* paths, uint32_t* colors, uint32_t nThreads,
                      hdo_stats* stats)
{
    if (!dag || !paths || !colors) return 1;
    const F3 sun = sun_direction();
    const F3 sunInv = { 1.0f / sun.x, 1.0f / sun.y, 1.0f / sun.z };
    const float fd = fogDensity * 0.00001f;  // tracer.cu:564
    parallel_rows(H, nThreads, [&](u32 y, Counters& c) {
        if (y == ~0u) { add_stats(stats, c); return; }
        Dag d(*dag, c);
        for (u32 x = 0; x < W; ++x) {
            const u32* p = paths + (u64(y) * W + x) * 4;
            const F3 ro = { float(p[0]), float(p[1]), float(p[2]) };
            const D3 dir = primary_direction(cam, rmin, ddx, ddy, x, H - 1 - y);
            // setColor, tracer.cu:604-619 + applyFog :551-572
            auto shade = [&](float lightScale, double distance, D3 rd) {
                const F3 col = rgb888_to_float3(colors[u64(y) * W + x]);
                const F3 lit = { col.x * lightScale, col.y * lightScale, col.z * lightScale };
                const double fogAmount = 1.0 - std::exp(-distance * double(fd));
                const double dotp = std::fma(rd.z, double(sun.z), std::fma(rd.x, double(sun.x), rd.y * double(sun.y)));
                const double sunAmount = double(1.01f) * fmax(dotp, 0.0);
                const float pw = float(std::pow(sunAmount, 30.0)), q = 1.f - pw;
                const F3 fog = { fmaf(q, 187 / 255.f, pw), fmaf(q, 242 / 255.f, pw), fmaf(q, 250 / 255.f, pw) };
                const float g = clampf(float(fogAmount), 0.f, 1.f), h = 1.f - g;
                const F3 res = { fmaf(lit.x, h, g * fog.x), fmaf(lit.y, h, g * fog.y), fmaf(lit.z, h, g * fog.z) };
                colors[u64(y) * W + x] = float3_to_rgb888(res);
            };
            if (p[0] == 0 && p[1] == 0 && p[2] == 0) { shade(1.0f, 1e9, dir); continue; }
            // ray_box_intersection (tracer.cu:574-587) against the voxel [p, p+1]
            const double bo[3] = { double(ro.x), double(ro.y), double(ro.z) }, dv[3] = { dir.x, dir.y, dir.z };
            double rm[3];
            for (int k = 0; k < 3; ++k) {
                const double t0 = (bo[k] - cam[k]) / dv[k], t1 = ((bo[k] + 1.0) - cam[k]) / dv[k];
                rm[k] = (t0 < t1) ? t0 : t1;
            }
            const double maxmin = fmax(fmax(rm[0], rm[1]), rm[2]);
            const F3 start = { float(std::fma(dv[0], maxmin, cam[0])), float(std::fma(dv[1], maxmin, cam[1])), float(std::fma(dv[2], maxmin, cam[2])) };
            const F3 so = { fmaf(shadowBias, sun.x, start.x), fmaf(shadowBias, sun.y, start.y), fmaf(shadowBias, sun.z, start.z) };
            u32 hx, hy, hz;
            const bool shadowed = traverse<false>(d, so, sun, sunInv, 0, hx, hy, hz);
            ++c.hit;
            const double vx = bo[0] - cam[0], vy = bo[1] - cam[1], vz = bo[2] - cam[2];
            const double dist = std::sqrt(std::fma(vz, vz, std::fma(vx, vx, vy * vy)));
            shade(shadowed ? 0.5f : 1.0f, dist, { vx / dist, vy / dist, vz / dist });
        }
    });
    return 0;
}

int hdo_get_value(const hdo_dag* dag, uint32_t x, uint32_t y, uint32_t z)
{
    Counters c;
    Dag d(*dag, c);
    u32 nodeIndex = d.first();
    for (u32 level = 0; level < d.levels(); ++level) {
        if (level < d.leaf_level()) {
            const u8 childMask = u8(d.get_node(nodeIndex) & 0xFF);
            const u32 sh = d.levels() - (level + 1);
            const u8 child = u8((((x >> sh) & 1) ? 4 : 0) | (((y >> sh) & 1) ? 2 : 0) | (((z >> sh) & 1) ? 1 : 0));
            if (!(childMask & (1u << child))) return 0;
            nodeIndex = d.get_child_index(nodeIndex, childMask, child);
        } else {
            const u64 leaf = d.get_leaf(nodeIndex);
            const u8 bit = u8(((x & 1) ? 4 : 0) | ((y & 1) ? 2 : 0) | ((z & 1) ? 1 : 0) | ((x & 2) ? 32 : 0) | ((y & 2) ? 16 : 0) | ((z & 2) ? 8 : 0));
            return (leaf >> bit) & 1;
        }
    }
    return 1;
}

uint32_t hdo_decode_color(const hdo_color_leaf* leaf, uint64_t index)
{
    Counters c;
    ColorLeafView v; v.l = leaf; v.offset = 0;
    return float3_to_rgb888(v.get_color(index, c).get_color());
}

}  // extern "C"
