// Synthetic voxel-scene builder (see scene_builder.h).  Everything is integer / fixed-point so
// the same seed gives the same DAG on every machine (golden fixtures depend on that).
//
// Pipeline:  implicit scene --top-down recursion with conservative bounds, thread pool,
//            sharded dedupe tables--> unique nodes --canonical DFS renumber--> BasicDAG words
//            --restated hash_dag_factory.cpp:5-99--> HashDAG pool + page table
//            --DFS voxel walk per macro block--> compressed colour leaf (+ hash colour tree)
#include "scene_builder.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using i64 = int64_t;

constexpr u32 NONE = 0xFFFFFFFFu;

inline u32 fmix32(u32 h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
inline u64 fmix64(u64 h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
inline u32 hash2(i32 x, i32 z, u32 seed)
{
    return fmix32(fmix32(u32(x) * 0x9E3779B1u + seed) ^ (u32(z) * 0x85EBCA77u + 0x165667B1u));
}
inline u32 hash3(i32 x, i32 y, i32 z, u32 seed)
{
    return fmix32(hash2(x, z, seed) ^ (u32(y) * 0xC2B2AE3Du + 0x27D4EB2Fu));
}

// ---------------------------------------------------------------------------------------------
// Hashes the HashDAG uses to pick a bucket.  Same functions as the reference so that a DAG built
// here lands in the buckets the reference's find_or_add would search
// (/root/reference/src/utils.h:68-110, hash_table.h:80-89).
// ---------------------------------------------------------------------------------------------
inline u32 bucket_hash_interior(const u32* words, u32 n)
{
    u32 h = 0;
    for (u32 i = 0; i < n; ++i) {
        u32 k = words[i];
        k *= 0xcc9e2d51u; k = (k << 15) | (k >> 17); k *= 0x1b873593u;
        h ^= k; h = (h << 13) | (h >> 19); h = h * 5 + 0xe6546b64u;
    }
    h ^= n;
    return fmix32(h);
}
inline u32 bucket_hash_leaf(u64 leaf) { return u32(fmix64(leaf)); }

// ---------------------------------------------------------------------------------------------
// Terrain: fixed-point fractal value noise.  Lattices are aligned to the world origin, so an
// octree node no larger than an octave's cell lies inside exactly one cell, where the noise is
// multilinear in the (smoothstepped) local coordinates and attains its extrema at the corners.
// ---------------------------------------------------------------------------------------------
struct Terrain {
    u32 L = 0, F = 0, seed = 0, rough = 0;
    i32 fx0 = 0, fz0 = 0, fsize = 0;  // footprint
    i32 baseY = 0;
    int cmin = 2, cmax = 2;
    i64 amp[32] = {};      // amplitude per octave in 1/256 voxel
    i32 lipschitz = 0;     // bound on |H(x,z) - H(x+-1,z+-1)| in voxels
    i32 relief = 0;        // bound on |H - baseY|

    void init(u32 levels, u32 footLog2, u32 s, u32 roughness, u32 finest)
    {
        L = levels; F = footLog2; seed = s; rough = roughness;
        fsize = i32(1u << F);
        fx0 = fz0 = (F == L) ? 0 : i32((1u << (L - 1)) - (1u << (F - 1)));
        baseY = i32(1u << (L - 1));
        cmin = int(finest);
        cmax = std::max(cmin, int(F) - 2);
        i64 lip256 = 0, rel256 = 0;
        for (int c = cmin; c <= cmax; ++c) {
            amp[c] = (i64(1) << c) * 8 * (2 + (c - cmin));
            lip256 += 3 * amp[c] / (i64(1) << c) + 1;
            rel256 += amp[c];
        }
        lipschitz = i32((lip256 + 255) / 256) + 2 * i32(rough) + 2;
        relief = i32((rel256 + 255) / 256) + i32(rough) + 2;
    }

    static inline i64 smooth16(i64 t)  // t in [0,65536] -> smoothstep in [0,65536]
    {
        return (t * t * (3 * 65536 - 2 * t)) >> 32;
    }
    // Octave value in [-32768, 32767] inside lattice cell (ix,iz) at local offset (fx,fz) in [0,C].
    inline i64 octave_in_cell(int c, i32 ix, i32 iz, i32 fx, i32 fz) const
    {
        const u32 so = seed * 0x9E3779B1u + u32(c) * 0x7F4A7C15u;
        const i64 v00 = hash2(ix, iz, so) & 0xFFFF, v10 = hash2(ix + 1, iz, so) & 0xFFFF;
        const i64 v01 = hash2(ix, iz + 1, so) & 0xFFFF, v11 = hash2(ix + 1, iz + 1, so) & 0xFFFF;
        const i64 sx = smooth16(i64(fx) << (16 - c)), sz = smooth16(i64(fz) << (16 - c));
        const i64 a = v00 + (((v10 - v00) * sx) >> 16);
        const i64 b = v01 + (((v11 - v01) * sx) >> 16);
        return a + (((b - a) * sz) >> 16) - 32768;
    }
    inline i32 height(i32 x, i32 z) const
    {
        i64 acc = 0;
        for (int c = cmin; c <= cmax; ++c) {
            const i32 m = (1 << c) - 1;
            acc += (amp[c] * octave_in_cell(c, x >> c, z >> c, x & m, z & m)) >> 15;
        }
        i32 h = baseY + i32(acc >> 8);
        if (rough) h += i32(hash2(x, z, seed ^ 0xA5A5F00Du) % 3u) - 1;
        return h;
    }
    inline bool inside(i32 x, i32 z) const
    {
        return x >= fx0 && x < fx0 + fsize && z >= fz0 && z < fz0 + fsize;
    }
    // Conservative [lo,hi] of H over the closed square [x0,x0+s]^2 (s power of two, aligned).
    void bounds(i32 x0, i32 z0, i32 s, i32& lo, i32& hi) const
    {
        i64 mn = 0, mx = 0;
        for (int c = cmin; c <= cmax; ++c) {
            const i32 C = 1 << c;
            if (s > C) { mn -= amp[c]; mx += amp[c]; continue; }
            const i32 ix = x0 >> c, iz = z0 >> c, fx = x0 & (C - 1), fz = z0 & (C - 1);
            i64 v[4] = { octave_in_cell(c, ix, iz, fx, fz), octave_in_cell(c, ix, iz, fx + s, fz),
                         octave_in_cell(c, ix, iz, fx, fz + s), octave_in_cell(c, ix, iz, fx + s, fz + s) };
            i64 a = std::min(std::min(v[0], v[1]), std::min(v[2], v[3])) - 4;  // fixed-point rounding slack
            i64 b = std::max(std::max(v[0], v[1]), std::max(v[2], v[3])) + 4;
            mn += (amp[c] * a) >> 15;
            mx += ((amp[c] * b) >> 15) + 1;
        }
        lo = baseY + i32(mn >> 8) - 1 - i32(rough);
        hi = baseY + i32(mx >> 8) + 1 + i32(rough);
    }
};

struct Sphere { i64 cx, cy, cz, r, t; };

// ---------------------------------------------------------------------------------------------
// Dedupe tables (one per level, sharded, each shard behind a mutex).
// ---------------------------------------------------------------------------------------------
constexpr u32 SHARD_BITS = 6, N_SHARDS = 1u << SHARD_BITS, IDX_BITS = 32 - SHARD_BITS;

struct NodeShard {
    std::mutex m;
    std::vector<u32> slots;   // open addressing, value = idx+1
    std::vector<u32> words;   // per node: mask, count_lo, count_hi, children...
    std::vector<u32> offs;    // node idx -> offset in words
    NodeShard() : slots(1024, 0) {}
};
struct LeafShard {
    std::mutex m;
    std::vector<u32> slots;
    std::vector<u64> vals;
    LeafShard() : slots(1024, 0) {}
};

struct NodeTable {
    NodeShard sh[N_SHARDS];
    u32 find_or_add(u8 mask, u64 count, const u32* children, u32 n)
    {
        u32 h = mask * 0x9E3779B1u;
        for (u32 i = 0; i < n; ++i) h = fmix32(h ^ children[i]) + 0x7F4A7C15u;
        h = fmix32(h);
        NodeShard& s = sh[h >> IDX_BITS];
        std::lock_guard<std::mutex> g(s.m);
        if ((s.offs.size() + 1) * 2 > s.slots.size()) grow(s);
        size_t msk = s.slots.size() - 1, p = (h * 0x9E3779B1u) & msk;
        for (;; p = (p + 1) & msk) {
            const u32 v = s.slots[p];
            if (!v) break;
            const u32* w = &s.words[s.offs[v - 1]];
            if (w[0] == mask && !memcmp(w + 3, children, n * 4)) return ((h >> IDX_BITS) << IDX_BITS) | (v - 1);
        }
        const u32 idx = u32(s.offs.size());
        if (idx >= (1u << IDX_BITS)) { fprintf(stderr, "scene_builder: node shard overflow\n"); abort(); }
        s.offs.push_back(u32(s.words.size()));
        s.words.push_back(mask); s.words.push_back(u32(count)); s.words.push_back(u32(count >> 32));
        s.words.insert(s.words.end(), children, children + n);
        s.slots[p] = idx + 1;
        return ((h >> IDX_BITS) << IDX_BITS) | idx;
    }
    static u32 rehash(const u32* w)
    {
        u32 h = w[0] * 0x9E3779B1u;
        const u32 n = u32(__builtin_popcount(w[0]));
        for (u32 i = 0; i < n; ++i) h = fmix32(h ^ w[3 + i]) + 0x7F4A7C15u;
        return fmix32(h);
    }
    static void grow(NodeShard& s)
    {
        std::vector<u32> ns(s.slots.size() * 4, 0);
        size_t msk = ns.size() - 1;
        for (u32 i = 0; i < s.offs.size(); ++i) {
            size_t p = (rehash(&s.words[s.offs[i]]) * 0x9E3779B1u) & msk;
            while (ns[p]) p = (p + 1) & msk;
            ns[p] = i + 1;
        }
        s.slots.swap(ns);
    }
    const u32* node(u32 id) const { const NodeShard& s = sh[id >> IDX_BITS]; return &s.words[s.offs[id & ((1u << IDX_BITS) - 1)]]; }
};

struct LeafTable {
    LeafShard sh[N_SHARDS];
    u32 find_or_add(u64 v)
    {
        const u32 h = u32(fmix64(v ^ 0x9E3779B97F4A7C15ull));
        LeafShard& s = sh[h >> IDX_BITS];
        std::lock_guard<std::mutex> g(s.m);
        if ((s.vals.size() + 1) * 2 > s.slots.size()) {
            std::vector<u32> ns(s.slots.size() * 4, 0);
            size_t msk = ns.size() - 1;
            for (u32 i = 0; i < s.vals.size(); ++i) {
                size_t p = (u32(fmix64(s.vals[i] ^ 0x9E3779B97F4A7C15ull)) * 0x9E3779B1u) & msk;
                while (ns[p]) p = (p + 1) & msk;
                ns[p] = i + 1;
            }
            s.slots.swap(ns);
        }
        size_t msk = s.slots.size() - 1, p = (h * 0x9E3779B1u) & msk;
        for (;; p = (p + 1) & msk) {
            const u32 e = s.slots[p];
            if (!e) break;
            if (s.vals[e - 1] == v) return ((h >> IDX_BITS) << IDX_BITS) | (e - 1);
        }
        const u32 idx = u32(s.vals.size());
        if (idx >= (1u << IDX_BITS)) { fprintf(stderr, "scene_builder: leaf shard overflow\n"); abort(); }
        s.vals.push_back(v);
        s.slots[p] = idx + 1;
        return ((h >> IDX_BITS) << IDX_BITS) | idx;
    }
    u64 leaf(u32 id) const { return sh[id >> IDX_BITS].vals[id & ((1u << IDX_BITS) - 1)]; }
};

struct Ref { u32 id = NONE; u64 count = 0; };

// Per-thread cache of exact column extents for a 16x16 tile: lo[x][z] <= y < hi[x][z] is solid.
struct Tile {
    i32 x0 = INT32_MIN, z0 = INT32_MIN;
    i32 lo[16][16], hi[16][16];
    i32 minLo, maxHi;
};

}  // namespace

struct hds_scene {
    hds_params P{};
    Terrain T;
    std::vector<Sphere> spheres;
    u32 L = 0, leafLevel = 0, topLevels = 0;

    std::vector<NodeTable*> tables;  // levels 0 .. L-3
    LeafTable leaves;                // level L-2

    // outputs
    std::vector<u32> basic;
    std::vector<u64> enclosed;
    std::vector<u32> pool, pageTable;
    u32 poolTop = 1, firstNodeIndex = 0;
    std::vector<u32> weights;
    std::vector<u64> blocks, macroBlocks;
    std::vector<u32> uncompressed;
    std::vector<u32> colorNodes;
    std::vector<u64> colorOffsets;
    u64 nVoxels = 0;
    u64 nodesPerLevel[32] = {};
    double seconds = 0;

    ~hds_scene() { for (auto* t : tables) delete t; }

    // ------------------------------------------------------------------ implicit scene
    void fill_tile(Tile& t, i32 x0, i32 z0) const
    {
        // Heights of the 18x18 columns around the tile.  Same value as Terrain::height() per
        // column, but the lattice hashes and smoothstep weights of each octave are hoisted
        // out of the per-column loop.
        i64 acc[18][18] = {};
        for (int c = T.cmin; c <= T.cmax; ++c) {
            const u32 so = T.seed * 0x9E3779B1u + u32(c) * 0x7F4A7C15u;
            const i32 m = (1 << c) - 1;
            const i32 ixa = (x0 - 1) >> c, iza = (z0 - 1) >> c;
            const int nx = int(((x0 + 16) >> c) - ixa) + 2, nz = int(((z0 + 16) >> c) - iza) + 2;  // <= 7
            i64 lat[8][8];
            for (int i = 0; i < nx; ++i)
                for (int k = 0; k < nz; ++k) lat[i][k] = hash2(ixa + i, iza + k, so) & 0xFFFF;
            i64 sx[18], sz[18]; int cx[18], cz[18];
            for (int i = 0; i < 18; ++i) {
                const i32 x = x0 - 1 + i, z = z0 - 1 + i;
                sx[i] = Terrain::smooth16(i64(x & m) << (16 - c)); cx[i] = int((x >> c) - ixa);
                sz[i] = Terrain::smooth16(i64(z & m) << (16 - c)); cz[i] = int((z >> c) - iza);
            }
            const i64 A = T.amp[c];
            for (int i = 0; i < 18; ++i)
                for (int k = 0; k < 18; ++k) {
                    const i64 v00 = lat[cx[i]][cz[k]], v10 = lat[cx[i] + 1][cz[k]];
                    const i64 v01 = lat[cx[i]][cz[k] + 1], v11 = lat[cx[i] + 1][cz[k] + 1];
                    const i64 a = v00 + (((v10 - v00) * sx[i]) >> 16);
                    const i64 b = v01 + (((v11 - v01) * sx[i]) >> 16);
                    acc[i][k] += (A * (a + (((b - a) * sz[k]) >> 16) - 32768)) >> 15;
                }
        }
        i32 h[18][18];
        for (int i = 0; i < 18; ++i)
            for (int k = 0; k < 18; ++k) {
                const i32 x = x0 - 1 + i, z = z0 - 1 + k;
                if (!T.inside(x, z)) { h[i][k] = INT32_MIN; continue; }
                i32 v = T.baseY + i32(acc[i][k] >> 8);
                if (T.rough) v += i32(hash2(x, z, T.seed ^ 0xA5A5F00Du) % 3u) - 1;
                h[i][k] = v;
            }
        t.x0 = x0; t.z0 = z0; t.minLo = INT32_MAX; t.maxHi = INT32_MIN;
        for (int i = 0; i < 16; ++i)
            for (int k = 0; k < 16; ++k) {
                const i32 H = h[i + 1][k + 1];
                i32 lo, hi;
                if (H == INT32_MIN) { lo = 0; hi = 0; }
                else {
                    // outside the footprint the neighbour reads as a 1-voxel step (keeps the rim thin)
                    i32 nb = H - 1;
                    const i32 n4[4] = { h[i][k + 1], h[i + 2][k + 1], h[i + 1][k], h[i + 1][k + 2] };
                    for (i32 v : n4) if (v != INT32_MIN) nb = std::min(nb, v);
                    lo = nb; hi = H;
                }
                t.lo[i][k] = lo; t.hi[i][k] = hi;
                if (hi > lo) { t.minLo = std::min(t.minLo, lo); t.maxHi = std::max(t.maxHi, hi); }
            }
    }
    const Tile& tile_for(Tile* cache, i32 x0, i32 z0) const
    {
        Tile& t = cache[((x0 >> 4) & 1) * 2 + ((z0 >> 4) & 1)];
        if (t.x0 != x0 || t.z0 != z0) fill_tile(t, x0, z0);
        return t;
    }
    bool terrain_overlaps(Tile* cache, i32 x0, i32 y0, i32 z0, i32 s) const
    {
        // footprint is aligned to 2^(F-1): a node of that size or smaller is inside or outside
        if (s <= T.fsize / 2 || T.F == T.L) {
            if (!T.inside(x0, z0)) return false;
        } else if (x0 + s <= T.fx0 || x0 >= T.fx0 + T.fsize || z0 + s <= T.fz0 || z0 >= T.fz0 + T.fsize) {
            return false;
        }
        if (s <= 16) {
            const Tile& t = tile_for(cache, x0 & ~15, z0 & ~15);
            if (s == 16) return t.maxHi > y0 && t.minLo < y0 + s;
            const int bx = x0 & 15, bz = z0 & 15;
            for (int i = 0; i < s; ++i)
                for (int k = 0; k < s; ++k)
                    if (t.hi[bx + i][bz + k] > y0 && t.lo[bx + i][bz + k] < y0 + s && t.hi[bx + i][bz + k] > t.lo[bx + i][bz + k]) return true;
            return false;
        }
        if (y0 >= T.baseY + T.relief || y0 + s <= T.baseY - T.relief - T.lipschitz) return false;
        i32 lo, hi;
        T.bounds(x0, z0, s, lo, hi);
        return y0 < hi && y0 + s > lo - 1 - T.lipschitz;
    }
    static bool sphere_overlaps(const Sphere& sp, i64 x0, i64 y0, i64 z0, i64 s)
    {
        auto axis = [](i64 c, i64 a, i64 b, i64& dmin, i64& dmax) {  // voxels a..b inclusive
            dmin = c < a ? a - c : (c > b ? c - b : 0);
            dmax = std::max(c > a ? c - a : a - c, c > b ? c - b : b - c);
        };
        i64 nx, fx, ny, fy, nz, fz;
        axis(sp.cx, x0, x0 + s - 1, nx, fx); axis(sp.cy, y0, y0 + s - 1, ny, fy); axis(sp.cz, z0, z0 + s - 1, nz, fz);
        const i64 dmin2 = nx * nx + ny * ny + nz * nz, dmax2 = fx * fx + fy * fy + fz * fz;
        const i64 ri = sp.r - sp.t;
        return dmin2 <= sp.r * sp.r && dmax2 >= ri * ri;
    }
    static bool sphere_voxel(const Sphere& sp, i64 x, i64 y, i64 z)
    {
        const i64 d2 = (x - sp.cx) * (x - sp.cx) + (y - sp.cy) * (y - sp.cy) + (z - sp.cz) * (z - sp.cz);
        const i64 ri = sp.r - sp.t;
        return d2 <= sp.r * sp.r && d2 >= ri * ri;
    }
    u64 spheres_overlapping(i32 x0, i32 y0, i32 z0, i32 s, u64 candidates) const
    {
        u64 out = 0;
        for (u64 m = candidates; m; m &= m - 1) {
            const int i = __builtin_ctzll(m);
            if (sphere_overlaps(spheres[i], x0, y0, z0, s)) out |= u64(1) << i;
        }
        return out;
    }

    // ------------------------------------------------------------------ recursion
    Ref build(Tile* cache, u32 level, i32 x0, i32 y0, i32 z0, u64 sphereCand)
    {
        const i32 s = i32(1u << (L - level));
        const bool terr = terrain_overlaps(cache, x0, y0, z0, s);
        const u64 sph = spheres_overlapping(x0, y0, z0, s, sphereCand);
        if (!terr && !sph) return {};
        if (level == leafLevel) return build_leaf(cache, terr, x0, y0, z0, sph);
        u32 children[8]; u32 n = 0; u8 mask = 0; u64 count = 0;
        const i32 h = s / 2;
        for (u32 c = 0; c < 8; ++c) {
            const Ref r = build(cache, level + 1, x0 + ((c & 4) ? h : 0), y0 + ((c & 2) ? h : 0), z0 + ((c & 1) ? h : 0), sph);
            if (r.id != NONE) { children[n++] = r.id; mask |= u8(1u << c); count += r.count; }
        }
        if (!n) return {};
        return { tables[level]->find_or_add(mask, count, children, n), count };
    }
    Ref build_leaf(Tile* cache, bool terr, i32 x0, i32 y0, i32 z0, u64 sph)
    {
        u64 bits = 0;
        const Tile* t = terr ? &tile_for(cache, x0 & ~15, z0 & ~15) : nullptr;
        const int bx = x0 & 15, bz = z0 & 15;
        for (int x = 0; x < 4; ++x)
            for (int z = 0; z < 4; ++z) {
                const i32 lo = t ? t->lo[bx + x][bz + z] : 0, hi = t ? t->hi[bx + x][bz + z] : 0;
                for (int y = 0; y < 4; ++y) {
                    bool solid = (y0 + y) >= lo && (y0 + y) < hi;
                    for (u64 m = sph; m && !solid; m &= m - 1)
                        solid = sphere_voxel(spheres[__builtin_ctzll(m)], x0 + x, y0 + y, z0 + z);
                    if (solid) {
                        // leaf bit = child1*8 + child2 (/root/reference/src/dags/dag_utils.h:160-167)
                        const int c1 = ((x & 2) ? 4 : 0) | ((y & 2) ? 2 : 0) | ((z & 2) ? 1 : 0);
                        const int c2 = ((x & 1) ? 4 : 0) | ((y & 1) ? 2 : 0) | ((z & 1) ? 1 : 0);
                        bits |= u64(1) << (c1 * 8 + c2);
                    }
                }
            }
        if (!bits) return {};
        return { leaves.find_or_add(bits), u64(__builtin_popcountll(bits)) };
    }

    struct Task { i32 x, y, z; };
    void enumerate_tasks(Tile* cache, u32 level, u32 splitLevel, i32 x0, i32 y0, i32 z0, u64 cand, std::vector<Task>& out,
                         std::vector<u64>& outCand)
    {
        const i32 s = i32(1u << (L - level));
        const bool terr = terrain_overlaps(cache, x0, y0, z0, s);
        const u64 sph = spheres_overlapping(x0, y0, z0, s, cand);
        if (!terr && !sph) return;
        if (level == splitLevel) { out.push_back({ x0, y0, z0 }); outCand.push_back(sph); return; }
        const i32 h = s / 2;
        for (u32 c = 0; c < 8; ++c)
            enumerate_tasks(cache, level + 1, splitLevel, x0 + ((c & 4) ? h : 0), y0 + ((c & 2) ? h : 0), z0 + ((c & 1) ? h : 0), sph, out, outCand);
    }
    Ref assemble(Tile* cache, u32 level, u32 splitLevel, i32 x0, i32 y0, i32 z0, u64 cand, const std::vector<Ref>& results, size_t& cursor)
    {
        const i32 s = i32(1u << (L - level));
        const bool terr = terrain_overlaps(cache, x0, y0, z0, s);
        const u64 sph = spheres_overlapping(x0, y0, z0, s, cand);
        if (!terr && !sph) return {};
        if (level == splitLevel) return results[cursor++];
        u32 children[8]; u32 n = 0; u8 mask = 0; u64 count = 0;
        const i32 h = s / 2;
        for (u32 c = 0; c < 8; ++c) {
            const Ref r = assemble(cache, level + 1, splitLevel, x0 + ((c & 4) ? h : 0), y0 + ((c & 2) ? h : 0), z0 + ((c & 1) ? h : 0), sph, results, cursor);
            if (r.id != NONE) { children[n++] = r.id; mask |= u8(1u << c); count += r.count; }
        }
        if (!n) return {};
        return { tables[level]->find_or_add(mask, count, children, n), count };
    }

    // ------------------------------------------------------------------ BasicDAG serialisation
    // Canonical numbering = first visit of a pre-order DFS from the root (children in index
    // order), so the word layout does not depend on thread interleaving.  Layout is level-major.
    void serialise_basic(Ref root)
    {
        const u32 nLevels = leafLevel + 1;
        std::vector<std::vector<u32>> newId(nLevels);       // per level: old (shard-packed, compacted) -> canonical
        std::vector<std::vector<u32>> order(nLevels);       // canonical -> old id
        std::vector<std::vector<u32>> shardBase(nLevels, std::vector<u32>(N_SHARDS + 1, 0));
        for (u32 l = 0; l < nLevels; ++l) {
            for (u32 s = 0; s < N_SHARDS; ++s) {
                const size_t n = (l == leafLevel) ? leaves.sh[s].vals.size() : tables[l]->sh[s].offs.size();
                shardBase[l][s + 1] = shardBase[l][s] + u32(n);
            }
            newId[l].assign(shardBase[l][N_SHARDS], NONE);
            order[l].reserve(shardBase[l][N_SHARDS]);
        }
        auto compact = [&](u32 l, u32 id) { return shardBase[l][id >> IDX_BITS] + (id & ((1u << IDX_BITS) - 1)); };

        struct Frame { u32 level, id, next; };
        std::vector<Frame> stack;
        auto visit = [&](u32 l, u32 id) {
            u32& slot = newId[l][compact(l, id)];
            if (slot != NONE) return false;
            slot = u32(order[l].size());
            order[l].push_back(id);
            return true;
        };
        visit(0, root.id);
        stack.push_back({ 0, root.id, 0 });
        while (!stack.empty()) {
            Frame& f = stack.back();
            const u32* w = tables[f.level]->node(f.id);
            const u32 n = u32(__builtin_popcount(w[0]));
            if (f.next == n) { stack.pop_back(); continue; }
            const u32 child = w[3 + f.next++];
            const u32 cl = f.level + 1;
            if (visit(cl, child) && cl < leafLevel) stack.push_back({ cl, child, 0 });
        }

        // word offsets
        std::vector<u64> levelBase(nLevels + 1, 0);
        std::vector<std::vector<u32>> nodeOff(nLevels);
        u64 total = 0;
        for (u32 l = 0; l < nLevels; ++l) {
            levelBase[l] = total;
            nodesPerLevel[l] = order[l].size();
            if (l == leafLevel) { total += 2 * u64(order[l].size()); continue; }
            nodeOff[l].resize(order[l].size());
            for (size_t i = 0; i < order[l].size(); ++i) {
                nodeOff[l][i] = u32(total - levelBase[l]);
                total += 1 + u32(__builtin_popcount(tables[l]->node(order[l][i])[0]));
            }
        }
        if (total >= (u64(1) << 32)) { fprintf(stderr, "scene_builder: BasicDAG exceeds 2^32 words\n"); abort(); }
        basic.assign(total, 0);
        enclosed.clear();
        for (u32 l = 0; l < leafLevel; ++l) {
            for (size_t i = 0; i < order[l].size(); ++i) {
                const u32* w = tables[l]->node(order[l][i]);
                const u64 count = u64(w[1]) | (u64(w[2]) << 32);
                u32 upper;
                if (l < topLevels) { upper = u32(enclosed.size()); enclosed.push_back(count); }
                else upper = u32(count);
                if (upper >= (1u << 24)) { fprintf(stderr, "scene_builder: 24-bit count overflow at level %u\n", l); abort(); }
                u32* dst = &basic[levelBase[l] + nodeOff[l][i]];
                dst[0] = w[0] | (upper << 8);
                const u32 n = u32(__builtin_popcount(w[0]));
                for (u32 c = 0; c < n; ++c) {
                    const u32 cid = newId[l + 1][compact(l + 1, w[3 + c])];
                    dst[1 + c] = (l + 1 == leafLevel) ? u32(levelBase[l + 1] + 2 * u64(cid)) : u32(levelBase[l + 1] + nodeOff[l + 1][cid]);
                }
            }
        }
        for (size_t i = 0; i < order[leafLevel].size(); ++i) {
            const u64 v = leaves.leaf(order[leafLevel][i]);
            basic[levelBase[leafLevel] + 2 * i] = u32(v);
            basic[levelBase[leafLevel] + 2 * i + 1] = u32(v >> 32);
        }
        if (enclosed.empty()) enclosed.push_back(0);
    }

    // ------------------------------------------------------------------ HashDAG
    // Virtual address space and insertion rules: hash_table.h:45-71 (make_ptr), :359-400
    // (add_leaf_node), :401-468 (add_interior_node, nodes never straddle a page), :796-812
    // (allocate_page: physical pages handed out in first-touch order starting at 1).
    static constexpr u32 PAGE = 512, TOP_LEVELS = 9, TOP_BUCKETS = 1024, LOW_BUCKETS = 65536, TOP_BSIZE = 1024, LOW_BSIZE = 4096;
    std::vector<u32> bucketSizes;
    u32 total_pages() const { return TOP_LEVELS * TOP_BUCKETS * (TOP_BSIZE / PAGE) + (L - TOP_LEVELS) * LOW_BUCKETS * (LOW_BSIZE / PAGE); }
    static u32 make_ptr(u32 level, u32 bucket, u32 pos)
    {
        if (level < TOP_LEVELS) return (level * TOP_BUCKETS + bucket) * TOP_BSIZE + pos;
        return TOP_LEVELS * TOP_BUCKETS * TOP_BSIZE + ((level - TOP_LEVELS) * LOW_BUCKETS + bucket) * LOW_BSIZE + pos;
    }
    static u32 bucket_global(u32 level, u32 bucket)
    {
        return level < TOP_LEVELS ? level * TOP_BUCKETS + bucket : TOP_LEVELS * TOP_BUCKETS + (level - TOP_LEVELS) * LOW_BUCKETS + bucket;
    }
    u32* sys_ptr(u32 vptr)
    {
        u32& pe = pageTable[vptr / PAGE];
        if (!pe) {
            pe = poolTop++;
            pool.resize(size_t(poolTop) * PAGE, 0);
        }
        return &pool[size_t(pe) * PAGE + vptr % PAGE];
    }
    u32 add_leaf_node(u64 leaf)
    {
        const u32 bucket = bucket_hash_leaf(leaf) & (LOW_BUCKETS - 1);
        u32& bs = bucketSizes[bucket_global(leafLevel, bucket)];
        const u32 ptr = make_ptr(leafLevel, bucket, bs);
        u32* d = sys_ptr(ptr);
        d[0] = u32(leaf); d[1] = u32(leaf >> 32);
        bs += 2;
        if (bs >= LOW_BSIZE) { fprintf(stderr, "scene_builder: leaf bucket overflow\n"); abort(); }
        return ptr;
    }
    u32 add_interior_node(u32 level, const u32* node, u32 n)
    {
        const u32 nb = level < TOP_LEVELS ? TOP_BUCKETS : LOW_BUCKETS, cap = level < TOP_LEVELS ? TOP_BSIZE : LOW_BSIZE;
        const u32 bucket = bucket_hash_interior(node, n) & (nb - 1);
        u32& bs = bucketSizes[bucket_global(level, bucket)];
        const u32 left = PAGE - (bs % PAGE);
        if (left != PAGE && left < n) bs += left;
        const u32 ptr = make_ptr(level, bucket, bs);
        if (bs + n >= cap) { fprintf(stderr, "scene_builder: bucket overflow on level %u\n", level); abort(); }
        u32* d = sys_ptr(ptr);
        memcpy(d, node, n * 4);
        bs += n;
        return ptr;
    }
    void build_hash()
    {
        if (L < 10 || L > 24) { fprintf(stderr, "scene_builder: HashDAG needs 10 <= levels <= 24\n"); abort(); }
        pageTable.assign(total_pages(), 0);
        bucketSizes.assign(TOP_LEVELS * TOP_BUCKETS + (L - TOP_LEVELS) * LOW_BUCKETS, 0);
        pool.assign(PAGE, 0);  // physical page 0 stays unused (hash_table.h:821)
        pool.reserve(basic.size() + basic.size() / 4 + PAGE);
        poolTop = 1;
        // full nodes first (hash_dag_factory.cpp:162-169, hash_table.h:749-789)
        {
            u32 below = add_leaf_node(~u64(0));
            for (u32 level = leafLevel - 1; level > 0; --level) {
                const u32 vox = level < 10 ? 0u : (1u << (3 * (L - level)));
                u32 nb[9] = { (vox << 8) | 0xFF };
                for (int c = 0; c < 8; ++c) nb[1 + c] = below;
                below = add_interior_node(level, nb, 9);
            }
        }
        // post-order DFS with memo over the BasicDAG (hash_dag_factory.cpp:5-99)
        std::vector<u32> memo(basic.size(), 0);
        struct Frame { u32 level, index, next, n; u32 buf[9]; };
        std::vector<Frame> st;
        auto open = [&](u32 level, u32 index) {
            Frame f{}; f.level = level; f.index = index; f.next = 0;
            f.buf[0] = basic[index]; f.n = 1;
            st.push_back(f);
        };
        open(0, 0);
        u32 result = 0;
        while (!st.empty()) {
            Frame& f = st.back();
            const u32 mask = f.buf[0] & 0xFF;
            bool descended = false;
            while (f.next < 8) {
                const u32 c = f.next;
                if (!(mask & (1u << c))) { ++f.next; continue; }
                const u32 childIndex = basic[f.index + 1 + __builtin_popcount(mask & ((1u << c) - 1))];
                if (memo[childIndex]) { f.buf[f.n++] = memo[childIndex] - 1; ++f.next; continue; }
                if (f.level + 1 == leafLevel) {
                    const u64 leaf = u64(basic[childIndex]) | (u64(basic[childIndex + 1]) << 32);
                    const u32 p = add_leaf_node(leaf);
                    memo[childIndex] = p + 1;
                    f.buf[f.n++] = p; ++f.next; continue;
                }
                open(f.level + 1, childIndex);  // invalidates f
                descended = true;
                break;
            }
            if (descended) continue;
            const u32 p = add_interior_node(f.level, f.buf, f.n);
            const u32 idx = f.index;
            st.pop_back();
            if (st.empty()) result = p;
            else { memo[idx] = p + 1; /* parent picks it up from memo on its next loop turn */ }
        }
        firstNodeIndex = result;
    }

    // ------------------------------------------------------------------ colours
    u64 node_count(u32 level, u32 hdr) const { return level < topLevels ? enclosed[hdr >> 8] : (hdr >> 8); }

    struct VoxelColor { u32 colorBits; u8 bpw; u8 weight; u32 rgb888; };
    VoxelColor voxel_color(i32 x, i32 y, i32 z) const
    {
        static const u8 bpwTable[8] = { 0, 1, 2, 3, 4, 2, 3, 1 };
        const u32 hc = hash3(x >> 3, y >> 3, z >> 3, P.seed ^ 0xC0105EEDu);
        const u8 bpw = bpwTable[hc & 7];
        // height ramp: sand -> grass -> rock -> snow
        static const i32 ramp[5][3] = { { 194, 178, 128 }, { 80, 150, 60 }, { 50, 110, 45 }, { 125, 115, 105 }, { 240, 240, 250 } };
        i64 rel = (i64(y) - (T.baseY - T.relief / 2)) * 1024 / std::max<i64>(1, T.relief);  // 0..1024 over the band
        rel = std::min<i64>(1023, std::max<i64>(0, rel));
        const int seg = int(rel >> 8), f = int(rel & 255);
        i32 base[3];
        for (int k = 0; k < 3; ++k) base[k] = (ramp[seg][k] * (256 - f) + ramp[seg + 1][k] * f) >> 8;
        for (int k = 0; k < 3; ++k) base[k] = std::min(255, std::max(0, base[k] + i32((hc >> (8 + 5 * k)) & 31) - 16));
        VoxelColor vc{};
        vc.bpw = bpw;
        if (bpw == 0) {
            vc.colorBits = u32(base[0] * 1023 / 255) | (u32(base[1] * 4095 / 255) << 10) | (u32(base[2] * 1023 / 255) << 22);
            vc.weight = 0;
            vc.rgb888 = 0xFF000000u | u32(base[0]) | (u32(base[1]) << 8) | (u32(base[2]) << 16);
        } else {
            i32 lo[3], hi[3];
            for (int k = 0; k < 3; ++k) { lo[k] = std::max(0, base[k] - 28); hi[k] = std::min(255, base[k] + 28); }
            auto to565 = [](const i32* c) { return u32(c[0] * 31 / 255) | (u32(c[1] * 63 / 255) << 5) | (u32(c[2] * 31 / 255) << 11); };
            vc.colorBits = to565(lo) | (to565(hi) << 16);
            const u32 maxw = (1u << bpw) - 1;
            vc.weight = u8(hash3(x, y, z, P.seed ^ 0x5EED5EEDu) & maxw);
            i32 c[3];
            for (int k = 0; k < 3; ++k) c[k] = lo[k] + (hi[k] - lo[k]) * i32(vc.weight) / i32(maxw);
            // one cell in 16 is deliberately off so the error view has both outcomes
            if (((hc >> 24) & 15) == 0) { c[0] = 255 - c[0]; c[2] = 255 - c[2]; }
            vc.rgb888 = 0xFF000000u | u32(c[0]) | (u32(c[1]) << 8) | (u32(c[2]) << 16);
        }
        return vc;
    }

    // Emit voxels [skip, skip+want) of the subtree in colour order (children ascending, leaf
    // bits ascending: tracer.cu:388-408 counts exactly this order).
    template <class F>
    void walk(u32 level, u32 index, i32 x, i32 y, i32 z, u64& skip, u64& want, F& emit) const
    {
        if (level == leafLevel) {
            const u64 bits = u64(basic[index]) | (u64(basic[index + 1]) << 32);
            for (u64 m = bits; m && want; m &= m - 1) {
                if (skip) { --skip; continue; }
                const int b = __builtin_ctzll(m), c1 = b >> 3, c2 = b & 7;
                emit(x + ((c1 & 4) ? 2 : 0) + ((c2 & 4) ? 1 : 0), y + ((c1 & 2) ? 2 : 0) + ((c2 & 2) ? 1 : 0), z + ((c1 & 1) ? 2 : 0) + ((c2 & 1) ? 1 : 0));
                --want;
            }
            return;
        }
        const u32 hdr = basic[index], mask = hdr & 0xFF;
        const i32 h = i32(1u << (L - level - 1));
        u32 k = 0;
        for (u32 c = 0; c < 8 && want; ++c) {
            if (!(mask & (1u << c))) continue;
            const u32 child = basic[index + 1 + k++];
            const u64 cnt = (level + 1 == leafLevel) ? u64(__builtin_popcountll(u64(basic[child]) | (u64(basic[child + 1]) << 32)))
                                                     : node_count(level + 1, basic[child]);
            if (skip >= cnt) { skip -= cnt; continue; }
            walk(level + 1, child, x + ((c & 4) ? h : 0), y + ((c & 2) ? h : 0), z + ((c & 1) ? h : 0), skip, want, emit);
        }
    }

    void build_colors(u32 nThreads)
    {
        constexpr u64 MACRO = 16 * 1024;  // variable_weight_size_colors.h:8
        const u64 nMacro = (nVoxels + MACRO - 1) / MACRO;
        struct MacroOut { std::vector<u64> blocks; std::vector<u32> weights; u32 nbits = 0; };
        std::vector<MacroOut> outs(nMacro);
        if (P.build_uncompressed) uncompressed.assign(nVoxels, 0);
        std::atomic<u64> next{ 0 };
        auto worker = [&]() {
            for (;;) {
                const u64 m = next.fetch_add(1);
                if (m >= nMacro) break;
                MacroOut& o = outs[m];
                u64 skip = m * MACRO, want = std::min<u64>(MACRO, nVoxels - m * MACRO);
                u32 local = 0, lastBits = 0, lastBpw = 0xFF;
                u64 acc = 0; int accBits = 0;  // MSB-first bit accumulator
                u64 outIdx = m * MACRO;
                auto emit = [&](i32 x, i32 y, i32 z) {
                    VoxelColor vc = voxel_color(x, y, z);
                    // new block on a macro boundary or when the colour pair / weight width changes
                    // (same rule as ColorLeafBuilder::add, vwsc.h:600-618)
                    if (local == 0 || vc.colorBits != lastBits || vc.bpw != lastBpw) {
                        // 0xFFFF in the offset field means "no weights" (vwsc.h:20-30); dodge it
                        if (vc.bpw && o.nbits == 0xFFFF) { vc.bpw = 0; vc.weight = 0; }
                        const u32 hdr = vc.bpw ? ((o.nbits << 16) | (u32(vc.bpw - 1) << 14) | local) : ((0xFFFFu << 16) | local);
                        o.blocks.push_back((u64(vc.colorBits) << 32) | hdr);
                        lastBits = vc.colorBits; lastBpw = vc.bpw;
                    }
                    if (vc.bpw) {
                        acc = (acc << vc.bpw) | vc.weight; accBits += vc.bpw; o.nbits += vc.bpw;
                        while (accBits >= 32) { o.weights.push_back(u32(acc >> (accBits - 32))); accBits -= 32; }
                    }
                    if (P.build_uncompressed) uncompressed[outIdx] = vc.rgb888;
                    ++outIdx; ++local;
                };
                walk(0, 0, 0, 0, 0, skip, want, emit);
                // flush to a 32-bit boundary: each macro block starts word-aligned here, which the
                // format allows (macroBlocks[2m+1] is an arbitrary bit offset, vwsc.h:187-191)
                if (accBits) { o.weights.push_back(u32(acc << (32 - accBits))); o.nbits += u32(32 - accBits); accBits = 0; }
            }
        };
        std::vector<std::thread> th;
        for (u32 i = 0; i < nThreads; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
        blocks.clear(); weights.clear(); macroBlocks.assign(2 * nMacro, 0);
        u64 nb = 0, nw = 0;
        for (auto& o : outs) { nb += o.blocks.size(); nw += o.weights.size(); }
        blocks.reserve(nb); weights.reserve(nw + 1);
        for (u64 m = 0; m < nMacro; ++m) {
            macroBlocks[2 * m] = blocks.size();
            macroBlocks[2 * m + 1] = u64(weights.size()) * 32;
            blocks.insert(blocks.end(), outs[m].blocks.begin(), outs[m].blocks.end());
            // stored byte-swapped (CFG_COLOR_SWAP_BYTE_ORDER, color_utils.h:84-105, vwsc.h:662-669)
            for (u32 w : outs[m].weights) weights.push_back(__builtin_bswap32(w));
            outs[m] = MacroOut();
        }
        weights.push_back(0);  // extract_bits reads 2 bytes at the last bit's byte
    }

    // hash_dag_factory.cpp:101-150
    u32 build_hash_color_tree(u32 level, u32 index, u64 leavesCount)  // by value, like the reference
    {
        const u32 hdr = basic[index], mask = hdr & 0xFF;
        const u32 colorIndex = u32(colorNodes.size());
        u32 k = 0;
        if (level == 10 - 1) {
            for (u32 c = 0; c < 8; ++c) {
                colorNodes.push_back(u32(colorOffsets.size()));
                colorOffsets.push_back(leavesCount);
                if (mask & (1u << c)) leavesCount += node_count(level + 1, basic[basic[index + 1 + k++]]);
            }
        } else {
            for (u32 c = 0; c < 8; ++c) colorNodes.push_back(0);
            for (u32 c = 0; c < 8; ++c) {
                if (!(mask & (1u << c))) continue;
                const u32 child = basic[index + 1 + k++];
                const u32 ci = build_hash_color_tree(level + 1, child, leavesCount);
                colorNodes[colorIndex + c] = ci;
                leavesCount += node_count(level + 1, basic[child]);
            }
        }
        return colorIndex;
    }

    // ------------------------------------------------------------------ driver
    bool run()
    {
        const auto t0 = std::chrono::steady_clock::now();
        L = P.levels; leafLevel = L - 2;
        if (L < 6 || L > 24 || P.footprint_log2 > L || P.footprint_log2 < 5) return false;
        topLevels = std::min<u32>(10, L > 7 ? L - 7 : 0);
        T.init(L, P.footprint_log2, P.seed, P.roughness, P.finest_cell_log2 ? P.finest_cell_log2 : 2);
        u32 nThreads = P.n_threads ? P.n_threads : std::max(1u, std::thread::hardware_concurrency());

        // floating sphere shells above the terrain
        for (u32 i = 0; i < std::min<u32>(P.n_spheres, 64); ++i) {
            const u32 h = hash2(i32(i), 77, P.seed ^ 0x51E2E5u), g = hash2(i32(i), 99, P.seed ^ 0x0B5E55EDu);
            Sphere sp;
            const i64 margin = T.fsize / 8;
            sp.cx = T.fx0 + margin + i64(h % u32(T.fsize - 2 * margin));
            sp.cz = T.fz0 + margin + i64((h >> 8) % u32(T.fsize - 2 * margin));
            sp.r = std::max<i64>(6, (T.fsize >> 7) + i64(g % u32(std::max(1, T.fsize >> 6))));
            sp.t = 2;
            sp.cy = T.height(i32(sp.cx), i32(sp.cz)) + 2 * sp.r + i64((g >> 12) % u32(std::max<i64>(1, 4 * sp.r)));
            sp.cy = std::min<i64>(sp.cy, (i64(1) << L) - sp.r - 2);
            spheres.push_back(sp);
        }
        const u64 allSpheres = spheres.empty() ? 0 : (spheres.size() == 64 ? ~u64(0) : ((u64(1) << spheres.size()) - 1));

        for (u32 l = 0; l < leafLevel; ++l) tables.push_back(new NodeTable());

        // split level: node size 2^(F-4) (>= 16), at least level 1
        u32 splitSize = std::max<u32>(4, P.footprint_log2 >= 4 ? P.footprint_log2 - 4 : 4);
        u32 splitLevel = std::min(leafLevel, std::max<u32>(1, L - splitSize));
        Tile cache0[4];
        std::vector<Task> tasks; std::vector<u64> taskCand;
        enumerate_tasks(cache0, 0, splitLevel, 0, 0, 0, allSpheres, tasks, taskCand);
        std::vector<Ref> results(tasks.size());
        std::atomic<size_t> next{ 0 };
        auto worker = [&]() {
            Tile cache[4];
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= tasks.size()) break;
                results[i] = build(cache, splitLevel, tasks[i].x, tasks[i].y, tasks[i].z, taskCand[i]);
            }
        };
        {
            std::vector<std::thread> th;
            for (u32 i = 0; i < nThreads; ++i) th.emplace_back(worker);
            for (auto& t : th) t.join();
        }
        const bool verbose = getenv("HDS_VERBOSE") != nullptr;
        auto lap = [&](const char* what) {
            if (verbose) fprintf(stderr, "[scene] %-12s %.2fs\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        };
        lap("build");
        size_t cursor = 0;
        Ref root = assemble(cache0, 0, splitLevel, 0, 0, 0, allSpheres, results, cursor);
        if (root.id == NONE) { fprintf(stderr, "scene_builder: empty scene\n"); return false; }
        nVoxels = root.count;
        serialise_basic(root);
        lap("serialise");
        for (auto* t : tables) delete t;
        tables.clear();
        for (auto& sh : leaves.sh) { std::vector<u32>().swap(sh.slots); std::vector<u64>().swap(sh.vals); }

        if (P.build_hash) { build_hash(); lap("hash"); }
        if (P.build_colors) {
            build_colors(nThreads);
            if (leafLevel > 10) build_hash_color_tree(0, 0, 0);
            lap("colors");
        }
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }
};

extern "C" {

hds_scene* hds_build(const hds_params* p)
{
    auto* s = new hds_scene();
    s->P = *p;
    if (!s->run()) { delete s; return nullptr; }
    return s;
}
void hds_free(hds_scene* s) { delete s; }
void hds_get_info(const hds_scene* s, hds_info* o)
{
    memset(o, 0, sizeof(*o));
    o->levels = s->L; o->top_levels = s->topLevels; o->n_voxels = s->nVoxels;
    o->basic_words = s->basic.size(); o->enclosed_leaves = s->enclosed.size();
    o->hash_page_table_size = u32(s->pageTable.size()); o->hash_pool_top = s->pool.empty() ? 0 : s->poolTop;
    o->hash_first_node_index = s->firstNodeIndex; o->has_hash_colors = !s->colorNodes.empty();
    o->n_weight_words = s->weights.size(); o->n_blocks = s->blocks.size(); o->n_macro_words = s->macroBlocks.size();
    o->n_color_nodes = s->colorNodes.size(); o->n_color_offsets = s->colorOffsets.size();
    o->n_uncompressed = s->uncompressed.size(); o->build_seconds = s->seconds;
    memcpy(o->nodes_per_level, s->nodesPerLevel, sizeof(o->nodes_per_level));
}
const uint32_t* hds_basic_data(const hds_scene* s) { return s->basic.data(); }
const uint64_t* hds_enclosed_leaves(const hds_scene* s) { return s->enclosed.data(); }
const uint32_t* hds_hash_pool(const hds_scene* s) { return s->pool.data(); }
const uint32_t* hds_hash_page_table(const hds_scene* s) { return s->pageTable.data(); }
// fill count of every bucket, HashDagUtils::get_bucket_global_index order (hash_table.h:18-35)
const uint32_t* hds_hash_bucket_sizes(const hds_scene* s) { return s->bucketSizes.data(); }
uint64_t hds_hash_bucket_count(const hds_scene* s) { return s->bucketSizes.size(); }
const uint32_t* hds_color_weights(const hds_scene* s) { return s->weights.data(); }
const uint64_t* hds_color_blocks(const hds_scene* s) { return s->blocks.data(); }
const uint64_t* hds_color_macro_blocks(const hds_scene* s) { return s->macroBlocks.data(); }
const uint32_t* hds_color_uncompressed(const hds_scene* s) { return s->uncompressed.data(); }
const uint32_t* hds_hash_color_nodes(const hds_scene* s) { return s->colorNodes.data(); }
const uint64_t* hds_hash_color_offsets(const hds_scene* s) { return s->colorOffsets.data(); }
int32_t hds_terrain_height(const hds_scene* s, int32_t x, int32_t z) { return s->T.inside(x, z) ? s->T.height(x, z) : -1; }

}  // extern "C"
