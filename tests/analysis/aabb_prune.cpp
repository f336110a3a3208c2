// ANALYSIS TOOL (test infrastructure, not shipped): how many node visits of the reference's ordered DFS would a
// conservative content-box test save?  On top of the CPU oracle.  Build + run: python tests/analysis/aabb_prune.py [footprint_log2]
//
// Per node (level, index) the exact box of its voxels in node-local coordinates is computed bottom-up (memoised: the DAG
// shares nodes).  The DFS of tracer.cu:166-249 is then replayed per pixel with a pruning rule and the visits counted:
//   mode 0  no pruning (the reference)
//   mode 1  a node's own box is tested on arrival (the visit is paid, the subtree below is saved)
//   mode 2  the boxes of the candidate children are tested at the parent (free pruning: an upper bound of what a
//           per-child record next to the child pointers could save)
// quant: boxes rounded outwards to 1/16 of the node size (what fits 24 bits: 6 x 4) or exact (0).  margin: voxels the box is
// inflated by before the slab test.  Every variant must reproduce the reference's hit voxel (mismatches are counted).
#include "../../oracle/hdo_oracle.cpp"
#include <unordered_map>

namespace {
struct Box { u32 lo[3], hi[3]; };   // voxel units, node-local, hi exclusive
struct BoxCache {
    const Dag& d; u32 levels, leafLevel;
    std::vector<std::unordered_map<u32, Box>> memo;
    BoxCache(const Dag& dd) : d(dd), levels(dd.levels()), leafLevel(dd.leaf_level()), memo(dd.levels()) {}
    static Box leaf_box(u64 leaf)
    {
        Box b{{4, 4, 4}, {0, 0, 0}};
        for (u32 c1 = 0; c1 < 8; ++c1) for (u32 c2 = 0; c2 < 8; ++c2) if ((leaf >> (c1 * 8 + c2)) & 1) {
            const u32 p[3] = { ((c1 >> 2) & 1) * 2 + ((c2 >> 2) & 1), ((c1 >> 1) & 1) * 2 + ((c2 >> 1) & 1), (c1 & 1) * 2 + (c2 & 1) };
            for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(b.lo[k], p[k]); b.hi[k] = std::max(b.hi[k], p[k] + 1); }
        }
        return b;
    }
    Box get(u32 level, u32 index)   // node at `level` (0 = root); level == leafLevel: index addresses a 64-bit leaf
    {
        auto it = memo[level].find(index);
        if (it != memo[level].end()) return it->second;
        Box b;
        if (level == leafLevel) b = leaf_box(d.get_leaf(index));
        else {
            const u32 half = 1u << (levels - level - 1);
            const u8 cm = u8(d.get_node(index) & 0xFF);
            b = Box{{~0u, ~0u, ~0u}, {0, 0, 0}};
            for (u8 c = 0; c < 8; ++c) if (cm & (1u << c)) {
                const Box cb = get(level + 1, d.get_child_index(index, cm, c));
                const u32 off[3] = { (c & 4u) ? half : 0, (c & 2u) ? half : 0, (c & 1u) ? half : 0 };
                for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(b.lo[k], cb.lo[k] + off[k]); b.hi[k] = std::max(b.hi[k], cb.hi[k] + off[k]); }
            }
        }
        memo[level][index] = b;
        return b;
    }
};

inline bool ray_hits_box(const F3& o, const F3& inv, const float lo[3], const float hi[3])
{
    const float oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
    float tn = 0.0f, tf = 3.0e38f;
    for (int k = 0; k < 3; ++k) {
        const float t1 = (lo[k] - oo[k]) * ii[k], t2 = (hi[k] - oo[k]) * ii[k];
        tn = std::max(tn, std::min(t1, t2));
        tf = std::min(tf, std::max(t1, t2));
    }
    return tn <= tf;
}

struct Prune {
    int mode, quant; float margin;
    // world box of node (level, path) with content box b
    bool passes(const Box& b, u32 level, u32 levels, u32 px, u32 py, u32 pz, const F3& o, const F3& inv) const
    {
        const u32 shift = levels - level;
        const u32 p[3] = { px, py, pz };
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            u32 l = b.lo[k], h = b.hi[k];
            if (quant && shift > 4) { const u32 q = 1u << (shift - 4); l = l / q * q; h = (h + q - 1) / q * q; }
            lo[k] = float((p[k] << shift) + l) - margin;
            hi[k] = float((p[k] << shift) + h) + margin;
        }
        return ray_hits_box(o, inv, lo, hi);
    }
};
}  // namespace

// out: [0] visits, [1] hits, [2] mismatches vs mode 0 (filled by the caller comparing results), result voxels in `res` (3 u32 per pixel)
extern "C" int ana_prune(const hdo_dag* dag, uint32_t W, uint32_t H, uint32_t stride, const double cam[3], const double rmin[3], const double ddx[3],
                         const double ddy[3], int mode, int quant, float margin, int shadow, const uint32_t* paths /* shadow: hit voxels per pixel */,
                         uint64_t* out, uint32_t* res)
{
    Counters c;
    Dag d(*dag, c);
    static BoxCache* cache = nullptr;
    static const void* cacheFor = nullptr;
    if (cacheFor != dag->data) { delete cache; cache = new BoxCache(d); cacheFor = dag->data; }
    // the cache holds a Dag reference with a dangling Counters otherwise: rebuild the accessor each call
    BoxCache bc(d);
    bc.memo.swap(cache->memo);
    const Prune pr{ mode, quant, margin };
    const u32 levels = d.levels(), leafLevel = d.leaf_level();
    const float sl = std::sqrt(0.3f * 0.3f + 1.0f + 0.5f * 0.5f);
    const F3 sun = { 0.3f / sl, 1.0f / sl, 0.5f / sl };
    u64 visits = 0, hits = 0;
    for (u32 y = 0; y < H; y += stride) for (u32 x = 0; x < W; x += stride) {
        F3 o, dir;
        if (!shadow) {
            const D3 dd = primary_direction(cam, rmin, ddx, ddy, x, H - 1 - y);
            o = { float(cam[0]), float(cam[1]), float(cam[2]) };
            dir = { float(dd.x), float(dd.y), float(dd.z) };
        } else {
            const u32* p = paths + (size_t(y) * W + x) * 4;
            if (!(p[0] | p[1] | p[2])) { res[(size_t(y) * W + x) * 3] = 2; continue; }
            // approximate shadow origin (voxel centre top + bias): visit statistics only, not a parity check of the exact f64 hit point
            o = { float(p[0]) + 0.5f + sun.x, float(p[1]) + 1.0f + sun.y, float(p[2]) + 0.5f + sun.z };
            dir = sun;
        }
        const F3 inv = { 1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z };
        const u8 order = u8((dir.x < 0 ? 4 : 0) + (dir.y < 0 ? 2 : 0) + (dir.z < 0 ? 1 : 0));
        u32 level = 0, px = 0, py = 0, pz = 0;
        StackEntry stack[32]; StackEntry cur; u64 cachedLeaf = 0;
        u32 leafIndex = 0;
        cur.index = d.first();
        cur.childMask = u8(d.get_node(cur.index) & 0xFF);
        cur.visitMask = cur.childMask & intersection_mask<true>(0, levels, px, py, pz, o, dir, inv);
        bool hit = false;
        auto prune_children = [&](u32 lvl, const StackEntry& e, u32 ppx, u32 ppy, u32 ppz) -> u8 {
            // mode 2: drop candidate children of the node `e` (at level lvl) whose content box the ray misses
            u8 vm = e.visitMask;
            if (pr.mode != 2 || lvl >= leafLevel) return vm;
            for (u8 ch = 0; ch < 8; ++ch) if (vm & (1u << ch)) {
                const u32 ci = d.get_child_index(e.index, e.childMask, ch);
                const Box cb = bc.get(lvl + 1, ci);
                if (!pr.passes(cb, lvl + 1, levels, (ppx << 1) | ((ch >> 2) & 1), (ppy << 1) | ((ch >> 1) & 1), (ppz << 1) | (ch & 1), o, inv)) vm &= u8(~(1u << ch));
            }
            return vm;
        };
        cur.visitMask = prune_children(0, cur, 0, 0, 0);
        for (;;) {
            u32 newLevel = level;
            while (newLevel > 0 && !cur.visitMask) { newLevel--; cur = stack[newLevel]; }
            if (newLevel == 0 && !cur.visitMask) break;
            px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
            level = newLevel;
            u8 nextChild;
            if (!shadow) nextChild = next_child_ordered(order, cur.visitMask);
            else nextChild = u8(31 - __builtin_clz(u32(cur.visitMask)));
            cur.visitMask &= u8(~(1u << nextChild));
            px = (px << 1) | ((nextChild & 4u) >> 2); py = (py << 1) | ((nextChild & 2u) >> 1); pz = (pz << 1) | (nextChild & 1u);
            stack[level] = cur;
            level++; ++visits;
            if (level == levels) { hit = true; break; }
            if (level < leafLevel) {
                cur.index = d.get_child_index(cur.index, cur.childMask, nextChild);
                cur.childMask = u8(d.get_node(cur.index) & 0xFF);
            } else if (level == leafLevel) {
                leafIndex = d.get_child_index(cur.index, cur.childMask, nextChild);
                cachedLeaf = d.get_leaf(leafIndex);
                cur.childMask = first_child_mask(cachedLeaf);
            } else cur.childMask = second_child_mask(cachedLeaf, nextChild);
            cur.visitMask = cur.childMask & intersection_mask<false>(level, levels, px, py, pz, o, dir, inv);
            if (pr.mode == 1 && level <= leafLevel && cur.visitMask) {
                const Box b = bc.get(level, level == leafLevel ? leafIndex : cur.index);
                if (!pr.passes(b, level, levels, px, py, pz, o, inv)) cur.visitMask = 0;
            }
            if (pr.mode == 2) cur.visitMask = prune_children(level, cur, px, py, pz);
        }
        u32* r = res + (size_t(y) * W + x) * 3;
        if (hit) { r[0] = px; r[1] = py; r[2] = pz; ++hits; } else { r[0] = r[1] = r[2] = 0; }
        if (shadow) { r[0] = hit ? 1 : 0; r[1] = r[2] = 0; }
    }
    cache->memo.swap(bc.memo);
    out[0] = visits; out[1] = hits;
    return 0;
}
