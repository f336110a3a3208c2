// ANALYSIS TOOL (test infrastructure, not shipped): node visits of the reference DFS per level, per ray and
// per tile, on top of the CPU oracle.  Build + run: python tests/analysis/visit_histogram.py [footprint_log2]
#include "../../oracle/hdo_oracle.cpp"
extern "C" int ana_paths(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                    const double ddy[3], uint64_t* visitsByLevel /*32*/, uint64_t* onPathByLevel, uint64_t* raysBySteps /*1024*/, uint64_t* childCount /*9*/, uint64_t* vmCount)
{
    Counters c;
    Dag d(*dag, c);
    const u32 levels = d.levels(), leafLevel = d.leaf_level();
    for (u32 y = 0; y < H; ++y) for (u32 x = 0; x < W; ++x) {
        const D3 dd = primary_direction(cam, rmin, ddx, ddy, x, H-1-y);  // orientation irrelevant
        F3 o = { float(cam[0]), float(cam[1]), float(cam[2]) };
        F3 dir = { float(dd.x), float(dd.y), float(dd.z) };
        F3 inv = { 1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z };
        u8 order = u8((dir.x < 0 ? 4 : 0) + (dir.y < 0 ? 2 : 0) + (dir.z < 0 ? 1 : 0));
        u32 level = 0, px = 0, py = 0, pz = 0;
        StackEntry stack[32]; StackEntry cache; u64 cachedLeaf = 0;
        u64 vis[32] = {0};
        cache.index = d.first();
        cache.childMask = u8(d.get_node(cache.index) & 0xFF);
        cache.visitMask = cache.childMask & intersection_mask<true>(0, levels, px, py, pz, o, dir, inv);
        u32 steps = 0; bool hit = false;
        for (;;) {
            u32 newLevel = level;
            while (newLevel > 0 && !cache.visitMask) { newLevel--; cache = stack[newLevel]; }
            if (newLevel == 0 && !cache.visitMask) break;
            px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
            level = newLevel;
            const u8 nextChild = next_child_ordered(order, cache.visitMask);
            cache.visitMask &= u8(~(1u << nextChild));
            px = (px << 1) | ((nextChild & 4u) >> 2); py = (py << 1) | ((nextChild & 2u) >> 1); pz = (pz << 1) | (nextChild & 1u);
            stack[level] = cache;
            level++; ++steps; vis[level]++;
            if (level == levels) { hit = true; break; }
            if (level < leafLevel) {
                cache.index = d.get_child_index(cache.index, cache.childMask, nextChild);
                cache.childMask = u8(d.get_node(cache.index) & 0xFF);
            } else if (level == leafLevel) {
                const u32 addr = d.get_child_index(cache.index, cache.childMask, nextChild);
                cachedLeaf = d.get_leaf(addr);
                cache.childMask = first_child_mask(cachedLeaf);
            } else cache.childMask = second_child_mask(cachedLeaf, nextChild);
            u8 m = intersection_mask<false>(level, levels, px, py, pz, o, dir, inv);
            cache.visitMask = cache.childMask & m;
            childCount[__builtin_popcount(cache.childMask)]++;
            vmCount[__builtin_popcount(cache.visitMask)]++;
        }
        for (u32 l = 0; l < 32; ++l) { visitsByLevel[l] += vis[l]; if (hit && vis[l]) onPathByLevel[l] += 1; }
        raysBySteps[steps < 1023 ? steps : 1023]++;
    }
    return 0;
}

// Beam-coherence upper bound: per TWxTH tile, length of the common prefix of the rays' DFS visit sequences
// (same node, same vm).  out[0] = total visits, out[1] = visits inside common prefix (summed over rays), out[2] = tiles, out[3]=sum prefix len
extern "C" int ana_beam(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                    const double ddy[3], uint32_t TW, uint32_t TH, uint64_t* out, uint64_t* prefixLevelHist)
{
    Counters c;
    Dag d(*dag, c);
    const u32 levels = d.levels(), leafLevel = d.leaf_level();
    struct Vis { u32 level, px, py, pz; u8 vm; };
    std::vector<std::vector<Vis>> seqs(TW * TH);
    for (u32 ty = 0; ty + TH <= H; ty += TH) for (u32 tx = 0; tx + TW <= W; tx += TW) {
        for (u32 i = 0; i < TW * TH; ++i) {
            const u32 x = tx + i % TW, y = ty + i / TW;
            auto& seq = seqs[i]; seq.clear();
            const D3 dd = primary_direction(cam, rmin, ddx, ddy, x, H-1-y);
            F3 o = { float(cam[0]), float(cam[1]), float(cam[2]) };
            F3 dir = { float(dd.x), float(dd.y), float(dd.z) };
            F3 inv = { 1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z };
            u8 order = u8((dir.x < 0 ? 4 : 0) + (dir.y < 0 ? 2 : 0) + (dir.z < 0 ? 1 : 0));
            u32 level = 0, px = 0, py = 0, pz = 0;
            StackEntry stack[32]; StackEntry cache; u64 cachedLeaf = 0;
            cache.index = d.first();
            cache.childMask = u8(d.get_node(cache.index) & 0xFF);
            cache.visitMask = cache.childMask & intersection_mask<true>(0, levels, px, py, pz, o, dir, inv);
            seq.push_back({0,0,0,0,u8(cache.visitMask ^ (order<<0)*0)});
            for (;;) {
                u32 newLevel = level;
                while (newLevel > 0 && !cache.visitMask) { newLevel--; cache = stack[newLevel]; }
                if (newLevel == 0 && !cache.visitMask) break;
                px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
                level = newLevel;
                const u8 nextChild = next_child_ordered(order, cache.visitMask);
                cache.visitMask &= u8(~(1u << nextChild));
                px = (px << 1) | ((nextChild & 4u) >> 2); py = (py << 1) | ((nextChild & 2u) >> 1); pz = (pz << 1) | (nextChild & 1u);
                stack[level] = cache;
                level++;
                if (level == levels) { seq.push_back({level,px,py,pz,0}); break; }
                if (level < leafLevel) {
                    cache.index = d.get_child_index(cache.index, cache.childMask, nextChild);
                    cache.childMask = u8(d.get_node(cache.index) & 0xFF);
                } else if (level == leafLevel) {
                    const u32 addr = d.get_child_index(cache.index, cache.childMask, nextChild);
                    cachedLeaf = d.get_leaf(addr);
                    cache.childMask = first_child_mask(cachedLeaf);
                } else cache.childMask = second_child_mask(cachedLeaf, nextChild);
                cache.visitMask = cache.childMask & intersection_mask<false>(level, levels, px, py, pz, o, dir, inv);
                seq.push_back({level,px,py,pz,cache.visitMask});
            }
        }
        size_t pre = 0;
        for (;; ++pre) {
            bool same = true;
            for (u32 i = 0; i < TW*TH && same; ++i) {
                if (pre >= seqs[i].size()) { same = false; break; }
                const Vis& a = seqs[0][pre]; const Vis& b = seqs[i][pre];
                same = a.level==b.level && a.px==b.px && a.py==b.py && a.pz==b.pz && a.vm==b.vm;
            }
            if (!same) break;
        }
        u32 lvl = pre ? seqs[0][pre-1].level : 0;
        prefixLevelHist[lvl]++;
        out[2]++; out[3] += pre;
        for (u32 i = 0; i < TW*TH; ++i) { out[0] += seqs[i].size(); out[1] += pre; }
    }
    return 0;
}
