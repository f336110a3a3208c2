// ANALYSIS (not a test): what would a second level of beams buy?  After the 8x4 tile beam of hdt_beam.cuh stops at the first
// node its 32 rays may disagree about, sub-beams (4x2, 2x2, ...) RESUME from that state with their own, tighter intervals and
// walk on until THEIR rays may disagree; the rays then resume from the sub-beam's state.  Counts interval-mask visits per level of
// the hierarchy and per-ray visits, and checks every pixel against the plain per-ray DFS.  Built on tests/cpp/beam_model.cpp.
#include "../cpp/beam_model.cpp"

namespace {
// the beam DFS of beam_model.cpp, started from a saved state instead of the root
void beam_dfs_from(const Dag& dag, const BeamRay& br, u8 order, const State& in, State& st, u64& beamVisits)
{
    const u32 levels = dag.levels(), leafLevel = dag.leaf_level();
    st = in;
    if (in.done) return;
    u32 level = in.level, px = in.px, py = in.py, pz = in.pz;
    StackEntry* stack = st.stack; StackEntry cache; u64 cachedLeaf = in.leaf;
    cache.index = in.index; cache.childMask = in.cm; cache.visitMask = 0;
    auto centre = [&](u32 lvl, u32 p) { const u32 sh = levels - lvl; return float(1u << (sh-1)) + float(p << sh); };
    auto save = [&](int done) { st.level = level; st.px = px; st.py = py; st.pz = pz; st.index = cache.index; st.cm = cache.childMask; st.leaf = cachedLeaf; st.done = done; };
    {
        int rs = 0; ++beamVisits;
        u32 m;
        if (level == 0) {
            m = interval_mask<true>(centre(0,0), centre(0,0), centre(0,0), float(1u << (levels-1)), br, &rs);
            if (rs == 1) { save(2); return; }
            if (rs == 2) { save(0); return; }
        } else m = interval_mask<false>(centre(level,px), centre(level,py), centre(level,pz), float(1u << (levels-level-1)), br, &rs);
        if (((m >> 8) ^ m) & 0xFF & cache.childMask) { save(0); return; }
        cache.visitMask = cache.childMask & (m & 0xFF);
    }
    for (;;) {
        u32 newLevel = level;
        while (newLevel > 0 && !cache.visitMask) { newLevel--; cache = stack[newLevel]; }
        if (newLevel == 0 && !cache.visitMask) { save(2); return; }
        px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
        level = newLevel;
        const u8 nextChild = next_child_ordered(order, cache.visitMask);
        cache.visitMask &= u8(~(1u << nextChild));
        px = (px << 1) | ((nextChild & 4u) >> 2); py = (py << 1) | ((nextChild & 2u) >> 1); pz = (pz << 1) | (nextChild & 1u);
        stack[level] = cache;
        level++;
        if (level == levels) { save(1); return; }
        if (level < leafLevel) {
            cache.index = dag.get_child_index(cache.index, cache.childMask, nextChild);
            cache.childMask = u8(dag.get_node(cache.index) & 0xFF);
        } else if (level == leafLevel) {
            const u32 addr = dag.get_child_index(cache.index, cache.childMask, nextChild);
            cachedLeaf = dag.get_leaf(addr);
            cache.childMask = first_child_mask(cachedLeaf);
        } else cache.childMask = second_child_mask(cachedLeaf, nextChild);
        ++beamVisits;
        int rs;
        u32 m = interval_mask<false>(centre(level,px), centre(level,py), centre(level,pz), float(1u << (levels-level-1)), br, &rs);
        if (((m >> 8) ^ m) & 0xFF & cache.childMask) { save(0); return; }
        cache.visitMask = cache.childMask & (m & 0xFF);
    }
}
}

// tiles[2*k], tiles[2*k+1] = width, height of hierarchy level k (level 0 = 8x4 tile, each next level must divide the previous).
// out: [0] ray visits, no beams; [1] ray visits after the whole hierarchy; [2+k] beam visits of level k; [10] mismatching pixels;
//      [11] sum over 8x4 tiles of max-over-rays visits after the hierarchy; [12] the same without beams
extern "C" int hier_paths(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                          const double ddy[3], const uint32_t* tiles, uint32_t nLevels, uint64_t* out)
{
    Counters c; Dag d(*dag, c);
    const u32 TW = tiles[0], TH = tiles[1];
    std::vector<RayS> rays(TW*TH);
    std::vector<State> rayState(TW*TH);
    std::vector<char> hasState(TW*TH);
    for (u32 ty = 0; ty < H; ty += TH) for (u32 tx = 0; tx < W; tx += TW) {
        for (u32 i = 0; i < TW*TH; ++i) {
            const u32 x = tx + i % TW, y = ty + i / TW;
            RayS& r = rays[i]; r.active = x < W && y < H; if (!r.active) continue;
            const D3 dd = primary_direction(cam, rmin, ddx, ddy, x, H-1-y);
            r.o = { float(cam[0]), float(cam[1]), float(cam[2]) };
            r.d = { float(dd.x), float(dd.y), float(dd.z) };
            r.inv = { 1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z };
            r.order = u8((r.d.x < 0 ? 4 : 0) + (r.d.y < 0 ? 2 : 0) + (r.d.z < 0 ? 1 : 0));
        }
        std::fill(hasState.begin(), hasState.end(), 0);
        // recursive refinement: process (x0,y0,w,h) sub-rectangles level by level
        struct Job { u32 x0, y0, lvl; State st; bool has; };
        std::vector<Job> jobs;
        { Job j; j.x0 = 0; j.y0 = 0; j.lvl = 0; j.has = false; jobs.push_back(j); }
        while (!jobs.empty()) {
            Job j = jobs.back(); jobs.pop_back();
            const u32 w = tiles[2*j.lvl], h = tiles[2*j.lvl+1];
            std::vector<RayS> sub;
            for (u32 yy = 0; yy < h; ++yy) for (u32 xx = 0; xx < w; ++xx) sub.push_back(rays[(j.y0+yy)*TW + j.x0+xx]);
            BeamRay br; u8 order = 0; State st; bool ok = make_beam(sub.data(), w*h, br, order);
            bool have = false;
            if (ok) {
                u64 bv = 0;
                if (j.has) { beam_dfs_from(d, br, order, j.st, st, bv); have = true; }
                else if (j.lvl == 0) { beam_dfs<true>(d, br, order, st, bv); have = true; }
                else { beam_dfs<true>(d, br, order, st, bv); have = true; }   // parent had no beam (mixed signs / untame): start at the root
                out[2 + j.lvl] += bv;
            } else if (j.has) { st = j.st; have = true; }   // cannot form a tighter beam: keep the parent's state
            if (j.lvl + 1 < nLevels && !(have && st.done)) {
                const u32 nw = tiles[2*(j.lvl+1)], nh = tiles[2*(j.lvl+1)+1];
                for (u32 yy = 0; yy < h; yy += nh) for (u32 xx = 0; xx < w; xx += nw) { Job n; n.x0 = j.x0+xx; n.y0 = j.y0+yy; n.lvl = j.lvl+1; n.st = st; n.has = have; jobs.push_back(n); }
            } else {
                for (u32 yy = 0; yy < h; ++yy) for (u32 xx = 0; xx < w; ++xx) { const u32 i = (j.y0+yy)*TW + j.x0+xx; hasState[i] = have; if (have) rayState[i] = st; }
            }
        }
        u64 mxA = 0, mxB = 0;
        for (u32 i = 0; i < TW*TH; ++i) {
            if (!rays[i].active) continue;
            u32 a0,a1,a2,b0,b1,b2; u64 va = 0, vb = 0;
            ray_dfs<true>(d, rays[i], nullptr, a0,a1,a2, va);
            ray_dfs<true>(d, rays[i], hasState[i] ? &rayState[i] : nullptr, b0,b1,b2, vb);
            out[0] += va; out[1] += vb; mxA = std::max(mxA, va); mxB = std::max(mxB, vb);
            if (a0!=b0||a1!=b1||a2!=b2) out[10]++;
        }
        out[11] += mxB; out[12] += mxA;
    }
    return 0;
}
