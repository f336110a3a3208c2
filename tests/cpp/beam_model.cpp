// TEST INFRASTRUCTURE: a host model of the beam pre-pass of hashdag_b200/csrc/hdt_beam.cuh, built on the CPU
// oracle (the only place its internals are reused: this file is compiled by tests/test_beam_model_cpu.py).
//
// For every 8x4 (or TWxTH) tile it encloses the tile's rays in intervals, walks the reference DFS with the
// interval form of compute_intersection_mask (tracer.cu:19-136) until the rays may disagree, then lets each
// ray resume from that state -- and checks pixel by pixel that resuming gives what the plain per-ray DFS of
// the oracle gives.  It also counts node visits with and without the pre-pass (the numbers in DESIGN.md §6).
// The arithmetic argument (monotone correctly-rounded operations => exact enclosures) is in hdt_beam.cuh.
#include "../../oracle/hdo_oracle.cpp"
#include <cassert>
#include <cstdio>
#include <algorithm>
namespace {
struct Iv { float lo, hi; };
inline Iv imul(Iv a, Iv b) { // interval product with the rounded multiply (monotone => exact enclosure)
    const float p0 = a.lo*b.lo, p1 = a.lo*b.hi, p2 = a.hi*b.lo, p3 = a.hi*b.hi;
    return { std::min(std::min(p0,p1), std::min(p2,p3)), std::max(std::max(p0,p1), std::max(p2,p3)) };
}
struct BeamRay { Iv o[3], d[3], inv[3], ainv[3]; };
// returns def mask in low byte, poss mask in second byte; rootMiss: 0 = all pass, 1 = all miss, 2 = ambiguous
template <bool isRoot>
inline u32 interval_mask(float cx, float cy, float cz, float radius, const BeamRay& br, int* rootState)
{
    const float c[3] = {cx,cy,cz};
    Iv r[3], t[3], a[3], b[3];
    for (int k=0;k<3;++k) {
        r[k] = { c[k] - br.o[k].hi, c[k] - br.o[k].lo };
        t[k] = imul(r[k], br.inv[k]);
        a[k] = { fmaf(-radius, br.ainv[k].hi, t[k].lo), fmaf(-radius, br.ainv[k].lo, t[k].hi) };
        b[k] = { fmaf(radius, br.ainv[k].lo, t[k].lo), fmaf(radius, br.ainv[k].hi, t[k].hi) };
    }
    Iv tmin = { fmaxf(fmaxf(a[0].lo,a[1].lo), fmaxf(a[2].lo,0.f)), fmaxf(fmaxf(a[0].hi,a[1].hi), fmaxf(a[2].hi,0.f)) };
    Iv tmax = { fminf(fminf(b[0].lo,b[1].lo), b[2].lo), fminf(fminf(b[0].hi,b[1].hi), b[2].hi) };
    if (isRoot) {
        if (tmin.lo >= tmax.hi) { *rootState = 1; return 0; }
        if (tmin.hi < tmax.lo) *rootState = 0; else { *rootState = 2; return 0xFF00; }
    }
    Iv h = { 0.5f*(tmin.lo+tmax.lo), 0.5f*(tmin.hi+tmax.hi) };
    u32 def = 0, poss = 0;
    {
        u32 bitsDef = 0, bitsAmb = 0;
        const u32 w[3] = {4,2,1};
        for (int k=0;k<3;++k) {
            Iv q = imul(h, br.d[k]);
            if (q.lo >= r[k].hi) bitsDef |= w[k];
            else if (q.hi < r[k].lo) {}
            else bitsAmb |= w[k];
        }
        if (!bitsAmb) { def |= 1u<<bitsDef; poss |= 1u<<bitsDef; }
        else { for (u32 s = 0; s < 8; ++s) if ((s & ~bitsAmb) == bitsDef) poss |= 1u<<s; }
    }
    const float eps = 1e-4f;
    Iv rm[3], rp[3];
    for (int k=0;k<3;++k) { rm[k] = { r[k].lo - eps, r[k].hi - eps }; rp[k] = { r[k].lo + eps, r[k].hi + eps }; }
    const u32 HI[3] = {0xF0,0xCC,0xAA}, LO[3] = {0x0F,0x33,0x55};
    for (int p=0;p<3;++p) {
        const bool pinDef = tmin.hi <= t[p].lo && t[p].hi <= tmax.lo;
        const bool pinPoss = tmin.lo <= t[p].hi && t[p].lo <= tmax.hi;
        if (!pinPoss) continue;
        u32 mdef = 0xFF, mposs = 0xFF;
        for (int k=0;k<3;++k) if (k != p) {
            Iv q = imul(t[p], br.d[k]);
            u32 Ad = 0, Ap = 0;
            if (q.lo >= rm[k].hi) Ad |= HI[k];
            if (q.hi >= rm[k].lo) Ap |= HI[k];
            if (q.hi <= rp[k].lo) Ad |= LO[k];
            if (q.lo <= rp[k].hi) Ap |= LO[k];
            mdef &= Ad; mposs &= Ap;
        }
        if (pinDef) def |= mdef;
        poss |= mposs;
    }
    return def | (poss << 8);
}

struct RayS { F3 o, d, inv; u8 order; bool active; };
struct State { u32 level, px, py, pz, index; u8 cm, vm; u64 leaf; u32 pending; StackEntry stack[32]; int done; /*0 = continue at node (vm to compute), 1 = hit, 2 = miss*/ };

// per-ray DFS, optionally starting from a beam state; counts visits (mask evaluations)
template <bool ordered>
bool ray_dfs(const Dag& dag, const RayS& ry, const State* st, u32& ox, u32& oy, u32& oz, u64& visits)
{
    const u32 levels = dag.levels(), leafLevel = dag.leaf_level();
    u32 level = 0, px = 0, py = 0, pz = 0;
    StackEntry stack[32]; StackEntry cache; u64 cachedLeaf = 0;
    if (st) {
        if (st->done == 1) { ox = st->px; oy = st->py; oz = st->pz; return true; }
        if (st->done == 2) { ox = oy = oz = 0; return false; }
        level = st->level; px = st->px; py = st->py; pz = st->pz; cachedLeaf = st->leaf;
        for (u32 l = 0; l < level; ++l) { stack[l] = st->stack[l]; }
        cache.index = st->index; cache.childMask = st->cm;
        ++visits;
        if (level == 0) cache.visitMask = cache.childMask & intersection_mask<true>(0, levels, px, py, pz, ry.o, ry.d, ry.inv);
        else cache.visitMask = cache.childMask & intersection_mask<false>(level, levels, px, py, pz, ry.o, ry.d, ry.inv);
    } else {
        cache.index = dag.first();
        cache.childMask = u8(dag.get_node(cache.index) & 0xFF);
        ++visits;
        cache.visitMask = cache.childMask & intersection_mask<true>(0, levels, px, py, pz, ry.o, ry.d, ry.inv);
    }
    for (;;) {
        u32 newLevel = level;
        while (newLevel > 0 && !cache.visitMask) { newLevel--; cache = stack[newLevel]; }
        if (newLevel == 0 && !cache.visitMask) { ox = oy = oz = 0; return false; }
        px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
        level = newLevel;
        const u8 nextChild = ordered ? next_child_ordered(ry.order, cache.visitMask) : u8(31 - __builtin_clz(u32(cache.visitMask)));
        cache.visitMask &= u8(~(1u << nextChild));
        px = (px << 1) | ((nextChild & 4u) >> 2); py = (py << 1) | ((nextChild & 2u) >> 1); pz = (pz << 1) | (nextChild & 1u);
        stack[level] = cache;
        level++;
        if (level == levels) { ox = px; oy = py; oz = pz; return true; }
        if (level < leafLevel) {
            cache.index = dag.get_child_index(cache.index, cache.childMask, nextChild);
            cache.childMask = u8(dag.get_node(cache.index) & 0xFF);
        } else if (level == leafLevel) {
            const u32 addr = dag.get_child_index(cache.index, cache.childMask, nextChild);
            cachedLeaf = dag.get_leaf(addr);
            cache.childMask = first_child_mask(cachedLeaf);
        } else cache.childMask = second_child_mask(cachedLeaf, nextChild);
        ++visits;
        cache.visitMask = cache.childMask & intersection_mask<false>(level, levels, px, py, pz, ry.o, ry.d, ry.inv);
    }
}

// beam DFS: same loop, interval masks; stops at first ambiguous node
template <bool ordered>
void beam_dfs(const Dag& dag, const BeamRay& br, u8 order, State& st, u64& beamVisits)
{
    const u32 levels = dag.levels(), leafLevel = dag.leaf_level();
    u32 level = 0, px = 0, py = 0, pz = 0;
    StackEntry* stack = st.stack; StackEntry cache; u64 cachedLeaf = 0;
    auto centre = [&](u32 lvl, u32 p) { const u32 sh = levels - lvl; return float(1u << (sh-1)) + float(p << sh); };
    auto save = [&](int done) { st.level = level; st.px = px; st.py = py; st.pz = pz; st.index = cache.index; st.cm = cache.childMask; st.leaf = cachedLeaf; st.done = done; };
    cache.index = dag.first();
    cache.childMask = u8(dag.get_node(cache.index) & 0xFF);
    {
        int rs = 0; ++beamVisits;
        u32 m = interval_mask<true>(centre(0,0), centre(0,0), centre(0,0), float(1u << (levels-1)), br, &rs);
        if (rs == 1) { save(2); return; }
        if (rs == 2 || (((m >> 8) ^ m) & 0xFF & cache.childMask)) { save(0); return; }
        cache.visitMask = cache.childMask & (m & 0xFF);
    }
    for (;;) {
        u32 newLevel = level;
        while (newLevel > 0 && !cache.visitMask) { newLevel--; cache = stack[newLevel]; }
        if (newLevel == 0 && !cache.visitMask) { save(2); return; }
        px >>= (level - newLevel); py >>= (level - newLevel); pz >>= (level - newLevel);
        level = newLevel;
        const u8 nextChild = ordered ? next_child_ordered(order, cache.visitMask) : u8(31 - __builtin_clz(u32(cache.visitMask)));
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

inline bool tame(const RayS& r) {
    const float lim = 1.2676506e30f;
    return std::fabs(r.inv.x) <= lim && std::fabs(r.inv.y) <= lim && std::fabs(r.inv.z) <= lim && std::fabs(r.o.x) <= 3e38f && std::fabs(r.o.y) <= 3e38f && std::fabs(r.o.z) <= 3e38f;
}
bool make_beam(const RayS* rays, u32 n, BeamRay& br, u8& order)
{
    bool first = true;
    for (u32 i = 0; i < n; ++i) {
        const RayS& r = rays[i];
        if (!r.active) continue;
        if (!tame(r)) return false;
        const float o[3] = {r.o.x,r.o.y,r.o.z}, d[3] = {r.d.x,r.d.y,r.d.z}, iv[3] = {r.inv.x,r.inv.y,r.inv.z};
        if (first) { order = r.order; for (int k=0;k<3;++k) { br.o[k] = {o[k],o[k]}; br.d[k] = {d[k],d[k]}; br.inv[k] = {iv[k],iv[k]}; } first = false; continue; }
        if (r.order != order) return false;
        for (int k=0;k<3;++k) {
            br.o[k].lo = std::min(br.o[k].lo,o[k]); br.o[k].hi = std::max(br.o[k].hi,o[k]);
            br.d[k].lo = std::min(br.d[k].lo,d[k]); br.d[k].hi = std::max(br.d[k].hi,d[k]);
            br.inv[k].lo = std::min(br.inv[k].lo,iv[k]); br.inv[k].hi = std::max(br.inv[k].hi,iv[k]);
        }
    }
    if (first) return false;
    for (int k=0;k<3;++k) {
        // same sign within the beam (order equal => d signs equal, but d == 0 gives inv = inf -> not tame)
        if ((br.inv[k].lo < 0) != (br.inv[k].hi < 0)) return false;
        const float a = std::fabs(br.inv[k].lo), b = std::fabs(br.inv[k].hi);
        br.ainv[k] = { std::min(a,b), std::max(a,b) };
    }
    return true;
}
}

// out: [0] ray visits without beam, [1] ray visits with beam, [2] beam visits, [3] mismatches, [4] beams built, [5] tiles, [6] beams fully resolved
extern "C" int beam_paths(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                    const double ddy[3], uint32_t TW, uint32_t TH, uint32_t* paths, uint64_t* out)
{
    Counters c; Dag d(*dag, c);
    std::vector<RayS> rays(TW*TH);
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
        BeamRay br; u8 order = 0; State st; out[5]++;
        const bool ok = make_beam(rays.data(), TW*TH, br, order);
        if (ok) { out[4]++; u64 bv = 0; beam_dfs<true>(d, br, order, st, bv); out[2] += bv; out[16 + std::min<u64>(bv, 127)]++; if (st.done) out[6]++; }
        u64 mxA = 0, mxB = 0;
        for (u32 i = 0; i < TW*TH; ++i) {
            if (!rays[i].active) continue;
            const u32 x = tx + i % TW, y = ty + i / TW;
            u32 a0,a1,a2,b0,b1,b2;
            u64 va = 0, vb = 0;
            ray_dfs<true>(d, rays[i], nullptr, a0,a1,a2, va);
            ray_dfs<true>(d, rays[i], ok ? &st : nullptr, b0,b1,b2, vb);
            out[0] += va; out[1] += vb; mxA = std::max(mxA, va); mxB = std::max(mxB, vb);
            if (i + 1 == TW*TH) {}
            if (a0!=b0||a1!=b1||a2!=b2) out[3]++;
            u32* p = paths + (u64(y)*W + x)*4; p[0]=a0;p[1]=a1;p[2]=a2;p[3]=0;
        }
        out[7] += mxB; out[8] += mxA;
    }
    return 0;
}

extern "C" int beam_shadows(const hdo_dag* dag, uint32_t W, uint32_t H, const double cam[3], const double rmin[3], const double ddx[3],
                    const double ddy[3], uint32_t TW, uint32_t TH, const uint32_t* paths, float shadowBias, uint64_t* out)
{
    Counters c; Dag d(*dag, c);
    const F3 sun = sun_direction();
    const F3 sunInv = { 1.0f / sun.x, 1.0f / sun.y, 1.0f / sun.z };
    std::vector<RayS> rays(TW*TH);
    for (u32 ty = 0; ty < H; ty += TH) for (u32 tx = 0; tx < W; tx += TW) {
        for (u32 i = 0; i < TW*TH; ++i) {
            const u32 x = tx + i % TW, y = ty + i / TW;
            RayS& r = rays[i]; r.active = x < W && y < H; if (!r.active) continue;
            const u32* p = paths + (u64(y) * W + x) * 4;
            if (p[0] == 0 && p[1] == 0 && p[2] == 0) { r.active = false; continue; }
            const F3 ro = { float(p[0]), float(p[1]), float(p[2]) };
            const D3 dir = primary_direction(cam, rmin, ddx, ddy, x, H - 1 - y);
            const double bo[3] = { double(ro.x), double(ro.y), double(ro.z) }, dv[3] = { dir.x, dir.y, dir.z };
            double rm[3];
            for (int k = 0; k < 3; ++k) {
                const double t0 = (bo[k] - cam[k]) / dv[k], t1 = ((bo[k] + 1.0) - cam[k]) / dv[k];
                rm[k] = (t0 < t1) ? t0 : t1;
            }
            const double maxmin = fmax(fmax(rm[0], rm[1]), rm[2]);
            const F3 start = { float(std::fma(dv[0], maxmin, cam[0])), float(std::fma(dv[1], maxmin, cam[1])), float(std::fma(dv[2], maxmin, cam[2])) };
            r.o = { fmaf(shadowBias, sun.x, start.x), fmaf(shadowBias, sun.y, start.y), fmaf(shadowBias, sun.z, start.z) };
            r.d = sun; r.inv = sunInv; r.order = 0;
        }
        BeamRay br; u8 order = 0; State st; out[5]++;
        const bool ok = make_beam(rays.data(), TW*TH, br, order);
        if (ok) { out[4]++; beam_dfs<false>(d, br, order, st, out[2]); if (st.done) out[6]++; }
        for (u32 i = 0; i < TW*TH; ++i) {
            if (!rays[i].active) continue;
            u32 a0,a1,a2,b0,b1,b2;
            const bool ha = ray_dfs<false>(d, rays[i], nullptr, a0,a1,a2, out[0]);
            const bool hb = ray_dfs<false>(d, rays[i], ok ? &st : nullptr, b0,b1,b2, out[1]);
            if (ha != hb) out[3]++;
        }
    }
    return 0;
}
