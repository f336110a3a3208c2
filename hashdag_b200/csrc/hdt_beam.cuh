// Beam pre-pass: one thread walks the DFS once for a whole 8x4-pixel tile (= one warp of the
// per-ray kernels) for as long as every ray of the tile provably does exactly the same thing.
//
// Why: the per-ray kernels are bound by instruction issue (profiles/), and about half of all node
// visits of a frame are visits every ray of a warp makes identically (top levels, empty space in
// front of the surface).  A warp pays a full issue slot per instruction no matter how many of its
// lanes agree, so those visits are only saved by doing them ONCE PER TILE in a separate, small
// kernel (one thread per tile) and letting the 32 rays resume from the state it leaves behind.
//
// How it stays bit-exact: the reference's intersection mask (tracer.cu:19-136) is evaluated in
// INTERVAL form over the tile's rays.  Every operation of the mask is a single correctly-rounded
// IEEE operation, and correctly-rounded +, -, *, fma, min, max are monotone in each argument, so
// evaluating the same operations on the end points of the rays' (origin, direction, 1/direction)
// ranges encloses the value every individual ray computes -- no error margins involved.  Each of
// the mask's comparisons is then "true for all rays", "false for all rays" or undecided; the mask
// logic is monotone in the comparison results, which gives two masks  def <= mask(ray) <= poss  for
// every ray of the tile.  While (def ^ poss) & childMask == 0 all rays have the same visit mask,
// hence the same DFS state; at the first node where that fails the beam stops and stores the DFS
// state (BeamState); each ray re-evaluates that node with its own exact mask and carries on alone.
// Only "tame" rays (hdt_device.cuh: no NaN/inf anywhere in the mask) with equal direction signs
// form a beam; any other tile starts at the root as before.
//
// Scheduling: the beam kernel is thinly populated (one thread per tile) and latency-bound, so it runs
// on a second stream CONCURRENTLY with the per-ray kernel it serves.  Each BeamState carries a
// release-stored word (launch tag << 2 | status); a warp of the per-ray kernel acquire-loads it once
// when it starts: if its tile's beam of THIS launch is finished it resumes from it, otherwise it
// starts at the root.  Both ways produce the same pixels, so the race only decides how much work is
// saved, never the result.
#pragma once
#include "hdt_device.cuh"

namespace hdt {

struct Iv { float lo, hi; };

// Enclosure of the rays of one tile.  Primary rays share the origin (o.lo == o.hi), shadow rays
// share the direction; the code below does not care.
struct BeamRays { Iv o[3], d[3], inv[3], ainv[3]; };

enum : u32 { kBeamNone = 0, kBeamResume = 1, kBeamHit = 2, kBeamMiss = 3 };

__device__ __forceinline__ void store_release(u32* p, u32 v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ u32 load_acquire(const u32* p)
{
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// DFS state of a tile at the point where its rays stop agreeing.  256 bytes, read by all 32 lanes
// of the tile's warp (broadcast loads).
struct __align__(16) BeamState {
    u32 status, level, pending, handle;      // status = launch tag << 2 | kBeam*; kBeamHit: level/pending/handle = voxel x/y/z
    u32 cm; float radius, cx, cy;
    float cz; u32 pad0; uint2 leaf;
    uint2 stack[kMaxLevels];                 // entries of the levels in `pending`
    u32 pad1[4];
};
static_assert(sizeof(BeamState) == 256, "BeamState is 256 bytes");

// What the per-pixel seed kernels leave for the per-tile beam kernels: the ranges over the tile's
// rays of the quantity that varies (primary rays: direction; shadow rays: origin), and whether the
// tile can form a beam at all (>= 1 ray, all tame, one direction-sign pattern).
struct __align__(16) BeamSeed { float lo[3]; u32 valid; float hi[3]; u32 order; };
static_assert(sizeof(BeamSeed) == 32, "BeamSeed is 32 bytes");

// float <-> unsigned key with the same order (for redux.sync min/max)
__device__ __forceinline__ u32 float_key(float f) { const u32 b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ float key_float(u32 k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

// Warp-wide ranges of v[0..2] over the lanes with `active`; lane 0 writes the tile's seed.
__device__ __forceinline__ void write_seed(BeamSeed* __restrict__ seed, const float v[3], bool active, bool tame, u32 order)
{
    u32 lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const u32 key = float_key(v[k]);
        lo[k] = __reduce_min_sync(0xFFFFFFFFu, active ? key : 0xFFFFFFFFu);
        hi[k] = __reduce_max_sync(0xFFFFFFFFu, active ? key : 0u);
    }
    const u32 any = __ballot_sync(0xFFFFFFFFu, active);
    const u32 bad = __ballot_sync(0xFFFFFFFFu, active && !tame);
    const u32 omin = __reduce_min_sync(0xFFFFFFFFu, active ? order : 8u), omax = __reduce_max_sync(0xFFFFFFFFu, active ? order : 0u);
    if ((threadIdx.x & 31) == 0) {
        BeamSeed s;
#pragma unroll
        for (int k = 0; k < 3; ++k) { s.lo[k] = key_float(lo[k]); s.hi[k] = key_float(hi[k]); }
        s.valid = (any != 0 && bad == 0 && omin == omax) ? 1u : 0u;
        s.order = omin & 7u;
        *reinterpret_cast<uint4*>(seed) = *reinterpret_cast<const uint4*>(&s);
        *(reinterpret_cast<uint4*>(seed) + 1) = *(reinterpret_cast<const uint4*>(&s) + 1);
    }
}

__device__ __forceinline__ Iv iv_mul(Iv a, Iv b)
{
    const float p0 = __fmul_rn(a.lo, b.lo), p1 = __fmul_rn(a.lo, b.hi), p2 = __fmul_rn(a.hi, b.lo), p3 = __fmul_rn(a.hi, b.hi);
    return { fminf(fminf(p0, p1), fminf(p2, p3)), fmaxf(fmaxf(p0, p1), fmaxf(p2, p3)) };
}
// one factor is a point (lo == hi): two products suffice
__device__ __forceinline__ Iv iv_mul_point(Iv a, float b)
{
    const float p0 = __fmul_rn(a.lo, b), p1 = __fmul_rn(a.hi, b);
    return { fminf(p0, p1), fmaxf(p0, p1) };
}

// Interval form of intersection_mask<isRoot, TAME=true>.  Returns def | poss << 8.
// rootState (isRoot only): 0 = every ray enters the root, 1 = every ray misses it, 2 = undecided.
// POINT_O: all rays share the origin (primary rays); POINT_D: all share the direction (shadow rays).
// The specialisations only drop products whose two factors' end points coincide.
template <bool isRoot, bool POINT_O, bool POINT_D>
__device__ __forceinline__ u32 interval_mask(float cx, float cy, float cz, float radius, const BeamRays& br, int& rootState)
{
    const float c[3] = { cx, cy, cz };
    Iv r[3], t[3], a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r[k] = { __fsub_rn(c[k], br.o[k].hi), __fsub_rn(c[k], br.o[k].lo) };
        t[k] = POINT_O ? iv_mul_point(br.inv[k], r[k].lo) : POINT_D ? iv_mul_point(r[k], br.inv[k].lo) : iv_mul(r[k], br.inv[k]);
        a[k] = { __fmaf_rn(-radius, br.ainv[k].hi, t[k].lo), __fmaf_rn(-radius, br.ainv[k].lo, t[k].hi) };
        b[k] = { __fmaf_rn(radius, br.ainv[k].lo, t[k].lo), __fmaf_rn(radius, br.ainv[k].hi, t[k].hi) };
    }
    const Iv tmin = { fmaxf(fmaxf(a[0].lo, a[1].lo), fmaxf(a[2].lo, 0.0f)), fmaxf(fmaxf(a[0].hi, a[1].hi), fmaxf(a[2].hi, 0.0f)) };
    const Iv tmax = { fminf(fminf(b[0].lo, b[1].lo), b[2].lo), fminf(fminf(b[0].hi, b[1].hi), b[2].hi) };
    rootState = 0;
    if (isRoot) {
        if (tmin.lo >= tmax.hi) { rootState = 1; return 0; }
        if (!(tmin.hi < tmax.lo)) { rootState = 2; return 0xFF00u; }
    }
    const Iv h = { __fmul_rn(0.5f, __fadd_rn(tmin.lo, tmax.lo)), __fmul_rn(0.5f, __fadd_rn(tmin.hi, tmax.hi)) };
    u32 def = 0, poss = 0;
    {
        u32 bitsDef = 0, bitsAmb = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const u32 w = 4u >> k;
            const Iv q = POINT_D ? iv_mul_point(h, br.d[k].lo) : iv_mul(h, br.d[k]);
            if (q.lo >= r[k].hi) bitsDef |= w;
            else if (!(q.hi < r[k].lo)) bitsAmb |= w;
        }
        if (!bitsAmb) { def = poss = 1u << bitsDef; }
        else {
#pragma unroll
            for (u32 s = 0; s < 8; ++s) if ((s & ~bitsAmb) == bitsDef) poss |= 1u << s;
        }
    }
    const float eps = 1e-4f;
    Iv rm[3], rp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        rm[k] = { __fsub_rn(r[k].lo, eps), __fsub_rn(r[k].hi, eps) };
        rp[k] = { __fadd_rn(r[k].lo, eps), __fadd_rn(r[k].hi, eps) };
    }
    const u32 HI[3] = { 0xF0u, 0xCCu, 0xAAu }, LO[3] = { 0x0Fu, 0x33u, 0x55u };
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const bool pinDef = tmin.hi <= t[p].lo && t[p].hi <= tmax.lo;
        const bool pinPoss = tmin.lo <= t[p].hi && t[p].lo <= tmax.hi;
        u32 mdef = 0xFFu, mposs = 0xFFu;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k == p) continue;
            const Iv q = POINT_D ? iv_mul_point(t[p], br.d[k].lo) : iv_mul(t[p], br.d[k]);
            u32 Ad = 0, Ap = 0;
            if (q.lo >= rm[k].hi) Ad |= HI[k];
            if (q.hi >= rm[k].lo) Ap |= HI[k];
            if (q.hi <= rp[k].lo) Ad |= LO[k];
            if (q.lo <= rp[k].hi) Ap |= LO[k];
            mdef &= Ad; mposs &= Ap;
        }
        if (pinDef) def |= mdef;
        if (pinPoss) poss |= mposs;
    }
    return def | (poss << 8);
}

// Finish a BeamRays from the ranges of origin and direction: 1/d ranges (rcp.rn is monotone),
// |1/d| ranges, and the validity test (tame, one sign per axis).
__device__ __forceinline__ bool finish_beam(BeamRays& br)
{
    const float lim = 1.2676506e30f;  // 2^100, ray_is_tame
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ok = ok && (br.d[k].lo > 0.0f || br.d[k].hi < 0.0f);        // no zero, no sign change
        ok = ok && fabsf(br.o[k].lo) <= 3.0e38f && fabsf(br.o[k].hi) <= 3.0e38f;
        br.inv[k] = { __frcp_rn(br.d[k].hi), __frcp_rn(br.d[k].lo) };
        const float x = fabsf(br.inv[k].lo), y = fabsf(br.inv[k].hi);
        br.ainv[k] = { fminf(x, y), fmaxf(x, y) };
        ok = ok && br.ainv[k].hi <= lim;
    }
    return ok;
}

// The DFS of Walker::step (hdt_device.cuh) with interval masks.  Writes the tile's BeamState.
// ORDERED beams are primary rays (shared origin), unordered ones shadow rays (shared direction).
// maxVisits bounds the walk (and with it the latency of this thinly populated kernel): the DFS state
// is a valid hand-over point before any node's mask, so the beam may stop wherever it likes.
template <class DAG, bool ORDERED>
__device__ __forceinline__ void beam_traverse(const DAG& dag, const u32 levels, const BeamRays& br, const TraverseTables& tab, const u32 order,
                                              const u32 maxVisits, const u32 tag, BeamState* __restrict__ out)
{
    constexpr bool PO = ORDERED, PD = !ORDERED;
    const u32 leafLevel = levels - 2;
    auto publish = [&](u32 status) { store_release(&out->status, (tag << 2) | status); };
    WalkStack stack;
    u32 level = 0, pending = 0;
    uint2 leaf = make_uint2(0, 0);
    float radius = __uint_as_float((127u + levels - 1u) << 23);
    float cx = radius, cy = radius, cz = radius;
    u32 handle = dag.root();
    u32 cm = dag.header(handle) & 0xFF;
    u32 vm, visits = 0;
    int rootState;
    {
        const u32 m = interval_mask<true, PO, PD>(cx, cy, cz, radius, br, rootState);
        if (rootState == 1) { publish(kBeamMiss); return; }
        if (rootState == 2 || (((m >> 8) ^ m) & cm)) { publish(kBeamNone); return; }
        vm = cm & m & 0xFF;
    }
    for (;;) {
        if (vm == 0) {
            if (pending == 0) { publish(kBeamMiss); return; }
            const u32 nl = 31 - __clz(pending);
            pending ^= 1u << nl;
            const uint2 e = stack[nl];
            handle = e.x; cm = e.y & 0xFF; vm = e.y >> 8;
            const float ra = __uint_as_float(__float_as_uint(radius) + ((level - nl) << 23));
            const float magic = __fmul_rn(ra, 25165824.0f);
            cx = __fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(cx, ra), magic), magic), ra);
            cy = __fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(cy, ra), magic), magic), ra);
            cz = __fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(cz, ra), magic), magic), ra);
            radius = ra;
            level = nl;
        }
        const u32 child = ORDERED ? u32(tab.o[order].child[vm]) : (31 - __clz(vm));
        const float4 st = tab.o[0].step[child];
        vm &= ~__float_as_uint(st.w);
        if (ORDERED || vm) stack[level] = make_uint2(handle, cm | (vm << 8));   // ORDERED: every ancestor, like Walker::step
        if (vm) pending |= 1u << level;
        radius = __fmul_rn(radius, 0.5f);
        cx = __fmaf_rn(st.x, radius, cx); cy = __fmaf_rn(st.y, radius, cy); cz = __fmaf_rn(st.z, radius, cz);
        ++level;
        if (level == levels) {
            out->level = __float2uint_rz(cx); out->pending = __float2uint_rz(cy); out->handle = __float2uint_rz(cz);
            if (ORDERED) {   // the ancestors trace_paths hands on to trace_colors (ancestor_words)
                out->leaf = leaf;
                for (u32 l = kColorTreeDepth; l + 3 <= levels; ++l) out->stack[l] = stack[l];
            }
            publish(kBeamHit);
            return;
        }
        if (level <= leafLevel) {
            const u32 next = dag.child(handle, __popc(cm & (__float_as_uint(st.w) - 1u)) + 1);
            if (level < leafLevel) {
                handle = next;
                cm = dag.header(next) & 0xFF;
            } else {
                leaf = dag.leaf(next);
                cm = first_child_mask(leaf);
            }
        } else {
            cm = second_child_mask(leaf, child);
        }
        if (++visits >= maxVisits) break;  // enough: hand over here
        const u32 m = interval_mask<false, PO, PD>(cx, cy, cz, radius, br, rootState);
        if (((m >> 8) ^ m) & cm) break;    // the rays disagree about this node: hand over
        vm = cm & m & 0xFF;
    }
    out->level = level; out->pending = pending; out->handle = handle;
    out->cm = cm; out->radius = radius; out->cx = cx; out->cy = cy; out->cz = cz; out->leaf = leaf;
    // the pending entries; for primary rays also every ancestor below the colour tree (pending or not)
    const u32 keep = ORDERED ? (pending | (((1u << level) - 1u) & ~((1u << kColorTreeDepth) - 1u))) : pending;
    for (u32 m = keep; m; m &= m - 1) {
        const u32 l = __ffs(m) - 1;
        out->stack[l] = stack[l];
    }
    publish(kBeamResume);
}

// Per-ray side, called by all 32 lanes of the tile's warp before they diverge: the tile's beam status
// if the beam of launch `tag` has been published and every lane has seen it, else kBeamNone.
__device__ __forceinline__ u32 beam_status(const BeamState* __restrict__ bs, const u32 tag)
{
    if (!bs) return kBeamNone;
    const u32 word = load_acquire(&bs->status);
    const bool ready = (word >> 2) == tag;
    if (!__all_sync(0xFFFFFFFFu, ready)) return kBeamNone;
    return word & 3u;
}

// Per-ray side: load a BeamState (status == kBeamResume) into a walker; walk() (hdt_device.cuh) carries on from there.
// The ray is tame by construction.
template <class DAG, bool ORDERED>
__device__ __forceinline__ void resume_from(Walker<DAG>& w, WalkStack& stack, const Ray& ray, const BeamState* __restrict__ bs)
{
    // written by a concurrently running kernel: read through L2 (ld.cg), never the non-coherent path
    const uint4 h0 = __ldcg(reinterpret_cast<const uint4*>(bs));
    const uint4 h1 = __ldcg(reinterpret_cast<const uint4*>(bs) + 1);
    const uint4 h2 = __ldcg(reinterpret_cast<const uint4*>(bs) + 2);
    w.level = h0.y; w.pending = h0.z; w.handle = h0.w;
    w.cm = h1.x; w.radius = __uint_as_float(h1.y); w.cx = __uint_as_float(h1.z); w.cy = __uint_as_float(h1.w);
    w.cz = __uint_as_float(h2.x); w.leaf = make_uint2(h2.z, h2.w);
    const u32 keep = ORDERED ? (w.pending | (((1u << w.level) - 1u) & ~((1u << kColorTreeDepth) - 1u))) : w.pending;
    for (u32 m = keep; m; m &= m - 1) {
        const u32 l = __ffs(m) - 1;
        stack[l] = __ldcg(&bs->stack[l]);
    }
    w.vm = w.cm & intersection_mask<false, true>(w.cx, w.cy, w.cz, w.radius, ray);
}

}  // namespace hdt
