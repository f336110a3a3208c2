// Device-side building blocks of the tracer: DAG accessors, ray/node intersection, traversal.
//
// Arithmetic contract: every operation below is a single correctly-rounded IEEE
// operation written with an explicit intrinsic, and every fused multiply-add the reference's
// compiled kernels contain (read off its PTX/SASS) is an explicit __fma_rn/__fmaf_rn.  The file is
// compiled with -fmad=false so the compiler adds none of its own.
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace hdt {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

// Any-hit walks (trace_shadows) may visit children in any order -- the result is "is there a voxel", not which one.
// 0: highest child first like the reference (tracer.cu:493); 1: lowest child first (= front to back for the sun's direction,
// whose components are all positive: occluders near the surface point are met first).  A/B on B200, 1080p depth 17
// (profiles/r2_ab.md): shadows pass 0.328 ms against 0.346 ms, frames identical.
#ifndef HDT_ANYHIT_LOW_FIRST
#define HDT_ANYHIT_LOW_FIRST 1
#endif
constexpr u32 kMaxLevels = 24;   // float node centres stay exact below 2^24
constexpr u32 kPageWords = 512;  // C_pageSize, hash_dag_globals.h:10

// ---------------------------------------------------------------------------------------------
// Where pixel (x, y) of the frame lives in this rank's buffers.  world == 1: plain row-major.
// world > 1: tiles of (1<<tileLog2)^2 pixels, tile t owned by rank t % world, owned tiles stored
// back to back (slot = t / world), each tile row-major.
// ---------------------------------------------------------------------------------------------
struct PixelMap {
    u32 width, height, tileLog2, tilesX, tilesY, world, rank;
    __host__ __device__ __forceinline__ u64 index(u32 x, u32 y) const
    {
        if (world == 1) return u64(y) * width + x;
        const u32 t = (y >> tileLog2) * tilesX + (x >> tileLog2), m = (1u << tileLog2) - 1;
        return (u64(t / world) << (2 * tileLog2)) + (u64(y & m) << tileLog2) + (x & m);
    }
};

// ---------------------------------------------------------------------------------------------
// DAG accessors.  A "handle" is whatever addresses a node's first word cheaply:
//   BasicDAG: the word index itself            (basic_dag.h:20-35)
//   HashDAG : the PHYSICAL word index, i.e. the virtual pointer already pushed through the page
//             table (hash_table.h:156-173).  Nodes never straddle a page (hash_table.h:416-442),
//             so header and child pointers of one node share one translation.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 first_child_mask(uint2 leaf);

struct BasicDagDev {
    const u32* __restrict__ data;
    __device__ __forceinline__ u32 root() const { return 0; }
    __device__ __forceinline__ u32 raw_root() const { return 0; }
    __device__ __forceinline__ u32 header(u32 h) const { return __ldg(data + h); }
    // h + off is formed in 32 bits (a node lies inside an array of < 2^32 words): one IMAD.WIDE instead of a 64-bit add chain
    __device__ __forceinline__ u32 raw_child(u32 h, u32 off) const { return __ldg(data + u32(h + off)); }
    __device__ __forceinline__ u32 to_handle(u32 ptr) const { return ptr; }
    __device__ __forceinline__ u32 child(u32 h, u32 off) const { return __ldg(data + u32(h + off)); }
    __device__ __forceinline__ uint2 leaf(u32 h) const { return make_uint2(__ldg(data + h), __ldg(data + h + 1)); }
    __device__ __forceinline__ u32 leaf_first_mask(u32, uint2 l) const { return first_child_mask(l); }
};

struct HashDagDev {
    const u32* __restrict__ pool;
    const u32* __restrict__ pageTable;
    u32 firstNodeIndex;
    __device__ __forceinline__ u32 to_handle(u32 vptr) const { return __ldg(pageTable + (vptr >> 9)) * kPageWords + (vptr & (kPageWords - 1)); }
    __device__ __forceinline__ u32 root() const { return to_handle(firstNodeIndex); }
    __device__ __forceinline__ u32 raw_root() const { return firstNodeIndex; }
    __device__ __forceinline__ u32 header(u32 h) const { return __ldg(pool + h); }
    __device__ __forceinline__ u32 raw_child(u32 h, u32 off) const { return __ldg(pool + u32(h + off)); }
    __device__ __forceinline__ u32 child(u32 h, u32 off) const { return to_handle(__ldg(pool + u32(h + off))); }
    // leaves sit at even bucket positions in 512-word pages: 8-byte aligned (hash_table.h:367-392)
    __device__ __forceinline__ uint2 leaf(u32 h) const { return __ldg(reinterpret_cast<const uint2*>(pool + h)); }
    __device__ __forceinline__ u32 leaf_first_mask(u32, uint2 l) const { return first_child_mask(l); }
};

// A HashDAG whose child words have been pushed through the page table ONCE (hdt_hash_dag_resolve, hdt_resolve.cuh):
// `pool` is a library-format copy of the caller's pool with the same physical layout in which every child pointer
// holds the physical word index of the child.  A descent is then two dependent loads (child word, child header) like
// BasicDAG's instead of three.  The caller's pool and page table stay at hand for the two debug views that display the
// reference's virtual indices.
// PREFIX: the DAG comes with a prefix pool (hdt_resolve.cuh), whose word at a pointer to a 64-bit leaf also carries the leaf's
// first-level child mask in its top byte: the traversal reads it (one load, issued beside the pointer load) instead of
// reducing the 64 bits itself (11 instructions that every lane of a warp pays for whenever one lane is at that level).
template <bool PREFIX>
struct HashDagResolvedDevT {
    const u32* __restrict__ pool;        // resolved copy
    const u32* __restrict__ vpool;       // the caller's pool (virtual child pointers)
    const u32* __restrict__ pageTable;
    const u32* __restrict__ prefix;      // per child-pointer word: voxels under the node's earlier children | leaf mask << 24; null unless PREFIX
    u32 firstNodeIndex;
    __device__ __forceinline__ u32 to_handle(u32 vptr) const { return __ldg(pageTable + (vptr >> 9)) * kPageWords + (vptr & (kPageWords - 1)); }
    __device__ __forceinline__ u32 root() const { return to_handle(firstNodeIndex); }
    __device__ __forceinline__ u32 raw_root() const { return firstNodeIndex; }
    __device__ __forceinline__ u32 header(u32 h) const { return __ldg(pool + h); }
    __device__ __forceinline__ u32 raw_child(u32 h, u32 off) const { return __ldg(vpool + u32(h + off)); }
    __device__ __forceinline__ u32 child(u32 h, u32 off) const { return __ldg(pool + u32(h + off)); }
    __device__ __forceinline__ uint2 leaf(u32 h) const { return __ldg(reinterpret_cast<const uint2*>(pool + h)); }
    __device__ __forceinline__ u32 leaf_first_mask(u32 pointerWord, uint2 l) const { return PREFIX ? (__ldg(prefix + pointerWord) >> 24) : first_child_mask(l); }
};
using HashDagResolvedDev = HashDagResolvedDevT<false>;
using HashDagPrefixDev = HashDagResolvedDevT<true>;

// Bit k of the result = byte k of the 64-bit leaf is non-zero.  Per word: the carry of (byte & 0x7F) + 0x7F ORed with the
// byte's own top bit marks a non-zero byte at bit 7 of the byte; the multiplication gathers bits 7/15/23/31 into bits
// 28..31 (no two partial products meet above bit 23).
__device__ __forceinline__ u32 first_child_mask(uint2 leaf)
{
    const u32 a = (((leaf.x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | leaf.x) & 0x80808080u;
    const u32 b = (((leaf.y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | leaf.y) & 0x80808080u;
    return ((a * 0x00204081u) >> 28) | (((b * 0x00204081u) >> 24) & 0xF0u);
}
__device__ __forceinline__ u32 second_child_mask(uint2 leaf, u32 firstChild)
{
    return (((firstChild & 4) ? leaf.y : leaf.x) >> ((firstChild & 3) * 8)) & 0xFF;
}

struct Ray { float ox, oy, oz, dx, dy, dz, ix, iy, iz; };

// index of the highest set bit (x != 0): one FLO instead of the 31 - clz(x) pair
__device__ __forceinline__ u32 top_bit(u32 x) { u32 r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x)); return r; }

// Comparison -> all-ones / zero word (one FSET instead of FSETP + SEL).  Ordered compares: false
// on NaN, like the C++ comparisons of the reference.
__device__ __forceinline__ u32 set_ge(float a, float b) { u32 d; asm("set.ge.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ u32 set_le(float a, float b) { u32 d; asm("set.le.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }

// A ray is "tame" when no NaN can reach the slab min/max below: finite origin and |1/d| <= 2^100
// on every axis (so radius*|1/d| is finite and centre*1/d cannot be 0*inf or inf-inf).  Only then
// may the reference's ternary max/min (cuda_math.h:42-44) be replaced by fmaxf/fminf, which differ
// from it solely when the second operand is NaN.  Axis-parallel rays take the exact-ternary path.
__device__ __forceinline__ bool ray_is_tame(const Ray& r)
{
    const float lim = 1.2676506e30f;  // 2^100
    return fabsf(r.ix) <= lim && fabsf(r.iy) <= lim && fabsf(r.iz) <= lim && fabsf(r.ox) <= 3.0e38f && fabsf(r.oy) <= 3.0e38f && fabsf(r.oz) <= 3.0e38f;
}

// One of the three axis-plane tests of tracer.cu:91-133: if the ray meets the node's mid-plane at t within [tmin, tmax],
// OR into `mask` the octants on the side(s) of the crossing point q = (q1, q2), each side decided with an epsilon band
// (lo = centre - eps, hi = centre + eps).  The reference forms, per coordinate, A = (q >= lo ? HI : 0) + (q <= hi ? LO : 0)
// with LO = ~HI, and ORs A1 & A2.  Here the COMPLEMENT of each A is built in two instructions -- n = (q <= hi ? 0 : LO),
// then |= HI unless q >= lo -- and one LOP3 does mask | ~(n1 | n2).  The range test rides on the .AND input of the first
// pair of compares: out of range (or NaN) makes n1 = 0xFF, so the plane contributes nothing.  6 FSETP + 2 SEL + 2 predicated
// OR + 1 LOP3, no branches.  Bits above bit 7 of the result are junk (the caller masks).
template <u32 HI1, u32 HI2>
__device__ __forceinline__ u32 plane_or(u32 mask, float tmin, float tmax, float t, float q1, float lo1, float hi1, float q2, float lo2, float hi2)
{
    u32 out;
    asm("{\n\t"
        ".reg .pred pin, pa, pb, pc, pd;\n\t"
        ".reg .u32 n1, n2;\n\t"
        "setp.le.f32 pin, %1, %3;\n\t"
        "setp.le.and.f32 pin, %3, %2, pin;\n\t"
        "setp.ge.and.f32 pa, %4, %5, pin;\n\t"
        "setp.le.and.f32 pb, %4, %6, pin;\n\t"
        "setp.ge.f32 pc, %7, %8;\n\t"
        "setp.le.f32 pd, %7, %9;\n\t"
        "selp.u32 n1, 0, %11, pb;\n\t"
        "@!pa or.b32 n1, n1, %10;\n\t"
        "selp.u32 n2, 0, %13, pd;\n\t"
        "@!pc or.b32 n2, n2, %12;\n\t"
        "lop3.b32 %0, %14, n1, n2, 0xF1;\n\t"
        "}"
        : "=r"(out)
        : "f"(tmin), "f"(tmax), "f"(t), "f"(q1), "f"(lo1), "f"(hi1), "f"(q2), "f"(lo2), "f"(hi2), "n"(HI1), "n"(HI1 ^ 0xFFu), "n"(HI2), "n"(HI2 ^ 0xFFu), "r"(mask));
    return out;
}

// 1 << (4a + 2b + c) for three comparisons a = (ha >= ra), ...: the octant of the ray's mid-point (tracer.cu:57-63).
__device__ __forceinline__ u32 octant_bit(float hx, float rx, float hy, float ry, float hz, float rz)
{
    u32 m;
    asm("{\n\t"
        ".reg .pred p4, p2, p1;\n\t"
        "setp.ge.f32 p4, %1, %2;\n\t"
        "setp.ge.f32 p2, %3, %4;\n\t"
        "setp.ge.f32 p1, %5, %6;\n\t"
        "selp.u32 %0, 16, 1, p4;\n\t"
        "@p2 shl.b32 %0, %0, 2;\n\t"
        "@p1 shl.b32 %0, %0, 1;\n\t"
        "}"
        : "=r"(m)
        : "f"(hx), "f"(rx), "f"(hy), "f"(ry), "f"(hz), "f"(rz));
    return m;
}

// tracer.cu:19-136 for the node with centre (cx,cy,cz) and half-size `radius`.
//  * Centre and radius are exact floats (multiples of 0.5 below 2^24, radius a power of two), so
//    radius*|inv| is exact and tmid -+ radius*|inv| equals fma(-+radius, |inv|, tmid) bit for bit.
//  * The three plane tests are evaluated without branches: every comparison becomes a 0/~0 word and
//    the reference's byte constants are merged with LOP3.  Bits above bit 7 of the result are junk;
//    the caller ANDs with an 8-bit child mask.
template <bool isRoot, bool TAME>
__device__ __forceinline__ u32 intersection_mask(float cx, float cy, float cz, float radius, const Ray& r)
{
    const float rx = __fsub_rn(cx, r.ox), ry = __fsub_rn(cy, r.oy), rz = __fsub_rn(cz, r.oz);
    const float tx = __fmul_rn(rx, r.ix), ty = __fmul_rn(ry, r.iy), tz = __fmul_rn(rz, r.iz);
    const float nr = -radius;
    const float ax = __fmaf_rn(nr, fabsf(r.ix), tx), ay = __fmaf_rn(nr, fabsf(r.iy), ty), az = __fmaf_rn(nr, fabsf(r.iz), tz);
    const float bx = __fmaf_rn(radius, fabsf(r.ix), tx), by = __fmaf_rn(radius, fabsf(r.iy), ty), bz = __fmaf_rn(radius, fabsf(r.iz), tz);
    float tmin, tmax;
    if (TAME) {
        tmin = fmaxf(fmaxf(ax, ay), fmaxf(az, 0.0f));
        tmax = fminf(fminf(bx, by), bz);
    } else {
        const float ayz = (ay > az) ? ay : az;
        const float a3 = (ax > ayz) ? ax : ayz;
        tmin = fmaxf(a3, 0.0f);
        const float byz = (by < bz) ? by : bz;
        tmax = (bx < byz) ? bx : byz;
    }
    if (isRoot && (tmin >= tmax)) return 0;

    const float h = __fmul_rn(0.5f, __fadd_rn(tmin, tmax));
    u32 mask = octant_bit(__fmul_rn(h, r.dx), rx, __fmul_rn(h, r.dy), ry, __fmul_rn(h, r.dz), rz);

    // Plane tests without branches (see plane_mask): an out-of-range plane contributes nothing.
    const float eps = 1e-4f;
    const float rxm = __fsub_rn(rx, eps), rxp = __fadd_rn(rx, eps);
    const float rym = __fsub_rn(ry, eps), ryp = __fadd_rn(ry, eps);
    const float rzm = __fsub_rn(rz, eps), rzp = __fadd_rn(rz, eps);
    mask = plane_or<0xCC, 0xAA>(mask, tmin, tmax, tx, __fmul_rn(tx, r.dy), rym, ryp, __fmul_rn(tx, r.dz), rzm, rzp);
    mask = plane_or<0xF0, 0xAA>(mask, tmin, tmax, ty, __fmul_rn(ty, r.dx), rxm, rxp, __fmul_rn(ty, r.dz), rzm, rzp);
    mask = plane_or<0xF0, 0xCC>(mask, tmin, tmax, tz, __fmul_rn(tz, r.dx), rxm, rxp, __fmul_rn(tz, r.dy), rym, ryp);
    return mask;
}

// Shared-memory tables, filled once per CTA.  One 384-byte block per ray order (tracer.cu:7-17: children are visited in
// the order child ^ order, child = 0..7):
//   o[order].child[mask] = next_child(order, mask): first set bit of mask in that order;
//   o[order].step[c]     = (+-1, +-1, +-1, 1<<c): direction of child c's centre from its parent's (child bit 4 -> x,
//                          2 -> y, 1 -> z, path.h:20-28) and its mask bit (the same in every block).
// A ray keeps ONE register -- the shared-memory address of its block -- and reaches both tables from it with immediate
// offsets (table_child / table_step below).
struct TraverseTables {
    struct Block { u8 child[256]; float4 step[8]; } o[8];
};

static_assert(sizeof(TraverseTables::Block) == 384 && sizeof(TraverseTables) == 3072, "table layout is addressed by hand");
static_assert(sizeof(TraverseTables) % 16 == 0, "copied as uint4");

// Host-side initialisation (once per context); CTAs copy the table with 16-byte loads.
inline void build_tables(TraverseTables& t)
{
    for (u32 order = 0; order < 8; ++order) {
        for (u32 mask = 0; mask < 256; ++mask) {
            u32 res = 0;
            for (int child = 7; child >= 0; --child) {
                const u32 c = u32(child) ^ order;
                if (mask & (1u << c)) res = c;
            }
            t.o[order].child[mask] = u8(res);
        }
        for (u32 c = 0; c < 8; ++c) {
            const u32 bit = 1u << c;
            float w;
            memcpy(&w, &bit, 4);
            t.o[order].step[c] = make_float4((c & 4) ? 1.f : -1.f, (c & 2) ? 1.f : -1.f, (c & 1) ? 1.f : -1.f, w);
        }
    }
}

__device__ __forceinline__ void load_tables(TraverseTables& dst, const TraverseTables* __restrict__ src)
{
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(&dst);
    for (u32 i = threadIdx.x; i < sizeof(TraverseTables) / 16; i += blockDim.x) d[i] = __ldg(s + i);
}

// Shared-window address of the block of ray order `order`.  The empty asm keeps the compiler from re-deriving the
// address from the ray's direction signs inside the traversal loop (it did: 8 instructions per node visit).
__device__ __forceinline__ u32 table_block(const TraverseTables& tab, u32 order)
{
    u32 a = u32(__cvta_generic_to_shared(&tab)) + order * u32(sizeof(TraverseTables::Block));
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ u32 table_child(u32 block, u32 mask)
{
    u32 c;
    asm("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(block + mask));
    return c;
}
__device__ __forceinline__ float4 table_step(u32 block, u32 child)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+256];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(block + child * 16u));
    return v;
}

// Stack DFS shared by trace_paths (ORDERED: children in ray order, first voxel wins,
// tracer.cu:166-249) and trace_shadows (!ORDERED: highest child first, any voxel,
// tracer.cu:458-542).  Same nodes visited in the same order as the reference's loop; what differs
// is bookkeeping, chosen because the loop is ALU-pipe bound on sm_100a (profiles/):
//   * the node is tracked by its float centre and radius instead of the integer Path: descending
//     adds +-radius/2 per axis (one table load + three FFMA), ascending re-aligns the centre with
//     four exact FADDs per axis (round-to-multiple by the 1.5*2^23 trick); every value involved is
//     a multiple of 0.5 below 2^24, so all of it is exact and equals float(path << shift) + radius;
//   * `pending` has bit L set iff stack[L] still has unvisited children: ascending is one FLO
//     instead of a chain of dependent local loads, and exhausted entries are never stored;
//   * one intersection_mask call site after the (divergent) child fetch instead of three;
//   * HashDAG handles are physical: one page-table lookup per node instead of one per word.
// Per-ray DFS state.  Scalars only, so it stays in registers; the stack (one uint2 per level:
// .x = handle, .y = childMask | visitMask << 8) is a separate local-memory array.
using WalkStack = uint2[kMaxLevels];

template <class DAG>
struct Walker {
    u32 level, pending, handle, cm, vm;
    uint2 leaf;
    float radius, cx, cy, cz;

    template <bool TAME>
    __device__ __forceinline__ void start(const DAG& dag, const u32 levels, const Ray& ray)
    {
        level = 0; pending = 0; leaf = make_uint2(0, 0);
        radius = __uint_as_float((127u + levels - 1u) << 23);   // 2^(levels-1)
        cx = cy = cz = radius;
        handle = dag.root();
        cm = dag.header(handle) & 0xFF;
        vm = cm & intersection_mask<true, TAME>(cx, cy, cz, radius, ray);
    }

    // One descent (preceded by an ascent if the current node is exhausted).
    // Returns 0: keep going, 1: reached a voxel (cx,cy,cz = its centre), 2: left the DAG.
    // `block`: the ray's table block (table_block).  ORDERED walks store every level's entry, pending or not: when the
    // walk ends on a voxel, stack[d] holds the handle and child mask of the voxel's ancestor of every depth d, which
    // trace_paths hands to trace_colors (ancestor_words below).
    template <bool ORDERED, bool TAME>
    __device__ __forceinline__ int step(const DAG& dag, const u32 levels, const Ray& ray, const u32 block, WalkStack& stack)
    {
        const u32 leafLevel = levels - 2;
        if (vm == 0) {
            if (pending == 0) return 2;
            const u32 nl = top_bit(pending);
            pending ^= 1u << nl;
            const uint2 e = stack[nl];
            handle = e.x; cm = e.y & 0xFF; vm = e.y >> 8;
            // centre of the ancestor `level - nl` levels up: the cell of size S = 2*ra that contains the current centre.  Adding
            // 1.5*2^23*S (one ulp = S) with rounding towards -inf drops the centre's offset inside that cell exactly
            // (coordinates are far below 2^22*S), subtracting it again leaves the cell's corner, + ra its centre.
            const float ra = __uint_as_float(__float_as_uint(radius) + ((level - nl) << 23));
            const float magic = __fmul_rn(ra, 25165824.0f);   // 1.5 * 2^24 * ra = 1.5 * 2^23 * S
            cx = __fadd_rn(__fsub_rn(__fadd_rd(cx, magic), magic), ra);
            cy = __fadd_rn(__fsub_rn(__fadd_rd(cy, magic), magic), ra);
            cz = __fadd_rn(__fsub_rn(__fadd_rd(cz, magic), magic), ra);
            radius = ra;
            level = nl;
        }
        const u32 child = ORDERED ? table_child(block, vm) : HDT_ANYHIT_LOW_FIRST ? u32(__ffs(vm) - 1) : top_bit(vm);
        const float4 st = table_step(block, child);
        vm &= ~__float_as_uint(st.w);
        if (ORDERED || vm) stack[level] = make_uint2(handle, (cm & 0xFF) | (vm << 8));
        if (vm) pending |= 1u << level;
        radius = __fmul_rn(radius, 0.5f);
        cx = __fmaf_rn(st.x, radius, cx); cy = __fmaf_rn(st.y, radius, cy); cz = __fmaf_rn(st.z, radius, cz);
        ++level;
        if (level == levels) return 1;
        // The child's header is loaded here but not touched before the ~90 instructions of mask arithmetic (which need
        // nothing from memory) have been issued: `cm` keeps the whole header word (count24 rides along above bit 7) and the
        // first instruction that reads it is the LOP3 that also needs the finished mask, so the load's latency hides behind
        // the arithmetic.  (With `cm = header & 0xFF` the AND sat right behind the LDG in the SASS and the warp stalled
        // there, in front of the arithmetic.)  Users of cm: popc(cm & (bit - 1)) with bit <= 0x80, and the stack push,
        // which masks it.
        if (level <= leafLevel) {
            const u32 off = __popc(cm & (__float_as_uint(st.w) - 1u)) + 1;
            const u32 next = dag.child(handle, off);
            if (level < leafLevel) {
                handle = next;
                cm = dag.header(next);
            } else {
                leaf = dag.leaf(next);
                cm = dag.leaf_first_mask(handle + off, leaf);
            }
        } else {
            cm = second_child_mask(leaf, child);
        }
        vm = cm & intersection_mask<false, TAME>(cx, cy, cz, radius, ray) & 0xFF;
        return 0;
    }
    // voxel coordinates once step() returned 1: centre = corner + 0.5
    __device__ __forceinline__ void voxel(u32& x, u32& y, u32& z) const { x = __float2uint_rz(cx); y = __float2uint_rz(cy); z = __float2uint_rz(cz); }
};

// Finish a walk that has been started (Walker::start) or resumed (traverse_from, hdt_beam.cuh).
template <class DAG, bool ORDERED, bool TAME>
__device__ __forceinline__ bool walk(Walker<DAG>& w, WalkStack& stack, const DAG& dag, const u32 levels, const Ray& ray, const u32 block,
                                     u32& outx, u32& outy, u32& outz)
{
    for (;;) {
        const int r = w.template step<ORDERED, TAME>(dag, levels, ray, block, stack);
        if (r == 1) { w.voxel(outx, outy, outz); return true; }
        if (r == 2) { outx = outy = outz = 0; return false; }
    }
}

// What trace_paths leaves for trace_colors about a hit voxel (x, y, z): for every ancestor of depth d in [10, levels-3] --
// the DAG levels below the colour tree (hash_dag_globals.h:7), whose earlier siblings trace_colors has to count
// (tracer.cu:391-420) -- the physical index of the child-pointer word the path follows out of that ancestor, and the 64-bit
// leaf the voxel lies in.  `entry(d)` returns the walk's stack entry of depth d (handle, childMask | ...).
constexpr u32 kColorTreeDepth = 10;   // C_colorTreeLevels
constexpr u32 kMaxAncestorWords = 6;  // depths 10 .. 15: DAGs of up to 18 levels
struct AncestorRecord { uint4 a, b; };   // a = words of depths 10..13, b = { depth 14, depth 15, leaf.x, leaf.y }

template <class Entry>
__device__ __forceinline__ AncestorRecord ancestor_words(const u32 levels, const u32 x, const u32 y, const u32 z, const uint2 leaf, Entry entry)
{
    u32 w[kMaxAncestorWords];
#pragma unroll
    for (u32 i = 0; i < kMaxAncestorWords; ++i) {
        const u32 d = kColorTreeDepth + i;
        w[i] = 0;
        if (d + 3 <= levels) {
            const uint2 e = entry(d);
            const u32 sh = levels - 1 - d;
            const u32 child = (((x >> sh) & 1) << 2) | (((y >> sh) & 1) << 1) | ((z >> sh) & 1);
            w[i] = e.x + 1 + __popc(e.y & 0xFFu & ((1u << child) - 1u));
        }
    }
    AncestorRecord r;
    r.a = make_uint4(w[0], w[1], w[2], w[3]);
    r.b = make_uint4(w[4], w[5], leaf.x, leaf.y);
    return r;
}


// TracePathsParams (tracer.h:81-91) + the float camera position the kernels would otherwise
// re-convert from double at every use (make_float3(cameraPosition), tracer.cu:157).
struct CameraParams { double cam[3], rayMin[3], ddx[3], ddy[3]; float camf[3]; float pad; };

// tracer.cu:158 / :622 with the reference's contraction (see oracle/hdo_oracle.cpp primary_direction)
__device__ __forceinline__ void primary_direction(const CameraParams& p, u32 px, u32 cameraRow, double& dx, double& dy, double& dz)
{
    const double fx = __uint2double_rn(px), fy = __uint2double_rn(cameraRow);
    const double vx = __dsub_rn(__fma_rn(fy, p.ddy[0], __fma_rn(fx, p.ddx[0], p.rayMin[0])), p.cam[0]);
    const double vy = __dsub_rn(__fma_rn(fy, p.ddy[1], __fma_rn(fx, p.ddx[1], p.rayMin[1])), p.cam[1]);
    const double vz = __dsub_rn(__fma_rn(fy, p.ddy[2], __fma_rn(fx, p.ddx[2], p.rayMin[2])), p.cam[2]);
    const double len = __dsqrt_rn(__fma_rn(vz, vz, __fma_rn(vx, vx, __dmul_rn(vy, vy))));
    const double r = __drcp_rn(len);
    dx = __dmul_rn(vx, r); dy = __dmul_rn(vy, r); dz = __dmul_rn(vz, r);
}

}  // namespace hdt
