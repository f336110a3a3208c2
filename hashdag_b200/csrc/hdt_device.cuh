// Device-side building blocks of the tracer: DAG accessors, ray/node intersection, traversal.
//
// Arithmetic contract: every operation below is a single correctly-rounded IEEE
// operation written with an explicit intrinsic, and every fused multiply-add the reference's
// compiled kernels contain (read off its PTX/SASS) is an explicit __fma_rn/__fmaf_rn.  The file is
// compiled with -fmad=false so the compiler adds none of its own.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hdt {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

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
struct BasicDagDev {
    const u32* __restrict__ data;
    __device__ __forceinline__ u32 root() const { return 0; }
    __device__ __forceinline__ u32 raw_root() const { return 0; }
    __device__ __forceinline__ u32 header(u32 h) const { return __ldg(data + h); }
    __device__ __forceinline__ u32 raw_child(u32 h, u32 off) const { return __ldg(data + h + off); }
    __device__ __forceinline__ u32 to_handle(u32 ptr) const { return ptr; }
    __device__ __forceinline__ u32 child(u32 h, u32 off) const { return __ldg(data + h + off); }
    __device__ __forceinline__ uint2 leaf(u32 h) const { return make_uint2(__ldg(data + h), __ldg(data + h + 1)); }
};

struct HashDagDev {
    const u32* __restrict__ pool;
    const u32* __restrict__ pageTable;
    u32 firstNodeIndex;
    __device__ __forceinline__ u32 to_handle(u32 vptr) const { return __ldg(pageTable + (vptr >> 9)) * kPageWords + (vptr & (kPageWords - 1)); }
    __device__ __forceinline__ u32 root() const { return to_handle(firstNodeIndex); }
    __device__ __forceinline__ u32 raw_root() const { return firstNodeIndex; }
    __device__ __forceinline__ u32 header(u32 h) const { return __ldg(pool + h); }
    __device__ __forceinline__ u32 raw_child(u32 h, u32 off) const { return __ldg(pool + h + off); }
    __device__ __forceinline__ u32 child(u32 h, u32 off) const { return to_handle(__ldg(pool + h + off)); }
    // leaves sit at even bucket positions in 512-word pages: 8-byte aligned (hash_table.h:367-392)
    __device__ __forceinline__ uint2 leaf(u32 h) const { return __ldg(reinterpret_cast<const uint2*>(pool + h)); }
};

// base_dag.h:16-58
__device__ __forceinline__ u32 first_child_mask(uint2 leaf)
{
    u32 a = leaf.x | (leaf.x >> 4); a |= a >> 2; a |= a >> 1; a &= 0x01010101u;
    u32 b = leaf.y | (leaf.y >> 4); b |= b >> 2; b |= b >> 1; b &= 0x01010101u;
    return ((a * 0x01020408u) >> 24) | (((b * 0x01020408u) >> 24) << 4);
}
__device__ __forceinline__ u32 second_child_mask(uint2 leaf, u32 firstChild)
{
    return (((firstChild & 4) ? leaf.y : leaf.x) >> ((firstChild & 3) * 8)) & 0xFF;
}

struct Ray { float ox, oy, oz, dx, dy, dz, ix, iy, iz; };

// tracer.cu:19-136.  Centre and radius are exact floats (integers / halves below 2^24, power-of-two
// radius), so radius*|inv| is exact and pmin/pmax are the same whether or not they are fused.
template <bool isRoot>
__device__ __forceinline__ u32 intersection_mask(u32 level, u32 levels, u32 px, u32 py, u32 pz, const Ray& r)
{
    const u32 shift = levels - level;
    const float radius = __uint2float_rn(1u << (shift - 1));
    const float cx = __fadd_rn(radius, __uint2float_rn(px << shift));
    const float cy = __fadd_rn(radius, __uint2float_rn(py << shift));
    const float cz = __fadd_rn(radius, __uint2float_rn(pz << shift));
    const float rx = __fsub_rn(cx, r.ox), ry = __fsub_rn(cy, r.oy), rz = __fsub_rn(cz, r.oz);
    const float tx = __fmul_rn(rx, r.ix), ty = __fmul_rn(ry, r.iy), tz = __fmul_rn(rz, r.iz);
    const float sx = __fmul_rn(radius, fabsf(r.ix)), sy = __fmul_rn(radius, fabsf(r.iy)), sz = __fmul_rn(radius, fabsf(r.iz));

    const float ax = __fsub_rn(tx, sx), ay = __fsub_rn(ty, sy), az = __fsub_rn(tz, sz);
    const float ayz = (ay > az) ? ay : az;           // ternary max keeps the reference's NaN behaviour
    const float a3 = (ax > ayz) ? ax : ayz;
    const float tmin = fmaxf(a3, 0.0f);
    const float bx = __fadd_rn(tx, sx), by = __fadd_rn(ty, sy), bz = __fadd_rn(tz, sz);
    const float byz = (by < bz) ? by : bz;
    const float tmax = (bx < byz) ? bx : byz;
    if (isRoot && (tmin >= tmax)) return 0;

    u32 mask;
    {
        const float h = __fmul_rn(0.5f, __fadd_rn(tmin, tmax));
        const float qx = __fmul_rn(h, r.dx), qy = __fmul_rn(h, r.dy), qz = __fmul_rn(h, r.dz);
        mask = 1u << (((qx >= rx) ? 4u : 0u) + ((qy >= ry) ? 2u : 0u) + ((qz >= rz) ? 1u : 0u));
    }
    const float eps = 1e-4f;
    const float rxm = __fsub_rn(rx, eps), rxp = __fadd_rn(rx, eps);
    const float rym = __fsub_rn(ry, eps), ryp = __fadd_rn(ry, eps);
    const float rzm = __fsub_rn(rz, eps), rzp = __fadd_rn(rz, eps);
    if (tmin <= tx && tx <= tmax) {
        const float qy = __fmul_rn(tx, r.dy), qz = __fmul_rn(tx, r.dz);
        const u32 A = ((qy >= rym) ? 0xCCu : 0u) | ((qy <= ryp) ? 0x33u : 0u);
        const u32 B = ((qz >= rzm) ? 0xAAu : 0u) | ((qz <= rzp) ? 0x55u : 0u);
        mask |= A & B;
    }
    if (tmin <= ty && ty <= tmax) {
        const float qx = __fmul_rn(ty, r.dx), qz = __fmul_rn(ty, r.dz);
        const u32 C = ((qx >= rxm) ? 0xF0u : 0u) | ((qx <= rxp) ? 0x0Fu : 0u);
        const u32 D = ((qz >= rzm) ? 0xAAu : 0u) | ((qz <= rzp) ? 0x55u : 0u);
        mask |= C & D;
    }
    if (tmin <= tz && tz <= tmax) {
        const float qx = __fmul_rn(tz, r.dx), qy = __fmul_rn(tz, r.dy);
        const u32 E = ((qx >= rxm) ? 0xF0u : 0u) | ((qx <= rxp) ? 0x0Fu : 0u);
        const u32 F = ((qy >= rym) ? 0xCCu : 0u) | ((qy <= ryp) ? 0x33u : 0u);
        mask |= E & F;
    }
    return mask;
}

// next_child (tracer.cu:7-17) as a table: lut[order*256 + mask] = first set bit of mask in the
// order child ^ order, child = 0..7.  2 KB of shared memory, filled once per CTA.
__device__ __forceinline__ void fill_next_child_lut(u8* lut)
{
    for (u32 i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
        const u32 order = i >> 8, mask = i & 255;
        u32 res = 0;
        for (int child = 7; child >= 0; --child) {
            const u32 c = u32(child) ^ order;
            if (mask & (1u << c)) res = c;
        }
        lut[i] = u8(res);
    }
}

// Stack DFS shared by trace_paths (ORDERED: children in ray order, first voxel wins,
// tracer.cu:166-249) and trace_shadows (!ORDERED: highest child first, any voxel,
// tracer.cu:458-542).  Differences from the reference's loop, none of which change what is
// visited or in which order:
//   * `pending` has bit L set iff stack[L] still has unvisited children, so ascending is one
//     clz instead of a chain of dependent local-memory loads, and empty entries are never stored;
//   * one intersection_mask call site after the (divergent) child fetch instead of three;
//   * HashDAG handles are physical, one page-table lookup per node instead of one per word.
template <class DAG, bool ORDERED>
__device__ __forceinline__ bool traverse(const DAG& dag, const u32 levels, const Ray& ray, const u8* __restrict__ lut, const u32 order,
                                         u32& outx, u32& outy, u32& outz)
{
    const u32 leafLevel = levels - 2;
    uint2 stack[kMaxLevels];  // .x = handle, .y = childMask | visitMask << 8
    u32 px = 0, py = 0, pz = 0, level = 0, pending = 0;
    uint2 leaf = make_uint2(0, 0);

    u32 handle = dag.root();
    u32 cm = dag.header(handle) & 0xFF;
    u32 vm = cm & intersection_mask<true>(0, levels, 0, 0, 0, ray);

    for (;;) {
        if (vm == 0) {
            if (pending == 0) { outx = outy = outz = 0; return false; }
            const u32 nl = 31 - __clz(pending);
            pending ^= 1u << nl;
            const uint2 e = stack[nl];
            handle = e.x; cm = e.y & 0xFF; vm = e.y >> 8;
            const u32 up = level - nl;
            px >>= up; py >>= up; pz >>= up;
            level = nl;
        }
        const u32 child = ORDERED ? u32(lut[(order << 8) | vm]) : (31 - __clz(vm));
        vm &= ~(1u << child);
        if (vm) { stack[level] = make_uint2(handle, cm | (vm << 8)); pending |= 1u << level; }
        px = (px << 1) | (child >> 2); py = (py << 1) | ((child >> 1) & 1); pz = (pz << 1) | (child & 1);
        ++level;
        if (level == levels) { outx = px; outy = py; outz = pz; return true; }

        if (level < leafLevel) {
            handle = dag.child(handle, __popc(cm & ((1u << child) - 1u)) + 1);
            cm = dag.header(handle) & 0xFF;
        } else if (level == leafLevel) {
            leaf = dag.leaf(dag.child(handle, __popc(cm & ((1u << child) - 1u)) + 1));
            cm = first_child_mask(leaf);
        } else {
            cm = second_child_mask(leaf, child);
        }
        vm = cm & intersection_mask<false>(level, levels, px, py, pz, ray);
    }
}

struct CameraParams { double cam[3], rayMin[3], ddx[3], ddy[3]; };  // TracePathsParams, tracer.h:81-91

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
