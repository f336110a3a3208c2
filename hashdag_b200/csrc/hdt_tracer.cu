// libhashdag_b200.so: kernels + C ABI of the B200-native DAG tracer (include/hashdag_b200.h).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
//        -Xcompiler -fPIC -shared   (see __graft_entry__.build()).
#include "../../include/hashdag_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <vector>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "hdt_beam.cuh"
#include "hdt_color_leaf.cuh"
#include "hdt_colors.cuh"
#include "hdt_device.cuh"
#include "hdt_exchange.cuh"
#include "hdt_hash_table.cuh"
#include "hdt_region.cuh"
#include "hdt_resolve.cuh"

using namespace hdt;

static_assert(sizeof(hdt_basic_dag) == 16, "BasicDAG layout");
static_assert(sizeof(hdt_hash_dag) == 32, "HashDAG layout");
static_assert(sizeof(hdt_resolved_hash_dag) == 48, "hdt_resolved_hash_dag layout");
static_assert(sizeof(hdt_color_leaf) == 104, "CompressedColorLeaf layout");
static_assert(sizeof(hdt_basic_compressed_colors) == 128, "BasicDAGCompressedColors layout");
static_assert(sizeof(hdt_basic_uncompressed_colors) == 40, "BasicDAGUncompressedColors layout");
static_assert(sizeof(hdt_basic_color_errors) == 288, "BasicDAGColorErrors layout");
static_assert(sizeof(hdt_hash_colors) == 248, "HashDAGColors layout");
static_assert(sizeof(hdt_tool_info) == 44, "ToolInfo layout");
static_assert(sizeof(hdt_color_op) == 32, "hdt_color_op layout");

namespace {

// A warp is an 8x4-pixel patch (one beam of hdt_beam.cuh); a CTA is kBlockW x kBlockH pixels.
#ifndef HDT_BLOCK_W
#define HDT_BLOCK_W 16
#endif
#ifndef HDT_BLOCK_H
#define HDT_BLOCK_H 8
#endif
// Resident CTAs per SM the register allocation is held to (A/B on B200, profiles/r1_ab_occupancy.md): the per-ray
// traversal kernels at 12 (40 / 37 registers instead of 48 / 44, no spills: paths -2 %), the colour walk -- bound by the
// latency of dependent gathers, not by issue slots -- at 16 (32 registers, 128 B of spills, all 64 warps resident: -14 %).
#ifndef HDT_MIN_BLOCKS
#define HDT_MIN_BLOCKS 12
#endif
#ifndef HDT_MIN_BLOCKS_SHADOWS
#define HDT_MIN_BLOCKS_SHADOWS HDT_MIN_BLOCKS
#endif
#ifndef HDT_MIN_BLOCKS_COLORS
#define HDT_MIN_BLOCKS_COLORS 16
#endif
#ifndef HDT_MIN_BLOCKS_COLORS_RECORDED
#define HDT_MIN_BLOCKS_COLORS_RECORDED 16
#endif
constexpr u32 kBlockW = HDT_BLOCK_W, kBlockH = HDT_BLOCK_H, kWarpsX = kBlockW / 8;
constexpr u32 kBlockThreads = kBlockW * kBlockH;
static_assert(kBlockW % 8 == 0 && kBlockH % 4 == 0 && kBlockW <= 32 && kBlockH <= 32 && kBlockThreads <= 1024, "CTA shape");

// Which pixel does lane `lane` of warp `warp` of CTA `block` own?  CTAs are numbered tile-major over
// the screen tiles this rank owns; a warp covers 8x4 pixels (one beam of hdt_beam.cuh).
__device__ __forceinline__ bool warp_pixel(const PixelMap& m, u32 block, u32 warp, u32 lane, u32& x, u32& y)
{
    const u32 T = 1u << m.tileLog2, blocksX = T / kBlockW, blocksPerTile = blocksX * (T / kBlockH);
    // Units are launched in storage order, one CTA each, placed by the hardware scheduler.  Measured and rejected on B200
    // (profiles/r2_ab.md): spreading the launch order over the screen (slot * k mod nOwned, to keep the horizon's expensive
    // tiles out of the last wave): +10 % on both traversal kernels -- the CTAs resident together lose the cache lines they
    // share; resident CTAs fetching units from per-SM runs of neighbouring tiles (SM-affine persistent kernel): +27 %.
    const u32 slot = block / blocksPerTile, b = block % blocksPerTile;
    const u32 t = m.rank + slot * m.world;
    const u32 tx = t % m.tilesX, ty = t / m.tilesX;
    x = tx * T + (b % blocksX) * kBlockW + (warp % kWarpsX) * 8 + (lane & 7);
    y = ty * T + (b / blocksX) * kBlockH + (warp / kWarpsX) * 4 + (lane >> 3);
    return x < m.width && y < m.height;
}
__device__ __forceinline__ bool thread_pixel(const PixelMap& m, u32& x, u32& y)
{
    return warp_pixel(m, blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, x, y);
}

// Float ray of pixel (x, cameraRow): tracer.cu:157-165.
__device__ __forceinline__ void primary_ray(const CameraParams& cam, u32 x, u32 cameraRow, Ray& ray)
{
    double ddx, ddy, ddz;
    primary_direction(cam, x, cameraRow, ddx, ddy, ddz);
    ray.ox = cam.camf[0]; ray.oy = cam.camf[1]; ray.oz = cam.camf[2];
    ray.dx = __double2float_rn(ddx); ray.dy = __double2float_rn(ddy); ray.dz = __double2float_rn(ddz);
    ray.ix = __frcp_rn(ray.dx); ray.iy = __frcp_rn(ray.dy); ray.iz = __frcp_rn(ray.dz);
}

// ---------------------------------------------------------------------------------------------
// trace_paths (tracer.cu:145-252).  Output pixel row r holds camera row height-1-r (:251).
// Three kernels: ray setup (one thread per pixel), beam pre-pass (one thread per 8x4 tile, hdt_beam.cuh,
// concurrent with) the per-ray traversal.
// ---------------------------------------------------------------------------------------------
#ifdef HDT_BEAM_DEBUG
__device__ u32 g_beamDebug[8];   // warps of the per-ray kernels by the beam status they saw: [0..3] paths, [4..7] shadows
#endif

// A frame of rays as three planes of floats, indexed like the paths buffer.
struct RayPlanes {
    float* __restrict__ base; u64 n;
    __device__ __forceinline__ void store(u64 i, float a, float b, float c) const { base[i] = a; base[n + i] = b; base[2 * n + i] = c; }
    __device__ __forceinline__ void load(u64 i, float& a, float& b, float& c) const { a = base[i]; b = base[n + i]; c = base[2 * n + i]; }
};

__device__ __forceinline__ u32 ray_order(const Ray& r) { return (r.dx < 0.f ? 4u : 0u) + (r.dy < 0.f ? 2u : 0u) + (r.dz < 0.f ? 1u : 0u); }

// Ray setup: the double-precision normalisation of tracer.cu:158 once per pixel; the float direction
// goes to `dirs` for the per-ray kernel, its range over each 8x4 tile to `seeds` for the beam kernel.
__global__ void __launch_bounds__(kBlockThreads) setup_paths_kernel(const CameraParams cam, const PixelMap map, const RayPlanes dirs,
                                                                    BeamSeed* __restrict__ seeds)
{
    u32 x, y;
    const bool active = thread_pixel(map, x, y);
    Ray ray;
    primary_ray(cam, active ? x : 0, active ? map.height - 1 - y : 0, ray);
    if (active) dirs.store(map.index(x, y), ray.dx, ray.dy, ray.dz);
    const float d[3] = { ray.dx, ray.dy, ray.dz };
    write_seed(seeds + (blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5)), d, active, ray_is_tame(ray), ray_order(ray));
}

// Beam pre-pass: one thread per tile walks the DFS with interval masks.
template <class DAG>
__global__ void __launch_bounds__(32) beam_paths_kernel(const CameraParams cam, const DAG dag, const u32 levels, const BeamSeed* __restrict__ seeds,
                                                        BeamState* __restrict__ beams, const u32 nBeams, const u32 maxVisits, const u32 tag,
                                                        const TraverseTables* __restrict__ tables)
{
    __shared__ TraverseTables tab;
    load_tables(tab, tables);
    __syncthreads();
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nBeams) return;
    BeamState* out = beams + g;
    const uint4 s0 = __ldg(reinterpret_cast<const uint4*>(seeds + g)), s1 = __ldg(reinterpret_cast<const uint4*>(seeds + g) + 1);
    BeamRays br;
    br.d[0] = { __uint_as_float(s0.x), __uint_as_float(s1.x) };
    br.d[1] = { __uint_as_float(s0.y), __uint_as_float(s1.y) };
    br.d[2] = { __uint_as_float(s0.z), __uint_as_float(s1.z) };
#pragma unroll
    for (int k = 0; k < 3; ++k) br.o[k] = { cam.camf[k], cam.camf[k] };
    if (!s0.w || !finish_beam(br)) { store_release(&out->status, (tag << 2) | kBeamNone); return; }
    beam_traverse<DAG, true>(dag, levels, br, tab, s1.w & 7u, maxVisits, tag, out);
}

// Where trace_paths leaves the ancestor records of hit pixels for trace_colors (two planes indexed like the paths buffer;
// null = not wanted).
struct AncestorPlanes { uint4* __restrict__ a; uint4* __restrict__ b; };

template <class DAG>
__global__ void __launch_bounds__(kBlockThreads, HDT_MIN_BLOCKS) trace_paths_kernel(const CameraParams cam, const DAG dag, const u32 levels, const PixelMap map,
                                                                    const RayPlanes dirs, uint4* __restrict__ paths,
                                                                    const TraverseTables* __restrict__ tables, const BeamState* __restrict__ beams,
                                                                    const u32 tag, const AncestorPlanes anc)
{
    __shared__ TraverseTables tab;
    load_tables(tab, tables);
    const BeamState* bs = beams ? beams + (blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5)) : nullptr;
    const u32 status = beam_status(bs, tag);   // whole warp, before anybody leaves
#ifdef HDT_BEAM_DEBUG
    if (beams && (threadIdx.x & 31) == 0) atomicAdd(&g_beamDebug[status], 1u);
#endif
    __syncthreads();
    u32 x, y;
    if (!thread_pixel(map, x, y)) return;
    const u64 idx = map.index(x, y);

    u32 px, py, pz;
    if (status == kBeamHit) {
        px = __ldcg(&bs->level); py = __ldcg(&bs->pending); pz = __ldcg(&bs->handle);
        if (anc.a && (px | py | pz)) {
            const AncestorRecord r = ancestor_words(levels, px, py, pz, __ldcg(&bs->leaf), [&](u32 d) { return __ldcg(&bs->stack[d]); });
            anc.a[idx] = r.a; anc.b[idx] = r.b;
        }
    }
    else if (status == kBeamMiss) px = py = pz = 0;
    else {
        Ray ray;
        ray.ox = cam.camf[0]; ray.oy = cam.camf[1]; ray.oz = cam.camf[2];
        dirs.load(idx, ray.dx, ray.dy, ray.dz);
        ray.ix = __frcp_rn(ray.dx); ray.iy = __frcp_rn(ray.dy); ray.iz = __frcp_rn(ray.dz);
        const u32 block = table_block(tab, ray_order(ray));
        Walker<DAG> w;
        WalkStack stack;
        bool hit;
        if (status == kBeamResume) {
            resume_from<DAG, true>(w, stack, ray, bs);
            hit = walk<DAG, true, true>(w, stack, dag, levels, ray, block, px, py, pz);
        } else if (ray_is_tame(ray)) {
            w.template start<true>(dag, levels, ray);
            hit = walk<DAG, true, true>(w, stack, dag, levels, ray, block, px, py, pz);
        } else {
            w.template start<false>(dag, levels, ray);
            hit = walk<DAG, true, false>(w, stack, dag, levels, ray, block, px, py, pz);
        }
        if (anc.a && hit && (px | py | pz)) {
            const AncestorRecord r = ancestor_words(levels, px, py, pz, w.leaf, [&](u32 d) { return stack[d]; });
            anc.a[idx] = r.a; anc.b[idx] = r.b;
        }
    }
    paths[idx] = make_uint4(px, py, pz, 0);
}

// ---------------------------------------------------------------------------------------------
// trace_colors (tracer.cu:254-451)
// ---------------------------------------------------------------------------------------------
template <class DAG>
__global__ void __launch_bounds__(kBlockThreads, HDT_MIN_BLOCKS_COLORS) trace_colors_kernel(const DAG dag, const ColorsDev colors, const u32 levels, const ColorsParams prm,
                                                                     const PixelMap map, const uint4* __restrict__ paths, u32* __restrict__ out)
{
    u32 x, y;
    if (!thread_pixel(map, x, y)) return;
    const u64 idx = map.index(x, y);
    const uint4 p = paths[idx];
    out[idx] = color_pixel(dag, colors, levels, prm, p.x, p.y, p.z);
}

// trace_colors from the ancestor records of the paths pass (color_pixel_recorded, hdt_colors.cuh).
__global__ void __launch_bounds__(kBlockThreads, HDT_MIN_BLOCKS_COLORS_RECORDED) trace_colors_recorded_kernel(const u32* __restrict__ prefix, const ColorsDev colors, const u32 levels,
                                                                     const ColorsParams prm, const PixelMap map, const uint4* __restrict__ paths,
                                                                     const uint4* __restrict__ ancA, const uint4* __restrict__ ancB, u32* __restrict__ out)
{
    u32 x, y;
    if (!thread_pixel(map, x, y)) return;
    const u64 idx = map.index(x, y);
    const uint4 p = paths[idx];
    if ((p.x | p.y | p.z) == 0) { out[idx] = sky_color(); return; }
    out[idx] = color_pixel_recorded(prefix, colors, levels, prm, p.x, p.y, p.z, __ldg(ancA + idx), __ldg(ancB + idx));
}

// ---------------------------------------------------------------------------------------------
// trace_shadows (tracer.cu:589-697): exact hit point, any-hit ray towards the sun, shade, fog.
// ---------------------------------------------------------------------------------------------
struct ShadowParams { float shadowBias, fogDensity; float sunX, sunY, sunZ; };

// Shadow ray of a hit pixel (tracer.cu:622-655): origin = exact hit point of the primary ray on the
// voxel [p, p+1] (ray_box_intersection, tracer.cu:574-587, in double) + shadowBias * sun.
__device__ __forceinline__ void shadow_ray(const CameraParams& cam, const ShadowParams& sp, const uint4 p, const double dirx, const double diry,
                                           const double dirz, Ray& ray)
{
    const double bo[3] = { double(__uint2float_rn(p.x)), double(__uint2float_rn(p.y)), double(__uint2float_rn(p.z)) };
    const double dv[3] = { dirx, diry, dirz };
    double rm[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // tracer.cu:578-581 takes the smaller of t0 = (lo - cam) / d and t1 = (lo + 1 - cam) / d.  Correctly rounded subtraction and
        // division are monotone, and the numerators differ by one while |d| <= 1, so t0 < t1 exactly when d > 0 and t1 < t0 when
        // d < 0: one division gives the minimum.  (d = 0 or NaN: both, as written there.)
        const double n0 = __dsub_rn(bo[k], cam.cam[k]), n1 = __dsub_rn(__dadd_rn(bo[k], 1.0), cam.cam[k]);
        if (dv[k] > 0.0) rm[k] = __ddiv_rn(n0, dv[k]);
        else if (dv[k] < 0.0) rm[k] = __ddiv_rn(n1, dv[k]);
        else {
            const double t0 = __ddiv_rn(n0, dv[k]), t1 = __ddiv_rn(n1, dv[k]);
            rm[k] = (t0 < t1) ? t0 : t1;
        }
    }
    const double maxmin = fmax(fmax(rm[0], rm[1]), rm[2]);
    ray.ox = __fmaf_rn(sp.shadowBias, sp.sunX, __double2float_rn(__fma_rn(dv[0], maxmin, cam.cam[0])));
    ray.oy = __fmaf_rn(sp.shadowBias, sp.sunY, __double2float_rn(__fma_rn(dv[1], maxmin, cam.cam[1])));
    ray.oz = __fmaf_rn(sp.shadowBias, sp.sunZ, __double2float_rn(__fma_rn(dv[2], maxmin, cam.cam[2])));
    ray.dx = sp.sunX; ray.dy = sp.sunY; ray.dz = sp.sunZ;
    ray.ix = __frcp_rn(ray.dx); ray.iy = __frcp_rn(ray.dy); ray.iz = __frcp_rn(ray.dz);
}

// Ray setup of trace_shadows, once per pixel: the shadow ray's origin goes to `origins` (hit pixels
// only), its range over each 8x4 tile to `seeds`.
__global__ void __launch_bounds__(kBlockThreads) setup_shadows_kernel(const CameraParams cam, const ShadowParams sp, const PixelMap map,
                                                                      const uint4* __restrict__ paths, const RayPlanes origins,
                                                                      BeamSeed* __restrict__ seeds)
{
    u32 x, y;
    bool active = thread_pixel(map, x, y);
    u64 idx = 0;
    uint4 p = make_uint4(0, 0, 0, 0);
    if (active) { idx = map.index(x, y); p = paths[idx]; }
    active = active && (p.x | p.y | p.z) != 0;
    Ray ray;
    ray.ox = ray.oy = ray.oz = 0.f; ray.dx = ray.dy = ray.dz = ray.ix = ray.iy = ray.iz = 1.f;
    if (active) {
        double dirx, diry, dirz;
        primary_direction(cam, x, map.height - 1 - y, dirx, diry, dirz);   // the row flip of tracer.cu:622 cancels the one of :251
        shadow_ray(cam, sp, p, dirx, diry, dirz, ray);
        origins.store(idx, ray.ox, ray.oy, ray.oz);
    }
    const float o[3] = { ray.ox, ray.oy, ray.oz };
    write_seed(seeds + (blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5)), o, active, ray_is_tame(ray), 0);
}

// Beam pre-pass of trace_shadows: the shadow rays of a tile share the direction; their origins span a box.
template <class DAG>
__global__ void __launch_bounds__(32) beam_shadows_kernel(const ShadowParams sp, const DAG dag, const u32 levels, const BeamSeed* __restrict__ seeds,
                                                          BeamState* __restrict__ beams, const u32 nBeams, const u32 maxVisits, const u32 tag,
                                                          const TraverseTables* __restrict__ tables)
{
    __shared__ TraverseTables tab;
    load_tables(tab, tables);
    __syncthreads();
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nBeams) return;
    BeamState* out = beams + g;
    const uint4 s0 = __ldg(reinterpret_cast<const uint4*>(seeds + g)), s1 = __ldg(reinterpret_cast<const uint4*>(seeds + g) + 1);
    BeamRays br;
    br.o[0] = { __uint_as_float(s0.x), __uint_as_float(s1.x) };
    br.o[1] = { __uint_as_float(s0.y), __uint_as_float(s1.y) };
    br.o[2] = { __uint_as_float(s0.z), __uint_as_float(s1.z) };
    br.d[0] = { sp.sunX, sp.sunX }; br.d[1] = { sp.sunY, sp.sunY }; br.d[2] = { sp.sunZ, sp.sunZ };
    if (!s0.w || !finish_beam(br)) { store_release(&out->status, (tag << 2) | kBeamNone); return; }
    beam_traverse<DAG, false>(dag, levels, br, tab, 0, maxVisits, tag, out);
}

template <class DAG>
__device__ __forceinline__ void shadow_pixel(const CameraParams& cam, const ShadowParams& sp, const DAG& dag, const u32 levels, const PixelMap& map,
                                             const uint4* __restrict__ paths, const RayPlanes& origins, u32* __restrict__ colors, const TraverseTables& tab,
                                             const BeamState* bs, const u32 status, const u32 x, const u32 y, u32* __restrict__ xFrame)
{
    const u64 idx = map.index(x, y);
    const uint4 p = paths[idx];
    auto put = [&](u32 rgba) {
        colors[idx] = rgba;
        if (xFrame) xFrame[u64(y) * map.width + x] = rgba;
    };

    // setColor (tracer.cu:604-619) + applyFog (:551-572); contraction as in the reference PTX.
    // hit == false: sky pixel, distance 1e9 along the primary direction (tracer.cu:644-648).
    const float fd = __fmul_rn(sp.fogDensity, 0.00001f);
    auto shade = [&](float lightScale, bool hit) {
        const float3 c = rgb888_to_float3(colors[idx]);
        const float lx = __fmul_rn(c.x, lightScale), ly = __fmul_rn(c.y, lightScale), lz = __fmul_rn(c.z, lightScale);
        if (fd == 0.0f) {
            // No fog: fogAmount = 1 - exp(-0) = 0 exactly, so lerp(lit, fogColor, 0) = fma(lit, 1, 0*fogColor) = lit
            // bit for bit (fogColor is finite); distance, direction, exp and pow of applyFog cannot change the
            // result and are not computed.
            put(float3_to_rgb888(lx, ly, lz));
            return;
        }
        double distance = 1e9, rdx, rdy, rdz;
        if (hit) {
            const double vx = __dsub_rn(double(__uint2float_rn(p.x)), cam.cam[0]), vy = __dsub_rn(double(__uint2float_rn(p.y)), cam.cam[1]),
                         vz = __dsub_rn(double(__uint2float_rn(p.z)), cam.cam[2]);
            distance = __dsqrt_rn(__fma_rn(vz, vz, __fma_rn(vx, vx, __dmul_rn(vy, vy))));
            rdx = __ddiv_rn(vx, distance); rdy = __ddiv_rn(vy, distance); rdz = __ddiv_rn(vz, distance);
        } else {
            primary_direction(cam, x, map.height - 1 - y, rdx, rdy, rdz);
        }
        const double fogAmount = __dsub_rn(1.0, exp(__dmul_rn(-distance, double(fd))));
        const double dotp = __fma_rn(rdz, double(sp.sunZ), __fma_rn(rdx, double(sp.sunX), __dmul_rn(rdy, double(sp.sunY))));
        const double sunAmount = __dmul_rn(double(1.01f), fmax(dotp, 0.0));
        const float pw = __double2float_rn(pow(sunAmount, 30.0)), q = __fsub_rn(1.f, pw);
        const float fx = __fmaf_rn(q, __fdiv_rn(187.f, 255.f), pw), fy = __fmaf_rn(q, __fdiv_rn(242.f, 255.f), pw), fz = __fmaf_rn(q, __fdiv_rn(250.f, 255.f), pw);
        const float g = clampf(__double2float_rn(fogAmount), 0.f, 1.f), h = __fsub_rn(1.f, g);
        put(float3_to_rgb888(__fmaf_rn(lx, h, __fmul_rn(g, fx)), __fmaf_rn(ly, h, __fmul_rn(g, fy)), __fmaf_rn(lz, h, __fmul_rn(g, fz))));
    };
    if ((p.x | p.y | p.z) == 0) { shade(1.0f, false); return; }

    bool shadowed;
    if (status == kBeamHit) shadowed = true;
    else if (status == kBeamMiss) shadowed = false;
    else {
        Ray ray;
        origins.load(idx, ray.ox, ray.oy, ray.oz);
        ray.dx = sp.sunX; ray.dy = sp.sunY; ray.dz = sp.sunZ;
        ray.ix = __frcp_rn(ray.dx); ray.iy = __frcp_rn(ray.dy); ray.iz = __frcp_rn(ray.dz);
        u32 hx, hy, hz;
        const u32 block = table_block(tab, 0);
        Walker<DAG> w;
        WalkStack stack;
        if (status == kBeamResume) {
            resume_from<DAG, false>(w, stack, ray, bs);
            shadowed = walk<DAG, false, true>(w, stack, dag, levels, ray, block, hx, hy, hz);
        } else if (ray_is_tame(ray)) {
            w.template start<true>(dag, levels, ray);
            shadowed = walk<DAG, false, true>(w, stack, dag, levels, ray, block, hx, hy, hz);
        } else {
            w.template start<false>(dag, levels, ray);
            shadowed = walk<DAG, false, false>(w, stack, dag, levels, ray, block, hx, hy, hz);
        }
    }
    shade(shadowed ? 0.5f : 1.0f, true);
}

template <class DAG>
__global__ void __launch_bounds__(kBlockThreads, HDT_MIN_BLOCKS_SHADOWS) trace_shadows_kernel(const CameraParams cam, const ShadowParams sp, const DAG dag, const u32 levels,
                                                                      const PixelMap map, const uint4* __restrict__ paths, const RayPlanes origins,
                                                                      u32* __restrict__ colors, const TraverseTables* __restrict__ tables,
                                                                      const BeamState* __restrict__ beams, const u32 tag, const ExchangeOut xo)
{
    __shared__ TraverseTables tab;
    load_tables(tab, tables);
    const BeamState* bs = beams ? beams + (blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5)) : nullptr;
    const u32 status = beam_status(bs, tag);   // whole warp, before anybody leaves
#ifdef HDT_BEAM_DEBUG
    if (beams && (threadIdx.x & 31) == 0) atomicAdd(&g_beamDebug[4 + status], 1u);
#endif
    __syncthreads();
    u32 x, y;
    if (thread_pixel(map, x, y)) shadow_pixel<DAG>(cam, sp, dag, levels, map, paths, origins, colors, tab, bs, status, x, y, xo.frame);
    // Fused framebuffer exchange (hdt_exchange.cuh): the pixel above also went into rank 0's row-major frame; the last CTA
    // of the launch tells rank 0 that this rank's tiles are complete.
    if (xo.frame) exchange_signal_last_cta(xo.ctasDone, xo.arrived, xo.seq);
}

// ---------------------------------------------------------------------------------------------
// Small plumbing kernels
// ---------------------------------------------------------------------------------------------
// Scatter compact per-rank tile buffers (world of them, each maxTiles tiles) into a row-major frame.
template <class T>
__global__ void assemble_kernel(const T* __restrict__ gathered, T* __restrict__ frame, PixelMap map, u64 tilePixelsPerRank)
{
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= map.width || y >= map.height) return;
    const u32 t = (y >> map.tileLog2) * map.tilesX + (x >> map.tileLog2);
    PixelMap m = map;
    m.world = map.world; m.rank = t % map.world;
    const u64 src = u64(t % map.world) * tilePixelsPerRank + (map.world == 1 ? (u64(y) * map.width + x) : m.index(x, y));
    frame[u64(y) * map.width + x] = gathered[src];
}

// Copy this rank's compact buffer into a row-major frame, leaving other ranks' pixels untouched.
template <class T>
__global__ void untile_own_kernel(const T* __restrict__ compact, T* __restrict__ frame, PixelMap map)
{
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= map.width || y >= map.height) return;
    const u32 t = (y >> map.tileLog2) * map.tilesX + (x >> map.tileLog2);
    if (t % map.world != map.rank) return;
    frame[u64(y) * map.width + x] = compact[map.index(x, y)];
}

__global__ void count_hits_kernel(const uint4* __restrict__ paths, u64 n, unsigned long long* __restrict__ out)
{
    u32 local = 0;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x) {
        const uint4 p = paths[i];
        local += (p.x | p.y | p.z) != 0;
    }
    local = __reduce_add_sync(0xFFFFFFFFu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

// Histogram of the last beam pre-pass: out[status] += 1, out[4] += level at which resuming tiles hand over.
__global__ void beam_stats_kernel(const BeamState* __restrict__ beams, u32 n, unsigned long long* __restrict__ out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 st = beams[i].status & 3u;
    atomicAdd(out + st, 1ull);
    if (st == kBeamResume) atomicAdd(out + 4, (unsigned long long)beams[i].level);
}

__global__ void apply_ranges_kernel(u32* __restrict__ dst, const u32* __restrict__ payload, const hdt_range* __restrict__ ranges, u32 nRanges)
{
    // one CTA per range, coalesced copy
    for (u32 r = blockIdx.x; r < nRanges; r += gridDim.x) {
        const hdt_range rg = ranges[r];
        for (u64 i = threadIdx.x; i < rg.n_words; i += blockDim.x) dst[rg.dst_word + i] = payload[rg.src_word + i];
    }
}

thread_local std::string g_lastError;

int fail(int code, const char* what)
{
    g_lastError = what;
    return code;
}
int cuda_fail(cudaError_t e, const char* where)
{
    g_lastError = std::string(where) + ": " + cudaGetErrorString(e);
    return HDT_ERR_CUDA + int(e);
}
#define HDT_CUDA(call)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);   \
    } while (0)

}  // namespace

struct hdt_ctx {
    int device = 0;
    u32 levels = 0;
    PixelMap map{};
    u32 nOwnedTiles = 0, maxTilesPerRank = 0;
    cudaStream_t stream = nullptr;      // where work is enqueued (ownStream unless hdt_set_stream)
    cudaStream_t ownStream = nullptr;
    cudaEvent_t ev[4] = {};
    cudaEvent_t timer[2] = {};
    unsigned long long* hitCounter = nullptr;  // device
    uint4* paths = nullptr;      // compact (tiled) when world > 1, row-major otherwise
    u32* colors = nullptr;
    uint4* framePaths = nullptr; // row-major staging for read-back / assembly when world > 1
    u32* frameColors = nullptr;
    u32* pathCache = nullptr;    // pinned, 4 words
    u64 launches = 0;
    u64 recordedColorPasses = 0;   // colour passes that read the ancestor records instead of walking the DAG
    TraverseTables* tables = nullptr;   // device copy of the traversal tables
    // Beam pre-pass (hdt_beam.cuh): per pass (0 = paths, 1 = shadows) one BeamState + BeamSeed per 8x4-pixel
    // tile (= warp of the per-ray kernels) and three float planes of per-pixel ray data (directions / origins).
    BeamState* beams[2] = {};
    BeamSeed* seeds[2] = {};
    float* rays[2] = {};
    cudaStream_t side = nullptr;        // beam kernels run here, concurrently with the per-ray kernels
    cudaEvent_t fork[2] = {}, setupDone[2] = {}, join[2] = {}, traceDone[2] = {};   // per pass; traceDone[0] also gates the prefetch
    bool useBeams = true;
    bool l2Persist = false;             // HDT_OPT_L2_PERSIST: page table of a plain HashDAG pinned in L2 (access policy window); measured slower, off
    const void* l2Window = nullptr;     // what the streams' access policy window currently covers
    size_t l2WindowBytes = 0;
    bool beamPrefetch = false;
    bool beamSerial = false;            // diagnostics: per-ray kernels wait for the beam kernel
    u32 beamMaxVisits = 32;
    u32 beamTag = 0;                    // bumped per beam launch; per-ray kernels ignore states of other launches
    int lastBeamPass = 0;
    // framebuffer exchange over peer memory (hdt_exchange.cuh)
    u32* xBlock = nullptr;              // [frame W*H][ExchangeCounters]: own allocation (root) or a mapping of the root's
    bool xOwned = false, xIpc = false;
    void* xHostBlock = nullptr;         // hdt_exchange_attach_host: the pinned host block xBlock aliases (unregistered at destroy if this context registered it)
    bool xHostRegistered = false;
    u32 xSeq = 0;                       // frames exchanged on this context
    bool xFused = false;                // HDT_OPT_EXCHANGE_FUSED: shadow passes store into rank 0's frame themselves
    u32 xFusedSeq = 0;                  // sequence number the last fused shadow pass stored for
    u32* xTimedOutDev = nullptr;        // device alias of xTimedOut
    u32* xCtasDone = nullptr;           // device, 2 words: the scatter kernel's last-CTA counter, the abort flag of a wait that gave up
    unsigned long long xWaitCycles = 40000000000ull;   // ~20 s of SM clock; 0 = wait for ever (HDT_OPT_EXCHANGE_TIMEOUT_MS)
    u32* xTimedOut = nullptr;           // pinned + mapped: raised by a wait kernel that gave up
    char* stagingHost = nullptr;        // hdt_apply_ranges_host: pinned + device staging, bump-allocated, reset when full
    char* stagingDev = nullptr;
    size_t stagingCap = 0, stagingUsed = 0;
    // Ancestor records of the last paths frame (AncestorRecord, hdt_device.cuh), written for HDT_DAG_HASH_RESOLVED DAGs with a
    // prefix pool; ancFor* say which DAG that frame was traced in (trace_colors takes the short route only for the same one).
    uint4* anc[2] = {};
    bool useRecorded = true;            // HDT_OPT_COLORS_RECORDED
    bool ancValid = false;
    const u32* ancForPool = nullptr; const u32* ancForPrefix = nullptr; u32 ancForRoot = 0;
    u32* physToVirt = nullptr;          // hdt_hash_dag_resolve: physical page -> virtual page (grow-only)
    size_t physToVirtPages = 0;
    void* foaScratch = nullptr;         // hdt_find_or_add: keys, flags, staging (grow-only)
    size_t foaScratchBytes = 0;
    void* comm = nullptr;               // ncclComm_t (hdt_comm_init, hdt_multi.cuh); null while world == 1
    u32 commRank = 0, commWorld = 1;
    void* rebuildScratch = nullptr;     // hdt_rebuild_color_leaf: ops, per-macro-block sums (grow-only)
    size_t rebuildScratchBytes = 0;

    RayPlanes ray_planes(int pass) const { return RayPlanes{ rays[pass], buffer_pixels() }; }

    u32 n_beams() const { return grid_blocks() * (kBlockThreads / 32); }

    u64 buffer_pixels() const { return map.world == 1 ? u64(map.width) * map.height : (u64(maxTilesPerRank) << (2 * map.tileLog2)); }
    u32 grid_blocks() const
    {
        const u32 T = 1u << map.tileLog2;
        return nOwnedTiles * (T / kBlockW) * (T / kBlockH);
    }
};

namespace {

int configure(hdt_ctx* c, u32 rank, u32 world, u32 tileLog2)
{
    if (world == 0 || rank >= world || tileLog2 < 5 || tileLog2 > 10) return fail(HDT_ERR_ARG, "bad partition");
    HDT_CUDA(cudaSetDevice(c->device));
    if (c->side) HDT_CUDA(cudaStreamSynchronize(c->side));
    if (c->paths) { cudaFree(c->paths); c->paths = nullptr; }
    if (c->colors) { cudaFree(c->colors); c->colors = nullptr; }
    if (c->framePaths) { cudaFree(c->framePaths); c->framePaths = nullptr; }
    if (c->frameColors) { cudaFree(c->frameColors); c->frameColors = nullptr; }
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->beams[i]); c->beams[i] = nullptr;
        cudaFree(c->seeds[i]); c->seeds[i] = nullptr;
        cudaFree(c->rays[i]); c->rays[i] = nullptr;
        cudaFree(c->anc[i]); c->anc[i] = nullptr;
    }
    c->ancValid = false;
    PixelMap& m = c->map;
    m.tileLog2 = tileLog2; m.world = world; m.rank = rank;
    const u32 T = 1u << tileLog2;
    m.tilesX = (m.width + T - 1) / T; m.tilesY = (m.height + T - 1) / T;
    const u32 nTiles = m.tilesX * m.tilesY;
    c->nOwnedTiles = (nTiles > rank) ? (nTiles - rank + world - 1) / world : 0;
    c->maxTilesPerRank = (nTiles + world - 1) / world;
    const u64 n = c->buffer_pixels();
    HDT_CUDA(cudaMalloc(&c->paths, n * sizeof(uint4)));
    HDT_CUDA(cudaMalloc(&c->colors, n * sizeof(u32)));
    for (int i = 0; i < 2 && c->n_beams(); ++i) {
        HDT_CUDA(cudaMalloc(&c->beams[i], size_t(c->n_beams()) * sizeof(BeamState)));
        HDT_CUDA(cudaMemsetAsync(c->beams[i], 0xFF, size_t(c->n_beams()) * sizeof(BeamState), c->stream));   // no launch has tag 2^30-1
        HDT_CUDA(cudaMalloc(&c->seeds[i], size_t(c->n_beams()) * sizeof(BeamSeed)));
        HDT_CUDA(cudaMalloc(&c->rays[i], n * 3 * sizeof(float)));
    }
    // ancestor records: DAGs with HashDAGColors (levels - 2 > 10 colour-tree levels) of at most 18 levels
    if (c->levels >= kColorTreeDepth + 3 && c->levels <= kColorTreeDepth + 2 + kMaxAncestorWords && n)
        for (int i = 0; i < 2; ++i) HDT_CUDA(cudaMalloc(&c->anc[i], n * sizeof(uint4)));
    HDT_CUDA(cudaMemsetAsync(c->paths, 0, n * sizeof(uint4), c->stream));
    HDT_CUDA(cudaMemsetAsync(c->colors, 0, n * sizeof(u32), c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    return HDT_OK;
}

struct DagArg {
    int kind; BasicDagDev basic; HashDagDev hash; HashDagResolvedDev resolved;
    u32 pageTableSize = 0;   // HashDAG kinds: entries of the page table
    // a resolved HashDAG with a prefix pool, seen through the accessor that reads leaf masks from it (per-ray traversal kernels)
#ifdef HDT_NO_PREFIX_LEAF_MASK   // A/B switch: traverse through the plain resolved accessor even when a prefix pool is there
    bool has_prefix() const { return false; }
#else
    bool has_prefix() const { return kind == HDT_DAG_HASH_RESOLVED && resolved.prefix != nullptr; }
#endif
    HashDagPrefixDev prefixed() const { return HashDagPrefixDev{ resolved.pool, resolved.vpool, resolved.pageTable, resolved.prefix, resolved.firstNodeIndex }; }
};

int parse_dag(int kind, const void* pod, size_t size, DagArg& out)
{
    if (!pod) return fail(HDT_ERR_ARG, "null DAG");
    out.kind = kind;
    if (kind == HDT_DAG_BASIC) {
        if (size != sizeof(hdt_basic_dag)) return fail(HDT_ERR_POD_SIZE, "BasicDAG: expected 16 bytes");
        hdt_basic_dag d; memcpy(&d, pod, sizeof(d));
        if (!d.data.data) return fail(HDT_ERR_ARG, "BasicDAG: null data");
        out.basic.data = static_cast<const u32*>(d.data.data);
        return HDT_OK;
    }
    if (kind == HDT_DAG_HASH) {
        if (size != sizeof(hdt_hash_dag)) return fail(HDT_ERR_POD_SIZE, "HashDAG: expected 32 bytes");
        hdt_hash_dag d; memcpy(&d, pod, sizeof(d));
        if (!d.pool || !d.page_table) return fail(HDT_ERR_ARG, "HashDAG: null pool / page table");
        if (u64(d.pool_top) * kPageWords > (u64(1) << 32)) return fail(HDT_ERR_ARG, "HashDAG: pool beyond 2^32 words");
        out.hash.pool = d.pool; out.hash.pageTable = d.page_table; out.hash.firstNodeIndex = d.first_node_index;
        out.pageTableSize = d.page_table_size;
        return HDT_OK;
    }
    if (kind == HDT_DAG_HASH_RESOLVED) {
        if (size != sizeof(hdt_resolved_hash_dag)) return fail(HDT_ERR_POD_SIZE, "hdt_resolved_hash_dag: expected 48 bytes");
        hdt_resolved_hash_dag d; memcpy(&d, pod, sizeof(d));
        if (!d.dag.pool || !d.dag.page_table || !d.resolved_pool) return fail(HDT_ERR_ARG, "resolved HashDAG: null pool / page table / resolved pool");
        if (u64(d.dag.pool_top) * kPageWords > (u64(1) << 32)) return fail(HDT_ERR_ARG, "HashDAG: pool beyond 2^32 words");
        if ((d.dag.first_node_index >> 9) >= d.dag.page_table_size) return fail(HDT_ERR_ARG, "resolved HashDAG: root outside the page table");
        out.resolved.pool = d.resolved_pool; out.resolved.vpool = d.dag.pool; out.resolved.pageTable = d.dag.page_table;
        out.resolved.prefix = d.prefix_pool;
        out.resolved.firstNodeIndex = d.dag.first_node_index;
        return HDT_OK;
    }
    return fail(HDT_ERR_ARG, "unknown DAG kind");
}

ColorLeafDev leaf_dev(const hdt_color_leaf& l)
{
    ColorLeafDev d;
    d.offset = l.offset;
    d.weights = static_cast<const u32*>(l.weights_gpu.data); d.nWeights = l.weights_gpu.size;
    d.blocks = static_cast<const u64*>(l.blocks_gpu.data); d.nBlocks = l.blocks_gpu.size;
    d.macroBlocks = static_cast<const u64*>(l.macro_blocks_gpu.data); d.nMacroWords = l.macro_blocks_gpu.size;
    return d;
}

int parse_colors(int kind, const void* pod, size_t size, ColorsDev& out)
{
    if (!pod) return fail(HDT_ERR_ARG, "null colours");
    memset(&out, 0, sizeof(out));
    out.kind = kind;
    switch (kind) {
    case HDT_COLORS_COMPRESSED: {
        if (size != sizeof(hdt_basic_compressed_colors)) return fail(HDT_ERR_POD_SIZE, "BasicDAGCompressedColors: expected 128 bytes");
        hdt_basic_compressed_colors c; memcpy(&c, pod, sizeof(c));
        out.topLevels = c.base.top_levels; out.enclosedLeaves = static_cast<const u64*>(c.base.enclosed_leaves.data);
        out.leaf = leaf_dev(c.leaf);
        return HDT_OK;
    }
    case HDT_COLORS_UNCOMPRESSED: {
        if (size != sizeof(hdt_basic_uncompressed_colors)) return fail(HDT_ERR_POD_SIZE, "BasicDAGUncompressedColors: expected 40 bytes");
        hdt_basic_uncompressed_colors c; memcpy(&c, pod, sizeof(c));
        out.topLevels = c.base.top_levels; out.enclosedLeaves = static_cast<const u64*>(c.base.enclosed_leaves.data);
        out.uncompressed = static_cast<const u32*>(c.colors.data); out.nUncompressed = c.colors.size;
        return HDT_OK;
    }
    case HDT_COLORS_ERRORS: {
        if (size != sizeof(hdt_basic_color_errors)) return fail(HDT_ERR_POD_SIZE, "BasicDAGColorErrors: expected 288 bytes");
        hdt_basic_color_errors c; memcpy(&c, pod, sizeof(c));
        out.topLevels = c.compressed.base.top_levels; out.enclosedLeaves = static_cast<const u64*>(c.compressed.base.enclosed_leaves.data);
        out.leaf = leaf_dev(c.compressed.leaf);
        out.uncompressed = static_cast<const u32*>(c.uncompressed.colors.data); out.nUncompressed = c.uncompressed.colors.size;
        return HDT_OK;
    }
    case HDT_COLORS_HASH: {
        if (size != sizeof(hdt_hash_colors)) return fail(HDT_ERR_POD_SIZE, "HashDAGColors: expected 248 bytes");
        hdt_hash_colors c; memcpy(&c, pod, sizeof(c));
        if (!c.nodes_gpu.data) return fail(HDT_ERR_ARG, "HashDAGColors: null nodes");
        out.nodes = static_cast<const u32*>(c.nodes_gpu.data);
        out.leaves = static_cast<const ColorLeafPod*>(c.leaves_gpu.data);
        out.offsets = static_cast<const u64*>(c.offsets_gpu.data);
        out.leaf = leaf_dev(c.main_leaf);
        return HDT_OK;
    }
    }
    return fail(HDT_ERR_ARG, "unknown colours kind");
}

CameraParams make_cam(const double cam[3], const double rmin[3], const double ddx[3], const double ddy[3])
{
    CameraParams p;
    for (int k = 0; k < 3; ++k) { p.cam[k] = cam[k]; p.rayMin[k] = rmin[k]; p.ddx[k] = ddx[k]; p.ddy[k] = ddy[k]; p.camf[k] = float(cam[k]); }
    p.pad = 0.f;
    return p;
}

ShadowParams make_shadow(float bias, float fog)
{
    // sun_direction(), tracer.cu:546-549: normalize(float3(0.3, 1, 0.5)) in IEEE single without
    // contraction -- the same values the reference compiler folded into its SASS immediates.
    volatile float x = 0.3f, y = 1.f, z = 0.5f;
    volatile float xx = x * x, yy = y * y, zz = z * z;
    volatile float s1 = xx + yy;
    volatile float s2 = s1 + zz;
    const float len = sqrtf(s2);
    volatile float r = 1.0f / len;
    ShadowParams sp;
    sp.shadowBias = bias; sp.fogDensity = fog;
    volatile float sx = r * x, sy = r * y, sz = r * z;
    sp.sunX = sx; sp.sunY = sy; sp.sunZ = sz;
    return sp;
}

// Launch tags run through 0 .. 2^30-2; the state buffers are initialised with 2^30-1.
u32 next_beam_tag(hdt_ctx* c) { return c->beamTag = (c->beamTag + 1) % 0x3FFFFFFFu; }

// A plain HashDAG (HDT_DAG_HASH, what a drop-in caller passes) is traced through its page table: one 4-byte entry per node
// visited, scattered over 16 MiB at depth 17 (hash_table.h:156-173).  With HDT_OPT_L2_PERSIST the table is declared persisting
// in L2 for the tracer's streams (cudaAccessPolicyWindow), so the streaming node and leaf loads of a frame cannot evict it; the
// window is (re)set only when the table's address or size changes.  MEASURED on B200 (profiles/r2_ab.md): +4.6 % on the paths
// pass, +4.9 % shadows -- the 126 MB L2 keeps the table resident anyway and the set-aside only shrinks what the nodes can use --
// so the option is OFF by default.  (The resolved pool does not read the page table while traversing.)
int pin_page_table(hdt_ctx* c, const DagArg& d, uint32_t pageTableEntries)
{
    const void* base = (d.kind == HDT_DAG_HASH && c->l2Persist) ? static_cast<const void*>(d.hash.pageTable) : nullptr;
    const size_t bytes = base ? size_t(pageTableEntries) * 4 : 0;
    if (base == c->l2Window && bytes == c->l2WindowBytes) return HDT_OK;
    if (base && !c->l2WindowBytes) {   // first use: set aside L2 for persisting lines (device-wide limit; errors here are not fatal)
        int maxPersist = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
        if (maxPersist > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(size_t(maxPersist), size_t(64) << 20));
        cudaGetLastError();
    }
    cudaStreamAttrValue attr{};
    int maxWindow = 0;
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    attr.accessPolicyWindow.num_bytes = std::min(bytes, size_t(maxWindow > 0 ? maxWindow : 0));
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = base ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (!base) attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    for (cudaStream_t st : { c->stream, c->side })
        if (st && cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) { cudaGetLastError(); break; }
    c->l2Window = base; c->l2WindowBytes = bytes;
    return HDT_OK;
}

// Every launch helper returns HDT_OK or the code of the first failing CUDA call (named in hdt_last_error()).
#define HDT_LAUNCHED(what)                                       \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);     \
        ++c->launches;                                           \
    } while (0)

int launch_paths(hdt_ctx* c, const DagArg& d, const CameraParams& cam)
{
    const dim3 grid(c->grid_blocks()), block(kBlockThreads);
    c->ancValid = false;
    if (!grid.x) return HDT_OK;
    if (int rc = pin_page_table(c, d, d.pageTableSize)) return rc;
    // Ray setup and beams on the side stream.  Normally they are ordered after everything enqueued on the
    // main stream so far.  With HDT_OPT_BEAM_PREFETCH the caller promises that the DAG is not modified by
    // work queued on the tracer's stream, and they only wait for the previous paths kernel (the last
    // reader of their buffers): enqueued right behind the previous frame, they then run beside its
    // colours / shadows kernels, and the per-ray kernel below finds every beam finished.
    HDT_CUDA(cudaEventRecord(c->fork[0], c->stream));
    HDT_CUDA(cudaStreamWaitEvent(c->side, c->beamPrefetch ? c->traceDone[0] : c->fork[0], 0));
    setup_paths_kernel<<<grid, block, 0, c->side>>>(cam, c->map, c->ray_planes(0), c->seeds[0]);
    HDT_LAUNCHED("setup_paths_kernel");
    HDT_CUDA(cudaEventRecord(c->setupDone[0], c->side));
    const BeamState* beams = nullptr;
    u32 tag = 0;
    if (c->useBeams) {
        beams = c->beams[0];
        tag = next_beam_tag(c);
        const u32 nb = c->n_beams();
        const dim3 g((nb + 31) / 32), b(32);
        if (d.kind == HDT_DAG_BASIC) beam_paths_kernel<BasicDagDev><<<g, b, 0, c->side>>>(cam, d.basic, c->levels, c->seeds[0], c->beams[0], nb, c->beamMaxVisits, tag, c->tables);
        else if (d.kind == HDT_DAG_HASH) beam_paths_kernel<HashDagDev><<<g, b, 0, c->side>>>(cam, d.hash, c->levels, c->seeds[0], c->beams[0], nb, c->beamMaxVisits, tag, c->tables);
        else beam_paths_kernel<HashDagResolvedDev><<<g, b, 0, c->side>>>(cam, d.resolved, c->levels, c->seeds[0], c->beams[0], nb, c->beamMaxVisits, tag, c->tables);
        HDT_LAUNCHED("beam_paths_kernel");
        c->lastBeamPass = 0;
    }
    HDT_CUDA(cudaEventRecord(c->join[0], c->side));
    HDT_CUDA(cudaStreamWaitEvent(c->stream, c->beamSerial ? c->join[0] : c->setupDone[0], 0));   // the directions
    // ancestor records for trace_colors: only a resolved HashDAG with a prefix pool can use them
    AncestorPlanes anc{ nullptr, nullptr };
    if (d.kind == HDT_DAG_HASH_RESOLVED && d.resolved.prefix && c->anc[0] && c->useRecorded) anc = AncestorPlanes{ c->anc[0], c->anc[1] };
    if (d.has_prefix()) trace_paths_kernel<HashDagPrefixDev><<<grid, block, 0, c->stream>>>(cam, d.prefixed(), c->levels, c->map, c->ray_planes(0), c->paths, c->tables, beams, tag, anc);
    else if (d.kind == HDT_DAG_BASIC) trace_paths_kernel<BasicDagDev><<<grid, block, 0, c->stream>>>(cam, d.basic, c->levels, c->map, c->ray_planes(0), c->paths, c->tables, beams, tag, anc);
    else if (d.kind == HDT_DAG_HASH) trace_paths_kernel<HashDagDev><<<grid, block, 0, c->stream>>>(cam, d.hash, c->levels, c->map, c->ray_planes(0), c->paths, c->tables, beams, tag, anc);
    else trace_paths_kernel<HashDagResolvedDev><<<grid, block, 0, c->stream>>>(cam, d.resolved, c->levels, c->map, c->ray_planes(0), c->paths, c->tables, beams, tag, anc);
    HDT_LAUNCHED("trace_paths_kernel");
    if (anc.a) { c->ancValid = true; c->ancForPool = d.resolved.pool; c->ancForPrefix = d.resolved.prefix; c->ancForRoot = d.resolved.firstNodeIndex; }
    HDT_CUDA(cudaEventRecord(c->traceDone[0], c->stream));
    HDT_CUDA(cudaStreamWaitEvent(c->stream, c->join[0], 0));   // a synchronisation of the main stream covers the beam kernel too
    return HDT_OK;
}

// Can trace_colors take the short route (color_pixel_recorded)?  Only for the DAG the current paths frame was traced in
// (same resolved pool, prefix pool and root: the records hold physical word indices of that pool), HashDAGColors, and a
// view that decodes a colour (the index / position / colour-tree debug views show what only the full walk computes).
bool colors_recorded_ok(const hdt_ctx* c, const DagArg& d, const ColorsDev& col, const ColorsParams& prm)
{
    return c->useRecorded && c->ancValid && d.kind == HDT_DAG_HASH_RESOLVED && d.resolved.pool == c->ancForPool && d.resolved.prefix &&
           d.resolved.prefix == c->ancForPrefix && d.resolved.firstNodeIndex == c->ancForRoot && col.kind == HDT_COLORS_HASH &&
           prm.debugColors != HDT_DEBUG_INDEX && prm.debugColors != HDT_DEBUG_POSITION && prm.debugColors != HDT_DEBUG_COLOR_TREE;
}

int launch_colors(hdt_ctx* c, const DagArg& d, const ColorsDev& col, const ColorsParams& prm)
{
    const dim3 grid(c->grid_blocks()), block(kBlockThreads);
    if (!grid.x) return HDT_OK;
    if (int rc = pin_page_table(c, d, d.pageTableSize)) return rc;
    if (colors_recorded_ok(c, d, col, prm))
    {
        trace_colors_recorded_kernel<<<grid, block, 0, c->stream>>>(d.resolved.prefix, col, c->levels, prm, c->map, c->paths, c->anc[0], c->anc[1], c->colors);
        ++c->recordedColorPasses;
    }
    else if (d.kind == HDT_DAG_BASIC) trace_colors_kernel<BasicDagDev><<<grid, block, 0, c->stream>>>(d.basic, col, c->levels, prm, c->map, c->paths, c->colors);
    else if (d.kind == HDT_DAG_HASH) trace_colors_kernel<HashDagDev><<<grid, block, 0, c->stream>>>(d.hash, col, c->levels, prm, c->map, c->paths, c->colors);
    else trace_colors_kernel<HashDagResolvedDev><<<grid, block, 0, c->stream>>>(d.resolved, col, c->levels, prm, c->map, c->paths, c->colors);
    HDT_LAUNCHED("trace_colors_kernel");
    return HDT_OK;
}
// trace_shadows in two halves so that a whole-frame call can enqueue the first one (ray setup + beams, which
// only need the paths frame; side stream) before the colours kernel and the second one after it.
struct ShadowPrep { const BeamState* beams = nullptr; u32 tag = 0; bool valid = false; };

int prepare_shadows(hdt_ctx* c, const DagArg& d, const CameraParams& cam, const ShadowParams& sp, ShadowPrep& prep)
{
    prep = ShadowPrep{};
    const dim3 grid(c->grid_blocks()), block(kBlockThreads);
    if (!grid.x) return HDT_OK;
    prep.valid = true;
    if (int rc = pin_page_table(c, d, d.pageTableSize)) return rc;
    HDT_CUDA(cudaEventRecord(c->fork[1], c->stream));
    HDT_CUDA(cudaStreamWaitEvent(c->side, c->fork[1], 0));
    setup_shadows_kernel<<<grid, block, 0, c->side>>>(cam, sp, c->map, c->paths, c->ray_planes(1), c->seeds[1]);
    HDT_LAUNCHED("setup_shadows_kernel");
    HDT_CUDA(cudaEventRecord(c->setupDone[1], c->side));
    if (c->useBeams) {
        prep.beams = c->beams[1];
        prep.tag = next_beam_tag(c);
        const u32 nb = c->n_beams();
        const dim3 g((nb + 31) / 32), b(32);
        if (d.kind == HDT_DAG_BASIC) beam_shadows_kernel<BasicDagDev><<<g, b, 0, c->side>>>(sp, d.basic, c->levels, c->seeds[1], c->beams[1], nb, c->beamMaxVisits, prep.tag, c->tables);
        else if (d.kind == HDT_DAG_HASH) beam_shadows_kernel<HashDagDev><<<g, b, 0, c->side>>>(sp, d.hash, c->levels, c->seeds[1], c->beams[1], nb, c->beamMaxVisits, prep.tag, c->tables);
        else beam_shadows_kernel<HashDagResolvedDev><<<g, b, 0, c->side>>>(sp, d.resolved, c->levels, c->seeds[1], c->beams[1], nb, c->beamMaxVisits, prep.tag, c->tables);
        HDT_LAUNCHED("beam_shadows_kernel");
        c->lastBeamPass = 1;
    }
    HDT_CUDA(cudaEventRecord(c->join[1], c->side));
    return HDT_OK;
}
int finish_shadows(hdt_ctx* c, const DagArg& d, const CameraParams& cam, const ShadowParams& sp, const ShadowPrep& prep)
{
    if (!prep.valid) return HDT_OK;
    const dim3 grid(c->grid_blocks()), block(kBlockThreads);
    HDT_CUDA(cudaStreamWaitEvent(c->stream, c->beamSerial ? c->join[1] : c->setupDone[1], 0));   // the origins
    // Fused framebuffer exchange: this pass writes the final colours, so it can store them into rank 0's frame itself.
    ExchangeOut xo{ nullptr, nullptr, nullptr, 0 };
    if (c->xFused && c->xBlock) {
        // arrived[rank] = seq is a plain store: a shadows pass repeated before hdt_exchange_frame (a re-render, a standalone
        // hdt_resolve_shadows) stores the same frame number again and cannot run the root ahead
        const u32 seq = c->xSeq + 1;
        ExchangeCounters* k = reinterpret_cast<ExchangeCounters*>(reinterpret_cast<char*>(c->xBlock) + ((size_t(c->map.width) * c->map.height * 4 + 255) & ~size_t(255)));
        if (c->map.rank != 0) {   // rank 0 must have consumed the previous frame of this lane
            exchange_wait_kernel<<<1, 1, 0, c->stream>>>(&k->credit, 1, seq - 1, c->xWaitCycles, c->xTimedOutDev, c->xCtasDone + 1);
            HDT_LAUNCHED("exchange_wait_kernel");
        }
        xo = ExchangeOut{ c->xBlock, c->xCtasDone, c->map.rank != 0 ? &k->arrived[c->map.rank] : nullptr, seq };
        c->xFusedSeq = seq;
    }
    if (d.has_prefix()) trace_shadows_kernel<HashDagPrefixDev><<<grid, block, 0, c->stream>>>(cam, sp, d.prefixed(), c->levels, c->map, c->paths, c->ray_planes(1), c->colors, c->tables, prep.beams, prep.tag, xo);
    else if (d.kind == HDT_DAG_BASIC) trace_shadows_kernel<BasicDagDev><<<grid, block, 0, c->stream>>>(cam, sp, d.basic, c->levels, c->map, c->paths, c->ray_planes(1), c->colors, c->tables, prep.beams, prep.tag, xo);
    else if (d.kind == HDT_DAG_HASH) trace_shadows_kernel<HashDagDev><<<grid, block, 0, c->stream>>>(cam, sp, d.hash, c->levels, c->map, c->paths, c->ray_planes(1), c->colors, c->tables, prep.beams, prep.tag, xo);
    else trace_shadows_kernel<HashDagResolvedDev><<<grid, block, 0, c->stream>>>(cam, sp, d.resolved, c->levels, c->map, c->paths, c->ray_planes(1), c->colors, c->tables, prep.beams, prep.tag, xo);
    HDT_LAUNCHED("trace_shadows_kernel");
    HDT_CUDA(cudaEventRecord(c->traceDone[1], c->stream));
    HDT_CUDA(cudaStreamWaitEvent(c->stream, c->join[1], 0));
    return HDT_OK;
}
int launch_shadows(hdt_ctx* c, const DagArg& d, const CameraParams& cam, const ShadowParams& sp)
{
    ShadowPrep prep;
    if (int rc = prepare_shadows(c, d, cam, sp, prep)) return rc;
    return finish_shadows(c, d, cam, sp, prep);
}

int check_combo(int dagKind, int colorsKind)
{
    // the instantiations the reference provides (tracer.cu:705-711)
    if ((dagKind == HDT_DAG_HASH || dagKind == HDT_DAG_HASH_RESOLVED) != (colorsKind == HDT_COLORS_HASH)) return fail(HDT_ERR_ARG, "DAG / colours combination not provided by the tracer");
    return HDT_OK;
}

// After a host synchronisation: did a framebuffer-exchange wait give up since the last check?  (The stream is idle here.)
int check_exchange(hdt_ctx* c)
{
    if (c->xTimedOut && *reinterpret_cast<volatile u32*>(c->xTimedOut)) {
        *reinterpret_cast<volatile u32*>(c->xTimedOut) = 0;
        if (c->xCtasDone) cudaMemsetAsync(c->xCtasDone, 0, 2 * sizeof(u32), c->stream);
        return fail(HDT_ERR_STATE, "framebuffer exchange: a rank did not arrive (or rank 0 did not release) within the exchange timeout; the frame was dropped");
    }
    return HDT_OK;
}

int finish_timed(hdt_ctx* c, cudaEvent_t a, cudaEvent_t b, float* ms)
{
    HDT_CUDA(cudaEventSynchronize(b));
    HDT_CUDA(cudaGetLastError());
    if (int rc = check_exchange(c)) return rc;
    if (ms) HDT_CUDA(cudaEventElapsedTime(ms, a, b));
    return HDT_OK;
}

}  // namespace

extern "C" {

const char* hdt_last_error(void) { return g_lastError.c_str(); }
int hdt_version(void) { return 1; }
uint64_t hdt_launch_count(const hdt_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t hdt_recorded_color_passes(const hdt_ctx* ctx) { return ctx ? ctx->recordedColorPasses : 0; }

int hdt_create(uint32_t width, uint32_t height, uint32_t levels, int device, hdt_ctx** out)
{
    if (!out || !width || !height) return fail(HDT_ERR_ARG, "hdt_create: bad arguments");
    if (levels < 3 || levels > kMaxLevels) return fail(HDT_ERR_ARG, "hdt_create: levels must be in [3, 24]");
    int n = 0;
    HDT_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(HDT_ERR_ARG, "hdt_create: no such CUDA device");
    HDT_CUDA(cudaSetDevice(device));
    hdt_ctx* c = new hdt_ctx();
    c->device = device; c->levels = levels;
    c->map.width = width; c->map.height = height;
    if (const char* env = getenv("HDT_BEAMS")) c->useBeams = atoi(env) != 0;
    if (const char* env = getenv("HDT_BEAM_PREFETCH")) c->beamPrefetch = atoi(env) != 0;
    if (const char* env = getenv("HDT_COLORS_RECORDED")) c->useRecorded = atoi(env) != 0;
    if (const char* env = getenv("HDT_L2_PERSIST")) c->l2Persist = atoi(env) != 0;
    if (const char* env = getenv("HDT_BEAM_MAX_VISITS")) c->beamMaxVisits = u32(atoi(env) > 0 ? atoi(env) : 1);
    cudaError_t e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking);
    c->stream = c->ownStream;
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->timer[i]);
    {
        int lo = 0, hi = 0;
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi);   // hi = greatest priority
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->fork[i]);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->join[i]);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->setupDone[i]);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->traceDone[i]);
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->hitCounter, 8 * sizeof(unsigned long long));   // hit count / beam statistics
    if (e == cudaSuccess) e = cudaMallocHost(&c->pathCache, 4 * sizeof(u32));
    if (e != cudaSuccess) { hdt_destroy(c); return cuda_fail(e, "hdt_create"); }
    {
        TraverseTables host;
        build_tables(host);
        e = cudaMalloc(&c->tables, sizeof(TraverseTables));
        if (e == cudaSuccess) e = cudaMemcpy(c->tables, &host, sizeof(host), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { hdt_destroy(c); return cuda_fail(e, "hdt_create: tables"); }
    }
    const int rc = configure(c, 0, 1, 6);
    if (rc) { hdt_destroy(c); return rc; }
    *out = c;
    return HDT_OK;
}

int hdt_destroy(hdt_ctx* c)
{
    if (!c) return HDT_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->ownStream) cudaStreamSynchronize(c->ownStream);
    cudaFree(c->paths); cudaFree(c->colors); cudaFree(c->framePaths); cudaFree(c->frameColors);
    if (c->pathCache) cudaFreeHost(c->pathCache);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->timer) if (e) cudaEventDestroy(e);
    cudaFree(c->hitCounter);
    cudaFree(c->tables);
    if (c->xBlock && c->xOwned) cudaFree(c->xBlock);
    if (c->xBlock && c->xIpc) cudaIpcCloseMemHandle(c->xBlock);
    if (c->xHostBlock && c->xHostRegistered) cudaHostUnregister(c->xHostBlock);
    cudaFree(c->xCtasDone);
    if (c->xTimedOut) cudaFreeHost(c->xTimedOut);
    hdt_comm_destroy(c);
    cudaFree(c->physToVirt);
    cudaFree(c->foaScratch);
    cudaFree(c->rebuildScratch);
    if (c->stagingHost) cudaFreeHost(c->stagingHost);
    cudaFree(c->stagingDev);
    if (c->side) cudaStreamSynchronize(c->side);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->beams[i]); cudaFree(c->seeds[i]); cudaFree(c->rays[i]); cudaFree(c->anc[i]);
        if (c->fork[i]) cudaEventDestroy(c->fork[i]);
        if (c->join[i]) cudaEventDestroy(c->join[i]);
        if (c->setupDone[i]) cudaEventDestroy(c->setupDone[i]);
        if (c->traceDone[i]) cudaEventDestroy(c->traceDone[i]);
    }
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
    return HDT_OK;
}

int hdt_set_option(hdt_ctx* c, int option, int value)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (option == HDT_OPT_BEAMS) { c->useBeams = value != 0; return HDT_OK; }
    if (option == HDT_OPT_BEAM_PREFETCH) { c->beamPrefetch = value != 0; return HDT_OK; }
    if (option == HDT_OPT_BEAM_SERIAL) { c->beamSerial = value != 0; return HDT_OK; }
    if (option == HDT_OPT_EXCHANGE_FUSED) { c->xFused = value != 0; return HDT_OK; }
    if (option == HDT_OPT_L2_PERSIST) { c->l2Persist = value != 0; return HDT_OK; }
    if (option == HDT_OPT_COLORS_RECORDED) { c->useRecorded = value != 0; c->ancValid = false; return HDT_OK; }
    if (option == HDT_OPT_EXCHANGE_TIMEOUT_MS) {
        if (value < 0) return fail(HDT_ERR_ARG, "exchange timeout must be >= 0 ms (0 = wait for ever)");
        c->xWaitCycles = (unsigned long long)value * 2000000ull;   // SM clock <= 2 GHz: at least `value` ms
        return HDT_OK;
    }
    if (option == HDT_OPT_BEAM_MAX_VISITS) { if (value < 1) return fail(HDT_ERR_ARG, "beam visit cap must be >= 1"); c->beamMaxVisits = u32(value); return HDT_OK; }
    return fail(HDT_ERR_ARG, "unknown option");
}

int hdt_set_partition(hdt_ctx* c, uint32_t rank, uint32_t world, uint32_t tile_log2)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    return configure(c, rank, world, tile_log2);
}

int hdt_resolve_paths(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, const double cam[3], const double ray_min[3],
                      const double ray_ddx[3], const double ray_ddy[3], float* ms)
{
    if (!c || !cam || !ray_min || !ray_ddx || !ray_ddy) return fail(HDT_ERR_ARG, "hdt_resolve_paths: null argument");
    DagArg d;
    if (int rc = parse_dag(dag_kind, dag_pod, dag_pod_size, d)) return rc;
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (int rc = launch_paths(c, d, make_cam(cam, ray_min, ray_ddx, ray_ddy))) return rc;
    HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    return finish_timed(c, c->ev[0], c->ev[1], ms);
}

int hdt_resolve_colors(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, int colors_kind, const void* colors_pod,
                       size_t colors_pod_size, int debug_colors, uint32_t debug_level, const hdt_tool_info* tool, int overlay, float* ms)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (debug_colors < 0 || debug_colors > HDT_DEBUG_WEIGHT) return fail(HDT_ERR_ARG, "unknown debug colour mode");
    DagArg d; ColorsDev col;
    if (int rc = parse_dag(dag_kind, dag_pod, dag_pod_size, d)) return rc;
    if (int rc = parse_colors(colors_kind, colors_pod, colors_pod_size, col)) return rc;
    if (int rc = check_combo(dag_kind, colors_kind)) return rc;
    ColorsParams prm{};
    prm.debugColors = debug_colors; prm.debugIndexLevel = debug_level; prm.overlay = (overlay && tool) ? 1 : 0;
    if (tool) prm.tool = *tool;
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (int rc = launch_colors(c, d, col, prm)) return rc;
    HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    return finish_timed(c, c->ev[0], c->ev[1], ms);
}

int hdt_resolve_shadows(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, const double cam[3], const double ray_min[3],
                        const double ray_ddx[3], const double ray_ddy[3], float shadow_bias, float fog_density, float* ms)
{
    if (!c || !cam || !ray_min || !ray_ddx || !ray_ddy) return fail(HDT_ERR_ARG, "hdt_resolve_shadows: null argument");
    DagArg d;
    if (int rc = parse_dag(dag_kind, dag_pod, dag_pod_size, d)) return rc;
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (int rc = launch_shadows(c, d, make_cam(cam, ray_min, ray_ddx, ray_ddy), make_shadow(shadow_bias, fog_density))) return rc;
    HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    return finish_timed(c, c->ev[0], c->ev[1], ms);
}

static int enqueue_frame(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, int colors_kind, const void* colors_pod,
                         size_t colors_pod_size, const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3],
                         float shadow_bias, float fog_density, int with_shadows, uint32_t* host_colors, bool events)
{
    if (!c || !cam || !ray_min || !ray_ddx || !ray_ddy) return fail(HDT_ERR_ARG, "hdt_resolve_frame: null argument");
    DagArg d; ColorsDev col;
    if (int rc = parse_dag(dag_kind, dag_pod, dag_pod_size, d)) return rc;
    if (int rc = parse_colors(colors_kind, colors_pod, colors_pod_size, col)) return rc;
    if (int rc = check_combo(dag_kind, colors_kind)) return rc;
    if (host_colors && c->map.world != 1) return fail(HDT_ERR_STATE, "hdt_resolve_frame: host read-back needs an unpartitioned context");
    const CameraParams cp = make_cam(cam, ray_min, ray_ddx, ray_ddy);
    ColorsParams prm{};
    HDT_CUDA(cudaSetDevice(c->device));
    if (events) HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (int rc = launch_paths(c, d, cp)) return rc;
    if (events) HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    // the shadow pass' ray setup and beams only need the paths frame: start them beside the colours kernel
    const ShadowParams sp = make_shadow(shadow_bias, fog_density);
    ShadowPrep prep;
    if (with_shadows) if (int rc = prepare_shadows(c, d, cp, sp, prep)) return rc;
    if (int rc = launch_colors(c, d, col, prm)) return rc;
    if (events) HDT_CUDA(cudaEventRecord(c->ev[2], c->stream));
    if (with_shadows) if (int rc = finish_shadows(c, d, cp, sp, prep)) return rc;
    if (events) HDT_CUDA(cudaEventRecord(c->ev[3], c->stream));
    if (host_colors)
        HDT_CUDA(cudaMemcpyAsync(host_colors, c->colors, u64(c->map.width) * c->map.height * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    return HDT_OK;
}

int hdt_resolve_frame(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, int colors_kind, const void* colors_pod,
                      size_t colors_pod_size, const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3],
                      float shadow_bias, float fog_density, int with_shadows, uint32_t* host_colors, float ms[3])
{
    if (int rc = enqueue_frame(c, dag_kind, dag_pod, dag_pod_size, colors_kind, colors_pod, colors_pod_size, cam, ray_min, ray_ddx, ray_ddy,
                               shadow_bias, fog_density, with_shadows, host_colors, true)) return rc;
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    if (int rc = check_exchange(c)) return rc;
    if (ms) {
        HDT_CUDA(cudaEventElapsedTime(&ms[0], c->ev[0], c->ev[1]));
        HDT_CUDA(cudaEventElapsedTime(&ms[1], c->ev[1], c->ev[2]));
        HDT_CUDA(cudaEventElapsedTime(&ms[2], c->ev[2], c->ev[3]));
    }
    return HDT_OK;
}

int hdt_resolve_frame_async(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, int colors_kind, const void* colors_pod,
                            size_t colors_pod_size, const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3],
                            float shadow_bias, float fog_density, int with_shadows, uint32_t* host_colors)
{
    return enqueue_frame(c, dag_kind, dag_pod, dag_pod_size, colors_kind, colors_pod, colors_pod_size, cam, ray_min, ray_ddx, ray_ddy,
                         shadow_bias, fog_density, with_shadows, host_colors, false);
}

int hdt_sync(hdt_ctx* c)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    return check_exchange(c);
}

int hdt_timer_begin(hdt_ctx* c)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaEventRecord(c->timer[0], c->stream));
    return HDT_OK;
}

int hdt_timer_end(hdt_ctx* c, float* ms)
{
    if (!c || !ms) return fail(HDT_ERR_ARG, "null argument");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaEventRecord(c->timer[1], c->stream));
    HDT_CUDA(cudaEventSynchronize(c->timer[1]));
    HDT_CUDA(cudaGetLastError());
    HDT_CUDA(cudaEventElapsedTime(ms, c->timer[0], c->timer[1]));
    return check_exchange(c);
}

int hdt_count_hits(hdt_ctx* c, uint64_t* n_hits)
{
    if (!c || !n_hits) return fail(HDT_ERR_ARG, "null argument");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaMemsetAsync(c->hitCounter, 0, sizeof(unsigned long long), c->stream));
    const u64 n = c->map.world == 1 ? u64(c->map.width) * c->map.height : (u64(c->nOwnedTiles) << (2 * c->map.tileLog2));
    count_hits_kernel<<<592, 256, 0, c->stream>>>(c->paths, n, c->hitCounter);
    ++c->launches;
    unsigned long long v = 0;
    HDT_CUDA(cudaMemcpyAsync(&v, c->hitCounter, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    *n_hits = v;
    return HDT_OK;
}

#ifdef HDT_BEAM_DEBUG
extern "C" int hdt_debug_beam_counters(uint32_t out[8], int reset)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_beamDebug, 8 * sizeof(u32)) != cudaSuccess) return 1;
    if (reset) { const u32 z[8] = {}; cudaMemcpyToSymbol(g_beamDebug, z, sizeof(z)); }
    return 0;
}
#endif

int hdt_beam_stats(hdt_ctx* c, uint64_t out[5])
{
    if (!c || !out) return fail(HDT_ERR_ARG, "null argument");
    for (int i = 0; i < 5; ++i) out[i] = 0;
    const u32 n = c->n_beams();
    const BeamState* beams = c->beams[c->lastBeamPass];
    if (!beams || !n) return HDT_OK;
    HDT_CUDA(cudaSetDevice(c->device));
    unsigned long long* dev = c->hitCounter;   // 8 words of scratch
    HDT_CUDA(cudaMemsetAsync(dev, 0, 5 * sizeof(unsigned long long), c->stream));
    beam_stats_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(beams, n, dev);
    HDT_LAUNCHED("beam_stats_kernel");
    unsigned long long host[5] = {};
    HDT_CUDA(cudaMemcpyAsync(host, dev, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 5; ++i) out[i] = host[i];
    return HDT_OK;
}

int hdt_pass_timeline(hdt_ctx* c, int pass, float ms[3])
{
    if (!c || !ms || pass < 0 || pass > 1) return fail(HDT_ERR_ARG, "bad argument");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->side));
    HDT_CUDA(cudaEventElapsedTime(&ms[0], c->fork[pass], c->setupDone[pass]));
    HDT_CUDA(cudaEventElapsedTime(&ms[1], c->fork[pass], c->join[pass]));
    HDT_CUDA(cudaEventElapsedTime(&ms[2], c->fork[pass], c->traceDone[pass]));
    return HDT_OK;
}

int hdt_get_path(hdt_ctx* c, uint32_t x, uint32_t y, uint32_t out[3])
{
    if (!c || !out) return fail(HDT_ERR_ARG, "null argument");
    if (x >= c->map.width || y >= c->map.height) return fail(HDT_ERR_ARG, "pixel outside the frame");
    const u32 t = (y >> c->map.tileLog2) * c->map.tilesX + (x >> c->map.tileLog2);
    if (t % c->map.world != c->map.rank) return fail(HDT_ERR_STATE, "pixel belongs to another rank");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaMemcpyAsync(c->pathCache, c->paths + c->map.index(x, y), sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    out[0] = c->pathCache[0]; out[1] = c->pathCache[1]; out[2] = c->pathCache[2];
    return HDT_OK;
}

static int read_frame(hdt_ctx* c, bool paths, void* host)
{
    if (!c || !host) return fail(HDT_ERR_ARG, "null argument");
    HDT_CUDA(cudaSetDevice(c->device));
    const u64 n = u64(c->map.width) * c->map.height;
    if (c->map.world == 1) {
        if (paths) HDT_CUDA(cudaMemcpyAsync(host, c->paths, n * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
        else HDT_CUDA(cudaMemcpyAsync(host, c->colors, n * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
        HDT_CUDA(cudaStreamSynchronize(c->stream));
        return check_exchange(c);
    }
    const dim3 block(32, 8), grid((c->map.width + 31) / 32, (c->map.height + 7) / 8);
    if (paths) {
        if (!c->framePaths) HDT_CUDA(cudaMalloc(&c->framePaths, n * sizeof(uint4)));
        HDT_CUDA(cudaMemsetAsync(c->framePaths, 0, n * sizeof(uint4), c->stream));
        untile_own_kernel<uint4><<<grid, block, 0, c->stream>>>(c->paths, c->framePaths, c->map);
        HDT_CUDA(cudaMemcpyAsync(host, c->framePaths, n * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
    } else {
        if (!c->frameColors) HDT_CUDA(cudaMalloc(&c->frameColors, n * sizeof(u32)));
        HDT_CUDA(cudaMemsetAsync(c->frameColors, 0, n * sizeof(u32), c->stream));
        untile_own_kernel<u32><<<grid, block, 0, c->stream>>>(c->colors, c->frameColors, c->map);
        HDT_CUDA(cudaMemcpyAsync(host, c->frameColors, n * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    }
    ++c->launches;
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    return check_exchange(c);
}

int hdt_read_paths(hdt_ctx* c, uint32_t* host) { return read_frame(c, true, host); }
int hdt_read_colors(hdt_ctx* c, uint32_t* host) { return read_frame(c, false, host); }

int hdt_partition_buffers(hdt_ctx* c, void** paths_dev, void** colors_dev, uint64_t* n_owned, uint64_t* max_tiles)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (paths_dev) *paths_dev = c->paths;
    if (colors_dev) *colors_dev = c->colors;
    if (n_owned) *n_owned = c->nOwnedTiles;
    if (max_tiles) *max_tiles = c->map.world == 1 ? 0 : c->maxTilesPerRank;
    return HDT_OK;
}

int hdt_assemble_colors(hdt_ctx* c, const uint32_t* gathered_dev, uint32_t* frame_dev)
{
    if (!c || !gathered_dev) return fail(HDT_ERR_ARG, "null argument");
    HDT_CUDA(cudaSetDevice(c->device));
    const u64 n = u64(c->map.width) * c->map.height;
    if (!frame_dev) {
        if (!c->frameColors) HDT_CUDA(cudaMalloc(&c->frameColors, n * sizeof(u32)));
        frame_dev = c->frameColors;
    }
    const dim3 block(32, 8), grid((c->map.width + 31) / 32, (c->map.height + 7) / 8);
    assemble_kernel<u32><<<grid, block, 0, c->stream>>>(gathered_dev, frame_dev, c->map, c->buffer_pixels());
    ++c->launches;
    HDT_CUDA(cudaGetLastError());
    return HDT_OK;
}

int hdt_set_stream(hdt_ctx* c, void* cuda_stream)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->ownStream;
    return HDT_OK;
}

int hdt_apply_ranges(hdt_ctx* c, uint32_t* dst_dev, const uint32_t* payload_dev, const hdt_range* ranges_dev, uint32_t n_ranges)
{
    if (!c || !dst_dev || !payload_dev || (!ranges_dev && n_ranges)) return fail(HDT_ERR_ARG, "null argument");
    if (!n_ranges) return HDT_OK;
    HDT_CUDA(cudaSetDevice(c->device));
    apply_ranges_kernel<<<n_ranges < 1184 ? n_ranges : 1184, 128, 0, c->stream>>>(dst_dev, payload_dev, ranges_dev, n_ranges);
    ++c->launches;
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    return HDT_OK;
}

// ---- framebuffer exchange over peer memory (hdt_exchange.cuh) ---------------------------------------
namespace {
// frame part of an exchange block: row-major W x H (device block), or -- host block -- every rank's compact tile buffer
// (maxTilesPerRank tiles) back to back
size_t exchange_frame_bytes(const hdt_ctx* c)
{
    if (c->xHostBlock) return (size_t(c->map.world) * c->maxTilesPerRank * (size_t(4) << (2 * c->map.tileLog2)) + 255) & ~size_t(255);
    return (size_t(c->map.width) * c->map.height * 4 + 255) & ~size_t(255);
}
ExchangeCounters* exchange_counters(const hdt_ctx* c) { return reinterpret_cast<ExchangeCounters*>(reinterpret_cast<char*>(c->xBlock) + exchange_frame_bytes(c)); }
int exchange_common(hdt_ctx* c)
{
    if (c->map.world > kMaxExchangeRanks) return fail(HDT_ERR_ARG, "framebuffer exchange: at most 64 ranks");
    if (!c->xCtasDone) {
        HDT_CUDA(cudaMalloc(&c->xCtasDone, 2 * sizeof(u32)));
        HDT_CUDA(cudaMemset(c->xCtasDone, 0, 2 * sizeof(u32)));
    }
    if (!c->xTimedOut) {
        HDT_CUDA(cudaHostAlloc(&c->xTimedOut, sizeof(u32), cudaHostAllocMapped));
        *c->xTimedOut = 0;
        HDT_CUDA(cudaHostGetDevicePointer(&c->xTimedOutDev, c->xTimedOut, 0));
    }
    c->xSeq = 0; c->xFusedSeq = 0;
    return HDT_OK;
}
}  // namespace

int hdt_exchange_create(hdt_ctx* c, uint8_t ipc_handle_out[HDT_IPC_HANDLE_BYTES], void** frame_dev_out)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (c->map.rank != 0) return fail(HDT_ERR_STATE, "hdt_exchange_create: only rank 0 owns the frame");
    if (c->xBlock) return fail(HDT_ERR_STATE, "hdt_exchange_create: the context already has an exchange");
    static_assert(sizeof(cudaIpcMemHandle_t) == HDT_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    HDT_CUDA(cudaSetDevice(c->device));
    const size_t bytes = exchange_frame_bytes(c) + sizeof(ExchangeCounters);
    HDT_CUDA(cudaMalloc(&c->xBlock, bytes));
    HDT_CUDA(cudaMemset(c->xBlock, 0, bytes));
    c->xOwned = true;
    if (int rc = exchange_common(c)) return rc;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        HDT_CUDA(cudaIpcGetMemHandle(&h, c->xBlock));
        memcpy(ipc_handle_out, &h, sizeof(h));
    }
    if (frame_dev_out) *frame_dev_out = c->xBlock;
    return HDT_OK;
}

int hdt_exchange_open(hdt_ctx* c, const uint8_t ipc_handle[HDT_IPC_HANDLE_BYTES])
{
    if (!c || !ipc_handle) return fail(HDT_ERR_ARG, "null argument");
    if (c->map.rank == 0) return fail(HDT_ERR_STATE, "hdt_exchange_open: rank 0 creates the exchange");
    if (c->xBlock) return fail(HDT_ERR_STATE, "hdt_exchange_open: the context already has an exchange");
    HDT_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    HDT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->xBlock = static_cast<u32*>(p);
    c->xIpc = true;
    return exchange_common(c);
}

int hdt_exchange_block(hdt_ctx* c, void** block_dev_out)
{
    if (!c || !block_dev_out) return fail(HDT_ERR_ARG, "null argument");
    if (!c->xBlock) return fail(HDT_ERR_STATE, "no exchange on this context");
    *block_dev_out = c->xBlock;
    return HDT_OK;
}

int hdt_exchange_attach(hdt_ctx* c, void* block_dev)
{
    if (!c || !block_dev) return fail(HDT_ERR_ARG, "null argument");
    if (c->map.rank == 0) return fail(HDT_ERR_STATE, "hdt_exchange_attach: rank 0 creates the exchange");
    if (c->xBlock) return fail(HDT_ERR_STATE, "hdt_exchange_attach: the context already has an exchange");
    HDT_CUDA(cudaSetDevice(c->device));
    c->xBlock = static_cast<u32*>(block_dev);
    return exchange_common(c);
}

int hdt_exchange_block_bytes(hdt_ctx* c, uint64_t* bytes)
{
    if (!c || !bytes) return fail(HDT_ERR_ARG, "null argument");
    *bytes = ((size_t(c->map.world) * c->maxTilesPerRank * (size_t(4) << (2 * c->map.tileLog2)) + 255) & ~size_t(255)) + sizeof(ExchangeCounters);
    return HDT_OK;
}

int hdt_exchange_attach_host(hdt_ctx* c, void* host_block, uint64_t bytes)
{
    if (!c || !host_block) return fail(HDT_ERR_ARG, "null argument");
    if (c->xBlock) return fail(HDT_ERR_STATE, "hdt_exchange_attach_host: the context already has an exchange");
    uint64_t need = 0;
    hdt_exchange_block_bytes(c, &need);
    if (bytes < need) return fail(HDT_ERR_CAPACITY, "hdt_exchange_attach_host: block smaller than hdt_exchange_block_bytes");
    HDT_CUDA(cudaSetDevice(c->device));
    // pin + map the caller's (shared) memory; a second context of this process finds it registered already
    cudaError_t e = cudaHostRegister(host_block, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) cudaGetLastError();
    else if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister(exchange block)");
    else c->xHostRegistered = true;
    void* dev = nullptr;
    HDT_CUDA(cudaHostGetDevicePointer(&dev, host_block, 0));
    c->xBlock = static_cast<u32*>(dev);     // device alias: the counters behind the frame are polled / written by kernels
    c->xHostBlock = host_block;
    if (c->xFused) return fail(HDT_ERR_STATE, "hdt_exchange_attach_host: not with HDT_OPT_EXCHANGE_FUSED");
    return exchange_common(c);
}

int hdt_exchange_frame(hdt_ctx* c)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (!c->xBlock) return fail(HDT_ERR_STATE, "hdt_exchange_frame: no exchange (hdt_exchange_create / open / attach first)");
    HDT_CUDA(cudaSetDevice(c->device));
    const u32 seq = ++c->xSeq;
    ExchangeCounters* k = exchange_counters(c);
    u32* timedOutDev = c->xTimedOutDev;
    const u32 T = 1u << c->map.tileLog2;
    const bool root = c->map.rank == 0;
    u32* abortDev = c->xCtasDone + 1;
    if (c->xHostBlock) {
        // host block: this rank's compact tile buffer goes to its slice of the shared host frame in ONE copy-engine transfer
        // over the rank's own PCIe link; the arrival word follows in stream order
        const size_t tileBytes = size_t(4) << (2 * c->map.tileLog2), slice = size_t(c->maxTilesPerRank) * tileBytes;
        if (!root) {
            exchange_wait_kernel<<<1, 1, 0, c->stream>>>(&k->credit, 1, seq - 1, c->xWaitCycles, timedOutDev, abortDev);
            HDT_LAUNCHED("exchange_wait_kernel");
        }
        if (c->nOwnedTiles)
            HDT_CUDA(cudaMemcpyAsync(static_cast<char*>(c->xHostBlock) + size_t(c->map.rank) * slice, c->colors, size_t(c->nOwnedTiles) * tileBytes, cudaMemcpyDeviceToHost, c->stream));
        if (!root) {
            exchange_signal_kernel<<<1, 1, 0, c->stream>>>(&k->arrived[c->map.rank], seq, abortDev);
            HDT_LAUNCHED("exchange_signal_kernel");
        } else if (c->map.world > 1) {
            exchange_wait_kernel<<<1, 1, 0, c->stream>>>(&k->arrived[1], c->map.world - 1, seq, c->xWaitCycles, timedOutDev, abortDev);
            HDT_LAUNCHED("exchange_wait_kernel");
        }
        return HDT_OK;
    }
    const bool fused = c->xFusedSeq == seq && c->grid_blocks() > 0;   // the last shadow pass already stored (and signalled) this frame
    const u32 grid = fused ? 0 : c->nOwnedTiles * (T / 16);
    if (!root && !fused) {   // the root must have consumed the previous frame of this lane before it is overwritten
        exchange_wait_kernel<<<1, 1, 0, c->stream>>>(&k->credit, 1, seq - 1, c->xWaitCycles, timedOutDev, abortDev);
        HDT_LAUNCHED("exchange_wait_kernel");
    }
    if (grid) {
        exchange_scatter_kernel<<<grid, 256, 0, c->stream>>>(c->colors, c->xBlock, c->map, c->xCtasDone, root ? nullptr : &k->arrived[c->map.rank], seq, abortDev);
        HDT_LAUNCHED("exchange_scatter_kernel");
    } else if (!root && !fused) {
        exchange_signal_kernel<<<1, 1, 0, c->stream>>>(&k->arrived[c->map.rank], seq, abortDev);
        HDT_LAUNCHED("exchange_signal_kernel");
    }
    if (root && c->map.world > 1) {
        exchange_wait_kernel<<<1, 1, 0, c->stream>>>(&k->arrived[1], c->map.world - 1, seq, c->xWaitCycles, timedOutDev, abortDev);
        HDT_LAUNCHED("exchange_wait_kernel");
    }
    return HDT_OK;
}

int hdt_exchange_release(hdt_ctx* c)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (!c->xBlock || c->map.rank != 0) return fail(HDT_ERR_STATE, "hdt_exchange_release: rank 0 with an exchange only");
    HDT_CUDA(cudaSetDevice(c->device));
    exchange_publish_kernel<<<1, 1, 0, c->stream>>>(&exchange_counters(c)->credit, c->xSeq);
    ++c->launches;
    HDT_CUDA(cudaGetLastError());
    return HDT_OK;
}

int hdt_apply_ranges_host(hdt_ctx* c, uint32_t* dst_dev, const uint32_t* payload_host, uint64_t n_payload_words, const hdt_range* ranges_host, uint32_t n_ranges)
{
    if (!c || !dst_dev || (!ranges_host && n_ranges) || (!payload_host && n_payload_words)) return fail(HDT_ERR_ARG, "hdt_apply_ranges_host: null argument");
    if (!n_ranges) return HDT_OK;
    for (u32 i = 0; i < n_ranges; ++i)
        if (ranges_host[i].src_word > n_payload_words || ranges_host[i].n_words > n_payload_words - ranges_host[i].src_word)
            return fail(HDT_ERR_ARG, "hdt_apply_ranges_host: a range reads beyond the payload");
    HDT_CUDA(cudaSetDevice(c->device));
    const size_t rangeBytes = (size_t(n_ranges) * sizeof(hdt_range) + 255) & ~size_t(255);
    const size_t need = rangeBytes + ((size_t(n_payload_words) * 4 + 255) & ~size_t(255));
    if (c->stagingUsed + need > c->stagingCap) {
        HDT_CUDA(cudaStreamSynchronize(c->stream));           // everything staged so far has been consumed
        c->stagingUsed = 0;
        if (need > c->stagingCap) {
            const size_t cap = need * 2 > (size_t(8) << 20) ? need * 2 : (size_t(8) << 20);
            if (c->stagingHost) cudaFreeHost(c->stagingHost);
            cudaFree(c->stagingDev);
            c->stagingHost = nullptr; c->stagingDev = nullptr; c->stagingCap = 0;
            HDT_CUDA(cudaMallocHost(&c->stagingHost, cap));
            HDT_CUDA(cudaMalloc(&c->stagingDev, cap));
            c->stagingCap = cap;
        }
    }
    char* h = c->stagingHost + c->stagingUsed;
    char* d = c->stagingDev + c->stagingUsed;
    memcpy(h, ranges_host, size_t(n_ranges) * sizeof(hdt_range));
    memcpy(h + rangeBytes, payload_host, size_t(n_payload_words) * 4);
    HDT_CUDA(cudaMemcpyAsync(d, h, need, cudaMemcpyHostToDevice, c->stream));
    c->stagingUsed += need;
    apply_ranges_kernel<<<n_ranges < 1184 ? n_ranges : 1184, 128, 0, c->stream>>>(dst_dev, reinterpret_cast<const u32*>(d + rangeBytes),
                                                                                 reinterpret_cast<const hdt_range*>(d), n_ranges);
    ++c->launches;
    HDT_CUDA(cudaGetLastError());
    return HDT_OK;
}

int hdt_hash_dag_resolve(hdt_ctx* c, const hdt_hash_dag* dag, size_t dag_size, uint32_t* resolved_pool_dev, uint32_t* prefix_pool_dev, uint64_t capacity_words,
                         const hdt_range* ranges_host, uint32_t n_ranges)
{
    if (!c || !dag || !resolved_pool_dev) return fail(HDT_ERR_ARG, "hdt_hash_dag_resolve: null argument");
    if (dag_size != sizeof(hdt_hash_dag)) return fail(HDT_ERR_POD_SIZE, "HashDAG: expected 32 bytes");
    if (!dag->pool || !dag->page_table) return fail(HDT_ERR_ARG, "HashDAG: null pool / page table");
    const u64 poolWords = u64(dag->pool_top) * kPageWords;
    if (poolWords > (u64(1) << 32)) return fail(HDT_ERR_ARG, "HashDAG: pool beyond 2^32 words");
    if (capacity_words < poolWords) return fail(HDT_ERR_CAPACITY, "hdt_hash_dag_resolve: resolved pool smaller than pool_top pages");
    if (!ranges_host && n_ranges) return fail(HDT_ERR_ARG, "hdt_hash_dag_resolve: null ranges");
    if (!dag->pool_top) return HDT_OK;
    // pages to refresh: all of them, or those the spans touch
    std::vector<u32> pages;
    if (ranges_host) {
        for (u32 i = 0; i < n_ranges; ++i) {
            const hdt_range& r = ranges_host[i];
            if (!r.n_words) continue;
            if (r.dst_word + r.n_words > poolWords) return fail(HDT_ERR_ARG, "hdt_hash_dag_resolve: a range lies beyond pool_top");
            for (u64 p = r.dst_word / kPageWords; p <= (r.dst_word + r.n_words - 1) / kPageWords; ++p)
                if (pages.empty() || pages.back() != u32(p)) pages.push_back(u32(p));
        }
        if (pages.empty()) return HDT_OK;
    }
    HDT_CUDA(cudaSetDevice(c->device));
    if (c->physToVirtPages < dag->pool_top) {
        HDT_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->physToVirt); c->physToVirt = nullptr; c->physToVirtPages = 0;
        const size_t n = size_t(dag->pool_top) + size_t(dag->pool_top) / 4 + 1024;
        HDT_CUDA(cudaMalloc(&c->physToVirt, n * sizeof(u32)));
        c->physToVirtPages = n;
    }
    HDT_CUDA(cudaMemsetAsync(c->physToVirt, 0xFF, size_t(dag->pool_top) * sizeof(u32), c->stream));
    map_pages_kernel<<<(dag->page_table_size + 255) / 256, 256, 0, c->stream>>>(dag->page_table, dag->page_table_size, c->physToVirt, dag->pool_top);
    HDT_LAUNCHED("map_pages_kernel");
    const u32* dPages = nullptr;
    u32 nPages = dag->pool_top;
    if (ranges_host) {
        // the page list travels through the same pinned staging as hdt_apply_ranges_host
        const size_t need = (pages.size() * sizeof(u32) + 255) & ~size_t(255);
        if (c->stagingUsed + need > c->stagingCap) {
            HDT_CUDA(cudaStreamSynchronize(c->stream));
            c->stagingUsed = 0;
            if (need > c->stagingCap) {
                const size_t cap = need * 2 > (size_t(8) << 20) ? need * 2 : (size_t(8) << 20);
                if (c->stagingHost) cudaFreeHost(c->stagingHost);
                cudaFree(c->stagingDev);
                c->stagingHost = nullptr; c->stagingDev = nullptr; c->stagingCap = 0;
                HDT_CUDA(cudaMallocHost(&c->stagingHost, cap));
                HDT_CUDA(cudaMalloc(&c->stagingDev, cap));
                c->stagingCap = cap;
            }
        }
        memcpy(c->stagingHost + c->stagingUsed, pages.data(), pages.size() * sizeof(u32));
        HDT_CUDA(cudaMemcpyAsync(c->stagingDev + c->stagingUsed, c->stagingHost + c->stagingUsed, need, cudaMemcpyHostToDevice, c->stream));
        dPages = reinterpret_cast<const u32*>(c->stagingDev + c->stagingUsed);
        c->stagingUsed += need;
        nPages = u32(pages.size());
    }
    resolve_pages_kernel<<<(nPages + 3) / 4, 128, 0, c->stream>>>(dag->pool, dag->page_table, c->physToVirt, dPages, nPages,
                                                                  HashLayoutDev{ c->levels, dag->page_table_size, dag->pool_top }, resolved_pool_dev, prefix_pool_dev);
    HDT_LAUNCHED("resolve_pages_kernel");
    return HDT_OK;
}

int hdt_rebuild_color_leaf(hdt_ctx* c, const hdt_color_leaf* old_leaf, size_t old_leaf_size, const hdt_color_op* ops, uint64_t n_ops,
                           uint32_t* weights_out, uint64_t weights_capacity, uint64_t* blocks_out, uint64_t blocks_capacity,
                           uint64_t* macro_blocks_out, uint64_t macro_blocks_capacity, uint64_t counts_out[4], float* ms)
{
    if (!c || (!ops && n_ops) || !counts_out) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: null argument");
    if (old_leaf && old_leaf_size != sizeof(hdt_color_leaf)) return fail(HDT_ERR_POD_SIZE, "CompressedColorLeaf: expected 104 bytes");
    if (ms) *ms = 0.f;
    // op list -> device form: empty ops dropped, exclusive prefix of the counts, one sentinel
    std::vector<ColorOpDev> dev;
    dev.reserve(n_ops + 1);
    u64 n = 0, piecesBound = 0;       // pieces: runs of one old block / one FILL op (cut again at the new macro blocks: + one per segment)
    bool copies = false;
    for (u64 i = 0; i < n_ops; ++i) {
        const hdt_color_op& o = ops[i];
        if (o.kind != HDT_COLOR_OP_COPY && o.kind != HDT_COLOR_OP_FILL) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: unknown op kind");
        if (o.kind == HDT_COLOR_OP_FILL && (o.bits_per_weight > 4 || (o.bits_per_weight && o.weight >> o.bits_per_weight) || (!o.bits_per_weight && o.weight)))
            return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: weight does not fit bits_per_weight (0..4)");
        if (!o.count) continue;
        if (o.kind == HDT_COLOR_OP_COPY) {
            // the old leaf addresses whole macro blocks (the length of its last block is not stored)
            const u64 have = old_leaf ? (old_leaf->macro_blocks_gpu.size / 2) * kColorsPerMacroBlock : 0;
            const u64 off = (old_leaf && old_leaf->offset != kUniqueOffset) ? old_leaf->offset : 0;
            if (off > have || o.src_start > have - off || o.count > have - off - o.src_start)
                return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: COPY op reaches beyond the old leaf's macro blocks");
        }
        copies |= o.kind == HDT_COLOR_OP_COPY;
        piecesBound += o.kind == HDT_COLOR_OP_FILL ? 1 : std::min<u64>(o.count, old_leaf ? old_leaf->blocks_gpu.size : 0);
        dev.push_back(ColorOpDev{ n, o.src_start, o.kind, o.bits_per_weight, o.color_bits, o.weight });
        n += o.count;
    }
    counts_out[0] = n; counts_out[1] = counts_out[2] = counts_out[3] = 0;
    if (n == 0) return HDT_OK;
    if (n >= (u64(1) << 30)) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: more colours than the reference's 32-bit weight offsets can address");
    if (dev.size() >= (u64(1) << 32) - 1) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: too many ops");
    ColorLeafDev leaf{};
    if (copies) {
        if (!old_leaf || !old_leaf->blocks_gpu.data || !old_leaf->macro_blocks_gpu.data) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: COPY ops need the old leaf");
        leaf = leaf_dev(*old_leaf);
    }
    dev.push_back(ColorOpDev{ n, 0, HDT_COLOR_OP_FILL, 0, 0, 0 });
    const u32 nTiles = u32((n + kColorsPerMacroBlock - 1) / kColorsPerMacroBlock);
    HDT_CUDA(cudaSetDevice(c->device));
    // scratch: [ops][per macro block: pieces, offsets, segments][totals, group sums, stage top][segments, their weight ranges][staged block entries]
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t nSlots = dev.size() + nTiles;               // segment slots (color_segments_kernel)
    const u64 stageEntries = piecesBound + nSlots;           // a macro block stages at most one block entry per piece
    if (stageEntries >= (u64(1) << 32)) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: too many pieces");
    const size_t offOps = 0, offTiles = align(dev.size() * sizeof(ColorOpDev));
    const size_t offOffsets = offTiles + align(size_t(nTiles) * sizeof(TilePieces)), offTileSegs = offOffsets + align((size_t(nTiles) + 1) * sizeof(ulonglong2));
    const size_t offTotals = offTileSegs + align(size_t(nTiles) * sizeof(TileSegments)), offGroups = offTotals + 256, offStageTop = offGroups + 256 * sizeof(u64);
    const size_t offSegs = offStageTop + 256, offSegWeights = offSegs + align(nSlots * sizeof(SegmentDev)), offStage = offSegWeights + align(nSlots * sizeof(SegmentWeights));
    const size_t need = offStage + align(stageEntries * sizeof(u64));
    if (need > c->rebuildScratchBytes) {
        HDT_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->rebuildScratch); c->rebuildScratch = nullptr; c->rebuildScratchBytes = 0;
        HDT_CUDA(cudaMalloc(&c->rebuildScratch, need + need / 8));
        c->rebuildScratchBytes = need + need / 8;
    }
    char* base = static_cast<char*>(c->rebuildScratch);
    ColorOpDev* dOps = reinterpret_cast<ColorOpDev*>(base + offOps);
    TilePieces* dTiles = reinterpret_cast<TilePieces*>(base + offTiles);
    ulonglong2* dOffsets = reinterpret_cast<ulonglong2*>(base + offOffsets);
    TileSegments* dTileSegs = reinterpret_cast<TileSegments*>(base + offTileSegs);
    u64* dTotals = reinterpret_cast<u64*>(base + offTotals);
    u64* dGroups = reinterpret_cast<u64*>(base + offGroups);     // sums of {blocks, bits} over groups of 256 macro blocks
    u32* dStageTop = reinterpret_cast<u32*>(base + offStageTop);
    SegmentDev* dSegs = reinterpret_cast<SegmentDev*>(base + offSegs);
    SegmentWeights* dSegWeights = reinterpret_cast<SegmentWeights*>(base + offSegWeights);
    u64* dStage = reinterpret_cast<u64*>(base + offStage);
    HDT_CUDA(cudaMemcpyAsync(dOps, dev.data(), dev.size() * sizeof(ColorOpDev), cudaMemcpyHostToDevice, c->stream));
    HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    HDT_CUDA(cudaMemsetAsync(dGroups, 0, 256 * sizeof(u64) + 256, c->stream));   // group sums and the stage top
    color_segments_kernel<<<(nTiles + 7) / 8, 256, 0, c->stream>>>(dOps, u32(dev.size() - 1), leaf, n, nTiles, dSegs, dTileSegs);
    color_pieces_kernel<<<nTiles, kPieceThreads, 0, c->stream>>>(dSegs, dTileSegs, leaf, dTiles, dGroups, dStage, dStageTop, dSegWeights);
    scan_color_tiles_kernel<<<(nTiles + kTileGroup - 1) / kTileGroup, kTileGroup, 0, c->stream>>>(dTiles, nTiles, dGroups, dOffsets, dTotals, weights_out, weights_out ? weights_capacity : 0);
    HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->launches += 3;
    u64 totals[2] = { 0, 0 };
    HDT_CUDA(cudaMemcpyAsync(totals, dTotals, sizeof(totals), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    if (totals[1] >= (u64(1) << 32)) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: weight stream beyond the reference's 32-bit bit offsets");
    counts_out[1] = (totals[1] + 31) / 32; counts_out[2] = totals[0]; counts_out[3] = 2 * u64(nTiles);
    if (counts_out[1] > weights_capacity || counts_out[2] > blocks_capacity || counts_out[3] > macro_blocks_capacity)
        return fail(HDT_ERR_CAPACITY, "hdt_rebuild_color_leaf: an output buffer is too small (see counts_out)");
    if ((counts_out[1] && !weights_out) || !blocks_out || !macro_blocks_out) return fail(HDT_ERR_ARG, "hdt_rebuild_color_leaf: null output buffer");
    HDT_CUDA(cudaEventRecord(c->ev[2], c->stream));
    color_emit_kernel<<<nTiles, kEmitThreads, 0, c->stream>>>(dSegWeights, dTileSegs, dTiles, dStage, dOffsets, leaf, weights_out, blocks_out, macro_blocks_out);
    HDT_CUDA(cudaEventRecord(c->ev[3], c->stream));
    ++c->launches;
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    if (ms) {
        float a = 0.f, b = 0.f;
        HDT_CUDA(cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
        HDT_CUDA(cudaEventElapsedTime(&b, c->ev[2], c->ev[3]));
        *ms = a + b;
    }
    return HDT_OK;
}

int hdt_find_or_add(hdt_ctx* c, hdt_hash_table* table, uint32_t level, int leaves, const uint32_t* words_dev, const uint64_t* offsets_dev, uint32_t n_nodes,
                    uint32_t* ptrs_out_dev, uint32_t counts_out[2])
{
    if (!c || !table || (!words_dev && n_nodes) || (!offsets_dev && n_nodes) || (!ptrs_out_dev && n_nodes)) return fail(HDT_ERR_ARG, "hdt_find_or_add: null argument");
    if (!table->pool || !table->page_table || !table->bucket_sizes) return fail(HDT_ERR_ARG, "hdt_find_or_add: hash table without pool / page table / bucket sizes");
    if (level + 2 > table->levels || (leaves != 0) != (level + 2 == table->levels)) return fail(HDT_ERR_ARG, "hdt_find_or_add: leaves live at level levels-2, interior nodes above it");
    if (counts_out) counts_out[0] = counts_out[1] = 0;
    if (!n_nodes) return HDT_OK;
    if (HashTableDev::bucket_global_index(level, HashTableDev::buckets_per_level(level) - 1) >= table->n_buckets) return fail(HDT_ERR_ARG, "hdt_find_or_add: bucket_sizes shorter than the level's buckets");
    if (HashTableDev::make_ptr(level, HashTableDev::buckets_per_level(level) - 1, HashTableDev::bucket_capacity(level) - 1) / kPageWords >= table->page_table_size)
        return fail(HDT_ERR_ARG, "hdt_find_or_add: page table shorter than the level's buckets");
    HDT_CUDA(cudaSetDevice(c->device));
    HashTableDev t{ table->pool, table->page_table, table->bucket_sizes, table->page_table_size, table->levels, table->pool_top, table->pool_capacity_words };
    // scratch: [keys in][keys out][isStart][segFirst][ptr flags: added, openedPage, pageRank, segBucket, segSize][staging 17 n][totals][cub temp]
    const size_t n = n_nodes;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t cubSort = 0, cubSelect = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cubSort, static_cast<u64*>(nullptr), static_cast<u64*>(nullptr), int(n), 0, 64, c->stream);
    cub::DeviceSelect::Flagged(nullptr, cubSelect, thrust::counting_iterator<u32>(0), static_cast<u32*>(nullptr), static_cast<u32*>(nullptr), static_cast<u32*>(nullptr), int(n), c->stream);
    const size_t oKeys = 0, oKeys2 = oKeys + align(n * 8), oStart = oKeys2 + align(n * 8), oSeg = oStart + align(n * 4), oAdded = oSeg + align(n * 4);
    const size_t oOpened = oAdded + align(n * 4), oRank = oOpened + align(n * 4), oSegBucket = oRank + align(n * 4), oSegSize = oSegBucket + align(n * 4);
    const size_t oStage = oSegSize + align(n * 4), oTotals = oStage + align(n * 17 * 4), oCub = oTotals + 256, need = oCub + align(std::max(cubSort, cubSelect));
    if (need > c->foaScratchBytes) {
        HDT_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->foaScratch); c->foaScratch = nullptr; c->foaScratchBytes = 0;
        HDT_CUDA(cudaMalloc(&c->foaScratch, need + need / 2));
        c->foaScratchBytes = need + need / 2;
    }
    char* base = static_cast<char*>(c->foaScratch);
    u64* keys = reinterpret_cast<u64*>(base + oKeys);
    u64* keys2 = reinterpret_cast<u64*>(base + oKeys2);
    u32* isStart = reinterpret_cast<u32*>(base + oStart);
    u32* segFirst = reinterpret_cast<u32*>(base + oSeg);
    u32* pageRank = reinterpret_cast<u32*>(base + oRank);
    u32* staging = reinterpret_cast<u32*>(base + oStage);
    u32* totals = reinterpret_cast<u32*>(base + oTotals);   // [0] pages opened, [1] nodes added, [2] segments, [3] errors
    FoaOut out{ ptrs_out_dev, reinterpret_cast<u32*>(base + oAdded), reinterpret_cast<u32*>(base + oOpened), totals + 3,
                reinterpret_cast<u32*>(base + oSegBucket), reinterpret_cast<u32*>(base + oSegSize) };
    const u32 blocks = u32((n + 255) / 256);
    HDT_CUDA(cudaMemsetAsync(totals, 0, 256, c->stream));
    HDT_CUDA(cudaMemsetAsync(staging, 0, n * 17 * 4, c->stream));
    foa_hash_kernel<<<blocks, 256, 0, c->stream>>>(words_dev, offsets_dev, n_nodes, level, leaves != 0, keys, totals + 3);
    HDT_LAUNCHED("foa_hash_kernel");
    size_t tmp = cubSort;
    HDT_CUDA(cub::DeviceRadixSort::SortKeys(base + oCub, tmp, keys, keys2, int(n), 0, 64, c->stream));
    foa_segments_kernel<<<blocks, 256, 0, c->stream>>>(keys2, n_nodes, isStart);
    HDT_LAUNCHED("foa_segments_kernel");
    tmp = cubSelect;
    HDT_CUDA(cub::DeviceSelect::Flagged(base + oCub, tmp, thrust::counting_iterator<u32>(0), isStart, segFirst, totals + 2, int(n), c->stream));
    u32 host[4] = { 0, 0, 0, 0 };
    HDT_CUDA(cudaMemcpyAsync(host, totals, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    if (host[3] & 2) return fail(HDT_ERR_ARG, "hdt_find_or_add: a candidate is not a node (leaf: 2 words; interior: 1 + popc(child mask) words)");
    const u32 nSeg = host[2];
    foa_bucket_kernel<<<(nSeg + 63) / 64, 64, 0, c->stream>>>(t, level, leaves != 0, words_dev, offsets_dev, keys2, n_nodes, segFirst, nSeg, staging, out);
    HDT_LAUNCHED("foa_bucket_kernel");
    foa_pages_kernel<<<1, 1024, 0, c->stream>>>(out.openedPage, n_nodes, pageRank, totals);
    HDT_LAUNCHED("foa_pages_kernel");
    HDT_CUDA(cudaMemcpyAsync(host, totals, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    if (host[3]) return fail(HDT_ERR_CAPACITY, "hdt_find_or_add: a bucket would overflow (hash_table.h:461); nothing was inserted");
    if ((u64(table->pool_top) + host[0]) * kPageWords > table->pool_capacity_words) return fail(HDT_ERR_CAPACITY, "hdt_find_or_add: the pool has no room for the pages the batch opens; nothing was inserted");
    foa_apply_kernel<<<blocks, 256, 0, c->stream>>>(t, n_nodes, nSeg, out, pageRank);
    HDT_LAUNCHED("foa_apply_kernel");
    foa_commit_kernel<<<blocks, 256, 0, c->stream>>>(t, words_dev, offsets_dev, n_nodes, out, totals);
    HDT_LAUNCHED("foa_commit_kernel");
    HDT_CUDA(cudaMemcpyAsync(host, totals, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    HDT_CUDA(cudaGetLastError());
    table->pool_top += host[0];
    if (counts_out) { counts_out[0] = host[1]; counts_out[1] = host[0]; }
    c->ancValid = false;
    return HDT_OK;
}

int hdt_get_values(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, const uint32_t start[3], const uint32_t size[3],
                   uint8_t* values_dev, float* ms)
{
    if (!c || !start || !size) return fail(HDT_ERR_ARG, "hdt_get_values: null argument");
    if (ms) *ms = 0.f;
    DagArg dag;
    if (int rc = parse_dag(dag_kind, dag_pod, dag_pod_size, dag)) return rc;
    RegionParams rp;
    for (int a = 0; a < 3; ++a) {
        rp.start[a] = start[a]; rp.size[a] = size[a];
        if (!size[a]) return HDT_OK;                       // nothing to write
        if (u64(start[a]) + size[a] > (u64(1) << c->levels)) return fail(HDT_ERR_ARG, "hdt_get_values: region beyond the DAG's 2^levels voxels");
        rp.cell0[a] = start[a] >> 2;
        rp.nCells[a] = ((start[a] + size[a] - 1) >> 2) - rp.cell0[a] + 1;
    }
    if (!values_dev) return fail(HDT_ERR_ARG, "hdt_get_values: null output");
    const u64 nCells = u64(rp.nCells[0]) * rp.nCells[1] * rp.nCells[2];
    if ((nCells + 127) / 128 > 0x7FFFFFFFull) return fail(HDT_ERR_ARG, "hdt_get_values: region too large for one launch");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    const u32 grid = u32((nCells + 127) / 128);
    if (dag.kind == HDT_DAG_BASIC) get_values_kernel<<<grid, 128, 0, c->stream>>>(dag.basic, c->levels, rp, values_dev);
    else if (dag.kind == HDT_DAG_HASH) get_values_kernel<<<grid, 128, 0, c->stream>>>(dag.hash, c->levels, rp, values_dev);
    else get_values_kernel<<<grid, 128, 0, c->stream>>>(dag.resolved, c->levels, rp, values_dev);
    HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    ++c->launches;
    return finish_timed(c, c->ev[0], c->ev[1], ms);
}

int hdt_is_empty(hdt_ctx* c, int dag_kind, const void* dag_pod, size_t dag_pod_size, uint32_t max_level, const uint32_t start[3],
                 const uint32_t size[3], int* empty, float* ms)
{
    if (!c || !start || !size || !empty) return fail(HDT_ERR_ARG, "hdt_is_empty: null argument");
    if (ms) *ms = 0.f;
    if (max_level > c->levels - 2) return fail(HDT_ERR_ARG, "hdt_is_empty: max_level must be <= levels - 2 (dag_utils.h:264)");
    DagArg dag;
    if (int rc = parse_dag(dag_kind, dag_pod, dag_pod_size, dag)) return rc;
    if (max_level == 0) { *empty = 0; return HDT_OK; }    // the recursion returns false at the root (dag_utils.h:185-188)
    // nodes of level max_level-1 are cells of 2^shift voxels; candidates: bmin < start+size and bmin + 2^shift - 1 > start
    const u32 steps = max_level - 1, shift = c->levels - steps;
    RegionParams rp;
    for (int a = 0; a < 3; ++a) {
        rp.start[a] = start[a]; rp.size[a] = size[a];
        const u64 inMax = u64(start[a]) + size[a];
        if (inMax > (u64(1) << c->levels)) return fail(HDT_ERR_ARG, "hdt_is_empty: region beyond the DAG's 2^levels voxels");
        if (inMax == 0) { *empty = 1; return HDT_OK; }
        const u64 lo = (u64(start[a]) + 1) >> shift, hi = (inMax - 1) >> shift;
        if (lo > hi || lo >= (u64(1) << steps)) { *empty = 1; return HDT_OK; }
        rp.cell0[a] = u32(lo);
        rp.nCells[a] = u32(hi - lo + 1);
    }
    const u64 nCells = u64(rp.nCells[0]) * rp.nCells[1] * rp.nCells[2];
    if ((nCells + 127) / 128 > 0x7FFFFFFFull) return fail(HDT_ERR_ARG, "hdt_is_empty: region too large for one launch");
    HDT_CUDA(cudaSetDevice(c->device));
    u32* found = reinterpret_cast<u32*>(c->hitCounter);
    HDT_CUDA(cudaMemsetAsync(found, 0, sizeof(u32), c->stream));
    HDT_CUDA(cudaEventRecord(c->ev[0], c->stream));
    const u32 grid = u32((nCells + 127) / 128);
    if (dag.kind == HDT_DAG_BASIC) is_empty_kernel<<<grid, 128, 0, c->stream>>>(dag.basic, steps, rp, found);
    else if (dag.kind == HDT_DAG_HASH) is_empty_kernel<<<grid, 128, 0, c->stream>>>(dag.hash, steps, rp, found);
    else is_empty_kernel<<<grid, 128, 0, c->stream>>>(dag.resolved, steps, rp, found);
    HDT_CUDA(cudaEventRecord(c->ev[1], c->stream));
    ++c->launches;
    u32 host = 0;
    HDT_CUDA(cudaMemcpyAsync(&host, found, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = finish_timed(c, c->ev[0], c->ev[1], ms)) return rc;
    *empty = host ? 0 : 1;
    return HDT_OK;
}

}  // extern "C"

#include "hdt_multi.cuh"
