// C++ host shim: the reference's DAGTracer interface (/root/reference/src/dag_tracer.h:9-43) over
// the C ABI of libhashdag_b200.so (include/hashdag_b200.h).  Header-only, no CUDA headers needed.
//
// Drop-in use inside the reference tree:
//     #include "dags/basic_dag/basic_dag.h"          // the reference's own types
//     #include "dags/hash_dag/hash_dag.h"
//     #include "dags/hash_dag/hash_dag_colors.h"
//     #include "dag_tracer_b200.h"
//     HDT_DECLARE_DAG(BasicDAG, HDT_DAG_BASIC)        // tell the shim which kernel family a type selects
//     HDT_DECLARE_DAG(HashDAG, HDT_DAG_HASH)
//     HDT_DECLARE_COLORS(BasicDAGUncompressedColors, HDT_COLORS_UNCOMPRESSED)
//     HDT_DECLARE_COLORS(BasicDAGCompressedColors, HDT_COLORS_COMPRESSED)
//     HDT_DECLARE_COLORS(BasicDAGColorErrors, HDT_COLORS_ERRORS)
//     HDT_DECLARE_COLORS(HashDAGColors, HDT_COLORS_HASH)
//     using DAGTracer = hashdag_b200::DAGTracer;      // engine.cpp keeps compiling unchanged
//
// The structs are handed to the library exactly as the reference hands them to its kernels: by
// value, as bytes.  Same method names, argument order and meaning, same return value (kernel time
// in milliseconds), same error behaviour (print + abort, cuda_error_check.h:42-51).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/hashdag_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace hashdag_b200 {

template <class T> struct dag_kind;      // ::value = HDT_DAG_*
template <class T> struct colors_kind;   // ::value = HDT_COLORS_*

#define HDT_DECLARE_DAG(Type, Kind) \
    namespace hashdag_b200 { template <> struct dag_kind<Type> { static constexpr int value = Kind; }; }
#define HDT_DECLARE_COLORS(Type, Kind) \
    namespace hashdag_b200 { template <> struct colors_kind<Type> { static constexpr int value = Kind; }; }

template <> struct dag_kind<hdt_basic_dag> { static constexpr int value = HDT_DAG_BASIC; };
template <> struct dag_kind<hdt_hash_dag> { static constexpr int value = HDT_DAG_HASH; };
template <> struct colors_kind<hdt_basic_uncompressed_colors> { static constexpr int value = HDT_COLORS_UNCOMPRESSED; };
template <> struct colors_kind<hdt_basic_compressed_colors> { static constexpr int value = HDT_COLORS_COMPRESSED; };
template <> struct colors_kind<hdt_basic_color_errors> { static constexpr int value = HDT_COLORS_ERRORS; };
template <> struct colors_kind<hdt_hash_colors> { static constexpr int value = HDT_COLORS_HASH; };

// Stand-ins for gmath's Vector3 / Matrix3x3 and for camera_view.h / dag_info.h, for hosts that do
// not have the reference headers.  The shim's templates accept the reference's own types as well.
struct Vector3 { double X, Y, Z; };
struct CameraView {
    static constexpr float fov = 60.f;   // camera_view.h:10
    Vector3 position{ 0, 0, 0 };
    double rotation[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    Vector3 right() const { return { -rotation[0][0], -rotation[0][1], -rotation[0][2] }; }   // camera_view.h:15-19
    Vector3 up() const { return { rotation[1][0], rotation[1][1], rotation[1][2] }; }
    Vector3 forward() const { return { rotation[2][0], rotation[2][1], rotation[2][2] }; }
};
struct DAGInfo { Vector3 boundsAABBMin, boundsAABBMax; };
// What get_path returns: converts to CUDA's uint3 (or anything with x, y, z) so that `config.path = tracer->get_path(..)`
// (engine.cpp:597) compiles unchanged, without this header needing the CUDA headers.
struct uint3_t {
    uint32_t x, y, z;
    template <class T, class = decltype(T{}.x, T{}.y, T{}.z)>
    operator T() const { T t{}; t.x = x; t.y = y; t.z = z; return t; }
};

// The reference decides at compile time whether trace_colors draws the tool overlay (TOOL_OVERLAY, typedefs.h:70-72;
// off under BENCHMARK); a host that defines the macro gets the same behaviour from the shim by default.
#if defined(TOOL_OVERLAY)
#define HDT_SHIM_TOOL_OVERLAY (TOOL_OVERLAY != 0)
#else
#define HDT_SHIM_TOOL_OVERLAY false
#endif

struct TraceParams { double cam[3], rayMin[3], rayDDx[3], rayDDy[3]; };

// get_trace_params, dag_tracer.cu:71-113: same operations in the same order, in double.
template <class TCamera, class TInfo>
inline TraceParams get_trace_params(const TCamera& camera, uint32_t levels, const TInfo& dagInfo, uint32_t imageWidth, uint32_t imageHeight)
{
    const auto p = camera.position; const auto d = camera.forward(); const auto u = camera.up(); const auto r = camera.right();
    const double position[3] = { double(p.X), double(p.Y), double(p.Z) }, direction[3] = { double(d.X), double(d.Y), double(d.Z) };
    const double up[3] = { double(u.X), double(u.Y), double(u.Z) }, right[3] = { double(r.X), double(r.Y), double(r.Z) };
    const double bmin[3] = { double(dagInfo.boundsAABBMin.X), double(dagInfo.boundsAABBMin.Y), double(dagInfo.boundsAABBMin.Z) };
    const double bmax[3] = { double(dagInfo.boundsAABBMax.X), double(dagInfo.boundsAABBMax.Y), double(dagInfo.boundsAABBMax.Z) };
    const double fov = double(camera.fov) / 2.0 * (double(M_PI) / 180.);
    const double aspect_ratio = double(imageWidth) / double(imageHeight);
    const double s = std::sin(fov), c = std::cos(fov);
    TraceParams out;
    for (int k = 0; k < 3; ++k) {
        const double X = right[k] * s * aspect_ratio, Y = up[k] * s, Z = direction[k] * c;
        const double bottomLeft = position[k] + Z - Y - X, bottomRight = position[k] + Z - Y + X, topLeft = position[k] + Z + Y - X;
        const double translation = -bmin[k];
        const double scale = double(1 << levels) / (bmax[k] - bmin[k]);
        const double finalPosition = (position[k] + translation) * scale, finalBottomLeft = (bottomLeft + translation) * scale;
        const double finalTopLeft = (topLeft + translation) * scale, finalBottomRight = (bottomRight + translation) * scale;
        out.cam[k] = finalPosition;
        out.rayMin[k] = finalBottomLeft;
        out.rayDDx[k] = (finalBottomRight - finalBottomLeft) * (1.0 / imageWidth);
        out.rayDDy[k] = (finalTopLeft - finalBottomLeft) * (1.0 / imageHeight);
    }
    return out;
}

class DAGTracer {
public:
    const bool headLess;

    // The reference fixes resolution and depth at compile time (typedefs.h:517,683-684); here they
    // are constructor arguments defaulting to the reference's values.
    explicit DAGTracer(bool headLess, uint32_t imageWidth = 1920, uint32_t imageHeight = 1080, uint32_t levels = 17, int device = 0)
        : headLess(headLess), width_(imageWidth), height_(imageHeight), levels_(levels)
    {
        require_ok(hdt_create(imageWidth, imageHeight, levels, device, &ctx_), "hdt_create");
    }
    ~DAGTracer() { hdt_destroy(ctx_); }
    DAGTracer(const DAGTracer&) = delete;
    DAGTracer& operator=(const DAGTracer&) = delete;

    // There is no GL path: the colour frame lives in a linear device buffer (hdt_partition_buffers).
    inline unsigned get_colors_image() const { return 0; }

    template <typename TDAG, typename TCamera, typename TInfo>
    float resolve_paths(const TCamera& camera, const TInfo& dagInfo, const TDAG& dag)
    {
        const TraceParams p = get_trace_params(camera, levels_, dagInfo, width_, height_);
        float ms = 0;
        require_ok(hdt_resolve_paths(ctx_, dag_kind<TDAG>::value, &dag, sizeof(TDAG), p.cam, p.rayMin, p.rayDDx, p.rayDDy, &ms), "resolve_paths");
        return ms;
    }

    template <typename TDAG, typename TDAGColors, typename TDebugColors, typename TToolInfo>
    float resolve_colors(const TDAG& dag, const TDAGColors& colors, TDebugColors debugColors, uint32_t debugColorsIndexLevel, TToolInfo toolInfo,
                         bool toolOverlay = HDT_SHIM_TOOL_OVERLAY)
    {
        static_assert(sizeof(TToolInfo) == sizeof(hdt_tool_info), "ToolInfo layout (tracer.h:33-39)");
        float ms = 0;
        require_ok(hdt_resolve_colors(ctx_, dag_kind<TDAG>::value, &dag, sizeof(TDAG), colors_kind<TDAGColors>::value, &colors, sizeof(TDAGColors),
                                 int(debugColors), debugColorsIndexLevel, reinterpret_cast<const hdt_tool_info*>(&toolInfo), toolOverlay ? 1 : 0, &ms),
              "resolve_colors");
        return ms;
    }

    template <typename TDAG, typename TCamera, typename TInfo>
    float resolve_shadows(const TCamera& camera, const TInfo& dagInfo, const TDAG& dag, float shadowBias, float fogDensity)
    {
        const TraceParams p = get_trace_params(camera, levels_, dagInfo, width_, height_);
        float ms = 0;
        require_ok(hdt_resolve_shadows(ctx_, dag_kind<TDAG>::value, &dag, sizeof(TDAG), p.cam, p.rayMin, p.rayDDx, p.rayDDy, shadowBias, fogDensity, &ms),
              "resolve_shadows");
        return ms;
    }

    uint3_t get_path(uint32_t posX, uint32_t posY)
    {
        uint32_t v[3] = { 0, 0, 0 };
        if (headLess) return { 0, 0, 0 };   // dag_tracer.cu:244
        require_ok(hdt_get_path(ctx_, posX, posY, v), "get_path");
        return { v[0], v[1], v[2] };
    }

    // Additions the reference lacks: full-frame read-back (its harness reads the cudaArrays).
    void read_paths(uint32_t* host) { require_ok(hdt_read_paths(ctx_, host), "read_paths"); }
    void read_colors(uint32_t* host) { require_ok(hdt_read_colors(ctx_, host), "read_colors"); }
    hdt_ctx* context() { return ctx_; }

private:
    // (not called `check`: the reference defines a macro of that name, typedefs.h)
    static void require_ok(int rc, const char* what)
    {
        if (rc != HDT_OK) {   // the reference prints and aborts on any CUDA error
            std::fprintf(stderr, "ERROR hashdag_b200 %s: %d: %s\n", what, rc, hdt_last_error());
            std::abort();
        }
    }
    hdt_ctx* ctx_ = nullptr;
    uint32_t width_, height_, levels_;
};

}  // namespace hashdag_b200
