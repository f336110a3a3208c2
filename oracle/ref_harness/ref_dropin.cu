// Compile-time and run-time proof that hashdag_b200's C++ shim is a drop-in for the reference's DAGTracer
// (/root/reference/src/dag_tracer.h:9-43) at its call sites in Engine::tick (engine.cpp:575-648).
// TEST INFRASTRUCTURE: built by oracle/build_ref.py into the oracle/_ref harness libraries.
//
// This translation unit includes the reference's OWN headers, untouched (no `#define private public` here), declares
// which kernel family each reference type selects, aliases DAGTracer to the shim the way INTEGRATION.md tells a
// maintainer to, and then makes the calls of Engine::tick with the reference's real BasicDAG / HashDAG / *Colors /
// CameraView / DAGInfo / ToolInfo / EDebugColors objects.  The structs cross into libhashdag_b200.so as bytes, exactly
// as the reference hands them to its kernels by value; the static_asserts pin the layouts the library mirrors.
#include "ref_harness_std.h"

#include "tracer.h"
#include "camera_view.h"
#include "dag_info.h"
#include "dags/basic_dag/basic_dag.h"
#include "dags/hash_dag/hash_dag.h"
#include "dags/hash_dag/hash_dag_colors.h"
#include "ref_harness_globals.h"

#include "dag_tracer_b200.h"

HDT_DECLARE_DAG(BasicDAG, HDT_DAG_BASIC)
HDT_DECLARE_DAG(HashDAG, HDT_DAG_HASH)
HDT_DECLARE_COLORS(BasicDAGUncompressedColors, HDT_COLORS_UNCOMPRESSED)
HDT_DECLARE_COLORS(BasicDAGCompressedColors, HDT_COLORS_COMPRESSED)
HDT_DECLARE_COLORS(BasicDAGColorErrors, HDT_COLORS_ERRORS)
HDT_DECLARE_COLORS(HashDAGColors, HDT_COLORS_HASH)
using DAGTracer = hashdag_b200::DAGTracer;   // engine.cpp keeps compiling unchanged

static_assert(sizeof(BasicDAG) == sizeof(hdt_basic_dag), "BasicDAG (basic_dag.h:11-13)");
static_assert(sizeof(HashDAG) == sizeof(hdt_hash_dag), "HashDAG (hash_dag.h:214-221, hash_table.h:818-826)");
static_assert(sizeof(CompressedColorLeaf) == sizeof(hdt_color_leaf), "CompressedColorLeaf (variable_weight_size_colors.h:157-191)");
static_assert(sizeof(BasicDAGCompressedColors) == sizeof(hdt_basic_compressed_colors), "BasicDAGCompressedColors (basic_dag.h:91-95)");
static_assert(sizeof(BasicDAGUncompressedColors) == sizeof(hdt_basic_uncompressed_colors), "BasicDAGUncompressedColors (basic_dag.h:122-177)");
static_assert(sizeof(BasicDAGColorErrors) == sizeof(hdt_basic_color_errors), "BasicDAGColorErrors (basic_dag.h:179-203)");
static_assert(sizeof(HashDAGColors) == sizeof(hdt_hash_colors), "HashDAGColors (hash_dag_colors.h:65-73)");
static_assert(sizeof(ToolInfo) == sizeof(hdt_tool_info), "ToolInfo (tracer.h:33-39)");
static_assert(offsetof(hdt_hash_dag, page_table) == 8 && offsetof(hdt_hash_dag, pool) == 16 && offsetof(hdt_hash_dag, first_node_index) == 24, "HashDAG field offsets");
static_assert(int(EDebugColors::None) == HDT_DEBUG_NONE && int(EDebugColors::Index) == HDT_DEBUG_INDEX && int(EDebugColors::Position) == HDT_DEBUG_POSITION &&
              int(EDebugColors::ColorTree) == HDT_DEBUG_COLOR_TREE && int(EDebugColors::ColorBits) == HDT_DEBUG_COLOR_BITS &&
              int(EDebugColors::MinColor) == HDT_DEBUG_MIN_COLOR && int(EDebugColors::MaxColor) == HDT_DEBUG_MAX_COLOR &&
              int(EDebugColors::Weight) == HDT_DEBUG_WEIGHT, "EDebugColors (tracer.h:7-17)");

namespace {
std::unique_ptr<DAGTracer> tracer;   // Engine::tracer (engine.cpp:765)

// engine.cpp's EDag (engine.h) selects the branch of each switch below
enum class EDag { BasicDagUncompressedColors = 0, BasicDagCompressedColors = 1, BasicDagColorErrors = 2, HashDag = 3 };
}  // namespace

extern "C" {

int ref_dropin_init(int device)
{
    // headLess = false so that get_path answers (dag_tracer.cu:244 returns 0 when head-less); the shim has no GL path either way
    tracer = std::make_unique<DAGTracer>(false, imageWidth, imageHeight, MAX_LEVELS, device);
    return 0;
}
int ref_dropin_shutdown()
{
    tracer.reset();   // Engine::destroy (engine.cpp:1134)
    return 0;
}

// One Engine::tick worth of tracer calls (engine.cpp:575-648), statement for statement, on the harness' scene.
// times = { pathsTime, colorsTime, shadowsTime }; path = what engine.cpp stores in config.path.
int ref_dropin_tick(int currentDag, const double pos[3], const double rot[9], const double bmin[3], const double bmax[3], int debugColors,
                    uint32_t debugColorsIndexLevel, int shadows, float shadowBias, float fogDensity, uint32_t posX, uint32_t posY, double times[3], uint32_t path[3])
{
    using namespace refh;
    CameraView view;
    view.position = Vector3(pos[0], pos[1], pos[2]);
    view.rotation = Matrix3x3(rot[0], rot[1], rot[2], rot[3], rot[4], rot[5], rot[6], rot[7], rot[8]);
    DAGInfo dagInfo;
    dagInfo.boundsAABBMin = Vector3(bmin[0], bmin[1], bmin[2]);
    dagInfo.boundsAABBMax = Vector3(bmax[0], bmax[1], bmax[2]);
    const BasicDAG& basicDag = g_basic;
    const HashDAG& hashDag = g_hash;
    const BasicDAGUncompressedColors& basicDagUncompressedColors = g_uncompressed;
    const BasicDAGCompressedColors& basicDagCompressedColors = g_compressed;
    const HashDAGColors& hashDagColors = g_hashColors;
    BasicDAGColorErrors basicDagColorErrors = g_errors;
    basicDagColorErrors.compressedColors = g_compressed;
    basicDagColorErrors.uncompressedColors = g_uncompressed;
    struct { EDag currentDag; uint3 path; EDebugColors debugColors; ETool tool; float radius; uint3 copySourcePath, copyDestPath; } config{};
    config.currentDag = EDag(currentDag);
    config.debugColors = EDebugColors(debugColors);
    config.tool = ETool::Sphere;
    config.radius = 10.f;

    double pathsTime = 0;
    switch (config.currentDag)
    {
    case EDag::BasicDagUncompressedColors:
    case EDag::BasicDagCompressedColors:
    case EDag::BasicDagColorErrors:
        pathsTime = tracer->resolve_paths(view, dagInfo, basicDag);
        break;
    case EDag::HashDag:
        pathsTime = tracer->resolve_paths(view, dagInfo, hashDag);
        break;
    }

    config.path = tracer->get_path(posX, posY);

    double colorsTime = 0;
    const ToolInfo toolInfo
    {
        config.tool,
        config.path,
        config.radius,
        config.copySourcePath,
        config.copyDestPath
    };
    switch (config.currentDag)
    {
    case EDag::BasicDagUncompressedColors:
        colorsTime = tracer->resolve_colors(basicDag, basicDagUncompressedColors, config.debugColors,
                                            debugColorsIndexLevel, toolInfo);
        break;
    case EDag::BasicDagCompressedColors:
        colorsTime = tracer->resolve_colors(basicDag, basicDagCompressedColors, config.debugColors, debugColorsIndexLevel,
                                            toolInfo);
        break;
    case EDag::BasicDagColorErrors:
        colorsTime = tracer->resolve_colors(basicDag, basicDagColorErrors, config.debugColors,
                                            debugColorsIndexLevel, toolInfo);
        break;
    case EDag::HashDag:
        colorsTime = tracer->resolve_colors(hashDag, hashDagColors, config.debugColors, debugColorsIndexLevel, toolInfo);
        break;
    }

    double shadowsTime = 0;
    if (shadows)
    {
        switch (config.currentDag)
        {
            case EDag::BasicDagUncompressedColors:
            case EDag::BasicDagCompressedColors:
            case EDag::BasicDagColorErrors:
                shadowsTime = tracer->resolve_shadows(view, dagInfo, basicDag, shadowBias, fogDensity);
                break;
            case EDag::HashDag:
                shadowsTime = tracer->resolve_shadows(view, dagInfo, hashDag, shadowBias, fogDensity);
                break;
        }
    }
    times[0] = pathsTime; times[1] = colorsTime; times[2] = shadowsTime;
    path[0] = config.path.x; path[1] = config.path.y; path[2] = config.path.z;
    return 0;
}

int ref_dropin_read_paths(uint32_t* out) { tracer->read_paths(out); return 0; }
int ref_dropin_read_colors(uint32_t* out) { tracer->read_colors(out); return 0; }

}  // extern "C"
