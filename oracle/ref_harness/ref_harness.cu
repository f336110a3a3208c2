// Harness around the UNMODIFIED reference tracer, compiled against the reference's own headers.
// TEST INFRASTRUCTURE: built by oracle/build_ref.py into oracle/_ref/ (git-ignored), never linked
// into the product.  It feeds the reference's DAGTracer / HashDAGFactory from memory buffers and
// reads its frames back, so the oracle and the product can be compared with the real kernels.
//
// Depth and resolution are compile-time in the reference (typedefs.h:517,683-684), hence one
// library per (depth, width, height).
#include "ref_harness_shared.h"

namespace refh {
BasicDAG g_basic;
BasicDAGCompressedColors g_compressed;
BasicDAGUncompressedColors g_uncompressed;
BasicDAGColorErrors g_errors;
HashDAG g_hash;
HashDAGColors g_hashColors;
bool g_hasHash = false, g_hasHashColors = false;
}
using namespace refh;

namespace {
std::unique_ptr<DAGTracer> g_tracer;

CameraView make_camera(const double pos[3], const double rot[9])
{
    CameraView v;
    v.position = Vector3(pos[0], pos[1], pos[2]);
    v.rotation = Matrix3x3(rot[0], rot[1], rot[2], rot[3], rot[4], rot[5], rot[6], rot[7], rot[8]);
    return v;
}
DAGInfo make_info(const double bmin[3], const double bmax[3])
{
    DAGInfo i;
    i.boundsAABBMin = Vector3(bmin[0], bmin[1], bmin[2]);
    i.boundsAABBMax = Vector3(bmax[0], bmax[1], bmax[2]);
    return i;
}
}  // namespace

extern "C" {

int ref_levels() { return MAX_LEVELS; }
int ref_width() { return imageWidth; }
int ref_height() { return imageHeight; }

int ref_init(int device)
{
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    g_tracer = std::make_unique<DAGTracer>(true);
    return 0;
}

// Free everything the reference allocated, the way main.cpp does before returning, so that its
// leak check (Memory::~Memory, memory.cpp:145-157: abort() on leaks) stays quiet at process exit.
int ref_shutdown()
{
    if (!g_tracer) return 0;
    cudaDeviceSynchronize();
    if (g_hasHashColors) { g_hashColors.free(); g_hasHashColors = false; }
    if (g_hasHash) { g_hash.free(); g_hasHash = false; }
    if (g_compressed.is_valid()) g_compressed.free();
    else if (g_compressed.enclosedLeaves.is_valid()) g_compressed.enclosedLeaves.free();
    if (g_uncompressed.is_valid()) g_uncompressed.free();
    if (g_basic.is_valid()) g_basic.free();
    g_tracer.reset();
    return 0;
}

int ref_set_basic_dag(const uint32_t* words, uint64_t n)
{
    // like BasicDAGFactory::load_dag_from_file: managed memory (basic_dag.cpp:104-117)
    g_basic.data = StaticArray<uint32>::allocate("basic dag nodes", n, EMemoryType::GPU_Managed);
    std::memcpy(g_basic.data.data(), words, n * sizeof(uint32));
    return 0;
}

int ref_set_compressed_colors(uint32_t topLevels, const uint64_t* enclosed, uint64_t nEnclosed, const uint32_t* weights, uint64_t nWeights,
                              const uint64_t* blocks, uint64_t nBlocks, const uint64_t* macro, uint64_t nMacro)
{
    // like load_compressed_colors_from_file (basic_dag.cpp:55-89); weights arrive already in the
    // in-memory (byte-swapped) order
    g_compressed.topLevels = topLevels;
    g_compressed.enclosedLeaves = StaticArray<EnclosedLeavesType>::allocate("basic dag enclosed leaves", nEnclosed, EMemoryType::GPU_Managed);
    std::memcpy(g_compressed.enclosedLeaves.data(), enclosed, nEnclosed * sizeof(uint64));
    g_compressed.leaf.weights_CPU = StaticArray<uint32>::allocate("basic dag weights", nWeights, EMemoryType::CPU);
    std::memcpy(g_compressed.leaf.weights_CPU.data(), weights, nWeights * sizeof(uint32));
    g_compressed.leaf.blocks_CPU = StaticArray<uint64>::allocate("basic dag blocks", nBlocks, EMemoryType::CPU);
    std::memcpy(g_compressed.leaf.blocks_CPU.data(), blocks, nBlocks * sizeof(uint64));
    g_compressed.leaf.macroBlocks_CPU = StaticArray<uint64>::allocate("basic dag macro blocks", nMacro, EMemoryType::CPU);
    std::memcpy(g_compressed.leaf.macroBlocks_CPU.data(), macro, nMacro * sizeof(uint64));
    g_compressed.leaf.set_as_unique();
    g_compressed.leaf.upload_to_gpu();
    return 0;
}

int ref_set_uncompressed_colors(uint32_t topLevels, const uint64_t* enclosed, uint64_t nEnclosed, const uint32_t* colors, uint64_t n)
{
    g_uncompressed.topLevels = topLevels;
    g_uncompressed.enclosedLeaves = StaticArray<EnclosedLeavesType>::allocate("basic dag enclosed leaves", nEnclosed, EMemoryType::GPU_Managed);
    std::memcpy(g_uncompressed.enclosedLeaves.data(), enclosed, nEnclosed * sizeof(uint64));
    g_uncompressed.leaf.colors = StaticArray<uint32>::allocate("basic dag uncompressed colors", n, EMemoryType::GPU_Managed);
    std::memcpy(g_uncompressed.leaf.colors.data(), colors, n * sizeof(uint32));
    return 0;
}

int ref_build_hash_dag(uint32_t poolPages, int withColors)
{
    HashDAGFactory::load_from_DAG(g_hash, g_basic, poolPages);
    g_hasHash = true;
    if (withColors) {
        HashDAGFactory::load_colors_from_DAG(g_hashColors, g_basic, g_compressed);
        // Under BENCHMARK the factory "reserves" nodes_GPU/leaves_GPU right after the first upload
        // (hash_dag_colors.h:244-250); that goes through Memory::realloc_impl, whose GPU branch calls
        // cuda_memcpy_impl(..., cudaMemcpyDefault) -- a kind that function silently ignores
        // (memory.cpp:159-188, :268-273) -- so the device copy of the colour tree is lost.  The engine
        // repairs it with the upload it issues after every edit (hash_dag.h:398); do the same here.
        g_hashColors.upload_to_gpu(false);
        g_hasHashColors = true;
    }
    return 0;
}

int ref_hash_info(uint32_t* firstNode, uint32_t* poolTop, uint32_t* pageTableSize)
{
    if (!g_hasHash) return 1;
    *firstNode = g_hash.firstNodeIndex; *poolTop = g_hash.data.poolTop; *pageTableSize = g_hash.data.pageTableSize;
    return 0;
}
int ref_hash_copy(uint32_t* pool, uint32_t* pageTable)
{
    if (!g_hasHash) return 1;
    std::memcpy(pool, HashTable::cpuData.cpuPool, size_t(g_hash.data.poolTop) * C_pageSize * sizeof(uint32));
    std::memcpy(pageTable, HashTable::cpuData.cpuPageTable, size_t(g_hash.data.pageTableSize) * sizeof(uint32));
    return 0;
}
// Zero-copy access for the host-side dirty tracker (hashdag_b200/edits.py): the reference's own host arrays and the
// per-bucket fill counts its upload_to_gpu walks (hash_table.cpp:160-183).
int ref_hash_pointers(uint64_t* pool, uint64_t* pageTable, uint64_t* bucketSizes, uint32_t* nBuckets)
{
    if (!g_hasHash) return 1;
    *pool = reinterpret_cast<uint64_t>(HashTable::cpuData.cpuPool);
    *pageTable = reinterpret_cast<uint64_t>(HashTable::cpuData.cpuPageTable);
    *bucketSizes = reinterpret_cast<uint64_t>(HashTable::cpuData.bucketsSizes);
    *nBuckets = C_totalNumberOfBuckets;
    return 0;
}
int ref_hash_colors_info(uint64_t* nNodes, uint64_t* nOffsets)
{
    if (!g_hasHashColors) return 1;
    *nNodes = g_hashColors.nodes_CPU.size(); *nOffsets = g_hashColors.offsets_CPU.size();
    return 0;
}
int ref_hash_colors_copy(uint32_t* nodes, uint64_t* offsets)
{
    if (!g_hasHashColors) return 1;
    std::memcpy(nodes, g_hashColors.nodes_CPU.data(), g_hashColors.nodes_CPU.size() * sizeof(uint32));
    std::memcpy(offsets, g_hashColors.offsets_CPU.data(), g_hashColors.offsets_CPU.size() * sizeof(uint64));
    return 0;
}

float ref_resolve_paths(int dagKind, const double pos[3], const double rot[9], const double bmin[3], const double bmax[3])
{
    const CameraView v = make_camera(pos, rot);
    const DAGInfo info = make_info(bmin, bmax);
    return dagKind == 0 ? g_tracer->resolve_paths(v, info, g_basic) : g_tracer->resolve_paths(v, info, g_hash);
}

float ref_resolve_colors(int dagKind, int colorsKind, int debugColors, uint32_t debugLevel)
{
    const ToolInfo tool{};
    const EDebugColors dbg = EDebugColors(debugColors);
    if (dagKind == 1) return g_tracer->resolve_colors(g_hash, g_hashColors, dbg, debugLevel, tool);
    if (colorsKind == 0) return g_tracer->resolve_colors(g_basic, g_uncompressed, dbg, debugLevel, tool);
    if (colorsKind == 1) return g_tracer->resolve_colors(g_basic, g_compressed, dbg, debugLevel, tool);
    g_errors.compressedColors = g_compressed;
    g_errors.uncompressedColors = g_uncompressed;
    return g_tracer->resolve_colors(g_basic, g_errors, dbg, debugLevel, tool);
}

// Tracer::trace_colors with a tool overlay (tracer.cu:276-286; compiled in only when TOOL_OVERLAY is set,
// i.e. in the "_overlay" variant of this library).  toolKind: ETool as int, then ToolInfo's fields.
int ref_tool_overlay_compiled() { return TOOL_OVERLAY ? 1 : 0; }
float ref_resolve_colors_tool(int dagKind, int colorsKind, int debugColors, uint32_t debugLevel, int toolKind, const uint32_t position[3], float radius,
                              const uint32_t copySource[3], const uint32_t copyDest[3])
{
    const ToolInfo tool(ETool(toolKind), make_uint3(position[0], position[1], position[2]), radius,
                        make_uint3(copySource[0], copySource[1], copySource[2]), make_uint3(copyDest[0], copyDest[1], copyDest[2]));
    const EDebugColors dbg = EDebugColors(debugColors);
    if (dagKind == 1) return g_tracer->resolve_colors(g_hash, g_hashColors, dbg, debugLevel, tool);
    if (colorsKind == 0) return g_tracer->resolve_colors(g_basic, g_uncompressed, dbg, debugLevel, tool);
    return g_tracer->resolve_colors(g_basic, g_compressed, dbg, debugLevel, tool);
}

float ref_resolve_shadows(int dagKind, const double pos[3], const double rot[9], const double bmin[3], const double bmax[3], float bias, float fog)
{
    const CameraView v = make_camera(pos, rot);
    const DAGInfo info = make_info(bmin, bmax);
    return dagKind == 0 ? g_tracer->resolve_shadows(v, info, g_basic, bias, fog) : g_tracer->resolve_shadows(v, info, g_hash, bias, fog);
}

int ref_read_paths(uint32_t* out)
{
    return cudaMemcpy2DFromArray(out, imageWidth * 16, g_tracer->pathArray, 0, 0, imageWidth * 16, imageHeight, cudaMemcpyDeviceToHost) != cudaSuccess;
}
int ref_read_colors(uint32_t* out)
{
    return cudaMemcpy2DFromArray(out, imageWidth * 4, g_tracer->colorsArray, 0, 0, imageWidth * 4, imageHeight, cudaMemcpyDeviceToHost) != cudaSuccess;
}
int ref_write_colors(const uint32_t* in)
{
    return cudaMemcpy2DToArray(g_tracer->colorsArray, 0, 0, in, imageWidth * 4, imageWidth * 4, imageHeight, cudaMemcpyHostToDevice) != cudaSuccess;
}

}  // extern "C"
