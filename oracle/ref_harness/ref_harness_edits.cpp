// The reference's CPU edit code behind the harness (host C++: see ref_harness_shared.h for why it is a separate file).
// TEST INFRASTRUCTURE: built by oracle/build_ref.py into oracle/_ref/, never linked into the product.
#include "ref_harness_shared.h"
#define private public
#include "dags/hash_dag/hash_dag_editors.h"
#undef private

using namespace refh;

namespace {
HashDAGUndoRedo g_undoRedo;
StatsRecorder g_statsRecorder;
double g_lastEditMs = 0, g_lastUploadMs = 0;   // harness stopwatch around the reference's two calls
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" {

// The reference's own CPU edit of the HashDAG, as Engine::edit drives it (engine.h:78-103): build the
// editor, HashDAG::edit_threads (hash_dag.h:255-406), HashTable::upload_to_gpu (hash_table.cpp:120-184).
int ref_edit_sphere(float x, float y, float z, float radius, int adding)
{
    if (!g_hasHash || !g_hasHashColors) return 1;
    const double t0 = now_ms();
    if (adding) {
        const auto tool = SphereEditor<true>(make_float3(x, y, z), radius);
        g_hash.edit_threads(tool, g_hashColors, g_undoRedo, g_statsRecorder);
    } else {
        const auto tool = SphereEditor<false>(make_float3(x, y, z), radius);
        g_hash.edit_threads(tool, g_hashColors, g_undoRedo, g_statsRecorder);
    }
    const double t1 = now_ms();
    g_hash.data.upload_to_gpu();
    cudaDeviceSynchronize();
    g_lastEditMs = t1 - t0; g_lastUploadMs = now_ms() - t1;
    return 0;
}
// Host wall time of the last ref_edit_sphere: [0] HashDAG::edit_threads (includes its colour-leaf rebuilds and their
// uploads, hash_dag.h:384-398), [1] HashTable::upload_to_gpu (hash_table.cpp:120-184).
void ref_last_edit_ms(double* out) { out[0] = g_lastEditMs; out[1] = g_lastUploadMs; }

// Unique colour leaves created by edits (HashDAGColors::leaves, hash_dag_colors.h:65-73).
uint64_t ref_color_leaf_count() { return g_hasHashColors ? g_hashColors.leaves_CPU.size() : 0; }
int ref_color_leaf_info(uint64_t i, uint64_t* nWeights, uint64_t* nBlocks, uint64_t* nMacro)
{
    if (!g_hasHashColors || i >= g_hashColors.leaves_CPU.size()) return 1;
    const CompressedColorLeaf& l = g_hashColors.leaves_CPU[i];
    *nWeights = l.weights_CPU.size(); *nBlocks = l.blocks_CPU.size(); *nMacro = l.macroBlocks_CPU.size();
    return 0;
}
int ref_color_leaf_copy(uint64_t i, uint32_t* weights, uint64_t* blocks, uint64_t* macro)
{
    if (!g_hasHashColors || i >= g_hashColors.leaves_CPU.size()) return 1;
    const CompressedColorLeaf& l = g_hashColors.leaves_CPU[i];
    if (l.weights_CPU.size()) std::memcpy(weights, l.weights_CPU.data(), l.weights_CPU.size() * sizeof(uint32));
    if (l.blocks_CPU.size()) std::memcpy(blocks, l.blocks_CPU.data(), l.blocks_CPU.size() * sizeof(uint64));
    if (l.macroBlocks_CPU.size()) std::memcpy(macro, l.macroBlocks_CPU.data(), l.macroBlocks_CPU.size() * sizeof(uint64));
    return 0;
}

// The copy tool's read-only walks over the (possibly edited) HashDAG, on the host like the reference runs them
// (hash_dag_editors.h:401, :476, :570): DAGUtils::get_values<5> (dag_utils.h:363-411) and DAGUtils::is_empty (:255-266).
int ref_get_values(const uint32_t* start, const uint32_t* size, uint8_t* values)
{
    if (!g_hasHash) return 1;
    static_assert(sizeof(bool) == 1, "values are bytes");
    DAGUtils::get_values<5>(g_hash, reinterpret_cast<bool*>(values), make_uint3(start[0], start[1], start[2]), make_uint3(size[0], size[1], size[2]));
    return 0;
}
int ref_is_empty(uint32_t maxLevel, const uint32_t* start, const uint32_t* size)
{
    if (!g_hasHash) return -1;
    return DAGUtils::is_empty(g_hash, maxLevel, make_uint3(start[0], start[1], start[2]), make_uint3(size[0], size[1], size[2])) ? 1 : 0;
}


// The reference's own insert, one node after the other (HashTable::find_or_add_interior_node / find_or_add_leaf_node,
// hash_table.h:470-560), on its host table: node i = words[offsets[i] .. offsets[i+1]); the checker of hdt_find_or_add.
int ref_find_or_add(uint32_t level, int leaves, const uint32_t* words, const uint64_t* offsets, uint32_t n, uint32_t* ptrs)
{
    if (!g_hasHash) return 1;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t* w = words + offsets[i];
        if (leaves) ptrs[i] = g_hash.data.find_or_add_leaf_node(level, uint64_t(w[0]) | (uint64_t(w[1]) << 32));
        else ptrs[i] = g_hash.data.find_or_add_interior_node(level, uint32_t(offsets[i + 1] - offsets[i]), w);
    }
    return 0;
}
}  // extern "C"
