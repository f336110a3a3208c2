// Synthetic voxel-scene builder: procedural terrain shell + sphere shells -> the
// reference's on-disk / in-memory DAG formats.  Host-only C++ (std::thread); it feeds
// the tests and bench.py, it is not on the traced path.
//
// Formats produced (layouts documented in DESIGN.md §3, all cite the reference):
//   * BasicDAG word array            (/root/reference/src/dags/basic_dag/basic_dag.h:11-45)
//   * enclosedLeaves + topLevels     (basic_dag.h:47-76)
//   * HashDAG pool + page table      (hash_table.h:45-71,156-173,359-468; built the way
//                                     hash_dag_factory.cpp:5-99,152-184 builds it)
//   * compressed colour leaf         (variable_weight_size_colors.h:157-191,557-721)
//   * HashDAGColors nodes + offsets  (hash_dag_factory.cpp:101-150, hash_dag_colors.h:65-73)
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hds_scene hds_scene;

typedef struct hds_params {
    uint32_t levels;          // DAG depth L: world is (2^L)^3 voxels
    uint32_t footprint_log2;  // terrain covers a (2^F)^2 footprint centred in x/z (F <= L)
    uint32_t seed;
    uint32_t n_spheres;       // floating sphere shells above the terrain (shadow casters)
    uint32_t roughness;       // 0: smooth, 1: +-1 voxel hash noise on the height
    uint32_t n_threads;       // 0: hardware_concurrency
    uint32_t build_hash;      // also build HashDAG pool/page table
    uint32_t build_colors;    // also build compressed colours (+ hash colour tree if possible)
    uint32_t build_uncompressed; // also emit RGB888-per-voxel array (small scenes only)
    uint32_t finest_cell_log2;  // finest noise octave cell (default 2 -> 4 voxels)
} hds_params;

typedef struct hds_info {
    uint32_t levels;
    uint32_t top_levels;            // BasicDAGColorsBase::topLevels
    uint64_t n_voxels;
    uint64_t basic_words;
    uint64_t enclosed_leaves;
    // HashDAG
    uint32_t hash_page_table_size;  // entries (C_totalPages for this depth)
    uint32_t hash_pool_top;         // pages in use (incl. the unused page 0)
    uint32_t hash_first_node_index;
    uint32_t has_hash_colors;       // 1 if levels-2 > 10 (hash_dag_factory.cpp:113)
    // colours
    uint64_t n_weight_words, n_blocks, n_macro_words;
    uint64_t n_color_nodes, n_color_offsets;
    uint64_t n_uncompressed;
    double   build_seconds;
    uint64_t nodes_per_level[32];
} hds_info;

hds_scene* hds_build(const hds_params* p);
void hds_free(hds_scene* s);
void hds_get_info(const hds_scene* s, hds_info* out);

const uint32_t* hds_basic_data(const hds_scene* s);
const uint64_t* hds_enclosed_leaves(const hds_scene* s);
const uint32_t* hds_hash_pool(const hds_scene* s);        // hash_pool_top * 512 words
const uint32_t* hds_hash_page_table(const hds_scene* s);
const uint32_t* hds_hash_bucket_sizes(const hds_scene* s);   /* hds_hash_bucket_count() fill counts, hash_table.h:18-35 order */
uint64_t hds_hash_bucket_count(const hds_scene* s);
const uint32_t* hds_color_weights(const hds_scene* s);
const uint64_t* hds_color_blocks(const hds_scene* s);
const uint64_t* hds_color_macro_blocks(const hds_scene* s);
const uint32_t* hds_color_uncompressed(const hds_scene* s);
const uint32_t* hds_hash_color_nodes(const hds_scene* s);
const uint64_t* hds_hash_color_offsets(const hds_scene* s);

// A point on the terrain surface (for camera placement): height at (x,z).
int32_t hds_terrain_height(const hds_scene* s, int32_t x, int32_t z);

#ifdef __cplusplus
}
#endif
