/* hashdag_b200 -- C ABI of the B200-native DAG tracer.
 *
 * Drop-in boundary for the reference's DAGTracer (/root/reference/src/dag_tracer.h:9-43,
 * dag_tracer.cu:116-256): the C++ shim in hashdag_b200/cpp/dag_tracer_b200.h keeps the
 * reference's class shape and forwards to these entry points.  Plain pointers and sizes only.
 *
 * DAG and colour structures cross the boundary as the reference passes them to its kernels: by
 * value, as the bytes of the struct (const void* + sizeof).  The hdt_* POD types below mirror the
 * reference layouts; every entry point checks the size it is handed.
 *
 * All functions return 0 on success, a non-zero code otherwise (CUDA errors are returned as
 * 1000 + cudaError_t); hdt_last_error() describes the last failure on the calling thread.
 * All resolve_* calls are synchronous like the reference's (dag_tracer.cu:130-133): they return
 * after the kernel has finished and report its device time in milliseconds.
 */
#ifndef HASHDAG_B200_H
#define HASHDAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hdt_ctx hdt_ctx;

enum { HDT_DAG_BASIC = 0, HDT_DAG_HASH = 1,
       HDT_DAG_HASH_RESOLVED = 2 /* hdt_resolved_hash_dag: a HashDAG + its resolved pool (hdt_hash_dag_resolve); same results, shorter load chains */ };
enum {
    HDT_COLORS_UNCOMPRESSED = 0, /* BasicDAGUncompressedColors, basic_dag.h:122-177 */
    HDT_COLORS_COMPRESSED = 1,   /* BasicDAGCompressedColors,   basic_dag.h:91-120  */
    HDT_COLORS_ERRORS = 2,       /* BasicDAGColorErrors,        basic_dag.h:179-242 */
    HDT_COLORS_HASH = 3          /* HashDAGColors,              hash_dag_colors.h:9-73 */
};
/* EDebugColors, tracer.h:7-17 */
enum { HDT_DEBUG_NONE = 0, HDT_DEBUG_INDEX, HDT_DEBUG_POSITION, HDT_DEBUG_COLOR_TREE, HDT_DEBUG_COLOR_BITS,
       HDT_DEBUG_MIN_COLOR, HDT_DEBUG_MAX_COLOR, HDT_DEBUG_WEIGHT };

enum { HDT_OK = 0, HDT_ERR_ARG = 1, HDT_ERR_POD_SIZE = 2, HDT_ERR_STATE = 3, HDT_ERR_CAPACITY = 4, HDT_ERR_NCCL = 5, HDT_ERR_CUDA = 1000 };

/* ---- POD mirrors (device pointers unless the field says _cpu) ------------------------------ */
typedef struct hdt_array { const void* data; uint64_t size; } hdt_array;                       /* StaticArray<T>,  array.h:8-129  */
typedef struct hdt_dyn_array { const void* data; uint64_t size; uint64_t allocated; } hdt_dyn_array; /* DynamicArray<T>, array.h:131-227 */

typedef struct hdt_basic_dag { hdt_array data; } hdt_basic_dag;                                /* BasicDAG, basic_dag.h:11-13: 16 B */

typedef struct hdt_hash_dag {                                                                  /* HashDAG, hash_dag.h:214-221 + hash_table.h:818-826: 32 B */
    uint32_t page_table_size;
    uint32_t pool_top;
    const uint32_t* page_table;   /* gpuPageTable */
    const uint32_t* pool;         /* gpuPool */
    uint32_t first_node_index;
    uint32_t _pad;
} hdt_hash_dag;

typedef struct hdt_resolved_hash_dag {      /* library format, 48 B: the HashDAG as the reference passes it + a copy of its pool in which every */
    hdt_hash_dag dag;                        /* child pointer already holds the child's physical word index (hdt_hash_dag_resolve)               */
    const uint32_t* resolved_pool;           /* device, pool_top * 512 words, caller-owned */
    const uint32_t* prefix_pool;             /* device, same size, caller-owned, or NULL: per child-pointer word of the levels below the colour
                                                tree, the voxels under the node's earlier children (what trace_colors otherwise adds up by
                                                loading every preceding sibling, tracer.cu:391-420) */
} hdt_resolved_hash_dag;

typedef struct hdt_color_leaf {                                                                /* CompressedColorLeaf, vwsc.h:157-191: 104 B */
    uint64_t offset;              /* offset into the shared leaf; UINT64_MAX = unique */
    hdt_array weights_gpu, blocks_gpu, macro_blocks_gpu;
    hdt_array weights_cpu, blocks_cpu, macro_blocks_cpu;   /* ignored by the tracer */
} hdt_color_leaf;

typedef struct hdt_basic_colors_base { uint32_t top_levels; uint32_t _pad; hdt_array enclosed_leaves; } hdt_basic_colors_base; /* basic_dag.h:47-51: 24 B */
typedef struct hdt_basic_compressed_colors { hdt_basic_colors_base base; hdt_color_leaf leaf; } hdt_basic_compressed_colors;   /* 128 B */
typedef struct hdt_basic_uncompressed_colors { hdt_basic_colors_base base; hdt_array colors; } hdt_basic_uncompressed_colors;   /* 40 B */
typedef struct hdt_basic_color_errors {                                                        /* basic_dag.h:179-203: 288 B */
    hdt_color_leaf leaf_compressed; hdt_array leaf_uncompressed;   /* the unused `leaf` member */
    hdt_basic_compressed_colors compressed; hdt_basic_uncompressed_colors uncompressed;
} hdt_basic_color_errors;
typedef struct hdt_hash_colors {                                                               /* HashDAGColors, hash_dag_colors.h:65-73: 248 B */
    hdt_dyn_array nodes_gpu, leaves_gpu, offsets_gpu;
    hdt_color_leaf main_leaf;
    hdt_dyn_array nodes_cpu, leaves_cpu, offsets_cpu;      /* ignored by the tracer */
} hdt_hash_colors;

typedef struct hdt_tool_info {                                                                 /* ToolInfo, tracer.h:33-39: 44 B */
    int32_t tool; uint32_t position[3]; float radius; uint32_t copy_source[3]; uint32_t copy_dest[3];
} hdt_tool_info;

typedef struct hdt_range { uint64_t dst_word; uint64_t src_word; uint64_t n_words; } hdt_range; /* one dirty span, hash_table.cpp:129-180 */

/* One step of a colour-leaf rebuild, in voxel order: what the editor hands to ColorLeafBuilder while it walks an
 * edited colour leaf (hash_dag_edits.h:381-396, :430-437, :488-516). */
enum {
    HDT_COLOR_OP_COPY = 0,  /* CompressedColorLeaf::copy_colors (vwsc.h:416-542): `count` colours of the old leaf starting at its
                               colour `src_start` (the editor's oldLeavesCount; a shared leaf's offset is added by the library) */
    HDT_COLOR_OP_FILL = 1   /* `count` times ColorLeafBuilder::add(color) (vwsc.h:582-613); with bits_per_weight == 0 this is
                               add_large_single_color (vwsc.h:614-641) */
};
typedef struct hdt_color_op {
    uint64_t src_start; uint64_t count; uint32_t kind;
    uint32_t bits_per_weight; uint32_t color_bits; uint32_t weight;   /* CompressedColor (vwsc.h:86-90), FILL only */
} hdt_color_op;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces DAGTracer::DAGTracer(headLess=true) (dag_tracer.cu:9-45); width/height/levels are
 * runtime values here (compile-time imageWidth/imageHeight/MAX_LEVELS in typedefs.h:517,683). */
int hdt_create(uint32_t width, uint32_t height, uint32_t levels, int device, hdt_ctx** out);
int hdt_destroy(hdt_ctx* ctx);                                  /* DAGTracer::~DAGTracer, dag_tracer.cu:48-69 */
const char* hdt_last_error(void);

/* Screen-space partition for one-process-per-GPU runs: the frame is cut into (1<<tile_log2)^2
 * pixel tiles, tile t belongs to rank t % world.  Default: rank 0 of 1 (whole frame). */
int hdt_set_partition(hdt_ctx* ctx, uint32_t rank, uint32_t world, uint32_t tile_log2);

/* Tuning switches (results never depend on them).  HDT_OPT_BEAMS (default 1): run the per-tile beam
 * pre-pass of trace_paths / trace_shadows (csrc/hdt_beam.cuh) before the per-ray kernels; 0 makes
 * every ray start at the root like the reference.  The environment variable HDT_BEAMS=0|1 sets the
 * default of new contexts. */
enum {
    HDT_OPT_BEAMS = 1,
    HDT_OPT_BEAM_MAX_VISITS = 2, /* nodes a beam may visit before handing over (default 32; env HDT_BEAM_MAX_VISITS) */
    HDT_OPT_BEAM_PREFETCH = 3,   /* default 0.  1: the ray setup + beam kernels of a paths pass wait only for the previous paths
                                    kernel instead of for everything queued on the tracer's stream, so that with several frames
                                    in flight (hdt_resolve_frame_async) they run beside the previous frame's colours / shadows.
                                    Only valid while nothing queued on the tracer's stream modifies the DAG (static scene, or
                                    edits applied after hdt_sync()).  env HDT_BEAM_PREFETCH */
    HDT_OPT_BEAM_SERIAL = 4,     /* diagnostics, default 0.  1: the per-ray kernels wait for the beam kernel instead of racing it */
    HDT_OPT_EXCHANGE_FUSED = 5,  /* default 0.  1 (needs an exchange, hdt_exchange_*): every shadows pass -- the pass that writes a frame's final
                                    colours -- also stores them into rank 0's frame and its last CTA signals the arrival, so hdt_exchange_frame has
                                    nothing left to copy (a pass repeated before hdt_exchange_frame stores the same frame number again). */
    HDT_OPT_COLORS_RECORDED = 6, /* default 1 (env HDT_COLORS_RECORDED).  For a HDT_DAG_HASH_RESOLVED DAG with a prefix pool, trace_paths also
                                    records, per hit pixel, where the path leaves each ancestor below the colour tree, and trace_colors of the
                                    SAME DAG (same resolved pool, prefix pool and root) reads the voxel's colour index off those records
                                    instead of walking the DAG again.  0: always the full walk of tracer.cu:300-430. */
    HDT_OPT_L2_PERSIST = 8,      /* default 0 (env HDT_L2_PERSIST).  1: the page table of a plain HashDAG (HDT_DAG_HASH; 16 MiB at depth 17,
                                    hash_table.h:156-173) is declared persisting in L2 for the tracer's streams (cudaAccessPolicyWindow).
                                    Measured on B200: 4-5 % SLOWER (the 126 MB L2 keeps the table resident without help), hence off. */
    HDT_OPT_EXCHANGE_TIMEOUT_MS = 7 /* how long a framebuffer-exchange wait polls before it gives up and drops the frame (default 20000;
                                    0 = for ever).  A timeout is reported (HDT_ERR_STATE) by the next host-synchronising call. */
};
int hdt_set_option(hdt_ctx* ctx, int option, int value);
/* Diagnostics of the last beam pre-pass: out[0..3] = tiles that start at the root / resume below it /
 * were resolved as a common hit / as a common miss; out[4] = sum of the hand-over levels of resuming tiles. */
int hdt_beam_stats(hdt_ctx* ctx, uint64_t out[5]);
/* Device timeline of the last paths (pass 0) or shadows (pass 1) pass, in ms from the moment the pass was
 * enqueued: ms[0] = ray setup finished, ms[1] = beam kernel finished, ms[2] = per-ray kernel finished.
 * Synchronises both streams of the context. */
int hdt_pass_timeline(hdt_ctx* ctx, int pass, float ms[3]);

/* ---- the path ------------------------------------------------------------------------------ */
/* DAGTracer::resolve_paths<TDAG> (dag_tracer.cu:116-143) -> Tracer::trace_paths (tracer.cu:145-252).
 * cam/ray_min/ray_ddx/ray_ddy are TracePathsParams (tracer.h:81-91), i.e. get_trace_params' output. */
int hdt_resolve_paths(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size,
                      const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3], float* ms);

/* DAGTracer::resolve_colors<TDAG,TDAGColors> (dag_tracer.cu:145-179) -> Tracer::trace_colors
 * (tracer.cu:254-451).  tool_info may be NULL; tool_overlay mirrors the TOOL_OVERLAY build flag. */
int hdt_resolve_colors(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size,
                       int colors_kind, const void* colors_pod, size_t colors_pod_size,
                       int debug_colors, uint32_t debug_colors_index_level, const hdt_tool_info* tool_info, int tool_overlay, float* ms);

/* DAGTracer::resolve_shadows<TDAG> (dag_tracer.cu:181-219) -> Tracer::trace_shadows (tracer.cu:589-697). */
int hdt_resolve_shadows(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size,
                        const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3],
                        float shadow_bias, float fog_density, float* ms);

/* One frame = paths + colours + shadows enqueued back to back on the tracer's stream with a single
 * synchronisation, optionally followed by the colour frame's copy into host memory (pinned or not).
 * ms[3] receives the three kernel times.  Same results as the three calls above. */
int hdt_resolve_frame(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size,
                      int colors_kind, const void* colors_pod, size_t colors_pod_size,
                      const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3],
                      float shadow_bias, float fog_density, int with_shadows, uint32_t* host_colors_or_null, float ms[3]);

/* Pipelined use (SURVEY.md §8b "optional hdt_*_async + hdt_sync pair"): enqueue a frame on the
 * tracer's stream without waiting; hdt_sync() blocks until everything enqueued has finished and
 * reports the first error.  hdt_timer_begin/end bracket any number of enqueued frames with CUDA
 * events on that stream (device time, in milliseconds). */
int hdt_resolve_frame_async(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size,
                            int colors_kind, const void* colors_pod, size_t colors_pod_size,
                            const double cam[3], const double ray_min[3], const double ray_ddx[3], const double ray_ddy[3],
                            float shadow_bias, float fog_density, int with_shadows, uint32_t* host_colors_or_null);
int hdt_sync(hdt_ctx* ctx);
int hdt_timer_begin(hdt_ctx* ctx);
int hdt_timer_end(hdt_ctx* ctx, float* ms);

/* Number of pixels of the last paths frame with a non-null path (= shadow rays cast). */
int hdt_count_hits(hdt_ctx* ctx, uint64_t* n_hits);

/* DAGTracer::get_path (dag_tracer.cu:240-256): voxel under pixel (x, y) of the last paths frame. */
int hdt_get_path(hdt_ctx* ctx, uint32_t x, uint32_t y, uint32_t out_xyz[3]);

/* Full-frame read-back (the reference has none; its harness reads the cudaArrays): row-major,
 * reference orientation (paths row r = camera row height-1-r, tracer.cu:251). */
int hdt_read_paths(hdt_ctx* ctx, uint32_t* host_w_h_4);
int hdt_read_colors(hdt_ctx* ctx, uint32_t* host_w_h);

/* ---- multi-GPU plumbing (the collectives themselves are issued by the host layer) ----------- */
/* This rank's compact tile buffers (owned tiles back to back, each tile row-major). */
int hdt_partition_buffers(hdt_ctx* ctx, void** paths_dev, void** colors_dev, uint64_t* n_owned_tiles, uint64_t* max_tiles_per_rank);
/* Rank 0: scatter `world` gathered compact colour buffers (each max_tiles_per_rank tiles) into the row-major frame.
 * Asynchronous: enqueued on the tracer's stream, complete after hdt_sync(). */
int hdt_assemble_colors(hdt_ctx* ctx, const uint32_t* gathered_dev, uint32_t* frame_dev_or_null);
/* Framebuffer exchange over peer memory (NVLink), the alternative to gather + hdt_assemble_colors: rank 0's row-major
 * colour frame is mapped into every rank and each rank stores its own tiles straight into it (csrc/hdt_exchange.cuh).
 * Call after hdt_set_partition.  Rank 0: hdt_exchange_create allocates the frame (+ two counters behind it) and returns
 * its CUDA IPC handle (64 bytes) for the other processes, which call hdt_exchange_open; contexts of the SAME process
 * attach by pointer instead (hdt_exchange_block of the root -> hdt_exchange_attach).
 * Per frame, every rank calls hdt_exchange_frame after its passes (asynchronous, tracer's stream): other ranks wait for
 * rank 0's credit, scatter their tiles and signal; rank 0 scatters its own tiles and waits for world-1 arrivals, after
 * which work queued on its stream may read the frame.  Rank 0 calls hdt_exchange_release once that work is queued; it
 * publishes the credit that lets the others overwrite the frame.  A peer that never arrives makes the wait give up after
 * HDT_OPT_EXCHANGE_TIMEOUT_MS: the frame is dropped (nothing is stored or signalled for it) and the next host-synchronising
 * call (hdt_sync, hdt_resolve_*, hdt_read_*, hdt_timer_end) returns HDT_ERR_STATE. */
#define HDT_IPC_HANDLE_BYTES 64
int hdt_exchange_create(hdt_ctx* ctx, uint8_t ipc_handle_out[HDT_IPC_HANDLE_BYTES], void** frame_dev_out);
int hdt_exchange_open(hdt_ctx* ctx, const uint8_t ipc_handle[HDT_IPC_HANDLE_BYTES]);
int hdt_exchange_block(hdt_ctx* ctx, void** block_dev_out);
int hdt_exchange_attach(hdt_ctx* ctx, void* block_dev);
/* The same exchange with the frame in HOST memory: `host_block` is hdt_exchange_block_bytes() bytes of zero-initialised memory
 * every rank's process has mapped (POSIX shared memory, an mmap'ed file ...).  EVERY rank, rank 0 included, attaches it; the
 * library pins and maps it (cudaHostRegister).  hdt_exchange_frame then copies each rank's compact tile buffer into its
 * slice of the block in ONE copy-engine transfer over the rank's own PCIe link -- the frame never funnels through rank
 * 0's GPU or its single link -- followed by the rank's arrival word; rank 0's stream waits, as before, until every rank
 * has arrived: after hdt_sync() on rank 0 the host owns the frame.  Layout of the block's frame part: rank r's slice
 * starts at byte r * max_tiles_per_rank * tile_bytes and holds its owned tiles back to back, each tile row-major --
 * tile t of the screen (row-major tile numbering) is slot t / world of rank t % world, the layout of
 * hdt_partition_buffers.  hdt_exchange_release as before.  Not combinable with HDT_OPT_EXCHANGE_FUSED. */
int hdt_exchange_block_bytes(hdt_ctx* ctx, uint64_t* bytes);
int hdt_exchange_attach_host(hdt_ctx* ctx, void* host_block, uint64_t bytes);
int hdt_exchange_frame(hdt_ctx* ctx);
int hdt_exchange_release(hdt_ctx* ctx);

/* Run the tracer on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL = back to the
 * context's own stream), so that frames, the NCCL gather and the assembly queue up on one stream
 * without host synchronisation in between. */
int hdt_set_stream(hdt_ctx* ctx, void* cuda_stream);
/* Apply edit-dirtied spans to a replica: dst[range.dst_word + i] = payload[range.src_word + i]. */
int hdt_apply_ranges(hdt_ctx* ctx, uint32_t* dst_dev, const uint32_t* payload_dev, const hdt_range* ranges_dev, uint32_t n_ranges);

/* ---- next to the path: colour-leaf rebuild ------------------------------------ */
/* Replaces ColorLeafBuilder::add / add_large_single_color / build (vwsc.h:549-721) fed by
 * CompressedColorLeaf::copy_colors (vwsc.h:416-542), i.e. the re-encoding of a colour leaf an edit touched
 * (hash_dag.h:384-398).  `ops` (HOST memory) lists, in voxel order, what the new leaf is made of; `old_leaf`
 * (device arrays; may be NULL when there is no COPY op) is the leaf being replaced.  The three output arrays are
 * DEVICE buffers owned by the caller; worst-case sizes for n = sum of counts colours are n blocks, 2*ceil(n/16384)
 * macro-block words and ceil(4n/32) weight words.  counts_out = {colours, weight words, blocks, macro-block words}
 * actually produced (HOST); the arrays are bit-identical to weights_CPU / blocks_CPU / macroBlocks_CPU after
 * ColorLeafBuilder::build.  HDT_ERR_CAPACITY (counts_out filled in, nothing written) if a buffer is too small.
 * Synchronous; ms = device time of the kernels. */
int hdt_rebuild_color_leaf(hdt_ctx* ctx, const hdt_color_leaf* old_leaf, size_t old_leaf_size, const hdt_color_op* ops, uint64_t n_ops,
                           uint32_t* weights_out, uint64_t weights_capacity, uint64_t* blocks_out, uint64_t blocks_capacity,
                           uint64_t* macro_blocks_out, uint64_t macro_blocks_capacity, uint64_t counts_out[4], float* ms);

/* ---- next to the path: region queries of the copy tool -------------------------- */
/* DAGUtils::get_values (dag_utils.h:268-411): values_dev[x + size.x*(y + size.y*z)] (DEVICE, one byte per voxel, all
 * size.x*size.y*size.z of them written) = 1 iff voxel start+(x,y,z) exists and x, y, z > 0 -- the reference's box
 * test (dag_utils.h:278-296) is strict on the low side, so the three planes through `start` stay 0.
 * Synchronous; ms = device time. */
int hdt_get_values(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size, const uint32_t start[3], const uint32_t size[3],
                   uint8_t* values_dev, float* ms);
/* DAGUtils::is_empty (dag_utils.h:175-266), max_level <= levels-2 like its checkAlways: *empty = 0 iff max_level == 0 or
 * a node of level max_level-1 exists whose voxel box [bmin, bmax] has bmin < start+size and bmax > start on every axis. */
int hdt_is_empty(hdt_ctx* ctx, int dag_kind, const void* dag_pod, size_t dag_pod_size, uint32_t max_level, const uint32_t start[3],
                 const uint32_t size[3], int* empty, float* ms);

/* GPU batch insert into the hash table: HashTable::find_or_add_interior_node / find_or_add_leaf_node
 * (hash_table.h:470-560) for n_nodes candidate nodes of ONE level, given in insertion order -- node i is
 * words_dev[offsets_dev[i] .. offsets_dev[i+1]) (DEVICE; an interior node = [header][virtual child pointers], 2..9 words; a leaf
 * = its 64 bits as two words; `leaves` != 0 iff level == levels-2).  ptrs_out_dev[i] (DEVICE) receives the node's virtual pointer,
 * and pool, page table, bucket fill counts and pool_top end up exactly as if the reference had been handed the nodes one
 * after the other (same search order and node-boundary rule, same page padding, physical pages in insertion order; its
 * Bloom filter never changes a result and has no counterpart).  Pool words no node occupies are taken to be zero.
 * `table`: device arrays owned by the caller; bucket_sizes in HashDagUtils::get_bucket_global_index order (hash_table.h:18-35).
 * counts_out = {nodes added, pages opened}.  HDT_ERR_CAPACITY if a bucket or the pool would overflow: nothing is inserted.
 * Synchronous. */
typedef struct hdt_hash_table {
    uint32_t* pool; uint64_t pool_capacity_words;
    uint32_t* page_table; uint32_t page_table_size;
    uint32_t* bucket_sizes; uint32_t n_buckets;
    uint32_t pool_top;          /* in/out */
    uint32_t levels;
} hdt_hash_table;
int hdt_find_or_add(hdt_ctx* ctx, hdt_hash_table* table, uint32_t level, int leaves, const uint32_t* words_dev, const uint64_t* offsets_dev, uint32_t n_nodes,
                    uint32_t* ptrs_out_dev, uint32_t counts_out[2]);

/* hdt_apply_ranges with `ranges` and `payload` in HOST memory (pageable is fine): both are staged through pinned memory
 * owned by the context, copied and applied in stream order.  Returns once the copy and the kernel are enqueued --
 * frames enqueued afterwards see the edit, no host synchronisation is involved (hdt_sync() reports errors).
 * Replaces the page-table memcpy + one cudaMemcpyAsync per grown bucket of HashTable::upload_to_gpu
 * (hash_table.cpp:120-184) and its closing cudaDeviceSynchronize. */
int hdt_apply_ranges_host(hdt_ctx* ctx, uint32_t* dst_dev, const uint32_t* payload_host, uint64_t n_payload_words,
                          const hdt_range* ranges_host, uint32_t n_ranges);

/* Fill (ranges_host == NULL: all pool_top pages) or refresh (the pages the n_ranges pool spans of an edit touch, as
 * given to hdt_apply_ranges[_host]) `resolved_pool_dev`, a caller-owned device buffer of pool_top * 512 words
 * (capacity_words >= that): a copy of the HashDAG's pool with the same physical layout whose child pointers have been
 * pushed through the page table once -- and, if `prefix_pool_dev` is not NULL (same size), the prefix pool described at
 * hdt_resolved_hash_dag.  Pass them with the DAG as hdt_resolved_hash_dag / HDT_DAG_HASH_RESOLVED wherever a HashDAG is
 * accepted: a descent then costs two dependent loads instead of three and trace_colors no longer walks the DAG; frames
 * are identical.  Call it after the pool and page table of an edit have been applied; asynchronous (tracer's stream,
 * ordered before later frames).
 * Nothing is assumed about pool words no node occupies (page tails, unused pages): they may hold anything, the
 * reference does not clear its pool (hash_table.cpp:60-76); every index read from the pool is bounds-checked. */
int hdt_hash_dag_resolve(hdt_ctx* ctx, const hdt_hash_dag* dag, size_t dag_size, uint32_t* resolved_pool_dev, uint32_t* prefix_pool_dev_or_null,
                         uint64_t capacity_words, const hdt_range* ranges_host, uint32_t n_ranges);

/* ---- multi-GPU host layer: replicas that follow edits ------------------------------- */
/* One edit of a HashDAG as the spans it dirtied: what HashTable::upload_to_gpu (hash_table.cpp:120-184) would copy, as
 * {dst_word, src_word, n_words} records + payload for the pool (physical word indices) and the page table, plus the new
 * root and pool top (HashDAG::firstNodeIndex, HashTable::poolTop).  All pointers are HOST memory. */
typedef struct hdt_dag_delta {
    uint32_t first_node_index, pool_top;
    const hdt_range* pool_ranges;  uint32_t n_pool_ranges;  const uint32_t* pool_payload;  uint64_t n_pool_payload;
    const hdt_range* table_ranges; uint32_t n_table_ranges; const uint32_t* table_payload; uint64_t n_table_payload;
} hdt_dag_delta;

/* The dirty tracker: finds an edit's delta from the hash table's own bookkeeping -- per bucket, its fill count now
 * (HashTable::cpuData.bucketsSizes) against the count at the previous delta (the reference's lastBucketsSizes,
 * hash_table.cpp:146-183; the table only appends) -- without comparing arrays: O(#buckets).  `levels` = DAG depth;
 * bucket_sizes holds at least hdt_tracker_bucket_count() entries in HashDagUtils::get_bucket_global_index order
 * (hash_table.h:18-35).  hdt_tracker_snapshot records the state of the initial upload; hdt_tracker_delta builds the delta
 * since the previous snapshot / delta from the host pool and page table (HashTable::cpuData.cpuPool / cpuPageTable) and
 * advances the snapshot.  The arrays a delta points to belong to the tracker and stay valid until its next call.
 * A bucket that shrank (undo, garbage collection) is HDT_ERR_STATE: re-upload and take a new snapshot. */
typedef struct hdt_dirty_tracker hdt_dirty_tracker;
int hdt_tracker_create(uint32_t levels, hdt_dirty_tracker** out);
int hdt_tracker_destroy(hdt_dirty_tracker* tracker);
uint32_t hdt_tracker_bucket_count(const hdt_dirty_tracker* tracker);
int hdt_tracker_snapshot(hdt_dirty_tracker* tracker, const uint32_t* bucket_sizes, uint32_t n_buckets);
int hdt_tracker_delta(hdt_dirty_tracker* tracker, const uint32_t* bucket_sizes, uint32_t n_buckets, const uint32_t* cpu_pool, const uint32_t* cpu_page_table,
                      uint32_t first_node_index, uint32_t pool_top, hdt_dag_delta* out);

/* One process per GPU: an NCCL communicator per context.  libnccl.so.2 is loaded at run time (dlopen; HDT_NCCL_LIB overrides
 * the name) -- no link-time dependency, and world == 1 never touches it.  Rank 0 obtains the id, the host distributes its
 * 128 bytes by whatever it has (MPI, sockets, a file), every rank calls hdt_comm_init (collective). */
#define HDT_COMM_ID_BYTES 128
int hdt_comm_unique_id(uint8_t id_out[HDT_COMM_ID_BYTES]);
int hdt_comm_init(hdt_ctx* ctx, const uint8_t id[HDT_COMM_ID_BYTES], uint32_t rank, uint32_t world);
int hdt_comm_destroy(hdt_ctx* ctx);
/* Initial replication: broadcast n_bytes of device memory from `root`'s buffer into every rank's buffer,
 * on the tracer's stream (collective; asynchronous). */
int hdt_replicate(hdt_ctx* ctx, void* dev_buffer, uint64_t n_bytes, uint32_t root);

/* One GPU's copy of a HashDAG that follows edits.  Device pointers, owned by the caller; resolved_pool / prefix_pool may be
 * NULL (then only pool and page table are maintained). */
typedef struct hdt_replica {
    uint32_t* pool; uint64_t pool_capacity_words;
    uint32_t* page_table; uint32_t page_table_size;
    uint32_t first_node_index, pool_top;          /* updated by hdt_broadcast_dirty */
    uint32_t* resolved_pool; uint32_t* prefix_pool;
} hdt_replica;
/* One edit, on every rank (collective): `root` passes the edit's delta, the other ranks NULL.  The delta travels as one
 * packed buffer (a 32-byte size header first), and every rank applies the page-table and pool spans to ITS replica with
 * apply_ranges_kernel, re-derives the resolved / prefix pools of the touched pages and updates the replica's root and pool
 * top -- all enqueued on the tracer's stream, in order with the frames around it (the ranks other than root synchronise
 * their stream once, to learn the sizes).  With world == 1 (no communicator needed) it simply applies the delta.
 * Replaces, for N GPUs, HashTable::upload_to_gpu (hash_table.cpp:120-184; Engine::edit, engine.h:77-103). */
int hdt_broadcast_dirty(hdt_ctx* ctx, uint32_t root, const hdt_dag_delta* delta_or_null, hdt_replica* replica);
/* The building block, for any replicated array of 32-bit words (the HashDAGColors tree, hash_dag_colors.h:102-120): `root`
 * passes spans + payload (HOST; the others NULL / 0), every rank applies them to its dst_dev.  Collective, tracer's stream. */
int hdt_broadcast_ranges(hdt_ctx* ctx, uint32_t root, uint32_t* dst_dev, uint64_t dst_capacity_words, const uint32_t* payload_host, uint64_t n_payload_words,
                         const hdt_range* ranges_host, uint32_t n_ranges);

/* Kernel launches issued by this context since creation (bench bookkeeping). */
uint64_t hdt_launch_count(const hdt_ctx* ctx);
/* Colour passes of this context that read the ancestor records of the paths pass (HDT_OPT_COLORS_RECORDED) instead of
 * walking the DAG (diagnostics: which route did hdt_resolve_colors take?). */
uint64_t hdt_recorded_color_passes(const hdt_ctx* ctx);
int hdt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HASHDAG_B200_H */
