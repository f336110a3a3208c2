// Host layer of the multi-GPU path behind the C ABI: one process per GPU, the DAG replicated, an edit's
// dirty spans broadcast and applied on every replica.  Included at the end of hdt_tracer.cu (it needs hdt_ctx).
//
//   hdt_dirty_tracker      what changed since the last upload, from the hash table's own bookkeeping: the reference keeps, per
//                          bucket, its fill count now and at the last upload and only ever appends (hash_table.cpp:146-183), so a
//                          delta is the walk of HashTable::upload_to_gpu turned into page-split, physically sorted, merged spans
//                          + the touched page-table entries.  O(#buckets), no array comparison.  (C++ twin of
//                          hashdag_b200/edits.py::delta_from_bucket_sizes, which stays as its checker.)
//   hdt_comm_*             an NCCL communicator per context.  libnccl is loaded at run time (dlopen): the library has no
//                          link-time dependency on it and single-GPU hosts never touch it.
//   hdt_replicate          ncclBroadcast of a device buffer (initial replication of pool, page table, colours).
//   hdt_broadcast_dirty    one edit: the root's delta -> packed staging buffer -> ncclBroadcast -> apply_ranges_kernel on every
//                          rank's replica, page table and pool, then the resolved / prefix pools of the touched pages -- all on
//                          the tracer's stream, in order with the frames around it.  Replaces HashTable::upload_to_gpu's
//                          whole-page-table copy + one cudaMemcpyAsync per grown bucket (hash_table.cpp:120-184).
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

namespace {

// ---- libnccl, resolved at run time --------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };
static_assert(sizeof(NcclUniqueId) == HDT_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
typedef struct ncclComm* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int /* ncclDataType_t */, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};
constexpr int kNcclUint8 = 1;   // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1 (nccl.h, stable since NCCL 2.0)

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { getenv("HDT_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = "libnccl.so.2 not found (set HDT_NCCL_LIB to its path)"; return; }
        auto sym = [&](const char* s) { void* p = dlsym(api.lib, s); if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + s; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}
int nccl_fail(int rc, const char* where)
{
    g_lastError = std::string(where) + ": NCCL error " + std::to_string(rc) + " (" + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "?") + ")";
    return HDT_ERR_NCCL;
}
#define HDT_NCCL(call)                                 \
    do {                                               \
        const int r__ = (call);                        \
        if (r__ != 0) return nccl_fail(r__, #call);    \
    } while (0)

// Virtual address space of the reference's HashTable (hash_dag_globals.h:7-38 with the defaults of typedefs.h:201-236,
// hash_table.h:18-63): levels 0-8 have 1024 buckets of 1024 words, deeper levels 65536 buckets of 4096 words.
struct HashLayoutHost {
    u32 levels;
    static constexpr u32 kTopLevels = 9, kTopBuckets = 1024, kTopBucketWords = 1024, kLowBuckets = 65536, kLowBucketWords = 4096;
    u32 n_buckets() const { return std::min(levels, kTopLevels) * kTopBuckets + (levels > kTopLevels ? levels - kTopLevels : 0) * kLowBuckets; }
    u64 bucket_base(u32 bucket) const   // HashDagUtils::make_ptr(level, bucket, 0) of global bucket index `bucket`
    {
        const u32 nTop = std::min(levels, kTopLevels) * kTopBuckets;
        if (bucket < nTop) return u64(bucket) * kTopBucketWords;
        return u64(nTop) * kTopBucketWords + u64(bucket - nTop) * kLowBucketWords;
    }
    u64 n_pages() const { return bucket_base(n_buckets()) / kPageWords; }
};

}  // namespace

struct hdt_dirty_tracker {
    HashLayoutHost layout;
    u32 mergeGap = 32;
    std::vector<u32> last;                       // bucket fill counts at the last delta (lastBucketsSizes, hash_table.cpp:161)
    // storage the delta handed out points into (valid until the next hdt_tracker_delta)
    std::vector<hdt_range> poolRanges, tableRanges;
    std::vector<u32> poolPayload, tablePayload;
    struct Piece { u64 phys; u32 n; };
    std::vector<Piece> pieces;
    std::vector<u32> pages;
};

extern "C" {

int hdt_tracker_create(uint32_t levels, hdt_dirty_tracker** out)
{
    if (!out || levels < 3 || levels > kMaxLevels) return fail(HDT_ERR_ARG, "hdt_tracker_create: bad arguments");
    hdt_dirty_tracker* t = new hdt_dirty_tracker();
    t->layout.levels = levels;
    t->last.assign(t->layout.n_buckets(), 0);
    *out = t;
    return HDT_OK;
}
int hdt_tracker_destroy(hdt_dirty_tracker* t) { delete t; return HDT_OK; }
uint32_t hdt_tracker_bucket_count(const hdt_dirty_tracker* t) { return t ? u32(t->last.size()) : 0; }

int hdt_tracker_snapshot(hdt_dirty_tracker* t, const uint32_t* bucket_sizes, uint32_t n_buckets)
{
    if (!t || !bucket_sizes) return fail(HDT_ERR_ARG, "null argument");
    if (n_buckets < t->last.size()) return fail(HDT_ERR_ARG, "hdt_tracker_snapshot: fewer bucket sizes than the DAG's levels have buckets");
    std::copy(bucket_sizes, bucket_sizes + t->last.size(), t->last.begin());
    return HDT_OK;
}

int hdt_tracker_delta(hdt_dirty_tracker* t, const uint32_t* bucket_sizes, uint32_t n_buckets, const uint32_t* cpu_pool, const uint32_t* cpu_page_table,
                      uint32_t first_node_index, uint32_t pool_top, hdt_dag_delta* out)
{
    if (!t || !bucket_sizes || !cpu_pool || !cpu_page_table || !out) return fail(HDT_ERR_ARG, "hdt_tracker_delta: null argument");
    if (n_buckets < t->last.size()) return fail(HDT_ERR_ARG, "hdt_tracker_delta: fewer bucket sizes than the DAG's levels have buckets");
    const u64 nPages = t->layout.n_pages();
    t->pieces.clear(); t->pages.clear();
    t->poolRanges.clear(); t->tableRanges.clear(); t->poolPayload.clear(); t->tablePayload.clear();
    // the walk of HashTable::upload_to_gpu (hash_table.cpp:158-183): per grown bucket the words between its size then and now,
    // split at page boundaries and pushed through the page table
    for (u32 b = 0; b < t->last.size(); ++b) {
        const u32 was = t->last[b], now = bucket_sizes[b];
        if (now == was) continue;
        if (now < was) return fail(HDT_ERR_STATE, "hdt_tracker_delta: a bucket shrank -- not an append-only edit (undo / GC are outside the tracker)");
        const u64 base = t->layout.bucket_base(b);
        u64 pos = base + was;
        const u64 end = base + now;
        while (pos < end) {
            const u64 page = pos / kPageWords;
            const u32 inPage = u32(std::min<u64>(kPageWords - pos % kPageWords, end - pos));
            if (page >= nPages) return fail(HDT_ERR_ARG, "hdt_tracker_delta: a bucket reaches beyond the page table");
            const u32 phys = cpu_page_table[page];
            if (phys == 0 || phys >= pool_top) return fail(HDT_ERR_ARG, "hdt_tracker_delta: a grown bucket lies in an unallocated page");
            t->pieces.push_back({ u64(phys) * kPageWords + pos % kPageWords, inPage });
            t->pages.push_back(u32(page));
            pos += inPage;
        }
        t->last[b] = now;
    }
    std::sort(t->pieces.begin(), t->pieces.end(), [](const auto& a, const auto& b) { return a.phys < b.phys; });
    // merge neighbouring pieces (gap <= mergeGap words: what lies between is unchanged on both sides)
    for (size_t i = 0; i < t->pieces.size();) {
        const u64 d0 = t->pieces[i].phys;
        u64 d1 = d0 + t->pieces[i].n;
        size_t j = i + 1;
        while (j < t->pieces.size() && t->pieces[j].phys - d1 <= t->mergeGap) { d1 = t->pieces[j].phys + t->pieces[j].n; ++j; }
        t->poolRanges.push_back(hdt_range{ d0, t->poolPayload.size(), d1 - d0 });
        t->poolPayload.insert(t->poolPayload.end(), cpu_pool + d0, cpu_pool + d1);
        i = j;
    }
    // page table: the entries of every touched page, as runs of consecutive pages (rewriting an unchanged entry is harmless)
    std::sort(t->pages.begin(), t->pages.end());
    t->pages.erase(std::unique(t->pages.begin(), t->pages.end()), t->pages.end());
    for (size_t i = 0; i < t->pages.size();) {
        size_t j = i + 1;
        while (j < t->pages.size() && t->pages[j] == t->pages[j - 1] + 1) ++j;
        const u32 p0 = t->pages[i], n = u32(j - i);
        t->tableRanges.push_back(hdt_range{ p0, t->tablePayload.size(), n });
        t->tablePayload.insert(t->tablePayload.end(), cpu_page_table + p0, cpu_page_table + p0 + n);
        i = j;
    }
    out->first_node_index = first_node_index; out->pool_top = pool_top;
    out->pool_ranges = t->poolRanges.data(); out->n_pool_ranges = u32(t->poolRanges.size());
    out->pool_payload = t->poolPayload.data(); out->n_pool_payload = t->poolPayload.size();
    out->table_ranges = t->tableRanges.data(); out->n_table_ranges = u32(t->tableRanges.size());
    out->table_payload = t->tablePayload.data(); out->n_table_payload = t->tablePayload.size();
    return HDT_OK;
}

// ---- communicator --------------------------------------------------------------------------------------------------------
int hdt_comm_unique_id(uint8_t id_out[HDT_COMM_ID_BYTES])
{
    if (!id_out) return fail(HDT_ERR_ARG, "null argument");
    NcclApi& n = nccl();
    if (!n.error.empty()) return fail(HDT_ERR_NCCL, n.error.c_str());
    NcclUniqueId id;
    HDT_NCCL(n.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return HDT_OK;
}

int hdt_comm_init(hdt_ctx* c, const uint8_t id[HDT_COMM_ID_BYTES], uint32_t rank, uint32_t world)
{
    if (!c || !id || !world || rank >= world) return fail(HDT_ERR_ARG, "hdt_comm_init: bad arguments");
    if (c->comm) return fail(HDT_ERR_STATE, "hdt_comm_init: the context already has a communicator");
    c->commRank = rank; c->commWorld = world;
    if (world == 1) return HDT_OK;   // nothing to talk to: broadcasts degenerate to local applies
    NcclApi& n = nccl();
    if (!n.error.empty()) return fail(HDT_ERR_NCCL, n.error.c_str());
    HDT_CUDA(cudaSetDevice(c->device));
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NcclComm comm = nullptr;
    HDT_NCCL(n.CommInitRank(&comm, int(world), uid, int(rank)));
    c->comm = comm;
    return HDT_OK;
}

int hdt_comm_destroy(hdt_ctx* c)
{
    if (!c) return fail(HDT_ERR_ARG, "null context");
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        nccl().CommDestroy(static_cast<NcclComm>(c->comm));
        c->comm = nullptr;
    }
    c->commWorld = 1; c->commRank = 0;
    return HDT_OK;
}

int hdt_replicate(hdt_ctx* c, void* dev_buffer, uint64_t n_bytes, uint32_t root)
{
    if (!c || (!dev_buffer && n_bytes)) return fail(HDT_ERR_ARG, "hdt_replicate: null argument");
    if (root >= c->commWorld) return fail(HDT_ERR_ARG, "hdt_replicate: no such root");
    if (!n_bytes || c->commWorld == 1) return HDT_OK;
    if (!c->comm) return fail(HDT_ERR_STATE, "hdt_replicate: hdt_comm_init first");
    HDT_CUDA(cudaSetDevice(c->device));
    HDT_NCCL(nccl().Broadcast(dev_buffer, dev_buffer, size_t(n_bytes), kNcclUint8, int(root), static_cast<NcclComm>(c->comm), c->stream));
    return HDT_OK;
}

namespace {
// staging of the context grown to `need` bytes (pinned host + device twins); everything staged so far has been consumed
// once the stream is synchronised
int staging_reserve(hdt_ctx* c, size_t need)
{
    if (c->stagingUsed + need <= c->stagingCap) return HDT_OK;
    HDT_CUDA(cudaStreamSynchronize(c->stream));
    c->stagingUsed = 0;
    if (need > c->stagingCap) {
        const size_t cap = std::max(need * 2, size_t(8) << 20);
        if (c->stagingHost) cudaFreeHost(c->stagingHost);
        cudaFree(c->stagingDev);
        c->stagingHost = nullptr; c->stagingDev = nullptr; c->stagingCap = 0;
        HDT_CUDA(cudaMallocHost(&c->stagingHost, cap));
        HDT_CUDA(cudaMalloc(&c->stagingDev, cap));
        c->stagingCap = cap;
    }
    return HDT_OK;
}
size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// what travels first: the sizes of the packed delta (so that the other ranks can post the second broadcast)
struct DeltaHeader { u32 firstNodeIndex, poolTop, nPoolRanges, nTableRanges; u64 nPoolPayload, nTablePayload; };
}  // namespace

int hdt_broadcast_ranges(hdt_ctx* c, uint32_t root, uint32_t* dst_dev, uint64_t dst_capacity_words, const uint32_t* payload_host, uint64_t n_payload_words,
                         const hdt_range* ranges_host, uint32_t n_ranges)
{
    if (!c || !dst_dev) return fail(HDT_ERR_ARG, "hdt_broadcast_ranges: null argument");
    if (root >= c->commWorld) return fail(HDT_ERR_ARG, "hdt_broadcast_ranges: no such root");
    const bool isRoot = c->commRank == root, multi = c->commWorld > 1;
    if (multi && !c->comm) return fail(HDT_ERR_STATE, "hdt_broadcast_ranges: hdt_comm_init first");
    if (isRoot && ((!ranges_host && n_ranges) || (!payload_host && n_payload_words))) return fail(HDT_ERR_ARG, "hdt_broadcast_ranges: null arrays");
    HDT_CUDA(cudaSetDevice(c->device));
    NcclComm comm = static_cast<NcclComm>(c->comm);
    DeltaHeader h{};
    if (isRoot) {
        for (u32 i = 0; i < n_ranges; ++i) {
            const hdt_range& r = ranges_host[i];
            if (r.src_word > n_payload_words || r.n_words > n_payload_words - r.src_word || r.dst_word + r.n_words > dst_capacity_words)
                return fail(HDT_ERR_ARG, "hdt_broadcast_ranges: a range lies outside its payload or beyond the destination");
        }
        h.nPoolRanges = n_ranges; h.nPoolPayload = n_payload_words;
    }
    if (multi) {
        if (int rc = staging_reserve(c, 256)) return rc;
        char* hh = c->stagingHost + c->stagingUsed;
        char* hd = c->stagingDev + c->stagingUsed;
        c->stagingUsed += 256;
        if (isRoot) { memcpy(hh, &h, sizeof(h)); HDT_CUDA(cudaMemcpyAsync(hd, hh, sizeof(h), cudaMemcpyHostToDevice, c->stream)); }
        HDT_NCCL(nccl().Broadcast(hd, hd, sizeof(h), kNcclUint8, int(root), comm, c->stream));
        if (!isRoot) {
            HDT_CUDA(cudaMemcpyAsync(hh, hd, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            HDT_CUDA(cudaStreamSynchronize(c->stream));
            memcpy(&h, hh, sizeof(h));
        }
    }
    if (!h.nPoolRanges) return HDT_OK;
    const size_t oP = align256(size_t(h.nPoolRanges) * sizeof(hdt_range)), total = oP + align256(size_t(h.nPoolPayload) * 4);
    if (int rc = staging_reserve(c, total)) return rc;
    char* bh = c->stagingHost + c->stagingUsed;
    char* bd = c->stagingDev + c->stagingUsed;
    c->stagingUsed += total;
    if (isRoot) {
        memcpy(bh, ranges_host, size_t(h.nPoolRanges) * sizeof(hdt_range));
        memcpy(bh + oP, payload_host, size_t(h.nPoolPayload) * 4);
        HDT_CUDA(cudaMemcpyAsync(bd, bh, total, cudaMemcpyHostToDevice, c->stream));
    }
    if (multi) HDT_NCCL(nccl().Broadcast(bd, bd, total, kNcclUint8, int(root), comm, c->stream));
    apply_ranges_kernel<<<std::min(h.nPoolRanges, 1184u), 128, 0, c->stream>>>(dst_dev, reinterpret_cast<const u32*>(bd + oP), reinterpret_cast<const hdt_range*>(bd), h.nPoolRanges);
    HDT_LAUNCHED("apply_ranges_kernel");
    return HDT_OK;
}

int hdt_broadcast_dirty(hdt_ctx* c, uint32_t root, const hdt_dag_delta* delta, hdt_replica* replica)
{
    if (!c || !replica) return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: null argument");
    if (root >= c->commWorld) return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: no such root");
    const bool isRoot = c->commRank == root, multi = c->commWorld > 1;
    if (isRoot && !delta) return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: the root passes the delta");
    if (multi && !c->comm) return fail(HDT_ERR_STATE, "hdt_broadcast_dirty: hdt_comm_init first");
    if (!replica->pool || !replica->page_table) return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: replica without pool / page table");
    HDT_CUDA(cudaSetDevice(c->device));
    NcclComm comm = static_cast<NcclComm>(c->comm);

    // 1. sizes
    DeltaHeader h{};
    if (isRoot) {
        if ((delta->n_pool_ranges && (!delta->pool_ranges || !delta->pool_payload)) || (delta->n_table_ranges && (!delta->table_ranges || !delta->table_payload)))
            return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: delta with null arrays");
        h = DeltaHeader{ delta->first_node_index, delta->pool_top, delta->n_pool_ranges, delta->n_table_ranges, delta->n_pool_payload, delta->n_table_payload };
    }
    if (multi) {
        if (int rc = staging_reserve(c, 256)) return rc;
        char* hh = c->stagingHost + c->stagingUsed;
        char* hd = c->stagingDev + c->stagingUsed;
        c->stagingUsed += 256;
        if (isRoot) { memcpy(hh, &h, sizeof(h)); HDT_CUDA(cudaMemcpyAsync(hd, hh, sizeof(h), cudaMemcpyHostToDevice, c->stream)); }
        HDT_NCCL(nccl().Broadcast(hd, hd, sizeof(h), kNcclUint8, int(root), comm, c->stream));
        if (!isRoot) {
            HDT_CUDA(cudaMemcpyAsync(hh, hd, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            HDT_CUDA(cudaStreamSynchronize(c->stream));
            memcpy(&h, hh, sizeof(h));
        }
    }
    if (u64(h.poolTop) * kPageWords > replica->pool_capacity_words) return fail(HDT_ERR_CAPACITY, "hdt_broadcast_dirty: the edit outgrew the replica's pool");

    // 2. body: [pool ranges][table ranges][pool payload][table payload], one buffer, one broadcast
    const size_t oPR = 0, oTR = oPR + align256(size_t(h.nPoolRanges) * sizeof(hdt_range)), oPP = oTR + align256(size_t(h.nTableRanges) * sizeof(hdt_range));
    const size_t oTP = oPP + align256(size_t(h.nPoolPayload) * 4), total = oTP + align256(size_t(h.nTablePayload) * 4);
    if (total) {
        if (int rc = staging_reserve(c, total)) return rc;
        char* bh = c->stagingHost + c->stagingUsed;
        char* bd = c->stagingDev + c->stagingUsed;
        c->stagingUsed += total;
        if (isRoot) {
            for (u32 i = 0; i < h.nPoolRanges; ++i) {
                const hdt_range& r = delta->pool_ranges[i];
                if (r.src_word > h.nPoolPayload || r.n_words > h.nPoolPayload - r.src_word || r.dst_word + r.n_words > u64(h.poolTop) * kPageWords)
                    return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: a pool range lies outside its payload or beyond pool_top");
            }
            for (u32 i = 0; i < h.nTableRanges; ++i) {
                const hdt_range& r = delta->table_ranges[i];
                if (r.src_word > h.nTablePayload || r.n_words > h.nTablePayload - r.src_word || r.dst_word + r.n_words > replica->page_table_size)
                    return fail(HDT_ERR_ARG, "hdt_broadcast_dirty: a page-table range lies outside its payload or beyond the page table");
            }
            memcpy(bh + oPR, delta->pool_ranges, size_t(h.nPoolRanges) * sizeof(hdt_range));
            memcpy(bh + oTR, delta->table_ranges, size_t(h.nTableRanges) * sizeof(hdt_range));
            memcpy(bh + oPP, delta->pool_payload, size_t(h.nPoolPayload) * 4);
            memcpy(bh + oTP, delta->table_payload, size_t(h.nTablePayload) * 4);
            HDT_CUDA(cudaMemcpyAsync(bd, bh, total, cudaMemcpyHostToDevice, c->stream));
        }
        if (multi) HDT_NCCL(nccl().Broadcast(bd, bd, total, kNcclUint8, int(root), comm, c->stream));
        // 3. apply, in stream order
        if (h.nTableRanges) {
            apply_ranges_kernel<<<std::min(h.nTableRanges, 1184u), 128, 0, c->stream>>>(replica->page_table, reinterpret_cast<const u32*>(bd + oTP),
                                                                                       reinterpret_cast<const hdt_range*>(bd + oTR), h.nTableRanges);
            HDT_LAUNCHED("apply_ranges_kernel");
        }
        if (h.nPoolRanges) {
            apply_ranges_kernel<<<std::min(h.nPoolRanges, 1184u), 128, 0, c->stream>>>(replica->pool, reinterpret_cast<const u32*>(bd + oPP),
                                                                                      reinterpret_cast<const hdt_range*>(bd + oPR), h.nPoolRanges);
            HDT_LAUNCHED("apply_ranges_kernel");
        }
        // 4. the library-format pools follow: the pages the pool spans touch
        if (replica->resolved_pool && h.nPoolRanges) {
            const hdt_range* ranges = isRoot ? delta->pool_ranges : nullptr;
            if (!isRoot) {   // the ranges arrived on the device; the page list is built on the host
                HDT_CUDA(cudaMemcpyAsync(bh + oPR, bd + oPR, size_t(h.nPoolRanges) * sizeof(hdt_range), cudaMemcpyDeviceToHost, c->stream));
                HDT_CUDA(cudaStreamSynchronize(c->stream));
                ranges = reinterpret_cast<const hdt_range*>(bh + oPR);
            }
            hdt_hash_dag d{};
            d.page_table_size = replica->page_table_size; d.pool_top = h.poolTop; d.page_table = replica->page_table; d.pool = replica->pool;
            d.first_node_index = h.firstNodeIndex;
            // copy: hdt_hash_dag_resolve stages its page list through the same staging buffer, which may move
            std::vector<hdt_range> keep(ranges, ranges + h.nPoolRanges);
            if (int rc = hdt_hash_dag_resolve(c, &d, sizeof(d), replica->resolved_pool, replica->prefix_pool, replica->pool_capacity_words, keep.data(), h.nPoolRanges)) return rc;
        }
    }
    replica->first_node_index = h.firstNodeIndex;
    replica->pool_top = h.poolTop;
    c->ancValid = false;
    return HDT_OK;
}

}  // extern "C"
