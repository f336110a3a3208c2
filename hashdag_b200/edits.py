"""Edit-dirtied spans of a HashDAG: find them, ship them, apply them to device replicas.

The reference re-uploads after every edit what `HashTable::upload_to_gpu`
(/root/reference/src/dags/hash_dag/hash_table.cpp:120-184) finds dirty: the whole page table (16 MiB at
depth 17) and, per bucket that grew, the new words -- one cudaMemcpyAsync each -- plus the colour tree
and every rebuilt colour leaf (hash_dag_colors.h:102-120).  Here one edit becomes one packed delta
(`{dst_word, src_word, n_words}` spans + payload) that is broadcast once to every GPU
(torch.distributed) and applied by one kernel launch per array (`hdt_apply_ranges`), on the tracer's
stream, in order with the frames around it.  Edits themselves (find_or_add into the hash table) stay with
the host editor; this module only needs the arrays before and after.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

RANGE_DTYPE = np.dtype([("dst_word", "<u8"), ("src_word", "<u8"), ("n_words", "<u8")])   # hdt_range


def dirty_spans(old: np.ndarray, new: np.ndarray, n_new: int | None = None, merge_gap: int = 32):
    """Spans of `new[:n_new]` that differ from `old` (everything beyond len(old) counts as changed).
    -> (ranges[RANGE_DTYPE], payload) with payload = the spans' words back to back.
    Spans closer than `merge_gap` words are merged: fewer, longer copies."""
    old = np.ascontiguousarray(old).reshape(-1)
    new = np.ascontiguousarray(new).reshape(-1)
    assert old.dtype == new.dtype
    n_new = new.size if n_new is None else int(n_new)
    common = min(old.size, n_new)
    idx = np.flatnonzero(old[:common] != new[:common])
    starts, ends = [], []
    if idx.size:
        cut = np.flatnonzero(np.diff(idx) > merge_gap)
        starts = idx[np.concatenate(([0], cut + 1))].tolist()
        ends = (idx[np.concatenate((cut, [idx.size - 1]))] + 1).tolist()
    if n_new > common:
        if ends and common - ends[-1] <= merge_gap:
            ends[-1] = n_new
        else:
            starts.append(common)
            ends.append(n_new)
    ranges = np.zeros(len(starts), dtype=RANGE_DTYPE)
    chunks, src = [], 0
    for i, (a, b) in enumerate(zip(starts, ends)):
        ranges[i] = (a, src, b - a)
        chunks.append(new[a:b])
        src += b - a
    payload = np.concatenate(chunks) if chunks else np.zeros(0, dtype=new.dtype)
    return ranges, payload


def apply_spans_host(dst: np.ndarray, ranges: np.ndarray, payload: np.ndarray) -> None:
    """Host mirror of apply_ranges_kernel (csrc/hdt_tracer.cu): dst[r.dst_word + i] = payload[r.src_word + i]."""
    for r in ranges:
        d, s, n = int(r["dst_word"]), int(r["src_word"]), int(r["n_words"])
        dst[d:d + n] = payload[s:s + n]


@dataclass
class ColorLeafArrays:
    """One unique colour leaf (CompressedColorLeaf, vwsc.h:157-191) as host arrays."""
    weights: np.ndarray
    blocks: np.ndarray
    macro_blocks: np.ndarray

    def same_as(self, o) -> bool:
        return o is not None and np.array_equal(self.weights, o.weights) and np.array_equal(self.blocks, o.blocks) and np.array_equal(self.macro_blocks, o.macro_blocks)


@dataclass
class DagDelta:
    """Everything a replica needs to follow one edit."""
    first_node_index: int
    pool_top: int
    pool_ranges: np.ndarray
    pool_payload: np.ndarray            # uint32
    table_ranges: np.ndarray
    table_payload: np.ndarray           # uint32
    color_node_ranges: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=RANGE_DTYPE))
    color_node_payload: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.uint32))
    n_color_nodes: int = 0
    color_leaves: dict = field(default_factory=dict)   # leaf index -> ColorLeafArrays (new or rebuilt leaves)
    n_color_leaves: int = 0

    @property
    def nbytes(self) -> int:
        n = sum(a.nbytes for a in (self.pool_ranges, self.pool_payload, self.table_ranges, self.table_payload, self.color_node_ranges, self.color_node_payload))
        return n + sum(l.weights.nbytes + l.blocks.nbytes + l.macro_blocks.nbytes for l in self.color_leaves.values())


def diff_hash_dag(old_pool, old_table, new_pool, new_table, first_node_index, pool_top,
                  old_color_nodes=None, new_color_nodes=None, old_leaves=None, new_leaves=None) -> DagDelta:
    """Delta between two versions of a HashDAG (+ its HashDAGColors tree and unique leaves)."""
    pr, pp = dirty_spans(old_pool, new_pool, int(pool_top) * 512)
    tr, tp = dirty_spans(old_table, new_table)
    d = DagDelta(int(first_node_index), int(pool_top), pr, pp, tr, tp)
    if new_color_nodes is not None:
        d.color_node_ranges, d.color_node_payload = dirty_spans(old_color_nodes if old_color_nodes is not None else np.zeros(0, np.uint32), new_color_nodes)
        d.n_color_nodes = int(new_color_nodes.size)
    if new_leaves is not None:
        old_leaves = old_leaves or []
        d.n_color_leaves = len(new_leaves)
        for i, leaf in enumerate(new_leaves):
            if not leaf.same_as(old_leaves[i] if i < len(old_leaves) else None):
                d.color_leaves[i] = leaf
    return d


# ---------------------------------------------------------------------------------------------
# the dirty tracker proper: what grew since the last upload, from the hash table's own bookkeeping
# ---------------------------------------------------------------------------------------------
class HashLayout:
    """Virtual address space of the reference's HashTable (hash_dag_globals.h:7-38 with the defaults of
    typedefs.h:201-236): levels 0-8 have 1024 buckets of 1024 words, deeper levels 65536 buckets of 4096 words;
    `bucket_base[i]` = HashDagUtils::make_ptr(level, bucket, 0) (hash_table.h:45-63) of global bucket i
    (get_bucket_global_index, hash_table.h:18-35)."""
    PAGE = 512

    def __init__(self, levels, top_levels=9, top_bits=10, low_bits=16, top_size=1024, low_size=4096):
        n_top = min(top_levels, levels) << top_bits
        n_low = max(levels - top_levels, 0) << low_bits
        self.levels, self.n_buckets = levels, n_top + n_low
        self.bucket_base = np.concatenate((np.arange(n_top, dtype=np.int64) * top_size, n_top * top_size + np.arange(n_low, dtype=np.int64) * low_size))
        self.n_pages = int((n_top * top_size + n_low * low_size) // self.PAGE)


def delta_from_bucket_sizes(layout: HashLayout, last_sizes, sizes, cpu_pool, page_table, first_node_index, pool_top, merge_gap: int = 32) -> "DagDelta":
    """The pool / page-table part of a delta WITHOUT comparing the arrays: the hash table only ever appends to its
    buckets, so what changed since the last upload is, per bucket, the words between its size then and now -- the very
    walk of HashTable::upload_to_gpu (hash_table.cpp:158-183).  `cpu_pool` / `page_table` are the host arrays (physical
    layout == the GPU pool's under MANUAL_VIRTUAL_MEMORY, hash_table.cpp:137-141).  O(#buckets) instead of O(pool)."""
    last_sizes = np.asarray(last_sizes)
    sizes = np.asarray(sizes)
    grown = np.flatnonzero(sizes != last_sizes)
    if grown.size == 0:
        e = np.zeros(0, dtype=RANGE_DTYPE)
        return DagDelta(int(first_node_index), int(pool_top), e, np.zeros(0, np.uint32), e.copy(), np.zeros(0, np.uint32))
    P = layout.PAGE
    start = layout.bucket_base[grown] + last_sizes[grown].astype(np.int64)     # virtual pointers
    end = layout.bucket_base[grown] + sizes[grown].astype(np.int64)
    assert (end > start).all(), "a bucket shrank: not an append-only edit (undo / GC are outside the tracker)"
    # split at page boundaries: a bucket spans at most low_size / PAGE pages
    first_page, last_page = start // P, (end - 1) // P
    n_pieces = (last_page - first_page + 1)
    idx = np.repeat(np.arange(grown.size), n_pieces)
    page = np.repeat(first_page, n_pieces) + (np.arange(n_pieces.sum()) - np.repeat(np.cumsum(n_pieces) - n_pieces, n_pieces))
    lo = np.maximum(start[idx], page * P)
    hi = np.minimum(end[idx], (page + 1) * P)
    phys = page_table[page].astype(np.int64) * P + (lo - page * P)
    order = np.argsort(phys, kind="stable")
    phys, n = phys[order], (hi - lo)[order]
    # merge neighbouring pieces (gap <= merge_gap words: the words in between are unchanged on both sides)
    brk = np.flatnonzero(phys[1:] - (phys[:-1] + n[:-1]) > merge_gap) + 1
    s = np.concatenate(([0], brk))
    e = np.concatenate((brk, [phys.size]))
    d0, d1 = phys[s], (phys + n)[e - 1]
    ranges = np.zeros(s.size, dtype=RANGE_DTYPE)
    ranges["dst_word"], ranges["n_words"] = d0, d1 - d0
    ranges["src_word"] = np.concatenate(([0], np.cumsum(d1 - d0)[:-1]))
    payload = np.concatenate([cpu_pool[a:b] for a, b in zip(d0.tolist(), d1.tolist())]).astype(np.uint32, copy=False)
    # page table: the entries of every touched page (rewriting an unchanged entry is harmless)
    pages = np.unique(page)
    pb = np.flatnonzero(np.diff(pages) > 1) + 1
    ps, pe = np.concatenate(([0], pb)), np.concatenate((pb, [pages.size]))
    tr = np.zeros(ps.size, dtype=RANGE_DTYPE)
    tr["dst_word"], tr["n_words"] = pages[ps], pages[pe - 1] - pages[ps] + 1
    tr["src_word"] = np.concatenate(([0], np.cumsum(tr["n_words"])[:-1]))
    tp = np.concatenate([page_table[a:b + 1] for a, b in zip(pages[ps].tolist(), pages[pe - 1].tolist())]).astype(np.uint32, copy=False)
    return DagDelta(int(first_node_index), int(pool_top), ranges, payload, tr, tp)


def resolve_pool_host(layout: HashLayout, pool, page_table, pool_top: int, pages=None) -> np.ndarray:
    """Host mirror of resolve_pages_kernel (csrc/hdt_resolve.cuh): the pool with every child pointer replaced by the child's
    physical word index.  Pages are parsed on their own: a virtual page belongs to one level (hash_table.h:45-63); interior
    nodes `[header][pointer] x popc(header & 0xFF)` are packed from the start of a page, never straddle one and the tail is
    zero padding (add_interior_node, hash_table.h:416-430); pages of level >= levels-2 hold 64-bit leaves and are copied.
    `pages`: physical pages to resolve (default all below pool_top); the others are returned unchanged."""
    P = layout.PAGE
    pool = np.asarray(pool)
    out = pool[: pool_top * P].copy()
    page_table = np.asarray(page_table)
    used = np.flatnonzero(page_table)
    used = used[page_table[used] < pool_top]
    virt_of = np.full(pool_top, -1, dtype=np.int64)
    virt_of[page_table[used]] = used
    top_pages = min(9, layout.levels) * 1024 * (1024 // P)
    for p in (range(pool_top) if pages is None else pages):
        v = int(virt_of[p])
        if v < 0:
            continue
        level = v // (1024 * (1024 // P)) if v < top_pages else 9 + (v - top_pages) // (65536 * (4096 // P))
        if level >= layout.levels - 2:
            continue
        words = pool[p * P:(p + 1) * P]
        pos = 0
        while pos < P:
            hdr = int(words[pos])
            if hdr & 0xFF == 0:
                break
            n = bin(hdr & 0xFF).count("1")
            ptrs = words[pos + 1: pos + 1 + n].astype(np.int64)
            out[p * P + pos + 1: p * P + pos + 1 + n] = (page_table[ptrs >> 9].astype(np.int64) * P + (ptrs & (P - 1))).astype(np.uint32)
            pos += 1 + n
    return out


def add_color_delta(d: "DagDelta", old_color_nodes, new_color_nodes, old_leaves, new_leaves) -> "DagDelta":
    """Colour part of a delta (tree nodes + rebuilt unique leaves), as in diff_hash_dag."""
    d.color_node_ranges, d.color_node_payload = dirty_spans(old_color_nodes if old_color_nodes is not None else np.zeros(0, np.uint32), new_color_nodes)
    d.n_color_nodes = int(new_color_nodes.size)
    old_leaves = old_leaves or []
    d.n_color_leaves = len(new_leaves)
    for i, leaf in enumerate(new_leaves):
        if not leaf.same_as(old_leaves[i] if i < len(old_leaves) else None):
            d.color_leaves[i] = leaf
    return d


# ---------------------------------------------------------------------------------------------
# shipping a delta to the other ranks
# ---------------------------------------------------------------------------------------------
def broadcast_delta(delta: DagDelta | None, src: int = 0, device="cpu") -> DagDelta:
    """Rank `src` passes its delta, the others None; everybody returns the delta.  One object broadcast for
    the sizes, one tensor broadcast for all the payload bytes (device="cuda:N" with the nccl backend)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    names = ("pool_ranges", "pool_payload", "table_ranges", "table_payload", "color_node_ranges", "color_node_payload")
    if rank == src:
        arrays = [np.ascontiguousarray(getattr(delta, n)) for n in names]
        leaf_ids = sorted(delta.color_leaves)
        for i in leaf_ids:
            l = delta.color_leaves[i]
            arrays += [np.ascontiguousarray(l.weights), np.ascontiguousarray(l.blocks), np.ascontiguousarray(l.macro_blocks)]
        head = {"first": delta.first_node_index, "top": delta.pool_top, "n_color_nodes": delta.n_color_nodes, "n_color_leaves": delta.n_color_leaves,
                "leaf_ids": leaf_ids, "sizes": [(a.dtype.str if a.dtype.names is None else "range", int(a.size)) for a in arrays]}
        box = [head]
    else:
        box = [None]
    dist.broadcast_object_list(box, src=src)
    head = box[0]
    nbytes = [n * (RANGE_DTYPE.itemsize if dt == "range" else np.dtype(dt).itemsize) for dt, n in head["sizes"]]
    offs = np.concatenate(([0], np.cumsum([(b + 7) // 8 * 8 for b in nbytes]))).astype(np.int64)
    total = int(offs[-1])
    if rank == src:
        flat = np.zeros(total, dtype=np.uint8)
        for a, o, b in zip(arrays, offs, nbytes):
            flat[o:o + b] = a.view(np.uint8).reshape(-1)
        t = torch.from_numpy(flat).to(device)
    else:
        t = torch.empty(total, dtype=torch.uint8, device=device)
    if total:
        dist.broadcast(t, src=src)
    if rank == src:
        return delta
    flat = t.cpu().numpy()
    out = []
    for (dt, n), o, b in zip(head["sizes"], offs, nbytes):
        out.append(flat[o:o + b].view(RANGE_DTYPE if dt == "range" else np.dtype(dt)).copy())
    d = DagDelta(head["first"], head["top"], out[0], out[1], out[2], out[3], out[4], out[5], head["n_color_nodes"], {}, head["n_color_leaves"])
    for k, i in enumerate(head["leaf_ids"]):
        d.color_leaves[i] = ColorLeafArrays(out[6 + 3 * k], out[7 + 3 * k], out[8 + 3 * k])
    return d


# ---------------------------------------------------------------------------------------------
# device replicas
# ---------------------------------------------------------------------------------------------
class HashDagReplica:
    """One GPU's copy of a HashDAG (+ colours) that follows edits through deltas.  Needs CUDA."""

    def __init__(self, tracer_obj, pool, page_table, pool_top, first_node_index, levels, pool_capacity_pages,
                 color_nodes=None, color_offsets=None, main_leaf=None, color_node_capacity=0, device="cuda:0", resolved=True, replicate_from=None,
                 color_leaves=None):
        """color_leaves: the unique colour leaves the colour tree already references (HashDAGColors::leaves, hash_dag_colors.h:65-73),
        a sequence of tracer.CompressedColorLeaf (or None for unused slots) on `device`; later deltas replace entries.
        replicate_from = r (tracer with a communicator): rank r's arrays are the truth -- the other ranks pass arrays of
        the same SIZES (contents ignored) and receive rank r's over hdt_replicate (initial replication)."""
        import torch
        from . import tracer as T
        self._T, self._torch, self.tracer, self.device, self.levels = T, torch, tracer_obj, device, levels
        cap = max(int(pool_capacity_pages), int(pool_top)) * 512
        self.pool = torch.zeros(cap, dtype=torch.int32, device=device)
        self.pool[: pool.size] = T._to_device(pool, device)
        self.page_table = T._to_device(page_table, device)
        self.pool_top, self.first_node_index = int(pool_top), int(first_node_index)
        if replicate_from is not None and tracer_obj.comm_world > 1:
            torch.cuda.synchronize()
            tracer_obj.replicate(self.pool[: int(pool_top) * 512], replicate_from)
            tracer_obj.replicate(self.page_table, replicate_from)
            tracer_obj.sync()
        # the resolved pool (child pointers pre-translated, csrc/hdt_resolve.cuh) follows every edit page by page
        # ... and so does the prefix pool (voxels under a node's earlier children: trace_colors without the DAG walk)
        self.resolved_pool = self.prefix_pool = None
        if resolved:
            self.resolved_pool = torch.zeros(cap, dtype=torch.int32, device=device)
            self.prefix_pool = torch.zeros(cap, dtype=torch.int32, device=device)
            torch.cuda.synchronize()                    # the uploads above ran on torch's stream, the resolve runs on the tracer's
            self.levels = levels
            tracer_obj.resolve_hash_dag(self._hash_dag(), self.resolved_pool, None, self.prefix_pool)
        self.has_colors = color_nodes is not None
        if self.has_colors:
            ncap = max(int(color_node_capacity), int(color_nodes.size))
            self.color_nodes = torch.zeros(ncap, dtype=torch.int32, device=device)
            self.color_nodes[: color_nodes.size] = T._to_device(color_nodes, device)
            self.n_color_nodes = int(color_nodes.size)
            self.color_offsets = T._to_device(color_offsets, device)
            self.main_leaf = main_leaf                  # tracer.CompressedColorLeaf on this device
            if replicate_from is not None and tracer_obj.comm_world > 1:
                torch.cuda.synchronize()
                for tns in (self.color_nodes[: self.n_color_nodes], self.color_offsets, main_leaf.weights, main_leaf.blocks, main_leaf.macro_blocks):
                    if tns is not None and tns.numel():
                        tracer_obj.replicate(tns, replicate_from)
                tracer_obj.sync()
            self.leaves = {}                            # index -> tracer.CompressedColorLeaf (keeps the tensors alive)
            self.leaf_pods = None                       # int64 tensor, 13 words per leaf (CompressedColorLeaf, 104 B)
            self._pods_host = None                      # the same on the host, updated row by row
            if color_leaves:
                self._pods_host = np.zeros((len(color_leaves), 13), dtype=np.uint64)
                for i, leaf in enumerate(color_leaves):
                    if leaf is not None:
                        self.leaves[i] = leaf
                        self._pods_host[i] = np.frombuffer(leaf.pod(), dtype=np.uint64)
                self.leaf_pods = T._to_device(self._pods_host.reshape(-1), device)

    def _replica_pod(self):
        T = self._T
        return T.ReplicaPod(self.pool.data_ptr(), self.pool.numel(), self.page_table.data_ptr(), self.page_table.numel(), self.first_node_index, self.pool_top,
                            T._ptr(self.resolved_pool), T._ptr(self.prefix_pool))

    def apply(self, delta: "DagDelta | None", root: int = 0, pod=None) -> None:
        """One edit on this rank's replica, through the C ABI's multi-GPU layer (hdt_broadcast_dirty / hdt_broadcast_ranges /
        hdt_replicate, csrc/hdt_multi.cuh): rank `root` passes the delta (and optionally `pod`, the hdt_dag_delta a
        tracer.DirtyTracker built, handed over without a copy), the other ranks None.  Everything is enqueued on the
        tracer's stream; with one rank no communicator is involved."""
        T, torch = self._T, self._torch
        tr = self.tracer
        is_root = tr.comm_rank == root
        multi = tr.comm_world > 1
        keep = None
        if is_root and pod is None:
            pod, keep = T.delta_pod_from_arrays(delta.first_node_index, delta.pool_top, delta.pool_ranges, delta.pool_payload, delta.table_ranges, delta.table_payload)
        rp = self._replica_pod()
        tr.broadcast_dirty(rp, pod if is_root else None, root)          # pool, page table, resolved + prefix pools
        self.pool_top, self.first_node_index = int(rp.pool_top), int(rp.first_node_index)
        del keep
        if not self.has_colors:
            return
        # colour tree + rebuilt unique leaves: sizes first (ranks other than root learn them here), then spans, then ONE packed leaf buffer
        ids, offs, flat = [], [], None
        if is_root:
            ids = sorted(delta.color_leaves)
            parts, total = [], 0
            for i in ids:
                l = delta.color_leaves[i]
                for a in (l.blocks, l.macro_blocks, l.weights):          # 8-byte arrays first: every part stays aligned
                    bts = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
                    offs.append((total, a.size))
                    parts.append(bts)
                    total += (bts.size + 7) // 8 * 8
            flat = np.zeros(total, dtype=np.uint8)
            for (o, _), bts in zip(offs, parts):
                flat[o:o + bts.size] = bts
        n_nodes = delta.n_color_nodes if is_root else 0
        n_leaves = delta.n_color_leaves if is_root else 0
        if multi:
            # [n_color_nodes, n_color_leaves, n_ids, blob_bytes, then per id: id, 3 x (offset, count)]  -- two hdt_replicate calls
            head = np.zeros(4, dtype=np.int64)
            meta = np.zeros(0, dtype=np.int64)
            if is_root:
                meta = np.array([v for k, i in enumerate(ids) for v in (i, *offs[3 * k], *offs[3 * k + 1], *offs[3 * k + 2])], dtype=np.int64)
                head[:] = (n_nodes, n_leaves, len(ids), flat.size)
            th = torch.from_numpy(head).to(self.device)
            torch.cuda.current_stream().synchronize()
            tr.replicate(th, root)
            tr.sync()
            n_nodes, n_leaves, n_ids, blob_bytes = (int(v) for v in th.cpu().tolist())
            if n_ids:
                tm = torch.from_numpy(meta).to(self.device) if is_root else torch.empty(7 * n_ids, dtype=torch.int64, device=self.device)
                torch.cuda.current_stream().synchronize()
                tr.replicate(tm, root)
                tr.sync()
                m = tm.cpu().numpy().reshape(n_ids, 7)
                ids = [int(v) for v in m[:, 0]]
                offs = [(int(m[k, 1 + 2 * j]), int(m[k, 2 + 2 * j])) for k in range(n_ids) for j in range(3)]
        if n_nodes:
            if n_nodes > self.color_nodes.numel():
                grown = torch.zeros(2 * n_nodes, dtype=torch.int32, device=self.device)
                tr.sync()
                grown[: self.color_nodes.numel()] = self.color_nodes
                torch.cuda.synchronize()
                self.color_nodes = grown
            if is_root:
                tr.broadcast_ranges(self.color_nodes, delta.color_node_payload, delta.color_node_ranges, root)
            else:
                tr.broadcast_ranges(self.color_nodes, None, None, root)
            self.n_color_nodes = n_nodes
            if ids:
                if is_root:
                    buf = torch.from_numpy(flat).to(self.device)
                else:
                    buf = torch.empty(blob_bytes, dtype=torch.uint8, device=self.device)
                if multi:
                    torch.cuda.current_stream().synchronize()
                    tr.replicate(buf, root)
                for k, i in enumerate(ids):
                    (ob, nb), (om, nm), (ow, nw) = offs[3 * k: 3 * k + 3]
                    blocks = buf[ob: ob + 8 * nb].view(torch.int64) if nb else None
                    macro = buf[om: om + 8 * nm].view(torch.int64) if nm else None
                    weights = buf[ow: ow + 4 * nw].view(torch.int32) if nw else None
                    self.leaves[i] = T.CompressedColorLeaf(weights, blocks, macro, T.UNIQUE_OFFSET)
            if n_leaves and (ids or self._pods_host is None or self._pods_host.shape[0] != n_leaves):
                if self._pods_host is None or self._pods_host.shape[0] < n_leaves:
                    grown = np.zeros((n_leaves, 13), dtype=np.uint64)
                    if self._pods_host is not None:
                        grown[: self._pods_host.shape[0]] = self._pods_host
                    self._pods_host = grown
                for i in ids:
                    self._pods_host[i] = np.frombuffer(self.leaves[i].pod(), dtype=np.uint64)
                tr.sync()                               # frames in flight may still read the previous POD array
                self.leaf_pods = T._to_device(self._pods_host[:n_leaves].reshape(-1), self.device)

    def _hash_dag(self):
        return self._T.HashDAG(self.pool, self.page_table, self.pool_top, self.first_node_index, self.levels)

    def dag(self):
        d = self._hash_dag()
        return d if self.resolved_pool is None else self._T.ResolvedHashDAG(d, self.resolved_pool, self.prefix_pool)

    def colors(self):
        T = self._T
        return T.HashDAGColors(self.color_nodes[: self.n_color_nodes], self.color_offsets, self.main_leaf, self.leaf_pods)
