"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT THE PRODUCT.  The reference's hash-table insert, one node at a time, on numpy arrays:
HashTable::find_or_add_interior_node / find_or_add_leaf_node with find_*_node_in_bucket, add_*_node and allocate_page
(/root/reference/src/dags/hash_dag/hash_table.h:196-560, :794-807), hashes from utils.h:68-110.  Its Bloom filter is a
pure accelerator (it only skips pages a node cannot be in) and is left out.  Pinned by tests/golden/ref_find_or_add_d13.npz,
written by the reference's own functions (tests/golden/make_find_or_add_golden.py)."""
import numpy as np

PAGE = 512
NONE = 0xFFFFFFFF
M32 = 0xFFFFFFFF


def murmur32xn(words):                       # utils.h:91-110
    h = 0
    for k in words:
        k = (int(k) * 0xcc9e2d51) & M32
        k = ((k << 15) | (k >> 17)) & M32
        k = (k * 0x1b873593) & M32
        h ^= k
        h = ((h << 13) | (h >> 19)) & M32
        h = (h * 5 + 0xe6546b64) & M32
    h ^= len(words)
    h ^= h >> 16; h = (h * 0x85ebca6b) & M32; h ^= h >> 13; h = (h * 0xc2b2ae35) & M32; h ^= h >> 16
    return h


def murmur64(h):                              # utils.h:77-85
    m = (1 << 64) - 1
    h ^= h >> 33; h = (h * 0xff51afd7ed558ccd) & m; h ^= h >> 33; h = (h * 0xc4ceb9fe1a85ec53) & m; h ^= h >> 33
    return h


def buckets_per_level(level):
    return 1024 if level < 9 else 65536


def bucket_capacity(level):
    return 1024 if level < 9 else 4096


def bucket_global_index(level, bucket):       # hash_table.h:18-35
    return level * 1024 + bucket if level < 9 else 9 * 1024 + (level - 9) * 65536 + bucket


def make_ptr(level, bucket, pos):             # hash_table.h:45-63
    if level < 9:
        return (level * 1024 + bucket) * 1024 + pos
    return 9 * 1024 * 1024 + ((level - 9) * 65536 + bucket) * 4096 + pos


class HashTable:
    """pool (uint32, capacity pages * 512), page_table, bucket_sizes: modified in place; pool_top advances."""

    def __init__(self, pool, page_table, bucket_sizes, pool_top, levels):
        self.pool, self.table, self.sizes, self.pool_top, self.levels = pool, page_table, bucket_sizes, int(pool_top), levels

    def _sys(self, ptr):
        return int(self.table[ptr // PAGE]) * PAGE + ptr % PAGE

    def _allocate_page(self, page):           # :794-807
        assert self.table[page] == 0
        self.table[page] = self.pool_top
        self.pool_top += 1
        assert self.pool_top * PAGE <= self.pool.size, "pool exhausted"

    def find_or_add_leaf(self, leaf):         # :470-513
        level = self.levels - 2
        bucket = murmur64(int(leaf)) & M32 & (buckets_per_level(level) - 1)
        g = bucket_global_index(level, bucket)
        size = int(self.sizes[g])
        w0, w1 = int(leaf) & M32, int(leaf) >> 32
        base = make_ptr(level, bucket, 0)
        for pindex in range(0, size, PAGE):   # find_leaf_node_in_bucket :196-262
            p = self._sys(base + pindex)
            for index in range(pindex, min(size, pindex + PAGE), 2):
                if self.pool[p] == w0 and self.pool[p + 1] == w1:
                    return base + index, False
                p += 2
        ptr = base + size                     # add_leaf_node :355-400
        if size % PAGE == 0 and self.table[ptr // PAGE] == 0:
            self._allocate_page(ptr // PAGE)
        p = self._sys(ptr)
        self.pool[p], self.pool[p + 1] = w0, w1
        self.sizes[g] = size + 2
        assert size + 2 < bucket_capacity(level)
        return ptr, True

    def find_or_add_interior(self, level, node):   # :514-560
        node = [int(x) for x in node]
        n = len(node)
        assert n > 1
        bucket = murmur32xn(node) & (buckets_per_level(level) - 1)
        g = bucket_global_index(level, bucket)
        size = int(self.sizes[g])
        base = make_ptr(level, bucket, 0)
        for pindex in range(0, size, PAGE):   # find_interior_node_in_bucket :264-358
            p = self._sys(base + pindex)
            page_end = min(size, pindex + PAGE)
            if pindex + n >= page_end:
                break                          # `return 0xFFFFFFFF`
            page_end -= n
            index = pindex
            while index < page_end:
                if [int(x) for x in self.pool[p:p + n]] == node:
                    return base + index, False
                length = bin(int(self.pool[p]) & 0xFF).count("1") + 1
                index += length
                p += length
        left = PAGE - size % PAGE             # add_interior_node :401-468
        if left == PAGE or left < n:
            if left != PAGE:
                size += left
            ptr = base + size
            if self.table[ptr // PAGE] == 0:
                self._allocate_page(ptr // PAGE)
        else:
            ptr = base + size
        p = self._sys(ptr)
        self.pool[p:p + n] = node
        self.sizes[g] = size + n
        assert size + n < bucket_capacity(level), "Bucket size too low"
        return ptr, True
