"""Batches for the hash-table insert (HashTable::find_or_add_*, hash_table.h:470-560), shared by the fixture generator
(tests/golden/make_find_or_add_golden.py), the oracle test and the GPU test.  Everything is derived from the recipe
scene and fixed seeds; the batches are applied one after the other to the same table.

What they cover: nodes that are already in the table, new ones, a new node several times in one batch, leaves,
buckets growing over one and several page boundaries (padding in front of a boundary, pages opened in insertion order)
and the reference's end-of-page rule (hash_table.h:327-329: a node sitting in the last words of a partly filled page
is not found again)."""
import numpy as np

RECIPE = "d13"
SPARE_PAGES = 4096   # pool pages beyond the scene's for the batches to open


def _hash32xn(rows):
    """Utils::murmurhash32xN (utils.h:91-110) of every row of a 2-D uint32 array."""
    rows = rows.astype(np.uint32)
    h = np.zeros(rows.shape[0], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for j in range(rows.shape[1]):
            k = rows[:, j] * np.uint32(0xcc9e2d51)
            k = (k << np.uint32(15)) | (k >> np.uint32(17))
            k = k * np.uint32(0x1b873593)
            h ^= k
            h = (h << np.uint32(13)) | (h >> np.uint32(19))
            h = h * np.uint32(5) + np.uint32(0xe6546b64)
        h ^= np.uint32(rows.shape[1])
        h ^= h >> np.uint32(16); h = h * np.uint32(0x85ebca6b); h ^= h >> np.uint32(13); h = h * np.uint32(0xc2b2ae35); h ^= h >> np.uint32(16)
    return h


def _hash64(v):
    """uint32(Utils::murmurhash64) (utils.h:77-85) of a uint64 array."""
    h = v.astype(np.uint64)
    with np.errstate(over="ignore"):
        h ^= h >> np.uint64(33); h = h * np.uint64(0xff51afd7ed558ccd); h ^= h >> np.uint64(33)
        h = h * np.uint64(0xc4ceb9fe1a85ec53); h ^= h >> np.uint64(33)
    return (h & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def nodes_by_level(pool, table, root, levels):
    """{level: sorted unique virtual pointers} of the nodes reachable from `root` (levels-2 = the 64-bit leaves)."""
    out, cur = {}, np.array([root], dtype=np.uint32)
    for level in range(levels - 1):
        out[level] = cur
        if level == levels - 2:
            break
        nxt = []
        for p in cur:
            s = int(table[p >> 9]) * 512 + (int(p) & 511)
            n = bin(int(pool[s]) & 0xFF).count("1")
            nxt.append(pool[s + 1:s + 1 + n])
        cur = np.unique(np.concatenate(nxt))
    return out


def read_node(pool, table, ptr, leaf):
    s = int(table[int(ptr) >> 9]) * 512 + (int(ptr) & 511)
    n = 2 if leaf else bin(int(pool[s]) & 0xFF).count("1") + 1
    return np.array(pool[s:s + n], dtype=np.uint32)


def _random_interior(rng, count, size, buckets, level_buckets, pointer_pool):
    """`count` random nodes of `size` words whose bucket index is below `buckets` (rejection sampling)."""
    out = []
    mask = {k: m for k, m in ((2, 0x10), (3, 0x41), (4, 0x0B), (5, 0x1E), (6, 0x9B), (7, 0x7E), (8, 0xFE), (9, 0xFF))}[size]
    while len(out) < count:
        rows = np.empty((4096, size), dtype=np.uint32)
        rows[:, 0] = mask | (rng.integers(1, 1 << 20, 4096).astype(np.uint32) << np.uint32(8))
        rows[:, 1:] = rng.choice(pointer_pool, (4096, size - 1))
        keep = rows[(_hash32xn(rows) & np.uint32(level_buckets - 1)) < buckets]
        out.extend(list(keep))
    return out[:count]


def cases(pool, table, root, levels):
    """-> [(name, level, leaves, [uint32 arrays])].  `pool`/`table`: the table BEFORE the first batch (not modified)."""
    rng = np.random.default_rng(20240607)
    by_level = nodes_by_level(pool, table, root, levels)
    leaf_level = levels - 2
    out = []

    # 1. parents of leaves: known nodes, nodes with one child swapped, some of the new ones twice
    lvl = leaf_level - 1
    have = [read_node(pool, table, p, False) for p in rng.choice(by_level[lvl], 300, replace=False)]
    leaf_ptrs = by_level[leaf_level]
    new = []
    for w in have:
        m = w.copy()
        m[1 + int(rng.integers(0, len(m) - 1))] = rng.choice(leaf_ptrs)
        new.append(m)
    batch = have + new + [new[i].copy() for i in rng.choice(len(new), 100, replace=False)]
    out.append(("leaf parents", lvl, False, [batch[i] for i in rng.permutation(len(batch))]))

    # 2. leaves: known, new, new twice
    have = [read_node(pool, table, p, True) for p in rng.choice(leaf_ptrs, 300, replace=False)]
    new = [rng.integers(1, 1 << 32, 2).astype(np.uint32) for _ in range(300)]
    batch = have + new + [new[i].copy() for i in rng.choice(len(new), 100, replace=False)]
    out.append(("leaves", leaf_level, True, [batch[i] for i in rng.permutation(len(batch))]))

    # 3./4. a level with 1024-word buckets: full-size nodes into 16 buckets until most of them cross their page boundary; then
    #       the same nodes again (all found -- except those the end-of-page rule hides, which are added a second time)
    crowd = _random_interior(rng, 1150, 9, 16, 1024, by_level[6])
    out.append(("crowded small buckets", 5, False, crowd))
    out.append(("crowded small buckets again", 5, False, [w.copy() for w in crowd]))

    # 5. a level with 4096-word buckets: nodes of every size into 8 buckets, several page boundaries each, padding of all widths
    sizes = rng.integers(2, 10, 3600)
    mixed = []
    for size in range(2, 10):
        mixed.extend(_random_interior(rng, int((sizes == size).sum()), size, 8, 65536, by_level[10]))
    mixed = [mixed[i] for i in rng.permutation(len(mixed))]
    out.append(("mixed sizes", 9, False, mixed))
    again = [mixed[i].copy() for i in rng.choice(len(mixed), 900, replace=False)]
    out.append(("mixed sizes again", 9, False, again))

    # 6. leaves into 4 buckets over two page boundaries
    many = []
    while len(many) < 2400:
        v = rng.integers(1, 1 << 63, 8192, dtype=np.uint64)
        keep = v[(_hash64(v) & np.uint32(65535)) < 4]
        many.extend(np.array([x & np.uint64(0xFFFFFFFF), x >> np.uint64(32)], dtype=np.uint64).astype(np.uint32) for x in keep)
    many = many[:2400]
    out.append(("crowded leaf buckets", leaf_level, True, many + [many[i].copy() for i in range(0, 2400, 7)]))
    return out


def overflow_case(levels):
    """A batch one 1024-word bucket cannot take (hash_table.h:461 `Bucket size on level %u too low`)."""
    rng = np.random.default_rng(5)
    return 4, False, _random_interior(rng, 130, 9, 1, 1024, np.arange(100, 4000, dtype=np.uint32))


def fresh_table(scene):
    """(pool with SPARE_PAGES zero pages behind the scene's, page table, bucket fill counts, pool top): copies."""
    pool = np.zeros((scene.hash_pool_top + SPARE_PAGES) * 512, dtype=np.uint32)
    pool[:scene.hash_pool.size] = scene.hash_pool
    return pool, scene.hash_page_table.copy(), scene.hash_bucket_sizes.copy(), scene.hash_pool_top


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
