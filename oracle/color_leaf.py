"""CPU oracle of the colour-leaf rebuild.  TEST INFRASTRUCTURE: only tests/, smoke() and the
`cpu_baseline` legs of the bench scripts may import this; the product (hashdag_b200/) never does.

Restates, for host arrays, /root/reference/src/variable_weight_size_colors.h:
  * `get_color_loop`      CompressedColorLeaf::get_color            :405-413  (binary_search_blocks :343-372,
                          get_color_for_block :374-403, ColorUtils::extract_bits color_utils.h:137-146)
  * `ColorLeafBuilderPort` ColorLeafBuilder::add / add_weight / add_large_single_color / build   :549-721,
                          one colour at a time like the reference (pure Python: small cases only)
  * `decode_range` / `encode` the same two things vectorised with numpy (what the parity tests run at size)
  * `expand_ops`          what hash_dag_edits.h:381-396 / :430-437 / :488-516 feed the builder: copy_colors of the
                          old leaf (vwsc.h:416-542) and constant colours

Pinning: the reference has no golden vectors for its colour codec.  tests/golden/ref_color_leaves_d13.npz holds
leaves written by the reference's own ColorLeafBuilder::build during SphereEditor edits (oracle/_ref on a B200,
tests/golden/make_color_leaf_golden.py); `encode(decode_range(leaf))` must reproduce every one of them byte for
byte (tests/test_color_leaf_cpu.py), and the two restatements must agree with each other.
"""
from __future__ import annotations

import numpy as np

COLORS_PER_MACRO = 16 * 1024          # CColorsPerMacroBlock, vwsc.h:8
OP_COPY, OP_FILL = 0, 1               # hdt_color_op.kind
OP_DTYPE = np.dtype([("src_start", "<u8"), ("count", "<u8"), ("kind", "<u4"), ("bits_per_weight", "<u4"), ("color_bits", "<u4"), ("weight", "<u4")])


# ------------------------------------------------------------------------------------------------------------
# scalar restatements (follow the reference statement by statement)
# ------------------------------------------------------------------------------------------------------------
def header_bits_per_weight(header: int) -> int:          # vwsc.h:20-30
    return 0 if (header >> 16) == 0xFFFF else ((header >> 14) & 3) + 1


def make_block_header(weight_offset: int, bits_per_weight: int, index: int) -> int:   # vwsc.h:32-52
    if bits_per_weight == 0:
        weight_offset = 0xFFFF
    else:
        bits_per_weight -= 1
    return ((weight_offset << 16) | (bits_per_weight << 14) | index) & 0xFFFFFFFF


def get_color_loop(weights, blocks, macro_blocks, color_index: int, offset=None):
    """-> (color_bits, weight, bits_per_weight) of colour `color_index` (vwsc.h:343-413)."""
    if offset is not None:
        color_index += int(offset)
    macro = color_index // COLORS_PER_MACRO
    local = color_index % COLORS_PER_MACRO
    lo = int(macro_blocks[2 * macro])
    hi = int(macro_blocks[2 * (macro + 1)]) - 1 if 2 * (macro + 1) < len(macro_blocks) else len(blocks) - 1
    pos = (lo + hi) // 2
    idx = int(blocks[pos]) & 0x3FFF
    while idx != local and lo <= hi:
        if idx > local:
            hi = pos - 1
        else:
            lo = pos + 1
        pos = (lo + hi) // 2
        idx = int(blocks[pos]) & 0x3FFF
    block = int(blocks[pos])
    header = block & 0xFFFFFFFF
    bpw = header_bits_per_weight(header)
    weight = 0
    if bpw:
        bit = int(macro_blocks[2 * macro + 1]) + (header >> 16) + (local - (header & 0x3FFF)) * bpw
        by = np.asarray(weights).view(np.uint8)
        nxt = int(by[(bit >> 3) + 1]) if (bit >> 3) + 1 < by.size else 0      # the reference's 16-bit load may run one byte past the array; those bits are shifted out
        be16 = (int(by[bit >> 3]) << 8) | nxt                                 # extract_bits, swapped byte order
        weight = (be16 >> (16 - bpw - (bit & 7))) & ((1 << bpw) - 1)
    return block >> 32, weight, bpw


class ColorLeafBuilderPort:
    """ColorLeafBuilder (vwsc.h:546-721), member for member."""

    def __init__(self):
        self.blocks, self.weights, self.macro_blocks = [], [], []   # BlockStruct / uint32 / MacroBlockStruct
        self.num_weights = 0
        self.target_bit_pos = 0
        self.color_index = 0
        self.last_block_bpw = 0
        self.last_block_color_bits = 0
        self.last_macro_start_color_index = 0
        self.last_macro_start_weight_index = 0

    def add_weight(self, weight, bpw):                   # :552-580 (uint32 targetBitPos wraps like the reference's)
        target = (self.target_bit_pos + 32 - bpw) % 64
        ww = (weight << target) & 0xFFFFFFFFFFFFFFFF
        if target > 32 - bpw:
            self.weights[-1] |= ww >> 32
            self.target_bit_pos = (self.target_bit_pos - bpw) & 0xFFFFFFFF
        if target < 32:
            self.weights.append(ww & 0xFFFFFFFF)
            self.target_bit_pos = target
        self.num_weights += bpw

    def add(self, color_bits, weight, bpw):              # :582-613
        if self.color_index % COLORS_PER_MACRO == 0:
            self.last_macro_start_color_index = self.color_index
            self.last_macro_start_weight_index = self.num_weights
            self.macro_blocks.append((self.num_weights, len(self.blocks)))
        if self.color_index % COLORS_PER_MACRO == 0 or self.last_block_color_bits != color_bits or self.last_block_bpw != bpw:
            self.last_block_color_bits, self.last_block_bpw = color_bits, bpw
            self.blocks.append((bpw, color_bits, self.num_weights - self.last_macro_start_weight_index, self.color_index - self.last_macro_start_color_index))
        if bpw > 0:
            self.add_weight(weight, bpw)
        self.color_index += 1

    def add_large_single_color(self, color_bits, num_voxels):   # :614-641 (colour = set_single_color: bpw 0)
        left = num_voxels
        while left > 0:
            in_macro = COLORS_PER_MACRO - (self.color_index % COLORS_PER_MACRO)
            n = min(in_macro, left)
            left -= n
            self.add(color_bits, 0, 0)
            self.color_index += n - 1

    def build(self):                                     # :643-685 -> (weights, blocks, macro_blocks) as stored in the leaf
        blocks = np.array([(cb << 32) | make_block_header(wo, bpw, idx) for bpw, cb, wo, idx in self.blocks], dtype=np.uint64)
        weights = np.array(self.weights, dtype=np.uint32).byteswap()
        macro = np.zeros(2 * len(self.macro_blocks), dtype=np.uint64)
        for i, (start_weight, start_block) in enumerate(self.macro_blocks):
            macro[2 * i], macro[2 * i + 1] = start_block, start_weight
        return weights, blocks, macro


# ------------------------------------------------------------------------------------------------------------
# vectorised
# ------------------------------------------------------------------------------------------------------------
def leaf_color_count(weights, blocks, macro_blocks) -> int:
    """The largest number of colours consistent with a UNIQUE leaf's arrays.  The format does not store the
    length of the last block: with weights it is bounded by the (word-padded) weight stream, without it is taken
    as 1.  Colours decoded from the zero padding re-encode to the same bytes, so `encode(decode_range(leaf, 0,
    leaf_color_count(leaf)))` reproduces the leaf exactly iff the encoder is right."""
    if len(blocks) == 0:
        return 0
    n_macro = len(macro_blocks) // 2
    header = int(blocks[-1]) & 0xFFFFFFFF
    start = (n_macro - 1) * COLORS_PER_MACRO + (header & 0x3FFF)
    bpw = header_bits_per_weight(header)
    if bpw == 0:
        return start + 1
    first_bit = int(macro_blocks[2 * n_macro - 1]) + (header >> 16)
    return min(start + (32 * len(weights) - first_bit) // bpw, n_macro * COLORS_PER_MACRO)


def decode_range(weights, blocks, macro_blocks, start: int, count: int, offset=None):
    """-> (color_bits u32[count], weight u32[count], bits_per_weight u32[count]) of colours start .. start+count-1."""
    if count == 0:
        z = np.zeros(0, np.uint32)
        return z, z.copy(), z.copy()
    blocks = np.asarray(blocks, dtype=np.uint64)
    macro_blocks = np.asarray(macro_blocks, dtype=np.uint64)
    idx = np.arange(count, dtype=np.int64) + int(start) + (int(offset) if offset is not None else 0)
    macro = idx // COLORS_PER_MACRO
    local = idx % COLORS_PER_MACRO
    first = macro_blocks[0::2].astype(np.int64)
    # macro block of every block, then a global sort key (macro, colour index): blocks are sorted by it
    bmacro = np.searchsorted(first, np.arange(len(blocks), dtype=np.int64), side="right") - 1
    hdr = (blocks & np.uint64(0xFFFFFFFF)).astype(np.int64)
    key = bmacro * COLORS_PER_MACRO + (hdr & 0x3FFF)
    pos = np.searchsorted(key, idx, side="right") - 1
    h = hdr[pos]
    cb = (blocks[pos] >> np.uint64(32)).astype(np.uint32)
    bpw = np.where((h >> 16) == 0xFFFF, 0, ((h >> 14) & 3) + 1).astype(np.int64)
    weight = np.zeros(count, dtype=np.int64)
    m = bpw > 0
    if m.any():
        by = np.asarray(weights).view(np.uint8)
        bit = macro_blocks[1::2].astype(np.int64)[macro[m]] + (h[m] >> 16) + (local[m] - (h[m] & 0x3FFF)) * bpw[m]
        by = np.concatenate((by, np.zeros(2, np.uint8)))
        be16 = (by[bit >> 3].astype(np.int64) << 8) | by[(bit >> 3) + 1].astype(np.int64)
        weight[m] = (be16 >> (16 - bpw[m] - (bit & 7))) & ((1 << bpw[m]) - 1)
    return cb, weight.astype(np.uint32), bpw.astype(np.uint32)


def encode(color_bits, weight, bits_per_weight):
    """ColorLeafBuilder fed colour by colour, then build(): -> (weights u32[], blocks u64[], macro_blocks u64[])."""
    cb = np.asarray(color_bits, dtype=np.uint64)
    w = np.asarray(weight, dtype=np.uint64)
    bpw = np.asarray(bits_per_weight, dtype=np.int64)
    n = cb.size
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    i = np.arange(n, dtype=np.int64)
    start = (i % COLORS_PER_MACRO) == 0
    start[1:] |= (cb[1:] != cb[:-1]) | (bpw[1:] != bpw[:-1])
    bit_off = np.concatenate(([0], np.cumsum(bpw)[:-1]))                     # numWeights before each colour
    macro_first = (i // COLORS_PER_MACRO) * COLORS_PER_MACRO
    s = np.flatnonzero(start)
    wo = (bit_off[s] - bit_off[macro_first[s]])
    f = np.where(bpw[s] == 0, 0xFFFF, wo)
    b = np.where(bpw[s] == 0, 0, bpw[s] - 1)
    header = ((f << 16) | (b << 14) | (s - macro_first[s])) & 0xFFFFFFFF
    blocks = (cb[s] << np.uint64(32)) | header.astype(np.uint64)
    block_index = np.cumsum(start) - 1
    m = np.arange(0, n, COLORS_PER_MACRO)
    macro = np.zeros(2 * m.size, dtype=np.uint64)
    macro[0::2] = block_index[m]
    macro[1::2] = bit_off[m]
    total_bits = int(bit_off[-1] + bpw[-1])
    n_words = (total_bits + 31) // 32
    bits = np.zeros(n_words * 32, dtype=np.uint8)
    for k in range(4):                                   # bit k (from the MSB) of every weight that has > k bits
        sel = bpw > k
        bits[bit_off[sel] + k] = ((w[sel] >> (bpw[sel] - 1 - k).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
    weights = np.packbits(bits).view(np.uint32).copy()   # big-endian bit stream == the byte-swapped words of build()
    return weights, blocks, macro


def expand_ops(ops, old_leaf=None):
    """ops: array of OP_DTYPE.  old_leaf: (weights, blocks, macro_blocks, offset_or_None).  -> colour stream."""
    cbs, ws, bs = [], [], []
    for o in np.asarray(ops, dtype=OP_DTYPE):
        n = int(o["count"])
        if n == 0:
            continue
        if int(o["kind"]) == OP_COPY:
            cb, w, b = decode_range(old_leaf[0], old_leaf[1], old_leaf[2], int(o["src_start"]), n, old_leaf[3] if len(old_leaf) > 3 else None)
        else:
            cb = np.full(n, int(o["color_bits"]), np.uint32)
            w = np.full(n, int(o["weight"]), np.uint32)
            b = np.full(n, int(o["bits_per_weight"]), np.uint32)
        cbs.append(cb); ws.append(w); bs.append(b)
    if not cbs:
        z = np.zeros(0, np.uint32)
        return z, z.copy(), z.copy()
    return np.concatenate(cbs), np.concatenate(ws), np.concatenate(bs)


def rebuild(ops, old_leaf=None):
    """The whole oracle: op list -> (weights, blocks, macro_blocks)."""
    return encode(*expand_ops(ops, old_leaf))
