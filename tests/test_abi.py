"""The C-ABI library loads and exports every symbol include/hashdag_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hashdag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hdt_[a-z_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for name in ("hdt_create", "hdt_destroy", "hdt_resolve_paths", "hdt_resolve_colors", "hdt_resolve_shadows", "hdt_get_path"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "hashdag_b200", "libhashdag_b200.so"))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/hashdag_b200.h but not exported"


def test_python_mirror_lists_the_same_symbols():
    from hashdag_b200 import tracer
    assert sorted(tracer.EXPORTS) == declared_symbols()


def test_pod_sizes_match_reference_layouts():
    # sizes probed from the reference structs; the .cu file static_asserts the same
    from hashdag_b200 import tracer
    import struct
    assert len(struct.pack("<QQ", 0, 0)) == 16                      # BasicDAG
    leaf = tracer.CompressedColorLeaf(None, None, None)
    assert len(leaf.pod()) == 104                                   # CompressedColorLeaf
    assert len(tracer.HashDAG(None, None, 1, 0, 17).pod()) == 32    # HashDAG
    comp = tracer.BasicDAGCompressedColors(10, None, leaf)
    unc = tracer.BasicDAGUncompressedColors(10, None, None)
    assert len(comp.pod()) == 128 and len(unc.pod()) == 40
    assert len(tracer.BasicDAGColorErrors(comp, unc).pod()) == 288
    assert len(tracer.HashDAGColors(None, None, leaf).pod()) == 248
    assert ctypes.sizeof(tracer.ToolInfo) == 44
    # library formats next to the path
    from hashdag_b200 import color_leaf, edits
    assert len(tracer.ResolvedHashDAG(tracer.HashDAG(None, None, 1, 0, 17), None).pod()) == 48     # hdt_resolved_hash_dag
    assert color_leaf.OP_DTYPE.itemsize == 32 and edits.RANGE_DTYPE.itemsize == 24                  # hdt_color_op, hdt_range


def test_no_cpu_fallback():
    """Without a GPU the product refuses to run instead of routing through the oracle."""
    import pytest
    import torch
    from hashdag_b200 import tracer
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tracer.TracerError):
        tracer.DAGTracer(True, 64, 64, 12)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hashdag_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "hdo_" not in text or "oracle/hdo_oracle.cpp" in text, f"{f} references the oracle"
                assert "from oracle" not in text and "import oracle" not in text, f"{f} imports the oracle"
