"""Build recipe for oracle/_ref/: the UNMODIFIED reference tracer compiled for sm_100a.

TEST INFRASTRUCTURE.  Sources are compiled from /root/reference (never copied into the repo): the
reference fixes depth and resolution at compile time and force-includes its own
script_definitions.h, so each (depth, width, height) variant is staged under /tmp, three things are
patched there (script_definitions.h, the two image-size constants in typedefs.h, a stub GL/glew.h
because GLEW is not installed and only the non-headless branch uses it) and nvcc writes one shared
library into oracle/_ref/.  That directory is git-ignored but travels to the GPU box.

    python oracle/build_ref.py            # all variants the tests and bench use
    python oracle/build_ref.py 17 1920 1080
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT_DIR = os.path.join(HERE, "_ref")
HARNESS = os.path.join(HERE, "ref_harness", "ref_harness.cu")
# compiled: the harness proper, the reference's CPU edit code, and the drop-in proof (the product's C++ shim driven with the
# reference's own types, linked against hashdag_b200/libhashdag_b200.so); the headers only count for staleness
HARNESS_FILES = [HARNESS, os.path.join(HERE, "ref_harness", "ref_harness_edits.cpp"), os.path.join(HERE, "ref_harness", "ref_dropin.cu"),
                 os.path.join(HERE, "ref_harness", "ref_harness_shared.h"), os.path.join(HERE, "ref_harness", "ref_harness_std.h"),
                 os.path.join(HERE, "ref_harness", "ref_harness_globals.h"),
                 os.path.join(os.path.dirname(HERE), "hashdag_b200", "cpp", "dag_tracer_b200.h"), os.path.join(os.path.dirname(HERE), "include", "hashdag_b200.h")]
PRODUCT_DIR = os.path.join(os.path.dirname(HERE), "hashdag_b200")

# (depth, width, height): golden/parity variants are small, the bench variant is the headline config
VARIANTS = [
    (12, 256, 256),
    (13, 256, 256),
    (17, 256, 256),
    (16, 1920, 1080),
    (17, 1920, 1080),
    (17, 3840, 2160),
]

REF_FILES = ["dag_tracer.cu", "tracer.cu", "dags/hash_dag/hash_table.cpp", "dags/hash_dag/hash_dag_factory.cpp",
             "dags/basic_dag/basic_dag.cpp", "memory.cpp", "stats.cpp"]

GLEW_STUB = """#pragma once
// Stand-in for GLEW: only DAGTracer's non-headless branch touches GL (dag_tracer.cu:24-41,58-64).
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef int GLsizei;
#define GL_TEXTURE_2D 0x0DE1
#define GL_RGBA32UI 0x8D70
#define GL_RGBA_INTEGER 0x8D99
#define GL_UNSIGNED_INT 0x1405
#define GL_RGBA 0x1908
#define GL_UNSIGNED_BYTE 0x1401
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_NEAREST 0x2600
inline void glGenTextures(GLsizei, GLuint*) {}
inline void glBindTexture(GLenum, GLuint) {}
inline void glTexImage2D(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void*) {}
inline void glTexParameteri(GLenum, GLenum, GLint) {}
inline void glDeleteTextures(GLsizei, const GLuint*) {}
"""


def lib_path(depth, width, height, overlay=False):
    return os.path.join(OUT_DIR, f"libhashdag_ref_d{depth}_{width}x{height}{'_overlay' if overlay else ''}.so")


def build_variant(depth, width, height, force=False, verbose=True, overlay=False):
    """overlay=True: the same build with TOOL_OVERLAY 1 (off under BENCHMARK, typedefs.h:70-72), to pin the tool overlay."""
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: prebuilt libraries only
    out = lib_path(depth, width, height, overlay)
    if os.path.exists(out) and not force and os.path.getmtime(out) >= max([os.path.getmtime(f) for f in HARNESS_FILES] + [os.path.getmtime(__file__)]):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    stage = f"/tmp/hashdag_ref_stage_d{depth}_{width}x{height}{'_overlay' if overlay else ''}"
    shutil.rmtree(stage, ignore_errors=True)
    shutil.copytree(REF_SRC, os.path.join(stage, "src"))
    os.makedirs(os.path.join(stage, "stub", "GL"))
    for name in ("glew.h", "gl.h"):
        with open(os.path.join(stage, "stub", "GL", name), "w") as f:
            f.write(GLEW_STUB)
    with open(os.path.join(stage, "src", "script_definitions.h"), "w") as f:
        f.write(f"#define SCENE_DEPTH {depth}\n#define REPLAY_DEPTH {depth}\n#define BENCHMARK 1\n#define HEADLESS 1\n#define ENABLE_CHECKS 0\n"
                # the library is dlopen()ed into python: keep the reference from replacing global new/delete
                "#define TRACK_GLOBAL_NEWDELETE 0\n" + ("#define TOOL_OVERLAY 1\n" if overlay else ""))
    tpath = os.path.join(stage, "src", "typedefs.h")
    text = open(tpath, encoding="utf-8-sig").read()
    text, n1 = re.subn(r"constexpr uint32 imageWidth = \d+;", f"constexpr uint32 imageWidth = {width};", text)
    text, n2 = re.subn(r"constexpr uint32 imageHeight = \d+;", f"constexpr uint32 imageHeight = {height};", text)
    assert n1 == 1 and n2 == 1, "typedefs.h image-size constants not found"
    open(tpath, "w").write(text)
    # .cu files (and the harness) are CUDA; the reference's .cpp files are host C++ that only
    # needs the CUDA headers -- the same split its CMakeLists.txt makes.
    src = [os.path.join(stage, "src", f) for f in REF_FILES] + HARNESS_FILES[:3]
    common = ["nvcc", "-std=c++17", "--expt-relaxed-constexpr", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-w", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-gnu-unique",  # several variants share one process
              "-I" + os.path.join(stage, "stub"), "-I" + os.path.join(stage, "src"), "-I" + os.path.join(HERE, "ref_harness"),
              "-I" + os.path.join(PRODUCT_DIR, "cpp")]
    if verbose:
        print(f"[build_ref] d{depth} {width}x{height} -> {os.path.relpath(out, HERE)}", flush=True)
    procs, objs = [], []
    for i, f in enumerate(src):
        obj = os.path.join(stage, f"obj{i}.o")
        objs.append(obj)
        procs.append((f, subprocess.Popen(common + ["-c", f, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for f, pr in procs:
        log, _ = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(log[-8000:])
            raise RuntimeError(f"reference build failed for d{depth} {width}x{height}: {f}")
    # the drop-in proof calls the product's C ABI: link the product library, found at run time relative to oracle/_ref/
    r = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs +
                       ["-L" + PRODUCT_DIR, "-lhashdag_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../hashdag_b200"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:] + r.stderr[-8000:])
        raise RuntimeError(f"reference link failed for d{depth} {width}x{height}")
    if not os.environ.get("HDT_KEEP_STAGE"):
        shutil.rmtree(stage, ignore_errors=True)
    return out


def build_all(force=False):
    return [build_variant(*v, force=force) for v in VARIANTS] + [build_variant(13, 256, 256, force=force, overlay=True)]


if __name__ == "__main__":
    if len(sys.argv) == 4:
        build_variant(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), force=True)
    else:
        build_all(force="--force" in sys.argv)
