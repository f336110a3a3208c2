// Headless C++ harness over the DAGTracer shim (hashdag_b200/cpp/dag_tracer_b200.h): the ~100-line
// stand-in for the reference's engine loop (engine.cpp:575-648).  Reads a scene dump written by
// tests/test_gpu_cpp_shim.py, renders one frame with BasicDAG and HashDAG, writes paths + colours.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../hashdag_b200/cpp/dag_tracer_b200.h"

using namespace hashdag_b200;

static std::vector<char> read_blob(FILE* f)
{
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); }
    std::vector<char> v(n);
    if (n && fread(v.data(), 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}
static const void* upload(const std::vector<char>& v)
{
    if (v.empty()) return nullptr;
    void* d = nullptr;
    if (cudaMalloc(&d, v.size()) != cudaSuccess || cudaMemcpy(d, v.data(), v.size(), cudaMemcpyHostToDevice) != cudaSuccess) { fprintf(stderr, "upload failed\n"); exit(3); }
    return d;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: shim_harness scene.bin out.bin\n"); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    uint32_t hdr[8];
    double cam[12 + 6];   // position(3) rotation(9) boundsMin(3) boundsMax(3)
    if (fread(hdr, 4, 8, f) != 8 || fread(cam, 8, 18, f) != 18) return 2;
    const uint32_t levels = hdr[0], W = hdr[1], H = hdr[2], firstNode = hdr[3], poolTop = hdr[4], topLevels = hdr[5];
    auto basic = read_blob(f), enclosed = read_blob(f), pool = read_blob(f), pageTable = read_blob(f);
    auto weights = read_blob(f), blocks = read_blob(f), macro = read_blob(f), colorNodes = read_blob(f), colorOffsets = read_blob(f);
    fclose(f);

    CameraView view;
    view.position = { cam[0], cam[1], cam[2] };
    for (int i = 0; i < 9; ++i) view.rotation[i / 3][i % 3] = cam[3 + i];
    DAGInfo info{ { cam[12], cam[13], cam[14] }, { cam[15], cam[16], cam[17] } };

    hdt_basic_dag basicDag{ { upload(basic), basic.size() / 4 } };
    hdt_hash_dag hashDag{ uint32_t(pageTable.size() / 4), poolTop, (const uint32_t*)upload(pageTable), (const uint32_t*)upload(pool), firstNode, 0 };
    hdt_color_leaf leaf{};
    leaf.offset = ~uint64_t(0);
    leaf.weights_gpu = { upload(weights), weights.size() / 4 };
    leaf.blocks_gpu = { upload(blocks), blocks.size() / 8 };
    leaf.macro_blocks_gpu = { upload(macro), macro.size() / 8 };
    hdt_basic_compressed_colors basicColors{ { topLevels, 0, { upload(enclosed), enclosed.size() / 8 } }, leaf };
    hdt_hash_colors hashColors{};
    hashColors.nodes_gpu = { upload(colorNodes), colorNodes.size() / 4, colorNodes.size() / 4 };
    hashColors.offsets_gpu = { upload(colorOffsets), colorOffsets.size() / 8, colorOffsets.size() / 8 };
    hashColors.main_leaf = leaf;

    DAGTracer tracer(false, W, H, levels);
    std::vector<uint32_t> paths(size_t(W) * H * 4), colors(size_t(W) * H);
    FILE* o = fopen(argv[2], "wb");
    const hdt_tool_info tool{};
    float ms[3];
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0) {
            ms[0] = tracer.resolve_paths(view, info, basicDag);
            ms[1] = tracer.resolve_colors(basicDag, basicColors, HDT_DEBUG_NONE, 0, tool);
            ms[2] = tracer.resolve_shadows(view, info, basicDag, 1.0f, 0.0f);
        } else {
            if (colorNodes.empty()) break;
            ms[0] = tracer.resolve_paths(view, info, hashDag);
            ms[1] = tracer.resolve_colors(hashDag, hashColors, HDT_DEBUG_NONE, 0, tool);
            ms[2] = tracer.resolve_shadows(view, info, hashDag, 1.0f, 0.0f);
        }
        tracer.read_paths(paths.data());
        tracer.read_colors(colors.data());
        const uint3_t centre = tracer.get_path(W / 2, H / 2);
        const uint32_t* c = &paths[(size_t(H / 2) * W + W / 2) * 4];
        if (centre.x != c[0] || centre.y != c[1] || centre.z != c[2]) { fprintf(stderr, "get_path mismatch\n"); return 4; }
        fwrite(paths.data(), 4, paths.size(), o);
        fwrite(colors.data(), 4, colors.size(), o);
        printf("pass %d: paths %.3f ms, colors %.3f ms, shadows %.3f ms\n", pass, ms[0], ms[1], ms[2]);
    }
    fclose(o);
    return 0;
}
