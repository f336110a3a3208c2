// The one scene a harness library holds (the reference keeps its HashTable in static storage: one scene per process and
// library).  Include after the reference's DAG headers.  TEST INFRASTRUCTURE, see ref_harness.cu.
#pragma once
namespace refh {
extern BasicDAG g_basic;
extern BasicDAGCompressedColors g_compressed;
extern BasicDAGUncompressedColors g_uncompressed;
extern BasicDAGColorErrors g_errors;
extern HashDAG g_hash;
extern HashDAGColors g_hashColors;
extern bool g_hasHash, g_hasHashColors;
}
