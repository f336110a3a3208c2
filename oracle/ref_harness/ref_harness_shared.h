// Shared state of the reference harness (ref_harness.cu: tracer side, compiled by nvcc; ref_harness_edits.cpp:
// the reference's CPU edit code, which only instantiates under a host compiler -- hash_dag_edits.h replaces its
// `if constexpr` by `if (true)` when __CUDACC__ is defined).  TEST INFRASTRUCTURE, see ref_harness.cu.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <mutex>
#include <atomic>
#include <unordered_map>
#include <array>
#include <random>
#include <set>
#include <limits>
#include <chrono>
#include <type_traits>
#include <cmath>
#include <sstream>
#include <fstream>
#include <iostream>
#include <iomanip>
#include <thread>
#include <unordered_set>
#include <map>
#include <functional>
#include <algorithm>
#include <numeric>
#include <condition_variable>
#include <future>
#include <queue>
#include <deque>
#include <list>
#include <cassert>
#include <cstdlib>
#include <filesystem>

// The frame surfaces, the HashTable pointers and the colour arrays are private in the reference
// and it has no full-frame read-back; the harness alone looks inside.
#define private public
#define protected public
#include "dag_tracer.h"
#include "dags/basic_dag/basic_dag.h"
#include "dags/hash_dag/hash_dag.h"
#include "dags/hash_dag/hash_dag_colors.h"
#include "dags/hash_dag/hash_dag_factory.h"
#include "memory.h"
#include "stats.h"
#undef private
#undef protected

namespace refh {
extern HashDAG g_hash;
extern HashDAGColors g_hashColors;
extern bool g_hasHash, g_hasHashColors;
}
