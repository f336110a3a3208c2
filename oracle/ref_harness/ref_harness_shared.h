// Shared state of the reference harness (ref_harness.cu: tracer side, compiled by nvcc; ref_harness_edits.cpp:
// the reference's CPU edit code, which only instantiates under a host compiler -- hash_dag_edits.h replaces its
// `if constexpr` by `if (true)` when __CUDACC__ is defined).  TEST INFRASTRUCTURE, see ref_harness.cu.
#pragma once
#include "ref_harness_std.h"

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

#include "ref_harness_globals.h"
