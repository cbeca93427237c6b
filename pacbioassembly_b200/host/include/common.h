// common.h -- limits and typedefs with the reference's names (src/common.h:31-59), for code that is compiled
// against this directory instead of the reference's src/.
#pragma once

#include <cstdio>
#include <list>

#define DBG
#ifdef DBG
#define LOG(...) std::fprintf(stderr, __VA_ARGS__)
#else
#define LOG(...)
#endif

#define MAX_SEQ_LEN 800000 // common.h:31
#define MAX_READ_LEN 20000 // common.h:33
#define MAX_DIFF_LEN 6000  // common.h:35
#define MAXR 0.3           // common.h:37
#ifndef OVERLAP_MIN
#define OVERLAP_MIN 64 // common.h:39 (overridable: the reference's test/ref_test.cpp aligns ~45-base strings and predates this value)
#endif

typedef unsigned t_seed;      // common.h:44
typedef unsigned char t_bseq; // common.h:49

#include "seed_index.h" // hash_table / sm_it (common.h:54,59) backed by the device index
