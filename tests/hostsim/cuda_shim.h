// TEST INFRASTRUCTURE ONLY — never built into or loaded by the riichienv_b200 package.
//
// Lets g++ compile the *device* sources (riichienv_b200/csrc/{tables,hand,game}.cuh) as
// ordinary host C++ so the kernel logic can be exercised against the oracle on a box
// without a GPU (this authoring container has none; GPU minutes are budgeted).  It is a
// debugging aid for the CUDA source, not a CPU fallback: the shipped library has no
// host execution path and fails with RV_ERR_CUDA when no device is present.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
using std::max;
using std::min;
