// cuda_shim.h — just enough of the CUDA device vocabulary to compile the line-walker sources
// (flowgger_b200/csrc/fg_r5fast.cuh, fg_common.cuh, ...) with g++ for the CPU test-suite.
// TEST INFRASTRUCTURE: the emulation build (tests/emu/libfg_emu.so) runs the DEVICE LOGIC one lane at a time so that
// `pytest -m "not gpu"` can compare it with the oracle; nothing in the product libraries includes or links this.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __noinline__
#define __constant__
#define __forceinline__ inline
#define FG_DEV static inline

struct uint4 { uint32_t x, y, z, w; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

// IEEE round-to-nearest-even is the host default; these are exact stand-ins for the device intrinsics
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t bits = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &bits, 8);
    return d;
}
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}

// host stand-ins for the types fg_kernels.cuh mentions and the warp intrinsics the round-1 scanner calls directly
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
static inline int __any_sync(unsigned, int p) { return p; }
static inline void __syncwarp() {}
static inline int __ffs(int x) { return x ? __builtin_ctz((unsigned)x) + 1 : 0; }
#define __grid_constant__
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
using std::isinf;
static inline double __ll2double_rn(long long v) { return (double)v; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
