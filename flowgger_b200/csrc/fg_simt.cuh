// fg_simt.cuh — the handful of warp / bit primitives the line walkers are written against.
//
// On the device they are the hardware instructions (VOTE, FLO/BREV, SHF.R.W).  tests/emu/ compiles the same walker
// sources with g++ and FG_HOST_EMU defined, where a "warp" is one lane: that build exists ONLY so the CPU test-suite can
// check the device logic against the oracle without a GPU; it is never linked into the product libraries.
#pragma once
#include <stdint.h>

#ifdef FG_HOST_EMU
#include "../../tests/emu/cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif

#ifndef FG_DEV
#define FG_DEV __device__ __forceinline__
#endif

namespace fg {

#ifdef FG_HOST_EMU
#ifdef FG_HOST_EMU_WARP
// tests/emu/emu_r3164.cpp built a second time with 32 host threads per warp: a vote is a rendezvous of the 32 lanes, and the
// harness checks that all of them arrive at the SAME vote (source line) — the convergence protocol of the lock-step walkers
bool fg_emu_vote(bool p, int site);
void fg_emu_sync(int site);
FG_DEV bool fg_any(bool p, int site = __builtin_LINE()) { return fg_emu_vote(p, site); }
FG_DEV void fg_syncwarp(int site = __builtin_LINE()) { fg_emu_sync(site); }
#else
FG_DEV bool fg_any(bool p) { return p; }
FG_DEV void fg_syncwarp() {}
#endif
FG_DEV int fg_ffs(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
FG_DEV int fg_clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
FG_DEV uint32_t fg_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
#else
FG_DEV bool fg_any(bool p) { return __any_sync(0xFFFFFFFFu, p) != 0; }
FG_DEV void fg_syncwarp() { __syncwarp(); }
FG_DEV int fg_ffs(uint32_t x) { return __ffs((int)x); }
FG_DEV int fg_clz(uint32_t x) { return __clz((int)x); }
FG_DEV uint32_t fg_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
#endif

}  // namespace fg
