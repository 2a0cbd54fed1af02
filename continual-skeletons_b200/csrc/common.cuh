// Shared device helpers: split-bf16 state format, tile geometry.
//
// Activation state in HBM is token-major: row = token (skeleton*V + vertex), columns = channels,
// stored as two bf16 planes "hi" and "lo" with  x ~= hi + lo  (hi = rn_bf16(x), lo = rn_bf16(x - hi)),
// i.e. 4 bytes per element like fp32 but directly usable as tcgen05 kind::f16 operands
// (3-product error-compensated MMA: hi*hi + lo*hi + hi*lo, fp32 accumulate).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cosk {

constexpr int kTileRows = 128;   // UMMA M: token rows per tile (only skel_per_tile*V of them are valid)
constexpr int kRingSlots = 9;    // temporal ring, per-step stepping: 8 history frames + the frame being written
constexpr int kOutSlots = 5;     // block-output ring, per-step stepping: newest + 4 delayed (residual alignment)
// With a time chunk of Tc > 1 frames (cosk_set_batch_ex) the rings hold 8 + Tc and 4 + Tc slots, so that a whole chunk of
// one block's executions can be in flight before the next block starts (module-by-module over time, as the library's
// forward_steps runs, models/base.py:187-190).

// Time-batched launches: one launch covers n_frames consecutive executions of a kernel and its work items become
// (frame, tile) pairs.  Frame f of the launch lives in ring slot (slot0 + f * step) % slots.
struct RingWalk {
  int slot0 = 0, step = 1, slots = 1;
  __host__ __device__ __forceinline__ int slot(int f) const { return (slot0 + f * step) % slots; }
};
constexpr int kTaps = 9;         // temporal kernel size (models/base.py:284,310 in the reference)
constexpr int kResDelay = 4;     // every residual kind reads the block input of 4 executions ago

// Programmatic dependent launch: every kernel of a step is launched with the programmatic-stream-
// serialization attribute, lets its successor start launching right away (pdl_trigger) and waits for
// its predecessor's results only after its own prologue (pdl_wait).  Both are no-ops without PDL.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

// pack two floats as bf16x2 (first argument in the low half)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ float bf16_lo_as_float(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_as_float(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

}  // namespace cosk
