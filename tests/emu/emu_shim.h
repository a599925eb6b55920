// TEST INFRASTRUCTURE - host stand-ins for the CUDA builtins used by csrc/epilogue_staged.cuh, so that the very same kernel
// source runs on the CPU (g++, one std::thread per CUDA thread of a CTA, CTAs one after the other) and can be checked
// against oracle/wave_spec.py without a GPU.  Never part of libelastic_b200.so, never loaded by the package.
//
// What is emulated: threadIdx / blockIdx / blockDim, __syncthreads (a barrier over the CTA's threads), static and dynamic
// shared memory (one CTA at a time), the mbarrier transaction count, cp.async.bulk.tensor.3d as a synchronous box copy
// with out-of-bounds zero fill, read-only / streaming loads, the round-to-nearest fp32 intrinsics (plain IEEE ops; build
// with -ffp-contract=off), fp16 / bf16 storage types.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>

#include "elastic_b200.h"

#define ED_HOST_EMU_ACTIVE 1
#define ED_DEVICE inline
#define ED_TMAP ed::EmuTensorMap
#define ED_DYN_SMEM(name) uint8_t* name = ed::emu_dyn_smem
#define __global__
#define __device__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

extern thread_local dim3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

struct __half { _Float16 v; };
struct __nv_bfloat16 { uint16_t v; };
static inline float __half2float(__half h) { return (float)h.v; }
static inline __half __float2half_rn(float f) { return __half{(_Float16)f}; }
static inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = (uint32_t)h.v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
void __syncthreads();

// (epilogue_half.cuh needs no barrier / TMA: its emulation runs the threads of a CTA one after the other)
// path-coverage counters: 0 staged threads, 1 unstaged threads, 2 RRG down-cell inside the boxes, 3 outside (global load),
// 4 vector view loads, 5 scalar single-view loads, 6 multi-cover walks
extern std::atomic<long long> ed_emu_counters[8];
#define ED_EMU_COUNT(i) (ed_emu_counters[(i)]++)

namespace ed {

extern uint8_t* emu_dyn_smem;

template <typename T> static inline float to_f32(T v);
template <> inline float to_f32<float>(float v) { return v; }
template <> inline float to_f32<__half>(__half v) { return __half2float(v); }
template <> inline float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> static inline float ld_ro(const T* p) { return to_f32<T>(*p); }

// 3-D tensor (d0 fastest) of `es`-byte elements with a fixed box, like a CUtensorMap encoded by encode_tmap_3d
struct EmuTensorMap {
  const uint8_t* base;
  long long d0, d1, d2;
  int b0, b1, b2, es;
};

// mbarrier emulation: bits 0..39 = pending transaction bytes of the current phase, bits 40..63 = number of completed
// phases.  One thread arms and (through the synchronous box copies) completes a phase; the others wait on a parity.
static inline std::atomic<uint64_t>& emu_bar(uint64_t* bar) { return *reinterpret_cast<std::atomic<uint64_t>*>(bar); }
static const uint64_t EMU_PENDING_MASK = (1ull << 40) - 1;
static inline void mbar_init(uint64_t* bar, uint32_t) { emu_bar(bar).store(0); }
static inline void fence_mbar_init() {}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_bar(bar).fetch_add(bytes); }
// true once the phase with parity `parity` has completed (= the barrier's current phase has the other parity)
static inline void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  while (((emu_bar(bar).load() >> 40) & 1u) == parity) std::this_thread::yield();
}
static inline void emu_complete_tx(uint64_t* bar, uint64_t bytes) {
  const uint64_t left = (emu_bar(bar).fetch_sub(bytes) - bytes) & EMU_PENDING_MASK;
  if (left == 0) emu_bar(bar).fetch_add(1ull << 40);   // phase flip
}
template <int BYTES> static inline void cp_async(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, BYTES); }
static inline void cp_async_wait_all() {}
static inline void tma_load_3d(void* smem_dst, const EmuTensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  uint8_t* dst = static_cast<uint8_t*>(smem_dst);
  for (int k = 0; k < m->b2; ++k)
    for (int j = 0; j < m->b1; ++j)
      for (int i = 0; i < m->b0; ++i) {
        const long long x = c0 + i, y = c1 + j, z = c2 + k;
        uint8_t* d = dst + ((size_t)(k * m->b1 + j) * m->b0 + i) * m->es;
        if (x < 0 || x >= m->d0 || y < 0 || y >= m->d1 || z < 0 || z >= m->d2) memset(d, 0, m->es);
        else memcpy(d, m->base + ((z * m->d1 + y) * m->d0 + x) * m->es, m->es);
      }
  emu_complete_tx(bar, (uint64_t)m->b0 * m->b1 * m->b2 * m->es);
}

}  // namespace ed
