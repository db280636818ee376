// emu_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile the __device__ phase functions of
// spectrogram-rs_b200/csrc/stft_kernels.cuh for the host, so the CPU-only test suite can run the
// kernel's index math / butterflies / resampling against the oracle without a GPU.
// Never part of libsbx_b200.so; the product has no CPU execution path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __restrict__
#define __launch_bounds__(x)

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }

struct EmuDim { int x; };
static thread_local EmuDim threadIdx = {0};

template <class T> static inline T __ldg(const T *p) { return *p; }
// compiled with -ffp-contract=off: plain operators are the round-to-nearest, unfused operations
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline void __syncthreads() {}
using std::max;
using std::min;
