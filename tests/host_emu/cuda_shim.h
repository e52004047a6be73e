// cuda_shim.h -- TEST INFRASTRUCTURE: just enough of the CUDA device vocabulary for g++ to compile the
// lane kernels of redux_b200/csrc/redux_lane_codec.cuh as ordinary functions.  The lane mapping has no
// barriers and no warp collectives (every lane is an independent stream), so running the "threads" one
// after another on the CPU executes exactly the arithmetic the GPU executes.  This lets the container
// without a GPU check every change to the kernels against the oracle before GPU minutes are spent.
// It is never linked into libredux_b200.so: the product has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define RDX_HOST_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__
#define __restrict__ __restrict

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }

struct EmuIdx { unsigned x, y, z; };
extern thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;

template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() {}      // the emulation runs the threads of a CTA one after another

static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(src >> (8 * (s & 7))) & 0xFF;
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
// funnel shifts, wrap mode (shift & 31), as the hardware SHF.W
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
// clamp mode (shift clamped to 32), SHF.L.CLAMP / SHF.R.CLAMP
static inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t sh)
{
    if (sh >= 32) return lo;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh)
{
    if (sh >= 32) return hi;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __ull2float_rn(unsigned long long x) { return (float)x; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
