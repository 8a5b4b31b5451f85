// Core device helpers of the hash-grid kernels (corner location + hashing, row loads, vector REDs), shared by
// tn_encode.cu and the fused field kernels.  Including translation units are compiled with -fmad=false.
#pragma once
#include "tn_common.cuh"

namespace tn {

struct LevelScales {
  float s[TN_MAX_LEVELS];
};

constexpr uint32_t kPrimeY = 2654435761u;  // encodings.py:413
constexpr uint32_t kPrimeZ = 805459861u;
constexpr int kPts = 128;                  // points per CTA

template <int F, bool HALF>
__device__ __forceinline__ void load_row(const void* __restrict__ table, uint32_t row, float (&f)[F]) {
  if constexpr (!HALF) {
    const float* t = reinterpret_cast<const float*>(table) + (size_t)row * F;
    if constexpr (F == 1) {
      f[0] = __ldg(t);
    } else if constexpr (F == 2) {
      float2 v = __ldg(reinterpret_cast<const float2*>(t));
      f[0] = v.x; f[1] = v.y;
    } else {
#pragma unroll
      for (int i = 0; i < F; i += 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(t + i));
        f[i] = v.x; f[i + 1] = v.y; f[i + 2] = v.z; f[i + 3] = v.w;
      }
    }
  } else {
    const __half* t = reinterpret_cast<const __half*>(table) + (size_t)row * F;
    if constexpr (F == 1) {
      f[0] = __half2float(__ldg(t));
    } else {
#pragma unroll
      for (int i = 0; i < F; i += 2) {
        float2 v = __half22float2(__ldg(reinterpret_cast<const __half2*>(t + i)));
        f[i] = v.x; f[i + 1] = v.y;
      }
    }
  }
}

// Corner rows in the reference order hashed_0..hashed_7 (encodings.py:431-438) and the three offsets.
struct Cell {
  uint32_t idx[8];
  float ox, oy, oz;
  uint64_t key;  // identifies (floor, ceil) triple: equal keys <=> identical eight rows
};

__device__ __forceinline__ Cell locate(float x0, float x1, float x2, float scale, uint32_t mask, uint32_t base) {
  Cell c;
  const float s0 = x0 * scale, s1 = x1 * scale, s2 = x2 * scale;  // encodings.py:425
  const int lo0 = (int)floorf(s0), lo1 = (int)floorf(s1), lo2 = (int)floorf(s2);  // :427
  const int hi0 = (int)ceilf(s0), hi1 = (int)ceilf(s1), hi2 = (int)ceilf(s2);     // :426
  c.ox = s0 - (float)lo0;  // :429
  c.oy = s1 - (float)lo1;
  c.oz = s2 - (float)lo2;
  // int32 * int64 primes, xor, mod 2^k (:413-417)  ==  uint32 wrap-around arithmetic, masked
  const uint32_t xc = (uint32_t)hi0, xf = (uint32_t)lo0;
  const uint32_t yc = (uint32_t)hi1 * kPrimeY, yf = (uint32_t)lo1 * kPrimeY;
  const uint32_t zc = (uint32_t)hi2 * kPrimeZ, zf = (uint32_t)lo2 * kPrimeZ;
  c.idx[0] = ((xc ^ yc ^ zc) & mask) + base;
  c.idx[1] = ((xc ^ yf ^ zc) & mask) + base;
  c.idx[2] = ((xf ^ yf ^ zc) & mask) + base;
  c.idx[3] = ((xf ^ yc ^ zc) & mask) + base;
  c.idx[4] = ((xc ^ yc ^ zf) & mask) + base;
  c.idx[5] = ((xc ^ yf ^ zf) & mask) + base;
  c.idx[6] = ((xf ^ yf ^ zf) & mask) + base;
  c.idx[7] = ((xf ^ yc ^ zf) & mask) + base;
  c.key = (uint64_t)(uint32_t)(lo0 & 0xFFFFF) | ((uint64_t)(uint32_t)(lo1 & 0xFFFFF) << 20) |
          ((uint64_t)(uint32_t)(lo2 & 0xFFFFF) << 40) | ((uint64_t)(hi0 != lo0) << 60) |
          ((uint64_t)(hi1 != lo1) << 61) | ((uint64_t)(hi2 != lo2) << 62);
  return c;
}

template <int F>
__device__ __forceinline__ void red_row(float* __restrict__ dtable, uint32_t row, const float (&g)[F]) {
  float* a = dtable + (size_t)row * F;
  if constexpr (F == 1) {
    atomicAdd(a, g[0]);
  } else if constexpr (F == 2) {
    red_add_v2(a, g[0], g[1]);
  } else {
#pragma unroll
    for (int i = 0; i < F; i += 4) red_add_v4(a + i, g[i], g[i + 1], g[i + 2], g[i + 3]);
  }
}

}  // namespace tn
