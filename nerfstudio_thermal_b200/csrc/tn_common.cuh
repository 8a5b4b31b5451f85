// Shared helpers for libtn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtn_b200 targets sm_100a (Blackwell B200) only"
#endif

namespace tn {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TN_ECUDA;
  }
  return TN_OK;
}

#define TN_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::tn::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int kNumSMs = 148;  // B200

// Tuning knobs (host): the largest level scale whose table REDs are aggregated over equal cells of a warp.  The
// environment is consulted ONCE, when the library first needs the value (no per-launch getenv, no state that changes
// after that): TN_AGG_ENC / TN_AGG_ENC_PATCH (encode backward, plain / patch tiles), TN_AGG_PROP (proposal backward).
float agg_threshold_enc();
float agg_threshold_enc_patch();
float agg_threshold_prop();
float pair_threshold_enc();  // TN_PAIR_ENC: level scale from which gathers use the lane-pair access (tn_encode_core.cuh)
int pair_reds();             // TN_PAIR_RED: 0 switches the lane-pair REDs off (A/B measurements)
float jac_threshold_enc();   // TN_JAC_ENC: level scale from which a saved Jacobian (jac_out / jac) covers the level

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// vectorised no-return global reduction (sm_90+): one RED.64 instead of two RED.32
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// torch.nan_to_num defaults: nan -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX
__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

}  // namespace tn
