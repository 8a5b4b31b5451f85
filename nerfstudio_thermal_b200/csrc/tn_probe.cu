// L2 gather / reduction throughput probes: the denominators of the hash-grid kernels' L2 rooflines (SURVEY 8d asks
// for a MEASURED L2 peak).  The encode kernels issue per-lane 8-byte row gathers and 8-byte vector REDs at
// pseudo-random rows of a table that lives in L2; these probes issue exactly that access pattern and nothing else,
// so their rate is the ceiling any schedule of the encode kernels can reach on this device.
#include "tn_common.cuh"

namespace tn {

__device__ __forceinline__ uint32_t mix(uint32_t a) {  // cheap integer hash (2 multiplies)
  a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16;
  return a;
}

// mode 0: 8-byte gathers, every lane its own row          mode 1: lanes 2i / 2i+1 read rows r, r^1 (one 16-byte pair)
// mode 2: 8-byte v2 REDs, every lane its own row          mode 3: lanes 2i / 2i+1 add to rows r, r^1
// mode 4: 16-byte v4 REDs (pair of rows per lane)         mode 5: 4-byte scalar REDs
// mode 6: 8-byte gathers, 4 consecutive lanes read rows r..r+3 (one 32-byte sector)
// mode 7 / 8: gathers, lane pairs read rows r, r^3 (same sector, other 16-byte half) / r, r^7 (same line, other sector)
// mode 9 / 10: v2 REDs with the pairings of modes 7 / 8
template <int MODE>
__global__ void __launch_bounds__(256) l2_probe_kernel(float* __restrict__ table, uint32_t row_mask, int iters,
                                                       float* __restrict__ sink) {
  const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
#pragma unroll 1
  for (int it = 0; it < iters; it += 8) {
    uint32_t rows[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t id = gt;
      if (MODE == 1 || MODE == 3 || MODE >= 7) id = gt >> 1;
      if (MODE == 6) id = gt >> 2;
      uint32_t r = mix(id * 0x9e3779b9u + (uint32_t)(it + k) * 0x85ebca6bu) & row_mask;
      if (MODE == 1 || MODE == 3) r = (r & ~1u) | (gt & 1u);
      if (MODE == 6) r = (r & ~3u) | (gt & 3u);
      if (MODE == 7 || MODE == 9) r ^= (gt & 1u) * 3u;
      if (MODE == 8 || MODE == 10) r ^= (gt & 1u) * 7u;
      if (MODE == 4) r &= ~1u;
      rows[k] = r;
    }
    if constexpr (MODE == 0 || MODE == 1 || MODE == 6 || MODE == 7 || MODE == 8) {
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float2*>(table) + rows[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y;
    } else if constexpr (MODE == 2 || MODE == 3 || MODE == 9 || MODE == 10) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red_add_v2(table + (size_t)rows[k] * 2, 1.f, 2.f);
    } else if constexpr (MODE == 4) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red_add_v4(table + (size_t)rows[k] * 2, 1.f, 2.f, 3.f, 4.f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(table + (size_t)rows[k] * 2, 1.f);
    }
  }
  if (acc == 123.456f) sink[0] = acc;  // keeps the loads alive
}

}  // namespace tn

using namespace tn;

extern "C" int tn_l2_probe(int mode, float* table, int log2_rows, int iters, int ctas, float* sink, void* stream) {
  TN_REQUIRE(table && sink, TN_EINVAL, "l2_probe: null pointer");
  TN_REQUIRE(mode >= 0 && mode <= 10, TN_EINVAL, "l2_probe: mode=%d", mode);
  TN_REQUIRE(log2_rows >= 4 && log2_rows <= 30 && iters > 0 && iters % 8 == 0 && ctas > 0, TN_EINVAL,
             "l2_probe: log2_rows=%d iters=%d ctas=%d", log2_rows, iters, ctas);
  TN_REQUIRE(aligned(table, 16), TN_EALIGN, "l2_probe: table must be 16-byte aligned");
  const uint32_t mask = (1u << log2_rows) - 1u;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: l2_probe_kernel<0><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 1: l2_probe_kernel<1><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 2: l2_probe_kernel<2><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 3: l2_probe_kernel<3><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 4: l2_probe_kernel<4><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 5: l2_probe_kernel<5><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 6: l2_probe_kernel<6><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 7: l2_probe_kernel<7><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 8: l2_probe_kernel<8><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    case 9: l2_probe_kernel<9><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
    default: l2_probe_kernel<10><<<ctas, 256, 0, st>>>(table, mask, iters, sink); break;
  }
  return check_launch("l2_probe_kernel");
}
