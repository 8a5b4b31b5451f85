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
constexpr int kPts = 128;                  // points per CTA (lane = consecutive point)
constexpr int kMaxTile = 256;              // largest encode tile (patch mode: 4 rays x S samples, S <= 64)
// Measured on B200 (profiles/): fetching / reducing adjacent x-pairs of corners as one 16-byte access lowers the L2
// sector count by 25 % but is SLOWER (bwd main grid 0.633 -> 0.654 ms/step, fused proposal bwd 0.475 -> 0.557): the
// atomic path is bound by REDs per lane-address, and unmerged lanes pay a padded 16-byte RED plus the 8-byte one.
constexpr bool kMergeXPairs = false;

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
  bool xmerge;   // floor(x) even and ceil(x) = floor(x)+1: each x-pair of corners is one aligned pair of rows
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
  // hash = x ^ (y*P1) ^ (z*P2): for even floor(x) the ceil corner's row is the floor corner's row with bit 0
  // flipped, i.e. the two rows of an x-pair are adjacent and 2-row aligned
  c.xmerge = ((lo0 & 1) == 0) && (hi0 == lo0 + 1);
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

// The eight corner rows of a cell.  For F = 2 fp32 tables the four x-pairs {3,0} {2,1} {7,4} {6,5} (floor-x corner,
// ceil-x corner) are fetched as ONE 16-byte access each when the pair is adjacent (Cell::xmerge, half of all cells):
// 6 L2 sector accesses per cell on average instead of 8.  Values are identical to eight separate loads.
template <int F, bool HALF>
__device__ __forceinline__ void load_cell(const void* __restrict__ table, const Cell& c, float (&f)[8][F]) {
  if constexpr (kMergeXPairs && F == 2 && !HALF) {
    const float* t = reinterpret_cast<const float*>(table);
    constexpr int pf[4] = {3, 2, 7, 6}, pc[4] = {0, 1, 4, 5};
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(t + (size_t)(c.idx[pf[q]] & ~1u) * 2));
    float2 w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      w[q] = make_float2(0.f, 0.f);
      if (!c.xmerge) w[q] = __ldg(reinterpret_cast<const float2*>(t + (size_t)c.idx[pc[q]] * 2));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool odd = c.idx[pf[q]] & 1u;
      f[pf[q]][0] = odd ? v[q].z : v[q].x;
      f[pf[q]][1] = odd ? v[q].w : v[q].y;
      f[pc[q]][0] = c.xmerge ? (odd ? v[q].x : v[q].z) : w[q].x;
      f[pc[q]][1] = c.xmerge ? (odd ? v[q].y : v[q].w) : w[q].y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) load_row<F, HALF>(table, c.idx[k], f[k]);
  }
}

// scatter-add the eight corner gradients; same pairing: one 16-byte RED per adjacent x-pair
template <int F>
__device__ __forceinline__ void red_cell(float* __restrict__ dtable, const Cell& c, const float (&g)[8][F]) {
  if constexpr (kMergeXPairs && F == 2) {
    constexpr int pf[4] = {3, 2, 7, 6}, pc[4] = {0, 1, 4, 5};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t rf = c.idx[pf[q]];
      const bool odd = rf & 1u;
      const float cx = c.xmerge ? g[pc[q]][0] : 0.f, cy = c.xmerge ? g[pc[q]][1] : 0.f;
      float* a = dtable + (size_t)(rf & ~1u) * 2;
      if (odd) red_add_v4(a, cx, cy, g[pf[q]][0], g[pf[q]][1]);
      else red_add_v4(a, g[pf[q]][0], g[pf[q]][1], cx, cy);
    }
    if (!c.xmerge) {
#pragma unroll
      for (int q = 0; q < 4; ++q) red_add_v2(dtable + (size_t)c.idx[pc[q]] * 2, g[pc[q]][0], g[pc[q]][1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) red_row<F>(dtable, c.idx[k], g[k]);
  }
}

// ---- lane-pair access of x-neighbour corners -----------------------------------------------------------------
// Measured (tools/probe_l2.py, profiles/r02_l2_probe.json): the L2 serves ~240 G sector requests/s to gathers and
// ~160 G/s to REDs whatever their width, and two lanes of ONE instruction that touch the same sector cost one
// request.  The floor-x and ceil-x corners of a cell sit at rows (x ^ h) and ((x+1) ^ h): the same 32-byte sector
// for 3 of 4 cells (same 16 bytes when x is even, same sector when x = 1 mod 4).  So instead of "every lane
// touches corner k of its own cell" (32 sectors per instruction), lanes 2i and 2i+1 touch the floor-x and the
// ceil-x corner of the SAME cell -- first the even lane's cell, then the odd lane's -- and swap the results with
// one shuffle: 1.25 sectors per corner pair instead of 2, for the same number of memory instructions.
// Corner pairs (floor-x corner, ceil-x corner) in the reference's corner numbering:
__device__ constexpr int kPf[4] = {3, 2, 7, 6};
__device__ constexpr int kPc[4] = {0, 1, 4, 5};

struct PairRows {
  uint32_t other[4];  // even lane: the odd lane's floor-x rows; odd lane: the even lane's ceil-x rows
};

// (whole warp, converged)
__device__ __forceinline__ PairRows exchange_rows(const Cell& c, bool odd) {
  PairRows pr;
#pragma unroll
  for (int q = 0; q < 4; ++q) pr.other[q] = __shfl_xor_sync(0xffffffffu, odd ? c.idx[kPf[q]] : c.idx[kPc[q]], 1);
  return pr;
}

// The pairwise fetch in two steps, so that a caller can put independent work between the loads and their first use:
// issue: raw[2q] = this lane's share of the even lane's pair q, raw[2q+1] = of the odd lane's pair q
template <int F, bool HALF>
__device__ __forceinline__ void load_cell_paired_issue(const void* __restrict__ table, const Cell& c,
                                                       const PairRows& pr, bool odd, float (&raw)[8][F]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    load_row<F, HALF>(table, odd ? pr.other[q] : c.idx[kPf[q]], raw[2 * q]);      // even lane's cell: floor-x | ceil-x
    load_row<F, HALF>(table, odd ? c.idx[kPc[q]] : pr.other[q], raw[2 * q + 1]);  // odd lane's cell
  }
}
// finish: swap with the neighbouring lane; f[k] = corner k of this lane's cell (whole warp, converged)
template <int F>
__device__ __forceinline__ void load_cell_paired_finish(const float (&raw)[8][F], bool odd, float (&f)[8][F]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const float recv = __shfl_xor_sync(0xffffffffu, odd ? raw[2 * q][j] : raw[2 * q + 1][j], 1);
      f[kPf[q]][j] = odd ? recv : raw[2 * q][j];
      f[kPc[q]][j] = odd ? raw[2 * q + 1][j] : recv;
    }
}

// the eight corner rows of this lane's cell, fetched pairwise with the neighbouring lane (values identical to
// load_cell); whole warp, converged
template <int F, bool HALF>
__device__ __forceinline__ void load_cell_paired(const void* __restrict__ table, const Cell& c, const PairRows& pr,
                                                 bool odd, float (&f)[8][F]) {
  float raw[8][F];
  load_cell_paired_issue<F, HALF>(table, c, pr, odd, raw);
  load_cell_paired_finish<F>(raw, odd, f);
}

// scatter-add of the eight corner gradients with the same pairing.  `issue`: this lane's cell is to be written
// (false on lanes without a point and on lanes whose gradients were folded into another lane).  Whole warp, converged.
template <int F>
__device__ __forceinline__ void red_cell_paired(float* __restrict__ dtable, const Cell& c, const PairRows& pr, int lane,
                                                bool issue, const float (&g)[8][F]) {
  const bool odd = lane & 1;
  const unsigned issuing = __ballot_sync(0xffffffffu, issue);
  const bool issue_even = (issuing >> (lane & ~1)) & 1u, issue_odd = (issuing >> (lane | 1)) & 1u;
  float recv[4][F];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < F; ++j) recv[q][j] = __shfl_xor_sync(0xffffffffu, odd ? g[kPf[q]][j] : g[kPc[q]][j], 1);
  if (issue_even) {  // the even lane's cell: even lane adds to the floor-x rows, odd lane to the ceil-x rows
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[F];
#pragma unroll
      for (int j = 0; j < F; ++j) v[j] = odd ? recv[q][j] : g[kPf[q]][j];
      red_row<F>(dtable, odd ? pr.other[q] : c.idx[kPf[q]], v);
    }
  }
  if (issue_odd) {  // the odd lane's cell
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[F];
#pragma unroll
      for (int j = 0; j < F; ++j) v[j] = odd ? g[kPc[q]][j] : recv[q][j];
      red_row<F>(dtable, odd ? c.idx[kPc[q]] : pr.other[q], v);
    }
  }
}

}  // namespace tn
