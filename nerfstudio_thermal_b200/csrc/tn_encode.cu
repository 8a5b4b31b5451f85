// Multiresolution hash-grid encoding, forward and backward (sm_100a).
//
// Semantics follow the reference's torch branch, field_components/encodings.py:401-461: every level is
// hashed, T rows per level, corners are ceil/floor of x*scale, the interpolation weight sits on the CEIL
// corner.  This file is compiled with -fmad=false: every product/sum below is individually rounded exactly
// like the reference's chain of separate torch ops, which makes corner indices AND interpolated features
// bit-identical to the reference for identical inputs.
//
// Mapping: one thread per point, a tile of consecutive points per CTA, levels looped inside the thread (unrolled by
// two so sixteen independent 8-byte gathers are in flight per thread).  The [tile x L*F] output tile is staged in
// shared memory (rows rotated to dodge bank conflicts) and leaves the SM as one contiguous, fully coalesced run.
// Tables live in L2 (64 MB fp32 / 32 MB fp16 at T=2^19 vs 126 MB of L2).
//
// Which point a LANE owns is a permutation inside the tile.  When the caller says the points are ray samples
// (samples_per_ray = S, S % 8 == 0), a tile is 4 rays x S samples (one 2x2 pixel patch of the reference's patch
// sampler, data/pixel_samplers.py:389-438, or four neighbouring pixels of a render chunk) and a warp owns 8
// consecutive samples of each of the 4 rays: those 32 points share far more grid cells than 32 consecutive samples
// of one ray (measured on the bench batch: 0.57 distinct cells per point and level instead of 0.82), so the
// gathers of a warp coalesce into fewer L1 wavefronts and the backward issues fewer REDs.
//
// Backward: the same index math, then one vectorised RED (red.global.add.v2.f32) per corner.  The kernel is bound
// by RED lanes (L2 atomic units), not bytes: lanes of a warp that hold the SAME cell are found with
// match.any on the cell key -- anywhere in the warp, not only neighbouring lanes -- their 8 corner gradients are
// folded into the lowest such lane by a peer-tree of shuffles, and only that lane issues the 8 REDs
// (warp-aggregated atomics).  Levels finer than the aggregation threshold have no duplicates worth the shuffles.
#include "tn_encode_core.cuh"

namespace tn {

// position of (row, level) inside the rotated tile
__device__ __forceinline__ int tile_pos(int row, int l, int L, int F) {
  int r = l + row % L;
  if (r >= L) r -= L;
  return (row * L + r) * F;
}

constexpr int kLevelBatch = 2;  // levels whose 8-corner gathers are issued back to back (16 loads in flight/thread)

// JAC: also store d(feature)/d(x) (times the level scale) of the levels l >= jac_from, PLANAR: jac[l][F*3][N], so
// that every store (and the backward's load) of a warp is whole 32-byte sectors.  The backward then needs no
// second gather of those levels' corner rows for dL/dx: 12*F bytes per level of streaming instead of 8 L2 gathers
// -- worth it exactly on the fine levels, where a gather instruction touches ~20 distinct sectors.
// tile row owned by a thread: identity, or (patch mode) warp w / lane -> ray lane>>3, sample 8w + (lane&7)
__device__ __forceinline__ int tile_row(int tid, int patch_S) {
  return patch_S ? ((tid & 31) >> 3) * patch_S + ((tid >> 5) << 3) + (tid & 7) : tid;
}

template <int F, bool HALF, bool WRITE_IDX, bool JAC = false>
__global__ void __launch_bounds__(kMaxTile, 3) hash_fwd_kernel(const float* __restrict__ x,
                                                               const void* __restrict__ table, LevelScales sc,
                                                               int64_t N, int L, int log2T, int patch_S,
                                                               int pair_from, float* __restrict__ out,
                                                               int32_t* __restrict__ idx_out, float* __restrict__ jac,
                                                               int jac_from) {
  extern __shared__ float4 smem4[];
  __shared__ float s_scale[TN_MAX_LEVELS];
  float* tile = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, kTile = blockDim.x;
  if (tid < TN_MAX_LEVELS) s_scale[tid] = sc.s[tid];
  const int64_t base_pt = (int64_t)blockIdx.x * kTile;
  const int row_t = tile_row(tid, patch_S);
  const int64_t p = base_pt + row_t;
  const bool valid = p < N;
  const uint32_t T = 1u << log2T, mask = T - 1u;
  float x0 = 0.f, x1 = 0.f, x2 = 0.f;
  if (valid) {
    x0 = __ldg(x + 3 * p); x1 = __ldg(x + 3 * p + 1); x2 = __ldg(x + 3 * p + 2);
  }
  const int rowmod = row_t % L;
  __syncthreads();
#pragma unroll 1
  for (int l0 = 0; l0 < L; l0 += kLevelBatch) {
    // phase 1: all corner rows of the batch, phase 2: all gathers, phase 3: blends.  Keeping the phases apart
    // is what puts 8*kLevelBatch independent loads in flight per thread (the gathers are latency-bound).
    Cell c[kLevelBatch];
    float f[kLevelBatch][8][F];
#pragma unroll
    for (int b = 0; b < kLevelBatch; ++b) {
      const int l = min(l0 + b, L - 1);
      c[b] = locate(x0, x1, x2, s_scale[l], mask, (uint32_t)l * T);
    }
    if (l0 >= pair_from) {  // warp-uniform: fine levels fetch x-neighbour corners by lane pairs (tn_encode_core.cuh)
      const bool odd = tid & 1;
      PairRows pr[kLevelBatch];
#pragma unroll
      for (int b = 0; b < kLevelBatch; ++b) pr[b] = exchange_rows(c[b], odd);
#pragma unroll
      for (int b = 0; b < kLevelBatch; ++b) load_cell_paired<F, HALF>(table, c[b], pr[b], odd, f[b]);
    } else {
#pragma unroll
      for (int b = 0; b < kLevelBatch; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) load_row<F, HALF>(table, c[b].idx[k], f[b][k]);
    }
#pragma unroll
    for (int b = 0; b < kLevelBatch; ++b) {
      const int l = l0 + b;
      if (l < L) {
        if constexpr (WRITE_IDX) {
          if (valid) {
#pragma unroll
            for (int k = 0; k < 8; ++k) idx_out[(p * L + l) * 8 + k] = (int32_t)c[b].idx[k];
          }
        }
        const float ox = c[b].ox, oy = c[b].oy, oz = c[b].oz;
        const float mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
        int r = l + rowmod;
        if (r >= L) r -= L;
        float* dst = tile + (row_t * L + r) * F;
#pragma unroll
        for (int j = 0; j < F; ++j) {
          // encodings.py:449-459, same association
          const float f03 = f[b][0][j] * ox + f[b][3][j] * mx;
          const float f12 = f[b][1][j] * ox + f[b][2][j] * mx;
          const float f56 = f[b][5][j] * ox + f[b][6][j] * mx;
          const float f47 = f[b][4][j] * ox + f[b][7][j] * mx;
          const float f0312 = f03 * oy + f12 * my;
          const float f4756 = f47 * oy + f56 * my;
          dst[j] = f0312 * oz + f4756 * mz;
          if constexpr (JAC) {
            if (valid && l >= jac_from) {
              const float sl = s_scale[l];
              const float e03 = f[b][0][j] - f[b][3][j], e12 = f[b][1][j] - f[b][2][j];
              const float e56 = f[b][5][j] - f[b][6][j], e47 = f[b][4][j] - f[b][7][j];
              float* jd = jac + ((size_t)l * (F * 3) + j * 3) * N + p;
              jd[0] = ((e03 * oy + e12 * my) * oz + (e47 * oy + e56 * my) * mz) * sl;
              jd[N] = ((f03 - f12) * oz + (f47 - f56) * mz) * sl;
              jd[2 * N] = (f0312 - f4756) * sl;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // contiguous, coalesced write-out of the tile (rows base_pt .. base_pt+rows-1 are adjacent in `out`)
  const int rows = (int)min((int64_t)kTile, N - base_pt);
  float* gout = out + base_pt * (int64_t)(L * F);
  for (int i = tid; i < rows * L; i += kTile) {
    const int row = i / L, l = i - row * L;
    const float* src = tile + tile_pos(row, l, L, F);
    if constexpr (F == 2) {
      *reinterpret_cast<float2*>(gout + (size_t)i * 2) = *reinterpret_cast<const float2*>(src);
    } else if constexpr (F == 4 || F == 8) {
#pragma unroll
      for (int j = 0; j < F; j += 4)
        *reinterpret_cast<float4*>(gout + (size_t)i * F + j) = *reinterpret_cast<const float4*>(src + j);
    } else {
      gout[i] = src[0];
    }
  }
}

// Fold the corner gradients of all lanes that hold the same cell into the lowest such lane (peer tree: log2 of the
// largest group rounds, each a shuffle per value).  Returns true on the lane that must issue the REDs.
template <int F>
__device__ __forceinline__ bool aggregate_equal_cells(uint64_t key, int lane, float (&gc)[8][F]) {
  const unsigned all = 0xffffffffu;
  const unsigned peers_all = __match_any_sync(all, key);
  int rel = __popc(peers_all & ((1u << lane) - 1u));  // peers in lower lanes
  const bool leader = rel == 0;
  unsigned peers = peers_all & (0xfffffffeu << lane);  // peers in higher lanes
  while (__any_sync(all, peers != 0u)) {
    const int next = __ffs(peers);  // 1-based lane of the nearest remaining higher peer, 0: none
    const int src = next ? next - 1 : lane;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < F; ++j) {
        const float o = __shfl_sync(all, gc[k][j], src);
        if (next) gc[k][j] += o;
      }
    // lanes at odd positions among their remaining peers have just been absorbed by their predecessor
    peers &= ~__ballot_sync(all, rel & 1);
    rel >>= 1;
  }
  return leader;
}

template <int F, bool HALF, bool NEED_DX, bool JAC = false>
__global__ void __launch_bounds__(kMaxTile, 3) hash_bwd_kernel(const float* __restrict__ x,
                                                               const void* __restrict__ table, LevelScales sc,
                                                               const float* __restrict__ dy, int64_t N, int L,
                                                               int log2T, int n_agg, int patch_S, int pair_from,
                                                               int pair_red, float* __restrict__ dtable,
                                                               float* __restrict__ dx,
                                                               const float* __restrict__ jac, int jac_from) {
  extern __shared__ float4 smem4[];
  float* tile = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, lane = tid & 31, kTile = blockDim.x;
  const int64_t base_pt = (int64_t)blockIdx.x * kTile;
  const int row_t = tile_row(tid, patch_S);
  const int64_t p = base_pt + row_t;
  const bool valid = p < N;
  const uint32_t T = 1u << log2T, mask = T - 1u;
  // coalesced load of the dy tile into the rotated layout
  const int rows = (int)min((int64_t)kTile, N - base_pt);
  const float* gdy = dy + base_pt * (int64_t)(L * F);
  for (int i = tid; i < rows * L; i += kTile) {
    const int row = i / L, l = i - row * L;
    float* dst = tile + tile_pos(row, l, L, F);
    if constexpr (F == 2) {
      *reinterpret_cast<float2*>(dst) = __ldg(reinterpret_cast<const float2*>(gdy + (size_t)i * 2));
    } else if constexpr (F == 4 || F == 8) {
#pragma unroll
      for (int j = 0; j < F; j += 4)
        *reinterpret_cast<float4*>(dst + j) = __ldg(reinterpret_cast<const float4*>(gdy + (size_t)i * F + j));
    } else {
      dst[0] = __ldg(gdy + i);
    }
  }
  float x0 = 0.f, x1 = 0.f, x2 = 0.f;
  if (valid) {
    x0 = __ldg(x + 3 * p); x1 = __ldg(x + 3 * p + 1); x2 = __ldg(x + 3 * p + 2);
  }
  __syncthreads();
  const int rowmod = row_t % L;
  float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
  // dL/dx of the levels whose Jacobian the forward kept: a streaming pre-pass, three levels (9*F independent,
  // coalesced loads) in flight per thread -- the scatter loop below then has no load to wait for on those levels.
  if constexpr (NEED_DX && JAC) {
    constexpr int JB = 3;
#pragma unroll 1
    for (int l0 = jac_from; l0 < L; l0 += JB) {
      float jv[JB][F * 3];
#pragma unroll
      for (int b = 0; b < JB; ++b) {
        const int l = min(l0 + b, L - 1);
        const float* js = jac + (size_t)l * (F * 3) * N + (valid ? p : 0);
#pragma unroll
        for (int q = 0; q < F * 3; ++q) jv[b][q] = __ldg(js + (size_t)q * N);
      }
#pragma unroll
      for (int b = 0; b < JB; ++b) {
        const int l = l0 + b;
        if (l < L && valid) {
          int r = l + rowmod;
          if (r >= L) r -= L;
#pragma unroll
          for (int j = 0; j < F; ++j) {
            const float gj = tile[(row_t * L + r) * F + j];
            dx0 += gj * jv[b][j * 3]; dx1 += gj * jv[b][j * 3 + 1]; dx2 += gj * jv[b][j * 3 + 2];
          }
        }
      }
    }
  }
  // One level per iteration.  (Prefetching the next level's corner rows one iteration ahead -- 16 gathers in flight --
  // was measured and lost, DESIGN.md "experiments that lost".)
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const float scale = sc.s[l];
    const Cell c = locate(x0, x1, x2, scale, mask, (uint32_t)l * T);
    int r = l + rowmod;
    if (r >= L) r -= L;
    float g[F];
#pragma unroll
    for (int j = 0; j < F; ++j) g[j] = valid ? tile[(row_t * L + r) * F + j] : 0.f;
    const float mx = 1.f - c.ox, my = 1.f - c.oy, mz = 1.f - c.oz;
    float gc[8][F];
    float f[8][F];
    const bool use_jac = JAC && l >= jac_from;  // warp-uniform: this level's dL/dx came from the forward's Jacobian
    const bool odd = lane & 1;
    PairRows pr;
    if (pair_red || l >= pair_from) pr = exchange_rows(c, odd);  // (warp-uniform conditions)
    // The gathers (for dL/dx) are issued first and consumed LAST: the gradient scatter below -- corner weights,
    // warp aggregation, REDs -- does not depend on them and runs while they are in flight.
    if (NEED_DX && !use_jac) {
      if (l >= pair_from) {
        load_cell_paired_issue<F, HALF>(table, c, pr, odd, f);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) load_row<F, HALF>(table, c.idx[k], f[k]);
      }
    }
    float gq[F][4];  // g03, g12, g56, g47 per feature (reused by dL/dx)
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const float g0312 = g[j] * c.oz, g4756 = g[j] * mz;
      const float g03 = g0312 * c.oy, g12 = g0312 * my;
      const float g47 = g4756 * c.oy, g56 = g4756 * my;
      gq[j][0] = g03; gq[j][1] = g12; gq[j][2] = g56; gq[j][3] = g47;
      gc[0][j] = g03 * c.ox; gc[3][j] = g03 * mx;
      gc[1][j] = g12 * c.ox; gc[2][j] = g12 * mx;
      gc[5][j] = g56 * c.ox; gc[6][j] = g56 * mx;
      gc[4][j] = g47 * c.ox; gc[7][j] = g47 * mx;
    }
    bool issue = valid;
    if (l < n_agg) {  // warp-uniform branch; lanes without a point get a key no other lane has
      const uint64_t key = valid ? c.key : ((1ull << 63) | (uint64_t)lane);
      issue = aggregate_equal_cells<F>(key, lane, gc) && valid;
    }
    if (pair_red) {
      red_cell_paired<F>(dtable, c, pr, lane, issue, gc);
    } else if (issue) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red_row<F>(dtable, c.idx[k], gc[k]);
    }
    // ---- dL/dx of this level
    if (NEED_DX && use_jac) {
      // (pre-pass above)
    } else if constexpr (NEED_DX) {
      if (l >= pair_from) {
        float ff[8][F];
        load_cell_paired_finish<F>(f, odd, ff);
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int j = 0; j < F; ++j) f[k][j] = ff[k][j];
      }
      float dox = 0.f, doy = 0.f, doz = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
        const float f03 = f[0][j] * c.ox + f[3][j] * mx;
        const float f12 = f[1][j] * c.ox + f[2][j] * mx;
        const float f56 = f[5][j] * c.ox + f[6][j] * mx;
        const float f47 = f[4][j] * c.ox + f[7][j] * mx;
        const float f0312 = f03 * c.oy + f12 * my;
        const float f4756 = f47 * c.oy + f56 * my;
        dox += gq[j][0] * (f[0][j] - f[3][j]) + gq[j][1] * (f[1][j] - f[2][j]) + gq[j][2] * (f[5][j] - f[6][j]) +
               gq[j][3] * (f[4][j] - f[7][j]);
        doy += (g[j] * c.oz) * (f03 - f12) + (g[j] * mz) * (f47 - f56);
        doz += g[j] * (f0312 - f4756);
      }
      dx0 += dox * scale; dx1 += doy * scale; dx2 += doz * scale;
    }
  }
  if constexpr (NEED_DX) {
    if (valid) {
      dx[3 * p] = dx0; dx[3 * p + 1] = dx1; dx[3 * p + 2] = dx2;
    }
  }
}

static int check_common(const float* x, const void* table, const float* scales_host, int64_t N, int L, int F,
                        int log2_T, int table_dtype) {
  TN_REQUIRE((x || N == 0) && table && scales_host, TN_EINVAL, "hash_encode: null pointer");
  TN_REQUIRE(N >= 0 && N < (int64_t(1) << 40), TN_EINVAL, "hash_encode: bad N=%lld", (long long)N);
  TN_REQUIRE(L >= 1 && L <= TN_MAX_LEVELS, TN_EINVAL, "hash_encode: L=%d out of range [1,%d]", L, TN_MAX_LEVELS);
  TN_REQUIRE(F == 1 || F == 2 || F == 4 || F == 8, TN_EINVAL, "hash_encode: F=%d not in {1,2,4,8}", F);
  TN_REQUIRE(log2_T >= 1 && log2_T <= 26 && ((int64_t)L << log2_T) < (int64_t(1) << 32), TN_EINVAL,
             "hash_encode: log2_T=%d unsupported", log2_T);
  TN_REQUIRE(table_dtype == 0 || table_dtype == 1, TN_EINVAL, "hash_encode: table_dtype=%d", table_dtype);
  TN_REQUIRE(aligned(table, 16) && aligned(x, 4), TN_EALIGN, "hash_encode: table must be 16-byte aligned");
  return TN_OK;
}

// Tile geometry: patch mode (4 rays x S samples per CTA, see the file header) when the points are whole 4-ray
// groups of S samples and the tile fits a CTA; otherwise 128 consecutive points, lane = point.
struct TileShape {
  int threads, patch_S;
};
static TileShape tile_shape(int64_t N, int samples_per_ray) {
  const int S = samples_per_ray;
  if (S >= 8 && S % 8 == 0 && 4 * S <= kMaxTile && N % (4 * (int64_t)S) == 0) return {4 * S, S};
  return {kPts, 0};
}

// first level (a multiple of the forward's level batch) whose gathers use the lane-pair access
static int pair_from_level(const LevelScales& sc, int L) {
  const float thr = pair_threshold_enc();
  int l = 0;
  while (l < L && sc.s[l] < thr) ++l;
  return (l + kLevelBatch - 1) / kLevelBatch * kLevelBatch;
}

// first level whose Jacobian the forward keeps for the backward (levels below it are gathered again: coarse cells
// are shared by many lanes, their gathers are cheap)
static int jac_from_level(const LevelScales& sc, int L) {
  const float thr = jac_threshold_enc();
  int l = 0;
  while (l < L && sc.s[l] < thr) ++l;
  return l;
}

template <int F>
static int launch_fwd(const float* x, const void* table, int table_dtype, const LevelScales& sc, int64_t N, int L,
                      int log2_T, int samples_per_ray, float* out, int32_t* idx_out, float* jac, cudaStream_t st) {
  const TileShape ts = tile_shape(N, samples_per_ray);
  const int pair_from = pair_from_level(sc, L), jac_from = jac_from_level(sc, L);
  const unsigned grid = (unsigned)((N + ts.threads - 1) / ts.threads);
  const size_t smem = (size_t)ts.threads * L * F * sizeof(float);
#define TN_FWD(H, W)                                                                                         \
  do {                                                                                                       \
    auto k = hash_fwd_kernel<F, H, W>;                                                                       \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    k<<<grid, ts.threads, smem, st>>>(x, table, sc, N, L, log2_T, ts.patch_S, pair_from, out, idx_out, jac,   \
                                      jac_from);                                                            \
  } while (0)
  if (jac) {  // (no index dump on this path: checked by the caller)
    auto k = table_dtype == 0 ? hash_fwd_kernel<F, false, false, true> : hash_fwd_kernel<F, true, false, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, ts.threads, smem, st>>>(x, table, sc, N, L, log2_T, ts.patch_S, pair_from, out, nullptr, jac, jac_from);
  } else if (table_dtype == 0) {
    if (idx_out) TN_FWD(false, true); else TN_FWD(false, false);
  } else {
    if (idx_out) TN_FWD(true, true); else TN_FWD(true, false);
  }
#undef TN_FWD
  return check_launch("hash_fwd_kernel");
}

template <int F>
static int launch_bwd(const float* x, const void* table, int table_dtype, const LevelScales& sc, const float* dy,
                      int64_t N, int L, int log2_T, int n_agg, int samples_per_ray, float* dtable, float* dx,
                      const float* jac, cudaStream_t st) {
  const TileShape ts = tile_shape(N, samples_per_ray);
  const int pair_from = pair_from_level(sc, L), pair_red = pair_reds(), jac_from = jac_from_level(sc, L);
  const unsigned grid = (unsigned)((N + ts.threads - 1) / ts.threads);
  const size_t smem = (size_t)ts.threads * L * F * sizeof(float);
#define TN_BWD(H, D)                                                                                         \
  do {                                                                                                       \
    auto k = hash_bwd_kernel<F, H, D>;                                                                       \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    k<<<grid, ts.threads, smem, st>>>(x, table, sc, dy, N, L, log2_T, n_agg, ts.patch_S, pair_from, pair_red, \
                                      dtable, dx, jac, jac_from);                                            \
  } while (0)
  if (jac && dx) {
    auto k = table_dtype == 0 ? hash_bwd_kernel<F, false, true, true> : hash_bwd_kernel<F, true, true, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, ts.threads, smem, st>>>(x, table, sc, dy, N, L, log2_T, n_agg, ts.patch_S, pair_from, pair_red, dtable,
                                      dx, jac, jac_from);
  } else if (table_dtype == 0) {
    if (dx) TN_BWD(false, true); else TN_BWD(false, false);
  } else {
    if (dx) TN_BWD(true, true); else TN_BWD(true, false);
  }
#undef TN_BWD
  return check_launch("hash_bwd_kernel");
}

}  // namespace tn

using namespace tn;

extern "C" int tn_hash_encode_fwd(const float* x, const void* table, int table_dtype, const float* scales_host,
                                  int64_t N, int L, int F, int log2_T, int samples_per_ray, float* out,
                                  int32_t* idx_out, float* jac_out, void* stream) {
  int rc = check_common(x, table, scales_host, N, L, F, log2_T, table_dtype);
  if (rc) return rc;
  TN_REQUIRE(out || N == 0, TN_EINVAL, "hash_encode_fwd: out is null");
  TN_REQUIRE(aligned(out, 16), TN_EALIGN, "hash_encode_fwd: out must be 16-byte aligned");
  TN_REQUIRE((size_t)kMaxTile * L * F * 4 <= 200 * 1024, TN_EINVAL, "hash_encode_fwd: L*F=%d too large", L * F);
  TN_REQUIRE(samples_per_ray >= 0, TN_EINVAL, "hash_encode_fwd: samples_per_ray=%d", samples_per_ray);
  TN_REQUIRE(!(jac_out && idx_out), TN_EINVAL, "hash_encode_fwd: idx_out and jac_out are exclusive");
  if (N == 0) return TN_OK;
  LevelScales sc;
  for (int l = 0; l < TN_MAX_LEVELS; ++l) sc.s[l] = l < L ? scales_host[l] : 0.f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (F) {
    case 1: return launch_fwd<1>(x, table, table_dtype, sc, N, L, log2_T, samples_per_ray, out, idx_out, jac_out, st);
    case 2: return launch_fwd<2>(x, table, table_dtype, sc, N, L, log2_T, samples_per_ray, out, idx_out, jac_out, st);
    case 4: return launch_fwd<4>(x, table, table_dtype, sc, N, L, log2_T, samples_per_ray, out, idx_out, jac_out, st);
    default: return launch_fwd<8>(x, table, table_dtype, sc, N, L, log2_T, samples_per_ray, out, idx_out, jac_out, st);
  }
}

extern "C" int tn_hash_encode_bwd(const float* x, const void* table, int table_dtype, const float* scales_host,
                                  const float* dy, int64_t N, int L, int F, int log2_T, int samples_per_ray,
                                  float* dtable, float* dx, const float* jac, void* stream) {
  int rc = check_common(x, table, scales_host, N, L, F, log2_T, table_dtype);
  if (rc) return rc;
  TN_REQUIRE((dy || N == 0) && dtable, TN_EINVAL, "hash_encode_bwd: null pointer");
  TN_REQUIRE(aligned(dy, 16) && aligned(dtable, 16), TN_EALIGN, "hash_encode_bwd: dy/dtable must be 16-byte aligned");
  TN_REQUIRE((size_t)kMaxTile * L * F * 4 <= 200 * 1024, TN_EINVAL, "hash_encode_bwd: L*F=%d too large", L * F);
  TN_REQUIRE(samples_per_ray >= 0, TN_EINVAL, "hash_encode_bwd: samples_per_ray=%d", samples_per_ray);
  if (N == 0) return TN_OK;
  LevelScales sc;
  int n_agg = 0;
  // levels coarse enough for the lanes of a warp to share cells: aggregate there.  Measured on the bench batch
  // (distinct cells per point in a warp): 0.10 at scale 16 ... 0.63 at 212, 0.74 at 294, 0.91 at 561 (patch tiles);
  // lane = consecutive sample of one ray: 0.29 at 16 ... 0.82 at 80, 1.0 from 212 on.
  const TileShape ts = tile_shape(N, samples_per_ray);
  const float thr = ts.patch_S ? agg_threshold_enc_patch() : agg_threshold_enc();
  for (int l = 0; l < TN_MAX_LEVELS; ++l) {
    sc.s[l] = l < L ? scales_host[l] : 0.f;
    if (l < L && l == n_agg && scales_host[l] <= thr) ++n_agg;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (F) {
    case 1: return launch_bwd<1>(x, table, table_dtype, sc, dy, N, L, log2_T, n_agg, samples_per_ray, dtable, dx, jac, st);
    case 2: return launch_bwd<2>(x, table, table_dtype, sc, dy, N, L, log2_T, n_agg, samples_per_ray, dtable, dx, jac, st);
    case 4: return launch_bwd<4>(x, table, table_dtype, sc, dy, N, L, log2_T, n_agg, samples_per_ray, dtable, dx, jac, st);
    default: return launch_bwd<8>(x, table, table_dtype, sc, dy, N, L, log2_T, n_agg, samples_per_ray, dtable, dx, jac, st);
  }
}
