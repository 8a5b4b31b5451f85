// Fused proposal density field: ray samples -> density in ONE kernel (and one kernel for the whole backward).
//
//   positions (cameras/rays.py:55) -> L-inf contraction, (p+2)/4, selector (fields/density_fields.py:96-103)
//   -> hash grid, L<=8 levels x 2 features (field_components/encodings.py:401-461)
//   -> Linear(2L,16) + ReLU -> Linear(16,1)   (fields/density_fields.py:81-92, field_components/mlp.py:159-178)
//   -> average_init_density * trunc_exp(.) * selector   (fields/density_fields.py:116-117)
//
// The proposal networks see 704 of the 896 points of every ray but do almost no arithmetic (176 MACs/point),
// so the unfused chain is pure launch + HBM round-trip overhead (positions, 40-byte feature rows, raw density).
// Here a thread owns a point end to end: features, hidden units and the raw density never leave registers.
// The MLP is SIMT on purpose: a 10x16 layer is far below one tcgen05 tile, and the kernel is bound by the
// 40 table gathers (forward) / 40 vector REDs (backward) per point.
//
// Backward per point: recompute features and hidden units, back-propagate to the features, scatter into the
// table gradient (run-aggregated REDs as in tn_encode.cu), optionally dL/dx -> contraction backward -> per-ray
// dL/d(origin, direction).  Weight gradients: each 128-point tile stages (dh, features, g*h) in shared memory
// and 128 threads reduce the 193 entries of (dW1, db1, dW2, db2) into register accumulators that live across the
// tiles of a persistent CTA; one atomicAdd per entry per CTA at the end.
// Compiled with -fmad=false (index math must round like the reference); the MLP uses explicit fmaf().
#include "tn_encode_core.cuh"
#include "tn_geometry.cuh"

namespace tn {

constexpr int PH = 16;       // hidden width of the proposal MLP
constexpr int PMAXL = 8;     // levels supported by the fused kernel (F = 2 -> at most 16 input features)

struct PropWeights {  // shared-memory image
  float w1[PH][2 * PMAXL];   // [hidden][input]
  float b1[PH];
  float w2[PH];
  float b2;
};

__device__ __forceinline__ void load_prop_weights(PropWeights& sw, const float* __restrict__ w1,
                                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                                  const float* __restrict__ b2, int in_dim) {
  for (int i = threadIdx.x; i < PH * 2 * PMAXL; i += blockDim.x) {
    const int j = i / (2 * PMAXL), k = i % (2 * PMAXL);
    sw.w1[j][k] = k < in_dim ? __ldg(w1 + j * in_dim + k) : 0.f;
  }
  if (threadIdx.x < PH) {
    sw.b1[threadIdx.x] = __ldg(b1 + threadIdx.x);
    sw.w2[threadIdx.x] = __ldg(w2 + threadIdx.x);
  }
  if (threadIdx.x == 0) sw.b2 = __ldg(b2);
}

// features of one point: all levels, 8 gathers each, two levels' gathers in flight together.
// JAC additionally returns d(feat[k])/d(x) (already times the level scale), so the backward kernel never gathers
// the corner rows a second time.
template <int L, bool JAC>
__device__ __forceinline__ void prop_features(const float* __restrict__ table, const float* s_scale, uint32_t mask,
                                              uint32_t T, float x0, float x1, float x2, float (&feat)[2 * L],
                                              float (&jac)[JAC ? 2 * L : 1][3]) {
#pragma unroll
  for (int l0 = 0; l0 < L; l0 += 2) {
    Cell c[2];
    float f[2][8][2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int l = (l0 + b < L) ? l0 + b : L - 1;
      c[b] = locate(x0, x1, x2, s_scale[l], mask, (uint32_t)l * T);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) load_cell<2, false>(table, c[b], f[b]);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (l0 + b < L) {
        const float ox = c[b].ox, oy = c[b].oy, oz = c[b].oz;
        const float mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float f03 = f[b][0][j] * ox + f[b][3][j] * mx;
          const float f12 = f[b][1][j] * ox + f[b][2][j] * mx;
          const float f56 = f[b][5][j] * ox + f[b][6][j] * mx;
          const float f47 = f[b][4][j] * ox + f[b][7][j] * mx;
          const float f0312 = f03 * oy + f12 * my;
          const float f4756 = f47 * oy + f56 * my;
          feat[(l0 + b) * 2 + j] = f0312 * oz + f4756 * mz;
          if constexpr (JAC) {
            const float sl = s_scale[l0 + b];
            const float e03 = f[b][0][j] - f[b][3][j], e12 = f[b][1][j] - f[b][2][j];
            const float e56 = f[b][5][j] - f[b][6][j], e47 = f[b][4][j] - f[b][7][j];
            jac[(l0 + b) * 2 + j][0] = ((e03 * oy + e12 * my) * oz + (e47 * oy + e56 * my) * mz) * sl;
            jac[(l0 + b) * 2 + j][1] = ((f03 - f12) * oz + (f47 - f56) * mz) * sl;
            jac[(l0 + b) * 2 + j][2] = (f0312 - f4756) * sl;
          }
        }
      }
    }
  }
}

template <int L>
__device__ __forceinline__ float prop_mlp(const PropWeights& sw, const float (&feat)[2 * L], float (&h)[PH]) {
  float raw = sw.b2;
#pragma unroll
  for (int j = 0; j < PH; ++j) {
    float a = sw.b1[j];
#pragma unroll
    for (int k = 0; k < 2 * L; ++k) a = fmaf(sw.w1[j][k], feat[k], a);
    h[j] = fmaxf(a, 0.f);
    raw = fmaf(sw.w2[j], h[j], raw);
  }
  return raw;
}

template <int L>
__global__ void __launch_bounds__(kPts, 4) prop_fwd_kernel(const float* __restrict__ origins,
                                                           const float* __restrict__ dirs,
                                                           const float* __restrict__ ebins, const float* __restrict__ table,
                                                           LevelScales sc, int log2T, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, float scale, int64_t R, int S,
                                                           float* __restrict__ density) {
  __shared__ PropWeights sw;
  __shared__ float s_scale[TN_MAX_LEVELS];
  if (threadIdx.x < TN_MAX_LEVELS) s_scale[threadIdx.x] = sc.s[threadIdx.x];
  load_prop_weights(sw, w1, b1, w2, b2, 2 * L);
  __syncthreads();
  const uint32_t T = 1u << log2T, mask = T - 1u;
  const int64_t N = R * S;
  for (int64_t p = (int64_t)blockIdx.x * kPts + threadIdx.x; p < N; p += (int64_t)gridDim.x * kPts) {
    const int64_t r = p / S;
    const int s = (int)(p - r * S);
    const float st = __ldg(ebins + r * (S + 1) + s), en = __ldg(ebins + r * (S + 1) + s + 1);
    float x0 = sample_pos(__ldg(origins + 3 * r), __ldg(dirs + 3 * r), st, en);
    float x1 = sample_pos(__ldg(origins + 3 * r + 1), __ldg(dirs + 3 * r + 1), st, en);
    float x2 = sample_pos(__ldg(origins + 3 * r + 2), __ldg(dirs + 3 * r + 2), st, en);
    const float sel = contract_normalise(x0, x1, x2);
    float feat[2 * L], h[PH], nojac[1][3];
    prop_features<L, false>(table, s_scale, mask, T, x0, x1, x2, feat, nojac);
    const float raw = prop_mlp<L>(sw, feat, h);
    density[p] = scale * expf(raw) * sel;
  }
}

constexpr int kStageRows = 2 * PH + 2 * PMAXL + 1;  // dh[16] | g*h[16] | feat[<=16] | g
constexpr int kStageStride = kPts + 4;              // rows 16-byte aligned; 8 rows x 4 floats cover all 32 banks

template <int L, bool NEED_DX>
__global__ void __launch_bounds__(kPts, 4) prop_bwd_kernel(
    const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ ebins,
    const float* __restrict__ table, LevelScales sc, int log2T, int n_coarse, const float* __restrict__ w1,
    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, float scale,
    const float* __restrict__ d_density, int64_t R, int S, float* __restrict__ dtable, float* __restrict__ dw1,
    float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ d_origins,
    float* __restrict__ d_dirs) {
  __shared__ PropWeights sw;
  __shared__ float s_scale[TN_MAX_LEVELS];
  __shared__ __align__(16) float stage[kStageRows * kStageStride];
  constexpr int IN = 2 * L;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < TN_MAX_LEVELS) s_scale[tid] = sc.s[tid];
  load_prop_weights(sw, w1, b1, w2, b2, IN);
  __syncthreads();
  const uint32_t T = 1u << log2T, mask = T - 1u;
  const int64_t N = R * S;
  const int64_t tiles = (N + kPts - 1) / kPts;
  // weight-gradient entries owned by this thread: e0 = tid, e1 = tid + 128 over
  //   [0, PH*IN): dW1[j][k] ; then db1[PH] ; dW2[PH] ; db2
  constexpr int NE = PH * IN + 2 * PH + 1;
  float acc0 = 0.f, acc1 = 0.f;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t p = t * kPts + tid;
    const bool valid = p < N;
    const int64_t r = valid ? p / S : 0;
    const int s = valid ? (int)(p - r * S) : 0;
    float st = 0.f, en = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (valid) {
      st = __ldg(ebins + r * (S + 1) + s); en = __ldg(ebins + r * (S + 1) + s + 1);
      o0 = __ldg(origins + 3 * r); o1 = __ldg(origins + 3 * r + 1); o2 = __ldg(origins + 3 * r + 2);
      d0 = __ldg(dirs + 3 * r); d1 = __ldg(dirs + 3 * r + 1); d2 = __ldg(dirs + 3 * r + 2);
    }
    const float p0 = sample_pos(o0, d0, st, en), p1 = sample_pos(o1, d1, st, en), p2 = sample_pos(o2, d2, st, en);
    float x0 = p0, x1 = p1, x2 = p2;
    const float sel = contract_normalise(x0, x1, x2);
    float feat[IN], h[PH], jac[NEED_DX ? IN : 1][3];
    prop_features<L, NEED_DX>(table, s_scale, mask, T, x0, x1, x2, feat, jac);
    const float raw = prop_mlp<L>(sw, feat, h);
    // d(density)/d(raw) = scale * sel * exp(clamp(raw, -15, 15))   (activations.py:41)
    const float g = valid ? __ldg(d_density + p) * scale * sel * expf(fminf(fmaxf(raw, -15.f), 15.f)) : 0.f;
    float dfeat[IN];
#pragma unroll
    for (int k = 0; k < IN; ++k) dfeat[k] = 0.f;
    __syncthreads();  // previous tile's reduction has finished reading the stage
#pragma unroll
    for (int j = 0; j < PH; ++j) {
      const float dh = h[j] > 0.f ? g * sw.w2[j] : 0.f;
      stage[j * kStageStride + tid] = dh;
      stage[(PH + j) * kStageStride + tid] = g * h[j];
#pragma unroll
      for (int k = 0; k < IN; ++k) dfeat[k] = fmaf(dh, sw.w1[j][k], dfeat[k]);
    }
#pragma unroll
    for (int k = 0; k < IN; ++k) stage[(2 * PH + k) * kStageStride + tid] = feat[k];
    stage[(2 * PH + 2 * PMAXL) * kStageStride + tid] = g;
    // ---- dL/dx from the Jacobian of the first pass, then the table scatter
    float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
    if constexpr (NEED_DX) {
#pragma unroll
      for (int k = 0; k < IN; ++k) {
        dx0 = fmaf(dfeat[k], jac[k][0], dx0);
        dx1 = fmaf(dfeat[k], jac[k][1], dx1);
        dx2 = fmaf(dfeat[k], jac[k][2], dx2);
      }
    }
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const Cell c = locate(x0, x1, x2, s_scale[l], mask, (uint32_t)l * T);
      const float mx = 1.f - c.ox, my = 1.f - c.oy, mz = 1.f - c.oz;
      float gc[8][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float gj = dfeat[l * 2 + j];
        const float g0312 = gj * c.oz, g4756 = gj * mz;
        const float g03 = g0312 * c.oy, g12 = g0312 * my;
        const float g47 = g4756 * c.oy, g56 = g4756 * my;
        gc[0][j] = g03 * c.ox; gc[3][j] = g03 * mx;
        gc[1][j] = g12 * c.ox; gc[2][j] = g12 * mx;
        gc[5][j] = g56 * c.ox; gc[6][j] = g56 * mx;
        gc[4][j] = g47 * c.ox; gc[7][j] = g47 * mx;
      }
      bool issue = valid;
      if (l < n_coarse) {  // warp-uniform: fold contiguous runs of equal cells into their head lane
        const uint64_t key = valid ? c.key : ~0ull;
        const uint64_t pkey = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (pkey != key);
        const int run = __popc(__ballot_sync(0xffffffffu, head) & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int orun = __shfl_down_sync(0xffffffffu, run, d);
          const bool take = (lane + d < 32) && (orun == run);
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float o = __shfl_down_sync(0xffffffffu, gc[k][j], d);
              if (take) gc[k][j] += o;
            }
        }
        issue = valid && head;
      }
      if (issue) red_cell<2>(dtable, c, gc);
    }
    // ---- dL/dx -> contraction backward -> per-ray (origin, direction) gradients, folded over runs of one ray
    if constexpr (NEED_DX) {
      float g0 = dx0, g1 = dx1, g2 = dx2;
      contract_normalise_bwd(p0, p1, p2, g0, g1, g2);
      const float tmid = (st + en) / 2.f;
      float v[6] = {g0, g1, g2, g0 * tmid, g1 * tmid, g2 * tmid};
      const int64_t key = valid ? r : -1;
      const int64_t pkey = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = (lane == 0) || (pkey != key);
      const int run = __popc(__ballot_sync(0xffffffffu, head) & (0xffffffffu >> (31 - lane)));
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int orun = __shfl_down_sync(0xffffffffu, run, d);
        const bool take = (lane + d < 32) && (orun == run);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float o = __shfl_down_sync(0xffffffffu, v[k], d);
          if (take) v[k] += o;
        }
      }
      if (valid && head) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          atomicAdd(d_origins + 3 * r + k, v[k]);
          atomicAdd(d_dirs + 3 * r + k, v[3 + k]);
        }
      }
    }
    // ---- weight gradients of this tile: column sums / dot products over the 128 staged points
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int e = tid + half * kPts;
      if (e < NE) {
        const float* a;
        const float* b = nullptr;
        if (e < PH * IN) {
          a = stage + (e / IN) * kStageStride;                 // dh[j]
          b = stage + (2 * PH + e % IN) * kStageStride;        // feat[k]
        } else if (e < PH * IN + PH) {
          a = stage + (e - PH * IN) * kStageStride;            // db1[j] = sum dh[j]
        } else if (e < PH * IN + 2 * PH) {
          a = stage + (PH + e - PH * IN - PH) * kStageStride;  // dW2[j] = sum g*h[j]
        } else {
          a = stage + (2 * PH + 2 * PMAXL) * kStageStride;     // db2 = sum g
        }
        // rows are read four points at a time (LDS.128): two loads feed four FMAs
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* a4 = reinterpret_cast<const float4*>(a);
        if (b) {
          const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll 8
          for (int q = 0; q < kPts / 4; ++q) {
            const float4 x = a4[q], y = b4[q];
            s4.x = fmaf(x.x, y.x, s4.x); s4.y = fmaf(x.y, y.y, s4.y);
            s4.z = fmaf(x.z, y.z, s4.z); s4.w = fmaf(x.w, y.w, s4.w);
          }
        } else {
#pragma unroll 8
          for (int q = 0; q < kPts / 4; ++q) {
            const float4 x = a4[q];
            s4.x += x.x; s4.y += x.y; s4.z += x.z; s4.w += x.w;
          }
        }
        const float sum = (s4.x + s4.y) + (s4.z + s4.w);
        if (half == 0) acc0 += sum; else acc1 += sum;
      }
    }
  }
  // ---- flush the weight-gradient accumulators
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int e = tid + half * kPts;
    const float v = half == 0 ? acc0 : acc1;
    if (e < PH * IN) atomicAdd(dw1 + e, v);  // [j][k] row-major with k < IN: same flattening as nn.Linear
    else if (e < PH * IN + PH) atomicAdd(db1 + (e - PH * IN), v);
    else if (e < PH * IN + 2 * PH) atomicAdd(dw2 + (e - PH * IN - PH), v);
    else if (e < NE) atomicAdd(db2, v);
  }
}

static int prop_check(const float* origins, const float* dirs, const float* ebins, const void* table,
                      const float* scales_host, int64_t R, int S, int L, int log2_T, int hidden) {
  TN_REQUIRE(origins && dirs && ebins && table && scales_host, TN_EINVAL, "prop_density: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "prop_density: bad R=%lld S=%d", (long long)R, S);
  TN_REQUIRE(L >= 1 && L <= PMAXL, TN_EINVAL, "prop_density: L=%d not in [1,%d]", L, PMAXL);
  TN_REQUIRE(hidden == PH, TN_EINVAL, "prop_density: hidden width %d (fused kernel handles %d)", hidden, PH);
  TN_REQUIRE(log2_T >= 1 && log2_T <= 26, TN_EINVAL, "prop_density: log2_T=%d", log2_T);
  TN_REQUIRE(aligned(table, 16), TN_EALIGN, "prop_density: table must be 16-byte aligned");
  return TN_OK;
}

}  // namespace tn

using namespace tn;

#define TN_PROP_L(FN, ...)            \
  switch (L) {                        \
    case 1: FN(1, __VA_ARGS__); break; \
    case 2: FN(2, __VA_ARGS__); break; \
    case 3: FN(3, __VA_ARGS__); break; \
    case 4: FN(4, __VA_ARGS__); break; \
    case 5: FN(5, __VA_ARGS__); break; \
    case 6: FN(6, __VA_ARGS__); break; \
    case 7: FN(7, __VA_ARGS__); break; \
    default: FN(8, __VA_ARGS__); break; \
  }

extern "C" int tn_prop_density_fwd(const float* origins, const float* directions, const float* ebins, const float* table,
                                   const float* scales_host, int64_t R, int S, int L, int log2_T, int hidden,
                                   const float* w1, const float* b1, const float* w2, const float* b2,
                                   float density_scale, float* density_out, void* stream) {
  int rc = prop_check(origins, directions, ebins, table, scales_host, R, S, L, log2_T, hidden);
  if (rc) return rc;
  TN_REQUIRE(w1 && b1 && w2 && b2 && density_out, TN_EINVAL, "prop_density_fwd: null pointer");
  if (R == 0) return TN_OK;
  LevelScales sc;
  for (int l = 0; l < TN_MAX_LEVELS; ++l) sc.s[l] = l < L ? scales_host[l] : 0.f;
  const int64_t N = R * S;
  const unsigned grid = (unsigned)min((N + kPts - 1) / kPts, (int64_t)kNumSMs * 16);
  cudaStream_t st = (cudaStream_t)stream;
#define TN_PF(LL, ...) prop_fwd_kernel<LL><<<grid, kPts, 0, st>>>(__VA_ARGS__)
  TN_PROP_L(TN_PF, origins, directions, ebins, table, sc, log2_T, w1, b1, w2, b2, density_scale, R, S, density_out);
#undef TN_PF
  return check_launch("prop_fwd_kernel");
}

extern "C" int tn_prop_density_bwd(const float* origins, const float* directions, const float* ebins, const float* table,
                                   const float* scales_host, int64_t R, int S, int L, int log2_T, int hidden,
                                   const float* w1, const float* b1, const float* w2, const float* b2,
                                   float density_scale, const float* d_density, float* dtable, float* dw1, float* db1,
                                   float* dw2, float* db2, float* d_origins, float* d_directions, void* stream) {
  int rc = prop_check(origins, directions, ebins, table, scales_host, R, S, L, log2_T, hidden);
  if (rc) return rc;
  TN_REQUIRE(w1 && b1 && w2 && b2 && d_density && dtable && dw1 && db1 && dw2 && db2, TN_EINVAL,
             "prop_density_bwd: null pointer");
  TN_REQUIRE((d_origins == nullptr) == (d_directions == nullptr), TN_EINVAL,
             "prop_density_bwd: d_origins and d_directions must both be given or both be NULL");
  TN_REQUIRE(aligned(dtable, 16), TN_EALIGN, "prop_density_bwd: dtable must be 16-byte aligned");
  if (R == 0) return TN_OK;
  LevelScales sc;
  int n_coarse = 0;
  for (int l = 0; l < TN_MAX_LEVELS; ++l) {
    sc.s[l] = l < L ? scales_host[l] : 0.f;
    if (l < L && l == n_coarse && scales_host[l] <= agg_threshold_prop()) ++n_coarse;
  }
  const int64_t N = R * S;
  const unsigned grid = (unsigned)min((N + kPts - 1) / kPts, (int64_t)kNumSMs * 4);
  cudaStream_t st = (cudaStream_t)stream;
#define TN_PB(LL, ...)                                                                 \
  do {                                                                                 \
    if (d_origins) prop_bwd_kernel<LL, true><<<grid, kPts, 0, st>>>(__VA_ARGS__);      \
    else prop_bwd_kernel<LL, false><<<grid, kPts, 0, st>>>(__VA_ARGS__);               \
  } while (0)
  TN_PROP_L(TN_PB, origins, directions, ebins, table, sc, log2_T, n_coarse, w1, b1, w2, b2, density_scale, d_density, R,
            S, dtable, dw1, db1, dw2, db2, d_origins, d_directions);
#undef TN_PB
  return check_launch("prop_bwd_kernel");
}
