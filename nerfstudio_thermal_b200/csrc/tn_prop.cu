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
// dL/d(origin, direction).  Weight gradients: the kernel is bound by the SM's load/store pipe (gathers, REDs,
// shuffles and shared-memory traffic share it: ncu l1tex 80 % of peak at 45 % issue), so the 193 sums over points
// (dW1, db1, dW2, db2) run on the tensor cores instead of as shared-memory dot products: every warp stages
// (dh, g*h, features) of its 32 points once (conflict-free 4-byte rows) and accumulates
//     [dh ; g*h]^T (32 x points)  .  [features | 1] (points x 11)
// with mma.sync.m16n8k8 (tf32 operands split in two terms, three MMAs per product: ~2^-21 relative) in register
// accumulators that live across the tiles of a persistent CTA; the warps' partial sums are combined in shared
// memory and leave as one atomicAdd per entry per CTA.
// Compiled with -fmad=false (index math must round like the reference); the MLP uses explicit fmaf().
#include "tn_encode_core.cuh"
#include "tn_geometry.cuh"

namespace tn {

#ifndef TN_PROP_FWD_MB
#define TN_PROP_FWD_MB 4
#endif
#ifndef TN_PROP_BWD_MB
#define TN_PROP_BWD_MB 4
#endif
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
          // trilinear weights (encodings.py:443-461).  The products may be contracted here (explicit fmaf in a
          // -fmad=false unit): the proposal densities only steer the sampling, 1-ulp differences are far inside
          // the 1e-5 parity bar of this field, and the kernels are short of issue slots, not of precision
          const float f03 = fmaf(f[b][0][j], ox, f[b][3][j] * mx);
          const float f12 = fmaf(f[b][1][j], ox, f[b][2][j] * mx);
          const float f56 = fmaf(f[b][5][j], ox, f[b][6][j] * mx);
          const float f47 = fmaf(f[b][4][j], ox, f[b][7][j] * mx);
          const float f0312 = fmaf(f03, oy, f12 * my);
          const float f4756 = fmaf(f47, oy, f56 * my);
          feat[(l0 + b) * 2 + j] = fmaf(f0312, oz, f4756 * mz);
          if constexpr (JAC) {
            const float sl = s_scale[l0 + b];
            const float e03 = f[b][0][j] - f[b][3][j], e12 = f[b][1][j] - f[b][2][j];
            const float e56 = f[b][5][j] - f[b][6][j], e47 = f[b][4][j] - f[b][7][j];
            jac[(l0 + b) * 2 + j][0] = fmaf(fmaf(e03, oy, e12 * my), oz, fmaf(e47, oy, e56 * my) * mz) * sl;
            jac[(l0 + b) * 2 + j][1] = fmaf(f03 - f12, oz, (f47 - f56) * mz) * sl;
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
__global__ void __launch_bounds__(kPts, TN_PROP_FWD_MB) prop_fwd_kernel(const float* __restrict__ origins,
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

constexpr int kStageRows = 2 * PH + 2 * PMAXL;  // per warp: dh[16] | g*h[16] | feat[<=16], one column per lane
constexpr int kStageStride = 36;                // = 4 mod 32: the MMA fragment reads (row g, column t) hit 32 banks

// D(16x8) += A(16x8, row) . B(8x8, col), tf32 operands, fp32 accumulate.  Fragments (g = lane/4, t = lane%4):
// a = {A[g][t], A[g+8][t], A[g][t+4], A[g+8][t+4]}, b = {B[t][g], B[t+4][g]}, d = {D[g][2t], D[g][2t+1], D[g+8][2t], ..}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo exactly; hi holds the top 11 significand bits (a tf32), the tensor core reads the top 11 of lo
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
template <int NV>
__device__ __forceinline__ void split_frag(const float (&x)[NV], uint32_t (&hi)[NV], uint32_t (&lo)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) split_tf32(x[i], hi[i], lo[i]);
}
__device__ __forceinline__ void mma_split(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

template <int L, bool NEED_DX>
__global__ void __launch_bounds__(kPts, TN_PROP_BWD_MB) prop_bwd_kernel(
    const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ ebins,
    const float* __restrict__ table, LevelScales sc, int log2T, int n_coarse, const float* __restrict__ w1,
    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, float scale,
    const float* __restrict__ d_density, int64_t R, int S, float* __restrict__ dtable, float* __restrict__ dw1,
    float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ d_origins,
    float* __restrict__ d_dirs) {
  __shared__ PropWeights sw;
  __shared__ float s_scale[TN_MAX_LEVELS];
  __shared__ float stage[(kPts / 32) * kStageRows * kStageStride];
  constexpr int IN = 2 * L;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < TN_MAX_LEVELS) s_scale[tid] = sc.s[tid];
  load_prop_weights(sw, w1, b1, w2, b2, IN);
  __syncthreads();
  const uint32_t T = 1u << log2T, mask = T - 1u;
  const int64_t N = R * S;
  const int64_t tiles = (N + kPts - 1) / kPts;
  // weight-gradient accumulators of this warp (MMA D fragments): dh x feature tile q (the "ones" column, whose sums
  // are db1, sits at column IN % 8 of tile IN / 8), g*h x the ones tile (column IN % 8 = dW2), and sum of g (db2)
  constexpr int NTL = IN / 8 + 1, OT = IN / 8, OC = IN % 8;
  float accW[NTL][4], accV[4] = {0.f, 0.f, 0.f, 0.f}, acc_g = 0.f;
#pragma unroll
  for (int q = 0; q < NTL; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) accW[q][e] = 0.f;
  float* wst = stage + (tid >> 5) * kStageRows * kStageStride;  // this warp's staging rows
  const int fg = lane >> 2, ft = lane & 3;                      // fragment coordinates

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t p = t * kPts + tid;
    const bool valid = p < N;
    // The interlevel loss is a hinge (losses.py:87-103): rays whose proposal histogram already bounds the fine
    // weights send exactly zero down here.  A warp whose 32 samples all carry a zero upstream gradient contributes
    // nothing to any output (table, weights, rays): it skips the tile before the first gather.
    const float up = valid ? __ldg(d_density + p) : 0.f;
    if (!__any_sync(0xffffffffu, up != 0.f)) continue;
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
    const float g = valid ? up * scale * sel * expf(fminf(fmaxf(raw, -15.f), 15.f)) : 0.f;
    float dfeat[IN];
#pragma unroll
    for (int k = 0; k < IN; ++k) dfeat[k] = 0.f;
    __syncwarp();  // the previous tile's MMA fragments have been read
#pragma unroll
    for (int j = 0; j < PH; ++j) {
      const float dh = h[j] > 0.f ? g * sw.w2[j] : 0.f;
      wst[j * kStageStride + lane] = dh;
      wst[(PH + j) * kStageStride + lane] = g * h[j];
#pragma unroll
      for (int k = 0; k < IN; ++k) dfeat[k] = fmaf(dh, sw.w1[j][k], dfeat[k]);
    }
#pragma unroll
    for (int k = 0; k < IN; ++k) wst[(2 * PH + k) * kStageStride + lane] = feat[k];
    acc_g += g;
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
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
        // step d of the tree is needed only if some run is longer than d, i.e. if the non-head lanes contain a
        // streak of >= d consecutive bits (`need`); fine levels with short runs stop after one or two steps
        unsigned need = ~heads;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          if (need == 0u) break;
          const int orun = __shfl_down_sync(0xffffffffu, run, d);
          const bool take = (lane + d < 32) && (orun == run);
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float o = __shfl_down_sync(0xffffffffu, gc[k][j], d);
              if (take) gc[k][j] += o;
            }
          need &= need >> d;
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
    // ---- weight gradients of this warp's 32 points: four k-steps of 8 points
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float* col = wst + ks * 8 + ft;
      float a1[4], a2[4];
      a1[0] = col[fg * kStageStride]; a1[1] = col[(fg + 8) * kStageStride];
      a1[2] = col[fg * kStageStride + 4]; a1[3] = col[(fg + 8) * kStageStride + 4];
      a2[0] = col[(PH + fg) * kStageStride]; a2[1] = col[(PH + fg + 8) * kStageStride];
      a2[2] = col[(PH + fg) * kStageStride + 4]; a2[3] = col[(PH + fg + 8) * kStageStride + 4];
      uint32_t a1h[4], a1l[4], a2h[4], a2l[4];
      split_frag<4>(a1, a1h, a1l);
      split_frag<4>(a2, a2h, a2l);
#pragma unroll
      for (int q = 0; q < NTL; ++q) {
        const int fr = q * 8 + fg;  // this lane's column of tile q: feature fr, the ones column, or padding
        const int frc = fr < IN ? fr : 0;
        float b[2];
        b[0] = col[(2 * PH + frc) * kStageStride];
        b[1] = col[(2 * PH + frc) * kStageStride + 4];
        if (fr >= IN) b[0] = b[1] = (q == OT && fg == OC) ? 1.f : 0.f;
        uint32_t bh[2], bl[2];
        split_frag<2>(b, bh, bl);
        mma_split(accW[q], a1h, a1l, bh, bl);
        if (q == OT) mma_split(accV, a2h, a2l, bh, bl);
      }
    }
  }
  // ---- flush: the warps' fragments -> one [NE] vector per warp in shared memory -> one atomicAdd per entry per CTA
  //   entries: [0, PH*IN): dW1[j][k] (nn.Linear flattening) ; then db1[PH] ; dW2[PH] ; db2
  constexpr int NE = PH * IN + 2 * PH + 1;
  static_assert(NE <= kStageRows * kStageStride, "flush vector must fit in a warp's staging rows");
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NTL; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = fg + (e >= 2 ? 8 : 0), k = q * 8 + 2 * ft + (e & 1);
      if (k < IN) wst[j * IN + k] = accW[q][e];
      else if (q == OT && k - q * 8 == OC) wst[PH * IN + j] = accW[q][e];
    }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (2 * ft + (e & 1) == OC) wst[PH * IN + PH + fg + (e >= 2 ? 8 : 0)] = accV[e];
  const float gsum = warp_sum(acc_g);
  if (lane == 0) wst[PH * IN + 2 * PH] = gsum;
  __syncthreads();
  for (int e = tid; e < NE; e += kPts) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kPts / 32; ++w) v += stage[w * kStageRows * kStageStride + e];
    if (e < PH * IN) atomicAdd(dw1 + e, v);
    else if (e < PH * IN + PH) atomicAdd(db1 + (e - PH * IN), v);
    else if (e < PH * IN + 2 * PH) atomicAdd(dw2 + (e - PH * IN - PH), v);
    else atomicAdd(db2, v);
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
  const unsigned grid = (unsigned)min((N + kPts - 1) / kPts, (int64_t)kNumSMs * TN_PROP_BWD_MB);
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
