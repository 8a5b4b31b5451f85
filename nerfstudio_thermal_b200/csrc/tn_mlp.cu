// Fully fused small MLPs, float32 SIMT path (exact-fp32 parity path of the density / colour / proposal
// heads; the tcgen05 tensor-core path lives in tn_mlp_tc.cu).
//
// One CTA (256 threads) owns a tile of P = 64 points; all layers of the network run back to back on the
// tile with weights and activations in shared memory, so hidden activations never touch HBM.  Activations
// are kept feature-major ([feature][point], row stride PS = 68 floats) so that every product in forward
// AND backward is one of two register-blocked shared-memory GEMM forms:
//   TN:  O[j][p]  = sum_k B1[k][j] * B2[k][p]      (layer forward with B1 = W^T; dA = W . dZ with B1 = W)
//   NT:  C[j][k] += sum_p B1[j][p] * B2[k][p]      (weight gradients, accumulated in registers across tiles)
// The backward kernel recomputes the forward activations from x (cheaper than saving 2 x 256 B/point),
// accumulates dW/db in registers over a persistent grid-stride loop and flushes them with one atomicAdd
// per weight per CTA.
#include "tn_common.cuh"

namespace tn {

constexpr int P = 64;     // points per tile
constexpr int PS = 68;    // padded row stride (floats) of feature-major activation buffers
constexpr int NT = 256;   // threads per CTA

enum { ACT_NONE = 0, ACT_SIGMOID = 1, ACT_TRUNC_EXP = 2 };

struct MlpParams {
  const float* w[3];
  const float* b[3];
  float* dw[3];
  float* db[3];
  int in_dim, out_dim, n_layers, out_act;
};

// O[j][p] = epi( bias[j] + sum_k B1[k*NOUT + j] * B2[k*PS + p] )
template <int K, int NOUT, bool BIAS, bool RELU, bool MASK>
__device__ __forceinline__ void gemm_tn(const float* __restrict__ B1, const float* __restrict__ bias,
                                        const float* __restrict__ B2, float* __restrict__ O,
                                        const float* __restrict__ mask) {
  constexpr int JB = NOUT >= 16 ? NOUT / 16 : 1;  // 64 -> 4, 32 -> 2, 16 -> 1, 4 -> 1
  constexpr int TJ = NOUT / JB;
  static_assert(TJ <= 16 && (JB == 1 || JB == 2 || JB == 4), "gemm_tn: unsupported NOUT");
  const int tp = threadIdx.x & 15, tj = threadIdx.x >> 4;
  if (tj < TJ) {
    float acc[JB][4];
#pragma unroll
    for (int a = 0; a < JB; ++a) {
      const float bv = BIAS ? bias[tj * JB + a] : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[a][q] = bv;
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(B2 + k * PS + 4 * tp);
      float wv[JB];
      if constexpr (JB == 4) {
        const float4 t = *reinterpret_cast<const float4*>(B1 + k * NOUT + 4 * tj);
        wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
      } else if constexpr (JB == 2) {
        const float2 t = *reinterpret_cast<const float2*>(B1 + k * NOUT + 2 * tj);
        wv[0] = t.x; wv[1] = t.y;
      } else {
        wv[0] = B1[k * NOUT + tj];
      }
#pragma unroll
      for (int a = 0; a < JB; ++a) {
        acc[a][0] = fmaf(wv[a], av.x, acc[a][0]);
        acc[a][1] = fmaf(wv[a], av.y, acc[a][1]);
        acc[a][2] = fmaf(wv[a], av.z, acc[a][2]);
        acc[a][3] = fmaf(wv[a], av.w, acc[a][3]);
      }
    }
#pragma unroll
    for (int a = 0; a < JB; ++a) {
      const int j = tj * JB + a;
      float4 o = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      if constexpr (RELU) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      if constexpr (MASK) {
        const float4 m = *reinterpret_cast<const float4*>(mask + j * PS + 4 * tp);
        o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f;
        o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
      }
      *reinterpret_cast<float4*>(O + j * PS + 4 * tp) = o;
    }
  }
}

// thread decomposition of an [NJ][NK] weight-gradient block
template <int NJ, int NK>
struct NtShape {
  static constexpr int TJ = NJ < 16 ? NJ : 16;
  static constexpr int JB = NJ / TJ;
  static constexpr int TKmax = NT / TJ;
  static constexpr int TK = NK < TKmax ? NK : (TKmax > 16 && NK >= 64 && TJ <= 4 ? 64 : 16);
  static constexpr int KB = NK / TK;
  static constexpr int NACC = JB * KB;
};

// acc[a][b] += sum_p B1[(tj + TJ*a)*PS + p] * B2[(tk + TK*b)*PS + p]
template <int NJ, int NK>
__device__ __forceinline__ void gemm_nt_acc(const float* __restrict__ B1, const float* __restrict__ B2,
                                            float (&acc)[NtShape<NJ, NK>::JB][NtShape<NJ, NK>::KB]) {
  using S = NtShape<NJ, NK>;
  const int tk = threadIdx.x % S::TK, tj = threadIdx.x / S::TK;
  if (tj < S::TJ) {
#pragma unroll 2
    for (int p = 0; p < P; p += 4) {
      float4 u[S::JB], v[S::KB];
#pragma unroll
      for (int a = 0; a < S::JB; ++a) u[a] = *reinterpret_cast<const float4*>(B1 + (tj + S::TJ * a) * PS + p);
#pragma unroll
      for (int b = 0; b < S::KB; ++b) v[b] = *reinterpret_cast<const float4*>(B2 + (tk + S::TK * b) * PS + p);
#pragma unroll
      for (int a = 0; a < S::JB; ++a)
#pragma unroll
        for (int b = 0; b < S::KB; ++b) {
          acc[a][b] = fmaf(u[a].x, v[b].x, acc[a][b]);
          acc[a][b] = fmaf(u[a].y, v[b].y, acc[a][b]);
          acc[a][b] = fmaf(u[a].z, v[b].z, acc[a][b]);
          acc[a][b] = fmaf(u[a].w, v[b].w, acc[a][b]);
        }
    }
  }
}

template <int NJ, int NK>
__device__ __forceinline__ void flush_dw(float* __restrict__ dw, int nj_real, int nk_real,
                                         const float (&acc)[NtShape<NJ, NK>::JB][NtShape<NJ, NK>::KB]) {
  using S = NtShape<NJ, NK>;
  const int tk = threadIdx.x % S::TK, tj = threadIdx.x / S::TK;
  if (tj < S::TJ) {
#pragma unroll
    for (int a = 0; a < S::JB; ++a)
#pragma unroll
      for (int b = 0; b < S::KB; ++b) {
        const int j = tj + S::TJ * a, k = tk + S::TK * b;
        if (j < nj_real && k < nk_real) atomicAdd(dw + (size_t)j * nk_real + k, acc[a][b]);
      }
  }
}

// load nn.Linear weight [out_real][in_real] into smem as orig [NOUT][NIN] and/or transposed [NIN][NOUT], zero padded
template <int NIN, int NOUT>
__device__ __forceinline__ void load_weight(const float* __restrict__ w, int in_real, int out_real,
                                            float* __restrict__ s_orig, float* __restrict__ s_t) {
  for (int i = threadIdx.x; i < NIN * NOUT; i += NT) {
    const int j = i / NIN, k = i - j * NIN;
    const float v = (j < out_real && k < in_real) ? __ldg(w + (size_t)j * in_real + k) : 0.f;
    if (s_orig) s_orig[j * NIN + k] = v;
    if (s_t) s_t[k * NOUT + j] = v;
  }
}
template <int NOUT>
__device__ __forceinline__ void load_bias(const float* __restrict__ b, int out_real, float* __restrict__ s) {
  for (int i = threadIdx.x; i < NOUT; i += NT) s[i] = i < out_real ? __ldg(b + i) : 0.f;
}

// global row-major [rows][dim_real] -> smem feature-major [DIM][PS] (zero padded)
template <int DIM>
__device__ __forceinline__ void load_tile_t(const float* __restrict__ g, int64_t row0, int rows, int dim_real,
                                            float* __restrict__ s) {
  const float* src = g + row0 * dim_real;
  const int n = rows * dim_real;
  for (int i = threadIdx.x; i < P * DIM; i += NT) {  // zero fill (covers padding rows / features)
    const int k = i / P, p = i - k * P;
    s[k * PS + p] = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += NT) {
    const int p = i / dim_real, k = i - p * dim_real;
    s[k * PS + p] = __ldg(src + i);
  }
}
template <int DIM>
__device__ __forceinline__ void store_tile_t(float* __restrict__ g, int64_t row0, int rows, int dim_real,
                                             const float* __restrict__ s) {
  float* dst = g + row0 * dim_real;
  const int n = rows * dim_real;
  for (int i = threadIdx.x; i < n; i += NT) {
    const int p = i / dim_real, k = i - p * dim_real;
    dst[i] = s[k * PS + p];
  }
}

__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  if (act == ACT_TRUNC_EXP) return expf(z);
  return z;
}
// derivative given pre-activation z (trunc_exp backward clamps, activations.py:41)
__device__ __forceinline__ float act_bwd(float z, int act) {
  if (act == ACT_SIGMOID) {
    const float s = 1.f / (1.f + expf(-z));
    return s * (1.f - s);
  }
  if (act == ACT_TRUNC_EXP) return expf(fminf(fmaxf(z, -15.f), 15.f));
  return 1.f;
}

template <int IN, int W, int OUT, int NL>
struct Smem {
  // forward: transposed weights; backward additionally the original layout
  static constexpr int w1t = 0;
  static constexpr int w2t = w1t + IN * W;
  static constexpr int w3t = w2t + (NL == 3 ? W * W : 0);
  static constexpr int bias = w3t + W * OUT;                 // b1[W] b2[W] b3[OUT] (b2 unused if NL==2)
  static constexpr int a0 = bias + ((2 * W + OUT + 3) / 4) * 4;
  static constexpr int a1 = a0 + IN * PS;
  static constexpr int a2 = a1 + W * PS;
  static constexpr int ao = a2 + (NL == 3 ? W * PS : 0);
  static constexpr int fwd_total = ao + OUT * PS;
  // backward extras
  static constexpr int w1 = fwd_total;
  static constexpr int w2 = w1 + IN * W;
  static constexpr int w3 = w2 + (NL == 3 ? W * W : 0);
  static constexpr int dza = w3 + W * OUT;                   // [max(W,IN)][PS]
  static constexpr int dzb = dza + (W > IN ? W : IN) * PS;   // [max(W,IN)][PS]
  static constexpr int bwd_total = dzb + (W > IN ? W : IN) * PS;
};

template <int IN, int W, int OUT, int NL>
__device__ __forceinline__ void forward_tile(float* sm, int out_act, bool apply_act) {
  using L = Smem<IN, W, OUT, NL>;
  gemm_tn<IN, W, true, true, false>(sm + L::w1t, sm + L::bias, sm + L::a0, sm + L::a1, nullptr);
  __syncthreads();
  const float* last = sm + L::a1;
  if constexpr (NL == 3) {
    gemm_tn<W, W, true, true, false>(sm + L::w2t, sm + L::bias + W, sm + L::a1, sm + L::a2, nullptr);
    __syncthreads();
    last = sm + L::a2;
  }
  gemm_tn<W, OUT, true, false, false>(sm + L::w3t, sm + L::bias + 2 * W, last, sm + L::ao, nullptr);
  __syncthreads();
  if (apply_act && out_act != ACT_NONE) {
    for (int i = threadIdx.x; i < OUT * P; i += NT) {
      const int j = i / P, p = i - j * P;
      sm[L::ao + j * PS + p] = act_fwd(sm[L::ao + j * PS + p], out_act);
    }
    __syncthreads();
  }
}

template <int IN, int W, int OUT, int NL>
__global__ void __launch_bounds__(NT) mlp_fwd_kernel(const float* __restrict__ x, int64_t N, MlpParams prm,
                                                     float* __restrict__ y) {
  extern __shared__ float4 smem4[];
  float* sm = reinterpret_cast<float*>(smem4);
  using L = Smem<IN, W, OUT, NL>;
  load_weight<IN, W>(prm.w[0], prm.in_dim, W, nullptr, sm + L::w1t);
  if constexpr (NL == 3) load_weight<W, W>(prm.w[1], W, W, nullptr, sm + L::w2t);
  load_weight<W, OUT>(prm.w[NL - 1], W, prm.out_dim, nullptr, sm + L::w3t);
  load_bias<W>(prm.b[0], W, sm + L::bias);
  if constexpr (NL == 3) load_bias<W>(prm.b[1], W, sm + L::bias + W);
  load_bias<OUT>(prm.b[NL - 1], prm.out_dim, sm + L::bias + 2 * W);
  const int64_t tiles = (N + P - 1) / P;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * P;
    const int rows = (int)min((int64_t)P, N - row0);
    __syncthreads();
    load_tile_t<IN>(x, row0, rows, prm.in_dim, sm + L::a0);
    __syncthreads();
    forward_tile<IN, W, OUT, NL>(sm, prm.out_act, true);
    store_tile_t<OUT>(y, row0, rows, prm.out_dim, sm + L::ao);
  }
}

template <int IN, int W, int OUT, int NL, bool NEED_DX>
__global__ void __launch_bounds__(NT) mlp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                     int64_t N, MlpParams prm, float* __restrict__ dx) {
  extern __shared__ float4 smem4[];
  float* sm = reinterpret_cast<float*>(smem4);
  using L = Smem<IN, W, OUT, NL>;
  load_weight<IN, W>(prm.w[0], prm.in_dim, W, sm + L::w1, sm + L::w1t);
  if constexpr (NL == 3) load_weight<W, W>(prm.w[1], W, W, sm + L::w2, sm + L::w2t);
  load_weight<W, OUT>(prm.w[NL - 1], W, prm.out_dim, sm + L::w3, sm + L::w3t);
  load_bias<W>(prm.b[0], W, sm + L::bias);
  if constexpr (NL == 3) load_bias<W>(prm.b[1], W, sm + L::bias + W);
  load_bias<OUT>(prm.b[NL - 1], prm.out_dim, sm + L::bias + 2 * W);

  float dw1[NtShape<W, IN>::JB][NtShape<W, IN>::KB] = {};
  float dw2[NtShape<W, W>::JB][NtShape<W, W>::KB] = {};
  float dw3[NtShape<OUT, W>::JB][NtShape<OUT, W>::KB] = {};
  float db1 = 0.f, db2 = 0.f, db3 = 0.f;  // thread j < W (or OUT) owns bias j

  const int64_t tiles = (N + P - 1) / P;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * P;
    const int rows = (int)min((int64_t)P, N - row0);
    __syncthreads();
    load_tile_t<IN>(x, row0, rows, prm.in_dim, sm + L::a0);
    __syncthreads();
    forward_tile<IN, W, OUT, NL>(sm, prm.out_act, false);  // ao holds pre-activations z
    // dZ_out = dy * act'(z), feature-major into dza[OUT][PS]
    load_tile_t<OUT>(dy, row0, rows, prm.out_dim, sm + L::dza);
    __syncthreads();
    if (prm.out_act != ACT_NONE) {
      for (int i = threadIdx.x; i < OUT * P; i += NT) {
        const int j = i / P, p = i - j * P;
        sm[L::dza + j * PS + p] *= act_bwd(sm[L::ao + j * PS + p], prm.out_act);
      }
      __syncthreads();
    }
    const float* a_last = NL == 3 ? sm + L::a2 : sm + L::a1;
    // last layer: dW3 += dZ3 . A_last^T ; db3 ; dA_last = W3^T-form
    gemm_nt_acc<OUT, W>(sm + L::dza, a_last, dw3);
    if (threadIdx.x < OUT) {
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += sm[L::dza + threadIdx.x * PS + p];
      db3 += s;
    }
    // dZ_prev[k][p] = relu'(A_last[k][p]) * sum_j W3[j][k] dZ3[j][p]      (B1 = W3 orig [OUT][W])
    gemm_tn<OUT, W, false, false, true>(sm + L::w3, nullptr, sm + L::dza, sm + L::dzb, a_last);
    __syncthreads();
    float* dz1 = sm + L::dzb;  // gradient at the output of layer 1 (pre-ReLU)
    if constexpr (NL == 3) {
      gemm_nt_acc<W, W>(sm + L::dzb, sm + L::a1, dw2);
      if (threadIdx.x < W) {
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += sm[L::dzb + threadIdx.x * PS + p];
        db2 += s;
      }
      gemm_tn<W, W, false, false, true>(sm + L::w2, nullptr, sm + L::dzb, sm + L::dza, sm + L::a1);
      __syncthreads();
      dz1 = sm + L::dza;
    }
    gemm_nt_acc<W, IN>(dz1, sm + L::a0, dw1);
    if (threadIdx.x < W) {
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += dz1[threadIdx.x * PS + p];
      db1 += s;
    }
    if constexpr (NEED_DX) {
      float* dxs = (dz1 == sm + L::dza) ? sm + L::dzb : sm + L::dza;
      gemm_tn<W, IN, false, false, false>(sm + L::w1, nullptr, dz1, dxs, nullptr);
      __syncthreads();
      store_tile_t<IN>(dx, row0, rows, prm.in_dim, dxs);
    }
  }
  flush_dw<W, IN>(prm.dw[0], W, prm.in_dim, dw1);
  if constexpr (NL == 3) flush_dw<W, W>(prm.dw[1], W, W, dw2);
  flush_dw<OUT, W>(prm.dw[NL - 1], prm.out_dim, W, dw3);
  if (threadIdx.x < W) atomicAdd(prm.db[0] + threadIdx.x, db1);
  if constexpr (NL == 3) {
    if (threadIdx.x < W) atomicAdd(prm.db[1] + threadIdx.x, db2);
  }
  if (threadIdx.x < prm.out_dim) atomicAdd(prm.db[NL - 1] + threadIdx.x, db3);
}

template <int IN, int W, int OUT, int NL>
static int launch_fwd(const float* x, int64_t N, const MlpParams& prm, float* y, cudaStream_t st) {
  using L = Smem<IN, W, OUT, NL>;
  const size_t smem = (size_t)L::fwd_total * sizeof(float);
  auto k = mlp_fwd_kernel<IN, W, OUT, NL>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t tiles = (N + P - 1) / P;
  const int per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / (smem + 1024)));
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * per_sm);
  k<<<grid, NT, smem, st>>>(x, N, prm, y);
  return check_launch("mlp_fwd_kernel");
}

template <int IN, int W, int OUT, int NL>
static int launch_bwd(const float* x, const float* dy, int64_t N, const MlpParams& prm, float* dx, cudaStream_t st) {
  using L = Smem<IN, W, OUT, NL>;
  const size_t smem = (size_t)L::bwd_total * sizeof(float);
  const int64_t tiles = (N + P - 1) / P;
  const int per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / (smem + 1024)));
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * per_sm);
  if (dx) {
    auto k = mlp_bwd_kernel<IN, W, OUT, NL, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, NT, smem, st>>>(x, dy, N, prm, dx);
  } else {
    auto k = mlp_bwd_kernel<IN, W, OUT, NL, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, NT, smem, st>>>(x, dy, N, prm, dx);
  }
  return check_launch("mlp_bwd_kernel");
}

static int fill_params(MlpParams& prm, int in_dim, int width, int out_dim, int n_layers, const float* const* w,
                       const float* const* b, int out_act) {
  TN_REQUIRE(w && b, TN_EINVAL, "mlp: null weight pointer table");
  TN_REQUIRE(n_layers == 2 || n_layers == 3, TN_EINVAL, "mlp: n_layers=%d not in {2,3}", n_layers);
  TN_REQUIRE(in_dim >= 1 && in_dim <= 64, TN_EINVAL, "mlp: in_dim=%d not in [1,64]", in_dim);
  TN_REQUIRE(width == 16 || width == 64, TN_EINVAL, "mlp: width=%d not in {16,64}", width);
  TN_REQUIRE(out_dim >= 1 && out_dim <= 16, TN_EINVAL, "mlp: out_dim=%d not in [1,16]", out_dim);
  TN_REQUIRE(out_act >= 0 && out_act <= 2, TN_EINVAL, "mlp: out_act=%d", out_act);
  for (int i = 0; i < n_layers; ++i) {
    TN_REQUIRE(w[i] && b[i], TN_EINVAL, "mlp: null weight/bias for layer %d", i);
    prm.w[i] = w[i];
    prm.b[i] = b[i];
  }
  prm.in_dim = in_dim; prm.out_dim = out_dim; prm.n_layers = n_layers; prm.out_act = out_act;
  return TN_OK;
}

}  // namespace tn

using namespace tn;

#define TN_DISPATCH(FN, ...)                                                       \
  do {                                                                             \
    const int inp = in_dim <= 16 ? 16 : (in_dim <= 32 ? 32 : 64);                  \
    const int outp = out_dim <= 4 ? 4 : 16;                                        \
    const int key = (inp << 16) | (width << 8) | (outp << 2) | n_layers;           \
    switch (key) {                                                                 \
      case (16 << 16) | (16 << 8) | (4 << 2) | 2: return FN<16, 16, 4, 2>(__VA_ARGS__);   \
      case (16 << 16) | (16 << 8) | (4 << 2) | 3: return FN<16, 16, 4, 3>(__VA_ARGS__);   \
      case (16 << 16) | (16 << 8) | (16 << 2) | 2: return FN<16, 16, 16, 2>(__VA_ARGS__); \
      case (16 << 16) | (64 << 8) | (4 << 2) | 2: return FN<16, 64, 4, 2>(__VA_ARGS__);   \
      case (16 << 16) | (64 << 8) | (16 << 2) | 2: return FN<16, 64, 16, 2>(__VA_ARGS__); \
      case (32 << 16) | (16 << 8) | (4 << 2) | 2: return FN<32, 16, 4, 2>(__VA_ARGS__);   \
      case (32 << 16) | (64 << 8) | (4 << 2) | 2: return FN<32, 64, 4, 2>(__VA_ARGS__);   \
      case (32 << 16) | (64 << 8) | (16 << 2) | 2: return FN<32, 64, 16, 2>(__VA_ARGS__); \
      case (32 << 16) | (64 << 8) | (16 << 2) | 3: return FN<32, 64, 16, 3>(__VA_ARGS__); \
      case (64 << 16) | (64 << 8) | (4 << 2) | 2: return FN<64, 64, 4, 2>(__VA_ARGS__);   \
      case (64 << 16) | (64 << 8) | (4 << 2) | 3: return FN<64, 64, 4, 3>(__VA_ARGS__);   \
      case (64 << 16) | (64 << 8) | (16 << 2) | 2: return FN<64, 64, 16, 2>(__VA_ARGS__); \
      case (64 << 16) | (64 << 8) | (16 << 2) | 3: return FN<64, 64, 16, 3>(__VA_ARGS__); \
      default:                                                                     \
        tn::set_error("mlp: shape in=%d width=%d out=%d layers=%d has no kernel instance", in_dim, width, \
                      out_dim, n_layers);                                          \
        return TN_EINVAL;                                                          \
    }                                                                              \
  } while (0)

extern "C" int tn_mlp_fwd(const float* x, int64_t N, int in_dim, int width, int out_dim, int n_layers,
                          const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act, float* y,
                          void* stream) {
  MlpParams prm = {};
  int rc = fill_params(prm, in_dim, width, out_dim, n_layers, w_host_ptrs, b_host_ptrs, out_act);
  if (rc) return rc;
  if (N == 0) return TN_OK;
  TN_REQUIRE(x && y && N > 0, TN_EINVAL, "mlp_fwd: bad x/y/N");
  cudaStream_t st = (cudaStream_t)stream;
  TN_DISPATCH(launch_fwd, x, N, prm, y, st);
}

extern "C" int tn_mlp_bwd(const float* x, const float* dy, int64_t N, int in_dim, int width, int out_dim, int n_layers,
                          const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act, float* dx,
                          float* const* dw_host_ptrs, float* const* db_host_ptrs, void* stream) {
  MlpParams prm = {};
  int rc = fill_params(prm, in_dim, width, out_dim, n_layers, w_host_ptrs, b_host_ptrs, out_act);
  if (rc) return rc;
  if (N == 0) return TN_OK;
  TN_REQUIRE(x && dy && N > 0 && dw_host_ptrs && db_host_ptrs, TN_EINVAL, "mlp_bwd: bad x/dy/N/grad tables");
  for (int i = 0; i < n_layers; ++i) {
    TN_REQUIRE(dw_host_ptrs[i] && db_host_ptrs[i], TN_EINVAL, "mlp_bwd: null grad pointer for layer %d", i);
    prm.dw[i] = dw_host_ptrs[i];
    prm.db[i] = db_host_ptrs[i];
  }
  if (N == 0) return TN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TN_DISPATCH(launch_bwd, x, dy, N, prm, dx, st);
}
