// Fully fused small MLPs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One CTA = 128 threads = one 128-point tile (UMMA M = 128): thread t owns point t of the tile, which is also
// TMEM lane t, so every epilogue (bias, ReLU, activation, re-quantisation for the next layer) is a plain
// thread-per-point loop over the row that `tcgen05.ld.32x32b` hands to that thread.  Hidden activations go
// TMEM -> registers -> shared memory (as the next layer's A operand) and never reach HBM.
//
// Precision: operands are bf16 PAIRS (hi + lo, tn_tc.cuh) and each product is three MMAs (hi*hi + lo*hi +
// hi*lo) accumulated in fp32, i.e. ~16 mantissa bits per factor (relative error ~2e-5) -- this keeps the
// "fp32" parity bar (1e-3) with a wide margin while running on the bf16 tensor pipe; the MLPs are <1 % of the
// step's time either way, the point of the tensor cores here is that hidden layers cost no SIMT issue slots.
//
// The reference math: field_components/mlp.py:159-178 (nn.Linear stack, ReLU, optional Sigmoid).
#include "tn_tc.cuh"

namespace tn {

using namespace tc;

constexpr int TP = 128;      // points per tile == threads per CTA == UMMA M
constexpr int OUTP = 16;     // padded width of the output layer (UMMA N must be a multiple of 16 at M = 128)

struct TcParams {
  const float* w[3];
  const float* b[3];
  float* dw[3];
  float* db[3];
  int in_dim, out_dim, out_act;
};

// nn.Linear weight [n_real][k_real] (row-major fp32) -> hi/lo operand tiles with NP rows, KP features
template <int NP, int KP>
__device__ __forceinline__ void load_weight_split(const float* __restrict__ w, int n_real, int k_real, uint8_t* hi,
                                                  uint8_t* lo) {
  for (int i = threadIdx.x; i < NP * KP; i += TP) {
    const int n = i / KP, k = i - n * KP;
    const float v = (n < n_real && k < k_real) ? __ldg(w + (size_t)n * k_real + k) : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    const size_t off = (size_t)(k >> 3) * NP * 16 + (size_t)n * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo + off) = l;
  }
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T with split operands: 3 * K/16 MMAs, issued by ONE thread
template <int N, int K>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, N, false, false);
  uint32_t acc = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const uint32_t a = (s == 1) ? a_lo : a_hi;
    const uint32_t b = (s == 2) ? b_lo : b_hi;
#pragma unroll
    for (int k = 0; k < K / 16; ++k) {
      // K-major tiles: chunk stride (LBO) = rows*16 bytes, 8-row group stride (SBO) = 128 bytes
      const uint64_t ad = smem_desc(a + k * 2 * (TP * 16), TP * 16, 128);
      const uint64_t bd = smem_desc(b + k * 2 * (N * 16), N * 16, 128);
      mma_bf16(tmem_d, ad, bd, idesc, acc);
      acc = 1;
    }
  }
}

template <int IN, int W, int NL>
struct TcSmem {
  static constexpr int w1 = 0;                                   // hi, then lo
  static constexpr int w2 = w1 + 2 * IN * W * 2;
  static constexpr int w3 = w2 + (NL == 3 ? 2 * W * W * 2 : 0);
  static constexpr int bias = w3 + 2 * W * OUTP * 2;             // fp32: b1[W] b2[W] b3[OUTP]
  static constexpr int a0 = bias + (2 * W + OUTP) * 4;           // A0 hi, lo : IN*256 bytes each
  static constexpr int h = a0 + 2 * IN * TP * 2;                 // H hi, lo  : W*256 bytes each (also the fp32 stage)
  static constexpr int stage_bytes = TP * 65 * 4;                // fp32 [128][in_dim|1] input staging
  static constexpr int h_bytes = (2 * W * TP * 2 > stage_bytes) ? 2 * W * TP * 2 : stage_bytes;
  static constexpr int bar = h + h_bytes;                        // mbarrier (8 B) + tmem slot (4 B)
  static constexpr int total = bar + 16;
};

template <int W>
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float* __restrict__ bias, uint8_t* h_hi,
                                                uint8_t* h_lo, int row) {
#pragma unroll
  for (int c0 = 0; c0 < W; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = fmaxf(v[half * 8 + i] + bias[c0 + half * 8 + i], 0.f);
      store_chunk_split(h_hi, h_lo, TP, (c0 >> 3) + half, row, u);
    }
  }
}

template <int IN, int W, int NL>
__global__ void __launch_bounds__(TP) mlp_tc_fwd_kernel(const float* __restrict__ x, int64_t N, TcParams prm,
                                                        float* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t sm[];
  using L = TcSmem<IN, W, NL>;
  constexpr int TCOLS = W >= 64 ? 64 : 32;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* w1h = sm + L::w1; uint8_t* w1l = w1h + IN * W * 2;
  uint8_t* w2h = sm + L::w2; uint8_t* w2l = w2h + W * W * 2;
  uint8_t* w3h = sm + L::w3; uint8_t* w3l = w3h + W * OUTP * 2;
  float* bias = reinterpret_cast<float*>(sm + L::bias);
  uint8_t* a0h = sm + L::a0; uint8_t* a0l = a0h + IN * TP * 2;
  uint8_t* hh = sm + L::h;   uint8_t* hl = hh + W * TP * 2;
  float* stage = reinterpret_cast<float*>(sm + L::h);
  const uint32_t bar = smem_u32(sm + L::bar);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L::bar + 8);

  load_weight_split<W, IN>(prm.w[0], W, prm.in_dim, w1h, w1l);
  if constexpr (NL == 3) load_weight_split<W, W>(prm.w[1], W, W, w2h, w2l);
  load_weight_split<OUTP, W>(prm.w[NL - 1], prm.out_dim, W, w3h, w3l);
  for (int i = tid; i < W; i += TP) {
    bias[i] = __ldg(prm.b[0] + i);
    if constexpr (NL == 3) bias[W + i] = __ldg(prm.b[1] + i);
  }
  if (tid < OUTP) bias[2 * W + tid] = tid < prm.out_dim ? __ldg(prm.b[NL - 1] + tid) : 0.f;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<TCOLS>(smem_u32((const void*)tmem_slot));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 lanes
  uint32_t phase = 0;
  const int sstride = prm.in_dim | 1;  // odd row stride of the fp32 stage: conflict-free per-thread row reads

  const int64_t tiles = (N + TP - 1) / TP;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * TP;
    const int rows = (int)min((int64_t)TP, N - row0);
    // ---- input tile: coalesced global -> fp32 stage -> per-thread row -> split -> A0 operand tiles
    {
      const float* src = x + row0 * prm.in_dim;
      const int n = rows * prm.in_dim;
      for (int i = tid; i < n; i += TP) {
        const int p = i / prm.in_dim, k = i - p * prm.in_dim;
        stage[p * sstride + k] = __ldg(src + i);
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < IN / 8; ++c) {
      float u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = c * 8 + i;
        u[i] = (tid < rows && k < prm.in_dim) ? stage[tid * sstride + k] : 0.f;
      }
      store_chunk_split(a0h, a0l, TP, c, tid, u);
    }
    fence_async_smem();
    __syncthreads();  // A0 visible to the async proxy; the stage (aliasing H) is free again
    // ---- layer 1
    if (tid == 0) {
      tc_fence_after();
      issue_layer<W, IN>(tmem, smem_u32(a0h), smem_u32(a0l), smem_u32(w1h), smem_u32(w1l));
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    hidden_epilogue<W>(tmem_row, bias, hh, hl, tid);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2 (3-layer networks)
    if constexpr (NL == 3) {
      if (tid == 0) {
        tc_fence_after();
        issue_layer<W, W>(tmem, smem_u32(hh), smem_u32(hl), smem_u32(w2h), smem_u32(w2l));
        mma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      hidden_epilogue<W>(tmem_row, bias + W, hh, hl, tid);  // the MMAs that read H have completed
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    // ---- output layer
    if (tid == 0) {
      tc_fence_after();
      issue_layer<OUTP, W>(tmem, smem_u32(hh), smem_u32(hl), smem_u32(w3h), smem_u32(w3l));
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      float v[16];
      tmem_ld16(tmem_row, v);
      tmem_ld_wait();
      if (tid < rows) {
        float* dst = y + (row0 + tid) * prm.out_dim;
        for (int j = 0; j < prm.out_dim; ++j) {
          float z = v[j] + bias[2 * W + j];
          if (prm.out_act == 1) z = 1.f / (1.f + expf(-z));
          else if (prm.out_act == 2) z = expf(z);
          dst[j] = z;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM and the stage/H region are reused by the next tile
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem);
  }
}

template <int IN, int W, int NL>
static int launch_tc_fwd(const float* x, int64_t N, const TcParams& prm, float* y, cudaStream_t st) {
  using L = TcSmem<IN, W, NL>;
  auto k = mlp_tc_fwd_kernel<IN, W, NL>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  const int64_t tiles = (N + TP - 1) / TP;
  const int per_sm = max(1, min(4, (220 * 1024) / (L::total + 1024)));
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * per_sm);
  k<<<grid, TP, L::total, st>>>(x, N, prm, y);
  return check_launch("mlp_tc_fwd_kernel");
}

// ------------------------------------------------------------------------------------------------ backward
//
// Per 128-point tile the backward kernel recomputes the forward (keeping every layer's activations as operand
// tiles), then runs all gradient products on the tensor cores as well:
//     dH_{l-1} = dZ_l . W_l            A = dZ_l tile (K-major),            B = W_l tile viewed MN-major
//     dW_l^T  += A_{l-1}^T . dZ_l      A = activation tile viewed MN-major, B = dZ_l tile viewed MN-major, K = points
// The dW^T accumulators stay in TMEM across all tiles of the persistent CTA (lane = input feature, column =
// output feature) and are flushed once with atomics.  Each activation tile carries one extra "ones" chunk, so
// the row after the last input feature of dW^T is the bias gradient -- no separate reduction.
template <int IN, int W, int NL>
struct TcBwdSmem {
  static constexpr int CH = 2048;                                // bytes of one 16-byte chunk column (128 rows)
  static constexpr int w1 = 0;
  static constexpr int w2 = w1 + 2 * IN * W * 2;
  static constexpr int w3 = w2 + (NL == 3 ? 2 * W * W * 2 : 0);
  static constexpr int bias = w3 + 2 * W * OUTP * 2;
  static constexpr int a0 = bias + (2 * W + OUTP) * 4;           // hi then lo, (IN/8 + 1) chunks each
  static constexpr int a0_bytes = (IN / 8 + 1) * CH;
  static constexpr int h1 = a0 + 2 * a0_bytes;                   // hi then lo, (W/8 + 1) chunks each
  static constexpr int h_bytes = (W / 8 + 1) * CH;
  static constexpr int h2 = h1 + 2 * h_bytes;
  static constexpr int g = h2 + (NL == 3 ? 2 * h_bytes : 0);     // dZ tile hi then lo (8 chunks each) / fp32 stage
  static constexpr int g_half = 8 * CH;
  static constexpr int g_bytes = 2 * g_half + 2048;              // >= 128*65*4 stage, and M=128 over-read slack
  static constexpr int bar = g + g_bytes;
  static constexpr int total = bar + 16;
};

// D[128 x N] (+)= Act^T-view . dZ^T-view over the 128 points of the tile (both operands MN-major)
template <int N>
__device__ __forceinline__ void issue_dw(uint32_t tmem_d, uint32_t act_hi, uint32_t act_lo, uint32_t dz_hi,
                                         uint32_t dz_lo, uint32_t accumulate) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, N, true, true);
  uint32_t acc = accumulate;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const uint32_t a = (s == 1) ? act_lo : act_hi;
    const uint32_t b = (s == 2) ? dz_lo : dz_hi;
#pragma unroll
    for (int j = 0; j < TP / 16; ++j) {  // 16 points per MMA
      const uint64_t ad = smem_desc(a + j * 256, 128, 2048);
      const uint64_t bd = smem_desc(b + j * 256, 128, 2048);
      mma_bf16(tmem_d, ad, bd, idesc, acc);
      acc = 1;
    }
  }
}

// D[128 x NIN] = dZ[128 x KOUT] . Wtile  (Wtile has NP rows = out features, NIN input features)
template <int NIN, int KOUT, int NP>
__device__ __forceinline__ void issue_dh(uint32_t tmem_d, uint32_t dz_hi, uint32_t dz_lo, uint32_t w_hi, uint32_t w_lo) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, NIN, false, true);
  uint32_t acc = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const uint32_t a = (s == 1) ? dz_lo : dz_hi;
    const uint32_t b = (s == 2) ? w_lo : w_hi;
#pragma unroll
    for (int kk = 0; kk < KOUT / 16; ++kk) {
      const uint64_t ad = smem_desc(a + kk * 2 * 2048, 2048, 128);          // K-major dZ tile
      const uint64_t bd = smem_desc(b + kk * 256, 128, NP * 16);            // W tile, MN-major view
      mma_bf16(tmem_d, ad, bd, idesc, acc);
      acc = 1;
    }
  }
}

__device__ __forceinline__ void store_ones_chunk(uint8_t* hi, uint8_t* lo, int chunk, int row) {
  const size_t off = (size_t)chunk * 2048 + (size_t)row * 16;
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(0x00003F80u, 0u, 0u, 0u);  // bf16 {1,0,0,0,0,0,0,0}
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(0u, 0u, 0u, 0u);
}

// dZ_prev = dH (TMEM) masked by relu'(H_prev) -> split -> G tile
template <int W>
__device__ __forceinline__ void dh_epilogue(uint32_t tmem_row, const uint8_t* h_hi, uint8_t* g_hi, uint8_t* g_lo, int row) {
#pragma unroll
  for (int c0 = 0; c0 < W; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int chunk = (c0 >> 3) + half;
      const uint4 hv = *reinterpret_cast<const uint4*>(h_hi + (size_t)chunk * 2048 + (size_t)row * 16);
      const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&hv);
      float u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = __bfloat162float(hb[i]) > 0.f ? v[half * 8 + i] : 0.f;
      store_chunk_split(g_hi, g_lo, TP, chunk, row, u);
    }
  }
}

template <int IN, int W, int NL, bool NEED_DX>
__global__ void __launch_bounds__(TP) mlp_tc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        int64_t N, TcParams prm, float* __restrict__ dx) {
  extern __shared__ __align__(128) uint8_t sm[];
  using L = TcBwdSmem<IN, W, NL>;
  constexpr int TCOLS = 256;
  constexpr int C_ACC = 0;        // forward accumulators / dH / dX   (64 columns)
  constexpr int C_DW1 = 64;       // dW1^T  [IN(+1) lanes x W cols]
  constexpr int C_DW2 = 128;      // dW2^T  [W(+1) lanes x W cols]
  constexpr int C_DW3 = 192;      // dWlast^T [W(+1) lanes x OUTP cols]
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* w1h = sm + L::w1; uint8_t* w1l = w1h + IN * W * 2;
  uint8_t* w2h = sm + L::w2; uint8_t* w2l = w2h + W * W * 2;
  uint8_t* w3h = sm + L::w3; uint8_t* w3l = w3h + W * OUTP * 2;
  float* bias = reinterpret_cast<float*>(sm + L::bias);
  uint8_t* a0h = sm + L::a0; uint8_t* a0l = a0h + L::a0_bytes;
  uint8_t* h1h = sm + L::h1; uint8_t* h1l = h1h + L::h_bytes;
  uint8_t* h2h = sm + L::h2; uint8_t* h2l = h2h + L::h_bytes;
  uint8_t* gh = sm + L::g;   uint8_t* gl = gh + L::g_half;
  float* stage = reinterpret_cast<float*>(sm + L::g);
  uint8_t* hlast_h = NL == 3 ? h2h : h1h;
  uint8_t* hlast_l = NL == 3 ? h2l : h1l;
  const uint32_t bar = smem_u32(sm + L::bar);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L::bar + 8);

  // zero the whole operand area once: padding chunks / over-read regions must hold finite values
  for (int i = tid; i < (L::bar - L::a0) / 16; i += TP) reinterpret_cast<uint4*>(sm + L::a0)[i] = make_uint4(0, 0, 0, 0);
  load_weight_split<W, IN>(prm.w[0], W, prm.in_dim, w1h, w1l);
  if constexpr (NL == 3) load_weight_split<W, W>(prm.w[1], W, W, w2h, w2l);
  load_weight_split<OUTP, W>(prm.w[NL - 1], prm.out_dim, W, w3h, w3l);
  for (int i = tid; i < W; i += TP) {
    bias[i] = __ldg(prm.b[0] + i);
    if constexpr (NL == 3) bias[W + i] = __ldg(prm.b[1] + i);
  }
  if (tid < OUTP) bias[2 * W + tid] = tid < prm.out_dim ? __ldg(prm.b[NL - 1] + tid) : 0.f;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<TCOLS>(smem_u32((const void*)tmem_slot));
  __syncthreads();
  // "ones" chunks (bias-gradient rows) are constant across tiles
  store_ones_chunk(a0h, a0l, IN / 8, tid);
  store_ones_chunk(h1h, h1l, W / 8, tid);
  if constexpr (NL == 3) store_ones_chunk(h2h, h2l, W / 8, tid);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  uint32_t dw_acc = 0;  // 0 on the first tile (overwrite), 1 afterwards (accumulate)
  const int sstride = prm.in_dim | 1;

  const int64_t tiles = (N + TP - 1) / TP;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * TP;
    const int rows = (int)min((int64_t)TP, N - row0);
    {
      const float* src = x + row0 * prm.in_dim;
      const int n = rows * prm.in_dim;
      for (int i = tid; i < n; i += TP) {
        const int p = i / prm.in_dim, k = i - p * prm.in_dim;
        stage[p * sstride + k] = __ldg(src + i);
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < IN / 8; ++c) {
      float u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = c * 8 + i;
        u[i] = (tid < rows && k < prm.in_dim) ? stage[tid * sstride + k] : 0.f;
      }
      store_chunk_split(a0h, a0l, TP, c, tid, u);
    }
    fence_async_smem();
    __syncthreads();
    // ---------------- forward recompute
    if (tid == 0) {
      tc_fence_after();
      issue_layer<W, IN>(tmem + C_ACC, smem_u32(a0h), smem_u32(a0l), smem_u32(w1h), smem_u32(w1l));
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    hidden_epilogue<W>(tmem_row + C_ACC, bias, h1h, h1l, tid);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if constexpr (NL == 3) {
      if (tid == 0) {
        tc_fence_after();
        issue_layer<W, W>(tmem + C_ACC, smem_u32(h1h), smem_u32(h1l), smem_u32(w2h), smem_u32(w2l));
        mma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      hidden_epilogue<W>(tmem_row + C_ACC, bias + W, h2h, h2l, tid);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    if (tid == 0) {
      tc_fence_after();
      issue_layer<OUTP, W>(tmem + C_ACC, smem_u32(hlast_h), smem_u32(hlast_l), smem_u32(w3h), smem_u32(w3l));
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    // ---------------- dZ_out = dy * act'(z)  -> G tile (16 features = 2 chunks)
    {
      float z[16];
      tmem_ld16(tmem_row + C_ACC, z);
      tmem_ld_wait();
      float g[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float gv = 0.f;
        if (tid < rows && j < prm.out_dim) {
          gv = __ldg(dy + (row0 + tid) * prm.out_dim + j);
          const float zz = z[j] + bias[2 * W + j];
          if (prm.out_act == 1) {
            const float sg = 1.f / (1.f + expf(-zz));
            gv *= sg * (1.f - sg);
          } else if (prm.out_act == 2) {
            gv *= expf(fminf(fmaxf(zz, -15.f), 15.f));
          }
        }
        g[j] = gv;
      }
      float u0[8], u1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { u0[i] = g[i]; u1[i] = g[8 + i]; }
      store_chunk_split(gh, gl, TP, 0, tid, u0);
      store_chunk_split(gh, gl, TP, 1, tid, u1);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---------------- last layer: dWlast^T, dH_last
    if (tid == 0) {
      tc_fence_after();
      issue_dw<OUTP>(tmem + C_DW3, smem_u32(hlast_h), smem_u32(hlast_l), smem_u32(gh), smem_u32(gl), dw_acc);
      issue_dh<W, OUTP, OUTP>(tmem + C_ACC, smem_u32(gh), smem_u32(gl), smem_u32(w3h), smem_u32(w3l));
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    dh_epilogue<W>(tmem_row + C_ACC, hlast_h, gh, gl, tid);  // G now holds dZ of the last hidden layer
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if constexpr (NL == 3) {
      if (tid == 0) {
        tc_fence_after();
        issue_dw<W>(tmem + C_DW2, smem_u32(h1h), smem_u32(h1l), smem_u32(gh), smem_u32(gl), dw_acc);
        issue_dh<W, W, W>(tmem + C_ACC, smem_u32(gh), smem_u32(gl), smem_u32(w2h), smem_u32(w2l));
        mma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      dh_epilogue<W>(tmem_row + C_ACC, h1h, gh, gl, tid);  // dZ1
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    // ---------------- first layer: dW1^T (and dX)
    if (tid == 0) {
      tc_fence_after();
      issue_dw<W>(tmem + C_DW1, smem_u32(a0h), smem_u32(a0l), smem_u32(gh), smem_u32(gl), dw_acc);
      if constexpr (NEED_DX) issue_dh<IN, W, W>(tmem + C_ACC, smem_u32(gh), smem_u32(gl), smem_u32(w1h), smem_u32(w1l));
      mma_commit(bar);
    }
    dw_acc = 1;
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    if constexpr (NEED_DX) {
      // dX rows: TMEM -> fp32 stage (G is free: its readers completed) -> coalesced global store
#pragma unroll
      for (int c0 = 0; c0 < IN; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_row + C_ACC + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < prm.in_dim) stage[tid * sstride + c0 + i] = v[i];
      }
      __syncthreads();
      float* dst = dx + row0 * prm.in_dim;
      const int n = rows * prm.in_dim;
      for (int i = tid; i < n; i += TP) {
        const int p = i / prm.in_dim, k = i - p * prm.in_dim;
        dst[i] = stage[p * sstride + k];
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  // ---------------- flush dW^T / db accumulators (lane = input feature, ones row = bias gradient)
  if (dw_acc) {
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < W; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_row + C_DW1 + c0, v);
      tmem_ld_wait();
      if (tid < prm.in_dim) {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(prm.dw[0] + (size_t)(c0 + i) * prm.in_dim + tid, v[i]);
      } else if (tid == IN) {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(prm.db[0] + c0 + i, v[i]);
      }
      if constexpr (NL == 3) {
        tmem_ld16(tmem_row + C_DW2 + c0, v);
        tmem_ld_wait();
        if (tid < W) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(prm.dw[1] + (size_t)(c0 + i) * W + tid, v[i]);
        } else if (tid == W) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(prm.db[1] + c0 + i, v[i]);
        }
      }
    }
    float v[16];
    tmem_ld16(tmem_row + C_DW3, v);
    tmem_ld_wait();
    if (tid < W) {
      for (int o = 0; o < prm.out_dim; ++o) atomicAdd(prm.dw[NL - 1] + (size_t)o * W + tid, v[o]);
    } else if (tid == W) {
      for (int o = 0; o < prm.out_dim; ++o) atomicAdd(prm.db[NL - 1] + o, v[o]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem);
  }
}

template <int IN, int W, int NL>
static int launch_tc_bwd(const float* x, const float* dy, int64_t N, const TcParams& prm, float* dx, cudaStream_t st) {
  using L = TcBwdSmem<IN, W, NL>;
  static_assert(L::total <= 227 * 1024, "backward tile set does not fit in shared memory");
  const int64_t tiles = (N + TP - 1) / TP;
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs);
  if (dx) {
    auto k = mlp_tc_bwd_kernel<IN, W, NL, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    k<<<grid, TP, L::total, st>>>(x, dy, N, prm, dx);
  } else {
    auto k = mlp_tc_bwd_kernel<IN, W, NL, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    k<<<grid, TP, L::total, st>>>(x, dy, N, prm, dx);
  }
  return check_launch("mlp_tc_bwd_kernel");
}

static int fill_tc(TcParams& prm, int in_dim, int width, int out_dim, int n_layers, const float* const* w,
                   const float* const* b, int out_act) {
  TN_REQUIRE(w && b, TN_EINVAL, "mlp_tc: null weight pointer table");
  TN_REQUIRE(n_layers == 2 || n_layers == 3, TN_EINVAL, "mlp_tc: n_layers=%d not in {2,3}", n_layers);
  TN_REQUIRE(in_dim >= 1 && in_dim <= 64, TN_EINVAL, "mlp_tc: in_dim=%d not in [1,64]", in_dim);
  TN_REQUIRE(width == 16 || width == 64, TN_EINVAL, "mlp_tc: width=%d not in {16,64}", width);
  TN_REQUIRE(out_dim >= 1 && out_dim <= 16, TN_EINVAL, "mlp_tc: out_dim=%d not in [1,16]", out_dim);
  TN_REQUIRE(out_act >= 0 && out_act <= 2, TN_EINVAL, "mlp_tc: out_act=%d", out_act);
  for (int i = 0; i < n_layers; ++i) {
    TN_REQUIRE(w[i] && b[i], TN_EINVAL, "mlp_tc: null weight/bias for layer %d", i);
    prm.w[i] = w[i];
    prm.b[i] = b[i];
  }
  prm.in_dim = in_dim; prm.out_dim = out_dim; prm.out_act = out_act;
  return TN_OK;
}

}  // namespace tn

using namespace tn;

#define TN_TC_DISPATCH(FN, ...)                                                          \
  do {                                                                                   \
    const int inp = in_dim <= 16 ? 16 : (in_dim <= 32 ? 32 : 64);                        \
    const int key = (inp << 16) | (width << 8) | n_layers;                               \
    switch (key) {                                                                       \
      case (16 << 16) | (16 << 8) | 2: return FN<16, 16, 2>(__VA_ARGS__);                \
      case (16 << 16) | (16 << 8) | 3: return FN<16, 16, 3>(__VA_ARGS__);                \
      case (32 << 16) | (16 << 8) | 2: return FN<32, 16, 2>(__VA_ARGS__);                \
      case (16 << 16) | (64 << 8) | 2: return FN<16, 64, 2>(__VA_ARGS__);                \
      case (32 << 16) | (64 << 8) | 2: return FN<32, 64, 2>(__VA_ARGS__);                \
      case (32 << 16) | (64 << 8) | 3: return FN<32, 64, 3>(__VA_ARGS__);                \
      case (64 << 16) | (64 << 8) | 2: return FN<64, 64, 2>(__VA_ARGS__);                \
      case (64 << 16) | (64 << 8) | 3: return FN<64, 64, 3>(__VA_ARGS__);                \
      default:                                                                           \
        tn::set_error("mlp_tc: shape in=%d width=%d layers=%d has no kernel instance", in_dim, width, n_layers); \
        return TN_EINVAL;                                                                \
    }                                                                                    \
  } while (0)

extern "C" int tn_mlp_tc_fwd(const float* x, int64_t N, int in_dim, int width, int out_dim, int n_layers,
                             const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act, float* y,
                             void* stream) {
  TcParams prm = {};
  int rc = fill_tc(prm, in_dim, width, out_dim, n_layers, w_host_ptrs, b_host_ptrs, out_act);
  if (rc) return rc;
  if (N == 0) return TN_OK;
  TN_REQUIRE(x && y && N > 0, TN_EINVAL, "mlp_tc_fwd: bad x/y/N");
  cudaStream_t st = (cudaStream_t)stream;
  TN_TC_DISPATCH(launch_tc_fwd, x, N, prm, y, st);
}

extern "C" int tn_mlp_tc_bwd(const float* x, const float* dy, int64_t N, int in_dim, int width, int out_dim, int n_layers,
                             const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act, float* dx,
                             float* const* dw_host_ptrs, float* const* db_host_ptrs, void* stream) {
  TcParams prm = {};
  int rc = fill_tc(prm, in_dim, width, out_dim, n_layers, w_host_ptrs, b_host_ptrs, out_act);
  if (rc) return rc;
  if (N == 0) return TN_OK;
  TN_REQUIRE(x && dy && N > 0 && dw_host_ptrs && db_host_ptrs, TN_EINVAL, "mlp_tc_bwd: bad x/dy/N/grad tables");
  for (int i = 0; i < n_layers; ++i) {
    TN_REQUIRE(dw_host_ptrs[i] && db_host_ptrs[i], TN_EINVAL, "mlp_tc_bwd: null grad pointer for layer %d", i);
    prm.dw[i] = dw_host_ptrs[i];
    prm.db[i] = db_host_ptrs[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  TN_TC_DISPATCH(launch_tc_bwd, x, dy, N, prm, dx, st);
}
