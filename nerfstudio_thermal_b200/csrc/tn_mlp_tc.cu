// Fully fused small MLPs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One CTA = 256 threads = one 128-point tile (UMMA M = 128).  Thread t works on point (t & 127), which is TMEM
// lane (t & 127); the two threads of a point (warps w and w+4 address the same 32 TMEM lanes) split the COLUMNS of
// every epilogue (bias, ReLU, activation, re-quantisation for the next layer) between them, so a tile keeps eight
// warps busy.  Hidden activations go TMEM -> registers -> shared memory (as the next layer's A operand, written
// over the previous layer's operand: it is dead once the layer's MMAs have committed) and never reach HBM.  The
// next tile's inputs are fetched into registers while the current tile's MMAs run.
//
// Precision.  Operands are sums of bf16 terms and a product is a few MMAs accumulated in fp32 by the tensor core:
//   * forward: THREE terms per operand (x = x1+x2+x3, 24 mantissa bits), six MMAs per product (all term pairs
//     down to 2^-16): pre-activations are fp32-faithful, so the ReLU masks agree with the fp32 reference (a
//     2^-16 forward flips enough units at the kink to push gradients past the 1e-3 parity bar).  The forward
//     optionally stores the ReLU masks (W bits per point and hidden layer);
//   * backward: two terms, three MMAs (hi*hi + lo*hi + hi*lo, ~2e-5 relative) for the recomputed activations,
//     dH = dZ.W and dW^T += A^T.dZ; ReLU gating uses the forward's masks.
// The kernels are bound by instruction issue and latency, not by the tensor pipe: what the tensor cores buy is
// that the layer products cost no SIMT issue slots and no shared-memory operand traffic.
//
// The reference math: field_components/mlp.py:159-178 (nn.Linear stack, ReLU, optional Sigmoid).
#include "tn_tc.cuh"

namespace tn {

using namespace tc;

constexpr int TP = 128;      // points per tile == UMMA M
constexpr int NTH = 256;     // threads per CTA: two per point (column halves)
constexpr int OUTP = 16;     // padded width of the output layer (UMMA N must be a multiple of 16 at M = 128)
constexpr int CH = TP * 16;  // bytes of one chunk column of an activation tile (8 features x 128 rows)

struct TcParams {
  const float* w[3];
  const float* b[3];
  float* dw[3];
  float* db[3];
  int in_dim, out_dim, out_act, x_stride;
  const float* row_mul;  // optional [N]: every output row is multiplied by out_scale * row_mul[p] after the activation
  float out_scale;       // (density = average_init_density * trunc_exp(z) * selector as the epilogue of the density MLP)
};

// x = t[0] + t[1] + ... as bf16 terms.  Every term but the last of a two-term split is a TRUNCATION (the top 16
// bits of the running remainder: integer ops only, and the subtraction that forms the next remainder is exact);
// three truncated terms hold 8+8+8 = all 24 mantissa bits of an fp32, so the forward split is exact.  The last
// term of the two-term (backward) split is rounded to nearest instead (x ~ 17 bits).
template <int NT>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t (&w)[NT]) {
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    if (k == NT - 1 && NT < 3) {
      const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);  // .x (low half, lower address) = a
      w[k] = *reinterpret_cast<const uint32_t*>(&p);
    } else {
      const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
      w[k] = __byte_perm(ua, ub, 0x7632);  // {hi16(a) in the low half, hi16(b) in the high half}
      if (k + 1 < NT) {
        a -= __uint_as_float(ua & 0xffff0000u);
        b -= __uint_as_float(ub & 0xffff0000u);
      }
    }
  }
}

// 8 consecutive features (one 16-byte chunk) of row r into the NT term tiles of a 128-row operand
template <int NT>
__device__ __forceinline__ void store_chunk_terms(uint8_t* base, int term_stride, int chunk, int r, const float* v) {
  uint32_t w[4][NT];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair<NT>(v[2 * i], v[2 * i + 1], w[i]);
  const size_t off = (size_t)chunk * CH + (size_t)r * 16;
#pragma unroll
  for (int k = 0; k < NT; ++k)
    *reinterpret_cast<uint4*>(base + (size_t)k * term_stride + off) = make_uint4(w[0][k], w[1][k], w[2][k], w[3][k]);
}

// nn.Linear weight [n_real][k_real] (row-major fp32) -> NT operand tiles with NP rows, KP features.
// A thread converts whole 16-byte chunks (8 consecutive input features of one output row): two 16-byte loads when
// the rows allow it, one 16-byte shared-memory store per term (consecutive threads -> consecutive rows, no bank
// conflicts).  All global loads of a matrix are issued before the first is consumed (one L2 round trip per matrix).
template <int NP, int KP, int NT>
__device__ __forceinline__ void load_weight_terms(const float* __restrict__ w, int n_real, int k_real, uint8_t* base) {
  constexpr int CHUNKS = NP * KP / 8;
  constexpr int ITERS = (CHUNKS + NTH - 1) / NTH;
  const bool vec = ((k_real & 3) == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  float v[ITERS][8];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = threadIdx.x + it * NTH;
    const int kc = c / NP, n = c - kc * NP, k0 = kc * 8;
    const float* src = w + (size_t)n * k_real + k0;
    if (c < CHUNKS && n < n_real && vec) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + 4 * h < k_real) q = __ldg(reinterpret_cast<const float4*>(src + 4 * h));
        v[it][4 * h] = q.x; v[it][4 * h + 1] = q.y; v[it][4 * h + 2] = q.z; v[it][4 * h + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[it][i] = (c < CHUNKS && n < n_real && k0 + i < k_real) ? __ldg(src + i) : 0.f;
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = threadIdx.x + it * NTH;
    if (c < CHUNKS) {
      const int kc = c / NP, n = c - kc * NP;
      uint32_t t[4][NT];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_pair<NT>(v[it][2 * i], v[it][2 * i + 1], t[i]);
#pragma unroll
      for (int q = 0; q < NT; ++q)
        *reinterpret_cast<uint4*>(base + (size_t)q * NP * KP * 2 + (size_t)kc * NP * 16 + (size_t)n * 16) =
            make_uint4(t[0][q], t[1][q], t[2][q], t[3][q]);
    }
  }
}

// D[128 x N] = A[128 x K] * B[N x K]^T, both K-major, NT-term operands: all term pairs (i,j) with i+j < NT
template <int N, int K, int NT>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_base, uint32_t a_term, uint32_t b_base) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, N, false, false);
  constexpr uint32_t b_term = N * K * 2;
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < NT - i; ++j)
#pragma unroll
      for (int k = 0; k < K / 16; ++k) {
        // K-major tiles: chunk stride (LBO) = rows*16 bytes, 8-row group stride (SBO) = 128 bytes
        const uint64_t ad = smem_desc(a_base + i * a_term + k * 2 * CH, CH, 128);
        const uint64_t bd = smem_desc(b_base + j * b_term + k * 2 * (N * 16), N * 16, 128);
        mma_bf16(tmem_d, ad, bd, idesc, acc);
        acc = 1;
      }
}

// CNT consecutive columns (from col0) of one row of x[N][stride] into registers; columns >= in_dim read as zero.
// All loads are independent (issued back to back).
template <int CNT>
__device__ __forceinline__ void load_row_part(const float* __restrict__ x, int64_t row, int stride, int in_dim,
                                              bool valid, int col0, float (&v)[CNT]) {
  const float* src = x + row * stride + col0;
  const bool vec = ((stride & 3) == 0) && (col0 + CNT <= stride) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (!valid) {
#pragma unroll
    for (int k = 0; k < CNT; ++k) v[k] = 0.f;
  } else if (vec) {
#pragma unroll
    for (int k = 0; k < CNT; k += 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(src + k));
      v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
    }
    if (col0 + CNT > in_dim) {  // padding columns of the caller's rows: never trusted to be finite
#pragma unroll
      for (int k = 0; k < CNT; ++k)
        if (col0 + k >= in_dim) v[k] = 0.f;
    }
  } else {
#pragma unroll
    for (int k = 0; k < CNT; ++k) v[k] = (col0 + k < in_dim) ? __ldg(src + k) : 0.f;
  }
}

// Head mode input: the colour head's 63-wide row [SH16 | geo15 | appearance32] (fields/nerfacto_field.py:335-344) is
// never materialised -- each thread assembles its 32 columns from the per-ray SH basis, the density-MLP output and
// the per-ray appearance embedding (all whole 16-byte loads).
struct HeadIn {
  const float* sh;   // [R,16]
  const float* h;    // [N,16]: column 0 = raw density, 1..15 = geometry features
  const float* emb;  // [R,32]
  int S;             // samples per ray
};
// columns half*32 .. half*32+31 of the head input of point p; h0 receives h[p,0] (half 0 only)
__device__ __forceinline__ void load_head_part(const HeadIn& g, int64_t p, bool valid, int half, float (&v)[32],
                                               float& h0) {
  h0 = 0.f;
  if (!valid) {
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
    return;
  }
  const int64_t r = p / g.S;
  if (half == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(g.sh + r * 16);
    const float4* h4 = reinterpret_cast<const float4*>(g.h + p * 16);
    float t[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = __ldg(s4 + i), b = __ldg(h4 + i);
      v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
      t[4 * i] = b.x; t[4 * i + 1] = b.y; t[4 * i + 2] = b.z; t[4 * i + 3] = b.w;
    }
    h0 = t[0];
#pragma unroll
    for (int k = 0; k < 15; ++k) v[16 + k] = t[k + 1];
    v[31] = __ldg(g.emb + r * 32);
  } else {
    const float4* e4 = reinterpret_cast<const float4*>(g.emb + r * 32);
    float t[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = __ldg(e4 + i);
      t[4 * i] = a.x; t[4 * i + 1] = a.y; t[4 * i + 2] = a.z; t[4 * i + 3] = a.w;
    }
#pragma unroll
    for (int k = 0; k < 31; ++k) v[k] = t[k + 1];
    v[31] = 0.f;
  }
}

// columns of a W-wide layer handled by one of the two threads of a point
template <int W>
struct ColSplit {
  static_assert(W == 16 || W == 32 || W == 64, "layer width");
  static constexpr int cols = W >= 32 ? W / 2 : W;  // per active thread (multiples of 16: tcgen05.ld.x16)
  __device__ static __forceinline__ bool active(int half) { return W >= 32 || half == 0; }
  __device__ static __forceinline__ int base(int half) { return W >= 32 ? half * cols : 0; }
};

// hidden layer epilogue: TMEM row -> +bias, ReLU -> NT-term operand tile of the next layer (+ optional mask bits).
// mask_words: this point's words of this layer (W/32 of them, or 1), or null.
template <int W, int NT>
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float* __restrict__ bias, uint8_t* h_base,
                                                int h_term, int row, int half, uint32_t* __restrict__ mask_words) {
  using CS = ColSplit<W>;
  static_assert(W != 32, "mask words of a 32-wide layer would be shared by the two column halves");
  if (!CS::active(half)) return;
  const int cb = CS::base(half);
  float v[CS::cols];
#pragma unroll
  for (int c0 = 0; c0 < CS::cols; c0 += 16) tmem_ld16(tmem_row + cb + c0, v + c0);
  tmem_ld_wait();
  uint32_t bits = 0;
#pragma unroll
  for (int c = 0; c < CS::cols / 8; ++c) {
    float u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float z = v[c * 8 + i] + bias[cb + c * 8 + i];
      u[i] = fmaxf(z, 0.f);
      bits |= (z > 0.f ? 1u : 0u) << (c * 8 + i);
    }
    store_chunk_terms<NT>(h_base, h_term, (cb >> 3) + c, row, u);
  }
  if (mask_words) mask_words[W >= 64 ? half : 0] = bits;
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int FT = 3;  // bf16 terms per operand in the forward

template <int IN, int W, int NL>
struct TcSmem {
  static constexpr int KMAX = IN > W ? IN : W;
  static constexpr int w1 = 0;                                   // FT term tiles each
  static constexpr int w2 = w1 + FT * IN * W * 2;
  static constexpr int w3 = w2 + (NL == 3 ? FT * W * W * 2 : 0);
  static constexpr int bias = w3 + FT * W * OUTP * 2;            // fp32: b1[W] b2[W] b3[OUTP]
  static constexpr int act = bias + (2 * W + OUTP) * 4;          // the ONE activation operand, rewritten per layer
  static constexpr int act_term = (KMAX / 8) * CH;
  static constexpr int bar = act + FT * act_term;                // mbarrier (8 B) + tmem slot (4 B)
  static constexpr int total = bar + 16;
  // resident CTAs per SM (shared memory bound); also the register budget the compiler is held to
  static constexpr int per_sm_raw = (224 * 1024) / (total + 1024);
  static constexpr int per_sm = per_sm_raw < 1 ? 1 : (per_sm_raw > 4 ? 4 : per_sm_raw);
};

// forward-side extras of head mode: the density (trunc_exp * selector, fields/nerfacto_field.py:227-228) leaves from
// the same kernel, since the thread that loads h[p, 1:16] holds h[p, 0] anyway
struct HeadFwd {
  HeadIn in;
  const float* sel;    // [N]
  float* density_out;  // [N]
  float scale;
};

template <int IN, int W, int NL, bool HEAD = false>
__global__ void __launch_bounds__(NTH, TcSmem<IN, W, NL>::per_sm) mlp_tc_fwd_kernel(
    const float* __restrict__ x, int64_t N, TcParams prm, float* __restrict__ y, uint32_t* __restrict__ relu_mask,
    HeadFwd hf) {
  static_assert(!HEAD || (IN == 64 && W == 64 && NL == 3), "head mode is the 63-64-64-c colour head");
  extern __shared__ __align__(128) uint8_t sm[];
  using L = TcSmem<IN, W, NL>;
  constexpr int TCOLS = W >= 64 ? 64 : 32;
  constexpr int MW = W >= 32 ? W / 32 : 1;  // mask words per point and hidden layer
  constexpr int XP = IN / 2;                // input columns per thread
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TP - 1), half = tid >> 7;
  float* bias = reinterpret_cast<float*>(sm + L::bias);
  uint8_t* act = sm + L::act;
  const uint32_t bar = smem_u32(sm + L::bar);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L::bar + 8);

  const int64_t tiles = (N + TP - 1) / TP;
  float xr[XP];  // this thread's half of its point's input row: fetched one tile ahead
  float xh0 = 0.f;  // head mode: h[p, 0] of that row
  float xmul = 1.f;  // row multiplier of that row (column half 0 writes the output)
  auto fetch = [&](int64_t t) {
    const int64_t p = t * TP + row;
    if constexpr (HEAD) {
      load_head_part(hf.in, p, p < N, half, xr, xh0);
      if (half == 0) xmul = p < N ? __ldg(hf.sel + p) : 0.f;  // head mode: the selector of the density epilogue
    } else {
      load_row_part<XP>(x, p, prm.x_stride, prm.in_dim, p < N, half * XP, xr);
      if (prm.row_mul && half == 0) xmul = p < N ? __ldg(prm.row_mul + p) : 0.f;
    }
  };
  fetch(blockIdx.x);

  load_weight_terms<W, IN, FT>(prm.w[0], W, prm.in_dim, sm + L::w1);
  if constexpr (NL == 3) load_weight_terms<W, W, FT>(prm.w[1], W, W, sm + L::w2);
  load_weight_terms<OUTP, W, FT>(prm.w[NL - 1], prm.out_dim, W, sm + L::w3);
  for (int i = tid; i < W; i += NTH) {
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
  const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 lanes
  uint32_t phase = 0;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * TP;
    const int rows = (int)min((int64_t)TP, N - row0);
    uint32_t* mrow = (relu_mask && row < rows) ? relu_mask + (row0 + row) * (NL - 1) * MW : nullptr;
    // ---- input tile: split the prefetched half row and write the A0 operand chunks
#pragma unroll
    for (int c = 0; c < XP / 8; ++c) store_chunk_terms<FT>(act, L::act_term, half * (XP / 8) + c, row, xr + c * 8);
    if constexpr (HEAD) {
      if (half == 0 && row < rows)
        hf.density_out[row0 + row] = hf.scale * expf(xh0) * xmul;
    }
    fence_async_smem();
    __syncthreads();  // A0 visible to the async proxy
    // ---- layer 1
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_layer<W, IN, FT>(tmem, smem_u32(act), L::act_term, smem_u32(sm + L::w1));
      mma_commit(bar);
    }
    // this tile's output multiplier, read before the prefetch overwrites it (head mode: xmul is the selector of
    // the density epilogue above, the head's own output is not scaled)
    const float om = HEAD ? prm.out_scale : prm.out_scale * xmul;
    if (t + gridDim.x < tiles) fetch(t + gridDim.x);  // next tile's input rows: in flight while this tile computes
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    hidden_epilogue<W, FT>(tmem_row, bias, act, L::act_term, row, half, mrow);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2 (3-layer networks)
    if constexpr (NL == 3) {
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        issue_layer<W, W, FT>(tmem, smem_u32(act), L::act_term, smem_u32(sm + L::w2));
        mma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      hidden_epilogue<W, FT>(tmem_row, bias + W, act, L::act_term, row, half, mrow ? mrow + MW : nullptr);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    // ---- output layer
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_layer<OUTP, W, FT>(tmem, smem_u32(act), L::act_term, smem_u32(sm + L::w3));
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (half == 0) {
      float v[16];
      tmem_ld16(tmem_row, v);
      tmem_ld_wait();
      if (row < rows) {
        float* dst = y + (row0 + row) * prm.out_dim;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float z = v[j] + bias[2 * W + j];
          if (prm.out_act == 1) z = 1.f / (1.f + expf(-z));
          else if (prm.out_act == 2) z = expf(z);
          v[j] = z * om;
        }
        if (prm.out_dim == 16 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < prm.out_dim) dst[j] = v[j];
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM and the activation operand are reused by the next tile
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem);
  }
}

template <int IN, int W, int NL>
static int launch_tc_fwd(const float* x, int64_t N, const TcParams& prm, float* y, uint32_t* mask, cudaStream_t st) {
  using L = TcSmem<IN, W, NL>;
  static_assert(L::total <= 227 * 1024, "forward tile set does not fit in shared memory");
  auto k = mlp_tc_fwd_kernel<IN, W, NL>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  const int64_t tiles = (N + TP - 1) / TP;
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * L::per_sm);
  k<<<grid, NTH, L::total, st>>>(x, N, prm, y, mask, HeadFwd{});
  return check_launch("mlp_tc_fwd_kernel");
}

static int launch_head_fwd(int64_t N, const TcParams& prm, float* y, uint32_t* mask, const HeadFwd& hf,
                           cudaStream_t st) {
  using L = TcSmem<64, 64, 3>;
  auto k = mlp_tc_fwd_kernel<64, 64, 3, true>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  const int64_t tiles = (N + TP - 1) / TP;
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * L::per_sm);
  k<<<grid, NTH, L::total, st>>>(nullptr, N, prm, y, mask, hf);
  return check_launch("mlp_tc_fwd_kernel(head)");
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
constexpr int BT = 2;  // bf16 terms per operand in the backward

template <int IN, int W, int NL>
struct TcBwdSmem {
  static constexpr int w1 = 0;
  static constexpr int w2 = w1 + BT * IN * W * 2;
  static constexpr int w3 = w2 + (NL == 3 ? BT * W * W * 2 : 0);
  static constexpr int bias = w3 + BT * W * OUTP * 2;
  static constexpr int oh = bias + (2 * W + OUTP) * 4;           // one-hot "ray slot" chunk (head mode), 1 term
  static constexpr int a0 = oh + (IN == 64 && NL == 3 ? CH : 0); // BT terms, (IN/8 + 1) chunks each
  static constexpr int a0_term = (IN / 8 + 1) * CH;
  static constexpr int h1 = a0 + BT * a0_term;                   // BT terms, (W/8 + 1) chunks each
  static constexpr int h_term = (W / 8 + 1) * CH;
  static constexpr int h2 = h1 + BT * h_term;
  // dZ tiles.  Each layer's dW^T MMAs run BEHIND the critical path (they are committed to their own mbarrier and
  // only awaited at the end of the tile), so a dZ tile must stay intact while the next one is written: the output
  // layer's dZ (16 features = 2 chunks) has its own tile, the hidden layers' dZ alternate between two tiles.
  static constexpr int gout = h2 + (NL == 3 ? BT * h_term : 0);  // dZ_out: BT terms of 2 chunks
  static constexpr int gout_term = 2 * CH;
  static constexpr int g2 = gout + BT * gout_term;               // dZ of the FIRST hidden layer of 3-layer networks
  static constexpr int g = g2 + (NL == 3 ? BT * 8 * CH : 0);     // dZ of the last hidden layer: BT terms of 8 chunks
  static constexpr int g_term = 8 * CH;
  static constexpr int g_bytes = BT * g_term + 2048;             // + M=128 over-read slack
  static constexpr int bar = g + g_bytes;                        // 4 mbarriers (critical path, dW3, dW2, dW1) + tmem slot
  static constexpr int total = bar + 64;
  // TMEM columns: forward accumulators / dH / dX, then the three dW^T accumulators
  static constexpr int c_acc = 0;
  static constexpr int acc_cols = (W > IN ? W : IN) > 16 ? (W > IN ? W : IN) : 16;
  static constexpr int c_dw1 = c_acc + acc_cols;
  static constexpr int c_dw2 = c_dw1 + W;
  static constexpr int c_dw3 = c_dw2 + (NL == 3 ? W : 0);
  static constexpr int c_ray = c_dw3 + OUTP;                     // head mode: per-ray sums of dZ1 (lane = ray slot)
  static constexpr int cols_used = c_ray + (IN == 64 && NL == 3 ? W : 0);
  static constexpr int tcols = cols_used <= 32 ? 32 : cols_used <= 64 ? 64 : cols_used <= 128 ? 128
                               : cols_used <= 256 ? 256 : 512;
  // resident CTAs per SM: bounded by shared memory and by the 512 TMEM columns (and held to in registers)
  static constexpr int per_sm_smem = (224 * 1024) / (total + 1024);
  static constexpr int per_sm_tmem = 512 / tcols;
  static constexpr int per_sm_raw = per_sm_smem < per_sm_tmem ? per_sm_smem : per_sm_tmem;
  static constexpr int per_sm = per_sm_raw < 1 ? 1 : (per_sm_raw > 2 ? 2 : per_sm_raw);
};

// D[128 x N] (+)= Act^T-view . dZ^T-view over the 128 points of the tile (both operands MN-major)
template <int N>
__device__ __forceinline__ void issue_dw(uint32_t tmem_d, uint32_t act_base, uint32_t act_term, uint32_t dz_base,
                                         uint32_t dz_term, uint32_t accumulate) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, N, true, true);
  uint32_t acc = accumulate;
#pragma unroll
  for (int i = 0; i < BT; ++i)
#pragma unroll
    for (int j = 0; j < BT - i; ++j)
#pragma unroll
      for (int q = 0; q < TP / 16; ++q) {  // 16 points per MMA
        const uint64_t ad = smem_desc(act_base + i * act_term + q * 256, 128, CH);
        const uint64_t bd = smem_desc(dz_base + j * dz_term + q * 256, 128, CH);
        mma_bf16(tmem_d, ad, bd, idesc, acc);
        acc = 1;
      }
}

// D[128 x N] = OneHot^T-view . dZ^T-view: row s of D = sum of dZ over the tile's points whose ray slot is s.
// The one-hot operand is a single chunk (8 "features" = slots); the MMA's remaining 120 M rows over-read the
// operand tiles that follow it in shared memory (finite values, results unused).
template <int N>
__device__ __forceinline__ void issue_ray_sums(uint32_t tmem_d, uint32_t onehot_base, uint32_t dz_base,
                                               uint32_t dz_term) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, N, true, true);
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < BT; ++j)
#pragma unroll
    for (int q = 0; q < TP / 16; ++q) {
      const uint64_t ad = smem_desc(onehot_base + q * 256, 128, CH);
      const uint64_t bd = smem_desc(dz_base + j * dz_term + q * 256, 128, CH);
      mma_bf16(tmem_d, ad, bd, idesc, acc);
      acc = 1;
    }
}

// D[128 x NIN] = dZ[128 x KOUT] . Wtile  (Wtile has NP rows = out features; NIN of its WIN input features, starting
// at the feature the caller offset w_base to)
template <int NIN, int KOUT, int NP, int WIN = NIN>
__device__ __forceinline__ void issue_dh(uint32_t tmem_d, uint32_t dz_base, uint32_t dz_term, uint32_t w_base) {
  constexpr uint32_t idesc = instr_desc_bf16(TP, NIN, false, true);
  constexpr uint32_t w_term = NP * WIN * 2;
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < BT; ++i)
#pragma unroll
    for (int j = 0; j < BT - i; ++j)
#pragma unroll
      for (int kk = 0; kk < KOUT / 16; ++kk) {
        const uint64_t ad = smem_desc(dz_base + i * dz_term + kk * 2 * CH, CH, 128);        // K-major dZ tile
        const uint64_t bd = smem_desc(w_base + j * w_term + kk * 256, 128, NP * 16);       // W tile, MN-major view
        mma_bf16(tmem_d, ad, bd, idesc, acc);
        acc = 1;
      }
}

__device__ __forceinline__ void store_ones_chunk(uint8_t* base, int term_stride, int chunk, int row) {
  const size_t off = (size_t)chunk * CH + (size_t)row * 16;
  *reinterpret_cast<uint4*>(base + off) = make_uint4(0x00003F80u, 0u, 0u, 0u);  // bf16 {1,0,0,0,0,0,0,0}
#pragma unroll
  for (int t = 1; t < BT; ++t) *reinterpret_cast<uint4*>(base + (size_t)t * term_stride + off) = make_uint4(0, 0, 0, 0);
}

// dZ_prev = dH (TMEM) gated by the ReLU mask of the previous layer -> split -> G tile (this thread's columns)
template <int W>
__device__ __forceinline__ void dh_epilogue(uint32_t tmem_row, bool have_mask, uint32_t word, const uint8_t* h_hi,
                                            uint8_t* g_base, int g_term, int row, int half) {
  using CS = ColSplit<W>;
  if (!CS::active(half)) return;
  const int cb = CS::base(half);
  float v[CS::cols];
#pragma unroll
  for (int c0 = 0; c0 < CS::cols; c0 += 16) tmem_ld16(tmem_row + cb + c0, v + c0);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < CS::cols / 8; ++c) {
    const int chunk = (cb >> 3) + c;
    float u[8];
    if (have_mask) {
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = ((word >> (c * 8 + i)) & 1u) ? v[c * 8 + i] : 0.f;
    } else {  // no saved masks: gate on the recomputed activation
      const uint4 hv = *reinterpret_cast<const uint4*>(h_hi + (size_t)chunk * CH + (size_t)row * 16);
      const __nv_bfloat16* hbv = reinterpret_cast<const __nv_bfloat16*>(&hv);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = __bfloat162float(hbv[i]) > 0.f ? v[c * 8 + i] : 0.f;
    }
    store_chunk_terms<BT>(g_base, g_term, chunk, row, u);
  }
}

// Head mode (the colour head of a NerfactoField, fields/nerfacto_field.py:221-228, 335-348): instead of dX the
// kernel emits what the layers before the head need --
//   dh[p, 0]    = d_density[p] * scale * sel[p] * exp(clamp(h[p,0]))      (trunc_exp * selector backward)
//   dh[p, 1:16] = dX[p, 16:31]                                            (the geometry features' slice of the input)
//   dz1_ray[r]  += sum over the samples p of ray r of dZ1[p]              (first-layer pre-activation gradients; the
//                  appearance-embedding gradient is dz1_ray . W1[:, 31:63], one tiny product per step)
// so neither dX[N,64] nor a separate split/reduce pass touches HBM.
struct HeadIO {
  HeadIn in;               // the head input is re-assembled, not read
  const float* h;          // [N,16] density-MLP output
  const float* sel;        // [N]
  const float* d_density;  // [N] or null
  float* dh;               // [N,16]
  float* dz1_ray;          // [R,64], accumulated
  float scale;
  int S;                   // samples per ray (>= 19: a 128-point tile touches at most 8 rays)
};

template <int IN, int W, int NL, bool NEED_DX, bool HEAD = false>
__global__ void __launch_bounds__(NTH, TcBwdSmem<IN, W, NL>::per_sm) mlp_tc_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ dy, const uint32_t* __restrict__ relu_mask, int64_t N,
    TcParams prm, float* __restrict__ dx, HeadIO hd) {
  static_assert(!HEAD || (IN == 64 && W == 64 && NL == 3), "head mode is the 63-64-64-c colour head");
  extern __shared__ __align__(128) uint8_t sm[];
  using L = TcBwdSmem<IN, W, NL>;
  constexpr int MW = W >= 32 ? W / 32 : 1;
  constexpr int XP = IN / 2;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TP - 1), half = tid >> 7;
  float* bias = reinterpret_cast<float*>(sm + L::bias);
  uint8_t* a0 = sm + L::a0;
  uint8_t* h1 = sm + L::h1;
  uint8_t* h2 = sm + L::h2;
  uint8_t* gb = sm + L::g;                    // dZ of the last hidden layer
  uint8_t* gfirst = NL == 3 ? sm + L::g2 : gb;  // dZ1 (the same tile in 2-layer networks)
  uint8_t* gout = sm + L::gout;
  uint8_t* hlast = NL == 3 ? h2 : h1;
  const uint32_t bar = smem_u32(sm + L::bar);  // critical path; bar + 8 / 16 / 24: the dW^T groups of layers 3 / 2 / 1
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L::bar + 32);

  const int64_t tiles = (N + TP - 1) / TP;
  // operands fetched one tile ahead: this thread's half input row, and (column half 0) the point's dy row
  float xr[XP];
  float dyr[OUTP];
  float xh0 = 0.f;   // head mode: h[p, 0] of the prefetched row
  float xmul = 1.f;  // row multiplier of the prefetched row
  auto prefetch = [&](int64_t t) {
    const int64_t r = t * TP + row;
    if constexpr (HEAD) load_head_part(hd.in, r, r < N, half, xr, xh0);
    else load_row_part<XP>(x, r, prm.x_stride, prm.in_dim, r < N, half * XP, xr);
    if (half == 0) {
      load_row_part<OUTP>(dy, r, prm.out_dim, prm.out_dim, r < N, 0, dyr);
      if (prm.row_mul) xmul = r < N ? __ldg(prm.row_mul + r) : 0.f;
    }
  };
  prefetch(blockIdx.x);

  // zero the whole operand area once: padding chunks / over-read regions must hold finite values
  for (int i = tid; i < (L::bar - L::oh) / 16; i += NTH) reinterpret_cast<uint4*>(sm + L::oh)[i] = make_uint4(0, 0, 0, 0);
  load_weight_terms<W, IN, BT>(prm.w[0], W, prm.in_dim, sm + L::w1);
  if constexpr (NL == 3) load_weight_terms<W, W, BT>(prm.w[1], W, W, sm + L::w2);
  load_weight_terms<OUTP, W, BT>(prm.w[NL - 1], prm.out_dim, W, sm + L::w3);
  for (int i = tid; i < W; i += NTH) {
    bias[i] = __ldg(prm.b[0] + i);
    if constexpr (NL == 3) bias[W + i] = __ldg(prm.b[1] + i);
  }
  if (tid < OUTP) bias[2 * W + tid] = tid < prm.out_dim ? __ldg(prm.b[NL - 1] + tid) : 0.f;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(bar + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<L::tcols>(smem_u32((const void*)tmem_slot));
  __syncthreads();
  // "ones" chunks (bias-gradient rows) are constant across tiles
  if (half == 0) {
    store_ones_chunk(a0, L::a0_term, IN / 8, row);
    store_ones_chunk(h1, L::h_term, W / 8, row);
    if constexpr (NL == 3) store_ones_chunk(h2, L::h_term, W / 8, row);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0, phase_dw = 0;
  uint32_t dw_acc = 0;  // 0 on the first tile (overwrite), 1 afterwards (accumulate)
  // the output layer's pre-activations are only needed for the derivative of its activation
  const bool need_z = prm.out_act != 0;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * TP;
    const int rows = (int)min((int64_t)TP, N - row0);
    if constexpr (HEAD) {
      // A tile whose 128 rows all carry a zero colour gradient needs no product at all: dX, dW and the per-ray sums
      // are exactly zero and only the density backward (column 0 of dh) is left.  This is not a corner case: the RGB
      // loss is masked to the rays of RGB cameras (models/thermal_nerfacto.py:315-318), so in the RGB branch every
      // sample of a thermal camera's ray arrives here with dy == 0 (41-50 % of the tiles of that launch).
      bool nz = false;
      if (half == 0) {
#pragma unroll
        for (int j = 0; j < OUTP; ++j) nz |= dyr[j] != 0.f;
      }
      if (!__syncthreads_or(nz)) {
        if (half == 0 && row < rows) {
          const int64_t p = row0 + row;
          const float sel = __ldg(hd.sel + p);
          const float dd = hd.d_density ? __ldg(hd.d_density + p) : 0.f;
          const float dh0 = dd * hd.scale * sel * expf(fminf(fmaxf(xh0, -15.f), 15.f));
          float4* dst = reinterpret_cast<float4*>(hd.dh + p * 16);
          dst[0] = make_float4(dh0, 0.f, 0.f, 0.f);
          dst[1] = dst[2] = dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (t + gridDim.x < tiles) prefetch(t + gridDim.x);
        continue;  // no barrier phase, no TMEM accumulator and no operand tile was touched
      }
    }
    // this thread's word (its 32 columns) of the ReLU masks of hidden layer 1 and of the last hidden layer
    uint32_t m_first = 0, m_last = 0;
    const bool have_mask = relu_mask != nullptr;
    if (have_mask && row < rows) {
      const uint32_t* mp = relu_mask + (row0 + row) * (NL - 1) * MW + (W >= 64 ? half : 0);
      m_first = __ldg(mp);
      m_last = __ldg(mp + (NL - 2) * MW);
    }
#pragma unroll
    for (int c = 0; c < XP / 8; ++c) store_chunk_terms<BT>(a0, L::a0_term, half * (XP / 8) + c, row, xr + c * 8);
    float hd_h0 = 0.f, hd_sel = 0.f, hd_dd = 0.f;
    if constexpr (HEAD) {
      if (half == 0) {
        uint4 oh = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows) {
          const int64_t p = row0 + row;
          hd_h0 = xh0;
          hd_sel = __ldg(hd.sel + p);
          hd_dd = hd.d_density ? __ldg(hd.d_density + p) : 0.f;
          const int slot = (int)(p / hd.S - row0 / hd.S);  // 0..7
          const uint32_t one = 0x3F80u << (16 * (slot & 1));  // bf16 1.0 in the slot's half word
          oh.x = (slot >> 1) == 0 ? one : 0u; oh.y = (slot >> 1) == 1 ? one : 0u;
          oh.z = (slot >> 1) == 2 ? one : 0u; oh.w = (slot >> 1) == 3 ? one : 0u;
        }
        *reinterpret_cast<uint4*>(sm + L::oh + (size_t)row * 16) = oh;
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---------------- forward recompute (activation VALUES; gating below uses the forward's masks)
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_layer<W, IN, BT>(tmem + L::c_acc, smem_u32(a0), L::a0_term, smem_u32(sm + L::w1));
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    hidden_epilogue<W, BT>(tmem_row + L::c_acc, bias, h1, L::h_term, row, half, nullptr);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if constexpr (NL == 3) {
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        issue_layer<W, W, BT>(tmem + L::c_acc, smem_u32(h1), L::h_term, smem_u32(sm + L::w2));
        mma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      hidden_epilogue<W, BT>(tmem_row + L::c_acc, bias + W, h2, L::h_term, row, half, nullptr);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    if (need_z) {
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        issue_layer<OUTP, W, BT>(tmem + L::c_acc, smem_u32(hlast), L::h_term, smem_u32(sm + L::w3));
        mma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
    }
    // ---------------- dZ_out = dy * act'(z)  -> its own tile (16 features = 2 chunks)
    if (half == 0) {
      float z[16];
      if (need_z) {
        tmem_ld16(tmem_row + L::c_acc, z);
        tmem_ld_wait();
      }
      float u[16];
      const float om = prm.out_scale * xmul;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float gv = dyr[j] * om;  // zero beyond out_dim and for rows past N
        if (need_z) {
          const float zz = z[j] + bias[2 * W + j];
          if (prm.out_act == 1) {
            const float sg = 1.f / (1.f + expf(-zz));
            gv *= sg * (1.f - sg);
          } else if (prm.out_act == 2) {
            gv *= expf(fminf(fmaxf(zz, -15.f), 15.f));
          }
        }
        u[j] = gv;
      }
      store_chunk_terms<BT>(gout, L::gout_term, 0, row, u);
      store_chunk_terms<BT>(gout, L::gout_term, 1, row, u + 8);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---------------- last layer: dH_last first (the epilogue waits for it), dWlast^T behind it
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_dh<W, OUTP, OUTP>(tmem + L::c_acc, smem_u32(gout), L::gout_term, smem_u32(sm + L::w3));
      mma_commit(bar);
      issue_dw<OUTP>(tmem + L::c_dw3, smem_u32(hlast), L::h_term, smem_u32(gout), L::gout_term, dw_acc);
      mma_commit(bar + 8);
    }
    if (t + gridDim.x < tiles) prefetch(t + gridDim.x);  // xr and dyr are both consumed by now
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    dh_epilogue<W>(tmem_row + L::c_acc, have_mask, m_last, hlast, gb, L::g_term, row, half);  // G = dZ of the last hidden layer
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if constexpr (NL == 3) {
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        issue_dh<W, W, W>(tmem + L::c_acc, smem_u32(gb), L::g_term, smem_u32(sm + L::w2));
        mma_commit(bar);
        issue_dw<W>(tmem + L::c_dw2, smem_u32(h1), L::h_term, smem_u32(gb), L::g_term, dw_acc);
        mma_commit(bar + 16);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      dh_epilogue<W>(tmem_row + L::c_acc, have_mask, m_first, h1, gfirst, L::g_term, row, half);  // dZ1 (second tile)
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    // ---------------- first layer: what the epilogue reads (dX / the head's sums) first, dW1^T behind it
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      if constexpr (HEAD) {
        issue_ray_sums<W>(tmem + L::c_ray, smem_u32(sm + L::oh), smem_u32(gfirst), L::g_term);
        // dX restricted to input features 16..31 (geometry features + the first embedding column)
        issue_dh<16, W, W, IN>(tmem + L::c_acc, smem_u32(gfirst), L::g_term, smem_u32(sm + L::w1) + 2 * (W * 16));
        mma_commit(bar);
      } else if constexpr (NEED_DX) {
        issue_dh<IN, W, W>(tmem + L::c_acc, smem_u32(gfirst), L::g_term, smem_u32(sm + L::w1));
        mma_commit(bar);
      }
      issue_dw<W>(tmem + L::c_dw1, smem_u32(a0), L::a0_term, smem_u32(gfirst), L::g_term, dw_acc);
      mma_commit(bar + 24);
    }
    dw_acc = 1;
    if constexpr (HEAD || NEED_DX) {
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
    }
    if constexpr (HEAD) {
      if (half == 0) {  // dh row = [density backward | dX[16:31]]
        float v[16];
        tmem_ld16(tmem_row + L::c_acc, v);
        tmem_ld_wait();
        if (row < rows) {
          const float dh0 = hd_dd * hd.scale * hd_sel * expf(fminf(fmaxf(hd_h0, -15.f), 15.f));
          float4* dst = reinterpret_cast<float4*>(hd.dh + (row0 + row) * 16);
          dst[0] = make_float4(dh0, v[0], v[1], v[2]);
          dst[1] = make_float4(v[3], v[4], v[5], v[6]);
          dst[2] = make_float4(v[7], v[8], v[9], v[10]);
          dst[3] = make_float4(v[11], v[12], v[13], v[14]);
        }
      }
      if ((warp & 3) == 0) {  // warps 0 and 4 read TMEM lanes 0..31 = ray slots; each takes 32 of the 64 columns
        float r[32];
        tmem_ld16(tmem + L::c_ray + half * 32, r);
        tmem_ld16(tmem + L::c_ray + half * 32 + 16, r + 16);
        tmem_ld_wait();
        const int64_t r0 = row0 / hd.S, r_last = (row0 + rows - 1) / hd.S;
        if (row < 8 && r0 + row <= r_last) {
          float* dst = hd.dz1_ray + (r0 + row) * W + half * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + i, r[i]);
        }
      }
    } else if constexpr (NEED_DX) {
      // dX rows: TMEM -> registers -> this thread's columns of its dx row (whole 64-byte runs)
      using CX = ColSplit<IN>;
      if (CX::active(half)) {
        const int cb = CX::base(half);
        float* dst = dx + (row0 + row) * prm.x_stride + cb;
        const bool vec = (prm.x_stride & 3) == 0 && IN <= prm.x_stride && (reinterpret_cast<uintptr_t>(dx) & 15) == 0;
        float v[CX::cols];
#pragma unroll
        for (int c0 = 0; c0 < CX::cols; c0 += 16) tmem_ld16(tmem_row + L::c_acc + cb + c0, v + c0);
        tmem_ld_wait();
        if (row < rows) {
          if (vec) {
#pragma unroll
            for (int i = 0; i < CX::cols; i += 4)
              *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < CX::cols; ++i)
              if (cb + i < prm.in_dim) dst[i] = v[i];
          }
        }
      }
    }
    // the dW^T groups have read their operand tiles (a0, h1, h2 and the dZ tiles are rewritten by the next tile)
    mbar_wait(bar + 8, phase_dw);
    if constexpr (NL == 3) mbar_wait(bar + 16, phase_dw);
    mbar_wait(bar + 24, phase_dw);
    phase_dw ^= 1;
    tc_fence_before();
    __syncthreads();
  }
  // ---------------- flush dW^T / db accumulators (lane = input feature, ones row = bias gradient)
  if (dw_acc) {
    tc_fence_after();
    using CS = ColSplit<W>;
    if (CS::active(half)) {
      const int cb = CS::base(half);
#pragma unroll 1
      for (int c0 = cb; c0 < cb + CS::cols; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_row + L::c_dw1 + c0, v);
        tmem_ld_wait();
        if (row < prm.in_dim) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(prm.dw[0] + (size_t)(c0 + i) * prm.in_dim + row, v[i]);
        } else if (row == IN) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(prm.db[0] + c0 + i, v[i]);
        }
        if constexpr (NL == 3) {
          tmem_ld16(tmem_row + L::c_dw2 + c0, v);
          tmem_ld_wait();
          if (row < W) {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(prm.dw[1] + (size_t)(c0 + i) * W + row, v[i]);
          } else if (row == W) {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(prm.db[1] + c0 + i, v[i]);
          }
        }
      }
    }
    if (half == 0) {
      float v[16];
      tmem_ld16(tmem_row + L::c_dw3, v);
      tmem_ld_wait();
      if (row < W) {
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (o < prm.out_dim) atomicAdd(prm.dw[NL - 1] + (size_t)o * W + row, v[o]);
      } else if (row == W) {
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (o < prm.out_dim) atomicAdd(prm.db[NL - 1] + o, v[o]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<L::tcols>(tmem);
  }
}

template <int IN, int W, int NL>
static int launch_tc_bwd(const float* x, const float* dy, const uint32_t* mask, int64_t N, const TcParams& prm,
                         float* dx, cudaStream_t st) {
  using L = TcBwdSmem<IN, W, NL>;
  static_assert(L::total <= 227 * 1024, "backward tile set does not fit in shared memory");
  const int64_t tiles = (N + TP - 1) / TP;
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * L::per_sm);
  const HeadIO none = {};
  if (dx) {
    auto k = mlp_tc_bwd_kernel<IN, W, NL, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    k<<<grid, NTH, L::total, st>>>(x, dy, mask, N, prm, dx, none);
  } else {
    auto k = mlp_tc_bwd_kernel<IN, W, NL, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    k<<<grid, NTH, L::total, st>>>(x, dy, mask, N, prm, dx, none);
  }
  return check_launch("mlp_tc_bwd_kernel");
}

static int launch_head_bwd(const float* x, const float* dy, const uint32_t* mask, int64_t N, const TcParams& prm,
                           const HeadIO& hd, cudaStream_t st) {
  using L = TcBwdSmem<64, 64, 3>;
  auto k = mlp_tc_bwd_kernel<64, 64, 3, true, true>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  const int64_t tiles = (N + TP - 1) / TP;
  const unsigned grid = (unsigned)min(tiles, (int64_t)kNumSMs * L::per_sm);
  k<<<grid, NTH, L::total, st>>>(x, dy, mask, N, prm, nullptr, hd);
  return check_launch("mlp_tc_bwd_kernel(head)");
}

static int fill_tc(TcParams& prm, int in_dim, int x_stride, int width, int out_dim, int n_layers,
                   const float* const* w, const float* const* b, int out_act) {
  TN_REQUIRE(w && b, TN_EINVAL, "mlp_tc: null weight pointer table");
  TN_REQUIRE(n_layers == 2 || n_layers == 3, TN_EINVAL, "mlp_tc: n_layers=%d not in {2,3}", n_layers);
  TN_REQUIRE(in_dim >= 1 && in_dim <= 64, TN_EINVAL, "mlp_tc: in_dim=%d not in [1,64]", in_dim);
  TN_REQUIRE(x_stride == 0 || x_stride >= in_dim, TN_EINVAL, "mlp_tc: x_stride=%d < in_dim=%d", x_stride, in_dim);
  TN_REQUIRE(width == 16 || width == 64, TN_EINVAL, "mlp_tc: width=%d not in {16,64}", width);
  TN_REQUIRE(out_dim >= 1 && out_dim <= 16, TN_EINVAL, "mlp_tc: out_dim=%d not in [1,16]", out_dim);
  TN_REQUIRE(out_act >= 0 && out_act <= 2, TN_EINVAL, "mlp_tc: out_act=%d", out_act);
  for (int i = 0; i < n_layers; ++i) {
    TN_REQUIRE(w[i] && b[i], TN_EINVAL, "mlp_tc: null weight/bias for layer %d", i);
    prm.w[i] = w[i];
    prm.b[i] = b[i];
  }
  prm.in_dim = in_dim; prm.out_dim = out_dim; prm.out_act = out_act;
  prm.x_stride = x_stride ? x_stride : in_dim;
  prm.row_mul = nullptr;
  prm.out_scale = 1.f;
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

extern "C" int tn_mlp_tc_fwd(const float* x, int64_t N, int in_dim, int x_stride, int width, int out_dim,
                             int n_layers, const float* const* w_host_ptrs, const float* const* b_host_ptrs,
                             int out_act, const float* row_mul, float out_scale, float* y, uint32_t* relu_mask_out,
                             void* stream) {
  TcParams prm = {};
  int rc = fill_tc(prm, in_dim, x_stride, width, out_dim, n_layers, w_host_ptrs, b_host_ptrs, out_act);
  if (rc) return rc;
  prm.row_mul = row_mul;
  prm.out_scale = out_scale;
  if (N == 0) return TN_OK;
  TN_REQUIRE(x && y && N > 0, TN_EINVAL, "mlp_tc_fwd: bad x/y/N");
  cudaStream_t st = (cudaStream_t)stream;
  TN_TC_DISPATCH(launch_tc_fwd, x, N, prm, y, relu_mask_out, st);
}

extern "C" int tn_mlp_tc_bwd(const float* x, const float* dy, const uint32_t* relu_mask, int64_t N, int in_dim,
                             int x_stride, int width, int out_dim, int n_layers, const float* const* w_host_ptrs,
                             const float* const* b_host_ptrs, int out_act, const float* row_mul, float out_scale,
                             float* dx, float* const* dw_host_ptrs, float* const* db_host_ptrs, void* stream) {
  TcParams prm = {};
  int rc = fill_tc(prm, in_dim, x_stride, width, out_dim, n_layers, w_host_ptrs, b_host_ptrs, out_act);
  if (rc) return rc;
  prm.row_mul = row_mul;
  prm.out_scale = out_scale;
  if (N == 0) return TN_OK;
  TN_REQUIRE(x && dy && N > 0 && dw_host_ptrs && db_host_ptrs, TN_EINVAL, "mlp_tc_bwd: bad x/dy/N/grad tables");
  for (int i = 0; i < n_layers; ++i) {
    TN_REQUIRE(dw_host_ptrs[i] && db_host_ptrs[i], TN_EINVAL, "mlp_tc_bwd: null grad pointer for layer %d", i);
    prm.dw[i] = dw_host_ptrs[i];
    prm.db[i] = db_host_ptrs[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  TN_TC_DISPATCH(launch_tc_bwd, x, dy, relu_mask, N, prm, dx, st);
}

static int fill_head(TcParams& prm, int out_dim, const float* const* w, const float* const* b, int out_act, int64_t R,
                     int S, const float* h, const float* sel, const float* sh, const float* emb_ray) {
  int rc = fill_tc(prm, 63, 64, 64, out_dim, 3, w, b, out_act);
  if (rc) return rc;
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "field_head: bad R=%lld S=%d", (long long)R, S);
  TN_REQUIRE(R == 0 || (h && sel && sh && emb_ray), TN_EINVAL, "field_head: null pointer");
  TN_REQUIRE(aligned(h, 16) && aligned(sh, 16) && aligned(emb_ray, 16), TN_EALIGN,
             "field_head: h / sh / emb_ray must be 16-byte aligned");
  return TN_OK;
}

extern "C" int tn_field_head_fwd(const float* h, const float* sel, const float* sh, const float* emb_ray, int64_t R,
                                 int S, int out_dim, float density_scale, const float* const* w_host_ptrs,
                                 const float* const* b_host_ptrs, int out_act, float* density_out, float* y,
                                 uint32_t* relu_mask_out, void* stream) {
  TcParams prm = {};
  int rc = fill_head(prm, out_dim, w_host_ptrs, b_host_ptrs, out_act, R, S, h, sel, sh, emb_ray);
  if (rc) return rc;
  if (R == 0) return TN_OK;
  TN_REQUIRE(density_out && y, TN_EINVAL, "field_head_fwd: null output");
  HeadFwd hf = {{sh, h, emb_ray, S}, sel, density_out, density_scale};
  return launch_head_fwd(R * (int64_t)S, prm, y, relu_mask_out, hf, (cudaStream_t)stream);
}

extern "C" int tn_field_head_bwd(const float* dy, const uint32_t* relu_mask, const float* h, const float* sel,
                                 const float* sh, const float* emb_ray, const float* d_density, int64_t R, int S,
                                 int out_dim, float density_scale, const float* const* w_host_ptrs,
                                 const float* const* b_host_ptrs, int out_act, float* dh_out, float* dz1_ray,
                                 float* const* dw_host_ptrs, float* const* db_host_ptrs, void* stream) {
  TcParams prm = {};
  int rc = fill_head(prm, out_dim, w_host_ptrs, b_host_ptrs, out_act, R, S, h, sel, sh, emb_ray);
  if (rc) return rc;
  TN_REQUIRE(S >= 19, TN_EINVAL, "field_head_bwd: S=%d < 19 (a 128-point tile must touch at most 8 rays)", S);
  if (R == 0) return TN_OK;
  TN_REQUIRE(dy && dh_out && dz1_ray && dw_host_ptrs && db_host_ptrs, TN_EINVAL, "field_head_bwd: null pointer");
  TN_REQUIRE(aligned(dh_out, 16), TN_EALIGN, "field_head_bwd: dh_out must be 16-byte aligned");
  for (int i = 0; i < 3; ++i) {
    TN_REQUIRE(dw_host_ptrs[i] && db_host_ptrs[i], TN_EINVAL, "field_head_bwd: null grad pointer for layer %d", i);
    prm.dw[i] = dw_host_ptrs[i];
    prm.db[i] = db_host_ptrs[i];
  }
  HeadIO hd = {{sh, h, emb_ray, S}, h, sel, d_density, dh_out, dz1_ray, density_scale, S};
  return launch_head_bwd(nullptr, dy, relu_mask, R * (int64_t)S, prm, hd, (cudaStream_t)stream);
}
