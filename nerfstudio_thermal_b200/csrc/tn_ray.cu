// Warp-per-ray kernels: sample weights (alpha compositing), renderer reductions, inverse-CDF resampling.
//
// One warp owns one ray; samples are visited in rounds of 32 consecutive samples (coalesced 128-byte
// loads), the transmittance / CDF prefix sums are shuffle scans with a carry between rounds.  Prefix sums
// accumulate in double: the reference's CPU cumsum does the same (at::acc_type<float> is double), and the
// searchsorted() steps downstream (PDF resampling, median depth) are discontinuous in the sums.
// Compiled with -fmad=false (each product/sum rounds like the reference's separate torch ops).
#include <math_constants.h>

#include "tn_common.cuh"

namespace tn {

constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ RaySamples.get_weights (rays.py:128-150)
__global__ void __launch_bounds__(32 * kWarpsPerCta) weights_fwd_kernel(const float* __restrict__ sigma,
                                                                       const float* __restrict__ deltas, int64_t R,
                                                                       int S, float* __restrict__ w) {
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  double carry = 0.0;
  // the next round's operands are requested before the current round's scan: the scan's shuffle chain would
  // otherwise sit between two dependent trips to memory
  float dl_n = lane < S ? __ldg(deltas + r * S + lane) : 0.f, sg_n = lane < S ? __ldg(sigma + r * S + lane) : 0.f;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const float dd = dl_n * sg_n;
    const int sn = s + 32;
    dl_n = sn < S ? __ldg(deltas + r * S + sn) : 0.f;
    sg_n = sn < S ? __ldg(sigma + r * S + sn) : 0.f;
    const double incl = warp_incl_scan((double)dd, lane) + carry;
    // exclusive cumsum (shifted inclusive scan: subtracting dd back would turn inf into nan), rounded to
    // float like torch.cumsum's output, then exp(-.)
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    const float excl = (float)(lane == 0 ? carry : prev);
    if (s < S) {
      const float alpha = 1.f - expf(-dd);
      const float trans = expf(-excl);
      w[r * S + s] = nan_to_num(alpha * trans);
    }
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
}

// dL/dsigma_i = delta_i * ( g_i (1-alpha_i) T_i  -  sum_{j>i} g_j w_j )
__global__ void __launch_bounds__(32 * kWarpsPerCta) weights_bwd_kernel(const float* __restrict__ sigma,
                                                                       const float* __restrict__ deltas,
                                                                       const float* __restrict__ dw, int64_t R, int S,
                                                                       float* __restrict__ dsigma) {
  extern __shared__ float smem[];
  float* tr = smem + (size_t)(threadIdx.x >> 5) * S;  // transmittance per sample of this warp's ray
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  double carry = 0.0;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const float dd = s < S ? __ldg(deltas + r * S + s) * __ldg(sigma + r * S + s) : 0.f;
    const double incl = warp_incl_scan((double)dd, lane) + carry;
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (s < S) tr[s] = expf(-(float)(lane == 0 ? carry : prev));
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  double suffix = 0.0;  // sum over samples after the current round
  const int last_base = ((S - 1) / 32) * 32;
  for (int base = last_base; base >= 0; base -= 32) {
    const int s = base + lane;
    float gw = 0.f, one_minus_alpha = 0.f, t = 0.f, dl = 0.f, g = 0.f;
    if (s < S) {
      dl = __ldg(deltas + r * S + s);
      const float dd = dl * __ldg(sigma + r * S + s);
      one_minus_alpha = expf(-dd);
      t = tr[s];
      const float wv = (1.f - one_minus_alpha) * t;
      g = __ldg(dw + r * S + s);
      if (!isfinite(wv)) g = 0.f;  // nan_to_num passes no gradient where it replaced the value
      gw = g * wv;
    }
    // reverse inclusive scan inside the round
    double v = (double)gw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_down_sync(0xffffffffu, v, o);
      if (lane + o < 32) v += u;
    }
    const double after = v - (double)gw + suffix;  // sum_{j>s} g_j w_j
    if (s < S) dsigma[r * S + s] = dl * (g * one_minus_alpha * t - (float)after);
    suffix += __shfl_sync(0xffffffffu, v, 0);
  }
}

// ------------------------------------------------------------------ renderers (renderers.py)
template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta) render_fwd_kernel(
    const float* __restrict__ w, const float* __restrict__ col, const float* __restrict__ starts,
    const float* __restrict__ ends, int64_t R, int S, int TS, int bg_mode, float4 bg, int eval_mode,
    float* __restrict__ rgb_out,
    float* __restrict__ acc_out, float* __restrict__ med_out, float* __restrict__ exp_out, float* __restrict__ minmax) {
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float comp[C > 0 ? C : 1];
#pragma unroll
  for (int c = 0; c < C; ++c) comp[c] = 0.f;
  float acc = 0.f, num = 0.f, smin = CUDART_INF_F, smax = -CUDART_INF_F;
  double carry = 0.0;
  int below_half = 0;  // number of cumulative weights < 0.5 == searchsorted(cum, 0.5, left)
  const bool need_steps = starts != nullptr;
  // operands of the next round are requested before the current round's reductions and scan
  float w_n = lane < S ? __ldg(w + r * S + lane) : 0.f;
  float st_n = (need_steps && lane < S) ? __ldg(starts + r * TS + lane) : 0.f;
  float en_n = (need_steps && lane < S) ? __ldg(ends + r * TS + lane) : 0.f;
  float c_n[C > 0 ? C : 1];
#pragma unroll
  for (int c = 0; c < C; ++c) c_n[c] = lane < S ? __ldg(col + (r * S + lane) * C + c) : 0.f;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const bool ok = s < S;
    const float wv = w_n, st = st_n, en = en_n;
    float cv[C > 0 ? C : 1];
#pragma unroll
    for (int c = 0; c < C; ++c) cv[c] = c_n[c];
    const int sn = s + 32;
    const bool okn = sn < S;
    w_n = okn ? __ldg(w + r * S + sn) : 0.f;
    if (need_steps) {
      st_n = okn ? __ldg(starts + r * TS + sn) : 0.f;
      en_n = okn ? __ldg(ends + r * TS + sn) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) c_n[c] = okn ? __ldg(col + (r * S + sn) * C + c) : 0.f;
    if (C > 0 && ok) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v = cv[c];
        if (eval_mode) v = nan_to_num(v);
        comp[c] += wv * v;
      }
    }
    acc += wv;
    if (need_steps) {
      const float step = ok ? (st + en) / 2.f : 0.f;
      num += wv * step;
      if (ok) { smin = fminf(smin, step); smax = fmaxf(smax, step); }
      const double incl = warp_incl_scan((double)wv, lane) + carry;
      const unsigned m = __ballot_sync(0xffffffffu, ok && ((float)incl < 0.5f));
      below_half += __popc(m);
      carry = __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  acc = warp_sum(acc);
#pragma unroll
  for (int c = 0; c < C; ++c) comp[c] = warp_sum(comp[c]);
  if (need_steps) {
    num = warp_sum(num);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
  }
  if (lane == 0) {
    if (C > 0 && rgb_out) {
      const float bgv[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v = comp[c];
        if (bg_mode == 1) {
          float last = __ldg(col + (r * S + S - 1) * C + c);
          if (eval_mode) last = nan_to_num(last);
          v = v + last * (1.f - acc);
        } else if (bg_mode == 2) {
          v = v + bgv[c] * (1.f - acc);
        }
        if (eval_mode) v = fminf(fmaxf(v, 0.f), 1.f);
        rgb_out[r * C + c] = v;
      }
    }
    if (acc_out) acc_out[r] = acc;
    if (need_steps) {
      if (med_out) {
        const int idx = min(below_half, S - 1);
        med_out[r] = (__ldg(starts + r * TS + idx) + __ldg(ends + r * TS + idx)) / 2.f;
      }
      if (exp_out) exp_out[r] = num / (acc + 1e-10f);
      if (minmax) {
        // Launch-wide extrema: every ray would hit the same two words.  Read them first (they only ever move
        // outwards, so a stale value just means one redundant atomic) and touch them only when this ray improves
        // them -- a handful of rays per launch instead of all of them.
        const float cur_min = *reinterpret_cast<volatile float*>(minmax);
        const float cur_max = *reinterpret_cast<volatile float*>(minmax + 1);
        if (smin < cur_min) {  // steps are >= 0 in practice; the int trick below is valid for any sign
          atomicMin(reinterpret_cast<int*>(minmax), smin >= 0.f ? __float_as_int(smin) : (int)0x80000000);
          if (smin < 0.f) atomicMax(reinterpret_cast<unsigned*>(minmax), __float_as_uint(smin));
        }
        if (smax > cur_max) {
          if (smax >= 0.f) atomicMax(reinterpret_cast<int*>(minmax + 1), __float_as_int(smax));
          else atomicMin(reinterpret_cast<unsigned*>(minmax + 1), __float_as_uint(smax));
        }
      }
    }
  }
}

template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta) render_bwd_kernel(
    const float* __restrict__ w, const float* __restrict__ col, const float* __restrict__ starts,
    const float* __restrict__ ends, const float* __restrict__ d_rgb, const float* __restrict__ d_acc,
    const float* __restrict__ d_depth, int64_t R, int S, int TS, int bg_mode, float4 bg, float* __restrict__ dw,
    float* __restrict__ dcol) {
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float acc = 0.f, num = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float wv = __ldg(w + r * S + s);
    acc += wv;
    if (d_depth) num += wv * ((__ldg(starts + r * TS + s) + __ldg(ends + r * TS + s)) / 2.f);
  }
  acc = warp_sum(acc);
  num = warp_sum(num);
  float g[C > 0 ? C : 1], bgc[C > 0 ? C : 1];
  const float bgv[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
  for (int c = 0; c < C; ++c) {
    g[c] = d_rgb ? __ldg(d_rgb + r * C + c) : 0.f;
    bgc[c] = bg_mode == 1 ? __ldg(col + (r * S + S - 1) * C + c) : (bg_mode == 2 ? bgv[c] : 0.f);
  }
  const float ga = d_acc ? __ldg(d_acc + r) : 0.f;
  const float gd = d_depth ? __ldg(d_depth + r) : 0.f;
  const float den = acc + 1e-10f;
  for (int s = lane; s < S; s += 32) {
    const float wv = __ldg(w + r * S + s);
    float gw = ga;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float cv = __ldg(col + (r * S + s) * C + c);
      gw += g[c] * (cv - bgc[c]);
      float gc = g[c] * wv;
      if (bg_mode == 1 && s == S - 1) gc += g[c] * (1.f - acc);
      if (dcol) dcol[(r * S + s) * C + c] = gc;
    }
    if (d_depth) {
      const float step = (__ldg(starts + r * TS + s) + __ldg(ends + r * TS + s)) / 2.f;
      gw += gd * (step / den - num / (den * den));
    }
    if (dw) dw[r * S + s] = gw;
  }
}

// ------------------------------------------------------------------ PDFSampler (ray_samplers.py:301-372)
__device__ __forceinline__ float spacing_fn(float x) { return x < 1.f ? x / 2.f : 1.f - 1.f / (2.f * x); }
__device__ __forceinline__ float spacing_inv(float y) { return y < 0.5f ? 2.f * y : 1.f / (2.f - 2.f * y); }

__global__ void __launch_bounds__(32 * kWarpsPerCta) pdf_sample_kernel(
    const float* __restrict__ weights, const float* __restrict__ sbins_old, const float* __restrict__ nears,
    const float* __restrict__ fars, const float* __restrict__ u_base, const float* __restrict__ jitter,
    int jitter_per_sample, const float* __restrict__ anneal_dev, int64_t R, int S_old, int S_new, float pad, float eps,
    float* __restrict__ sbins_new, float* __restrict__ ebins_new) {
  extern __shared__ float smem[];
  float* cdf = smem + (size_t)(threadIdx.x >> 5) * 2 * (S_old + 1);
  float* bins = cdf + (S_old + 1);
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  // weights + histogram_padding, their sum, the zero-weight guard (:305-311)
  // proposal-weight annealing (ray_samplers.py:602: weights ** anneal) with the exponent read from the device, so
  // a captured graph follows the schedule; pow(w, 1) == w is taken literally
  const float anneal = anneal_dev ? __ldg(anneal_dev) : 1.f;
  double sum_d = 0.0;
  for (int s = lane; s < S_old; s += 32) {
    float w0 = __ldg(weights + r * S_old + s);
    if (anneal != 1.f) w0 = powf(w0, anneal);
    const float wv = w0 + pad;
    cdf[s + 1] = wv;  // stash
    sum_d += (double)wv;
  }
  for (int s = lane; s <= S_old; s += 32) bins[s] = __ldg(sbins_old + r * (S_old + 1) + s);
  float w_sum = (float)warp_sum_d(sum_d);
  const float padding = fmaxf(eps - w_sum, 0.f);
  const float add = padding / (float)S_old;
  w_sum = w_sum + padding;
  __syncwarp();
  double carry = 0.0;
  for (int base = 0; base < S_old; base += 32) {
    const int s = base + lane;
    const float pdf = s < S_old ? (cdf[s + 1] + add) / w_sum : 0.f;  // :313
    const double incl = warp_incl_scan((double)pdf, lane) + carry;
    if (s < S_old) cdf[s + 1] = fminf(1.f, (float)incl);  // :314
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) cdf[0] = 0.f;  // :315
  __syncwarp();
  const int nb = S_new + 1;
  const float sn = spacing_fn(__ldg(nears + r)), sf = spacing_fn(__ldg(fars + r));
  const float jit_ray = (jitter && !jitter_per_sample) ? __ldg(jitter + r) / (float)nb : 0.f;  // :322
  for (int i = lane; i < nb; i += 32) {
    const float jit = (jitter && jitter_per_sample) ? __ldg(jitter + r * nb + i) / (float)nb : jit_ray;  // :324
    const float u = jitter ? __ldg(u_base + i) + jit : __ldg(u_base + i);  // :325 / :329 (host adds 1/(2 nb))
    // searchsorted(cdf, u, side="right"): first index with cdf[idx] > u
    int lo = 0, hi = S_old + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = min(max(lo - 1, 0), S_old), above = min(max(lo, 0), S_old);
    const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float t = (u - c0) / (c1 - c0);
    if (isnan(t)) t = 0.f;  // nan_to_num(., 0); +-inf are clipped below
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float b = b0 + t * (b1 - b0);  // :354
    sbins_new[r * nb + i] = b;
    ebins_new[r * nb + i] = spacing_inv(b * sf + (1.f - b) * sn);
  }
}

}  // namespace tn

using namespace tn;

static inline unsigned ray_blocks(int64_t R) { return (unsigned)((R + kWarpsPerCta - 1) / kWarpsPerCta); }

extern "C" int tn_weights_fwd(const float* sigma, const float* deltas, int64_t R, int S, float* weights, void* stream) {
  TN_REQUIRE(sigma && deltas && weights, TN_EINVAL, "weights_fwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "weights_fwd: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  weights_fwd_kernel<<<ray_blocks(R), 32 * kWarpsPerCta, 0, (cudaStream_t)stream>>>(sigma, deltas, R, S, weights);
  return check_launch("weights_fwd_kernel");
}

extern "C" int tn_weights_bwd(const float* sigma, const float* deltas, const float* dw, int64_t R, int S, float* dsigma,
                              void* stream) {
  TN_REQUIRE(sigma && deltas && dw && dsigma, TN_EINVAL, "weights_bwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1 && S <= 8192, TN_EINVAL, "weights_bwd: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  const size_t smem = (size_t)kWarpsPerCta * S * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(weights_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  weights_bwd_kernel<<<ray_blocks(R), 32 * kWarpsPerCta, smem, (cudaStream_t)stream>>>(sigma, deltas, dw, R, S, dsigma);
  return check_launch("weights_bwd_kernel");
}

extern "C" int tn_render_fwd(const float* weights, const float* colour, const float* starts, const float* ends,
                             int64_t R, int S, int t_stride, int C, int bg_mode, const float* bg_host, int eval_mode,
                             float* rgb_out,
                             float* acc_out, float* depth_median_out, float* depth_expected_out,
                             float* steps_minmax_out, void* stream) {
  TN_REQUIRE(weights, TN_EINVAL, "render_fwd: weights is null");
  TN_REQUIRE(R >= 0 && S >= 1 && C >= 0 && C <= 4, TN_EINVAL, "render_fwd: bad R=%lld S=%d C=%d", (long long)R, S, C);
  TN_REQUIRE(C == 0 || colour, TN_EINVAL, "render_fwd: colour is null");
  TN_REQUIRE((starts == nullptr) == (ends == nullptr), TN_EINVAL, "render_fwd: starts/ends must both be given");
  TN_REQUIRE(starts || !(depth_median_out || depth_expected_out || steps_minmax_out), TN_EINVAL,
             "render_fwd: depth outputs need starts/ends");
  TN_REQUIRE(bg_mode >= 0 && bg_mode <= 2 && (bg_mode != 2 || bg_host), TN_EINVAL, "render_fwd: bad bg_mode");
  TN_REQUIRE(t_stride == 0 || t_stride >= S, TN_EINVAL, "render_fwd: t_stride=%d < S=%d", t_stride, S);
  const int TS = t_stride ? t_stride : S;
  if (R == 0) return TN_OK;
  float4 bg = make_float4(0, 0, 0, 0);
  if (bg_mode == 2) {
    float t[4] = {0, 0, 0, 0};
    for (int c = 0; c < C; ++c) t[c] = bg_host[c];
    bg = make_float4(t[0], t[1], t[2], t[3]);
  }
  cudaStream_t st = (cudaStream_t)stream;
#define TN_RF(CC)                                                                                                     \
  render_fwd_kernel<CC><<<ray_blocks(R), 32 * kWarpsPerCta, 0, st>>>(weights, colour, starts, ends, R, S, TS, bg_mode, bg, \
                                                                    eval_mode, rgb_out, acc_out, depth_median_out,     \
                                                                    depth_expected_out, steps_minmax_out)
  switch (C) {
    case 0: TN_RF(0); break;
    case 1: TN_RF(1); break;
    case 2: TN_RF(2); break;
    case 3: TN_RF(3); break;
    default: TN_RF(4); break;
  }
#undef TN_RF
  return check_launch("render_fwd_kernel");
}

extern "C" int tn_render_bwd(const float* weights, const float* colour, const float* starts, const float* ends,
                             const float* d_rgb, const float* d_acc, const float* d_depth, int64_t R, int S,
                             int t_stride, int C, int bg_mode, const float* bg_host, float* dweights, float* dcolour, void* stream) {
  TN_REQUIRE(weights, TN_EINVAL, "render_bwd: weights is null");
  TN_REQUIRE(R >= 0 && S >= 1 && C >= 0 && C <= 4, TN_EINVAL, "render_bwd: bad R=%lld S=%d C=%d", (long long)R, S, C);
  TN_REQUIRE(C == 0 || colour, TN_EINVAL, "render_bwd: colour is null");
  TN_REQUIRE(!d_depth || (starts && ends), TN_EINVAL, "render_bwd: d_depth needs starts/ends");
  TN_REQUIRE(bg_mode >= 0 && bg_mode <= 2 && (bg_mode != 2 || bg_host), TN_EINVAL, "render_bwd: bad bg_mode");
  TN_REQUIRE(t_stride == 0 || t_stride >= S, TN_EINVAL, "render_bwd: t_stride=%d < S=%d", t_stride, S);
  const int TS = t_stride ? t_stride : S;
  if (R == 0) return TN_OK;
  float4 bg = make_float4(0, 0, 0, 0);
  if (bg_mode == 2) {
    float t[4] = {0, 0, 0, 0};
    for (int c = 0; c < C; ++c) t[c] = bg_host[c];
    bg = make_float4(t[0], t[1], t[2], t[3]);
  }
  cudaStream_t st = (cudaStream_t)stream;
#define TN_RB(CC)                                                                                                   \
  render_bwd_kernel<CC><<<ray_blocks(R), 32 * kWarpsPerCta, 0, st>>>(weights, colour, starts, ends, d_rgb, d_acc,    \
                                                                    d_depth, R, S, TS, bg_mode, bg, dweights, dcolour)
  switch (C) {
    case 0: TN_RB(0); break;
    case 1: TN_RB(1); break;
    case 2: TN_RB(2); break;
    case 3: TN_RB(3); break;
    default: TN_RB(4); break;
  }
#undef TN_RB
  return check_launch("render_bwd_kernel");
}

extern "C" int tn_pdf_sample(const float* weights, const float* sbins_old, const float* nears, const float* fars,
                             const float* u_base, const float* jitter, int jitter_per_sample,
                             const float* anneal_dev, int64_t R, int S_old, int S_new, float histogram_padding,
                             float eps, float* sbins_new, float* ebins_new, void* stream) {
  TN_REQUIRE(weights && sbins_old && nears && fars && u_base && sbins_new && ebins_new, TN_EINVAL,
             "pdf_sample: null pointer");
  TN_REQUIRE(R >= 0 && S_old >= 1 && S_old <= 1024 && S_new >= 1, TN_EINVAL, "pdf_sample: bad R=%lld S_old=%d S_new=%d",
             (long long)R, S_old, S_new);
  if (R == 0) return TN_OK;
  const size_t smem = (size_t)kWarpsPerCta * 2 * (S_old + 1) * sizeof(float);
  pdf_sample_kernel<<<ray_blocks(R), 32 * kWarpsPerCta, smem, (cudaStream_t)stream>>>(
      weights, sbins_old, nears, fars, u_base, jitter, jitter_per_sample, anneal_dev, R, S_old, S_new, histogram_padding, eps,
      sbins_new, ebins_new);
  return check_launch("pdf_sample_kernel");
}
