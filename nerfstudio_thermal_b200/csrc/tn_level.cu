// One launch per sampling level: the warp-per-ray composite + resample kernels of the north star.
//
//   tn_level_resample   proposal level:  density -> get_weights -> (median depth) -> PDF inverse-CDF resampling
//                       (cameras/rays.py:128-150, renderers.py:547-557, ray_samplers.py:301-372, :602)
//   tn_ray_heads_fwd    final level:     density, colour -> get_weights -> RGB/RGBT + accumulation + median and
//                       expected depth (renderers.py:118-133, :509, :547-576) + distortion loss (losses.py:139-158)
//                       + the interlevel loss against every proposal histogram (losses.py:57-135)
//   tn_ray_heads_bwd    everything the ray-level backward has to do, for all levels of a branch, in one launch:
//                       renderer gradients + distortion gradient -> get_weights backward -> d density, d colour;
//                       interlevel gradient -> get_weights backward -> d density of each proposal level
//
// A warp owns a ray; samples are visited in rounds of 32 consecutive samples (coalesced), the transmittance and CDF
// prefix sums are shuffle scans with a double carry between rounds (the reference's CPU cumsum accumulates in
// double, and the searchsorted() steps downstream are discontinuous in the sums).  The arithmetic of every stage is
// that of the single-purpose kernels in tn_ray.cu / tn_fused.cu (same operations in the same order, file compiled
// with -fmad=false), so the fused launches return the same values; what disappears is the weights / loss
// round trips through HBM, 16 of the 24 ray-kernel launches of a train step and the torch glue between them.
#include <math_constants.h>

#include "tn_common.cuh"

namespace tn {

constexpr int kLevelWarps = 4;

__device__ __forceinline__ double scan_incl_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lv_spacing_fn(float x) { return x < 1.f ? x / 2.f : 1.f - 1.f / (2.f * x); }
__device__ __forceinline__ float lv_spacing_inv(float y) { return y < 0.5f ? 2.f * y : 1.f / (2.f - 2.f * y); }

// get_weights of one ray (cameras/rays.py:128-150) from its density row and euclidean bin edges.  Every round hands
// (sample index, weight) to `sink`; returns nothing.  Same operations as weights_fwd_kernel (tn_ray.cu).
template <typename Sink>
__device__ __forceinline__ void ray_weights(const float* __restrict__ sigma_row, const float* __restrict__ eb_row, int S,
                                            int lane, Sink&& sink) {
  double carry = 0.0;
  float e0_n = lane < S ? __ldg(eb_row + lane) : 0.f, e1_n = lane < S ? __ldg(eb_row + lane + 1) : 0.f;
  float sg_n = lane < S ? __ldg(sigma_row + lane) : 0.f;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const float dl = e1_n - e0_n;  // deltas = bin_ends - bin_starts (rays.py:274)
    const float dd = dl * sg_n;
    const int sn = s + 32;
    e0_n = sn < S ? __ldg(eb_row + sn) : 0.f;
    e1_n = sn < S ? __ldg(eb_row + sn + 1) : 0.f;
    sg_n = sn < S ? __ldg(sigma_row + sn) : 0.f;
    const double incl = scan_incl_d((double)(s < S ? dd : 0.f), lane) + carry;
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    const float excl = (float)(lane == 0 ? carry : prev);
    float wv = 0.f;
    if (s < S) {
      const float alpha = 1.f - expf(-dd);
      const float trans = expf(-excl);
      wv = nan_to_num(alpha * trans);
    }
    sink(s, wv);  // all lanes call it (s may be >= S, wv = 0 then)
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
}

// get_weights backward of one ray: gw[s] = dL/dw_s in shared memory -> dsigma row.  tr: S floats of scratch.
// Same operations as weights_bwd_kernel.
__device__ __forceinline__ void ray_weights_bwd(const float* __restrict__ sigma_row, const float* __restrict__ eb_row,
                                                const float* gw_smem, float* tr, int S, int lane,
                                                float* __restrict__ dsigma_row) {
  double carry = 0.0;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const float dd = s < S ? (__ldg(eb_row + s + 1) - __ldg(eb_row + s)) * __ldg(sigma_row + s) : 0.f;
    const double incl = scan_incl_d((double)dd, lane) + carry;
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (s < S) tr[s] = expf(-(float)(lane == 0 ? carry : prev));
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  double suffix = 0.0;
  const int last_base = ((S - 1) / 32) * 32;
  for (int base = last_base; base >= 0; base -= 32) {
    const int s = base + lane;
    float gwv = 0.f, one_minus_alpha = 0.f, t = 0.f, dl = 0.f, g = 0.f;
    if (s < S) {
      dl = __ldg(eb_row + s + 1) - __ldg(eb_row + s);
      const float dd = dl * __ldg(sigma_row + s);
      one_minus_alpha = expf(-dd);
      t = tr[s];
      const float wv = (1.f - one_minus_alpha) * t;
      g = gw_smem[s];
      if (!isfinite(wv)) g = 0.f;
      gwv = g * wv;
    }
    double v = (double)gwv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_down_sync(0xffffffffu, v, o);
      if (lane + o < 32) v += u;
    }
    const double after = v - (double)gwv + suffix;
    if (s < S) dsigma_row[s] = dl * (g * one_minus_alpha * t - (float)after);
    suffix += __shfl_sync(0xffffffffu, v, 0);
  }
}

// ------------------------------------------------------------------------------------------------ proposal level
// smem per warp: cdf[S+1] | bins[S+1]
__global__ void __launch_bounds__(32 * kLevelWarps) level_resample_kernel(
    const float* __restrict__ sigma, const float* __restrict__ ebins, const float* __restrict__ sbins,
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ u_base,
    const float* __restrict__ jitter, int jitter_per_sample, const float* __restrict__ anneal_dev, int64_t R, int S,
    int S_new, float pad, float eps, float* __restrict__ w_out, float* __restrict__ med_out,
    float* __restrict__ sbins_new, float* __restrict__ ebins_new) {
  extern __shared__ float smem[];
  float* cdf = smem + (size_t)(threadIdx.x >> 5) * 2 * (S + 1);
  float* bins = cdf + (S + 1);
  const int64_t r = (int64_t)blockIdx.x * kLevelWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* eb = ebins + r * (S + 1);
  // ---- weights (+ median-depth search, renderers.py:547-557: first cumulative weight >= 0.5)
  double wcarry = 0.0;
  int below_half = 0;
  ray_weights(sigma + r * S, eb, S, lane, [&](int s, float wv) {
    if (s < S) {
      w_out[r * S + s] = wv;
      cdf[s + 1] = wv;  // stash for the histogram below
    }
    if (med_out) {
      const double incl = scan_incl_d((double)wv, lane) + wcarry;
      below_half += __popc(__ballot_sync(0xffffffffu, s < S && ((float)incl < 0.5f)));
      wcarry = __shfl_sync(0xffffffffu, incl, 31);
    }
  });
  if (med_out && lane == 0) {
    const int idx = min(below_half, S - 1);
    med_out[r] = (__ldg(eb + idx) + __ldg(eb + idx + 1)) / 2.f;
  }
  if (!sbins_new) return;  // weights only
  __syncwarp();
  // ---- PDFSampler (same operations as pdf_sample_kernel)
  const float anneal = anneal_dev ? __ldg(anneal_dev) : 1.f;
  double sum = 0.0;
  for (int s = lane; s < S; s += 32) {
    float w0 = cdf[s + 1];
    if (anneal != 1.f) w0 = powf(w0, anneal);
    const float wv = w0 + pad;
    cdf[s + 1] = wv;
    sum += (double)wv;
  }
  for (int s = lane; s <= S; s += 32) bins[s] = __ldg(sbins + r * (S + 1) + s);
  float w_sum = (float)sum_d(sum);
  const float padding = fmaxf(eps - w_sum, 0.f);
  const float add = padding / (float)S;
  w_sum = w_sum + padding;
  __syncwarp();
  double carry = 0.0;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const float pdf = s < S ? (cdf[s + 1] + add) / w_sum : 0.f;
    const double incl = scan_incl_d((double)pdf, lane) + carry;
    if (s < S) cdf[s + 1] = fminf(1.f, (float)incl);
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  const int nb = S_new + 1;
  const float sn = lv_spacing_fn(__ldg(nears + r)), sf = lv_spacing_fn(__ldg(fars + r));
  const float jit_ray = (jitter && !jitter_per_sample) ? __ldg(jitter + r) / (float)nb : 0.f;
  for (int i = lane; i < nb; i += 32) {
    const float jit = (jitter && jitter_per_sample) ? __ldg(jitter + r * nb + i) / (float)nb : jit_ray;
    const float u = jitter ? __ldg(u_base + i) + jit : __ldg(u_base + i);
    int lo = 0, hi = S + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = min(max(lo - 1, 0), S), above = min(max(lo, 0), S);
    const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float t = (u - c0) / (c1 - c0);
    if (isnan(t)) t = 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float b = b0 + t * (b1 - b0);
    sbins_new[r * nb + i] = b;
    ebins_new[r * nb + i] = lv_spacing_inv(b * sf + (1.f - b) * sn);
  }
}

// ------------------------------------------------------------------------------------------------ final level
struct PropHist {  // one proposal histogram of the interlevel loss
  const float* w;      // [R,Sp]
  const float* sbins;  // [R,Sp+1]
  float* dw;           // [R,Sp] out: d(loss of this ray) / d w (NULL: not wanted)
  int Sp;
};
struct HeadsFwdArgs {
  const float* sigma;  // [R,S]
  const float* col;    // [R,S,C]
  const float* ebins;  // [R,S+1]
  const float* sbins;  // [R,S+1]
  int64_t R;
  int S, bg_mode, eval_mode, n_prop, smem_floats_per_warp;
  float4 bg;
  PropHist prop[2];
  float* w_out;      // [R,S]
  float* rgb_out;    // [R,C]
  float* acc_out;    // [R]
  float* med_out;    // [R]
  float* exp_out;    // [R] (before the batch-global clip)
  float* minmax;     // [2]
  float* loss_acc;   // [2]: += sum over rays of the distortion loss, += sum of the interlevel loss terms (NULL: skip)
  float* dw_dist;    // [R,S] d(distortion of this ray) / d w (NULL: not wanted)
  // self-resetting launch scratch {min, max, dist sum, inter sum, ticket}: when given, the atomics go there and the
  // CTA that finishes last writes the FINAL minmax / scaled losses / clipped expected depth and re-arms the scratch
  float* scratch;
  float* exp_clip_out;  // [R] expected depth after the launch-wide clip (renderers.py:574)
  float scale_dist, scale_inter;
};

// smem per warp: sw[S] | su[S] | (interlevel) cy[Sp+1] | edges[Sp+1] | diff[Sp+1]
template <int C>
__global__ void __launch_bounds__(32 * kLevelWarps) ray_heads_fwd_kernel(const HeadsFwdArgs a) {
  extern __shared__ float smem[];
  __shared__ float cta_loss[2][kLevelWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sw = smem + (size_t)warp * a.smem_floats_per_warp;
  float* su = sw + a.S;
  float* cy = su + a.S;
  const int64_t r = (int64_t)blockIdx.x * kLevelWarps + warp;
  const int S = a.S;
  float dist_ray = 0.f, inter_ray = 0.f;
  float* const mm = a.minmax ? (a.scratch ? a.scratch : a.minmax) : nullptr;          // atomics' targets
  float* const lacc = a.loss_acc ? (a.scratch ? a.scratch + 2 : a.loss_acc) : nullptr;
  if (r < a.R) {
    const float* eb = a.ebins + r * (S + 1);
    const float* sb = a.sbins + r * (S + 1);
    const float* col = a.col ? a.col + r * S * C : nullptr;
    float comp[C > 0 ? C : 1];
#pragma unroll
    for (int c = 0; c < C; ++c) comp[c] = 0.f;
    float acc = 0.f, num = 0.f, smin = CUDART_INF_F, smax = -CUDART_INF_F;
    double wcarry = 0.0;
    int below_half = 0;
    // ---- weights and, round by round, the renderer reductions (same operations as render_fwd_kernel)
    ray_weights(a.sigma + r * S, eb, S, lane, [&](int s, float wv) {
      const bool ok = s < S;
      if (ok) {
        a.w_out[r * S + s] = wv;
        sw[s] = wv;
        if (C > 0) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            float v = __ldg(col + s * C + c);
            if (a.eval_mode) v = nan_to_num(v);
            comp[c] += wv * v;
          }
        }
      }
      acc += wv;
      const float step = ok ? (__ldg(eb + s) + __ldg(eb + s + 1)) / 2.f : 0.f;
      num += wv * step;
      if (ok) { smin = fminf(smin, step); smax = fmaxf(smax, step); }
      const double incl = scan_incl_d((double)wv, lane) + wcarry;
      below_half += __popc(__ballot_sync(0xffffffffu, ok && ((float)incl < 0.5f)));
      wcarry = __shfl_sync(0xffffffffu, incl, 31);
    });
    acc = warp_sum(acc);
#pragma unroll
    for (int c = 0; c < C; ++c) comp[c] = warp_sum(comp[c]);
    num = warp_sum(num);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
    if (lane == 0) {
      if (C > 0 && a.rgb_out) {
        const float bgv[4] = {a.bg.x, a.bg.y, a.bg.z, a.bg.w};
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float v = comp[c];
          if (a.bg_mode == 1) {
            float last = __ldg(col + (S - 1) * C + c);
            if (a.eval_mode) last = nan_to_num(last);
            v = v + last * (1.f - acc);
          } else if (a.bg_mode == 2) {
            v = v + bgv[c] * (1.f - acc);
          }
          if (a.eval_mode) v = fminf(fmaxf(v, 0.f), 1.f);
          a.rgb_out[r * C + c] = v;
        }
      }
      if (a.acc_out) a.acc_out[r] = acc;
      if (a.med_out) {
        const int idx = min(below_half, S - 1);
        a.med_out[r] = (__ldg(eb + idx) + __ldg(eb + idx + 1)) / 2.f;
      }
      if (a.exp_out) a.exp_out[r] = num / (acc + 1e-10f);
      if (mm) {  // launch-wide extrema of the sample midpoints (renderers.py:574 clips with them)
        const float cur_min = *reinterpret_cast<volatile float*>(mm);
        const float cur_max = *reinterpret_cast<volatile float*>(mm + 1);
        if (smin < cur_min) {
          atomicMin(reinterpret_cast<int*>(mm), smin >= 0.f ? __float_as_int(smin) : (int)0x80000000);
          if (smin < 0.f) atomicMax(reinterpret_cast<unsigned*>(mm), __float_as_uint(smin));
        }
        if (smax > cur_max) {
          if (smax >= 0.f) atomicMax(reinterpret_cast<int*>(mm + 1), __float_as_int(smax));
          else atomicMin(reinterpret_cast<unsigned*>(mm + 1), __float_as_uint(smax));
        }
      }
    }
    if (a.loss_acc) {
      // ---- distortion loss (same operations as distortion_kernel)
      for (int i = lane; i < S; i += 32) su[i] = (__ldg(sb + i + 1) + __ldg(sb + i)) / 2.f;
      __syncwarp();
      float total = 0.f;
      for (int i = lane; i < S; i += 32) {
        const float wi = sw[i], ui = su[i];
        float inner = 0.f;
        for (int j = 0; j < S; ++j) inner += sw[j] * fabsf(ui - su[j]);
        const float delta = __ldg(sb + i + 1) - __ldg(sb + i);
        total += wi * inner + wi * wi * delta / 3.f;
        if (a.dw_dist) a.dw_dist[r * S + i] = 2.f * inner + 2.f * wi * delta / 3.f;
      }
      dist_ray = warp_sum(total);
      // ---- interlevel loss against every proposal histogram (same operations as interlevel_kernel)
      for (int q = 0; q < a.n_prop; ++q) {
        const PropHist ph = a.prop[q];
        const int Sp = ph.Sp;
        float* edges = cy + (Sp + 1);
        float* diff = edges + (Sp + 1);
        __syncwarp();
        for (int k = lane; k <= Sp; k += 32) {
          edges[k] = __ldg(ph.sbins + r * (Sp + 1) + k);
          diff[k] = 0.f;
        }
        double carry = 0.0;
        if (lane == 0) cy[0] = 0.f;
        for (int base = 0; base < Sp; base += 32) {
          const int k = base + lane;
          const double v = scan_incl_d(k < Sp ? (double)__ldg(ph.w + r * Sp + k) : 0.0, lane) + carry;
          if (k < Sp) cy[k + 1] = (float)v;
          carry = __shfl_sync(0xffffffffu, v, 31);
        }
        __syncwarp();
        float tot = 0.f;
        for (int i = lane; i < S; i += 32) {
          const float t0s = __ldg(sb + i), t0e = __ldg(sb + i + 1);
          int lo = 0, hi = Sp;
          while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (edges[m] <= t0s) lo = m + 1; else hi = m;
          }
          const int idx_lo = min(max(lo - 1, 0), Sp - 1);
          lo = 0; hi = Sp;
          while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (edges[m + 1] <= t0e) lo = m + 1; else hi = m;
          }
          const int idx_hi = min(max(lo, 0), Sp - 1);
          const float w_outer = cy[idx_hi + 1] - cy[idx_lo];
          const float wi = sw[i];
          const float ex = fmaxf(wi - w_outer, 0.f);
          tot += ex * ex / (wi + 1.0e-7f);
          if (ph.dw && ex > 0.f) {
            const float g = -2.f * ex / (wi + 1.0e-7f);
            atomicAdd(diff + idx_lo, g);
            atomicAdd(diff + idx_hi + 1, -g);
          }
        }
        inter_ray += warp_sum(tot);
        if (ph.dw) {
          __syncwarp();
          float run = 0.f;
          for (int base = 0; base < Sp; base += 32) {
            const int k = base + lane;
            float v = k < Sp ? diff[k] : 0.f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const float u = __shfl_up_sync(0xffffffffu, v, o);
              if (lane >= o) v += u;
            }
            v += run;
            if (k < Sp) ph.dw[r * Sp + k] = v;
            run = __shfl_sync(0xffffffffu, v, 31);
          }
        }
      }
    }
  }
  if (a.loss_acc) {  // one pair of atomics per CTA
    if (lane == 0) {
      cta_loss[0][warp] = dist_ray;
      cta_loss[1][warp] = inter_ray;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kLevelWarps; ++w) t += cta_loss[threadIdx.x][w];
      atomicAdd(lacc + threadIdx.x, t);
    }
  }
  if (a.scratch) {
    // The CTA that takes the last ticket sees every other CTA's atomics and stores (fence before the ticket): it turns
    // the launch-wide accumulators into the final outputs -- what used to be a clone, a multiply and a clamp launch.
    __shared__ bool last_cta;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      last_cta = atomicAdd(reinterpret_cast<unsigned*>(a.scratch + 4), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_cta) {
      __threadfence();
      const float mn = __ldcg(a.scratch), mx = __ldcg(a.scratch + 1);
      __syncthreads();  // every thread holds the extrema before thread 0 re-arms the scratch
      if (a.exp_clip_out && a.exp_out && a.minmax) {
        // one CTA walks all R rays: 16-byte accesses, eight independent loads in flight per thread (a scalar loop
        // of dependent L2 round trips took 100 us at 32768 rays)
        auto clip = [&](float v) { return v < mn ? mn : (v > mx ? mx : v); };  // torch.clamp: NaN stays NaN
        const bool vec = ((reinterpret_cast<uintptr_t>(a.exp_out) | reinterpret_cast<uintptr_t>(a.exp_clip_out)) & 15) == 0;
        const int64_t n4 = vec ? a.R / 4 : 0;
        const float4* src = reinterpret_cast<const float4*>(a.exp_out);
        float4* dst = reinterpret_cast<float4*>(a.exp_clip_out);
        constexpr int U = 8;
        for (int64_t i0 = threadIdx.x; i0 < n4; i0 += (int64_t)blockDim.x * U) {
          float4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            v[u] = i < n4 ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            if (i < n4) dst[i] = make_float4(clip(v[u].x), clip(v[u].y), clip(v[u].z), clip(v[u].w));
          }
        }
        for (int64_t i = 4 * n4 + threadIdx.x; i < a.R; i += blockDim.x) a.exp_clip_out[i] = clip(__ldcg(a.exp_out + i));
      }
      if (threadIdx.x == 0) {
        if (a.minmax) { a.minmax[0] = mn; a.minmax[1] = mx; }
        if (a.loss_acc) {
          a.loss_acc[0] = __ldcg(a.scratch + 2) * a.scale_dist;
          a.loss_acc[1] = __ldcg(a.scratch + 3) * a.scale_inter;
        }
        a.scratch[0] = CUDART_INF_F; a.scratch[1] = -CUDART_INF_F; a.scratch[2] = 0.f; a.scratch[3] = 0.f;
        reinterpret_cast<unsigned*>(a.scratch)[4] = 0u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
struct PropBwd {
  const float* sigma;  // [R,Sp]
  const float* ebins;  // [R,Sp+1]
  const float* dw;     // [R,Sp] from the forward (PropHist::dw)
  float* dsigma;       // [R,Sp] out
  int Sp;
};
struct HeadsBwdArgs {
  const float* sigma;
  const float* col;
  const float* ebins;
  const float* w;        // [R,S] forward weights
  const float* dw_dist;  // [R,S] or NULL
  const float* d_rgb;    // [R,C] or NULL
  const float* d_acc;    // [R] or NULL
  const float* d_exp;    // [R] or NULL (gradient of the UNCLIPPED expected depth)
  const float* g_dist;   // device scalar: dL/d(mean distortion loss) (NULL: 0)
  const float* g_inter;  // device scalar: dL/d(mean interlevel loss) (NULL: 0)
  int64_t R;
  int S, bg_mode, n_prop, smem_floats_per_warp;
  float dist_scale, inter_scale;  // 1/R and 1/(R*S): the means the two losses are
  float4 bg;
  float* dsigma;  // [R,S] out
  float* dcol;    // [R,S,C] out (NULL: not wanted)
  PropBwd prop[2];
};

// blockIdx.y = 0: the final level; 1 + q: proposal level q.  smem per warp: gw[Smax] | tr[Smax]
template <int C>
__global__ void __launch_bounds__(32 * kLevelWarps) ray_heads_bwd_kernel(const HeadsBwdArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* gw = smem + (size_t)warp * a.smem_floats_per_warp;
  const int64_t r = (int64_t)blockIdx.x * kLevelWarps + warp;
  if (r >= a.R) return;
  if (blockIdx.y > 0) {
    const PropBwd p = a.prop[blockIdx.y - 1];
    const int Sp = p.Sp;
    const float gs = a.g_inter ? __ldg(a.g_inter) * a.inter_scale : 0.f;
    for (int s = lane; s < Sp; s += 32) gw[s] = __ldg(p.dw + r * Sp + s) * gs;
    __syncwarp();
    ray_weights_bwd(p.sigma + r * Sp, p.ebins + r * (Sp + 1), gw, gw + Sp, Sp, lane, p.dsigma + r * Sp);
    return;
  }
  const int S = a.S;
  const float* eb = a.ebins + r * (S + 1);
  const float* w = a.w + r * S;
  const float* col = a.col ? a.col + r * S * C : nullptr;
  // renderer gradients (same operations as render_bwd_kernel)
  float acc = 0.f, num = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float wv = __ldg(w + s);
    acc += wv;
    if (a.d_exp) num += wv * ((__ldg(eb + s) + __ldg(eb + s + 1)) / 2.f);
  }
  acc = warp_sum(acc);
  num = warp_sum(num);
  float g[C > 0 ? C : 1], bgc[C > 0 ? C : 1];
  const float bgv[4] = {a.bg.x, a.bg.y, a.bg.z, a.bg.w};
#pragma unroll
  for (int c = 0; c < C; ++c) {
    g[c] = a.d_rgb ? __ldg(a.d_rgb + r * C + c) : 0.f;
    bgc[c] = a.bg_mode == 1 ? __ldg(col + (S - 1) * C + c) : (a.bg_mode == 2 ? bgv[c] : 0.f);
  }
  const float ga = a.d_acc ? __ldg(a.d_acc + r) : 0.f;
  const float gd = a.d_exp ? __ldg(a.d_exp + r) : 0.f;
  const float gdist = (a.g_dist && a.dw_dist) ? __ldg(a.g_dist) * a.dist_scale : 0.f;
  const float den = acc + 1e-10f;
  for (int s = lane; s < S; s += 32) {
    const float wv = __ldg(w + s);
    float gwv = ga;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float cv = __ldg(col + s * C + c);
      gwv += g[c] * (cv - bgc[c]);
      float gc = g[c] * wv;
      if (a.bg_mode == 1 && s == S - 1) gc += g[c] * (1.f - acc);
      if (a.dcol) a.dcol[(r * S + s) * C + c] = gc;
    }
    if (a.d_exp) {
      const float step = (__ldg(eb + s) + __ldg(eb + s + 1)) / 2.f;
      gwv += gd * (step / den - num / (den * den));
    }
    if (a.dw_dist) gwv += __ldg(a.dw_dist + r * S + s) * gdist;
    gw[s] = gwv;
  }
  __syncwarp();
  ray_weights_bwd(a.sigma + r * S, eb, gw, gw + S, S, lane, a.dsigma + r * S);
}

}  // namespace tn

using namespace tn;

static inline unsigned level_blocks(int64_t R) { return (unsigned)((R + kLevelWarps - 1) / kLevelWarps); }

extern "C" int tn_level_resample(const float* sigma, const float* ebins, const float* sbins, const float* nears,
                                 const float* fars, const float* u_base, const float* jitter, int jitter_per_sample,
                                 const float* anneal_dev, int64_t R, int S, int S_new, float histogram_padding,
                                 float eps, float* weights_out, float* depth_median_out, float* sbins_new,
                                 float* ebins_new, void* stream) {
  TN_REQUIRE(sigma && ebins && weights_out, TN_EINVAL, "level_resample: null pointer");
  TN_REQUIRE((sbins_new == nullptr) == (ebins_new == nullptr), TN_EINVAL,
             "level_resample: sbins_new and ebins_new must both be given or both be NULL");
  TN_REQUIRE(!sbins_new || (sbins && nears && fars && u_base && S_new >= 1), TN_EINVAL,
             "level_resample: resampling needs sbins, nears, fars, u_base and S_new >= 1");
  TN_REQUIRE(R >= 0 && S >= 1 && S <= 1024, TN_EINVAL, "level_resample: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  const size_t smem = (size_t)kLevelWarps * 2 * (S + 1) * sizeof(float);
  level_resample_kernel<<<level_blocks(R), 32 * kLevelWarps, smem, (cudaStream_t)stream>>>(
      sigma, ebins, sbins, nears, fars, u_base, jitter, jitter_per_sample, anneal_dev, R, S, S_new, histogram_padding,
      eps, weights_out, depth_median_out, sbins_new, ebins_new);
  return check_launch("level_resample_kernel");
}

static float4 bg4(int bg_mode, int C, const float* bg_host) {
  float t[4] = {0, 0, 0, 0};
  if (bg_mode == 2)
    for (int c = 0; c < C; ++c) t[c] = bg_host[c];
  return make_float4(t[0], t[1], t[2], t[3]);
}

extern "C" int tn_ray_heads_fwd(const float* sigma, const float* colour, const float* ebins, const float* sbins,
                                int64_t R, int S, int C, int bg_mode, const float* bg_host, int eval_mode, int n_prop,
                                const float* const* prop_w_host_ptrs, const float* const* prop_sbins_host_ptrs,
                                const int* prop_S_host, float* const* prop_dw_host_ptrs, float* weights_out,
                                float* rgb_out, float* acc_out, float* depth_median_out, float* depth_expected_out,
                                float* steps_minmax_out, float* loss_acc, float* dw_distortion_out, float* scratch5,
                                float loss_scale_distortion, float loss_scale_interlevel,
                                float* depth_expected_clipped_out, void* stream) {
  TN_REQUIRE(sigma && ebins && weights_out, TN_EINVAL, "ray_heads_fwd: null pointer");
  TN_REQUIRE(scratch5 || !depth_expected_clipped_out, TN_EINVAL,
             "ray_heads_fwd: the clipped expected depth is made by the scratch protocol's last CTA");
  TN_REQUIRE(R >= 0 && S >= 1 && S <= 2048 && C >= 0 && C <= 4, TN_EINVAL, "ray_heads_fwd: bad R=%lld S=%d C=%d",
             (long long)R, S, C);
  TN_REQUIRE(C == 0 || colour, TN_EINVAL, "ray_heads_fwd: colour is null");
  TN_REQUIRE(bg_mode >= 0 && bg_mode <= 2 && (bg_mode != 2 || bg_host), TN_EINVAL, "ray_heads_fwd: bad bg_mode");
  TN_REQUIRE(n_prop >= 0 && n_prop <= 2, TN_EINVAL, "ray_heads_fwd: n_prop=%d (0..2)", n_prop);
  TN_REQUIRE(!loss_acc || sbins, TN_EINVAL, "ray_heads_fwd: the losses need sbins");
  TN_REQUIRE(loss_acc || (n_prop == 0 && !dw_distortion_out), TN_EINVAL, "ray_heads_fwd: loss outputs without loss_acc");
  if (R == 0) return TN_OK;
  HeadsFwdArgs a{};
  a.sigma = sigma; a.col = colour; a.ebins = ebins; a.sbins = sbins; a.R = R; a.S = S; a.bg_mode = bg_mode;
  a.eval_mode = eval_mode; a.n_prop = n_prop; a.bg = bg4(bg_mode, C, bg_host);
  int sp_max = 0;
  for (int q = 0; q < n_prop; ++q) {
    TN_REQUIRE(prop_w_host_ptrs && prop_sbins_host_ptrs && prop_S_host && prop_w_host_ptrs[q] && prop_sbins_host_ptrs[q] &&
                   prop_S_host[q] >= 1 && prop_S_host[q] <= 2048,
               TN_EINVAL, "ray_heads_fwd: bad proposal histogram %d", q);
    a.prop[q] = PropHist{prop_w_host_ptrs[q], prop_sbins_host_ptrs[q], prop_dw_host_ptrs ? prop_dw_host_ptrs[q] : nullptr,
                         prop_S_host[q]};
    sp_max = prop_S_host[q] > sp_max ? prop_S_host[q] : sp_max;
  }
  a.smem_floats_per_warp = 2 * S + (n_prop ? 3 * (sp_max + 1) : 0);
  a.w_out = weights_out; a.rgb_out = rgb_out; a.acc_out = acc_out; a.med_out = depth_median_out;
  a.exp_out = depth_expected_out; a.minmax = steps_minmax_out; a.loss_acc = loss_acc; a.dw_dist = dw_distortion_out;
  a.scratch = scratch5; a.exp_clip_out = depth_expected_clipped_out;
  a.scale_dist = loss_scale_distortion; a.scale_inter = loss_scale_interlevel;
  const size_t smem = (size_t)kLevelWarps * a.smem_floats_per_warp * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define TN_HF(CC)                                                                                       \
  do {                                                                                                  \
    if (smem > 48 * 1024)                                                                               \
      cudaFuncSetAttribute(ray_heads_fwd_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    ray_heads_fwd_kernel<CC><<<level_blocks(R), 32 * kLevelWarps, smem, st>>>(a);                        \
  } while (0)
  switch (C) {
    case 0: TN_HF(0); break;
    case 1: TN_HF(1); break;
    case 2: TN_HF(2); break;
    case 3: TN_HF(3); break;
    default: TN_HF(4); break;
  }
#undef TN_HF
  return check_launch("ray_heads_fwd_kernel");
}

extern "C" int tn_ray_heads_bwd(const float* sigma, const float* colour, const float* ebins, const float* weights,
                                const float* dw_distortion, const float* d_rgb, const float* d_acc,
                                const float* d_depth_expected, const float* g_distortion_dev,
                                const float* g_interlevel_dev, int64_t R, int S, int C, int bg_mode,
                                const float* bg_host, int n_prop, const float* const* prop_sigma_host_ptrs,
                                const float* const* prop_ebins_host_ptrs, const float* const* prop_dw_host_ptrs,
                                const int* prop_S_host, float* dsigma_out, float* dcolour_out,
                                float* const* prop_dsigma_host_ptrs, void* stream) {
  TN_REQUIRE(sigma && ebins && weights && dsigma_out, TN_EINVAL, "ray_heads_bwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1 && S <= 2048 && C >= 0 && C <= 4, TN_EINVAL, "ray_heads_bwd: bad R=%lld S=%d C=%d",
             (long long)R, S, C);
  TN_REQUIRE(C == 0 || colour, TN_EINVAL, "ray_heads_bwd: colour is null");
  TN_REQUIRE(bg_mode >= 0 && bg_mode <= 2 && (bg_mode != 2 || bg_host), TN_EINVAL, "ray_heads_bwd: bad bg_mode");
  TN_REQUIRE(n_prop >= 0 && n_prop <= 2, TN_EINVAL, "ray_heads_bwd: n_prop=%d (0..2)", n_prop);
  if (R == 0) return TN_OK;
  HeadsBwdArgs a{};
  a.sigma = sigma; a.col = colour; a.ebins = ebins; a.w = weights; a.dw_dist = dw_distortion; a.d_rgb = d_rgb;
  a.d_acc = d_acc; a.d_exp = d_depth_expected; a.g_dist = g_distortion_dev; a.g_inter = g_interlevel_dev; a.R = R;
  a.S = S; a.bg_mode = bg_mode; a.n_prop = n_prop; a.bg = bg4(bg_mode, C, bg_host);
  a.dist_scale = 1.f / (float)R;
  a.inter_scale = 1.f / ((float)R * (float)S);
  a.dsigma = dsigma_out; a.dcol = dcolour_out;
  int s_max = S;
  for (int q = 0; q < n_prop; ++q) {
    TN_REQUIRE(prop_sigma_host_ptrs && prop_ebins_host_ptrs && prop_dw_host_ptrs && prop_S_host && prop_dsigma_host_ptrs &&
                   prop_sigma_host_ptrs[q] && prop_ebins_host_ptrs[q] && prop_dw_host_ptrs[q] && prop_dsigma_host_ptrs[q] &&
                   prop_S_host[q] >= 1 && prop_S_host[q] <= 2048,
               TN_EINVAL, "ray_heads_bwd: bad proposal level %d", q);
    a.prop[q] = PropBwd{prop_sigma_host_ptrs[q], prop_ebins_host_ptrs[q], prop_dw_host_ptrs[q], prop_dsigma_host_ptrs[q],
                        prop_S_host[q]};
    s_max = prop_S_host[q] > s_max ? prop_S_host[q] : s_max;
  }
  a.smem_floats_per_warp = 2 * s_max;
  const size_t smem = (size_t)kLevelWarps * a.smem_floats_per_warp * sizeof(float);
  const dim3 grid(level_blocks(R), 1 + n_prop);
  cudaStream_t st = (cudaStream_t)stream;
#define TN_HB(CC)                                                                                       \
  do {                                                                                                  \
    if (smem > 48 * 1024)                                                                               \
      cudaFuncSetAttribute(ray_heads_bwd_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    ray_heads_bwd_kernel<CC><<<grid, 32 * kLevelWarps, smem, st>>>(a);                                   \
  } while (0)
  switch (C) {
    case 0: TN_HB(0); break;
    case 1: TN_HB(1); break;
    case 2: TN_HB(2); break;
    case 3: TN_HB(3); break;
    default: TN_HB(4); break;
  }
#undef TN_HB
  return check_launch("ray_heads_bwd_kernel");
}
