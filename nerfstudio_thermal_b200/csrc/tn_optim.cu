// Fused multi-group Adam over the flat parameter / gradient buffers.
//
// The reference steps one torch.optim.Adam per parameter group (engine/optimizers.py:67-95, 172-180; groups and
// hyper-parameters in configs/method_configs.py:274-301) and one LambdaLR exponential-decay scheduler per group
// (engine/schedulers.py:109-142).  Here every parameter, gradient and moment lives in one flat fp32 buffer, the
// groups are contiguous ranges of it, and ONE launch applies the dense Adam update to all of them.  The learning
// rate of each group is evaluated on the device from the step counter (the scheduler's closed form, in double),
// so the launch can sit inside a captured CUDA graph: nothing on the host changes from step to step.
//
// Pure streaming work: 4 reads + 3 writes (+1 when the gradient is cleared) of 4 bytes per parameter, float4 wide.
#include "tn_common.cuh"

namespace tn {

struct AdamGroup {
  int64_t begin, end;                 // element range in the flat buffers
  float lr_init, lr_final, lr_pre_warmup, eps, weight_decay;
  int warmup_steps, max_steps, ramp;  // max_steps == 0: constant lr_init;  ramp: 0 linear, 1 cosine
};
struct AdamGroups {
  AdamGroup g[TN_ADAM_MAX_GROUPS];
  int n;
};

// lr_init * lr_lambda(k) with lr_lambda of schedulers.py:124-139 (k = scheduler steps taken so far)
__device__ double scheduled_lr(const AdamGroup& g, int k) {
  const double lr_init = (double)g.lr_init;
  if (g.max_steps <= 0) return lr_init;
  const double lr_final = g.lr_final > 0.f ? (double)g.lr_final : lr_init;
  double lr;
  if (k < g.warmup_steps) {
    const double pre = (double)g.lr_pre_warmup;
    if (g.ramp == 1) {
      double t = (double)k / (double)g.warmup_steps;
      t = fmin(fmax(t, 0.0), 1.0);
      lr = pre + (lr_init - pre) * sin(0.5 * 3.141592653589793 * t);
    } else {
      lr = pre + (lr_init - pre) * (double)k / (double)g.warmup_steps;
    }
  } else {
    double t = (double)(k - g.warmup_steps) / (double)(g.max_steps - g.warmup_steps);
    t = fmin(fmax(t, 0.0), 1.0);
    lr = exp(log(lr_init) * (1.0 - t) + log(lr_final) * t);
  }
  return lr_init * (lr / lr_init);  // LambdaLR multiplies the initial rate by the returned ratio
}

struct AdamCoef {
  float step_size, bc2_sqrt, eps, wd;
};

// the python scalars 1 - beta are formed in double and only then rounded to fp32 (1.f - 0.999f is off by 1e-5)
struct AdamBetas {
  float b2, omb1, omb2;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamCoef& c,
                                            const AdamBetas& b) {
  // torch/optim/adam.py, single-tensor path, amsgrad off, maximize off
  if (c.wd != 0.f) g = g + c.wd * p;
  m = m + (g - m) * b.omb1;                     // exp_avg.lerp_(grad, 1 - beta1)
  v = v * b.b2 + b.omb2 * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
  p = p - c.step_size * (m / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
}

constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                            float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                            AdamGroups groups, double beta1, double beta2,
                                                            const int32_t* __restrict__ step_dev, int step_host,
                                                            int per_group_steps, uint32_t active_mask,
                                                            const float* __restrict__ inv_scale_dev,
                                                            const float* __restrict__ found_inf_dev, int zero_grads) {
  __shared__ AdamCoef coef[TN_ADAM_MAX_GROUPS];
  __shared__ int64_t gbegin[TN_ADAM_MAX_GROUPS], gend[TN_ADAM_MAX_GROUPS];
  // step_dev[0]: 1-based number of this ITERATION (the schedulers have stepped iteration-1 times);
  // step_dev[1+k] (per_group_steps): 1-based number of group k's Adam step -- torch's state['step'], which lags the
  // iteration count for groups that sat out steps without a gradient
  const int iteration = step_dev ? step_dev[0] : step_host;
  if (threadIdx.x < groups.n) {
    const AdamGroup& gr = groups.g[threadIdx.x];
    const int step = (step_dev && per_group_steps) ? step_dev[1 + threadIdx.x] : iteration;
    const double lr = scheduled_lr(gr, iteration - 1);
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    coef[threadIdx.x] = {(float)(lr / bc1), (float)sqrt(bc2), gr.eps, gr.weight_decay};
    // a group that is not active this step (no gradient: torch.optim.Adam skips parameters whose grad is None,
    // engine/optimizers.py:165-170) covers no elements
    const bool on = (active_mask >> threadIdx.x) & 1u;
    gbegin[threadIdx.x] = gr.begin;
    gend[threadIdx.x] = on ? gr.end : gr.begin;
  }
  __syncthreads();
  const bool skip = found_inf_dev && *found_inf_dev != 0.f;  // GradScaler.step: no update on inf/nan gradients
  const float inv_scale = inv_scale_dev ? *inv_scale_dev : 1.f;
  const int ng = groups.n;
  const AdamBetas bt = {(float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2)};
  auto group_of = [&](int64_t i) {
    int k = -1;
    for (int q = 0; q < ng; ++q)
      if (i >= gbegin[q] && i < gend[q]) k = q;
    return k;
  };
  const int64_t n4 = n >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * kAdamThreads) {
    const int64_t i = i4 << 2;
    const int k0 = group_of(i), k3 = group_of(i + 3);
    float4 gq = reinterpret_cast<const float4*>(g)[i4];
    if (skip || (k0 < 0 && k3 < 0)) {  // alignment padding between parameters of no group
      if (zero_grads) reinterpret_cast<float4*>(g)[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float4 pq = reinterpret_cast<const float4*>(p)[i4];
    float4 mq = reinterpret_cast<const float4*>(m)[i4];
    float4 vq = reinterpret_cast<const float4*>(v)[i4];
    float* pp = &pq.x; float* gg = &gq.x; float* mm = &mq.x; float* vv = &vq.x;
    if (k0 == k3) {
      const AdamCoef c = coef[k0];
#pragma unroll
      for (int e = 0; e < 4; ++e) adam_update(pp[e], gg[e] * inv_scale, mm[e], vv[e], c, bt);
    } else {  // a group boundary inside the vector
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = group_of(i + e);
        if (k >= 0) adam_update(pp[e], gg[e] * inv_scale, mm[e], vv[e], coef[k], bt);
      }
    }
    reinterpret_cast<float4*>(p)[i4] = pq;
    reinterpret_cast<float4*>(m)[i4] = mq;
    reinterpret_cast<float4*>(v)[i4] = vq;
    // cleared last: a store issued right behind the load of the same line stalls on the line still in flight
    if (zero_grads) reinterpret_cast<float4*>(g)[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const int k = group_of(i);
    const float gv = g[i] * inv_scale;
    if (zero_grads) g[i] = 0.f;
    if (!skip && k >= 0) adam_update(p[i], gv, m[i], v[i], coef[k], bt);
  }
}

// GradScaler.unscale_ + inf check over the flat gradient: found_inf = 1 if any g * inv_scale is not finite
__global__ void __launch_bounds__(256) grad_check_kernel(const float* __restrict__ g, int64_t n,
                                                         const float* __restrict__ inv_scale_dev,
                                                         float* __restrict__ found_inf) {
  const float inv_scale = inv_scale_dev ? *inv_scale_dev : 1.f;
  bool bad = false;
  const int64_t n4 = n >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * 256 + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * 256) {
    const float4 q = reinterpret_cast<const float4*>(g)[i4];
    bad |= !isfinite(q.x * inv_scale) | !isfinite(q.y * inv_scale) | !isfinite(q.z * inv_scale) |
           !isfinite(q.w * inv_scale);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) bad |= !isfinite(g[(n4 << 2) + threadIdx.x] * inv_scale);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.f;
}

__global__ void counter_add_kernel(int32_t* c, int v) { *c += v; }

// counters[0] += 1 (iteration), counters[1+k] += 1 for every active group k
__global__ void step_counters_tick_kernel(int32_t* c, int n_groups, uint32_t active_mask) {
  const int t = threadIdx.x;
  if (t == 0) c[0] += 1;
  else if (t <= n_groups && ((active_mask >> (t - 1)) & 1u)) c[t] += 1;
}

// ---- gradient exchange over NVLink peer memory (parallel.PeerExchange): the reduction step.
// dst[i] = (dst[i] + sum_k src[k*stride + i]) * scale: this rank's shard plus the copies of the same shard that the
// copy engines pulled from the peers, averaged.  Streaming, float4, grid-stride; the only SM work of the exchange.
__global__ void __launch_bounds__(256) shard_mean_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                         int n_src, int64_t stride, int64_t n4, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(dst)[i];
    for (int k = 0; k < n_src; ++k) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * stride) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                            const int64_t* group_begin_host, const int64_t* group_end_host,
                            const float* group_hyper_host, int n_groups, double beta1, double beta2,
                            const int32_t* step_dev, int step_host, int per_group_steps, uint32_t active_mask,
                            const float* inv_scale_dev, const float* found_inf_dev, int zero_grads, void* stream) {
  TN_REQUIRE(n >= 0 && n_groups >= 0 && n_groups <= TN_ADAM_MAX_GROUPS, TN_EINVAL, "adam_step: n=%lld groups=%d",
             (long long)n, n_groups);
  if (n == 0 || n_groups == 0) return TN_OK;
  TN_REQUIRE(params && grads && exp_avg && exp_avg_sq && group_begin_host && group_end_host && group_hyper_host,
             TN_EINVAL, "adam_step: null pointer");
  TN_REQUIRE(aligned(params, 16) && aligned(grads, 16) && aligned(exp_avg, 16) && aligned(exp_avg_sq, 16), TN_EALIGN,
             "adam_step: buffers must be 16-byte aligned");
  TN_REQUIRE(step_dev || step_host >= 1, TN_EINVAL, "adam_step: step=%d (1-based)", step_host);
  TN_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, TN_EINVAL, "adam_step: betas");
  AdamGroups gs = {};
  gs.n = n_groups;
  for (int i = 0; i < n_groups; ++i) {
    const float* h = group_hyper_host + (size_t)i * TN_ADAM_HYPER;
    AdamGroup& a = gs.g[i];
    a.begin = group_begin_host[i];
    a.end = group_end_host[i];
    TN_REQUIRE(a.begin >= 0 && a.begin <= a.end && a.end <= n, TN_EINVAL, "adam_step: group %d range [%lld,%lld)", i,
               (long long)a.begin, (long long)a.end);
    TN_REQUIRE(i == 0 || a.begin >= gs.g[i - 1].end, TN_EINVAL, "adam_step: groups must be ordered and disjoint");
    a.lr_init = h[0]; a.lr_final = h[1]; a.lr_pre_warmup = h[2]; a.eps = h[3]; a.weight_decay = h[4];
    a.warmup_steps = (int)h[5]; a.max_steps = (int)h[6]; a.ramp = (int)h[7];
    TN_REQUIRE(a.lr_init > 0.f && a.eps >= 0.f && a.warmup_steps >= 0 && a.max_steps >= 0 &&
                   (a.max_steps == 0 || a.max_steps > a.warmup_steps),
               TN_EINVAL, "adam_step: group %d hyper-parameters", i);
  }
  const int64_t n4 = (n + 3) / 4;
  const unsigned grid = (unsigned)min((n4 + kAdamThreads - 1) / kAdamThreads, (int64_t)kNumSMs * 8);
  adam_kernel<<<grid, kAdamThreads, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, gs, beta1, beta2,
                                                               step_dev, step_host, per_group_steps, active_mask,
                                                               inv_scale_dev, found_inf_dev, zero_grads);
  return check_launch("adam_kernel");
}

extern "C" int tn_grad_unscale_check(const float* grads, int64_t n, const float* inv_scale_dev, float* found_inf_dev,
                                     void* stream) {
  TN_REQUIRE(n >= 0 && found_inf_dev && (grads || n == 0), TN_EINVAL, "grad_unscale_check: bad arguments");
  cudaError_t e = cudaMemsetAsync(found_inf_dev, 0, sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("grad_unscale_check: %s", cudaGetErrorString(e));
    return TN_ECUDA;
  }
  if (n == 0) return TN_OK;
  TN_REQUIRE(aligned(grads, 16), TN_EALIGN, "grad_unscale_check: buffer must be 16-byte aligned");
  const int64_t n4 = (n + 3) / 4;
  const unsigned grid = (unsigned)min((n4 + 255) / 256, (int64_t)kNumSMs * 8);
  grad_check_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grads, n, inv_scale_dev, found_inf_dev);
  return check_launch("grad_check_kernel");
}

extern "C" int tn_step_counters_tick(int32_t* counters_dev, int n_groups, uint32_t active_mask, void* stream) {
  TN_REQUIRE(counters_dev && n_groups >= 0 && n_groups <= TN_ADAM_MAX_GROUPS, TN_EINVAL,
             "step_counters_tick: bad arguments");
  step_counters_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counters_dev, n_groups, active_mask);
  return check_launch("step_counters_tick_kernel");
}

extern "C" int tn_counter_add(int32_t* counter_dev, int value, void* stream) {
  TN_REQUIRE(counter_dev, TN_EINVAL, "counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, value);
  return check_launch("counter_add_kernel");
}

extern "C" int tn_shard_mean(float* dst, const float* src, int n_src, int64_t src_stride, int64_t n, float scale,
                             int max_ctas, void* stream) {
  TN_REQUIRE(dst && (src || n_src == 0), TN_EINVAL, "shard_mean: null pointer");
  TN_REQUIRE(n >= 0 && n % 4 == 0 && n_src >= 0 && src_stride % 4 == 0 && src_stride >= n, TN_EINVAL,
             "shard_mean: n=%lld (multiple of 4), n_src=%d, stride=%lld", (long long)n, n_src, (long long)src_stride);
  TN_REQUIRE(aligned(dst, 16) && aligned(src, 16), TN_EALIGN, "shard_mean: pointers must be 16-byte aligned");
  if (n == 0) return TN_OK;
  const int64_t n4 = n / 4;
  int64_t ctas = (n4 + 255) / 256;
  const int64_t cap = max_ctas > 0 ? max_ctas : 2 * kNumSMs;
  if (ctas > cap) ctas = cap;
  shard_mean_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(dst, src, n_src, src_stride, n4, scale);
  return check_launch("shard_mean_kernel");
}
