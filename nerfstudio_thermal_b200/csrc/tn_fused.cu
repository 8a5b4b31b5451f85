// Glue fusions around the field kernels and the per-ray losses.
//
//  * field_split: one pass turns the density MLP's output h[N,16] into the density (trunc_exp * selector,
//    fields/nerfacto_field.py:221-228) and the colour head's input row [SH16 | geo15 | appearance32]
//    (fields/nerfacto_field.py:335-344); its backward routes d(density) and d(head input) back into dH and reduces
//    the appearance-embedding gradient over the samples of each ray (no atomics: one warp owns one ray).
//  * density_act: the proposal fields' trunc_exp * selector (fields/density_fields.py:116-117).
//  * distortion / interlevel losses (model_components/losses.py:57-158): one warp per ray, forward value and
//    the gradient w.r.t. the weights in the same launch.
#include "tn_common.cuh"

namespace tn {

constexpr int kWarps = 4;

__device__ __forceinline__ float trunc_exp_grad(float x) { return expf(fminf(fmaxf(x, -15.f), 15.f)); }

// ------------------------------------------------------------------------------------------------ field split
// one warp per row: lane l writes columns l and l+32 of the head input (coalesced 252-byte rows)
__global__ void __launch_bounds__(32 * kWarps) field_split_fwd_kernel(
    const float* __restrict__ h, const float* __restrict__ sel, const float* __restrict__ sh,
    const float* __restrict__ emb_ray, int64_t N, int S, int hw, int geo, int emb_dim, int xs, float scale,
    float* __restrict__ density, float* __restrict__ xout) {
  const int lane = threadIdx.x & 31;
  const int in_dim = 16 + geo + emb_dim;  // xs >= in_dim: row stride of xout, padding columns written as zero
  for (int64_t p = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); p < N; p += (int64_t)gridDim.x * kWarps) {
    const int64_t r = p / S;
    if (lane == 0) density[p] = scale * expf(__ldg(h + p * hw)) * __ldg(sel + p);
    for (int c = lane; c < xs; c += 32) {
      float v = 0.f;
      if (c < 16) v = __ldg(sh + r * 16 + c);
      else if (c < 16 + geo) v = __ldg(h + p * hw + 1 + (c - 16));
      else if (c < in_dim) v = __ldg(emb_ray + r * emb_dim + (c - 16 - geo));
      xout[p * xs + c] = v;
    }
  }
}

// one warp per ray: dH[p] = [d_density * scale * sel * exp(clamp(h0)) | dX[p][16:16+geo]],
// d_emb_ray[r] = sum_s dX[r,s][16+geo:]
__global__ void __launch_bounds__(32 * kWarps) field_split_bwd_kernel(
    const float* __restrict__ h, const float* __restrict__ sel, const float* __restrict__ d_density,
    const float* __restrict__ dx, int64_t R, int S, int hw, int geo, int emb_dim, int xs, float scale,
    float* __restrict__ dh, float* __restrict__ demb_ray) {
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const int in_dim = xs;  // row stride of dx (>= 16 + geo + emb_dim)
  float acc0 = 0.f, acc1 = 0.f;  // lane's appearance columns: 16+geo+lane and 16+geo+lane+32
  for (int s = 0; s < S; ++s) {
    const int64_t p = r * S + s;
    if (lane == 0) {
      const float g = d_density ? __ldg(d_density + p) : 0.f;
      dh[p * hw] = g * scale * __ldg(sel + p) * trunc_exp_grad(__ldg(h + p * hw));
    }
    if (dx) {
      if (lane < geo) dh[p * hw + 1 + lane] = __ldg(dx + p * in_dim + 16 + lane);
      if (lane < emb_dim) acc0 += __ldg(dx + p * in_dim + 16 + geo + lane);
      if (lane + 32 < emb_dim) acc1 += __ldg(dx + p * in_dim + 16 + geo + lane + 32);
    } else if (lane < geo) {
      dh[p * hw + 1 + lane] = 0.f;
    }
    if (lane >= geo && lane + 1 < hw) dh[p * hw + 1 + lane] = 0.f;  // columns beyond 1+geo (none for hw = 16)
  }
  if (demb_ray && dx) {
    if (lane < emb_dim) demb_ray[r * emb_dim + lane] = acc0;
    if (lane + 32 < emb_dim) demb_ray[r * emb_dim + lane + 32] = acc1;
  }
}

// raw / draw are columns of row-major matrices: element i lives at [i * stride]
__global__ void density_act_fwd_kernel(const float* __restrict__ raw, int stride, const float* __restrict__ sel,
                                       int64_t N, float scale, float* __restrict__ density) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) density[i] = scale * expf(__ldg(raw + i * stride)) * __ldg(sel + i);
}
__global__ void density_act_bwd_kernel(const float* __restrict__ raw, int stride, const float* __restrict__ sel,
                                       const float* __restrict__ dd, int64_t N, float scale, float* __restrict__ draw,
                                       int dstride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) draw[i * dstride] = __ldg(dd + i) * scale * __ldg(sel + i) * trunc_exp_grad(__ldg(raw + i * stride));
}

// ------------------------------------------------------------------------------------------------ distortion loss
// lossfun_distortion (losses.py:139-150): sum_i w_i sum_j w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1}-t_i)/3, u = bin midpoints
__global__ void __launch_bounds__(32 * kWarps) distortion_kernel(const float* __restrict__ w, const float* __restrict__ t,
                                                                 int64_t R, int S, float* __restrict__ loss_ray,
                                                                 float* __restrict__ dw) {
  extern __shared__ float smem[];
  float* sw = smem + (size_t)(threadIdx.x >> 5) * 2 * S;
  float* su = sw + S;
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  for (int i = lane; i < S; i += 32) {
    const float a = __ldg(t + r * (S + 1) + i), b = __ldg(t + r * (S + 1) + i + 1);
    sw[i] = __ldg(w + r * S + i);
    su[i] = (b + a) / 2.f;
  }
  __syncwarp();
  float total = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float wi = sw[i], ui = su[i];
    float inner = 0.f;
    for (int j = 0; j < S; ++j) inner += sw[j] * fabsf(ui - su[j]);
    const float delta = __ldg(t + r * (S + 1) + i + 1) - __ldg(t + r * (S + 1) + i);
    total += wi * inner + wi * wi * delta / 3.f;
    if (dw) dw[r * S + i] = 2.f * inner + 2.f * wi * delta / 3.f;
  }
  total = warp_sum(total);
  if (lane == 0) loss_ray[r] = total;
}

// ------------------------------------------------------------------------------------------------ interlevel loss
// lossfun_outer (losses.py:57-103) of the fine histogram (c, w) against one proposal histogram (cp, wp):
//   w_outer_i = cumsum(wp)[hi_i] - cumsum(wp)[lo_i - 1],  loss_i = max(0, w_i - w_outer_i)^2 / (w_i + 1e-7)
__global__ void __launch_bounds__(32 * kWarps) interlevel_kernel(
    const float* __restrict__ w, const float* __restrict__ c, const float* __restrict__ wp, const float* __restrict__ cp,
    int64_t R, int Sf, int Sp, float* __restrict__ loss_ray, float* __restrict__ dwp) {
  extern __shared__ float smem[];
  float* cy = smem + (size_t)(threadIdx.x >> 5) * 3 * (Sp + 1);  // [0, cumsum(wp)]
  float* edges = cy + (Sp + 1);                                   // cp
  float* diff = edges + (Sp + 1);                                 // gradient difference array
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  for (int k = lane; k <= Sp; k += 32) {
    edges[k] = __ldg(cp + r * (Sp + 1) + k);
    diff[k] = 0.f;
  }
  // inclusive scan of wp in rounds of 32 (double carry, like the reference's CPU cumsum)
  double carry = 0.0;
  if (lane == 0) cy[0] = 0.f;
  for (int base = 0; base < Sp; base += 32) {
    const int k = base + lane;
    double v = k < Sp ? (double)__ldg(wp + r * Sp + k) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    v += carry;
    if (k < Sp) cy[k + 1] = (float)v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();
  float total = 0.f;
  for (int i = lane; i < Sf; i += 32) {
    const float t0s = __ldg(c + r * (Sf + 1) + i), t0e = __ldg(c + r * (Sf + 1) + i + 1);
    // searchsorted(cp[:-1], t0s, right) - 1 : number of starts <= t0s, minus one
    int lo = 0, hi = Sp;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (edges[m] <= t0s) lo = m + 1; else hi = m;
    }
    const int idx_lo = min(max(lo - 1, 0), Sp - 1);
    // searchsorted(cp[1:], t0e, right) : number of ends <= t0e
    lo = 0; hi = Sp;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (edges[m + 1] <= t0e) lo = m + 1; else hi = m;
    }
    const int idx_hi = min(max(lo, 0), Sp - 1);
    const float w_outer = cy[idx_hi + 1] - cy[idx_lo];
    const float wi = __ldg(w + r * Sf + i);
    const float ex = fmaxf(wi - w_outer, 0.f);
    total += ex * ex / (wi + 1.0e-7f);
    if (dwp && ex > 0.f) {
      const float g = -2.f * ex / (wi + 1.0e-7f);  // d loss_i / d w_outer_i
      atomicAdd(diff + idx_lo, g);
      atomicAdd(diff + idx_hi + 1, -g);
    }
  }
  total = warp_sum(total);
  if (lane == 0) loss_ray[r] = total;
  if (dwp) {
    __syncwarp();
    float run = 0.f;  // prefix sum of the difference array, again in rounds of 32
    for (int base = 0; base < Sp; base += 32) {
      const int k = base + lane;
      float v = k < Sp ? diff[k] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      v += run;
      if (k < Sp) dwp[r * Sp + k] = v;
      run = __shfl_sync(0xffffffffu, v, 31);
    }
  }
}

}  // namespace tn

using namespace tn;

static inline unsigned warp_blocks(int64_t n) { return (unsigned)((n + kWarps - 1) / kWarps); }

extern "C" int tn_field_split_fwd(const float* h, const float* sel, const float* sh, const float* emb_ray, int64_t R,
                                  int S, int h_width, int geo_dim, int emb_dim, int x_stride, float density_scale,
                                  float* density_out, float* x_out, void* stream) {
  TN_REQUIRE(h && sel && sh && density_out && x_out && (emb_ray || emb_dim == 0), TN_EINVAL, "field_split_fwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1 && h_width >= 1 + geo_dim && geo_dim >= 0 && geo_dim <= 31 && emb_dim >= 0 && emb_dim <= 64,
             TN_EINVAL, "field_split_fwd: bad sizes");
  if (x_stride == 0) x_stride = 16 + geo_dim + emb_dim;
  TN_REQUIRE(x_stride >= 16 + geo_dim + emb_dim, TN_EINVAL, "field_split_fwd: x_stride=%d too small", x_stride);
  if (R == 0) return TN_OK;
  const int64_t N = R * S;
  const unsigned grid = (unsigned)min((int64_t)kNumSMs * 32, (N + kWarps - 1) / kWarps);
  field_split_fwd_kernel<<<grid, 32 * kWarps, 0, (cudaStream_t)stream>>>(h, sel, sh, emb_ray, N, S, h_width, geo_dim,
                                                                       emb_dim, x_stride, density_scale, density_out,
                                                                       x_out);
  return check_launch("field_split_fwd_kernel");
}

extern "C" int tn_field_split_bwd(const float* h, const float* sel, const float* d_density, const float* dx, int64_t R,
                                  int S, int h_width, int geo_dim, int emb_dim, int x_stride, float density_scale,
                                  float* dh_out, float* demb_ray_out, void* stream) {
  TN_REQUIRE(h && sel && dh_out, TN_EINVAL, "field_split_bwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1 && h_width >= 1 + geo_dim && geo_dim >= 0 && geo_dim <= 31 && emb_dim >= 0 && emb_dim <= 64,
             TN_EINVAL, "field_split_bwd: bad sizes");
  if (x_stride == 0) x_stride = 16 + geo_dim + emb_dim;
  TN_REQUIRE(x_stride >= 16 + geo_dim + emb_dim, TN_EINVAL, "field_split_bwd: x_stride=%d too small", x_stride);
  if (R == 0) return TN_OK;
  field_split_bwd_kernel<<<warp_blocks(R), 32 * kWarps, 0, (cudaStream_t)stream>>>(
      h, sel, d_density, dx, R, S, h_width, geo_dim, emb_dim, x_stride, density_scale, dh_out, demb_ray_out);
  return check_launch("field_split_bwd_kernel");
}

extern "C" int tn_density_act_fwd(const float* raw, int raw_stride, const float* sel, int64_t N, float scale,
                                  float* density_out, void* stream) {
  TN_REQUIRE(N == 0 || (raw && sel && density_out && raw_stride >= 1), TN_EINVAL, "density_act_fwd: bad arguments");
  if (N <= 0) return N == 0 ? TN_OK : TN_EINVAL;
  density_act_fwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw, raw_stride, sel, N, scale,
                                                                                      density_out);
  return check_launch("density_act_fwd_kernel");
}

extern "C" int tn_density_act_bwd(const float* raw, int raw_stride, const float* sel, const float* d_density, int64_t N,
                                  float scale, float* draw_out, int draw_stride, void* stream) {
  TN_REQUIRE(N == 0 || (raw && sel && d_density && draw_out && raw_stride >= 1 && draw_stride >= 1), TN_EINVAL,
             "density_act_bwd: bad arguments");
  if (N <= 0) return N == 0 ? TN_OK : TN_EINVAL;
  density_act_bwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw, raw_stride, sel, d_density,
                                                                                      N, scale, draw_out, draw_stride);
  return check_launch("density_act_bwd_kernel");
}

extern "C" int tn_distortion_loss(const float* weights, const float* sbins, int64_t R, int S, float* loss_ray_out,
                                  float* dweights_out, void* stream) {
  TN_REQUIRE(weights && sbins && loss_ray_out, TN_EINVAL, "distortion_loss: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1 && S <= 2048, TN_EINVAL, "distortion_loss: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  const size_t smem = (size_t)kWarps * 2 * S * sizeof(float);
  distortion_kernel<<<warp_blocks(R), 32 * kWarps, smem, (cudaStream_t)stream>>>(weights, sbins, R, S, loss_ray_out,
                                                                               dweights_out);
  return check_launch("distortion_kernel");
}

extern "C" int tn_interlevel_loss(const float* w_fine, const float* sbins_fine, const float* w_prop,
                                  const float* sbins_prop, int64_t R, int S_fine, int S_prop, float* loss_ray_out,
                                  float* dw_prop_out, void* stream) {
  TN_REQUIRE(w_fine && sbins_fine && w_prop && sbins_prop && loss_ray_out, TN_EINVAL, "interlevel_loss: null pointer");
  TN_REQUIRE(R >= 0 && S_fine >= 1 && S_prop >= 1 && S_prop <= 2048, TN_EINVAL, "interlevel_loss: bad sizes");
  if (R == 0) return TN_OK;
  const size_t smem = (size_t)kWarps * 3 * (S_prop + 1) * sizeof(float);
  interlevel_kernel<<<warp_blocks(R), 32 * kWarps, smem, (cudaStream_t)stream>>>(
      w_fine, sbins_fine, w_prop, sbins_prop, R, S_fine, S_prop, loss_ray_out, dw_prop_out);
  return check_launch("interlevel_kernel");
}
