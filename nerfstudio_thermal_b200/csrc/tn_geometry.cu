// Sample positions, L-inf scene contraction and grid normalisation; SH basis; initial sampler bins.
// Compiled with -fmad=false so each operation rounds like the reference's separate torch ops: the grid
// coordinates feed floor()/ceil() in the hash encoder and must not drift by an ulp.
#include "tn_geometry.cuh"

namespace tn {

__global__ void sample_positions_fwd_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                            const float* __restrict__ ebins, int64_t R, int S, float* __restrict__ x,
                                            float* __restrict__ sel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const int64_t r = i / S;
  const int s = (int)(i - r * S);
  const float st = __ldg(ebins + r * (S + 1) + s), en = __ldg(ebins + r * (S + 1) + s + 1);
  float p0 = sample_pos(__ldg(o + 3 * r), __ldg(d + 3 * r), st, en);
  float p1 = sample_pos(__ldg(o + 3 * r + 1), __ldg(d + 3 * r + 1), st, en);
  float p2 = sample_pos(__ldg(o + 3 * r + 2), __ldg(d + 3 * r + 2), st, en);
  const float sv = contract_normalise(p0, p1, p2);
  x[3 * i] = p0; x[3 * i + 1] = p1; x[3 * i + 2] = p2;
  sel[i] = sv;
}

// one warp per ray: reduce the per-sample position gradients into d_origin / d_direction
__global__ void sample_positions_bwd_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                            const float* __restrict__ ebins, const float* __restrict__ dx, int64_t R,
                                            int S, float* __restrict__ dorig, float* __restrict__ ddir) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float o0 = __ldg(o + 3 * r), o1 = __ldg(o + 3 * r + 1), o2 = __ldg(o + 3 * r + 2);
  const float d0 = __ldg(d + 3 * r), d1 = __ldg(d + 3 * r + 1), d2 = __ldg(d + 3 * r + 2);
  float ao0 = 0, ao1 = 0, ao2 = 0, ad0 = 0, ad1 = 0, ad2 = 0;
  for (int s = lane; s < S; s += 32) {
    const float st = __ldg(ebins + r * (S + 1) + s), en = __ldg(ebins + r * (S + 1) + s + 1);
    const float p0 = sample_pos(o0, d0, st, en), p1 = sample_pos(o1, d1, st, en), p2 = sample_pos(o2, d2, st, en);
    const int64_t i = r * S + s;
    float g0 = __ldg(dx + 3 * i), g1 = __ldg(dx + 3 * i + 1), g2 = __ldg(dx + 3 * i + 2);
    contract_normalise_bwd(p0, p1, p2, g0, g1, g2);
    const float t = (st + en) / 2.f;
    ao0 += g0; ao1 += g1; ao2 += g2;
    ad0 += g0 * t; ad1 += g1 * t; ad2 += g2 * t;
  }
  ao0 = warp_sum(ao0); ao1 = warp_sum(ao1); ao2 = warp_sum(ao2);
  ad0 = warp_sum(ad0); ad1 = warp_sum(ad1); ad2 = warp_sum(ad2);
  if (lane == 0) {
    dorig[3 * r] = ao0; dorig[3 * r + 1] = ao1; dorig[3 * r + 2] = ao2;
    ddir[3 * r] = ad0; ddir[3 * r + 1] = ad1; ddir[3 * r + 2] = ad2;
  }
}

__global__ void contract_points_fwd_kernel(const float* __restrict__ p, int64_t N, float* __restrict__ x,
                                           float* __restrict__ sel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float p0 = __ldg(p + 3 * i), p1 = __ldg(p + 3 * i + 1), p2 = __ldg(p + 3 * i + 2);
  const float sv = contract_normalise(p0, p1, p2);
  x[3 * i] = p0; x[3 * i + 1] = p1; x[3 * i + 2] = p2;
  sel[i] = sv;
}

__global__ void contract_points_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dx, int64_t N,
                                           float* __restrict__ dp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float g0 = __ldg(dx + 3 * i), g1 = __ldg(dx + 3 * i + 1), g2 = __ldg(dx + 3 * i + 2);
  contract_normalise_bwd(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2), g0, g1, g2);
  dp[3 * i] = g0; dp[3 * i + 1] = g1; dp[3 * i + 2] = g2;
}

// utils/math.py:29-95, levels = 4
__global__ void sh4_kernel(const float* __restrict__ d, int64_t N, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float x = __ldg(d + 3 * i), y = __ldg(d + 3 * i + 1), z = __ldg(d + 3 * i + 2);
  const float xx = x * x, yy = y * y, zz = z * z;
  float c[16];
  c[0] = 0.28209479177387814f;
  c[1] = 0.4886025119029199f * y;
  c[2] = 0.4886025119029199f * z;
  c[3] = 0.4886025119029199f * x;
  c[4] = 1.0925484305920792f * x * y;
  c[5] = 1.0925484305920792f * y * z;
  c[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
  c[7] = 1.0925484305920792f * x * z;
  c[8] = 0.5462742152960396f * (xx - yy);
  c[9] = 0.5900435899266435f * y * (3.f * xx - yy);
  c[10] = 2.890611442640554f * x * y * z;
  c[11] = 0.4570457994644658f * y * (5.f * zz - 1.f);
  c[12] = 0.3731763325901154f * z * (5.f * zz - 3.f);
  c[13] = 0.4570457994644658f * x * (5.f * zz - 1.f);
  c[14] = 1.445305721320277f * z * (xx - yy);
  c[15] = 0.5900435899266435f * x * (xx - 3.f * yy);
  float4* o4 = reinterpret_cast<float4*>(out + 16 * i);
#pragma unroll
  for (int k = 0; k < 4; ++k) o4[k] = make_float4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
}

// UniformLinDispPiecewiseSampler spacing (model_components/ray_samplers.py:244-245)
__device__ __forceinline__ float spacing_fn(float x) { return x < 1.f ? x / 2.f : 1.f - 1.f / (2.f * x); }
__device__ __forceinline__ float spacing_inv(float y) { return y < 0.5f ? 2.f * y : 1.f / (2.f - 2.f * y); }

__global__ void piecewise_bins_kernel(const float* __restrict__ unit, const float* __restrict__ nears,
                                      const float* __restrict__ fars, const float* __restrict__ jitter,
                                      int jitter_per_sample, int64_t R, int S, float* __restrict__ sbins,
                                      float* __restrict__ ebins) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * (S + 1)) return;
  const int64_t r = i / (S + 1);
  const int k = (int)(i - r * (S + 1));
  float b = __ldg(unit + k);
  if (jitter) {  // ray_samplers.py:103-111: one draw per ray (single_jitter) or per bin edge
    const float lower = k == 0 ? b : (b + __ldg(unit + k - 1)) / 2.f;
    const float upper = k == S ? b : (__ldg(unit + k + 1) + b) / 2.f;
    b = lower + (upper - lower) * __ldg(jitter + (jitter_per_sample ? i : r));
  }
  const float sn = spacing_fn(__ldg(nears + r)), sf = spacing_fn(__ldg(fars + r));
  sbins[i] = b;
  ebins[i] = spacing_inv(b * sf + (1.f - b) * sn);  // :115-116
}

}  // namespace tn

using namespace tn;

static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" int tn_sample_positions_fwd(const float* origins, const float* directions, const float* ebins, int64_t R,
                                       int S, float* x_out, float* selector_out, void* stream) {
  TN_REQUIRE(origins && directions && ebins && x_out && selector_out, TN_EINVAL, "sample_positions_fwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "sample_positions_fwd: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  sample_positions_fwd_kernel<<<blocks_for(R * S, 256), 256, 0, (cudaStream_t)stream>>>(origins, directions, ebins, R,
                                                                                       S, x_out, selector_out);
  return check_launch("sample_positions_fwd_kernel");
}

extern "C" int tn_sample_positions_bwd(const float* origins, const float* directions, const float* ebins,
                                       const float* dx, int64_t R, int S, float* d_origins, float* d_directions,
                                       void* stream) {
  TN_REQUIRE(origins && directions && ebins && dx && d_origins && d_directions, TN_EINVAL,
             "sample_positions_bwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "sample_positions_bwd: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  sample_positions_bwd_kernel<<<blocks_for(R * 32, 128), 128, 0, (cudaStream_t)stream>>>(origins, directions, ebins, dx,
                                                                                        R, S, d_origins, d_directions);
  return check_launch("sample_positions_bwd_kernel");
}

extern "C" int tn_contract_points_fwd(const float* p, int64_t N, float* x_out, float* selector_out, void* stream) {
  TN_REQUIRE(p && x_out && selector_out, TN_EINVAL, "contract_points_fwd: null pointer");
  if (N <= 0) return N == 0 ? TN_OK : TN_EINVAL;
  contract_points_fwd_kernel<<<blocks_for(N, 256), 256, 0, (cudaStream_t)stream>>>(p, N, x_out, selector_out);
  return check_launch("contract_points_fwd_kernel");
}

extern "C" int tn_contract_points_bwd(const float* p, const float* dx, int64_t N, float* dp, void* stream) {
  TN_REQUIRE(p && dx && dp, TN_EINVAL, "contract_points_bwd: null pointer");
  if (N <= 0) return N == 0 ? TN_OK : TN_EINVAL;
  contract_points_bwd_kernel<<<blocks_for(N, 256), 256, 0, (cudaStream_t)stream>>>(p, dx, N, dp);
  return check_launch("contract_points_bwd_kernel");
}

extern "C" int tn_sh4(const float* d, int64_t N, float* out, void* stream) {
  TN_REQUIRE(d && out, TN_EINVAL, "sh4: null pointer");
  TN_REQUIRE(aligned(out, 16), TN_EALIGN, "sh4: out must be 16-byte aligned");
  if (N <= 0) return N == 0 ? TN_OK : TN_EINVAL;
  sh4_kernel<<<blocks_for(N, 256), 256, 0, (cudaStream_t)stream>>>(d, N, out);
  return check_launch("sh4_kernel");
}

extern "C" int tn_piecewise_bins(const float* unit_bins, const float* nears, const float* fars, const float* jitter,
                                 int jitter_per_sample, int64_t R, int S, float* sbins_out, float* ebins_out,
                                 void* stream) {
  TN_REQUIRE(unit_bins && nears && fars && sbins_out && ebins_out, TN_EINVAL, "piecewise_bins: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "piecewise_bins: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  piecewise_bins_kernel<<<blocks_for(R * (S + 1), 256), 256, 0, (cudaStream_t)stream>>>(
      unit_bins, nears, fars, jitter, jitter_per_sample, R, S, sbins_out, ebins_out);
  return check_launch("piecewise_bins_kernel");
}
