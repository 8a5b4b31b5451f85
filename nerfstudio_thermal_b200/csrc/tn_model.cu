// Per-ray model glue as single launches: camera-pose correction of the ray bundle and the pixel / density losses.
//
//  * camera_opt: CameraOptimizer.apply_to_raybundle with mode SO3xR3 (cameras/camera_optimizers.py:132-176,
//    cameras/lie_groups.py:24-59): o' = o + t_c, d' = R(w_c) d with the Rodrigues map, identity for frozen cameras;
//    backward reduces dL/d(o', d') into dL/d(pose_adjustment).  The torch expression is ~40 tiny kernels forward and
//    ~60 backward per ray bundle.
//  * pixel_losses: rgb / thermal MSE, pixel-wise thermal TV and cross-channel gradient losses of
//    ThermalNerfactoModel.get_loss_dict (models/thermal_nerfacto.py:286-354, model_components/losses.py:603-651,
//    utils/rgbt_utils.py:6-33) for patch-ordered batches, forward values and gradients.
//  * density_l1: the four asymmetric L1 terms of the cross-field density regulariser (thermal_nerfacto.py:328-344).
#include "tn_common.cuh"

namespace tn {

__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// ------------------------------------------------------------------------------------------------ camera optimizer
__global__ void camera_opt_fwd_kernel(const float* __restrict__ pose, const uint8_t* __restrict__ frozen,
                                      const int64_t* __restrict__ cam, const float* __restrict__ o,
                                      const float* __restrict__ d, int64_t R, int shared_pose, float* __restrict__ oo,
                                      float* __restrict__ dd) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t c = cam[r];
  float o3[3] = {o[3 * r], o[3 * r + 1], o[3 * r + 2]};
  float d3[3] = {d[3 * r], d[3 * r + 1], d[3 * r + 2]};
  if (!(frozen && frozen[c])) {
    const float* p = pose + (shared_pose ? 0 : c * 6);
    const float w[3] = {p[3], p[4], p[5]};
    const float th = sqrtf(fmaxf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2], 1e-4f));
    const float inv = 1.f / th;
    const float a = inv * sinf(th), b = inv * inv * (1.f - cosf(th));
    float u[3], v[3];
    cross3(w, d3, u);
    cross3(w, u, v);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o3[k] += p[k];
      d3[k] = d3[k] + a * u[k] + b * v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    oo[3 * r + k] = o3[k];
    dd[3 * r + k] = d3[k];
  }
}

// dpose[c] += (d_o', dL/dw) ; atomics (4096 rays onto <= a few hundred cameras)
__global__ void camera_opt_bwd_kernel(const float* __restrict__ pose, const uint8_t* __restrict__ frozen,
                                      const int64_t* __restrict__ cam, const float* __restrict__ d,
                                      const float* __restrict__ g_o, const float* __restrict__ g_d, int64_t R,
                                      int shared_pose, float* __restrict__ dpose) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t c = cam[r];
  if (frozen && frozen[c]) return;
  const int64_t row = shared_pose ? 0 : c;
  const float* p = pose + row * 6;
  const float w[3] = {p[3], p[4], p[5]};
  const float d3[3] = {d[3 * r], d[3 * r + 1], d[3 * r + 2]};
  const float g[3] = {g_d ? g_d[3 * r] : 0.f, g_d ? g_d[3 * r + 1] : 0.f, g_d ? g_d[3 * r + 2] : 0.f};
  const float n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const float th = sqrtf(fmaxf(n2, 1e-4f));
  const float inv = 1.f / th;
  const float s = sinf(th), co = cosf(th);
  const float a = inv * s, b = inv * inv * (1.f - co);
  float u[3], v[3], gxw[3], h[3], dxh[3], uxg[3];
  cross3(w, d3, u);
  cross3(w, u, v);
  const float dLda = g[0] * u[0] + g[1] * u[1] + g[2] * u[2];
  const float dLdb = g[0] * v[0] + g[1] * v[1] + g[2] * v[2];
  cross3(g, w, gxw);
#pragma unroll
  for (int k = 0; k < 3; ++k) h[k] = a * g[k] + b * gxw[k];  // dL/du
  cross3(d3, h, dxh);                                        // via u = w x d
  cross3(u, g, uxg);                                         // via v = w x u (direct w dependence)
  float dth = 0.f;
  if (n2 > 1e-4f) {
    const float da = (th * co - s) * inv * inv;                       // d(sin t / t)/dt
    const float db = (th * s - 2.f * (1.f - co)) * inv * inv * inv;   // d((1-cos t)/t^2)/dt
    dth = dLda * da + dLdb * db;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (g_o) atomicAdd(dpose + row * 6 + k, g_o[3 * r + k]);
    atomicAdd(dpose + row * 6 + 3 + k, dxh[k] + b * uxg[k] + dth * w[k] * inv);
  }
}

// ------------------------------------------------------------------------------------------------ pixel losses
// losses[0..3] = rgb MSE, thermal MSE (unscaled), tv_pixel, cross_channel.  One CTA; R is a few thousand.
// mults = upstream multipliers of the four terms for the gradient (d_rgb, d_thermal); values are un-multiplied.
constexpr int kLossThreads = 1024;

__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(kLossThreads) pixel_losses_kernel(
    const float* __restrict__ rgb, const float* __restrict__ thermal, const float* __restrict__ image,
    const float* __restrict__ is_thermal, int64_t R, const float* __restrict__ upstream, float* __restrict__ losses,
    float* __restrict__ d_rgb, float* __restrict__ d_thermal) {
  __shared__ float red[5][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_thermal = thermal != nullptr;
  // pass 1 needs the number of RGB patches for the masked means
  float n_rgb_patch = 0.f;
  for (int64_t q = tid; q < R / 4; q += kLossThreads) n_rgb_patch += 1.f - is_thermal[4 * q];
  n_rgb_patch = warp_sum(n_rgb_patch);
  if (lane == 0) red[4][warp] = n_rgb_patch;
  __syncthreads();
  float npatch = 0.f;
  for (int i = 0; i < kLossThreads / 32; ++i) npatch += red[4][i];
  const float up0 = upstream ? upstream[0] : 0.f, up1 = upstream ? upstream[1] : 0.f;
  const float up2 = upstream ? upstream[2] : 0.f, up3 = upstream ? upstream[3] : 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t q = tid; q < R / 4; q += kLossThreads) {  // one 2x2 patch (four consecutive rays of one camera)
    float pt[4], gray[4];
    float th = is_thermal[4 * q];
    const float isrgb = 1.f - th;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = 4 * q + k;
      const float t_r = is_thermal[r], rgb_r = 1.f - t_r;
      float gsum = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // gt = image * is_rgb (rgb_to_rgbt_image), both sides masked by is_rgb again (thermal_nerfacto.py:315-318)
        const float gt = image[3 * r + c] * rgb_r;
        const float diff = gt * rgb_r - rgb[3 * r + c] * rgb_r;
        acc[0] += diff * diff;
        if (d_rgb) d_rgb[3 * r + c] = up0 * (-2.f * diff * rgb_r) / (float)(3 * R);
        gsum += gt;
      }
      gray[k] = gsum / 3.f;
      pt[k] = has_thermal ? thermal[r] : 0.f;
      if (has_thermal) {
        const float gtT = image[3 * r] * t_r;  // thermal frames carry T in channel 0 (rgbt_utils.py:31)
        const float diff = gtT * t_r - pt[k] * t_r;
        acc[1] += diff * diff;
        if (d_thermal) d_thermal[r] = up1 * (-2.f * diff * t_r) / (float)R;
      }
    }
    if (has_thermal) {
      // tv (losses.py:603-620) and cross-channel (losses.py:623-651) on RGB-camera patches
      const float e01 = pt[0] - pt[1], e02 = pt[0] - pt[2], e13 = pt[1] - pt[3], e23 = pt[2] - pt[3];
      acc[2] += isrgb * (fabsf(e01) + fabsf(e02) + fabsf(e13) + fabsf(e23));
      const float c0 = (pt[1] - pt[0]) - (gray[1] - gray[0]), c1 = (pt[2] - pt[0]) - (gray[2] - gray[0]);
      const float c2 = (pt[3] - pt[1]) - (gray[3] - gray[1]), c3 = (pt[3] - pt[2]) - (gray[3] - gray[2]);
      acc[3] += isrgb * (fabsf(c0) + fabsf(c1) + fabsf(c2) + fabsf(c3));
      if (d_thermal && npatch > 0.f) {
        const float kt = up2 * 0.25f * isrgb / npatch, kc = up3 * 0.25f * isrgb / npatch;
        d_thermal[4 * q + 0] += kt * (sgn(e01) + sgn(e02)) + kc * (-sgn(c0) - sgn(c1));
        d_thermal[4 * q + 1] += kt * (-sgn(e01) + sgn(e13)) + kc * (sgn(c0) - sgn(c2));
        d_thermal[4 * q + 2] += kt * (-sgn(e02) + sgn(e23)) + kc * (sgn(c1) - sgn(c3));
        d_thermal[4 * q + 3] += kt * (-sgn(e13) - sgn(e23)) + kc * (sgn(c2) + sgn(c3));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (tid < 4) {
    float v = 0.f;
    for (int i = 0; i < kLossThreads / 32; ++i) v += red[tid][i];
    if (tid == 0) v /= (float)(3 * R);
    else if (tid == 1) v /= (float)R;
    else v = npatch > 0.f ? 0.25f * v / npatch : 0.f;
    if (losses) losses[tid] = v;
  }
}

// ------------------------------------------------------------------------------------------------ density L1
// loss = m*[ L1(d2*, dt) + L1(d*, d2t) ] + r*m*[ L1(d2, dt*) + L1(d, d2t*) ]   (* = detached), L1 = mean |a-b|
// => d(loss)/d(dt) = -m sgn(d2-dt)/N, d/d(d2t) = -m sgn(d-d2t)/N, d/d(d2) = r m sgn(d2-dt)/N, d/d(d) = r m sgn(d-d2t)/N
__global__ void density_l1_kernel(const float* __restrict__ d, const float* __restrict__ d2, const float* __restrict__ dt,
                                  const float* __restrict__ d2t, int64_t N, float vm, float m, float rm,
                                  const float* __restrict__ upstream, float* __restrict__ partial,
                                  unsigned int* __restrict__ ticket, float* __restrict__ g_d, float* __restrict__ g_d2,
                                  float* __restrict__ g_dt, float* __restrict__ g_d2t) {
  float acc = 0.f;
  const float invn = 1.f / (float)N;
  const float up = upstream ? __ldg(upstream) : 1.f;  // gradients = upstream * d(loss)/d(.) (one launch, no scaling pass)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = d2[i] - dt[i], b = d[i] - d2t[i];
    acc += vm * (fabsf(a) + fabsf(b));
    if (g_d) {
      g_dt[i] = -m * sgn(a) * invn * up;
      g_d2t[i] = -m * sgn(b) * invn * up;
      g_d2[i] = rm * sgn(a) * invn * up;
      g_d[i] = rm * sgn(b) * invn * up;
    }
  }
  if (!partial) return;
  acc = warp_sum(acc);
  __shared__ float red[32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[i];
    partial[blockIdx.x] = v * invn;
    last = false;
    if (ticket) {  // the CTA that finishes last adds the partial sums up, in index order (deterministic)
      __threadfence();
      last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    if (last) {
      __threadfence();
      float t = 0.f;
      for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partial + b);
      partial[gridDim.x] = t;
      *ticket = 0u;  // ready for the next launch
    }
  }
}


// ------------------------------------------------------------------------------------------------ camera regulariser
// cameras/camera_optimizers.py:188-204: out[0] = (mean_i |t_i| * trans_pen + mean_i |w_i| * rot_pen) * scale,
// out[1] = |T|_F, out[2] = |W|_F (the two metrics).  One CTA: the pose table is a few hundred rows.
__global__ void __launch_bounds__(256) camera_reg_fwd_kernel(const float* __restrict__ pose, int C, float trans_pen,
                                                             float rot_pen, float scale, float* __restrict__ out) {
  __shared__ float red[4][8];
  float st = 0.f, sw = 0.f, qt = 0.f, qw = 0.f;
  for (int i = threadIdx.x; i < C; i += 256) {
    const float* p = pose + 6 * i;
    const float t2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2], w2 = p[3] * p[3] + p[4] * p[4] + p[5] * p[5];
    st += sqrtf(t2); sw += sqrtf(w2); qt += t2; qw += w2;
  }
  st = warp_sum(st); sw = warp_sum(sw); qt = warp_sum(qt); qw = warp_sum(qw);
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    red[0][w] = st; red[1][w] = sw; red[2][w] = qt; red[3][w] = qw;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; d += red[3][w]; }
    out[0] = (a / (float)C * trans_pen + b / (float)C * rot_pen) * scale;
    out[1] = sqrtf(c);
    out[2] = sqrtf(d);
  }
}
// dpose[i] (overwritten, or atomically added onto) = g * scale / C * (trans_pen * t_i/|t_i| , rot_pen * w_i/|w_i|), zero
// where the norm is zero
__global__ void camera_reg_bwd_kernel(const float* __restrict__ pose, const float* __restrict__ g, int C,
                                      float trans_pen, float rot_pen, float scale, int accumulate,
                                      float* __restrict__ dpose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float* p = pose + 6 * i;
  const float gs = g[0] * scale / (float)C;
  const float nt = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  const float nw = sqrtf(p[3] * p[3] + p[4] * p[4] + p[5] * p[5]);
  const float ct = nt > 0.f ? gs * trans_pen / nt : 0.f, cw = nw > 0.f ? gs * rot_pen / nw : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (accumulate) {  // the bundle's pose gradient (camera_opt_bwd_kernel) lands in the same buffer, on another stream
      atomicAdd(dpose + 6 * i + k, ct * p[k]);
      atomicAdd(dpose + 6 * i + 3 + k, cw * p[3 + k]);
    } else {
      dpose[6 * i + k] = ct * p[k];
      dpose[6 * i + 3 + k] = cw * p[3 + k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ appearance embedding
// Gradient of the per-camera appearance embedding from the colour head's per-ray first-layer gradient:
// d emb[cam[r], j] += sum_k dz1_ray[r, k] * W0[k, col0 + j]  (the [R,64] x [64,E] product and the index_add_ of
// field_components/embedding.py's autograd in one launch).  A warp owns kRaysPerWarp consecutive rays, lane = j;
// consecutive rays of a patch-ordered batch share their camera, so the warp adds them up and issues one atomic row per
// run.  dz1_ray rows are cleared after use when `clear` is set (the caller keeps one self-cleaning buffer).
constexpr int kEmbRaysPerWarp = 4;  // one 2x2 patch: the rays of a patch share their camera
template <int WIDTH>
__global__ void __launch_bounds__(128) embed_bwd_kernel(float* __restrict__ dz1_ray, const float* __restrict__ w0,
                                                        const int64_t* __restrict__ cam, int64_t R, int in_dim,
                                                        int col0, int E, int clear, float* __restrict__ dweight) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r0 = warp * kEmbRaysPerWarp;
  if (r0 >= R) return;
  const int64_t r1 = min(r0 + (int64_t)kEmbRaysPerWarp, R);
  for (int j0 = 0; j0 < E; j0 += 32) {
    const int j = min(j0 + lane, E - 1);  // lanes past E compute a duplicate column and do not store
    float wcol[WIDTH];                    // this lane's column of W0: loaded once, all loads in flight together
#pragma unroll
    for (int k = 0; k < WIDTH; ++k) wcol[k] = __ldg(w0 + (int64_t)k * in_dim + col0 + j);
    float acc = 0.f;
    int64_t run_cam = cam[r0];
    for (int64_t r = r0; r < r1; ++r) {
      const int64_t c = cam[r];
      if (c != run_cam) {
        if (j0 + lane < E) atomicAdd(dweight + run_cam * E + j, acc);
        acc = 0.f;
        run_cam = c;
      }
      const float4* z4 = reinterpret_cast<const float4*>(dz1_ray + r * WIDTH);  // same address in every lane: broadcast
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < WIDTH / 4; ++k) {
        const float4 z = z4[k];
        t += z.x * wcol[4 * k] + z.y * wcol[4 * k + 1] + z.z * wcol[4 * k + 2] + z.w * wcol[4 * k + 3];
      }
      acc += t;
    }
    if (j0 + lane < E) atomicAdd(dweight + run_cam * E + j, acc);
  }
  if (clear) {
    __syncwarp();
    float4* z4 = reinterpret_cast<float4*>(dz1_ray + r0 * WIDTH);
    for (int64_t i = lane; i < (r1 - r0) * (WIDTH / 4); i += 32) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ loss assembly
// out[0] = sum_k scale[k] * *term[k]; out[1 + j] = sum over the terms of dictionary entry j (slot[k] == j)
// (K scalar loss terms living anywhere on the device)
struct TermPtrs {
  const float* p[TN_MAX_LOSS_TERMS];
  float scale[TN_MAX_LOSS_TERMS];
  int slot[TN_MAX_LOSS_TERMS];
};
__global__ void loss_sum_kernel(TermPtrs t, int K, int n_slots, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int j = 0; j < n_slots; ++j) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k)
        if (t.slot[k] == j) acc += t.scale[k] * t.p[k][0];
      out[1 + j] = acc;
      total += acc;
    }
    out[0] = total;
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_camera_opt_fwd(const float* pose, const uint8_t* frozen, const int64_t* camera_indices,
                                 const float* origins, const float* directions, int64_t R, int shared_pose,
                                 float* origins_out, float* directions_out, void* stream) {
  TN_REQUIRE(pose && camera_indices && origins && directions && origins_out && directions_out, TN_EINVAL,
             "camera_opt_fwd: null pointer");
  if (R <= 0) return R == 0 ? TN_OK : TN_EINVAL;
  camera_opt_fwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pose, frozen, camera_indices, origins, directions, R, shared_pose, origins_out, directions_out);
  return check_launch("camera_opt_fwd_kernel");
}

extern "C" int tn_camera_opt_bwd(const float* pose, const uint8_t* frozen, const int64_t* camera_indices,
                                 const float* directions, const float* d_origins_out, const float* d_directions_out,
                                 int64_t R, int shared_pose, float* dpose, void* stream) {
  TN_REQUIRE(pose && camera_indices && directions && dpose, TN_EINVAL, "camera_opt_bwd: null pointer");
  if (R <= 0) return R == 0 ? TN_OK : TN_EINVAL;
  camera_opt_bwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pose, frozen, camera_indices, directions, d_origins_out, d_directions_out, R, shared_pose, dpose);
  return check_launch("camera_opt_bwd_kernel");
}

extern "C" int tn_pixel_losses(const float* rgb, const float* thermal, const float* image, const float* is_thermal,
                               int64_t R, const float* upstream, float* losses_out, float* d_rgb_out,
                               float* d_thermal_out, void* stream) {
  TN_REQUIRE(rgb && image && is_thermal, TN_EINVAL, "pixel_losses: null pointer");
  TN_REQUIRE(R >= 0 && R % 4 == 0, TN_EINVAL, "pixel_losses: R=%lld must be a multiple of 4 (2x2 patches)", (long long)R);
  TN_REQUIRE(!(d_rgb_out || d_thermal_out) || upstream, TN_EINVAL, "pixel_losses: gradients need upstream multipliers");
  if (R == 0) return TN_OK;
  pixel_losses_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(rgb, thermal, image, is_thermal, R, upstream,
                                                                  losses_out, d_rgb_out, d_thermal_out);
  return check_launch("pixel_losses_kernel");
}

extern "C" int tn_density_l1(const float* d, const float* d2, const float* dt, const float* d2t, int64_t N,
                             float value_mult, float thermal_grad_mult, float rgb_grad_mult, const float* upstream_dev,
                             float* partial_out, int n_partial, uint32_t* ticket, float* g_d, float* g_d2, float* g_dt,
                             float* g_d2t, void* stream) {
  TN_REQUIRE(d && d2 && dt && d2t && n_partial >= 1 && n_partial <= 1024, TN_EINVAL, "density_l1: bad arguments");
  TN_REQUIRE(partial_out || g_d, TN_EINVAL, "density_l1: nothing to compute");
  TN_REQUIRE(!ticket || partial_out, TN_EINVAL, "density_l1: a ticket counter without partial_out");
  TN_REQUIRE((g_d != nullptr) == (g_d2 != nullptr) && (g_d != nullptr) == (g_dt != nullptr) &&
                 (g_d != nullptr) == (g_d2t != nullptr), TN_EINVAL, "density_l1: give all four gradient outputs or none");
  if (N <= 0) return N == 0 ? TN_OK : TN_EINVAL;
  density_l1_kernel<<<n_partial, 256, 0, (cudaStream_t)stream>>>(d, d2, dt, d2t, N, value_mult, thermal_grad_mult,
                                                               rgb_grad_mult, upstream_dev, partial_out, ticket, g_d,
                                                               g_d2, g_dt, g_d2t);
  return check_launch("density_l1_kernel");
}

extern "C" int tn_camera_reg_fwd(const float* pose, int num_cameras, float trans_penalty, float rot_penalty,
                                 float penalty_scale, float* out3, void* stream) {
  TN_REQUIRE(pose && out3 && num_cameras >= 1, TN_EINVAL, "camera_reg_fwd: bad arguments");
  camera_reg_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pose, num_cameras, trans_penalty, rot_penalty,
                                                            penalty_scale, out3);
  return check_launch("camera_reg_fwd_kernel");
}

extern "C" int tn_camera_reg_bwd(const float* pose, const float* upstream_dev, int num_cameras, float trans_penalty,
                                 float rot_penalty, float penalty_scale, int accumulate, float* dpose_out,
                                 void* stream) {
  TN_REQUIRE(pose && upstream_dev && dpose_out && num_cameras >= 1, TN_EINVAL, "camera_reg_bwd: bad arguments");
  camera_reg_bwd_kernel<<<(num_cameras + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      pose, upstream_dev, num_cameras, trans_penalty, rot_penalty, penalty_scale, accumulate, dpose_out);
  return check_launch("camera_reg_bwd_kernel");
}

extern "C" int tn_embed_bwd(float* dz1_ray, const float* w0, const int64_t* camera_indices, int64_t R, int width,
                            int in_dim, int col0, int emb_dim, int clear, float* dweight, void* stream) {
  TN_REQUIRE(dz1_ray && w0 && camera_indices && dweight, TN_EINVAL, "embed_bwd: null pointer");
  TN_REQUIRE(width == 64 && emb_dim >= 1 && col0 >= 0 && col0 + emb_dim <= in_dim, TN_EINVAL,
             "embed_bwd: bad shape width=%d (64) in_dim=%d col0=%d emb_dim=%d", width, in_dim, col0, emb_dim);
  if (R <= 0) return R == 0 ? TN_OK : TN_EINVAL;
  const int64_t warps = (R + kEmbRaysPerWarp - 1) / kEmbRaysPerWarp;
  embed_bwd_kernel<64><<<(unsigned)((warps + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
      dz1_ray, w0, camera_indices, R, in_dim, col0, emb_dim, clear, dweight);
  return check_launch("embed_bwd_kernel");
}

extern "C" int tn_loss_sum(const float* const* term_host_ptrs, const float* scale_host, const int* slot_host,
                           int n_terms, int n_slots, float* out, void* stream) {
  TN_REQUIRE(term_host_ptrs && scale_host && slot_host && out && n_terms >= 1 && n_terms <= TN_MAX_LOSS_TERMS &&
                 n_slots >= 1 && n_slots <= n_terms,
             TN_EINVAL, "loss_sum: n_terms=%d (max %d) n_slots=%d", n_terms, TN_MAX_LOSS_TERMS, n_slots);
  TermPtrs t = {};
  for (int k = 0; k < n_terms; ++k) {
    TN_REQUIRE(term_host_ptrs[k] && slot_host[k] >= 0 && slot_host[k] < n_slots, TN_EINVAL,
               "loss_sum: term %d is null or has slot outside [0,%d)", k, n_slots);
    t.p[k] = term_host_ptrs[k];
    t.scale[k] = scale_host[k];
    t.slot[k] = slot_host[k];
  }
  loss_sum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, n_terms, n_slots, out);
  return check_launch("loss_sum_kernel");
}
