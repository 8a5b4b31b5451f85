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
                                            int S, int accumulate, float* __restrict__ dorig,
                                            float* __restrict__ ddir) {
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
    if (accumulate) {  // one warp owns the ray: plain read-modify-write onto the gradient the caller already holds
      ao0 += dorig[3 * r]; ao1 += dorig[3 * r + 1]; ao2 += dorig[3 * r + 2];
      ad0 += ddir[3 * r]; ad1 += ddir[3 * r + 1]; ad2 += ddir[3 * r + 2];
    }
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
__device__ __forceinline__ void sh4_components(float x, float y, float z, float* c) {
  const float xx = x * x, yy = y * y, zz = z * z;
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
}

__global__ void sh4_kernel(const float* __restrict__ d, int64_t N, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float x = __ldg(d + 3 * i), y = __ldg(d + 3 * i + 1), z = __ldg(d + 3 * i + 2);
  float c[16];
  sh4_components(x, y, z, c);
  float4* o4 = reinterpret_cast<float4*>(out + 16 * i);
#pragma unroll
  for (int k = 0; k < 4; ++k) o4[k] = make_float4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
}


// Per-ray inputs of a NerfactoField's colour head in ONE launch (fields/nerfacto_field.py:284-290, 335-344):
// SH basis of the normalised directions (d+1)/2 (fields/base_field.py:136-142: the add and the divide round like the
// reference's two torch ops) and the appearance-embedding row of the ray's camera (field_components/embedding.py:48-55).
__global__ void ray_features_kernel(const float* __restrict__ d, const float* __restrict__ emb,
                                    const int64_t* __restrict__ cam, int64_t R, int E, float* __restrict__ sh_out,
                                    float* __restrict__ emb_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const float x = (__ldg(d + 3 * i) + 1.0f) / 2.0f, y = (__ldg(d + 3 * i + 1) + 1.0f) / 2.0f,
              z = (__ldg(d + 3 * i + 2) + 1.0f) / 2.0f;
  float c[16];
  sh4_components(x, y, z, c);
  float4* o4 = reinterpret_cast<float4*>(sh_out + 16 * i);
#pragma unroll
  for (int k = 0; k < 4; ++k) o4[k] = make_float4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
  if (emb_out) {
    const float* row = emb + cam[i] * E;
    for (int k = 0; k < E; ++k) emb_out[i * E + k] = __ldg(row + k);
  }
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


// ------------------------------------------------------------------------------------------------ ray generation (8f-3)
// PatchPixelSampler.sample_method without a mask (data/pixel_samplers.py:417-438): patch p, pixel (yy, xx) ->
// (floor(u0*num_images), floor(u1*(H-patch) + yy), floor(u2*(W-patch) + xx)); fp32 products as in the reference.
__global__ void patch_indices_kernel(const float* __restrict__ u, int64_t P, int patch, int num_images, int H, int W,
                                     int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int pp = patch * patch;
  if (i >= P * pp) return;
  const int64_t p = i / pp;
  const int k = (int)(i - p * pp), yy = k / patch, xx = k - yy * patch;
  const float c = __ldg(u + 3 * p) * (float)num_images;
  const float y = __ldg(u + 3 * p + 1) * (float)(H - patch) + (float)yy;
  const float x = __ldg(u + 3 * p + 2) * (float)(W - patch) + (float)xx;
  out[3 * i] = (int64_t)floorf(c);
  out[3 * i + 1] = (int64_t)floorf(y);
  out[3 * i + 2] = (int64_t)floorf(x);
}

// collate (pixel_samplers.py:239-256): image[r] = images[c, y, x, :], is_thermal[r] = is_thermal_per_image[c],
// indices[r, 0] = image_idx[c]
template <typename PIX>
__global__ void gather_pixels_kernel(const PIX* __restrict__ images, int64_t n_img, int H, int W, int C,
                                     int64_t* __restrict__ indices, const int64_t* __restrict__ image_idx,
                                     const float* __restrict__ thermal_per_image, int64_t R, float* __restrict__ image_out,
                                     float* __restrict__ thermal_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t c = indices[3 * r], y = indices[3 * r + 1], x = indices[3 * r + 2];
  const PIX* src = images + ((c * H + y) * W + x) * C;
  for (int k = 0; k < C; ++k) {
    if constexpr (sizeof(PIX) == 1) image_out[r * C + k] = (float)src[k] / 255.0f;
    else image_out[r * C + k] = (float)src[k];
  }
  if (thermal_out) thermal_out[r] = thermal_per_image ? __ldg(thermal_per_image + c) : 0.f;
  if (image_idx) indices[3 * r] = __ldg(image_idx + c);
}

// RayGenerator.forward + Cameras._generate_rays_from_coords for undistorted PERSPECTIVE cameras
// (model_components/ray_generators.py:40-55, cameras/cameras.py:504-905): pixel centre (y+0.5, x+0.5),
// camera-space directions for the pixel and its +1 x / +1 y neighbours, rotation by c2w, normalisation
// (camera_utils.py:286-298), pixel_area = |d - d_x| * |d - d_y|.
__global__ void generate_rays_kernel(const int64_t* __restrict__ indices, const float* __restrict__ c2w,
                                     const float* __restrict__ intr, int64_t R, float* __restrict__ origins,
                                     float* __restrict__ directions, float* __restrict__ pixel_area,
                                     float* __restrict__ dir_norm, int64_t* __restrict__ cam_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t c = indices[3 * r];
  const float y = (float)indices[3 * r + 1] + 0.5f, x = (float)indices[3 * r + 2] + 0.5f;
  const float fx = __ldg(intr + 4 * c), fy = __ldg(intr + 4 * c + 1), cx = __ldg(intr + 4 * c + 2),
              cy = __ldg(intr + 4 * c + 3);
  const float* m = c2w + 12 * c;  // [3,4] row-major
  const float eps = 8.881784197001252e-16f;  // np.finfo(float).eps * 4
  // the three image-plane points; the y axis flips from OpenCV to OpenGL (cameras.py:654)
  const float px[3] = {(x - cx) / fx, (x - cx + 1.f) / fx, (x - cx) / fx};
  const float py[3] = {-((y - cy) / fy), -((y - cy) / fy), -((y - cy + 1.f) / fy)};
  float d[3][3];
  float n0 = 0.f;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    float v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = (px[q] * m[4 * i] + py[q] * m[4 * i + 1]) + (-1.f) * m[4 * i + 2];
    const float nrm = fmaxf(sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]), eps);
    if (q == 0) n0 = nrm;
#pragma unroll
    for (int i = 0; i < 3; ++i) d[q][i] = v[i] / nrm;
  }
  float dx2 = 0.f, dy2 = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = d[0][i] - d[1][i], b = d[0][i] - d[2][i];
    dx2 += a * a;
    dy2 += b * b;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    origins[3 * r + i] = m[4 * i + 3];
    directions[3 * r + i] = d[0][i];
  }
  pixel_area[r] = sqrtf(dx2) * sqrtf(dy2);
  if (dir_norm) dir_norm[r] = n0;
  if (cam_out) cam_out[r] = c;
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
                                       const float* dx, int64_t R, int S, int accumulate, float* d_origins,
                                       float* d_directions, void* stream) {
  TN_REQUIRE(origins && directions && ebins && dx && d_origins && d_directions, TN_EINVAL,
             "sample_positions_bwd: null pointer");
  TN_REQUIRE(R >= 0 && S >= 1, TN_EINVAL, "sample_positions_bwd: bad R=%lld S=%d", (long long)R, S);
  if (R == 0) return TN_OK;
  sample_positions_bwd_kernel<<<blocks_for(R * 32, 128), 128, 0, (cudaStream_t)stream>>>(origins, directions, ebins, dx,
                                                                                        R, S, accumulate, d_origins,
                                                                                        d_directions);
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

extern "C" int tn_ray_features(const float* directions, const float* embedding, const int64_t* camera_indices,
                               int64_t R, int emb_dim, float* sh_out, float* emb_out, void* stream) {
  TN_REQUIRE(directions && sh_out, TN_EINVAL, "ray_features: null pointer");
  TN_REQUIRE(!emb_out || (embedding && camera_indices && emb_dim >= 1), TN_EINVAL,
             "ray_features: emb_out needs the embedding table, the camera indices and emb_dim >= 1");
  if (R <= 0) return R == 0 ? TN_OK : TN_EINVAL;
  ray_features_kernel<<<blocks_for(R, 128), 128, 0, (cudaStream_t)stream>>>(directions, embedding, camera_indices, R,
                                                                            emb_dim, sh_out, emb_out);
  return check_launch("ray_features_kernel");
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

extern "C" int tn_patch_pixel_indices(const float* u, int64_t num_patches, int patch_size, int num_images,
                                      int image_height, int image_width, int64_t* indices_out, void* stream) {
  TN_REQUIRE(num_patches >= 0 && patch_size >= 1 && num_images >= 1 && image_height >= patch_size &&
                 image_width >= patch_size, TN_EINVAL, "patch_pixel_indices: bad sizes");
  if (num_patches == 0) return TN_OK;
  TN_REQUIRE(u && indices_out, TN_EINVAL, "patch_pixel_indices: null pointer");
  const int64_t n = num_patches * patch_size * patch_size;
  patch_indices_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(u, num_patches, patch_size, num_images,
                                                                           image_height, image_width, indices_out);
  return check_launch("patch_indices_kernel");
}

extern "C" int tn_gather_pixels(const void* images, int pixel_dtype, int64_t num_images, int image_height,
                                int image_width, int channels, int64_t* indices, const int64_t* image_idx,
                                const float* is_thermal_per_image, int64_t R, float* image_out, float* is_thermal_out,
                                void* stream) {
  TN_REQUIRE(R >= 0 && num_images >= 1 && image_height >= 1 && image_width >= 1 && channels >= 1 && channels <= 4,
             TN_EINVAL, "gather_pixels: bad sizes");
  TN_REQUIRE(pixel_dtype == 0 || pixel_dtype == 1, TN_EINVAL, "gather_pixels: pixel_dtype=%d (0 float32, 1 uint8)",
             pixel_dtype);
  if (R == 0) return TN_OK;
  TN_REQUIRE(images && indices && image_out, TN_EINVAL, "gather_pixels: null pointer");
  if (pixel_dtype == 0)
    gather_pixels_kernel<float><<<blocks_for(R, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)images, num_images, image_height, image_width, channels, indices, image_idx,
        is_thermal_per_image, R, image_out, is_thermal_out);
  else
    gather_pixels_kernel<uint8_t><<<blocks_for(R, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)images, num_images, image_height, image_width, channels, indices, image_idx,
        is_thermal_per_image, R, image_out, is_thermal_out);
  return check_launch("gather_pixels_kernel");
}

extern "C" int tn_generate_rays(const int64_t* ray_indices, const float* camera_to_worlds, const float* intrinsics,
                                int64_t num_cameras, int64_t R, float* origins_out, float* directions_out,
                                float* pixel_area_out, float* directions_norm_out, int64_t* camera_indices_out,
                                void* stream) {
  TN_REQUIRE(R >= 0 && num_cameras >= 1, TN_EINVAL, "generate_rays: bad sizes");
  if (R == 0) return TN_OK;
  TN_REQUIRE(ray_indices && camera_to_worlds && intrinsics && origins_out && directions_out && pixel_area_out,
             TN_EINVAL, "generate_rays: null pointer");
  generate_rays_kernel<<<blocks_for(R, 256), 256, 0, (cudaStream_t)stream>>>(
      ray_indices, camera_to_worlds, intrinsics, R, origins_out, directions_out, pixel_area_out, directions_norm_out,
      camera_indices_out);
  return check_launch("generate_rays_kernel");
}
