// Device helpers shared by the geometry kernels and the fused field kernels.  Translation units that include
// this header are compiled with -fmad=false: every operation rounds like the reference's separate torch ops
// (the resulting grid coordinates feed floor()/ceil() in the hash encoder and must not drift by an ulp).
#pragma once
#include "tn_common.cuh"

namespace tn {

// Frustums.get_positions (cameras/rays.py:55): o + d * (start + end) / 2
__device__ __forceinline__ float sample_pos(float o, float d, float st, float en) { return o + (d * (st + en)) / 2.f; }

// SceneContraction(order=inf) (field_components/spatial_distortions.py:66-69), then (p+2)/4, selector and
// zeroing (fields/nerfacto_field.py:207-215).  Returns selector.
__device__ __forceinline__ float contract_normalise(float& p0, float& p1, float& p2) {
  const float mag = fmaxf(fabsf(p0), fmaxf(fabsf(p1), fabsf(p2)));
  if (!(mag < 1.f)) {
    const float k = 2.f - (1.f / mag);
    p0 = k * (p0 / mag); p1 = k * (p1 / mag); p2 = k * (p2 / mag);
  }
  p0 = (p0 + 2.f) / 4.f; p1 = (p1 + 2.f) / 4.f; p2 = (p2 + 2.f) / 4.f;
  const bool in = (p0 > 0.f) && (p0 < 1.f) && (p1 > 0.f) && (p1 < 1.f) && (p2 > 0.f) && (p2 < 1.f);
  const float sel = in ? 1.f : 0.f;
  p0 *= sel; p1 *= sel; p2 *= sel;
  return sel;
}

// gradient of contract_normalise w.r.t. the un-contracted position (autograd-equivalent, ties of the
// max-norm share the gradient like torch.linalg.vector_norm's backward)
__device__ __forceinline__ void contract_normalise_bwd(float p0, float p1, float p2, float& g0, float& g1, float& g2) {
  float q0 = p0, q1 = p1, q2 = p2;
  const float sel = contract_normalise(q0, q1, q2);
  g0 = g0 * sel / 4.f; g1 = g1 * sel / 4.f; g2 = g2 * sel / 4.f;
  const float a0 = fabsf(p0), a1 = fabsf(p1), a2 = fabsf(p2);
  const float m = fmaxf(a0, fmaxf(a1, a2));
  if (m < 1.f) return;
  const float k = 2.f - 1.f / m;
  const float inv = 1.f / m;
  const float dk = g0 * (p0 * inv) + g1 * (p1 * inv) + g2 * (p2 * inv);
  const float dq0 = g0 * k, dq1 = g1 * k, dq2 = g2 * k;
  const float dm = dk * inv * inv - (dq0 * p0 + dq1 * p1 + dq2 * p2) * inv * inv;
  const float t0 = a0 == m ? 1.f : 0.f, t1 = a1 == m ? 1.f : 0.f, t2 = a2 == m ? 1.f : 0.f;
  const float share = dm / (t0 + t1 + t2);
  g0 = dq0 * inv + share * t0 * copysignf(1.f, p0);
  g1 = dq1 * inv + share * t1 * copysignf(1.f, p1);
  g2 = dq2 * inv + share * t2 * copysignf(1.f, p2);
}

}  // namespace tn
