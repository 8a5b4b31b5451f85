// Library identification and the thread-local error string of the C ABI (include/tn_b200.h).
#include <stdarg.h>

#include "tn_common.cuh"

namespace tn {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Tuning knobs: the environment is read once, on first use (C++11 magic statics: thread-safe, immutable afterwards).
static float env_float(const char* name, float dflt) {
  const char* e = getenv(name);
  return e ? (float)atof(e) : dflt;
}
float agg_threshold_enc() {
  static const float v = env_float("TN_AGG_ENC", 96.f);
  return v;
}
float agg_threshold_enc_patch() {
  static const float v = env_float("TN_AGG_ENC_PATCH", 300.f);
  return v;
}
// smallest level scale from which the gathers use the lane-pair access (finer levels: few duplicate cells in a warp)
float pair_threshold_enc() {
  static const float v = env_float("TN_PAIR_ENC", 150.f);
  return v;
}
// 1: REDs use the lane-pair access on every level
int pair_reds() {
  static const int v = (int)env_float("TN_PAIR_RED", 1.f);
  return v;
}
// levels at least this fine keep their Jacobian when the caller passes jac_out (tn_hash_encode_fwd)
float jac_threshold_enc() {
  static const float v = env_float("TN_JAC_ENC", 400.f);
  return v;
}
float agg_threshold_prop() {
  static const float v = env_float("TN_AGG_PROP", 300.f);
  return v;
}
}  // namespace tn

extern "C" int tn_version(void) { return 100; }
extern "C" const char* tn_last_error_string(void) { return tn::g_err; }
extern "C" const char* tn_build_arch(void) { return "sm_100a"; }
