// tcgen05 / TMEM / mbarrier building blocks for the fused MLP kernels (sm_100a inline PTX).
//
// Operand tiles live in shared memory in the no-swizzle UMMA canonical layout, built from 16-byte chunks
// (8 bf16 along the "feature" axis) so that ONE buffer serves three roles without any transposition:
//
//     element (row r, feature j) of a tile with R rows   ->   byte (j/8) * (R*16) + r*16 + (j%8)*2
//
//   * K-major operand  (rows = M or N, features = K):   core matrix = 8 rows x 16 B,  SBO = 128 B (next 8 rows),
//                                                        LBO = R*16 B (next 8 features)
//   * MN-major operand (features = M or N, rows = K):   core matrix = 8 K-rows x 16 B, SBO = R*16 B (next 8
//                                                        features), LBO = 128 B (next 8 rows)
//
// (bit layouts follow CUTLASS' cute/arch/mma_sm100_desc.hpp: SmemDescriptor / InstrDescriptor.)
#pragma once
#include <cuda_bf16.h>

#include "tn_common.cuh"

namespace tn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (SWIZZLE_NONE, version 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// ---- instruction descriptor: kind::f16, BF16 x BF16 -> F32, dense
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One lane of a CONVERGED warp (call under a warp-uniform condition).  Issuing the MMAs under `warp == 0 &&
// elect_one()` instead of `tid == 0` matters: behind a divergent predicate ptxas wraps every UTCHMMA (a uniform-
// datapath instruction) in its own ELECT / BRA.U.ANY loop, ~13 instructions and a branch per MMA, all of it on the
// tile's critical path (measured: the issuing warp spent a third of the backward kernel in that code).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {  // one full warp
  static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "TMEM columns");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bf16 split: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi); products hi*hi + lo*hi + hi*lo keep
// ~16 mantissa bits (relative error ~2^-16), accumulated in fp32 by the tensor core
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// store 8 consecutive features (one 16-byte chunk) of row r into hi / lo tiles
__device__ __forceinline__ void store_chunk_split(uint8_t* hi_tile, uint8_t* lo_tile, int rows, int chunk, int r,
                                                  const float (&v)[8]) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(v[i], h[i], l[i]);
  const size_t off = (size_t)chunk * rows * 16 + (size_t)r * 16;
  *reinterpret_cast<uint4*>(hi_tile + off) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo_tile + off) = *reinterpret_cast<const uint4*>(l);
}

}  // namespace tc
}  // namespace tn
