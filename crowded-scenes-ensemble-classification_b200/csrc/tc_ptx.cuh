// PTX wrappers shared by the tcgen05 kernels (conv_tc.cu: single-CTA modes, conv_tc2.cu: CTA-pair mode).
#pragma once
#include "common.cuh"

namespace cse {

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("cse conv_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// Warp-uniform issue: all 32 lanes of the producer / MMA warp run the loops with warp-uniform
// operands (so ptxas keeps them in uniform registers - no per-instruction R2UR waterfall), and the
// asynchronous instruction itself is predicated on the lane chosen once by elect.sync.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_expect_tx_p(uint32_t pred, uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
               "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes), "r"(pred) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t pred, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %8, 0;\n\t"
      "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t pred, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t pred, uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(pred)
               : "memory");
}
// NK consecutive K=16 MMAs of one pipeline stage chunk in ONE asm block: the descriptors only differ in their
// 14-bit start address (+2 per 32-byte K step), so they are rebuilt from a 32-bit low word and the constant
// high word inside the block.  The MMA warp then spends ~3 instructions per MMA instead of ~12 (64-bit adds,
// vector->uniform moves and predicate votes per call), which matters for N = 64 tiles (32-48 cycles each).
#define CSE_MMA_STEP                                                        \
  "add.u32 al, al, 2;\n\tadd.u32 bl, bl, 2;\n\t"                           \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                     \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t"
template <int NK>
__device__ __forceinline__ void tc_mma_k(uint32_t pred, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi32,
                                         uint32_t idesc, uint32_t accumulate_first) {
  static_assert(NK == 1 || NK == 2 || NK == 4, "K steps per stage chunk");
#define CSE_MMA_HEAD                                                        \
  "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"   \
  "setp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t" \
  "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"                                 \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                     \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
  if (NK == 1)
    asm volatile(CSE_MMA_HEAD "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi32), "r"(idesc), "r"(accumulate_first), "r"(pred) : "memory");
  else if (NK == 2)
    asm volatile(CSE_MMA_HEAD CSE_MMA_STEP "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi32), "r"(idesc), "r"(accumulate_first), "r"(pred) : "memory");
  else
    asm volatile(CSE_MMA_HEAD CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_STEP "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi32), "r"(idesc), "r"(accumulate_first), "r"(pred) : "memory");
#undef CSE_MMA_HEAD
}
// Four taps x NK K-steps (the kh taps of an h-halo stage: tap t reads A at a_lo + t*a_step, B at b_lo + t*b_step)
// in one asm block: 4*NK back-to-back MMAs with one operand set-up.
#define CSE_MMA_TAP(FIRST_PRED)                                                   \
  "mov.b32 al, ab;\n\tmov.b32 bl, bb;\n\t"                                       \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                           \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, " FIRST_PRED ";\n\t"
#define CSE_MMA_NEXT_TAP "add.u32 ab, ab, %7;\n\tadd.u32 bb, bb, %8;\n\t"
template <int NK>
__device__ __forceinline__ void tc_mma_taps4(uint32_t pred, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi32,
                                             uint32_t idesc, uint32_t accumulate_first, uint32_t a_step, uint32_t b_step) {
  static_assert(NK == 1 || NK == 2 || NK == 4, "K steps per stage chunk");
#define CSE_MMA_HEAD4                                                             \
  "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl, ab, bb;\n\t" \
  "setp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t"     \
  "mov.b32 ab, %1;\n\tmov.b32 bb, %2;\n\t"
#define CSE_MMA_ARGS ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi32), "r"(idesc), "r"(accumulate_first), "r"(pred), "r"(a_step), "r"(b_step) : "memory"
  if (NK == 1)
    asm volatile(CSE_MMA_HEAD4 CSE_MMA_TAP("p") CSE_MMA_NEXT_TAP CSE_MMA_TAP("t") CSE_MMA_NEXT_TAP CSE_MMA_TAP("t")
                 CSE_MMA_NEXT_TAP CSE_MMA_TAP("t") "}" CSE_MMA_ARGS);
  else if (NK == 2)
    asm volatile(CSE_MMA_HEAD4 CSE_MMA_TAP("p") CSE_MMA_STEP CSE_MMA_NEXT_TAP CSE_MMA_TAP("t") CSE_MMA_STEP CSE_MMA_NEXT_TAP
                 CSE_MMA_TAP("t") CSE_MMA_STEP CSE_MMA_NEXT_TAP CSE_MMA_TAP("t") CSE_MMA_STEP "}" CSE_MMA_ARGS);
  else
    asm volatile(CSE_MMA_HEAD4 CSE_MMA_TAP("p") CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_NEXT_TAP
                 CSE_MMA_TAP("t") CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_NEXT_TAP
                 CSE_MMA_TAP("t") CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_NEXT_TAP
                 CSE_MMA_TAP("t") CSE_MMA_STEP CSE_MMA_STEP CSE_MMA_STEP "}" CSE_MMA_ARGS);
#undef CSE_MMA_ARGS
#undef CSE_MMA_HEAD4
}
#undef CSE_MMA_NEXT_TAP
#undef CSE_MMA_TAP
#undef CSE_MMA_STEP
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (sm_100 format, version 1).
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte
//   offset >> 4 (distance between 8-row groups), [46,48) version = 1, [61,64) layout type.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// Fast epilogue math for one EC-wide chunk of a row (no residual, no second output):
// y = acc*scale + shift -> (ReLU) -> bf16 -> swizzled staging row.
template <int EC, bool SCALE, bool RELU>
__device__ __forceinline__ void epi_chunk_fast(const uint32_t (&r)[EC], const float (*par)[256], int c0, uint32_t s0,
                                               uint32_t swz) {
#pragma unroll
  for (int g8 = 0; g8 < EC / 8; ++g8) {
    const float4 sh_a = *reinterpret_cast<const float4*>(&par[1][c0 + g8 * 8]);
    const float4 sh_b = *reinterpret_cast<const float4*>(&par[1][c0 + g8 * 8 + 4]);
    const float shv[8] = {sh_a.x, sh_a.y, sh_a.z, sh_a.w, sh_b.x, sh_b.y, sh_b.z, sh_b.w};
    float y[8];
    if (SCALE) {
      const float4 sc_a = *reinterpret_cast<const float4*>(&par[0][c0 + g8 * 8]);
      const float4 sc_b = *reinterpret_cast<const float4*>(&par[0][c0 + g8 * 8 + 4]);
      const float scv[8] = {sc_a.x, sc_a.y, sc_a.z, sc_a.w, sc_b.x, sc_b.y, sc_b.z, sc_b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(__uint_as_float(r[g8 * 8 + j]), scv[j], shv[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[g8 * 8 + j]) + shv[j];
    }
    uint32_t p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = RELU ? pack_bf16x2_relu(y[2 * j], y[2 * j + 1]) : pack_bf16x2(y[2 * j], y[2 * j + 1]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + (((uint32_t)g8 ^ swz) << 4)), "r"(p[0]), "r"(p[1]),
                 "r"(p[2]), "r"(p[3]) : "memory");
  }
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

}  // namespace cse
