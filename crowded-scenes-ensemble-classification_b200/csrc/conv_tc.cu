// tcgen05 implicit-GEMM Conv3D for sm_100a.
//
//   GEMM view:  D[M, N] += A[M, K] * B[N, K]^T
//     M = output pixels, tiled as 128-row "bricks" (bn x bd x bh x bw output pixels),
//     N = output channels (tile BN <= 256), K = taps x input channels (chunk KC per stage).
//   A is never materialised: for every filter tap (fd,fh,fw) and channel chunk the producer
//   issues ONE 5-D TMA load (dims C,W,H,D,N of the NDHWC activation) whose box is the brick
//   shifted by the tap offset; out-of-bounds coordinates are zero-filled by the TMA unit,
//   which is exactly Conv3D's zero padding (symmetric or TF-'same' asymmetric alike).  The box
//   lands in shared memory as a K-major, 128B/64B/32B-swizzled [rows][KC] tile that
//   tcgen05.mma consumes directly through a shared-memory descriptor.
//   B (weights) is pre-packed [Cout_pad][taps*kchunks*KC] bf16, K-major, loaded by 2-D TMA.
//   Accumulators live in TMEM (2 x BN fp32 columns, double buffered) so the epilogue of tile i
//   (tcgen05.ld -> scale/shift -> +residual -> ReLU -> bf16 -> global) overlaps the main loop
//   of tile i+1.  Persistent grid: one CTA per SM, static round-robin over tiles.
//
//   Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
//   warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
//
// Replaces the cuDNN FP32 Conv3D the reference reaches through Keras/TF
// (train.py:653-658, 1230-1258, 1294-1298 ...), with the fused bias / BatchNormalization /
// ReLU / residual-add / second BN-ReLU output / channel-offset (concat) write epilogue.
#include "common.cuh"

namespace cse {

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;
constexpr int TC_MAX_STAGES = 8;
constexpr uint32_t TC_TMEM_COLS = 512;

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
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
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

struct ConvTcArgs {
  // geometry
  int Do, Ho, Wo, Co;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int kchunks;
  int bn, n_tiles_n;
  int b_n, b_d, b_h, b_w;          // brick
  int tiles_d, tiles_h, tiles_w;
  int n_batch;                     // clips in this launch
  long long num_tiles;             // m_tiles * n_tiles_n
  int stages;
  uint32_t a_bytes, b_bytes;       // TMA bytes per stage
  int out_ld;
  Epilogue ep;
};

// KC = channels per stage (16 -> SWIZZLE_32B, 32 -> SWIZZLE_64B, 64 -> SWIZZLE_128B)
template <int KC>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const ConvTcArgs a) {
  constexpr uint32_t ROW_BYTES = KC * 2;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : (KC == 32 ? 4u : 6u);
  constexpr uint32_t A_STAGE = TC_BM * ROW_BYTES;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  // 1024-byte aligned base for the swizzled tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_stage_bytes = (uint32_t)a.bn * ROW_BYTES;
  const uint32_t stage_bytes = A_STAGE + ((b_stage_bytes + 1023u) & ~1023u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full_bar[b]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[b]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int taps = a.kd * a.kh * a.kw;
  const int ksteps = taps * a.kchunks;
  const int tiles_per_n = a.tiles_d * a.tiles_h * a.tiles_w;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % a.n_tiles_n);
        long long mt = tile / a.n_tiles_n;
        const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
        const int th = (int)(mt % a.tiles_h); mt /= a.tiles_h;
        const int td = (int)(mt % a.tiles_d);
        const int tn = (int)(mt / a.tiles_d);
        const int iw0 = tw * a.b_w * a.sw - a.pw;
        const int ih0 = th * a.b_h * a.sh - a.ph;
        const int id0 = td * a.b_d * a.sd - a.pd;
        const int n0 = tn * a.b_n;
        int kstep = 0;
        for (int fd = 0; fd < a.kd; ++fd)
          for (int fh = 0; fh < a.kh; ++fh)
            for (int fw = 0; fw < a.kw; ++fw)
              for (int ch = 0; ch < a.kchunks; ++ch, ++kstep) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                mbar_expect_tx(fb, a.a_bytes + a.b_bytes);
                const uint32_t sa = smem_base + stage * stage_bytes;
                tma_load_5d(sa, &tmap_a, fb, ch * KC, iw0 + fw, ih0 + fh, id0 + fd, n0);
                tma_load_2d(sa + A_STAGE, &tmap_b, fb, kstep * KC, nt * a.bn);
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
              }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 @17, M>>4 @24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) |
                           ((uint32_t)(TC_BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase[2] = {0u, 0u};
    int buf = 0;
    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&tmem_empty_bar[buf]), acc_phase[buf] ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + A_STAGE;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * 32, SBO, LAYOUT);
            const uint64_t bd = make_smem_desc(sb + k * 32, SBO, LAYOUT);
            tc_mma_bf16(d_tmem, ad, bd, idesc, (ks > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[stage]));       // frees the smem stage when the MMAs retire
          if (ks == ksteps - 1) tc_commit(smem_u32(&tmem_full_bar[buf]));
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
      acc_phase[buf] ^= 1u;
      buf ^= 1;
    }
  } else {
    // =============================== epilogue ===================================
    const int quad = warp % 4;                      // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;               // tile row = output pixel inside the brick
    const int rw = row % a.b_w;
    const int rh = (row / a.b_w) % a.b_h;
    const int rd = (row / (a.b_w * a.b_h)) % a.b_d;
    const int rn = row / (a.b_w * a.b_h * a.b_d);
    uint32_t acc_phase[2] = {0u, 0u};
    int buf = 0;
    __nv_bfloat16* out0 = reinterpret_cast<__nv_bfloat16*>(a.ep.out0);
    __nv_bfloat16* out1 = reinterpret_cast<__nv_bfloat16*>(a.ep.out1);
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(a.ep.res);
    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int nt = (int)(tile % a.n_tiles_n);
      long long mt = tile / a.n_tiles_n;
      const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
      const int th = (int)(mt % a.tiles_h); mt /= a.tiles_h;
      const int td = (int)(mt % a.tiles_d);
      const int tn = (int)(mt / a.tiles_d);
      const int ow = tw * a.b_w + rw, oh = th * a.b_h + rh, od = td * a.b_d + rd, on = tn * a.b_n + rn;
      const bool valid = (rn < a.b_n) && ow < a.Wo && oh < a.Ho && od < a.Do && on < a.n_batch;
      const long long pix = (((long long)on * a.Do + od) * a.Ho + oh) * a.Wo + ow;
      const int col_base = nt * a.bn;

      mbar_wait(smem_u32(&tmem_full_bar[buf]), acc_phase[buf]);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * 256u;
      for (int c0 = 0; c0 < a.bn; c0 += 16) {
        uint32_t r[16];
        tc_ld16(t_row + (uint32_t)c0, r);
        tc_wait_ld();
        const int col = col_base + c0;
        if (valid && col < a.Co) {
          float y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = __uint_as_float(r[j]);
            const int cj = min(col + j, a.Co - 1);       // clamp: tile may overhang Cout
            if (a.ep.scale0) v *= __ldg(a.ep.scale0 + cj);
            if (a.ep.shift0) v += __ldg(a.ep.shift0 + cj);
            y[j] = v;
          }
          if (res) {
            const uint4* rp = reinterpret_cast<const uint4*>(res + pix * a.ep.res_ld + col);
            uint4 q0 = rp[0], q1 = make_uint4(0u, 0u, 0u, 0u);
            if (col + 8 < a.Co) q1 = rp[1];
            const __nv_bfloat16* e0 = reinterpret_cast<const __nv_bfloat16*>(&q0);
            const __nv_bfloat16* e1 = reinterpret_cast<const __nv_bfloat16*>(&q1);
#pragma unroll
            for (int j = 0; j < 8; ++j) { y[j] += __bfloat162float(e0[j]); y[8 + j] += __bfloat162float(e1[j]); }
          }
          {
            __align__(16) __nv_bfloat16 o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float2bfloat16_rn(a.ep.relu0 ? fmaxf(y[j], 0.f) : y[j]);
            uint4* op = reinterpret_cast<uint4*>(out0 + pix * a.out_ld + col);
            op[0] = reinterpret_cast<const uint4*>(o)[0];
            if (col + 8 < a.Co) op[1] = reinterpret_cast<const uint4*>(o)[1];
          }
          if (out1) {
            __align__(16) __nv_bfloat16 o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float z = y[j];
              const int cj = min(col + j, a.Co - 1);
              if (a.ep.scale1) z *= __ldg(a.ep.scale1 + cj);
              if (a.ep.shift1) z += __ldg(a.ep.shift1 + cj);
              o[j] = __float2bfloat16_rn(a.ep.relu1 ? fmaxf(z, 0.f) : z);
            }
            uint4* op = reinterpret_cast<uint4*>(out1 + pix * a.ep.out1_ld + col);
            op[0] = reinterpret_cast<const uint4*>(o)[0];
            if (col + 8 < a.Co) op[1] = reinterpret_cast<const uint4*>(o)[1];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
      acc_phase[buf] ^= 1u;
      buf ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int conv_tc_build(ConvTcDesc* d, const void* in, const void* w_packed, int max_batch, const WinGeom& g,
                  int kc, int bn, const int brick[4]) {
  CSE_REQUIRE(kc == 16 || kc == 32 || kc == 64, "conv_tc: kc=%d must be 16/32/64", kc);
  CSE_REQUIRE(bn >= 16 && bn <= 256 && bn % 16 == 0, "conv_tc: bn=%d must be a multiple of 16 in [16,256]", bn);
  CSE_REQUIRE(g.Ci % 8 == 0 && g.in_ld % 8 == 0, "conv_tc: Cin=%d / ld=%d must be multiples of 8", g.Ci, g.in_ld);
  CSE_REQUIRE(g.in_wpitch == 0 || g.in_wpitch * g.in_ld >= (g.Wi - 1) * g.in_ld + g.Ci,
              "conv_tc: row pitch %d too small for W=%d, window %d", g.in_wpitch, g.Wi, g.Ci);
  CSE_REQUIRE(g.Co % 8 == 0 && g.out_ld % 8 == 0, "conv_tc: Cout=%d / ld=%d must be multiples of 8", g.Co, g.out_ld);
  CSE_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)w_packed % 16) == 0, "conv_tc: pointers must be 16B aligned");
  CSE_REQUIRE(g.sd >= 1 && g.sd <= 8 && g.sh >= 1 && g.sh <= 8 && g.sw >= 1 && g.sw <= 8, "conv_tc: stride out of range");
  const int rows = brick[0] * brick[1] * brick[2] * brick[3];
  CSE_REQUIRE(rows >= 1 && rows <= TC_BM, "conv_tc: brick %dx%dx%dx%d exceeds 128 rows", brick[0], brick[1], brick[2], brick[3]);
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return CSE_ERR_CUDA;

  d->g = g; d->kc = kc; d->bn = bn; d->max_batch = max_batch;
  d->n_tiles_n = ceil_div(g.Co, bn);
  d->kchunks = ceil_div(g.Ci, kc);
  for (int i = 0; i < 4; ++i) d->brick[i] = brick[i];
  d->tiles_d = ceil_div(g.Do, brick[1]);
  d->tiles_h = ceil_div(g.Ho, brick[2]);
  d->tiles_w = ceil_div(g.Wo, brick[3]);

  // A: 5-D map over the NDHWC activation (dims C,W,H,D,N).  With stride s the box spans b*s
  // input positions and elementStrides = s picks every s-th one (b elements land in smem).
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di, (cuuint64_t)max_batch};
    // in_wpitch > Wi: rows of a W-padded tensor; with Ci > in_ld the channel window of a pixel
    // overlaps its right neighbours (packed stem: 4 pixels x 8 channels = one 32-wide K chunk).
    const cuuint64_t wp = (cuuint64_t)(g.in_wpitch > 0 ? g.in_wpitch : g.Wi);
    cuuint64_t strides[4] = {(cuuint64_t)g.in_ld * 2, wp * g.in_ld * 2, (cuuint64_t)g.Hi * wp * g.in_ld * 2,
                             (cuuint64_t)g.Di * g.Hi * wp * g.in_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)(brick[3] * g.sw), (cuuint32_t)(brick[2] * g.sh),
                         (cuuint32_t)(brick[1] * g.sd), (cuuint32_t)brick[0]};
    cuuint32_t estr[5] = {1, (cuuint32_t)g.sw, (cuuint32_t)g.sh, (cuuint32_t)g.sd, 1};
    CUresult r = enc(&d->tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A) failed: %d (dims %d,%d,%d,%d,%d ld %d box %u,%u,%u,%u,%u)", (int)r, g.Ci, g.Wi,
                g.Hi, g.Di, max_batch, g.in_ld, box[0], box[1], box[2], box[3], box[4]);
      return CSE_ERR_CUDA;
    }
  }
  // B: [Cout_pad][Ktot] bf16, K-major
  {
    const long long ktot = (long long)g.kd * g.kh * g.kw * d->kchunks * kc;
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(d->n_tiles_n * bn)};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&d->tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B) failed: %d (ktot %lld rows %d box %d,%d)", (int)r, ktot, d->n_tiles_n * bn, kc, bn);
      return CSE_ERR_CUDA;
    }
  }
  const size_t a_stage = (size_t)TC_BM * kc * 2;
  const size_t b_stage = (((size_t)bn * kc * 2) + 1023) & ~(size_t)1023;
  const size_t stage = a_stage + b_stage;
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / stage);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  CSE_REQUIRE(stages >= 2, "conv_tc: tile too large for shared memory");
  d->stages = stages;
  d->smem_bytes = stage * stages + 1024;   // + alignment slack
  return CSE_OK;
}

template <int KC>
static int launch_tc_t(const ConvTcDesc& d, const ConvTcArgs& args, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CSE_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  conv_tc_kernel<KC><<<grid, TC_THREADS, d.smem_bytes, st>>>(d.tmap_a, d.tmap_b, args);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

int launch_conv_tc(const ConvTcDesc& d, int n, const Epilogue& ep, int sm_count, cudaStream_t st) {
  CSE_REQUIRE(n >= 0 && n <= d.max_batch, "conv_tc: n=%d exceeds max_batch=%d", n, d.max_batch);
  if (n == 0) return CSE_OK;
  const WinGeom& g = d.g;
  ConvTcArgs a;
  a.Do = g.Do; a.Ho = g.Ho; a.Wo = g.Wo; a.Co = g.Co;
  a.kd = g.kd; a.kh = g.kh; a.kw = g.kw; a.sd = g.sd; a.sh = g.sh; a.sw = g.sw;
  a.pd = g.pd; a.ph = g.ph; a.pw = g.pw;
  a.kchunks = d.kchunks; a.bn = d.bn; a.n_tiles_n = d.n_tiles_n;
  a.b_n = d.brick[0]; a.b_d = d.brick[1]; a.b_h = d.brick[2]; a.b_w = d.brick[3];
  a.tiles_d = d.tiles_d; a.tiles_h = d.tiles_h; a.tiles_w = d.tiles_w;
  a.n_batch = n;
  const long long m_tiles = (long long)ceil_div(n, d.brick[0]) * d.tiles_d * d.tiles_h * d.tiles_w;
  a.num_tiles = m_tiles * d.n_tiles_n;
  a.stages = d.stages;
  a.a_bytes = (uint32_t)(d.brick[0] * d.brick[1] * d.brick[2] * d.brick[3] * d.kc * 2);
  a.b_bytes = (uint32_t)(d.bn * d.kc * 2);
  a.out_ld = g.out_ld;
  a.ep = ep;
  int grid = (int)(a.num_tiles < (long long)sm_count ? a.num_tiles : (long long)sm_count);
  switch (d.kc) {
    case 64: return launch_tc_t<64>(d, a, grid, st);
    case 32: return launch_tc_t<32>(d, a, grid, st);
    default: return launch_tc_t<16>(d, a, grid, st);
  }
}

}  // namespace cse
