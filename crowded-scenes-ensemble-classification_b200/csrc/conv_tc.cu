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
//   Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
//   warps 2..9 = epilogue (TMEM lane quadrant = warp_idx % 4, two warps per quadrant).
//
// Replaces the cuDNN FP32 Conv3D the reference reaches through Keras/TF
// (train.py:653-658, 1230-1258, 1294-1298 ...), with the fused bias / BatchNormalization /
// ReLU / residual-add / second BN-ReLU output / channel-offset (concat) write epilogue.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

constexpr int TC_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_BM = 128;
constexpr int TC_MAX_STAGES = 24;         // small stages (kc = 16 / 32) need a deep ring: stage time = TMA latency / stages
// mbarrier byte offsets inside `bars`: full[stage] at 0, then
constexpr uint32_t BAR_EMPTY = 8u * TC_MAX_STAGES;            // empty[stage]
constexpr uint32_t BAR_TMEM_FULL = 16u * TC_MAX_STAGES;       // tmem_full[8]
constexpr uint32_t BAR_TMEM_EMPTY = BAR_TMEM_FULL + 64u;      // tmem_empty[8]
constexpr uint32_t BAR_B_RESIDENT = BAR_TMEM_EMPTY + 64u;     // resident-B full
constexpr uint32_t BAR_B_FULL = BAR_B_RESIDENT + 8u;          // shared-B ring full[3]
constexpr uint32_t BAR_B_EMPTY = BAR_B_FULL + 24u;            // shared-B ring empty[3]
constexpr int BAR_COUNT = (int)(BAR_B_EMPTY + 24u) / 8;
constexpr uint32_t TC_TMEM_COLS = 512;

struct ConvTcArgs {
  // geometry
  int Do, Ho, Wo, Co;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int kchunks;
  int bn, n_tiles_n;
  int b_n, b_d, b_h, b_w;          // brick
  int tiles_d, tiles_h, tiles_w;
  int n_batch;                     // clips in this launch
  int num_tiles;                   // m_tiles * n_tiles_n
  int stages;
  uint32_t a_bytes, b_bytes;       // TMA bytes per stage
  uint32_t a_stage, stage_bytes;   // smem bytes of the A region / of a whole stage (1024-aligned)
  int halo;                        // 1: one stage per tile holds the (kd,kh)-halo'd A brick + all taps of B
  int b_resident;                  // halo mode, single N tile: B is loaded once per CTA, stages hold A only
  uint32_t b_region;               // smem offset of the resident B tile
  uint32_t stage_region;           // bytes of the A/B pipeline region (1024-aligned)
  int nslots;                      // staging slots for the TMA-store epilogue
  uint32_t slot_bytes;             // bytes of one staging slot (full tile [+ out1 tile] [+ pooled tile])
  int pool_d, pool_h, pool_w;      // fused MaxPooling3D window (= stride); 0 = no pooling
  int pool_zero;                   // rows outside the conv output count as 0 (ZeroPadding3D before the pool)
  int out_split, out_split2;       // fused sibling 1x1 convs: columns >= out_split go to tmap_o1, columns >= out_split2 to tmap_o2
  int bshare;                      // h-halo mode: groups of `bshare` M tiles share every weight stage (B ring of b_slots slots)
  int b_slots;
  uint32_t b_stage;                // bshare: bytes of one slot of the B ring (1024-aligned); the ring starts at b_region
  int nbuf;                        // TMEM accumulators in flight (2 / 4 twin / 2*bshare), acc_cols columns apart
  uint32_t acc_cols;
  int stepG[5];                    // bshare * gridDim.x as mixed-radix digits
  int twin;                        // twin-tile mode: two M tiles share every B stage (4 TMEM accumulators of bn <= 128 columns)
  int pair_pool;                   // pair-packed stem: GEMM row = 2 output pixels (N = 2*Cout), (1,2,2) max-pool in registers
  int step1[5], step2[5];          // gridDim.x and 2*gridDim.x as mixed-radix digits (nt, tw, th, td, tn)
  int stepE[5];                    // (epilogue groups)*gridDim.x: the stride of one epilogue group's tile walk
  // split-K (generic mode, layers with fewer tiles than SMs): the N-tile digit of the tile walk carries the K split too
  // (n_tiles_n = real N tiles x ksplit, nt = digit / ksplit, split = digit % ksplit); every split accumulates
  // ksteps_per k-steps and stores its fp32 accumulator tile to partial[split][m_row][col]; splitk_reduce_kernel sums the
  // splits in order and applies the epilogue
  int ksplit, ksteps_per;
  float* partial;
  int part_ld;
  long long part_rows;
  Epilogue ep;
};

// Tile coordinates of a CTA's round-robin walk, advanced by digit-wise addition with carry
// instead of five integer divisions per tile.
struct TileIter {
  int tile, nt, tw, th, td, tn;
  __device__ __forceinline__ void init(const ConvTcArgs& a, int t0) {
    tile = t0;
    nt = t0 % a.n_tiles_n;
    int mt = t0 / a.n_tiles_n;
    tw = mt % a.tiles_w; mt /= a.tiles_w;
    th = mt % a.tiles_h; mt /= a.tiles_h;
    td = mt % a.tiles_d;
    tn = mt / a.tiles_d;
  }
  __device__ __forceinline__ void advance(const ConvTcArgs& a, const int (&st)[5], int stride) {
    tile += stride;
    nt += st[0]; int c = nt >= a.n_tiles_n; nt -= c ? a.n_tiles_n : 0;
    tw += st[1] + c; c = tw >= a.tiles_w; tw -= c ? a.tiles_w : 0;
    th += st[2] + c; c = th >= a.tiles_h; th -= c ? a.tiles_h : 0;
    td += st[3] + c; c = td >= a.tiles_d; td -= c ? a.tiles_d : 0;
    tn += st[4] + c;
  }
};

// KC = channels per pipeline stage (16 -> SWIZZLE_32B, 32 -> SWIZZLE_64B, 64 -> SWIZZLE_128B)
// EC = output channels per epilogue chunk = TMA-store box width (16/32/64, same swizzle family)
// NG = epilogue groups of 4 warps (one TMEM accumulator each): 2 everywhere; a 4-group instantiation exists for the
// pair-packed C3D stem (four 128-column accumulators drained in turn; opt-in, see conv_tc_build).
template <int KC, int EC, int NG = 2>
__global__ void __launch_bounds__(64 + 128 * NG, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_o0, const __grid_constant__ CUtensorMap tmap_o1,
               const __grid_constant__ CUtensorMap tmap_o2,
               const ConvTcArgs a) {
  constexpr uint32_t ROW_BYTES = KC * 2;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : (KC == 32 ? 4u : 6u);
  constexpr uint32_t STG_BYTES = TC_BM * EC * 2;       // one staged [128][EC] bf16 tile

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // mbarriers: see the BAR_* offsets
  __shared__ __align__(8) uint64_t bars[BAR_COUNT];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_par[NG][4][256];                  // per epilogue group: scale0, shift0, scale1, shift1

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  // 1024-byte aligned base for the swizzled tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t A_STAGE = a.a_stage;
  const uint32_t stage_bytes = a.stage_bytes;
  const uint32_t stg_base = smem_base + a.stage_region;

  const uint32_t bar_base = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(bar_base + 8u * s, 1);
      mbar_init(bar_base + BAR_EMPTY + 8u * s, 1);
    }
    for (int b = 0; b < 8; ++b) {
      mbar_init(bar_base + BAR_TMEM_FULL + 8u * b, 1);
      mbar_init(bar_base + BAR_TMEM_EMPTY + 8u * b, 128);
    }
    mbar_init(bar_base + BAR_B_RESIDENT, 1);
    for (int b = 0; b < 3; ++b) {
      mbar_init(bar_base + BAR_B_FULL + 8u * b, 1);
      mbar_init(bar_base + BAR_B_EMPTY + 8u * b, 1);
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
  // The CTA owns the SM (1 CTA/SM) and allocates all 512 columns, so the allocation starts at
  // TMEM address 0; treating it as a constant keeps every TMEM operand warp-uniform.
  if (tmem_base_smem != 0u) {
    if (threadIdx.x == 0) printf("cse conv_tc: unexpected TMEM base %u\n", tmem_base_smem);
    __trap();
  }
  const uint32_t tmem_base = 0u;

  const int taps = a.kd * a.kh * a.kw;
  const int ksteps = taps * a.kchunks;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    const uint32_t leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    if (a.b_resident) {
      // the whole (single N tile) weight matrix stays in shared memory for the CTA's lifetime
      const uint32_t bb = bar_base + BAR_B_RESIDENT;
      mbar_expect_tx_p(leader, bb, a.b_bytes);
      const uint32_t b_tap = (uint32_t)a.bn * ROW_BYTES;        // resident B: one box per (fd,fh) tap
      for (int t = 0; t < a.kd * a.kh; ++t)
        tma_load_2d(leader, smem_base + a.b_region + t * b_tap, &tmap_b, bb, 0, t * a.bn);
    }
    if (a.twin) {
      // twin-tile mode (single N tile): the CTA's tiles 2j and 2j+1 run in lock-step; every stage
      // holds both A tiles and ONE B tile, which halves the weight traffic per MMA
      TileIter t0, t1;
      t0.init(a, blockIdx.x);
      t1.init(a, blockIdx.x + gridDim.x);
      for (; t0.tile < a.num_tiles; t0.advance(a, a.step2, 2 * gridDim.x), t1.advance(a, a.step2, 2 * gridDim.x)) {
        const bool two = t1.tile < a.num_tiles;
        const int iw0 = t0.tw * a.b_w * a.sw - a.pw, ih0 = t0.th * a.b_h * a.sh - a.ph, id0 = t0.td * a.b_d * a.sd - a.pd;
        const int iw1 = t1.tw * a.b_w * a.sw - a.pw, ih1 = t1.th * a.b_h * a.sh - a.ph, id1 = t1.td * a.b_d * a.sd - a.pd;
        const int n0 = t0.tn * a.b_n, n1 = t1.tn * a.b_n;
        int kcoord = 0;
        for (int fd = 0; fd < a.kd; ++fd)
          for (int fh = 0; fh < a.kh; ++fh)
            for (int fw = 0; fw < a.kw; ++fw)
              for (int ch = 0; ch < a.kchunks; ++ch, kcoord += KC) {
                mbar_wait(bar_base + BAR_EMPTY + 8u * stage, phase ^ 1u);
                const uint32_t fb = bar_base + 8u * stage;
                mbar_expect_tx_p(leader, fb, (two ? 2u : 1u) * a.a_bytes + a.b_bytes);
                const uint32_t sa = smem_base + stage * stage_bytes;
                tma_load_5d(leader, sa, &tmap_a, fb, ch * KC, iw0 + fw, ih0 + fh, id0 + fd, n0);
                if (two) tma_load_5d(leader, sa + A_STAGE, &tmap_a, fb, ch * KC, iw1 + fw, ih1 + fh, id1 + fd, n1);
                tma_load_2d(leader, sa + 2u * A_STAGE, &tmap_b, fb, kcoord, 0);
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
              }
      }
    }
    if (a.bshare) {
      // shared-B h-halo mode (7x7x7/2 stems, single N tile): the CTA's tiles run in groups of G; per
      // (fd, chunk) ONE weight block is loaded into the B ring and used by the G tiles' A boxes, so
      // the weights are streamed from L2 once per G tiles instead of once per tile
      const int G = a.bshare;
      TileIter tg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tg[j].init(a, blockIdx.x + j * gridDim.x);
      int bs = 0;
      uint32_t bphase = 0;
      while (tg[0].tile < a.num_tiles) {
        for (int fd = 0; fd < a.kd; ++fd)
          for (int ch = 0; ch < a.kchunks; ++ch) {
            mbar_wait(bar_base + BAR_B_EMPTY + 8u * bs, bphase ^ 1u);
            const uint32_t bf = bar_base + BAR_B_FULL + 8u * bs;
            mbar_expect_tx_p(leader, bf, a.b_bytes);
            tma_load_2d(leader, smem_base + a.b_region + (uint32_t)bs * a.b_stage, &tmap_b, bf, 0,
                        (fd * a.kchunks + ch) * a.kh * a.bn);
            if (++bs == a.b_slots) { bs = 0; bphase ^= 1u; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < G && tg[j].tile < a.num_tiles) {
                const int iw0 = tg[j].tw * a.b_w * a.sw - a.pw, ih0 = tg[j].th * a.b_h * a.sh - a.ph;
                const int id0 = tg[j].td * a.b_d * a.sd - a.pd, n0 = tg[j].tn * a.b_n;
                mbar_wait(bar_base + BAR_EMPTY + 8u * stage, phase ^ 1u);
                const uint32_t fb = bar_base + 8u * stage;
                mbar_expect_tx_p(leader, fb, a.a_bytes);
                tma_load_5d(leader, smem_base + stage * stage_bytes, &tmap_a, fb, ch * KC, iw0, ih0, id0 + fd, n0);
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
              }
            }
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) tg[j].advance(a, a.stepG, G * gridDim.x);
      }
    }
    TileIter ti;
    for (ti.init(a, (a.twin || a.bshare) ? a.num_tiles : blockIdx.x); ti.tile < a.num_tiles; ti.advance(a, a.step1, gridDim.x)) {
      const int nt = ti.nt;
      const int iw0 = ti.tw * a.b_w * a.sw - a.pw;
      const int ih0 = ti.th * a.b_h * a.sh - a.ph;
      const int id0 = ti.td * a.b_d * a.sd - a.pd;
      const int n0 = ti.tn * a.b_n;
      const int bcol = nt * a.bn;
      if (a.halo == 2) {
        // h-halo mode: one stage per (fd, channel chunk); the A box carries kh-1 extra rows, every fh
        // tap is a swizzle-atom-aligned row offset into it; B holds the kh taps of that (fd, chunk)
        for (int fd = 0; fd < a.kd; ++fd)
          for (int ch = 0; ch < a.kchunks; ++ch) {
            mbar_wait(bar_base + BAR_EMPTY + 8u * stage, phase ^ 1u);
            const uint32_t fb = bar_base + 8u * stage;
            mbar_expect_tx_p(leader, fb, a.a_bytes + a.b_bytes);
            const uint32_t sa = smem_base + stage * stage_bytes;
            tma_load_5d(leader, sa, &tmap_a, fb, ch * KC, iw0, ih0, id0 + fd, n0);
            tma_load_2d(leader, sa + A_STAGE, &tmap_b, fb, 0, ((nt * a.kd + fd) * a.kchunks + ch) * a.kh * a.bn);
            if (++stage == a.stages) { stage = 0; phase ^= 1u; }
          }
        continue;
      }
      if (a.halo) {
        // halo mode: the A box carries kd-1 / kh-1 extra planes / rows; every (fd,fh) tap is a
        // swizzle-atom-aligned row offset into it, so one load feeds all taps of the tile
        mbar_wait(bar_base + BAR_EMPTY + 8u * stage, phase ^ 1u);
        const uint32_t fb = bar_base + 8u * stage;
        mbar_expect_tx_p(leader, fb, a.b_resident ? a.a_bytes : a.a_bytes + a.b_bytes);
        const uint32_t sa = smem_base + stage * stage_bytes;
        tma_load_5d(leader, sa, &tmap_a, fb, 0, iw0, ih0, id0, n0);
        if (!a.b_resident) {
          const uint32_t b_fd = (uint32_t)(a.kh * a.bn) * ROW_BYTES;
          for (int fd = 0; fd < a.kd; ++fd)
            tma_load_2d(leader, sa + A_STAGE + fd * b_fd, &tmap_b, fb, 0, (nt * a.kd + fd) * a.kh * a.bn);
        }
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        continue;
      }
      if (a.ksplit > 1) {
        // this work item = k-steps [k0, k1) of the tile's K loop (flattened tap-major, chunk-minor like the weights)
        const int bcol_s = (nt / a.ksplit) * a.bn;
        const int k0 = (nt % a.ksplit) * a.ksteps_per, k1 = min(k0 + a.ksteps_per, ksteps);
        for (int kidx = k0; kidx < k1; ++kidx) {
          const int ch = kidx % a.kchunks;
          int tap = kidx / a.kchunks;
          const int fw = tap % a.kw; tap /= a.kw;
          const int fh = tap % a.kh;
          const int fd = tap / a.kh;
          mbar_wait(bar_base + BAR_EMPTY + 8u * stage, phase ^ 1u);
          const uint32_t fb = bar_base + 8u * stage;
          mbar_expect_tx_p(leader, fb, a.a_bytes + a.b_bytes);
          const uint32_t sa = smem_base + stage * stage_bytes;
          tma_load_5d(leader, sa, &tmap_a, fb, ch * KC, iw0 + fw, ih0 + fh, id0 + fd, n0);
          tma_load_2d(leader, sa + A_STAGE, &tmap_b, fb, kidx * KC, bcol_s);
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        continue;
      }
      int kcoord = 0;
      for (int fd = 0; fd < a.kd; ++fd)
        for (int fh = 0; fh < a.kh; ++fh)
          for (int fw = 0; fw < a.kw; ++fw)
            for (int ch = 0; ch < a.kchunks; ++ch, kcoord += KC) {
              mbar_wait(bar_base + BAR_EMPTY + 8u * stage, phase ^ 1u);          // empty[stage]
              const uint32_t fb = bar_base + 8u * stage;                   // full[stage]
              mbar_expect_tx_p(leader, fb, a.a_bytes + a.b_bytes);
              const uint32_t sa = smem_base + stage * stage_bytes;
              tma_load_5d(leader, sa, &tmap_a, fb, ch * KC, iw0 + fw, ih0 + fh, id0 + fd, n0);
              tma_load_2d(leader, sa + A_STAGE, &tmap_b, fb, kcoord, bcol);
              if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 @17, M>>4 @24
    const uint32_t leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) |
                           ((uint32_t)(TC_BM >> 4) << 24);
    // descriptor template: everything but the 14-bit start address
    const uint32_t desc_hi32 = (uint32_t)(make_smem_desc(0, SBO, LAYOUT) >> 32);
    auto dlo = [](uint32_t saddr) -> uint32_t { return ((saddr >> 4) & 0x3FFFu) | 0x10000u; };   // start address + LBO = 1
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    if (a.b_resident) {
      mbar_wait(bar_base + BAR_B_RESIDENT, 0u);
      tc_fence_after();
    }
    if (a.twin) {
      uint32_t eph[2] = {0u, 0u};                 // tmem_empty phase per buffer pair
      int pp = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += 2 * gridDim.x) {
        const bool two = tile + (int)gridDim.x < a.num_tiles;
        const uint32_t b0 = (uint32_t)(pp * 2), b1 = b0 + 1u;
        const uint32_t ep_ = pp ? eph[1] : eph[0];
        mbar_wait(bar_base + BAR_TMEM_EMPTY + 8u * b0, ep_ ^ 1u);
        mbar_wait(bar_base + BAR_TMEM_EMPTY + 8u * b1, ep_ ^ 1u);
        tc_fence_after();
        const uint32_t d0 = b0 * 128u, d1 = b1 * 128u;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(bar_base + 8u * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t bl = dlo(sa + 2u * A_STAGE);
          tc_mma_k<KC / 16>(leader, d0, dlo(sa), bl, desc_hi32, idesc, ks > 0 ? 1u : 0u);
          if (two) tc_mma_k<KC / 16>(leader, d1, dlo(sa + A_STAGE), bl, desc_hi32, idesc, ks > 0 ? 1u : 0u);
          tc_commit(leader, bar_base + BAR_EMPTY + 8u * stage);
          if (ks == ksteps - 1) {
            tc_commit(leader, bar_base + BAR_TMEM_FULL + 8u * b0);
            if (two) tc_commit(leader, bar_base + BAR_TMEM_FULL + 8u * b1);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        if (pp) eph[1] ^= 1u; else eph[0] ^= 1u;
        pp ^= 1;
      }
    }
    if (a.bshare) {
      const int G = a.bshare;
      const int nst = a.kd * a.kchunks;
      const uint32_t a_fh = ((uint32_t)a.b_w * ROW_BYTES) >> 4;          // one brick row of pixels
      const uint32_t b_tap = ((uint32_t)a.bn * ROW_BYTES) >> 4;
      const uint32_t bmask = (uint32_t)a.nbuf - 1u;
      int bs = 0;
      uint32_t bphase = 0;
      uint32_t i0 = 0;                             // sequence index of the group's first tile in this CTA
      for (int tile0 = blockIdx.x; tile0 < a.num_tiles; tile0 += G * (int)gridDim.x, i0 += (uint32_t)G) {
        int nvalid = 0;
        for (int j = 0; j < G; ++j)
          if (tile0 + j * (int)gridDim.x < a.num_tiles) nvalid = j + 1;
        for (int j = 0; j < nvalid; ++j) {
          const uint32_t i = i0 + (uint32_t)j;
          mbar_wait(bar_base + BAR_TMEM_EMPTY + 8u * (i & bmask), ((i / (uint32_t)a.nbuf) & 1u) ^ 1u);
        }
        tc_fence_after();
        for (int st = 0; st < nst; ++st) {
          mbar_wait(bar_base + BAR_B_FULL + 8u * bs, bphase);
          tc_fence_after();
          const uint32_t sb = smem_base + a.b_region + (uint32_t)bs * a.b_stage;
          const uint32_t bl0 = dlo(sb);
          for (int j = 0; j < nvalid; ++j) {
            const uint32_t bj = (i0 + (uint32_t)j) & bmask;
            const uint32_t d_tmem = bj * a.acc_cols;
            mbar_wait(bar_base + 8u * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t al0 = dlo(sa);
            if (a.kh == 4) {
              tc_mma_taps4<KC / 16>(leader, d_tmem, al0, bl0, desc_hi32, idesc, st > 0 ? 1u : 0u, a_fh, b_tap);
            } else {
              for (int fh = 0; fh < a.kh; ++fh)
                tc_mma_k<KC / 16>(leader, d_tmem, al0 + (uint32_t)fh * a_fh, bl0 + (uint32_t)fh * b_tap, desc_hi32, idesc,
                                  (st > 0 || fh > 0) ? 1u : 0u);
            }
            tc_commit(leader, bar_base + BAR_EMPTY + 8u * stage);
            if (st == nst - 1) tc_commit(leader, bar_base + BAR_TMEM_FULL + 8u * bj);
            if (++stage == a.stages) { stage = 0; phase ^= 1u; }
          }
          tc_commit(leader, bar_base + BAR_B_EMPTY + 8u * bs);
          if (++bs == a.b_slots) { bs = 0; bphase ^= 1u; }
        }
      }
    }
    uint32_t it = 0;                   // the CTA's it-th tile lives in accumulator it % nbuf (2 x 256 or 4 x 128 columns)
    for (int tile = (a.twin || a.bshare) ? a.num_tiles : blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t nbm = (uint32_t)a.nbuf - 1u;
      buf = (int)(it & nbm);
      const uint32_t aph = (it / (uint32_t)a.nbuf) & 1u;
      mbar_wait(bar_base + BAR_TMEM_EMPTY + 8u * buf, aph ^ 1u);                   // tmem_empty[buf]
      tc_fence_after();
      const uint32_t d_tmem = (uint32_t)buf * a.acc_cols;                 // TMEM base is 0 (asserted)
      if (a.halo == 2) {
        const int nst = a.kd * a.kchunks;
        const uint32_t a_fh = ((uint32_t)a.b_w * ROW_BYTES) >> 4;          // one brick row of pixels
        const uint32_t b_tap = ((uint32_t)a.bn * ROW_BYTES) >> 4;
        uint32_t acc = 0u;
        for (int st = 0; st < nst; ++st) {
          mbar_wait(bar_base + 8u * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t al0 = dlo(sa), bl0 = dlo(sa + A_STAGE);
          if (a.kh == 4) {
            tc_mma_taps4<KC / 16>(leader, d_tmem, al0, bl0, desc_hi32, idesc, acc, a_fh, b_tap);
            acc = 1u;
          } else {
            for (int fh = 0; fh < a.kh; ++fh) {
              tc_mma_k<KC / 16>(leader, d_tmem, al0 + (uint32_t)fh * a_fh, bl0 + (uint32_t)fh * b_tap, desc_hi32, idesc, acc);
              acc = 1u;
            }
          }
          tc_commit(leader, bar_base + BAR_EMPTY + 8u * stage);
          if (st == nst - 1) tc_commit(leader, bar_base + BAR_TMEM_FULL + 8u * buf);
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        continue;
      }
      if (a.halo) {
        mbar_wait(bar_base + 8u * stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t sb = a.b_resident ? (smem_base + a.b_region) : (sa + A_STAGE);
        const uint32_t al0 = dlo(sa), bl0 = dlo(sb);
        const uint32_t a_fh = ((uint32_t)a.b_w * ROW_BYTES) >> 4;                       // one brick row of pixels
        const uint32_t a_fd = ((uint32_t)(a.b_h + a.kh - 1) * a.b_w * ROW_BYTES) >> 4;  // one halo plane
        const uint32_t b_tap = ((uint32_t)a.bn * ROW_BYTES) >> 4;
        uint32_t tap = 0;
        for (int fd = 0; fd < a.kd; ++fd)
          for (int fh = 0; fh < a.kh; ++fh, ++tap)
            tc_mma_k<KC / 16>(leader, d_tmem, al0 + (uint32_t)fd * a_fd + (uint32_t)fh * a_fh, bl0 + tap * b_tap, desc_hi32, idesc,
                              tap > 0 ? 1u : 0u);
        tc_commit(leader, bar_base + BAR_EMPTY + 8u * stage);
        tc_commit(leader, bar_base + BAR_TMEM_FULL + 8u * buf);
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        continue;
      }
      int nks = ksteps;
      if (a.ksplit > 1) {                       // split-K: this item's share of the K loop (N-tile digit % ksplit)
        const int k0 = ((tile % a.n_tiles_n) % a.ksplit) * a.ksteps_per;
        nks = min(a.ksteps_per, ksteps - k0);
      }
      for (int ks = 0; ks < nks; ++ks) {
        mbar_wait(bar_base + 8u * stage, phase);                           // full[stage]
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        tc_mma_k<KC / 16>(leader, d_tmem, dlo(sa), dlo(sa + A_STAGE), desc_hi32, idesc, ks > 0 ? 1u : 0u);
        tc_commit(leader, bar_base + BAR_EMPTY + 8u * stage);                    // frees the smem stage when the MMAs retire
        if (ks == nks - 1) tc_commit(leader, bar_base + BAR_TMEM_FULL + 8u * buf);   // tmem_full[buf]
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // =============================== epilogue ===================================
    // TMEM -> registers -> scale/shift (+residual) -> ReLU -> bf16 -> swizzled smem staging tile
    // -> one 5-D TMA store per EC-channel chunk (box = the output brick, clipped at the tensor
    // edges by the TMA unit, written at the op's channel offset / leading dimension).
    // Two independent epilogue groups of 4 warps (one warp per TMEM lane quadrant each): group g
    // owns accumulator buffer g and drains the CTA's tiles g, g+2, ... so the latency chain of
    // one tile's epilogue (barrier -> tcgen05.ld -> math -> staging -> TMA store) overlaps the
    // other group's, and both overlap the MMAs of the following tiles.
    const int grp = (warp - 2) / 4;
    const int quad = warp % 4;                      // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;               // tile row = output pixel inside the brick
    const int et = threadIdx.x - 64 - grp * 128;    // 0..127 within the group
    const bool store_thread = (et == 0);
    const int rw = row % a.b_w;
    const int rh = (row / a.b_w) % a.b_h;
    const int rd = (row / (a.b_w * a.b_h)) % a.b_d;
    const int rn = row / (a.b_w * a.b_h * a.b_d);
    // swizzle of the 16-byte chunk index inside a staged row (Swizzle<B,4,3> on byte addresses)
    const uint32_t swz = (EC == 64) ? (row & 7) : (EC == 32 ? ((row >> 1) & 3) : ((row >> 2) & 1));
    const bool has_out1 = a.ep.out1 != nullptr;
    const bool has_scale0 = a.ep.scale0 != nullptr;
    const bool relu0 = a.ep.relu0 != 0, relu1 = a.ep.relu1 != 0;
    const uint32_t slot_bytes = a.slot_bytes;
    const uint32_t my_stg = stg_base + (uint32_t)grp * (uint32_t)a.nslots * slot_bytes;
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(a.ep.res);
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    float (*par)[256] = s_par[grp];
    const bool fast = (res == nullptr) && !has_out1;            // no residual / second output: templated math
    // fused MaxPooling3D (window == stride, windows never straddle a brick)
    const bool pooled = a.pool_d > 0;
    const uint32_t pool_stg_off = STG_BYTES;        // pooled tile lives right after the full tile of a slot
    // The pooled tile is [prows][EC]; its 16-byte vectors are spread over the group's 128 threads.
    // Everything that does not depend on the tile is computed once here: for each of this
    // thread's (<= 4) vectors the first source row of its window and the destination offset.
    constexpr int V = EC / 8;                       // 16-byte vectors per row
    constexpr int PV_MAX = 4;
    int pv_n = 0;
    bool pool_precomputed = false;
    int pv_r2_0 = 0, pv_r2_1 = 0, pv_r2_2 = 0, pv_r2_3 = 0;
    uint32_t pv_cd_0 = 0, pv_cd_1 = 0, pv_cd_2 = 0, pv_cd_3 = 0;     // (dst offset << 4) | cv
    if (pooled) {
      const int pb_w = a.b_w / a.pool_w, pb_h = a.b_h / a.pool_h, pb_d = a.b_d / a.pool_d;
      const int nvec = a.b_n * pb_d * pb_h * pb_w * V;
      pool_precomputed = nvec <= PV_MAX * 128;
      auto pre = [&](int k, int& r2_out, uint32_t& cd_out) {
        const int v = et + k * 128;
        if (v < nvec && pool_precomputed) {
          const int p = v / V, cv = v % V;
          int q = p;
          const int pw_ = q % pb_w; q /= pb_w;
          const int ph_ = q % pb_h; q /= pb_h;
          const int pd_ = q % pb_d; const int pn_ = q / pb_d;
          r2_out = ((pn_ * a.b_d + pd_ * a.pool_d) * a.b_h + ph_ * a.pool_h) * a.b_w + pw_ * a.pool_w;
          const uint32_t swp = (EC == 64) ? (p & 7) : (EC == 32 ? ((p >> 1) & 3) : ((p >> 2) & 1));
          cd_out = (((uint32_t)p * (EC * 2) + (((uint32_t)cv ^ swp) << 4)) << 4) | (uint32_t)cv;
          pv_n = k + 1;
        }
      };
      pre(0, pv_r2_0, pv_cd_0); pre(1, pv_r2_1, pv_cd_1); pre(2, pv_r2_2, pv_cd_2); pre(3, pv_r2_3, pv_cd_3);
    }
    const uint32_t bar_id = 1u + (uint32_t)grp;
    // The CTA's i-th tile (i = grp, grp + 2, ...) lives in accumulator i % nbuf, acc_cols columns apart:
    // normal mode 2 x 256 columns (group g owns buffer g), twin mode 4 x 128, shared-B mode 2G x 512/(2G).
    uint32_t seq = (uint32_t)grp;
    const uint32_t bmask = (uint32_t)a.nbuf - 1u;
    int slot = 0;
    int last_nt = -1;
    TileIter ti;
    for (ti.init(a, blockIdx.x + grp * gridDim.x); ti.tile < a.num_tiles; ti.advance(a, a.stepE, NG * gridDim.x)) {
      const int buf = (int)(seq & bmask);
      const uint32_t acc_phase = (seq / (uint32_t)a.nbuf) & 1u;
      const uint32_t buf_col = (uint32_t)buf * a.acc_cols;
      seq += (uint32_t)NG;
      const int nt = ti.nt;
      const int ow0 = ti.tw * a.b_w, oh0 = ti.th * a.b_h, od0 = ti.td * a.b_d, on0 = ti.tn * a.b_n;
      const int col_base = nt * a.bn;
      bool use_res = false, row_valid = true;
      long long pix = 0;
      if (res != nullptr || pooled) {               // rows that exist in the conv output
        const int ow = ow0 + rw, oh = oh0 + rh, od = od0 + rd, on = on0 + rn;
        row_valid = (rn < a.b_n) && ow < a.Wo && oh < a.Ho && od < a.Do && on < a.n_batch;
        use_res = row_valid && res != nullptr;
        pix = (((long long)on * a.Do + od) * a.Ho + oh) * a.Wo + ow;
      }

      // epilogue parameters of this N tile -> smem (this group's previous readers are past their
      // last barrier); only when the N tile changed
      if (nt != last_nt) {
        for (int i = et; i < a.bn; i += 128) {
          const int c = a.pair_pool ? (i % (a.bn >> 1)) : min(col_base + i, a.Co - 1);
          par[0][i] = has_scale0 ? __ldg(a.ep.scale0 + c) : 1.f;
          par[1][i] = a.ep.shift0 ? __ldg(a.ep.shift0 + c) : 0.f;
          par[2][i] = a.ep.scale1 ? __ldg(a.ep.scale1 + c) : 1.f;
          par[3][i] = a.ep.shift1 ? __ldg(a.ep.shift1 + c) : 0.f;
        }
        last_nt = nt;
      }

      mbar_wait(bar_base + BAR_TMEM_FULL + 8u * buf, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + buf_col;
      if (a.ksplit > 1) {
        // split-K: raw fp32 accumulator rows -> partial[split][m_tile * 128 + row][N tile columns]
        const int split = nt % a.ksplit, ntile = nt / a.ksplit;
        const long long m_tile = (((long long)ti.tn * a.tiles_d + ti.td) * a.tiles_h + ti.th) * a.tiles_w + ti.tw;
        float* dst = a.partial + ((long long)split * a.part_rows + m_tile * TC_BM + row) * a.part_ld + ntile * a.bn;
        for (int c0 = 0; c0 < a.bn; c0 += 16) {
          uint32_t r[16];
          tc_ld16(t_row + (uint32_t)c0, r);
          tc_wait_ld();
          if (c0 + 16 >= a.bn) {
            tc_fence_before();
            mbar_arrive(bar_base + BAR_TMEM_EMPTY + 8u * buf);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dst + c0 + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        continue;
      }
      if (EC == 64 && a.pair_pool) {
        // Pair-packed stem + MaxPooling3D (1,2,2): row = (h, pixel pair), columns [0,64) = left pixel,
        // [64,128) = right pixel.  max over the pair in registers, + bias, (ReLU) -> bf16, max with the
        // partner row (h ^ 1 = lane ^ b_w, same warp) by shuffle; max commutes with the monotone
        // bias/ReLU/rounding, so the values equal pool(relu(conv + bias)) computed in bf16.
        if (store_thread) {
          switch (a.nslots) {
            case 1: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
            case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
            case 3: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
            default: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const int prow = (rh >> 1) * a.b_w + rw;                         // row of the pooled tile
        const uint32_t pdst = my_stg + (uint32_t)slot * slot_bytes + (uint32_t)prow * 128u;
        const uint32_t pswz = (uint32_t)(prow & 7);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t ra[16], rb[16];
          tc_ld16(t_row + (uint32_t)(q * 16), ra);
          tc_ld16(t_row + (uint32_t)(64 + q * 16), rb);
          tc_wait_ld();
          if (q == 3) {
            tc_fence_before();
            mbar_arrive(bar_base + BAR_TMEM_EMPTY + 8u * buf);
          }
          uint32_t pk[8];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 sh = *reinterpret_cast<const float4*>(&par[1][q * 16 + i4 * 4]);
            const float y0 = fmaxf(__uint_as_float(ra[i4 * 4 + 0]), __uint_as_float(rb[i4 * 4 + 0])) + sh.x;
            const float y1 = fmaxf(__uint_as_float(ra[i4 * 4 + 1]), __uint_as_float(rb[i4 * 4 + 1])) + sh.y;
            const float y2 = fmaxf(__uint_as_float(ra[i4 * 4 + 2]), __uint_as_float(rb[i4 * 4 + 2])) + sh.z;
            const float y3 = fmaxf(__uint_as_float(ra[i4 * 4 + 3]), __uint_as_float(rb[i4 * 4 + 3])) + sh.w;
            pk[i4 * 2 + 0] = relu0 ? pack_bf16x2_relu(y0, y1) : pack_bf16x2(y0, y1);
            pk[i4 * 2 + 1] = relu0 ? pack_bf16x2_relu(y2, y3) : pack_bf16x2(y2, y3);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, pk[i], a.b_w);
            const __nv_bfloat162 t = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&pk[i]),
                                             *reinterpret_cast<const __nv_bfloat162*>(&other));
            pk[i] = *reinterpret_cast<const uint32_t*>(&t);
          }
          if ((q >> 1) == (rh & 1)) {                                   // each lane of the row pair stages half the channels
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pdst + (((uint32_t)(2 * q) ^ pswz) << 4)), "r"(pk[0]),
                         "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pdst + (((uint32_t)(2 * q + 1) ^ pswz) << 4)), "r"(pk[4]),
                         "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (store_thread) {
          tma_store_5d(&tmap_o0, my_stg + (uint32_t)slot * slot_bytes, 0, ow0, oh0 >> 1, od0, on0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (++slot == a.nslots) slot = 0;
        continue;
      }
      if constexpr (NG == 2)          // (the 4-group instantiation only serves the pair-packed stem above)
      for (int c0 = 0; c0 < a.bn; c0 += EC) {
        // the staging slot we are about to overwrite must have been read by its TMA store
        if (store_thread) {
          switch (a.nslots) {
            case 1: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
            case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
            case 3: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
            default: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
          }
        }
        uint32_t r[EC];
#pragma unroll
        for (int q = 0; q < EC / 16; ++q) tc_ld16(t_row + (uint32_t)(c0 + q * 16), *reinterpret_cast<uint32_t(*)[16]>(&r[q * 16]));
        tc_wait_ld();
        if (c0 + EC >= a.bn) {
          // the accumulator now lives in registers: hand the TMEM buffer back to the MMA warp
          // before the math / staging / store of this last chunk
          tc_fence_before();
          mbar_arrive(bar_base + BAR_TMEM_EMPTY + 8u * buf);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const uint32_t s0 = my_stg + (uint32_t)slot * slot_bytes + (uint32_t)row * (EC * 2);
        if (fast) {
          if (has_scale0) {
            if (relu0) epi_chunk_fast<EC, true, true>(r, par, c0, s0, swz);
            else epi_chunk_fast<EC, true, false>(r, par, c0, s0, swz);
          } else {
            if (relu0) epi_chunk_fast<EC, false, true>(r, par, c0, s0, swz);
            else epi_chunk_fast<EC, false, false>(r, par, c0, s0, swz);
          }
          if (pooled && !row_valid) {               // rows outside the tensor: 0 (ZeroPadding3D) or -inf
            const uint32_t kv = a.pool_zero ? 0u : 0xFF80FF80u;
#pragma unroll
            for (int g8 = 0; g8 < EC / 8; ++g8)
              asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s0 + ((uint32_t)g8 << 4)), "r"(kv) : "memory");
          }
        } else {
#pragma unroll
          for (int g8 = 0; g8 < EC / 8; ++g8) {        // groups of 8 columns = one 16-byte staged chunk
            float y[8];
            const float4 sh_a = *reinterpret_cast<const float4*>(&par[1][c0 + g8 * 8]);
            const float4 sh_b = *reinterpret_cast<const float4*>(&par[1][c0 + g8 * 8 + 4]);
            const float shv[8] = {sh_a.x, sh_a.y, sh_a.z, sh_a.w, sh_b.x, sh_b.y, sh_b.z, sh_b.w};
            if (has_scale0) {
              const float4 sc_a = *reinterpret_cast<const float4*>(&par[0][c0 + g8 * 8]);
              const float4 sc_b = *reinterpret_cast<const float4*>(&par[0][c0 + g8 * 8 + 4]);
              const float scv[8] = {sc_a.x, sc_a.y, sc_a.z, sc_a.w, sc_b.x, sc_b.y, sc_b.z, sc_b.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = fmaf(__uint_as_float(r[g8 * 8 + j]), scv[j], shv[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[g8 * 8 + j]) + shv[j];
            }
            const int col = col_base + c0 + g8 * 8;
            if (use_res && col < a.Co) {
              const uint4 q0 = *reinterpret_cast<const uint4*>(res + pix * a.ep.res_ld + col);
              const __nv_bfloat16* e0 = reinterpret_cast<const __nv_bfloat16*>(&q0);
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] += __bfloat162float(e0[j]);
            }
            const uint32_t c16 = (uint32_t)g8;
            {
              uint32_t p[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
                if (relu0) h = __hmax2(h, zero2);
                p[j] = *reinterpret_cast<uint32_t*>(&h);
                if (pooled && !row_valid) p[j] = a.pool_zero ? 0u : 0xFF80FF80u;   // 0 or -inf for rows outside the tensor
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + ((c16 ^ swz) << 4)), "r"(p[0]), "r"(p[1]),
                           "r"(p[2]), "r"(p[3]) : "memory");
            }
            if (has_out1) {
              const float4 s1a = *reinterpret_cast<const float4*>(&par[2][c0 + g8 * 8]);
              const float4 s1b = *reinterpret_cast<const float4*>(&par[2][c0 + g8 * 8 + 4]);
              const float4 t1a = *reinterpret_cast<const float4*>(&par[3][c0 + g8 * 8]);
              const float4 t1b = *reinterpret_cast<const float4*>(&par[3][c0 + g8 * 8 + 4]);
              const float s1[8] = {s1a.x, s1a.y, s1a.z, s1a.w, s1b.x, s1b.y, s1b.z, s1b.w};
              const float t1[8] = {t1a.x, t1a.y, t1a.z, t1a.w, t1b.x, t1b.y, t1b.z, t1b.w};
              uint32_t p[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(fmaf(y[2 * j], s1[2 * j], t1[2 * j]),
                                                         fmaf(y[2 * j + 1], s1[2 * j + 1], t1[2 * j + 1]));
                if (relu1) h = __hmax2(h, zero2);
                p[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + STG_BYTES + ((c16 ^ swz) << 4)), "r"(p[0]),
                           "r"(p[1]), "r"(p[2]), "r"(p[3]) : "memory");
            }
          }
        }
        if (pooled) {
          // window max over the staged tile -> pooled tile [prows][EC] (same swizzle family), then store that
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          const uint32_t full = my_stg + (uint32_t)slot * slot_bytes;
          const uint32_t pst = full + pool_stg_off;
          if (pool_precomputed) {
            auto pool_vec = [&](int r2b, uint32_t cd) {
              const uint32_t cv = cd & 15u, dst = cd >> 4;
              uint32_t m0 = 0xFF80FF80u, m1 = 0xFF80FF80u, m2 = 0xFF80FF80u, m3 = 0xFF80FF80u;   // -inf, -inf
              for (int i = 0; i < a.pool_d; ++i)
                for (int j = 0; j < a.pool_h; ++j) {
                  const int rb = r2b + (i * a.b_h + j) * a.b_w;
                  for (int l = 0; l < a.pool_w; ++l) {
                    const int r2 = rb + l;
                    const uint32_t sw2 = (EC == 64) ? (r2 & 7) : (EC == 32 ? ((r2 >> 1) & 3) : ((r2 >> 2) & 1));
                    uint32_t x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                                 : "r"(full + (uint32_t)r2 * (EC * 2) + ((cv ^ sw2) << 4)));
                    __nv_bfloat162 t;
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m0), *reinterpret_cast<__nv_bfloat162*>(&x0)); m0 = *reinterpret_cast<uint32_t*>(&t);
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m1), *reinterpret_cast<__nv_bfloat162*>(&x1)); m1 = *reinterpret_cast<uint32_t*>(&t);
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m2), *reinterpret_cast<__nv_bfloat162*>(&x2)); m2 = *reinterpret_cast<uint32_t*>(&t);
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m3), *reinterpret_cast<__nv_bfloat162*>(&x3)); m3 = *reinterpret_cast<uint32_t*>(&t);
                  }
                }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pst + dst), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
                           : "memory");
            };
            if (pv_n > 0) pool_vec(pv_r2_0, pv_cd_0);
            if (pv_n > 1) pool_vec(pv_r2_1, pv_cd_1);
            if (pv_n > 2) pool_vec(pv_r2_2, pv_cd_2);
            if (pv_n > 3) pool_vec(pv_r2_3, pv_cd_3);
          } else {
            const int pb_w = a.b_w / a.pool_w, pb_h = a.b_h / a.pool_h, pb_d = a.b_d / a.pool_d;
            const int prows = a.b_n * pb_d * pb_h * pb_w;
            for (int v = et; v < prows * V; v += 128) {
              const int p = v / V, cv = v % V;
              int q = p;
              const int pw_ = q % pb_w; q /= pb_w;
              const int ph_ = q % pb_h; q /= pb_h;
              const int pd_ = q % pb_d; const int pn_ = q / pb_d;
              uint32_t m0 = 0xFF80FF80u, m1 = 0xFF80FF80u, m2 = 0xFF80FF80u, m3 = 0xFF80FF80u;   // -inf, -inf
              for (int i = 0; i < a.pool_d; ++i)
                for (int j = 0; j < a.pool_h; ++j)
                  for (int l = 0; l < a.pool_w; ++l) {
                    const int r2 = ((pn_ * a.b_d + pd_ * a.pool_d + i) * a.b_h + ph_ * a.pool_h + j) * a.b_w + pw_ * a.pool_w + l;
                    const uint32_t sw2 = (EC == 64) ? (r2 & 7) : (EC == 32 ? ((r2 >> 1) & 3) : ((r2 >> 2) & 1));
                    uint32_t x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                                 : "r"(full + (uint32_t)r2 * (EC * 2) + (((uint32_t)cv ^ sw2) << 4)));
                    __nv_bfloat162 t;
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m0), *reinterpret_cast<__nv_bfloat162*>(&x0)); m0 = *reinterpret_cast<uint32_t*>(&t);
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m1), *reinterpret_cast<__nv_bfloat162*>(&x1)); m1 = *reinterpret_cast<uint32_t*>(&t);
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m2), *reinterpret_cast<__nv_bfloat162*>(&x2)); m2 = *reinterpret_cast<uint32_t*>(&t);
                    t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m3), *reinterpret_cast<__nv_bfloat162*>(&x3)); m3 = *reinterpret_cast<uint32_t*>(&t);
                  }
              const uint32_t swp = (EC == 64) ? (p & 7) : (EC == 32 ? ((p >> 1) & 3) : ((p >> 2) & 1));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pst + (uint32_t)p * (EC * 2) + (((uint32_t)cv ^ swp) << 4)),
                           "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (store_thread) {
          const uint32_t src = my_stg + (uint32_t)slot * slot_bytes;
          if (col_base + c0 < a.Co) {
            const int col = col_base + c0;
            if (pooled) {
              tma_store_5d(&tmap_o0, src + pool_stg_off, col, ow0 / a.pool_w, oh0 / a.pool_h, od0 / a.pool_d, on0);
            } else if (a.out_split > 0 && col >= a.out_split) {
              if (a.out_split2 > 0 && col >= a.out_split2) tma_store_5d(&tmap_o2, src, col - a.out_split2, ow0, oh0, od0, on0);
              else tma_store_5d(&tmap_o1, src, col - a.out_split, ow0, oh0, od0, on0);
            } else {
              tma_store_5d(&tmap_o0, src, col, ow0, oh0, od0, on0);
              if (has_out1) tma_store_5d(&tmap_o1, src + STG_BYTES, col, ow0, oh0, od0, on0);
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (++slot == a.nslots) slot = 0;
      }
    }
    if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// Split-K tail: out[m, c] = epilogue(sum over splits, in order, of partial[split][m][c]).  One thread per (output row, 8
// columns); rows are decoded back to NDHWC pixels through the brick tiling of the main kernel.
struct SplitKReduceArgs {
  const float* partial;
  int ksplit, part_ld;
  long long part_rows;
  int Do, Ho, Wo, Co;
  int b_n, b_d, b_h, b_w, tiles_d, tiles_h, tiles_w, n_batch;
  int out_ld;
  Epilogue ep;
};

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(SplitKReduceArgs a) {
  const int cv = (a.Co + 7) / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.part_rows * cv) return;
  const int c0 = (int)(idx % cv) * 8;
  const long long m = idx / cv;
  const int row = (int)(m % TC_BM);
  long long mt = m / TC_BM;
  const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
  const int th = (int)(mt % a.tiles_h); mt /= a.tiles_h;
  const int td = (int)(mt % a.tiles_d);
  const int tn = (int)(mt / a.tiles_d);
  const int rw = row % a.b_w, rh = (row / a.b_w) % a.b_h, rd = (row / (a.b_w * a.b_h)) % a.b_d, rn = row / (a.b_w * a.b_h * a.b_d);
  const int ow = tw * a.b_w + rw, oh = th * a.b_h + rh, od = td * a.b_d + rd, on = tn * a.b_n + rn;
  if (rn >= a.b_n || ow >= a.Wo || oh >= a.Ho || od >= a.Do || on >= a.n_batch) return;
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = 0.f;
  const float* p = a.partial + m * a.part_ld + c0;
  for (int sidx = 0; sidx < a.ksplit; ++sidx, p += a.part_rows * a.part_ld) {
    const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
    y[0] += v0.x; y[1] += v0.y; y[2] += v0.z; y[3] += v0.w; y[4] += v1.x; y[5] += v1.y; y[6] += v1.z; y[7] += v1.w;
  }
  const long long pix = (((long long)on * a.Do + od) * a.Ho + oh) * a.Wo + ow;
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(a.ep.res);
  __align__(16) __nv_bfloat16 o0[8], o1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = min(c0 + j, a.Co - 1);
    float v = a.ep.scale0 ? fmaf(y[j], __ldg(a.ep.scale0 + c), a.ep.shift0 ? __ldg(a.ep.shift0 + c) : 0.f)
                          : y[j] + (a.ep.shift0 ? __ldg(a.ep.shift0 + c) : 0.f);
    if (res) v += __bfloat162float(res[pix * a.ep.res_ld + c]);
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    if (a.ep.relu0) h = __hmax(h, __float2bfloat16_rn(0.f));
    o0[j] = h;
    if (a.ep.out1) {
      __nv_bfloat16 h1 = __float2bfloat16_rn(fmaf(v, a.ep.scale1 ? __ldg(a.ep.scale1 + c) : 1.f, a.ep.shift1 ? __ldg(a.ep.shift1 + c) : 0.f));
      if (a.ep.relu1) h1 = __hmax(h1, __float2bfloat16_rn(0.f));
      o1[j] = h1;
    }
  }
  // Co is a multiple of 8 (conv_tc_build): whole 16-byte vectors
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.ep.out0) + pix * a.out_ld + c0) = *reinterpret_cast<const uint4*>(o0);
  if (a.ep.out1)
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.ep.out1) + pix * a.ep.out1_ld + c0) = *reinterpret_cast<const uint4*>(o1);
}

// ----------------------------------------------------------------------------- host side
// Tile-count thresholds of the shared-B / twin / CTA-pair modes and the epilogue groups of the pair-packed stem; -1 = the
// built-in default.  Set through cse_tune() (tests force the modes on small shapes); never read from the environment on
// the launch path.
struct TcTune { int bshare_min_tiles = -1, twin_min_tiles = -1, pair_min_tiles = -1, stem_groups = -1; };
static TcTune tc_tune;

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

static int encode_out_map(PFN_encodeTiled enc, CUtensorMap* m, void* out, int ld, const WinGeom& g, int max_batch,
                          int ec, const int brick[4], const int pool[3] = nullptr, const int pdims[3] = nullptr) {
  // 5-D map over the NDHWC output (dims C,W,H,D,N), box = EC channels x the output brick
  // (with a fused pool: the pooled tensor and the pooled brick)
  const int Do = pool ? pdims[0] : g.Do, Ho = pool ? pdims[1] : g.Ho, Wo = pool ? pdims[2] : g.Wo;
  const int bd = pool ? brick[1] / pool[0] : brick[1], bh = pool ? brick[2] / pool[1] : brick[2],
            bw = pool ? brick[3] / pool[2] : brick[3];
  cuuint64_t dims[5] = {(cuuint64_t)g.Co, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)Do, (cuuint64_t)max_batch};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)Wo * ld * 2, (cuuint64_t)Ho * Wo * ld * 2,
                           (cuuint64_t)Do * Ho * Wo * ld * 2};
  cuuint32_t box[5] = {(cuuint32_t)ec, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, (cuuint32_t)brick[0]};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(ec), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out) failed: %d (dims %d,%d,%d,%d,%d ld %d box %u,%u,%u,%u,%u)", (int)r, g.Co, g.Wo,
              g.Ho, g.Do, max_batch, ld, box[0], box[1], box[2], box[3], box[4]);
    return CSE_ERR_CUDA;
  }
  return CSE_OK;
}

int conv_tc_build(ConvTcDesc* d, const void* in, const void* w_packed, void* out0, void* out1, int out1_ld,
                  int max_batch, const WinGeom& g, int kc, int bn, const int brick[4], int halo, const int pool[3],
                  const int pool_dims[3], int pool_zero, int pair_pool, int out_split, int out_split2, void* out2, int out2_ld,
                  int ksplit, void* partial, size_t partial_bytes) {
  CSE_REQUIRE(kc == 16 || kc == 32 || kc == 64, "conv_tc: kc=%d must be 16/32/64", kc);
  CSE_REQUIRE(bn >= 16 && bn <= 256 && bn % 16 == 0, "conv_tc: bn=%d must be a multiple of 16 in [16,256]", bn);
  CSE_REQUIRE(g.Ci % 8 == 0 && g.in_ld % 8 == 0, "conv_tc: Cin=%d / ld=%d must be multiples of 8", g.Ci, g.in_ld);
  CSE_REQUIRE(g.in_wpitch == 0 || g.in_wpitch * g.in_ld >= (g.Wi - 1) * g.in_ld + g.Ci,
              "conv_tc: row pitch %d too small for W=%d, window %d", g.in_wpitch, g.Wi, g.Ci);
  CSE_REQUIRE(g.Co % 8 == 0 && g.out_ld % 8 == 0, "conv_tc: Cout=%d / ld=%d must be multiples of 8", g.Co, g.out_ld);
  CSE_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)w_packed % 16) == 0 && ((uintptr_t)out0 % 16) == 0 &&
                  ((uintptr_t)out1 % 16) == 0 && out1_ld % 8 == 0,
              "conv_tc: pointers must be 16B aligned");
  CSE_REQUIRE(g.sd >= 1 && g.sd <= 8 && g.sh >= 1 && g.sh <= 8 && g.sw >= 1 && g.sw <= 8, "conv_tc: stride out of range");
  const int rows = brick[0] * brick[1] * brick[2] * brick[3];
  CSE_REQUIRE(rows >= 1 && rows <= TC_BM, "conv_tc: brick %dx%dx%dx%d exceeds 128 rows", brick[0], brick[1], brick[2], brick[3]);
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return CSE_ERR_CUDA;

  d->g = g; d->kc = kc; d->bn = bn; d->max_batch = max_batch;
  d->ksplit = 1; d->partial = nullptr;
  d->n_tiles_n = ceil_div(g.Co, bn);
  d->kchunks = ceil_div(g.Ci, kc);
  for (int i = 0; i < 4; ++i) d->brick[i] = brick[i];
  d->tiles_d = ceil_div(g.Do, brick[1]);
  d->tiles_h = ceil_div(g.Ho, brick[2]);
  d->tiles_w = ceil_div(g.Wo, brick[3]);
  d->ec = (bn % 64 == 0) ? 64 : (bn % 32 == 0 ? 32 : 16);
  d->out_split = out_split; d->out_split2 = out_split > 0 ? out_split2 : 0;
  if (out_split > 0) {
    // fused sibling 1x1 convs: columns [0, split) -> out0, [split, split2 or Co) -> out1, [split2, Co) -> out2; a store
    // chunk must not straddle a split
    CSE_REQUIRE(out_split % 16 == 0 && out_split < g.Co && out1 != nullptr && ((uintptr_t)out1 % 16) == 0 && out1_ld % 8 == 0 &&
                    !(pool && pool[0] > 0) && !pair_pool,
                "conv_tc: bad output split %d for Cout=%d", out_split, g.Co);
    CSE_REQUIRE(out_split2 == 0 || (out_split2 % 16 == 0 && out_split2 > out_split && out_split2 < g.Co && out2 != nullptr &&
                                    ((uintptr_t)out2 % 16) == 0 && out2_ld % 8 == 0),
                "conv_tc: bad second output split %d (first %d, Cout=%d)", out_split2, out_split, g.Co);
    while (out_split % d->ec || out_split2 % d->ec) d->ec /= 2;
  }
  d->has_out1 = out1 != nullptr && out_split == 0;
  d->halo = halo;
  d->pair_pool = pair_pool;
  if (pair_pool) {
    // GEMM view: g.Wo = pixel pairs, g.Co = 2*Cout = bn = 128; pool_dims = D, H/2, W/2 (pairs)
    CSE_REQUIRE(halo == 1 && bn == 128 && g.Co == 128 && d->n_tiles_n == 1 && out1 == nullptr && pool && pool[0] == 1 &&
                    pool[1] == 2 && pool[2] == 1 && (brick[3] == 8 || brick[3] == 16) && brick[2] % 2 == 0 &&
                    brick[2] * brick[3] == TC_BM && g.Ho % 2 == 0 && g.out_ld % 8 == 0,
                "conv_tc: pair-pool stem needs halo mode, N = 128, brick (1,1,h,8|16) of 128 rows, even H");
  }
  const int taps = g.kd * g.kh * g.kw;
  if (halo == 1) {
    CSE_REQUIRE(g.kw == 1 && d->kchunks == 1 && g.sd == 1 && g.sh == 1 && g.sw == 1 && brick[0] == 1 && brick[1] == 1 &&
                    brick[3] % 8 == 0 && (g.kh * bn <= 256 || d->n_tiles_n == 1),
                "conv_tc: halo mode needs kw=1, Cin<=kc, stride 1, brick (1,1,h,w%%8==0), kh*bn<=256");
  } else if (halo == 2) {
    CSE_REQUIRE(g.kw == 1 && g.sh == 1 && g.sw == 1 && brick[0] == 1 && brick[1] == 1 && brick[3] % 8 == 0 &&
                    g.kh * bn <= 256,
                "conv_tc: h-halo mode needs kw=1, stride 1 in H/W, brick (1,1,h,w%%8==0), kh*bn<=256");
  } else if (halo == 3) {
    // h-halo with kw taps, CTA-pair kernel only (conv_tc2.cu): one stage per (fd, fw, chunk), the haloed box shifted by fw
    CSE_REQUIRE(g.sh == 1 && g.sw == 1 && brick[0] == 1 && brick[1] == 1 && brick[3] % 8 == 0 && d->n_tiles_n == 1 &&
                    bn <= 128 && !(pool && pool[0] > 0) && d->out_split2 == 0 && !pair_pool,
                "conv_tc: pair h-halo mode needs stride 1 in H/W, brick (1,1,h,w%%8==0), a single N tile <= 128, no pool, at most one split");
  } else {
    CSE_REQUIRE(halo == 0, "conv_tc: unknown halo mode %d", halo);
  }

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
    if (halo == 1) { box[2] = (cuuint32_t)(brick[2] + g.kh - 1); box[3] = (cuuint32_t)(brick[1] + g.kd - 1); }
    if (halo == 2 || halo == 3) box[2] = (cuuint32_t)(brick[2] + g.kh - 1);
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
  // B: [Cout_pad][Ktot] bf16, K-major (the pair-only mode loads B through tmap_bh, built below)
  if (halo != 3) {
    // halo mode packs B as [n_tile][tap][bn][kc]: one box = the kh taps of one fd plane
    // h-halo mode packs B as [n_tile][fd][chunk][fh][bn][kc]: one box = the kh taps of one (fd, chunk)
    const long long ktot = halo ? kc : (long long)taps * d->kchunks * kc;
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(d->n_tiles_n * bn) * (halo == 1 ? taps : (halo == 2 ? taps * d->kchunks : 1))};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    const bool resident = (halo == 1 && d->n_tiles_n == 1);       // resident B is loaded tap by tap
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)((halo && !resident) ? g.kh * bn : bn)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&d->tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B) failed: %d (ktot %lld rows %d box %d,%d)", (int)r, ktot, d->n_tiles_n * bn, kc, bn);
      return CSE_ERR_CUDA;
    }
  }
  const bool pooled = pool && pool[0] > 0;
  for (int i = 0; i < 3; ++i) { d->pool[i] = pooled ? pool[i] : 0; d->pool_dims[i] = pooled ? pool_dims[i] : 0; }
  d->pool_zero = pooled ? pool_zero : 0;
  if (pooled) {
    CSE_REQUIRE(out1 == nullptr, "conv_tc: fused pooling does not support a second output");
    CSE_REQUIRE(brick[1] % pool[0] == 0 && brick[2] % pool[1] == 0 && brick[3] % pool[2] == 0,
                "conv_tc: brick %dx%dx%d is not a multiple of the pool window %dx%dx%d", brick[1], brick[2], brick[3],
                pool[0], pool[1], pool[2]);
  }
  int rc;
  if (pair_pool) {
    // pooled tensor [N, D, H/2, W/2, Cout]; box = Cout channels x (brick_w pairs) x (brick_h / 2) rows
    WinGeom pg = g;
    pg.Co = g.Co / 2;
    const int pbrick[4] = {1, 1, brick[2], brick[3]};
    const int ppool[3] = {1, 2, 1};
    rc = encode_out_map(enc, &d->tmap_o0, out0, g.out_ld, pg, max_batch, 64, pbrick, ppool, pool_dims);
    if (rc) return rc;
    d->tmap_o1 = d->tmap_o0;
    d->tmap_o2 = d->tmap_o0;
  } else if (out_split > 0) {
    WinGeom og = g;
    og.Co = out_split;
    if ((rc = encode_out_map(enc, &d->tmap_o0, out0, g.out_ld, og, max_batch, d->ec, brick))) return rc;
    og.Co = (d->out_split2 > 0 ? d->out_split2 : g.Co) - out_split;
    if ((rc = encode_out_map(enc, &d->tmap_o1, out1, out1_ld, og, max_batch, d->ec, brick))) return rc;
    d->tmap_o2 = d->tmap_o1;
    if (d->out_split2 > 0) {
      og.Co = g.Co - d->out_split2;
      if ((rc = encode_out_map(enc, &d->tmap_o2, out2, out2_ld, og, max_batch, d->ec, brick))) return rc;
    }
  } else {
    rc = pooled ? encode_out_map(enc, &d->tmap_o0, out0, g.out_ld, g, max_batch, d->ec, brick, pool, pool_dims)
                : encode_out_map(enc, &d->tmap_o0, out0, g.out_ld, g, max_batch, d->ec, brick);
    if (rc) return rc;
    rc = encode_out_map(enc, &d->tmap_o1, d->has_out1 ? out1 : out0, d->has_out1 ? out1_ld : g.out_ld, g, max_batch, d->ec,
                        brick);
    if (rc) return rc;
    d->tmap_o2 = d->tmap_o0;
  }

  // shared memory: [pipeline stages][epilogue staging slots]; 227 KB per CTA minus static + alignment slack
  size_t a_stage = (size_t)TC_BM * kc * 2;
  size_t b_stage = (((size_t)bn * kc * 2) + 1023) & ~(size_t)1023;
  d->a_bytes = (uint32_t)(rows * kc * 2);
  d->b_bytes = (uint32_t)(bn * kc * 2);
  if (halo == 2 || halo == 3) {
    const size_t halo_rows = (size_t)brick[3] * (brick[2] + g.kh - 1);
    d->a_bytes = (uint32_t)(halo_rows * kc * 2);
    d->b_bytes = (uint32_t)((size_t)g.kh * bn * kc * 2);
    a_stage = (d->a_bytes + 1023) & ~(size_t)1023;
    b_stage = (d->b_bytes + 1023) & ~(size_t)1023;
    CSE_REQUIRE(((size_t)bn * kc * 2) % 1024 == 0, "conv_tc: h-halo B taps must be 1024-byte multiples");
  } else if (halo) {
    const size_t halo_rows = (size_t)brick[3] * (brick[2] + g.kh - 1) * (brick[1] + g.kd - 1);
    d->a_bytes = (uint32_t)(halo_rows * kc * 2);
    d->b_bytes = (uint32_t)((size_t)taps * bn * kc * 2);
    // the last tap reads 128 rows starting (kd-1) planes + (kh-1) rows in: stays inside the halo box
    a_stage = (d->a_bytes + 1023) & ~(size_t)1023;
    b_stage = (d->b_bytes + 1023) & ~(size_t)1023;
    CSE_REQUIRE(((size_t)bn * kc * 2) % 1024 == 0 || taps == 1, "conv_tc: halo B taps must be 1024-byte multiples");
  }
  d->a_stage = (uint32_t)a_stage;
  d->b_resident = (halo == 1 && d->n_tiles_n == 1) ? 1 : 0;
  const size_t resident = d->b_resident ? b_stage : 0;
  const size_t stage = d->b_resident ? a_stage : a_stage + b_stage;
  d->stage_bytes = (uint32_t)stage;
  size_t slot = (size_t)TC_BM * d->ec * 2 * (d->has_out1 ? 2 : 1);
  if (pooled) slot += (((size_t)TC_BM / (pool[0] * pool[1] * pool[2])) * d->ec * 2 + 1023) & ~(size_t)1023;
  if (pair_pool) slot = (size_t)(TC_BM / 2) * 64 * 2;           // only the pooled [64 rows][64 ch] tile is staged
  // staging slots per epilogue group (two groups): as many as fit (<= 4) while the load pipeline
  // keeps >= 4 stages (3 for the widest tiles); a TMA store only releases its slot once it has
  // read it, so more slots = more stores in flight
  d->slot_bytes = (uint32_t)slot;
  // epilogue groups: 2.  The pair-packed C3D stem can run 4 (cse_tune("stem_groups", 4)); measured no gain (conv1 5.99
  // vs 6.00 ms per 4 x 256 clips): the tile's 128 x 128 fp32 accumulator (64 KB) leaves TMEM at ~64 B/clk = 1024 cycles
  // against 576 cycles of MMA, and MMA accumulation and tcgen05.ld share the TMEM port - conv1 is TMEM-drain bound.
  d->groups = (pair_pool && tc_tune.stem_groups == 4) ? 4 : 2;
  const size_t groups = (size_t)d->groups;
  auto layout = [&](size_t stage_sz, int* out_stages, int* out_nslots, size_t* out_staging) -> bool {
    const int want_stages = (stage_sz >= 48 * 1024) ? 3 : 4;
    int stages = 0, nslots = 0;
    size_t staging = 0;
    for (int ns = 4; ns >= 1; --ns) {
      staging = slot * ns * groups;
      if (staging + resident + 2 * stage_sz > (214 - 8 * (groups - 2) / 2) * 1024) continue;   // s_par grows with the groups
      stages = (int)((214 * 1024 - staging - resident) / stage_sz);
      nslots = ns;
      if (stages >= want_stages) break;
    }
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    *out_stages = stages; *out_nslots = nslots; *out_staging = staging;
    return stages >= 2;
  };
  int stages = 0;
  size_t staging = 0;
  const bool fits = layout(stage, &stages, &d->nslots, &staging);
  CSE_REQUIRE(fits || halo == 3, "conv_tc: tile too large for shared memory");
  d->stages = stages;
  d->b_region = (uint32_t)(stage * stages);
  d->stage_region = (uint32_t)(stage * stages + resident);
  // twin-tile layout (two A tiles + one B tile per stage): usable when there is a single N tile of
  // <= 128 columns (four accumulators fit in TMEM) on the generic (non-halo) path
  d->twin_ok = 0;
  if (halo == 0 && !pair_pool && d->n_tiles_n == 1 && bn <= 128) {
    const size_t tstage = 2 * a_stage + b_stage;
    int tst = 0, tns = 0;
    size_t tstaging = 0;
    if (layout(tstage, &tst, &tns, &tstaging) && tst >= 3) {
      d->twin_ok = 1;
      d->tw_stage_bytes = (uint32_t)tstage;
      d->tw_stages = tst;
      d->tw_nslots = tns;
      d->tw_stage_region = (uint32_t)(tstage * tst);
      d->tw_smem_bytes = tstage * tst + tstaging + 1024;
    }
  }
  // shared-B layout for the h-halo mode (single N tile): an A ring of `bs_stages` haloed boxes, a B ring of
  // two weight blocks each used by G consecutive tiles of the CTA, 2G accumulators of 512/(2G) TMEM columns
  d->bs_group = 0;
  if (halo == 2 && d->n_tiles_n == 1 && bn <= 128 && kc <= 32) {
    // (kc = 64 stems are bound by the MMA operand reads, not by the weight stream: measured slower there)
    const int G = bn <= 64 ? 4 : 2;
    for (int nb_ = 3; nb_ >= 2 && !d->bs_group; --nb_)
      for (int ns = 2; ns >= 1 && !d->bs_group; --ns) {
        const size_t stg = slot * ns * 2;
        if (stg + nb_ * b_stage + (size_t)(G + 2) * a_stage > 214 * 1024) continue;
        int st = (int)((214 * 1024 - stg - nb_ * b_stage) / a_stage);
        if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
        d->bs_group = G;
        d->bs_slots = nb_;
        d->bs_stages = st;
        d->bs_nslots = ns;
        d->bs_b_region = (uint32_t)(a_stage * st);
        d->bs_b_stage = (uint32_t)b_stage;
        d->bs_stage_region = (uint32_t)(a_stage * st + nb_ * b_stage);
        d->bs_smem_bytes = a_stage * st + nb_ * b_stage + stg + 1024;
      }
  }
  if (ksplit > 1) {
    // split-K: generic mode only, plain tensors (no fused pool / sibling split / pair-packed stem)
    const int ksteps = taps * d->kchunks;
    CSE_REQUIRE(halo == 0 && !pooled && out_split == 0 && !pair_pool && ksplit <= 16 && ksplit <= ksteps && partial != nullptr &&
                    ((uintptr_t)partial % 16) == 0,
                "conv_tc: split-K (%d) needs the generic mode, plain outputs and a 16-byte aligned fp32 partial buffer", ksplit);
    const long long m_tiles = (long long)ceil_div(max_batch, brick[0]) * d->tiles_d * d->tiles_h * d->tiles_w;
    d->part_rows = m_tiles * TC_BM;
    d->part_ld = d->n_tiles_n * bn;
    CSE_REQUIRE((size_t)ksplit * d->part_rows * d->part_ld * sizeof(float) <= partial_bytes,
                "conv_tc: split-K partial buffer of %zu bytes is too small", partial_bytes);
    d->ksplit = ksplit;
    d->partial = reinterpret_cast<float*>(partial);
    d->twin_ok = 0;
  }
  // CTA-pair layout (conv_tc2.cu): h-halo mode, single N tile, no pool / second output / split: every CTA of a
  // 2-CTA cluster stages its own A box and HALF of the weight taps (bn/2 rows each)
  d->pair_ok = 0;
  if ((halo == 2 || halo == 3) && d->n_tiles_n == 1 && bn <= 128 && bn % 16 == 0 && (kc == 64 || kc == 32) && !pooled &&
      (out_split == 0 || (halo == 3 && d->out_split2 == 0)) && !pair_pool && brick[0] == 1 && brick[1] == 1 &&
      (((size_t)(bn / 2) * kc * 2) % (kc == 64 ? 1024 : 512)) == 0) {
    cuuint64_t dims[2] = {(cuuint64_t)kc, (cuuint64_t)bn * taps * d->kchunks};
    cuuint64_t strides[1] = {(cuuint64_t)kc * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(bn / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&d->tmap_bh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B half) failed: %d", (int)r);
      return CSE_ERR_CUDA;
    }
    const size_t bh_stage = ((size_t)g.kh * (bn / 2) * kc * 2 + 1023) & ~(size_t)1023;
    const size_t pstage = a_stage + bh_stage;
    for (int ns = 4; ns >= 1 && !d->pair_ok; --ns) {
      const size_t stg = slot * ns * 2;
      if (stg + 3 * pstage > 214 * 1024) continue;
      int st_ = (int)((214 * 1024 - stg) / pstage);
      if (st_ > 8) st_ = 8;
      if (st_ < 4 && ns > 1) continue;                   // prefer a deeper load ring over more store slots
      d->pair_ok = 1;
      d->p2_stages = st_;
      d->p2_nslots = ns;
      d->p2_stage_bytes = (uint32_t)pstage;
      d->p2_stage_region = (uint32_t)(pstage * st_);
      d->p2_smem_bytes = pstage * st_ + stg + 1024;
    }
  }
  if (halo == 3) {
    CSE_REQUIRE(d->pair_ok, "conv_tc: pair h-halo mode does not fit (kc=%d bn=%d brick %dx%d)", kc, bn, brick[2], brick[3]);
    d->tmap_b = d->tmap_bh;
  }
  d->smem_bytes = stage * stages + resident + staging + 1024;   // + alignment slack
  return CSE_OK;
}

template <int KC, int EC, int NG = 2>
static int launch_tc_t(const ConvTcDesc& d, const ConvTcArgs& args, int grid, size_t smem_bytes, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.need()) {
    CSE_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KC, EC, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (217 - 8 * (NG - 2) / 2) * 1024));
    once.mark();
  }
  conv_tc_kernel<KC, EC, NG><<<grid, 64 + 128 * NG, smem_bytes, st>>>(d.tmap_a, d.tmap_b, d.tmap_o0, d.tmap_o1, d.tmap_o2, args);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

template <int KC>
static int launch_tc_kc(const ConvTcDesc& d, const ConvTcArgs& a, int grid, size_t smem_bytes, cudaStream_t st) {
  switch (d.ec) {
    case 64: return launch_tc_t<KC, 64>(d, a, grid, smem_bytes, st);
    case 32: return launch_tc_t<KC, 32>(d, a, grid, smem_bytes, st);
    default: return launch_tc_t<KC, 16>(d, a, grid, smem_bytes, st);
  }
}

int conv_tc_tune(const char* key, int value) {
  if (!strcmp(key, "bshare_min_tiles")) { tc_tune.bshare_min_tiles = value; return CSE_OK; }
  if (!strcmp(key, "twin_min_tiles")) { tc_tune.twin_min_tiles = value; return CSE_OK; }
  if (!strcmp(key, "pair_min_tiles")) { tc_tune.pair_min_tiles = value; return CSE_OK; }
  if (!strcmp(key, "stem_groups")) { tc_tune.stem_groups = value; return CSE_OK; }     // read at plan finalize
  return CSE_ERR_INVALID;
}

int launch_conv_tc(const ConvTcDesc& d, int n, const Epilogue& ep, int sm_count, cudaStream_t st) {
  CSE_REQUIRE(n >= 0 && n <= d.max_batch, "conv_tc: n=%d exceeds max_batch=%d", n, d.max_batch);
  CSE_REQUIRE((ep.out1 != nullptr) == d.has_out1, "conv_tc: second output does not match the built descriptor");
  if (n == 0) return CSE_OK;
  const WinGeom& g = d.g;
  {
    // CTA-pair (cta_group::2) path for the small-N h-halo layers when there are enough tiles to fill every SM pair
    const long long tiles = (long long)n * d.tiles_d * d.tiles_h * d.tiles_w;
    const int pair_min = tc_tune.pair_min_tiles >= 0 ? tc_tune.pair_min_tiles : 2 * sm_count;
    if (d.pair_ok && (d.halo == 3 || (pair_min > 0 && tiles >= pair_min)))
      return launch_conv_tc_pair(d, n, ep, sm_count, st);
  }
  ConvTcArgs a;
  a.Do = g.Do; a.Ho = g.Ho; a.Wo = g.Wo; a.Co = g.Co;
  a.kd = g.kd; a.kh = g.kh; a.kw = g.kw; a.sd = g.sd; a.sh = g.sh; a.sw = g.sw;
  a.pd = g.pd; a.ph = g.ph; a.pw = g.pw;
  a.kchunks = d.kchunks; a.bn = d.bn; a.n_tiles_n = d.n_tiles_n;
  a.b_n = d.brick[0]; a.b_d = d.brick[1]; a.b_h = d.brick[2]; a.b_w = d.brick[3];
  a.tiles_d = d.tiles_d; a.tiles_h = d.tiles_h; a.tiles_w = d.tiles_w;
  a.n_batch = n;
  const long long m_tiles = (long long)ceil_div(n, d.brick[0]) * d.tiles_d * d.tiles_h * d.tiles_w;
  CSE_REQUIRE(m_tiles * d.n_tiles_n * d.ksplit < (1LL << 31), "conv_tc: too many tiles");
  a.ksplit = d.ksplit; a.partial = d.partial; a.part_ld = d.part_ld; a.part_rows = d.part_rows;
  a.ksteps_per = ceil_div(g.kd * g.kh * g.kw * d.kchunks, d.ksplit);
  if (d.ksplit > 1) a.n_tiles_n = d.n_tiles_n * d.ksplit;        // the N-tile digit of the tile walk carries the split
  a.num_tiles = (int)(m_tiles * a.n_tiles_n);
  a.stages = d.stages;
  a.a_bytes = d.a_bytes; a.b_bytes = d.b_bytes;
  a.a_stage = d.a_stage; a.stage_bytes = d.stage_bytes;
  a.halo = d.halo; a.b_resident = d.b_resident; a.b_region = d.b_region;
  a.pool_d = d.pool[0]; a.pool_h = d.pool[1]; a.pool_w = d.pool[2]; a.pool_zero = d.pool_zero;
  a.pair_pool = d.pair_pool;
  a.out_split = d.out_split; a.out_split2 = d.out_split2;
  a.stage_region = d.stage_region;
  a.nslots = d.nslots; a.slot_bytes = d.slot_bytes;
  a.ep = ep;
  int grid = a.num_tiles < sm_count ? a.num_tiles : sm_count;
  size_t smem_bytes = d.smem_bytes;
  a.twin = 0;
  a.bshare = 0; a.b_slots = 2; a.b_stage = 0; a.nbuf = 2; a.acc_cols = 256;
  if (d.groups == 4) { a.nbuf = 4; a.acc_cols = 128; }        // pair-packed stem: four 128-column accumulators
  const int bs_min = tc_tune.bshare_min_tiles >= 0 ? tc_tune.bshare_min_tiles : 8 * sm_count;
  if (d.bs_group && bs_min > 0 && a.num_tiles >= bs_min) {
    a.bshare = d.bs_group; a.b_slots = d.bs_slots;
    a.nbuf = 2 * d.bs_group; a.acc_cols = 512u / (uint32_t)a.nbuf;
    a.stages = d.bs_stages; a.stage_bytes = d.a_stage; a.b_region = d.bs_b_region; a.b_stage = d.bs_b_stage;
    a.stage_region = d.bs_stage_region; a.nslots = d.bs_nslots;
    smem_bytes = d.bs_smem_bytes;
  }
  const int twin_min = tc_tune.twin_min_tiles >= 0 ? tc_tune.twin_min_tiles : 2 * sm_count;
  if (d.twin_ok && twin_min > 0 && a.num_tiles >= twin_min) {
    // enough tiles to keep every SM busy with tile pairs
    a.twin = 1;
    a.nbuf = 4; a.acc_cols = 128;
    a.stages = d.tw_stages; a.stage_bytes = d.tw_stage_bytes; a.stage_region = d.tw_stage_region; a.nslots = d.tw_nslots;
    smem_bytes = d.tw_smem_bytes;
    const int pairs = (a.num_tiles + 1) / 2;
    grid = pairs < sm_count ? pairs : sm_count;
  }
  {
    const int radix[4] = {a.n_tiles_n, a.tiles_w, a.tiles_h, a.tiles_d};
    for (int which = 0; which < 4; ++which) {
      int v = grid * (which == 3 ? d.groups : (which == 2 ? (a.bshare ? a.bshare : 1) : which + 1));
      int* st = which == 3 ? a.stepE : (which == 2 ? a.stepG : (which ? a.step2 : a.step1));
      for (int i = 0; i < 4; ++i) { st[i] = v % radix[i]; v /= radix[i]; }
      st[4] = v;
    }
  }
  if (d.groups == 4) {
    CSE_REQUIRE(d.pair_pool && d.kc == 16 && d.ec == 64, "conv_tc: four epilogue groups are only built for the pair-packed stem");
    return launch_tc_t<16, 64, 4>(d, a, grid, smem_bytes, st);
  }
  int rc;
  switch (d.kc) {
    case 64: rc = launch_tc_kc<64>(d, a, grid, smem_bytes, st); break;
    case 32: rc = launch_tc_kc<32>(d, a, grid, smem_bytes, st); break;
    default: rc = launch_tc_kc<16>(d, a, grid, smem_bytes, st); break;
  }
  if (rc || d.ksplit <= 1) return rc;
  SplitKReduceArgs r;
  r.partial = d.partial; r.ksplit = d.ksplit; r.part_ld = d.part_ld; r.part_rows = d.part_rows;
  r.Do = g.Do; r.Ho = g.Ho; r.Wo = g.Wo; r.Co = g.Co;
  r.b_n = d.brick[0]; r.b_d = d.brick[1]; r.b_h = d.brick[2]; r.b_w = d.brick[3];
  r.tiles_d = d.tiles_d; r.tiles_h = d.tiles_h; r.tiles_w = d.tiles_w; r.n_batch = n;
  r.out_ld = g.out_ld; r.ep = ep;
  const long long rows = m_tiles * TC_BM;           // rows of the tiles this launch ran (n <= max_batch)
  r.part_rows = d.part_rows;
  const long long items = rows * ((g.Co + 7) / 8);
  SplitKReduceArgs rr = r;
  splitk_reduce_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(rr);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

}  // namespace cse
