// CTA-pair (tcgen05 cta_group::2) variant of the h-halo implicit-GEMM Conv3D, for the layers whose N tile is
// small (Cout = 64: the 7x7x7 / stride-2 stems of I3D and R3D, train.py:1026, 1481).
//
// Why: with N = 64 a 128x64x16 MMA is 32 cycles of math but reads 4 KB of A + 2 KB of B from shared memory,
// 48 cycles at the 128 B/clk port (profiles/r1_i3d_stem_ncu.csv: tensor pipe 51 %, tc smem wavefronts 75 %).
// Two CTAs of a cluster (the two SMs of a TPC) issue ONE 256 x N x 16 MMA: every CTA still feeds its own 128
// rows of A, but only HALF of B (N/2 weight rows) - the other half comes from the peer's shared memory - so the
// per-SM operand read drops to 5 KB = 40 cycles.  Each CTA also only streams half of the weights from L2.
//
//   cluster = 2 CTAs (rank 0 = leader).  Tile pair tp (tiles 2tp, 2tp+1) -> CTA rank r runs tile 2tp + r.
//   warp 0 (both CTAs)  TMA producer: own A box + own half of the weight taps; every load completes on the
//                       LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2, remote mbarrier)
//   warp 1 (leader)     MMA issuer: tcgen05.mma.cta_group::2, M = 256; tcgen05.commit ... multicast::cluster
//                       releases the smem stage in BOTH CTAs and publishes the accumulators to both epilogues
//   warps 2..9 (both)   epilogue of the CTA's own 128 accumulator rows (its own TMEM): scale/shift, ReLU,
//                       bf16, swizzled staging, 5-D TMA store; hands the TMEM buffer back by a remote arrive
//                       on the leader's tmem_empty barrier
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

constexpr int P2_THREADS = 320;
constexpr int P2_MAX_STAGES = 8;
constexpr uint32_t P2_BAR_EMPTY = 8u * P2_MAX_STAGES;
constexpr uint32_t P2_BAR_TMEM_FULL = 16u * P2_MAX_STAGES;
constexpr uint32_t P2_BAR_TMEM_EMPTY = P2_BAR_TMEM_FULL + 16u;
constexpr int P2_BAR_COUNT = (int)(P2_BAR_TMEM_EMPTY + 16u) / 8;

struct ConvTc2Args {
  int Do, Ho, Wo, Co;
  int kd, kh, kw, sd, pd, ph, pw;
  int kchunks, bn;
  int b_h, b_w, tiles_d, tiles_h, tiles_w;
  int n_batch, num_tiles;
  int stages;
  uint32_t a_bytes, bh_bytes;      // TMA bytes per stage and CTA: haloed A box / kh taps of bn/2 weight rows
  uint32_t a_stage, stage_bytes;   // shared-memory bytes of the A region / of a whole stage
  uint32_t stage_region;           // bytes of the pipeline region (staging slots follow)
  int nslots;
  uint32_t slot_bytes;
  const float* scale0;
  const float* shift0;
  int relu0;
  // residual add / second BN-ReLU output (pre-activation ResNet blocks, train.py:1346, 1278)
  const float* scale1;
  const float* shift1;
  const void* res;
  int res_ld, has_out1, relu1;
  int out_split;                   // > 0: columns >= out_split are stored through tmap_o1 (two members' stems in one GEMM)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's shared memory, the bytes are counted on the barrier at
// `bar` (a shared::cluster address - the leader's barrier for both CTAs).
__device__ __forceinline__ void tma2_load_5d(uint32_t pred, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                             int c1, int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %8, 0;\n\t"
      "@q cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t pred, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                             int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(pred)
      : "memory");
}
// tcgen05.commit of the pair: arrives (once the MMAs issued so far have retired) on the barrier at the same
// shared-memory offset in BOTH CTAs.
__device__ __forceinline__ void tc2_commit(uint32_t pred, uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 m;\n\tsetp.ne.b32 q, %1, 0;\n\tmov.b16 m, 3;\n\t"
               "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
               ::"r"(bar), "r"(pred) : "memory");
}
// kh taps x NK K-steps of one stage in one asm block (cf. tc_mma_taps4): tap t reads A at a_lo + t*a_step (one brick
// row of the haloed box further down) and B at b_lo + t*b_step.
#define CSE2_MMA(ACC) "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t" \
                      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, " ACC ";\n\t"
#define CSE2_STEP "add.u32 al, al, 2;\n\tadd.u32 bl, bl, 2;\n\t" CSE2_MMA("t")
#define CSE2_TAP_FIRST "mov.b32 al, ab;\n\tmov.b32 bl, bb;\n\t" CSE2_MMA("p")
#define CSE2_TAP_NEXT "add.u32 ab, ab, %7;\n\tadd.u32 bb, bb, %8;\n\tmov.b32 al, ab;\n\tmov.b32 bl, bb;\n\t" CSE2_MMA("t")
template <int NK>
__device__ __forceinline__ void tc2_mma_tap(uint32_t pred, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi32,
                                            uint32_t idesc, uint32_t accumulate_first) {
  static_assert(NK == 2 || NK == 4, "K steps per stage chunk");
#define CSE2_HEAD                                                                   \
  "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl, ab, bb;\n\t"  \
  "setp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t"      \
  "mov.b32 ab, %1;\n\tmov.b32 bb, %2;\n\t"
#define CSE2_ARGS ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi32), "r"(idesc), "r"(accumulate_first), "r"(pred) : "memory"
  if (NK == 2)
    asm volatile(CSE2_HEAD CSE2_TAP_FIRST CSE2_STEP "}" CSE2_ARGS);
  else
    asm volatile(CSE2_HEAD CSE2_TAP_FIRST CSE2_STEP CSE2_STEP CSE2_STEP "}" CSE2_ARGS);
#undef CSE2_ARGS
#undef CSE2_HEAD
}
#undef CSE2_TAP_NEXT
#undef CSE2_TAP_FIRST
#undef CSE2_STEP
#undef CSE2_MMA

template <int KC, int EC>
__global__ void __launch_bounds__(P2_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bh,
                    const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o1,
                    const ConvTc2Args a) {
  constexpr uint32_t ROW_BYTES = KC * 2;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  constexpr uint32_t STG_BYTES = 128 * EC * 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[P2_BAR_COUNT];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_par[4][256];                  // scale0, shift0, scale1, shift1 (single N tile: loaded once)

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_pairs = (a.num_tiles + 1) >> 1;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t half_rows = (uint32_t)a.bn >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(bar_base + 8u * s, 1);                           // full: the leader's arrive.expect_tx (both CTAs' bytes)
      mbar_init(bar_base + P2_BAR_EMPTY + 8u * s, 1);            // empty: one multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_base + P2_BAR_TMEM_FULL + 8u * b, 1);
      mbar_init(bar_base + P2_BAR_TMEM_EMPTY + 8u * b, 8);       // 4 epilogue warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < a.bn; i += P2_THREADS) {
    const int c = min(i, a.Co - 1);
    s_par[0][i] = a.scale0 ? __ldg(a.scale0 + c) : 1.f;
    s_par[1][i] = a.shift0 ? __ldg(a.shift0 + c) : 0.f;
    s_par[2][i] = a.scale1 ? __ldg(a.scale1 + c) : 1.f;
    s_par[3][i] = a.shift1 ? __ldg(a.shift1 + c) : 0.f;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  if (tmem_base_smem != 0u) {         // 1 CTA per SM, all 512 columns: the allocation starts at TMEM address 0
    if (threadIdx.x == 0) printf("cse conv_tc_pair: unexpected TMEM base %u\n", tmem_base_smem);
    __trap();
  }

  const int nst = a.kd * a.kw * a.kchunks;   // pipeline stages per tile: one per (fd, fw, channel chunk)
  auto decode = [&](int tile, int& tw, int& th, int& td, int& tn) {
    int mt = tile;
    tw = mt % a.tiles_w; mt /= a.tiles_w;
    th = mt % a.tiles_h; mt /= a.tiles_h;
    td = mt % a.tiles_d;
    tn = mt / a.tiles_d;
  };

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    const uint32_t leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tap_bytes = half_rows * ROW_BYTES;
    for (int tp = cluster_id; tp < num_pairs; tp += num_clusters) {
      const int tile = min(2 * tp + (int)rank, a.num_tiles - 1);        // odd tile count: the last CTA re-runs a tile, unsaved
      int tw, th, td, tn;
      decode(tile, tw, th, td, tn);
      const int iw0 = tw * a.b_w - a.pw, ih0 = th * a.b_h - a.ph, id0 = td * a.sd - a.pd;
      for (int fd = 0; fd < a.kd; ++fd)
        for (int fw = 0; fw < a.kw; ++fw)                     // kw > 1: the same haloed box, shifted by the W tap
          for (int ch = 0; ch < a.kchunks; ++ch) {
            mbar_wait(bar_base + P2_BAR_EMPTY + 8u * stage, phase ^ 1u);
            const uint32_t fb = mapa_shared(bar_base + 8u * stage, 0u);   // the LEADER's full[stage]
            if (rank == 0) mbar_expect_tx_p(leader, bar_base + 8u * stage, 2u * (a.a_bytes + a.bh_bytes));
            const uint32_t sa = smem_base + stage * a.stage_bytes;
            tma2_load_5d(leader, sa, &tmap_a, fb, ch * KC, iw0 + fw, ih0, id0 + fd, tn);
            const int row0 = ((fd * a.kw + fw) * a.kchunks + ch) * a.kh * a.bn + (int)(rank * half_rows);
            for (int fh = 0; fh < a.kh; ++fh)
              tma2_load_2d(leader, sa + a.a_stage + fh * tap_bytes, &tmap_bh, fb, 0, row0 + fh * a.bn);
            if (++stage == a.stages) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0) {
      const uint32_t leader = elect_one();
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 @17, M>>4 @24 with M = 256 (128 rows per CTA)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t desc_hi32 = (uint32_t)(make_smem_desc(0, SBO, LAYOUT) >> 32);
      auto dlo = [](uint32_t saddr) -> uint32_t { return ((saddr >> 4) & 0x3FFFu) | 0x10000u; };
      const uint32_t a_fh = ((uint32_t)a.b_w * ROW_BYTES) >> 4;          // one brick row of pixels
      const uint32_t b_tap = (half_rows * ROW_BYTES) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t i = 0;
      for (int tp = cluster_id; tp < num_pairs; tp += num_clusters, ++i) {
        const uint32_t buf = i & 1u;
        mbar_wait(bar_base + P2_BAR_TMEM_EMPTY + 8u * buf, ((i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = buf * 256u;
        uint32_t acc = 0u;
        for (int st = 0; st < nst; ++st) {
          mbar_wait(bar_base + 8u * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * a.stage_bytes;
          const uint32_t al0 = dlo(sa), bl0 = dlo(sa + a.a_stage);
          for (int fh = 0; fh < a.kh; ++fh) {
            tc2_mma_tap<KC / 16>(leader, d_tmem, al0 + (uint32_t)fh * a_fh, bl0 + (uint32_t)fh * b_tap, desc_hi32, idesc, acc);
            acc = 1u;
          }
          tc2_commit(leader, bar_base + P2_BAR_EMPTY + 8u * stage);                  // frees the stage in both CTAs
          if (st == nst - 1) tc2_commit(leader, bar_base + P2_BAR_TMEM_FULL + 8u * buf);
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // =============================== epilogue (both CTAs, own accumulator rows) ===============================
    const int grp = (warp - 2) / 4;
    const int quad = warp % 4;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 64 - grp * 128;
    const bool store_thread = (et == 0);
    const uint32_t swz = (EC == 64) ? (row & 7) : (EC == 32 ? ((row >> 1) & 3) : ((row >> 2) & 1));
    const bool has_scale0 = a.scale0 != nullptr;
    const bool relu0 = a.relu0 != 0;
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(a.res);
    const bool fast = res == nullptr && !a.has_out1;
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    const uint32_t my_stg = smem_base + a.stage_region + (uint32_t)grp * (uint32_t)a.nslots * a.slot_bytes;
    const uint32_t bar_id = 1u + (uint32_t)grp;
    const uint32_t buf = (uint32_t)grp;                       // group g drains the pair's tiles g, g + 2, ... = buffer g
    const uint32_t tmem_empty_leader = mapa_shared(bar_base + P2_BAR_TMEM_EMPTY + 8u * buf, 0u);
    int slot = 0;
    uint32_t seq = 0;
    for (int tp = cluster_id + grp * num_clusters; tp < num_pairs; tp += 2 * num_clusters, ++seq) {
      const int tile = 2 * tp + (int)rank;
      const bool valid = tile < a.num_tiles;
      int tw, th, td, tn;
      decode(valid ? tile : a.num_tiles - 1, tw, th, td, tn);
      const int ow0 = tw * a.b_w, oh0 = th * a.b_h;
      mbar_wait(bar_base + P2_BAR_TMEM_FULL + 8u * buf, seq & 1u);
      tc_fence_after();
      const uint32_t t_row = ((uint32_t)(quad * 32) << 16) + buf * 256u;
      for (int c0 = 0; c0 < a.bn; c0 += EC) {
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
          // accumulator is in registers: hand the TMEM buffer (of both CTAs) back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const uint32_t s0 = my_stg + (uint32_t)slot * a.slot_bytes + (uint32_t)row * (EC * 2);
        if (fast) {
          if (has_scale0) {
            if (relu0) epi_chunk_fast<EC, true, true>(r, s_par, c0, s0, swz);
            else epi_chunk_fast<EC, true, false>(r, s_par, c0, s0, swz);
          } else {
            if (relu0) epi_chunk_fast<EC, false, true>(r, s_par, c0, s0, swz);
            else epi_chunk_fast<EC, false, false>(r, s_par, c0, s0, swz);
          }
        } else {
          // residual add (+ second output y*scale1 + shift1 -> ReLU): the two tensors of a pre-activation ResNet block
          const int ow = ow0 + (row % a.b_w), oh = oh0 + (row / a.b_w);
          const bool use_res = res != nullptr && valid && row < a.b_h * a.b_w && ow < a.Wo && oh < a.Ho;
          const long long pix = (((long long)tn * a.Do + td) * a.Ho + oh) * a.Wo + ow;
#pragma unroll
          for (int g8 = 0; g8 < EC / 8; ++g8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float acc = __uint_as_float(r[g8 * 8 + j]);
              y[j] = has_scale0 ? fmaf(acc, s_par[0][c0 + g8 * 8 + j], s_par[1][c0 + g8 * 8 + j]) : acc + s_par[1][c0 + g8 * 8 + j];
            }
            const int col = c0 + g8 * 8;
            if (use_res && col < a.Co) {
              const uint4 q0 = *reinterpret_cast<const uint4*>(res + pix * a.res_ld + col);
              const __nv_bfloat16* e0 = reinterpret_cast<const __nv_bfloat16*>(&q0);
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] += __bfloat162float(e0[j]);
            }
            uint32_t p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
              if (relu0) h = __hmax2(h, zero2);
              p[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + (((uint32_t)g8 ^ swz) << 4)), "r"(p[0]), "r"(p[1]),
                         "r"(p[2]), "r"(p[3]) : "memory");
            if (a.has_out1) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(
                    fmaf(y[2 * j], s_par[2][c0 + g8 * 8 + 2 * j], s_par[3][c0 + g8 * 8 + 2 * j]),
                    fmaf(y[2 * j + 1], s_par[2][c0 + g8 * 8 + 2 * j + 1], s_par[3][c0 + g8 * 8 + 2 * j + 1]));
                if (a.relu1) h = __hmax2(h, zero2);
                p[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + STG_BYTES + (((uint32_t)g8 ^ swz) << 4)), "r"(p[0]),
                           "r"(p[1]), "r"(p[2]), "r"(p[3]) : "memory");
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (store_thread) {
          if (valid && c0 < a.Co) {
            if (a.out_split > 0 && c0 >= a.out_split) tma_store_5d(&tmap_o1, my_stg + (uint32_t)slot * a.slot_bytes, c0 - a.out_split, ow0, oh0, td, tn);
            else tma_store_5d(&tmap_o, my_stg + (uint32_t)slot * a.slot_bytes, c0, ow0, oh0, td, tn);
            if (a.has_out1) tma_store_5d(&tmap_o1, my_stg + (uint32_t)slot * a.slot_bytes + STG_BYTES, c0, ow0, oh0, td, tn);
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
  cluster_sync_all();                 // the peer's shared memory / TMEM stay alive until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
template <int KC, int EC>
static int launch_pair_t(const ConvTcDesc& d, const ConvTc2Args& args, int grid, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.need()) {
    CSE_CUDA(cudaFuncSetAttribute(conv_tc_pair_kernel<KC, EC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 217 * 1024));
    once.mark();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(P2_THREADS);
  cfg.dynamicSmemBytes = d.p2_smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSE_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel<KC, EC>, d.tmap_a, d.tmap_bh, d.tmap_o0, d.tmap_o1, args));
  return CSE_OK;
}

int launch_conv_tc_pair(const ConvTcDesc& d, int n, const Epilogue& ep, int sm_count, cudaStream_t st) {
  const WinGeom& g = d.g;
  ConvTc2Args a;
  a.Do = g.Do; a.Ho = g.Ho; a.Wo = g.Wo; a.Co = g.Co;
  a.kd = g.kd; a.kh = g.kh; a.kw = g.kw; a.sd = g.sd; a.pd = g.pd; a.ph = g.ph; a.pw = g.pw;
  a.kchunks = d.kchunks; a.bn = d.bn;
  a.b_h = d.brick[2]; a.b_w = d.brick[3];
  a.tiles_d = d.tiles_d; a.tiles_h = d.tiles_h; a.tiles_w = d.tiles_w;
  a.n_batch = n;
  const long long m_tiles = (long long)n * d.tiles_d * d.tiles_h * d.tiles_w;
  CSE_REQUIRE(m_tiles < (1LL << 30), "conv_tc_pair: too many tiles");
  a.num_tiles = (int)m_tiles;
  a.stages = d.p2_stages;
  a.a_bytes = d.a_bytes; a.bh_bytes = d.b_bytes / 2;
  a.a_stage = d.a_stage; a.stage_bytes = d.p2_stage_bytes; a.stage_region = d.p2_stage_region;
  a.nslots = d.p2_nslots; a.slot_bytes = d.slot_bytes;
  a.scale0 = ep.scale0; a.shift0 = ep.shift0; a.relu0 = ep.relu0;
  a.scale1 = ep.scale1; a.shift1 = ep.shift1; a.res = ep.res; a.res_ld = ep.res_ld;
  a.has_out1 = ep.out1 != nullptr ? 1 : 0; a.relu1 = ep.relu1;
  a.out_split = d.out_split;
  const int pairs = (a.num_tiles + 1) / 2;
  int grid = 2 * (pairs < sm_count / 2 ? pairs : sm_count / 2);
  if (d.kc == 64) {
    switch (d.ec) {
      case 64: return launch_pair_t<64, 64>(d, a, grid, st);
      case 32: return launch_pair_t<64, 32>(d, a, grid, st);
      default: return launch_pair_t<64, 16>(d, a, grid, st);
    }
  }
  switch (d.ec) {
    case 64: return launch_pair_t<32, 64>(d, a, grid, st);
    case 32: return launch_pair_t<32, 32>(d, a, grid, st);
    default: return launch_pair_t<32, 16>(d, a, grid, st);
  }
}

}  // namespace cse
