// Shared helpers for libcse_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/cse.h"

namespace cse {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define CSE_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) return ::cse::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define CSE_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      ::cse::set_error(__VA_ARGS__);              \
      return CSE_ERR_INVALID;                     \
    }                                             \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute (opt-in dynamic shared memory) is a per-DEVICE setting: one flag per device ordinal and
// per kernel instantiation.  A racing second thread at worst sets the attribute twice.
struct PerDeviceOnce {
  bool done[64] = {};
  int dev = 0;
  bool need() {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    dev &= 63;
    return !done[dev];
  }
  void mark() { done[dev] = true; }
};
static inline size_t dtype_size(int dt) { return dt == CSE_F32 ? 4 : (dt == CSE_BF16 ? 2 : 1); }

// ---- dtype load/store helpers ------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Geometry of one conv / pool launch (per clip dims; batch n is separate).
struct WinGeom {
  int in_wpitch = 0;     // input row pitch in pixels (0 = Wi): rows of a W-padded tensor
  int Di, Hi, Wi, Ci, in_ld;
  int Do, Ho, Wo, Co, out_ld;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
};

// Epilogue description shared by both conv engines.
struct Epilogue {
  const float* scale0;   // [Co] or null (=1)
  const float* shift0;   // [Co] or null (=0)
  const float* scale1;   // [Co] or null
  const float* shift1;   // [Co] or null
  const void* res;       // residual (same dtype as out), or null
  int res_ld;
  void* out0;
  void* out1;            // null = none
  int out1_ld;
  int relu0, relu1;
};

// ---- launchers implemented in the .cu files ------------------------------------
int launch_conv_direct(int in_dt, int w_dt, int out_dt, const void* in, const void* w, int n,
                       const WinGeom& g, const Epilogue& ep, cudaStream_t st);
int launch_pool(int dt, bool is_max, bool pad_is_zero, const void* in, void* out, int n,
                const WinGeom& g, cudaStream_t st);
int launch_affine(int dt, const void* in, int in_ld, void* out, int out_ld, long long pixels, int C,
                  const float* scale, const float* shift, int relu, cudaStream_t st);
int launch_add(int dt, const void* a, int a_ld, const void* b, int b_ld, void* out, int out_ld,
               long long pixels, int C, cudaStream_t st);
int launch_softmax(const float* in, float* out, int rows, int C, cudaStream_t st);
int launch_preprocess(const uint8_t* src, int n, int T, int H, int W, int C, int t0, int h0, int w0,
                      int To, int Ho, int Wo, const float* mean3, const float* scale3, void* out,
                      int out_dt, int out_ld, cudaStream_t st, int wpitch = 0, int wpad = 0, int unroll_w = 0, int s2d = 0,
                      int src_dt = CSE_U8);

// tcgen05 engine
struct ConvTcDesc {            // built once at plan finalize
  CUtensorMap tmap_a, tmap_b, tmap_o0, tmap_o1, tmap_o2;
  int ec, nslots;              // epilogue chunk width (channels per TMA store) and staging slots
  uint32_t slot_bytes;
  bool has_out1;
  int out_split, out_split2;   // fused sibling 1x1 convs: columns >= split go to out1, columns >= split2 to out2
  int pair_pool;               // pair-packed stem with the (1,2,2) max-pool done in registers
  int groups;                  // epilogue groups of 4 warps (2; 4 for the pair-packed stem)
  int ksplit;                  // split-K factor (1 = none); partial = fp32 [ksplit][part_rows][part_ld] in the workspace
  float* partial;
  int part_ld;
  long long part_rows;
  int bs_group;                // shared-B h-halo layout: tiles per weight stage (0 = not available)
  int bs_stages, bs_nslots, bs_slots;
  uint32_t bs_b_region, bs_b_stage, bs_stage_region;
  size_t bs_smem_bytes;
  CUtensorMap tmap_bh;         // CTA-pair mode: weight map with a box of bn/2 rows (each CTA loads half of every tap)
  int pair_ok;                 // CTA-pair (cta_group::2) layout available: h-halo mode, single N tile, plain epilogue
  int p2_stages, p2_nslots;
  uint32_t p2_stage_bytes, p2_stage_region;
  size_t p2_smem_bytes;
  int twin_ok;                 // twin-tile layout available (two M tiles share every B stage)
  uint32_t tw_stage_bytes, tw_stage_region;
  int tw_stages, tw_nslots;
  size_t tw_smem_bytes;
  int halo;                    // (kd,kh)-halo'd A brick: one pipeline stage per tile
  int b_resident;              // halo + single N tile: weights stay in smem for the CTA's lifetime
  int pool[3], pool_dims[3], pool_zero;   // fused MaxPooling3D (window == stride) and its output dims
  uint32_t stage_region, a_bytes, b_bytes, a_stage, stage_bytes, b_region;
  WinGeom g;
  int kc, bn, n_tiles_n;       // K chunk (channels), N tile, number of N tiles
  int kchunks;                 // chunks per tap = ceil(Ci/kc)
  int brick[4];                // n,d,h,w
  int tiles_d, tiles_h, tiles_w;
  int stages;
  size_t smem_bytes;
  int max_batch;
};
int conv_tc_build(ConvTcDesc* d, const void* in, const void* w_packed, void* out0, void* out1, int out1_ld,
                  int max_batch, const WinGeom& g, int kc, int bn, const int brick[4], int halo, const int pool[3],
                  const int pool_dims[3], int pool_zero, int pair_pool, int out_split, int out_split2, void* out2, int out2_ld,
                  int ksplit = 1, void* partial = nullptr, size_t partial_bytes = 0);
int launch_conv_tc(const ConvTcDesc& d, int n, const Epilogue& ep, int sm_count, cudaStream_t st);
int launch_conv_tc_pair(const ConvTcDesc& d, int n, const Epilogue& ep, int sm_count, cudaStream_t st);   // conv_tc2.cu

}  // namespace cse
