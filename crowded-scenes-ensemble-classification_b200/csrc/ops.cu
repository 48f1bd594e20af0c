// HBM-bound kernels (preprocess, pooling, affine, add, softmax) and the CUDA-core
// implicit-GEMM Conv3D ("direct" engine: any shape, fp32 accumulate; it is the FP32
// parity path and the fallback for shapes the tcgen05 engine does not take).
//
// Semantics follow Keras 2.2.4 / TF 1.15 channels_last as used by the reference:
//   Conv3D            train.py:653-658, 1230-1258, 1294-1298   cross-correlation, zero padding
//   MaxPooling3D      train.py:1029-1187, 1233-1261, 1487      padded taps ignored
//   AveragePooling3D  train.py:1215-1217, 1504-1507            'valid' mean
//   BatchNormalization train.py:665, 1280                      x*inv + (beta-mean*inv)
#include "common.cuh"

namespace cse {

// ============================================================================
// direct conv
// ============================================================================
constexpr int DBM = 64, DBN = 64, DBK = 16;

template <typename TI, typename TW, typename TO>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const TI* __restrict__ in, const TW* __restrict__ w, long long M, int K,
                   WinGeom g, Epilogue ep) {
  __shared__ float As[DBK][DBM + 4];
  __shared__ __align__(16) float Bs[DBK][DBN + 4];
  __shared__ int r_n[DBM], r_d[DBM], r_h[DBM], r_w[DBM];

  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * DBM;
  const int n0 = blockIdx.y * DBN;

  if (tid < DBM) {
    long long m = m0 + tid;
    if (m < M) {
      int ow = (int)(m % g.Wo); long long t = m / g.Wo;
      int oh = (int)(t % g.Ho); t /= g.Ho;
      int od = (int)(t % g.Do); int nn = (int)(t / g.Do);
      r_n[tid] = nn; r_d[tid] = od * g.sd - g.pd; r_h[tid] = oh * g.sh - g.ph; r_w[tid] = ow * g.sw - g.pw;
    } else {
      r_n[tid] = -1; r_d[tid] = 0; r_h[tid] = 0; r_w[tid] = 0;
    }
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tx = tid % 16, ty = tid / 16;
  const int a_kk = tid % 16, a_r0 = tid / 16;
  const int b_kk = tid / 16, b_c = (tid % 16) * 4;
  const int khw = g.kh * g.kw;

  for (int k0 = 0; k0 < K; k0 += DBK) {
    // ---- A tile (gather) ----
    {
      int k = k0 + a_kk;
      bool kval = k < K;
      int tap = kval ? k / g.Ci : 0;
      int c = kval ? k - tap * g.Ci : 0;
      int fd = tap / khw; int rem = tap - fd * khw; int fh = rem / g.kw; int fw = rem - fh * g.kw;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r = a_r0 + 16 * j;
        float v = 0.f;
        int nn = r_n[r];
        if (kval && nn >= 0) {
          int id = r_d[r] + fd, ih = r_h[r] + fh, iw = r_w[r] + fw;
          if ((unsigned)id < (unsigned)g.Di && (unsigned)ih < (unsigned)g.Hi && (unsigned)iw < (unsigned)g.Wi) {
            long long pix = (((long long)nn * g.Di + id) * g.Hi + ih) * g.Wi + iw;
            v = to_f32(in[pix * g.in_ld + c]);
          }
        }
        As[a_kk][r] = v;
      }
    }
    // ---- B tile ----
    {
      int k = k0 + b_kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int col = n0 + b_c + j;
        float v = 0.f;
        if (k < K && col < g.Co) v = to_f32(w[(long long)k * g.Co + col]);
        Bs[b_kk][b_c + j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  TO* out0 = reinterpret_cast<TO*>(ep.out0);
  TO* out1 = reinterpret_cast<TO*>(ep.out1);
  const TO* res = reinterpret_cast<const TO*>(ep.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col >= g.Co) continue;
      float y = acc[i][j];
      if (ep.scale0) y *= ep.scale0[col];
      if (ep.shift0) y += ep.shift0[col];
      if (res) y += to_f32(res[m * ep.res_ld + col]);
      float y0 = ep.relu0 ? fmaxf(y, 0.f) : y;
      out0[m * g.out_ld + col] = from_f32<TO>(y0);
      if (out1) {
        float z = y;
        if (ep.scale1) z *= ep.scale1[col];
        if (ep.shift1) z += ep.shift1[col];
        if (ep.relu1) z = fmaxf(z, 0.f);
        out1[m * ep.out1_ld + col] = from_f32<TO>(z);
      }
    }
  }
}

// Dense with a handful of outputs (the 11-class heads: fc8 / predictions / dense_1,
// train.py:1268, 840, 1007, 1510): one block per clip row, threads stride over K reading the
// Keras [K][Co] kernel rows contiguously, block-wide reduction of Co partial sums.
constexpr int DS_MAX_CO = 16;
template <typename TI, typename TW, typename TO>
__global__ void __launch_bounds__(256)
dense_small_kernel(const TI* __restrict__ in, const TW* __restrict__ w, int K, int Co, int in_ld, int out_ld,
                   Epilogue ep) {
  __shared__ float red[8][DS_MAX_CO];
  const int row = blockIdx.x;
  const TI* x = in + (long long)row * in_ld;
  float acc[DS_MAX_CO];
#pragma unroll
  for (int j = 0; j < DS_MAX_CO; ++j) acc[j] = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float xv = to_f32(x[k]);
    const TW* wr = w + (long long)k * Co;
#pragma unroll
    for (int j = 0; j < DS_MAX_CO; ++j)
      if (j < Co) acc[j] = fmaf(xv, to_f32(wr[j]), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < DS_MAX_CO; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < DS_MAX_CO; ++j) red[warp][j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < Co) {
    const int col = threadIdx.x;
    float y = 0.f;
    for (int wi = 0; wi < 8; ++wi) y += red[wi][col];
    if (ep.scale0) y *= ep.scale0[col];
    if (ep.shift0) y += ep.shift0[col];
    if (ep.res) y += to_f32(reinterpret_cast<const TO*>(ep.res)[(long long)row * ep.res_ld + col]);
    if (ep.relu0) y = fmaxf(y, 0.f);
    reinterpret_cast<TO*>(ep.out0)[(long long)row * out_ld + col] = from_f32<TO>(y);
  }
}

template <typename TI, typename TW, typename TO>
static int conv_direct_t(const void* in, const void* w, int n, const WinGeom& g, const Epilogue& ep,
                         cudaStream_t st) {
  long long M = (long long)n * g.Do * g.Ho * g.Wo;
  int K = g.kd * g.kh * g.kw * g.Ci;
  if (M == 0) return CSE_OK;
  if (g.kd * g.kh * g.kw == 1 && g.Do * g.Ho * g.Wo == 1 && g.Di * g.Hi * g.Wi == 1 && g.Co <= DS_MAX_CO &&
      ep.out1 == nullptr) {
    dense_small_kernel<TI, TW, TO><<<(unsigned)n, 256, 0, st>>>(
        reinterpret_cast<const TI*>(in), reinterpret_cast<const TW*>(w), K, g.Co, g.in_ld, g.out_ld, ep);
    CSE_CUDA(cudaGetLastError());
    return CSE_OK;
  }
  dim3 grid((unsigned)((M + DBM - 1) / DBM), (unsigned)ceil_div(g.Co, DBN));
  conv_direct_kernel<TI, TW, TO><<<grid, 256, 0, st>>>(
      reinterpret_cast<const TI*>(in), reinterpret_cast<const TW*>(w), M, K, g, ep);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

int launch_conv_direct(int in_dt, int w_dt, int out_dt, const void* in, const void* w, int n,
                       const WinGeom& g, const Epilogue& ep, cudaStream_t st) {
  using bf = __nv_bfloat16;
  if (in_dt == CSE_F32 && w_dt == CSE_F32 && out_dt == CSE_F32)
    return conv_direct_t<float, float, float>(in, w, n, g, ep, st);
  if (in_dt == CSE_BF16 && w_dt == CSE_BF16 && out_dt == CSE_BF16)
    return conv_direct_t<bf, bf, bf>(in, w, n, g, ep, st);
  if (in_dt == CSE_BF16 && w_dt == CSE_BF16 && out_dt == CSE_F32)
    return conv_direct_t<bf, bf, float>(in, w, n, g, ep, st);
  if (in_dt == CSE_BF16 && w_dt == CSE_F32 && out_dt == CSE_F32)
    return conv_direct_t<bf, float, float>(in, w, n, g, ep, st);
  set_error("conv direct: unsupported dtype combination in=%d w=%d out=%d", in_dt, w_dt, out_dt);
  return CSE_ERR_INVALID;
}

// ============================================================================
// pooling: one thread per (output pixel, V-channel vector)
// ============================================================================
template <typename T, int V> struct Vec;
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<float, 1> { using type = float; };
template <> struct Vec<__nv_bfloat16, 8> { using type = uint4; };
template <> struct Vec<__nv_bfloat16, 1> { using type = __nv_bfloat16; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[V]) {
  typename Vec<T, V>::type raw = *reinterpret_cast<const typename Vec<T, V>::type*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = to_f32(e[i]);
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[V]) {
  typename Vec<T, V>::type raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) e[i] = from_f32<T>(v[i]);
  *reinterpret_cast<typename Vec<T, V>::type*>(p) = raw;
}

template <typename T, int V, bool IS_MAX>
__global__ void __launch_bounds__(256)
pool_kernel(const T* __restrict__ in, T* __restrict__ out, long long total, WinGeom g, int pad_is_zero) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = g.Co / V;
  int c = (int)(idx % cv) * V; long long t = idx / cv;
  int ow = (int)(t % g.Wo); t /= g.Wo;
  int oh = (int)(t % g.Ho); t /= g.Ho;
  int od = (int)(t % g.Do); long long nn = t / g.Do;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = IS_MAX ? -INFINITY : 0.f;
  bool saw_pad = false;
  for (int fd = 0; fd < g.kd; ++fd) {
    int id = od * g.sd - g.pd + fd;
    for (int fh = 0; fh < g.kh; ++fh) {
      int ih = oh * g.sh - g.ph + fh;
      for (int fw = 0; fw < g.kw; ++fw) {
        int iw = ow * g.sw - g.pw + fw;
        if ((unsigned)id < (unsigned)g.Di && (unsigned)ih < (unsigned)g.Hi && (unsigned)iw < (unsigned)g.Wi) {
          long long pix = ((nn * g.Di + id) * g.Hi + ih) * g.Wi + iw;
          float v[V];
          load_vec<T, V>(in + pix * g.in_ld + c, v);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] = IS_MAX ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
        } else {
          saw_pad = true;
        }
      }
    }
  }
  if (IS_MAX) {
    if (saw_pad && pad_is_zero) {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
  } else {
    const float inv = 1.f / (float)(g.kd * g.kh * g.kw);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= inv;
  }
  long long opix = ((nn * g.Do + od) * g.Ho + oh) * g.Wo + ow;
  store_vec<T, V>(out + opix * g.out_ld + c, acc);
}

// bf16 max pooling, W-blocked: one thread produces WT consecutive output pixels of one 8-channel
// vector.  The (kd, kh) taps are reduced once per input column (packed bf16x2 max, exact - no fp32
// round trip), then each output combines its KW columns, so a 3x3x3 / stride-1 window costs
// (WT+2)*9 16-byte loads per WT outputs instead of WT*27.  -inf ('same' padding, padded taps
// ignored) or 0 (ZeroPadding3D) padding as in pool_kernel.
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162 t;
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.x), *reinterpret_cast<__nv_bfloat162*>(&b.x)); r.x = *reinterpret_cast<uint32_t*>(&t);
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.y), *reinterpret_cast<__nv_bfloat162*>(&b.y)); r.y = *reinterpret_cast<uint32_t*>(&t);
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.z), *reinterpret_cast<__nv_bfloat162*>(&b.z)); r.z = *reinterpret_cast<uint32_t*>(&t);
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.w), *reinterpret_cast<__nv_bfloat162*>(&b.w)); r.w = *reinterpret_cast<uint32_t*>(&t);
  return r;
}

template <int KW, int SW, int WT>
__global__ void __launch_bounds__(256)
maxpool_bf16_wblock_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long total,
                           WinGeom g, int pad_is_zero, int wblocks) {
  constexpr int NCOL = (WT - 1) * SW + KW;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = g.Co / 8;
  const int c = (int)(idx % cv) * 8; long long t = idx / cv;
  const int owb = (int)(t % wblocks); t /= wblocks;
  const int oh = (int)(t % g.Ho); t /= g.Ho;
  const int od = (int)(t % g.Do); const long long nn = t / g.Do;
  const int ow0 = owb * WT;
  const int iw0 = ow0 * SW - g.pw;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  uint4 col[NCOL];
#pragma unroll
  for (int j = 0; j < NCOL; ++j) col[j] = ninf;
  bool pad_dh = false;
  // column validity / clamped offsets are tap-independent
  bool wok[NCOL];
  long long woff[NCOL];
#pragma unroll
  for (int j = 0; j < NCOL; ++j) {
    const int iw = iw0 + j;
    wok[j] = (unsigned)iw < (unsigned)g.Wi;
    woff[j] = (long long)min(max(iw, 0), g.Wi - 1) * g.in_ld;
  }
  for (int fd = 0; fd < g.kd; ++fd) {
    const int id = od * g.sd - g.pd + fd;
    const bool okd = (unsigned)id < (unsigned)g.Di;
    const int idc = min(max(id, 0), g.Di - 1);
    for (int fh = 0; fh < g.kh; ++fh) {
      const int ih = oh * g.sh - g.ph + fh;
      const bool ok = okd && (unsigned)ih < (unsigned)g.Hi;
      const int ihc = min(max(ih, 0), g.Hi - 1);
      pad_dh = pad_dh || !ok;
      const __nv_bfloat16* row = in + (((nn * g.Di + idc) * g.Hi + ihc) * (long long)g.Wi) * g.in_ld + c;
      // all NCOL loads are issued before the first max (addresses are clamped, never predicated off)
      uint4 v[NCOL];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(row + woff[j]));
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const uint4 x = (ok && wok[j]) ? v[j] : ninf;
        col[j] = hmax8(col[j], x);
      }
    }
  }
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  __nv_bfloat16* orow = out + (((nn * g.Do + od) * g.Ho + oh) * (long long)g.Wo) * g.out_ld + c;
#pragma unroll
  for (int o = 0; o < WT; ++o) {
    const int ow = ow0 + o;
    if (ow >= g.Wo) break;
    uint4 m = col[o * SW];
    bool pad = pad_dh || (unsigned)(iw0 + o * SW) >= (unsigned)g.Wi;
#pragma unroll
    for (int fw = 1; fw < KW; ++fw) {
      m = hmax8(m, col[o * SW + fw]);
      pad = pad || (unsigned)(iw0 + o * SW + fw) >= (unsigned)g.Wi;
    }
    if (pad && pad_is_zero) m = hmax8(m, zero);
    *reinterpret_cast<uint4*>(orow + (long long)ow * g.out_ld) = m;
  }
}

// 3x3x3 / stride-1 max pooling (the branch-3 pools of the Inception blocks, train.py:1066-1187), bf16,
// register-blocked in all three dims: one thread produces a DT x HT x WT block of outputs of one
// 8-channel vector.  Each input row is loaded once ((WT+2) 16-byte loads, clamped addresses, masked
// to -inf outside the tensor), reduced along W, folded into the <= 3 output rows that contain it,
// and each hw-reduced plane into the <= 3 output planes that contain it: 6 loads per output instead
// of 27 (13.5 with W blocking only).
template <int DT, int HT, int WT>
__global__ void __launch_bounds__(128)
maxpool3s1_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long total,
                       WinGeom g, int wblocks, int hblocks, int dblocks) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = g.Co / 8;
  const int c = (int)(idx % cv) * 8; long long t = idx / cv;
  const int owb = (int)(t % wblocks); t /= wblocks;
  const int ohb = (int)(t % hblocks); t /= hblocks;
  const int odb = (int)(t % dblocks); const long long nn = t / dblocks;
  const int ow0 = owb * WT, oh0 = ohb * HT, od0 = odb * DT;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  bool wok[WT + 2];
  long long woff[WT + 2];
#pragma unroll
  for (int j = 0; j < WT + 2; ++j) {
    const int iw = ow0 - g.pw + j;
    wok[j] = (unsigned)iw < (unsigned)g.Wi;
    woff[j] = (long long)min(max(iw, 0), g.Wi - 1) * g.in_ld;
  }
  uint4 acc[DT][HT][WT];
#pragma unroll
  for (int a = 0; a < DT; ++a)
#pragma unroll
    for (int b = 0; b < HT; ++b)
#pragma unroll
      for (int o = 0; o < WT; ++o) acc[a][b][o] = ninf;
#pragma unroll
  for (int pd = 0; pd < DT + 2; ++pd) {
    const int id = od0 - g.pd + pd;
    const bool okd = (unsigned)id < (unsigned)g.Di;
    const int idc = min(max(id, 0), g.Di - 1);
    uint4 hw[HT][WT];
#pragma unroll
    for (int b = 0; b < HT; ++b)
#pragma unroll
      for (int o = 0; o < WT; ++o) hw[b][o] = ninf;
#pragma unroll
    for (int ph = 0; ph < HT + 2; ++ph) {
      const int ih = oh0 - g.ph + ph;
      const bool ok = okd && (unsigned)ih < (unsigned)g.Hi;
      const int ihc = min(max(ih, 0), g.Hi - 1);
      const __nv_bfloat16* row = in + (((nn * g.Di + idc) * g.Hi + ihc) * (long long)g.Wi) * g.in_ld + c;
      uint4 v[WT + 2];
#pragma unroll
      for (int j = 0; j < WT + 2; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(row + woff[j]));
#pragma unroll
      for (int j = 0; j < WT + 2; ++j) v[j] = (ok && wok[j]) ? v[j] : ninf;
#pragma unroll
      for (int o = 0; o < WT; ++o) {
        const uint4 wm = hmax8(hmax8(v[o], v[o + 1]), v[o + 2]);
#pragma unroll
        for (int b = 0; b < HT; ++b)
          if (ph - b >= 0 && ph - b <= 2) hw[b][o] = hmax8(hw[b][o], wm);
      }
    }
#pragma unroll
    for (int a = 0; a < DT; ++a)
      if (pd - a >= 0 && pd - a <= 2) {
#pragma unroll
        for (int b = 0; b < HT; ++b)
#pragma unroll
          for (int o = 0; o < WT; ++o) acc[a][b][o] = hmax8(acc[a][b][o], hw[b][o]);
      }
  }
#pragma unroll
  for (int a = 0; a < DT; ++a) {
    const int od = od0 + a;
    if (od >= g.Do) break;
#pragma unroll
    for (int b = 0; b < HT; ++b) {
      const int oh = oh0 + b;
      if (oh >= g.Ho) break;
      __nv_bfloat16* orow = out + (((nn * g.Do + od) * g.Ho + oh) * (long long)g.Wo) * g.out_ld + c;
#pragma unroll
      for (int o = 0; o < WT; ++o)
        if (ow0 + o < g.Wo) *reinterpret_cast<uint4*>(orow + (long long)(ow0 + o) * g.out_ld) = acc[a][b][o];
    }
  }
}

// Plane-sweep variant (Mixed_3b..5c: 28x28, 14x14, 7x7 planes): a CTA owns one clip x one chunk of
// 8*vpc channels (vpc = 8: a full 128-byte line per pixel) x one strip of hs output rows x one segment of
// output planes, and walks the input planes of that segment once.  Each plane strip (+ one halo row above
// and below) goes global -> shared memory with cp.async, three stages deep.  A thread owns NG groups of
// WT neighbouring pixels of one 8-channel vector: per plane it reads 3 x (WT + 2) vectors from shared
// memory (a neighbour outside the plane is replaced by the pixel itself, which leaves the maximum
// unchanged), reduces them along H then W, and keeps the last two hw-reduced planes in registers, so
// every input element is read from HBM once (halo rows / planes are re-read through L2).
constexpr int SWEEP_STAGES = 3;
template <int WT, int NG>
__global__ void __launch_bounds__(224, 3)
maxpool3s1_sweep_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, WinGeom g,
                        int vpc, int nchunks, int hs, int nstrips, int dsegs, int dlen) {
  extern __shared__ uint4 sweep_smem[];
  const int rowv = g.Wi * vpc;                       // vectors per plane row of this chunk
  const int gpr = g.Wi / WT;                         // pixel groups per row
  int b = blockIdx.x;
  const int chunk = b % nchunks; b /= nchunks;
  const int strip = b % nstrips; b /= nstrips;
  const int seg = b % dsegs;
  const long long nn = b / dsegs;
  const int h0 = strip * hs, h1 = min(h0 + hs, g.Hi);
  const int lo = max(h0 - 1, 0), hi = min(h1 + 1, g.Hi);          // rows loaded: [lo, hi)
  const int d_lo = seg * dlen, d_hi = min(d_lo + dlen, g.Di);
  const int ngroups = (h1 - h0) * gpr * vpc;
  const int stage_v = (hs + 2) * rowv;                // smem row r holds plane row h0 - 1 + r
  const int cbase = chunk * vpc * 8;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  // per group: first pixel | channel << 16, its smem index, and 5 flag bits (active, has up / down / left /
  // right neighbour)
  uint32_t pc[NG], ci[NG];
  uint32_t flags = 0;
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int gi = threadIdx.x + k * blockDim.x;
    const int v = gi % vpc, t = gi / vpc;
    const int wg = t % gpr, hh = t / gpr;
    const int w0 = wg * WT, h = h0 + hh;
    const int c = cbase + v * 8;
    const bool act = gi < ngroups && c < g.Co;
    pc[k] = act ? ((uint32_t)(h * g.Wi + w0) | ((uint32_t)c << 16)) : 0u;
    ci[k] = act ? (uint32_t)(((hh + 1) * g.Wi + w0) * vpc + v) : 0u;
    flags |= ((act ? 1u : 0u) | (h > 0 ? 2u : 0u) | (h < g.Hi - 1 ? 4u : 0u) | (w0 > 0 ? 8u : 0u) |
              (w0 + WT < g.Wi ? 16u : 0u)) << (5 * k);
  }
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(sweep_smem);
  const int nload = (hi - lo) * rowv;
  auto load_plane = [&](int id, int stage) {          // rows [lo, hi) of plane id -> stage
    if (id >= 0 && id < g.Di && id <= d_hi) {
      const __nv_bfloat16* pl = in + ((nn * g.Di + id) * g.Hi + lo) * (long long)g.Wi * g.in_ld + cbase;
      const uint32_t dst = smem_base + (uint32_t)(stage * stage_v + (lo - (h0 - 1)) * rowv) * 16u;
      for (int j = threadIdx.x; j < nload; j += blockDim.x) {
        const int p = j / vpc, v = j - p * vpc;
        if (cbase + v * 8 < g.Co)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)j * 16u),
                       "l"(pl + (long long)p * g.in_ld + v * 8) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint4 r0[NG][WT], r1[NG][WT];
#pragma unroll
  for (int k = 0; k < NG; ++k)
#pragma unroll
    for (int o = 0; o < WT; ++o) { r0[k][o] = ninf; r1[k][o] = ninf; }
  load_plane(d_lo - 1, 0);
  load_plane(d_lo, 1);
  int stage = 0;
  for (int id = d_lo - 1; id <= d_hi; ++id) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                   // plane id has landed; everyone is done with plane id - 1
    load_plane(id + 2, stage == 0 ? SWEEP_STAGES - 1 : stage - 1);
    const bool valid = id >= 0 && id < g.Di;
    const uint4* sp = sweep_smem + stage * stage_v;
    const bool emit = id - 1 >= d_lo && id - 1 < d_hi;
    __nv_bfloat16* opl = out + (nn * g.Do + (id - 1)) * (long long)g.Hi * g.Wi * g.out_ld;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const uint32_t f = flags >> (5 * k);
      uint4 hw[WT];
#pragma unroll
      for (int o = 0; o < WT; ++o) hw[o] = ninf;
      if (valid && (f & 1u)) {
        const uint4* q = sp + ci[k];
        const int ou = (f & 2u) ? -rowv : 0, od = (f & 4u) ? rowv : 0;
        uint4 col[WT + 2];
#pragma unroll
        for (int j = 0; j < WT + 2; ++j) {
          const int cj = j == 0 ? ((f & 8u) ? -vpc : 0) : (j == WT + 1 ? ((f & 16u) ? WT * vpc : (WT - 1) * vpc) : (j - 1) * vpc);
          col[j] = hmax8(hmax8(q[ou + cj], q[cj]), q[od + cj]);
        }
#pragma unroll
        for (int o = 0; o < WT; ++o) hw[o] = hmax8(hmax8(col[o], col[o + 1]), col[o + 2]);
      }
      if (emit && (f & 1u)) {
        __nv_bfloat16* op = opl + (long long)(pc[k] & 0xFFFFu) * g.out_ld + (pc[k] >> 16);
#pragma unroll
        for (int o = 0; o < WT; ++o)
          *reinterpret_cast<uint4*>(op + (long long)o * g.out_ld) = hmax8(hmax8(r0[k][o], r1[k][o]), hw[o]);
      }
#pragma unroll
      for (int o = 0; o < WT; ++o) { r0[k][o] = r1[k][o]; r1[k][o] = hw[o]; }
    }
    stage = stage == SWEEP_STAGES - 1 ? 0 : stage + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int WT, int NG>
static int maxpool3s1_sweep_launch(const void* in, void* out, int n, const WinGeom& g, int vpc, cudaStream_t st) {
  static PerDeviceOnce once;
  // strip height: <= 224 threads x NG groups of WT pixels, equal strips
  const int per_row = (g.Wi / WT) * vpc;
  int hs = max(1, min(g.Hi, 224 * NG / per_row));
  hs = ceil_div(g.Hi, ceil_div(g.Hi, hs));
  const int threads = ceil_div(ceil_div(hs * per_row, NG), 32) * 32;
  const size_t smem = (size_t)SWEEP_STAGES * (hs + 2) * g.Wi * vpc * sizeof(uint4);
  if (threads > 224 || smem > 72 * 1024) return -1;
  if (once.need()) {
    CSE_CUDA(cudaFuncSetAttribute(maxpool3s1_sweep_kernel<WT, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    once.mark();
  }
  const int nchunks = ceil_div(g.Co / 8, vpc);
  const int nstrips = ceil_div(g.Hi, hs);
  // segments of >= 8 output planes, as many as it takes to give every SM a few CTAs
  int dsegs = 1;
  while (dsegs * 2 * 8 <= g.Di && (long long)n * nchunks * nstrips * dsegs < 8 * 148) dsegs *= 2;
  const int dlen = ceil_div(g.Di, dsegs);
  dsegs = ceil_div(g.Di, dlen);
  maxpool3s1_sweep_kernel<WT, NG><<<(unsigned)(n * nchunks * nstrips * dsegs), threads, smem, st>>>(
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, g, vpc, nchunks, hs, nstrips, dsegs, dlen);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

static int maxpool3s1(const void* in, void* out, int n, const WinGeom& g, cudaStream_t st) {
  const int vpc = min(8, g.Co / 8);
  if (n > 0 && g.Hi * g.Wi <= 65535 && g.Co <= 65528 &&
      (long long)g.Hi * g.Wi * max(g.in_ld, g.out_ld) < (1ll << 31)) {
    int rc;
    if (g.Wi % 4 == 0) rc = maxpool3s1_sweep_launch<4, 1>(in, out, n, g, vpc, st);
    else if (g.Wi % 2 == 0) rc = maxpool3s1_sweep_launch<2, 2>(in, out, n, g, vpc, st);
    else rc = maxpool3s1_sweep_launch<1, 4>(in, out, n, g, vpc, st);
    if (rc >= 0) return rc;                           // -1: the plane strip does not fit -> register-blocked kernel
  }
  constexpr int DT = 2, HT = 2, WT = 4;
  const int wb = ceil_div(g.Wo, WT), hb = ceil_div(g.Ho, HT), db = ceil_div(g.Do, DT);
  const long long total = (long long)n * db * hb * wb * (g.Co / 8);
  if (total == 0) return CSE_OK;
  maxpool3s1_bf16_kernel<DT, HT, WT><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, total, g, wb, hb, db);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

template <int KW, int SW>
static int maxpool_wblock(bool pad_is_zero, const void* in, void* out, int n, const WinGeom& g, cudaStream_t st) {
  constexpr int WT = 4;
  const int wblocks = ceil_div(g.Wo, WT);
  const long long total = (long long)n * g.Do * g.Ho * wblocks * (g.Co / 8);
  if (total == 0) return CSE_OK;
  maxpool_bf16_wblock_kernel<KW, SW, WT><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, total, g, pad_is_zero ? 1 : 0, wblocks);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

template <typename T, int V>
static int pool_t(bool is_max, bool pad_is_zero, const void* in, void* out, int n, const WinGeom& g,
                  cudaStream_t st) {
  long long total = (long long)n * g.Do * g.Ho * g.Wo * (g.Co / V);
  if (total == 0) return CSE_OK;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (is_max)
    pool_kernel<T, V, true><<<blocks, 256, 0, st>>>((const T*)in, (T*)out, total, g, pad_is_zero ? 1 : 0);
  else
    pool_kernel<T, V, false><<<blocks, 256, 0, st>>>((const T*)in, (T*)out, total, g, 0);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

static bool vec_ok(int C, int in_ld, int out_ld, int V, const void* a, const void* b, size_t es) {
  return C % V == 0 && in_ld % V == 0 && out_ld % V == 0 &&
         ((uintptr_t)a % (V * es)) == 0 && ((uintptr_t)b % (V * es)) == 0;
}

int launch_pool(int dt, bool is_max, bool pad_is_zero, const void* in, void* out, int n,
                const WinGeom& g, cudaStream_t st) {
  CSE_REQUIRE(g.Ci == g.Co, "pool: channel mismatch %d vs %d", g.Ci, g.Co);
  if (dt == CSE_F32) {
    if (vec_ok(g.Co, g.in_ld, g.out_ld, 4, in, out, 4)) return pool_t<float, 4>(is_max, pad_is_zero, in, out, n, g, st);
    return pool_t<float, 1>(is_max, pad_is_zero, in, out, n, g, st);
  }
  if (dt == CSE_BF16) {
    if (is_max && vec_ok(g.Co, g.in_ld, g.out_ld, 8, in, out, 2)) {
      if (!pad_is_zero && g.kd == 3 && g.kh == 3 && g.kw == 3 && g.sd == 1 && g.sh == 1 && g.sw == 1 &&
          g.Do == g.Di && g.Ho == g.Hi && g.Wo == g.Wi)
        return maxpool3s1(in, out, n, g, st);
      if (g.kw == 3 && g.sw == 1) return maxpool_wblock<3, 1>(pad_is_zero, in, out, n, g, st);
      if (g.kw == 3 && g.sw == 2) return maxpool_wblock<3, 2>(pad_is_zero, in, out, n, g, st);
      if (g.kw == 2 && g.sw == 2) return maxpool_wblock<2, 2>(pad_is_zero, in, out, n, g, st);
    }
    if (vec_ok(g.Co, g.in_ld, g.out_ld, 8, in, out, 2))
      return pool_t<__nv_bfloat16, 8>(is_max, pad_is_zero, in, out, n, g, st);
    return pool_t<__nv_bfloat16, 1>(is_max, pad_is_zero, in, out, n, g, st);
  }
  set_error("pool: unsupported dtype %d", dt);
  return CSE_ERR_INVALID;
}

// ============================================================================
// affine (stand-alone BN +/- ReLU) and add
// ============================================================================
template <typename T, int V>
__global__ void __launch_bounds__(256)
affine_kernel(const T* __restrict__ in, int in_ld, T* __restrict__ out, int out_ld, long long total, int C,
              const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = C / V;
  int c = (int)(idx % cv) * V; long long pix = idx / cv;
  float v[V];
  load_vec<T, V>(in + pix * in_ld + c, v);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float y = v[i];
    if (scale) y *= scale[c + i];
    if (shift) y += shift[c + i];
    v[i] = relu ? fmaxf(y, 0.f) : y;
  }
  store_vec<T, V>(out + pix * out_ld + c, v);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
add_kernel(const T* __restrict__ a, int a_ld, const T* __restrict__ b, int b_ld, T* __restrict__ out,
           int out_ld, long long total, int C) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = C / V;
  int c = (int)(idx % cv) * V; long long pix = idx / cv;
  float x[V], y[V];
  load_vec<T, V>(a + pix * a_ld + c, x);
  load_vec<T, V>(b + pix * b_ld + c, y);
#pragma unroll
  for (int i = 0; i < V; ++i) x[i] += y[i];
  store_vec<T, V>(out + pix * out_ld + c, x);
}

int launch_affine(int dt, const void* in, int in_ld, void* out, int out_ld, long long pixels, int C,
                  const float* scale, const float* shift, int relu, cudaStream_t st) {
  if (pixels == 0) return CSE_OK;
  if (dt == CSE_F32) {
    if (vec_ok(C, in_ld, out_ld, 4, in, out, 4)) {
      long long total = pixels * (C / 4);
      affine_kernel<float, 4><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const float*)in, in_ld, (float*)out, out_ld, total, C, scale, shift, relu);
    } else {
      long long total = pixels * C;
      affine_kernel<float, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const float*)in, in_ld, (float*)out, out_ld, total, C, scale, shift, relu);
    }
  } else if (dt == CSE_BF16) {
    using bf = __nv_bfloat16;
    if (vec_ok(C, in_ld, out_ld, 8, in, out, 2)) {
      long long total = pixels * (C / 8);
      affine_kernel<bf, 8><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const bf*)in, in_ld, (bf*)out, out_ld, total, C, scale, shift, relu);
    } else {
      long long total = pixels * C;
      affine_kernel<bf, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const bf*)in, in_ld, (bf*)out, out_ld, total, C, scale, shift, relu);
    }
  } else {
    set_error("affine: unsupported dtype %d", dt);
    return CSE_ERR_INVALID;
  }
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

int launch_add(int dt, const void* a, int a_ld, const void* b, int b_ld, void* out, int out_ld,
               long long pixels, int C, cudaStream_t st) {
  if (pixels == 0) return CSE_OK;
  if (dt == CSE_F32) {
    bool v = vec_ok(C, a_ld, out_ld, 4, a, out, 4) && b_ld % 4 == 0 && ((uintptr_t)b % 16) == 0;
    if (v) {
      long long total = pixels * (C / 4);
      add_kernel<float, 4><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const float*)a, a_ld, (const float*)b, b_ld, (float*)out, out_ld, total, C);
    } else {
      long long total = pixels * C;
      add_kernel<float, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const float*)a, a_ld, (const float*)b, b_ld, (float*)out, out_ld, total, C);
    }
  } else if (dt == CSE_BF16) {
    using bf = __nv_bfloat16;
    bool v = vec_ok(C, a_ld, out_ld, 8, a, out, 2) && b_ld % 8 == 0 && ((uintptr_t)b % 16) == 0;
    if (v) {
      long long total = pixels * (C / 8);
      add_kernel<bf, 8><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const bf*)a, a_ld, (const bf*)b, b_ld, (bf*)out, out_ld, total, C);
    } else {
      long long total = pixels * C;
      add_kernel<bf, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
          (const bf*)a, a_ld, (const bf*)b, b_ld, (bf*)out, out_ld, total, C);
    }
  } else {
    set_error("add: unsupported dtype %d", dt);
    return CSE_ERR_INVALID;
  }
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

// ============================================================================
// softmax: one warp per row, fp32, stable form (Keras softmax over the last axis)
// ============================================================================
__global__ void __launch_bounds__(128)
softmax_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int C) {
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* x = in + (long long)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(x[c] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int c = lane; c < C; c += 32) out[(long long)row * C + c] = expf(x[c] - mx) / sum;
}

int launch_softmax(const float* in, float* out, int rows, int C, cudaStream_t st) {
  if (rows == 0) return CSE_OK;
  softmax_kernel<<<ceil_div(rows, 4), 128, 0, st>>>(in, out, rows, C);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

// ============================================================================
// clip pre-processing: uint8 NDHWC -> float (bf16 / fp32), optional crop + mean/scale,
// output channel count padded with zeros up to out_ld.
// Reference: frames are stored raw into an np.float32 batch (train.py:466-478):
// crop = none, mean = 0, scale = 1.
// ============================================================================
struct PreArgs {
  int T, H, W, C, t0, h0, w0, To, Ho, Wo, out_ld;
  int wpitch, wpad;      // output row pitch in pixels and zero columns on the left (>= Wo + wpad)
  float mean[4], scale[4];
};

template <typename TO, int CO, typename TS>
__global__ void __launch_bounds__(256)
preprocess_kernel(const TS* __restrict__ src, TO* __restrict__ out, long long total, PreArgs a) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int wp = (int)(idx % a.wpitch); long long t = idx / a.wpitch;
  int h = (int)(t % a.Ho); t /= a.Ho;
  int d = (int)(t % a.To); long long nn = t / a.To;
  const int w = wp - a.wpad;
  const bool real = (w >= 0 && w < a.Wo);         // pad columns of the row are written as zeros
  long long spix = ((nn * a.T + (d + a.t0)) * a.H + (h + a.h0)) * a.W + (real ? (w + a.w0) : 0);
  const TS* s = src + spix * a.C;
  __align__(16) TO v[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    float f = 0.f;
    if (real && c < a.C) f = ((float)s[c] - a.mean[c]) * a.scale[c];
    v[c] = from_f32<TO>(f);
  }
  TO* o = out + idx * a.out_ld;
  if (CO * sizeof(TO) == 16) {
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(v);
  } else {
#pragma unroll
    for (int c = 0; c < CO; ++c) o[c] = v[c];
  }
}

// ---------------------------------------------------------------------------------------------------
// Stem layouts (packed / pair-packed C3D stem, 2-D and 3-D space-to-depth for the stride-2 7x7x7 stems).
// One CTA converts PRE_RH output rows of one output plane: the source rows it needs are staged in shared
// memory with coalesced 16-byte loads (uint8 rows: W*C bytes, read exactly once), then every thread
// builds 16-byte chunks (8 bf16 channels) of output positions from the staged bytes and stores them
// with fully coalesced 128-bit stores.
//
//   MODE 3 / 4  packed stem: output pixel w carries its horizontal neighbours w-1, w, w+1 (MODE 3) or the
//               pixel pair p = (2p, 2p+1) carries pixels 2p-1 .. 2p+2 (MODE 4, a.Wo = pairs); channel
//               index j*C + c, zero-padded to 16; pixels outside the row are zeros.  The kw taps of a 3x3x3
//               stem conv (train.py:1230) then are one contiguous K chunk per position.
//   MODE 20     2x2 space-to-depth over (H, W) (7x7x7 / stride 2 stems, train.py:1026, 1481): output cell
//               (h2, w2) carries the input pixels (2*h2+ph, 2*w2+pw) x C channels at (ph*2+pw)*C + c.
//   MODE 21     2x2x2 space-to-depth over (T, H, W): cell (d2, h2, w2) carries (2*d2+pd, 2*h2+ph, 2*w2+pw)
//               at ((pd*2+ph)*2+pw)*C + c: 8*C channels without padding (C = 3: 24 channels = 48 bytes), so
//               the stride-2 stem becomes a stride-1 4x4x4-cell conv with K = 512*C instead of 7*4*4*16.
//   Pixels outside the frame, the wpad / right pad positions of a row and channels beyond the packed ones
//   are exact zeros (they meet zero weights or stand for the conv's zero padding).
// ---------------------------------------------------------------------------------------------------
constexpr int PRE_THREADS = 256;

struct PreRowArgs {
  int T, H, W;             // source clip
  int To, Ho, Wo;          // output grid: planes, rows, real positions per row
  int wpitch, wpad;        // output row: wpitch positions, the first wpad and everything beyond wpad + Wo are zeros
  int t0, h0, w0;          // crop (packed stem modes)
  int rh;                  // output rows per CTA
  int hgroups;             // ceil(Ho / rh)
  int row_bytes, row_pitch;// W*C and its 16-byte rounded smem pitch
  int aligned;             // source rows start 16-byte aligned
  float mean[4], scale[4];
};

// uint8 -> float without the conversion unit: 0x4B000000 | b is the float 2^23 + b, exactly.
__device__ __forceinline__ float u8_to_f32(uint32_t b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }

// source element -> float: uint8 frames (the reference's decoded BGR / TV-L1 gray values) or float32 (the dense flow
// of the FarneBack_onTheFly variant, train.py:294-332)
__device__ __forceinline__ float src_to_f32(const uint8_t* p) { return u8_to_f32(*p); }
__device__ __forceinline__ float src_to_f32(const float* p) { return *p; }

template <int MODE, int C, int CL, bool IDENT, typename TS>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_rows_kernel(const TS* __restrict__ src_t, __nv_bfloat16* __restrict__ out, PreRowArgs a) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(src_t);
  constexpr int ES = (int)sizeof(TS);                     // bytes per source element
  constexpr int NPL = (MODE == 21) ? 2 : 1;               // source planes per output plane
  constexpr int NR = (MODE >= 20) ? 2 : 1;                // source rows per output row
  constexpr int Q = CL / 8;                               // 16-byte chunks per output position
  constexpr int KREAL = (MODE == 21) ? 8 * C : (MODE == 20 ? 4 * C : MODE * C);
  // An output position is the concatenation of KREAL / L segments of L consecutive source bytes (one segment per
  // staged source row), then zeros: packed stem = ONE run of NB*C bytes, s2d = 2C bytes from each of its 2 / 4 rows.
  constexpr int L = (MODE >= 20) ? 2 * C : MODE * C;
  extern __shared__ __align__(16) uint8_t rows[];         // [rh][NPL][NR][row_pitch]
  int b = blockIdx.x;
  const int hg = b % a.hgroups; b /= a.hgroups;
  const int d = b % a.To;
  const long long n = b / a.To;
  const int h_first = hg * a.rh;
  const int nrows = min(a.rh, a.Ho - h_first);
  // ---- stage the source rows (each source byte leaves HBM once, 16 bytes per load) ----
  const int staged = nrows * NPL * NR;
  if (a.aligned) {
    const int v_per_row = a.row_bytes >> 4;
    for (int i = threadIdx.x; i < staged * v_per_row; i += PRE_THREADS) {
      const int r = i / v_per_row, v = i - r * v_per_row;
      const int rr = r % NR, pl = (r / NR) % NPL, rh = r / (NR * NPL);
      const int ts = (MODE == 21) ? 2 * d + pl : d + a.t0;
      const int hs = (MODE >= 20) ? 2 * (h_first + rh) + rr : h_first + rh + a.h0;
      if (ts < a.T && hs < a.H) {
        const uint4* g = reinterpret_cast<const uint4*>(src + (((n * a.T + ts) * a.H + hs) * (long long)a.W) * (C * ES));
        reinterpret_cast<uint4*>(rows + (size_t)r * a.row_pitch)[v] = __ldg(g + v);
      }
    }
  } else {
    for (int i = threadIdx.x; i < staged * a.row_bytes; i += PRE_THREADS) {
      const int r = i / a.row_bytes, v = i - r * a.row_bytes;
      const int rr = r % NR, pl = (r / NR) % NPL, rh = r / (NR * NPL);
      const int ts = (MODE == 21) ? 2 * d + pl : d + a.t0;
      const int hs = (MODE >= 20) ? 2 * (h_first + rh) + rr : h_first + rh + a.h0;
      if (ts < a.T && hs < a.H) rows[(size_t)r * a.row_pitch + v] = __ldg(src + (((n * a.T + ts) * a.H + hs) * (long long)a.W) * (C * ES) + v);
    }
  }
  __syncthreads();
  // ---- one thread per output position: Q 16-byte stores ----
  __nv_bfloat16* obase = out + (((n * a.To + d) * a.Ho + h_first) * (long long)a.wpitch) * CL;
  const bool planes_ok = (MODE == 21) ? (2 * d + 1 < a.T) : true;
  for (int i = threadIdx.x; i < nrows * a.wpitch; i += PRE_THREADS) {
    const int rh = i / a.wpitch, x = i - rh * a.wpitch;
    const int xo = x - a.wpad;                                  // real output position
    const int h = h_first + rh;
    uint32_t pk[CL / 2];
    // first source pixel of the position and whether all of its pixels lie inside the frame
    int p0;
    bool interior;
    if (MODE >= 20) {
      p0 = 2 * xo;
      interior = xo >= 0 && xo < a.Wo && p0 + 1 < a.W && 2 * h + 1 < a.H && planes_ok;
    } else {
      p0 = (MODE == 4 ? 2 * xo : xo) - 1;
      const int wlim = (MODE == 4) ? (a.W - a.w0) : a.Wo;      // pixels available in the (cropped) row
      interior = xo >= 0 && xo < a.Wo && p0 >= 0 && p0 + MODE <= wlim;
    }
    const uint8_t* seg0 = rows + (size_t)(rh * NPL * NR) * a.row_pitch + ((MODE >= 20) ? p0 : p0 + a.w0) * (C * ES);
    if (interior) {
#pragma unroll
      for (int k2 = 0; k2 < CL / 2; ++k2) {
        float f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          constexpr int dummy = 0; (void)dummy;
          const int k = 2 * k2 + e;
          float val = 0.f;
          if (k < KREAL) {
            const int sg = k / L, j = k - sg * L, c = j % C;      // compile-time after unrolling
            val = src_to_f32(reinterpret_cast<const TS*>(seg0 + (size_t)sg * a.row_pitch) + j);
            if (!IDENT) val = (val - a.mean[c]) * a.scale[c];
          }
          f[e] = val;
        }
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[0], f[1]);
        pk[k2] = *reinterpret_cast<uint32_t*>(&h2);
      }
    } else {
      // frame edges, pad positions of the row, missing rows / planes: per-element checks
      const bool xreal = xo >= 0 && xo < a.Wo;
#pragma unroll
      for (int k2 = 0; k2 < CL / 2; ++k2) {
        float f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * k2 + e;
          float val = 0.f;
          if (k < KREAL && xreal) {
            const int sg = k / L, j = k - sg * L, c = j % C, px = j / C;
            bool ok;
            if (MODE >= 20) {
              const int pd = (MODE == 21) ? (sg >> 1) : 0, ph = sg & 1;
              ok = (2 * h + ph < a.H) && (p0 + px < a.W) && ((MODE == 21) ? (2 * d + pd < a.T) : true);
            } else {
              const int wlim = (MODE == 4) ? (a.W - a.w0) : a.Wo;
              ok = p0 + px >= 0 && p0 + px < wlim;
            }
            if (ok) {
              val = src_to_f32(reinterpret_cast<const TS*>(seg0 + (size_t)sg * a.row_pitch) + j);
              if (!IDENT) val = (val - a.mean[c]) * a.scale[c];
            }
          }
          f[e] = val;
        }
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[0], f[1]);
        pk[k2] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(obase + (size_t)i * CL);
#pragma unroll
    for (int q = 0; q < Q; ++q) o[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

template <int MODE, int C, int CL, typename TS>
static int preprocess_rows_launch_t(const TS* src, void* out, int n, PreRowArgs a, cudaStream_t st) {
  a.row_bytes = a.W * C * (int)sizeof(TS);
  a.row_pitch = (a.row_bytes + 15) & ~15;
  a.aligned = (a.row_bytes % 16 == 0) && (((uintptr_t)src) % 16 == 0);
  constexpr int NPL = (MODE == 21) ? 2 : 1, NR = (MODE >= 20) ? 2 : 1;
  // rows per CTA: ~8 chunk stores per thread, <= 40 KB of staged rows
  const int per_row = a.wpitch;
  int rh = max(1, (4 * PRE_THREADS) / max(per_row, 1));
  rh = min(rh, max(1, (40 * 1024) / (NPL * NR * a.row_pitch)));
  rh = min(rh, a.Ho);
  a.rh = rh;
  a.hgroups = ceil_div(a.Ho, rh);
  const size_t smem = (size_t)rh * NPL * NR * a.row_pitch;
  CSE_REQUIRE(smem <= 48 * 1024, "preprocess: source row of %d bytes is too long to stage", a.row_bytes);
  const long long blocks = (long long)n * a.To * a.hgroups;
  CSE_REQUIRE(blocks < (1ll << 31), "preprocess: too many blocks");
  if (blocks == 0) return CSE_OK;
  bool ident = true;          // the reference's behaviour (train.py:466-478): raw 0..255, no mean / scale
  for (int c = 0; c < C; ++c) ident = ident && a.mean[c] == 0.f && a.scale[c] == 1.f;
  if (ident)
    preprocess_rows_kernel<MODE, C, CL, true, TS><<<(unsigned)blocks, PRE_THREADS, smem, st>>>(src, (__nv_bfloat16*)out, a);
  else
    preprocess_rows_kernel<MODE, C, CL, false, TS><<<(unsigned)blocks, PRE_THREADS, smem, st>>>(src, (__nv_bfloat16*)out, a);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

// src_f32: the clip holds float32 values (on-the-fly Farneback flow) instead of uint8 frames.  Only the 1- / 2-channel
// space-to-depth stems are instantiated for it (a flow volume has 2 channels).
template <int MODE, int C, int CL>
static int preprocess_rows_launch(const uint8_t* src, void* out, int n, const PreRowArgs& a, cudaStream_t st, bool src_f32) {
  if (!src_f32) return preprocess_rows_launch_t<MODE, C, CL, uint8_t>(src, out, n, a, st);
  if constexpr (MODE >= 20 && C <= 2) {
    return preprocess_rows_launch_t<MODE, C, CL, float>(reinterpret_cast<const float*>(src), out, n, a, st);
  } else {
    set_error("preprocess: float32 clips are supported for the 1- / 2-channel space-to-depth stems only");
    return CSE_ERR_INVALID;
  }
}

int launch_preprocess(const uint8_t* src, int n, int T, int H, int W, int C, int t0, int h0, int w0,
                      int To, int Ho, int Wo, const float* mean, const float* scale, void* out,
                      int out_dt, int out_ld, cudaStream_t st, int wpitch, int wpad, int unroll_w, int s2d, int src_dt) {
  CSE_REQUIRE(src_dt == CSE_U8 || src_dt == CSE_F32, "preprocess: source dtype %d must be u8 or f32", src_dt);
  const bool src_f32 = src_dt == CSE_F32;
  if (s2d) {
    // s2d = 1: To,Ho,Wo = T, ceil(H/2), ceil(W/2); s2d = 2: ceil(T/2), ceil(H/2), ceil(W/2) - the space-to-depth grid
    const int cells = s2d == 2 ? 8 : 4;
    CSE_REQUIRE((s2d == 1 || s2d == 2) && out_dt == CSE_BF16 && C >= 1 && C <= 4 && out_ld % 8 == 0 && cells * C <= out_ld &&
                    out_ld <= 32 && unroll_w == 0,
                "preprocess: s2d needs bf16 and out_ld in {8,16,24,32} >= %d*C (C=%d, out_ld=%d)", cells, C, out_ld);
    CSE_REQUIRE(t0 == 0 && h0 == 0 && w0 == 0 && To == (s2d == 2 ? (T + 1) / 2 : T) && Ho == (H + 1) / 2 && Wo == (W + 1) / 2,
                "preprocess: s2d output grid (%d,%d,%d) does not match clip (%d,%d,%d)", To, Ho, Wo, T, H, W);
    if (wpitch <= 0) { wpitch = Wo; wpad = 0; }
    CSE_REQUIRE(wpad >= 0 && wpitch >= Wo + wpad, "preprocess: row pitch %d < Wo %d + pad %d", wpitch, Wo, wpad);
    PreRowArgs a;
    a.T = T; a.H = H; a.W = W; a.t0 = 0; a.h0 = 0; a.w0 = 0;
    a.To = To; a.Ho = Ho; a.Wo = Wo; a.wpitch = wpitch; a.wpad = wpad;
    for (int c = 0; c < 4; ++c) {
      a.mean[c] = (mean && c < C) ? mean[c] : 0.f;
      a.scale[c] = (scale && c < C) ? scale[c] : 1.f;
    }
#define S2D_CASE(M_, C_, CL_) return preprocess_rows_launch<M_, C_, CL_>(src, out, n, a, st, src_f32)
    if (s2d == 2) {
      if (C == 1 && out_ld == 8) S2D_CASE(21, 1, 8);
      if (C == 2 && out_ld == 16) S2D_CASE(21, 2, 16);
      if (C == 3 && out_ld == 24) S2D_CASE(21, 3, 24);
      if (C == 4 && out_ld == 32) S2D_CASE(21, 4, 32);
    } else {
      if (C == 1 && out_ld == 8) S2D_CASE(20, 1, 8);
      if (C == 2 && out_ld == 8) S2D_CASE(20, 2, 8);
      if (C == 1 && out_ld == 16) S2D_CASE(20, 1, 16);
      if (C == 2 && out_ld == 16) S2D_CASE(20, 2, 16);
      if (C == 3 && out_ld == 16) S2D_CASE(20, 3, 16);
      if (C == 4 && out_ld == 16) S2D_CASE(20, 4, 16);
    }
#undef S2D_CASE
    set_error("preprocess: s2d=%d with C=%d, out_ld=%d is not instantiated", s2d, C, out_ld);
    return CSE_ERR_INVALID;
  }
  if (unroll_w > 0) {
    CSE_REQUIRE((unroll_w == 3 || unroll_w == 4) && out_dt == CSE_BF16 && out_ld == 16 && C >= 1 && C <= 4 && C * unroll_w <= 16 &&
                    wpitch <= 0,
                "preprocess: unroll_w supports 3 (pixel) / 4 (pixel pair), bf16, out_ld=16, C*unroll_w<=16");
    CSE_REQUIRE(t0 >= 0 && h0 >= 0 && w0 >= 0 && t0 + To <= T && h0 + Ho <= H &&
                    w0 + (unroll_w == 4 ? 2 * Wo - 1 : Wo) <= W,
                "preprocess: crop outside clip");
    PreRowArgs a;
    a.T = T; a.H = H; a.W = W; a.t0 = t0; a.h0 = h0; a.w0 = w0;
    a.To = To; a.Ho = Ho; a.Wo = Wo; a.wpitch = Wo; a.wpad = 0;
    for (int c = 0; c < 4; ++c) {
      a.mean[c] = (mean && c < C) ? mean[c] : 0.f;
      a.scale[c] = (scale && c < C) ? scale[c] : 1.f;
    }
#define UNR_CASE(NB_, C_) return preprocess_rows_launch<NB_, C_, 16>(src, out, n, a, st, src_f32)
    if (unroll_w == 4) {
      switch (C) { case 1: UNR_CASE(4, 1); case 2: UNR_CASE(4, 2); case 3: UNR_CASE(4, 3); default: UNR_CASE(4, 4); }
    } else {
      switch (C) { case 1: UNR_CASE(3, 1); case 2: UNR_CASE(3, 2); case 3: UNR_CASE(3, 3); default: UNR_CASE(3, 4); }
    }
#undef UNR_CASE
  }
  CSE_REQUIRE(C >= 1 && C <= 4, "preprocess: C=%d not in 1..4", C);
  CSE_REQUIRE(t0 >= 0 && h0 >= 0 && w0 >= 0 && t0 + To <= T && h0 + Ho <= H && w0 + Wo <= W,
              "preprocess: crop (%d,%d,%d)+(%d,%d,%d) outside clip (%d,%d,%d)", t0, h0, w0, To, Ho, Wo, T, H, W);
  CSE_REQUIRE(out_ld >= C && out_ld <= 8, "preprocess: out_ld=%d must be in [C,8]", out_ld);
  PreArgs a;
  a.T = T; a.H = H; a.W = W; a.C = C; a.t0 = t0; a.h0 = h0; a.w0 = w0;
  a.To = To; a.Ho = Ho; a.Wo = Wo; a.out_ld = out_ld;
  if (wpitch <= 0) { wpitch = Wo; wpad = 0; }
  CSE_REQUIRE(wpad >= 0 && wpitch >= Wo + wpad, "preprocess: row pitch %d < Wo %d + pad %d", wpitch, Wo, wpad);
  a.wpitch = wpitch; a.wpad = wpad;
  for (int c = 0; c < 4; ++c) {
    a.mean[c] = (mean && c < C) ? mean[c] : 0.f;
    a.scale[c] = (scale && c < C) ? scale[c] : 1.f;
  }
  long long total = (long long)n * To * Ho * wpitch;
  if (total == 0) return CSE_OK;
  unsigned blocks = (unsigned)((total + 255) / 256);
  using bf = __nv_bfloat16;
#define PRE_CASE(TT, CO)                                                                                          \
  do {                                                                                                            \
    if (src_f32) preprocess_kernel<TT, CO, float><<<blocks, 256, 0, st>>>((const float*)src, (TT*)out, total, a); \
    else preprocess_kernel<TT, CO, uint8_t><<<blocks, 256, 0, st>>>(src, (TT*)out, total, a);                     \
  } while (0)
  if (out_dt == CSE_BF16) {
    switch (out_ld) {
      case 8: PRE_CASE(bf, 8); break;
      case 4: PRE_CASE(bf, 4); break;
      case 3: PRE_CASE(bf, 3); break;
      case 2: PRE_CASE(bf, 2); break;
      default: set_error("preprocess: bf16 out_ld=%d unsupported (2,3,4,8)", out_ld); return CSE_ERR_INVALID;
    }
  } else if (out_dt == CSE_F32) {
    switch (out_ld) {
      case 4: PRE_CASE(float, 4); break;
      case 3: PRE_CASE(float, 3); break;
      case 2: PRE_CASE(float, 2); break;
      case 8: PRE_CASE(float, 8); break;
      default: set_error("preprocess: f32 out_ld=%d unsupported (2,3,4,8)", out_ld); return CSE_ERR_INVALID;
    }
  } else {
    set_error("preprocess: unsupported output dtype %d", out_dt);
    return CSE_ERR_INVALID;
  }
#undef PRE_CASE
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

}  // namespace cse
