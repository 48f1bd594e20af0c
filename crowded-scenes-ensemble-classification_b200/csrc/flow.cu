// Dense optical flow (Gunnar Farneback's polynomial expansion) on the GPU, for the FarneBack_onTheFly TwoStream variant.
//
// The reference computes the flow volume of a clip in its loader with OpenCV (opticalflow_FarneBack_extractor,
// train.py:294-332):  cv2.calcOpticalFlowFarneback(prev_gray, gray, None, pyr_scale=0.5, levels=5, winsize=11,
// iterations=5, poly_n=5, poly_sigma=1.1, flags=0)  between consecutive gray frames.  OpenCV (4.x, video/optflowgf.cpp,
// BSD) is a third-party dependency that is not part of the reference tree; its published algorithm is restated here
// kernel by kernel, in the same order of operations:
//   per pyramid level (coarse -> fine; level k works at scale pyr_scale^k, skipping levels smaller than 32 pixels):
//     flow   = resize(previous level's flow) / pyr_scale                     (zeros at the coarsest level)
//     I_i    = resize(GaussianBlur(float(img_i), sigma = (1/scale - 1)/2))   for both frames
//     R_i    = polynomial expansion of I_i (separable (2n+1) Gaussian-weighted fits -> 5 coefficients per pixel)
//     M      = UpdateMatrices(R_0, R_1 sampled at x + flow)                  (G11, G12, G22, h1, h2 per pixel)
//     repeat `iterations` times:  flow = solve(box-blur(M, winsize)),  M = UpdateMatrices(...) (not after the last)
// Arithmetic follows OpenCV's (float for images / coefficients, double for the horizontal polynomial sums and the box
// sums) with every multiply and add explicitly rounded (no FMA contraction), in the order oracle/farneback.py uses, so
// the kernels are BIT-IDENTICAL to that restatement; the restatement itself agrees with cv2 to <= 6e-6 pixel on flows
// of several pixels (cv2's IPP / AVX2 blur sums in another order, its box filter is a sliding window of float-rounded
// differences) - tests/test_oracle_farneback.py, tests/test_gpu_flow.py.  One call handles all the frames of a video:
// blur / resize / polynomial expansion once per FRAME and level, matrices and solves per consecutive PAIR (blockIdx.y).
// clips.ClipSequence(..., device=cuda) - the evaluation path's default - computes the flow here; CSE_CPU_FLOW=1 keeps the
// loader-side cv2 calls of the reference (bit-exact with its golden, tests/test_clips_farneback.py).
#include <math.h>

#include <vector>

#include "common.cuh"

namespace cse {

constexpr int FB_THREADS = 256;
constexpr int FB_MAX_N = 16;          // poly_n and blur radius limits (constant-size weight arrays in kernel arguments)

struct FbTaps { float g[2 * FB_MAX_N + 1]; int n; };

// Every kernel works on a batch of images: blockIdx.y is the image (frame or frame pair), blockIdx.x * blockDim.x +
// threadIdx.x the element inside it; the per-image strides are in elements of the respective type.
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv2.GaussianBlur on CV_32F, BORDER_REFLECT_101: separable (rows first), symmetric taps; TS = uint8_t fuses convertTo
template <typename TS>
__global__ void fb_blur_kernel(const TS* __restrict__ src_, float* __restrict__ dst_, int H, int W, FbTaps t, int vertical) {
  const long long px = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= px) return;
  const TS* src = src_ + blockIdx.y * px;
  const int x = (int)(i % W), y = (int)(i / W);
  float s = __fmul_rn((float)src[i], t.g[t.n]);
  for (int k = 1; k <= t.n; ++k) {
    float a, b;
    if (vertical) { a = (float)src[(long long)reflect101(y - k, H) * W + x]; b = (float)src[(long long)reflect101(y + k, H) * W + x]; }
    else { a = (float)src[(long long)y * W + reflect101(x - k, W)]; b = (float)src[(long long)y * W + reflect101(x + k, W)]; }
    s = __fadd_rn(s, __fmul_rn(__fadd_rn(a, b), t.g[t.n + k]));
  }
  dst_[blockIdx.y * px + i] = s;
}

// cv2.resize(..., INTER_LINEAR) on CV_32F with C channels: pixel-centre mapping, float weights, horizontal then vertical
// combination; columns clamp the fraction at the edges, rows keep it and clamp the row indices.  `mul` scales the
// result (flow *= 1 / pyr_scale)
__device__ __forceinline__ void fb_tap(int d, double scale, int src, bool horizontal, int& i0, int& i1, float& f) {
  f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (horizontal) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  i0 = clampi(s, 0, src - 1);
  i1 = clampi(s + 1, 0, src - 1);
}
__global__ void fb_resize_kernel(const float* __restrict__ src_, long long sstride, int Hs, int Ws, int C, float* __restrict__ dst_,
                                 long long dstride, int H, int W, double sy_, double sx_, float mul) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W * C) return;
  const float* src = src_ + blockIdx.y * sstride;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % W), y = (int)(i / ((long long)C * W));
  int x0, x1, y0, y1;
  float fx, fy;
  fb_tap(x, sx_, Ws, true, x0, x1, fx);
  fb_tap(y, sy_, Hs, false, y0, y1, fy);
  const float* r0 = src + ((long long)y0 * Ws) * C + c;
  const float* r1 = src + ((long long)y1 * Ws) * C + c;
  const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
  const float h0 = __fadd_rn(__fmul_rn(r0[(long long)x0 * C], a0), __fmul_rn(r0[(long long)x1 * C], a1));
  const float h1 = __fadd_rn(__fmul_rn(r1[(long long)x0 * C], a0), __fmul_rn(r1[(long long)x1 * C], a1));
  dst_[blockIdx.y * dstride + i] = __fmul_rn(__fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1)), mul);
}

struct FbPoly { float g[FB_MAX_N + 1], xg[FB_MAX_N + 1], xxg[FB_MAX_N + 1]; int n; double ig11, ig03, ig33, ig55; };

// FarnebackPolyExp, vertical part: row[x] = (sum g*p, sum xg*(p(y+k) - p(y-k)), sum xxg*p) over rows clamped to the frame
__global__ void fb_poly_v_kernel(const float* __restrict__ src_, long long sstride, float* __restrict__ row3_, long long rstride, int H,
                                 int W, FbPoly p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W) return;
  const float* src = src_ + blockIdx.y * sstride;
  float* row3 = row3_ + blockIdx.y * rstride;
  const int x = (int)(i % W), y = (int)(i / W);
  float t0 = __fmul_rn(src[i], p.g[0]), t1 = 0.f, t2 = 0.f;
  for (int k = 1; k <= p.n; ++k) {
    const float s0 = src[(long long)max(y - k, 0) * W + x], s1 = src[(long long)min(y + k, H - 1) * W + x];
    const float q = __fadd_rn(s0, s1);
    t0 = __fadd_rn(t0, __fmul_rn(p.g[k], q));
    t1 = __fadd_rn(t1, __fmul_rn(p.xg[k], __fsub_rn(s1, s0)));
    t2 = __fadd_rn(t2, __fmul_rn(p.xxg[k], q));
  }
  row3[i * 3] = t0; row3[i * 3 + 1] = t1; row3[i * 3 + 2] = t2;
}

// FarnebackPolyExp, horizontal part (double accumulators; the 1 and x^2 sums take double products, the others float
// products widened afterwards, as the C expression types give; columns replicated at the frame edge) -> R[y][x][5]
__global__ void fb_poly_h_kernel(const float* __restrict__ row3_, long long rstride, float* __restrict__ R_, long long Rstride, int H, int W,
                                 FbPoly p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W) return;
  const int x = (int)(i % W), y = (int)(i / W);
  const float* row = row3_ + blockIdx.y * rstride + (long long)y * W * 3;
  const float g0 = p.g[0];
  double b1 = (double)__fmul_rn(row[x * 3], g0), b2 = 0, b3 = (double)__fmul_rn(row[x * 3 + 1], g0), b4 = 0,
         b5 = (double)__fmul_rn(row[x * 3 + 2], g0), b6 = 0;
  for (int k = 1; k <= p.n; ++k) {
    const float* rp = row + clampi(x + k, 0, W - 1) * 3;
    const float* rm = row + clampi(x - k, 0, W - 1) * 3;
    const double tg = (double)__fadd_rn(rp[0], rm[0]);
    const float gk = p.g[k], xgk = p.xg[k];
    b1 = __dadd_rn(b1, __dmul_rn(tg, (double)gk));
    b4 = __dadd_rn(b4, __dmul_rn(tg, (double)p.xxg[k]));
    b2 = __dadd_rn(b2, (double)__fmul_rn(__fsub_rn(rp[0], rm[0]), xgk));
    b3 = __dadd_rn(b3, (double)__fmul_rn(__fadd_rn(rp[1], rm[1]), gk));
    b6 = __dadd_rn(b6, (double)__fmul_rn(__fsub_rn(rp[1], rm[1]), xgk));
    b5 = __dadd_rn(b5, (double)__fmul_rn(__fadd_rn(rp[2], rm[2]), gk));
  }
  float* d = R_ + blockIdx.y * Rstride + i * 5;
  d[1] = (float)__dmul_rn(b2, p.ig11);
  d[0] = (float)__dmul_rn(b3, p.ig11);
  d[3] = (float)__dadd_rn(__dmul_rn(b1, p.ig03), __dmul_rn(b4, p.ig33));
  d[2] = (float)__dadd_rn(__dmul_rn(b1, p.ig03), __dmul_rn(b5, p.ig33));
  d[4] = (float)__dmul_rn(b6, p.ig55);
}

// FarnebackUpdateMatrices for pair b = blockIdx.y: R0 = R[b], R1 = R[b + 1] sampled bilinearly at (x, y) + flow,
// combined into (G11, G12, G22, h1, h2)
__global__ void fb_update_matrices_kernel(const float* __restrict__ R_, long long Rstride, const float* __restrict__ flow_, long long fstride,
                                          float* __restrict__ M_, long long Mstride, int H, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W) return;
  const float* R0 = R_ + blockIdx.y * Rstride;
  const float* R1 = R0 + Rstride;
  const float* flow = flow_ + blockIdx.y * fstride;
  const int x = (int)(i % W), y = (int)(i / W);
  const float dx = flow[i * 2], dy = flow[i * 2 + 1];
  float fx = __fadd_rn((float)x, dx), fy = __fadd_rn((float)y, dy);
  const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
  fx = __fsub_rn(fx, (float)x1); fy = __fsub_rn(fy, (float)y1);
  const float* r0 = R0 + i * 5;
  float r2, r3, r4, r5, r6;
  if ((unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1)) {
    const float* p = R1 + ((long long)y1 * W + x1) * 5;
    const long long st = (long long)W * 5;
    const float ofx = __fsub_rn(1.f, fx), ofy = __fsub_rn(1.f, fy);
    const float a00 = __fmul_rn(ofx, ofy), a01 = __fmul_rn(fx, ofy), a10 = __fmul_rn(ofx, fy), a11 = __fmul_rn(fx, fy);
#define FB_BIL(c) __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a00, p[c]), __fmul_rn(a01, p[5 + c])), __fmul_rn(a10, p[st + c])), __fmul_rn(a11, p[st + 5 + c]))
    r2 = FB_BIL(0); r3 = FB_BIL(1); r4 = FB_BIL(2); r5 = FB_BIL(3); r6 = FB_BIL(4);
#undef FB_BIL
    r4 = __fmul_rn(__fadd_rn(r0[2], r4), 0.5f);
    r5 = __fmul_rn(__fadd_rn(r0[3], r5), 0.5f);
    r6 = __fmul_rn(__fadd_rn(r0[4], r6), 0.25f);
  } else {
    r2 = r3 = 0.f;
    r4 = r0[2]; r5 = r0[3]; r6 = __fmul_rn(r0[4], 0.5f);
  }
  r2 = __fmul_rn(__fsub_rn(r0[0], r2), 0.5f);
  r3 = __fmul_rn(__fsub_rn(r0[1], r3), 0.5f);
  r2 = __fadd_rn(r2, __fadd_rn(__fmul_rn(r4, dy), __fmul_rn(r6, dx)));
  r3 = __fadd_rn(r3, __fadd_rn(__fmul_rn(r6, dy), __fmul_rn(r5, dx)));
  const int BORDER = 5;
  if ((unsigned)(x - BORDER) >= (unsigned)(W - BORDER * 2) || (unsigned)(y - BORDER) >= (unsigned)(H - BORDER * 2)) {
    const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    float scale = (x < BORDER ? border[x] : 1.f);
    scale = __fmul_rn(scale, (x >= W - BORDER ? border[W - x - 1] : 1.f));
    scale = __fmul_rn(scale, (y < BORDER ? border[y] : 1.f));
    scale = __fmul_rn(scale, (y >= H - BORDER ? border[H - y - 1] : 1.f));
    r2 = __fmul_rn(r2, scale); r3 = __fmul_rn(r3, scale); r4 = __fmul_rn(r4, scale); r5 = __fmul_rn(r5, scale); r6 = __fmul_rn(r6, scale);
  }
  float* m = M_ + blockIdx.y * Mstride + i * 5;
  m[0] = __fadd_rn(__fmul_rn(r4, r4), __fmul_rn(r6, r6));
  m[1] = __fmul_rn(__fadd_rn(r4, r5), r6);
  m[2] = __fadd_rn(__fmul_rn(r5, r5), __fmul_rn(r6, r6));
  m[3] = __fadd_rn(__fmul_rn(r4, r2), __fmul_rn(r6, r3));
  m[4] = __fadd_rn(__fmul_rn(r6, r2), __fmul_rn(r5, r3));
}

// FarnebackUpdateFlow_Blur, vertical box sum (rows replicated at the frame edge), double, rows -m..m in that order
__global__ void fb_box_v_kernel(const float* __restrict__ M_, long long Mstride, double* __restrict__ vsum_, long long vstride, int H, int W, int m) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W * 5) return;
  const float* M = M_ + blockIdx.y * Mstride;
  const long long xc = i % ((long long)W * 5);
  const int y = (int)(i / ((long long)W * 5));
  double s = 0.0;
  for (int r = -m; r <= m; ++r) s = __dadd_rn(s, (double)M[(long long)clampi(y + r, 0, H - 1) * W * 5 + xc]);
  vsum_[blockIdx.y * vstride + i] = s;
}
// horizontal box sum (columns replicated), scaling, 2x2 solve with the 1e-3 regulariser -> flow
__global__ void fb_box_h_solve_kernel(const double* __restrict__ vsum_, long long vstride, float* __restrict__ flow_, long long fstride, int H,
                                      int W, int m, double scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W) return;
  const int x = (int)(i % W), y = (int)(i / W);
  const double* row = vsum_ + blockIdx.y * vstride + (long long)y * W * 5;
  double g11 = 0, g12 = 0, g22 = 0, h1 = 0, h2 = 0;
  for (int r = -m; r <= m; ++r) {
    const double* v = row + clampi(x + r, 0, W - 1) * 5;
    g11 = __dadd_rn(g11, v[0]); g12 = __dadd_rn(g12, v[1]); g22 = __dadd_rn(g22, v[2]); h1 = __dadd_rn(h1, v[3]); h2 = __dadd_rn(h2, v[4]);
  }
  g11 = __dmul_rn(g11, scale); g12 = __dmul_rn(g12, scale); g22 = __dmul_rn(g22, scale); h1 = __dmul_rn(h1, scale); h2 = __dmul_rn(h2, scale);
  const double idet = __ddiv_rn(1.0, __dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3));
  float* flow = flow_ + blockIdx.y * fstride + i * 2;
  flow[0] = (float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet);
  flow[1] = (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet);
}

// The two box sums and the solve in ONE pass over M for the reference's window (HALF = winsize / 2 = 5): a CTA owns a
// 32 x 16 tile of flow vectors.  One thread per (column, channel) of the tile plus halo loads its 26 rows of M straight
// into registers (coalesced: a row of the haloed tile is 210 consecutive floats; edge-replicated) and forms the 16 column
// sums with a register window, so every M value is fetched once instead of 11 times; the row sums are formed the same way
// per (row, channel, half row) from shared memory, then one thread per pixel scales and solves.  Each window is still
// summed directly in double in the order r = -HALF .. HALF, so the result is bit-identical to the two generic kernels
// above (which remain for other window sizes).  FP64-pipe work: 352 DADD per thread-column; the shared arrays are
// padded (213 / 165 doubles per row) so that the 64-bit accesses of a half-warp fall into distinct banks.
constexpr int FB_TW = 32, FB_TH = 16;
template <int HALF>
__global__ void __launch_bounds__(256)
fb_box_solve_tile_kernel(const float* __restrict__ M_, long long Mstride, float* __restrict__ flow_, long long fstride, int H, int W, double scale) {
  constexpr int WIN = 2 * HALF + 1, CW = (FB_TW + 2 * HALF) * 5, RH = FB_TH + 2 * HALF, VLD = CW + 3, SLD = FB_TW * 5 + 5;
  constexpr int SEG = FB_TW / 2;
  static_assert(CW <= 256 && FB_TH * 5 * 2 <= 256, "one thread per column-channel / per (row, channel, half row)");
  __shared__ double V[FB_TH * VLD];                      // column sums [FB_TH][(FB_TW + 2 HALF) * 5]
  __shared__ double S[FB_TH * SLD];                      // window sums [FB_TH][FB_TW * 5]
  const int tid = threadIdx.x;
  const int tiles_x = (W + FB_TW - 1) / FB_TW;
  const int x0 = (blockIdx.x % tiles_x) * FB_TW, y0 = (blockIdx.x / tiles_x) * FB_TH;
  if (tid < CW) {
    const int tx = tid / 5, c = tid - tx * 5;
    const float* src = M_ + blockIdx.y * Mstride + (long long)clampi(x0 + tx - HALF, 0, W - 1) * 5 + c;
    float in[RH];
#pragma unroll
    for (int ty = 0; ty < RH; ++ty) in[ty] = src[(long long)clampi(y0 + ty - HALF, 0, H - 1) * W * 5];
    double w[WIN];                                        // widened once on entry: conversions are a slow pipe
#pragma unroll
    for (int r = 0; r < WIN - 1; ++r) w[r + 1] = (double)in[r];
#pragma unroll
    for (int yo = 0; yo < FB_TH; ++yo) {
#pragma unroll
      for (int r = 0; r < WIN - 1; ++r) w[r] = w[r + 1];
      w[WIN - 1] = (double)in[yo + WIN - 1];
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < WIN; ++r) s = __dadd_rn(s, w[r]);
      V[yo * VLD + tid] = s;
    }
  }
  __syncthreads();
  if (tid < FB_TH * 5 * 2) {
    const int c = tid % 5, ty = (tid / 5) % FB_TH, seg = tid / (5 * FB_TH);
    const double* v = V + ty * VLD + seg * SEG * 5 + c;
    double* out = S + ty * SLD + seg * SEG * 5 + c;
    double w[WIN];
#pragma unroll
    for (int r = 0; r < WIN - 1; ++r) w[r + 1] = v[r * 5];
#pragma unroll
    for (int xo = 0; xo < SEG; ++xo) {
#pragma unroll
      for (int r = 0; r < WIN - 1; ++r) w[r] = w[r + 1];
      w[WIN - 1] = v[(xo + WIN - 1) * 5];
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < WIN; ++r) s = __dadd_rn(s, w[r]);
      out[xo * 5] = s;
    }
  }
  __syncthreads();
  float* flow = flow_ + blockIdx.y * fstride;
#pragma unroll
  for (int p = tid; p < FB_TH * FB_TW; p += 256) {
    const int ty = p / FB_TW, tx = p % FB_TW, gy = y0 + ty, gx = x0 + tx;
    if (gy >= H || gx >= W) continue;
    const double* v = S + ty * SLD + tx * 5;
    const double g11 = __dmul_rn(v[0], scale), g12 = __dmul_rn(v[1], scale), g22 = __dmul_rn(v[2], scale), h1 = __dmul_rn(v[3], scale),
                 h2 = __dmul_rn(v[4], scale);
    const double idet = __ddiv_rn(1.0, __dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3));
    float2 o;
    o.x = (float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet);
    o.y = (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet);
    *reinterpret_cast<float2*>(flow + ((long long)gy * W + gx) * 2) = o;
  }
}

// cv2.cvtColor(BGR2GRAY) on CV_8U: 15-bit fixed point (B 3735, G 19235, R 9798), round to nearest
__global__ void fb_bgr2gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = bgr + i * 3;
  gray[i] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15);
}

// ----------------------------------------------------------------------------- host side
static inline unsigned fb_blocks(long long n) { return (unsigned)((n + FB_THREADS - 1) / FB_THREADS); }
static inline int cv_round(double v) { return (int)nearbyint(v); }          // round half to even, like cvRound

static FbPoly fb_prepare_poly(int n, double sigma) {
  // FarnebackPrepareGaussian: normalised Gaussian g, x*g, x^2*g and the entries of the inverse moment matrix it needs
  FbPoly p;
  p.n = n;
  if (sigma < 1.1920929e-07) sigma = n * 0.3;
  std::vector<float> g(2 * n + 1);
  double s = 0.0;
  for (int x = -n; x <= n; ++x) { g[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
  s = 1.0 / s;
  std::vector<float> xg(2 * n + 1), xxg(2 * n + 1);
  for (int x = -n; x <= n; ++x) {
    g[x + n] = (float)(g[x + n] * s);
    xg[x + n] = (float)(x * g[x + n]);
    xxg[x + n] = (float)(x * x * g[x + n]);
  }
  double a = 0, b = 0, c = 0, d = 0;       // G(0,0), G(1,1), G(3,3), G(5,5)
  for (int y = -n; y <= n; ++y)
    for (int x = -n; x <= n; ++x) {
      const float w = g[y + n] * g[x + n], fx = (float)x, fy = (float)y;       // float products, as the C expression types give
      a += w; b += w * fx * fx; c += w * fx * fx * fx * fx; d += w * fx * fx * fy * fy;
    }
  // G = [[a,0,0,b,b,0],[0,b,0,0,0,0],[0,0,b,0,0,0],[b,0,0,c,d,0],[b,0,0,d,c,0],[0,0,0,0,0,d]]: inverse of the
  // {0,3,4} block [[a,b,b],[b,c,d],[b,d,c]] by cofactors
  const double det = a * (c * c - d * d) - b * (b * c - d * b) + b * (b * d - c * b);
  p.ig11 = 1.0 / b;
  p.ig55 = 1.0 / d;
  p.ig03 = -(b * c - b * d) / det;
  p.ig33 = (a * c - b * b) / det;
  for (int k = 0; k <= n; ++k) { p.g[k] = g[n + k]; p.xg[k] = xg[n + k]; p.xxg[k] = xxg[n + k]; }
  return p;
}

static FbTaps fb_gaussian_taps(int ksize, double sigma) {
  // cv::getGaussianKernel(ksize, sigma, CV_32F): fixed table for small kernels with sigma <= 0
  FbTaps t;
  t.n = ksize / 2;
  if (sigma <= 0 && ksize <= 9) {
    static const float tab[5][9] = {{1.f}, {0.25f, 0.5f, 0.25f}, {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
                                    {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f},
                                    {4 / 256.f, 13 / 256.f, 30 / 256.f, 51 / 256.f, 60 / 256.f, 51 / 256.f, 30 / 256.f, 13 / 256.f, 4 / 256.f}};
    for (int i = 0; i < ksize; ++i) t.g[i] = tab[ksize / 2][i];
    return t;
  }
  const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
  const double scale2x = -0.5 / (sx * sx);
  std::vector<double> w(ksize);
  double sum = 0.0;
  for (int i = 0; i < ksize; ++i) { const double x = i - (ksize - 1) * 0.5; w[i] = exp(scale2x * x * x); sum += w[i]; }
  for (int i = 0; i < ksize; ++i) t.g[i] = (float)(w[i] / sum);
  return t;
}

struct FbLevel { int w, h, smooth; double scale, sigma; };

static std::vector<FbLevel> fb_levels(int H, int W, double pyr_scale, int levels) {
  const int min_size = 32;
  int k = 0;
  double scale = 1;
  for (; k < levels; ++k) {
    scale *= pyr_scale;
    if (W * scale < min_size || H * scale < min_size) break;
  }
  std::vector<FbLevel> out;
  for (int lv = k; lv >= 0; --lv) {
    FbLevel L;
    L.scale = 1;
    for (int i = 0; i < lv; ++i) L.scale *= pyr_scale;
    L.sigma = (1.0 / L.scale - 1) * 0.5;
    L.smooth = cv_round(L.sigma * 5) | 1;
    if (L.smooth < 3) L.smooth = 3;
    L.w = cv_round(W * L.scale);
    L.h = cv_round(H * L.scale);
    out.push_back(L);
  }
  return out;
}

static size_t fb_workspace(int F, int H, int W) {
  const size_t px = (size_t)H * W, P = (size_t)F - 1;
  // per frame: two blur / level-image planes, the 3-channel vertical sums, the 5 polynomial coefficients; per pair: M (5),
  // two flow fields (2 each), the double column sums (5)
  return F * px * (2 + 3 + 5) * sizeof(float) + P * px * (5 + 4) * sizeof(float) + P * px * 5 * sizeof(double) + 256;
}

}  // namespace cse

using namespace cse;

extern "C" {

size_t cse_farneback_workspace_bytes(int n_frames, int H, int W) { return (n_frames >= 2 && H > 0 && W > 0) ? fb_workspace(n_frames, H, W) : 0; }

int cse_farneback(const uint8_t* d_gray, int n_frames, int H, int W, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                  double poly_sigma, float* d_flow, void* d_work, size_t work_bytes, void* stream) {
  CSE_REQUIRE(d_gray && d_flow && d_work, "farneback: NULL pointer");
  CSE_REQUIRE(n_frames >= 2 && n_frames <= 65535, "farneback: %d frames (2..65535: flow between consecutive frames)", n_frames);
  CSE_REQUIRE(H >= 2 && W >= 2 && pyr_scale > 0 && pyr_scale < 1 && levels >= 0 && iterations >= 1, "farneback: bad geometry / pyramid");
  CSE_REQUIRE(poly_n >= 1 && poly_n <= FB_MAX_N && winsize >= 1 && winsize <= 255, "farneback: poly_n 1..%d, winsize 1..255", FB_MAX_N);
  CSE_REQUIRE(work_bytes >= fb_workspace(n_frames, H, W) && ((uintptr_t)d_work % 8) == 0,
              "farneback: workspace of %zu bytes needed (8-byte aligned), %zu given", fb_workspace(n_frames, H, W), work_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const long long px = (long long)H * W;
  const unsigned F = (unsigned)n_frames, P = F - 1;
  double* vsum = reinterpret_cast<double*>(d_work);
  float* f = reinterpret_cast<float*>(vsum + (size_t)P * px * 5);
  float* planeA = f; f += (size_t)F * px;
  float* planeB = f; f += (size_t)F * px;
  float* row3 = f; f += (size_t)F * px * 3;
  float* R = f; f += (size_t)F * px * 5;
  float* M = f; f += (size_t)P * px * 5;
  float* flowbuf[2] = {f, f + (size_t)P * px * 2};
  const FbPoly poly = fb_prepare_poly(poly_n, poly_sigma);
  const std::vector<FbLevel> lv = fb_levels(H, W, pyr_scale, levels);
  float* prev_flow = nullptr;
  int pw = 0, ph = 0;
  for (size_t li = 0; li < lv.size(); ++li) {
    const FbLevel& L = lv[li];
    const bool last = li + 1 == lv.size();
    float* flow = last ? d_flow : flowbuf[li & 1];
    const long long lpx = (long long)L.w * L.h;
    if (!prev_flow) {
      CSE_CUDA(cudaMemsetAsync(flow, 0, sizeof(float) * 2 * lpx * P, st));
    } else {
      fb_resize_kernel<<<dim3(fb_blocks(lpx * 2), P), FB_THREADS, 0, st>>>(prev_flow, (long long)pw * ph * 2, ph, pw, 2, flow, lpx * 2, L.h, L.w,
                                                                          (double)ph / L.h, (double)pw / L.w, (float)(1.0 / pyr_scale));
    }
    CSE_REQUIRE(L.smooth / 2 <= FB_MAX_N, "farneback: smoothing kernel of %d taps is too wide", L.smooth);
    const FbTaps taps = fb_gaussian_taps(L.smooth, L.sigma);
    fb_blur_kernel<uint8_t><<<dim3(fb_blocks(px), F), FB_THREADS, 0, st>>>(d_gray, planeA, H, W, taps, 0);
    fb_blur_kernel<float><<<dim3(fb_blocks(px), F), FB_THREADS, 0, st>>>(planeA, planeB, H, W, taps, 1);
    fb_resize_kernel<<<dim3(fb_blocks(lpx), F), FB_THREADS, 0, st>>>(planeB, px, H, W, 1, planeA, lpx, L.h, L.w, (double)H / L.h, (double)W / L.w, 1.0f);
    fb_poly_v_kernel<<<dim3(fb_blocks(lpx), F), FB_THREADS, 0, st>>>(planeA, lpx, row3, lpx * 3, L.h, L.w, poly);
    fb_poly_h_kernel<<<dim3(fb_blocks(lpx), F), FB_THREADS, 0, st>>>(row3, lpx * 3, R, lpx * 5, L.h, L.w, poly);
    fb_update_matrices_kernel<<<dim3(fb_blocks(lpx), P), FB_THREADS, 0, st>>>(R, lpx * 5, flow, lpx * 2, M, lpx * 5, L.h, L.w);
    const int m = winsize / 2;
    const double scale = 1.0 / ((double)winsize * winsize);
    for (int it = 0; it < iterations; ++it) {
      if (m == 5) {
        const unsigned tiles = (unsigned)(((L.w + FB_TW - 1) / FB_TW) * ((L.h + FB_TH - 1) / FB_TH));
        fb_box_solve_tile_kernel<5><<<dim3(tiles, P), 256, 0, st>>>(M, lpx * 5, flow, lpx * 2, L.h, L.w, scale);
      } else {
        fb_box_v_kernel<<<dim3(fb_blocks(lpx * 5), P), FB_THREADS, 0, st>>>(M, lpx * 5, vsum, lpx * 5, L.h, L.w, m);
        fb_box_h_solve_kernel<<<dim3(fb_blocks(lpx), P), FB_THREADS, 0, st>>>(vsum, lpx * 5, flow, lpx * 2, L.h, L.w, m, scale);
      }
      if (it < iterations - 1)
        fb_update_matrices_kernel<<<dim3(fb_blocks(lpx), P), FB_THREADS, 0, st>>>(R, lpx * 5, flow, lpx * 2, M, lpx * 5, L.h, L.w);
    }
    prev_flow = flow; pw = L.w; ph = L.h;
  }
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

int cse_bgr2gray(const uint8_t* d_bgr, uint8_t* d_gray, long long pixels, void* stream) {
  CSE_REQUIRE(d_bgr && d_gray && pixels >= 0, "bgr2gray: bad argument");
  if (pixels == 0) return CSE_OK;
  fb_bgr2gray_kernel<<<fb_blocks(pixels), FB_THREADS, 0, (cudaStream_t)stream>>>(d_bgr, d_gray, pixels);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

int cse_resize_linear_f32(const float* d_src, int n, int Hs, int Ws, int C, float* d_dst, int H, int W, void* stream) {
  CSE_REQUIRE(d_src && d_dst && n >= 1 && n <= 65535 && Hs >= 1 && Ws >= 1 && H >= 1 && W >= 1 && C >= 1, "resize_linear_f32: bad argument");
  const long long e = (long long)H * W * C;
  fb_resize_kernel<<<dim3(fb_blocks(e), (unsigned)n), FB_THREADS, 0, (cudaStream_t)stream>>>(d_src, (long long)Hs * Ws * C, Hs, Ws, C, d_dst, e, H, W,
                                                                                             (double)Hs / H, (double)Ws / W, 1.0f);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

}  // extern "C"
