// Clip assembly on the GPU: frame selection + bilinear resize of decoded uint8 frames, bit-exact
// with what the reference does on the CPU before a clip reaches the network:
//   select_frames              train.py:132-145   frames[::step][:T], step = max(1, n // T)
//   cv2.resize(f, (W, H))      train.py:286, 209-214   OpenCV INTER_LINEAR on CV_8U
// OpenCV's 8-bit bilinear path (imgproc/resize.cpp) is fixed point with 11 coefficient bits:
//   fx = float((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx
//   horizontal: sx < 0 -> (0, fx = 0); sx >= Ws-1 -> (Ws-1, fx = 0); a1 = cvRound(fx * 2048), a0 = cvRound((1-fx) * 2048)
//   vertical:   fy is kept, the two row indices are clipped to [0, Hs-1]
//   r = a0 * S[sx] + a1 * S[sx+1]  (int32);  out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
// The double arithmetic uses explicitly rounded mul / add (no FMA contraction) so the taps equal
// the host library's.  HBM-bound: reads the touched source pixels, writes T*H*W*C bytes.
#include "common.cuh"

namespace cse {

__device__ __forceinline__ void linear_tap(int d, double scale, int src, bool horizontal, int& i0, int& i1, int& w0, int& w1) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (horizontal) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  i0 = min(max(s, 0), src - 1);
  i1 = min(max(s + 1, 0), src - 1);
}

template <int C>
__global__ void __launch_bounds__(128)
assemble_clip_kernel(const uint8_t* __restrict__ frames, int frame_step, int Hs, int Ws, uint8_t* __restrict__ clip,
                     int H, int W, double scale_x, double scale_y) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y, t = blockIdx.z;
  if (dx >= W) return;
  int x0, x1, a0, a1, y0, y1, b0, b1;
  linear_tap(dx, scale_x, Ws, true, x0, x1, a0, a1);
  linear_tap(dy, scale_y, Hs, false, y0, y1, b0, b1);
  const uint8_t* src = frames + (size_t)t * frame_step * Hs * Ws * C;
  const uint8_t* r0 = src + (size_t)y0 * Ws * C;
  const uint8_t* r1 = src + (size_t)y1 * Ws * C;
  uint8_t* dst = clip + (((size_t)t * H + dy) * W + dx) * C;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int top = a0 * (int)__ldg(r0 + x0 * C + c) + a1 * (int)__ldg(r0 + x1 * C + c);
    const int bot = a0 * (int)__ldg(r1 + x0 * C + c) + a1 * (int)__ldg(r1 + x1 * C + c);
    const int v = (((b0 * (top >> 4)) >> 16) + ((b1 * (bot >> 4)) >> 16) + 2) >> 2;
    dst[c] = (uint8_t)v;
  }
}

}  // namespace cse

using namespace cse;

extern "C" int cse_assemble_clip(const uint8_t* d_frames, int n_frames, int Hs, int Ws, int C, uint8_t* d_clip, int T,
                                 int H, int W, void* stream) {
  CSE_REQUIRE(d_frames && d_clip && n_frames >= 1 && Hs >= 1 && Ws >= 1 && T >= 1 && H >= 1 && W >= 1,
              "assemble_clip: bad arguments (frames %d %dx%d -> %dx%dx%d)", n_frames, Hs, Ws, T, H, W);
  CSE_REQUIRE(C >= 1 && C <= 4, "assemble_clip: C=%d channels (1..4 supported)", C);
  int step = n_frames / T;
  if (step == 0) step = 1;
  const int kept = (n_frames + step - 1) / step;          // len(frames[::step])
  CSE_REQUIRE(kept >= T, "assemble_clip: the video has %d frames, select_frames keeps %d < T=%d", n_frames, kept, T);
  CSE_REQUIRE(T <= 65535 && H <= 65535, "assemble_clip: T=%d / H=%d exceed the grid limits", T, H);
  const double scale_x = 1.0 / ((double)W / (double)Ws), scale_y = 1.0 / ((double)H / (double)Hs);
  const dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)T);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 1: assemble_clip_kernel<1><<<grid, 128, 0, st>>>(d_frames, step, Hs, Ws, d_clip, H, W, scale_x, scale_y); break;
    case 2: assemble_clip_kernel<2><<<grid, 128, 0, st>>>(d_frames, step, Hs, Ws, d_clip, H, W, scale_x, scale_y); break;
    case 3: assemble_clip_kernel<3><<<grid, 128, 0, st>>>(d_frames, step, Hs, Ws, d_clip, H, W, scale_x, scale_y); break;
    default: assemble_clip_kernel<4><<<grid, 128, 0, st>>>(d_frames, step, Hs, Ws, d_clip, H, W, scale_x, scale_y); break;
  }
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

// cv2.resize(img, (W, H)) (fx = fy = 0) or cv2.resize(img, None, fx=fx, fy=fy) (W, H = cvRound(size * factor), taps from
// 1 / factor - OpenCV uses the factor as given) for n uint8 images [n,Hs,Ws,C] -> [n,H,W,C]; the Farneback extractor's
// frame scaling (train.py:300-312)
extern "C" int cse_resize_u8(const uint8_t* d_src, int n, int Hs, int Ws, int C, uint8_t* d_dst, int H, int W, double fx, double fy,
                             void* stream) {
  CSE_REQUIRE(d_src && d_dst && n >= 1 && n <= 65535 && Hs >= 1 && Ws >= 1 && H >= 1 && H <= 65535 && W >= 1, "resize_u8: bad arguments");
  CSE_REQUIRE(C >= 1 && C <= 4, "resize_u8: C=%d channels (1..4 supported)", C);
  CSE_REQUIRE((fx > 0) == (fy > 0), "resize_u8: give both factors or neither");
  if (fx > 0)
    CSE_REQUIRE(W == (int)nearbyint(Ws * fx) && H == (int)nearbyint(Hs * fy), "resize_u8: %dx%d is not cvRound(%dx%d * (%g, %g))", H, W, Hs, Ws,
                fy, fx);
  const double scale_x = 1.0 / (fx > 0 ? fx : (double)W / (double)Ws), scale_y = 1.0 / (fy > 0 ? fy : (double)H / (double)Hs);
  const dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 1: assemble_clip_kernel<1><<<grid, 128, 0, st>>>(d_src, 1, Hs, Ws, d_dst, H, W, scale_x, scale_y); break;
    case 2: assemble_clip_kernel<2><<<grid, 128, 0, st>>>(d_src, 1, Hs, Ws, d_dst, H, W, scale_x, scale_y); break;
    case 3: assemble_clip_kernel<3><<<grid, 128, 0, st>>>(d_src, 1, Hs, Ws, d_dst, H, W, scale_x, scale_y); break;
    default: assemble_clip_kernel<4><<<grid, 128, 0, st>>>(d_src, 1, Hs, Ws, d_dst, H, W, scale_x, scale_y); break;
  }
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}
