// Graph -> device plan, in C++ (host only): fusion, engine selection, weight packing, BN folding, buffer planning.
// The native twin of cse_b200/lowering.py in its default configuration; tests/test_model_api.py holds the two to
// byte-identical `cse_op` lists and weight arenas for every architecture, so every decision below is the one documented
// (and measured) there.  See lowering.py for the rationale of each rule.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>

#include "../../include/cse.h"
#include "model_graph.h"

namespace cse {
namespace mdl {

constexpr long long ALIGN = 1024;
constexpr int SM_COUNT = 148;          // B200: persistent grid = one CTA per SM
constexpr float BN_EPS = 1e-3f;        // Keras BatchNormalization default epsilon

inline long long round_up(long long x, long long a) { return (x + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

inline uint16_t bf16_bits(float f) {   // round to nearest even (what torch's .to(bfloat16) does)
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x0040u);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

struct Tensor {                        // a Keras weight tensor (fp32, Keras layout)
  std::vector<int> shape;
  std::vector<float> data;
  bool set = false;
};

struct Buf {
  std::string name;
  long long nbytes = 0;
  int first = 1 << 30, last = -1;
  long long offset = -1;
};

struct TRef {                          // [n,D,H,W,C] view: channel slice [coff, coff+C) of a buffer with ld channels per pixel
  int buf = -1;
  int coff = 0, C = 0, ld = 0;
  int dims[3] = {0, 0, 0};
  int dtype = CSE_BF16;
  int wpitch = 0, wpad = 0, unroll_w = 0, s2d = 0, src_c = 0;
  bool valid() const { return buf >= 0; }
};

struct DevOp {
  int kind = 0;
  std::string name;
  TRef in0, in1, out0, out1, out2;
  int engine = 0, w_dtype = 0;
  int k[3] = {1, 1, 1}, s[3] = {1, 1, 1}, pad[3] = {0, 0, 0};
  int relu0 = 0, relu1 = 0, pad_is_zero = 0, ext_input = 0;
  int src_dims[4] = {0, 0, 0, 0};
  int kc = 0, bn = 0, brick[4] = {0, 0, 0, 0}, halo = 0, pair_pool = 0, ksplit = 1, part = -1, src_dtype = CSE_U8;
  int out_split = 0, out_split2 = 0;
  int pool_k[3] = {0, 0, 0}, pool_zero = 0;
  bool has_conv_out_dims = false;
  int conv_out_dims[3] = {0, 0, 0};
  int w_blob = -1, scale0 = -1, shift0 = -1, scale1 = -1, shift1 = -1;
  int softmax_C = 0;
};

struct Val {                           // value of a lowered layer: one view, or the list a 1-D concat carries to Dense
  TRef ref;
  std::vector<TRef> list;
  bool is_list = false;
};

struct Plan {
  std::vector<cse_op> ops;
  std::vector<uint8_t> arena;
  long long workspace_bytes = 0;
  long long logits_off = -1, probs_off = -1;
  int n_inputs = 0;
};

// ----------------------------------------------------------------------------- tile heuristics (lowering.py)
inline int choose_kc(int ci) {
  int best = 64;
  long long best_cost = -1;
  for (int kc : {64, 32, 16}) {
    const long long cost = (long long)cdiv(ci, kc) * (24 + kc);
    if (best_cost < 0 || cost < best_cost) { best = kc; best_cost = cost; }
  }
  return best;
}

inline void choose_bn(int co, int m_tiles, int* bn_out, int* nt_out) {
  const int min_t = cdiv(co, 256);
  bool have = false;
  double best_time = 0;
  int best_nt = 0, best_bn = 0;
  const int hi = std::max(std::max(min_t + 1, 9), co / 64 + 1);
  for (int n_tiles = min_t; n_tiles < hi; ++n_tiles) {
    const int bn = (int)round_up(cdiv(co, n_tiles), 16);
    if (n_tiles > min_t && bn < 64) break;
    if (m_tiles <= 0) { *bn_out = bn; *nt_out = n_tiles; return; }
    const double mma = 4 * std::max(bn / 2.0, 48.0);
    const double l2 = (128 * 128 + bn * 128) / 85.0;
    const long long waves = ((long long)m_tiles * n_tiles + SM_COUNT - 1) / SM_COUNT;
    const double t = waves * std::max(mma, l2);
    if (!have || t < best_time || (t == best_time && n_tiles < best_nt)) { have = true; best_time = t; best_nt = n_tiles; best_bn = bn; }
  }
  *bn_out = best_bn; *nt_out = best_nt;
}

inline void choose_brick(int nb, int d_, int h_, int w_, const int mult[3], int out[4]) {
  const int md = mult[0], mh = mult[1], mw = mult[2];
  bool have = false;
  long long bk[4] = {0, 0, 0, 0};
  for (int bw = mw; bw <= std::min((int)round_up(w_, mw), 128); bw += mw) {
    const int tw = cdiv(w_, bw);
    for (int bh = mh; bh <= std::min((int)round_up(h_, mh), 128 / bw); bh += mh) {
      const int th = cdiv(h_, bh);
      for (int bd = md; bd <= std::min((int)round_up(d_, md), 128 / (bw * bh)); bd += md) {
        const int td = cdiv(d_, bd);
        const int bn_ = std::max(1, std::min(nb, 128 / (bw * bh * bd)));
        const int tn = cdiv(nb, bn_);
        const long long key[4] = {(long long)tn * td * th * tw, bn_, -bw, -bh};
        bool less = !have;
        if (have) {
          for (int i = 0; i < 4; ++i) {
            if (key[i] != bk[i]) { less = key[i] < bk[i]; break; }
          }
        }
        if (less) { have = true; std::copy(key, key + 4, bk); out[0] = bn_; out[1] = bd; out[2] = bh; out[3] = bw; }
      }
    }
  }
}

inline void choose_brick_hhalo(int ho, int wo, int kh, int out[4]) {
  bool have = false;
  long long bk[3] = {0, 0, 0};
  for (int bw = 8; bw <= std::min((int)round_up(wo, 8), 128); bw += 8)
    for (int bh = 1; bh <= 128 / bw; ++bh) {
      const long long tiles = (long long)cdiv(ho, bh) * cdiv(wo, bw);
      const long long key[3] = {tiles * (bh + kh - 1) * bw, tiles, -bw};
      bool less = !have;
      if (have) for (int i = 0; i < 3; ++i) if (key[i] != bk[i]) { less = key[i] < bk[i]; break; }
      if (less) { have = true; std::copy(key, key + 3, bk); out[0] = 1; out[1] = 1; out[2] = bh; out[3] = bw; }
    }
}

inline bool choose_brick_hw(int ho, int wo, const int mult[3], int out[4]) {
  if (mult[0] != 1) return false;
  int a = 8, b = mult[2];
  while (b) { int t = a % b; a = b; b = t; }
  const int step_w = 8 * mult[2] / a;
  bool have = false;
  long long bk[2] = {0, 0};
  for (int bw = step_w; bw <= std::min((int)round_up(wo, step_w), 128); bw += step_w)
    for (int bh = mult[1]; bh <= 128 / bw; bh += mult[1]) {
      const long long key[2] = {(long long)cdiv(ho, bh) * cdiv(wo, bw), -bw};
      bool less = !have;
      if (have) for (int i = 0; i < 2; ++i) if (key[i] != bk[i]) { less = key[i] < bk[i]; break; }
      if (less) { have = true; std::copy(key, key + 2, bk); out[0] = 1; out[1] = 1; out[2] = bh; out[3] = bw; }
    }
  return have;
}

inline double choose_brick_pair_halo(int ho, int wo, int kh, int out[4]) {
  bool have = false;
  long long bk[2] = {0, 0};
  for (int bw = 8; bw <= std::min((int)round_up(wo, 8), 128); bw += 8)
    for (int bh = 1; bh <= 128 / bw; ++bh) {
      const long long tiles = (long long)cdiv(ho, bh) * cdiv(wo, bw);
      const long long key[2] = {tiles, tiles * (bh + kh - 1) * bw};
      bool less = !have;
      if (have) for (int i = 0; i < 2; ++i) if (key[i] != bk[i]) { less = key[i] < bk[i]; break; }
      if (less) { have = true; std::copy(key, key + 2, bk); out[0] = 1; out[1] = 1; out[2] = bh; out[3] = bw; }
    }
  return (double)ho * wo / (double)(bk[0] * 128);
}

// ----------------------------------------------------------------------------- the conv kernel as a dense 5-D array
struct Kernel5 {                       // [kd][kh][kw][ci][co] fp32
  int kd = 0, kh = 0, kw = 0, ci = 0, co = 0;
  std::vector<float> v;
  Kernel5() {}
  Kernel5(int a, int b, int c, int d, int e) : kd(a), kh(b), kw(c), ci(d), co(e), v((size_t)a * b * c * d * e, 0.f) {}
  float& at(int a, int b, int c, int d, int e) { return v[((((size_t)a * kh + b) * kw + c) * ci + d) * co + e]; }
  float at(int a, int b, int c, int d, int e) const { return v[((((size_t)a * kh + b) * kw + c) * ci + d) * co + e]; }
};

using Blob = std::vector<uint8_t>;
inline Blob blob_u16(const std::vector<uint16_t>& a) { Blob b(a.size() * 2); std::memcpy(b.data(), a.data(), b.size()); return b; }
inline Blob blob_f32(const std::vector<float>& a) { Blob b(a.size() * 4); std::memcpy(b.data(), a.data(), b.size()); return b; }

// Keras [kd,kh,kw,Ci,Co] -> [Co_pad][taps*kchunks*kc] bf16 (K-major rows)
inline Blob pack_tc_weights(const Kernel5& k, int kc, int bn, int n_tiles) {
  const int taps = k.kd * k.kh * k.kw, cip = (int)round_up(k.ci, kc), rows = n_tiles * bn;
  std::vector<uint16_t> w((size_t)rows * taps * cip, 0);
  for (int t = 0; t < taps; ++t)
    for (int c = 0; c < k.ci; ++c)
      for (int o = 0; o < k.co; ++o)
        w[((size_t)o * taps + t) * cip + c] = bf16_bits(k.v[((size_t)t * k.ci + c) * k.co + o]);
  return blob_u16(w);
}
// Keras [kd,kh,1,Ci,Co] -> [n_tile][tap][bn][kc]
inline Blob pack_tc_weights_halo(const Kernel5& k, int kc, int bn, int n_tiles) {
  const int taps = k.kd * k.kh;
  std::vector<uint16_t> w((size_t)n_tiles * taps * bn * kc, 0);
  for (int t = 0; t < taps; ++t)
    for (int c = 0; c < k.ci; ++c)
      for (int o = 0; o < k.co; ++o) {
        const int nt = o / bn, r = o % bn;
        w[(((size_t)nt * taps + t) * bn + r) * kc + c] = bf16_bits(k.v[((size_t)t * k.ci + c) * k.co + o]);
      }
  return blob_u16(w);
}
// Keras [kd,kh,kw,Ci,Co] -> [n_tile][fd][fw][chunk][fh][bn][kc]
inline Blob pack_tc_weights_hhalo(const Kernel5& k, int kc, int bn, int n_tiles) {
  const int nch = cdiv(k.ci, kc);
  std::vector<uint16_t> w((size_t)n_tiles * k.kd * k.kw * nch * k.kh * bn * kc, 0);
  for (int fd = 0; fd < k.kd; ++fd)
    for (int fh = 0; fh < k.kh; ++fh)
      for (int fw = 0; fw < k.kw; ++fw)
        for (int c = 0; c < k.ci; ++c)
          for (int o = 0; o < k.co; ++o) {
            const int nt = o / bn, r = o % bn, ch = c / kc, cc = c % kc;
            const size_t idx = ((((((size_t)nt * k.kd + fd) * k.kw + fw) * nch + ch) * k.kh + fh) * bn + r) * kc + cc;
            w[idx] = bf16_bits(k.at(fd, fh, fw, c, o));
          }
  return blob_u16(w);
}

// (conv bias, BN tensors) -> fp32 (scale, shift) with y = acc*scale + shift, folded like TF's non-fused BN
inline void fold_bn(const std::vector<float>* bias, const std::vector<Tensor>* bn, bool has_gamma, std::vector<float>* scale,
                    std::vector<float>* shift, bool* has_scale, bool* has_shift) {
  *has_scale = false; *has_shift = false;
  if (!bn) {
    if (bias) { *shift = *bias; *has_shift = true; }
    return;
  }
  const int o = has_gamma ? 1 : 0;
  const std::vector<float>* gamma = has_gamma ? &(*bn)[0].data : nullptr;
  const std::vector<float>& beta = (*bn)[o].data; const std::vector<float>& mean = (*bn)[o + 1].data; const std::vector<float>& var = (*bn)[o + 2].data;
  const size_t n = beta.size();
  scale->resize(n); shift->resize(n);
  for (size_t i = 0; i < n; ++i) {
    volatile float inv = 1.0f / std::sqrt(var[i] + BN_EPS);
    if (gamma) { volatile float t = inv * (*gamma)[i]; inv = t; }
    volatile float mi = mean[i] * inv;
    volatile float sh = beta[i] - mi;
    if (bias) { volatile float bi = (*bias)[i] * inv; volatile float t2 = bi + sh; sh = t2; }
    (*scale)[i] = inv; (*shift)[i] = sh;
  }
  *has_scale = true; *has_shift = true;
}

}  // namespace mdl
}  // namespace cse
