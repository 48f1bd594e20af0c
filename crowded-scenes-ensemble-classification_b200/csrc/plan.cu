// C ABI (include/cse.h): error state, device info, the member forward plan and the
// stand-alone entry points.  A plan is the device-side replacement of
// evaluate_load_model + predict_generator (train.py:1712-1772, evaluate_ensemble.py:1053-1056):
// a flat list of fused ops over one workspace arena and one weight arena, all launches issued
// from C++ on the caller's stream (capturable into a CUDA graph by the caller).
#include <stdarg.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace cse {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return CSE_ERR_CUDA;
}

int conv_tc_tune(const char* key, int value);
int vote_launch(const void* probs, int is_f64, const double* weights, int mode, int M, int N, int C,
                int32_t* pred, double* summed, cudaStream_t st);
int vote_search_launch(const double* probs, const double* weights, const int32_t* labels, int W, int M, int N,
                       int C, int32_t* correct, cudaStream_t st);

struct PlanOp {
  cse_op op;
  ConvTcDesc tc;
  bool has_tc = false;
};

}  // namespace cse

struct cse_plan {
  int max_batch = 0;
  int nb_classes = 0;
  std::vector<cse::PlanOp> ops;
  char* ws = nullptr;
  size_t ws_bytes = 0;
  const char* wts = nullptr;
  size_t wt_bytes = 0;
  int64_t logits_off = -1, probs_off = -1;
  bool finalized = false;
  int sm_count = 0;
  int last_launches = 0;
};

using namespace cse;

static WinGeom geom_of(const cse_op& o) {
  WinGeom g;
  g.Di = o.in_dims[0]; g.Hi = o.in_dims[1]; g.Wi = o.in_dims[2]; g.Ci = o.in_dims[3]; g.in_ld = o.in_ld;
  g.Do = o.out_dims[0]; g.Ho = o.out_dims[1]; g.Wo = o.out_dims[2]; g.Co = o.out_dims[3]; g.out_ld = o.out_ld;
  g.kd = o.k[0]; g.kh = o.k[1]; g.kw = o.k[2];
  g.sd = o.s[0]; g.sh = o.s[1]; g.sw = o.s[2];
  g.pd = o.pad[0]; g.ph = o.pad[1]; g.pw = o.pad[2];
  g.in_wpitch = o.in_wpitch;
  return g;
}

static int check_span(const cse_plan* p, int64_t off, long long elems, int dt, const char* what, bool weights) {
  if (off < 0) return CSE_OK;
  size_t lim = weights ? p->wt_bytes : p->ws_bytes;
  long long end = off + elems * (long long)dtype_size(dt);
  if ((size_t)end > lim) {
    set_error("%s span [%lld, %lld) exceeds %s arena of %zu bytes", what, (long long)off, end,
              weights ? "weight" : "workspace", lim);
    return CSE_ERR_INVALID;
  }
  return CSE_OK;
}

static long long tensor_span(int n, const int32_t dims[4], int ld, int wpitch = 0) {
  // elements from the first to one past the last addressed element of a [n,D,H,W,(C of ld)] slice
  long long pix = (long long)n * dims[0] * dims[1] * (wpitch > 0 ? wpitch : dims[2]);
  if (pix == 0) return 0;
  if (wpitch > 0) return pix * ld;           // padded rows: the whole pitch belongs to the tensor
  return (pix - 1) * ld + dims[3];
}

extern "C" {

int cse_abi_version(void) { return CSE_ABI_VERSION; }
const char* cse_last_error(void) { return g_err.c_str(); }

int cse_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  CSE_CUDA(cudaGetDevice(&dev));
  int sm = 0, mj = 0, mn = 0;
  CSE_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  CSE_CUDA(cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev));
  CSE_CUDA(cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sm;
  if (cc_major) *cc_major = mj;
  if (cc_minor) *cc_minor = mn;
  return CSE_OK;
}

int cse_malloc(void** d_ptr, size_t bytes) {
  CSE_REQUIRE(d_ptr != nullptr, "malloc: NULL argument");
  CSE_CUDA(cudaMalloc(d_ptr, bytes));
  return CSE_OK;
}
int cse_free(void* d_ptr) {
  CSE_CUDA(cudaFree(d_ptr));
  return CSE_OK;
}
int cse_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
  CSE_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return CSE_OK;
}
int cse_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
  CSE_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return CSE_OK;
}
int cse_stream_synchronize(void* stream) {
  CSE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return CSE_OK;
}

int cse_tune(const char* key, int value) {
  CSE_REQUIRE(key != nullptr, "tune: NULL key");
  int rc = conv_tc_tune(key, value);
  if (rc) set_error("tune: unknown key '%s'", key);
  return rc;
}

int cse_plan_create(cse_plan** out, int max_batch, int nb_classes) {
  CSE_REQUIRE(out != nullptr, "plan_create: out is NULL");
  CSE_REQUIRE(max_batch >= 1 && nb_classes >= 1, "plan_create: max_batch=%d nb_classes=%d", max_batch, nb_classes);
  cse_plan* p = new (std::nothrow) cse_plan();
  if (!p) { set_error("plan_create: out of host memory"); return CSE_ERR_NOMEM; }
  p->max_batch = max_batch;
  p->nb_classes = nb_classes;
  *out = p;
  return CSE_OK;
}

int cse_plan_add_op(cse_plan* p, const cse_op* op) {
  CSE_REQUIRE(p && op, "plan_add_op: NULL argument");
  if (p->finalized) { set_error("plan_add_op: plan already finalized"); return CSE_ERR_STATE; }
  CSE_REQUIRE(op->kind >= CSE_OP_PREPROCESS && op->kind <= CSE_OP_SOFTMAX, "plan_add_op: unknown kind %d", op->kind);
  if (op->kind == CSE_OP_PREPROCESS)
    CSE_REQUIRE(op->in_dtype == CSE_U8 || op->in_dtype == CSE_F32,
                "plan_add_op: PREPROCESS in_dtype must be CSE_U8 (decoded frames) or CSE_F32 (float flow), got %d", op->in_dtype);
  if (op->kind == CSE_OP_CONV3D)
    CSE_REQUIRE(op->engine == CSE_ENGINE_DIRECT || op->engine == CSE_ENGINE_TCGEN05,
                "plan_add_op: conv engine must be DIRECT or TCGEN05 (got %d)", op->engine);
  PlanOp po;
  po.op = *op;
  p->ops.push_back(po);
  return CSE_OK;
}

int cse_plan_finalize(cse_plan* p, void* d_workspace, size_t workspace_bytes, const void* d_weights,
                      size_t weight_bytes, int64_t logits_off, int64_t probs_off) {
  CSE_REQUIRE(p != nullptr, "plan_finalize: NULL plan");
  if (p->finalized) { set_error("plan_finalize: called twice"); return CSE_ERR_STATE; }
  CSE_REQUIRE(d_workspace && ((uintptr_t)d_workspace % 1024) == 0, "plan_finalize: workspace must be 1024-byte aligned");
  CSE_REQUIRE(!d_weights || ((uintptr_t)d_weights % 256) == 0, "plan_finalize: weights must be 256-byte aligned");
  int rc = cse_device_info(&p->sm_count, nullptr, nullptr);
  if (rc) return rc;
  p->ws = (char*)d_workspace; p->ws_bytes = workspace_bytes;
  p->wts = (const char*)d_weights; p->wt_bytes = weight_bytes;
  p->logits_off = logits_off; p->probs_off = probs_off;
  const int nb = p->max_batch;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    PlanOp& po = p->ops[i];
    const cse_op& o = po.op;
    // bounds of every tensor the op touches
    if (o.kind != CSE_OP_PREPROCESS && o.kind != CSE_OP_SOFTMAX) {
      if ((rc = check_span(p, o.in0_off, tensor_span(nb, o.in_dims, o.in_ld, o.in_wpitch), o.in_dtype, "in0", false))) return rc;
    }
    if (o.kind == CSE_OP_SOFTMAX) {
      long long e = (long long)nb * o.in_dims[3];
      if ((rc = check_span(p, o.in0_off, e, CSE_F32, "softmax in", false))) return rc;
      if ((rc = check_span(p, o.out0_off, e, CSE_F32, "softmax out", false))) return rc;
      continue;
    }
    if (o.kind == CSE_OP_CONV3D && o.pool_k[0] > 0) {
      const int32_t pd[4] = {o.pool_dims[0], o.pool_dims[1], o.pool_dims[2],
                             o.tc_pair_pool ? o.out_dims[3] / 2 : o.out_dims[3]};
      if ((rc = check_span(p, o.out0_off, tensor_span(nb, pd, o.out_ld), o.out_dtype, "pooled out0", false))) return rc;
      CSE_REQUIRE(o.engine == CSE_ENGINE_TCGEN05, "op %zu: fused pooling needs the tcgen05 engine", i);
    } else if (o.kind == CSE_OP_CONV3D && o.out_split > 0) {
      // fused sibling convs: three column ranges, three tensors
      CSE_REQUIRE(o.engine == CSE_ENGINE_TCGEN05 && o.out_split < o.out_dims[3] && o.out1_off >= 0 &&
                      (o.out_split2 == 0 || (o.out_split2 > o.out_split && o.out_split2 < o.out_dims[3] && o.out2_off >= 0)),
                  "op %zu: split output needs the tcgen05 engine and out1 (/ out2) tensors", i);
      const int end1 = o.out_split2 > 0 ? o.out_split2 : o.out_dims[3];
      const int32_t d0[4] = {o.out_dims[0], o.out_dims[1], o.out_dims[2], o.out_split};
      const int32_t d1[4] = {o.out_dims[0], o.out_dims[1], o.out_dims[2], end1 - o.out_split};
      const int32_t d2[4] = {o.out_dims[0], o.out_dims[1], o.out_dims[2], o.out_dims[3] - end1};
      if ((rc = check_span(p, o.out0_off, tensor_span(nb, d0, o.out_ld), o.out_dtype, "out0 (split)", false))) return rc;
      if ((rc = check_span(p, o.out1_off, tensor_span(nb, d1, o.out1_ld), o.out_dtype, "out1 (split)", false))) return rc;
      if (o.out_split2 > 0 &&
          (rc = check_span(p, o.out2_off, tensor_span(nb, d2, o.out2_ld), o.out_dtype, "out2 (split)", false))) return rc;
    } else {
      if ((rc = check_span(p, o.out0_off, tensor_span(nb, o.out_dims, o.out_ld, o.kind == CSE_OP_PREPROCESS ? o.out_wpitch : 0),
                           o.out_dtype, "out0", false))) return rc;
      if (o.out1_off >= 0 &&
          (rc = check_span(p, o.out1_off, tensor_span(nb, o.out_dims, o.out1_ld), o.out_dtype, "out1", false))) return rc;
    }
    if (o.in1_off >= 0) {
      const int32_t* d = (o.kind == CSE_OP_CONV3D) ? o.out_dims : o.in_dims;
      int dt = (o.kind == CSE_OP_CONV3D) ? o.out_dtype : o.in_dtype;
      if ((rc = check_span(p, o.in1_off, tensor_span(nb, d, o.in1_ld), dt, "in1", false))) return rc;
    }
    const int co = (o.kind == CSE_OP_CONV3D && o.tc_pair_pool) ? o.out_dims[3] / 2 : o.out_dims[3];
    if (o.kind == CSE_OP_CONV3D && o.tc_pair_pool)
      CSE_REQUIRE(o.scale0_off < 0 && o.in1_off < 0 && o.out1_off < 0 && o.pool_k[0] == 1 && o.pool_k[1] == 2 && o.pool_k[2] == 1,
                  "op %zu: pair-pool stem takes bias (+ReLU) only and pool (1,2,1) in the pair view", i);
    if ((rc = check_span(p, o.scale0_off, co, CSE_F32, "scale0", true))) return rc;
    if ((rc = check_span(p, o.shift0_off, co, CSE_F32, "shift0", true))) return rc;
    if ((rc = check_span(p, o.scale1_off, co, CSE_F32, "scale1", true))) return rc;
    if ((rc = check_span(p, o.shift1_off, co, CSE_F32, "shift1", true))) return rc;
    if (o.kind == CSE_OP_CONV3D) {
      CSE_REQUIRE(o.w_off >= 0, "op %zu: conv without weights", i);
      WinGeom g = geom_of(o);
      if (o.engine == CSE_ENGINE_TCGEN05) {
        CSE_REQUIRE(o.in_dtype == CSE_BF16 && o.out_dtype == CSE_BF16 && o.w_dtype == CSE_BF16,
                    "op %zu: tcgen05 engine needs bf16 in/out/weights", i);
        const int kchunks = ceil_div(g.Ci, o.kc > 0 ? o.kc : 64);
        const long long ktot = (long long)g.kd * g.kh * g.kw * kchunks * o.kc;   // same element count in every packing
        const long long rows = (long long)ceil_div(g.Co, o.bn > 0 ? o.bn : 16) * o.bn;
        if ((rc = check_span(p, o.w_off, ktot * rows, CSE_BF16, "tc weights", true))) return rc;
        if (o.ksplit > 1) {
          CSE_REQUIRE(o.part_off >= 0 && o.part_bytes > 0, "op %zu: split-K without a partial buffer", i);
          if ((rc = check_span(p, o.part_off, o.part_bytes, CSE_U8, "split-K partials", false))) return rc;
        }
        rc = conv_tc_build(&po.tc, p->ws + o.in0_off, p->wts + o.w_off, p->ws + o.out0_off,
                           o.out1_off >= 0 ? p->ws + o.out1_off : nullptr, o.out1_ld, nb, g, o.kc, o.bn, o.brick,
                           o.tc_halo, o.pool_k, o.pool_dims, o.pool_zero, o.tc_pair_pool, o.out_split, o.out_split2,
                           o.out_split2 > 0 ? p->ws + o.out2_off : nullptr, o.out2_ld,
                           o.ksplit > 1 ? o.ksplit : 1, (o.ksplit > 1 && o.part_off >= 0) ? p->ws + o.part_off : nullptr,
                           o.ksplit > 1 ? (size_t)o.part_bytes : 0);
        if (rc) return rc;
        po.has_tc = true;
      } else {
        const long long kk = (long long)g.kd * g.kh * g.kw * g.Ci * g.Co;
        if ((rc = check_span(p, o.w_off, kk, o.w_dtype, "direct weights", true))) return rc;
      }
    }
  }
  p->finalized = true;
  return CSE_OK;
}

static int run_op(cse_plan* p, PlanOp& po, const uint8_t* rgb, const uint8_t* flow, int n, cudaStream_t st) {
  const cse_op& o = po.op;
  char* ws = p->ws;
  const char* wt = p->wts;
  auto wsp = [&](int64_t off) -> void* { return off >= 0 ? (void*)(ws + off) : nullptr; };
  auto wtf = [&](int64_t off) -> const float* { return off >= 0 ? (const float*)(wt + off) : nullptr; };
  switch (o.kind) {
    case CSE_OP_PREPROCESS: {
      const uint8_t* src = o.ext_input == 0 ? rgb : flow;
      CSE_REQUIRE(src != nullptr, "plan_run: external input %d is NULL", o.ext_input);
      return launch_preprocess(src, n, o.src_dims[0], o.src_dims[1], o.src_dims[2], o.src_dims[3], o.crop[0],
                               o.crop[1], o.crop[2], o.out_dims[0], o.out_dims[1], o.out_dims[2], o.pre_mean,
                               o.pre_scale, wsp(o.out0_off), o.out_dtype, o.out_ld, st, o.out_wpitch, o.out_wpad, o.pre_unroll_w,
                               o.pre_s2d, o.in_dtype);
    }
    case CSE_OP_CONV3D: {
      Epilogue ep;
      ep.scale0 = wtf(o.scale0_off); ep.shift0 = wtf(o.shift0_off);
      ep.scale1 = wtf(o.scale1_off); ep.shift1 = wtf(o.shift1_off);
      ep.res = wsp(o.in1_off); ep.res_ld = o.in1_ld;
      ep.out0 = wsp(o.out0_off); ep.out1 = o.out_split > 0 ? nullptr : wsp(o.out1_off); ep.out1_ld = o.out1_ld;
      ep.relu0 = o.relu0; ep.relu1 = o.relu1;
      if (po.has_tc) return launch_conv_tc(po.tc, n, ep, p->sm_count, st);
      return launch_conv_direct(o.in_dtype, o.w_dtype, o.out_dtype, wsp(o.in0_off), wt + o.w_off, n, geom_of(o), ep, st);
    }
    case CSE_OP_MAXPOOL3D:
      return launch_pool(o.in_dtype, true, o.pad_is_zero != 0, wsp(o.in0_off), wsp(o.out0_off), n, geom_of(o), st);
    case CSE_OP_AVGPOOL3D:
      return launch_pool(o.in_dtype, false, false, wsp(o.in0_off), wsp(o.out0_off), n, geom_of(o), st);
    case CSE_OP_AFFINE: {
      long long pix = (long long)n * o.in_dims[0] * o.in_dims[1] * o.in_dims[2];
      return launch_affine(o.in_dtype, wsp(o.in0_off), o.in_ld, wsp(o.out0_off), o.out_ld, pix, o.in_dims[3],
                           wtf(o.scale0_off), wtf(o.shift0_off), o.relu0, st);
    }
    case CSE_OP_ADD: {
      long long pix = (long long)n * o.in_dims[0] * o.in_dims[1] * o.in_dims[2];
      return launch_add(o.in_dtype, wsp(o.in0_off), o.in_ld, wsp(o.in1_off), o.in1_ld, wsp(o.out0_off), o.out_ld, pix,
                        o.in_dims[3], st);
    }
    case CSE_OP_SOFTMAX:
      return launch_softmax((const float*)wsp(o.in0_off), (float*)wsp(o.out0_off), n, o.in_dims[3], st);
  }
  set_error("plan_run: unknown op kind %d", o.kind);
  return CSE_ERR_INVALID;
}

int cse_plan_run_range(cse_plan* p, const uint8_t* d_rgb_u8, const uint8_t* d_flow_u8, int n, int first, int last,
                       void* stream) {
  CSE_REQUIRE(p != nullptr, "plan_run: NULL plan");
  if (!p->finalized) { set_error("plan_run: plan not finalized"); return CSE_ERR_STATE; }
  CSE_REQUIRE(n >= 0 && n <= p->max_batch, "plan_run: n=%d outside [0, %d]", n, p->max_batch);
  CSE_REQUIRE(first >= 0 && last <= (int)p->ops.size() && first <= last, "plan_run: op range [%d,%d)", first, last);
  cudaStream_t st = (cudaStream_t)stream;
  p->last_launches = 0;
  if (n == 0) return CSE_OK;
  for (int i = first; i < last; ++i) {
    int rc = run_op(p, p->ops[i], d_rgb_u8, d_flow_u8, n, st);
    if (rc) return rc;
    p->last_launches++;
  }
  return CSE_OK;
}

int cse_plan_run(cse_plan* p, const uint8_t* d_rgb_u8, const uint8_t* d_flow_u8, int n, float* d_logits,
                 float* d_probs, void* stream) {
  return cse_plan_run_from(p, d_rgb_u8, d_flow_u8, n, 0, d_logits, d_probs, stream);
}

int cse_plan_num_input_ops(const cse_plan* p) {
  int k = 0;
  if (p) while (k < (int)p->ops.size() && p->ops[k].op.kind == CSE_OP_PREPROCESS) ++k;
  return k;
}

int cse_plan_run_from(cse_plan* p, const uint8_t* d_rgb_u8, const uint8_t* d_flow_u8, int n, int first_op,
                      float* d_logits, float* d_probs, void* stream) {
  CSE_REQUIRE(p != nullptr, "plan_run: NULL plan");
  int rc = cse_plan_run_range(p, d_rgb_u8, d_flow_u8, n, first_op, (int)p->ops.size(), stream);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)n * p->nb_classes * sizeof(float);
  if (d_logits && bytes) {
    CSE_REQUIRE(p->logits_off >= 0, "plan_run: plan has no logits buffer");
    CSE_CUDA(cudaMemcpyAsync(d_logits, p->ws + p->logits_off, bytes, cudaMemcpyDeviceToDevice, st));
  }
  if (d_probs && bytes) {
    CSE_REQUIRE(p->probs_off >= 0, "plan_run: plan has no probabilities buffer");
    CSE_CUDA(cudaMemcpyAsync(d_probs, p->ws + p->probs_off, bytes, cudaMemcpyDeviceToDevice, st));
  }
  return CSE_OK;
}

int cse_plan_num_ops(const cse_plan* p) { return p ? (int)p->ops.size() : 0; }
int cse_plan_last_launches(const cse_plan* p) { return p ? p->last_launches : 0; }
void cse_plan_destroy(cse_plan* p) { delete p; }

int cse_preprocess(const uint8_t* d_clips, int n, int T, int H, int W, int C, int t0, int h0, int w0, int To, int Ho,
                   int Wo, const float* mean, const float* scale, void* d_out, int out_dtype, int out_ld, void* stream) {
  CSE_REQUIRE(d_clips && d_out, "preprocess: NULL pointer");
  return launch_preprocess(d_clips, n, T, H, W, C, t0, h0, w0, To, Ho, Wo, mean, scale, d_out, out_dtype, out_ld,
                           (cudaStream_t)stream);
}

int cse_vote(const void* d_probs, int probs_dtype_is_f64, const double* d_weights, int mode, int M, int N, int C,
             int32_t* d_pred, double* d_summed, void* stream) {
  CSE_REQUIRE(d_probs && d_pred, "vote: NULL pointer");
  return vote_launch(d_probs, probs_dtype_is_f64, d_weights, mode, M, N, C, d_pred, d_summed, (cudaStream_t)stream);
}

int cse_vote_search(const double* d_probs, const double* d_weights, const int32_t* d_labels, int W, int M, int N, int C,
                    int32_t* d_correct, void* stream) {
  CSE_REQUIRE(d_probs && d_weights && d_labels && d_correct, "vote_search: NULL pointer");
  return vote_search_launch(d_probs, d_weights, d_labels, W, M, N, C, d_correct, (cudaStream_t)stream);
}

}  // extern "C"
