/* cse.h - C ABI of libcse_b200.so: the B200-native replacement of the arithmetic that
 * MounirB/Crowded-scenes-Ensemble-classification reaches through Keras 2.2.4 / TensorFlow 1.15
 * on its ensemble-inference hot path.
 *
 * The reference has no FFI of its own (100 % Python); the boundary it crosses is
 *   model.predict_generator(...)            evaluate_ensemble.py:1053-1056, 1099-1102
 *   evaluate_load_model(...)                train.py:1712-1772   (graph build + load_weights)
 *   DataGenerator.__data_generation         train.py:466-478     (uint8 frames -> float32 batch)
 *   ensemble_predictions(...)               evaluate_ensemble.py:343-370  (np.tensordot + argmax)
 * Each entry point below cites the reference code it replaces.  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions: plain C types only; every pointer named d_* is a DEVICE pointer on the current
 * CUDA device; `stream` is a cudaStream_t passed as void*; every function returns 0 on success
 * and a negative cse_status otherwise, with a human-readable message in cse_last_error()
 * (thread-local).  Handles are not thread-safe; distinct handles are independent.  Nothing
 * here falls back to the CPU: without a CUDA device every compute entry point fails.
 */
#ifndef CSE_B200_H
#define CSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSE_ABI_VERSION 2

typedef enum {
  CSE_OK = 0,
  CSE_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  CSE_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  CSE_ERR_STATE = -3,       /* call order (e.g. run before finalize) */
  CSE_ERR_NOMEM = -4
} cse_status;

typedef enum { CSE_F32 = 0, CSE_BF16 = 1, CSE_U8 = 2 } cse_dtype;

/* conv engines */
typedef enum {
  CSE_ENGINE_AUTO = 0,
  CSE_ENGINE_DIRECT = 1,    /* CUDA-core implicit GEMM, fp32 accumulate (any shape; the FP32 parity path) */
  CSE_ENGINE_TCGEN05 = 2    /* tcgen05.mma + TMEM + 5-D TMA im2col staging, bf16 in / fp32 accumulate */
} cse_engine;

typedef enum {
  CSE_OP_PREPROCESS = 1,    /* uint8 NDHWC -> float: replaces train.py:466-478 (+ optional crop/mean/scale) */
  CSE_OP_CONV3D = 2,        /* Conv3D (+bias / folded BN, ReLU, residual add, 2nd BN-ReLU output, channel-offset write) */
  CSE_OP_MAXPOOL3D = 3,     /* MaxPooling3D ('same' = -inf padding; ZeroPadding3D folded in as 0-valued padding) */
  CSE_OP_AVGPOOL3D = 4,     /* AveragePooling3D 'valid' */
  CSE_OP_AFFINE = 5,        /* stand-alone BatchNormalization (+ReLU): y = relu?(x*scale+shift) */
  CSE_OP_ADD = 6,           /* keras.layers.add */
  CSE_OP_SOFTMAX = 7        /* softmax over the last axis (fp32 in, fp32 out) */
} cse_op_kind;

/* One fused device op of a member's forward plan.  Tensors are NDHWC; a tensor is addressed by a
 * byte offset into the plan's workspace (activations) or weight arena, plus a channel leading
 * dimension `*_ld` (elements per pixel of the underlying buffer) so an op can read or write a
 * channel slice of a wider buffer (Inception concat written in place, train.py:1048-1193).
 * Dense layers are CONV3D ops on [n,1,1,1,K] (Flatten is a view: row-major D,H,W,C). */
typedef struct cse_op {
  int32_t kind;              /* cse_op_kind */
  int32_t engine;            /* cse_engine (CONV3D only) */
  int32_t in_dtype;          /* cse_dtype of in0 / in1; PREPROCESS: dtype of the external clip - CSE_U8 (decoded frames, the
                                default of every reference path) or CSE_F32 (dense flow of the FarneBack_onTheFly TwoStream
                                variant, train.py:294-332, values are not integers) */
  int32_t out_dtype;         /* cse_dtype of out0 / out1 */
  int32_t w_dtype;           /* cse_dtype of the packed kernel */
  int32_t in_dims[4];        /* D,H,W,C of in0 per clip */
  int32_t out_dims[4];       /* D,H,W,C of out0 per clip */
  int32_t in_ld, in1_ld, out_ld, out1_ld;
  int32_t k[3], s[3], pad[3];/* window, stride, padding BEFORE (after-padding is implied by out_dims) */
  int32_t relu0, relu1;      /* apply ReLU to out0 / out1 */
  int32_t pad_is_zero;       /* MAXPOOL: 1 = padded taps count as 0 (ZeroPadding3D), 0 = ignored (-inf) */
  int32_t ext_input;         /* PREPROCESS: index of the external uint8 input (0 = rgb, 1 = flow) */
  int32_t crop[3];           /* PREPROCESS: t0,h0,w0 of the crop inside the source clip */
  int32_t src_dims[4];       /* PREPROCESS: T,H,W,C of the source uint8 clip */
  int32_t kc;                /* TCGEN05: channels per K-chunk (16/32/64) */
  int32_t bn;                /* TCGEN05: N tile (multiple of 16, <= 256) */
  int32_t brick[4];          /* TCGEN05: output-pixel brick (n,d,h,w) per 128-row M tile */
  float   pre_mean[4];       /* PREPROCESS: per-channel mean  (reference behaviour: 0) */
  float   pre_scale[4];      /* PREPROCESS: per-channel scale (reference behaviour: 1); out = (x-mean)*scale */
  int32_t in_wpitch;         /* row pitch of in0 in pixels (0 = W): reading a W-padded tensor (TCGEN05 packed stem) */
  int32_t out_wpitch;        /* PREPROCESS: output row pitch in pixels (0 = W); pad columns are written as zeros */
  int32_t out_wpad;          /* PREPROCESS: zero columns on the left of every output row */
  int32_t pre_unroll_w;      /* PREPROCESS: 3 = every output pixel carries its neighbours w-1,w,w+1 (zero outside
                                the row), C channels each, packed j*C+c and zero-padded to out_ld = 16 (packed stem);
                                4 = every output element is a pixel PAIR p carrying pixels 2p-1 .. 2p+2 (out W = W/2) */
  int32_t pool_k[3];         /* TCGEN05: fused MaxPooling3D window (= stride, 'valid'); 0 = none.  out0 is then the
                                pooled tensor [n, pool_dims, Cout]; out_dims stay the conv's own output dims */
  int32_t pool_dims[3];      /* D,H,W of the pooled output */
  int32_t pool_zero;         /* 1 = positions beyond the conv output count as 0 (ZeroPadding3D in front of the pool) */
  int32_t tc_halo;           /* TCGEN05: 1 = (kd,kh)-halo'd A brick, weights packed [n_tile][tap][bn][kc] (packed stem);
                                2 = kh-halo'd A brick, one stage per (fd, chunk), weights [n_tile][fd][chunk][fh][bn][kc];
                                3 = the same with kw taps (one stage per (fd, fw, chunk), box shifted by fw), weights
                                [fd][fw][chunk][fh][bn][kc], single N tile <= 128: CTA-pair (cta_group::2) kernel only */
  int32_t tc_pair_pool;      /* TCGEN05 pair-packed stem (C3D conv1 + pool1, train.py:1230-1233): in0 = pair-unrolled clip
                                [n,T,H,W/2,16], GEMM row = 2 neighbouring output pixels, out_dims = D,H,W/2,2*Cout, weights
                                [2*Cout][taps*16]; MaxPooling3D (1,2,2) is done in registers, pool_k = (1,2,1) in this
                                view, out0 = pooled [n, pool_dims, Cout]; bias (+ReLU) epilogue only */
  int32_t out_split, out_split2, out2_ld;
                             /* TCGEN05: out_split > 0 = horizontally fused sibling 1x1x1 convs of an Inception block
                                (train.py:1048-1193), one GEMM whose output columns go to up to three tensors: columns
                                [0, out_split) -> out0 (branch 0, a channel slice of the concat buffer); columns
                                [out_split, out_split2 or Cout) -> the tensor at out1_off / out1_ld (branch 1a), starting at
                                its channel 0; columns [out_split2, Cout) -> the tensor at out2_off / out2_ld (branch 2a)
                                when out_split2 > 0.  One epilogue (scale0 / shift0 / relu0) for all columns; out1 is
                                then NOT a second epilogue output.  With tc_halo = 3 and out_split = Cout / 2 = 64 the same
                                fields carry the 7x7x7 stems of TWO ensemble members that read the same clips (train.py:1026,
                                1481): columns [0, 64) -> this member's tensor, [64, 128) -> the peer member's */
  int32_t pre_s2d;           /* PREPROCESS: 1 = 2x2 space-to-depth over (H,W) for the stride-2 7x7x7 stems: out_dims =
                                T, ceil(H/2), ceil(W/2); cell channel (ph*2+pw)*C + c, zero-padded to out_ld = 8/16 */
  int64_t in0_off, in1_off;  /* workspace byte offsets (-1 = none); in1 = residual for CONV3D, 2nd addend for ADD */
  int64_t out0_off, out1_off, out2_off;
  int64_t w_off;             /* weight-arena byte offsets (-1 = none) */
  int64_t scale0_off, shift0_off;   /* fp32 [Cout]: y = acc*scale0 + shift0 (+ in1); out0 = relu0?(y) */
  int64_t scale1_off, shift1_off;   /* fp32 [Cout]: out1 = relu1?(y*scale1 + shift1) */
  int64_t part_off;          /* TCGEN05 split-K: workspace byte offset of the fp32 partial tiles (-1 = none) */
  int64_t part_bytes;        /* size of that buffer */
  int32_t ksplit;            /* TCGEN05: split the K loop of every tile over `ksplit` CTAs (layers with fewer tiles than SMs:
                                small batches, the 1-2 row Dense layers); 0 / 1 = no split.  The splits are summed in order
                                by a second kernel, so results do not depend on scheduling */
  int32_t reserved0;
} cse_op;

typedef struct cse_plan cse_plan;   /* opaque: one ensemble member's forward pass on one GPU */

/* ---- library ------------------------------------------------------------------------------ */
int         cse_abi_version(void);
const char* cse_last_error(void);
/* sm count / compute capability of the current device; fails without a CUDA device */
int         cse_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Tuning / test hook: overrides a launch heuristic for the whole process.  Keys: "twin_min_tiles", "bshare_min_tiles",
 * "pair_min_tiles" (tile counts from which the twin-tile / shared-B / CTA-pair conv modes are used; 0 = never,
 * -1 = built-in default).  Unknown key -> CSE_ERR_INVALID. */
int         cse_tune(const char* key, int value);

/* ---- device memory helpers (for hosts that do not link the CUDA runtime themselves: plain C, ctypes, JNI ...) --- */
int  cse_malloc(void** d_ptr, size_t bytes);                      /* cudaMalloc */
int  cse_free(void* d_ptr);
int  cse_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);    /* asynchronous on `stream` */
int  cse_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
int  cse_stream_synchronize(void* stream);                        /* NULL = the default stream */

/* ---- member plan: replaces evaluate_load_model + predict_generator ------------------------ */
/* (train.py:1712-1772, evaluate_ensemble.py:1053-1056) */
int  cse_plan_create(cse_plan** out, int max_batch, int nb_classes);
int  cse_plan_add_op(cse_plan* p, const cse_op* op);
/* Binds device memory (caller-owned; workspace 1024-byte, weights 256-byte aligned) and builds TMA descriptors.  No
 * allocation happens after this call. */
int  cse_plan_finalize(cse_plan* p, void* d_workspace, size_t workspace_bytes,
                       const void* d_weights, size_t weight_bytes,
                       int64_t logits_off, int64_t probs_off);
/* Runs the member on n <= max_batch clips.  d_rgb_u8 / d_flow_u8: uint8 NDHWC clips (flow may be
 * NULL for single-stream members; it points to float32 values when the flow PREPROCESS op was declared with
 * in_dtype = CSE_F32).  d_logits / d_probs: fp32 [n, nb_classes] (either may be NULL). */
int  cse_plan_run(cse_plan* p, const uint8_t* d_rgb_u8, const uint8_t* d_flow_u8, int n,
                  float* d_logits, float* d_probs, void* stream);
/* Same, starting at op `first_op`.  Members of one fold ensemble share the architecture, the
 * workspace and therefore the pre-processed clip tensor(s): the first member runs the whole plan,
 * the others start after the cse_plan_num_input_ops() leading PREPROCESS ops (the plans must have
 * been lowered with persistent input buffers). */
int  cse_plan_run_from(cse_plan* p, const uint8_t* d_rgb_u8, const uint8_t* d_flow_u8, int n, int first_op,
                       float* d_logits, float* d_probs, void* stream);
int  cse_plan_num_input_ops(const cse_plan* p);
/* Runs ops [first, last) only (per-layer parity tests, profiling). */
int  cse_plan_run_range(cse_plan* p, const uint8_t* d_rgb_u8, const uint8_t* d_flow_u8, int n,
                        int first, int last, void* stream);
int  cse_plan_num_ops(const cse_plan* p);
/* number of kernels the last cse_plan_run launched */
int  cse_plan_last_launches(const cse_plan* p);
void cse_plan_destroy(cse_plan* p);

/* ---- member model: graph construction and lowering inside the library --------------------------- */
/* The same replacement of evaluate_load_model + predict_generator (train.py:1712-1772, evaluate_ensemble.py:1053-1056)
 * one level up: the caller names the architecture and hands over the Keras weight tensors; fusion, engine / tile choice,
 * BatchNormalization folding, weight re-packing and buffer planning happen in the library (csrc/model*.{h,cu}), so a
 * binding in any language needs none of it.
 *   model_type : "C3D" | "I3D" | "TWOSTREAM_I3D" | "R3D_18" | "R3D_34" | "R3D_50" | "R3D_101" | "R3D_152"  (train.py:1712-1772)
 *   T, H, W    : clip geometry (define_input, train.py:1566-1616: 16x112x112 for C3D / R3D, 20x224x224 for I3D / TwoStream)
 *   dtype      : CSE_BF16 = tcgen05 tensor-core path, CSE_F32 = CUDA-core reference-precision path
 * Layers are addressed the way model.load_weights(path) (by_name = False) pairs them: index into the weighted layers in
 * Keras' model.layers order, tensors within a layer in Keras order (Conv3D / Dense: kernel, bias; BatchNormalization:
 * [gamma,] beta, moving_mean, moving_variance), host fp32 in Keras layout ([kd,kh,kw,Cin,Cout], [in,out]). */
typedef struct cse_model cse_model;
int  cse_model_create(cse_model** out, const char* model_type, int T, int H, int W, int nb_classes, int dtype, int max_batch);
/* options before lowering: "persist_input" = 1 keeps the pre-processed clip alive after a forward pass (members of one
 * fold share it, cse_model_forward_shared_input); "flow_input_f32" = 1: the flow clip is float32 (FarneBack_onTheFly) */
int  cse_model_set_option(cse_model* m, const char* key, int value);
int  cse_model_num_layers(const cse_model* m);
int  cse_model_layer_info(const cse_model* m, int layer, char* name, int name_cap, int* n_tensors);
int  cse_model_tensor_info(const cse_model* m, int layer, int tensor, int64_t* dims /*[5]*/, int* ndim, char* name, int name_cap);
int  cse_model_set_weight(cse_model* m, int layer, int tensor, const float* host, const int64_t* dims, int ndim);
/* Stem fusion across two members of one ensemble (I3D / TwoStream-I3D / R3D: 64-filter 7x7x7 stride-2 stems, train.py:1026,
 * 999-1009, 1481).  Members of a fold read the same clips, so `lead` runs the stems of BOTH as one N = 128 GEMM and stores the
 * follower's activations into a persistent buffer that `follow` reads instead of running its own stem (bit-identical results).
 * Call after both members' weights are set and before either is lowered; it turns "persist_input" on for both.  Finalize both
 * on ONE shared workspace of max(workspace_bytes) bytes, and per batch run cse_model_forward(lead, ...) first, then
 * cse_model_forward_shared_input(follow, ...). */
int  cse_model_pair_stems(cse_model* lead, cse_model* follow);
/* Host-only lowering (no device needed): after it the plan can be inspected. */
int  cse_model_lower(cse_model* m);
int  cse_model_num_ops(const cse_model* m);
int  cse_model_get_op(const cse_model* m, int i, cse_op* out);
size_t cse_model_workspace_bytes(const cse_model* m);
size_t cse_model_weight_bytes(const cse_model* m);
int  cse_model_copy_weight_arena(const cse_model* m, void* host_dst, size_t cap);
int64_t cse_model_logits_offset(const cse_model* m);
int64_t cse_model_probs_offset(const cse_model* m);
/* Lowers if needed, uploads the packed weights, allocates the workspace (or binds d_shared_workspace - members that run
 * back to back on one stream may share one activation arena of max(cse_model_workspace_bytes) bytes, 1024-byte aligned) and
 * builds the plan.  No allocation happens after this call. */
int  cse_model_finalize(cse_model* m, void* d_shared_workspace, size_t shared_workspace_bytes);
/* n <= max_batch clips: d_rgb uint8 [n,T,H,W,3]; d_flow uint8 (or float32) [n,T,H,W,2] for the two-stream model, else NULL.
 * d_logits / d_probs: fp32 [n, nb_classes], either may be NULL.  Asynchronous on `stream`. */
int  cse_model_forward(cse_model* m, const void* d_rgb, const void* d_flow, int n, float* d_logits, float* d_probs, void* stream);
/* Same, skipping the pre-processing ops: the clip tensor written by the previous member of the fold (same architecture,
 * same shared workspace, "persist_input") is reused. */
int  cse_model_forward_shared_input(cse_model* m, const void* d_rgb, const void* d_flow, int n, float* d_logits, float* d_probs,
                                    void* stream);
void cse_model_destroy(cse_model* m);

/* ---- stand-alone kernels ------------------------------------------------------------------- */
/* Clip pre-processing, replaces the uint8 -> float32 store of train.py:466-478.  Reference
 * behaviour = crop none, mean 0, scale 1 (raw BGR 0..255).  Output channels c >= C are zero. */
int  cse_preprocess(const uint8_t* d_clips, int n, int T, int H, int W, int C,
                    int t0, int h0, int w0, int To, int Ho, int Wo,
                    const float* mean /*host [C] or NULL*/, const float* scale /*host [C] or NULL*/,
                    void* d_out, int out_dtype, int out_ld, void* stream);

/* Soft vote, replaces ensemble_predictions (evaluate_ensemble.py:343-370).
 * probs: [M,N,C] fp32 or fp64 (the reference votes on fp64 values parsed from its CSV);
 * weights: fp64 [M] (NULL = ones = "SUM"); mode 0 = weighted sum -> argmax (first max wins),
 * mode 1 = "MAXIMUM" (argmax over the member-major [M*C] row, modulo C).
 * d_pred int32 [N]; d_summed fp64 [N,C] or NULL. */
int  cse_vote(const void* d_probs, int probs_dtype_is_f64, const double* d_weights, int mode,
              int M, int N, int C, int32_t* d_pred, double* d_summed, void* stream);

/* Batched weighted vote for the ensemble-weight searches (evaluate_ensemble.py:302-339):
 * W weight vectors [W,M] at once; d_correct int32 [W] = number of clips whose vote equals label. */
int  cse_vote_search(const double* d_probs /*[M,N,C]*/, const double* d_weights /*[W,M]*/,
                     const int32_t* d_labels /*[N]*/, int W, int M, int N, int C,
                     int32_t* d_correct, void* stream);

/* ---- clip assembly: replaces select_frames + cv2.resize of get_onestream_videoclip / get_twostream_videoclip ---- */
/* (train.py:132-145, 279-291, 196-221).  d_frames: the decoded frames of one video, uint8
 * [n_frames, Hs, Ws, C] (BGR, or gray flow frames), C <= 4.  Keeps frames t*step, t < T, step = max(1, n_frames / T)
 * and resizes each to H x W exactly like OpenCV's INTER_LINEAR on CV_8U (11-bit fixed point): d_clip uint8 [T, H, W, C]
 * is bit-identical to the reference's CPU result.  Fails when select_frames would keep fewer than T frames. */
int  cse_assemble_clip(const uint8_t* d_frames, int n_frames, int Hs, int Ws, int C,
                       uint8_t* d_clip, int T, int H, int W, void* stream);

/* ---- on-the-fly dense optical flow: replaces opticalflow_FarneBack_extractor (train.py:294-332) ---- */
/* cv2.resize on CV_8U images [n,Hs,Ws,C] -> [n,H,W,C], bit-identical: fx = fy = 0 is cv2.resize(img, (W, H));
 * fx, fy > 0 is cv2.resize(img, None, fx=, fy=) (the caller passes H = cvRound(Hs * fy), W = cvRound(Ws * fx); the
 * taps come from 1 / factor, as OpenCV computes them) - the extractor's frame scaling (train.py:300-312). */
int  cse_resize_u8(const uint8_t* d_src, int n, int Hs, int Ws, int C, uint8_t* d_dst, int H, int W,
                   double fx, double fy, void* stream);
/* cv2.cvtColor(COLOR_BGR2GRAY) on CV_8U, bit-identical: `pixels` BGR triples -> gray bytes. */
int  cse_bgr2gray(const uint8_t* d_bgr, uint8_t* d_gray, long long pixels, void* stream);
/* cv2.calcOpticalFlowFarneback(prev, next, None, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags=0)
 * between every pair of CONSECUTIVE frames of d_gray uint8 [n_frames,H,W] -> d_flow fp32 [n_frames-1,H,W,2] (dx, dy).
 * Bit-identical to oracle/farneback.py (the restatement of OpenCV's algorithm), which agrees with cv2 itself to
 * <= 1e-4 pixel (cv2's SIMD / IPP summation order is not reproducible).  d_work: cse_farneback_workspace_bytes(),
 * 8-byte aligned.  The reference calls it with (0.5, 5, 11, 5, 5, 1.1). */
size_t cse_farneback_workspace_bytes(int n_frames, int H, int W);
int  cse_farneback(const uint8_t* d_gray, int n_frames, int H, int W, double pyr_scale, int levels, int winsize,
                   int iterations, int poly_n, double poly_sigma, float* d_flow, void* d_work, size_t work_bytes,
                   void* stream);
/* cv2.resize on CV_32F images [n,Hs,Ws,C] -> [n,H,W,C] (INTER_LINEAR), bit-identical: the flow fields resized to the
 * network's input size (get_twostream_videoclip, train.py:223-239). */
int  cse_resize_linear_f32(const float* d_src, int n, int Hs, int Ws, int C, float* d_dst, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSE_B200_H */
