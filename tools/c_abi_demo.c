/* A host that is neither Python nor PyTorch: plain C against include/cse.h.
 *
 * Builds one C3D member (train.py:1224-1273) through cse_model_* - the library constructs the graph, folds / packs the
 * weights and plans the buffers - runs a batch of synthetic uint8 clips and soft-votes two "members" (the same network
 * with two seeds) exactly as ensemble_predictions does (evaluate_ensemble.py:343-370).  Prints one line per clip.
 *
 *   gcc -O2 -I include tools/c_abi_demo.c -L crowded-scenes-ensemble-classification_b200 -lcse_b200 \
 *       -Wl,-rpath,$PWD/crowded-scenes-ensemble-classification_b200 -lm -o /tmp/c_abi_demo && /tmp/c_abi_demo
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "cse.h"

#define CHECK(call)                                                       \
  do {                                                                    \
    int rc_ = (call);                                                     \
    if (rc_ != 0) {                                                       \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cse_last_error()); \
      return 1;                                                           \
    }                                                                     \
  } while (0)

static uint64_t rng_state = 88172645463325252ull;
static double uniform01(void) {                 /* xorshift64 */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}
static float gauss(void) { return (float)(sqrt(-2.0 * log(uniform01() + 1e-300)) * cos(6.283185307179586 * uniform01())); }

static int fill_weights(cse_model* m, float head_scale) {
  const int layers = cse_model_num_layers(m);
  for (int l = 0; l < layers; ++l) {
    char lname[128];
    int nt = 0;
    CHECK(cse_model_layer_info(m, l, lname, sizeof lname, &nt));
    for (int t = 0; t < nt; ++t) {
      int64_t dims[5];
      int nd = 0;
      char tname[160];
      CHECK(cse_model_tensor_info(m, l, t, dims, &nd, tname, sizeof tname));
      size_t n = 1, fan_in = 1;
      for (int i = 0; i < nd; ++i) { n *= (size_t)dims[i]; if (i + 1 < nd) fan_in *= (size_t)dims[i]; }
      float* w = (float*)malloc(n * sizeof(float));
      if (!w) return 1;
      const int is_kernel = nd > 1;
      const float sd = is_kernel ? (float)sqrt(2.0 / (double)fan_in) * (l + 1 == layers ? head_scale : 1.f) : 0.f;
      for (size_t i = 0; i < n; ++i) w[i] = is_kernel ? gauss() * sd : 0.f;     /* He-scaled kernels, zero biases */
      CHECK(cse_model_set_weight(m, l, t, w, dims, nd));
      free(w);
    }
  }
  return 0;
}

int main(void) {
  enum { N = 4, T = 16, H = 112, W = 112, CLASSES = 11, MEMBERS = 2 };
  int sm = 0, major = 0, minor = 0;
  CHECK(cse_device_info(&sm, &major, &minor));
  printf("libcse_b200 ABI %d on a cc %d.%d device with %d SMs\n", cse_abi_version(), major, minor, sm);

  const size_t clip_bytes = (size_t)N * T * H * W * 3;
  uint8_t* h_clips = (uint8_t*)malloc(clip_bytes);
  for (size_t i = 0; i < clip_bytes; ++i) h_clips[i] = (uint8_t)(uniform01() * 256.0);
  void *d_clips = NULL, *d_probs = NULL, *d_pred = NULL;
  CHECK(cse_malloc(&d_clips, clip_bytes));
  CHECK(cse_malloc(&d_probs, sizeof(float) * MEMBERS * N * CLASSES));
  CHECK(cse_malloc(&d_pred, sizeof(int32_t) * N));
  CHECK(cse_memcpy_h2d(d_clips, h_clips, clip_bytes, NULL));

  cse_model* members[MEMBERS];
  for (int j = 0; j < MEMBERS; ++j) {
    CHECK(cse_model_create(&members[j], "C3D", T, H, W, CLASSES, CSE_BF16, N));
    rng_state += 1000003ull * (uint64_t)(j + 1);
    if (fill_weights(members[j], 1e-3f)) return 1;
    CHECK(cse_model_finalize(members[j], NULL, 0));
    CHECK(cse_model_forward(members[j], d_clips, NULL, N, NULL, (float*)d_probs + (size_t)j * N * CLASSES, NULL));
  }
  CHECK(cse_vote(d_probs, 0, NULL, 0, MEMBERS, N, CLASSES, (int32_t*)d_pred, NULL, NULL));     /* SUM vote */

  float h_probs[MEMBERS * N * CLASSES];
  int32_t h_pred[N];
  CHECK(cse_memcpy_d2h(h_probs, d_probs, sizeof h_probs, NULL));
  CHECK(cse_memcpy_d2h(h_pred, d_pred, sizeof h_pred, NULL));
  CHECK(cse_stream_synchronize(NULL));
  int ok = 1;
  for (int n = 0; n < N; ++n) {
    double best = -1.0;
    int arg = 0;
    for (int c = 0; c < CLASSES; ++c) {
      double s = 0.0, row = 0.0;
      for (int j = 0; j < MEMBERS; ++j) s += (double)h_probs[((size_t)j * N + n) * CLASSES + c];
      for (int j = 0; j < 1; ++j) for (int cc = 0; cc < CLASSES; ++cc) row += h_probs[((size_t)j * N + n) * CLASSES + cc];
      if (fabs(row - 1.0) > 1e-4) ok = 0;
      if (s > best) { best = s; arg = c; }
    }
    printf("clip %d: ensemble class %d (host re-vote %d)\n", n, h_pred[n], arg);
    if (arg != h_pred[n]) ok = 0;
  }
  for (int j = 0; j < MEMBERS; ++j) cse_model_destroy(members[j]);
  cse_free(d_clips); cse_free(d_probs); cse_free(d_pred);
  free(h_clips);
  printf(ok ? "c_abi_demo ok\n" : "c_abi_demo FAILED\n");
  return ok ? 0 : 1;
}
