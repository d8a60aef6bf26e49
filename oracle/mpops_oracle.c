/*
 * mpops_oracle.c -- CPU restatement of the reference's aggregation operators.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it;
 * the product path (gammagl_b200/) never does.
 *
 * Every function restates one serial loop of /root/reference/gammagl/mpops/torch_ext (cited per function).
 * Compile with -O2 -ffp-contract=off (the reference ships -O2 without FMA targets; products and sums are
 * rounded separately).  Parity status: PINNED -- checked against the reference's own golden vectors
 * (tests/mpops/torch_ops.py, tests/layers/conv/test_message_passing.py, docstrings of mpops/torch.py) and
 * against the compiled reference extension oracle/_ref on randomized inputs (tests/test_oracle.py,
 * tests/golden/).
 *
 * Deliberate deviations from the reference text (documented quirks):
 *   - seg_mean: the reference sizes its count buffer by E and divides rows e < E
 *     (cpu/segment_mean_cpu.cpp:44,67-76), which is only well defined for E >= N; this restatement counts
 *     per segment and divides every row n < N -- identical whenever the reference is well defined.
 *   - seg_max_bwd skips empty segments; the reference scatters their gradient into edge row N
 *     (src/segment_max.cpp:55-59).
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LOWEST_float (-FLT_MAX)
#define LOWEST_double (-DBL_MAX)
#define LOWEST_int8_t INT8_MIN
#define LOWEST_int16_t INT16_MIN
#define LOWEST_int32_t INT32_MIN
#define LOWEST_int64_t INT64_MIN
#define LOWEST__Float16 ((_Float16)-65504.0f)

/* ---- segment ops: cpu/segment_sum_cpu.cpp:47-56, cpu/segment_mean_cpu.cpp:50-76, cpu/segment_max_cpu.cpp:40-64 ---- */
#define DEFINE_SEGMENT(T, SUF)                                                                              \
  void seg_sum_##SUF(const T* x, const int64_t* idx, int64_t E, int64_t K, int64_t N, T* out) {             \
    memset(out, 0, sizeof(T) * (size_t)(N * K));                                                            \
    for (int64_t e = 0; e < E; ++e) {                                                                       \
      int64_t i = idx[e];                                                                                   \
      for (int64_t k = 0; k < K; ++k) out[i * K + k] += x[e * K + k];                                       \
    }                                                                                                       \
  }                                                                                                         \
  void seg_mean_##SUF(const T* x, const int64_t* idx, int64_t E, int64_t K, int64_t N, T* out) {            \
    T* degree = (T*)calloc((size_t)(N > 0 ? N : 1), sizeof(T));                                             \
    memset(out, 0, sizeof(T) * (size_t)(N * K));                                                            \
    for (int64_t e = 0; e < E; ++e) {                                                                       \
      int64_t i = idx[e];                                                                                   \
      degree[i] += 1;                                                                                       \
      for (int64_t k = 0; k < K; ++k) out[i * K + k] += x[e * K + k];                                       \
    }                                                                                                       \
    for (int64_t n = 0; n < N; ++n)                                                                         \
      if (degree[n] > 1)                                                                                    \
        for (int64_t k = 0; k < K; ++k) out[n * K + k] /= degree[n];                                        \
    free(degree);                                                                                           \
  }                                                                                                         \
  void seg_max_##SUF(const T* x, const int64_t* idx, int64_t E, int64_t K, int64_t N, T* out, int64_t* arg) { \
    for (int64_t t = 0; t < N * K; ++t) { out[t] = LOWEST_##T; arg[t] = N; }                                \
    for (int64_t e = 0; e < E; ++e) {                                                                       \
      int64_t i = idx[e];                                                                                   \
      for (int64_t k = 0; k < K; ++k) {                                                                     \
        T cur = x[e * K + k];                                                                               \
        if (out[i * K + k] < cur) { out[i * K + k] = cur; arg[i * K + k] = e; }                             \
      }                                                                                                     \
    }                                                                                                       \
  }

DEFINE_SEGMENT(float, f32)
DEFINE_SEGMENT(double, f64)
DEFINE_SEGMENT(_Float16, f16)
DEFINE_SEGMENT(int8_t, i8)
DEFINE_SEGMENT(int16_t, i16)
DEFINE_SEGMENT(int32_t, i32)
DEFINE_SEGMENT(int64_t, i64)

/* SegmentSum::backward, src/segment_sum.cpp:43-54: grad_x = grad_out.index_select(0, index) */
void seg_sum_bwd_f32(const float* g, const int64_t* idx, int64_t E, int64_t K, float* gx) {
  for (int64_t e = 0; e < E; ++e) memcpy(gx + e * K, g + idx[e] * K, sizeof(float) * (size_t)K);
}
/* SegmentMean::backward, src/segment_mean.cpp:44-63: grad_out[index] / bincount(index)[index] */
void seg_mean_bwd_f32(const float* g, const int64_t* idx, int64_t E, int64_t K, int64_t N, float* gx) {
  int64_t* cnt = (int64_t*)calloc((size_t)(N > 0 ? N : 1), sizeof(int64_t));
  for (int64_t e = 0; e < E; ++e) cnt[idx[e]] += 1;
  for (int64_t e = 0; e < E; ++e)
    for (int64_t k = 0; k < K; ++k) gx[e * K + k] = g[idx[e] * K + k] / (float)cnt[idx[e]];
  free(cnt);
}
/* SegmentMax::backward, src/segment_max.cpp:48-61: zeros[E+1,K].scatter_(0, arg, grad_out)[:E]; empty rows skipped */
void seg_max_bwd_f32(const float* g, const int64_t* arg, const int64_t* idx, int64_t E, int64_t K, int64_t N, float* gx) {
  char* nonempty = (char*)calloc((size_t)(N > 0 ? N : 1), 1);
  for (int64_t e = 0; e < E; ++e) nonempty[idx[e]] = 1;
  memset(gx, 0, sizeof(float) * (size_t)(E * K));
  for (int64_t n = 0; n < N; ++n) {
    if (!nonempty[n]) continue;
    for (int64_t k = 0; k < K; ++k) gx[arg[n * K + k] * K + k] = g[n * K + k];
  }
  free(nonempty);
}

/* ---- gspmm: index[2,E] row 0 = src, row 1 = dst (read as index[e], index[e+E]) ---- */
/* cpu/spmm_sum_cpu.cpp:27-39 */
void spmm_sum_fwd(const int64_t* index, const float* w, const float* x, int64_t E, int64_t N, int64_t K, float* out) {
  memset(out, 0, sizeof(float) * (size_t)(N * K));
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t k = 0; k < K; ++k) out[dst * K + k] += w[e] * x[src * K + k];
  }
}
/* cpu/spmm_sum_cpu.cpp:66-77 */
void spmm_sum_bwd(const int64_t* index, const float* w, const float* g, int64_t E, int64_t N, int64_t K, float* gx) {
  memset(gx, 0, sizeof(float) * (size_t)(N * K));
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t k = 0; k < K; ++k) gx[src * K + k] += w[e] * g[dst * K + k];
  }
}
/* cpu/spmm_mean_cpu.cpp:34-58 */
void spmm_mean_fwd(const int64_t* index, const float* w, const float* x, int64_t E, int64_t N, int64_t K, float* out,
                   int64_t* count) {
  memset(out, 0, sizeof(float) * (size_t)(N * K));
  memset(count, 0, sizeof(int64_t) * (size_t)N);
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    count[dst]++;
    for (int64_t k = 0; k < K; ++k) out[dst * K + k] += w[e] * x[src * K + k];
  }
  for (int64_t n = 0; n < N; ++n)
    if (count[n] > 0)
      for (int64_t k = 0; k < K; ++k) out[n * K + k] /= (float)count[n];
}
/* cpu/spmm_mean_cpu.cpp:91-102 */
void spmm_mean_bwd(const int64_t* index, const float* w, const float* g, const int64_t* count, int64_t E, int64_t N,
                   int64_t K, float* gx) {
  memset(gx, 0, sizeof(float) * (size_t)(N * K));
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t k = 0; k < K; ++k) {
      float c = g[dst * K + k] / count[dst] * w[e];
      gx[src * K + k] += c;
    }
  }
}
/* cpu/spmm_max_cpu.cpp:21-53 */
void spmm_max_fwd(const int64_t* index, const float* w, const float* x, int64_t E, int64_t N, int64_t K, float* out,
                  int64_t* argsrc) {
  for (int64_t t = 0; t < N * K; ++t) { out[t] = -FLT_MAX; argsrc[t] = 0; }
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t k = 0; k < K; ++k) {
      float v = w[e] * x[src * K + k];
      if (out[dst * K + k] < v) { out[dst * K + k] = v; argsrc[dst * K + k] = src; }
    }
  }
}
/* cpu/spmm_max_cpu.cpp:84-96 */
void spmm_max_bwd(const int64_t* index, const float* w, const float* g, const int64_t* argsrc, int64_t E, int64_t N,
                  int64_t K, float* gx) {
  memset(gx, 0, sizeof(float) * (size_t)(N * K));
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t k = 0; k < K; ++k)
      if (argsrc[dst * K + k] == src) gx[src * K + k] += w[e] * g[dst * K + k];
  }
}
/* cpu/bspmm_sum_cpu.cpp:40-55 */
void bspmm_sum_fwd(const int64_t* index, const float* w, const float* x, int64_t E, int64_t N, int64_t H, int64_t C,
                   float* out) {
  memset(out, 0, sizeof(float) * (size_t)(N * H * C));
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t h = 0; h < H; ++h)
      for (int64_t k = 0; k < C; ++k) out[dst * C * H + h * C + k] += w[e * H + h] * x[src * C * H + h * C + k];
  }
}
/* cpu/bspmm_sum_cpu.cpp:90-108 */
void bspmm_sum_bwd(const int64_t* index, const float* w, const float* x, const float* g, int64_t E, int64_t N, int64_t H,
                   int64_t C, float* gx, float* gw) {
  memset(gx, 0, sizeof(float) * (size_t)(N * H * C));
  memset(gw, 0, sizeof(float) * (size_t)(E * H));
  for (int64_t e = 0; e < E; ++e) {
    int64_t src = index[e], dst = index[e + E];
    for (int64_t h = 0; h < H; ++h)
      for (int64_t k = 0; k < C; ++k) {
        gx[src * C * H + h * C + k] += w[e * H + h] * g[dst * C * H + h * C + k];
        gw[e * H + h] += x[src * C * H + h * C + k] * g[dst * C * H + h * C + k];
      }
  }
}

/* ind2ptr, gammagl/ops/sparse/cpu/convert.cpp:58-90 (sorted row ids -> CSR pointer) */
void ind2ptr_i64(const int64_t* ind, int64_t E, int64_t M, int64_t* ptr) {
  for (int64_t i = 0; i <= M; ++i) ptr[i] = 0;
  for (int64_t e = 0; e < E; ++e) ptr[ind[e] + 1]++;
  for (int64_t i = 0; i < M; ++i) ptr[i + 1] += ptr[i];
}

/* ptr2ind, gammagl/ops/sparse/cpu/convert.cpp:103-128 (CSR pointer -> row id of every element) */
void ptr2ind_i64(const int64_t* ptr, int64_t M, int64_t E, int64_t* ind) {
  for (int64_t i = 0; i < M; ++i)
    for (int64_t e = ptr[i]; e < ptr[i + 1] && e < E; ++e) ind[e] = i;
}
