// oneshot.cu -- SORT-FREE unsorted segment reduce: one streaming pass over the messages, reductions resolved by the L2
// (SURVEY.md 8b `gglb200_segment_reduce`).
//
//   out[ids[e], k]  (+|mean|max)=  x[e, k]                     e < E, x read once, coalesced, in its ORIGINAL order
//
// The CSR path (gglb200_csr_build + gglb200_spmm_csr(col = perm)) sorts the ids once and then gathers x[perm[i]] per pass:
// the right trade when the ids are reused (every layer of a full-graph model).  When they change per step (DropEdge, sampled
// sub-graphs, coalesce) the sort is 6 ms at 1.2e8 edges, more than the reduce itself; and for very narrow rows (degree():
// unsorted_segment_sum(ones[E]), /root/reference/gammagl/utils/degree.py:35-40, K = 1) the permuted 4-byte gather wastes
// 7/8 of every DRAM sector.  This path needs neither: replaces the reference's scalar-atomic kernels
// (mpops/torch_ext/cuda/segment_{sum,mean,max}_cuda.cu) with vector reductions and an exact arg-max.
//
//   SUM    red.global.add.v4.f32 / .v2 / scalar on out (zero-filled first).  Summation order = arrival order: results differ
//          from the serial reference by rounding only (<= 1e-6 relative; exact for integer-valued data such as degree());
//          callers that need run-to-run bit-stable fp32 sums use the CSR path.
//   MEAN   SUM + red.global.add.u32 on cnt[ids[e]], then rows with cnt > 1 divided (cpu/segment_mean_cpu.cpp:67-76).
//   MAX    DETERMINISTIC and bit-exact, arg included: atom.max.u64 on key = ordered(x) << 32 | ~e -- the largest value wins and
//          among equal values the SMALLEST edge id, i.e. the first maximal edge of the reference's ascending loop with `<`
//          (cpu/segment_max_cpu.cpp:54-62).  -0.0 and +0.0 compare equal like there; NaN never replaces (x < NaN is false).
//          A second pass decodes: out = x[arg] (exact bits), arg = edge id, empty rows = lowest() / arg_empty.
// ids outside [0, N) are skipped and counted in *n_bad (the host raises); nothing is clamped.
#include "common.cuh"
#include <cstdlib>

namespace ggl {

__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// monotone map float -> uint32 (total order of the finite values and infinities; -0.0 is folded onto +0.0 first)
__device__ __forceinline__ uint32_t ordered_bits(float v) {
  uint32_t b = __float_as_uint(v);
  if (b == 0x80000000u) b = 0u;
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

// A group of LPR lanes owns one edge row of KV = K / VEC vector elements (LPR = next power of two >= min(KV, 32)); a warp
// moves 32 / LPR edges per step, U steps in flight.
template <int VEC, int LPR, int MODE /* 0 sum, 1 max */, int U = 4>
__global__ void __launch_bounds__(256) oneshot_kernel(const float* __restrict__ x, const int64_t* __restrict__ ids, float* __restrict__ out,
                                                      unsigned long long* __restrict__ keys, uint32_t* __restrict__ cnt,
                                                      unsigned long long* __restrict__ n_bad, int64_t E, int32_t K, int64_t N,
                                                      int32_t rep_mask /* SUM: replicas - 1 (0 = none) */, int64_t rep_stride) {
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1), grp = lane / LPR;
  const int32_t KV = K / VEC;
  const int64_t span = (int64_t)gridDim.x * (blockDim.x >> 5) * (GPW * U);
  unsigned long long bad = 0;
  for (int64_t e0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (GPW * U); e0 < E; e0 += span) {
    int64_t r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * GPW + grp;
      r[u] = -1;
      if (e < E) {
        const int64_t id = __ldcs(ids + e);
        if (id >= 0 && id < N) r[u] = id;
        else if (sub == 0) ++bad;
      }
    }
    if (cnt != nullptr && sub == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r[u] >= 0) asm volatile("red.global.add.u32 [%0], 1;" ::"l"(cnt + r[u]) : "memory");
    }
    for (int32_t kv = sub; kv < KV; kv += LPR) {
      float v[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r[u] >= 0) {
          const float* p = x + (e0 + u * GPW + grp) * K + (int64_t)kv * VEC;
          if constexpr (VEC == 4) { const float4 t = __ldcs(reinterpret_cast<const float4*>(p)); v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w; }
          else if constexpr (VEC == 2) { const float2 t = __ldcs(reinterpret_cast<const float2*>(p)); v[u][0] = t.x; v[u][1] = t.y; }
          else v[u][0] = __ldcs(p);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r[u] < 0) continue;
        int64_t o = r[u] * K + (int64_t)kv * VEC;
        if constexpr (MODE == 0) {
          // hub segments: consecutive edges land in different replicas of the output, so reductions on one hot address do not
          // queue behind each other in a single L2 slice (folded in fixed order by oneshot_fold_kernel)
          o += ((e0 + u * GPW + grp) & rep_mask) * rep_stride;
          if constexpr (VEC == 4) red_add4(out + o, v[u][0], v[u][1], v[u][2], v[u][3]);
          else if constexpr (VEC == 2) red_add2(out + o, v[u][0], v[u][1]);
          else red_add(out + o, v[u][0]);
        } else {
          const uint32_t e32 = (uint32_t)(e0 + u * GPW + grp);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if (v[u][i] != v[u][i]) continue;  // NaN never replaces the running maximum (`out < x` is false)
            atomicMax(keys + o + i, ((unsigned long long)ordered_bits(v[u][i]) << 32) | (unsigned long long)(0xffffffffu - e32));
          }
        }
      }
    }
  }
  if (bad && n_bad) atomicAdd(n_bad, bad);  // error path only
}

// SUM with replicas: out[t] = rep[0][t] + rep[1][t] + ... (fixed order)
__global__ void __launch_bounds__(256) oneshot_fold_kernel(const float* __restrict__ rep, float* __restrict__ out, int64_t total, int32_t replicas) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    float a = rep[t];
    for (int32_t r = 1; r < replicas; ++r) a += rep[(int64_t)r * total + t];
    out[t] = a;
  }
}

// MEAN: rows with more than one message are divided by their count (cpu/segment_mean_cpu.cpp:67-76)
__global__ void __launch_bounds__(256) oneshot_mean_kernel(float* __restrict__ out, const uint32_t* __restrict__ cnt, int64_t total, int32_t K) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const uint32_t c = __ldg(cnt + t / K);
    if (c > 1) out[t] = __fdiv_rn(out[t], (float)c);
  }
}

// MAX: key -> (value bits re-read from x, edge id); untouched keys (0) are empty rows
__global__ void __launch_bounds__(256) oneshot_max_decode_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ x,
                                                                 float* __restrict__ out, int64_t* __restrict__ arg, int64_t total, int32_t K,
                                                                 int64_t arg_empty) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const unsigned long long key = keys[t];
    if (key == 0ull) {
      out[t] = Elem<float>::lowest();
      if (arg) arg[t] = arg_empty;
    } else {
      const int64_t e = (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
      out[t] = __ldg(x + e * K + (t % K));
      if (arg) arg[t] = e;
    }
  }
}

// ---- sort-free edge softmax (gammagl/utils/softmax.py:29-35) -------------------------------------------------------------------
// The CSR form gathers 4H-byte logit rows through `perm` (request-bound: 5.5 ms at 1.15e8 edges, H = 8).  Here every pass streams
// [E, H] in its original order: max by red.global.max.u32 on order-preserving keys, then sum(exp(x - m)) and the backward's
// sum(alpha * galpha) by vector reductions; the element-wise passes are the existing ones (gat.cu).  No CSR, no sort.
__global__ void __launch_bounds__(256) softmax_max_kernel(const float* __restrict__ x, const int64_t* __restrict__ ids, uint32_t* __restrict__ mkey,
                                                          int64_t total, int32_t H, int64_t N) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int64_t id = __ldg(ids + e);
    const float v = __ldcs(x + t);
    if (id < 0 || id >= N || v != v) continue;
    asm volatile("red.global.max.u32 [%0], %1;" ::"l"(mkey + id * H + (t - e * H)), "r"(ordered_bits(v)) : "memory");
  }
}
// keys -> floats in place (untouched keys = segments without edges: 0, never read by an edge)
__global__ void __launch_bounds__(256) softmax_decode_kernel(uint32_t* __restrict__ mkey, int64_t total) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const uint32_t k = mkey[t];
    const uint32_t b = k == 0u ? 0u : (k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
    mkey[t] = b;
  }
}
// s[ids[e], h] += exp(x[e,h] - m[ids[e],h]); p written to alpha when alpha != NULL (the H % 4 != 0 composition divides it in place)
template <int VEC>
__global__ void __launch_bounds__(256) softmax_sum_kernel(const float* __restrict__ x, const float* __restrict__ m, const int64_t* __restrict__ ids,
                                                          float* __restrict__ ssum, float* __restrict__ p_out, int64_t totalv, int32_t HV, int64_t N) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < totalv; t += stride) {
    const int64_t e = t / HV;
    const int64_t id = __ldg(ids + e);
    if (id < 0 || id >= N) continue;
    const int64_t o = (id * HV + (t - e * HV)) * VEC;
    if constexpr (VEC == 4) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + t);
      const float4 mm = __ldg(reinterpret_cast<const float4*>(m + o));
      red_add4(ssum + o, expf(v.x - mm.x), expf(v.y - mm.y), expf(v.z - mm.z), expf(v.w - mm.w));
    } else {
      const float pv = expf(__ldcs(x + t) - __ldg(m + o));
      if (p_out) p_out[t] = pv;
      red_add(ssum + o, pv);
    }
  }
}
// dot[ids[e], h] += alpha[e,h] * galpha[e,h]
template <int VEC>
__global__ void __launch_bounds__(256) softmax_dotsum_kernel(const float* __restrict__ a, const float* __restrict__ g, const int64_t* __restrict__ ids,
                                                             float* __restrict__ dot, int64_t totalv, int32_t HV, int64_t N) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < totalv; t += stride) {
    const int64_t e = t / HV;
    const int64_t id = __ldg(ids + e);
    if (id < 0 || id >= N) continue;
    const int64_t o = (id * HV + (t - e * HV)) * VEC;
    if constexpr (VEC == 4) {
      const float4 av = __ldcs(reinterpret_cast<const float4*>(a) + t), gv = __ldcs(reinterpret_cast<const float4*>(g) + t);
      red_add4(dot + o, av.x * gv.x, av.y * gv.y, av.z * gv.z, av.w * gv.w);
    } else {
      red_add(dot + o, __ldcs(a + t) * __ldcs(g + t));
    }
  }
}

static int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static int oneshot_unroll() { return env_int("GGLB200_ONESHOT_UNROLL", 4); }

template <int VEC, int MODE>
static void launch_oneshot(const float* x, const int64_t* ids, float* out, unsigned long long* keys, uint32_t* cnt, unsigned long long* n_bad,
                           int64_t E, int32_t K, int64_t N, cudaStream_t stream, int32_t rep_mask = 0, int64_t rep_stride = 0) {
  const int kv = K / VEC;
  int lpr = 1;
  while (lpr < 32 && lpr < kv) lpr <<= 1;
  const int64_t warps = (E * lpr + 127) / 128;
  int64_t grid = (warps + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (lpr == 1 && oneshot_unroll() == 8) {   // experiment knob (GGLB200_ONESHOT_UNROLL): 8 edges per lane in flight
    oneshot_kernel<VEC, 1, MODE, 8><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride);
    return;
  }
  switch (lpr) {
    case 1: oneshot_kernel<VEC, 1, MODE><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride); break;
    case 2: oneshot_kernel<VEC, 2, MODE><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride); break;
    case 4: oneshot_kernel<VEC, 4, MODE><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride); break;
    case 8: oneshot_kernel<VEC, 8, MODE><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride); break;
    case 16: oneshot_kernel<VEC, 16, MODE><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride); break;
    default: oneshot_kernel<VEC, 32, MODE><<<(int)grid, 256, 0, stream>>>(x, ids, out, keys, cnt, n_bad, E, K, N, rep_mask, rep_stride); break;
  }
}

}  // namespace ggl

using namespace ggl;

constexpr int kReplicas = 8;                       // output replicas of narrow sums (hub contention), when they stay L2-resident
constexpr size_t kReplicaBudget = 48u << 20;

static int sum_replicas(int64_t N, int64_t K) {
  int r = ggl::env_int("GGLB200_ONESHOT_REPLICAS", kReplicas);   // power of two, at most kReplicas (the workspace is sized for that)
  if (r < 1 || r > kReplicas || (r & (r - 1))) r = kReplicas;
  return (K <= 4 && (size_t)N * (size_t)K * 4 * kReplicas <= kReplicaBudget) ? r : 1;
}

extern "C" size_t gglb200_segment_reduce_workspace_bytes(int reduce, int64_t N, int64_t K) {
  if (N <= 0 || K <= 0) return 0;
  if (reduce == GGL_MAX) return (size_t)N * (size_t)K * 8 + 256;
  return (K <= 4 && (size_t)N * (size_t)K * 4 * kReplicas <= kReplicaBudget) ? (size_t)N * (size_t)K * 4 * kReplicas + 256 : 0;
}

extern "C" int gglb200_segment_reduce(int reduce, const void* x_, const int64_t* ids, void* out_, int64_t* arg, int64_t arg_empty,
                                      int32_t* cnt, int64_t* n_bad, int64_t E, int64_t K, int64_t N, int dtype, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(reduce == GGL_SUM || reduce == GGL_MEAN || reduce == GGL_MAX, GGL_EINVAL, "segment_reduce: unknown reduce %d", reduce);
  GGL_REQUIRE(E >= 0 && K >= 0 && N >= 0, GGL_EINVAL, "segment_reduce: negative size");
  GGL_REQUIRE(E < (int64_t)0xffffffffll && K < INT32_MAX, GGL_ERANGE, "segment_reduce: E / K exceed 32 bits");
  GGL_REQUIRE(dtype == GGL_F32, GGL_EUNSUPPORTED, "segment_reduce: float32 only (other dtypes take the CSR path)");
  GGL_REQUIRE(arg == nullptr || reduce == GGL_MAX, GGL_EINVAL, "segment_reduce: arg is only produced by GGL_MAX");
  GGL_REQUIRE(reduce == GGL_SUM || cnt != nullptr, GGL_EINVAL, "segment_reduce: MEAN / MAX need the cnt[N] output");
  if (n_bad) GGL_CUDA(cudaMemsetAsync(n_bad, 0, sizeof(int64_t), stream));
  if (N == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(out_ != nullptr && (E == 0 || (x_ != nullptr && ids != nullptr)), GGL_EINVAL, "segment_reduce: NULL pointer");
  const float* x = (const float*)x_;
  float* out = (float*)out_;
  if (cnt) GGL_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)N, stream));
  unsigned long long* keys = nullptr;
  if (reduce == GGL_MAX) {
    GGL_REQUIRE(workspace != nullptr && workspace_bytes >= gglb200_segment_reduce_workspace_bytes(reduce, N, K), GGL_EWORKSPACE,
                "segment_reduce: workspace %zu < %zu", workspace_bytes, gglb200_segment_reduce_workspace_bytes(reduce, N, K));
    keys = (unsigned long long*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    GGL_CUDA(cudaMemsetAsync(keys, 0, (size_t)N * (size_t)K * 8, stream));
  }
  // narrow sums into an L2-resident output: 8 replicas take the queueing out of hub segments (C4 degree(): 1.75 ms without)
  float* rep = nullptr;
  const int replicas = (reduce != GGL_MAX && workspace != nullptr && workspace_bytes >= gglb200_segment_reduce_workspace_bytes(reduce, N, K))
                           ? sum_replicas(N, K) : 1;
  if (reduce != GGL_MAX) {
    if (replicas > 1) {
      rep = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
      GGL_CUDA(cudaMemsetAsync(rep, 0, (size_t)N * (size_t)K * 4 * replicas, stream));
    } else {
      GGL_CUDA(cudaMemsetAsync(out, 0, (size_t)N * (size_t)K * 4, stream));
    }
  }
  if (E > 0) {
    const uintptr_t al = (uintptr_t)x | (uintptr_t)out;
    const int vec = (K % 4 == 0 && al % 16 == 0) ? 4 : ((K % 2 == 0 && al % 8 == 0) ? 2 : 1);
    unsigned long long* nb = (unsigned long long*)n_bad;
    uint32_t* c = (uint32_t*)cnt;
    if (reduce == GGL_MAX) {
      if (vec == 4) launch_oneshot<4, 1>(x, ids, out, keys, c, nb, E, (int32_t)K, N, stream);
      else if (vec == 2) launch_oneshot<2, 1>(x, ids, out, keys, c, nb, E, (int32_t)K, N, stream);
      else launch_oneshot<1, 1>(x, ids, out, keys, c, nb, E, (int32_t)K, N, stream);
    } else {
      float* acc = rep ? rep : out;
      const int32_t rm = replicas - 1;
      const int64_t rs = N * K;
      if (vec == 4) launch_oneshot<4, 0>(x, ids, acc, keys, c, nb, E, (int32_t)K, N, stream, rm, rs);
      else if (vec == 2) launch_oneshot<2, 0>(x, ids, acc, keys, c, nb, E, (int32_t)K, N, stream, rm, rs);
      else launch_oneshot<1, 0>(x, ids, acc, keys, c, nb, E, (int32_t)K, N, stream, rm, rs);
    }
    GGL_LAUNCHED("oneshot_segment");
  }
  if (rep) {
    oneshot_fold_kernel<<<ew_grid(N * K), 256, 0, stream>>>(rep, out, N * K, replicas);
    GGL_LAUNCHED("oneshot_fold");
  }
  if (reduce == GGL_MEAN) {
    oneshot_mean_kernel<<<ew_grid(N * K), 256, 0, stream>>>(out, (const uint32_t*)cnt, N * K, (int32_t)K);
    GGL_LAUNCHED("oneshot_mean");
  } else if (reduce == GGL_MAX) {
    oneshot_max_decode_kernel<<<ew_grid(N * K), 256, 0, stream>>>(keys, x, out, arg, N * K, (int32_t)K, arg_empty);
    GGL_LAUNCHED("oneshot_max_decode");
  }
  return GGL_OK;
}

// statistics of the sort-free edge softmax: stat_m[N,H] (exact segment maxima, 0 for segments without edges) and stat_s[N,H] =
// sum exp(x - m); `p_or_null` receives exp(x - m) per edge when H % 4 != 0 (the caller divides it in place).
extern "C" int gglb200_edge_softmax_stats_oneshot(const int64_t* ids, const float* logit, float* stat_m, float* stat_s, float* p_or_null,
                                                  int64_t n_rows, int64_t E, int32_t H, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1, GGL_EINVAL, "edge_softmax_stats_oneshot: bad sizes");
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(ids && logit && stat_m && stat_s, GGL_EINVAL, "edge_softmax_stats_oneshot: NULL pointer");
  GGL_CUDA(cudaMemsetAsync(stat_m, 0, (size_t)n_rows * H * 4, stream));
  GGL_CUDA(cudaMemsetAsync(stat_s, 0, (size_t)n_rows * H * 4, stream));
  softmax_max_kernel<<<ew_grid(E * H), 256, 0, stream>>>(logit, ids, (uint32_t*)stat_m, E * H, H, n_rows);
  GGL_LAUNCHED("softmax_max");
  softmax_decode_kernel<<<ew_grid(n_rows * H), 256, 0, stream>>>((uint32_t*)stat_m, n_rows * H);
  GGL_LAUNCHED("softmax_decode");
  if (H % 4 == 0 && p_or_null == nullptr && (((uintptr_t)logit | (uintptr_t)stat_m | (uintptr_t)stat_s) & 15) == 0)
    softmax_sum_kernel<4><<<ew_grid(E * (H / 4)), 256, 0, stream>>>(logit, stat_m, ids, stat_s, nullptr, E * (H / 4), H / 4, n_rows);
  else
    softmax_sum_kernel<1><<<ew_grid(E * H), 256, 0, stream>>>(logit, stat_m, ids, stat_s, p_or_null, E * H, H, n_rows);
  GGL_LAUNCHED("softmax_sum");
  return GGL_OK;
}

// dot[N,H] = segment sum of alpha * galpha (the row term of the softmax Jacobian), sort-free
extern "C" int gglb200_edge_softmax_dot_oneshot(const int64_t* ids, const float* alpha, const float* galpha, float* dot, int64_t n_rows,
                                                int64_t E, int32_t H, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1, GGL_EINVAL, "edge_softmax_dot_oneshot: bad sizes");
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(ids && alpha && galpha && dot, GGL_EINVAL, "edge_softmax_dot_oneshot: NULL pointer");
  GGL_CUDA(cudaMemsetAsync(dot, 0, (size_t)n_rows * H * 4, stream));
  if (H % 4 == 0 && (((uintptr_t)alpha | (uintptr_t)galpha | (uintptr_t)dot) & 15) == 0)
    softmax_dotsum_kernel<4><<<ew_grid(E * (H / 4)), 256, 0, stream>>>(alpha, galpha, ids, dot, E * (H / 4), H / 4, n_rows);
  else
    softmax_dotsum_kernel<1><<<ew_grid(E * H), 256, 0, stream>>>(alpha, galpha, ids, dot, E * H, H, n_rows);
  GGL_LAUNCHED("softmax_dotsum");
  return GGL_OK;
}
