// spmm_impl.cuh -- the CSR-row-partitioned aggregation kernels (gather . weight . reduce).
//
//   out[r,k] = REDUCE_{i in row r}  w(i,k) * x[col[i],k]
//
// One kernel family serves every reference operator on the path:
//   c_spmm_sum/mean/max   (/root/reference/gammagl/mpops/torch_ext/cpu/spmm_{sum,mean,max}_cpu.cpp)   col = src, w = edge weight
//   c_bspmm_sum           (cpu/bspmm_sum_cpu.cpp:7-56)                                                 w = per-head weight
//   c_segment_sum/mean/max(cpu/segment_{sum,mean,max}_cpu.cpp)                                         col = perm (edge ids), no w
// and their backward passes wrt x (the same kernel on the transposed CSR).
//
// Mapping (HBM-bound gather; no tensor cores -- arithmetic intensity 0.5 flop/B):
//   * a group of LPR lanes owns one output row; each lane owns VEC consecutive features and moves
//     them with one 16-byte (or 8/4-byte) ld.global.nc per edge -- a K=100 fp32 row is 25 float4
//     lanes of one warp, fully coalesced 400 B;
//   * the group reads its col/weight stream cooperatively (one coalesced load of LPR indices), then
//     broadcasts index j with a shuffle and issues U=8 independent row gathers before any of them
//     is consumed (memory-level parallelism: 8 x 400 B in flight per warp, ~32 warps/SM);
//   * each lane accumulates sequentially in edge order (stable CSR => ascending original edge id),
//     multiply and add rounded separately (no FMA contraction) => fp32 sums are bit-identical to the
//     reference's serial CPU loop; max keeps the first winner (strict <);
//   * rows longer than GGLB200_LONG_ROW are skipped by the row kernel and reduced by the chunk kernel
//     (one group per GGLB200_LONG_CHUNK edges -> partials) + finalize kernel (ordered combine), which
//     bounds the tail on power-law graphs without float atomics.
//   * K wider than LPR*VEC is covered by blockIdx.y column tiles.
#pragma once
#include "common.cuh"

namespace ggl {

struct SpmmArgs {
  int reduce;               // GGL_SUM / GGL_MEAN / GGL_MAX
  const int32_t* rowptr;
  const int32_t* col;
  const int32_t* plan;      // NULL => no row splitting
  const float* w;           // NULL => unweighted
  const int32_t* wperm;     // NULL => w already in CSR order
  int32_t heads;            // 1 => w[E]; H => w[E,H]
  const void* x;
  void* out;
  int64_t* arg;             // MAX only, may be NULL
  int64_t arg_empty;
  int64_t n_rows, E, K;
  void* part_val;           // workspace: [max_chunks, K] acc_t
  int32_t* part_arg;        // workspace: [max_chunks, K] int32 (MAX)
  cudaStream_t stream;
};

// ---- arithmetic with explicit rounding (never contracted into FMA) --------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ int64_t add_rn(int64_t a, int64_t b) { return a + b; }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <int LPR> __device__ __forceinline__ unsigned group_mask(int lane) {
  if constexpr (LPR == 32) return 0xffffffffu;
  else return ((1u << LPR) - 1u) << (lane & ~(LPR - 1));
}

// Reduce edges [start, start+deg) of one row into the lane's VEC accumulators.
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__device__ __forceinline__ void reduce_range(typename Elem<T>::acc_t (&acc)[VEC], int32_t (&argv)[VEC],
                                             const SpmmArgs& p, const T* __restrict__ xk /* x + kcol */,
                                             int32_t head, bool col_active, int32_t start, int32_t deg, int lane) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int U = (LPR >= 8) ? 8 : (LPR == 1 ? 8 : LPR);
  const int sub = lane & (LPR - 1);
  const int gbase = lane & ~(LPR - 1);
  const unsigned gmask = group_mask<LPR>(lane);
  const int64_t K = p.K;

  constexpr int STEP = (LPR == 1) ? U : LPR;
  for (int base = 0; base < deg; base += STEP) {
    const int n = min(STEP, deg - base);
    int32_t mycol = 0, myeid = 0;
    float myw = 1.f;
    if constexpr (LPR > 1) {
      if (sub < n) {
        const int32_t idx = start + base + sub;
        mycol = ld_stream_i32(p.col + idx);
        if constexpr (WMODE != 0) myeid = p.wperm ? ld_stream_i32(p.wperm + idx) : idx;
        if constexpr (WMODE == 1) myw = p.wperm ? __ldg(p.w + myeid) : ld_stream_f32(p.w + myeid);
      }
    }
#pragma unroll 1
    for (int j = 0; j < n; j += U) {
      Pack<T, VEC> v[U];
      int32_t c[U];
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u;
        if constexpr (LPR > 1) {
          c[u] = __shfl_sync(gmask, mycol, gbase + (jj & (LPR - 1)));
          if constexpr (WMODE == 1) wv[u] = __shfl_sync(gmask, myw, gbase + (jj & (LPR - 1)));
          if constexpr (WMODE == 2) {
            int32_t eid = __shfl_sync(gmask, myeid, gbase + (jj & (LPR - 1)));
            wv[u] = (jj < n) ? __ldg(p.w + (int64_t)eid * p.heads + head) : 0.f;
          }
        } else {
          c[u] = 0; wv[u] = 1.f;
          if (jj < n) {
            const int32_t idx = start + base + jj;
            c[u] = ld_stream_i32(p.col + idx);
            if constexpr (WMODE != 0) {
              const int32_t eid = p.wperm ? ld_stream_i32(p.wperm + idx) : idx;
              wv[u] = (WMODE == 1) ? __ldg(p.w + eid) : __ldg(p.w + (int64_t)eid * p.heads + head);
            }
          }
        }
        if (jj < n && col_active) v[u] = ld_gather<T, VEC>(xk + (int64_t)c[u] * K);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < n && col_active) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            acc_t xv = Elem<T>::to_acc(v[u].v[i]);
            if constexpr (WMODE != 0 && Elem<T>::is_float) xv = mul_rn((acc_t)wv[u], xv);
            if constexpr (ISMAX) {
              if (acc[i] < xv) { acc[i] = xv; argv[i] = c[u]; }
            } else {
              acc[i] = add_rn(acc[i], xv);
            }
          }
        }
      }
    }
  }
}

template <typename T, int VEC, bool ISMAX>
__device__ __forceinline__ void finish_row(typename Elem<T>::acc_t (&acc)[VEC], int32_t (&argv)[VEC], const SpmmArgs& p,
                                           int64_t row, int32_t kcol, int32_t deg) {
  using acc_t = typename Elem<T>::acc_t;
  Pack<T, VEC> o;
  if (p.reduce == GGL_MEAN) {
    if constexpr (Elem<T>::is_float) {
      if (deg > 1) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = div_rn(acc[i], (acc_t)deg);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = Elem<T>::from_acc(acc[i]);
    } else {
      // integer mean: sum and count both live in scalar_t in the reference (cpu/segment_mean_cpu.cpp:44-76)
      T cnt = (T)deg;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        T s = Elem<T>::from_acc(acc[i]);
        if (cnt > (T)1) s = (T)(s / cnt);
        o.v[i] = s;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = Elem<T>::from_acc(acc[i]);
  }
  T* op = reinterpret_cast<T*>(p.out) + row * p.K + kcol;
  st_stream<T, VEC>(op, o);
  if constexpr (ISMAX) {
    if (p.arg) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) p.arg[row * p.K + kcol + i] = argv[i] < 0 ? p.arg_empty : (int64_t)argv[i];
    }
  }
}

constexpr int kWarpsPerBlock = 8;

// ---- kernel 1: one LPR-lane group per row ----------------------------------------------------------
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spmm_rows_kernel(const SpmmArgs p) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (LPR - 1);
  const int64_t row = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * RPW + lane / LPR;
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  const bool col_active = kcol < p.K;
  const bool valid = row < p.n_rows;
  int32_t start = 0, deg = 0;
  if (valid) {
    start = __ldg(p.rowptr + row);
    deg = __ldg(p.rowptr + row + 1) - start;
  }
  const bool is_long = p.plan != nullptr && deg > GGLB200_LONG_ROW;
  acc_t acc[VEC];
  int32_t argv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { acc[i] = ISMAX ? Elem<T>::lowest() : (acc_t)0; argv[i] = -1; }
  const int32_t head = (WMODE == 2 && col_active) ? kcol / (int32_t)(p.K / p.heads) : 0;
  const T* xk = reinterpret_cast<const T*>(p.x) + kcol;
  reduce_range<T, VEC, LPR, WMODE, ISMAX>(acc, argv, p, xk, head, col_active, start, is_long ? 0 : deg, lane);
  if (valid && !is_long && col_active) finish_row<T, VEC, ISMAX>(acc, argv, p, row, kcol, deg);
}

// ---- kernel 2: one group per chunk of a long row -> partials ---------------------------------------
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spmm_chunks_kernel(const SpmmArgs p) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (LPR - 1);
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  const bool col_active = kcol < p.K;
  const int32_t n_chunks = min(__ldg(p.plan + 0), __ldg(p.plan + 2));
  const int32_t* chunk_tab = p.plan + 4;
  const int32_t head = (WMODE == 2 && col_active) ? kcol / (int32_t)(p.K / p.heads) : 0;
  const T* xk = reinterpret_cast<const T*>(p.x) + kcol;
  const int64_t groups = (int64_t)gridDim.x * kWarpsPerBlock * RPW;
  for (int64_t c = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * RPW + lane / LPR; c < n_chunks; c += groups) {
    const int32_t row = __ldg(chunk_tab + 2 * c), first = __ldg(chunk_tab + 2 * c + 1);
    const int32_t end = min(first + GGLB200_LONG_CHUNK, __ldg(p.rowptr + row + 1));
    acc_t acc[VEC];
    int32_t argv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { acc[i] = ISMAX ? Elem<T>::lowest() : (acc_t)0; argv[i] = -1; }
    reduce_range<T, VEC, LPR, WMODE, ISMAX>(acc, argv, p, xk, head, col_active, first, end - first, lane);
    if (col_active) {
      acc_t* pv = reinterpret_cast<acc_t*>(p.part_val) + c * p.K + kcol;
#pragma unroll
      for (int i = 0; i < VEC; ++i) pv[i] = acc[i];
      if constexpr (ISMAX) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p.part_arg[c * p.K + kcol + i] = argv[i];
      }
    }
  }
}

// ---- kernel 3: ordered combine of the partials of each long row ------------------------------------
template <typename T, int VEC, int LPR, bool ISMAX>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spmm_finalize_kernel(const SpmmArgs p) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (LPR - 1);
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  if (kcol >= p.K) return;
  const int32_t max_chunks = __ldg(p.plan + 2);
  const int32_t n_long = min(__ldg(p.plan + 1), __ldg(p.plan + 3));
  const int32_t* long_tab = p.plan + 4 + 2 * (int64_t)max_chunks;
  const int64_t groups = (int64_t)gridDim.x * kWarpsPerBlock * RPW;
  for (int64_t l = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * RPW + lane / LPR; l < n_long; l += groups) {
    const int32_t row = __ldg(long_tab + 3 * l), c0 = __ldg(long_tab + 3 * l + 1), nch = __ldg(long_tab + 3 * l + 2);
    acc_t acc[VEC];
    int32_t argv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { acc[i] = ISMAX ? Elem<T>::lowest() : (acc_t)0; argv[i] = -1; }
    for (int32_t j = 0; j < nch; ++j) {
      const acc_t* pv = reinterpret_cast<const acc_t*>(p.part_val) + (int64_t)(c0 + j) * p.K + kcol;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        acc_t v = pv[i];
        if constexpr (ISMAX) {
          if (acc[i] < v) { acc[i] = v; argv[i] = p.part_arg[(int64_t)(c0 + j) * p.K + kcol + i]; }
        } else {
          acc[i] = add_rn(acc[i], v);
        }
      }
    }
    const int32_t deg = __ldg(p.rowptr + row + 1) - __ldg(p.rowptr + row);
    finish_row<T, VEC, ISMAX>(acc, argv, p, row, kcol, deg);
  }
}

// ---- host-side launch ------------------------------------------------------------------------------
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
int launch_config(const SpmmArgs& p) {
  constexpr int RPW = 32 / LPR;
  const int tiles = (int)((p.K + (int64_t)LPR * VEC - 1) / ((int64_t)LPR * VEC));
  const int64_t rows_per_block = (int64_t)kWarpsPerBlock * RPW;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  if (blocks > 0) {
    GGL_REQUIRE(blocks < INT32_MAX && tiles <= 65535, GGL_EUNSUPPORTED, "spmm: grid too large (%lld x %d)", (long long)blocks, tiles);
    spmm_rows_kernel<T, VEC, LPR, WMODE, ISMAX><<<dim3((unsigned)blocks, tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
    GGL_LAUNCHED("spmm_rows");
  }
  if (p.plan != nullptr && p.E > GGLB200_LONG_ROW) {
    const int g = kSMs * 4;
    spmm_chunks_kernel<T, VEC, LPR, WMODE, ISMAX><<<dim3(g, tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
    GGL_LAUNCHED("spmm_chunks");
    spmm_finalize_kernel<T, VEC, LPR, ISMAX><<<dim3(kSMs, tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
    GGL_LAUNCHED("spmm_finalize");
  }
  return GGL_OK;
}

template <typename T, int VEC, int LPR>
int launch_modes(const SpmmArgs& p, bool allow_weight) {
  const int wmode = p.w == nullptr ? 0 : (p.heads > 1 ? 2 : 1);
  const bool ismax = p.reduce == GGL_MAX;
  if (wmode == 0) return ismax ? launch_config<T, VEC, LPR, 0, true>(p) : launch_config<T, VEC, LPR, 0, false>(p);
  if constexpr (Elem<T>::is_float) {
    if (allow_weight) {
      if (wmode == 1) return ismax ? launch_config<T, VEC, LPR, 1, true>(p) : launch_config<T, VEC, LPR, 1, false>(p);
      return ismax ? launch_config<T, VEC, LPR, 2, true>(p) : launch_config<T, VEC, LPR, 2, false>(p);
    }
  }
  return set_error(GGL_EUNSUPPORTED, "spmm: edge weights are only supported for floating dtypes");
}

static inline int lpr_for(int64_t K, int vec) {
  int64_t need = (K + vec - 1) / vec;
  int l = 1;
  while (l < 32 && l < need) l <<= 1;
  return l;
}

// entry points per element type (explicitly instantiated in spmm_f32.cu / spmm_bf16.cu / spmm_misc.cu)
template <typename T> int launch_spmm(const SpmmArgs& p);

}  // namespace ggl
