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
//   * a group of LPR lanes owns one output row at a time; each lane owns VEC consecutive features and moves
//     them 16 bytes (or 8/4) at a time -- a K=100 fp32 row is 25 float4 lanes of one warp, coalesced 400 B;
//   * work is cut into EDGE TILES (GGLB200_TILE_EDGES edges; a row belongs to the tile its first edge falls
//     in), so every lane group streams about the same number of edges whatever the degree distribution;
//     the group walks its edges as ONE FLAT STREAM and flushes its accumulator at row boundaries;
//   * the gather is latency-bound, not issue-bound (first profile: ~20 KB in flight per SM, DRAM 35 %): every
//     floating configuration therefore stages rows through a shared-memory ring filled by bulk async copies
//     (cp.async.bulk / UBLKCP, one instruction per row, mbarrier per slot) -- 16 rows per warp in flight without
//     holding registers, one warp per block so that the hardware block scheduler balances tiles
//     (spmm_ring_kernel);
//   * each lane accumulates sequentially in edge order (stable CSR => ascending original edge id),
//     multiply and add rounded separately (no FMA contraction) => fp32 sums are bit-identical to the
//     reference's serial CPU loop; max keeps the first winner (strict <);
//   * rows longer than GGLB200_LONG_ROW are reduced chunk-wise (GGLB200_LONG_CHUNK edges -> partials, in the
//     same launch for the ring kernel) + a finalize kernel (ordered combine): bounded tail on power-law
//     graphs, deterministic, no float atomics.  Such a row is always the LAST row of its tile (it outruns
//     the tile window; flagged in the tile table), so the tile stream simply stops in front of it.
//   * peer mode (x_table != NULL): col carries owner << shift | local row and the bulk copy reads the owner
//     GPU's HBM over NVLink -- the 1-D node partition needs no pack / all-to-all / halo buffer.
//   * K wider than LPR*VEC is covered by blockIdx.y column tiles.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace ggl {

struct SpmmArgs {
  int reduce;               // GGL_SUM / GGL_MEAN / GGL_MAX
  const int32_t* rowptr;
  const int32_t* col;
  const int32_t* plan;      // plan array of gglb200_csr_build (tile table + long-row work list)
  const int32_t* tile_row;  // plan + tile table offset
  int64_t n_tiles;
  bool has_long;            // false => caller asserts no row exceeds GGLB200_LONG_ROW (workspace == NULL)
  int32_t ws_chunks;        // number of chunk partials the workspace can hold
  const float* w;           // NULL => unweighted
  const int32_t* wperm;     // NULL => w already in CSR order
  int32_t heads;            // 1 => w[E]; H => w[E,H]
  const void* x;
  const void* const* x_table;  // NULL, or device array of per-rank base pointers of x (peer memory over NVLink)
  int32_t owner_shift;      // peer mode: col = owner << owner_shift | local row
  void* out;
  int64_t* arg;             // MAX only, may be NULL
  int64_t arg_empty;
  bool arg_eid;             // MAX: arg records wperm[i] (the ORIGINAL EDGE ID of the winning edge) instead of col[i]; needs wperm
  const float* out_scale;   // NULL, or [n_rows] fp32: out[r,:] = out_scale[r] * reduce(...) (folded GCN normalisation)
  const float* bias;        // NULL, or [K] fp32 added after out_scale (GCNConv's bias, gcn_conv.py:104-106); SUM, floating out
  int32_t relu;             // != 0: out = max(out, 0) after the bias (the activation the models apply next)
  int64_t n_rows, E, K;
  int64_t n_src;            // rows of x (bounds the superset copies of unaligned rows)
  void* part_val;           // workspace: [ws_chunks, K] acc_t
  int32_t* part_arg;        // workspace: [ws_chunks, K] int32 (MAX)
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

// WMODE: 0 = unweighted, 1 = w[E], 2 = w[E,H] with every lane inside one head, 3 = w[E,H] with lanes that may span TWO heads
// (head width C not a multiple of VEC, C >= VEC: elements [0, nlo) take wv, the rest wv_hi).
template <typename T, int VEC, int WMODE, bool ISMAX>
__device__ __forceinline__ void accumulate(typename Elem<T>::acc_t (&acc)[VEC], int32_t (&argv)[VEC], const Pack<T, VEC>& v,
                                           float wv, int32_t c, float wv_hi = 0.f, int nlo = VEC) {
  using acc_t = typename Elem<T>::acc_t;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    acc_t xv = Elem<T>::to_acc(v.v[i]);
    if constexpr (WMODE == 3) {
      if (i >= nlo) wv = wv_hi;
    }
    if constexpr (WMODE != 0 && Elem<T>::is_float && sizeof(T) == 2 && !ISMAX) {
      // 16-bit storage has no bit-exact reference (the reference is fp32-only here; north_star tolerance 1e-2): one fused
      // multiply-add per element instead of two separately rounded operations -- these rows are issue-bound, not byte-bound
      acc[i] = fmaf((acc_t)wv, xv, acc[i]);
      continue;
    }
    if constexpr (WMODE != 0 && Elem<T>::is_float) xv = mul_rn((acc_t)wv, xv);
    if constexpr (ISMAX) {
      if (acc[i] < xv) { acc[i] = xv; argv[i] = c; }
    } else {
      acc[i] = add_rn(acc[i], xv);
    }
  }
}

template <typename T, int VEC, bool ISMAX>
__device__ __forceinline__ void reset_acc(typename Elem<T>::acc_t (&acc)[VEC], int32_t (&argv)[VEC]) {
  using acc_t = typename Elem<T>::acc_t;
#pragma unroll
  for (int i = 0; i < VEC; ++i) { acc[i] = ISMAX ? Elem<T>::lowest() : (acc_t)0; argv[i] = -1; }
}

// Reduce edges [start, start+deg) of one row into the lane's VEC accumulators (register-staged gathers).
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__device__ __forceinline__ void reduce_range(typename Elem<T>::acc_t (&acc)[VEC], int32_t (&argv)[VEC],
                                             const SpmmArgs& p, const T* __restrict__ xk /* x + kcol */,
                                             int32_t head, bool col_active, int32_t start, int32_t deg, int lane) {
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
        if constexpr (WMODE != 0 || ISMAX) {
          if (WMODE != 0 || p.arg_eid) myeid = p.wperm ? ld_stream_i32(p.wperm + idx) : idx;
        }
        if constexpr (WMODE == 1) myw = p.wperm ? __ldg(p.w + myeid) : ld_stream_f32(p.w + myeid);
      }
    }
#pragma unroll 1
    for (int j = 0; j < n; j += U) {
      Pack<T, VEC> v[U];
      int32_t c[U];
      int32_t ai[U];
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u;
        if constexpr (LPR > 1) {
          c[u] = __shfl_sync(gmask, mycol, gbase + (jj & (LPR - 1)));
          if constexpr (ISMAX) ai[u] = p.arg_eid ? __shfl_sync(gmask, myeid, gbase + (jj & (LPR - 1))) : c[u];
          wv[u] = 1.f;
          if constexpr (WMODE == 1) wv[u] = __shfl_sync(gmask, myw, gbase + (jj & (LPR - 1)));
          if constexpr (WMODE == 2) {
            int32_t eid = __shfl_sync(gmask, myeid, gbase + (jj & (LPR - 1)));
            wv[u] = (jj < n) ? __ldg(p.w + (int64_t)eid * p.heads + head) : 0.f;
          }
        } else {
          c[u] = 0; wv[u] = 1.f;
          if constexpr (ISMAX) ai[u] = 0;
          if (jj < n) {
            const int32_t idx = start + base + jj;
            c[u] = ld_stream_i32(p.col + idx);
            if constexpr (ISMAX) ai[u] = p.arg_eid ? (p.wperm ? ld_stream_i32(p.wperm + idx) : idx) : c[u];
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
        if (j + u < n && col_active) accumulate<T, VEC, WMODE, ISMAX>(acc, argv, v[u], wv[u], ISMAX ? ai[u] : c[u]);
      }
    }
  }
}

template <typename T, int VEC, bool ISMAX>
__device__ __forceinline__ void finish_row(typename Elem<T>::acc_t (&acc)[VEC], int32_t (&argv)[VEC], const SpmmArgs& p,
                                           int64_t row, int32_t kcol, int32_t deg) {
  using acc_t = typename Elem<T>::acc_t;
  Pack<T, VEC> o;
  if constexpr (Elem<T>::is_float && sizeof(T) <= 4) {
    if (p.out_scale) {
      const float sc = __ldg(p.out_scale + row);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = mul_rn(acc[i], sc);
    }
    if (p.bias) {  // separately rounded add, like `out + bias` after the aggregation in the reference
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = add_rn(acc[i], (acc_t)__ldg(p.bias + kcol + i));
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = acc[i] < (acc_t)0 ? (acc_t)0 : acc[i];   // NaN stays NaN like torch.relu
    }
  }
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

// Row cursor of a tile: rows [r, r_end) with a register cache of LPR rowptr entries per group.
template <int LPR>
struct RowCursor {
  int32_t r, r_end, rb, my_end, row_start, cur_end;
  __device__ __forceinline__ void init(const int32_t* __restrict__ rowptr, int32_t r0, int32_t r1, int32_t e0, int sub, int gbase,
                                       unsigned gmask) {
    r = r0; r_end = r1; rb = r0; row_start = e0;
    my_end = (rb + 1 + sub <= r_end) ? __ldg(rowptr + rb + 1 + sub) : INT32_MAX;
    const int32_t first_end = __shfl_sync(gmask, my_end, gbase);
    cur_end = (r < r_end) ? first_end : INT32_MAX;
  }
  // step to the next row (the caller has flushed row r)
  __device__ __forceinline__ void advance(const int32_t* __restrict__ rowptr, int sub, int gbase, unsigned gmask) {
    row_start = cur_end;
    ++r;
    if (r < r_end) {
      if (r - rb == LPR) {
        rb = r;
        my_end = (rb + 1 + sub <= r_end) ? __ldg(rowptr + rb + 1 + sub) : INT32_MAX;
      }
      cur_end = __shfl_sync(gmask, my_end, gbase + ((r - rb) & (LPR - 1)));
    } else {
      cur_end = INT32_MAX;
    }
  }
};

// Tile prologue shared by both tile kernels: rows [r0, r1) and the flat edge range [e0, e1) of tile `tile`.
// A long row (has_long && deg > LONG_ROW) can only be the last row a tile owns; it is left to the chunk path.
__device__ __forceinline__ bool tile_ranges(const SpmmArgs& p, int64_t tile, int32_t& r0, int32_t& r1, int32_t& e0, int32_t& e1) {
  const int2 a = __ldg(reinterpret_cast<const int2*>(p.tile_row) + tile);       // {r0 | long_last << 31, e0}
  const int2 b = __ldg(reinterpret_cast<const int2*>(p.tile_row) + tile + 1);
  r0 = a.x & 0x7fffffff; e0 = a.y;
  r1 = b.x & 0x7fffffff; e1 = b.y;
  if (r0 >= r1) return false;
  if (a.x < 0 && p.has_long) {  // the last owned row is long: leave it to the chunk path
    --r1;
    e1 = __ldg(p.rowptr + r1);
  }
  return r0 < r1;
}

constexpr int kWarpsPerBlock = 8;

// ---- kernel 1a: register-staged tile kernel (all vector widths; used when a lane moves < 16 bytes) ----
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3) spmm_tiles_kernel(const SpmmArgs p) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int GPW = 32 / LPR;
  constexpr int U = (LPR >= 8) ? 8 : LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (LPR - 1);
  const int gbase = lane & ~(LPR - 1);
  const unsigned gmask = group_mask<LPR>(lane);
  const int64_t tile = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * GPW + lane / LPR;
  if (tile >= p.n_tiles) return;  // group-uniform
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  const bool col_active = kcol < p.K;
  const int32_t head = (WMODE == 2 && col_active) ? kcol / (int32_t)(p.K / p.heads) : 0;
  const T* xk = reinterpret_cast<const T*>(p.x) + kcol;
  const int64_t K = p.K;

  int32_t r0, r1, e, e_end;
  if (!tile_ranges(p, tile, r0, r1, e, e_end)) return;
  RowCursor<LPR> rc;
  rc.init(p.rowptr, r0, r1, e, sub, gbase, gmask);

  acc_t acc[VEC];
  int32_t argv[VEC];
  reset_acc<T, VEC, ISMAX>(acc, argv);
  auto next_row = [&]() {
    if (col_active) finish_row<T, VEC, ISMAX>(acc, argv, p, rc.r, kcol, rc.cur_end - rc.row_start);
    reset_acc<T, VEC, ISMAX>(acc, argv);
    rc.advance(p.rowptr, sub, gbase, gmask);
  };
  auto load_batch = [&](int32_t at, int32_t& bcol, int32_t& beid, float& bw) {
    bcol = 0; beid = 0; bw = 1.f;
    const int32_t idx = at + sub;
    if (idx < e_end) {
      bcol = ld_stream_i32(p.col + idx);
      if constexpr (WMODE != 0 || ISMAX) {
        if (WMODE != 0 || p.arg_eid) beid = p.wperm ? ld_stream_i32(p.wperm + idx) : idx;
      }
      if constexpr (WMODE == 1) bw = p.wperm ? __ldg(p.w + beid) : ld_stream_f32(p.w + beid);
    }
  };

  int32_t mycol, myeid; float myw;
  load_batch(e, mycol, myeid, myw);
  while (e < e_end) {
    const int nb = min(LPR, e_end - e);
    int32_t ncol, neid; float nw;
    load_batch(e + LPR, ncol, neid, nw);  // prefetch (predicated off past the tile end)
#pragma unroll 1
    for (int j = 0; j < nb; j += U) {
      Pack<T, VEC> v[U];
      int32_t c[U];
      int32_t ai[U];
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u;
        c[u] = __shfl_sync(gmask, mycol, gbase + (jj & (LPR - 1)));
        if constexpr (ISMAX) ai[u] = p.arg_eid ? __shfl_sync(gmask, myeid, gbase + (jj & (LPR - 1))) : c[u];
        wv[u] = 1.f;
        if constexpr (WMODE == 1) wv[u] = __shfl_sync(gmask, myw, gbase + (jj & (LPR - 1)));
        if constexpr (WMODE == 2) {
          const int32_t eid = __shfl_sync(gmask, myeid, gbase + (jj & (LPR - 1)));
          wv[u] = (jj < nb) ? __ldg(p.w + (int64_t)eid * p.heads + head) : 0.f;
        }
        if (jj < nb && col_active) v[u] = ld_gather<T, VEC>(xk + (int64_t)c[u] * K);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u;
        if (jj < nb) {
          while (e + jj == rc.cur_end) next_row();
          if (col_active) accumulate<T, VEC, WMODE, ISMAX>(acc, argv, v[u], wv[u], ISMAX ? ai[u] : c[u]);
        }
      }
    }
    e += nb;
    mycol = ncol; myeid = neid; myw = nw;
  }
  while (rc.r < rc.r_end) next_row();  // last row of the tile + trailing empty rows
}

// ---- kernel 1c: lean register-staged tile kernel for NARROW aligned fp32 rows (<= 256 bytes) --------------------------------
// tools/probe_tma.cu (profiles/probe_tma_r2_*.jsonl): a random gather of rows up to 256 bytes runs at 36-41 G rows/s with plain
// 16-byte ld.global.nc and many threads in flight, against 18-26 G rows/s through one bulk copy (UBLKCP) per row -- for narrow rows
// the TMA unit's per-copy cost binds, not bytes.  This kernel is that probe with the row bookkeeping added: a group of LPR lanes
// owns one edge tile, every lane reads its own copy of the col / weight stream (broadcast loads inside the group: no shuffles,
// no shared memory, no barriers), U independent 16-byte gathers in flight per lane, serial accumulation in edge order (same
// bits as the ring kernel and the reference loop).  All tiles hold the same number of edges, so the groups of a warp run the
// same trip counts and only part ways at row boundaries.
template <int LPR, int WMODE /* 0 | 1 */, bool ISMAX>
__global__ void __launch_bounds__(256) spmm_lean_kernel(const SpmmArgs p) {
  constexpr int VEC = 4, U = 4;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  if (tile >= p.n_tiles) return;  // group-uniform
  const int sub = threadIdx.x & (LPR - 1);
  const int32_t kcol = sub * VEC;
  const bool col_active = kcol < p.K;
  const float* xk = reinterpret_cast<const float*>(p.x) + kcol;
  const int64_t K = p.K;
  int32_t r0, r1, e, e_end;
  if (!tile_ranges(p, tile, r0, r1, e, e_end)) return;
  int32_t r = r0, row_start = e, row_end = __ldg(p.rowptr + r0 + 1);
  float acc[VEC];
  int32_t argv[VEC];
  reset_acc<float, VEC, ISMAX>(acc, argv);
  auto flush = [&]() {
    if (col_active) finish_row<float, VEC, ISMAX>(acc, argv, p, r, kcol, row_end - row_start);
    reset_acc<float, VEC, ISMAX>(acc, argv);
    row_start = row_end;
    ++r;
    row_end = r < r1 ? __ldg(p.rowptr + r + 1) : INT32_MAX;
  };
  for (; e < e_end; e += U) {
    int32_t c[U], ai[U];
    float wv[U];
    Pack<float, VEC> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = 0; ai[u] = 0; wv[u] = 1.f;
      if (e + u < e_end) {
        c[u] = __ldg(p.col + e + u);
        ai[u] = c[u];
        if constexpr (WMODE == 1 || ISMAX) {
          if (WMODE == 1 || p.arg_eid) {
            const int32_t eid = p.wperm ? __ldg(p.wperm + e + u) : e + u;
            if constexpr (WMODE == 1) wv[u] = __ldg(p.w + eid);
            if constexpr (ISMAX) { if (p.arg_eid) ai[u] = eid; }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u < e_end && col_active) v[u] = ld_gather<float, VEC>(xk + (int64_t)c[u] * K);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (e + u < e_end) {
        while (e + u == row_end) flush();
        if (col_active) accumulate<float, VEC, WMODE, ISMAX>(acc, argv, v[u], wv[u], ai[u]);
      }
    }
  }
  while (r < r1) flush();  // last row of the tile + trailing empty rows
}

int tune_narrow();  // 1 = lean register-staged kernel for aligned fp32 rows <= 256 bytes, 0 = bulk-copy ring (env GGLB200_NARROW), api.cu

// ---- kernel 1b: shared-memory ring tile kernel -------------------------------------------------------------------
// Each lane group owns a ring of STAGES slots x R feature rows in shared memory and keeps it LAG = STAGES-1 slots
// ahead of consumption, so up to LAG*R rows per group are in flight without holding registers.  One lane per row
// issues a bulk async copy (cp.async.bulk.shared::cluster.global -- the TMA-class UBLKCP path: one instruction moves
// the whole row; completion is counted in bytes on one mbarrier per stage).  Consumption is in stream order from
// shared memory (each lane reads its own VB-byte column), with a branch-free fast path for slots inside one row.
//
// Rows that are not 16-byte aligned / sized (bf16 K=100 -> 200 B, fp32 K=47 -> 188 B ...) are fetched as the aligned
// 16-byte SUPERSET [addr & ~15, +CB) with CB = roundup16(row bytes + 15); lanes then read at offset (addr & 15).  The
// last one or two rows of x, whose superset would cross the end of the tensor, are not staged: their lanes read
// global memory directly at consume time (`a0 > lim`).
//
// Blocks are ONE warp: the hardware block scheduler is the dynamic tile scheduler -- a finished or empty tile
// frees its ring at once, so skewed graphs do not park shared memory behind one long tile (RMAT 13.1 -> 8.4 ms).
// Groups [0, ws_chunks) of the launch reduce long-row chunks into partials, the remaining groups are edge tiles.
template <int LPR> struct RingCfg {
  static constexpr int R = LPR >= 8 ? 8 : LPR;                      // rows per slot
  static constexpr int STAGES = 2;                                  // ring depth of the SDDMM kernel; (STAGES-1)*R <= 2*LPR (index batches A,B,C)
  static constexpr int LAG = STAGES - 1;
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> lds_pack(uint32_t addr) {
  Pack<T, VEC> v;
  constexpr int B = sizeof(T) * VEC;
  uint32_t* q = reinterpret_cast<uint32_t*>(&v);
  if constexpr (B == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(addr));
  else if constexpr (B == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(q[0]), "=r"(q[1]) : "r"(addr));
  else if constexpr (B == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q[0]) : "r"(addr));
  else {
    uint16_t h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    *reinterpret_cast<uint16_t*>(&v) = h;
  }
  return v;
}

template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__global__ void __launch_bounds__(32) spmm_ring_kernel(const SpmmArgs p, const int32_t pitch /* bytes per row slot */,
                                                       const uint64_t x_end_al /* align_down(x + n_src*K*sizeof(T), 16) */,
                                                       const int32_t STAGES /* ring depth: 2 or 3 */) {
  using acc_t = typename Elem<T>::acc_t;
  using Cfg = RingCfg<LPR>;
  constexpr int VB = (int)sizeof(T) * VEC;   // bytes a lane owns per row
  constexpr bool ALIGNED = VB == 16;         // rows are 16-byte aligned and sized: no superset, no tail rows
  constexpr int GPW = 32 / LPR, R = Cfg::R;
  const int32_t LAG = STAGES - 1;
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPR - 1);
  const int gbase = lane & ~(LPR - 1);
  const unsigned gmask = group_mask<LPR>(lane);
  const int gib = lane / LPR;  // group index in the (one-warp) block
  const int64_t g = (int64_t)blockIdx.x * GPW + gib;
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  const bool col_active = kcol < p.K;
  const int32_t head = ((WMODE == 2 || WMODE == 3) && col_active) ? kcol / (int32_t)(p.K / p.heads) : 0;
  // WMODE 3: the lane's VEC elements may reach into the next head (head width not a multiple of VEC): elements [0, nlo) belong
  // to `head`, the rest to `head_hi`
  int32_t nlo = VEC, head_hi = head;
  if constexpr (WMODE == 3) {
    const int32_t Cw = (int32_t)(p.K / p.heads);
    nlo = min(VEC, (head + 1) * Cw - kcol);
    head_hi = min(head + 1, p.heads - 1);
  }
  const int64_t K = p.K;
  const T* xt = reinterpret_cast<const T*>(p.x) + (int64_t)blockIdx.y * LPR * VEC;  // first column of this column tile
  // bytes of one row inside this column tile, and the bytes copied per row
  const uint32_t rowb = (uint32_t)min((int64_t)LPR * VB, K * (int64_t)sizeof(T) - (int64_t)blockIdx.y * LPR * VB);
  const uint32_t cb = ALIGNED ? rowb : ((rowb + 15u + 15u) & ~15u);
  const uint64_t lim = x_end_al - cb;  // a superset copy starting above `lim` would cross the end of x

  // shared memory: data rings of all groups, then STAGES mbarriers per group
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(ring_smem);
  const uint32_t ring = smem0 + (uint32_t)(gib * (STAGES * R)) * pitch;
  const uint32_t bars = smem0 + (uint32_t)(GPW * STAGES * R) * pitch + (uint32_t)gib * STAGES * 8;
  if (sub == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(bars + i * 8, ALIGNED ? 1 : R);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp(gmask);

  // ---- what does this group stream? ----
  int32_t r0 = 0, r1 = 0, e0, e_end;
  int64_t chunk = -1;
  if (g < p.ws_chunks) {  // long-row chunk -> partial
    if (g >= min(__ldg(p.plan + 0), p.ws_chunks)) return;
    chunk = g;
    const int32_t* chunk_tab = p.plan + plan_chunk_off();
    const int32_t row = __ldg(chunk_tab + 2 * g);
    e0 = __ldg(chunk_tab + 2 * g + 1);
    e_end = min(e0 + GGLB200_LONG_CHUNK, __ldg(p.rowptr + row + 1));
  } else {
    const int64_t tile = g - p.ws_chunks;
    if (tile >= p.n_tiles) return;
    if (!tile_ranges(p, tile, r0, r1, e0, e_end)) return;
  }
  RowCursor<LPR> rc;
  rc.init(p.rowptr, r0, r1, e0, sub, gbase, gmask);  // chunk: r0 == r1 -> cur_end = INT32_MAX (no boundaries)

  acc_t acc[VEC];
  int32_t argv[VEC];
  reset_acc<T, VEC, ISMAX>(acc, argv);
  auto next_row = [&]() {
    if (col_active) finish_row<T, VEC, ISMAX>(acc, argv, p, rc.r, kcol, rc.cur_end - rc.row_start);
    reset_acc<T, VEC, ISMAX>(acc, argv);
    rc.advance(p.rowptr, sub, gbase, gmask);
  };

  const int32_t n = e_end - e0;                 // edges in the stream
  const int32_t nslots = (n + R - 1) / R;

  // index batches of LPR stream positions: A = batch bA (being consumed), B = bA+1, C = bA+2.  The weight gather
  // w[perm[i]] depends on the perm load, so it is issued one batch AFTER its index (when the batch moves from C to
  // B): an in-order warp never waits on the perm -> w chain.
  struct Batch { int32_t col, eid; float w; };
  auto load_index = [&](int32_t b) {
    Batch t; t.col = 0; t.eid = 0; t.w = 1.f;
    const int32_t pos = b * LPR + sub;
    if (pos < n) {
      const int32_t idx = e0 + pos;
      t.col = ld_stream_i32(p.col + idx);
      if constexpr (WMODE != 0 || ISMAX) {
        if (WMODE != 0 || p.arg_eid) t.eid = p.wperm ? ld_stream_i32(p.wperm + idx) : idx;
      }
    }
    return t;
  };
  auto load_weight = [&](Batch& t, int32_t b) {
    if constexpr (WMODE == 1) {
      if (b * LPR + sub < n) t.w = p.wperm ? __ldg(p.w + t.eid) : ld_stream_f32(p.w + t.eid);
    }
  };
  Batch A = load_index(0), B = load_index(1), C = load_index(2);
  int32_t bA = 0;

  int32_t istage = 0;  // ring stage of the next slot to issue (issue() is called once per slot, in order)
  auto issue = [&](int32_t s) {  // start the copies of slot s
    const int32_t stg = istage;
    istage = istage + 1 == STAGES ? 0 : istage + 1;
    if (s < nslots) {  // group-uniform
      const int32_t pos0 = s * R;
      const int32_t d = pos0 / LPR - bA;  // 0, 1 or 2
      const int32_t colsrc = d == 0 ? A.col : (d == 1 ? B.col : C.col);
      const uint32_t dst = ring + (uint32_t)(stg * R) * pitch;
      const uint32_t bar = bars + (uint32_t)stg * 8;
      const int32_t nv = min(R, n - pos0);
      const int32_t c = __shfl_sync(gmask, colsrc, gbase + ((pos0 + sub) & (LPR - 1)));  // lane u copies row u
      __syncwarp(gmask);  // every lane of the group has finished reading this stage (slot s - STAGES)
      if constexpr (ALIGNED) {
        if (sub == 0) mbar_arrive_expect_tx(bar, (uint32_t)nv * rowb);
        if (sub < nv) {
          const T* rowp;
          if (p.x_table) {  // 1-D node partition: the row lives in the HBM of GPU `owner`; fetched over NVLink by the same bulk copy
            const int32_t owner = c >> p.owner_shift, local = c & ((1 << p.owner_shift) - 1);
            rowp = reinterpret_cast<const T*>(__ldg(reinterpret_cast<const unsigned long long*>(p.x_table) + owner)) +
                   (int64_t)blockIdx.y * LPR * VEC + (int64_t)local * K;
          } else {
            rowp = xt + (int64_t)c * K;
          }
          bulk_copy_g2s(dst + (uint32_t)sub * pitch, rowp, rowb, bar);
        }
      } else if (sub < R) {
        const uint64_t a0 = (uint64_t)(xt + (int64_t)c * K) & ~(uint64_t)15;
        if (sub < nv && a0 <= lim) {
          mbar_arrive_expect_tx(bar, cb);
          bulk_copy_g2s(dst + (uint32_t)sub * pitch, (const void*)a0, cb, bar);
        } else {
          mbar_arrive(bar);
        }
      }
    }
  };

  for (int s = 0; s < LAG; ++s) issue(s);   // the first copies need only A.col / B.col
  load_weight(A, 0);
  load_weight(B, 1);
  int32_t cstage = 0;
  uint32_t cparity = 0;
  for (int32_t s = 0; s < nslots; ++s) {
    const int32_t pos0 = s * R;
    if (pos0 / LPR != bA) {  // the consumer enters the next index batch
      A = B; B = C; ++bA;
      load_weight(B, bA + 1);
      C = load_index(bA + 2);
    }
    issue(s + LAG);
    // per-edge scalars of this slot, fetched in convergent code
    float wv[R];
    float wv_hi[WMODE == 3 ? R : 1];
    int32_t cc[R];
    int32_t ai[R];   // MAX: what `arg` records for the edge (source id, or the original edge id with arg_eid)
#pragma unroll
    for (int u = 0; u < R; ++u) {
      wv[u] = 1.f; cc[u] = 0;
      if constexpr (ISMAX || !ALIGNED) cc[u] = __shfl_sync(gmask, A.col, gbase + ((pos0 + u) & (LPR - 1)));
      if constexpr (ISMAX) ai[u] = p.arg_eid ? __shfl_sync(gmask, A.eid, gbase + ((pos0 + u) & (LPR - 1))) : cc[u];
      if constexpr (WMODE == 1) wv[u] = __shfl_sync(gmask, A.w, gbase + ((pos0 + u) & (LPR - 1)));
      if constexpr (WMODE == 2 || WMODE == 3) {
        const int32_t eid = __shfl_sync(gmask, A.eid, gbase + ((pos0 + u) & (LPR - 1)));
        wv[u] = __ldg(p.w + (int64_t)eid * p.heads + head);
        if constexpr (WMODE == 3) wv_hi[u] = __ldg(p.w + (int64_t)eid * p.heads + head_hi);
      }
    }
    auto acc_edge = [&](int u, const Pack<T, VEC>& v) {
      if constexpr (WMODE == 3) accumulate<T, VEC, WMODE, ISMAX>(acc, argv, v, wv[u], ISMAX ? ai[u] : cc[u], wv_hi[u], nlo);
      else accumulate<T, VEC, WMODE, ISMAX>(acc, argv, v, wv[u], ISMAX ? ai[u] : cc[u]);
    };
    mbar_wait(bars + (uint32_t)cstage * 8, cparity);
    const uint32_t src = ring + (uint32_t)(cstage * R) * pitch + (uint32_t)sub * VB;
    if (++cstage == STAGES) { cstage = 0; cparity ^= 1u; }
    // the lane's VB bytes of the row at stream position pos0+u
    auto fetch = [&](int u) {
      if constexpr (ALIGNED) {
        return lds_pack<T, VEC>(src + u * pitch);
      } else {
        const T* rp = xt + (int64_t)cc[u] * K;
        const uint64_t a = (uint64_t)rp;
        if ((a & ~(uint64_t)15) <= lim) return lds_pack<T, VEC>(src + u * pitch + (uint32_t)(a & 15));
        return ld_gather<T, VEC>(rp + sub * VEC);  // tail row of x: not staged
      }
    };
    if (pos0 + R <= n && e0 + pos0 + R <= rc.cur_end) {
      // fast path: the whole slot lies inside the current row
      if (col_active) {
#pragma unroll
        for (int u = 0; u < R; ++u) acc_edge(u, fetch(u));
      }
    } else {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int32_t pos = pos0 + u;
        if (pos < n) {  // group-uniform
          while (e0 + pos == rc.cur_end) next_row();
          if (col_active) acc_edge(u, fetch(u));
        }
      }
    }
  }
  if (chunk >= 0) {
    if (col_active) {
      acc_t* pv = reinterpret_cast<acc_t*>(p.part_val) + chunk * K + kcol;
#pragma unroll
      for (int i = 0; i < VEC; ++i) pv[i] = acc[i];
      if constexpr (ISMAX) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p.part_arg[chunk * K + kcol + i] = argv[i];
      }
    }
  } else {
    while (rc.r < rc.r_end) next_row();  // last row of the tile + trailing empty rows
  }
}

// ---- kernel 2: one group per chunk of a long row -> partials (register-staged configurations) -----------
template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spmm_chunks_kernel(const SpmmArgs p) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (LPR - 1);
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  const bool col_active = kcol < p.K;
  const int32_t n_chunks = min(__ldg(p.plan + 0), p.ws_chunks);
  const int32_t* chunk_tab = p.plan + plan_chunk_off();
  const int32_t head = (WMODE == 2 && col_active) ? kcol / (int32_t)(p.K / p.heads) : 0;
  const T* xk = reinterpret_cast<const T*>(p.x) + kcol;
  const int64_t groups = (int64_t)gridDim.x * kWarpsPerBlock * RPW;
  for (int64_t c = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * RPW + lane / LPR; c < n_chunks; c += groups) {
    const int32_t row = __ldg(chunk_tab + 2 * c), first = __ldg(chunk_tab + 2 * c + 1);
    const int32_t end = min(first + GGLB200_LONG_CHUNK, __ldg(p.rowptr + row + 1));
    acc_t acc[VEC];
    int32_t argv[VEC];
    reset_acc<T, VEC, ISMAX>(acc, argv);
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
  const int32_t* long_tab = p.plan + plan_long_off(max_chunks);
  const int64_t groups = (int64_t)gridDim.x * kWarpsPerBlock * RPW;
  for (int64_t l = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * RPW + lane / LPR; l < n_long; l += groups) {
    const int32_t row = __ldg(long_tab + 3 * l), c0 = __ldg(long_tab + 3 * l + 1), nch = __ldg(long_tab + 3 * l + 2);
    if (c0 + nch > p.ws_chunks) continue;  // workspace contract violated by the caller: stay memory-safe
    acc_t acc[VEC];
    int32_t argv[VEC];
    reset_acc<T, VEC, ISMAX>(acc, argv);
    // A hub row has hundreds of chunks: fetch the partials of FB chunks at once (independent loads in flight), then
    // combine them in ascending chunk order -- the same order and rounding as a plain sequential loop.
    constexpr int FB = 8;
    for (int32_t j0 = 0; j0 < nch; j0 += FB) {
      acc_t v[FB][VEC];
#pragma unroll
      for (int u = 0; u < FB; ++u) {
        if (j0 + u < nch) {
          const acc_t* pv = reinterpret_cast<const acc_t*>(p.part_val) + (int64_t)(c0 + j0 + u) * p.K + kcol;
#pragma unroll
          for (int i = 0; i < VEC; ++i) v[u][i] = pv[i];
        }
      }
#pragma unroll
      for (int u = 0; u < FB; ++u) {
        if (j0 + u < nch) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if constexpr (ISMAX) {
              if (acc[i] < v[u][i]) { acc[i] = v[u][i]; argv[i] = p.part_arg[(int64_t)(c0 + j0 + u) * p.K + kcol + i]; }
            } else {
              acc[i] = add_rn(acc[i], v[u][i]);
            }
          }
        }
      }
    }
    const int32_t deg = __ldg(p.rowptr + row + 1) - __ldg(p.rowptr + row);
    finish_row<T, VEC, ISMAX>(acc, argv, p, row, kcol, deg);
  }
}

// ---- host-side launch ------------------------------------------------------------------------------
int tune_ring();         // tuning knob (env GGLB200_RING = bulk (default) | none -> register-staged kernels), api.cu
int tune_ring_stages();  // tuning knob (env GGLB200_RING_STAGES = 2 | 3; 0 = heuristic), api.cu

template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
int launch_ring(const SpmmArgs& p, int tiles) {
  constexpr int GPW = 32 / LPR;
  constexpr int VB = (int)sizeof(T) * VEC;
  using Cfg = RingCfg<LPR>;
  const int64_t rowbytes = p.K * (int64_t)sizeof(T);
  const int64_t tileb = rowbytes < (int64_t)LPR * VB ? rowbytes : (int64_t)LPR * VB;   // widest column tile
  const int32_t pitch = (int32_t)(VB == 16 ? tileb : ((tileb + 15 + 15) & ~(int64_t)15));
  // ring depth: 2 slots of 8 rows per group -- more resident one-warp blocks beat a deeper ring (C2/K=128: 2.69 vs
  // 2.96 ms, C3 RMAT 6.97 vs 7.13 ms, GAT forward 7.6 vs 9.7 ms; only the DRAM-bound uniform C3 graph is 1 % better at 3)
  const int32_t stages = tune_ring_stages() ? tune_ring_stages() : 2;
  const size_t smem = (size_t)GPW * stages * Cfg::R * pitch + (size_t)GPW * stages * 8;
  GGL_REQUIRE(smem <= 200 * 1024, GGL_EUNSUPPORTED, "spmm: ring of %zu bytes does not fit shared memory", smem);
  // function attributes are per device: set on every launch (a few hundred ns) so any current device works
  GGL_CUDA(cudaFuncSetAttribute(spmm_ring_kernel<T, VEC, LPR, WMODE, ISMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t groups = p.n_tiles + (p.has_long ? p.ws_chunks : 0);
  const int64_t blocks = (groups + GPW - 1) / GPW;
  GGL_REQUIRE(blocks < INT32_MAX, GGL_EUNSUPPORTED, "spmm: grid too large");
  SpmmArgs q = p;
  if (!p.has_long) q.ws_chunks = 0;
  const uint64_t x_end_al = ((uint64_t)p.x + (uint64_t)p.n_src * (uint64_t)rowbytes) & ~(uint64_t)15;
  spmm_ring_kernel<T, VEC, LPR, WMODE, ISMAX><<<dim3((unsigned)blocks, tiles), 32, smem, p.stream>>>(q, pitch, x_end_al, stages);
  GGL_LAUNCHED("spmm_ring");
  if (p.has_long) {
    spmm_finalize_kernel<T, VEC, LPR, ISMAX><<<dim3(sm_count(), tiles), kWarpsPerBlock * 32, 0, p.stream>>>(q);
    GGL_LAUNCHED("spmm_finalize");
  }
  return GGL_OK;
}

template <typename T, int VEC, int LPR, int WMODE, bool ISMAX>
int launch_config(const SpmmArgs& p) {
  constexpr int GPW = 32 / LPR;
  const int tiles = (int)((p.K + (int64_t)LPR * VEC - 1) / ((int64_t)LPR * VEC));
  GGL_REQUIRE(tiles <= 65535, GGL_EUNSUPPORTED, "spmm: K=%lld needs too many column tiles", (long long)p.K);
  // floating rows of at least 4 lanes stream through the shared-memory ring; the rest (K <= 4 elements per row,
  // f64, integers) through the register-staged kernels
  if constexpr (std::is_same<T, float>::value && VEC == 4 && LPR <= 16 && (WMODE == 0 || WMODE == 1)) {
    if (tune_narrow() != 0 && p.x_table == nullptr && tiles == 1 && (((uintptr_t)p.x | (uintptr_t)p.out) & 15) == 0) {
      const int64_t threads = p.n_tiles * LPR;
      const int64_t blocks = (threads + 255) / 256;
      GGL_REQUIRE(blocks < INT32_MAX, GGL_EUNSUPPORTED, "spmm: grid too large");
      spmm_lean_kernel<LPR, WMODE, ISMAX><<<(unsigned)blocks, 256, 0, p.stream>>>(p);
      GGL_LAUNCHED("spmm_lean");
      if (p.has_long) {   // rows above GGLB200_LONG_ROW: chunk partials + ordered combine, as on the register-staged path
        spmm_chunks_kernel<T, VEC, LPR, WMODE, ISMAX><<<dim3(sm_count() * 2, tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
        GGL_LAUNCHED("spmm_chunks");
        spmm_finalize_kernel<T, VEC, LPR, ISMAX><<<dim3(sm_count(), tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
        GGL_LAUNCHED("spmm_finalize");
      }
      return GGL_OK;
    }
  }
  if constexpr (WMODE == 3) {
    // spanning lanes exist on the bulk-copy ring only (fp32, 16-byte lanes)
    if constexpr (Elem<T>::is_float && sizeof(T) == 4 && VEC == 4 && LPR >= 4) return launch_ring<T, VEC, LPR, WMODE, ISMAX>(p, tiles);
    else return set_error(GGL_EUNSUPPORTED, "spmm: head width not a multiple of the lane width is fp32 / 16-byte lanes only");
  } else {
  if constexpr (Elem<T>::is_float && sizeof(T) <= 4 && LPR >= 4) {
    if (tune_ring() != 0 || p.x_table) {
      GGL_REQUIRE(p.x_table == nullptr || (sizeof(T) * VEC == 16 && !ISMAX), GGL_EUNSUPPORTED,
                  "spmm: peer-memory gather needs 16-byte aligned rows (K*sizeof(T) %% 16 == 0) and sum/mean");
      return launch_ring<T, VEC, LPR, WMODE, ISMAX>(p, tiles);
    }
  }
  GGL_REQUIRE(p.x_table == nullptr, GGL_EUNSUPPORTED, "spmm: peer-memory gather is not available for this dtype / row width");
  const int64_t groups_per_block = (int64_t)kWarpsPerBlock * GPW;
  const int64_t blocks = (p.n_tiles + groups_per_block - 1) / groups_per_block;
  GGL_REQUIRE(blocks < INT32_MAX, GGL_EUNSUPPORTED, "spmm: grid too large");
  spmm_tiles_kernel<T, VEC, LPR, WMODE, ISMAX><<<dim3((unsigned)blocks, tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
  GGL_LAUNCHED("spmm_tiles");
  if (p.has_long) {
    static int occ_chunks = 0;  // resident blocks per SM of the chunk kernel (persistent grid)
    if (occ_chunks == 0) {
      int o = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, spmm_chunks_kernel<T, VEC, LPR, WMODE, ISMAX>, kWarpsPerBlock * 32, 0) != cudaSuccess || o < 1) o = 2;
      occ_chunks = o;
    }
    spmm_chunks_kernel<T, VEC, LPR, WMODE, ISMAX><<<dim3(sm_count() * occ_chunks, tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
    GGL_LAUNCHED("spmm_chunks");
    spmm_finalize_kernel<T, VEC, LPR, ISMAX><<<dim3(sm_count(), tiles), kWarpsPerBlock * 32, 0, p.stream>>>(p);
    GGL_LAUNCHED("spmm_finalize");
  }
  return GGL_OK;
  }  // WMODE != 3
}

template <typename T, int VEC, int LPR>
int launch_modes(const SpmmArgs& p, bool allow_weight) {
  int wmode = p.w == nullptr ? 0 : (p.heads > 1 ? 2 : 1);
  if (wmode == 2 && (p.K / p.heads) % VEC != 0) wmode = 3;   // lanes span two heads (the caller guarantees C >= VEC)
  const bool ismax = p.reduce == GGL_MAX;
  if (wmode == 0) return ismax ? launch_config<T, VEC, LPR, 0, true>(p) : launch_config<T, VEC, LPR, 0, false>(p);
  if constexpr (Elem<T>::is_float) {
    if (allow_weight) {
      if (wmode == 1) return ismax ? launch_config<T, VEC, LPR, 1, true>(p) : launch_config<T, VEC, LPR, 1, false>(p);
      if (wmode == 3) {
        if constexpr (sizeof(T) == 4 && VEC == 4 && LPR >= 4)
          return ismax ? launch_config<T, VEC, LPR, 3, true>(p) : launch_config<T, VEC, LPR, 3, false>(p);
        else return set_error(GGL_EUNSUPPORTED, "spmm: head width %lld is not a multiple of the lane width", (long long)(p.K / p.heads));
      }
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

// vector-width dispatch shared by the 16-bit floating types
template <typename T>
int launch_16bit(const SpmmArgs& p) {
  const int64_t K = p.K;
  const int64_t C = p.heads > 1 ? K / p.heads : K;
  const uintptr_t al = (uintptr_t)p.x | (uintptr_t)p.out;
  int vec = 1;
  if (K > 4) {
    if (K % 8 == 0 && C % 8 == 0 && al % 16 == 0) vec = 8;
    else if (K % 4 == 0 && C % 4 == 0 && al % 8 == 0) vec = 4;
    else if (K % 2 == 0 && C % 2 == 0 && al % 4 == 0) vec = 2;
  }
  const int lpr = lpr_for(K, vec);
  if (vec == 8) {
    if (lpr <= 4) return launch_modes<T, 8, 4>(p, true);
    if (lpr == 8) return launch_modes<T, 8, 8>(p, true);
    if (lpr == 16) return launch_modes<T, 8, 16>(p, true);
    return launch_modes<T, 8, 32>(p, true);
  }
  if (vec == 4) return launch_modes<T, 4, 32>(p, true);
  if (vec == 2) return launch_modes<T, 2, 32>(p, true);
  if (lpr == 1) return launch_modes<T, 1, 1>(p, true);
  if (lpr <= 4) return launch_modes<T, 1, 4>(p, true);
  return launch_modes<T, 1, 32>(p, true);
}

// entry points per element type (explicitly instantiated in spmm_f32.cu / spmm_bf16.cu / spmm_f16.cu / spmm_misc.cu)
template <typename T> int launch_spmm(const SpmmArgs& p);

}  // namespace ggl
