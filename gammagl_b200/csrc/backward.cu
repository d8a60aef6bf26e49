// backward.cu -- backward helpers that are not an SpMM on the transposed CSR.
//
//   segment_gather_bwd   SegmentSum::backward  (src/segment_sum.cpp:43-54):  gx[e] = g[ids[e]]
//                        SegmentMean::backward (src/segment_mean.cpp:44-63): gx[e] = g[ids[e]] / bincount(ids)[ids[e]]
//   segment_max_bwd      SegmentMax::backward  (src/segment_max.cpp:48-61):  gx[arg[r,k],k] = g[r,k]
//   scale_rows_by_count  first factor of spmm_mean_cpu_backward (cpu/spmm_mean_cpu.cpp:96-98): g[d,k] / indeg[d]
//   spmm_max_bwd         spmm_max_cpu_backward (cpu/spmm_max_cpu.cpp:84-96) on the transposed CSR
// (paths relative to /root/reference/gammagl/mpops/torch_ext)
#include "common.cuh"

namespace ggl {

static int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Row gather, dtype-agnostic: rows are moved as W-byte words.  A group of LPR lanes (power of two >= words per row)
// copies one edge row, 32/LPR rows per warp step, so narrow rows (K=32 fp32 = 8 uint4 words) still use every lane.
template <typename W, int LPR>
__global__ void __launch_bounds__(256) gather_rows_kernel(const W* __restrict__ g, const int64_t* __restrict__ ids,
                                                          W* __restrict__ gx, int64_t E, int32_t words) {
  constexpr int GPW = 32 / LPR;  // rows per warp step
  constexpr int U = 4;           // independent index -> row chains in flight per group
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1), grp = lane / LPR;
  const int64_t span = (int64_t)gridDim.x * (blockDim.x >> 5) * (GPW * U);
  for (int64_t e0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (GPW * U); e0 < E; e0 += span) {
    int64_t r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * GPW + grp;
      r[u] = e < E ? __ldcs(ids + e) : (int64_t)-1;
    }
    for (int32_t k = sub; k < words; k += LPR) {
      W v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r[u] >= 0) v[u] = __ldg(g + r[u] * words + k);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r[u] >= 0) __stcs(gx + (e0 + u * GPW + grp) * words + k, v[u]);
    }
  }
}

template <typename W>
static void launch_gather_rows(const void* g, const int64_t* ids, void* gx, int64_t E, int64_t words, cudaStream_t stream) {
  int lpr = 1;
  while (lpr < 32 && lpr < words) lpr <<= 1;
  const int64_t warps = (E * lpr + 127) / 128;  // 4 rows per group per iteration
  int64_t grid = (warps + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const W* gp = (const W*)g;
  W* op = (W*)gx;
  switch (lpr) {
    case 1: gather_rows_kernel<W, 1><<<(int)grid, 256, 0, stream>>>(gp, ids, op, E, (int32_t)words); break;
    case 2: gather_rows_kernel<W, 2><<<(int)grid, 256, 0, stream>>>(gp, ids, op, E, (int32_t)words); break;
    case 4: gather_rows_kernel<W, 4><<<(int)grid, 256, 0, stream>>>(gp, ids, op, E, (int32_t)words); break;
    case 8: gather_rows_kernel<W, 8><<<(int)grid, 256, 0, stream>>>(gp, ids, op, E, (int32_t)words); break;
    case 16: gather_rows_kernel<W, 16><<<(int)grid, 256, 0, stream>>>(gp, ids, op, E, (int32_t)words); break;
    default: gather_rows_kernel<W, 32><<<(int)grid, 256, 0, stream>>>(gp, ids, op, E, (int32_t)words); break;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_mean_kernel(const T* __restrict__ g, const int64_t* __restrict__ ids,
                                                               const int32_t* __restrict__ rowptr, T* __restrict__ gx,
                                                               int64_t E, int32_t K) {
  using acc_t = typename Elem<T>::acc_t;
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < E; e += warps) {
    const int64_t r = __ldg(ids + e);
    const acc_t cnt = (acc_t)(__ldg(rowptr + r + 1) - __ldg(rowptr + r));
    const T* src = g + r * K;
    T* dst = gx + e * K;
    for (int32_t k = lane; k < K; k += 32) dst[k] = Elem<T>::from_acc(Elem<T>::to_acc(__ldg(src + k)) / cnt);
  }
}
template <> __global__ void gather_rows_mean_kernel<__nv_bfloat16>(const __nv_bfloat16* __restrict__ g, const int64_t* __restrict__ ids,
                                                                   const int32_t* __restrict__ rowptr, __nv_bfloat16* __restrict__ gx,
                                                                   int64_t E, int32_t K) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < E; e += warps) {
    const int64_t r = ids[e];
    const float cnt = (float)(rowptr[r + 1] - rowptr[r]);
    for (int32_t k = lane; k < K; k += 32) gx[e * K + k] = __float2bfloat16_rn(__bfloat162float(g[r * K + k]) / cnt);
  }
}
template <> __global__ void gather_rows_mean_kernel<__half>(const __half* __restrict__ g, const int64_t* __restrict__ ids,
                                                            const int32_t* __restrict__ rowptr, __half* __restrict__ gx,
                                                            int64_t E, int32_t K) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < E; e += warps) {
    const int64_t r = ids[e];
    const float cnt = (float)(rowptr[r + 1] - rowptr[r]);
    for (int32_t k = lane; k < K; k += 32) gx[e * K + k] = __float2half_rn(__half2float(g[r * K + k]) / cnt);
  }
}

// ---- row expansion through the CSR: the same backward as gather_rows, in scatter form ------------------------------------
// Sorted position p of the segment CSR belongs to row(p) and to message perm[p]:  gx[perm[p]] = g[row(p)].  Reading g in row
// order touches every gradient row once (N*K*s bytes instead of E*K*s gathered bytes); the writes are whole rows.  One warp
// per fixed tile of TILE_EDGES sorted positions (balanced whatever the degree distribution); row(p) is a binary search of
// rowptr between the tile-table rows of this tile and the next one (a handful of L1-resident steps).
__device__ __forceinline__ int32_t row_of_position(const int32_t* __restrict__ rowptr, int32_t lo, int32_t hi, int32_t p) {
  while (hi - lo > 1) {  // invariant: rowptr[lo] <= p, answer in [lo, hi)
    const int32_t m = (lo + hi) >> 1;
    if (__ldg(rowptr + m) <= p) lo = m; else hi = m;
  }
  return lo;
}

struct ExpandTile {
  int32_t p0, p1, lo, hi;
};
__device__ __forceinline__ ExpandTile expand_tile(const int2* __restrict__ tile_tab, int64_t t, int32_t te, int64_t E) {
  ExpandTile x;
  const int64_t b = t * te, e = b + te < E ? b + te : E;
  x.p0 = (int32_t)b; x.p1 = (int32_t)e;
  x.lo = (__ldg(&tile_tab[t].x) & 0x7fffffff) - 1;   // the row that straddles the tile start begins before r0(t)
  if (x.lo < 0) x.lo = 0;
  x.hi = __ldg(&tile_tab[t + 1].x) & 0x7fffffff;
  return x;
}

template <typename W, int LPR>
__global__ void __launch_bounds__(256) expand_rows_kernel(const W* __restrict__ g, const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ perm, const int2* __restrict__ tile_tab,
                                                          W* __restrict__ gx, int64_t E, int32_t words, int32_t te, int64_t n_tiles) {
  constexpr int GPW = 32 / LPR;
  constexpr int U = 4;
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1), grp = lane / LPR;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += warps) {
    const ExpandTile x = expand_tile(tile_tab, t, te, E);
    for (int32_t pb = x.p0; pb < x.p1; pb += GPW * U) {
      int32_t e[U], r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t p = pb + u * GPW + grp;
        e[u] = p < x.p1 ? __ldcs(perm + p) : -1;
        r[u] = p < x.p1 ? row_of_position(rowptr, x.lo, x.hi, p) : 0;
      }
      for (int32_t k = sub; k < words; k += LPR) {
        W v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e[u] >= 0) v[u] = __ldg(g + (int64_t)r[u] * words + k);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e[u] >= 0) __stcs(gx + (int64_t)e[u] * words + k, v[u]);
      }
    }
  }
}

// mean: gx[perm[p]] = g[row(p)] / count(row(p)), element-wise in the accumulation type like gather_rows_mean_kernel
template <typename T, int LPR>
__global__ void __launch_bounds__(256) expand_rows_mean_kernel(const T* __restrict__ g, const int32_t* __restrict__ rowptr,
                                                               const int32_t* __restrict__ perm, const int2* __restrict__ tile_tab,
                                                               T* __restrict__ gx, int64_t E, int32_t K, int32_t te, int64_t n_tiles) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1), grp = lane / LPR;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += warps) {
    const ExpandTile x = expand_tile(tile_tab, t, te, E);
    for (int32_t p = x.p0 + grp; p < x.p1; p += GPW) {
      const int64_t e = __ldcs(perm + p);
      const int64_t r = row_of_position(rowptr, x.lo, x.hi, p);
      const acc_t cnt = (acc_t)(__ldg(rowptr + r + 1) - __ldg(rowptr + r));
      for (int32_t k = sub; k < K; k += LPR) gx[e * K + k] = Elem<T>::from_acc(Elem<T>::to_acc(g[r * K + k]) / cnt);
    }
  }
}

// fp32 rows that are whole float4 words: the mean forms of the two kernels above at full vector width
template <int LPR>
__global__ void __launch_bounds__(256) gather_rows_mean4_kernel(const float4* __restrict__ g, const int64_t* __restrict__ ids,
                                                                const int32_t* __restrict__ rowptr, float4* __restrict__ gx,
                                                                int64_t E, int32_t words) {
  constexpr int GPW = 32 / LPR;
  constexpr int U = 4;
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1), grp = lane / LPR;
  const int64_t span = (int64_t)gridDim.x * (blockDim.x >> 5) * (GPW * U);
  for (int64_t e0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (GPW * U); e0 < E; e0 += span) {
    int64_t r[U];
    float c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * GPW + grp;
      r[u] = e < E ? __ldcs(ids + e) : (int64_t)-1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = r[u] >= 0 ? (float)(__ldg(rowptr + r[u] + 1) - __ldg(rowptr + r[u])) : 1.f;
    for (int32_t k = sub; k < words; k += LPR) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r[u] >= 0) v[u] = __ldg(g + r[u] * words + k);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r[u] >= 0) {
          float4 o;
          o.x = __fdiv_rn(v[u].x, c[u]); o.y = __fdiv_rn(v[u].y, c[u]); o.z = __fdiv_rn(v[u].z, c[u]); o.w = __fdiv_rn(v[u].w, c[u]);
          __stcs(gx + (e0 + u * GPW + grp) * words + k, o);
        }
    }
  }
}

template <int LPR>
__global__ void __launch_bounds__(256) expand_rows_mean4_kernel(const float4* __restrict__ g, const int32_t* __restrict__ rowptr,
                                                                const int32_t* __restrict__ perm, const int2* __restrict__ tile_tab,
                                                                float4* __restrict__ gx, int64_t E, int32_t words, int32_t te,
                                                                int64_t n_tiles) {
  constexpr int GPW = 32 / LPR;
  constexpr int U = 4;
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1), grp = lane / LPR;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += warps) {
    const ExpandTile x = expand_tile(tile_tab, t, te, E);
    for (int32_t pb = x.p0; pb < x.p1; pb += GPW * U) {
      int32_t e[U], r[U];
      float c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t p = pb + u * GPW + grp;
        e[u] = p < x.p1 ? __ldcs(perm + p) : -1;
        r[u] = p < x.p1 ? row_of_position(rowptr, x.lo, x.hi, p) : 0;
        c[u] = (float)(__ldg(rowptr + r[u] + 1) - __ldg(rowptr + r[u]));
      }
      for (int32_t k = sub; k < words; k += LPR) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e[u] >= 0) v[u] = __ldg(g + (int64_t)r[u] * words + k);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e[u] >= 0) {
            float4 o;
            o.x = __fdiv_rn(v[u].x, c[u]); o.y = __fdiv_rn(v[u].y, c[u]); o.z = __fdiv_rn(v[u].z, c[u]); o.w = __fdiv_rn(v[u].w, c[u]);
            __stcs(gx + (int64_t)e[u] * words + k, o);
          }
      }
    }
  }
}

static int lanes_for(int64_t n) {
  int lpr = 1;
  while (lpr < 32 && lpr < n) lpr <<= 1;
  return lpr;
}

#define GGL_BY_LPR(lpr, CALL)        \
  switch (lpr) {                     \
    case 1: { CALL(1); break; }      \
    case 2: { CALL(2); break; }      \
    case 4: { CALL(4); break; }      \
    case 8: { CALL(8); break; }      \
    case 16: { CALL(16); break; }    \
    default: { CALL(32); break; }    \
  }

template <typename W>
static void launch_expand_rows(const void* g, const int32_t* rowptr, const int32_t* perm, const int2* tab, void* gx, int64_t E,
                               int64_t words, int64_t n_tiles, cudaStream_t stream) {
  const int grid = grid_for(n_tiles * 32);
  const int32_t te = tile_edges();
#define GGL_EXPAND(L) expand_rows_kernel<W, L><<<grid, 256, 0, stream>>>((const W*)g, rowptr, perm, tab, (W*)gx, E, (int32_t)words, te, n_tiles)
  GGL_BY_LPR(lanes_for(words), GGL_EXPAND)
#undef GGL_EXPAND
}

template <typename T>
static void launch_expand_mean(const void* g, const int32_t* rowptr, const int32_t* perm, const int2* tab, void* gx, int64_t E,
                               int64_t K, int64_t n_tiles, cudaStream_t stream) {
  const int grid = grid_for(n_tiles * 32);
  const int32_t te = tile_edges();
#define GGL_EXPAND(L) expand_rows_mean_kernel<T, L><<<grid, 256, 0, stream>>>((const T*)g, rowptr, perm, tab, (T*)gx, E, (int32_t)K, te, n_tiles)
  GGL_BY_LPR(lanes_for(K), GGL_EXPAND)
#undef GGL_EXPAND
}

template <typename W>
__global__ void __launch_bounds__(256) scatter_arg_kernel(const W* __restrict__ g, const int64_t* __restrict__ arg,
                                                          const int32_t* __restrict__ rowptr, W* __restrict__ gx,
                                                          int64_t n_rows, int64_t E, int32_t K) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
    if (__ldg(rowptr + r + 1) == __ldg(rowptr + r)) continue;  // empty segment: constant output, zero gradient
    for (int32_t k = lane; k < K; k += 32) {
      const int64_t e = __ldg(arg + r * K + k);
      if (e >= 0 && e < E) gx[e * K + k] = __ldg(g + r * K + k);
    }
  }
}

__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ g, const int32_t* __restrict__ rowptr,
                                                         float* __restrict__ out, int64_t n_rows, int32_t K) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
    const int32_t cnt = __ldg(rowptr + r + 1) - __ldg(rowptr + r);
    const float c = (float)cnt;
    for (int32_t k = lane; k < K; k += 32) {
      float v = __ldg(g + r * K + k);
      out[r * K + k] = cnt > 0 ? __fdiv_rn(v, c) : v;
    }
  }
}

// rows = source nodes; sequential over the row's edges (ascending original edge id) like the reference loop
__global__ void __launch_bounds__(256) spmm_max_bwd_kernel(const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ col_t,
                                                           const float* __restrict__ w, const int32_t* __restrict__ wperm,
                                                           const float* __restrict__ g, const int64_t* __restrict__ argsrc,
                                                           float* __restrict__ gx, int64_t n_rows, int32_t K,
                                                           const int32_t* __restrict__ eid_t /* NULL: arg holds source ids */) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n_rows; s += warps) {
    const int32_t b = __ldg(rowptr_t + s), e = __ldg(rowptr_t + s + 1);
    for (int32_t k0 = 0; k0 < K; k0 += 32) {
      const int32_t k = k0 + lane;
      float acc = 0.f;
      for (int32_t i = b; i < e; ++i) {
        const int64_t d = __ldg(col_t + i);
        const float wi = w ? __ldg(w + (wperm ? __ldg(wperm + i) : i)) : 1.f;
        // edge-id form: only the ONE winning edge of (d, k) is credited (unsorted_segment_max semantics on the message,
        // src/segment_max.cpp:48-61); source-id form: every parallel edge of the winning source (cpu/spmm_max_cpu.cpp:84-96)
        const int64_t want = eid_t ? (int64_t)__ldg(eid_t + i) : s;
        if (k < K && __ldg(argsrc + d * K + k) == want) acc = __fadd_rn(acc, __fmul_rn(wi, __ldg(g + d * K + k)));
      }
      if (k < K) gx[s * K + k] = acc;
    }
  }
}

// out[r,:] = scale[r] * x[r,:]  (pre-scaling of the folded GCN normalisation; float4 fast path for fp32)
template <typename T>
__global__ void __launch_bounds__(256) scale_rows_by_vec_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                                T* __restrict__ out, int64_t total, int32_t K) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t r = t / K;
    out[t] = Elem<T>::from_acc(__fmul_rn(__ldg(scale + r), Elem<T>::to_acc(x[t])));
  }
}
__global__ void __launch_bounds__(256) scale_rows_by_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ scale,
                                                                 float4* __restrict__ out, int64_t total4, int32_t K4) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total4; t += stride) {
    const float s = __ldg(scale + t / K4);
    float4 v = __ldcs(x + t);
    v.x = __fmul_rn(s, v.x); v.y = __fmul_rn(s, v.y); v.z = __fmul_rn(s, v.z); v.w = __fmul_rn(s, v.w);
    out[t] = v;
  }
}
// out[r] = deg(r)^power for deg > 0, else 0   (deg = rowptr[r+1] - rowptr[r])
__global__ void __launch_bounds__(256) degree_power_kernel(const int32_t* __restrict__ rowptr, float* __restrict__ out, int64_t n,
                                                           float power) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; r < n; r += stride) {
    const float d = (float)(rowptr[r + 1] - rowptr[r]);
    out[r] = d > 0.f ? (power == -0.5f ? rsqrtf(d) : (power == -1.f ? __fdiv_rn(1.f, d) : powf(d, power))) : 0.f;
  }
}

static int elem_bytes(int dtype) {
  switch (dtype) {
    case GGL_F32: case GGL_I32: return 4;
    case GGL_BF16: case GGL_F16: case GGL_I16: return 2;
    case GGL_F64: case GGL_I64: return 8;
    case GGL_I8: case GGL_U8: return 1;
    default: return 0;
  }
}

}  // namespace ggl

using namespace ggl;

extern "C" int gglb200_segment_gather_bwd(int reduce, const void* g, const int64_t* ids, const int32_t* rowptr, void* gx,
                                          int64_t E, int64_t K, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(reduce == GGL_SUM || reduce == GGL_MEAN, GGL_EINVAL, "segment_gather_bwd: reduce must be SUM or MEAN");
  GGL_REQUIRE(E >= 0 && K >= 0, GGL_EINVAL, "segment_gather_bwd: negative size");
  const int eb = elem_bytes(dtype);
  GGL_REQUIRE(eb > 0, GGL_EUNSUPPORTED, "segment_gather_bwd: unknown dtype %d", dtype);
  if (E == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(g && ids && gx, GGL_EINVAL, "segment_gather_bwd: NULL pointer");
  GGL_REQUIRE(K < INT32_MAX / 8, GGL_ERANGE, "segment_gather_bwd: K too large");
  const int grid = grid_for(E * 32);
  if (reduce == GGL_SUM) {
    const int64_t rb = K * eb;
    const uintptr_t al = (uintptr_t)g | (uintptr_t)gx;
    if (rb % 16 == 0 && al % 16 == 0) launch_gather_rows<uint4>(g, ids, gx, E, rb / 16, stream);
    else if (rb % 8 == 0 && al % 8 == 0) launch_gather_rows<uint2>(g, ids, gx, E, rb / 8, stream);
    else if (rb % 4 == 0 && al % 4 == 0) launch_gather_rows<uint32_t>(g, ids, gx, E, rb / 4, stream);
    else if (rb % 2 == 0 && al % 2 == 0) launch_gather_rows<uint16_t>(g, ids, gx, E, rb / 2, stream);
    else launch_gather_rows<uint8_t>(g, ids, gx, E, rb, stream);
    GGL_LAUNCHED("gather_rows");
    return GGL_OK;
  }
  GGL_REQUIRE(rowptr, GGL_EINVAL, "segment_gather_bwd: MEAN needs rowptr");
  if (dtype == GGL_F32 && K % 4 == 0 && (((uintptr_t)g | (uintptr_t)gx) % 16 == 0)) {
    const int64_t words = K / 4;
    const int lpr = lanes_for(words);
    const int64_t warps = (E * lpr + 127) / 128;
    const int g4 = grid_for(warps * 32);
#define GGL_GM4(L) gather_rows_mean4_kernel<L><<<g4, 256, 0, stream>>>((const float4*)g, ids, rowptr, (float4*)gx, E, (int32_t)words)
    GGL_BY_LPR(lpr, GGL_GM4)
#undef GGL_GM4
    GGL_LAUNCHED("gather_rows_mean4");
    return GGL_OK;
  }
  switch (dtype) {
    case GGL_F32: gather_rows_mean_kernel<float><<<grid, 256, 0, stream>>>((const float*)g, ids, rowptr, (float*)gx, E, (int32_t)K); break;
    case GGL_F64: gather_rows_mean_kernel<double><<<grid, 256, 0, stream>>>((const double*)g, ids, rowptr, (double*)gx, E, (int32_t)K); break;
    case GGL_BF16: gather_rows_mean_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)g, ids, rowptr, (__nv_bfloat16*)gx, E, (int32_t)K); break;
    case GGL_F16: gather_rows_mean_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)g, ids, rowptr, (__half*)gx, E, (int32_t)K); break;
    default: return set_error(GGL_EUNSUPPORTED, "segment_gather_bwd: MEAN backward needs a floating dtype (got %d)", dtype);
  }
  GGL_LAUNCHED("gather_rows_mean");
  return GGL_OK;
}

extern "C" int gglb200_segment_expand_bwd(int reduce, const void* g, const int32_t* rowptr, const int32_t* perm,
                                          const int32_t* plan, void* gx, int64_t E, int64_t n_rows, int64_t K, int dtype,
                                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(reduce == GGL_SUM || reduce == GGL_MEAN, GGL_EINVAL, "segment_expand_bwd: reduce must be SUM or MEAN");
  GGL_REQUIRE(E >= 0 && K >= 0 && n_rows >= 0, GGL_EINVAL, "segment_expand_bwd: negative size");
  GGL_REQUIRE(E < INT32_MAX && n_rows < INT32_MAX, GGL_ERANGE, "segment_expand_bwd: n_rows/E exceed int32");
  const int eb = elem_bytes(dtype);
  GGL_REQUIRE(eb > 0, GGL_EUNSUPPORTED, "segment_expand_bwd: unknown dtype %d", dtype);
  if (E == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(g && rowptr && perm && plan && gx, GGL_EINVAL, "segment_expand_bwd: NULL pointer");
  GGL_REQUIRE(n_rows > 0, GGL_EINVAL, "segment_expand_bwd: E > 0 needs n_rows > 0");
  GGL_REQUIRE(K < INT32_MAX / 8, GGL_ERANGE, "segment_expand_bwd: K too large");
  const int64_t n_tiles = plan_n_tiles(E);
  const int2* tab = (const int2*)(plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E)));
  if (reduce == GGL_SUM) {
    const int64_t rb = K * eb;
    const uintptr_t al = (uintptr_t)g | (uintptr_t)gx;
    if (rb % 16 == 0 && al % 16 == 0) launch_expand_rows<uint4>(g, rowptr, perm, tab, gx, E, rb / 16, n_tiles, stream);
    else if (rb % 8 == 0 && al % 8 == 0) launch_expand_rows<uint2>(g, rowptr, perm, tab, gx, E, rb / 8, n_tiles, stream);
    else if (rb % 4 == 0 && al % 4 == 0) launch_expand_rows<uint32_t>(g, rowptr, perm, tab, gx, E, rb / 4, n_tiles, stream);
    else if (rb % 2 == 0 && al % 2 == 0) launch_expand_rows<uint16_t>(g, rowptr, perm, tab, gx, E, rb / 2, n_tiles, stream);
    else launch_expand_rows<uint8_t>(g, rowptr, perm, tab, gx, E, rb, n_tiles, stream);
    GGL_LAUNCHED("expand_rows");
    return GGL_OK;
  }
  if (dtype == GGL_F32 && K % 4 == 0 && (((uintptr_t)g | (uintptr_t)gx) % 16 == 0)) {
    const int64_t words = K / 4;
    const int g4 = grid_for(n_tiles * 32);
    const int32_t te = tile_edges();
#define GGL_EM4(L) expand_rows_mean4_kernel<L><<<g4, 256, 0, stream>>>((const float4*)g, rowptr, perm, tab, (float4*)gx, E, (int32_t)words, te, n_tiles)
    GGL_BY_LPR(lanes_for(words), GGL_EM4)
#undef GGL_EM4
    GGL_LAUNCHED("expand_rows_mean4");
    return GGL_OK;
  }
  switch (dtype) {
    case GGL_F32: launch_expand_mean<float>(g, rowptr, perm, tab, gx, E, K, n_tiles, stream); break;
    case GGL_F64: launch_expand_mean<double>(g, rowptr, perm, tab, gx, E, K, n_tiles, stream); break;
    case GGL_BF16: launch_expand_mean<__nv_bfloat16>(g, rowptr, perm, tab, gx, E, K, n_tiles, stream); break;
    case GGL_F16: launch_expand_mean<__half>(g, rowptr, perm, tab, gx, E, K, n_tiles, stream); break;
    default: return set_error(GGL_EUNSUPPORTED, "segment_expand_bwd: MEAN backward needs a floating dtype (got %d)", dtype);
  }
  GGL_LAUNCHED("expand_rows_mean");
  return GGL_OK;
}

extern "C" int gglb200_segment_max_bwd(const void* g, const int64_t* arg, const int32_t* rowptr, void* gx, int64_t E,
                                       int64_t n_rows, int64_t K, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(E >= 0 && K >= 0 && n_rows >= 0, GGL_EINVAL, "segment_max_bwd: negative size");
  const int eb = elem_bytes(dtype);
  GGL_REQUIRE(eb > 0, GGL_EUNSUPPORTED, "segment_max_bwd: unknown dtype %d", dtype);
  if (E == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(gx, GGL_EINVAL, "segment_max_bwd: gx is NULL");
  GGL_CUDA(cudaMemsetAsync(gx, 0, (size_t)E * (size_t)K * eb, stream));
  if (n_rows == 0) return GGL_OK;
  GGL_REQUIRE(g && arg && rowptr, GGL_EINVAL, "segment_max_bwd: NULL pointer");
  const int grid = grid_for(n_rows * 32);
  switch (eb) {
    case 1: scatter_arg_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)g, arg, rowptr, (uint8_t*)gx, n_rows, E, (int32_t)K); break;
    case 2: scatter_arg_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)g, arg, rowptr, (uint16_t*)gx, n_rows, E, (int32_t)K); break;
    case 4: scatter_arg_kernel<uint32_t><<<grid, 256, 0, stream>>>((const uint32_t*)g, arg, rowptr, (uint32_t*)gx, n_rows, E, (int32_t)K); break;
    default: scatter_arg_kernel<uint64_t><<<grid, 256, 0, stream>>>((const uint64_t*)g, arg, rowptr, (uint64_t*)gx, n_rows, E, (int32_t)K); break;
  }
  GGL_LAUNCHED("scatter_arg");
  return GGL_OK;
}

extern "C" int gglb200_scale_rows_by_count(const float* g, const int32_t* rowptr, float* out, int64_t n_rows, int64_t K,
                                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && K >= 0, GGL_EINVAL, "scale_rows: negative size");
  if (n_rows == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(g && rowptr && out, GGL_EINVAL, "scale_rows: NULL pointer");
  scale_rows_kernel<<<grid_for(n_rows * 32), 256, 0, stream>>>(g, rowptr, out, n_rows, (int32_t)K);
  GGL_LAUNCHED("scale_rows");
  return GGL_OK;
}

extern "C" int gglb200_scale_rows(const void* x, const float* scale, void* out, int64_t n_rows, int64_t K, int dtype,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && K >= 0, GGL_EINVAL, "scale_rows: negative size");
  if (n_rows == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(x && scale && out, GGL_EINVAL, "scale_rows: NULL pointer");
  const int64_t total = n_rows * K;
  if (dtype == GGL_F32 && K % 4 == 0 && (((uintptr_t)x | (uintptr_t)out) % 16 == 0)) {
    scale_rows_by_vec4_kernel<<<grid_for(total / 4), 256, 0, stream>>>((const float4*)x, scale, (float4*)out, total / 4, (int32_t)(K / 4));
  } else if (dtype == GGL_F32) {
    scale_rows_by_vec_kernel<float><<<grid_for(total), 256, 0, stream>>>((const float*)x, scale, (float*)out, total, (int32_t)K);
  } else if (dtype == GGL_BF16) {
    scale_rows_by_vec_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>((const __nv_bfloat16*)x, scale, (__nv_bfloat16*)out, total, (int32_t)K);
  } else if (dtype == GGL_F16) {
    scale_rows_by_vec_kernel<__half><<<grid_for(total), 256, 0, stream>>>((const __half*)x, scale, (__half*)out, total, (int32_t)K);
  } else {
    return set_error(GGL_EUNSUPPORTED, "scale_rows: floating dtypes only (got %d)", dtype);
  }
  GGL_LAUNCHED("scale_rows_by_vec");
  return GGL_OK;
}

extern "C" int gglb200_degree_power(const int32_t* rowptr, float* out, int64_t n_rows, float power, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0, GGL_EINVAL, "degree_power: negative size");
  if (n_rows == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && out, GGL_EINVAL, "degree_power: NULL pointer");
  degree_power_kernel<<<grid_for(n_rows), 256, 0, stream>>>(rowptr, out, n_rows, power);
  GGL_LAUNCHED("degree_power");
  return GGL_OK;
}

extern "C" int gglb200_spmm_max_bwd(const int32_t* rowptr_t, const int32_t* col_t, const float* w, const int32_t* wperm,
                                    const float* g, const int64_t* argsrc, float* gx, int64_t n_rows, int64_t K,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && K >= 0, GGL_EINVAL, "spmm_max_bwd: negative size");
  if (n_rows == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(rowptr_t && g && argsrc && gx, GGL_EINVAL, "spmm_max_bwd: NULL pointer");
  spmm_max_bwd_kernel<<<grid_for(n_rows * 32), 256, 0, stream>>>(rowptr_t, col_t, w, wperm, g, argsrc, gx, n_rows, (int32_t)K, nullptr);
  GGL_LAUNCHED("spmm_max_bwd");
  return GGL_OK;
}

extern "C" int gglb200_spmm_max_edge_bwd(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* perm_t, const float* w,
                                         const float* g, const int64_t* arg_eid, float* gx, int64_t n_rows, int64_t K,
                                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && K >= 0, GGL_EINVAL, "spmm_max_edge_bwd: negative size");
  if (n_rows == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(rowptr_t && perm_t && g && arg_eid && gx, GGL_EINVAL, "spmm_max_edge_bwd: NULL pointer");
  spmm_max_bwd_kernel<<<grid_for(n_rows * 32), 256, 0, stream>>>(rowptr_t, col_t, w, w ? perm_t : nullptr, g, arg_eid, gx, n_rows,
                                                                 (int32_t)K, perm_t);
  GGL_LAUNCHED("spmm_max_edge_bwd");
  return GGL_OK;
}
