// gat.cu -- the GAT-specific pieces of the path.
//
//   bspmm_bwd_w     grad of c_bspmm_sum wrt the per-head edge weight (SDDMM),
//                   /root/reference/gammagl/mpops/torch_ext/cpu/bspmm_sum_cpu.cpp:95-107
//   edge_softmax    segment_softmax over incoming edges, /root/reference/gammagl/utils/softmax.py:29-35,
//                   fused: max -> exp -> sum -> divide in one kernel over the forward CSR instead of
//                   5 passes over [E,H] + 2 segment reductions.
#include "common.cuh"

namespace ggl {

constexpr int kWarps = 8;

// ---- SDDMM: gw[perm[i], h] = sum_c x[col[i], h, c] * g[r, h, c] for i in row r -----------------------
// Fast path: C % 4 == 0, LH = C/4 a power of two <= 32, K = H*C <= 128*TILES: a lane owns 4 consecutive
// channels of one head per 128-column tile; the LH lanes of a head combine with xor-shuffles.
// The g row is loaded once per row; x rows are gathered U edges at a time.
template <int LH, int TILES>
__global__ void __launch_bounds__(kWarps * 32) sddmm_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                 const int32_t* __restrict__ perm, const float* __restrict__ x,
                                                                 const float* __restrict__ g, float* __restrict__ gw,
                                                                 int64_t n_rows, int32_t H, int32_t C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;  // warp-uniform
  const int32_t K = H * C;
  const int32_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  Pack<float, 4> gr[TILES];
#pragma unroll
  for (int t = 0; t < TILES; ++t) {
    const int32_t k = t * 128 + lane * 4;
    if (k < K) gr[t] = ld_gather<float, 4>(g + row * K + k);
    else gr[t].v[0] = gr[t].v[1] = gr[t].v[2] = gr[t].v[3] = 0.f;
  }
  constexpr int U = 4;
  for (int32_t base = b; base < e; base += 32) {
    const int n = min(32, e - base);
    int32_t mycol = 0, myeid = 0;
    if (lane < n) { mycol = __ldcs(col + base + lane); myeid = __ldcs(perm + base + lane); }
#pragma unroll 1
    for (int j = 0; j < n; j += U) {
      Pack<float, 4> xv[U][TILES];
      int32_t eid[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t c = __shfl_sync(0xffffffffu, mycol, (j + u) & 31);
        eid[u] = __shfl_sync(0xffffffffu, myeid, (j + u) & 31);
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const int32_t k = t * 128 + lane * 4;
          if (j + u < n && k < K) xv[u][t] = ld_gather<float, 4>(x + (int64_t)c * K + k);
          else xv[u][t].v[0] = xv[u][t].v[1] = xv[u][t].v[2] = xv[u][t].v[3] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          float s = __fmul_rn(xv[u][t].v[0], gr[t].v[0]);
          s = __fadd_rn(s, __fmul_rn(xv[u][t].v[1], gr[t].v[1]));
          s = __fadd_rn(s, __fmul_rn(xv[u][t].v[2], gr[t].v[2]));
          s = __fadd_rn(s, __fmul_rn(xv[u][t].v[3], gr[t].v[3]));
#pragma unroll
          for (int off = LH / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
          const int32_t k = t * 128 + lane * 4;
          if (j + u < n && k < K && (lane & (LH - 1)) == 0) gw[(int64_t)eid[u] * H + k / C] = s;
        }
      }
    }
  }
}

// Generic path (any H, C): per edge and head, lanes stride over channels, full-warp reduction.
__global__ void __launch_bounds__(kWarps * 32) sddmm_rows_generic_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                         const int32_t* __restrict__ perm, const float* __restrict__ x,
                                                                         const float* __restrict__ g, float* __restrict__ gw,
                                                                         int64_t n_rows, int32_t H, int32_t C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t K = (int64_t)H * C;
  const int32_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  for (int32_t i = b; i < e; ++i) {
    const int64_t c = __ldg(col + i), eid = __ldg(perm + i);
    for (int32_t h = 0; h < H; ++h) {
      float s = 0.f;
      for (int32_t ch = lane; ch < C; ch += 32) s += __ldg(x + c * K + h * C + ch) * __ldg(g + row * K + h * C + ch);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) gw[eid * H + h] = s;
    }
  }
}

// ---- fused edge softmax ---------------------------------------------------------------------------
// A group of LPH lanes (>= H, power of two) owns one destination row; lane h walks the row's logits of
// head h in edge order.  exp(l - max) is the accurate expf (the reference calls torch.exp); the sum is
// sequential in ascending edge id like the reference's segment_sum.
template <int LPH>
__global__ void __launch_bounds__(kWarps * 32) edge_softmax_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                                                       const float* __restrict__ logit, float* __restrict__ alpha,
                                                                       int64_t n_rows, int32_t H) {
  constexpr int RPW = 32 / LPH;
  const int lane = threadIdx.x & 31;
  const int h = lane & (LPH - 1);
  const int64_t row = ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * RPW + lane / LPH;
  if (row >= n_rows || h >= H) return;
  const int32_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  float m = -3.402823466e+38f;  // lowest(): unsorted_segment_max's start value (cpu/segment_max_cpu.cpp:40)
  for (int32_t i = b; i < e; ++i) {
    const float v = __ldg(logit + (int64_t)__ldg(perm + i) * H + h);
    if (m < v) m = v;
  }
  float s = 0.f;
  for (int32_t i = b; i < e; ++i) s = __fadd_rn(s, expf(__ldg(logit + (int64_t)__ldg(perm + i) * H + h) - m));
  const float den = s + 1e-16f;
  for (int32_t i = b; i < e; ++i) {
    const int64_t o = (int64_t)__ldg(perm + i) * H + h;
    alpha[o] = __fdiv_rn(expf(__ldg(logit + o) - m), den);
  }
}

template <int LPH>
__global__ void __launch_bounds__(kWarps * 32) edge_softmax_bwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                                                       const float* __restrict__ alpha, const float* __restrict__ galpha,
                                                                       float* __restrict__ glogit, int64_t n_rows, int32_t H) {
  constexpr int RPW = 32 / LPH;
  const int lane = threadIdx.x & 31;
  const int h = lane & (LPH - 1);
  const int64_t row = ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * RPW + lane / LPH;
  if (row >= n_rows || h >= H) return;
  const int32_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  float dot = 0.f;
  for (int32_t i = b; i < e; ++i) {
    const int64_t o = (int64_t)__ldg(perm + i) * H + h;
    dot += __ldg(alpha + o) * __ldg(galpha + o);
  }
  for (int32_t i = b; i < e; ++i) {
    const int64_t o = (int64_t)__ldg(perm + i) * H + h;
    glogit[o] = __ldg(alpha + o) * (__ldg(galpha + o) - dot);
  }
}

}  // namespace ggl

using namespace ggl;

extern "C" int gglb200_bspmm_bwd_w(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* x,
                                   const float* g, float* gw, int64_t n_rows, int64_t E, int32_t H, int32_t C,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1 && C >= 1, GGL_EINVAL, "bspmm_bwd_w: bad sizes");
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && col && perm && x && g && gw, GGL_EINVAL, "bspmm_bwd_w: NULL pointer");
  const int64_t K = (int64_t)H * C;
  const unsigned blocks = (unsigned)((n_rows + kWarps - 1) / kWarps);
  const int lh = C / 4;
  const bool fast = C % 4 == 0 && lh >= 1 && lh <= 32 && (lh & (lh - 1)) == 0 && K <= 512 &&
                    (((uintptr_t)x | (uintptr_t)g) % 16 == 0);
  if (fast) {
    const int tiles = (int)((K + 127) / 128);
#define GGL_SDDMM(LH_, T_) \
  sddmm_rows_kernel<LH_, T_><<<blocks, kWarps * 32, 0, stream>>>(rowptr, col, perm, x, g, gw, n_rows, H, C)
#define GGL_SDDMM_T(LH_)                              \
  do {                                                \
    if (tiles == 1) GGL_SDDMM(LH_, 1);                \
    else if (tiles == 2) GGL_SDDMM(LH_, 2);           \
    else GGL_SDDMM(LH_, 4);                           \
  } while (0)
    switch (lh) {
      case 1: GGL_SDDMM_T(1); break;
      case 2: GGL_SDDMM_T(2); break;
      case 4: GGL_SDDMM_T(4); break;
      case 8: GGL_SDDMM_T(8); break;
      case 16: GGL_SDDMM_T(16); break;
      default: GGL_SDDMM_T(32); break;
    }
#undef GGL_SDDMM_T
#undef GGL_SDDMM
    GGL_LAUNCHED("sddmm_rows");
  } else {
    sddmm_rows_generic_kernel<<<blocks, kWarps * 32, 0, stream>>>(rowptr, col, perm, x, g, gw, n_rows, H, C);
    GGL_LAUNCHED("sddmm_rows_generic");
  }
  return GGL_OK;
}

template <bool FWD>
static int launch_softmax(const int32_t* rowptr, const int32_t* perm, const float* a, const float* b, float* o,
                          int64_t n_rows, int32_t H, cudaStream_t stream) {
  int lph = 1;
  while (lph < H) lph <<= 1;
  GGL_REQUIRE(lph <= 32, GGL_EUNSUPPORTED, "edge_softmax: H=%d > 32 heads not supported", H);
  const int rpw = 32 / lph;
  const unsigned blocks = (unsigned)((n_rows + (int64_t)kWarps * rpw - 1) / ((int64_t)kWarps * rpw));
#define GGL_SM(L_)                                                                                             \
  do {                                                                                                         \
    if constexpr (FWD) edge_softmax_fwd_kernel<L_><<<blocks, kWarps * 32, 0, stream>>>(rowptr, perm, a, o, n_rows, H); \
    else edge_softmax_bwd_kernel<L_><<<blocks, kWarps * 32, 0, stream>>>(rowptr, perm, a, b, o, n_rows, H);    \
  } while (0)
  switch (lph) {
    case 1: GGL_SM(1); break;
    case 2: GGL_SM(2); break;
    case 4: GGL_SM(4); break;
    case 8: GGL_SM(8); break;
    case 16: GGL_SM(16); break;
    default: GGL_SM(32); break;
  }
#undef GGL_SM
  GGL_LAUNCHED("edge_softmax");
  return GGL_OK;
}

extern "C" int gglb200_edge_softmax_fwd(const int32_t* rowptr, const int32_t* perm, const float* logit, float* alpha,
                                        int64_t n_rows, int64_t E, int32_t H, void* stream) {
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1, GGL_EINVAL, "edge_softmax_fwd: bad sizes");
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && perm && logit && alpha, GGL_EINVAL, "edge_softmax_fwd: NULL pointer");
  return launch_softmax<true>(rowptr, perm, logit, nullptr, alpha, n_rows, H, (cudaStream_t)stream);
}

extern "C" int gglb200_edge_softmax_bwd(const int32_t* rowptr, const int32_t* perm, const float* alpha, const float* galpha,
                                        float* glogit, int64_t n_rows, int64_t E, int32_t H, void* stream) {
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1, GGL_EINVAL, "edge_softmax_bwd: bad sizes");
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && perm && alpha && galpha && glogit, GGL_EINVAL, "edge_softmax_bwd: NULL pointer");
  return launch_softmax<false>(rowptr, perm, alpha, galpha, glogit, n_rows, H, (cudaStream_t)stream);
}
