// gat.cu -- the GAT-specific pieces of the path.
//
//   bspmm_bwd_w     grad of c_bspmm_sum wrt the per-head edge weight (SDDMM),
//                   /root/reference/gammagl/mpops/torch_ext/cpu/bspmm_sum_cpu.cpp:95-107
//   edge_softmax    segment_softmax over incoming edges, /root/reference/gammagl/utils/softmax.py:29-35
#include <cstdlib>

#include "spmm_impl.cuh"

namespace ggl {

// ---- SDDMM: gw[perm[i], h] = sum_c x[col[i], h, c] * g[r, h, c] for i in row r -----------------------------------
// Same machinery as spmm_ring_kernel -- x rows are bulk-copied (UBLKCP) into a shared-memory ring and consumed in
// stream order -- but every edge is independent, so tiles are PURE edge ranges [t*TE, (t+1)*TE): perfectly balanced,
// no long-row path.  The g row of the current destination lives in registers (the next row's is prefetched); a lane
// owns 4 consecutive channels of one head and the C/4 lanes of a head combine with xor-shuffles.
// Requires C % 4 == 0 with C/4 a power of two <= LPR (heads never straddle a 128-column tile).
// SPAN = true lifts that: any head width C >= 4 with K = H*C a multiple of 4 and K <= 128 (one column tile).  A lane's four
// channels may then reach into the next head: it forms two partial dots (its own head | the next one), hands the second to its
// right neighbour (whose first channel always belongs to that head when C >= 4), and the lanes of a head -- a contiguous run of
// arbitrary length -- combine with a segmented shuffle scan; the last lane of each run writes gw[e, h].
template <int LPR, bool SPAN>
__global__ void __launch_bounds__(32) sddmm_ring_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                        const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_row,
                                                        const int64_t n_tiles, const int32_t tile_edges, const int32_t E,
                                                        const float* __restrict__ x, const float* __restrict__ g,
                                                        float* __restrict__ gw, const int32_t n_rows, const int32_t H,
                                                        const int32_t C, const int32_t pitch) {
  using Cfg = RingCfg<LPR>;
  constexpr int VEC = 4, GPW = 32 / LPR, R = Cfg::R, STAGES = Cfg::STAGES, LAG = Cfg::LAG;
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPR - 1);
  const int gbase = lane & ~(LPR - 1);
  const unsigned gmask = group_mask<LPR>(lane);
  const int gib = lane / LPR;
  const int64_t tile = (int64_t)blockIdx.x * GPW + gib;
  const int32_t K = H * C;
  const int32_t kcol = ((int32_t)blockIdx.y * LPR + sub) * VEC;
  const bool col_active = kcol < K;
  const int32_t head = col_active ? kcol / C : (SPAN ? -1 - sub : 0);
  const int LH = C / VEC;  // lanes per head (SPAN = false)
  int32_t nlo = 4;         // SPAN: channels [0, nlo) of this lane belong to `head`, the rest to head + 1
  bool run_last = false;   // SPAN: this lane is the last one of its head's run
  if constexpr (SPAN) {
    if (col_active) nlo = min(4, (head + 1) * C - kcol);
    const int32_t next_head = __shfl_down_sync(gmask, head, 1, LPR);
    run_last = col_active && (sub == LPR - 1 || next_head != head);
  }
  const float* xt = x + (int64_t)blockIdx.y * LPR * VEC;
  const uint32_t rowb = (uint32_t)min(LPR * 16, K * 4 - (int32_t)blockIdx.y * LPR * 16);

  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(ring_smem);
  const uint32_t ring = smem0 + (uint32_t)(gib * (STAGES * R)) * pitch;
  const uint32_t bars = smem0 + (uint32_t)(GPW * STAGES * R) * pitch + (uint32_t)gib * STAGES * 8;
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) mbar_init(bars + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp(gmask);
  if (tile >= n_tiles) return;
  const int32_t e0 = (int32_t)(tile * tile_edges);
  const int32_t e_end = min(e0 + tile_edges, E);
  if (e0 >= e_end) return;
  const int2 te = __ldg(reinterpret_cast<const int2*>(tile_row) + tile);  // {first row starting at/after e0 | flag, its first edge}
  int32_t r = te.x & 0x7fffffff;
  if (r >= n_rows || te.y > e0) --r;  // the row that contains edge e0 (empty rows are skipped below)
  RowCursor<LPR> rc;
  rc.init(rowptr, r, n_rows, e0, sub, gbase, gmask);

  Pack<float, 4> gv, gnext;
  int32_t gnext_row = -1;
  auto load_g = [&](int32_t row) {
    Pack<float, 4> t;
    t.v[0] = t.v[1] = t.v[2] = t.v[3] = 0.f;
    if (col_active && row < n_rows) t = ld_gather<float, 4>(g + (int64_t)row * K + kcol);
    return t;
  };
  bool g_valid = false;

  const int32_t n = e_end - e0;
  const int32_t nslots = (n + R - 1) / R;
  struct Batch { int32_t col, eid; };
  auto load_batch = [&](int32_t b) {
    Batch t; t.col = 0; t.eid = 0;
    const int32_t pos = b * LPR + sub;
    if (pos < n) { t.col = ld_stream_i32(col + e0 + pos); t.eid = ld_stream_i32(perm + e0 + pos); }
    return t;
  };
  Batch A = load_batch(0), B = load_batch(1), Cb = load_batch(2);
  int32_t bA = 0;
  auto issue = [&](int32_t s) {
    if (s < nslots) {
      const int32_t pos0 = s * R;
      const int32_t d = pos0 / LPR - bA;
      const int32_t colsrc = d == 0 ? A.col : (d == 1 ? B.col : Cb.col);
      const uint32_t dst = ring + (uint32_t)((s % STAGES) * R) * pitch;
      const uint32_t bar = bars + (uint32_t)(s % STAGES) * 8;
      const int32_t nv = min(R, n - pos0);
      const int32_t c = __shfl_sync(gmask, colsrc, gbase + ((pos0 + sub) & (LPR - 1)));
      __syncwarp(gmask);
      if (sub == 0) mbar_arrive_expect_tx(bar, (uint32_t)nv * rowb);
      if (sub < nv) bulk_copy_g2s(dst + (uint32_t)sub * pitch, xt + (int64_t)c * K, rowb, bar);
    }
  };
#pragma unroll
  for (int s = 0; s < LAG; ++s) issue(s);
  for (int32_t s = 0; s < nslots; ++s) {
    const int32_t pos0 = s * R;
    if (pos0 / LPR != bA) { A = B; B = Cb; ++bA; Cb = load_batch(bA + 2); }
    issue(s + LAG);
    int32_t eid[R];
#pragma unroll
    for (int u = 0; u < R; ++u) eid[u] = __shfl_sync(gmask, A.eid, gbase + ((pos0 + u) & (LPR - 1)));
    mbar_wait(bars + (uint32_t)(s % STAGES) * 8, (uint32_t)(s / STAGES) & 1u);
    const uint32_t src = ring + (uint32_t)((s % STAGES) * R) * pitch + (uint32_t)sub * 16;
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int32_t pos = pos0 + u;
      if (pos < n) {  // group-uniform
        if (e0 + pos == rc.cur_end || !g_valid) {
          while (e0 + pos == rc.cur_end) rc.advance(rowptr, sub, gbase, gmask);
          gv = (gnext_row == rc.r) ? gnext : load_g(rc.r);
          gnext_row = rc.r + 1;
          gnext = load_g(gnext_row);
          g_valid = true;
        }
        if constexpr (!SPAN) {
          float dot = 0.f;
          if (col_active) {
            const Pack<float, 4> v = lds_pack<float, 4>(src + u * pitch);
            dot = fmaf(v.v[3], gv.v[3], fmaf(v.v[2], gv.v[2], fmaf(v.v[1], gv.v[1], v.v[0] * gv.v[0])));
          }
          for (int off = LH >> 1; off > 0; off >>= 1) dot += __shfl_xor_sync(gmask, dot, off);
          if (col_active && (sub & (LH - 1)) == 0) gw[(int64_t)eid[u] * H + head] = dot;
        } else {
          float lo = 0.f, hi = 0.f;
          if (col_active) {
            const Pack<float, 4> v = lds_pack<float, 4>(src + u * pitch);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float pr = v.v[i] * gv.v[i];
              if (i < nlo) lo += pr; else hi += pr;
            }
          }
          const float from_left = __shfl_up_sync(gmask, hi, 1, LPR);      // the left neighbour's spill-over belongs to this lane's head
          float dot = lo + (sub > 0 ? from_left : 0.f);
#pragma unroll
          for (int d = 1; d < LPR; d <<= 1) {                         // segmented inclusive scan over runs of equal head
            const float t = __shfl_up_sync(gmask, dot, d, LPR);
            const int32_t hh = __shfl_up_sync(gmask, head, d, LPR);
            if (sub >= d && hh == head) dot += t;
          }
          if (run_last) gw[(int64_t)eid[u] * H + head] = dot;
        }
      }
    }
  }
}

// Generic path (any H, C): per edge and head, lanes stride over channels, full-warp reduction.
constexpr int kWarps = 8;
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

// ---- edge softmax: the reference composition (utils/softmax.py:29-35) on the fast segment kernels ------------------
//   m = unsorted_segment_max(logit);  p = exp(logit - m[ids]);  s = unsorted_segment_sum(p);  alpha = p / (s[ids] + 1e-16)
// The two segment reductions are spmm_ring launches over the segment CSR (col = perm); the element-wise steps are
// the kernels below (one thread per (edge, head), coalesced).  Same arithmetic order as the reference: the sums
// are sequential in ascending edge id, so alpha differs from the reference only through expf.
__global__ void __launch_bounds__(256) softmax_exp_kernel(const float* __restrict__ logit, const float* __restrict__ m,
                                                          const int64_t* __restrict__ ids, float* __restrict__ p, int64_t total, int32_t H) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int32_t h = (int32_t)(t - e * H);
    p[t] = expf(__ldcs(logit + t) - __ldg(m + __ldg(ids + e) * H + h));
  }
}
__global__ void __launch_bounds__(256) softmax_div_kernel(float* __restrict__ alpha, const float* __restrict__ s,
                                                          const int64_t* __restrict__ ids, int64_t total, int32_t H) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int32_t h = (int32_t)(t - e * H);
    alpha[t] = __fdiv_rn(alpha[t], __fadd_rn(__ldg(s + __ldg(ids + e) * H + h), 1e-16f));
  }
}
__global__ void __launch_bounds__(256) softmax_prod_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           float* __restrict__ o, int64_t total) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) o[t] = __ldcs(a + t) * __ldcs(b + t);
}
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ galpha,
                                                          const float* __restrict__ dot, const int64_t* __restrict__ ids,
                                                          float* __restrict__ gl, int64_t total, int32_t H) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int32_t h = (int32_t)(t - e * H);
    gl[t] = __ldcs(alpha + t) * (__ldcs(galpha + t) - __ldg(dot + __ldg(ids + e) * H + h));
  }
}

// alpha = exp(logit - m[ids]) / (s[ids] + 1e-16), four heads per thread (the second half of the fused forward)
__global__ void __launch_bounds__(256) softmax_norm4_kernel(const float4* __restrict__ logit, const float* __restrict__ m,
                                                            const float* __restrict__ s, const int64_t* __restrict__ ids,
                                                            float4* __restrict__ alpha, int64_t total4, int32_t H4) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total4; t += stride) {
    const int64_t e = t / H4;
    const int64_t off = (__ldg(ids + e) * H4 + (t - e * H4)) * 4;
    const float4 mm = __ldg(reinterpret_cast<const float4*>(m + off));
    const float4 ss = __ldg(reinterpret_cast<const float4*>(s + off));
    const float4 v = __ldcs(logit + t);
    float4 o;
    o.x = __fdiv_rn(expf(v.x - mm.x), __fadd_rn(ss.x, 1e-16f));
    o.y = __fdiv_rn(expf(v.y - mm.y), __fadd_rn(ss.y, 1e-16f));
    o.z = __fdiv_rn(expf(v.z - mm.z), __fadd_rn(ss.z, 1e-16f));
    o.w = __fdiv_rn(expf(v.w - mm.w), __fadd_rn(ss.w, 1e-16f));
    __stcs(alpha + t, o);
  }
}

int softmax_stats(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const float* logit, float* stat_m, float* stat_s,
                  int64_t n_rows, int64_t E, int32_t H, void* part_ws, size_t part_bytes, cudaStream_t stream);  // gat_fused.cu

int softmax_dot(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const float* alpha, const float* galpha, float* dot,
                int64_t n_rows, int64_t E, int32_t H, void* part_ws, size_t part_bytes, cudaStream_t stream);  // gat_fused.cu

// glogit = alpha * (galpha - dot[ids]), four heads per thread
__global__ void __launch_bounds__(256) softmax_bwd4_kernel(const float4* __restrict__ alpha, const float4* __restrict__ galpha,
                                                           const float* __restrict__ dot, const int64_t* __restrict__ ids,
                                                           float4* __restrict__ gl, int64_t total4, int32_t H4) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total4; t += stride) {
    const int64_t e = t / H4;
    const float4 d = __ldg(reinterpret_cast<const float4*>(dot + (__ldg(ids + e) * H4 + (t - e * H4)) * 4));
    const float4 a = __ldcs(alpha + t), g = __ldcs(galpha + t);
    __stcs(gl + t, make_float4(a.x * (g.x - d.x), a.y * (g.y - d.y), a.z * (g.z - d.z), a.w * (g.w - d.w)));
  }
}

static bool softmax_fused_enabled() {  // env GGLB200_SOFTMAX=composed keeps the four-kernel composition
  static const bool v = [] {
    const char* e = getenv("GGLB200_SOFTMAX");
    return !(e && e[0] == 'c');
  }();
  return v;
}

static int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// segment reduction of x[E,H] over the segment CSR (col = perm) into out[n_rows,H]
static int segment_reduce_f32(int reduce, const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const float* x,
                              float* out, int64_t n_rows, int64_t E, int32_t H, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return gglb200_spmm_csr(reduce, rowptr, perm, plan, nullptr, nullptr, 1, x, out, nullptr, 0, nullptr, n_rows, E, E, H, GGL_F32, ws,
                          ws_bytes, (void*)stream);
}

}  // namespace ggl

using namespace ggl;

extern "C" int gglb200_bspmm_bwd_w(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* plan,
                                   const float* x, const float* g, float* gw, int64_t n_rows, int64_t E, int32_t H,
                                   int32_t C, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1 && C >= 1, GGL_EINVAL, "bspmm_bwd_w: bad sizes");
  GGL_REQUIRE(n_rows < INT32_MAX && E < INT32_MAX, GGL_ERANGE, "bspmm_bwd_w: n_rows/E exceed int32");
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && col && perm && plan && x && g && gw, GGL_EINVAL, "bspmm_bwd_w: NULL pointer");
  const int64_t K = (int64_t)H * C;
  const int lh = C / 4;
  const bool aligned = (((uintptr_t)x | (uintptr_t)g) % 16 == 0);
  const bool fast = C % 4 == 0 && lh >= 1 && lh <= 32 && (lh & (lh - 1)) == 0 && aligned;
  // any other head width (H=4, C=25 ...): spanning lanes, one column tile
  const bool span = !fast && C >= 4 && K % 4 == 0 && K <= 128 && aligned;
  if (fast || span) {
    int lpr = lpr_for(K, 4);
    if (lpr < 4) lpr = 4;
    if (fast && lpr < lh) lpr = lh;
    const int tiles = (int)((K + (int64_t)lpr * 4 - 1) / ((int64_t)lpr * 4));
    const int64_t n_tiles = plan_n_tiles(E);
    const int32_t* tile_row = plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E));
    const int32_t pitch = (int32_t)(K * 4 < lpr * 16 ? K * 4 : lpr * 16);
#define GGL_SDDMM(L_)                                                                                                       \
  do {                                                                                                                      \
    using Cfg = RingCfg<L_>;                                                                                                \
    constexpr int GPW = 32 / L_;                                                                                            \
    const size_t smem = (size_t)GPW * Cfg::STAGES * Cfg::R * pitch + (size_t)GPW * Cfg::STAGES * 8;                          \
    const int64_t blocks = (n_tiles + GPW - 1) / GPW;                                                                       \
    if (span) {                                                                                                             \
      GGL_CUDA(cudaFuncSetAttribute(sddmm_ring_kernel<L_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));  \
      sddmm_ring_kernel<L_, true><<<dim3((unsigned)blocks, tiles), 32, smem, stream>>>(rowptr, col, perm, tile_row, n_tiles, \
                                                                                       tile_edges(), (int32_t)E, x, g, gw,  \
                                                                                       (int32_t)n_rows, H, C, pitch);       \
    } else {                                                                                                                \
      GGL_CUDA(cudaFuncSetAttribute(sddmm_ring_kernel<L_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
      sddmm_ring_kernel<L_, false><<<dim3((unsigned)blocks, tiles), 32, smem, stream>>>(rowptr, col, perm, tile_row, n_tiles, \
                                                                                        tile_edges(), (int32_t)E, x, g, gw, \
                                                                                        (int32_t)n_rows, H, C, pitch);      \
    }                                                                                                                       \
  } while (0)
    switch (lpr) {
      case 4: GGL_SDDMM(4); break;
      case 8: GGL_SDDMM(8); break;
      case 16: GGL_SDDMM(16); break;
      default: GGL_SDDMM(32); break;
    }
#undef GGL_SDDMM
    GGL_LAUNCHED("sddmm_ring");
  } else {
    const unsigned blocks = (unsigned)((n_rows + kWarps - 1) / kWarps);
    sddmm_rows_generic_kernel<<<blocks, kWarps * 32, 0, stream>>>(rowptr, col, perm, x, g, gw, n_rows, H, C);
    GGL_LAUNCHED("sddmm_rows_generic");
  }
  return GGL_OK;
}

extern "C" size_t gglb200_edge_softmax_workspace_bytes(int64_t n_rows, int32_t H, int64_t n_chunks) {
  if (n_rows < 0 || H < 1) return 0;
  const size_t stat = (((size_t)n_rows * H * 4) + 255) & ~(size_t)255;
  return 2 * stat + gglb200_spmm_workspace_bytes_for_chunks(n_chunks, H, GGL_MAX);
}

static int softmax_args(const char* who, const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const int64_t* ids,
                        int64_t n_rows, int64_t E, int32_t H, void* ws, size_t ws_bytes, float** m, float** s, void** rest,
                        size_t* rest_bytes) {
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1, GGL_EINVAL, "%s: bad sizes", who);
  GGL_REQUIRE(ids != nullptr && ((rowptr && perm && plan) || (!rowptr && !perm && !plan)), GGL_EINVAL,
              "%s: ids required; rowptr / perm / plan all set (CSR form) or all NULL (sort-free form)", who);
  const size_t stat = (((size_t)n_rows * H * 4) + 255) & ~(size_t)255;
  GGL_REQUIRE(ws != nullptr && ws_bytes >= 2 * stat, GGL_EWORKSPACE, "%s: workspace %zu < %zu", who, ws_bytes, 2 * stat);
  *m = (float*)ws;
  *s = (float*)((char*)ws + stat);
  *rest_bytes = ws_bytes - 2 * stat;
  *rest = *rest_bytes >= 256 + 12 * (size_t)H ? (char*)ws + 2 * stat : nullptr;
  if (*rest == nullptr) *rest_bytes = 0;
  return GGL_OK;
}

extern "C" int gglb200_edge_softmax_fwd(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const int64_t* ids,
                                        const float* logit, float* alpha, int64_t n_rows, int64_t E, int32_t H,
                                        void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(logit && alpha, GGL_EINVAL, "edge_softmax_fwd: NULL pointer");
  float *m, *s; void* rest; size_t rest_bytes;
  int st = softmax_args("edge_softmax_fwd", rowptr, perm, plan, ids, n_rows, E, H, workspace, workspace_bytes, &m, &s, &rest, &rest_bytes);
  if (st) return st;
  if (rowptr == nullptr) {
    // sort-free form (csrc/oneshot.cu): exact maxima by red.max on ordered keys, denominators by vector reductions, then the same
    // element-wise normalisation; sums are in arrival order (rounding only)
    const bool v4 = H % 4 == 0 && (((uintptr_t)logit | (uintptr_t)alpha) % 16 == 0);
    st = gglb200_edge_softmax_stats_oneshot(ids, logit, m, s, v4 ? nullptr : alpha, n_rows, E, H, stream_);
    if (st) return st;
    if (v4) {
      softmax_norm4_kernel<<<ew_grid(E * H / 4), 256, 0, stream>>>((const float4*)logit, m, s, ids, (float4*)alpha, E * (H / 4), H / 4);
      GGL_LAUNCHED("softmax_norm4");
    } else {
      softmax_div_kernel<<<ew_grid(E * H), 256, 0, stream>>>(alpha, s, ids, E * H, H);
      GGL_LAUNCHED("softmax_div");
    }
    return GGL_OK;
  }
  if (softmax_fused_enabled() && H % 4 == 0 && H <= 128 && (((uintptr_t)logit | (uintptr_t)alpha) % 16 == 0)) {
    // two passes: online (max, denominator) per segment over the bulk-copy ring, then one element-wise normalisation
    st = softmax_stats(rowptr, perm, plan, logit, m, s, n_rows, E, H, rest, rest_bytes, stream);
    if (st) return st;
    softmax_norm4_kernel<<<ew_grid(E * H / 4), 256, 0, stream>>>((const float4*)logit, m, s, ids, (float4*)alpha, E * (H / 4), H / 4);
    GGL_LAUNCHED("softmax_norm4");
    return GGL_OK;
  }
  st = segment_reduce_f32(GGL_MAX, rowptr, perm, plan, logit, m, n_rows, E, H, rest, rest_bytes, stream);
  if (st) return st;
  softmax_exp_kernel<<<ew_grid(E * H), 256, 0, stream>>>(logit, m, ids, alpha, E * H, H);
  GGL_LAUNCHED("softmax_exp");
  st = segment_reduce_f32(GGL_SUM, rowptr, perm, plan, alpha, s, n_rows, E, H, rest, rest_bytes, stream);
  if (st) return st;
  softmax_div_kernel<<<ew_grid(E * H), 256, 0, stream>>>(alpha, s, ids, E * H, H);
  GGL_LAUNCHED("softmax_div");
  return GGL_OK;
}

extern "C" int gglb200_edge_softmax_bwd(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const int64_t* ids,
                                        const float* alpha, const float* galpha, float* glogit, int64_t n_rows, int64_t E,
                                        int32_t H, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows == 0 || E == 0) return GGL_OK;
  GGL_REQUIRE(alpha && galpha && glogit, GGL_EINVAL, "edge_softmax_bwd: NULL pointer");
  float *dot, *unused; void* rest; size_t rest_bytes;
  int st = softmax_args("edge_softmax_bwd", rowptr, perm, plan, ids, n_rows, E, H, workspace, workspace_bytes, &dot, &unused, &rest, &rest_bytes);
  if (st) return st;
  if (rowptr == nullptr) {   // sort-free form
    st = gglb200_edge_softmax_dot_oneshot(ids, alpha, galpha, dot, n_rows, E, H, stream_);
    if (st) return st;
    if (H % 4 == 0 && (((uintptr_t)alpha | (uintptr_t)galpha | (uintptr_t)glogit) % 16 == 0)) {
      softmax_bwd4_kernel<<<ew_grid(E * H / 4), 256, 0, stream>>>((const float4*)alpha, (const float4*)galpha, dot, ids, (float4*)glogit,
                                                                   E * (H / 4), H / 4);
      GGL_LAUNCHED("softmax_bwd4");
    } else {
      softmax_bwd_kernel<<<ew_grid(E * H), 256, 0, stream>>>(alpha, galpha, dot, ids, glogit, E * H, H);
      GGL_LAUNCHED("softmax_bwd");
    }
    return GGL_OK;
  }
  if (softmax_fused_enabled() && H % 4 == 0 && H <= 128 && (((uintptr_t)alpha | (uintptr_t)galpha | (uintptr_t)glogit) % 16 == 0)) {
    // two passes: per-segment dot over the ring (both rows gathered per edge), then one element-wise pass
    st = softmax_dot(rowptr, perm, plan, alpha, galpha, dot, n_rows, E, H, rest, rest_bytes, stream);
    if (st) return st;
    softmax_bwd4_kernel<<<ew_grid(E * H / 4), 256, 0, stream>>>((const float4*)alpha, (const float4*)galpha, dot, ids, (float4*)glogit,
                                                                 E * (H / 4), H / 4);
    GGL_LAUNCHED("softmax_bwd4");
    return GGL_OK;
  }
  softmax_prod_kernel<<<ew_grid(E * H), 256, 0, stream>>>(alpha, galpha, glogit, E * H);
  GGL_LAUNCHED("softmax_prod");
  st = segment_reduce_f32(GGL_SUM, rowptr, perm, plan, glogit, dot, n_rows, E, H, rest, rest_bytes, stream);
  if (st) return st;
  softmax_bwd_kernel<<<ew_grid(E * H), 256, 0, stream>>>(alpha, galpha, dot, ids, glogit, E * H, H);
  GGL_LAUNCHED("softmax_bwd");
  return GGL_OK;
}
