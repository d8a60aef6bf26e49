// gat_fused.cu -- GAT aggregation without any [E, H] tensor.
//
// GATConv (/root/reference/gammagl/layers/conv/gat_conv.py:98-122) computes, per edge e = (j -> i) and head h,
//     pre_e = s_src[j,h] + s_dst[i,h]         (= sum_c concat(x_j, x_i) * att, split into its two node-level halves)
//     l_e = leaky_relu(pre_e);   alpha_e = softmax over the in-edges of i (utils/softmax.py:29-35)
//     out[i,h,:] = sum_e alpha_e * x[j,h,:]
// through 2 gathers of [E,H,2C], 5 passes over [E,H] and 3 passes over [E,H,C].  Here the whole aggregation is ONE
// streaming pass per direction over the CSR, built on the same bulk-copy ring as spmm_ring_kernel:
//   forward   rows = i: gather [x_j | s_src_j] (one bulk copy per edge), ONLINE softmax per (row, head) in registers
//             (running max m, denominator s, accumulator rescaled when the max moves); writes out, m, s.
//   backward  rows = j (transposed CSR): gather [gout_i | {s_dst, m, s, D}_i] with D_i = <out_i, gout_i> per head
//             (the softmax Jacobian's row term, known per node without an edge pass); alpha_e is recomputed,
//             galpha_e = <x_j, gout_i> by an xor-shuffle inside the head; accumulates grad x_j and grad s_src_j.
//             rows = i (forward CSR): gather [x_j | s_src_j] again, accumulates grad s_dst_i.
// Rows above GGLB200_LONG_ROW edges are chunked exactly as in spmm: partial (m, s, acc) / partial sums per chunk and
// an ordered combine.  Requires C % 4 == 0, C/4 a power of two, H*C <= 128.  fp32.
#include <cstdlib>

#include "spmm_impl.cuh"

namespace ggl {

struct GatStream {
  const int32_t* rowptr;
  const int32_t* col;
  const int32_t* plan;
  const int32_t* tile_tab;
  int64_t n_tiles;
  int32_t ws_chunks;
  bool has_long;
  const float* src;      // gathered matrix, row stride `src_stride` floats, `rowb` bytes copied per edge
  int64_t src_stride;
  uint32_t rowb;
  int32_t n_rows;
  int32_t stages;        // ring depth (2 or 3)
  const float* src2;     // optional second gathered matrix (same row index), copied behind the first inside a ring row
  int64_t src2_stride;
  uint32_t rowb2;        // 0 when unused
};

// Streams the edges of one tile (or one long-row chunk) through the shared-memory ring and drives a State:
//   st.row_begin(row)            before the first edge of a non-empty row (may load row-side registers)
//   st.edge(smem_row_addr)       for every edge, in CSR order; the gathered row sits at smem_row_addr
//   st.row_end(row, deg)         after the last edge of a row (also called for empty rows, deg == 0)
//   st.chunk_end(chunk, row)     chunk mode: after the last edge of the chunk
template <int LPR, typename State>
__device__ __forceinline__ void ring_stream(const GatStream& a, State& st, unsigned char* smem_base, int32_t pitch) {
  using Cfg = RingCfg<LPR>;
  constexpr int GPW = 32 / LPR, R = Cfg::R;
  const int32_t STAGES = a.stages, LAG = a.stages - 1;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPR - 1);
  const int gbase = lane & ~(LPR - 1);
  const unsigned gmask = group_mask<LPR>(lane);
  const int gib = lane / LPR;
  const int64_t g = (int64_t)blockIdx.x * GPW + gib;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_base);
  const uint32_t ring = smem0 + (uint32_t)(gib * (STAGES * R)) * pitch;
  const uint32_t bars = smem0 + (uint32_t)(GPW * STAGES * R) * pitch + (uint32_t)gib * STAGES * 8;
  if (sub == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(bars + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp(gmask);

  int32_t r0 = 0, r1 = 0, e0, e_end;
  int64_t chunk = -1;
  int32_t chunk_row = 0;
  if (g < a.ws_chunks) {
    if (g >= min(__ldg(a.plan + 0), a.ws_chunks)) return;
    chunk = g;
    const int32_t* chunk_tab = a.plan + plan_chunk_off();
    chunk_row = __ldg(chunk_tab + 2 * g);
    e0 = __ldg(chunk_tab + 2 * g + 1);
    e_end = min(e0 + GGLB200_LONG_CHUNK, __ldg(a.rowptr + chunk_row + 1));
  } else {
    const int64_t tile = g - a.ws_chunks;
    if (tile >= a.n_tiles) return;
    const int2 ta = __ldg(reinterpret_cast<const int2*>(a.tile_tab) + tile);
    const int2 tb = __ldg(reinterpret_cast<const int2*>(a.tile_tab) + tile + 1);
    r0 = ta.x & 0x7fffffff; e0 = ta.y;
    r1 = tb.x & 0x7fffffff; e_end = tb.y;
    if (r0 >= r1) return;
    if (ta.x < 0 && a.has_long) { --r1; e_end = __ldg(a.rowptr + r1); }
    if (r0 >= r1) return;
  }
  RowCursor<LPR> rc;
  rc.init(a.rowptr, r0, r1, e0, sub, gbase, gmask);
  bool started = false;
  if (chunk >= 0) { st.row_begin(chunk_row); started = true; }

  const int32_t n = e_end - e0;
  const int32_t nslots = (n + R - 1) / R;
  auto load_cols = [&](int32_t b) {
    const int32_t pos = b * LPR + sub;
    return pos < n ? ld_stream_i32(a.col + e0 + pos) : 0;
  };
  int32_t A = load_cols(0), B = load_cols(1), C = load_cols(2);
  int32_t bA = 0, istage = 0;
  auto issue = [&](int32_t s) {
    const int32_t stg = istage;
    istage = istage + 1 == STAGES ? 0 : istage + 1;
    if (s < nslots) {
      const int32_t pos0 = s * R;
      const int32_t d = pos0 / LPR - bA;
      const int32_t colsrc = d == 0 ? A : (d == 1 ? B : C);
      const uint32_t dst = ring + (uint32_t)(stg * R) * pitch;
      const uint32_t bar = bars + (uint32_t)stg * 8;
      const int32_t nv = min(R, n - pos0);
      const int32_t c = __shfl_sync(gmask, colsrc, gbase + ((pos0 + sub) & (LPR - 1)));
      __syncwarp(gmask);
      if (sub == 0) mbar_arrive_expect_tx(bar, (uint32_t)nv * (a.rowb + a.rowb2));
      if (sub < nv) {
        bulk_copy_g2s(dst + (uint32_t)sub * pitch, a.src + (int64_t)c * a.src_stride, a.rowb, bar);
        if (a.rowb2) bulk_copy_g2s(dst + (uint32_t)sub * pitch + a.rowb, a.src2 + (int64_t)c * a.src2_stride, a.rowb2, bar);
      }
    }
  };
  for (int s = 0; s < LAG; ++s) issue(s);
  int32_t cstage = 0;
  uint32_t cparity = 0;
  for (int32_t s = 0; s < nslots; ++s) {
    const int32_t pos0 = s * R;
    if (pos0 / LPR != bA) { A = B; B = C; ++bA; C = load_cols(bA + 2); }
    issue(s + LAG);
    mbar_wait(bars + (uint32_t)cstage * 8, cparity);
    const uint32_t src = ring + (uint32_t)(cstage * R) * pitch;
    if (++cstage == STAGES) { cstage = 0; cparity ^= 1u; }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int32_t pos = pos0 + u;
      if (pos < n) {  // group-uniform
        if (chunk < 0) {
          while (e0 + pos == rc.cur_end) {
            st.row_end(rc.r, rc.cur_end - rc.row_start);
            rc.advance(a.rowptr, sub, gbase, gmask);
            started = false;
          }
          if (!started) { st.row_begin(rc.r); started = true; }
        }
        st.edge(src + (uint32_t)u * pitch);
      }
    }
  }
  if (chunk >= 0) {
    st.chunk_end(chunk, chunk_row);
  } else {
    while (rc.r < rc.r_end) {
      st.row_end(rc.r, rc.cur_end - rc.row_start);
      rc.advance(a.rowptr, sub, gbase, gmask);
    }
  }
}

__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : slope * v; }

constexpr float kNegBig = -1e30f;
constexpr int kFB = 8;  // chunk partials fetched together by the finalize kernels (loads in flight on a hub row's chunk list)

struct GatDims {
  int32_t H, C, K, LH;   // LH = C/4 lanes per head
  int32_t W;             // floats per chunk partial of the running kernel (multiple of 4: float4 stores)
  float slope;
};

// ---- forward ----------------------------------------------------------------------------------------------------------
struct GatFwdState {
  // lane geometry
  int32_t kcol, head; bool active, writer; int sub; unsigned gmask;
  GatDims d;
  const float* s_dst; float* out; float* stat_m; float* stat_s;
  float* part;  // chunk partials: [chunk][K + 2H] = acc | m | s
  // row state
  float sd, sd_next; int32_t next_row;
  float m, s; float4 acc;
  __device__ __forceinline__ void reset() { m = kNegBig; s = 0.f; acc = make_float4(0, 0, 0, 0); }
  __device__ __forceinline__ void row_begin(int32_t row) {
    sd = (next_row == row) ? sd_next : (active ? __ldg(s_dst + (int64_t)row * d.H + head) : 0.f);
    next_row = row + 1;
    sd_next = active ? __ldg(s_dst + (int64_t)min(next_row, d_rows - 1) * d.H + head) : 0.f;
    reset();
  }
  int32_t d_rows;
  __device__ __forceinline__ void edge(uint32_t row_addr) {
    if (!active) return;
    const float4 v = lds4(row_addr + (uint32_t)kcol * 4);
    const float ss = lds1(row_addr + (uint32_t)(d.K + head) * 4);
    const float l = lrelu(ss + sd, d.slope);
    const float mn = fmaxf(m, l);
    const float sc = __expf(m - mn), p = __expf(l - mn);
    acc.x = fmaf(acc.x, sc, p * v.x); acc.y = fmaf(acc.y, sc, p * v.y);
    acc.z = fmaf(acc.z, sc, p * v.z); acc.w = fmaf(acc.w, sc, p * v.w);
    s = fmaf(s, sc, p);
    m = mn;
  }
  __device__ __forceinline__ void row_end(int32_t row, int32_t deg) {
    if (!active) return;
    float4 o = make_float4(0, 0, 0, 0);
    if (deg > 0) { const float inv = 1.f / s; o = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv); }
    *reinterpret_cast<float4*>(out + (int64_t)row * d.K + kcol) = o;
    if (writer) {
      stat_m[(int64_t)row * d.H + head] = deg > 0 ? m : 0.f;
      stat_s[(int64_t)row * d.H + head] = deg > 0 ? s : 1.f;
    }
    reset();
  }
  __device__ __forceinline__ void chunk_end(int64_t chunk, int32_t) {
    if (!active) return;
    float* pc = part + chunk * d.W;
    *reinterpret_cast<float4*>(pc + kcol) = acc;
    if (writer) { pc[d.K + head] = m; pc[d.K + d.H + head] = s; }
  }
};

template <int LPR>
__global__ void __launch_bounds__(32) gat_fwd_kernel(const GatStream a, const GatDims d, const float* __restrict__ s_dst,
                                                     float* __restrict__ out, float* __restrict__ stat_m,
                                                     float* __restrict__ stat_s, float* __restrict__ part, int32_t pitch) {
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
  GatFwdState st;
  st.kcol = sub * 4; st.active = st.kcol < d.K; st.head = st.active ? st.kcol / d.C : 0;
  st.writer = st.active && (sub & (d.LH - 1)) == 0; st.sub = sub; st.gmask = group_mask<LPR>(lane);
  st.d = d; st.s_dst = s_dst; st.out = out; st.stat_m = stat_m; st.stat_s = stat_s; st.part = part;
  st.next_row = -1; st.sd = st.sd_next = 0.f; st.d_rows = a.n_rows;
  st.reset();
  ring_stream<LPR>(a, st, ring_smem, pitch);
}

// ordered combine of the (m, s, acc) partials of each long row
__global__ void __launch_bounds__(256) gat_fwd_finalize_kernel(const int32_t* __restrict__ plan, const float* __restrict__ part,
                                                               float* __restrict__ out, float* __restrict__ stat_m,
                                                               float* __restrict__ stat_s, GatDims d, int32_t ws_chunks) {
  const int32_t max_chunks = plan[2];
  const int32_t n_long = min(plan[1], plan[3]);
  const int32_t* long_tab = plan + plan_long_off(max_chunks);
  const int32_t W = d.W;
  for (int32_t l = blockIdx.x; l < n_long; l += gridDim.x) {
    const int32_t row = long_tab[3 * l], c0 = long_tab[3 * l + 1], nch = long_tab[3 * l + 2];
    if (c0 + nch > ws_chunks) continue;
    for (int32_t k = threadIdx.x; k < d.K; k += blockDim.x) {
      const int32_t h = k / d.C;
      // partials of kFB chunks are fetched together (independent loads), then combined in ascending chunk order
      float M = kNegBig;
      for (int32_t j0 = 0; j0 < nch; j0 += kFB) {
        float mv[kFB];
#pragma unroll
        for (int u = 0; u < kFB; ++u) mv[u] = j0 + u < nch ? part[(int64_t)(c0 + j0 + u) * W + d.K + h] : kNegBig;
#pragma unroll
        for (int u = 0; u < kFB; ++u) M = fmaxf(M, mv[u]);
      }
      float S = 0.f, A = 0.f;
      for (int32_t j0 = 0; j0 < nch; j0 += kFB) {
        float mv[kFB], sv[kFB], av[kFB];
#pragma unroll
        for (int u = 0; u < kFB; ++u) {
          const bool ok = j0 + u < nch;
          const float* pc = part + (int64_t)(c0 + (ok ? j0 + u : 0)) * W;
          mv[u] = ok ? pc[d.K + h] : kNegBig; sv[u] = ok ? pc[d.K + d.H + h] : 0.f; av[u] = ok ? pc[k] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kFB; ++u) {
          if (j0 + u < nch) {
            const float f = __expf(mv[u] - M);
            S = fmaf(sv[u], f, S);
            A = fmaf(av[u], f, A);
          }
        }
      }
      out[(int64_t)row * d.K + k] = A / S;
      if (k % d.C == 0) { stat_m[(int64_t)row * d.H + h] = M; stat_s[(int64_t)row * d.H + h] = S; }
    }
  }
}

// ---- backward, rows = source nodes (transposed CSR): grad x and grad s_src -----------------------------------------------
// gathered row (per destination i): [gout_i (K) | {s_dst, m, 1/s, D}_i per head (4H)]
struct GatBwdSrcState {
  int32_t kcol, head; bool active, writer; int sub; unsigned gmask;
  GatDims d;
  const float* xcat; int64_t xstride; int32_t n_rows;
  float* gx; float* gs_src; float* part;  // partials: [chunk][K + H]
  float* gs_dst;      // merged backward: grad s_dst[i, h] accumulated by L2 reductions from this pass (NULL = separate pass)
  int32_t id_off;     // merged backward: float offset of the destination id stored behind the per-head statistics of a gathered row
  float4 xs, xs_next; float ss, ss_next; int32_t next_row;
  float4 acc; float gss;
  __device__ __forceinline__ void reset() { acc = make_float4(0, 0, 0, 0); gss = 0.f; }
  __device__ __forceinline__ void load_row(int32_t row, float4& x4, float& s1) {
    x4 = make_float4(0, 0, 0, 0); s1 = 0.f;
    if (active && row < n_rows) {
      x4 = ldg4(xcat + (int64_t)row * xstride + kcol);
      s1 = __ldg(xcat + (int64_t)row * xstride + d.K + head);
    }
  }
  __device__ __forceinline__ void row_begin(int32_t row) {
    if (next_row == row) { xs = xs_next; ss = ss_next; } else load_row(row, xs, ss);
    next_row = row + 1;
    load_row(next_row, xs_next, ss_next);
    reset();
  }
  __device__ __forceinline__ void edge(uint32_t row_addr) {
    float dotp = 0.f, alpha = 0.f, D = 0.f, pre = 0.f;
    float4 g4 = make_float4(0, 0, 0, 0);
    if (active) {
      g4 = lds4(row_addr + (uint32_t)kcol * 4);
      const float4 t = lds4(row_addr + (uint32_t)(d.K + 4 * head) * 4);  // {s_dst, m, s, D}
      pre = ss + t.x;
      alpha = __expf(lrelu(pre, d.slope) - t.y) * t.z;   // t.z = 1 / s
      D = t.w;
      dotp = fmaf(xs.w, g4.w, fmaf(xs.z, g4.z, fmaf(xs.y, g4.y, xs.x * g4.x)));
    }
    for (int off = d.LH >> 1; off > 0; off >>= 1) dotp += __shfl_xor_sync(gmask, dotp, off);
    if (active) {
      acc.x = fmaf(alpha, g4.x, acc.x); acc.y = fmaf(alpha, g4.y, acc.y);
      acc.z = fmaf(alpha, g4.z, acc.z); acc.w = fmaf(alpha, g4.w, acc.w);
      const float ge = alpha * (dotp - D) * (pre > 0.f ? 1.f : d.slope);   // d loss / d pre_e: the same scalar feeds s_src[j] and s_dst[i]
      gss += ge;
      if (gs_dst != nullptr && writer) {
        // merged backward: grad s_dst[i, h] += ge by a reduction in the L2 (one scalar per head and edge; [N, H] stays L2-resident),
        // instead of a second streaming pass over the forward CSR.  i rides in the gathered row behind the statistics.
        const int32_t i = __float_as_int(lds1(row_addr + (uint32_t)id_off * 4));
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gs_dst + (int64_t)i * d.H + head), "f"(ge) : "memory");
      }
    }
  }
  __device__ __forceinline__ void row_end(int32_t row, int32_t) {
    if (!active) return;
    *reinterpret_cast<float4*>(gx + (int64_t)row * d.K + kcol) = acc;
    if (writer) gs_src[(int64_t)row * d.H + head] = gss;
    reset();
  }
  __device__ __forceinline__ void chunk_end(int64_t chunk, int32_t) {
    if (!active) return;
    float* pc = part + chunk * d.W;
    *reinterpret_cast<float4*>(pc + kcol) = acc;
    if (writer) pc[d.K + head] = gss;
  }
};

template <int LPR>
__global__ void __launch_bounds__(32) gat_bwd_src_kernel(const GatStream a, const GatDims d, const float* __restrict__ xcat,
                                                         int64_t xstride, float* __restrict__ gx, float* __restrict__ gs_src,
                                                         float* __restrict__ part, float* __restrict__ gs_dst, int32_t id_off,
                                                         int32_t pitch) {
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
  GatBwdSrcState st;
  st.gs_dst = gs_dst; st.id_off = id_off;
  st.kcol = sub * 4; st.active = st.kcol < d.K; st.head = st.active ? st.kcol / d.C : 0;
  st.writer = st.active && (sub & (d.LH - 1)) == 0; st.sub = sub; st.gmask = group_mask<LPR>(lane);
  st.d = d; st.xcat = xcat; st.xstride = xstride; st.n_rows = a.n_rows; st.gx = gx; st.gs_src = gs_src; st.part = part;
  st.next_row = -1; st.xs = st.xs_next = make_float4(0, 0, 0, 0); st.ss = st.ss_next = 0.f;
  st.reset();
  ring_stream<LPR>(a, st, ring_smem, pitch);
}

// ---- backward, rows = destination nodes (forward CSR): grad s_dst ---------------------------------------------------------
// gathered row (per source j): [x_j (K) | s_src_j (H..)];  row side: gout_i piece + {s_dst, m, s, D}_i
struct GatBwdDstState {
  int32_t kcol, head; bool active, writer; int sub; unsigned gmask;
  GatDims d;
  const float* gcat; int64_t gstride; int32_t n_rows;
  float* gs_dst; float* part;  // partials: [chunk][H]
  float4 g4, g4_next, t, t_next; int32_t next_row;
  float gsd;
  __device__ __forceinline__ void load_row(int32_t row, float4& g, float4& tt) {
    g = make_float4(0, 0, 0, 0); tt = make_float4(0, 0, 1.f, 0);
    if (active && row < n_rows) {
      g = ldg4(gcat + (int64_t)row * gstride + kcol);
      tt = ldg4(gcat + (int64_t)row * gstride + d.K + 4 * head);
    }
  }
  __device__ __forceinline__ void row_begin(int32_t row) {
    if (next_row == row) { g4 = g4_next; t = t_next; } else load_row(row, g4, t);
    next_row = row + 1;
    load_row(next_row, g4_next, t_next);
    gsd = 0.f;
  }
  __device__ __forceinline__ void edge(uint32_t row_addr) {
    float dotp = 0.f, alpha = 0.f, pre = 0.f;
    if (active) {
      const float4 x4 = lds4(row_addr + (uint32_t)kcol * 4);
      const float ss = lds1(row_addr + (uint32_t)(d.K + head) * 4);
      pre = ss + t.x;
      alpha = __expf(lrelu(pre, d.slope) - t.y) * t.z;   // t.z = 1 / s
      dotp = fmaf(x4.w, g4.w, fmaf(x4.z, g4.z, fmaf(x4.y, g4.y, x4.x * g4.x)));
    }
    for (int off = d.LH >> 1; off > 0; off >>= 1) dotp += __shfl_xor_sync(gmask, dotp, off);
    if (active) gsd = fmaf(alpha * (dotp - t.w), pre > 0.f ? 1.f : d.slope, gsd);
  }
  __device__ __forceinline__ void row_end(int32_t row, int32_t) {
    if (writer) gs_dst[(int64_t)row * d.H + head] = gsd;
    gsd = 0.f;
  }
  __device__ __forceinline__ void chunk_end(int64_t chunk, int32_t) {
    if (writer) part[chunk * d.W + head] = gsd;
  }
};

template <int LPR>
__global__ void __launch_bounds__(32) gat_bwd_dst_kernel(const GatStream a, const GatDims d, const float* __restrict__ gcat,
                                                         int64_t gstride, float* __restrict__ gs_dst, float* __restrict__ part,
                                                         int32_t pitch) {
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
  GatBwdDstState st;
  st.kcol = sub * 4; st.active = st.kcol < d.K; st.head = st.active ? st.kcol / d.C : 0;
  st.writer = st.active && (sub & (d.LH - 1)) == 0; st.sub = sub; st.gmask = group_mask<LPR>(lane);
  st.d = d; st.gcat = gcat; st.gstride = gstride; st.n_rows = a.n_rows; st.gs_dst = gs_dst; st.part = part;
  st.next_row = -1; st.g4 = st.g4_next = make_float4(0, 0, 0, 0); st.t = st.t_next = make_float4(0, 0, 1.f, 0);
  st.gsd = 0.f;
  ring_stream<LPR>(a, st, ring_smem, pitch);
}

// sums the per-chunk partial rows (width W) of each long row into out rows: out[row, :wa] (stride sa) and out2[row, :W-wa]
__global__ void __launch_bounds__(256) gat_sum_finalize_kernel(const int32_t* __restrict__ plan, const float* __restrict__ part,
                                                               int32_t W /* used floats */, int32_t Wp /* partial stride */,
                                                               float* __restrict__ out_a, int32_t wa, int32_t sa,
                                                               float* __restrict__ out_b, int32_t sb, int32_t ws_chunks) {
  const int32_t max_chunks = plan[2];
  const int32_t n_long = min(plan[1], plan[3]);
  const int32_t* long_tab = plan + plan_long_off(max_chunks);
  for (int32_t l = blockIdx.x; l < n_long; l += gridDim.x) {
    const int32_t row = long_tab[3 * l], c0 = long_tab[3 * l + 1], nch = long_tab[3 * l + 2];
    if (c0 + nch > ws_chunks) continue;
    for (int32_t k = threadIdx.x; k < W; k += blockDim.x) {
      float S = 0.f;
      for (int32_t j0 = 0; j0 < nch; j0 += kFB) {
        float v[kFB];
#pragma unroll
        for (int u = 0; u < kFB; ++u) v[u] = j0 + u < nch ? part[(int64_t)(c0 + j0 + u) * Wp + k] : 0.f;
#pragma unroll
        for (int u = 0; u < kFB; ++u)
          if (j0 + u < nch) S += v[u];
      }
      if (k < wa) out_a[(int64_t)row * sa + k] = S;
      else out_b[(int64_t)row * sb + (k - wa)] = S;
    }
  }
}

static int gat_setup(const char* who, const int32_t* rowptr, const int32_t* col, const int32_t* plan, int64_t n_rows, int64_t E,
                     int32_t H, int32_t C, int32_t width /* floats per chunk partial */, void* ws, size_t ws_bytes, GatStream* a,
                     GatDims* d, int* lpr) {
  GGL_REQUIRE(n_rows >= 0 && E >= 0 && H >= 1 && C >= 4, GGL_EINVAL, "%s: bad sizes", who);
  GGL_REQUIRE(n_rows < INT32_MAX && E < INT32_MAX, GGL_ERANGE, "%s: n_rows/E exceed int32", who);
  GGL_REQUIRE(rowptr && col && plan, GGL_EINVAL, "%s: NULL CSR pointer", who);
  const int lh = C / 4;
  GGL_REQUIRE(C % 4 == 0 && (lh & (lh - 1)) == 0 && (int64_t)H * C <= 128, GGL_EUNSUPPORTED,
              "%s: needs C %% 4 == 0, C/4 a power of two and H*C <= 128 (H=%d C=%d)", who, H, C);
  d->H = H; d->C = C; d->K = H * C; d->LH = lh; d->W = (width + 3) & ~3;
  int l = lpr_for(d->K, 4);
  if (l < 8) l = 8;
  if (l < lh) l = lh;
  *lpr = l;
  a->rowptr = rowptr; a->col = col; a->plan = plan;
  a->src2 = nullptr; a->src2_stride = 0; a->rowb2 = 0;
  a->tile_tab = plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E));
  a->n_tiles = plan_n_tiles(E);
  a->n_rows = (int32_t)n_rows;
  a->has_long = ws != nullptr && E > GGLB200_LONG_ROW;
  {
    const char* e = getenv("GGLB200_GAT_STAGES");
    a->stages = (e && e[0] == '3') ? 3 : 2;   // 2 slots: more resident one-warp blocks (C4: fwd 9.7 -> 7.6 ms, bwd 29.1 -> 21.8 ms)
  }
  a->ws_chunks = 0;
  if (a->has_long) {
    int64_t fit = (int64_t)(ws_bytes / ((size_t)d->W * 4));
    if (fit > plan_max_chunks(E)) fit = plan_max_chunks(E);
    GGL_REQUIRE(fit >= 1, GGL_EWORKSPACE, "%s: workspace holds no chunk", who);
    a->ws_chunks = (int32_t)fit;
  }
  return GGL_OK;
}

template <typename Kern, typename... Args>
static int gat_launch(Kern kern, int lpr, const GatStream& a, size_t rowbytes, cudaStream_t stream, Args... args) {
  const int gpw = 32 / lpr;
  const int R = lpr >= 8 ? 8 : lpr, STAGES = a.stages;  // RingCfg<LPR>::R
  const int32_t pitch = (int32_t)rowbytes;
  const size_t smem = (size_t)gpw * STAGES * R * pitch + (size_t)gpw * STAGES * 8;
  GGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  const int64_t groups = a.n_tiles + a.ws_chunks;
  const int64_t blocks = (groups + gpw - 1) / gpw;
  kern<<<(unsigned)blocks, 32, smem, stream>>>(a, args..., pitch);
  return GGL_OK;
}

// ---- stand-alone edge softmax: per-(segment, head) max and denominator in ONE streaming pass ------------------------------
// utils/softmax.py:29-33 runs unsorted_segment_max, exp and unsorted_segment_sum as three passes over [E,H]; here the
// logit rows (H floats) of each segment stream through the same bulk-copy ring (col = perm of the segment CSR) and an
// online softmax per lane keeps (m, s) for 4 heads.  Chunk partials are (m, s) pairs combined by the finalize kernel.
struct SoftmaxStatsState {
  int32_t kcol, H; bool active;
  float* stat_m; float* stat_s; float* part;  // partials: [chunk][2H] = m | s
  float4 m, s;
  __device__ __forceinline__ void reset() { m = make_float4(kNegBig, kNegBig, kNegBig, kNegBig); s = make_float4(0, 0, 0, 0); }
  __device__ __forceinline__ void row_begin(int32_t) { reset(); }
  __device__ __forceinline__ static void upd(float& m_, float& s_, float v) {
    const float mn = fmaxf(m_, v);
    s_ = fmaf(s_, __expf(m_ - mn), __expf(v - mn));
    m_ = mn;
  }
  __device__ __forceinline__ void edge(uint32_t row_addr) {
    if (!active) return;
    const float4 v = lds4(row_addr + (uint32_t)kcol * 4);
    upd(m.x, s.x, v.x); upd(m.y, s.y, v.y); upd(m.z, s.z, v.z); upd(m.w, s.w, v.w);
  }
  __device__ __forceinline__ void row_end(int32_t row, int32_t deg) {
    if (!active) return;
    const bool any = deg > 0;
    *reinterpret_cast<float4*>(stat_m + (int64_t)row * H + kcol) = any ? m : make_float4(0, 0, 0, 0);
    *reinterpret_cast<float4*>(stat_s + (int64_t)row * H + kcol) = any ? s : make_float4(1.f, 1.f, 1.f, 1.f);
    reset();
  }
  __device__ __forceinline__ void chunk_end(int64_t chunk, int32_t) {
    if (!active) return;
    float* pc = part + chunk * (2 * (int64_t)H);
    *reinterpret_cast<float4*>(pc + kcol) = m;
    *reinterpret_cast<float4*>(pc + H + kcol) = s;
  }
};

template <int LPR>
__global__ void __launch_bounds__(32) softmax_stats_kernel(const GatStream a, int32_t H, float* __restrict__ stat_m,
                                                           float* __restrict__ stat_s, float* __restrict__ part, int32_t pitch) {
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
  SoftmaxStatsState st;
  st.kcol = sub * 4; st.H = H; st.active = st.kcol < H;
  st.stat_m = stat_m; st.stat_s = stat_s; st.part = part;
  st.reset();
  ring_stream<LPR>(a, st, ring_smem, pitch);
}

__global__ void __launch_bounds__(128) softmax_stats_finalize_kernel(const int32_t* __restrict__ plan, const float* __restrict__ part,
                                                                     float* __restrict__ stat_m, float* __restrict__ stat_s,
                                                                     int32_t H, int32_t ws_chunks) {
  const int32_t max_chunks = plan[2];
  const int32_t n_long = min(plan[1], plan[3]);
  const int32_t* long_tab = plan + plan_long_off(max_chunks);
  for (int32_t l = blockIdx.x; l < n_long; l += gridDim.x) {
    const int32_t row = long_tab[3 * l], c0 = long_tab[3 * l + 1], nch = long_tab[3 * l + 2];
    if (c0 + nch > ws_chunks) continue;
    for (int32_t h = threadIdx.x; h < H; h += blockDim.x) {
      float M = kNegBig;
      for (int32_t j0 = 0; j0 < nch; j0 += kFB) {
        float mv[kFB];
#pragma unroll
        for (int u = 0; u < kFB; ++u) mv[u] = j0 + u < nch ? part[(int64_t)(c0 + j0 + u) * 2 * H + h] : kNegBig;
#pragma unroll
        for (int u = 0; u < kFB; ++u) M = fmaxf(M, mv[u]);
      }
      float S = 0.f;
      for (int32_t j0 = 0; j0 < nch; j0 += kFB) {
        float mv[kFB], sv[kFB];
#pragma unroll
        for (int u = 0; u < kFB; ++u) {
          const bool ok = j0 + u < nch;
          const float* pc = part + (int64_t)(c0 + (ok ? j0 + u : 0)) * 2 * H;
          mv[u] = ok ? pc[h] : kNegBig; sv[u] = ok ? pc[H + h] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kFB; ++u)
          if (j0 + u < nch) S = fmaf(sv[u], __expf(mv[u] - M), S);
      }
      stat_m[(int64_t)row * H + h] = M;
      stat_s[(int64_t)row * H + h] = S;
    }
  }
}

// stat_m / stat_s [n_rows, H] of logit[E, H] over the segment CSR (rowptr, perm, plan).  H % 4 == 0, H <= 128, logit 16-byte aligned.
int softmax_stats(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const float* logit, float* stat_m, float* stat_s,
                  int64_t n_rows, int64_t E, int32_t H, void* part_ws, size_t part_bytes, cudaStream_t stream) {
  GatStream a;
  a.rowptr = rowptr; a.col = perm; a.plan = plan;
  a.tile_tab = plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E));
  a.n_tiles = plan_n_tiles(E);
  a.n_rows = (int32_t)n_rows;
  a.stages = 2;
  a.src = logit; a.src_stride = H; a.rowb = (uint32_t)H * 4;
  a.src2 = nullptr; a.src2_stride = 0; a.rowb2 = 0;
  a.ws_chunks = 0;
  a.has_long = part_ws != nullptr && E > GGLB200_LONG_ROW;
  if (a.has_long) {
    int64_t fit = (int64_t)(part_bytes / ((size_t)H * 8));
    if (fit > plan_max_chunks(E)) fit = plan_max_chunks(E);
    if (fit < 1) a.has_long = false; else a.ws_chunks = (int32_t)fit;
  }
  float* part = (float*)part_ws;
  const size_t rowbytes = (size_t)H * 4;
  int st;
  // lanes per segment row = float4 words per logit row (at least 2): H = 8 runs 16 segments per warp side by side
  if (H <= 8) st = gat_launch(softmax_stats_kernel<2>, 2, a, rowbytes, stream, H, stat_m, stat_s, part);
  else if (H <= 16) st = gat_launch(softmax_stats_kernel<4>, 4, a, rowbytes, stream, H, stat_m, stat_s, part);
  else if (H <= 32) st = gat_launch(softmax_stats_kernel<8>, 8, a, rowbytes, stream, H, stat_m, stat_s, part);
  else if (H <= 64) st = gat_launch(softmax_stats_kernel<16>, 16, a, rowbytes, stream, H, stat_m, stat_s, part);
  else st = gat_launch(softmax_stats_kernel<32>, 32, a, rowbytes, stream, H, stat_m, stat_s, part);
  if (st) return st;
  GGL_LAUNCHED("softmax_stats");
  if (a.has_long) {
    softmax_stats_finalize_kernel<<<sm_count() * 2, 128, 0, stream>>>(plan, part, stat_m, stat_s, H, a.ws_chunks);
    GGL_LAUNCHED("softmax_stats_finalize");
  }
  return GGL_OK;
}

// ---- edge softmax backward: dot[r, h] = sum over the segment of alpha * galpha, both rows gathered per edge -------------------
struct SoftmaxDotState {
  int32_t kcol, H; bool active;
  float* dot; float* part;  // partials: [chunk][H]
  float4 acc;
  __device__ __forceinline__ void row_begin(int32_t) { acc = make_float4(0, 0, 0, 0); }
  __device__ __forceinline__ void edge(uint32_t row_addr) {
    if (!active) return;
    const float4 a = lds4(row_addr + (uint32_t)kcol * 4);
    const float4 g = lds4(row_addr + (uint32_t)(H + kcol) * 4);
    acc.x = fmaf(a.x, g.x, acc.x); acc.y = fmaf(a.y, g.y, acc.y); acc.z = fmaf(a.z, g.z, acc.z); acc.w = fmaf(a.w, g.w, acc.w);
  }
  __device__ __forceinline__ void row_end(int32_t row, int32_t) {
    if (!active) return;
    *reinterpret_cast<float4*>(dot + (int64_t)row * H + kcol) = acc;
    acc = make_float4(0, 0, 0, 0);
  }
  __device__ __forceinline__ void chunk_end(int64_t chunk, int32_t) {
    if (!active) return;
    *reinterpret_cast<float4*>(part + chunk * (int64_t)H + kcol) = acc;
  }
};

template <int LPR>
__global__ void __launch_bounds__(32) softmax_dot_kernel(const GatStream a, int32_t H, float* __restrict__ dot,
                                                         float* __restrict__ part, int32_t pitch) {
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
  SoftmaxDotState st;
  st.kcol = sub * 4; st.H = H; st.active = st.kcol < H;
  st.dot = dot; st.part = part;
  st.acc = make_float4(0, 0, 0, 0);
  ring_stream<LPR>(a, st, ring_smem, pitch);
}

// dot[n_rows, H] = segment sums of alpha * galpha ([E, H] each) over the segment CSR.  H % 4 == 0, H <= 128, 16-byte aligned.
int softmax_dot(const int32_t* rowptr, const int32_t* perm, const int32_t* plan, const float* alpha, const float* galpha, float* dot,
                int64_t n_rows, int64_t E, int32_t H, void* part_ws, size_t part_bytes, cudaStream_t stream) {
  GatStream a;
  a.rowptr = rowptr; a.col = perm; a.plan = plan;
  a.tile_tab = plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E));
  a.n_tiles = plan_n_tiles(E);
  a.n_rows = (int32_t)n_rows;
  a.stages = 2;
  a.src = alpha; a.src_stride = H; a.rowb = (uint32_t)H * 4;
  a.src2 = galpha; a.src2_stride = H; a.rowb2 = (uint32_t)H * 4;
  a.ws_chunks = 0;
  a.has_long = part_ws != nullptr && E > GGLB200_LONG_ROW;
  if (a.has_long) {
    int64_t fit = (int64_t)(part_bytes / ((size_t)H * 4));
    if (fit > plan_max_chunks(E)) fit = plan_max_chunks(E);
    if (fit < 1) a.has_long = false; else a.ws_chunks = (int32_t)fit;
  }
  float* part = (float*)part_ws;
  const size_t rowbytes = (size_t)H * 8;
  int st;
  if (H <= 8) st = gat_launch(softmax_dot_kernel<2>, 2, a, rowbytes, stream, H, dot, part);
  else if (H <= 16) st = gat_launch(softmax_dot_kernel<4>, 4, a, rowbytes, stream, H, dot, part);
  else if (H <= 32) st = gat_launch(softmax_dot_kernel<8>, 8, a, rowbytes, stream, H, dot, part);
  else if (H <= 64) st = gat_launch(softmax_dot_kernel<16>, 16, a, rowbytes, stream, H, dot, part);
  else st = gat_launch(softmax_dot_kernel<32>, 32, a, rowbytes, stream, H, dot, part);
  if (st) return st;
  GGL_LAUNCHED("softmax_dot");
  if (a.has_long) {
    gat_sum_finalize_kernel<<<sm_count() * 2, 256, 0, stream>>>(plan, part, H, H, dot, H, H, nullptr, 0, a.ws_chunks);
    GGL_LAUNCHED("gat_sum_finalize");
  }
  return GGL_OK;
}

}  // namespace ggl

using namespace ggl;

extern "C" size_t gglb200_gat_workspace_bytes(int64_t n_chunks, int32_t H, int32_t C) {
  if (n_chunks <= 0) return 0;
  return (size_t)n_chunks * (size_t)((H * C + 2 * H + 3) & ~3) * 4 + 256;
}

#define GGL_GAT_DISPATCH(KERN, ...)                                                       \
  do {                                                                                    \
    int st__;                                                                             \
    switch (lpr) {                                                                        \
      case 8: st__ = gat_launch(KERN<8>, 8, a, rowbytes, stream, __VA_ARGS__); break;     \
      case 16: st__ = gat_launch(KERN<16>, 16, a, rowbytes, stream, __VA_ARGS__); break;  \
      default: st__ = gat_launch(KERN<32>, 32, a, rowbytes, stream, __VA_ARGS__); break;  \
    }                                                                                     \
    if (st__) return st__;                                                                \
  } while (0)

extern "C" int gglb200_gat_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* plan, const float* xcat,
                               int64_t xcat_stride, const float* s_dst, float slope, float* out, float* stat_m, float* stat_s,
                               int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace, size_t workspace_bytes,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows == 0) return GGL_OK;
  GatStream a; GatDims d; int lpr;
  int st = gat_setup("gat_fwd", rowptr, col, plan, n_rows, E, H, C, H * C + 2 * H, workspace, workspace_bytes, &a, &d, &lpr);
  if (st) return st;
  GGL_REQUIRE(xcat && s_dst && out && stat_m && stat_s, GGL_EINVAL, "gat_fwd: NULL pointer");
  GGL_REQUIRE(xcat_stride >= d.K + H && xcat_stride % 4 == 0 && ((uintptr_t)xcat | (uintptr_t)out) % 16 == 0, GGL_EINVAL,
              "gat_fwd: xcat rows must be [x | s_src | pad] with a stride that is a multiple of 4 floats, 16-byte aligned");
  d.slope = slope;
  a.src = xcat; a.src_stride = xcat_stride;
  const size_t rowbytes = (size_t)xcat_stride * 4;
  a.rowb = (uint32_t)rowbytes;
  float* part = (float*)workspace;
  GGL_GAT_DISPATCH(gat_fwd_kernel, d, s_dst, out, stat_m, stat_s, part);
  GGL_LAUNCHED("gat_fwd");
  if (a.has_long) {
    gat_fwd_finalize_kernel<<<sm_count() * 2, 256, 0, stream>>>(plan, part, out, stat_m, stat_s, d, a.ws_chunks);
    GGL_LAUNCHED("gat_fwd_finalize");
  }
  return GGL_OK;
}

extern "C" int gglb200_gat_bwd_src(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* plan_t, const float* gcat,
                                   int64_t gcat_stride, const float* xcat, int64_t xcat_stride, float slope, float* gx,
                                   float* gs_src, int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows == 0) return GGL_OK;
  GatStream a; GatDims d; int lpr;
  int st = gat_setup("gat_bwd_src", rowptr_t, col_t, plan_t, n_rows, E, H, C, H * C + H, workspace, workspace_bytes, &a, &d, &lpr);
  if (st) return st;
  GGL_REQUIRE(gcat && xcat && gx && gs_src, GGL_EINVAL, "gat_bwd_src: NULL pointer");
  GGL_REQUIRE(gcat_stride == d.K + 4 * H && xcat_stride % 4 == 0 && ((uintptr_t)gcat | (uintptr_t)xcat | (uintptr_t)gx) % 16 == 0,
              GGL_EINVAL, "gat_bwd_src: gcat rows must be [gout | {s_dst,m,s,D} x H], 16-byte aligned");
  d.slope = slope;
  a.src = gcat; a.src_stride = gcat_stride;
  const size_t rowbytes = (size_t)gcat_stride * 4;
  a.rowb = (uint32_t)rowbytes;
  float* part = (float*)workspace;
  GGL_GAT_DISPATCH(gat_bwd_src_kernel, d, xcat, xcat_stride, gx, gs_src, part, (float*)nullptr, 0);
  GGL_LAUNCHED("gat_bwd_src");
  if (a.has_long) {
    gat_sum_finalize_kernel<<<sm_count() * 2, 256, 0, stream>>>(plan_t, part, d.K + H, d.W, gx, d.K, d.K, gs_src, H, a.ws_chunks);
    GGL_LAUNCHED("gat_sum_finalize");
  }
  return GGL_OK;
}

// The whole backward in ONE streaming pass over the transposed CSR: as gglb200_gat_bwd_src, and every edge also adds its
// d loss / d pre_e to grad s_dst[i, h] with a reduction in the L2 (red.global.add.f32; gs_dst [N, H] is zero-filled here).  gcat rows
// carry the destination id behind the statistics: [gout (K) | {s_dst, m, 1/s, D} x H | int32 row id, 0, 0, 0].  grad s_dst is summed
// in arrival order (rounding differs run to run in the last bits); grad x and grad s_src are those of the two-pass form.
extern "C" int gglb200_gat_bwd(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* plan_t, const float* gcat,
                               int64_t gcat_stride, const float* xcat, int64_t xcat_stride, float slope, float* gx, float* gs_src,
                               float* gs_dst, int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows == 0) return GGL_OK;
  GatStream a; GatDims d; int lpr;
  int st = gat_setup("gat_bwd", rowptr_t, col_t, plan_t, n_rows, E, H, C, H * C + H, workspace, workspace_bytes, &a, &d, &lpr);
  if (st) return st;
  GGL_REQUIRE(gcat && xcat && gx && gs_src && gs_dst, GGL_EINVAL, "gat_bwd: NULL pointer");
  GGL_REQUIRE(gcat_stride == d.K + 4 * H + 4 && xcat_stride % 4 == 0 && ((uintptr_t)gcat | (uintptr_t)xcat | (uintptr_t)gx) % 16 == 0,
              GGL_EINVAL, "gat_bwd: gcat rows must be [gout | {s_dst,m,1/s,D} x H | row id, 0, 0, 0], 16-byte aligned");
  GGL_CUDA(cudaMemsetAsync(gs_dst, 0, (size_t)n_rows * H * sizeof(float), stream));
  d.slope = slope;
  a.src = gcat; a.src_stride = gcat_stride;
  const size_t rowbytes = (size_t)gcat_stride * 4;
  a.rowb = (uint32_t)rowbytes;
  float* part = (float*)workspace;
  GGL_GAT_DISPATCH(gat_bwd_src_kernel, d, xcat, xcat_stride, gx, gs_src, part, gs_dst, (int32_t)(d.K + 4 * H));
  GGL_LAUNCHED("gat_bwd_merged");
  if (a.has_long) {
    gat_sum_finalize_kernel<<<sm_count() * 2, 256, 0, stream>>>(plan_t, part, d.K + H, d.W, gx, d.K, d.K, gs_src, H, a.ws_chunks);
    GGL_LAUNCHED("gat_sum_finalize");
  }
  return GGL_OK;
}

extern "C" int gglb200_gat_bwd_dst(const int32_t* rowptr, const int32_t* col, const int32_t* plan, const float* xcat,
                                   int64_t xcat_stride, const float* gcat, int64_t gcat_stride, float slope, float* gs_dst,
                                   int64_t n_rows, int64_t E, int32_t H, int32_t C, void* workspace, size_t workspace_bytes,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows == 0) return GGL_OK;
  GatStream a; GatDims d; int lpr;
  int st = gat_setup("gat_bwd_dst", rowptr, col, plan, n_rows, E, H, C, H, workspace, workspace_bytes, &a, &d, &lpr);
  if (st) return st;
  GGL_REQUIRE(gcat && xcat && gs_dst, GGL_EINVAL, "gat_bwd_dst: NULL pointer");
  GGL_REQUIRE(gcat_stride == d.K + 4 * H && xcat_stride >= d.K + H && xcat_stride % 4 == 0 &&
                  ((uintptr_t)gcat | (uintptr_t)xcat) % 16 == 0, GGL_EINVAL, "gat_bwd_dst: bad row layouts");
  d.slope = slope;
  a.src = xcat; a.src_stride = xcat_stride;
  const size_t rowbytes = (size_t)xcat_stride * 4;
  a.rowb = (uint32_t)rowbytes;
  float* part = (float*)workspace;
  GGL_GAT_DISPATCH(gat_bwd_dst_kernel, d, gcat, gcat_stride, gs_dst, part);
  GGL_LAUNCHED("gat_bwd_dst");
  if (a.has_long) {
    gat_sum_finalize_kernel<<<sm_count() * 2, 256, 0, stream>>>(plan, part, H, d.W, gs_dst, H, H, nullptr, 0, a.ws_chunks);
    GGL_LAUNCHED("gat_sum_finalize");
  }
  return GGL_OK;
}
