// linear_tc.cu -- the dense feature x weight GEMM next to the aggregation on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   out[M, N] = act( A[M, K] @ W[K, N] + bias[N] )            fp32 in, fp32 out, act = identity | ReLU
//
// This is the Linear of GCNConv / SAGEConv / GINConv (/root/reference/gammagl/layers/conv/gcn_conv.py:79, sage_conv.py:80,
// gin_conv.py:57; tlx.layers.Linear = torch matmul on fp32 SIMT cores in the reference) applied to the aggregated rows when a
// layer aggregates first (F_in < F_out: `GCNConv(aggregate_first)`), with the bias + ReLU epilogue of gcn_conv.py:104-106 fused.
// A is the [N_nodes, F_in] output of the aggregation kernel, so F_in is small (100 / 128, up to 256 here): a 128-row tile is covered
// in one to four K passes of at most 64 columns whose operands live in shared memory at once; the kernel is bound by reading A and
// writing out, not by math.
//
// fp32 accuracy on TF32 tensor cores (3xTF32): a = a_hi + a_lo with a_hi = a truncated to TF32 (top 19 bits) and a_lo = a - a_hi
// (exact), likewise w; D = a_hi.w_hi + a_hi.w_lo + a_lo.w_hi in the fp32 TMEM accumulator.  The dropped a_lo.w_lo term and the
// truncation of the lo parts are O(2^-21) relative: parity with the fp32 reference at 1e-5 is tested, plain TF32 (2^-11) would
// not pass it.
//
// Two kernels share the operand layout and the W preparation: `linear_ws_kernel` (default) -- persistent, warp-specialised, one CTA per
// SM, two TMEM accumulators so the epilogue of a tile overlaps the MMAs of the next -- and `linear_tf32x3_kernel` (GGLB200_LINEAR=tile)
// -- one CTA per 128-row tile, two CTAs per SM, all warps run the phases in sequence.  Measured at C3 (100 -> 256, bias + ReLU):
// 0.93 ms and 1.56 ms (cuBLAS GEMM alone 2.95 ms); DESIGN.md 4.2 has the history and the probes.
//
// Layout / instructions (cta_group::1, 128 output rows per tile):
//   * operands in the canonical NO-SWIZZLE K-major layout of tcgen05 shared-memory descriptors: 8-row x 16-byte core matrices;
//     element (r, k) of a [R, K] operand sits at  (k / 4) * R * 16  +  r * 16  +  (k % 4) * 4   (K-chunk stride LBO = 16 R,
//     8-row-group stride SBO = 128).  Threads write A_hi/A_lo (lane <-> row: conflict-free 16-byte stores) and, per 64-column
//     chunk of the output, W_hi/W_lo transposed to [n, k];
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (UTCMMA; M = 128, N = 64, K = 8 per instruction), three per
//     K step, accumulating into TMEM columns [n0, n0 + 64); tcgen05.commit arrives on an mbarrier when the chunk's MMAs retire;
//   * epilogue: 8 warps read the 128 x N fp32 accumulator with tcgen05.ld.32x32b.x32 (warp w: TMEM lanes 32 (w % 4) .., column
//     half w / 4), add the bias, apply ReLU, and store 32-byte vectors (a lane owns 128 contiguous bytes of its row).
#include "common.cuh"
#include <cstdlib>

namespace ggl {

constexpr int kBM = 128;   // rows per CTA (UMMA M)
constexpr int kNC = 64;    // output columns per MMA chunk (UMMA N)

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (sm_100): start address >> 4 [0,14) | leading byte offset >> 4 [16,30) | stride byte offset >> 4 [32,46) |
  // version = 1 [46,48) | base offset 0 | lbo mode 0 | layout type SWIZZLE_NONE = 0 [61,64)
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);   // the 19 bits a TF32 operand keeps
  lo = v - hi;                                              // exact in fp32
}

// W prepared for the tensor core: for every (K pass, 64-column chunk) one contiguous block [W_hi tile | W_lo tile] in the canonical
// no-swizzle K-major layout (element (n, kc, j) of a tile at kc * 64 * 16 + n * 16 + j * 4), zero padded.  The GEMM kernel then fetches
// a block with ONE bulk async copy straight into shared memory -- no register staging, no split arithmetic per tile.
template <int KC>
__global__ void __launch_bounds__(256) linear_prep_w_kernel(const float* __restrict__ W, float* __restrict__ wc, int32_t K, int32_t N,
                                                            int32_t k_passes, int32_t n_chunks) {
  const int64_t total = (int64_t)k_passes * n_chunks * KC * kNC;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(t % kNC);
    const int kc = (int)((t / kNC) % KC);
    const int64_t blk = t / ((int64_t)kNC * KC);          // kp * n_chunks + c
    const int c = (int)(blk % n_chunks), kp = (int)(blk / n_chunks);
    float4 h, l;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = (kp * KC + kc) * 4 + j, col = c * kNC + n;
      v[j] = (k < K && col < N) ? __ldg(W + (int64_t)k * N + col) : 0.f;
    }
    split_tf32(v[0], h.x, l.x); split_tf32(v[1], h.y, l.y); split_tf32(v[2], h.z, l.z); split_tf32(v[3], h.w, l.w);
    float* base = wc + blk * (2 * KC * kNC * 4);
    *reinterpret_cast<float4*>(base + (kc * kNC + n) * 4) = h;
    *reinterpret_cast<float4*>(base + KC * kNC * 4 + (kc * kNC + n) * 4) = l;
  }
}

__device__ __forceinline__ bool elect_one() {   // one lane of a fully active warp
  uint32_t leader;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
  return leader != 0;
}

__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  // Bounded: a hand-off that never arrives (a protocol error) must end the launch with an error, not hang the GPU.  Every 4096
  // polls the global nanosecond timer is read; 10 s without progress -> trap.
  uint64_t t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if ((spins & 4095u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000ull) __trap();
    }
  }
}

// KC = 16-byte K chunks per K PASS (even, <= 14: one MMA K step = 2 chunks).  K is covered by k_passes passes, each re-staging its slice
// of A and accumulating into the same TMEM columns; with KC <= 14 a CTA needs 112 KB of dynamic + 128 B of static shared memory
// (A_hi | A_lo | two W blocks), so TWO CTAs are resident per SM -- one more KB of static shared memory costs the second one.
// Roles: all threads stage the A slice (fp32 -> hi / lo, one pass ahead in registers); warp 1 elects the W producer (one bulk async
// copy per (pass, chunk) into a 2-deep ring, mbarrier `full`); warp 0 waits for the block and ONE elected lane issues its MMAs and
// commits them to `done` (ring slot free), `pass` (A slice free) and `acc[c]` (accumulator chunk final); 8 warps run the epilogue,
// chunk by chunk as the acc barriers arrive.  The role warps are selected by a warp index that went through __shfl_sync: the compiler
// then knows the branch is warp-uniform, keeps the descriptors in uniform registers and issues the UTCHMMAs back to back; behind a plain
// `tid == 0` test every MMA is wrapped in an ELECT loop with five R2UR moves and costs ~129 cycles (tools/probe_linear.py).
// Every barrier is waited phase by phase, in order, by each thread that uses it: a parity wait means nothing to a thread that has not
// seen the previous phase.
template <int KC>
__global__ void __launch_bounds__(256, 2) linear_tf32x3_kernel(const float* __restrict__ A, const float* __restrict__ wc,
                                                               const float* __restrict__ bias, const int relu, float* __restrict__ out,
                                                               const int64_t M, const int32_t K, const int32_t N, const int32_t tmem_cols,
                                                               const int32_t k_passes, const int32_t dbg) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[9];   // full[0..1], done[0..1], pass (all threads wait EVERY phase of it in order:
                                                        // a parity wait is only meaningful to a thread that has seen the previous phase)
  __shared__ uint32_t tmem_base_slot;
  constexpr uint32_t A_BYTES = KC * kBM * 16;   // one of A_hi / A_lo
  constexpr uint32_t B_BYTES = KC * kNC * 16;   // one of W_hi / W_lo of a block
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + A_BYTES;
  unsigned char* b_ring = smem + 2 * A_BYTES;   // slot s at b_ring + s * 2 * B_BYTES: [W_hi | W_lo]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // the same value, provably warp-uniform: role branches stay convergent
  const int64_t row0 = (int64_t)blockIdx.x * kBM;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const int n_chunks = (N + kNC - 1) / kNC;
  const int total = k_passes * n_chunks;

  if (warp == 0) {   // TMEM allocation by one full warp; the address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_slot)),
                 "r"((uint32_t)tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }

  constexpr int AI = (4 * KC + 7) / 8;
  float4 va[AI];
  auto load_a = [&](int kp) {
#pragma unroll
    for (int i = 0; i < AI; ++i) {
      const int item = warp + 8 * i, band = item & 3, kc = item >> 2;
      const int r = band * 32 + lane;
      const int k = (kp * KC + kc) * 4;
      va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (item < 4 * KC && row0 + r < M && k < K && !(dbg & 4)) va[i] = __ldg(reinterpret_cast<const float4*>(A + (row0 + r) * K + k));
    }
  };
  auto store_a = [&]() {
#pragma unroll
    for (int i = 0; i < AI; ++i) {
      const int item = warp + 8 * i, band = item & 3, kc = item >> 2;
      const int r = band * 32 + lane;
      if (item < 4 * KC) {
        float4 h, l;
        split_tf32(va[i].x, h.x, l.x); split_tf32(va[i].y, h.y, l.y); split_tf32(va[i].z, h.z, l.z); split_tf32(va[i].w, h.w, l.w);
        const uint32_t off = (uint32_t)kc * (kBM * 16) + (uint32_t)r * 16;
        *reinterpret_cast<float4*>(a_hi + off) = h;
        *reinterpret_cast<float4*>(a_lo + off) = l;
      }
    }
  };
  // producer: block i -> ring slot i & 1 (after the MMAs of block i - 2 have released it)
  auto issue_copy = [&](int i) {
    const uint32_t slot = (uint32_t)i & 1u, full = bar0 + 8 * slot, done = bar0 + 16 + 8 * slot;
    if (i >= 2) mbar_wait_parity(done, (uint32_t)((i - 2) >> 1) & 1u);
    if (dbg & 1) {   // experiment: no W traffic
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
      return;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(2u * B_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(b_ring + slot * 2 * B_BYTES)), "l"(wc + (int64_t)i * (2 * B_BYTES / 4)),
                   "r"(2u * B_BYTES), "r"(full) : "memory");
  };

  load_a(0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();   // barriers initialised, TMEM address published
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  if (warp_u == 1) {
    if (elect_one()) {
      issue_copy(0);
      if (total > 1) issue_copy(1);
    }
    __syncwarp();
  }

  // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNC >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(a_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(a_lo);
  const uint32_t b_ring_s = (uint32_t)__cvta_generic_to_shared(b_ring);

  for (int kp = 0; kp < k_passes; ++kp) {
    // the MMAs of the previous pass have read the A slice: wait for that pass's commit before overwriting
    if (kp > 0) mbar_wait_parity(bar0 + 32, (uint32_t)(kp - 1) & 1u);
    store_a();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (kp + 1 < k_passes) load_a(kp + 1);   // next slice in flight while the tensor core works on this one
    if (warp_u == 0) {
      // MMA role: the whole warp waits for the W block, ONE elected lane issues (operands stay in uniform registers: a lone
      // `tid == 0` branch makes the compiler move five registers to the uniform file before every single MMA)
      for (int c = 0; c < n_chunks; ++c) {
        const int i = kp * n_chunks + c;
        const uint32_t slot = (uint32_t)i & 1u;
        mbar_wait_parity(bar0 + 8 * slot, (uint32_t)(i >> 1) & 1u);          // W block i has landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(c * kNC);
          const uint32_t bh_s = b_ring_s + slot * 2 * B_BYTES, bl_s = bh_s + B_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / 2; ++ks) {   // one MMA K step = 8 elements = 2 chunks
            const uint64_t ah = umma_desc(a_hi_s + (uint32_t)ks * 2 * (kBM * 16), kBM * 16, 128);
            const uint64_t al = umma_desc(a_lo_s + (uint32_t)ks * 2 * (kBM * 16), kBM * 16, 128);
            const uint64_t bh = umma_desc(bh_s + (uint32_t)ks * 2 * (kNC * 16), kNC * 16, 128);
            const uint64_t bl = umma_desc(bl_s + (uint32_t)ks * 2 * (kNC * 16), kNC * 16, 128);
            umma_tf32(d, al, bh, IDESC, (kp > 0 || ks > 0) ? 1u : 0u);   // small terms first
            umma_tf32(d, ah, bl, IDESC, 1u);
            umma_tf32(d, ah, bh, IDESC, 1u);
          }
          // arrives on `done[slot]` when every MMA issued so far has completed (implies tcgen05.fence::before_thread_sync)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 16 + 8 * slot) : "memory");
          if (kp + 1 == k_passes)   // chunk c of the accumulator is final: its epilogue may start while the next chunks multiply
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 40 + 8 * c) : "memory");
          if (c + 1 == n_chunks && kp + 1 < k_passes)   // pass done (the last pass is tracked per chunk: nothing may arrive after acc[])
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 32) : "memory");
        }
        __syncwarp();
      }
    } else if (warp_u == 1) {
      if (elect_one()) {   // W producer
        for (int c = 0; c < n_chunks; ++c) {
          const int i = kp * n_chunks + c + 2;
          if (i < total) issue_copy(i);
        }
      }
      __syncwarp();
    }
  }
  // ---- epilogue: TMEM -> registers -> bias / ReLU -> global.  Warp w reads TMEM lanes 32 (w % 4) .. (its rows) and the column half
  // w / 4, 32 columns per tcgen05.ld; a lane holds 128 contiguous bytes of its output row and stores them as four 32-byte vectors
  // (full sectors; 16-byte vectors when rows are not 32-byte aligned).  Chunk c is read as soon as ITS last MMAs have retired.
  {
    const int q = warp & 3, half = warp >> 2;
    const int64_t r = row0 + q * 32 + lane;
    const int n_cols = n_chunks * kNC;
    const int c_lo = half * (n_cols / 2), c_hi = c_lo + n_cols / 2;
    const bool wide = ((N & 7) == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0);
    for (int col = c_lo; col < c_hi; col += 32) {
      // the bias of these 32 columns, requested BEFORE the waits below so its latency hides behind them (static shared memory is
      // full: one more KB would cost the second resident CTA)
      float4 bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        bv[j] = (bias != nullptr && col + 4 * j < N) ? __ldg(reinterpret_cast<const float4*>(bias + col) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait_parity(bar0 + 40 + 8 * (uint32_t)(col / kNC), 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
          "%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr) : "memory");
      if (r < M && col < N && !(dbg & 2)) {
        float* orow = out + r * N + col;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (col + j >= N) break;
          const float4 b0 = bv[j / 4], b1 = bv[j / 4 + 1];
          float o[8] = {__uint_as_float(v[j]) + b0.x,     __uint_as_float(v[j + 1]) + b0.y, __uint_as_float(v[j + 2]) + b0.z,
                        __uint_as_float(v[j + 3]) + b0.w, __uint_as_float(v[j + 4]) + b1.x, __uint_as_float(v[j + 5]) + b1.y,
                        __uint_as_float(v[j + 6]) + b1.z, __uint_as_float(v[j + 7]) + b1.w};
          if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
          }
          if (wide) {
            asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + j), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]),
                         "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
          } else {
            __stcs(reinterpret_cast<float4*>(orow + j), make_float4(o[0], o[1], o[2], o[3]));
            if (col + j + 4 < N) __stcs(reinterpret_cast<float4*>(orow + j + 4), make_float4(o[4], o[5], o[6], o[7]));
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
  }
}

// ---- warp-specialised persistent form -------------------------------------------------------------------------------------------
// One CTA per SM (320 threads), tiles t = blockIdx.x, + gridDim.x, ...; everything that the one-tile form above does in sequence
// runs concurrently here, decoupled by mbarriers:
//   warps 0-3  A producers: thread p stages row p of the tile, slice by slice (fp32 -> hi / lo) into a 2-deep ring of A slices
//              (a_full: 128 arrivals after fence.proxy.async; a_empty: tcgen05.commit of the pass that read the slice);
//   warp 4     MMA role (one elected lane): waits a_full / w_full, issues 21 MMAs per (pass, chunk), commits w_done, a_empty and, after
//              the last pass, acc_full; the accumulators of tile i and i + 1 alternate between TMEM columns [0,256) and [256,512);
//   warp 5     W producer (one elected lane): bulk async copies of the pre-split blocks into a 3-deep ring (w_full / w_done);
//   warps 6-9  epilogue: wait acc_full, TMEM -> bias / ReLU -> 32-byte stores, arrive acc_empty (128 arrivals) -- while the tensor
//              core already multiplies the next tile.
// Barrier phases are derived from running counters (slice q, block i, tile it); every thread of a role waits every phase of its
// barriers in order; the FIRST wait on an `empty` barrier uses the inverted parity and passes at once.
constexpr int kWsRingMax = 6;   // W ring depth: 2 .. 6 slots (GGLB200_LINEAR_RING; default 2 for 56-column passes, 5 for the 32-column plan)
template <int KC>
__global__ void __launch_bounds__(352, 1) linear_ws_kernel(const float* __restrict__ A, const float* __restrict__ wc,
                                                           const float* __restrict__ bias, const int relu, float* __restrict__ out,
                                                           const int64_t M, const int32_t K, const int32_t N, const int32_t k_passes,
                                                           const int32_t ring, const int32_t staged) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(16) float bias_s[256];
  __shared__ __align__(8) unsigned long long bars[4 + 2 * kWsRingMax + 4 + 2];   // a_full[2] a_empty[2] w_full[S] w_done[S] acc_full[2] acc_empty[2] raw_full raw_empty
  __shared__ uint32_t tmem_base_slot;
  constexpr uint32_t A_BYTES = KC * kBM * 16;   // one of A_hi / A_lo of a slice
  constexpr uint32_t B_BYTES = KC * kNC * 16;   // one of W_hi / W_lo of a block
  unsigned char* a_ring = smem;                               // slice buffer b at a_ring + b * 2 * A_BYTES: [A_hi | A_lo]
  unsigned char* b_ring = smem + 4 * A_BYTES;                 // slot s at b_ring + s * 2 * B_BYTES: [W_hi | W_lo]
  unsigned char* raw = b_ring + (uint32_t)ring * 2 * B_BYTES; // staged form: the fp32 tile as it lies in global memory, [128, K]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t A_FULL = bar0, A_EMPTY = bar0 + 16, W_FULL = bar0 + 32, W_DONE = W_FULL + 8 * kWsRingMax, ACC_FULL = W_DONE + 8 * kWsRingMax,
                 ACC_EMPTY = ACC_FULL + 16, RAW_FULL = ACC_EMPTY + 16, RAW_EMPTY = RAW_FULL + 8;
  const int n_chunks = (N + kNC - 1) / kNC;
  const int total = k_passes * n_chunks;
  const int n_tiles = (int)((M + kBM - 1) / kBM);
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_all = my_tiles * total;

  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(A_FULL + 8 * i) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(A_EMPTY + 8 * i) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ACC_FULL + 8 * i) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(ACC_EMPTY + 8 * i) : "memory");
    }
    for (int i = 0; i < kWsRingMax; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W_FULL + 8 * i) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W_DONE + 8 * i) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(RAW_FULL) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(RAW_EMPTY) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < 256) bias_s[tid] = (bias != nullptr && tid < N) ? __ldg(bias + tid) : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp_u < 4) {
    // ---------------- A producers ----------------
    // Two slices in registers: the loads of slice q + 1 are in flight while slice q waits for its buffer and is split / stored, so
    // the DRAM latency never sits between two MMA passes.
    const int r = tid;   // row of the tile
    const int n_slices = my_tiles * k_passes;
    auto load_slice = [&](float4 (&v)[KC], int q) {
      const int it = q / k_passes, kp = q - it * k_passes;
      const int64_t row = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM + r;
      const float* arow = A + row * K;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const int k = (kp * KC + kc) * 4;
        v[kc] = (row < M && k < K) ? __ldg(reinterpret_cast<const float4*>(arow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_slice = [&](const float4 (&v)[KC], int q) {
      const uint32_t buf = (uint32_t)q & 1u;
      mbar_wait_parity(A_EMPTY + 8 * buf, (((uint32_t)q >> 1) & 1u) ^ 1u);   // the pass that read this buffer two slices ago has retired
      unsigned char* hi = a_ring + buf * 2 * A_BYTES + (uint32_t)r * 16;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        float4 h, l;
        split_tf32(v[kc].x, h.x, l.x); split_tf32(v[kc].y, h.y, l.y); split_tf32(v[kc].z, h.z, l.z); split_tf32(v[kc].w, h.w, l.w);
        *reinterpret_cast<float4*>(hi + (uint32_t)kc * (kBM * 16)) = h;
        *reinterpret_cast<float4*>(hi + A_BYTES + (uint32_t)kc * (kBM * 16)) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(A_FULL + 8 * buf) : "memory");
    };
    if (staged) {
      // Staged form (k_passes <= 2, or <= 4 with 32-column passes): the whole fp32 tile -- 128 consecutive rows = ONE contiguous piece of global memory -- arrives by a
      // single bulk async copy (warp 10); a thread takes its row into registers, hands the staging buffer back at once (the next
      // tile's copy overlaps this tile's split / store / MMA) and then feeds the slice ring from registers.  No L1, no per-thread
      // 16-byte requests: the loads were the limit of the form below (8 k cycles per 28 KB slice in flight).
      constexpr int PMAX = KC <= 8 ? 4 : 2;   // passes whose slices a thread keeps in registers (32-column plan: up to 4 x 8 chunks)
      float4 v[PMAX][KC];
      for (int it = 0; it < my_tiles; ++it) {
        const int64_t row = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM + r;
        mbar_wait_parity(RAW_FULL, (uint32_t)it & 1u);
        const unsigned char* rrow = raw + (size_t)r * (size_t)K * 4;
#pragma unroll
        for (int pp = 0; pp < PMAX; ++pp) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            const int k = (pp * KC + kc) * 4;
            v[pp][kc] = (row < M && k < K && pp < k_passes) ? *reinterpret_cast<const float4*>(rrow + k * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(RAW_EMPTY) : "memory");   // row is in registers
#pragma unroll
        for (int pp = 0; pp < PMAX; ++pp)
          if (pp < k_passes) store_slice(v[pp], it * k_passes + pp);
      }
    } else {
    float4 va[KC], vb[KC];
    load_slice(va, 0);
    for (int q = 0; q < n_slices; q += 2) {
      if (q + 1 < n_slices) load_slice(vb, q + 1);
      store_slice(va, q);
      if (q + 1 < n_slices) {
        if (q + 2 < n_slices) load_slice(va, q + 2);
        store_slice(vb, q + 1);
      }
    }
    }
  } else if (warp_u == 4) {
    // ---------------- MMA role ----------------
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNC >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    const uint32_t a_ring_s = (uint32_t)__cvta_generic_to_shared(a_ring), b_ring_s = (uint32_t)__cvta_generic_to_shared(b_ring);
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t ab = (uint32_t)it & 1u;
      mbar_wait_parity(ACC_EMPTY + 8 * ab, (((uint32_t)it >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kp = 0; kp < k_passes; ++kp) {
        const int q = it * k_passes + kp;
        const uint32_t buf = (uint32_t)q & 1u;
        mbar_wait_parity(A_FULL + 8 * buf, ((uint32_t)q >> 1) & 1u);
        for (int c = 0; c < n_chunks; ++c) {
          const int i = q * n_chunks + c;
          const uint32_t slot = (uint32_t)(i % ring);
          mbar_wait_parity(W_FULL + 8 * slot, (uint32_t)(i / ring) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t d = tmem_base + ab * 256u + (uint32_t)(c * kNC);
            const uint32_t a_hi_s = a_ring_s + buf * 2 * A_BYTES, a_lo_s = a_hi_s + A_BYTES;
            const uint32_t bh_s = b_ring_s + slot * 2 * B_BYTES, bl_s = bh_s + B_BYTES;
#pragma unroll
            for (int ks = 0; ks < KC / 2; ++ks) {
              const uint64_t ah = umma_desc(a_hi_s + (uint32_t)ks * 2 * (kBM * 16), kBM * 16, 128);
              const uint64_t al = umma_desc(a_lo_s + (uint32_t)ks * 2 * (kBM * 16), kBM * 16, 128);
              const uint64_t bh = umma_desc(bh_s + (uint32_t)ks * 2 * (kNC * 16), kNC * 16, 128);
              const uint64_t bl = umma_desc(bl_s + (uint32_t)ks * 2 * (kNC * 16), kNC * 16, 128);
              umma_tf32(d, al, bh, IDESC, (kp > 0 || ks > 0) ? 1u : 0u);
              umma_tf32(d, ah, bl, IDESC, 1u);
              umma_tf32(d, ah, bh, IDESC, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(W_DONE + 8 * slot) : "memory");
            if (c + 1 == n_chunks) {
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(A_EMPTY + 8 * buf) : "memory");
              if (kp + 1 == k_passes)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ACC_FULL + 8 * ab) : "memory");
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp_u == 5) {
    // ---------------- W producer ----------------
    if (elect_one()) {
      for (int i = 0; i < total_all; ++i) {
        const uint32_t slot = (uint32_t)(i % ring);
        mbar_wait_parity(W_DONE + 8 * slot, ((uint32_t)(i / ring) & 1u) ^ 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(W_FULL + 8 * slot), "r"(2u * B_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(b_ring + slot * 2 * B_BYTES)), "l"(wc + (int64_t)(i % total) * (2 * B_BYTES / 4)),
                       "r"(2u * B_BYTES), "r"(W_FULL + 8 * slot) : "memory");
      }
    }
    __syncwarp();
  } else if (warp_u == 10) {
    // ---------------- A tile copies (staged form) ----------------
    if (staged && elect_one()) {
      for (int it = 0; it < my_tiles; ++it) {
        const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM;
        const int64_t rows = (M - row0 < kBM) ? (M - row0) : kBM;
        const uint32_t bytes = (uint32_t)(rows * K * 4);
        mbar_wait_parity(RAW_EMPTY, ((uint32_t)it & 1u) ^ 1u);   // every producer thread holds its row of the previous tile
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(RAW_FULL), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(raw)), "l"(A + row0 * K), "r"(bytes), "r"(RAW_FULL) : "memory");
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (warps 6-9: TMEM lane quarter = warp % 4) ----------------
    const int q4 = warp & 3;
    const int n_cols = n_chunks * kNC;
    const bool wide = ((N & 7) == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0);
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t ab = (uint32_t)it & 1u;
      const int64_t r = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM + q4 * 32 + lane;
      mbar_wait_parity(ACC_FULL + 8 * ab, ((uint32_t)it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_lane = tmem_base + ab * 256u + ((uint32_t)(q4 * 32) << 16);
      for (int col = 0; col < n_cols; col += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
            "%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(t_lane + (uint32_t)col) : "memory");
        if (r < M && col < N) {
          float* orow = out + r * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col + j >= N) break;
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + col + j), b1 = *reinterpret_cast<const float4*>(bias_s + col + j + 4);
            float o[8] = {__uint_as_float(v[j]) + b0.x,     __uint_as_float(v[j + 1]) + b0.y, __uint_as_float(v[j + 2]) + b0.z,
                          __uint_as_float(v[j + 3]) + b0.w, __uint_as_float(v[j + 4]) + b1.x, __uint_as_float(v[j + 5]) + b1.y,
                          __uint_as_float(v[j + 6]) + b1.z, __uint_as_float(v[j + 7]) + b1.w};
            if (relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
            }
            if (wide) {
              asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + j), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]),
                           "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
            } else if ((N & 3) == 0) {
              __stcs(reinterpret_cast<float4*>(orow + j), make_float4(o[0], o[1], o[2], o[3]));
              if (col + j + 4 < N) __stcs(reinterpret_cast<float4*>(orow + j + 4), make_float4(o[4], o[5], o[6], o[7]));
            } else {   // N not a multiple of 4 (47 classes ...): rows are only 4-byte aligned
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (col + j + i < N) __stcs(orow + j + i, o[i]);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads done before the accumulator is handed back
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ACC_EMPTY + 8 * ab) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace ggl

using namespace ggl;

// 32-column plan of the warp-specialised kernel (K <= 128): passes of 8 chunks keep the A ring at 64 KB, which buys a 5-slot W ring
// (16 KB blocks) next to the staged tile -- the copy latency of a W block (~2.3 k cycles) is then covered by four blocks in flight.
static bool linear_plan_deep(int64_t K, int* k_passes, int* kc) {
  const int chunks = (int)((K + 3) / 4);
  if (chunks > 32) return false;
  *k_passes = (chunks + 7) / 8;
  int c = (chunks + *k_passes - 1) / *k_passes;
  *kc = (c + 1) & ~1;
  return true;
}
static void linear_plan(int64_t K, int* k_passes, int* kc) {
  // K passes: slices of at most 14 sixteen-byte chunks (56 columns), even, as equal as possible -- 112 KB of shared memory per CTA
  const int chunks = (int)((K + 3) / 4);
  *k_passes = (chunks + 13) / 14;
  int c = (chunks + *k_passes - 1) / *k_passes;
  *kc = (c + 1) & ~1;
}
static int linear_kc_class(int kc) { return kc <= 4 ? 4 : (kc <= 8 ? 8 : (kc <= 12 ? 12 : 14)); }

extern "C" size_t gglb200_linear_workspace_bytes(int64_t K, int64_t N) {
  if (K <= 0 || N <= 0) return 0;
  int k_passes, kc;
  linear_plan(K, &k_passes, &kc);
  const int KC = linear_kc_class(kc);
  const int64_t n_chunks = (N + kNC - 1) / kNC;
  size_t bytes = (size_t)k_passes * n_chunks * 2 * KC * kNC * 16;
  if (linear_plan_deep(K, &k_passes, &kc)) {   // whichever plan the launch picks must fit
    const size_t deep = (size_t)k_passes * n_chunks * 2 * linear_kc_class(kc) * kNC * 16;
    if (deep > bytes) bytes = deep;
  }
  return bytes + 256;
}

extern "C" int gglb200_linear_tf32x3(const float* A, const float* W, const float* bias, int relu, float* out, int64_t M, int64_t K,
                                     int64_t N, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(M >= 0 && K > 0 && N > 0, GGL_EINVAL, "linear: bad sizes M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
  if (M == 0) return GGL_OK;
  GGL_REQUIRE(A && W && out, GGL_EINVAL, "linear: NULL pointer");
  GGL_REQUIRE(K % 4 == 0 && K <= 256, GGL_EUNSUPPORTED, "linear: K = %lld (supported: multiples of 4 up to 256)", (long long)K);
  GGL_REQUIRE(N <= 256, GGL_EUNSUPPORTED, "linear: N = %lld (supported: up to 256)", (long long)N);
  GGL_REQUIRE((((uintptr_t)A | (uintptr_t)out | (uintptr_t)bias) & 15) == 0, GGL_EUNSUPPORTED, "linear: 16-byte aligned A / out / bias");
  GGL_REQUIRE((M + kBM - 1) / kBM < INT32_MAX, GGL_ERANGE, "linear: M too large");
  GGL_REQUIRE(workspace != nullptr && workspace_bytes >= gglb200_linear_workspace_bytes(K, N), GGL_EWORKSPACE,
              "linear: workspace %zu < %zu", workspace_bytes, gglb200_linear_workspace_bytes(K, N));
  const int n_chunks = (int)((N + kNC - 1) / kNC);
  int tmem_cols = 32;
  while (tmem_cols < n_chunks * kNC) tmem_cols <<= 1;
  int k_passes, kc;
  linear_plan(K, &k_passes, &kc);
  float* wc = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const char* dbg_env = getenv("GGLB200_LINEAR_DBG");   // timing experiments only (1: no W copies, 2: no stores, 4: no A loads)
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  const unsigned grid = (unsigned)((M + kBM - 1) / kBM);
  const char* form = getenv("GGLB200_LINEAR");   // "ws" (default): warp-specialised persistent kernel; "tile": one CTA per tile
  const bool ws = !(form && form[0] == 't') || (N % 4 != 0);   // the one-tile form stores 16-byte vectors only
  const unsigned grid_ws = (unsigned)((int64_t)grid < (int64_t)sm_count() ? grid : sm_count());
  const char* stage_env = getenv("GGLB200_LINEAR_STAGE");   // "0": keep the per-thread loads even where the staged form applies
  const bool no_stage = stage_env && stage_env[0] == '0';
  const char* ring_env = getenv("GGLB200_LINEAR_RING");
  int ring = (ring_env && atoi(ring_env) >= 2 && atoi(ring_env) <= kWsRingMax) ? atoi(ring_env) : 2;
  // GGLB200_LINEAR_PLAN=deep: 32-column passes + 5-slot W ring for the staged warp-specialised form (K <= 128)
  const char* plan_env = getenv("GGLB200_LINEAR_PLAN");
  bool deep = false;
  if (ws && !no_stage && plan_env && plan_env[0] == 'd') {
    int kp2, kc2;
    if (linear_plan_deep(K, &kp2, &kc2) && kc2 <= 8) {
      k_passes = kp2;
      kc = kc2;
      deep = true;
      if (!ring_env) ring = 5;
    }
  }
#define GGL_LIN(KC_)                                                                                                        \
  do {                                                                                                                      \
    const int64_t items = (int64_t)k_passes * n_chunks * KC_ * kNC;                                                         \
    linear_prep_w_kernel<KC_><<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(W, wc, (int32_t)K, (int32_t)N, k_passes,  \
                                                                                   n_chunks);                               \
    GGL_LAUNCHED("linear_prep_w");                                                                                          \
    if (ws) {                                                                                                               \
      const size_t smem_ring = (size_t)4 * KC_ * kBM * 16 + (size_t)ring * 2 * KC_ * kNC * 16;                             \
      const int staged = (!no_stage && k_passes <= ((deep && KC_ <= 8) ? 4 : 2) &&                                          \
                          smem_ring + (size_t)kBM * K * 4 + 2048 <= 232448) ? 1 : 0;                                       \
      const size_t smem_ws = smem_ring + (staged ? (size_t)kBM * K * 4 : 0);                                                 \
      GGL_CUDA(cudaFuncSetAttribute(linear_ws_kernel<KC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));     \
      linear_ws_kernel<KC_><<<grid_ws, 352, smem_ws, stream>>>(A, wc, bias, relu, out, M, (int32_t)K, (int32_t)N, k_passes,  \
                                                               ring, staged);                                                \
      break;                                                                                                                \
    }                                                                                                                       \
    const size_t smem = (size_t)2 * KC_ * kBM * 16 + (size_t)4 * KC_ * kNC * 16;                                            \
    GGL_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<KC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    linear_tf32x3_kernel<KC_><<<grid, 256, smem, stream>>>(A, wc, bias, relu, out, M, (int32_t)K, (int32_t)N, tmem_cols,    \
                                                           k_passes, dbg);                                                       \
  } while (0)
  switch (linear_kc_class(kc)) {
    case 4: GGL_LIN(4); break;
    case 8: GGL_LIN(8); break;
    case 12: GGL_LIN(12); break;
    default: GGL_LIN(14); break;
  }
#undef GGL_LIN
  GGL_LAUNCHED("linear_tf32x3");
  return GGL_OK;
}
