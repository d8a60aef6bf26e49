// probe_umma.cu -- how long does ONE tcgen05.mma take?  One CTA per SM, one thread issues `reps` back-to-back MMAs (same operand tiles
// in shared memory, canonical no-swizzle K-major layout, accumulator in TMEM), commits, waits; clock64 around it.
// Variants: kind::tf32 (K = 8) / kind::f16 with bf16 operands (K = 16), N = 64 / 128 / 256, M = 128, same or rotating accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma tools/probe_umma.cu ; run: ./probe_umma  (JSON lines)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

template <int KIND /* 0 tf32, 1 bf16 */>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if constexpr (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128) probe(int N, int reps, int rotate, int elect, long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < (128 * 32 + 256 * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(smem), b_s = a_s + 128 * 32;
  // A: [128 rows, 32 bytes of K] = 2 chunks, LBO = 128 * 16; B: [N rows, 32 bytes of K], LBO = N * 16
  const uint64_t ad = umma_desc(a_s, 128 * 16, 128), bd = umma_desc(b_s, (uint32_t)N * 16, 128);
  const uint32_t fmt = KIND == 0 ? 2u : 1u;   // tf32 = 2 (kind::tf32), bf16 = 1 (kind::f16)
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t acc_mask = rotate ? (uint32_t)(256 / N - 1) : 0u;
  if (!elect) {
    if (tid == 0) {   // "thread 0 issues": the compiler cannot prove uniformity and wraps every MMA in an ELECT loop
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) umma<KIND>(tmem + ((uint32_t)r & acc_mask) * (uint32_t)N, ad, bd, idesc);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tLW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra LD;\n\tbra LW;\n\tLD:\n\t}" ::"r"(bar_s) : "memory");
      const long long t1 = clock64();
      if (blockIdx.x == 0) *cycles = t1 - t0;
    }
  } else if (warp_u == 0) {   // whole warp in the role, one elected lane issues (warp-uniform operands)
    const long long t0 = clock64();
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
    if (is_leader) {
      for (int r = 0; r < reps; ++r) umma<KIND>(tmem + ((uint32_t)r & acc_mask) * (uint32_t)N, ad, bd, idesc);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s) : "memory");
    }
    __syncwarp();
    asm volatile("{\n\t.reg .pred p;\n\tLW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra LD;\n\tbra LW;\n\tLD:\n\t}" ::"r"(bar_s) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 128 * 32 + 256 * 32;
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {64, 128, 256})
      for (int rotate = 0; rotate < 2; ++rotate)
       for (int elect = 0; elect < 2; ++elect)
        for (int ctas : {148, 296}) {
          const int reps = 2000;
          for (int it = 0; it < 2; ++it) {
            if (kind == 0) probe<0><<<ctas, 128, smem>>>(N, reps, rotate, elect, d);
            else probe<1><<<ctas, 128, smem>>>(N, reps, rotate, elect, d);
          }
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("{\"kind\": \"%s\", \"M\": 128, \"N\": %d, \"K_bytes\": 32, \"rotate_accumulators\": %d, \"elected_lane_of_uniform_warp\": %d, \"ctas\": %d, \"cycles_per_mma\": %.1f, \"floor_guide\": %d, \"err\": \"%s\"}\n",
                 kind == 0 ? "tf32" : "bf16", N, rotate, elect, ctas, (double)c / reps, N / 2, cudaGetErrorString(e));
        }
  return 0;
}
