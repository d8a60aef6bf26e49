#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k(const __grid_constant__ CUtensorMap tm, const int* rows, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int r0 = rows[0], r1 = rows[1], r2 = rows[2], r3 = rows[3];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4 * 400) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(d), "l"(&tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(b) : "memory");
  }
  __syncthreads();
  out[threadIdx.x] = ((float*)smem)[threadIdx.x];
}
