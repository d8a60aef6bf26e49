// peer.cu -- halo-cache fill of the 1-D node partition (no counterpart in the reference, which has no
// multi-GPU code).
//
// gglb200_spmm_csr_peer pulls every remote source row over NVLink once PER EDGE.  Rows referenced by several local edges
// (hub rows on power-law graphs, ~half of the remote rows on a uniform graph at papers100M density) are instead pulled ONCE
// per pass into a cache region behind the rank's own shard -- a deduplicated halo filled by remote READS (no pack kernel on
// the owner, no all-to-all, no unpack) -- and the aggregation kernel then finds them in local HBM: `col` of such an edge
// carries owner = self, row = B + cache slot.  Rows referenced once keep their direct peer address.
//
//   dst[i, :] = table[ids[i] >> shift][(ids[i] & mask) * K : +K]           i < n
//
// One warp moves 4 rows per step as 16-byte words (ld.global.nc from the peer mapping, streaming store to local HBM);
// ids are sorted by (owner, row) at plan time so the remote reads walk each owner's shard in ascending order.
#include "common.cuh"

namespace ggl {

template <int U>
__global__ void __launch_bounds__(256) gather_rows_peer_kernel(const unsigned long long* __restrict__ table, int32_t shift,
                                                               const int32_t* __restrict__ ids, uint4* __restrict__ dst,
                                                               int64_t n, int32_t words /* 16-byte words per row */) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int32_t mask = (1 << shift) - 1;
  // Every warp walks its OWN contiguous segment of the id list (sorted by owner, row).  With a grid-stride walk all warps of all
  // ranks would read from owner 0 first, then owner 1 ...: one GPU's NVLink egress serves everybody while the others idle (measured:
  // 300 GB/s into each GPU at 8 GPUs).  Contiguous segments spread the reads of every rank over all owners at all times and keep
  // each segment sequential in the owner's HBM.
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t per = ((n + warps - 1) / warps + U - 1) / U * U;
  const int64_t seg_end = min(n, (w + 1) * per);
  for (int64_t i0 = w * per; i0 < seg_end; i0 += U) {
    const uint4* src[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      src[u] = nullptr;
      if (i0 + u < seg_end) {
        const int32_t c = __ldg(ids + i0 + u);
        src[u] = reinterpret_cast<const uint4*>(__ldg(table + (c >> shift))) + (int64_t)(c & mask) * words;
      }
    }
    for (int32_t k = lane; k < words; k += 32) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (src[u]) {
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src[u] + k));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (src[u]) __stcs(dst + (i0 + u) * words + k, v[u]);
    }
  }
}

}  // namespace ggl

using namespace ggl;

extern "C" int gglb200_gather_rows_peer(const void* const* x_table, int32_t owner_shift, const int32_t* ids, void* dst,
                                        int64_t n, int64_t row_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n >= 0 && row_bytes > 0, GGL_EINVAL, "gather_rows_peer: bad sizes");
  if (n == 0) return GGL_OK;
  GGL_REQUIRE(x_table && ids && dst, GGL_EINVAL, "gather_rows_peer: NULL pointer");
  GGL_REQUIRE(owner_shift >= 1 && owner_shift <= 30, GGL_EINVAL, "gather_rows_peer: bad owner_shift");
  GGL_REQUIRE(row_bytes % 16 == 0 && ((uintptr_t)dst & 15) == 0, GGL_EUNSUPPORTED,
              "gather_rows_peer: rows must be 16-byte multiples and 16-byte aligned (row_bytes=%lld)", (long long)row_bytes);
  constexpr int U = 4;
  int64_t grid = (n + 8 * U - 1) / (8 * U);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  gather_rows_peer_kernel<U><<<(int)grid, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(x_table), owner_shift, ids,
                                                            reinterpret_cast<uint4*>(dst), n, (int32_t)(row_bytes / 16));
  GGL_LAUNCHED("gather_rows_peer");
  return GGL_OK;
}
