// sampler.cu -- GPU neighbour sampling for the mini-batch path.
//
// Replaces, for CUDA tensors, the per-seed loops of
//   /root/reference/gammagl/ops/sparse/cpu/sample.cpp:10-135            (sample_adj: one hop, what NeighborSampler.sample calls,
//                                                                         loader/neighbor_sampler.py:76-100)
//   /root/reference/gammagl/ops/sparse/cpu/neighbor_sample.cpp:22-100   (neighbor_sample: the same selection rule per hop)
// and the hash-table kernels of ops/sparse/cuda/neighbor_sample.cu.  Selection rule per seed w with deg in-neighbours
// (identical to the reference):
//   fanout < 0, or without replacement and fanout >= deg      every neighbour
//   with replacement                                           fanout draws, uniform over the deg neighbours
//   without replacement, fanout < deg                          Robert Floyd's algorithm -- for j = deg - fanout .. deg - 1: t = uniform
//                                                              [0, j]; take t unless already taken, else take j
//                                                              (sample.cpp:78-96, neighbor_sample.cpp:73-90)
// Two kernels per hop: COUNT (edges per seed, then an exclusive scan gives the output rowptr) and PICK (thread per seed for
// small fanouts, the picked CSR positions are written in ascending pick order).  Random numbers come from a counter-based
// generator keyed by (seed, hop salt, seed node, draw): the result does not depend on the launch geometry.
// Relabelling global -> local node ids is integer sort / unique work done by the caller (gammagl_b200/sampler.py).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ggl {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
// uniform integer in [0, n) for n >= 1 (multiply-shift on 32 random bits; n < 2^31)
__device__ __forceinline__ int64_t rand_below(uint64_t key, uint64_t draw, int64_t n) {
  const uint64_t r = mix64(key ^ mix64(draw));
  return (int64_t)(((r >> 32) * (uint64_t)n) >> 32);
}

__global__ void __launch_bounds__(256) sample_count_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ idx, int64_t n_seeds,
                                                           int64_t n_nodes, int64_t fanout, int replace, int64_t* __restrict__ cnt,
                                                           unsigned long long* __restrict__ n_bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  const int64_t w = idx[i];
  if (w < 0 || w >= n_nodes) { cnt[i] = 0; atomicAdd(n_bad, 1ull); return; }
  const int64_t deg = rowptr[w + 1] - rowptr[w];
  cnt[i] = (fanout < 0 || (!replace && fanout >= deg)) ? deg : (deg == 0 ? 0 : fanout);
}

constexpr int64_t kCoopDegree = 64;   // "take every neighbour" rows above this degree are copied by a warp, not by one thread

__global__ void __launch_bounds__(128) sample_pick_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ col,
                                                          const int64_t* __restrict__ idx, int64_t n_seeds, int64_t n_nodes, int64_t fanout,
                                                          int replace, uint64_t seed, const int64_t* __restrict__ out_ptr,
                                                          int64_t* __restrict__ e_id, int64_t* __restrict__ nbr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  const int64_t w = idx[i];
  if (w < 0 || w >= n_nodes) return;
  const int64_t b = rowptr[w], deg = rowptr[w + 1] - b;
  const int64_t o = out_ptr[i], k = out_ptr[i + 1] - o;
  if (k == 0) return;
  if (k == deg && (fanout < 0 || !replace)) {            // every neighbour, in CSR order
    if (deg > kCoopDegree) return;                        // hub rows: copied by the cooperative kernel below
    for (int64_t j = 0; j < deg; ++j) { e_id[o + j] = b + j; nbr[o + j] = col[b + j]; }
    return;
  }
  const uint64_t key = mix64(seed ^ mix64((uint64_t)w * 0x9e3779b97f4a7c15ull + 1));
  if (replace) {
    for (int64_t j = 0; j < k; ++j) {
      const int64_t t = rand_below(key, (uint64_t)j, deg);
      e_id[o + j] = b + t; nbr[o + j] = col[b + t];
    }
    return;
  }
  // Floyd: the picks of this seed are re-read from e_id (already written, same thread) to test membership
  for (int64_t j = deg - k, n = 0; j < deg; ++j, ++n) {
    int64_t t = rand_below(key, (uint64_t)n, j + 1);
    for (int64_t q = 0; q < n; ++q)
      if (e_id[o + q] == b + t) { t = j; break; }
    e_id[o + n] = b + t; nbr[o + n] = col[b + t];
  }
}

// full-neighbourhood rows of hub seeds (fanout < 0 or fanout >= degree > kCoopDegree): one warp per seed, coalesced copy
__global__ void __launch_bounds__(256) sample_copy_hubs_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ col,
                                                               const int64_t* __restrict__ idx, int64_t n_seeds, int64_t n_nodes,
                                                               int64_t fanout, int replace, const int64_t* __restrict__ out_ptr,
                                                               int64_t* __restrict__ e_id, int64_t* __restrict__ nbr) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_seeds; i += warps) {
    const int64_t w = idx[i];
    if (w < 0 || w >= n_nodes) continue;
    const int64_t b = rowptr[w], deg = rowptr[w + 1] - b;
    const int64_t o = out_ptr[i], k = out_ptr[i + 1] - o;
    if (!(k == deg && (fanout < 0 || !replace)) || deg <= kCoopDegree) continue;
    for (int64_t j = lane; j < deg; j += 32) { e_id[o + j] = b + j; nbr[o + j] = col[b + j]; }
  }
}

}  // namespace ggl

using namespace ggl;

extern "C" size_t gglb200_sample_workspace_bytes(int64_t n_seeds) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (int64_t)(n_seeds > 0 ? n_seeds + 1 : 1), (cudaStream_t)0);
  return bytes + 256 + (size_t)(n_seeds + 1) * 8;
}

// Phase 1: out_ptr[n_seeds + 1] = exclusive scan of the per-seed edge counts (out_ptr[n_seeds] = total, read by the host to size
// the outputs); *n_bad counts seeds outside [0, n_nodes).
extern "C" int gglb200_sample_count(const int64_t* rowptr, const int64_t* idx, int64_t n_seeds, int64_t n_nodes, int64_t fanout,
                                    int replace, int64_t* out_ptr, int64_t* n_bad, void* workspace, size_t workspace_bytes,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_seeds >= 0 && n_nodes >= 0, GGL_EINVAL, "sample_count: negative size");
  GGL_REQUIRE(out_ptr != nullptr && n_bad != nullptr, GGL_EINVAL, "sample_count: NULL output");
  GGL_CUDA(cudaMemsetAsync(n_bad, 0, sizeof(int64_t), stream));
  if (n_seeds == 0) { GGL_CUDA(cudaMemsetAsync(out_ptr, 0, sizeof(int64_t), stream)); return GGL_OK; }
  GGL_REQUIRE(rowptr && idx, GGL_EINVAL, "sample_count: NULL pointer");
  GGL_REQUIRE(workspace && workspace_bytes >= gglb200_sample_workspace_bytes(n_seeds), GGL_EWORKSPACE, "sample_count: workspace too small");
  int64_t* cnt = (int64_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  void* cub_ws = cnt + (n_seeds + 1);
  size_t cub_bytes = workspace_bytes - ((char*)cub_ws - (char*)workspace);
  GGL_CUDA(cudaMemsetAsync(cnt + n_seeds, 0, sizeof(int64_t), stream));
  sample_count_kernel<<<(unsigned)((n_seeds + 255) / 256), 256, 0, stream>>>(rowptr, idx, n_seeds, n_nodes, fanout, replace, cnt,
                                                                              (unsigned long long*)n_bad);
  GGL_LAUNCHED("sample_count");
  GGL_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, (const int64_t*)cnt, out_ptr, (int64_t)(n_seeds + 1), stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return GGL_OK;
}

// Phase 2: e_id[total] = picked CSR positions (grouped by seed, out_ptr from phase 1), nbr[total] = col[e_id].
extern "C" int gglb200_sample_pick(const int64_t* rowptr, const int64_t* col, const int64_t* idx, int64_t n_seeds, int64_t n_nodes,
                                   int64_t fanout, int replace, uint64_t seed, const int64_t* out_ptr, int64_t* e_id, int64_t* nbr,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n_seeds >= 0, GGL_EINVAL, "sample_pick: negative size");
  if (n_seeds == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && col && idx && out_ptr && e_id && nbr, GGL_EINVAL, "sample_pick: NULL pointer");
  sample_pick_kernel<<<(unsigned)((n_seeds + 127) / 128), 128, 0, stream>>>(rowptr, col, idx, n_seeds, n_nodes, fanout, replace, seed,
                                                                            out_ptr, e_id, nbr);
  GGL_LAUNCHED("sample_pick");
  if (fanout < 0 || !replace) {   // rows taken whole: hubs are copied warp-wide
    int64_t grid = (n_seeds + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    sample_copy_hubs_kernel<<<(unsigned)grid, 256, 0, stream>>>(rowptr, col, idx, n_seeds, n_nodes, fanout, replace, out_ptr, e_id, nbr);
    GGL_LAUNCHED("sample_copy_hubs");
  }
  return GGL_OK;
}
