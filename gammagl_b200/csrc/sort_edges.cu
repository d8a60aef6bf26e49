// sort_edges.cu -- two-key edge sort and duplicate detection on the GPU (integer work, bit-exact).
//
// Replaces the host-side numpy/torch argsort of `major * n + minor` in
//   /root/reference/gammagl/utils/sort_edge_index.py:31-45   (sort_edge_index)
//   /root/reference/gammagl/utils/coalesce.py:41-78          (coalesce: sort, mask of first occurrences, merge ids)
//   /root/reference/gammagl/sparse/storage.py:124-140,200-217 (SparseStorage constructor sort, csr2csc)
//
// Pipeline (all on `stream`, no host sync, caller-owned workspace):
//   1. pack:    key[i] = major[i] << bits(n_minor) | minor[i]   (same order as major * n_minor + minor), id[i] = i
//   2. sort:    cub::DeviceRadixSort::SortPairs over bits(n_major) + bits(n_minor) bits; LSD radix = STABLE, so equal
//               pairs keep their input order (the reference's argsort leaves that order unspecified)
//   3. unpack:  sorted_major / sorted_minor decoded from the sorted keys (no gather), first[p] = key[p] != key[p-1]
//   4. scan:    inclusive sum of first[] -> group[p] = rank of the distinct pair at sorted position p; n_unique
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ggl {

static int bits_for(int64_t n) {  // smallest b >= 1 with 2^b >= n
  int b = 1;
  while (b < 31 && ((int64_t)1 << b) < n) ++b;
  return b;
}

static int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void __launch_bounds__(256) pack_pairs_kernel(const int64_t* __restrict__ major, const int64_t* __restrict__ minor,
                                                         uint64_t* __restrict__ key, int32_t* __restrict__ id, int64_t E,
                                                         int64_t major_max, int64_t minor_max, int minor_bits,
                                                         int64_t* __restrict__ n_bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (; i < E; i += stride) {
    int64_t a = __ldcs(major + i), b = __ldcs(minor + i);
    // ids outside [0, n) are a caller error: counted in *n_bad (the caller raises) and folded to stay in the key width
    if (a < 0 || a > major_max) { ++bad; a = a < 0 ? 0 : major_max; }
    if (b < 0 || b > minor_max) { ++bad; b = b < 0 ? 0 : minor_max; }
    key[i] = ((uint64_t)a << minor_bits) | (uint64_t)b;
    id[i] = (int32_t)i;
  }
  if (bad && n_bad) atomicAdd(reinterpret_cast<unsigned long long*>(n_bad), bad);  // error path only
}

__global__ void __launch_bounds__(256) unpack_pairs_kernel(const uint64_t* __restrict__ key, int64_t* __restrict__ major,
                                                           int64_t* __restrict__ minor, int32_t* __restrict__ first, int64_t E,
                                                           int minor_bits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint64_t mask = ((uint64_t)1 << minor_bits) - 1;
  for (; i < E; i += stride) {
    const uint64_t k = key[i];
    if (major) major[i] = (int64_t)(k >> minor_bits);
    if (minor) minor[i] = (int64_t)(k & mask);
    if (first) first[i] = (i == 0 || key[i - 1] != k) ? 1 : 0;
  }
}

// incl[p] = number of distinct pairs in sorted positions [0, p]  ->  group[p] = incl[p] - 1, n_unique = incl[E-1]
__global__ void __launch_bounds__(256) groups_kernel(const int32_t* __restrict__ incl, int64_t* __restrict__ group,
                                                     int64_t* __restrict__ n_unique, int64_t E) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) {
    const int32_t c = incl[i];
    if (group) group[i] = (int64_t)c - 1;
    if (n_unique && i == E - 1) *n_unique = (int64_t)c;
  }
}

// keep the first pair of every group: uniq[group[p]] = sorted[p] where group[p] != group[p-1]
__global__ void __launch_bounds__(256) unique_pairs_kernel(const int64_t* __restrict__ major, const int64_t* __restrict__ minor,
                                                           const int64_t* __restrict__ group, int64_t* __restrict__ umajor,
                                                           int64_t* __restrict__ uminor, int64_t E) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) {
    const int64_t g = group[i];
    if (i == 0 || group[i - 1] != g) {
      umajor[g] = major[i];
      uminor[g] = minor[i];
    }
  }
}

static size_t sort_temp_bytes(int64_t E, int bits) {
  size_t bytes = 0;
  cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
  cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int64_t)E, 0, bits, (cudaStream_t)0);
  return bytes;
}
static size_t scan_temp_bytes(int64_t E) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int64_t)E, (cudaStream_t)0);
  return bytes;
}

}  // namespace ggl

using namespace ggl;

extern "C" size_t gglb200_sort_edges_workspace_bytes(int64_t E, int64_t n_major, int64_t n_minor) {
  if (E < 1) E = 1;
  const int bits = bits_for(n_major) + bits_for(n_minor);
  const size_t a = sort_temp_bytes(E, bits), b = scan_temp_bytes(E);
  // key x2 (double buffer) + id alt buffer + first/scan buffer + cub temp
  return 2 * align256((size_t)E * 8) + 2 * align256((size_t)E * 4) + align256(a > b ? a : b) + 256;
}

extern "C" int gglb200_sort_edges(const int64_t* major, const int64_t* minor, int64_t E, int64_t n_major, int64_t n_minor,
                                  int64_t* sorted_major, int64_t* sorted_minor, int32_t* perm, int64_t* group,
                                  int64_t* n_unique, int64_t* n_bad, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(E >= 0 && n_major >= 0 && n_minor >= 0, GGL_EINVAL, "sort_edges: negative size");
  GGL_REQUIRE(E < INT32_MAX && n_major < INT32_MAX && n_minor < INT32_MAX, GGL_ERANGE, "sort_edges: E / n exceed int32");
  if (n_bad) GGL_CUDA(cudaMemsetAsync(n_bad, 0, sizeof(int64_t), stream));
  if (E == 0) {
    if (n_unique) GGL_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int64_t), stream));
    return GGL_OK;
  }
  GGL_REQUIRE(major && minor && perm, GGL_EINVAL, "sort_edges: major/minor/perm must not be NULL");
  GGL_REQUIRE(n_major > 0 && n_minor > 0, GGL_EINVAL, "sort_edges: E > 0 needs n_major, n_minor > 0");
  GGL_REQUIRE(workspace && workspace_bytes >= gglb200_sort_edges_workspace_bytes(E, n_major, n_minor), GGL_EWORKSPACE,
              "sort_edges: workspace %zu < %zu", workspace_bytes, gglb200_sort_edges_workspace_bytes(E, n_major, n_minor));
  const int minor_bits = bits_for(n_minor), bits = minor_bits + bits_for(n_major);

  char* ws = (char*)workspace;
  const size_t kb = align256((size_t)E * 8), ib = align256((size_t)E * 4);
  uint64_t* key_a = (uint64_t*)ws;
  uint64_t* key_b = (uint64_t*)(ws + kb);
  int32_t* id_b = (int32_t*)(ws + 2 * kb);
  int32_t* first = (int32_t*)(ws + 2 * kb + ib);
  void* cub_ws = ws + 2 * kb + 2 * ib;
  size_t cub_bytes = workspace_bytes - (2 * kb + 2 * ib);

  pack_pairs_kernel<<<grid_for(E), 256, 0, stream>>>(major, minor, key_a, perm, E, n_major - 1, n_minor - 1, minor_bits, n_bad);
  GGL_LAUNCHED("pack_pairs");
  cub::DoubleBuffer<uint64_t> dk(key_a, key_b);
  cub::DoubleBuffer<int32_t> dv(perm, id_b);
  GGL_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, dk, dv, (int64_t)E, 0, bits, stream));
  g_launches.fetch_add(3, std::memory_order_relaxed);  // cub: histogram + onesweep passes (approximate)
  if (dv.Current() != perm)
    GGL_CUDA(cudaMemcpyAsync(perm, dv.Current(), (size_t)E * 4, cudaMemcpyDeviceToDevice, stream));

  const bool want_groups = group != nullptr || n_unique != nullptr;
  if (sorted_major || sorted_minor || want_groups) {
    unpack_pairs_kernel<<<grid_for(E), 256, 0, stream>>>(dk.Current(), sorted_major, sorted_minor, want_groups ? first : nullptr, E,
                                                         minor_bits);
    GGL_LAUNCHED("unpack_pairs");
  }
  if (want_groups) {
    int32_t* incl = id_b;  // free again: the sorted ids were copied back to `perm` above (stream order)
    GGL_CUDA(cub::DeviceScan::InclusiveSum(cub_ws, cub_bytes, (const int32_t*)first, incl, (int64_t)E, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    groups_kernel<<<grid_for(E), 256, 0, stream>>>(incl, group, n_unique, E);
    GGL_LAUNCHED("groups");
  }
  return GGL_OK;
}

extern "C" int gglb200_unique_edges(const int64_t* sorted_major, const int64_t* sorted_minor, const int64_t* group, int64_t E,
                                    int64_t* uniq_major, int64_t* uniq_minor, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(E >= 0, GGL_EINVAL, "unique_edges: negative size");
  if (E == 0) return GGL_OK;
  GGL_REQUIRE(sorted_major && sorted_minor && group && uniq_major && uniq_minor, GGL_EINVAL, "unique_edges: NULL pointer");
  unique_pairs_kernel<<<grid_for(E), 256, 0, stream>>>(sorted_major, sorted_minor, group, uniq_major, uniq_minor, E);
  GGL_LAUNCHED("unique_pairs");
  return GGL_OK;
}
