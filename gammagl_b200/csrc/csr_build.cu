// csr_build.cu -- COO (int64, unsorted) -> CSR (int32, stable by row) ingest on the GPU.
//
// Replaces, for the aggregation path, SparseStorage's host-side numpy argsort + ind2ptr
// (/root/reference/gammagl/sparse/storage.py:124-140, gammagl/ops/sparse/cpu/convert.cpp:58-90) and the
// per-element binary-search kernels of gammagl/ops/sparse/cuda/convert.cu:41-60.
//
// Pipeline (all on `stream`, no host sync, caller-owned workspace):
//   1. narrow:   key32[i] = (int32) key[i],  id[i] = i; keys outside [0,N) are COUNTED in plan[6] (and folded to the
//                ends of the sort domain so that nothing is ever dereferenced out of range); the caller reads plan[6..7]
//                at its one host sync and raises -- the reference's CPU operators raise too (TORCH_CHECK_INDEX idx < N,
//                cpu/segment_max_cpu.cpp:50)
//   2. sort:     cub::DeviceRadixSort::SortPairs over ceil(log2 N) bits  -- LSD radix = STABLE, so
//                inside a row edge ids stay ascending (the order of the reference's serial loops)
//   3. rowptr:   boundary detection on the sorted keys (each thread fills the rowptr entries between
//                its key and its predecessor's -- handles empty rows without a scan)
//   4. col:      col[i] = (int32) other[perm[i]]
//   5. plan:     long-row work list (rows with more than GGLB200_LONG_ROW edges are cut into chunks of
//                GGLB200_LONG_CHUNK edges; see spmm.cu)
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ggl {

// plan layout: see common.cuh

__global__ void __launch_bounds__(256) narrow_keys_kernel(const int64_t* __restrict__ key, int32_t* __restrict__ key32,
                                                          int32_t* __restrict__ id, int64_t E, int32_t n_minus_1,
                                                          int32_t* __restrict__ bad_keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int32_t bad = 0;
  for (; i < E; i += stride) {
    int64_t k = key[i];
    if (k < 0 || k > n_minus_1) { ++bad; k = k < 0 ? 0 : n_minus_1; }
    key32[i] = (int32_t)k;
    id[i] = (int32_t)i;
  }
  if (bad) atomicAdd(bad_keys, bad);  // error path only
}

// rowptr[r] = first sorted position whose key >= r.  Thread i (0..E) owns rows (key[i-1], key[i]].
__global__ void __launch_bounds__(256) rowptr_kernel(const int32_t* __restrict__ skey, int32_t* __restrict__ rowptr,
                                                     int64_t E, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i <= E; i += stride) {
    int64_t lo = (i == 0) ? 0 : (int64_t)skey[i - 1] + 1;
    int64_t hi = (i == E) ? N : (int64_t)skey[i];
    for (int64_t r = lo; r <= hi; ++r) rowptr[r] = (int32_t)i;
  }
}

__global__ void __launch_bounds__(256) gather_col_kernel(const int64_t* __restrict__ other, const int32_t* __restrict__ perm,
                                                         int32_t* __restrict__ col, int64_t E, int64_t n_other,
                                                         int32_t* __restrict__ bad_other) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int32_t bad = 0;
  for (; i < E; i += stride) {
    int64_t c = other[perm[i]];
    if (n_other >= 0 && (c < 0 || c >= n_other)) { ++bad; c = 0; }  // never hand an out-of-range row id to the gather kernels
    col[i] = (int32_t)c;
  }
  if (bad) atomicAdd(bad_other, bad);  // error path only
}

__global__ void __launch_bounds__(256) plan_kernel(const int32_t* __restrict__ rowptr, int32_t* __restrict__ plan, int64_t N) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int32_t max_chunks = plan[2], max_long = plan[3];
  int32_t* chunk_tab = plan + plan_chunk_off();
  int32_t* long_tab = plan + plan_long_off(max_chunks);
  for (; r < N; r += stride) {
    int32_t s = rowptr[r], e = rowptr[r + 1];
    int32_t deg = e - s;
    if (deg > GGLB200_LONG_ROW) {
      int32_t nch = (deg + GGLB200_LONG_CHUNK - 1) / GGLB200_LONG_CHUNK;
      int32_t l = atomicAdd(&plan[1], 1);
      int32_t c0 = atomicAdd(&plan[0], nch);
      if (l < max_long && c0 + nch <= max_chunks) {  // always true by construction of the bounds
        long_tab[3 * l + 0] = (int32_t)r;
        long_tab[3 * l + 1] = c0;
        long_tab[3 * l + 2] = nch;
        for (int32_t j = 0; j < nch; ++j) {
          chunk_tab[2 * (c0 + j) + 0] = (int32_t)r;
          chunk_tab[2 * (c0 + j) + 1] = s + j * GGLB200_LONG_CHUNK;
        }
      }
    }
  }
}

__global__ void plan_init_kernel(int32_t* plan, int32_t max_chunks, int32_t max_long, int32_t n_tiles, int32_t te) {
  plan[0] = 0; plan[1] = 0; plan[2] = max_chunks; plan[3] = max_long;
  plan[4] = n_tiles; plan[5] = te; plan[6] = 0; plan[7] = 0;
}

// tile table entry t = {r0 | long_last << 31, rowptr[r0]} with r0 = lower_bound(rowptr[0..N], t * TILE_EDGES); entry n_tiles = {N, E}
__device__ __forceinline__ int64_t first_row_at(const int32_t* __restrict__ rowptr, int64_t N, int64_t target) {
  int64_t lo = 0, hi = N;  // first r in [0,N] with rowptr[r] >= target
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rowptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void __launch_bounds__(256) tile_rows_kernel(const int32_t* __restrict__ rowptr, int32_t* __restrict__ tile_tab,
                                                        int64_t N, int64_t n_tiles, int64_t te) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t <= n_tiles; t += stride) {
    if (t == n_tiles) { tile_tab[2 * t] = (int32_t)N; tile_tab[2 * t + 1] = rowptr[N]; continue; }
    const int64_t r0 = t == 0 ? 0 : first_row_at(rowptr, N, t * te);
    const int64_t r1 = t + 1 == n_tiles ? N : first_row_at(rowptr, N, (t + 1) * te);
    const bool long_last = r1 > r0 && rowptr[r1] - rowptr[r1 - 1] > GGLB200_LONG_ROW;
    tile_tab[2 * t] = (int32_t)r0 | (long_last ? (int32_t)0x80000000 : 0);
    tile_tab[2 * t + 1] = rowptr[r0];
  }
}

// ind2ptr / ptr2ind on int64, the types of gammagl/ops/sparse/sparse.py:38-75 (cpu/convert.cpp:58-128)
__global__ void __launch_bounds__(256) ind2ptr_i64_kernel(const int64_t* __restrict__ ind, int64_t* __restrict__ ptr, int64_t E, int64_t M) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i <= E; i += stride) {
    int64_t lo = (i == 0) ? 0 : ind[i - 1] + 1;
    int64_t hi = (i == E) ? M : ind[i];
    for (int64_t r = lo; r <= hi; ++r) ptr[r] = i;
  }
}
__global__ void __launch_bounds__(256) ptr2ind_i64_kernel(const int64_t* __restrict__ ptr, int64_t* __restrict__ ind, int64_t M, int64_t E) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < E; e += stride) {
    int64_t lo = 0, hi = M;  // last r in [0, M) with ptr[r] <= e
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (ptr[mid] <= e) lo = mid; else hi = mid;
    }
    ind[e] = lo;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) permute_rows_kernel(const T* __restrict__ in, const int32_t* __restrict__ perm,
                                                           T* __restrict__ out, int64_t n, int32_t width) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t total = n * width;
  for (; t < total; t += stride) {
    int64_t i = t / width;
    int32_t j = (int32_t)(t - i * width);
    out[t] = in[(int64_t)perm[i] * width + j];
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int sort_bits(int64_t N) {
  int b = 1;
  while (b < 31 && ((int64_t)1 << b) < N) ++b;
  return b;
}

static size_t cub_temp_bytes(int64_t E, int64_t N) {
  size_t bytes = 0;
  cub::DoubleBuffer<int32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int64_t)E, 0, sort_bits(N), (cudaStream_t)0);
  return bytes;
}

static int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace ggl

using namespace ggl;

extern "C" size_t gglb200_csr_plan_ints(int64_t E) {
  if (E < 0) E = 0;
  return (size_t)(plan_tile_off(plan_max_chunks(E), plan_max_long(E)) + 2 * (plan_n_tiles(E) + 1));
}

extern "C" size_t gglb200_csr_build_workspace_bytes(int64_t E, int64_t N) {
  if (E < 0) E = 0;
  if (N < 1) N = 1;
  // key32 x2 (double buffer) + id alt buffer + cub temp
  return 3 * align256((size_t)E * 4) + align256(cub_temp_bytes(E > 0 ? E : 1, N)) + 256;
}

extern "C" int gglb200_csr_build(const int64_t* key, const int64_t* other, int64_t E, int64_t N, int64_t n_other,
                                 int32_t* rowptr, int32_t* col, int32_t* perm, int32_t* plan,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(E >= 0 && N >= 0, GGL_EINVAL, "csr_build: negative size E=%lld N=%lld", (long long)E, (long long)N);
  GGL_REQUIRE(E < INT32_MAX && N < INT32_MAX, GGL_ERANGE, "csr_build: E=%lld / N=%lld exceed int32", (long long)E, (long long)N);
  GGL_REQUIRE(rowptr && plan, GGL_EINVAL, "csr_build: rowptr/plan must not be NULL");
  GGL_REQUIRE(E == 0 || (key && perm), GGL_EINVAL, "csr_build: key/perm must not be NULL");
  GGL_REQUIRE(col == nullptr || other != nullptr || E == 0, GGL_EINVAL, "csr_build: col requested but other is NULL");
  GGL_REQUIRE(workspace_bytes >= gglb200_csr_build_workspace_bytes(E, N), GGL_EWORKSPACE,
              "csr_build: workspace %zu < %zu", workspace_bytes, gglb200_csr_build_workspace_bytes(E, N));

  const int64_t n_tiles = plan_n_tiles(E);
  int32_t* tile_row = plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E));
  plan_init_kernel<<<1, 1, 0, stream>>>(plan, (int32_t)plan_max_chunks(E), (int32_t)plan_max_long(E), (int32_t)n_tiles, (int32_t)tile_edges());
  GGL_LAUNCHED("plan_init");
  if (E == 0) {
    GGL_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (size_t)(N + 1), stream));
    tile_rows_kernel<<<1, 256, 0, stream>>>(rowptr, tile_row, N, n_tiles, (int64_t)tile_edges());
    GGL_LAUNCHED("tile_rows");
    return GGL_OK;
  }
  GGL_REQUIRE(N > 0, GGL_EINVAL, "csr_build: E > 0 but N == 0");

  char* ws = (char*)workspace;
  size_t eb = align256((size_t)E * 4);
  int32_t* key_a = (int32_t*)ws;
  int32_t* key_b = (int32_t*)(ws + eb);
  int32_t* id_b = (int32_t*)(ws + 2 * eb);
  void* cub_ws = ws + 3 * eb;
  size_t cub_bytes = workspace_bytes - 3 * eb;

  // ids start in `perm` (buffer 0) and ping-pong with id_b; whichever side the sort ends on is copied back
  narrow_keys_kernel<<<grid_for(E), 256, 0, stream>>>(key, key_a, perm, E, (int32_t)(N - 1), plan + 6);
  GGL_LAUNCHED("narrow_keys");

  cub::DoubleBuffer<int32_t> dk(key_a, key_b), dv(perm, id_b);
  GGL_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, dk, dv, (int64_t)E, 0, sort_bits(N), stream));
  g_launches.fetch_add(3, std::memory_order_relaxed);  // cub: histogram + onesweep passes (approximate)
  if (dv.Current() != perm)
    GGL_CUDA(cudaMemcpyAsync(perm, dv.Current(), (size_t)E * 4, cudaMemcpyDeviceToDevice, stream));

  rowptr_kernel<<<grid_for(E + 1), 256, 0, stream>>>(dk.Current(), rowptr, E, N);
  GGL_LAUNCHED("rowptr");
  if (col) {
    gather_col_kernel<<<grid_for(E), 256, 0, stream>>>(other, perm, col, E, n_other, plan + 7);
    GGL_LAUNCHED("gather_col");
  }
  plan_kernel<<<grid_for(N), 256, 0, stream>>>(rowptr, plan, N);
  GGL_LAUNCHED("plan");
  tile_rows_kernel<<<grid_for(n_tiles + 1), 256, 0, stream>>>(rowptr, tile_row, N, n_tiles, (int64_t)tile_edges());
  GGL_LAUNCHED("tile_rows");
  return GGL_OK;
}

extern "C" int gglb200_permute_rows(const void* in, const int32_t* perm, void* out, int64_t n, int64_t row_bytes,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(n >= 0 && row_bytes > 0, GGL_EINVAL, "permute_rows: bad sizes");
  if (n == 0) return GGL_OK;
  GGL_REQUIRE(in && perm && out, GGL_EINVAL, "permute_rows: NULL pointer");
  if (row_bytes % 4 == 0) {
    permute_rows_kernel<uint32_t><<<grid_for(n * (row_bytes / 4)), 256, 0, stream>>>(
        (const uint32_t*)in, perm, (uint32_t*)out, n, (int32_t)(row_bytes / 4));
  } else {
    permute_rows_kernel<uint8_t><<<grid_for(n * row_bytes), 256, 0, stream>>>(
        (const uint8_t*)in, perm, (uint8_t*)out, n, (int32_t)row_bytes);
  }
  GGL_LAUNCHED("permute_rows");
  return GGL_OK;
}

extern "C" int gglb200_ind2ptr(const int64_t* ind, int64_t E, int64_t M, int64_t* ptr, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(E >= 0 && M >= 0, GGL_EINVAL, "ind2ptr: negative size");
  GGL_REQUIRE(ptr && (E == 0 || ind), GGL_EINVAL, "ind2ptr: NULL pointer");
  if (E == 0) {
    GGL_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int64_t) * (size_t)(M + 1), stream));
    return GGL_OK;
  }
  ind2ptr_i64_kernel<<<grid_for(E + 1), 256, 0, stream>>>(ind, ptr, E, M);
  GGL_LAUNCHED("ind2ptr");
  return GGL_OK;
}

extern "C" int gglb200_ptr2ind(const int64_t* ptr, int64_t M, int64_t E, int64_t* ind, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GGL_REQUIRE(E >= 0 && M >= 0, GGL_EINVAL, "ptr2ind: negative size");
  if (E == 0) return GGL_OK;
  GGL_REQUIRE(ptr && ind && M > 0, GGL_EINVAL, "ptr2ind: NULL pointer or no rows");
  ptr2ind_i64_kernel<<<grid_for(E), 256, 0, stream>>>(ptr, ind, M, E);
  GGL_LAUNCHED("ptr2ind");
  return GGL_OK;
}
