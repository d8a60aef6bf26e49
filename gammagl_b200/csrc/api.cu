// api.cu -- C ABI entry points of libgglb200.so: argument validation, dtype dispatch, error strings.
#include <cstdlib>

#include "spmm_impl.cuh"

namespace ggl {

std::atomic<int64_t> g_launches{0};

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int64_t plan_max_chunks_host(int64_t E) { return plan_max_chunks(E); }

int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int tile_edges() {
  static int v = [] {
    const char* e = getenv("GGLB200_TILE_EDGES");
    int t = e ? atoi(e) : GGLB200_TILE_EDGES;
    // a tile must not exceed GGLB200_LONG_ROW: a long row has to be the LAST row its tile owns (it outruns the tile window)
    if (t < 64 || t > GGLB200_LONG_ROW || (t & (t - 1))) t = GGLB200_TILE_EDGES;
    return t;
  }();
  return v;
}

int tune_ring_stages() {
  static int v = [] {
    const char* e = getenv("GGLB200_RING_STAGES");
    int t = e ? atoi(e) : 0;
    return (t == 2 || t == 3) ? t : 0;
  }();
  return v;
}

int tune_narrow() {  // aligned fp32 rows of at most 256 bytes: 1 = lean register-staged kernel, 0 = bulk-copy ring (env GGLB200_NARROW=ldg|ring)
  static int v = [] {
    const char* e = getenv("GGLB200_NARROW");
    if (e && e[0] == 'l') return 1;
    if (e && e[0] == 'r') return 0;
    return GGLB200_NARROW_DEFAULT;
  }();
  return v;
}

int tune_ring() {  // 1 = bulk-copy ring (default), 0 = register-staged kernels (env GGLB200_RING=none)
  static int v = [] {
    const char* e = getenv("GGLB200_RING");
    return (e && e[0] == 'n') ? 0 : 1;
  }();
  return v;
}

}  // namespace ggl

using namespace ggl;

extern "C" int gglb200_abi_version(void) { return GGLB200_ABI_VERSION; }
extern "C" const char* gglb200_last_error(void) { return error_buffer(); }
extern "C" int64_t gglb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static size_t bytes_per_chunk(int64_t K, int reduce) { return (size_t)K * (8 + (reduce == GGL_MAX ? 4 : 0)); }

extern "C" int gglb200_set_l2_fetch_granularity(int bytes, int* previous) {
  GGL_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, GGL_EINVAL, "l2 fetch granularity must be 32, 64 or 128");
  size_t old = 0;
  GGL_CUDA(cudaDeviceGetLimit(&old, cudaLimitMaxL2FetchGranularity));
  if (previous) *previous = (int)old;
  GGL_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return GGL_OK;
}

extern "C" size_t gglb200_spmm_workspace_bytes_for_chunks(int64_t n_chunks, int64_t K, int reduce) {
  if (n_chunks <= 0 || K <= 0) return 0;
  return (size_t)n_chunks * bytes_per_chunk(K, reduce) + 256;
}

extern "C" size_t gglb200_spmm_workspace_bytes(int64_t E, int64_t K, int reduce) {
  if (E <= GGLB200_LONG_ROW || K <= 0) return 0;
  return gglb200_spmm_workspace_bytes_for_chunks(plan_max_chunks_host(E), K, reduce);
}

static int spmm_entry(bool arg_eid, const void* const* x_table, int32_t owner_shift, int reduce, const int32_t* rowptr, const int32_t* col, const int32_t* plan,
                                const float* w, const int32_t* wperm, int32_t heads, const void* x, void* out,
                                int64_t* arg, int64_t arg_empty, const float* out_scale, const float* bias, int relu, int64_t n_rows, int64_t n_src,
                                int64_t E, int64_t K, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  GGL_REQUIRE(reduce == GGL_SUM || reduce == GGL_MEAN || reduce == GGL_MAX, GGL_EINVAL, "spmm: unknown reduce %d", reduce);
  GGL_REQUIRE(n_rows >= 0 && n_src >= 0 && E >= 0 && K >= 0, GGL_EINVAL, "spmm: negative size");
  GGL_REQUIRE(n_rows < INT32_MAX && E < INT32_MAX, GGL_ERANGE, "spmm: n_rows/E exceed int32");
  if (n_rows == 0 || K == 0) return GGL_OK;
  GGL_REQUIRE(rowptr && out && plan, GGL_EINVAL, "spmm: rowptr/out/plan must not be NULL");
  GGL_REQUIRE(E == 0 || (col && (x || x_table)), GGL_EINVAL, "spmm: col/x must not be NULL when E > 0");
  GGL_REQUIRE(heads >= 1, GGL_EINVAL, "spmm: heads must be >= 1");
  GGL_REQUIRE(heads == 1 || (w != nullptr && K % heads == 0), GGL_EINVAL, "spmm: heads=%d needs w and K %% heads == 0 (K=%lld)", heads, (long long)K);
  GGL_REQUIRE(arg == nullptr || reduce == GGL_MAX, GGL_EINVAL, "spmm: arg is only produced by GGL_MAX");
  GGL_REQUIRE(!arg_eid || (reduce == GGL_MAX && arg != nullptr && wperm != nullptr), GGL_EINVAL,
              "spmm: edge-id arg needs GGL_MAX, arg and the CSR's perm");
  GGL_REQUIRE(out_scale == nullptr || (reduce == GGL_SUM && (dtype == GGL_F32 || dtype == GGL_BF16 || dtype == GGL_F16)), GGL_EINVAL,
              "spmm: out_scale needs GGL_SUM and a floating dtype");
  GGL_REQUIRE((bias == nullptr && relu == 0) || (reduce == GGL_SUM && (dtype == GGL_F32 || dtype == GGL_BF16 || dtype == GGL_F16)), GGL_EINVAL,
              "spmm: the bias / relu epilogue needs GGL_SUM and a floating dtype");

  SpmmArgs p;
  p.reduce = reduce; p.rowptr = rowptr; p.col = col; p.plan = plan; p.w = w; p.wperm = wperm; p.heads = heads;
  p.x = x; p.x_table = x_table; p.owner_shift = owner_shift; p.out = out; p.arg = arg; p.arg_empty = arg_empty; p.arg_eid = arg_eid; p.out_scale = out_scale; p.bias = bias; p.relu = relu; p.n_rows = n_rows; p.E = E; p.K = K; p.n_src = n_src;
  p.part_val = nullptr; p.part_arg = nullptr; p.stream = (cudaStream_t)stream;
  p.n_tiles = plan_n_tiles(E);
  p.tile_row = plan + plan_tile_off(plan_max_chunks(E), plan_max_long(E));
  p.has_long = workspace != nullptr && E > GGLB200_LONG_ROW;
  p.ws_chunks = 0;
  if (p.has_long) {
    GGL_REQUIRE(workspace_bytes >= gglb200_spmm_workspace_bytes_for_chunks(1, K, reduce), GGL_EWORKSPACE,
                "spmm: workspace of %zu bytes holds no chunk", workspace_bytes);
    int64_t fit = (int64_t)((workspace_bytes - 256) / bytes_per_chunk(K, reduce));
    if (fit > plan_max_chunks(E)) fit = plan_max_chunks(E);
    p.ws_chunks = (int32_t)fit;
    p.part_val = workspace;
    p.part_arg = (int32_t*)((char*)workspace + (size_t)fit * (size_t)K * 8);
  }
  switch (dtype) {
    case GGL_F32: return launch_spmm<float>(p);
    case GGL_BF16: return launch_spmm<__nv_bfloat16>(p);
    case GGL_F16: return launch_spmm<__half>(p);
    case GGL_F64: return launch_spmm<double>(p);
    case GGL_I8: return launch_spmm<int8_t>(p);
    case GGL_U8: return launch_spmm<uint8_t>(p);
    case GGL_I16: return launch_spmm<int16_t>(p);
    case GGL_I32: return launch_spmm<int32_t>(p);
    case GGL_I64: return launch_spmm<int64_t>(p);
    default: return set_error(GGL_EUNSUPPORTED, "spmm: unknown dtype %d", dtype);
  }
}

extern "C" int gglb200_spmm_csr(int reduce, const int32_t* rowptr, const int32_t* col, const int32_t* plan,
                                const float* w, const int32_t* wperm, int32_t heads, const void* x, void* out,
                                int64_t* arg, int64_t arg_empty, const float* out_scale, int64_t n_rows, int64_t n_src,
                                int64_t E, int64_t K, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  return spmm_entry(false, nullptr, 0, reduce, rowptr, col, plan, w, wperm, heads, x, out, arg, arg_empty, out_scale, nullptr, 0, n_rows, n_src, E, K,
                    dtype, workspace, workspace_bytes, stream);
}

extern "C" int gglb200_spmm_csr_epilogue(const int32_t* rowptr, const int32_t* col, const int32_t* plan, const float* w,
                                         const int32_t* wperm, int32_t heads, const void* x, void* out, const float* out_scale,
                                         const float* bias, int relu, int64_t n_rows, int64_t n_src, int64_t E, int64_t K,
                                         int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  return spmm_entry(false, nullptr, 0, GGL_SUM, rowptr, col, plan, w, wperm, heads, x, out, nullptr, 0, out_scale, bias, relu, n_rows, n_src,
                    E, K, dtype, workspace, workspace_bytes, stream);
}

extern "C" int gglb200_spmm_csr_peer(int reduce, const int32_t* rowptr, const int32_t* col, const int32_t* plan,
                                     const float* w, const int32_t* wperm, int32_t heads, const void* const* x_table,
                                     int32_t owner_shift, void* out, const float* out_scale, int64_t n_rows, int64_t E,
                                     int64_t K, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  GGL_REQUIRE(x_table != nullptr && owner_shift >= 1 && owner_shift <= 30, GGL_EINVAL, "spmm_peer: bad x_table / owner_shift");
  GGL_REQUIRE(reduce == GGL_SUM || reduce == GGL_MEAN, GGL_EUNSUPPORTED, "spmm_peer: sum / mean only");
  return spmm_entry(false, x_table, owner_shift, reduce, rowptr, col, plan, w, wperm, heads, nullptr, out, nullptr, 0, out_scale, nullptr, 0, n_rows,
                    (int64_t)1 << owner_shift, E, K, dtype, workspace, workspace_bytes, stream);
}

extern "C" int gglb200_spmm_csr_max_edge(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* plan,
                                         const float* w, const void* x, void* out, int64_t* arg_eid, int64_t arg_empty,
                                         int64_t n_rows, int64_t n_src, int64_t E, int64_t K, int dtype, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  GGL_REQUIRE(perm != nullptr || E == 0, GGL_EINVAL, "spmm_max_edge: perm must not be NULL");
  return spmm_entry(true, nullptr, 0, GGL_MAX, rowptr, col, plan, w, perm, 1, x, out, arg_eid, arg_empty, nullptr, nullptr, 0, n_rows, n_src,
                    E, K, dtype, workspace, workspace_bytes, stream);
}
