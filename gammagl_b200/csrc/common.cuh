// common.cuh -- shared helpers of libgglb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <limits>

#include "../../include/gglb200.h"

namespace ggl {

// ---- error reporting (thread local, no exceptions cross the C ABI) -------------------------------
char* error_buffer();
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define GGL_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return ::ggl::set_error((code), __VA_ARGS__); \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors (no sync)
#define GGL_LAUNCHED(name)                                                           \
  do {                                                                               \
    ::ggl::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess)                                                          \
      return ::ggl::set_error((int)e__, "%s launch failed: %s", (name), cudaGetErrorString(e__)); \
  } while (0)

#define GGL_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return ::ggl::set_error((int)e__, "%s: %s", #call, cudaGetErrorString(e__));     \
  } while (0)

// default path of aligned fp32 rows <= 256 bytes (decided by measurement: tools/probe_narrow.py, profiles/README.md)
#ifndef GGLB200_NARROW_DEFAULT
#define GGLB200_NARROW_DEFAULT 0
#endif
int sm_count();  // SMs of the current device (148 on B200: 2 dies x 74), queried once per device; api.cu

// ---- plan layout (int32 array produced by gglb200_csr_build, consumed by the reduce kernels) --------
//   [0] n_chunks  [1] n_long_rows  [2] max_chunks (capacity)  [3] max_long (capacity)
//   [4] n_tiles   [5] GGLB200_TILE_EDGES  [6..7] reserved
//   [8 .. 8+2*max_chunks)            chunk table: {row, first_edge}          (rows > GGLB200_LONG_ROW edges)
//   [.. +3*max_long)                 long-row table: {row, first_chunk, n_chunks}
//   [.. +2*(n_tiles+1)) (8-byte aligned)  tile table, int2 per tile: {r0 | long_last << 31, e0} with r0 = first row whose
//                                    first edge is >= t*TILE_EDGES and e0 = rowptr[r0]; entry n_tiles = {N, E}.  Tile t
//                                    owns rows [r0(t), r0(t+1)) = edges [e0(t), e0(t+1)); long_last marks tiles whose
//                                    last owned row has more than GGLB200_LONG_ROW edges (left to the chunk path)
constexpr int kPlanHeader = 8;
__host__ __device__ inline int64_t plan_max_long(int64_t E) { return E / GGLB200_LONG_ROW + 1; }
__host__ __device__ inline int64_t plan_max_chunks(int64_t E) { return E / GGLB200_LONG_CHUNK + plan_max_long(E) + 1; }
int tile_edges();  // edges per tile: GGLB200_TILE_EDGES unless overridden by env GGLB200_TILE_EDGES (tuning), api.cu
inline int64_t plan_n_tiles(int64_t E) {
  const int64_t te = tile_edges();
  int64_t t = (E + te - 1) / te;
  return t < 1 ? 1 : t;
}
__host__ __device__ inline int64_t plan_chunk_off() { return kPlanHeader; }
__host__ __device__ inline int64_t plan_long_off(int64_t max_chunks) { return kPlanHeader + 2 * max_chunks; }
__host__ __device__ inline int64_t plan_tile_off(int64_t max_chunks, int64_t max_long) {
  return (kPlanHeader + 2 * max_chunks + 3 * max_long + 1) & ~(int64_t)1;
}

// ---- element traits ------------------------------------------------------------------------------
template <typename T> struct Elem;  // acc_t, to_acc, from_acc, lowest
template <> struct Elem<float> {
  using acc_t = float;
  static constexpr bool is_float = true;
  __device__ static float to_acc(float v) { return v; }
  __device__ static float from_acc(float v) { return v; }
  __device__ static float lowest() { return -3.402823466e+38f; }  // numeric_limits<float>::lowest()
};
template <> struct Elem<__nv_bfloat16> {
  using acc_t = float;
  static constexpr bool is_float = true;
  __device__ static float to_acc(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __nv_bfloat16 from_acc(float v) { return __float2bfloat16_rn(v); }
  __device__ static float lowest() { return -3.38953139e+38f; }  // bf16 lowest (0xFF7F)
};
template <> struct Elem<__half> {
  using acc_t = float;
  static constexpr bool is_float = true;
  __device__ static float to_acc(__half v) { return __half2float(v); }
  __device__ static __half from_acc(float v) { return __float2half_rn(v); }
  __device__ static float lowest() { return -65504.f; }
};
template <> struct Elem<double> {
  using acc_t = double;
  static constexpr bool is_float = true;
  __device__ static double to_acc(double v) { return v; }
  __device__ static double from_acc(double v) { return v; }
  __device__ static double lowest() { return -1.7976931348623157e+308; }
};
template <typename I> struct IntElem {
  using acc_t = int64_t;
  static constexpr bool is_float = false;
  __device__ static int64_t to_acc(I v) { return (int64_t)v; }
  __device__ static I from_acc(int64_t v) { return (I)v; }  // wraps like scalar_t += in the reference
  __device__ static int64_t lowest() { return (int64_t)std::numeric_limits<I>::lowest(); }
};
template <> struct Elem<int8_t> : IntElem<int8_t> {};
template <> struct Elem<uint8_t> : IntElem<uint8_t> {};
template <> struct Elem<int16_t> : IntElem<int16_t> {};
template <> struct Elem<int32_t> : IntElem<int32_t> {};
template <> struct Elem<int64_t> : IntElem<int64_t> {};

// ---- vector views: VEC consecutive elements of T moved with one load/store ------------------------
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

// read-only gather of feature rows: ld.global.nc, no L1 allocation (a row is used once per warp)
template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_gather(const T* p) {
  Pack<T, VEC> r;
  constexpr int B = sizeof(T) * VEC;
  if constexpr (B == 16) {
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    uint32_t* q = reinterpret_cast<uint32_t*>(&r);
    q[0] = a; q[1] = b; q[2] = c; q[3] = d;
  } else if constexpr (B == 8) {
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
    uint32_t* q = reinterpret_cast<uint32_t*>(&r);
    q[0] = a; q[1] = b;
  } else if constexpr (B == 4) {
    uint32_t a;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
    *reinterpret_cast<uint32_t*>(&r) = a;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __ldg(p + i);
  }
  return r;
}

// streaming store of finished output rows (written once, not re-read by this kernel)
template <typename T, int VEC>
__device__ __forceinline__ void st_stream(T* p, const Pack<T, VEC>& r) {
  constexpr int B = sizeof(T) * VEC;
  if constexpr (B == 16) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(&r);
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
  } else if constexpr (B == 8) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(&r);
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(q[0]), "r"(q[1]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
  }
}

// streaming (evict-first) loads of the index / weight streams: each element is read exactly once
__device__ __forceinline__ int32_t ld_stream_i32(const int32_t* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream_f32(const float* p) { return __ldcs(p); }

}  // namespace ggl
