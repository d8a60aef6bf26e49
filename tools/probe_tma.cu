// probe_tma.cu -- what bounds a random ROW GATHER on B200, per row width: the copy-issue rate of the TMA unit or DRAM?
// Three ways to bring row idx[i] of x[N, W bytes] into an SM, same ring structure as spmm_ring_kernel (one-warp blocks,
// 2 stages x 8 rows, mbarrier per stage, every lane reads its word of the row and accumulates a checksum):
//   bulk     cp.async.bulk.shared::cluster.global          one UBLKCP per row      (what the shipped kernels do)
//   gather4  cp.async.bulk.tensor.2d...tile::gather4       one UTMALDG.GATHER4 per FOUR rows (Blackwell only)
//   ldg      ld.global.nc.v4 into registers, 8 rows in flight per lane group       (no TMA)
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/probe_tma tools/probe_tma.cu
// Run:    build/probe_tma            (prints one JSON line per (method, row bytes))
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int R = 8, STAGES = 2;

// MODE 0 = bulk, 1 = gather4.  Block = one warp; block b streams rows idx[b*per .. (b+1)*per).
template <int MODE>
__global__ void __launch_bounds__(32) ring_probe(const __grid_constant__ CUtensorMap tm, const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                 int64_t per, int32_t wbytes, int32_t slot_bytes /* bytes of 4 rows, padded to 128 */,
                                                 float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[STAGES];
  const int lane = threadIdx.x;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(bars);
  if (lane == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(b0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const int32_t* my = idx + (int64_t)blockIdx.x * per;
  const int64_t nslots = per / R;
  const int words = wbytes / 4;
  const uint32_t stage_bytes = 2 * slot_bytes;   // 8 rows = two groups of 4
  auto issue = [&](int64_t s) {
    if (s >= nslots) return;
    const int stg = (int)(s % STAGES);
    const uint32_t dst = s0 + stg * stage_bytes, bar = b0 + 8 * stg;
    const int32_t r = lane < R ? __ldg(my + s * R + lane) : 0;
    __syncwarp();
    if (lane == 0) mbar_expect(bar, (uint32_t)R * wbytes);
    if (MODE == 0) {
      if (lane < R) {
        const uint32_t d = dst + (lane >> 2) * slot_bytes + (lane & 3) * wbytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(x + (int64_t)r * words), "r"((uint32_t)wbytes), "r"(bar) : "memory");
      }
    } else {
      const int g = lane;  // lanes 0 and 1 issue one gather4 each
      const int32_t r0 = __shfl_sync(0xffffffffu, r, (g & 1) * 4 + 0), r1 = __shfl_sync(0xffffffffu, r, (g & 1) * 4 + 1);
      const int32_t r2 = __shfl_sync(0xffffffffu, r, (g & 1) * 4 + 2), r3 = __shfl_sync(0xffffffffu, r, (g & 1) * 4 + 3);
      if (lane < 2) {
        const uint32_t d = dst + g * slot_bytes;
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(d), "l"(&tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
      }
    }
  };
  for (int s = 0; s < STAGES - 1; ++s) issue(s);
  float acc = 0.f;
  uint32_t parity = 0;
  for (int64_t s = 0; s < nslots; ++s) {
    issue(s + STAGES - 1);
    const int stg = (int)(s % STAGES);
    mbar_wait(b0 + 8 * stg, parity);
    if (stg == STAGES - 1) parity ^= 1u;
    const unsigned char* base = smem + stg * stage_bytes;
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const float* row = reinterpret_cast<const float*>(base + (u >> 2) * slot_bytes + (u & 3) * wbytes);
      for (int k = lane; k < words; k += 32) acc += row[k];
    }
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[blockIdx.x] = acc;
}

// register-staged: a group of LPR lanes (16 bytes per lane) owns a stream, 8 rows in flight
template <int LPR>
__global__ void __launch_bounds__(256) ldg_probe(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t per, int32_t words,
                                                 float* __restrict__ out) {
  const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int32_t* my = idx + grp * per;
  float acc = 0.f;
  for (int64_t s = 0; s < per; s += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int32_t r = __ldg(my + s + u);
      v[u] = make_float4(0, 0, 0, 0);
      if (sub * 4 < words) v[u] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)r * words) + sub);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (int o = LPR / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (sub == 0) out[grp] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  // argv[1]: which methods to run, a string of digits (0 = bulk, 1 = gather4, 2 = ldg; default "02"); argv[2]: box rows of the
  // gather4 tensor map (1 or 4).  A wrong tensor map kills the context, so gather4 runs in its own process.
  const char* which = argc > 1 ? argv[1] : "02";
  const int box_rows = argc > 2 ? atoi(argv[2]) : 1;
  const int64_t M = 64ll << 20;                 // rows gathered per run
  const size_t xbytes = 4ull << 30;             // 4 GB of rows: far beyond the 126 MB L2
  float* x; int32_t* idx; float* out;
  CK(cudaMalloc(&x, xbytes));
  CK(cudaMalloc(&idx, M * 4));
  CK(cudaMalloc(&out, 64 << 20));
  {  // fill x with small integers (exact sums whatever the order)
    std::vector<float> h(1 << 20);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u >> 20) & 7);
    for (size_t o = 0; o < xbytes; o += h.size() * 4) CK(cudaMemcpy((char*)x + o, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int widths[] = {16, 32, 64, 128, 256, 400, 512};
  for (int wbytes : widths) {
    const int words = wbytes / 4;
    const int64_t N = (int64_t)(xbytes / wbytes);
    {  // random row ids (LCG on the host: deterministic)
      std::vector<int32_t> h(M);
      uint64_t sd = 88172645463325252ull;
      for (int64_t i = 0; i < M; ++i) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; h[i] = (int32_t)(sd % (uint64_t)N); }
      CK(cudaMemcpy(idx, h.data(), M * 4, cudaMemcpyHostToDevice));
    }
    const int slot_bytes = (4 * wbytes + 127) & ~127;
    double sums[3] = {0, 0, 0};
    for (int mode = 0; mode < 3; ++mode) {
      if (!strchr(which, '0' + mode)) continue;
      CUtensorMap tm;
      memset(&tm, 0, sizeof(tm));
      bool ok = true;
      const int64_t per = 1024;                 // rows per block / group
      const int64_t blocks = M / per;
      if (mode == 1) {
        cuuint64_t dims[2] = {(cuuint64_t)words, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)wbytes};
        cuuint32_t box[2] = {(cuuint32_t)words, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("{\"method\":\"gather4\",\"row_bytes\":%d,\"error\":\"encode failed %d\"}\n", wbytes, (int)r); ok = false; }
      }
      if (!ok) continue;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      float best = 1e30f;
      const size_t smem = (size_t)STAGES * 2 * slot_bytes;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) ring_probe<0><<<(unsigned)blocks, 32, smem>>>(tm, x, idx, per, wbytes, slot_bytes, out);
        else if (mode == 1) ring_probe<1><<<(unsigned)blocks, 32, smem>>>(tm, x, idx, per, wbytes, slot_bytes, out);
        else {
          int lpr = 1; while (lpr < 32 && lpr * 4 < words) lpr <<= 1;
          const int64_t threads = blocks * lpr;
          const unsigned grid = (unsigned)((threads + 255) / 256);
          switch (lpr) {
            case 1: ldg_probe<1><<<grid, 256>>>(x, idx, per, words, out); break;
            case 2: ldg_probe<2><<<grid, 256>>>(x, idx, per, words, out); break;
            case 4: ldg_probe<4><<<grid, 256>>>(x, idx, per, words, out); break;
            case 8: ldg_probe<8><<<grid, 256>>>(x, idx, per, words, out); break;
            case 16: ldg_probe<16><<<grid, 256>>>(x, idx, per, words, out); break;
            default: ldg_probe<32><<<grid, 256>>>(x, idx, per, words, out); break;
          }
        }
        CK(cudaEventRecord(e1));
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { printf("{\"method\":%d,\"row_bytes\":%d,\"error\":\"%s\"}\n", mode, wbytes, cudaGetErrorString(err)); return 1; }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      std::vector<float> h(blocks);
      CK(cudaMemcpy(h.data(), out, blocks * 4, cudaMemcpyDeviceToHost));
      for (float v : h) sums[mode] += v;
      const char* name = mode == 0 ? "bulk (UBLKCP per row)" : mode == 1 ? "gather4 (UTMALDG per 4 rows)" : "ldg (registers)";
      printf("{\"method\":\"%s\",\"row_bytes\":%d,\"ms\":%.3f,\"rows_per_s\":%.4g,\"GBps\":%.1f,\"clk_per_row_per_sm\":%.2f,\"checksum\":%.0f}\n", name, wbytes,
             best, M / (best * 1e-3), (double)M * wbytes / (best * 1e-3) / 1e9, best * 1e-3 * 1.965e9 * sms / M, sums[mode]);
      fflush(stdout);
    }
    if ((sums[0] != 0 && sums[2] != 0 && sums[0] != sums[2]) || (sums[1] != 0 && sums[2] != 0 && sums[1] != sums[2]))
      printf("{\"row_bytes\":%d,\"checksum_mismatch\":[%.0f,%.0f,%.0f]}\n", wbytes, sums[0], sums[1], sums[2]);
  }
  return 0;
}
