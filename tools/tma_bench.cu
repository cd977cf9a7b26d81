// micro-benchmark: how should 32 lanes that each own a separate output stream write to HBM?
//  mode 0: per-lane ST.128 straight to global (32 lines per instruction)
//  mode 1: per-lane staging row in smem, drained by a per-lane cp.async.bulk (TMA) of S bytes
//  mode 2: per-lane staging row, drained cooperatively (2 rows per LDS.128/STG.128 pair)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int S, int MODE>
__global__ void __launch_bounds__(256) k(uint8_t* out, uint64_t stride, int chunks, uint64_t ngroups) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int wid = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint8_t* row = sm + (size_t)(wid * 32 + l) * (2 * S + 16);  // +16: rows on distinct bank groups
  const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(row);
  const uint64_t nw = (uint64_t)gridDim.x * 8;
  for (uint64_t g = (uint64_t)blockIdx.x * 8 + wid; g < ngroups; g += nw) {
    uint8_t* dst = out + (g * 32 + l) * stride;  // stride multiple of 16
    for (int c = 0; c < chunks; c++) {
      uint4 v = make_uint4((uint32_t)g, (uint32_t)c, l, 7);
      if (MODE == 0) {
        for (int i = 0; i < S; i += 16) { v.w = i; *(uint4*)(dst + (size_t)c * S + i) = v; }
      } else if (MODE == 1) {
        const uint32_t half = row_s + (c & 1) * S;
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        for (int i = 0; i < S; i += 16) { v.w = i; asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(half + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * S), "r"(half), "r"(S) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else {
        for (int i = 0; i < S; i += 16) { v.w = i; asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(row_s + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
        __syncwarp();
        constexpr int LPR = S / 16;          // lanes per row
        constexpr int RPI = 32 / LPR;        // rows per iteration
        const int sub = l / LPR, t = l % LPR;
        for (int j = 0; j < 32; j += RPI) {
          const int rj = j + sub;
          const uint32_t src = row_s + (uint32_t)((rj - l) * (2 * S + 16)) + t * 16;
          uint4 w;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(src) : "memory");
          uint8_t* d = out + (g * 32 + rj) * stride + (size_t)c * S + t * 16;
          *(uint4*)d = w;
        }
        __syncwarp();
      }
    }
  }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int S, int MODE>
int run(uint8_t* out, uint64_t bytes, const char* name) {
  const uint64_t stride = 2304;  // ~ one JSONL line
  const int chunks = 2048 / S;   // 2 KB written per stream
  const uint64_t ngroups = bytes / (stride * 32);
  const size_t smem = (size_t)256 * (2 * S + 16);
  CK(cudaFuncSetAttribute(k<S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<S, MODE>, 256, smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; it++) {
    cudaEventRecord(e0);
    k<S, MODE><<<148 * per_sm, 256, smem>>>(out, stride, chunks, ngroups);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double gb = (double)ngroups * 32 * 2048 / 1e9;
  printf("%-28s S=%4d ctas/sm=%d  %.3f ms  %.0f GB/s (payload)\n", name, S, per_sm, best, gb / best * 1e3);
  return 0;
}

int main() {
  const uint64_t bytes = 4ull << 30;
  uint8_t* out;
  CK(cudaMalloc(&out, bytes));
  CK(cudaMemset(out, 0, bytes));
  if (run<128, 0>(out, bytes, "direct ST.128")) return 1;
  if (run<128, 1>(out, bytes, "smem row + cp.async.bulk")) return 1;
  if (run<256, 1>(out, bytes, "smem row + cp.async.bulk")) return 1;
  if (run<512, 1>(out, bytes, "smem row + cp.async.bulk")) return 1;
  if (run<128, 2>(out, bytes, "smem row + coop copy-out")) return 1;
  if (run<256, 2>(out, bytes, "smem row + coop copy-out")) return 1;
  if (run<512, 2>(out, bytes, "smem row + coop copy-out")) return 1;
  return 0;
}
