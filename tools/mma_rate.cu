// Micro-benchmark: tcgen05.mma cycles per instruction for small-N tiles (M=128, K=16, fp16 -> fp32) as a function
// of N, the operand source of A (shared memory descriptor vs tensor memory) and the number of independent
// accumulators the instruction stream rotates over.  Not part of the product; it answers one design question for
// the quantised GEMM: is a chain of accumulating MMAs on ONE accumulator latency-bound at small N?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I lite_llama_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate && gpurun_out/mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "common.cuh"
void b2l_count_launch() {}
void b2l_set_error(const char*, ...) {}

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct Args { int ts; int n; int nacc; int count; int commit_every; long long* out; };

template <int TS, int NACC, int CE>
__global__ void __launch_bounds__(128) mma_rate_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 16384;         // 256 rows x 128 B
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc(1, 0, 0, 0, 0, 128, a.n);
    const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sA));
    const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sB));
    long long t0 = 0, t1 = 0, t2 = 0;
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      if (elect_one()) {
        t0 = clock64();
        for (int i0 = 0; i0 < a.count; i0 += 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t d = tmem + (uint32_t)((j % NACC) * a.n);
            const uint32_t k = (uint32_t)(j & 3) * 2;
            if (TS) umma_ts(d, tmem + 448 + (j & 3) * 8, bdesc + k, idesc, (i0 | j) >= NACC);
            else umma_f16(d, adesc + k, bdesc + k, idesc, (i0 | j) >= NACC);
            if (CE && (j + 1) % CE == 0) tc_commit(&bar2);  // never waited on
          }
        }
        t1 = clock64();
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase); phase ^= 1;
      t2 = clock64();
      tc_fence_after();
    }
    if (threadIdx.x == 0 || true) {
      if (t0 != 0) { a.out[0] = t1 - t0; a.out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int TS, int NACC, int CE> static void launch3(const Args& a) {
  cudaFuncSetAttribute(mma_rate_kernel<TS, NACC, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
  mma_rate_kernel<TS, NACC, CE><<<1, 128, 16384 + 32768 + 1024>>>(a);
}
template <int TS, int NACC> static void launch2(const Args& a) {
  if (a.commit_every == 0) launch3<TS, NACC, 0>(a);
  else if (a.commit_every == 8) launch3<TS, NACC, 8>(a);
  else launch3<TS, NACC, 4>(a);
}
template <int TS> static void launch1(const Args& a) {
  switch (a.nacc) {
    case 1: launch2<TS, 1>(a); break;
    case 2: launch2<TS, 2>(a); break;
    case 4: launch2<TS, 4>(a); break;
    default: launch2<TS, 8>(a); break;
  }
}
static void launch(const Args& a) { if (a.ts) launch1<1>(a); else launch1<0>(a); }

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  printf("%-4s %4s %5s %6s %7s | %10s %10s\n", "A", "N", "nacc", "count", "commit", "issue/mma", "total/mma");
  const int ns[] = {16, 32, 64, 128, 256};
  for (int ts = 0; ts <= 1; ++ts)
    for (int n : ns)
      for (int nacc : {1, 2, 4, 8}) {
        if (nacc * n > 448) continue;
        for (int commit_every : {0, 8, 4}) {
          Args a{ts, n, nacc, 256, commit_every, out};
          out[0] = out[1] = 0;
          launch(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          printf("%-4s %4d %5d %6d %7d | %10.1f %10.1f\n", ts ? "tmem" : "smem", n, nacc, a.count, commit_every,
                 out[0] / (double)a.count, out[1] / (double)a.count);
        }
      }
  return 0;
}
