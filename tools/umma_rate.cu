// Microbenchmark (diagnostics): cycles per tcgen05.mma (M=128, K=16 bf16 / K=32 s8) as a function of N and of the A source
// (shared memory vs TMEM), issued back to back by one thread per SM on all SMs.  Operand values are irrelevant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/umma_rate tools/umma_rate.cu && gpurun_out/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../morphik-core_b200/csrc/ptx.cuh"
using namespace bms;

template <int KIND, int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc_512(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(KIND, 128, N);
    const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(smem) + 32768);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // one "tile": 8 K-steps into one accumulator, alternate 2 accumulators
            const uint32_t d = tmem + 256 + (i & 1) * (N == 256 ? 0 : N);
            if (TS) umma_ts<KIND>(d, tmem + (k & 3) * 8, bd + (k & 3) * 2, idesc, k != 0);
            else umma_ss<KIND>(d, ad + (k & 3) * 2, bd + (k & 3) * 2, idesc, k != 0);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_512(tmem); }
}

template <int KIND, int N, bool TS>
void run(const char* name, long long* d_out) {
  const int iters = 4096, smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<KIND, N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<KIND, N, TS><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148.0;
  const double per_mma = avg / (iters * 8.0);
  const double macs = 128.0 * N * (KIND == 0 ? 16 : 32);
  printf("%-28s %s  cycles/MMA %.1f  MAC/clk/SM %.0f (nominal %d)\n", name, cudaGetErrorString(e), per_mma, macs / per_mma, KIND == 0 ? 4096 : 8192);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  run<0, 64, false>("bf16 SS N=64", d_out);
  run<0, 128, false>("bf16 SS N=128", d_out);
  run<0, 256, false>("bf16 SS N=256", d_out);
  run<0, 64, true>("bf16 TS N=64", d_out);
  run<0, 128, true>("bf16 TS N=128", d_out);
  run<0, 256, true>("bf16 TS N=256", d_out);
  run<1, 128, false>("s8 SS N=128", d_out);
  run<1, 256, false>("s8 SS N=256", d_out);
  return 0;
}
