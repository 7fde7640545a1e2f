// Probe (GPU box): cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, swizzle width and a row-shifted A start.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/umma_rate_probe tools/umma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../xview2_b200/csrc/tc_common.cuh"
using namespace xv2::tc;

__global__ void rate(int n, int rowb, int shift_rows, int iters, int distinct, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 65536, done = base + 131072, slot = done + 8;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint64_t proto = make_smem_desc(0, 16, 8 * rowb, rowb);
    const uint32_t hi = (uint32_t)(proto >> 32), lo = (uint32_t)proto;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        // `distinct` different A tiles (8 KB apart) like the taps of a row; each with the row shift
        const uint32_t a = sa + (i % distinct) * 8192 + shift_rows * rowb;
        umma_bf16_lohi(tmem, lo + (a >> 4), hi, lo + (sb >> 4), hi, idesc, 1);
        umma_bf16_lohi(tmem, lo + (a >> 4) + 2, hi, lo + (sb >> 4) + 2, hi, idesc, 1);
      }
      umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = 0;
    if (t0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  const int iters = 2000;
  for (int rowb : {64, 128})
    for (int n : {32, 64, 96, 128, 256})
      for (int shift : {0, 1, 2}) {
        rate<<<1, 128, 140000>>>(n, rowb, shift, iters, 6, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        printf("rowb=%3d N=%3d shift=%d: %.1f cycles per MMA (M=128,K=16)%s\n", rowb, n, shift, (double)cyc / (2.0 * iters),
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  // all SMs busy at once (power / clock behaviour): 148 CTAs
  for (int n : {32, 64, 256}) {
    rate<<<148, 128, 140000>>>(n, 128, 1, iters, 6, d);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("148 CTAs rowb=128 N=%3d shift=1: %.1f cycles per MMA\n", n, (double)cyc / (2.0 * iters));
  }
  return 0;
}
