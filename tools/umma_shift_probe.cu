// Probe (GPU box): does a UMMA K-major swizzled A descriptor whose start address is shifted by whole rows (a 1-pixel
// shift of an NHWC halo tile) read the rows TMA wrote there?  D[m][n] = sum_k A[m + shift][k] * I[n][k].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/umma_shift_probe tools/umma_shift_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../xview2_b200/csrc/tc_common.cuh"
using namespace xv2::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int C>  // channels per row: 32 -> SW64, 64 -> SW128
__global__ void probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out,
                      int shift_rows, int rows_a, int base_off) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr uint32_t ROWB = C * 2;
  const uint32_t sa = base, sb = base + 32768, bar = base + 49152, done = bar + 8, slot = bar + 16;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, rows_a * ROWB + 32 * ROWB);
    tma_load_2d(sa, &map_a, bar, 0, 0);
    tma_load_2d(sb, &map_b, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    uint64_t ad = make_smem_desc(sa + shift_rows * ROWB, 16, 8 * ROWB, ROWB);
    ad |= ((uint64_t)(base_off & 7)) << 49;
    const uint64_t bd = make_smem_desc(sb, 16, 8 * ROWB, ROWB);
    for (int k = 0; k < C / 16; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, k != 0);
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

template <int C> int run(EncodeTiledFn enc) {
  const int ROWS = 144;
  std::vector<__nv_bfloat16> ha(ROWS * C), hb(32 * C);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < C; ++c) ha[r * C + c] = __float2bfloat16((float)(r + 1) + (c < 32 ? 0.f : 0.f));
  // B[n][k] = 1 if k == n (n < 32): D[m][n] = A[m+shift][n]; encode the column too: A[r][c] = r + 1 + c/64
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < C; ++c) ha[r * C + c] = __float2bfloat16((float)((r + 1) % 200) + (float)c / 64.f);
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < C; ++k) hb[n * C + k] = __float2bfloat16(k == n ? 1.f : 0.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)ROWS};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)ROWS};
  cuuint32_t es[2] = {1, 1};
  CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t dimsb[2] = {(cuuint64_t)C, 32};
  cuuint32_t boxb[2] = {(cuuint32_t)C, 32};
  CUresult r2 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    printf("encode failed %d %d\n", (int)r, (int)r2);
    return 1;
  }
  cudaFuncSetAttribute(probe<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  std::vector<float> ho(128 * 32);
  for (int shift = 0; shift <= 10; ++shift) {
    for (int bo = 0; bo < 2; ++bo) {
      const int base_off = bo ? ((shift * C * 2) >> 7) & 7 : 0;
      if (bo && base_off == 0) continue;
      probe<C><<<1, 128, 65536>>>(ma, mb, dout, shift, ROWS, base_off);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("C=%d shift=%d base_off=%d: CUDA error %s\n", C, shift, base_off, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          const float exp = __bfloat162float(ha[(m + shift) * C + n]);
          if (ho[m * 32 + n] != exp) {
            if (first < 0) first = m * 32 + n;
            ++bad;
          }
        }
      printf("C=%d shift=%2d rows base_off=%d: %s (%d mismatches%s)\n", C, shift, base_off, bad ? "MISMATCH" : "ok", bad,
             bad ? "" : "");
      if (bad && first >= 0)
        printf("    first at m=%d n=%d got %.3f want %.3f\n", first / 32, first % 32, ho[first],
               __bfloat162float(ha[(first / 32 + shift) * C + first % 32]));
    }
  }
  return 0;
}

// ---- MN-major variant (weight gradient with a halo tile): D[c][n] = sum_p X[p + shift][c] * Y[p][n] ----------------
// X: two 64-channel atoms, each a TMA box [ROWS pixels][64 ch] (128B swizzle); Y: [128 pixels][32] (64B swizzle).
__global__ void probe_mn(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, float* out,
                         int shift_rows, int rows_x, int base_off) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t atom_bytes = rows_x * 128;  // multiple of 1024 (rows_x = 144)
  const uint32_t sx = base, sy = base + 2 * atom_bytes, bar = sy + 8192, done = bar + 8, slot = bar + 16;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 2 * atom_bytes + 128 * 64);
    tma_load_2d(sx, &map_x, bar, 0, 0);
    tma_load_2d(sx + atom_bytes, &map_x, bar, 64, 0);
    tma_load_2d(sy, &map_y, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 32, 1, 1);
    uint64_t ad = make_smem_desc(sx + shift_rows * 128, atom_bytes, 8 * 128, 128);
    ad |= ((uint64_t)(base_off & 7)) << 49;
    const uint64_t bd = make_smem_desc(sy, 8192, 8 * 64, 64);
    for (int k = 0; k < 8; ++k)
      umma_bf16(tmem, ad + (uint64_t)((k * 16 * 128) >> 4), bd + (uint64_t)((k * 16 * 64) >> 4), idesc, k != 0);
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int run_mn(EncodeTiledFn enc) {
  const int ROWS = 144, C = 128;
  std::vector<__nv_bfloat16> hx(ROWS * C), hy(128 * 32);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < C; ++c) hx[r * C + c] = __float2bfloat16((float)((r * 7 + c * 3) % 61) - 30.f);
  for (int p = 0; p < 128; ++p)
    for (int n = 0; n < 32; ++n) hy[p * 32 + n] = __float2bfloat16((p % 32) == n ? (float)(1 + p / 32) : 0.f);
  __nv_bfloat16 *dx, *dy;
  float* dout;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dy, hy.size() * 2);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, hy.data(), hy.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mx, my;
  cuuint32_t es[2] = {1, 1};
  cuuint64_t dx_dims[2] = {(cuuint64_t)C, (cuuint64_t)ROWS};
  cuuint64_t dx_str[1] = {(cuuint64_t)C * 2};
  cuuint32_t dx_box[2] = {64, (cuuint32_t)ROWS};
  CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dx_dims, dx_str, dx_box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t dy_dims[2] = {32, 128};
  cuuint64_t dy_str[1] = {64};
  cuuint32_t dy_box[2] = {32, 128};
  CUresult r2 = enc(&my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, dy_dims, dy_str, dy_box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    printf("mn encode failed %d %d\n", (int)r, (int)r2);
    return 1;
  }
  cudaFuncSetAttribute(probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  std::vector<float> ho(128 * 32);
  for (int shift = 0; shift <= 10; ++shift) {
    for (int bo = 0; bo < 2; ++bo) {
      const int base_off = bo ? shift & 7 : 0;
      if (bo && base_off == 0) continue;
      probe_mn<<<1, 128, 65536>>>(mx, my, dout, shift, ROWS, base_off);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("MN shift=%d base_off=%d: CUDA error %s\n", shift, base_off, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int c = 0; c < 128; ++c)
        for (int n = 0; n < 32; ++n) {
          float exp = 0.f;
          for (int j = 0; j < 4; ++j) exp += (float)(1 + j) * __bfloat162float(hx[(n + 32 * j + shift) * C + c]);
          if (ho[c * 32 + n] != exp) ++bad;
        }
      printf("MN-major shift=%2d rows base_off=%d: %s (%d mismatches)\n", shift, base_off, bad ? "MISMATCH" : "ok", bad);
    }
  }
  return 0;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  int rc = run<32>(enc);
  rc |= run<64>(enc);
  rc |= run_mn(enc);
  return rc;
}
