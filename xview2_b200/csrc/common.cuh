// Shared helpers for libxv2 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/xv2.h"

namespace xv2 {

void set_error(const char* fmt, ...);

#define XV2_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::xv2::set_error(__VA_ARGS__);    \
      return XV2_EINVAL;                \
    }                                   \
  } while (0)

#define XV2_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ::xv2::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return XV2_ECUDA;                                                                 \
    }                                                                                   \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t dtype_size(int dt) { return dt == XV2_BF16 ? 2 : 4; }

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// A training step is ~600 short dependent launches replayed from one CUDA graph; between two of them the GPU idles for the
// launch latency plus the successor's prologue (barrier init, TMEM allocation, descriptor prefetch, coefficient loads).
// Kernels launched through launch_pdl() carry cudaLaunchAttributeProgrammaticStreamSerialization: they may be scheduled as soon
// as every CTA of the predecessor has executed pdl_trigger() (first instruction of our kernels), run their prologue, and then
// block in pdl_wait() until the predecessor grid has COMPLETED and its writes are visible -- no global memory is touched before
// pdl_wait().  Under stream capture the edge becomes a programmatic graph dependency.
// MEASURED (B200, config-2 step replayed from its CUDA graph, profiles/r02_pdl_experiment.txt): 35.12 ms with the attribute on
// the conv / wgrad / BN / split-attention kernels vs 33.12 ms without -- the programmatic edges cost more than the ~2 us
// bubbles they hide -- so the attribute is OFF unless XV2_PDL=1 (griddepcontrol.* are no-ops for a plain launch).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- scalar / vector load-store with conversion to fp32 ---------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Vec<T>: 16-byte vector of T (8 bf16 or 4 fp32) for coalesced channel-contiguous access.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == XV2_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == XV2_ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  return v;
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == XV2_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  if (act == XV2_ACT_LRELU) return pre > 0.f ? 1.f : 0.01f;
  return 1.f;
}

// Dispatch a templated launcher on the activation dtype.
#define XV2_DISPATCH_DTYPE(dt, T, ...)                        \
  do {                                                        \
    if ((dt) == XV2_F32) {                                    \
      using T = float;                                        \
      __VA_ARGS__;                                            \
    } else if ((dt) == XV2_BF16) {                            \
      using T = __nv_bfloat16;                                \
      __VA_ARGS__;                                            \
    } else {                                                  \
      ::xv2::set_error("unknown dtype %d", (int)(dt));        \
      return XV2_EINVAL;                                      \
    }                                                         \
  } while (0)

}  // namespace xv2
