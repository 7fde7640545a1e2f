// NHWC max / average pooling, forward and backward (gather form: every output element is written exactly once, no
// atomics).  Replaces ATen/cuDNN pooling reached from unet.py:81 (MaxPool2d 3,2,1) and the ResNeSt avd / avg-down
// pools (unet.py:52).
#include "common.cuh"

namespace xv2 {

template <typename T, int VEC> __device__ __forceinline__ void pldv(const T* p, float* f) {
  if constexpr (VEC == 1) {
    f[0] = to_f(*p);
  } else {
    Vec<T> v;
    v.load(p);
    v.unpack(f);
  }
}
template <typename T, int VEC> __device__ __forceinline__ void pstv(T* p, const float* f) {
  if constexpr (VEC == 1) {
    *p = from_f<T>(f[0]);
  } else {
    Vec<T> v;
    v.pack(f);
    v.store(p);
  }
}

struct PoolGeom {
  int n, h, w, c, oh, ow, k, stride, pad, cip;
};

template <typename T, int VEC>
__global__ void maxpool_fwd_kernel(PoolGeom g, const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx) {
  const int cv = g.c / VEC;
  const long long total = (long long)g.n * g.oh * g.ow * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cvi = (int)(i % cv);
    long long t = i / cv;
    const int ow = (int)(t % g.ow);
    t /= g.ow;
    const int oh = (int)(t % g.oh);
    const int nb = (int)(t / g.oh);
    float m[VEC];
    uint8_t am[VEC];  // window position (r*k + s) of the FIRST maximum in scan order, like ATen
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      m[j] = -INFINITY;
      am[j] = 0;
    }
    for (int r = 0; r < g.k; ++r) {
      const int ih = oh * g.stride - g.pad + r;
      if (ih < 0 || ih >= g.h) continue;
      for (int s = 0; s < g.k; ++s) {
        const int iw = ow * g.stride - g.pad + s;
        if (iw < 0 || iw >= g.w) continue;
        float f[VEC];
        pldv<T, VEC>(x + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (f[j] > m[j]) {
            m[j] = f[j];
            am[j] = (uint8_t)(r * g.k + s);
          }
      }
    }
    const long long o = (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC;
    pstv<T, VEC>(y + o, m);
    if (idx) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) idx[o + j] = am[j];
    }
  }
}

// dx[n,ih,iw,c] = sum over the windows (oh,ow) containing (ih,iw) whose saved arg-max position is (ih,iw) of dy.
// Gather form: every dx element is written exactly once, no atomics; reads dy + the uint8 index map only.
template <typename T, int VEC>
__global__ void maxpool_bwd_kernel(PoolGeom g, const uint8_t* __restrict__ idx, const T* __restrict__ dy,
                                   T* __restrict__ dx) {
  const int cv = g.c / VEC;
  const long long total = (long long)g.n * g.h * g.w * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cvi = (int)(i % cv);
    long long t = i / cv;
    const int iw = (int)(t % g.w);
    t /= g.w;
    const int ih = (int)(t % g.h);
    const int nb = (int)(t / g.h);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    int oh_lo = (ih + g.pad - g.k + 1 + g.stride - 1);
    oh_lo = oh_lo < 0 ? 0 : oh_lo / g.stride;
    int oh_hi = (ih + g.pad) / g.stride;
    if (oh_hi > g.oh - 1) oh_hi = g.oh - 1;
    int ow_lo = (iw + g.pad - g.k + 1 + g.stride - 1);
    ow_lo = ow_lo < 0 ? 0 : ow_lo / g.stride;
    int ow_hi = (iw + g.pad) / g.stride;
    if (ow_hi > g.ow - 1) ow_hi = g.ow - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int mine = (ih - (oh * g.stride - g.pad)) * g.k + (iw - (ow * g.stride - g.pad));
        const long long o = (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC;
        float d[VEC];
        pldv<T, VEC>(dy + o, d);
        if constexpr (VEC == 8) {
          const uint2 pk = *reinterpret_cast<const uint2*>(idx + o);
          const uint32_t w2[2] = {pk.x, pk.y};
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += (int)((w2[j >> 2] >> (8 * (j & 3))) & 0xFF) == mine ? d[j] : 0.f;
        } else if constexpr (VEC == 4) {
          const uint32_t pk = *reinterpret_cast<const uint32_t*>(idx + o);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] += (int)((pk >> (8 * j)) & 0xFF) == mine ? d[j] : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] += (int)idx[o + j] == mine ? d[j] : 0.f;
        }
      }
    }
    pstv<T, VEC>(dx + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, acc);
  }
}

__device__ __forceinline__ float avg_divisor(const PoolGeom& g, int oh, int ow) {
  int hs = oh * g.stride - g.pad, ws = ow * g.stride - g.pad;
  int he = min(hs + g.k, g.h + g.pad), we = min(ws + g.k, g.w + g.pad);
  const int pool = (he - hs) * (we - ws);
  hs = max(hs, 0);
  ws = max(ws, 0);
  he = min(he, g.h);
  we = min(we, g.w);
  return (float)(g.cip ? pool : (he - hs) * (we - ws));
}

template <typename T, int VEC>
__global__ void avgpool_fwd_kernel(PoolGeom g, const T* __restrict__ x, T* __restrict__ y) {
  const int cv = g.c / VEC;
  const long long total = (long long)g.n * g.oh * g.ow * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cvi = (int)(i % cv);
    long long t = i / cv;
    const int ow = (int)(t % g.ow);
    t /= g.ow;
    const int oh = (int)(t % g.oh);
    const int nb = (int)(t / g.oh);
    float a[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) a[j] = 0.f;
    for (int r = 0; r < g.k; ++r) {
      const int ih = oh * g.stride - g.pad + r;
      if (ih < 0 || ih >= g.h) continue;
      for (int s = 0; s < g.k; ++s) {
        const int iw = ow * g.stride - g.pad + s;
        if (iw < 0 || iw >= g.w) continue;
        float f[VEC];
        pldv<T, VEC>(x + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[j] += f[j];
      }
    }
    const float inv = 1.0f / avg_divisor(g, oh, ow);
#pragma unroll
    for (int j = 0; j < VEC; ++j) a[j] *= inv;
    pstv<T, VEC>(y + (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC, a);
  }
}

template <typename T, int VEC>
__global__ void avgpool_bwd_kernel(PoolGeom g, const T* __restrict__ dy, T* __restrict__ dx) {
  const int cv = g.c / VEC;
  const long long total = (long long)g.n * g.h * g.w * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cvi = (int)(i % cv);
    long long t = i / cv;
    const int iw = (int)(t % g.w);
    t /= g.w;
    const int ih = (int)(t % g.h);
    const int nb = (int)(t / g.h);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    int oh_lo = (ih + g.pad - g.k + 1 + g.stride - 1);
    oh_lo = oh_lo < 0 ? 0 : oh_lo / g.stride;
    int oh_hi = (ih + g.pad) / g.stride;
    if (oh_hi > g.oh - 1) oh_hi = g.oh - 1;
    int ow_lo = (iw + g.pad - g.k + 1 + g.stride - 1);
    ow_lo = ow_lo < 0 ? 0 : ow_lo / g.stride;
    int ow_hi = (iw + g.pad) / g.stride;
    if (ow_hi > g.ow - 1) ow_hi = g.ow - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        float d[VEC];
        pldv<T, VEC>(dy + (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC, d);
        const float inv = 1.0f / avg_divisor(g, oh, ow);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = fmaf(d[j], inv, acc[j]);
      }
    pstv<T, VEC>(dx + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, acc);
  }
}

static int pool_blocks(long long total) {
  long long b = cdiv(total, 256);
  if (b > 16 * kNumSMs) b = 16 * kNumSMs;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace xv2

using namespace xv2;

#define XV2_POOL_LAUNCH(KERNEL, total_pixels, ...)                                              \
  do {                                                                                          \
    const int vecw = (dtype == XV2_BF16 ? 8 : 4);                                               \
    const bool use_vec = (c % vecw) == 0;                                                       \
    const long long total = (long long)(total_pixels) * (use_vec ? c / vecw : c);               \
    const int blocks = pool_blocks(total);                                                      \
    XV2_DISPATCH_DTYPE(dtype, T, {                                                              \
      if (use_vec) KERNEL<T, Vec<T>::N><<<blocks, 256, 0, as_stream(stream)>>>(__VA_ARGS__);    \
      else KERNEL<T, 1><<<blocks, 256, 0, as_stream(stream)>>>(__VA_ARGS__);                    \
    });                                                                                         \
    XV2_LAUNCH_CHECK();                                                                         \
  } while (0)

extern "C" int xv2_maxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t n, int32_t h, int32_t w, int32_t c,
                               int32_t oh, int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t dtype,
                               void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && k <= 15 && stride > 0, "maxpool: bad shape");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, 0};
  XV2_POOL_LAUNCH(maxpool_fwd_kernel, (long long)n * oh * ow, g, (const T*)x, (T*)y, idx);
  return XV2_OK;
}
extern "C" int xv2_maxpool_bwd(const uint8_t* idx, const void* dy, void* dx, int32_t n, int32_t h, int32_t w,
                               int32_t c, int32_t oh, int32_t ow, int32_t k, int32_t stride, int32_t pad,
                               int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && stride > 0, "maxpool: bad shape");
  XV2_REQUIRE(idx != nullptr, "maxpool_bwd: index map missing");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, 0};
  XV2_POOL_LAUNCH(maxpool_bwd_kernel, (long long)n * h * w, g, idx, (const T*)dy, (T*)dx);
  return XV2_OK;
}
extern "C" int xv2_avgpool_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                               int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t count_include_pad,
                               int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && stride > 0, "avgpool: bad shape");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, count_include_pad};
  XV2_POOL_LAUNCH(avgpool_fwd_kernel, (long long)n * oh * ow, g, (const T*)x, (T*)y);
  return XV2_OK;
}
extern "C" int xv2_avgpool_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                               int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t count_include_pad,
                               int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && stride > 0, "avgpool: bad shape");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, count_include_pad};
  XV2_POOL_LAUNCH(avgpool_bwd_kernel, (long long)n * h * w, g, (const T*)dy, (T*)dx);
  return XV2_OK;
}
