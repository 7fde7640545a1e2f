// NHWC max / average pooling, forward and backward (gather form: every output element is written exactly once, no
// atomics).  Replaces ATen/cuDNN pooling reached from unet.py:81 (MaxPool2d 3,2,1) and the ResNeSt avd / avg-down
// pools (unet.py:52).
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

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

template <typename T, int VEC, int KK, int SS>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(PoolGeom g, const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx) {
  const int cv = g.c / VEC;
  const int K = KK ? KK : g.k, S = SS ? SS : g.stride;  // compile-time for the shapes the U-Nets use
  const int rowsz = g.ow * cv;  // 32-bit index math: blockIdx.y = (image, row), x runs over (column, channel vector)
  const int nb = blockIdx.y / g.oh, oh = blockIdx.y - nb * g.oh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowsz; i += gridDim.x * blockDim.x) {
    const int ow = i / cv, cvi = i - ow * cv;
    float m[VEC];
    uint8_t am[VEC];  // window position (r*k + s) of the FIRST maximum in scan order, like ATen
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      m[j] = -INFINITY;
      am[j] = 0;
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int ih = oh * S - g.pad + r;
      if (ih < 0 || ih >= g.h) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int iw = ow * S - g.pad + s;
        if (iw < 0 || iw >= g.w) continue;
        float f[VEC];
        pldv<T, VEC>(x + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (f[j] > m[j]) {
            m[j] = f[j];
            am[j] = (uint8_t)(r * K + s);
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
template <typename T, int VEC, int KK, int SS>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(PoolGeom g, const uint8_t* __restrict__ idx, const T* __restrict__ dy,
                                   T* __restrict__ dx) {
  const int cv = g.c / VEC;
  const int K = KK ? KK : g.k, S = SS ? SS : g.stride;  // compile-time for the shapes the U-Nets use
  const int rowsz = g.w * cv;  // 32-bit index math: blockIdx.y = (image, row), x runs over (column, channel vector)
  const int nb = blockIdx.y / g.h, ih = blockIdx.y - nb * g.h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowsz; i += gridDim.x * blockDim.x) {
    const int iw = i / cv, cvi = i - iw * cv;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    int oh_lo = (ih + g.pad - K + 1 + S - 1);
    oh_lo = oh_lo < 0 ? 0 : oh_lo / S;
    int oh_hi = (ih + g.pad) / S;
    if (oh_hi > g.oh - 1) oh_hi = g.oh - 1;
    int ow_lo = (iw + g.pad - K + 1 + S - 1);
    ow_lo = ow_lo < 0 ? 0 : ow_lo / S;
    int ow_hi = (iw + g.pad) / S;
    if (ow_hi > g.ow - 1) ow_hi = g.ow - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int mine = (ih - (oh * S - g.pad)) * K + (iw - (ow * S - g.pad));
        const long long o = (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC;
        float d[VEC];
        pldv<T, VEC>(dy + o, d);
        if constexpr (VEC == 8) {
          const uint2 pk = *reinterpret_cast<const uint2*>(idx + o);
          const uint32_t pat = (uint32_t)mine * 0x01010101u;
          const uint32_t m2[2] = {__vcmpeq4(pk.x, pat), __vcmpeq4(pk.y, pat)};  // 0xFF per matching byte
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += (m2[j >> 2] >> (8 * (j & 3)) & 1u) ? d[j] : 0.f;
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

__device__ __forceinline__ float avg_divisor(const PoolGeom& g, int oh, int ow, int K, int S) {
  int hs = oh * S - g.pad, ws = ow * S - g.pad;
  int he = min(hs + K, g.h + g.pad), we = min(ws + K, g.w + g.pad);
  const int pool = (he - hs) * (we - ws);
  hs = max(hs, 0);
  ws = max(ws, 0);
  he = min(he, g.h);
  we = min(we, g.w);
  return (float)(g.cip ? pool : (he - hs) * (we - ws));
}

template <typename T, int VEC, int KK, int SS>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(PoolGeom g, const T* __restrict__ x, T* __restrict__ y) {
  const int cv = g.c / VEC;
  const int K = KK ? KK : g.k, S = SS ? SS : g.stride;  // compile-time for the shapes the U-Nets use
  const int rowsz = g.ow * cv;  // 32-bit index math: blockIdx.y = (image, row), x runs over (column, channel vector)
  const int nb = blockIdx.y / g.oh, oh = blockIdx.y - nb * g.oh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowsz; i += gridDim.x * blockDim.x) {
    const int ow = i / cv, cvi = i - ow * cv;
    float a[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) a[j] = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int ih = oh * S - g.pad + r;
      if (ih < 0 || ih >= g.h) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int iw = ow * S - g.pad + s;
        if (iw < 0 || iw >= g.w) continue;
        float f[VEC];
        pldv<T, VEC>(x + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[j] += f[j];
      }
    }
    const float inv = 1.0f / avg_divisor(g, oh, ow, K, S);
#pragma unroll
    for (int j = 0; j < VEC; ++j) a[j] *= inv;
    pstv<T, VEC>(y + (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC, a);
  }
}

template <typename T, int VEC, int KK, int SS>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(PoolGeom g, const T* __restrict__ dy, T* __restrict__ dx) {
  const int cv = g.c / VEC;
  const int K = KK ? KK : g.k, S = SS ? SS : g.stride;  // compile-time for the shapes the U-Nets use
  const int rowsz = g.w * cv;  // 32-bit index math: blockIdx.y = (image, row), x runs over (column, channel vector)
  const int nb = blockIdx.y / g.h, ih = blockIdx.y - nb * g.h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowsz; i += gridDim.x * blockDim.x) {
    const int iw = i / cv, cvi = i - iw * cv;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    int oh_lo = (ih + g.pad - K + 1 + S - 1);
    oh_lo = oh_lo < 0 ? 0 : oh_lo / S;
    int oh_hi = (ih + g.pad) / S;
    if (oh_hi > g.oh - 1) oh_hi = g.oh - 1;
    int ow_lo = (iw + g.pad - K + 1 + S - 1);
    ow_lo = ow_lo < 0 ? 0 : ow_lo / S;
    int ow_hi = (iw + g.pad) / S;
    if (ow_hi > g.ow - 1) ow_hi = g.ow - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        float d[VEC];
        pldv<T, VEC>(dy + (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC, d);
        const float inv = 1.0f / avg_divisor(g, oh, ow, K, S);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = fmaf(d[j], inv, acc[j]);
      }
    pstv<T, VEC>(dx + (((long long)nb * g.h + ih) * g.w + iw) * g.c + cvi * VEC, acc);
  }
}

// 3x3 / stride 2 / pad 1 backward on even-sized images, one thread per 2x2 input quad and channel vector: the quad
// (2a..2a+1, 2b..2b+1) is reached by the windows (a..a+1, b..b+1) only, so 4 dy loads serve 4 dx stores (the per-pixel gather
// above re-reads every dy element 2.25 times from 4x as many threads).  Same accumulation order as the generic kernels.
template <typename T, int VEC, bool MAX>
__global__ void __launch_bounds__(256) pool3s2_bwd_quad_kernel(PoolGeom g, const uint8_t* __restrict__ idx, const T* __restrict__ dy,
                                                               T* __restrict__ dx) {
  const int cv = g.c / VEC;
  const int rowsz = g.ow * cv;
  const int nb = blockIdx.y / g.oh, a = blockIdx.y - nb * g.oh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowsz; i += gridDim.x * blockDim.x) {
    const int b = i / cv, cvi = i - b * cv;
    float acc[4][VEC];  // quad pixel q = 2 * dr + dc
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int wr = 0; wr < 2; ++wr) {
#pragma unroll
      for (int wc = 0; wc < 2; ++wc) {
        const int oh = a + wr, ow = b + wc;
        if (oh >= g.oh || ow >= g.ow) continue;
        const long long o = (((long long)nb * g.oh + oh) * g.ow + ow) * g.c + cvi * VEC;
        float d[VEC];
        pldv<T, VEC>(dy + o, d);
        uint32_t pk[2] = {0u, 0u};
        float inv = 0.f;
        if constexpr (MAX) {
          if constexpr (VEC == 8) {
            const uint2 v = *reinterpret_cast<const uint2*>(idx + o);
            pk[0] = v.x;
            pk[1] = v.y;
          } else {
            pk[0] = *reinterpret_cast<const uint32_t*>(idx + o);
          }
        } else {
          inv = 1.0f / avg_divisor(g, oh, ow, 3, 2);
        }
        // window (oh, ow) covers input rows 2 oh - 1 .. 2 oh + 1: quad row dr (input row 2 a + dr) sits at window row 2 (a - oh) + dr + 1
#pragma unroll
        for (int dr = 0; dr < 2; ++dr) {
          const int r = dr + 1 - 2 * wr;
          if (r < 0) continue;
#pragma unroll
          for (int dc = 0; dc < 2; ++dc) {
            const int sx = dc + 1 - 2 * wc;
            if (sx < 0) continue;
            const int q = 2 * dr + dc;
            if constexpr (MAX) {
              const uint32_t mine = (uint32_t)(r * 3 + sx);
#pragma unroll
              for (int j = 0; j < VEC; ++j) acc[q][j] += ((pk[j >> 2] >> (8 * (j & 3))) & 0xFFu) == mine ? d[j] : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < VEC; ++j) acc[q][j] = fmaf(d[j], inv, acc[q][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
      for (int dc = 0; dc < 2; ++dc)
        pstv<T, VEC>(dx + (((long long)nb * g.h + 2 * a + dr) * g.w + 2 * b + dc) * g.c + cvi * VEC, acc[2 * dr + dc]);
  }
}

// 2x2 / stride 2 / pad 0 average pool backward on even images (the ResNeSt avg-down shortcut): every input pixel belongs to
// exactly one window -- one dy load, four dx stores per thread (same value as the generic gather: round(dy * 1/divisor))
template <typename T, int VEC>
__global__ void __launch_bounds__(256) avgpool2s2_bwd_quad_kernel(PoolGeom g, const T* __restrict__ dy, T* __restrict__ dx) {
  const int cv = g.c / VEC;
  const int rowsz = g.ow * cv;
  const int nb = blockIdx.y / g.oh, a = blockIdx.y - nb * g.oh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowsz; i += gridDim.x * blockDim.x) {
    const int b = i / cv, cvi = i - b * cv;
    float d[VEC];
    pldv<T, VEC>(dy + (((long long)nb * g.oh + a) * g.ow + b) * g.c + cvi * VEC, d);
    const float inv = 1.0f / avg_divisor(g, a, b, 2, 2);
#pragma unroll
    for (int j = 0; j < VEC; ++j) d[j] = fmaf(d[j], inv, 0.f);
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
      for (int dc = 0; dc < 2; ++dc) pstv<T, VEC>(dx + (((long long)nb * g.h + 2 * a + dr) * g.w + 2 * b + dc) * g.c + cvi * VEC, d);
  }
}

static int pool_blocks(long long total) {
  long long b = cdiv(total, 256);
  if (b > 16 * kNumSMs) b = 16 * kNumSMs;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace xv2

using namespace xv2;

#define XV2_POOL_LAUNCH(KERNEL, images, rows, cols, ...)                                        \
  do {                                                                                          \
    const int vecw = (dtype == XV2_BF16 ? 8 : 4);                                               \
    const bool use_vec = (c % vecw) == 0;                                                       \
    const long long rowsz = (long long)(cols) * (use_vec ? c / vecw : c);                       \
    XV2_REQUIRE((long long)(images) * (rows) <= 65535 && rowsz < (1LL << 30), "pool: tensor too large for the 2-D grid"); \
    const dim3 grid((unsigned)std::min<long long>(cdiv(rowsz, 256), 64), (unsigned)((images) * (rows)));            \
    XV2_DISPATCH_DTYPE(dtype, T, {                                                              \
      if (use_vec && k == 3 && stride == 2) KERNEL<T, Vec<T>::N, 3, 2><<<grid, 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
      else if (use_vec && k == 2 && stride == 2) KERNEL<T, Vec<T>::N, 2, 2><<<grid, 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
      else if (use_vec) KERNEL<T, Vec<T>::N, 0, 0><<<grid, 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
      else KERNEL<T, 1, 0, 0><<<grid, 256, 0, as_stream(stream)>>>(__VA_ARGS__);                \
    });                                                                                         \
    XV2_LAUNCH_CHECK();                                                                         \
  } while (0)

// quad kernel eligibility: 3x3 / 2 / 1 on even images with vectorised channels
static bool quad_ok(int h, int w, int c, int oh, int ow, int k, int stride, int pad, int dtype) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("XV2_NO_POOL_QUAD");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  const int vecw = dtype == XV2_BF16 ? 8 : 4;
  return enabled && k == 3 && stride == 2 && pad == 1 && h == 2 * oh && w == 2 * ow && c % vecw == 0;
}
template <bool MAX>
static int launch_quad(const PoolGeom& g, const uint8_t* idx, const void* dy, void* dx, int dtype, void* stream) {
  const int vecw = dtype == XV2_BF16 ? 8 : 4;
  const long long rowsz = (long long)g.ow * (g.c / vecw);
  XV2_REQUIRE((long long)g.n * g.oh <= 65535, "pool: tensor too large for the 2-D grid");
  const dim3 grid((unsigned)std::min<long long>(cdiv(rowsz, 256), 64), (unsigned)(g.n * g.oh));
  if (dtype == XV2_BF16)
    pool3s2_bwd_quad_kernel<__nv_bfloat16, 8, MAX><<<grid, 256, 0, as_stream(stream)>>>(g, idx, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx);
  else
    pool3s2_bwd_quad_kernel<float, 4, MAX><<<grid, 256, 0, as_stream(stream)>>>(g, idx, (const float*)dy, (float*)dx);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_maxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t n, int32_t h, int32_t w, int32_t c,
                               int32_t oh, int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t dtype,
                               void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && k <= 15 && stride > 0, "maxpool: bad shape");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, 0};
  XV2_POOL_LAUNCH(maxpool_fwd_kernel, n, oh, ow, g, (const T*)x, (T*)y, idx);
  return XV2_OK;
}
extern "C" int xv2_maxpool_bwd(const uint8_t* idx, const void* dy, void* dx, int32_t n, int32_t h, int32_t w,
                               int32_t c, int32_t oh, int32_t ow, int32_t k, int32_t stride, int32_t pad,
                               int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && stride > 0, "maxpool: bad shape");
  XV2_REQUIRE(idx != nullptr, "maxpool_bwd: index map missing");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, 0};
  if (quad_ok(h, w, c, oh, ow, k, stride, pad, dtype)) return launch_quad<true>(g, idx, dy, dx, dtype, stream);
  XV2_POOL_LAUNCH(maxpool_bwd_kernel, n, h, w, g, idx, (const T*)dy, (T*)dx);
  return XV2_OK;
}
extern "C" int xv2_avgpool_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                               int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t count_include_pad,
                               int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && stride > 0, "avgpool: bad shape");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, count_include_pad};
  XV2_POOL_LAUNCH(avgpool_fwd_kernel, n, oh, ow, g, (const T*)x, (T*)y);
  return XV2_OK;
}
extern "C" int xv2_avgpool_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                               int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t count_include_pad,
                               int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && k > 0 && stride > 0, "avgpool: bad shape");
  PoolGeom g{n, h, w, c, oh, ow, k, stride, pad, count_include_pad};
  if (quad_ok(h, w, c, oh, ow, k, stride, pad, dtype)) return launch_quad<false>(g, nullptr, dy, dx, dtype, stream);
  if (quad_ok(h, w, c, oh, ow, 3, stride, 1, dtype) && k == 2 && pad == 0 && (long long)n * oh <= 65535) {  // same gate, 2x2 / 2 / 0 window
    const int vecw = dtype == XV2_BF16 ? 8 : 4;
    const dim3 grid((unsigned)std::min<long long>(cdiv((long long)ow * (c / vecw), 256), 64), (unsigned)(n * oh));
    if (dtype == XV2_BF16)
      avgpool2s2_bwd_quad_kernel<__nv_bfloat16, 8><<<grid, 256, 0, as_stream(stream)>>>(g, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx);
    else
      avgpool2s2_bwd_quad_kernel<float, 4><<<grid, 256, 0, as_stream(stream)>>>(g, (const float*)dy, (float*)dx);
    XV2_LAUNCH_CHECK();
    return XV2_OK;
  }
  XV2_POOL_LAUNCH(avgpool_bwd_kernel, n, h, w, g, (const T*)dy, (T*)dx);
  return XV2_OK;
}
