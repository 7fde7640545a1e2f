// Split-attention (ResNeSt SplAtConv2d, radix 2, cardinality 1 -- call site unet.py:52) memory-bound pieces:
// radix-sum + global average pool, r-softmax, attention-weighted combine, and their backward passes.
// x is [n][hw][2C] with the two radix splits in channel halves [0,C) and [C,2C).
#include "common.cuh"

namespace xv2 {

template <typename T, int VEC> __device__ __forceinline__ void sldv(const T* p, float* f) {
  if constexpr (VEC == 1) {
    f[0] = to_f(*p);
  } else {
    Vec<T> v;
    v.load(p);
    v.unpack(f);
  }
}
template <typename T, int VEC> __device__ __forceinline__ void sstv(T* p, const float* f) {
  if constexpr (VEC == 1) {
    *p = from_f<T>(f[0]);
  } else {
    Vec<T> v;
    v.pack(f);
    v.store(p);
  }
}

// grid: (chunks of hw, n).  block 256 = lanes x cv (cv = C/VEC channel vectors, looped when cv > 256).
// MODE 0: gap[n][c] += sum_hw (x0 + x1) * inv_hw          (out: fp32 [n][C], atomics)
// MODE 1: datt[n][r*C + c] += sum_hw dout[c] * x[r*C + c] (out: fp32 [n][2C], atomics)
template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(256) splat_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                                           float* __restrict__ out, long long hw, int c,
                                                           float inv_hw) {
  __shared__ float sm[2][256 * VEC];
  const int cv = c / VEC;
  const int cvb = cv < 256 ? cv : 256;
  const int lanes = 256 / cvb;
  const int passes = (cv + cvb - 1) / cvb;
  const int tid = threadIdx.x, cvi0 = tid % cvb, lane = tid / cvb;
  const int nb = blockIdx.y;
  const T* xb = x + (long long)nb * hw * 2 * c;
  const T* db = MODE == 1 ? dout + (long long)nb * hw * c : nullptr;
  for (int pass = 0; pass < passes; ++pass) {
    const int cvi = pass * cvb + cvi0;
    float a0[VEC], a1[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) a0[i] = a1[i] = 0.f;
    if (lane < lanes && cvi < cv) {
      for (long long p = (long long)blockIdx.x * lanes + lane; p < hw; p += (long long)gridDim.x * lanes) {
        float f0[VEC], f1[VEC];
        sldv<T, VEC>(xb + p * 2 * c + cvi * VEC, f0);
        sldv<T, VEC>(xb + p * 2 * c + c + cvi * VEC, f1);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) a0[i] += f0[i] + f1[i];
        } else {
          float d[VEC];
          sldv<T, VEC>(db + p * c + cvi * VEC, d);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            a0[i] = fmaf(d[i], f0[i], a0[i]);
            a1[i] = fmaf(d[i], f1[i], a1[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sm[0][tid * VEC + i] = a0[i];
      sm[1][tid * VEC + i] = a1[i];
    }
    __syncthreads();
    if (lane == 0 && cvi < cv) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float t0 = 0.f, t1 = 0.f;
        for (int l = 0; l < lanes; ++l) {
          t0 += sm[0][(l * cvb + cvi0) * VEC + i];
          t1 += sm[1][(l * cvb + cvi0) * VEC + i];
        }
        if (MODE == 0) {
          atomicAdd(&out[(long long)nb * c + cvi * VEC + i], t0 * inv_hw);
        } else {
          atomicAdd(&out[(long long)nb * 2 * c + cvi * VEC + i], t0);
          atomicAdd(&out[(long long)nb * 2 * c + c + cvi * VEC + i], t1);
        }
      }
    }
    __syncthreads();
  }
}

__global__ void rsoftmax_fwd_kernel(const float* __restrict__ logits, float* __restrict__ att, int n, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int nb = i / c, ch = i - nb * c;
  const float a = logits[(long long)nb * 2 * c + ch], b = logits[(long long)nb * 2 * c + c + ch];
  const float m = fmaxf(a, b);
  const float ea = expf(a - m), eb = expf(b - m);
  const float inv = 1.f / (ea + eb);
  att[(long long)nb * 2 * c + ch] = ea * inv;
  att[(long long)nb * 2 * c + c + ch] = eb * inv;
}

__global__ void rsoftmax_bwd_kernel(const float* __restrict__ att, const float* __restrict__ datt,
                                    float* __restrict__ dlogits, int n, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int nb = i / c, ch = i - nb * c;
  const long long i0 = (long long)nb * 2 * c + ch, i1 = i0 + c;
  const float p0 = att[i0], p1 = att[i1], g0 = datt[i0], g1 = datt[i1];
  const float dot = p0 * g0 + p1 * g1;
  dlogits[i0] = p0 * (g0 - dot);
  dlogits[i1] = p1 * (g1 - dot);
}

// out[n][p][c] = att0*x0 + att1*x1
template <typename T, int VEC>
__global__ void __launch_bounds__(256) splat_combine_kernel(const T* __restrict__ x, const float* __restrict__ att,
                                                            T* __restrict__ out, long long hw, int c) {
  const int cv = c / VEC;
  const int nb = blockIdx.y;
  const long long total = hw * cv;
  const T* xb = x + (long long)nb * hw * 2 * c;
  T* ob = out + (long long)nb * hw * c;
  const float* ab = att + (long long)nb * 2 * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cvi = (int)(i % cv);
    const long long p = i / cv;
    float f0[VEC], f1[VEC], o[VEC];
    sldv<T, VEC>(xb + p * 2 * c + cvi * VEC, f0);
    sldv<T, VEC>(xb + p * 2 * c + c + cvi * VEC, f1);
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[j] = ab[cvi * VEC + j] * f0[j] + ab[c + cvi * VEC + j] * f1[j];
    sstv<T, VEC>(ob + p * c + cvi * VEC, o);
  }
}

// dx[n][p][r*C+c] = att[r*C+c]*dout[c] + dgap[c]*inv_hw
template <typename T, int VEC>
__global__ void __launch_bounds__(256) splat_bwd_x_kernel(const T* __restrict__ dout, const float* __restrict__ att,
                                                          const float* __restrict__ dgap, T* __restrict__ dx,
                                                          long long hw, int c, float inv_hw) {
  const int cv = c / VEC;
  const int nb = blockIdx.y;
  const long long total = hw * cv;
  const T* db = dout + (long long)nb * hw * c;
  T* xb = dx + (long long)nb * hw * 2 * c;
  const float* ab = att + (long long)nb * 2 * c;
  const float* gb = dgap + (long long)nb * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cvi = (int)(i % cv);
    const long long p = i / cv;
    float d[VEC], o0[VEC], o1[VEC];
    sldv<T, VEC>(db + p * c + cvi * VEC, d);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float gg = gb[cvi * VEC + j] * inv_hw;
      o0[j] = fmaf(ab[cvi * VEC + j], d[j], gg);
      o1[j] = fmaf(ab[c + cvi * VEC + j], d[j], gg);
    }
    sstv<T, VEC>(xb + p * 2 * c + cvi * VEC, o0);
    sstv<T, VEC>(xb + p * 2 * c + c + cvi * VEC, o1);
  }
}

static int splat_vec(int c, int dtype) {
  int v = dtype == XV2_BF16 ? 8 : 4;
  return (c % v == 0) ? v : 1;
}
static dim3 reduce_grid(int n, long long hw, int c, int vec) {
  int cv = c / vec;
  int cvb = cv < 256 ? cv : 256;
  int lanes = 256 / cvb;
  long long bx = cdiv(hw, (long long)lanes * 16);
  long long cap = cdiv(4LL * kNumSMs, n);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)n);
}
static dim3 ew_grid(int n, long long work) {
  long long bx = cdiv(work, 256 * 4);
  long long cap = cdiv(8LL * kNumSMs, n);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)n);
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_splat_gap(const void* x, float* gap, int32_t n, int64_t hw, int32_t c, int32_t dtype,
                             void* stream) {
  XV2_REQUIRE(n > 0 && hw > 0 && c > 0, "splat_gap: empty");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(gap, 0, sizeof(float) * (size_t)n * c, st);
  const int vec = splat_vec(c, dtype);
  dim3 grid = reduce_grid(n, hw, c, vec);
  const float inv = 1.0f / (float)hw;
  XV2_DISPATCH_DTYPE(dtype, T, {
    if (vec == 1) splat_reduce_kernel<T, 1, 0><<<grid, 256, 0, st>>>((const T*)x, nullptr, gap, hw, c, inv);
    else splat_reduce_kernel<T, Vec<T>::N, 0><<<grid, 256, 0, st>>>((const T*)x, nullptr, gap, hw, c, inv);
  });
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bwd_att(const void* x, const void* dout, float* datt, int32_t n, int64_t hw, int32_t c,
                                 int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && hw > 0 && c > 0, "splat_bwd_att: empty");
  cudaStream_t st = as_stream(stream);
  const int vec = splat_vec(c, dtype);
  dim3 grid = reduce_grid(n, hw, c, vec);
  XV2_DISPATCH_DTYPE(dtype, T, {
    if (vec == 1) splat_reduce_kernel<T, 1, 1><<<grid, 256, 0, st>>>((const T*)x, (const T*)dout, datt, hw, c, 1.f);
    else splat_reduce_kernel<T, Vec<T>::N, 1><<<grid, 256, 0, st>>>((const T*)x, (const T*)dout, datt, hw, c, 1.f);
  });
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_rsoftmax_fwd(const float* logits, float* att, int32_t n, int32_t c, void* stream) {
  XV2_REQUIRE(n > 0 && c > 0, "rsoftmax: empty");
  rsoftmax_fwd_kernel<<<(n * c + 255) / 256, 256, 0, as_stream(stream)>>>(logits, att, n, c);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
extern "C" int xv2_rsoftmax_bwd(const float* att, const float* datt, float* dlogits, int32_t n, int32_t c,
                                void* stream) {
  XV2_REQUIRE(n > 0 && c > 0, "rsoftmax: empty");
  rsoftmax_bwd_kernel<<<(n * c + 255) / 256, 256, 0, as_stream(stream)>>>(att, datt, dlogits, n, c);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_combine(const void* x, const float* att, void* out, int32_t n, int64_t hw, int32_t c,
                                 int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && hw > 0 && c > 0, "splat_combine: empty");
  const int vec = splat_vec(c, dtype);
  dim3 grid = ew_grid(n, hw * (c / vec));
  XV2_DISPATCH_DTYPE(dtype, T, {
    if (vec == 1) splat_combine_kernel<T, 1><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, att, (T*)out, hw, c);
    else splat_combine_kernel<T, Vec<T>::N><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, att, (T*)out, hw, c);
  });
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bwd_x(const void* dout, const float* att, const float* dgap, void* dx, int32_t n,
                               int64_t hw, int32_t c, int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && hw > 0 && c > 0, "splat_bwd_x: empty");
  const int vec = splat_vec(c, dtype);
  dim3 grid = ew_grid(n, hw * (c / vec));
  const float inv = 1.0f / (float)hw;
  XV2_DISPATCH_DTYPE(dtype, T, {
    if (vec == 1) splat_bwd_x_kernel<T, 1><<<grid, 256, 0, as_stream(stream)>>>((const T*)dout, att, dgap, (T*)dx, hw, c, inv);
    else splat_bwd_x_kernel<T, Vec<T>::N><<<grid, 256, 0, as_stream(stream)>>>((const T*)dout, att, dgap, (T*)dx, hw, c, inv);
  });
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
