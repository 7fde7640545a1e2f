// Batch normalisation for NHWC activations: statistics, finalize, apply(+activation, +residual) and the two-pass
// backward.  HBM-bound: every kernel streams [pixels][C] with 16-byte channel vectors, a thread owns one channel
// vector and walks pixels, so per-channel coefficients live in registers.  Replaces cuDNN BN reached from
// layers.py:93,72 and the encoder BNs (unet.py:52).
#include "common.cuh"

namespace xv2 {

// Thread mapping shared by all kernels here: block = 256 threads = lanes x cvb (channel vectors per block pass).
struct RowMap {
  int cv;      // channel vectors per pixel (C / VEC)
  int cvb;     // channel vectors handled per pass (<= 256)
  int lanes;   // pixel lanes per block
  int passes;  // ceil(cv / cvb)
};
static RowMap make_rowmap(int c, int vec) {
  RowMap m;
  m.cv = c / vec;
  m.cvb = m.cv < 256 ? m.cv : 256;
  m.lanes = 256 / m.cvb;
  m.passes = (m.cv + m.cvb - 1) / m.cvb;
  return m;
}
static int pick_vec(int c, int dtype) {
  int v = dtype == XV2_BF16 ? 8 : 4;
  return (c % v == 0) ? v : 1;
}
static int pick_blocks(int64_t pixels, const RowMap& m, int per_thread) {
  int64_t b = cdiv(pixels, (int64_t)m.lanes * per_thread);
  if (b < 1) b = 1;
  if (b > 8 * kNumSMs) b = 8 * kNumSMs;
  return (int)b;
}

template <typename T, int VEC> __device__ __forceinline__ void ldv(const T* p, float* f) {
  if constexpr (VEC == 1) {
    f[0] = to_f(*p);
  } else {
    Vec<T> v;
    v.load(p);
    v.unpack(f);
  }
}
template <typename T, int VEC> __device__ __forceinline__ void stv(T* p, const float* f) {
  if constexpr (VEC == 1) {
    *p = from_f<T>(f[0]);
  } else {
    Vec<T> v;
    v.pack(f);
    v.store(p);
  }
}

// ---- statistics ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long long pixels, int c, RowMap m,
                                                       double* __restrict__ stats) {
  __shared__ float sm[2][256 * (VEC > 4 ? 8 : (VEC > 1 ? 4 : 1))];
  const int tid = threadIdx.x;
  const int cvi0 = tid % m.cvb, lane = tid / m.cvb;
  for (int pass = 0; pass < m.passes; ++pass) {
    const int cvi = pass * m.cvb + cvi0;
    float s[VEC], q[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = q[i] = 0.f;
    if (lane < m.lanes && cvi < m.cv) {
      // 4 independent 16-byte loads in flight per thread (a single load per iteration leaves HBM at ~35 %)
      const long long stride = (long long)gridDim.x * m.lanes;
      long long p = (long long)blockIdx.x * m.lanes + lane;
      const T* xp = x + cvi * VEC;
      for (; p + 3 * stride < pixels; p += 4 * stride) {
        float f[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) ldv<T, VEC>(xp + (p + u * stride) * c, f[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            s[i] += f[u][i];
            q[i] = fmaf(f[u][i], f[u][i], q[i]);
          }
      }
      for (; p < pixels; p += stride) {
        float f[VEC];
        ldv<T, VEC>(xp + p * c, f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          s[i] += f[i];
          q[i] = fmaf(f[i], f[i], q[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sm[0][tid * VEC + i] = s[i];
      sm[1][tid * VEC + i] = q[i];
    }
    __syncthreads();
    if (lane == 0 && cvi < m.cv) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        double ts = 0.0, tq = 0.0;
        for (int l = 0; l < m.lanes; ++l) {
          ts += sm[0][(l * m.cvb + cvi0) * VEC + i];
          tq += sm[1][(l * m.cvb + cvi0) * VEC + i];
        }
        atomicAdd(&stats[cvi * VEC + i], ts);
        atomicAdd(&stats[c + cvi * VEC + i], tq);
      }
    }
    __syncthreads();
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long count, int c,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double n = (double)count;
  const double mu = stats[i] / n;
  double var = stats[c + i] / n - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  mean[i] = (float)mu;
  invstd[i] = is;
  scale[i] = g * is;
  shift[i] = b - (float)mu * g * is;
  if (rmean) {
    const double unbiased = count > 1 ? var * n / (n - 1.0) : var;
    rmean[i] = (1.f - momentum) * rmean[i] + momentum * (float)mu;
    rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coeffs_kernel(int c, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rmean, const float* __restrict__ rvar, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float is = 1.0f / sqrtf(rvar[i] + eps);
  const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  scale[i] = g * is;
  shift[i] = b - rmean[i] * g * is;
}

// ---- apply -----------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                       T* __restrict__ y, long long pixels, int c, RowMap m,
                                                       const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int act) {
  const int tid = threadIdx.x;
  const int cvi0 = tid % m.cvb, lane = tid / m.cvb;
  if (lane >= m.lanes) return;
  for (int pass = 0; pass < m.passes; ++pass) {
    const int cvi = pass * m.cvb + cvi0;
    if (cvi >= m.cv) break;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sc[i] = scale[cvi * VEC + i];
      sh[i] = shift[cvi * VEC + i];
    }
    const long long stride = (long long)gridDim.x * m.lanes;
    long long p = (long long)blockIdx.x * m.lanes + lane;
    for (; p + 3 * stride < pixels; p += 4 * stride) {
      float f[4][VEC], r[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long off = (p + u * stride) * c + cvi * VEC;
        ldv<T, VEC>(x + off, f[u]);
        if (res) ldv<T, VEC>(res + off, r[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float v = fmaf(f[u][i], sc[i], sh[i]);
          if (res) v += r[u][i];
          f[u][i] = apply_act(v, act);
        }
        stv<T, VEC>(y + (p + u * stride) * c + cvi * VEC, f[u]);
      }
    }
    for (; p < pixels; p += stride) {
      const long long off = p * c + cvi * VEC;
      float f[VEC], r[VEC];
      ldv<T, VEC>(x + off, f);
      if (res) ldv<T, VEC>(res + off, r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float u = fmaf(f[i], sc[i], sh[i]);
        if (res) u += r[i];
        f[i] = apply_act(u, act);
      }
      stv<T, VEC>(y + off, f);
    }
  }
}

// ---- backward pass 1: reductions ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const T* __restrict__ res, long long pixels, int c,
                                                            RowMap m, const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, int act,
                                                            double* __restrict__ red) {
  __shared__ float sm[2][256 * (VEC > 4 ? 8 : (VEC > 1 ? 4 : 1))];
  const int tid = threadIdx.x;
  const int cvi0 = tid % m.cvb, lane = tid / m.cvb;
  for (int pass = 0; pass < m.passes; ++pass) {
    const int cvi = pass * m.cvb + cvi0;
    float s1[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s1[i] = s2[i] = 0.f;
    if (lane < m.lanes && cvi < m.cv) {
      float sc[VEC], sh[VEC], mu[VEC], is[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        sc[i] = scale[cvi * VEC + i];
        sh[i] = shift[cvi * VEC + i];
        mu[i] = mean[cvi * VEC + i];
        is[i] = invstd[cvi * VEC + i];
      }
      const long long stride = (long long)gridDim.x * m.lanes;
      long long p = (long long)blockIdx.x * m.lanes + lane;
      auto accumulate = [&](const float* g, const float* f, const float* r) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float du = g[i];
          if (act != XV2_ACT_NONE) {
            float u = fmaf(f[i], sc[i], sh[i]);
            if (res) u += r[i];
            du *= act_grad(u, act);
          }
          s1[i] += du;
          s2[i] = fmaf(du, (f[i] - mu[i]) * is[i], s2[i]);
        }
      };
      for (; p + 1 * stride < pixels; p += 2 * stride) {  // 2 pixels x (dy, x[, res]) = 4-6 loads in flight
        float g[2][VEC], f[2][VEC], r[2][VEC];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const long long off = (p + u * stride) * c + cvi * VEC;
          ldv<T, VEC>(dy + off, g[u]);
          ldv<T, VEC>(x + off, f[u]);
          if (res) ldv<T, VEC>(res + off, r[u]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) accumulate(g[u], f[u], r[u]);
      }
      for (; p < pixels; p += stride) {
        const long long off = p * c + cvi * VEC;
        float g[VEC], f[VEC], r[VEC];
        ldv<T, VEC>(dy + off, g);
        ldv<T, VEC>(x + off, f);
        if (res) ldv<T, VEC>(res + off, r);
        accumulate(g, f, r);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sm[0][tid * VEC + i] = s1[i];
      sm[1][tid * VEC + i] = s2[i];
    }
    __syncthreads();
    if (lane == 0 && cvi < m.cv) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        double t1 = 0.0, t2 = 0.0;
        for (int l = 0; l < m.lanes; ++l) {
          t1 += sm[0][(l * m.cvb + cvi0) * VEC + i];
          t2 += sm[1][(l * m.cvb + cvi0) * VEC + i];
        }
        atomicAdd(&red[cvi * VEC + i], t1);
        atomicAdd(&red[c + cvi * VEC + i], t2);
      }
    }
    __syncthreads();
  }
}

// ---- backward pass 2 ----------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                           const T* __restrict__ res, T* __restrict__ dx,
                                                           T* __restrict__ dres, long long pixels, int c, RowMap m,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, int act,
                                                           const double* __restrict__ red, long long count,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           int accumulate) {
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && red != nullptr && dgamma != nullptr) {
    for (int i = tid; i < c; i += blockDim.x) {
      dbeta[i] = (accumulate ? dbeta[i] : 0.f) + (float)red[i];
      dgamma[i] = (accumulate ? dgamma[i] : 0.f) + (float)red[c + i];
    }
  }
  const int cvi0 = tid % m.cvb, lane = tid / m.cvb;
  if (lane >= m.lanes) return;
  const float inv_n = 1.0f / (float)count;
  for (int pass = 0; pass < m.passes; ++pass) {
    const int cvi = pass * m.cvb + cvi0;
    if (cvi >= m.cv) break;
    float sc[VEC], sh[VEC], mu[VEC], is[VEC], k0[VEC], k1[VEC], k2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int ch = cvi * VEC + i;
      sc[i] = scale[ch];
      sh[i] = shift[ch];
      if (red) {
        mu[i] = mean[ch];
        is[i] = invstd[ch];
        const float g = gamma ? gamma[ch] : 1.f;
        k0[i] = g * is[i];
        k1[i] = (float)(red[ch]) * inv_n;
        k2[i] = (float)(red[c + ch]) * inv_n;
      }
    }
    const long long stride = (long long)gridDim.x * m.lanes;
    long long p = (long long)blockIdx.x * m.lanes + lane;
    auto finish = [&](long long off, float* g, const float* f, const float* r) {
      float o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float du = g[i];
        if (act != XV2_ACT_NONE) {
          float u = fmaf(f[i], sc[i], sh[i]);
          if (res) u += r[i];
          du *= act_grad(u, act);
        }
        g[i] = du;
        if (red) {
          const float xh = (f[i] - mu[i]) * is[i];
          o[i] = k0[i] * (du - k1[i] - xh * k2[i]);
        } else {
          o[i] = du * sc[i];
        }
      }
      stv<T, VEC>(dx + off, o);
      if (dres) stv<T, VEC>(dres + off, g);
    };
    for (; p + 1 * stride < pixels; p += 2 * stride) {
      float g[2][VEC], f[2][VEC], r[2][VEC];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long off = (p + u * stride) * c + cvi * VEC;
        ldv<T, VEC>(dy + off, g[u]);
        ldv<T, VEC>(x + off, f[u]);
        if (res) ldv<T, VEC>(res + off, r[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) finish((p + u * stride) * c + cvi * VEC, g[u], f[u], r[u]);
    }
    for (; p < pixels; p += stride) {
      const long long off = p * c + cvi * VEC;
      float g[VEC], f[VEC], r[VEC];
      ldv<T, VEC>(dy + off, g);
      ldv<T, VEC>(x + off, f);
      if (res) ldv<T, VEC>(res + off, r);
      finish(off, g, f, r);
    }
  }
}

}  // namespace xv2

namespace xv2 {
bool bn_stream_ok(int64_t pixels, int c, int dtype);
int bn_stream_stats(const void* x, int64_t pixels, int c, double* stats, void* stream);
int bn_stream_apply(const void* x, const void* res, void* y, int64_t pixels, int c, const float* scale, const float* shift,
                    int act, void* stream);
int bn_stream_bwd_reduce(const void* dy, const void* dy2, const void* x, const void* res, void* du, int64_t pixels, int c,
                         const float* scale, const float* shift, const float* mean, const float* invstd, int act, double* red,
                         void* stream);
int bn_stream_train_apply(const void* x, const void* res, void* y, int64_t pixels, int c, const double* stats, int64_t count,
                          const float* gamma, const float* beta, float* rmean, float* rvar, float momentum, float eps,
                          float* coef, int act, void* stream);
int bn_stream_bwd_apply(const void* dy, const void* x, const void* res, void* dx, void* dres, int64_t pixels, int c,
                        const float* scale, const float* shift, const float* mean, const float* invstd, const float* gamma,
                        int act, const double* red, int64_t count, float* dgamma, float* dbeta, int accumulate, void* stream);
}  // namespace xv2

using namespace xv2;

#define XV2_DISPATCH_VEC(T, vec, KERNEL, ...)                 \
  do {                                                        \
    if ((vec) == 1) KERNEL<T, 1> __VA_ARGS__;                 \
    else KERNEL<T, Vec<T>::N> __VA_ARGS__;                    \
  } while (0)

extern "C" int xv2_bn_stats(const void* x, int64_t pixels, int32_t c, int32_t dtype, double* stats, void* stream) {
  XV2_REQUIRE(c > 0 && pixels > 0, "bn_stats: empty tensor");
  if (bn_stream_ok(pixels, c, dtype)) return bn_stream_stats(x, pixels, c, stats, stream);
  const int vec = pick_vec(c, dtype);
  RowMap m = make_rowmap(c, vec);
  int blocks = pick_blocks(pixels, m, 32);
  XV2_DISPATCH_DTYPE(dtype, T, XV2_DISPATCH_VEC(T, vec, bn_stats_kernel, <<<blocks, 256, 0, as_stream(stream)>>>(
                                                                         (const T*)x, pixels, c, m, stats)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bn_finalize(const double* stats, int64_t count, int32_t c, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, float* mean,
                               float* invstd, float* scale, float* shift, void* stream) {
  XV2_REQUIRE(c > 0 && count > 0, "bn_finalize: empty");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(stats, count, c, gamma, beta, running_mean,
                                                                     running_var, momentum, eps, mean, invstd, scale,
                                                                     shift);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bn_eval_coeffs(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift, void* stream) {
  XV2_REQUIRE(c > 0, "bn_eval_coeffs: empty");
  bn_eval_coeffs_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(c, gamma, beta, running_mean, running_var, eps,
                                                                        scale, shift);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bn_apply(const void* x, const void* residual, void* y, int64_t pixels, int32_t c, int32_t dtype,
                            const float* scale, const float* shift, int32_t act, void* stream) {
  XV2_REQUIRE(c > 0 && pixels > 0, "bn_apply: empty tensor");
  if (bn_stream_ok(pixels, c, dtype)) return bn_stream_apply(x, residual, y, pixels, c, scale, shift, act, stream);
  const int vec = pick_vec(c, dtype);
  RowMap m = make_rowmap(c, vec);
  int blocks = pick_blocks(pixels, m, 8);
  XV2_DISPATCH_DTYPE(dtype, T, XV2_DISPATCH_VEC(T, vec, bn_apply_kernel, <<<blocks, 256, 0, as_stream(stream)>>>(
                                                                         (const T*)x, (const T*)residual, (T*)y, pixels,
                                                                         c, m, scale, shift, act)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bn_train_apply(const void* x, const void* residual, void* y, int64_t pixels, int32_t c, int32_t dtype,
                                  const double* stats, int64_t count, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps, float* coef,
                                  int32_t act, void* stream) {
  XV2_REQUIRE(c > 0 && pixels > 0 && count > 0 && stats && coef, "bn_train_apply: bad argument");
  if (bn_stream_ok(pixels, c, dtype))
    return bn_stream_train_apply(x, residual, y, pixels, c, stats, count, gamma, beta, running_mean, running_var, momentum, eps,
                                 coef, act, stream);
  int rc = xv2_bn_finalize(stats, count, c, gamma, beta, running_mean, running_var, momentum, eps, coef, coef + c, coef + 2 * c,
                           coef + 3 * c, stream);
  if (rc) return rc;
  return xv2_bn_apply(x, residual, y, pixels, c, dtype, coef + 2 * c, coef + 3 * c, act, stream);
}

extern "C" int xv2_bn_bwd_reduce(const void* dy, const void* x, const void* residual, int64_t pixels, int32_t c,
                                 int32_t dtype, const float* scale, const float* shift, const float* mean,
                                 const float* invstd, int32_t act, double* red, void* stream) {
  XV2_REQUIRE(c > 0 && pixels > 0, "bn_bwd_reduce: empty tensor");
  if (bn_stream_ok(pixels, c, dtype))
    return bn_stream_bwd_reduce(dy, nullptr, x, residual, nullptr, pixels, c, scale, shift, mean, invstd, act, red, stream);
  const int vec = pick_vec(c, dtype);
  RowMap m = make_rowmap(c, vec);
  int blocks = pick_blocks(pixels, m, 32);
  XV2_DISPATCH_DTYPE(dtype, T,
                     XV2_DISPATCH_VEC(T, vec, bn_bwd_reduce_kernel, <<<blocks, 256, 0, as_stream(stream)>>>(
                                                                        (const T*)dy, (const T*)x, (const T*)residual,
                                                                        pixels, c, m, scale, shift, mean, invstd, act,
                                                                        red)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bn_bwd_reduce_du(const void* dy, const void* dy2, const void* x, const void* residual, void* du,
                                    int64_t pixels, int32_t c, int32_t dtype, const float* scale, const float* shift,
                                    const float* mean, const float* invstd, int32_t act, double* red, void* stream) {
  XV2_REQUIRE(c > 0 && pixels > 0 && dy && x && du && red, "bn_bwd_reduce_du: bad argument");
  if (!bn_stream_ok(pixels, c, dtype)) {
    set_error("bn_bwd_reduce_du: only the streaming bf16 path (2048 %% c == 0, >= 1 Mi elements) writes du");
    return XV2_EUNSUPPORTED;
  }
  return bn_stream_bwd_reduce(dy, dy2, x, residual, du, pixels, c, scale, shift, mean, invstd, act, red, stream);
}

extern "C" int xv2_bn_bwd_apply(const void* dy, const void* x, const void* residual, void* dx, void* dres,
                                int64_t pixels, int32_t c, int32_t dtype, const float* scale, const float* shift,
                                const float* mean, const float* invstd, const float* gamma, int32_t act,
                                const double* red, int64_t count, float* dgamma, float* dbeta, int32_t accumulate,
                                void* stream) {
  XV2_REQUIRE(c > 0 && pixels > 0, "bn_bwd_apply: empty tensor");
  if (bn_stream_ok(pixels, c, dtype))
    return bn_stream_bwd_apply(dy, x, residual, dx, dres, pixels, c, scale, shift, mean, invstd, gamma, act, red, count, dgamma,
                               dbeta, accumulate, stream);
  const int vec = pick_vec(c, dtype);
  RowMap m = make_rowmap(c, vec);
  int blocks = pick_blocks(pixels, m, 8);
  XV2_DISPATCH_DTYPE(dtype, T,
                     XV2_DISPATCH_VEC(T, vec, bn_bwd_apply_kernel, <<<blocks, 256, 0, as_stream(stream)>>>(
                                                                       (const T*)dy, (const T*)x, (const T*)residual,
                                                                       (T*)dx, (T*)dres, pixels, c, m, scale, shift,
                                                                       mean, invstd, gamma, act, red,
                                                                       count > 0 ? count : 1, dgamma, dbeta, accumulate)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
