// Memory-bound element-wise kernels: residual add + activation, activation backward, attention gate, flips (TTA),
// casts, the 1x1 output head, the uint8-tile normaliser of the loader and fused AdamW.
#include "common.cuh"

namespace xv2 {

static int ew_blocks(long long work) {
  long long b = cdiv(work, 256);
  if (b > 16 * kNumSMs) b = 16 * kNumSMs;
  return (int)(b < 1 ? 1 : b);
}

template <typename T>
__global__ void add_act_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, long long nvec,
                               long long numel, int act) {
  constexpr int V = Vec<T>::N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    Vec<T> va, vb;
    float fa[V], fb[V];
    va.load(a + i * V);
    vb.load(b + i * V);
    va.unpack(fa);
    vb.unpack(fb);
#pragma unroll
    for (int j = 0; j < V; ++j) fa[j] = apply_act(fa[j] + fb[j], act);
    va.pack(fa);
    va.store(y + i * V);
  }
  if (blockIdx.x == 0)
    for (long long i = nvec * V + threadIdx.x; i < numel; i += blockDim.x)
      y[i] = from_f<T>(apply_act(to_f(a[i]) + to_f(b[i]), act));
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long nvec,
                               long long numel, int act) {
  constexpr int V = Vec<T>::N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    Vec<T> vd, vy;
    float fd[V], fy[V];
    vd.load(dy + i * V);
    vy.load(y + i * V);
    vd.unpack(fd);
    vy.unpack(fy);
#pragma unroll
    for (int j = 0; j < V; ++j) fd[j] *= act_grad(fy[j], act);
    vd.pack(fd);
    vd.store(dx + i * V);
  }
  if (blockIdx.x == 0)
    for (long long i = nvec * V + threadIdx.x; i < numel; i += blockDim.x)
      dx[i] = from_f<T>(to_f(dy[i]) * act_grad(to_f(y[i]), act));
}

// one warp per pixel row chunk: out[p][c] = skip[p][c] * sigmoid(psi[p])
template <typename T>
__global__ void gate_fwd_kernel(const T* __restrict__ skip, const T* __restrict__ psi, T* __restrict__ out,
                                long long pixels, int c) {
  constexpr int V = Vec<T>::N;
  const int cv = c / V;
  const long long total = pixels * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / cv;
    const float sg = 1.f / (1.f + expf(-to_f(psi[p])));
    Vec<T> v;
    float f[V];
    v.load(skip + i * V);
    v.unpack(f);
#pragma unroll
    for (int j = 0; j < V; ++j) f[j] *= sg;
    v.pack(f);
    v.store(out + i * V);
  }
}

// one warp per pixel: dskip = dout*sig ; dpsi = sig(1-sig) * <dout, skip>
template <typename T>
__global__ void gate_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ skip, const T* __restrict__ psi,
                                T* __restrict__ dskip, T* __restrict__ dpsi, long long pixels, int c) {
  constexpr int V = Vec<T>::N;
  const int cv = c / V;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < pixels; p += nwarps) {
    const float sg = 1.f / (1.f + expf(-to_f(psi[p])));
    float dot = 0.f;
    for (int cvi = lane; cvi < cv; cvi += 32) {
      Vec<T> vd, vs;
      float fd[V], fs[V];
      vd.load(dout + p * c + cvi * V);
      vs.load(skip + p * c + cvi * V);
      vd.unpack(fd);
      vs.unpack(fs);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        dot = fmaf(fd[j], fs[j], dot);
        fd[j] *= sg;
      }
      vd.pack(fd);
      vd.store(dskip + p * c + cvi * V);
    }
    dot = warp_sum(dot);
    if (lane == 0) dpsi[p] = from_f<T>(dot * sg * (1.f - sg));
  }
}

template <typename T>
__global__ void flip_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c, int fh, int fw) {
  // element granularity keeps it general (c = 2/3/4/6 for images and logits)
  const long long total = (long long)n * h * w * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int ww = (int)(t % w);
    t /= w;
    const int hh = (int)(t % h);
    const int nb = (int)(t / h);
    const int sh = fh ? h - 1 - hh : hh, sw = fw ? w - 1 - ww : ww;
    y[i] = x[(((long long)nb * h + sh) * w + sw) * c + ch];
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long numel) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f<TD>(to_f(s[i]));
}

__global__ void mean4_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                             const float* __restrict__ d, float* __restrict__ out, long long numel) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x)
    out[i] = (((a[i] + b[i]) + c[i]) + d[i]) * 0.25f;  // same order as plt.py:44-47 (+= then /4)
}

// ---- 1x1 output head: a thread owns a pixel, weights live in shared memory ---------------------------------
template <typename T, int NCLS>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ logits,
                                                       long long pixels, int c) {
  extern __shared__ float sw[];  // [NCLS][c]
  for (int i = threadIdx.x; i < NCLS * c; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  constexpr int V = Vec<T>::N;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    float acc[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) acc[k] = b ? b[k] : 0.f;
    for (int c0 = 0; c0 < c; c0 += V) {
      Vec<T> v;
      float f[V];
      v.load(x + p * c + c0);
      v.unpack(f);
#pragma unroll
      for (int k = 0; k < NCLS; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[k] = fmaf(f[j], sw[k * c + c0 + j], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < NCLS; ++k) logits[p * NCLS + k] = acc[k];
  }
}

// dx[p][c] = sum_k dl[p][k] w[k][c];  dw[k][c] += sum_p dl[p][k] x[p][c];  db[k] += sum_p dl[p][k]
// block = 256 threads = lanes x cv channel vectors; thread owns a channel vector and walks pixels.
template <typename T, int NCLS>
__global__ void __launch_bounds__(256) head_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ dl, T* __restrict__ dx,
                                                       float* __restrict__ dw, float* __restrict__ db,
                                                       long long pixels, int c) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[256];
  const int cv = c / V;  // host guarantees cv <= 256 and 256 % cv == 0
  const int lanes = 256 / cv;
  const int tid = threadIdx.x, cvi = tid % cv, lane = tid / cv;
  float wr[NCLS][V], gw[NCLS][V], gb[NCLS];
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
    gb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      wr[k][j] = w[k * c + cvi * V + j];
      gw[k][j] = 0.f;
    }
  }
  for (long long p = (long long)blockIdx.x * lanes + lane; p < pixels; p += (long long)gridDim.x * lanes) {
    float g[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) g[k] = dl[p * NCLS + k];
    Vec<T> v;
    float f[V], o[V];
    v.load(x + p * c + cvi * V);
    v.unpack(f);
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = 0.f;
#pragma unroll
    for (int k = 0; k < NCLS; ++k) {
      gb[k] += g[k];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        o[j] = fmaf(g[k], wr[k][j], o[j]);
        gw[k][j] = fmaf(g[k], f[j], gw[k][j]);
      }
    }
    v.pack(o);
    v.store(dx + p * c + cvi * V);
  }
  // reduce gw over lanes (threads with equal cvi), then one atomic per (k, channel) per block
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      red[tid] = gw[k][j];
      __syncthreads();
      if (lane == 0) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[l * cv + cvi];
        atomicAdd(&dw[k * c + cvi * V + j], t);
      }
      __syncthreads();
    }
    red[tid] = (cvi == 0) ? gb[k] : 0.f;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += red[l * cv];
      atomicAdd(&db[k], t);
    }
    __syncthreads();
  }
}

// ---- loader: uint8 HWC (pre [+ post]) -> normalised NHWC 3 or 6 channels -------------------------------------
template <typename T>
__global__ void normalize_kernel(const uint8_t* __restrict__ pre, const uint8_t* __restrict__ post,
                                 T* __restrict__ out, long long pixels) {
  const float mean[3] = {0.485f * 255.f, 0.456f * 255.f, 0.406f * 255.f};
  const float inv[3] = {1.f / (0.229f * 255.f), 1.f / (0.224f * 255.f), 1.f / (0.225f * 255.f)};
  const int oc = post ? 6 : 3;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[p * oc + j] = from_f<T>(((float)pre[p * 3 + j] - mean[j]) * inv[j]);
    if (post) {
#pragma unroll
      for (int j = 0; j < 3; ++j) out[p * oc + 3 + j] = from_f<T>(((float)post[p * 3 + j] - mean[j]) * inv[j]);
    }
  }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long numel, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i] * (1.f - lr * wd);
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// Same update with the per-step scalars read from DEVICE memory (h = lr, beta1, beta2, eps, wd, 1 - beta1^t, sqrt(1 - beta2^t),
// grad scale): the launch carries no host scalar that changes between steps, so it can be replayed from a CUDA graph.
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long numel, const float* __restrict__ h) {
  const float lr = h[0], b1 = h[1], b2 = h[2], eps = h[3], wd = h[4], bc1 = h[5], bc2_sqrt = h[6], gscale = h[7];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i] * (1.f - lr * wd);
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}
// h = lr, momentum, grad scale, first (1.0 on the first step)
__global__ void sgd_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long numel,
                               const float* __restrict__ h) {
  const float lr = h[0], mu = h[1], gscale = h[2];
  const bool first = h[3] != 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float b = first ? gi : mu * buf[i] + gi;
    buf[i] = b;
    p[i] -= lr * b;
  }
}

// SGD with momentum (dampening 0, no nesterov): buf = mu*buf + g (buf = g on the first step); p -= lr*buf
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long numel,
                           float lr, float mu, float gscale, int first) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float b = first ? gi : mu * buf[i] + gi;
    buf[i] = b;
    p[i] -= lr * b;
  }
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_add_act(const void* a, const void* b, void* y, int64_t numel, int32_t dtype, int32_t act,
                           void* stream) {
  XV2_REQUIRE(numel > 0, "add_act: empty");
  XV2_DISPATCH_DTYPE(dtype, T, {
    const long long nvec = numel / Vec<T>::N;
    add_act_kernel<T><<<ew_blocks(nvec), 256, 0, as_stream(stream)>>>((const T*)a, (const T*)b, (T*)y, nvec, numel, act);
  });
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_act_bwd(const void* dy, const void* y, void* dx, int64_t numel, int32_t dtype, int32_t act,
                           void* stream) {
  XV2_REQUIRE(numel > 0, "act_bwd: empty");
  XV2_DISPATCH_DTYPE(dtype, T, {
    const long long nvec = numel / Vec<T>::N;
    act_bwd_kernel<T><<<ew_blocks(nvec), 256, 0, as_stream(stream)>>>((const T*)dy, (const T*)y, (T*)dx, nvec, numel, act);
  });
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_gate_fwd(const void* skip, const void* psi, void* out, int64_t pixels, int32_t c, int32_t dtype,
                            void* stream) {
  XV2_REQUIRE(pixels > 0 && c > 0, "gate: empty");
  XV2_REQUIRE(c % (dtype == XV2_BF16 ? 8 : 4) == 0, "gate: channels %d not a multiple of the vector width", c);
  XV2_DISPATCH_DTYPE(dtype, T, (gate_fwd_kernel<T><<<ew_blocks(pixels * (c / Vec<T>::N)), 256, 0, as_stream(stream)>>>(
                                   (const T*)skip, (const T*)psi, (T*)out, pixels, c)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_gate_bwd(const void* dout, const void* skip, const void* psi, void* dskip, void* dpsi,
                            int64_t pixels, int32_t c, int32_t dtype, void* stream) {
  XV2_REQUIRE(pixels > 0 && c > 0, "gate: empty");
  XV2_REQUIRE(c % (dtype == XV2_BF16 ? 8 : 4) == 0, "gate: channels %d not a multiple of the vector width", c);
  XV2_DISPATCH_DTYPE(dtype, T, (gate_bwd_kernel<T><<<ew_blocks(pixels * 32), 256, 0, as_stream(stream)>>>(
                                   (const T*)dout, (const T*)skip, (const T*)psi, (T*)dskip, (T*)dpsi, pixels, c)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_flip(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t flip_h,
                        int32_t flip_w, int32_t dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0, "flip: empty");
  XV2_DISPATCH_DTYPE(dtype, T, (flip_kernel<T><<<ew_blocks((long long)n * h * w * c), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (T*)y, n, h, w, c, flip_h, flip_w)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t numel,
                        void* stream) {
  XV2_REQUIRE(numel > 0, "cast: empty");
  const int blocks = ew_blocks(numel);
  cudaStream_t st = as_stream(stream);
  if (src_dtype == XV2_F32 && dst_dtype == XV2_BF16)
    cast_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, numel);
  else if (src_dtype == XV2_BF16 && dst_dtype == XV2_F32)
    cast_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, numel);
  else if (src_dtype == XV2_F32 && dst_dtype == XV2_F32)
    cast_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, (float*)dst, numel);
  else if (src_dtype == XV2_BF16 && dst_dtype == XV2_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, numel);
  else
    XV2_REQUIRE(false, "cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_mean4(const float* a, const float* b, const float* c, const float* d, float* out, int64_t numel,
                         void* stream) {
  XV2_REQUIRE(numel > 0, "mean4: empty");
  mean4_kernel<<<ew_blocks(numel), 256, 0, as_stream(stream)>>>(a, b, c, d, out, numel);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_head_fwd(const void* x, const float* w, const float* b, float* logits, int64_t pixels, int32_t c,
                            int32_t ncls, int32_t dtype, void* stream) {
  XV2_REQUIRE(pixels > 0 && c > 0, "head: empty");
  XV2_REQUIRE(ncls == 1 || ncls == 2 || ncls == 3 || ncls == 4, "head: ncls %d unsupported", ncls);
  XV2_REQUIRE(c % (dtype == XV2_BF16 ? 8 : 4) == 0, "head: channels %d not a multiple of the vector width", c);
  const int blocks = ew_blocks(pixels);
  const size_t sm = sizeof(float) * ncls * c;
  cudaStream_t st = as_stream(stream);
#define XV2_HEAD_F(N) head_fwd_kernel<T, N><<<blocks, 256, sm, st>>>((const T*)x, w, b, logits, pixels, c)
  XV2_DISPATCH_DTYPE(dtype, T, {
    if (ncls == 1) XV2_HEAD_F(1);
    else if (ncls == 2) XV2_HEAD_F(2);
    else if (ncls == 3) XV2_HEAD_F(3);
    else XV2_HEAD_F(4);
  });
#undef XV2_HEAD_F
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_head_bwd(const void* x, const float* w, const float* dlogits, void* dx, float* dw, float* db,
                            int64_t pixels, int32_t c, int32_t ncls, int32_t dtype, void* stream) {
  XV2_REQUIRE(pixels > 0 && c > 0, "head: empty");
  XV2_REQUIRE(ncls == 1 || ncls == 2 || ncls == 3 || ncls == 4, "head: ncls %d unsupported", ncls);
  const int vecw = dtype == XV2_BF16 ? 8 : 4;
  XV2_REQUIRE(c % vecw == 0 && (c / vecw) <= 256 && 256 % (c / vecw) == 0,
              "head_bwd: channels %d must give a power-of-two vector count <= 256", c);
  const int lanes = 256 / (c / vecw);
  long long blocks = cdiv(pixels, (long long)lanes * 64);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = as_stream(stream);
#define XV2_HEAD_B(N) head_bwd_kernel<T, N><<<(int)blocks, 256, 0, st>>>((const T*)x, w, dlogits, (T*)dx, dw, db, pixels, c)
  XV2_DISPATCH_DTYPE(dtype, T, {
    if (ncls == 1) XV2_HEAD_B(1);
    else if (ncls == 2) XV2_HEAD_B(2);
    else if (ncls == 3) XV2_HEAD_B(3);
    else XV2_HEAD_B(4);
  });
#undef XV2_HEAD_B
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_normalize_tiles(const uint8_t* pre, const uint8_t* post, void* out, int32_t n, int32_t h,
                                   int32_t w, int32_t out_dtype, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && pre != nullptr, "normalize_tiles: empty");
  const long long pixels = (long long)n * h * w;
  XV2_DISPATCH_DTYPE(out_dtype, T, (normalize_kernel<T><<<ew_blocks(pixels), 256, 0, as_stream(stream)>>>(
                                       pre, post, (T*)out, pixels)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_adamw(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream) {
  XV2_REQUIRE(numel > 0 && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<ew_blocks(numel), 256, 0, as_stream(stream)>>>(p, g, m, v, numel, lr, beta1, beta2, eps, weight_decay,
                                                                bc1, bc2s, grad_scale);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_adamw_dev(float* p, const float* g, float* m, float* v, int64_t numel, const float* hyper, void* stream) {
  XV2_REQUIRE(numel > 0 && hyper, "adamw_dev: bad arguments");
  adamw_dev_kernel<<<ew_blocks(numel), 256, 0, as_stream(stream)>>>(p, g, m, v, numel, hyper);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_sgd_dev(float* p, const float* g, float* buf, int64_t numel, const float* hyper, void* stream) {
  XV2_REQUIRE(numel > 0 && hyper, "sgd_dev: bad arguments");
  sgd_dev_kernel<<<ew_blocks(numel), 256, 0, as_stream(stream)>>>(p, g, buf, numel, hyper);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_sgd(float* p, const float* g, float* buf, int64_t numel, float lr, float momentum, float grad_scale,
                       int32_t step, void* stream) {
  XV2_REQUIRE(numel > 0 && step >= 1, "sgd: bad arguments");
  sgd_kernel<<<ew_blocks(numel), 256, 0, as_stream(stream)>>>(p, g, buf, numel, lr, momentum, grad_scale, step == 1);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
