// Resampling operators of the optional model parts on NHWC activations, and the ordinal (CORAL / MSE)
// damage heads:
//   * adaptive average pooling to bin x bin (PPM, layers.py:12-21) forward / backward;
//   * bilinear resize with align_corners=True (PPM layers.py:27, --dec_interp layers.py:154, --interpolate layers.py:186-188)
//     forward / backward -- same arithmetic as ATen's upsample_bilinear2d (fp32 source index = dst * (in-1)/(out-1),
//     lambda weights, h0*(w0*a + w1*b) + h1*(w0*c + w1*d));
//   * CORAL (loss.py:54-65) and MSE (loss.py:92-94, nn.MSELoss) losses with the `post` masking of loss.py:86-90: one reduction
//     pass + one backward pass; label decoding of both heads (utils/f1.py:7-15) for the F1 counters and Model.save.
// All streaming / HBM-bound; channel-contiguous 16-byte accesses where the channel count allows.
#include "common.cuh"

namespace xv2 {

// ---------------------------------------------------------------------------------------------------------------
// adaptive average pool: region of output index i over an axis of length L split in B bins = [floor(i L / B), ceil((i+1) L / B))
__device__ __forceinline__ int bin_lo(int i, int len, int bins) { return (i * len) / bins; }
__device__ __forceinline__ int bin_hi(int i, int len, int bins) { return ((i + 1) * len + bins - 1) / bins; }

template <typename T>
__global__ void __launch_bounds__(256) adaptive_pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h,
                                                                int w, int c, int bins) {
  const long long total = (long long)n * bins * bins * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int bw = (int)(t % bins);
    t /= bins;
    const int bh = (int)(t % bins);
    const int img = (int)(t / bins);
    const int h0 = bin_lo(bh, h, bins), h1 = bin_hi(bh, h, bins), w0 = bin_lo(bw, w, bins), w1 = bin_hi(bw, w, bins);
    float s = 0.f;
    for (int hh = h0; hh < h1; ++hh)
      for (int ww = w0; ww < w1; ++ww) s += to_f(x[(((long long)img * h + hh) * w + ww) * c + ch]);
    y[i] = from_f<T>(s / (float)((h1 - h0) * (w1 - w0)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) adaptive_pool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h,
                                                                int w, int c, int bins) {
  const long long total = (long long)n * h * w * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int ww = (int)(t % w);
    t /= w;
    const int hh = (int)(t % h);
    const int img = (int)(t / h);
    float s = 0.f;
    for (int bh = 0; bh < bins; ++bh) {
      const int h0 = bin_lo(bh, h, bins), h1 = bin_hi(bh, h, bins);
      if (hh < h0 || hh >= h1) continue;
      for (int bw = 0; bw < bins; ++bw) {
        const int w0 = bin_lo(bw, w, bins), w1 = bin_hi(bw, w, bins);
        if (ww < w0 || ww >= w1) continue;
        s += to_f(dy[(((long long)img * bins + bh) * bins + bw) * c + ch]) / (float)((h1 - h0) * (w1 - w0));
      }
    }
    dx[i] = from_f<T>(s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bilinear, align_corners = True
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_at(int o, int in, int out) {
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const float src = scale * (float)o;
  Lerp r;
  r.i0 = (int)src;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w,
                                                           int c, int oh, int ow) {
  const long long total = (long long)n * oh * ow * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int xo = (int)(t % ow);
    t /= ow;
    const int yo = (int)(t % oh);
    const int img = (int)(t / oh);
    const Lerp a = lerp_at(yo, h, oh), b = lerp_at(xo, w, ow);
    const T* base = x + (long long)img * h * w * c + ch;
    const float v00 = to_f(base[((long long)a.i0 * w + b.i0) * c]), v01 = to_f(base[((long long)a.i0 * w + b.i1) * c]);
    const float v10 = to_f(base[((long long)a.i1 * w + b.i0) * c]), v11 = to_f(base[((long long)a.i1 * w + b.i1) * c]);
    y[i] = from_f<T>(a.l0 * (b.l0 * v00 + b.l1 * v01) + a.l1 * (b.l0 * v10 + b.l1 * v11));
  }
}

// backward: scatter with fp32 atomics into a zero-filled fp32 accumulator (the caller casts when the activation is bf16)
template <typename T>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const T* __restrict__ dy, float* __restrict__ dx, int n, int h,
                                                           int w, int c, int oh, int ow) {
  const long long total = (long long)n * oh * ow * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int xo = (int)(t % ow);
    t /= ow;
    const int yo = (int)(t % oh);
    const int img = (int)(t / oh);
    const Lerp a = lerp_at(yo, h, oh), b = lerp_at(xo, w, ow);
    const float g = to_f(dy[i]);
    float* base = dx + (long long)img * h * w * c + ch;
    atomicAdd(base + ((long long)a.i0 * w + b.i0) * c, a.l0 * b.l0 * g);
    atomicAdd(base + ((long long)a.i0 * w + b.i1) * c, a.l0 * b.l1 * g);
    atomicAdd(base + ((long long)a.i1 * w + b.i0) * c, a.l1 * b.l0 * g);
    atomicAdd(base + ((long long)a.i1 * w + b.i1) * c, a.l1 * b.l1 * g);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ordinal heads.  mode 0 = MSE on relu(logit[0]) (1 logit), mode 1 = CORAL (3 rank logits).
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

template <int MODE>
__global__ void __launch_bounds__(256) ordinal_partials_kernel(const float* __restrict__ logits,
                                                               const uint8_t* __restrict__ labels, long long pixels,
                                                               int post, double* __restrict__ sums) {
  constexpr int NL = MODE == 0 ? 1 : 3;
  float acc = 0.f, cnt = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    int t = labels[p];
    if (post) {
      if (t == 0) continue;
      t -= 1;
    }
    if (MODE == 0) {
      const float d = fmaxf(logits[p], 0.f) - (float)t;
      acc += d * d;
    } else {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NL; ++c) {
        const float x = logits[p * NL + c], ls = log_sigmoid(x);
        s += (c < t) ? ls : (ls - x);  // levels[t][c] = (c < t)
      }
      acc -= s;
    }
    cnt += 1.f;
  }
  __shared__ double sm[8][2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double a = warp_sum((double)acc), b = warp_sum((double)cnt);
  if (lane == 0) {
    sm[wid][0] = a;
    sm[wid][1] = b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

__global__ void ordinal_finalize_kernel(const double* __restrict__ sums, float weight, float* __restrict__ loss,
                                        float* __restrict__ coef) {
  const double m = sums[1];
  loss[0] = (float)(weight * sums[0] / m);  // mean over kept pixels (NaN on an empty mask, like the reference)
  coef[0] = (float)(weight / m);
}

template <int MODE>
__global__ void __launch_bounds__(256) ordinal_backward_kernel(const float* __restrict__ logits,
                                                               const uint8_t* __restrict__ labels, long long pixels,
                                                               int post, const float* __restrict__ coef,
                                                               const float* __restrict__ upstream, float* __restrict__ dl) {
  constexpr int NL = MODE == 0 ? 1 : 3;
  const float k = coef[0] * upstream[0];
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    int t = labels[p];
    bool keep = true;
    if (post) {
      keep = t != 0;
      t -= 1;
    }
    if (MODE == 0) {
      const float x = logits[p];
      dl[p] = (keep && x > 0.f) ? 2.f * (x - (float)t) * k : 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < NL; ++c) {
        const float x = logits[p * NL + c];
        const float sig = 1.f / (1.f + expf(-x));
        dl[p * NL + c] = keep ? -((c < t ? 1.f : 0.f) - sig) * k : 0.f;
      }
    }
  }
}

// utils/f1.py:7-15 label decoding; optional F1 counters (tp | fp | fn for classes 1..4 over pixels with target > 0) and/or a
// label map: mode 0 -> round-half-even(relu(x0)) + 1 clamped to 4, mode 1 -> #(x_c > 0) + 1  (sigmoid(x) > 0.5 <=> x > 0)
template <int MODE>
__global__ void __launch_bounds__(256) ordinal_labels_kernel(const float* __restrict__ logits,
                                                             const uint8_t* __restrict__ labels, long long pixels,
                                                             int clamp4, unsigned long long* __restrict__ counters,
                                                             uint8_t* __restrict__ pred_u8, float* __restrict__ pred_f32) {
  constexpr int NL = MODE == 0 ? 1 : 3;
  __shared__ unsigned int cnt[12];
  if (threadIdx.x < 12) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    float lab;
    if (MODE == 0) {
      lab = rintf(fmaxf(logits[p], 0.f)) + 1.f;
      if (clamp4 && lab > 4.f) lab = 4.f;
    } else {
      int s = 1;
#pragma unroll
      for (int c = 0; c < NL; ++c) s += logits[p * NL + c] > 0.f ? 1 : 0;
      lab = (float)s;
    }
    if (pred_u8) pred_u8[p] = (uint8_t)fminf(lab, 255.f);
    if (pred_f32) pred_f32[p] = lab;
    if (counters) {
      const int t = labels[p];
      if (t > 0) {
        const int pr = (int)fminf(lab, 255.f);
#pragma unroll
        for (int c = 1; c <= 4; ++c) {
          if (pr == c && t == c) atomicAdd(&cnt[c - 1], 1u);
          if (pr == c && t != c) atomicAdd(&cnt[4 + c - 1], 1u);
          if (pr != c && t == c) atomicAdd(&cnt[8 + c - 1], 1u);
        }
      }
    }
  }
  __syncthreads();
  if (counters && threadIdx.x < 12 && cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

static int grid_for(long long total, int per_block = 256, int max_blocks = kNumSMs * 16) {
  long long b = (total + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (int)(b > max_blocks ? max_blocks : b);
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_adaptive_avgpool_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t bins,
                                        int32_t dtype, void* stream) {
  XV2_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && bins > 0, "adaptive_avgpool_fwd: bad argument");
  const long long total = (long long)n * bins * bins * c;
  XV2_DISPATCH_DTYPE(dtype, T, (adaptive_pool_fwd_kernel<T><<<grid_for(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (T*)y, n, h, w, c, bins)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_adaptive_avgpool_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t bins,
                                        int32_t dtype, void* stream) {
  XV2_REQUIRE(dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && bins > 0, "adaptive_avgpool_bwd: bad argument");
  const long long total = (long long)n * h * w * c;
  XV2_DISPATCH_DTYPE(dtype, T, (adaptive_pool_bwd_kernel<T><<<grid_for(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)dy, (T*)dx, n, h, w, c, bins)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bilinear_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh, int32_t ow,
                                int32_t dtype, void* stream) {
  XV2_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "bilinear_fwd: bad argument");
  const long long total = (long long)n * oh * ow * c;
  XV2_DISPATCH_DTYPE(dtype, T, (bilinear_fwd_kernel<T><<<grid_for(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (T*)y, n, h, w, c, oh, ow)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bilinear_bwd(const void* dy, float* dx_f32, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                                int32_t ow, int32_t dtype, void* stream) {
  XV2_REQUIRE(dy && dx_f32 && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "bilinear_bwd: bad argument");
  const long long total = (long long)n * oh * ow * c;
  XV2_DISPATCH_DTYPE(dtype, T, (bilinear_bwd_kernel<T><<<grid_for(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)dy, dx_f32, n, h, w, c, oh, ow)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_ordinal_loss_partials(const float* logits, const uint8_t* labels, int64_t pixels, int32_t mode,
                                         int32_t post, double* sums, void* stream) {
  XV2_REQUIRE(logits && labels && sums && pixels > 0 && (mode == 0 || mode == 1), "ordinal_loss_partials: bad argument");
  const int grid = grid_for(pixels);
  if (mode == 0) ordinal_partials_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(logits, labels, pixels, post, sums);
  else ordinal_partials_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(logits, labels, pixels, post, sums);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_ordinal_loss_finalize(const double* sums, float weight, float* loss, float* coef, void* stream) {
  XV2_REQUIRE(sums && loss && coef, "ordinal_loss_finalize: null argument");
  ordinal_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(sums, weight, loss, coef);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_ordinal_loss_backward(const float* logits, const uint8_t* labels, int64_t pixels, int32_t mode,
                                         int32_t post, const float* coef, const float* upstream, float* dlogits,
                                         void* stream) {
  XV2_REQUIRE(logits && labels && coef && upstream && dlogits && pixels > 0 && (mode == 0 || mode == 1),
              "ordinal_loss_backward: bad argument");
  const int grid = grid_for(pixels);
  if (mode == 0)
    ordinal_backward_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(logits, labels, pixels, post, coef, upstream, dlogits);
  else
    ordinal_backward_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(logits, labels, pixels, post, coef, upstream, dlogits);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_ordinal_labels(const float* logits, const uint8_t* labels, int64_t pixels, int32_t mode, int32_t clamp4,
                                  int64_t* counters, uint8_t* pred_u8, float* pred_f32, void* stream) {
  XV2_REQUIRE(logits && pixels > 0 && (mode == 0 || mode == 1), "ordinal_labels: bad argument");
  XV2_REQUIRE(!counters || labels, "ordinal_labels: counters need labels");
  const int grid = grid_for(pixels);
  auto cnt = reinterpret_cast<unsigned long long*>(counters);
  if (mode == 0) ordinal_labels_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(logits, labels, pixels, clamp4, cnt, pred_u8, pred_f32);
  else ordinal_labels_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(logits, labels, pixels, clamp4, cnt, pred_u8, pred_f32);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
