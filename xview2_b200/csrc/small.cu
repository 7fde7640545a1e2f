// Small-operand kernels that do not fit the tensor-core tiles: the Cin = 3 stride-2 stem convolution of the ResNeSt deep stem
// (unet.py:52 -> resnest conv1[0]: 2.6 % of the FLOPs, K = 27) and the per-image fully connected layers of split attention
// (fc1 / fc2 on [batch][C] vectors).  CUDA-core, fp32 accumulate; HBM- / latency-bound.
#include "common.cuh"

namespace xv2 {

// ---- stem conv forward: x (n, h, w, 3) bf16 -> y (n, h/2, w/2, K) bf16, 3x3 stride 2 pad 1, weights fp32 [K][3][3][3] ------------
template <int K>
__global__ void __launch_bounds__(256) stem_conv_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                            __nv_bfloat16* __restrict__ y, int n, int h, int wd, int oh, int ow) {
  __shared__ __align__(16) float ws[27][K];  // tap-major so that a thread reads 8 consecutive output channels per LDS.128 pair
  for (int i = threadIdx.x; i < 27 * K; i += blockDim.x) ws[i % 27][i / 27] = w[i];
  __syncthreads();
  const long long total = (long long)n * oh * ow;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % ow);
    long long t = p / ow;
    const int oy = (int)(t % oh);
    const int img = (int)(t / oh);
    float in[27];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = 2 * oy - 1 + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = 2 * ox - 1 + s;
        const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < wd;
        const __nv_bfloat16* px = x + (((long long)img * h + iy) * wd + ix) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) in[(r * 3 + s) * 3 + c] = ok ? __bfloat162float(px[c]) : 0.f;
      }
    }
    __nv_bfloat16* o = y + p * K;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int tpi = 0; tpi < 27; ++tpi) {
        const float4 wa = *reinterpret_cast<const float4*>(&ws[tpi][k0]);
        const float4 wb = *reinterpret_cast<const float4*>(&ws[tpi][k0 + 4]);
        const float v = in[tpi];
        acc[0] = fmaf(v, wa.x, acc[0]);
        acc[1] = fmaf(v, wa.y, acc[1]);
        acc[2] = fmaf(v, wa.z, acc[2]);
        acc[3] = fmaf(v, wa.w, acc[3]);
        acc[4] = fmaf(v, wb.x, acc[4]);
        acc[5] = fmaf(v, wb.y, acc[5]);
        acc[6] = fmaf(v, wb.z, acc[6]);
        acc[7] = fmaf(v, wb.w, acc[7]);
      }
      Vec<__nv_bfloat16> ov;
      ov.pack(acc);
      ov.store(o + k0);
    }
  }
}

// ---- stem conv weight gradient: dw[k][tap][c] += sum_p dy[p][k] * x[p*2 + tap][c] --------------------------------------------------
// lane = (tap, c) index (27 of 32 lanes): one x value per lane and pixel, dy[p][0..K) read at a warp-uniform address (broadcast).
template <int K>
__global__ void __launch_bounds__(256) stem_conv_wgrad_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int n,
                                                              int h, int wd, int oh, int ow) {
  constexpr int TP = 64;                       // output pixels per tile (a segment of one output row)
  constexpr int XW = (2 * TP + 1) * 3;         // input values per patch row: columns 2*ox0 - 1 .. 2*ox0 + 2*TP - 1, 3 channels
  __shared__ float red[27][K];
  __shared__ __align__(16) float dys[TP][K];  // dy tile converted to fp32 once, so the inner loop is LDS.128 + FFMA only
  __shared__ float xs[3][XW];
  for (int i = threadIdx.x; i < 27 * K; i += blockDim.x) red[i / K][i % K] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tp = lane < 27 ? lane : 26;
  const int r = tp / 9, s = (tp / 3) % 3, c = tp % 3;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  const int tiles_w = (ow + TP - 1) / TP;
  const long long tiles = (long long)n * oh * tiles_w;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const long long t2 = tile / tiles_w;
    const int oy = (int)(t2 % oh), img = (int)(t2 / oh);
    const int ox0 = tw * TP;
    const int npx = min(TP, ow - ox0);
    __syncthreads();  // previous tile consumed (and `red` initialised on the first pass)
    for (int i = threadIdx.x; i < TP * K / 8; i += blockDim.x) {
      const int px = i / (K / 8), v = i % (K / 8);
      uint4 q = make_uint4(0, 0, 0, 0);
      if (px < npx) q = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)img * oh + oy) * ow + ox0 + px) * K) + v);
      const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
      float4 lo = make_float4(__uint_as_float(wds[0] << 16), __uint_as_float(wds[0] & 0xffff0000u), __uint_as_float(wds[1] << 16),
                              __uint_as_float(wds[1] & 0xffff0000u));
      float4 hi = make_float4(__uint_as_float(wds[2] << 16), __uint_as_float(wds[2] & 0xffff0000u), __uint_as_float(wds[3] << 16),
                              __uint_as_float(wds[3] & 0xffff0000u));
      *reinterpret_cast<float4*>(&dys[px][v * 8]) = lo;
      *reinterpret_cast<float4*>(&dys[px][v * 8 + 4]) = hi;
    }
    for (int i = threadIdx.x; i < 3 * XW; i += blockDim.x) {
      const int rr = i / XW, j = i % XW;
      const int iy = 2 * oy - 1 + rr, ix = 2 * ox0 - 1 + j / 3;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < wd) v = __bfloat162float(x[(((long long)img * h + iy) * wd + ix) * 3 + j % 3]);
      xs[rr][j] = v;
    }
    __syncthreads();
    for (int px = warp; px < npx; px += 8) {
      const float xv = xs[r][(2 * px + s) * 3 + c];
      const float4* g = reinterpret_cast<const float4*>(&dys[px][0]);
#pragma unroll
      for (int v = 0; v < K / 4; ++v) {
        const float4 q = g[v];
        acc[v * 4] = fmaf(xv, q.x, acc[v * 4]);
        acc[v * 4 + 1] = fmaf(xv, q.y, acc[v * 4 + 1]);
        acc[v * 4 + 2] = fmaf(xv, q.z, acc[v * 4 + 2]);
        acc[v * 4 + 3] = fmaf(xv, q.w, acc[v * 4 + 3]);
      }
    }
  }
  __syncthreads();
  if (lane < 27) {
#pragma unroll
    for (int k = 0; k < K; ++k) atomicAdd(&red[lane][k], acc[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * K; i += blockDim.x) {
    const int k = i / 27, tpi = i % 27;
    atomicAdd(dw + i, red[tpi][k]);
  }
}

// Faster variant for even output widths: lane = OUTPUT CHANNEL (27 accumulators per lane), a block stages the three input rows
// of one output row in shared memory as fp32 (front-padded by 3 floats so that the patch of an even output pixel starts on a
// 16-byte boundary) and every warp walks pixel PAIRS: 12 broadcast LDS.128 + 2 coalesced dy loads feed 54 FFMAs per lane, so the
// FMA pipe, not the shared-memory port, bounds the loop (the (tap, c)-per-lane kernel above re-stages 64-pixel tiles and spends
// most of its time there: 7 TFLOP/s).
template <int K>
__global__ void __launch_bounds__(256) stem_conv_wgrad_rows_kernel(const __nv_bfloat16* __restrict__ x,
                                                                   const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                                                                   int n, int h, int wd, int oh, int ow) {
  extern __shared__ __align__(16) float xs_dyn[];
  const int row_len = ((3 + 3 * wd + 16) + 3) & ~3;  // floats per staged input row
  float* xs = xs_dyn;                                 // [3][row_len]
  __shared__ float red[27][K];
  for (int i = threadIdx.x; i < 27 * K; i += blockDim.x) red[i / K][i % K] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int KH = K / 32;
  float acc[KH][27];
#pragma unroll
  for (int a = 0; a < KH; ++a)
#pragma unroll
    for (int j = 0; j < 27; ++j) acc[a][j] = 0.f;
  const long long rows = (long long)n * oh;
  const int pairs = ow / 2;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int oy = (int)(row % oh), img = (int)(row / oh);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * row_len; i += blockDim.x) {
      const int rr = i / row_len, j = i - rr * row_len;
      const int iy = 2 * oy - 1 + rr;
      const int e = j - 3;  // element index within the (wd * 3)-long input row; < 0 is the left zero padding
      float v = 0.f;
      if (iy >= 0 && iy < h && e >= 0 && e < 3 * wd) v = __bfloat162float(x[((long long)img * h + iy) * wd * 3 + e]);
      xs[i] = v;
    }
    __syncthreads();
    const __nv_bfloat16* dyr = dy + row * ow * K;
    for (int pp = warp; pp < pairs; pp += 8) {
      float f[3][16];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const float4* src = reinterpret_cast<const float4*>(xs + rr * row_len + 12 * pp);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = src[q];
          f[rr][4 * q] = t.x; f[rr][4 * q + 1] = t.y; f[rr][4 * q + 2] = t.z; f[rr][4 * q + 3] = t.w;
        }
      }
#pragma unroll
      for (int a = 0; a < KH; ++a) {
        const float d0 = __bfloat162float(dyr[(long long)(2 * pp) * K + a * 32 + lane]);
        const float d1 = __bfloat162float(dyr[(long long)(2 * pp + 1) * K + a * 32 + lane]);
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int j = 0; j < 9; ++j) acc[a][rr * 9 + j] = fmaf(d0, f[rr][j], fmaf(d1, f[rr][6 + j], acc[a][rr * 9 + j]));
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < KH; ++a)
#pragma unroll
    for (int j = 0; j < 27; ++j) atomicAdd(&red[j][a * 32 + lane], acc[a][j]);
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * K; i += blockDim.x) {
    const int k = i / 27, tpi = i % 27;
    atomicAdd(dw + i, red[tpi][k]);
  }
}

// ---- fully connected on [n][c] vectors (n <= 32): y[n][k] = b[k] + sum_c x[n][c] w[k][c]; one warp per output channel ------------
constexpr int kFcRows = 8;
__global__ void __launch_bounds__(256) fc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ b, float* __restrict__ y, int n, int c, int k) {
  const int lane = threadIdx.x & 31;
  const int ko = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ko >= k) return;
  const float4* wr = reinterpret_cast<const float4*>(w + (long long)ko * c);
  const int c4 = c >> 2;
  for (int n0 = 0; n0 < n; n0 += kFcRows) {
    float acc[kFcRows];
#pragma unroll
    for (int i = 0; i < kFcRows; ++i) acc[i] = 0.f;
    for (int j = lane; j < c4; j += 32) {
      const float4 wv = __ldg(wr + j);
#pragma unroll
      for (int i = 0; i < kFcRows; ++i) {
        if (n0 + i < n) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + i) * c) + j);
          acc[i] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[i]))));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kFcRows; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0 && n0 + i < n) y[(long long)(n0 + i) * k + ko] = v + (b ? b[ko] : 0.f);
    }
  }
}

// dw[k][c] = sum_n dy[n][k] x[n][c] (written), db[k] = sum_n dy[n][k] (written)
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ dw, float* __restrict__ db, int n, int c, int k) {
  const int c4 = c >> 2;
  const long long total = (long long)k * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ko = (int)(i / c4), j = (int)(i % c4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float bs = 0.f;
    for (int r = 0; r < n; ++r) {
      const float g = __ldg(dy + (long long)r * k + ko);
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)r * c) + j);
      acc.x = fmaf(g, xv.x, acc.x);
      acc.y = fmaf(g, xv.y, acc.y);
      acc.z = fmaf(g, xv.z, acc.z);
      acc.w = fmaf(g, xv.w, acc.w);
      bs += g;
    }
    reinterpret_cast<float4*>(dw + (long long)ko * c)[j] = acc;
    if (j == 0 && db) db[ko] = bs;
  }
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_stem_conv_fwd(const void* x, const float* w, void* y, int32_t n, int32_t h, int32_t wd, int32_t k,
                                 void* stream) {
  XV2_REQUIRE(x && w && y && n > 0 && h > 0 && wd > 0, "stem_conv_fwd: bad argument");
  if (k != 32 && k != 64) {
    set_error("stem_conv_fwd: %d output channels not served", k);
    return XV2_EUNSUPPORTED;
  }
  const int oh = (h - 1) / 2 + 1, ow = (wd - 1) / 2 + 1;
  const long long total = (long long)n * oh * ow;
  const int blocks = (int)std::min<long long>(cdiv(total, 256), 16LL * kNumSMs);
  if (k == 32)
    stem_conv_fwd_kernel<32><<<blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, n, h, wd, oh, ow);
  else
    stem_conv_fwd_kernel<64><<<blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, n, h, wd, oh, ow);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_stem_conv_wgrad(const void* x, const void* dy, float* dw, int32_t n, int32_t h, int32_t wd, int32_t k,
                                   void* stream) {
  XV2_REQUIRE(x && dy && dw && n > 0 && h > 0 && wd > 0, "stem_conv_wgrad: bad argument");
  if (k != 32 && k != 64) {
    set_error("stem_conv_wgrad: %d output channels not served", k);
    return XV2_EUNSUPPORTED;
  }
  const int oh = (h - 1) / 2 + 1, ow = (wd - 1) / 2 + 1;
  if (ow % 2 == 0 && wd % 2 == 0 && wd <= 4096) {
    const size_t smem = 3 * (size_t)(((3 + 3 * wd + 16) + 3) & ~3) * sizeof(float);
    long long blocks2 = (long long)n * oh;
    if (blocks2 > 4LL * kNumSMs) blocks2 = 4LL * kNumSMs;
    cudaError_t e = cudaSuccess;
    if (k == 32) {
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(stem_conv_wgrad_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
        stem_conv_wgrad_rows_kernel<32><<<(unsigned)blocks2, 256, smem, as_stream(stream)>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, n, h, wd, oh, ow);
    } else {
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(stem_conv_wgrad_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
        stem_conv_wgrad_rows_kernel<64><<<(unsigned)blocks2, 256, smem, as_stream(stream)>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, n, h, wd, oh, ow);
    }
    if (e != cudaSuccess) {
      set_error("stem_conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return XV2_ECUDA;
    }
    XV2_LAUNCH_CHECK();
    return XV2_OK;
  }
  const int blocks = 6 * kNumSMs;
  if (k == 32)
    stem_conv_wgrad_kernel<32><<<blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, n, h,
                                                                      wd, oh, ow);
  else
    stem_conv_wgrad_kernel<64><<<blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, n, h,
                                                                      wd, oh, ow);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_fc_fwd(const float* x, const float* w, const float* b, float* y, int32_t n, int32_t c, int32_t k,
                          void* stream) {
  XV2_REQUIRE(x && w && y && n > 0 && c > 0 && k > 0 && c % 4 == 0, "fc_fwd: bad argument (c must be a multiple of 4)");
  fc_fwd_kernel<<<(k + 7) / 8, 256, 0, as_stream(stream)>>>(x, w, b, y, n, c, k);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_fc_wgrad(const float* x, const float* dy, float* dw, float* db, int32_t n, int32_t c, int32_t k,
                            void* stream) {
  XV2_REQUIRE(x && dy && dw && n > 0 && c > 0 && k > 0 && c % 4 == 0, "fc_wgrad: bad argument (c must be a multiple of 4)");
  const long long total = (long long)k * (c / 4);
  const int blocks = (int)std::min<long long>(cdiv(total, 256), 8LL * kNumSMs);
  fc_wgrad_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, dy, dw, db, n, c, k);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

// ====================================================================================================================
// Split-attention FC chain (SplAtConv2d tail on [n][C] vectors, n <= 32), fused to 2 forward + 3 backward launches:
//   forward  A: z1 = fc1(gap) -> BatchNorm over the n samples (train: batch statistics + running update) -> relu -> a1
//            B: z2 = fc2(a1) -> r-softmax over the radix pair (k, k + C) -> att
//   backward 1: dz2 = rsoftmax'(att, datt); dw2 = dz2^T a1; db2
//            2: da1 = dz2 w2 -> relu' -> BatchNorm backward over n -> dz1; dw1 = dz1^T gap; db1, dgamma, dbeta
//            3: dgap = dz1 w1
// One warp per output channel; lanes stride over the reduction axis with float4 loads; all n rows per pass.
// ====================================================================================================================
namespace xv2 {
constexpr int kSaMaxN = 32;

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}
// out[i] (all lanes) = sum_c x[i][c] * w[c] for i < n; c % 4 == 0
__device__ __forceinline__ void warp_matvec(const float* __restrict__ x, const float* __restrict__ w, int n, int c, int lane,
                                            float* out) {
  const int c4 = c >> 2;
  const float4* wr = reinterpret_cast<const float4*>(w);
  for (int n0 = 0; n0 < n; n0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j = lane; j < c4; j += 32) {
      const float4 wv = __ldg(wr + j);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n0 + i < n) acc[i] = dot4(wv, __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + i) * c) + j), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (n0 + i < n) out[n0 + i] = warp_sum(acc[i]);
  }
}

__global__ void __launch_bounds__(256) sa_fc1_bn_relu_kernel(const float* __restrict__ gap, const float* __restrict__ w1,
                                                             const float* __restrict__ b1, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ rmean,
                                                             float* __restrict__ rvar, float momentum, float eps, int training,
                                                             float* __restrict__ z1, float* __restrict__ a1,
                                                             float* __restrict__ coef, int n, int c, int inter) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= inter) return;
  float z[kSaMaxN];
  warp_matvec(gap, w1 + (long long)j * c, n, c, lane, z);
  const float bj = b1 ? b1[j] : 0.f;
  float mean, var;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < n; ++i) {
      z[i] += bj;
      s += z[i];
      q += (double)z[i] * z[i];
    }
    const double mu = s / n;
    double v = q / n - mu * mu;
    if (v < 0.0) v = 0.0;
    mean = (float)mu;
    var = (float)v;
    if (lane == 0 && rmean) {
      const double unbiased = n > 1 ? v * n / (n - 1.0) : v;
      rmean[j] = (1.f - momentum) * rmean[j] + momentum * mean;
      rvar[j] = (1.f - momentum) * rvar[j] + momentum * (float)unbiased;
    }
  } else {
    for (int i = 0; i < n; ++i) z[i] += bj;
    mean = rmean[j];
    var = rvar[j];
  }
  const float is = training ? (float)(1.0 / sqrt((double)var + (double)eps)) : 1.0f / sqrtf(var + eps);
  const float g = gamma ? gamma[j] : 1.f, b = beta ? beta[j] : 0.f;
  const float sc = g * is, sh = b - mean * g * is;
  if (lane == 0) {
    coef[j] = mean;
    coef[inter + j] = is;
    coef[2 * inter + j] = sc;
    coef[3 * inter + j] = sh;
  }
  for (int i = lane; i < n; i += 32) {
    z1[(long long)i * inter + j] = z[i];
    const float u = fmaf(z[i], sc, sh);
    a1[(long long)i * inter + j] = u > 0.f ? u : 0.f;
  }
}

__global__ void __launch_bounds__(256) sa_fc2_rsoftmax_kernel(const float* __restrict__ a1, const float* __restrict__ w2,
                                                              const float* __restrict__ b2, float* __restrict__ att, int n,
                                                              int c, int inter) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= c) return;
  float z0[kSaMaxN], z1[kSaMaxN];
  warp_matvec(a1, w2 + (long long)k * inter, n, inter, lane, z0);
  warp_matvec(a1, w2 + (long long)(k + c) * inter, n, inter, lane, z1);
  const float b0 = b2 ? b2[k] : 0.f, bb1 = b2 ? b2[k + c] : 0.f;
  for (int i = lane; i < n; i += 32) {
    const float a = z0[i] + b0, b = z1[i] + bb1;
    const float m = fmaxf(a, b);
    const float ea = expf(a - m), eb = expf(b - m);
    const float inv = 1.f / (ea + eb);
    att[(long long)i * 2 * c + k] = ea * inv;
    att[(long long)i * 2 * c + c + k] = eb * inv;
  }
}

// backward 1: warp per radix pair k: dz2[:, k], dz2[:, k + C]; dw2 rows; db2
// Optional inputs of the bn0-fused split attention (splat_fused.cu): with `part` (fp64 [4][n][2c] per-image partial sums A1 | A2 |
// M1 | M2) the first FC kernel derives datt = scale0 * A2 + shift0 * A1 itself and the last one finishes the two bn0 reductions
// (red fp64 [2][2c]) from the dgap column it has just produced -- the arithmetic of splat_bn_bwd_datt_kernel / splat_bn_bwd_red_kernel.
struct SaFused {
  const double* part;
  const float *scale0, *shift0, *mean0, *invstd0;
  double* red;
  float inv_hw;
  int accumulate;  // parameter gradients are ADDED to their destinations (the parameters' own slots in the flat gradient buffer)
};

__global__ void __launch_bounds__(256) sa_bwd_fc2_kernel(const float* __restrict__ att, const float* __restrict__ datt,
                                                         const float* __restrict__ a1, float* __restrict__ dz2,
                                                         float* __restrict__ dw2, float* __restrict__ db2, int n, int c,
                                                         int inter, SaFused fz) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= c) return;
  float d0[kSaMaxN], d1[kSaMaxN];
  float s0 = 0.f, s1 = 0.f;
  for (int i = 0; i < n; ++i) {
    const long long i0 = (long long)i * 2 * c + k, i1 = i0 + c;
    const float p0 = att[i0], p1 = att[i1];
    float g0, g1;
    if (fz.part != nullptr) {
      const long long plane = (long long)n * 2 * c;
      g0 = (float)((double)fz.scale0[k] * fz.part[plane + i0] + (double)fz.shift0[k] * fz.part[i0]);
      g1 = (float)((double)fz.scale0[k + c] * fz.part[plane + i1] + (double)fz.shift0[k + c] * fz.part[i1]);
    } else {
      g0 = datt[i0];
      g1 = datt[i1];
    }
    const float dot = p0 * g0 + p1 * g1;
    d0[i] = p0 * (g0 - dot);
    d1[i] = p1 * (g1 - dot);
    s0 += d0[i];
    s1 += d1[i];
  }
  if (lane == 0) {
    db2[k] = (fz.accumulate ? db2[k] : 0.f) + s0;
    db2[k + c] = (fz.accumulate ? db2[k + c] : 0.f) + s1;
  }
  for (int i = lane; i < n; i += 32) {
    dz2[(long long)i * 2 * c + k] = d0[i];
    dz2[(long long)i * 2 * c + c + k] = d1[i];
  }
  for (int j = lane; j < inter; j += 32) {
    float w0 = 0.f, w1 = 0.f;
    for (int i = 0; i < n; ++i) {
      const float a = a1[(long long)i * inter + j];
      w0 = fmaf(d0[i], a, w0);
      w1 = fmaf(d1[i], a, w1);
    }
    dw2[(long long)k * inter + j] = (fz.accumulate ? dw2[(long long)k * inter + j] : 0.f) + w0;
    dw2[(long long)(k + c) * inter + j] = (fz.accumulate ? dw2[(long long)(k + c) * inter + j] : 0.f) + w1;
  }
}

// backward 2: warp per fc1 output channel j.  w2t is fc2's weight transposed: [inter][2c]
__global__ void __launch_bounds__(256) sa_bwd_bn_fc1_kernel(const float* __restrict__ dz2, const float* __restrict__ w2t,
                                                            const float* __restrict__ z1, const float* __restrict__ coef,
                                                            const float* __restrict__ gamma, const float* __restrict__ gap,
                                                            int training, float* __restrict__ dz1, float* __restrict__ dw1,
                                                            float* __restrict__ db1, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int n, int c, int inter, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= inter) return;
  float da[kSaMaxN];
  warp_matvec(dz2, w2t + (long long)j * 2 * c, n, 2 * c, lane, da);
  const float mean = coef[j], is = coef[inter + j], sc = coef[2 * inter + j], sh = coef[3 * inter + j];
  const float g = gamma ? gamma[j] : 1.f;
  float r1 = 0.f, r2 = 0.f;
  float xh[kSaMaxN];
  for (int i = 0; i < n; ++i) {
    const float z = z1[(long long)i * inter + j];
    const float u = fmaf(z, sc, sh);
    da[i] = u > 0.f ? da[i] : 0.f;  // relu'
    xh[i] = (z - mean) * is;
    r1 += da[i];
    r2 = fmaf(da[i], xh[i], r2);
  }
  float d[kSaMaxN];
  float bsum = 0.f;
  const float inv_n = 1.f / (float)n;
  for (int i = 0; i < n; ++i) {
    d[i] = training ? g * is * (da[i] - r1 * inv_n - xh[i] * r2 * inv_n) : da[i] * sc;
    bsum += d[i];
  }
  if (lane == 0) {
    dbeta[j] = (accumulate ? dbeta[j] : 0.f) + r1;
    dgamma[j] = (accumulate ? dgamma[j] : 0.f) + r2;
    db1[j] = (accumulate ? db1[j] : 0.f) + bsum;
  }
  for (int i = lane; i < n; i += 32) dz1[(long long)i * inter + j] = d[i];
  for (int cc = lane; cc < c; cc += 32) {
    float w = 0.f;
    for (int i = 0; i < n; ++i) w = fmaf(d[i], gap[(long long)i * c + cc], w);
    dw1[(long long)j * c + cc] = (accumulate ? dw1[(long long)j * c + cc] : 0.f) + w;
  }
}

// backward 3: dgap[n][c] = sum_j dz1[n][j] w1t[c][j]  (w1t = fc1's weight transposed: [c][inter]); warp per channel c
__global__ void __launch_bounds__(256) sa_bwd_gap_kernel(const float* __restrict__ dz1, const float* __restrict__ w1t,
                                                         float* __restrict__ dgap, int n, int c, int inter,
                                                         const float* __restrict__ att, SaFused fz) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int cc = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cc >= c) return;
  float o[kSaMaxN];
  warp_matvec(dz1, w1t + (long long)cc * inter, n, inter, lane, o);
  for (int i = lane; i < n; i += 32) dgap[(long long)i * c + cc] = o[i];
  if (fz.red != nullptr && lane < 2) {
    // bn0 reductions of the radix halves ch = cc (lane 0) and cc + c (lane 1): both share this dgap column
    const int c2 = 2 * c, ch = cc + lane * c;
    const long long plane = (long long)n * c2;
    const double mu = fz.mean0[ch], is = fz.invstd0[ch];
    double r1 = 0.0, r2 = 0.0;
    for (int nb = 0; nb < n; ++nb) {
      const long long i = (long long)nb * c2 + ch;
      const double a = att[i], g = (double)o[nb] * fz.inv_hw;
      const double A1 = fz.part[i], A2 = fz.part[plane + i], M1 = fz.part[2 * plane + i], M2 = fz.part[3 * plane + i];
      r1 += a * A1 + g * M1;
      r2 += is * (a * (A2 - mu * A1) + g * (M2 - mu * M1));
    }
    fz.red[ch] = r1;
    fz.red[c2 + ch] = r2;
  }
}
}  // namespace xv2

extern "C" int xv2_splat_fc_fwd(const float* gap, const float* w1, const float* b1, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float momentum, float eps, int32_t training,
                                const float* w2, const float* b2, float* z1, float* a1, float* coef, float* att, int32_t n,
                                int32_t c, int32_t inter, void* stream) {
  XV2_REQUIRE(gap && w1 && w2 && z1 && a1 && coef && att, "splat_fc_fwd: null argument");
  XV2_REQUIRE(n >= 1 && n <= kSaMaxN && c % 4 == 0 && inter % 4 == 0, "splat_fc_fwd: n <= 32, channels multiples of 4 (n %d c %d inter %d)",
              n, c, inter);
  XV2_REQUIRE(training == 0 || n > 1, "Expected more than 1 value per channel when training");
  XV2_REQUIRE(training != 0 || (running_mean && running_var), "splat_fc_fwd: eval mode needs running statistics");
  launch_pdl(sa_fc1_bn_relu_kernel, dim3((inter + 7) / 8), dim3(256), 0, as_stream(stream), gap, w1, b1, gamma, beta, running_mean, running_var, momentum,
                                                                        eps, training, z1, a1, coef, n, c, inter);
  launch_pdl(sa_fc2_rsoftmax_kernel, dim3((c + 7) / 8), dim3(256), 0, as_stream(stream), a1, w2, b2, att, n, c, inter);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_fc_bwd(const float* att, const float* datt, const float* a1, const float* z1, const float* coef,
                                const float* gamma, const float* gap, const float* w2t, const float* w1t, int32_t training,
                                float* dz2, float* dz1, float* dw2, float* db2, float* dw1, float* db1, float* dgamma,
                                float* dbeta, float* dgap, int32_t n, int32_t c, int32_t inter, void* stream) {
  XV2_REQUIRE(att && datt && a1 && z1 && coef && gap && w2t && w1t && dz2 && dz1 && dw2 && db2 && dw1 && db1 && dgamma && dbeta &&
                  dgap,
              "splat_fc_bwd: null argument");
  XV2_REQUIRE(n >= 1 && n <= kSaMaxN && c % 4 == 0 && inter % 4 == 0, "splat_fc_bwd: n <= 32, channels multiples of 4");
  SaFused fz;
  memset(&fz, 0, sizeof(fz));
  launch_pdl(sa_bwd_fc2_kernel, dim3((c + 7) / 8), dim3(256), 0, as_stream(stream), att, datt, a1, dz2, dw2, db2, n, c, inter, fz);
  launch_pdl(sa_bwd_bn_fc1_kernel, dim3((inter + 7) / 8), dim3(256), 0, as_stream(stream), dz2, w2t, z1, coef, gamma, gap, training, dz1, dw1, db1, dgamma,
                                                                       dbeta, n, c, inter, 0);
  launch_pdl(sa_bwd_gap_kernel, dim3((c + 7) / 8), dim3(256), 0, as_stream(stream), dz1, w1t, dgap, n, c, inter, att, fz);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_fc_bwd_fused(const float* att, const double* part, const float* scale0, const float* shift0,
                                      const float* mean0, const float* invstd0, int64_t hw, const float* a1, const float* z1,
                                      const float* coef, const float* gamma, const float* gap, const float* w2t, const float* w1t,
                                      int32_t training, float* dz2, float* dz1, float* dw2, float* db2, float* dw1, float* db1,
                                      float* dgamma, float* dbeta, float* dgap, double* red, int32_t accumulate, int32_t n, int32_t c,
                                      int32_t inter, void* stream) {
  XV2_REQUIRE(att && part && scale0 && shift0 && mean0 && invstd0 && a1 && z1 && coef && gap && w2t && w1t && dz2 && dz1 && dw2 &&
                  db2 && dw1 && db1 && dgamma && dbeta && dgap && red && hw > 0,
              "splat_fc_bwd_fused: null argument");
  XV2_REQUIRE(n >= 1 && n <= kSaMaxN && c % 4 == 0 && inter % 4 == 0, "splat_fc_bwd_fused: n <= 32, channels multiples of 4");
  SaFused fz;
  fz.part = part;
  fz.scale0 = scale0;
  fz.shift0 = shift0;
  fz.mean0 = mean0;
  fz.invstd0 = invstd0;
  fz.red = red;
  fz.inv_hw = 1.0f / (float)hw;
  fz.accumulate = accumulate ? 1 : 0;
  launch_pdl(sa_bwd_fc2_kernel, dim3((c + 7) / 8), dim3(256), 0, as_stream(stream), att, (const float*)nullptr, a1, dz2, dw2, db2, n,
             c, inter, fz);
  launch_pdl(sa_bwd_bn_fc1_kernel, dim3((inter + 7) / 8), dim3(256), 0, as_stream(stream), dz2, w2t, z1, coef, gamma, gap, training,
             dz1, dw1, db1, dgamma, dbeta, n, c, inter, accumulate);
  launch_pdl(sa_bwd_gap_kernel, dim3((c + 7) / 8), dim3(256), 0, as_stream(stream), dz1, w1t, dgap, n, c, inter, att, fz);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
