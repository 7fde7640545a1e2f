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
  __shared__ __align__(16) __nv_bfloat16 dys[TP][K];
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
      *reinterpret_cast<uint4*>(&dys[px][v * 8]) = q;
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
      const uint4* g = reinterpret_cast<const uint4*>(&dys[px][0]);
#pragma unroll
      for (int v = 0; v < K / 8; ++v) {
        const uint4 q = g[v];
        const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[v * 8 + 2 * i] = fmaf(xv, __uint_as_float(wds[i] << 16), acc[v * 8 + 2 * i]);
          acc[v * 8 + 2 * i + 1] = fmaf(xv, __uint_as_float(wds[i] & 0xffff0000u), acc[v * 8 + 2 * i + 1]);
        }
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
