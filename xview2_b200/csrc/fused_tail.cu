// Fused tail of the localisation network: the last decoder ConvLayer's BatchNorm + LeakyReLU and the 1x1 output head
// (layers.py:96-100 followed by layers.py:180-183) WITHOUT materialising the 32-channel full-resolution activation between them.
//
//   forward : z (raw conv output, bf16 NHWC, c = 32 | 64) --read once--> y = bf16(act(scale z + shift)) in registers
//             --> logits[p][k] = sum_c w[k][c] y[c] + b[k]  (fp32)                      HBM: read z, write logits
//   backward: dy[c] = sum_k dl[k] w[k][c] is recomputed from the logit gradient (2-4 floats per pixel) instead of being written
//             and re-read twice; pass 1 reduces the BatchNorm sums (sum du, sum du xhat) AND the head's weight / bias gradients
//             (dW[k][c] = sum_p dl[k] y[c], db[k] = sum_p dl[k]); pass 2 writes dz.                HBM: 2 x read z, write dz
// versus the unfused chain (bn_train_apply, head_fwd, head_bwd, bn_bwd_reduce, bn_bwd_apply): 9 passes over a 537 MB tensor at
// BASELINE config 2 become 4.
//
// Thread mapping (all three kernels): a thread owns ONE 16-byte channel vector (8 bf16 channels) and walks pixels; the C/8
// threads of a pixel are neighbours in a warp, so a warp reads 512 contiguous bytes per load and the per-pixel dot products are
// finished with 2-3 xor-shuffles.  Per-channel constants live in registers.
#include "common.cuh"
#include <cstring>

namespace xv2 {

struct TailParams {
  const __nv_bfloat16* z;
  const float *scale, *shift, *mean, *invstd, *gamma;
  const float *hw, *hb;      // head weights [ncls][c], bias [ncls] (bias may be null)
  const float* dl;           // logit gradient [pixels][ncls]
  float* logits;             // [pixels][ncls]
  __nv_bfloat16* dz;
  double* red;               // [2c] BatchNorm reductions (accumulated)
  float *dhw, *dhb;          // head gradients (accumulated)
  float *dgamma, *dbeta;     // written / accumulated by block 0 of the apply pass
  long long pixels;
  int c, act, accumulate;
  float inv_n;
};

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ uint4 ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <int NCLS> struct Grad {
  float g[NCLS];
};
template <int NCLS> __device__ __forceinline__ Grad<NCLS> ldgrad(const float* dl, long long px) {
  Grad<NCLS> r;
  if constexpr (NCLS == 2) {
    const float2 v = *reinterpret_cast<const float2*>(dl + px * 2);
    r.g[0] = v.x;
    r.g[1] = v.y;
  } else if constexpr (NCLS == 4) {
    const float4 v = *reinterpret_cast<const float4*>(dl + px * 4);
    r.g[0] = v.x; r.g[1] = v.y; r.g[2] = v.z; r.g[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < NCLS; ++k) r.g[k] = dl[px * NCLS + k];
  }
  return r;
}
constexpr int kTailUnroll = 4;  // independent loads in flight per thread: these kernels are bound by memory-level parallelism

template <int CV, int NCLS>
__global__ void __launch_bounds__(256) tail_fwd_kernel(const TailParams p) {
  const int cv = threadIdx.x & (CV - 1), row = threadIdx.x / CV;
  constexpr int ROWS = 256 / CV;
  float sc[8], sh[8], w[NCLS][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = p.scale[cv * 8 + i];
    sh[i] = p.shift[cv * 8 + i];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) w[k][i] = p.hw[k * p.c + cv * 8 + i];
  }
  float b[NCLS];
#pragma unroll
  for (int k = 0; k < NCLS; ++k) b[k] = p.hb ? p.hb[k] : 0.f;
  const long long stride = (long long)gridDim.x * ROWS;
  auto emit = [&](long long px, const uint4& r, bool valid) {
    float f[8], acc[NCLS];
    unpack8(r, f);
#pragma unroll
    for (int k = 0; k < NCLS; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = bf16_round(apply_act(fmaf(f[i], sc[i], sh[i]), p.act));  // the activation the unfused path would store
#pragma unroll
      for (int k = 0; k < NCLS; ++k) acc[k] = fmaf(w[k][i], y, acc[k]);
    }
#pragma unroll
    for (int o = CV / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < NCLS; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (cv == 0 && valid) {
#pragma unroll
      for (int k = 0; k < NCLS; ++k) p.logits[px * NCLS + k] = acc[k] + b[k];
    }
  };
  // every thread of a block runs the same number of iterations (the bound is padded to whole blocks of rows) and every lane
  // takes part in the shuffles; rows past the end compute on zeros and skip the store
  const long long rows_total = (p.pixels + ROWS - 1) / ROWS * ROWS;
  long long px = (long long)blockIdx.x * ROWS + row;
  for (; px + (kTailUnroll - 1) * stride < rows_total; px += kTailUnroll * stride) {
    uint4 r[kTailUnroll];
#pragma unroll
    for (int u = 0; u < kTailUnroll; ++u) {
      const long long q = px + u * stride;
      r[u] = q < p.pixels ? ldraw(p.z + q * p.c + cv * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kTailUnroll; ++u) emit(px + u * stride, r[u], px + u * stride < p.pixels);
  }
  for (; px < rows_total; px += stride)
    emit(px, px < p.pixels ? ldraw(p.z + px * p.c + cv * 8) : make_uint4(0, 0, 0, 0), px < p.pixels);
}

template <int CV, int NCLS>
__global__ void __launch_bounds__(256) tail_bwd_reduce_kernel(const TailParams p) {
  const int cv = threadIdx.x & (CV - 1), row = threadIdx.x / CV;
  constexpr int ROWS = 256 / CV;
  float sc[8], sh[8], mu[8], is[8], w[NCLS][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = cv * 8 + i;
    sc[i] = p.scale[ch];
    sh[i] = p.shift[ch];
    mu[i] = p.mean[ch];
    is[i] = p.invstd[ch];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) w[k][i] = p.hw[k * p.c + ch];
  }
  float s1[8], s2[8], gw[NCLS][8], gb[NCLS];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[i] = s2[i] = 0.f;
#pragma unroll
    for (int k = 0; k < NCLS; ++k) gw[k][i] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < NCLS; ++k) gb[k] = 0.f;
  const long long stride = (long long)gridDim.x * ROWS;
  auto accum = [&](const uint4& r, const Grad<NCLS>& gr) {
    float f[8];
    unpack8(r, f);
#pragma unroll
    for (int k = 0; k < NCLS; ++k) gb[k] += gr.g[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float u = fmaf(f[i], sc[i], sh[i]);
      const float y = bf16_round(apply_act(u, p.act));
      float dy = 0.f;
#pragma unroll
      for (int k = 0; k < NCLS; ++k) {
        dy = fmaf(gr.g[k], w[k][i], dy);
        gw[k][i] = fmaf(gr.g[k], y, gw[k][i]);
      }
      const float du = dy * act_grad(u, p.act);
      s1[i] += du;
      s2[i] = fmaf(du, (f[i] - mu[i]) * is[i], s2[i]);
    }
  };
  long long px = (long long)blockIdx.x * ROWS + row;
  for (; px + (kTailUnroll - 1) * stride < p.pixels; px += kTailUnroll * stride) {
    uint4 r[kTailUnroll];
    Grad<NCLS> gr[kTailUnroll];
#pragma unroll
    for (int u = 0; u < kTailUnroll; ++u) {
      r[u] = ldraw(p.z + (px + u * stride) * p.c + cv * 8);
      gr[u] = ldgrad<NCLS>(p.dl, px + u * stride);
    }
#pragma unroll
    for (int u = 0; u < kTailUnroll; ++u) accum(r[u], gr[u]);
  }
  for (; px < p.pixels; px += stride) accum(ldraw(p.z + px * p.c + cv * 8), ldgrad<NCLS>(p.dl, px));
  // block reduction over the pixel rows that share a channel vector (CV < 32: xor-shuffles inside the warp first, then all
  // threads finish the 8 per-warp partials of every (vector, element) pair), one atomic per channel per block
  __shared__ float sm[8 * CV][9];
  auto reduce8 = [&](float* vals, auto&& sink) {
    __syncthreads();
#pragma unroll
    for (int off = CV; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < CV) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[warp * CV + lane][i] = vals[i];
    }
    __syncthreads();
    if (threadIdx.x < CV * 8) {
      const int v = threadIdx.x >> 3, i = threadIdx.x & 7;
      double t = 0.0;
#pragma unroll
      for (int wp = 0; wp < 8; ++wp) t += (double)sm[wp * CV + v][i];
      sink(v, i, t);
    }
  };
  reduce8(s1, [&](int v, int i, double t) { atomicAdd(&p.red[v * 8 + i], t); });
  reduce8(s2, [&](int v, int i, double t) { atomicAdd(&p.red[p.c + v * 8 + i], t); });
#pragma unroll
  for (int k = 0; k < NCLS; ++k) reduce8(gw[k], [&](int v, int i, double t) { atomicAdd(&p.dhw[k * p.c + v * 8 + i], (float)t); });
  // bias gradient: only the cv == 0 thread of each pixel contributes
  float gbv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gbv[i] = (i < NCLS && cv == 0) ? gb[i < NCLS ? i : 0] : 0.f;
  if (p.dhb) reduce8(gbv, [&](int v, int i, double t) {
    if (v == 0 && i < NCLS) atomicAdd(&p.dhb[i], (float)t);
  });
}

template <int CV, int NCLS>
__global__ void __launch_bounds__(256) tail_bwd_apply_kernel(const TailParams p) {
  if (blockIdx.x == 0 && p.dgamma != nullptr) {
    for (int i = threadIdx.x; i < p.c; i += blockDim.x) {
      p.dbeta[i] = (p.accumulate ? p.dbeta[i] : 0.f) + (float)p.red[i];
      p.dgamma[i] = (p.accumulate ? p.dgamma[i] : 0.f) + (float)p.red[p.c + i];
    }
  }
  const int cv = threadIdx.x & (CV - 1), row = threadIdx.x / CV;
  constexpr int ROWS = 256 / CV;
  float sc[8], sh[8], mu[8], is[8], k0[8], k1[8], k2[8], w[NCLS][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = cv * 8 + i;
    sc[i] = p.scale[ch];
    sh[i] = p.shift[ch];
    mu[i] = p.mean[ch];
    is[i] = p.invstd[ch];
    k0[i] = p.gamma[ch] * is[i];
    k1[i] = (float)p.red[ch] * p.inv_n;
    k2[i] = (float)p.red[p.c + ch] * p.inv_n;
#pragma unroll
    for (int k = 0; k < NCLS; ++k) w[k][i] = p.hw[k * p.c + ch];
  }
  const long long stride = (long long)gridDim.x * ROWS;
  auto emit = [&](long long q, const uint4& r, const Grad<NCLS>& gr) {
    float f[8], o[8];
    unpack8(r, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float u = fmaf(f[i], sc[i], sh[i]);
      float dy = 0.f;
#pragma unroll
      for (int k = 0; k < NCLS; ++k) dy = fmaf(gr.g[k], w[k][i], dy);
      const float du = dy * act_grad(u, p.act);
      const float xh = (f[i] - mu[i]) * is[i];
      o[i] = k0[i] * (du - k1[i] - xh * k2[i]);
    }
    Vec<__nv_bfloat16> ov;
    ov.pack(o);
    ov.store(p.dz + q * p.c + cv * 8);
  };
  long long px = (long long)blockIdx.x * ROWS + row;
  for (; px + (kTailUnroll - 1) * stride < p.pixels; px += kTailUnroll * stride) {
    uint4 r[kTailUnroll];
    Grad<NCLS> gr[kTailUnroll];
#pragma unroll
    for (int u = 0; u < kTailUnroll; ++u) {
      r[u] = ldraw(p.z + (px + u * stride) * p.c + cv * 8);
      gr[u] = ldgrad<NCLS>(p.dl, px + u * stride);
    }
#pragma unroll
    for (int u = 0; u < kTailUnroll; ++u) emit(px + u * stride, r[u], gr[u]);
  }
  for (; px < p.pixels; px += stride) emit(px, ldraw(p.z + px * p.c + cv * 8), ldgrad<NCLS>(p.dl, px));
}

static int tail_grid(long long pixels, int rows_per_block) {
  long long b = (pixels + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define XV2_TAIL_DISPATCH(KERNEL, p, st)                                                         \
  do {                                                                                           \
    const int cvn = (p).c / 8;                                                                   \
    const int grid = tail_grid((p).pixels, 256 / cvn);                                           \
    if (cvn == 4 && ncls == 1) KERNEL<4, 1><<<grid, 256, 0, st>>>(p);                            \
    else if (cvn == 4 && ncls == 2) KERNEL<4, 2><<<grid, 256, 0, st>>>(p);                       \
    else if (cvn == 4 && ncls == 3) KERNEL<4, 3><<<grid, 256, 0, st>>>(p);                       \
    else if (cvn == 4 && ncls == 4) KERNEL<4, 4><<<grid, 256, 0, st>>>(p);                       \
    else if (cvn == 8 && ncls == 1) KERNEL<8, 1><<<grid, 256, 0, st>>>(p);                       \
    else if (cvn == 8 && ncls == 2) KERNEL<8, 2><<<grid, 256, 0, st>>>(p);                       \
    else if (cvn == 8 && ncls == 3) KERNEL<8, 3><<<grid, 256, 0, st>>>(p);                       \
    else KERNEL<8, 4><<<grid, 256, 0, st>>>(p);                                                  \
  } while (0)

static int tail_check(int64_t pixels, int32_t c, int32_t ncls, const char* who) {
  if (pixels <= 0 || !(c == 32 || c == 64) || ncls < 1 || ncls > 4) {
    set_error("%s: serves bf16 tensors with 32 or 64 channels and 1-4 classes (got c %d, ncls %d)", who, c, ncls);
    return XV2_EUNSUPPORTED;
  }
  return XV2_OK;
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_bnact_head_fwd(const void* z, int64_t pixels, int32_t c, const float* scale, const float* shift, int32_t act,
                                  const float* head_w, const float* head_b, int32_t ncls, float* logits, void* stream) {
  XV2_REQUIRE(z && scale && shift && head_w && logits, "bnact_head_fwd: null argument");
  int rc = tail_check(pixels, c, ncls, "bnact_head_fwd");
  if (rc) return rc;
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.z = (const __nv_bfloat16*)z;
  p.scale = scale;
  p.shift = shift;
  p.hw = head_w;
  p.hb = head_b;
  p.logits = logits;
  p.pixels = pixels;
  p.c = c;
  p.act = act;
  XV2_TAIL_DISPATCH(tail_fwd_kernel, p, as_stream(stream));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bnact_head_bwd_reduce(const void* z, const float* dlogits, int64_t pixels, int32_t c, const float* scale,
                                         const float* shift, const float* mean, const float* invstd, int32_t act,
                                         const float* head_w, int32_t ncls, double* red, float* dhead_w, float* dhead_b,
                                         void* stream) {
  XV2_REQUIRE(z && dlogits && scale && shift && mean && invstd && head_w && red && dhead_w, "bnact_head_bwd_reduce: null argument");
  int rc = tail_check(pixels, c, ncls, "bnact_head_bwd_reduce");
  if (rc) return rc;
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.z = (const __nv_bfloat16*)z;
  p.dl = dlogits;
  p.scale = scale;
  p.shift = shift;
  p.mean = mean;
  p.invstd = invstd;
  p.hw = head_w;
  p.red = red;
  p.dhw = dhead_w;
  p.dhb = dhead_b;
  p.pixels = pixels;
  p.c = c;
  p.act = act;
  XV2_TAIL_DISPATCH(tail_bwd_reduce_kernel, p, as_stream(stream));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_bnact_head_bwd_apply(const void* z, const float* dlogits, void* dz, int64_t pixels, int32_t c,
                                        const float* scale, const float* shift, const float* mean, const float* invstd,
                                        const float* gamma, int32_t act, const float* head_w, int32_t ncls, const double* red,
                                        int64_t count, float* dgamma, float* dbeta, int32_t accumulate, void* stream) {
  XV2_REQUIRE(z && dlogits && dz && scale && shift && mean && invstd && gamma && head_w && red, "bnact_head_bwd_apply: null argument");
  int rc = tail_check(pixels, c, ncls, "bnact_head_bwd_apply");
  if (rc) return rc;
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.z = (const __nv_bfloat16*)z;
  p.dl = dlogits;
  p.dz = (__nv_bfloat16*)dz;
  p.scale = scale;
  p.shift = shift;
  p.mean = mean;
  p.invstd = invstd;
  p.gamma = gamma;
  p.hw = head_w;
  p.red = const_cast<double*>(red);
  p.dgamma = dgamma;
  p.dbeta = dbeta;
  p.accumulate = accumulate;
  p.inv_n = 1.0f / (float)(count > 0 ? count : 1);
  p.pixels = pixels;
  p.c = c;
  p.act = act;
  XV2_TAIL_DISPATCH(tail_bwd_apply_kernel, p, as_stream(stream));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
