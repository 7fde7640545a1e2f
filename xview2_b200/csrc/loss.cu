// Dice / Focal / CE(=OHEM as the reference computes it) loss over fp32 NHWC logits and uint8 labels, as ONE reduction
// pass + a scalar finalize + ONE backward pass (the reference takes 6-8 full-resolution ATen passes per term through
// MONAI: loss.py:7-21,78-101, plt.py:69-77).  Also the F1 counters (utils/f1.py:28-42) and the argmax / threshold
// post-process (utils/post_process.py:27-38, plt.py:126-131).  Warp-shuffle -> shared -> one fp64 atomic per block.
#include "common.cuh"

namespace xv2 {

constexpr int kMaxCls = 4;

template <int NCLS> struct PixelProb {
  float p[NCLS];
  float logpt;  // log p[target]
};

template <int NCLS>
__device__ __forceinline__ void softmax_px(const float* z, float* p, float& lse) {
  float m = z[0];
#pragma unroll
  for (int c = 1; c < NCLS; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCLS; ++c) {
    p[c] = expf(z[c] - m);
    s += p[c];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < NCLS; ++c) p[c] *= inv;
  lse = m + logf(s);
}

__device__ __forceinline__ int label_at(const uint8_t* labels, long long p, int h, int w, int ls) {
  if (ls == 1) return labels[p];
  const int ww = (int)(p % w);
  long long t = p / w;
  const int hh = (int)(t % h);
  const long long nb = t / h;
  return labels[(nb * (long long)(h * ls) + (long long)hh * ls) * (long long)(w * ls) + (long long)ww * ls];
}

template <int NCLS>
__global__ void __launch_bounds__(256) loss_partials_kernel(const float* __restrict__ logits,
                                                            const uint8_t* __restrict__ labels, long long pixels,
                                                            int h, int w, int ls, int post,
                                                            double* __restrict__ sums) {
  constexpr int NQ = 3 * NCLS + 3;
  float acc[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) acc[i] = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    int t = label_at(labels, p, h, w, ls);
    if (post) {
      if (t == 0) continue;
      t -= 1;
    }
    float z[NCLS], pr[NCLS], lse;
    if (NCLS == 2) {
      const float2 v = *reinterpret_cast<const float2*>(logits + p * 2);
      z[0] = v.x;
      z[1] = v.y;
    } else if (NCLS == 4) {
      const float4 v = *reinterpret_cast<const float4*>(logits + p * 4);
      z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
    } else {
#pragma unroll
      for (int c = 0; c < NCLS; ++c) z[c] = logits[p * NCLS + c];
    }
    softmax_px<NCLS>(z, pr, lse);
    float zt = z[0], pt = pr[0];
#pragma unroll
    for (int c = 0; c < NCLS; ++c) {
      const bool is_t = (c == t);
      if (is_t) { zt = z[c]; pt = pr[c]; }
      acc[c] += is_t ? pr[c] : 0.f;
      acc[NCLS + c] += pr[c];
      acc[2 * NCLS + c] += is_t ? 1.f : 0.f;
    }
    const float logpt = zt - lse;
    const float om = 1.f - expf(logpt);  // MONAI: pt = exp(logpt)
    (void)pt;
    acc[3 * NCLS + 0] += -(om * om) * logpt;
    acc[3 * NCLS + 1] += -logpt;
    acc[3 * NCLS + 2] += 1.f;
  }
  __shared__ double sm[8][NQ];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const double v = warp_sum((double)acc[i]);
    if (lane == 0) sm[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NQ) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ sums, int ncls, int terms, float weight,
                                     float* __restrict__ loss, float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double count = sums[3 * ncls + 2];
  double total = 0.0;
  const int c0 = (ncls == 2) ? 1 : 0;  // include_background=False for the 2-class head, loss.py:18-19
  const int nc = ncls - c0;
  for (int c = 0; c < ncls; ++c) coef[c] = coef[ncls + c] = 0.f;
  if (terms & XV2_LOSS_DICE) {
    double dsum = 0.0;
    for (int c = c0; c < ncls; ++c) {
      // MONAI computes in fp32; mirror its operand order: f = 1 - (2*I + s) / (G + P + s)
      const float I = (float)sums[c], P = (float)sums[ncls + c], G = (float)sums[2 * ncls + c];
      const float num = 2.0f * I + 1e-5f, den = (G + P) + 1e-5f;
      dsum += (double)(1.0f - num / den);
      coef[c] = weight * (-2.0f / den) / (float)nc;
      coef[ncls + c] = weight * (num / (den * den)) / (float)nc;
    }
    total += dsum / nc;
  }
  if (terms & XV2_LOSS_FOCAL) total += sums[3 * ncls + 0] / count;
  if (terms & XV2_LOSS_CE) total += sums[3 * ncls + 1] / count;
  coef[2 * ncls + 0] = (float)((double)weight / count);
  coef[2 * ncls + 1] = (float)count;
  loss[0] += weight * (float)total;
}

template <int NCLS>
__global__ void __launch_bounds__(256) loss_backward_kernel(const float* __restrict__ logits,
                                                            const uint8_t* __restrict__ labels, long long pixels,
                                                            int h, int w, int ls, int post, int terms,
                                                            const float* __restrict__ coef,
                                                            const float* __restrict__ dloss,
                                                            float* __restrict__ dlogits) {
  float ca[NCLS], cb[NCLS];
#pragma unroll
  for (int c = 0; c < NCLS; ++c) {
    ca[c] = coef[c];
    cb[c] = coef[NCLS + c];
  }
  const float wn = coef[2 * NCLS];
  const float up = dloss[0];
  const int ce_mult = ((terms & XV2_LOSS_CE) ? 1 : 0);
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    int t = label_at(labels, p, h, w, ls);
    float g[NCLS];
    bool skip = false;
    if (post) {
      if (t == 0) skip = true;
      t -= 1;
    }
    if (skip) {
#pragma unroll
      for (int c = 0; c < NCLS; ++c) g[c] = 0.f;
    } else {
      float z[NCLS], pr[NCLS], lse;
#pragma unroll
      for (int c = 0; c < NCLS; ++c) z[c] = logits[p * NCLS + c];
      softmax_px<NCLS>(z, pr, lse);
      float zt = z[0];
#pragma unroll
      for (int c = 0; c < NCLS; ++c)
        if (c == t) zt = z[c];
      const float logpt = zt - lse;
      const float pt = expf(logpt);
      // dice: dL/dp_c = ca_c * t_c + cb_c ; softmax jacobian
      float dp[NCLS], dot = 0.f;
#pragma unroll
      for (int c = 0; c < NCLS; ++c) {
        dp[c] = (terms & XV2_LOSS_DICE) ? (ca[c] * (c == t ? 1.f : 0.f) + cb[c]) : 0.f;
        dot = fmaf(dp[c], pr[c], dot);
      }
      // focal + ce share the (delta_jt - p_j) direction
      float k = 0.f;
      if (terms & XV2_LOSS_FOCAL) {
        const float om = 1.f - pt;
        k += (2.f * om * pt * logpt - om * om) * wn;
      }
      k -= (float)ce_mult * wn;
#pragma unroll
      for (int c = 0; c < NCLS; ++c) {
        const float delta = (c == t) ? 1.f : 0.f;
        g[c] = up * (pr[c] * (dp[c] - dot) + k * (delta - pr[c]));
      }
    }
#pragma unroll
    for (int c = 0; c < NCLS; ++c) dlogits[p * NCLS + c] = g[c];
  }
}

template <int NCLS>
__device__ __forceinline__ int argmax_px(const float* z) {
  int best = 0;
  float bv = z[0];
#pragma unroll
  for (int c = 1; c < NCLS; ++c)
    if (z[c] > bv) {  // strict: ties keep the lowest index (torch / numpy behaviour)
      bv = z[c];
      best = c;
    }
  return best;
}

// counters: [tp(nm) | fp(nm) | fn(nm)], nm = ncls_metric - 1
template <int NCLS_LOGIT, int NM>
__global__ void __launch_bounds__(256) f1_update_kernel(const float* __restrict__ logits,
                                                        const uint8_t* __restrict__ labels, long long pixels,
                                                        unsigned long long* __restrict__ counters,
                                                        uint8_t* __restrict__ pred_map) {
  int tp[NM], fp[NM], fn[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) tp[i] = fp[i] = fn[i] = 0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    float z[NCLS_LOGIT];
#pragma unroll
    for (int c = 0; c < NCLS_LOGIT; ++c) z[c] = logits[p * NCLS_LOGIT + c];
    int pred = argmax_px<NCLS_LOGIT>(z);
    const int t = labels[p];
    if (NM == 4) {  // damage: classes 1..4, only building pixels (f1.py:31-35)
      pred += 1;
      if (pred_map) pred_map[p] = (uint8_t)pred;
      if (t == 0) continue;
    } else {
      if (pred_map) pred_map[p] = (uint8_t)pred;
    }
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      const int cls = i + 1;
      tp[i] += (pred == cls && t == cls);
      fn[i] += (pred != cls && t == cls);
      fp[i] += (pred == cls && t != cls);
    }
  }
  __shared__ int sm[8][3 * NM];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    int a = tp[i], b = fp[i], c = fn[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
      sm[wid][i] = a;
      sm[wid][NM + i] = b;
      sm[wid][2 * NM + i] = c;
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 * NM) {
    unsigned long long t = 0;
    for (int k = 0; k < 8; ++k) t += (unsigned long long)sm[k][threadIdx.x];
    atomicAdd(&counters[threadIdx.x], t);
  }
}

__device__ __forceinline__ void post_rule(float loc, int post, uint8_t& pre_o, uint8_t& post_o) {
  const bool pre = (loc > 0.3f) || ((loc > 0.1f) && (post > 1));  // post_process.py:35
  pre_o = pre ? 1 : 0;
  post_o = pre ? (uint8_t)post : 0;  // post_process.py:38
}

__global__ void post_process_logits_kernel(const float* __restrict__ loc_logits, const float* __restrict__ dmg_logits,
                                           long long pixels, uint8_t* __restrict__ pre_map,
                                           uint8_t* __restrict__ post_map) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    const float loc = 1.f / (1.f + expf(-loc_logits[p * 2 + 1]));  // plt.py:128 sigmoid(pred[:,1])
    const float4 v = *reinterpret_cast<const float4*>(dmg_logits + p * 4);
    const float z[4] = {v.x, v.y, v.z, v.w};
    const int post = argmax_px<4>(z) + 1;  // softmax is monotonic: argmax of logits (post_process.py:32)
    post_rule(loc, post, pre_map[p], post_map[p]);
  }
}

__global__ void post_process_probs_kernel(const float* __restrict__ loc, const float* __restrict__ dmg,
                                          long long pixels, uint8_t* __restrict__ pre_map,
                                          uint8_t* __restrict__ post_map) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    const float z[4] = {dmg[p], dmg[pixels + p], dmg[2 * pixels + p], dmg[3 * pixels + p]};
    const int post = argmax_px<4>(z) + 1;
    post_rule(loc[p], post, pre_map[p], post_map[p]);
  }
}

// Model.save (plt.py:126-131): logits NHWC [n][hw][ncls] -> probabilities as the reference stores them per tile:
//   ncls == 2: out[n][hw]      = sigmoid(logit[.., 1])
//   ncls == 4: out[n][4][hw]   = softmax over the 4 classes (planar, the layout np.save receives)
template <int NCLS>
__global__ void save_probs_kernel(const float* __restrict__ logits, long long hw, int n, float* __restrict__ out) {
  const long long pixels = (long long)n * hw;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    if (NCLS == 2) {
      out[p] = 1.f / (1.f + expf(-logits[p * 2 + 1]));
    } else {
      const float4 v = *reinterpret_cast<const float4*>(logits + p * 4);
      const float z[4] = {v.x, v.y, v.z, v.w};
      float pr[4], lse;
      softmax_px<4>(z, pr, lse);
      const long long nb = p / hw, q = p - nb * hw;
#pragma unroll
      for (int c = 0; c < 4; ++c) out[(nb * 4 + c) * hw + q] = pr[c];
    }
  }
}

static int loss_blocks(long long pixels) {
  long long b = cdiv(pixels, 256 * 8);
  if (b > 4 * kNumSMs) b = 4 * kNumSMs;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_loss_partials(const float* logits, const uint8_t* labels, int32_t n, int32_t h, int32_t w,
                                 int32_t ncls, int32_t lstride, int32_t post, double* sums, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && lstride >= 1, "loss: empty");
  XV2_REQUIRE(ncls == 2 || ncls == 4, "loss: ncls %d unsupported (2 or 4)", ncls);
  const long long pixels = (long long)n * h * w;
  const int blocks = loss_blocks(pixels);
  if (ncls == 2)
    loss_partials_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, pixels, h, w, lstride, post, sums);
  else
    loss_partials_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, pixels, h, w, lstride, post, sums);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_loss_finalize(const double* sums, int32_t ncls, int32_t terms, float weight, float* loss,
                                 float* coef, void* stream) {
  XV2_REQUIRE(ncls == 2 || ncls == 4, "loss: ncls %d unsupported (2 or 4)", ncls);
  XV2_REQUIRE(terms != 0 && (terms & ~7) == 0, "loss: bad terms mask %d", terms);
  loss_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, ncls, terms, weight, loss, coef);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_loss_backward(const float* logits, const uint8_t* labels, int32_t n, int32_t h, int32_t w,
                                 int32_t ncls, int32_t lstride, int32_t post, int32_t terms, const float* coef,
                                 const float* dloss, float* dlogits, void* stream) {
  XV2_REQUIRE(n > 0 && h > 0 && w > 0 && lstride >= 1, "loss: empty");
  XV2_REQUIRE(ncls == 2 || ncls == 4, "loss: ncls %d unsupported (2 or 4)", ncls);
  const long long pixels = (long long)n * h * w;
  const int blocks = loss_blocks(pixels);
  if (ncls == 2)
    loss_backward_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, pixels, h, w, lstride, post, terms,
                                                                   coef, dloss, dlogits);
  else
    loss_backward_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, pixels, h, w, lstride, post, terms,
                                                                   coef, dloss, dlogits);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_f1_update(const float* logits, const uint8_t* labels, int64_t pixels, int32_t ncls_metric,
                             int64_t* counters, uint8_t* pred_map, void* stream) {
  XV2_REQUIRE(pixels > 0, "f1: empty");
  XV2_REQUIRE(ncls_metric == 2 || ncls_metric == 5, "f1: n_class %d unsupported (2 or 5)", ncls_metric);
  const int blocks = loss_blocks(pixels);
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(counters);
  if (ncls_metric == 2)
    f1_update_kernel<2, 1><<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, pixels, ctr, pred_map);
  else
    f1_update_kernel<4, 4><<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, pixels, ctr, pred_map);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_post_process(const float* loc_logits, const float* dmg_logits, int64_t pixels, uint8_t* pre_map,
                                uint8_t* post_map, void* stream) {
  XV2_REQUIRE(pixels > 0, "post_process: empty");
  post_process_logits_kernel<<<loss_blocks(pixels), 256, 0, as_stream(stream)>>>(loc_logits, dmg_logits, pixels, pre_map,
                                                                                post_map);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_post_process_probs(const float* loc, const float* dmg, int64_t pixels, uint8_t* pre_map,
                                      uint8_t* post_map, void* stream) {
  XV2_REQUIRE(pixels > 0, "post_process: empty");
  post_process_probs_kernel<<<loss_blocks(pixels), 256, 0, as_stream(stream)>>>(loc, dmg, pixels, pre_map, post_map);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_save_probs(const float* logits, int32_t n, int64_t hw, int32_t ncls, float* out, void* stream) {
  XV2_REQUIRE(n > 0 && hw > 0, "save_probs: empty");
  XV2_REQUIRE(ncls == 2 || ncls == 4, "save_probs: ncls %d unsupported (2 or 4)", ncls);
  const int blocks = loss_blocks((long long)n * hw);
  if (ncls == 2) save_probs_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(logits, hw, n, out);
  else save_probs_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(logits, hw, n, out);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
