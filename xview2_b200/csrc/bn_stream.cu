// Streaming (HBM-bound) batch-norm kernels for bf16 NHWC activations, staged through shared memory with bulk async copies.
//
// The register-resident kernels in bn.cu top out near 4 TB/s: per-channel coefficients plus the loads in flight cost ~128
// registers per thread, so only 16 warps per SM are resident and too few bytes are in flight for HBM3e.  Here a producer
// lane streams 8 KB chunks of each operand into a shared-memory ring with cp.async.bulk (1-D TMA, completion on an mbarrier),
// so 96-190 KB per SM are in flight no matter how many registers the 256 consumer threads use; consumers read their 16-byte
// channel vectors back with conflict-free LDS.128 and write results with coalesced 16-byte stores.
//
// A thread always meets the same channel vector because 2048 % C == 0 (every BN layer of the ResNet / ResNeSt U-Nets:
// C = 32 ... 2048); other channel counts and fp32 use bn.cu.  Arithmetic is identical to bn.cu.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

namespace xv2 {
using namespace tc;

constexpr int kChunkElems = 4096;  // bf16 elements per chunk and operand (8 KB)
constexpr int kChunkBytes = kChunkElems * 2;

enum { BN_STATS = 0, BN_APPLY = 1, BN_BWD_REDUCE = 2, BN_BWD_APPLY = 3 };

struct BnStreamParams {
  const __nv_bfloat16* in0;   // stats/apply: x;  backward: dy
  const __nv_bfloat16* in1;   // backward: x (the BN input);  apply: residual (or null)
  const __nv_bfloat16* in2;   // backward: residual (or the second dy addend when there is no residual)
  const __nv_bfloat16* in3;   // bwd_reduce: second dy addend when a residual is present
  int res_slot, dy2_slot;     // backward: staging slot of the residual / of the second dy addend (0 = absent)
  __nv_bfloat16* out0;        // apply: y;  bwd_apply: dx;  bwd_reduce: du = (dy [+ dy2]) * act'(u), bf16 (or null)
  __nv_bfloat16* out1;        // bwd_apply: dres (or null)
  long long elems;
  int c, act, n_in, stages;
  const float *scale, *shift, *mean, *invstd, *gamma;
  const double* red_in;       // bwd_apply (training): (sum du, sum du*xhat)
  double* red_out;            // stats: (sum, sumsq);  bwd_reduce: (sum du, sum du*xhat)
  float inv_n;
  float *dgamma, *dbeta;
  int accumulate;             // bwd_apply: dgamma / dbeta are added to (the parameter's .grad in the flat buffer) instead of written
  // BN_APPLY with the finalize step fused (training): coefficients come from `fin_stats` instead of scale / shift
  const double* fin_stats;    // fp64 [2c] (sum, sum of squares) over fin_count values per channel, or null
  const float *fin_gamma, *fin_beta;
  float *fin_rmean, *fin_rvar, *fin_coef;  // running statistics (updated by block 0), coef = [4][c] mean|invstd|scale|shift
  float fin_momentum, fin_eps;
  long long fin_count;
};

// nn.BatchNorm2d training statistics of one channel (same arithmetic as bn_finalize_kernel in bn.cu)
__device__ __forceinline__ void bn_finalize_channel(const BnStreamParams& p, int ch, float* mean, float* invstd, float* scale,
                                                    float* shift, double* var_out) {
  const double n = (double)p.fin_count;
  const double mu = p.fin_stats[ch] / n;
  double var = p.fin_stats[p.c + ch] / n - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)p.fin_eps));
  const float g = p.fin_gamma ? p.fin_gamma[ch] : 1.f, b = p.fin_beta ? p.fin_beta[ch] : 0.f;
  *mean = (float)mu;
  *invstd = is;
  *scale = g * is;
  *shift = b - (float)mu * g * is;
  *var_out = var;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, float* f) {
  uint32_t w[4];
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void stg128(__nv_bfloat16* p, const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int MODE>
__global__ void __launch_bounds__(288, 2) bn_stream_kernel(const BnStreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * 8];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
  const uint32_t stage_bytes = (uint32_t)p.n_in * kChunkBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nchunks = (p.elems + kChunkElems - 1) / kChunkElems;
  const uint32_t S = (uint32_t)p.stages;
  pdl_trigger();

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 8);
    }
    fence_barrier_init();
  }
  pdl_wait();  // the statistics / coefficients read below come from the preceding kernels
  if (MODE == BN_APPLY && p.fin_stats != nullptr && blockIdx.x == 0) {
    // fused finalize: block 0 publishes the saved coefficients and updates the running statistics
    for (int i = tid; i < p.c; i += blockDim.x) {
      float mean, invstd, scale, shift;
      double var;
      bn_finalize_channel(p, i, &mean, &invstd, &scale, &shift, &var);
      p.fin_coef[i] = mean;
      p.fin_coef[p.c + i] = invstd;
      p.fin_coef[2 * p.c + i] = scale;
      p.fin_coef[3 * p.c + i] = shift;
      if (p.fin_rmean) {
        const double n = (double)p.fin_count;
        const double unbiased = p.fin_count > 1 ? var * n / (n - 1.0) : var;
        p.fin_rmean[i] = (1.f - p.fin_momentum) * p.fin_rmean[i] + p.fin_momentum * mean;
        p.fin_rvar[i] = (1.f - p.fin_momentum) * p.fin_rvar[i] + p.fin_momentum * (float)unbiased;
      }
    }
  }
  if (MODE == BN_BWD_APPLY && blockIdx.x == 0 && p.red_in != nullptr && p.dgamma != nullptr) {
    for (int i = tid; i < p.c; i += blockDim.x) {
      p.dbeta[i] = (p.accumulate ? p.dbeta[i] : 0.f) + (float)p.red_in[i];
      p.dgamma[i] = (p.accumulate ? p.dgamma[i] : 0.f) + (float)p.red_in[p.c + i];
    }
  }
  __syncthreads();

  if (warp == 8) {
    // ===================== producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const long long e0 = chunk * kChunkElems;
        const long long left = p.elems - e0;
        const uint32_t bytes = left >= kChunkElems ? (uint32_t)kChunkBytes : (uint32_t)(left * 2);
        const uint32_t fb = full0 + 8 * s;
        mbar_expect_tx(fb, bytes * p.n_in);
        const uint32_t dst = base + s * stage_bytes;
        bulk_g2s(dst, p.in0 + e0, bytes, fb);
        if (p.n_in > 1) bulk_g2s(dst + kChunkBytes, p.in1 + e0, bytes, fb);
        if (p.n_in > 2) bulk_g2s(dst + 2 * kChunkBytes, p.in2 + e0, bytes, fb);
        if (p.n_in > 3) bulk_g2s(dst + 3 * kChunkBytes, p.in3 + e0, bytes, fb);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ===================== consumers (256 threads) =====================
  const int cbase = (tid * 8) % p.c;  // first channel of this thread's vector (same for every chunk: 2048 % c == 0)
  float sc[8], sh[8], mu[8], is[8], k0[8], k1[8], k2[8];
  float a1[8], a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a1[i] = a2[i] = 0.f;
    if (MODE == BN_APPLY && p.fin_stats != nullptr) {
      float mean, invstd;
      double var;
      bn_finalize_channel(p, cbase + i, &mean, &invstd, &sc[i], &sh[i], &var);
    } else if (MODE != BN_STATS) {
      sc[i] = p.scale[cbase + i];
      sh[i] = p.shift[cbase + i];
    }
    if (MODE == BN_BWD_REDUCE || (MODE == BN_BWD_APPLY && p.red_in)) {
      mu[i] = p.mean[cbase + i];
      is[i] = p.invstd[cbase + i];
    }
    if (MODE == BN_BWD_APPLY && p.red_in) {
      const float g = p.gamma ? p.gamma[cbase + i] : 1.f;
      k0[i] = g * is[i];
      k1[i] = (float)(p.red_in[cbase + i]) * p.inv_n;
      k2[i] = (float)(p.red_in[p.c + cbase + i]) * p.inv_n;
    }
  }
  const bool has_res = (MODE == BN_APPLY) ? (p.n_in > 1) : (p.res_slot != 0);
  uint32_t s = 0, ph = 0;
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    mbar_wait(full0 + 8 * s, ph);
    const uint32_t st = base + s * stage_bytes;
    const long long e0 = chunk * kChunkElems;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int v = tid + 256 * j;
      const long long e = e0 + (long long)v * 8;
      if (e < p.elems) {
        float f0[8], f1[8], f2[8];
        lds128(st + v * 16, f0);
        if (p.n_in > 1) lds128(st + kChunkBytes + v * 16, f1);
        if (p.n_in > 2) lds128(st + 2 * kChunkBytes + v * 16, f2);
        if (MODE == BN_BWD_REDUCE && p.dy2_slot) {  // gradient arriving from two consumers: summed here, not by autograd
          float f3[8];
          if (p.dy2_slot == 3) lds128(st + 3 * kChunkBytes + v * 16, f3);
#pragma unroll
          for (int i = 0; i < 8; ++i) f0[i] = bf16_round(f0[i] + (p.dy2_slot == 3 ? f3[i] : f2[i]));
        }
        if (MODE == BN_STATS) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            a1[i] += f0[i];
            a2[i] = fmaf(f0[i], f0[i], a2[i]);
          }
        } else if (MODE == BN_APPLY) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float u = fmaf(f0[i], sc[i], sh[i]);
            if (has_res) u += f1[i];
            o[i] = apply_act(u, p.act);
          }
          stg128(p.out0 + e, o);
        } else {
          // backward: f0 = dy, f1 = x, f2 = residual
          float du[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            du[i] = f0[i];
            if (p.act != XV2_ACT_NONE) {
              float u = fmaf(f1[i], sc[i], sh[i]);
              if (has_res) u += f2[i];
              du[i] *= act_grad(u, p.act);
            }
          }
          if (MODE == BN_BWD_REDUCE) {
            if (p.out0) {  // du is what bwd_apply (act = none) and the residual branch read: sums follow the rounded values
#pragma unroll
              for (int i = 0; i < 8; ++i) du[i] = bf16_round(du[i]);
              stg128(p.out0 + e, du);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              a1[i] += du[i];
              a2[i] = fmaf(du[i], (f1[i] - mu[i]) * is[i], a2[i]);
            }
          } else {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (p.red_in) {
                const float xh = (f1[i] - mu[i]) * is[i];
                o[i] = k0[i] * (du[i] - k1[i] - xh * k2[i]);
              } else {
                o[i] = du[i] * sc[i];
              }
            }
            stg128(p.out0 + e, o);
            if (p.out1) stg128(p.out1 + e, du);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
    if (++s == S) { s = 0; ph ^= 1; }
  }
  if (MODE == BN_STATS || MODE == BN_BWD_REDUCE) {
    // threads tid, tid + c/8, tid + 2c/8, ... hold the same channels
    // the stage ring is idle now: reuse its first 16 KB for the cross-thread reduction
    float(*red_sm)[256 * 8] = reinterpret_cast<float(*)[256 * 8]>(smem_raw + (base - smem_u32(smem_raw)));
    asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red_sm[0][tid * 8 + i] = a1[i];
      red_sm[1][tid * 8 + i] = a2[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int cv = p.c >> 3;  // channel vectors (<= 256)
    if (tid < cv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double t1 = 0.0, t2 = 0.0;
        for (int t = tid; t < 256; t += cv) {
          t1 += red_sm[0][t * 8 + i];
          t2 += red_sm[1][t * 8 + i];
        }
        atomicAdd(&p.red_out[tid * 8 + i], t1);
        atomicAdd(&p.red_out[p.c + tid * 8 + i], t2);
      }
    }
  }
}

bool bn_stream_ok(int64_t pixels, int c, int dtype) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("XV2_NO_BN_STREAM");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return enabled && dtype == XV2_BF16 && c >= 8 && c <= 2048 && (2048 % c) == 0 && pixels * c >= (1 << 20);
}

template <int MODE> static int launch_stream(BnStreamParams& p, void* stream) {
  p.stages = p.n_in == 1 ? 8 : (p.n_in == 2 ? 6 : (p.n_in == 3 ? 4 : 3));
  const size_t smem = (size_t)p.stages * p.n_in * kChunkBytes + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bn_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kChunkBytes * 2);
    if (e != cudaSuccess) {
      set_error("bn_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return XV2_ECUDA;
    }
    attr_set = true;
  }
  const long long nchunks = (p.elems + kChunkElems - 1) / kChunkElems;
  long long grid = 2LL * kNumSMs;
  if (grid > nchunks) grid = nchunks;
  cudaError_t le = launch_pdl(bn_stream_kernel<MODE>, dim3((unsigned)grid), dim3(288), smem, as_stream(stream), p);
  if (le != cudaSuccess) {
    set_error("bn_stream: launch: %s", cudaGetErrorString(le));
    return XV2_ECUDA;
  }
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

int bn_stream_stats(const void* x, int64_t pixels, int c, double* stats, void* stream) {
  BnStreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const __nv_bfloat16*)x;
  p.elems = pixels * c;
  p.c = c;
  p.n_in = 1;
  p.red_out = stats;
  return launch_stream<BN_STATS>(p, stream);
}
int bn_stream_apply(const void* x, const void* res, void* y, int64_t pixels, int c, const float* scale, const float* shift,
                    int act, void* stream) {
  BnStreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const __nv_bfloat16*)x;
  p.in1 = (const __nv_bfloat16*)res;
  p.out0 = (__nv_bfloat16*)y;
  p.elems = pixels * c;
  p.c = c;
  p.act = act;
  p.n_in = res ? 2 : 1;
  p.scale = scale;
  p.shift = shift;
  return launch_stream<BN_APPLY>(p, stream);
}
int bn_stream_train_apply(const void* x, const void* res, void* y, int64_t pixels, int c, const double* stats, int64_t count,
                          const float* gamma, const float* beta, float* rmean, float* rvar, float momentum, float eps,
                          float* coef, int act, void* stream) {
  BnStreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const __nv_bfloat16*)x;
  p.in1 = (const __nv_bfloat16*)res;
  p.out0 = (__nv_bfloat16*)y;
  p.elems = pixels * c;
  p.c = c;
  p.act = act;
  p.n_in = res ? 2 : 1;
  p.fin_stats = stats;
  p.fin_gamma = gamma;
  p.fin_beta = beta;
  p.fin_rmean = rmean;
  p.fin_rvar = rvar;
  p.fin_coef = coef;
  p.fin_momentum = momentum;
  p.fin_eps = eps;
  p.fin_count = count;
  return launch_stream<BN_APPLY>(p, stream);
}
int bn_stream_bwd_reduce(const void* dy, const void* dy2, const void* x, const void* res, void* du, int64_t pixels, int c,
                         const float* scale, const float* shift, const float* mean, const float* invstd, int act, double* red,
                         void* stream) {
  BnStreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const __nv_bfloat16*)dy;
  p.in1 = (const __nv_bfloat16*)x;
  p.n_in = 2;
  if (res) {
    p.in2 = (const __nv_bfloat16*)res;
    p.res_slot = p.n_in++;
  }
  if (dy2) {
    (p.n_in == 2 ? p.in2 : p.in3) = (const __nv_bfloat16*)dy2;
    p.dy2_slot = p.n_in++;
  }
  p.out0 = (__nv_bfloat16*)du;
  p.elems = pixels * c;
  p.c = c;
  p.act = act;
  p.scale = scale;
  p.shift = shift;
  p.mean = mean;
  p.invstd = invstd;
  p.red_out = red;
  return launch_stream<BN_BWD_REDUCE>(p, stream);
}
int bn_stream_bwd_apply(const void* dy, const void* x, const void* res, void* dx, void* dres, int64_t pixels, int c,
                        const float* scale, const float* shift, const float* mean, const float* invstd, const float* gamma,
                        int act, const double* red, int64_t count, float* dgamma, float* dbeta, int accumulate, void* stream) {
  BnStreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const __nv_bfloat16*)dy;
  p.in1 = (const __nv_bfloat16*)x;
  p.in2 = (const __nv_bfloat16*)res;
  p.res_slot = res ? 2 : 0;
  p.out0 = (__nv_bfloat16*)dx;
  p.out1 = (__nv_bfloat16*)dres;
  p.elems = pixels * c;
  p.c = c;
  p.act = act;
  p.n_in = res ? 3 : 2;
  p.scale = scale;
  p.shift = shift;
  p.mean = mean;
  p.invstd = invstd;
  p.gamma = gamma;
  p.red_in = red;
  p.inv_n = 1.0f / (float)(count > 0 ? count : 1);
  p.dgamma = dgamma;
  p.dbeta = dbeta;
  p.accumulate = accumulate;
  return launch_stream<BN_BWD_APPLY>(p, stream);
}

}  // namespace xv2
