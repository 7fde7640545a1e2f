// ResNeSt SplAtConv2d tail (call site unet.py:52) with its BatchNorm (bn0) + ReLU FOLDED IN: the post-BN activation of the
// radix convolution -- the widest tensor of every bottleneck -- is never written.
//
//   z = raw radix-conv output, bf16 [n][hw][2C] (radix halves in channel halves), y_r = relu(scale_r z_r + shift_r) in registers.
//
//   forward   gap[n][c]  = mean_hw (y_0 + y_1)                                   reads z            (xv2_splat_bn_gap)
//             att        = r-softmax(fc2(relu(bn1(fc1(gap)))))                   [n][C] vectors     (xv2_splat_fc_fwd, small.cu)
//             out        = att_0 y_0 + att_1 y_1                                 reads z, writes C  (xv2_splat_bn_combine)
//   backward  per-image partial sums over hw, with m_r = [y_r > 0]:
//             A1 = sum dout m_r, A2 = sum dout m_r z_r, M1 = sum m_r, M2 = sum m_r z_r          reads z, dout (xv2_splat_bn_bwd_partials)
//             datt[n][r,c] = sum dout y_r = scale A2 + shift A1                                  (xv2_splat_bn_bwd_datt)
//             dgap         = FC chain backward                                                    (xv2_splat_fc_bwd, small.cu)
//             du_r = (att_r dout + dgap / hw) m_r ;  the two BatchNorm reductions follow from the partial sums WITHOUT another pass:
//             sum du_r = sum_n att A1 + dgap/hw M1 ;  sum du_r xhat_r = invstd sum_n [att (A2 - mean A1) + dgap/hw (M2 - mean M1)]
//                                                                                                 (xv2_splat_bn_bwd_red)
//             dz_r = gamma invstd (du_r - mean(du) - xhat mean(du xhat))         reads z, dout, writes 2C (xv2_splat_bn_bwd_apply)
//
// HBM passes in units of one C-channel tensor: forward 5 (was 9: bn apply 4, gap 2, combine 3), backward 8 (was 16: bwd_att 3,
// bwd_x 3, bn reduce 4, bn apply 6).
#include "common.cuh"
#include <cstring>

namespace xv2 {

// Thread map: `cvt` 16-byte channel vectors per pixel row (a power of two <= 256); 256 / cvt pixel lanes per block;
// grid = (chunks of hw, n).
struct SplatMap {
  int cvt, lanes;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) {
  Vec<__nv_bfloat16> v;
  v.load(p);
  v.unpack(f);
}
__device__ __forceinline__ uint4 ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
constexpr int kUnroll = 4;  // independent 16-byte loads in flight per thread and operand: these kernels are bound by
                            // memory-level parallelism (bytes in flight per SM), not by issue slots

__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) {
  Vec<__nv_bfloat16> v;
  v.pack(f);
  v.store(p);
}

// Sums `vals[8]` over all threads of the block that own channel vector `cvi` (= tid % cvt) and hands every (vector, element)
// total to `sink(cv, i, total)` exactly once.  Stage 1: xor-shuffles inside the warp (when several pixel lanes of a vector share
// a warp); stage 2: shared memory, with ALL threads summing (a serial sum by one thread per vector cost as much as the streaming
// loop of a small block).
template <typename Sink>
__device__ __forceinline__ void block_reduce_vec8(float (*sm)[8], float* vals, int cvt, Sink&& sink) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int rows;  // partial rows per channel vector left in shared memory
  __syncthreads();
  if (cvt < 32) {
    for (int off = cvt; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], off);
    }
    if (lane < cvt) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[warp * cvt + lane][i] = vals[i];
    }
    rows = 8;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[tid][i] = vals[i];
    rows = 256 / cvt;
  }
  __syncthreads();
  for (int pair = tid; pair < cvt * 8; pair += 256) {
    const int cv = pair >> 3, i = pair & 7;
    float t = 0.f;
    for (int r = 0; r < rows; ++r) t += sm[r * cvt + cv][i];
    sink(cv, i, t);
  }
}

// ---- forward 1: gap ------------------------------------------------------------------------------------------
// nn.BatchNorm2d training statistics of one channel from the fp64 (sum, sum of squares): the arithmetic of bn_finalize_kernel (bn.cu)
struct SplatFin {
  const double* stats;        // fp64 [2 * 2c]; null: scale / shift are read from memory
  const float *gamma, *beta;
  float *rmean, *rvar;        // running statistics (may be null), updated by block (0, 0)
  float* coef;                // [4][2c] mean | invstd | scale | shift, written by block (0, 0)
  float momentum, eps;
  long long count;
};
__device__ __forceinline__ void splat_finalize_channel(const SplatFin& f, int c2, int ch, float* mean, float* invstd, float* scale,
                                                       float* shift, double* var_out) {
  const double n = (double)f.count;
  const double mu = f.stats[ch] / n;
  double var = f.stats[c2 + ch] / n - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)f.eps));
  const float g = f.gamma ? f.gamma[ch] : 1.f, b = f.beta ? f.beta[ch] : 0.f;
  *mean = (float)mu;
  *invstd = is;
  *scale = g * is;
  *shift = b - (float)mu * g * is;
  *var_out = var;
}

__global__ void __launch_bounds__(256) splat_bn_gap_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float* __restrict__ gap,
                                                           long long hw, int c, float inv_hw, SplatMap m, SplatFin fin,
                                                           double* __restrict__ gap_acc) {
  __shared__ float sm[256][8];
  const int cvi = threadIdx.x % m.cvt, lane = threadIdx.x / m.cvt;
  const int ch0 = cvi * 8, nb = blockIdx.y;
  if (fin.stats != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    // fused finalize: this block publishes the coefficients the later kernels read and updates the running statistics
    for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) {
      float mean, invstd, scl, sft;
      double var;
      splat_finalize_channel(fin, 2 * c, i, &mean, &invstd, &scl, &sft, &var);
      fin.coef[i] = mean;
      fin.coef[2 * c + i] = invstd;
      fin.coef[4 * c + i] = scl;
      fin.coef[6 * c + i] = sft;
      if (fin.rmean) {
        const double n = (double)fin.count;
        const double unbiased = fin.count > 1 ? var * n / (n - 1.0) : var;
        fin.rmean[i] = (1.f - fin.momentum) * fin.rmean[i] + fin.momentum * mean;
        fin.rvar[i] = (1.f - fin.momentum) * fin.rvar[i] + fin.momentum * (float)unbiased;
      }
    }
  }
  float sc[8], sh[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (fin.stats != nullptr) {
      float mean, invstd;
      double var;
      splat_finalize_channel(fin, 2 * c, ch0 + i, &mean, &invstd, &sc[i], &sh[i], &var);
    } else {
      sc[i] = scale[ch0 + i];
      sh[i] = shift[ch0 + i];
    }
    acc[i] = 0.f;
  }
  const __nv_bfloat16* zb = z + (long long)nb * hw * 2 * c + ch0;
  const long long stride = (long long)gridDim.x * m.lanes;
  long long p = (long long)blockIdx.x * m.lanes + lane;
  for (; p + (kUnroll - 1) * stride < hw; p += kUnroll * stride) {
    uint4 r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) r[u] = ldraw(zb + (p + u * stride) * 2 * c);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float f0[8];
      unpack8(r[u], f0);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += fmaxf(fmaf(f0[i], sc[i], sh[i]), 0.f);
    }
  }
  for (; p < hw; p += stride) {
    float f0[8];
    load8(zb + p * 2 * c, f0);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += fmaxf(fmaf(f0[i], sc[i], sh[i]), 0.f);
  }
  block_reduce_vec8(sm, acc, m.cvt, [&](int cv, int i, float t) {
    const int ch = cv * 8 + i;
    const long long gi = (long long)nb * c + (ch >= c ? ch - c : ch);  // both radix halves add into one gap channel
    // fp64 accumulation: the order of the per-block contributions no longer shows in the fp32 result (an fp32 atomic sum made
    // two identical runs differ by rounding flips that the BatchNorm over the n pooled vectors then amplified)
    if (gap_acc != nullptr) atomicAdd(&gap_acc[gi], (double)t);
    else atomicAdd(&gap[gi], t * inv_hw);
  });
}

__global__ void splat_gap_finish_kernel(const double* __restrict__ gap_acc, float* __restrict__ gap, int total, double inv_hw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) gap[i] = (float)(gap_acc[i] * inv_hw);
}

// ---- forward 2: combine (a thread owns BOTH radix halves of its 8 channels) ---------------------------------------
__global__ void __launch_bounds__(256) splat_bn_combine_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, const float* __restrict__ att,
                                                               __nv_bfloat16* __restrict__ out, long long hw, int c, SplatMap m) {
  pdl_trigger();
  pdl_wait();
  const int cvi = threadIdx.x % m.cvt, lane = threadIdx.x / m.cvt;
  const int ch0 = cvi * 8, nb = blockIdx.y;
  float s0[8], h0[8], s1[8], h1[8], a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s0[i] = scale[ch0 + i];
    h0[i] = shift[ch0 + i];
    s1[i] = scale[c + ch0 + i];
    h1[i] = shift[c + ch0 + i];
    a0[i] = att[(long long)nb * 2 * c + ch0 + i];
    a1[i] = att[(long long)nb * 2 * c + c + ch0 + i];
  }
  const __nv_bfloat16* zb = z + (long long)nb * hw * 2 * c + ch0;
  __nv_bfloat16* ob = out + (long long)nb * hw * c + ch0;
  const long long stride = (long long)gridDim.x * m.lanes;
  long long p = (long long)blockIdx.x * m.lanes + lane;
  auto emit = [&](long long px, const uint4& r0, const uint4& r1) {
    float f0[8], f1[8], o[8];
    unpack8(r0, f0);
    unpack8(r1, f1);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o[i] = a0[i] * fmaxf(fmaf(f0[i], s0[i], h0[i]), 0.f) + a1[i] * fmaxf(fmaf(f1[i], s1[i], h1[i]), 0.f);
    store8(ob + px * c, o);
  };
  for (; p + stride < hw; p += 2 * stride) {
    const uint4 r00 = ldraw(zb + p * 2 * c), r01 = ldraw(zb + p * 2 * c + c);
    const uint4 r10 = ldraw(zb + (p + stride) * 2 * c), r11 = ldraw(zb + (p + stride) * 2 * c + c);
    emit(p, r00, r01);
    emit(p + stride, r10, r11);
  }
  for (; p < hw; p += stride) emit(p, ldraw(zb + p * 2 * c), ldraw(zb + p * 2 * c + c));
}

// ---- backward 1: per-image partial sums -------------------------------------------------------------------------
// part = fp64 [4][n][2c]: A1 | A2 | M1 | M2 (accumulated; caller zero-fills)
__global__ void __launch_bounds__(256, 3) splat_bn_bwd_partials_kernel(const __nv_bfloat16* __restrict__ z,
                                                                    const __nv_bfloat16* __restrict__ dout,
                                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                                    double* __restrict__ part, long long hw, int c, int n,
                                                                    SplatMap m) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[256][8];
  const int cvi = threadIdx.x % m.cvt, lane = threadIdx.x / m.cvt;
  const int ch0 = cvi * 8, nb = blockIdx.y;
  const int chd = ch0 >= c ? ch0 - c : ch0;
  float sc[8], sh[8], a1[8], a2[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = scale[ch0 + i];
    sh[i] = shift[ch0 + i];
    a1[i] = a2[i] = m1[i] = m2[i] = 0.f;
  }
  const __nv_bfloat16* zb = z + (long long)nb * hw * 2 * c + ch0;
  const __nv_bfloat16* db = dout + (long long)nb * hw * c + chd;
  const long long stride = (long long)gridDim.x * m.lanes;
  auto accum = [&](const uint4& rz, const uint4& rd) {
    float f[8], d[8];
    unpack8(rz, f);
    unpack8(rd, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool on = fmaf(f[i], sc[i], sh[i]) > 0.f;
      const float dm = on ? d[i] : 0.f, fm = on ? f[i] : 0.f;
      a1[i] += dm;
      a2[i] = fmaf(dm, f[i], a2[i]);
      m1[i] += on ? 1.f : 0.f;
      m2[i] += fm;
    }
  };
  long long p = (long long)blockIdx.x * m.lanes + lane;
  for (; p + (kUnroll - 1) * stride < hw; p += kUnroll * stride) {
    uint4 rz[kUnroll], rd[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      rz[u] = ldraw(zb + (p + u * stride) * 2 * c);
      rd[u] = ldraw(db + (p + u * stride) * c);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) accum(rz[u], rd[u]);
  }
  for (; p < hw; p += stride) accum(ldraw(zb + p * 2 * c), ldraw(db + p * c));
  const long long plane = (long long)n * 2 * c;
  double* dst = part + (long long)nb * 2 * c;
  block_reduce_vec8(sm, a1, m.cvt, [&](int cv, int i, float t) { atomicAdd(dst + cv * 8 + i, (double)t); });
  block_reduce_vec8(sm, a2, m.cvt, [&](int cv, int i, float t) { atomicAdd(dst + plane + cv * 8 + i, (double)t); });
  block_reduce_vec8(sm, m1, m.cvt, [&](int cv, int i, float t) { atomicAdd(dst + 2 * plane + cv * 8 + i, (double)t); });
  block_reduce_vec8(sm, m2, m.cvt, [&](int cv, int i, float t) { atomicAdd(dst + 3 * plane + cv * 8 + i, (double)t); });
}

// datt[n][2c] = scale * A2 + shift * A1
__global__ void splat_bn_bwd_datt_kernel(const double* __restrict__ part, const float* __restrict__ scale,
                                         const float* __restrict__ shift, float* __restrict__ datt, int n, int c2) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c2) return;
  const int ch = i % c2;
  const long long plane = (long long)n * c2;
  datt[i] = (float)((double)scale[ch] * part[plane + i] + (double)shift[ch] * part[i]);
}

// red fp64 [2][2c]: (sum du, sum du * xhat) over the whole batch, from the per-image partial sums
__global__ void splat_bn_bwd_red_kernel(const double* __restrict__ part, const float* __restrict__ att,
                                        const float* __restrict__ dgap, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, double* __restrict__ red, int n, int c, float inv_hw) {
  pdl_trigger();
  pdl_wait();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = 2 * c;
  if (ch >= c2) return;
  const int chg = ch >= c ? ch - c : ch;
  const long long plane = (long long)n * c2;
  const double mu = mean[ch], is = invstd[ch];
  double r1 = 0.0, r2 = 0.0;
  for (int nb = 0; nb < n; ++nb) {
    const long long i = (long long)nb * c2 + ch;
    const double a = att[i], g = (double)dgap[(long long)nb * c + chg] * inv_hw;
    const double A1 = part[i], A2 = part[plane + i], M1 = part[2 * plane + i], M2 = part[3 * plane + i];
    r1 += a * A1 + g * M1;
    r2 += is * (a * (A2 - mu * A1) + g * (M2 - mu * M1));
  }
  red[ch] = r1;
  red[c2 + ch] = r2;
}

// ---- backward 2: dz -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splat_bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ z,
                                                                 const __nv_bfloat16* __restrict__ dout,
                                                                 const float* __restrict__ att, const float* __restrict__ dgap,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 const float* __restrict__ gamma, const double* __restrict__ red,
                                                                 __nv_bfloat16* __restrict__ dz, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, int accumulate, long long hw, int c,
                                                                 float inv_hw, float inv_count, SplatMap m) {
  pdl_trigger();
  pdl_wait();
  const int c2 = 2 * c;
  if (blockIdx.x == 0 && blockIdx.y == 0 && dgamma != nullptr) {
    for (int i = threadIdx.x; i < c2; i += blockDim.x) {
      dbeta[i] = (accumulate ? dbeta[i] : 0.f) + (float)red[i];
      dgamma[i] = (accumulate ? dgamma[i] : 0.f) + (float)red[c2 + i];
    }
  }
  const int cvi = threadIdx.x % m.cvt, lane = threadIdx.x / m.cvt;
  const int ch0 = cvi * 8, nb = blockIdx.y;
  const int chd = ch0 >= c ? ch0 - c : ch0;
  float sc[8], sh[8], mu[8], is[8], a[8], g[8], k0[8], k1[8], k2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = ch0 + i;
    sc[i] = scale[ch];
    sh[i] = shift[ch];
    mu[i] = mean[ch];
    is[i] = invstd[ch];
    a[i] = att[(long long)nb * c2 + ch];
    g[i] = dgap[(long long)nb * c + chd + i] * inv_hw;
    k0[i] = gamma[ch] * is[i];
    k1[i] = (float)red[ch] * inv_count;
    k2[i] = (float)red[c2 + ch] * inv_count;
  }
  const __nv_bfloat16* zb = z + (long long)nb * hw * c2 + ch0;
  const __nv_bfloat16* db = dout + (long long)nb * hw * c + chd;
  __nv_bfloat16* ob = dz + (long long)nb * hw * c2 + ch0;
  const long long stride = (long long)gridDim.x * m.lanes;
  auto emit = [&](long long px, const uint4& rz, const uint4& rd) {
    float f[8], d[8], o[8];
    unpack8(rz, f);
    unpack8(rd, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float du = fmaf(f[i], sc[i], sh[i]) > 0.f ? fmaf(a[i], d[i], g[i]) : 0.f;
      const float xh = (f[i] - mu[i]) * is[i];
      o[i] = k0[i] * (du - k1[i] - xh * k2[i]);
    }
    store8(ob + px * c2, o);
  };
  long long p = (long long)blockIdx.x * m.lanes + lane;
  for (; p + (kUnroll - 1) * stride < hw; p += kUnroll * stride) {
    uint4 rz[kUnroll], rd[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      rz[u] = ldraw(zb + (p + u * stride) * c2);
      rd[u] = ldraw(db + (p + u * stride) * c);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) emit(p + u * stride, rz[u], rd[u]);
  }
  for (; p < hw; p += stride) emit(p, ldraw(zb + p * c2), ldraw(db + p * c));
}

static bool splat_map(int vectors, SplatMap* m) {
  if (vectors < 1 || vectors > 256 || (vectors & (vectors - 1))) return false;
  m->cvt = vectors;
  m->lanes = 256 / vectors;
  return true;
}
// `waves_of`: resident CTAs per SM of the kernel -- the grid is capped at a whole number of waves (ncu, profiles/r02_ncu_kernels.md:
// the backward partial-sum kernel holds 3 CTAs per SM, so a 4-per-SM grid ran 1.33 waves at 35 % of the DRAM peak)
static dim3 splat_grid(int n, long long hw, int lanes, int per_lane, int waves_of = 4) {
  long long bx = cdiv(hw, (long long)lanes * per_lane);
  long long cap = waves_of == 4 ? cdiv(4LL * kNumSMs, n) : (long long)waves_of * kNumSMs / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)n);
}

}  // namespace xv2

using namespace xv2;

#define XV2_SPLAT_SHAPE(who, vectors)                                                                          \
  SplatMap m;                                                                                                  \
  if (n <= 0 || hw <= 0 || c <= 0 || c % 8 || !splat_map((vectors), &m)) {                                     \
    set_error(who ": serves bf16 tensors whose channel count / 8 is a power of two <= 256 (got c %d)", (int)c); \
    return XV2_EUNSUPPORTED;                                                                                   \
  }

extern "C" int xv2_splat_bn_gap(const void* z, const float* scale, const float* shift, float* gap, int32_t n, int64_t hw,
                                int32_t c, void* stream) {
  XV2_REQUIRE(z && scale && shift && gap, "splat_bn_gap: null argument");
  XV2_SPLAT_SHAPE("splat_bn_gap", 2 * c / 8)
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(gap, 0, sizeof(float) * (size_t)n * c, st);
  SplatFin fin;
  memset(&fin, 0, sizeof(fin));
  splat_bn_gap_kernel<<<splat_grid(n, hw, m.lanes, 16), 256, 0, st>>>((const __nv_bfloat16*)z, scale, shift, gap, hw, c,
                                                                     1.0f / (float)hw, m, fin, nullptr);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bn_gap_fin(const void* z, const double* stats, int64_t count, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, float momentum, float eps, float* coef, float* gap,
                                    double* gap_acc, int32_t n, int64_t hw, int32_t c, void* stream) {
  XV2_REQUIRE(z && stats && coef && gap && gap_acc && count > 0, "splat_bn_gap_fin: bad argument");
  XV2_SPLAT_SHAPE("splat_bn_gap_fin", 2 * c / 8)
  cudaStream_t st = as_stream(stream);
  SplatFin fin;
  fin.stats = stats;
  fin.gamma = gamma;
  fin.beta = beta;
  fin.rmean = running_mean;
  fin.rvar = running_var;
  fin.coef = coef;
  fin.momentum = momentum;
  fin.eps = eps;
  fin.count = count;
  splat_bn_gap_kernel<<<splat_grid(n, hw, m.lanes, 16), 256, 0, st>>>((const __nv_bfloat16*)z, nullptr, nullptr, gap, hw, c,
                                                                     1.0f / (float)hw, m, fin, gap_acc);
  splat_gap_finish_kernel<<<(n * c + 255) / 256, 256, 0, st>>>(gap_acc, gap, n * c, 1.0 / (double)hw);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bn_combine(const void* z, const float* scale, const float* shift, const float* att, void* out,
                                    int32_t n, int64_t hw, int32_t c, void* stream) {
  XV2_REQUIRE(z && scale && shift && att && out, "splat_bn_combine: null argument");
  XV2_SPLAT_SHAPE("splat_bn_combine", c / 8)
  launch_pdl(splat_bn_combine_kernel, dim3(splat_grid(n, hw, m.lanes, 8)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)z, scale, shift, att, (__nv_bfloat16*)out, hw, c, m);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bn_bwd_partials(const void* z, const void* dout, const float* scale, const float* shift, double* part,
                                         int32_t n, int64_t hw, int32_t c, void* stream) {
  XV2_REQUIRE(z && dout && scale && shift && part, "splat_bn_bwd_partials: null argument");
  XV2_SPLAT_SHAPE("splat_bn_bwd_partials", 2 * c / 8)
  launch_pdl(splat_bn_bwd_partials_kernel, dim3(splat_grid(n, hw, m.lanes, 16, 3)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)z, (const __nv_bfloat16*)dout, scale, shift, part, hw, c, n, m);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bn_bwd_datt(const double* part, const float* scale, const float* shift, float* datt, int32_t n,
                                     int32_t c, void* stream) {
  XV2_REQUIRE(part && scale && shift && datt && n > 0 && c > 0, "splat_bn_bwd_datt: bad argument");
  const int total = n * 2 * c;
  launch_pdl(splat_bn_bwd_datt_kernel, dim3((total + 255) / 256), dim3(256), 0, as_stream(stream), part, scale, shift, datt, n, 2 * c);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bn_bwd_red(const double* part, const float* att, const float* dgap, const float* mean,
                                    const float* invstd, double* red, int32_t n, int64_t hw, int32_t c, void* stream) {
  XV2_REQUIRE(part && att && dgap && mean && invstd && red && n > 0 && c > 0 && hw > 0, "splat_bn_bwd_red: bad argument");
  launch_pdl(splat_bn_bwd_red_kernel, dim3((2 * c + 127) / 128), dim3(128), 0, as_stream(stream), part, att, dgap, mean, invstd, red, n, c,
                                                                             1.0f / (float)hw);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_splat_bn_bwd_apply(const void* z, const void* dout, const float* att, const float* dgap, const float* scale,
                                      const float* shift, const float* mean, const float* invstd, const float* gamma,
                                      const double* red, void* dz, float* dgamma, float* dbeta, int32_t accumulate, int32_t n,
                                      int64_t hw, int32_t c, void* stream) {
  XV2_REQUIRE(z && dout && att && dgap && scale && shift && mean && invstd && gamma && red && dz, "splat_bn_bwd_apply: null argument");
  XV2_SPLAT_SHAPE("splat_bn_bwd_apply", 2 * c / 8)
  launch_pdl(splat_bn_bwd_apply_kernel, dim3(splat_grid(n, hw, m.lanes, 8)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)z, (const __nv_bfloat16*)dout, att, dgap, scale, shift, mean, invstd, gamma, red,
      (__nv_bfloat16*)dz, dgamma, dbeta, accumulate, hw, c, 1.0f / (float)hw, 1.0f / ((float)n * (float)hw), m);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
