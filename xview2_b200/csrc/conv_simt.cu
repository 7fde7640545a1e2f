// CUDA-core (SIMT) implicit-GEMM convolution: the general path (any stride / dilation / groups / channel count,
// fp32 or bf16 storage, fp32 accumulation).  Used for fp32 parity runs, for shapes the tcgen05 path does not take
// (3-channel strided stem, torchvision-ResNet strided convs, the tiny SplAt FCs) and as the on-device cross-check of
// the tensor-core kernels.  Replaces cuDNN calls reached from layers.py:83,92,71,180 and unet.py:52.
#include "common.cuh"
#include <cstdlib>

namespace xv2 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("XV2_PDL");  // opt-in: measured SLOWER on B200 inside the captured step (see common.cuh)
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

struct GatherGeom {
  int n, h, w, c, oh, ow, k, r, s, stride, pad, dil, ups, groups;
  int cg, kg, ktot;  // channels per group (in / out), reduction length r*s*cg
  long long pixels;  // n*oh*ow
};

constexpr int BM = 64, BN = 64, BK = 16;

// out[pixel][k] tile = A[pixel][(tap,c)] * W[k][(tap,c)]^T
template <typename T, typename TO>
__global__ void __launch_bounds__(256) conv_gather_kernel(GatherGeom g, const T* __restrict__ src,
                                                          const T* __restrict__ wgt, const float* __restrict__ bias,
                                                          TO* __restrict__ out) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int n_tiles_g = (g.kg + BN - 1) / BN;
  const int grp = blockIdx.y / n_tiles_g;
  const int n0 = (blockIdx.y % n_tiles_g) * BN;  // within group
  const long long m0 = (long long)blockIdx.x * BM;

  // loader mapping: k_local = tid % 16, rows tid/16 + 16*j
  const int lk = tid & 15, lr = tid >> 4;
  int p_n[4], p_h[4], p_w[4];
  bool p_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    long long m = m0 + lr + 16 * j;
    p_ok[j] = m < g.pixels;
    long long mm = p_ok[j] ? m : 0;
    int ow_ = (int)(mm % g.ow);
    long long t = mm / g.ow;
    int oh_ = (int)(t % g.oh);
    p_n[j] = (int)(t / g.oh);
    p_h[j] = oh_ * g.stride - g.pad;
    p_w[j] = ow_ * g.stride - g.pad;
  }
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4] = {};

  for (int k0 = 0; k0 < g.ktot; k0 += BK) {
    const int kk = k0 + lk;
    const bool kvalid = kk < g.ktot;
    int tap = 0, ci = 0, tr = 0, ts = 0;
    if (kvalid) {
      tap = kk / g.cg;
      ci = kk - tap * g.cg;
      tr = tap / g.s;
      ts = tap - tr * g.s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = 0.f;
      if (kvalid && p_ok[j]) {
        int nh = p_h[j] + tr * g.dil, nw = p_w[j] + ts * g.dil;
        bool ok = nh >= 0 && nw >= 0;
        if (g.ups > 1) {
          ok = ok && (nh % g.ups == 0) && (nw % g.ups == 0);
          nh /= g.ups;
          nw /= g.ups;
        }
        if (ok && nh < g.h && nw < g.w)
          v = to_f(src[(((long long)p_n[j] * g.h + nh) * g.w + nw) * g.c + grp * g.cg + ci]);
      }
      As[lk][lr + 16 * j] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + lr + 16 * j;
      float v = 0.f;
      if (kvalid && nn < g.kg) v = to_f(wgt[(long long)(grp * g.kg + nn) * g.ktot + kk]);
      Bs[lk][lr + 16 * j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + tm + i;
    if (m >= g.pixels) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + tn + j;
      if (nn >= g.kg) continue;
      int ko = grp * g.kg + nn;
      float v = acc[i][j] + (bias ? bias[ko] : 0.f);
      out[m * g.k + ko] = from_f<TO>(v);
    }
  }
}

// dw[k][(tap,c)] += sum_pixels dout[pixel][k] * src[pixel@tap][c]     (split over pixels, fp32 atomics)
template <typename T, int TM, int TN>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(GatherGeom g, const T* __restrict__ src,
                                                         const T* __restrict__ dout, float* __restrict__ dw,
                                                         long long pix_per_split) {
  constexpr int WM = TM * 16, WN = TN * 16;  // block tile: 16x16 threads, each TM x TN
  __shared__ float As[BK][WM + 4];           // [pixel][k-out]
  __shared__ float Bs[BK][WN + 4];           // [pixel][(tap,c)]
  const int tid = threadIdx.x;
  const int m_tiles_g = (g.kg + WM - 1) / WM;
  const int grp = blockIdx.y / m_tiles_g;
  const int m0 = (blockIdx.y % m_tiles_g) * WM;  // out channel within group
  const int n0 = blockIdx.x * WN;                // (tap,c) flattened
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > g.pixels) p_end = g.pixels;

  // B loader: column nn = n0 + (tid % WN) ... handled by loops below
  const int tm = (tid >> 4) * TM, tn = (tid & 15) * TN;
  float acc[TM][TN] = {};

  for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
    // A: BK pixels x WM out-channels; consecutive threads -> consecutive channels
    for (int e = tid; e < BK * WM; e += 256) {
      int col = e % WM, row = e / WM;
      long long p = p0 + row;
      float v = 0.f;
      if (p < p_end && m0 + col < g.kg) v = to_f(dout[p * g.k + grp * g.kg + m0 + col]);
      As[row][col] = v;
    }
    for (int e = tid; e < BK * WN; e += 256) {
      int col = e % WN, row = e / WN;
      long long p = p0 + row;
      int kk = n0 + col;
      float v = 0.f;
      if (p < p_end && kk < g.ktot) {
        int tap = kk / g.cg, ci = kk - tap * g.cg;
        int tr = tap / g.s, ts = tap - tr * g.s;
        int ow_ = (int)(p % g.ow);
        long long t = p / g.ow;
        int oh_ = (int)(t % g.oh);
        int nb = (int)(t / g.oh);
        int nh = oh_ * g.stride - g.pad + tr * g.dil, nw = ow_ * g.stride - g.pad + ts * g.dil;
        bool ok = nh >= 0 && nw >= 0;
        if (g.ups > 1) {
          ok = ok && (nh % g.ups == 0) && (nw % g.ups == 0);
          nh /= g.ups;
          nw /= g.ups;
        }
        if (ok && nh < g.h && nw < g.w) v = to_f(src[(((long long)nb * g.h + nh) * g.w + nw) * g.c + grp * g.cg + ci]);
      }
      Bs[row][col] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[k][tm + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tn + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int mm = m0 + tm + i;
    if (mm >= g.kg) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int kk = n0 + tn + j;
      if (kk >= g.ktot) continue;
      atomicAdd(&dw[(long long)(grp * g.kg + mm) * g.ktot + kk], acc[i][j]);
    }
  }
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long pixels, int k, float* __restrict__ out) {
  // block handles a pixel range; thread t handles channel t % k style striding
  extern __shared__ float sm[];
  const int lanes = blockDim.x / k;  // pixel lanes per block (k <= blockDim.x)
  const int ch = threadIdx.x % k, lane = threadIdx.x / k;
  float s = 0.f;
  if (lane < lanes)
    for (long long p = (long long)blockIdx.x * lanes + lane; p < pixels; p += (long long)gridDim.x * lanes)
      s += to_f(x[p * k + ch]);
  sm[threadIdx.x] = lane < lanes ? s : 0.f;
  __syncthreads();
  if (threadIdx.x < k) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += sm[l * k + threadIdx.x];
    atomicAdd(&out[threadIdx.x], t);
  }
}

template <typename TD>
__global__ void pack_weight_kernel(const float* __restrict__ src, TD* __restrict__ dst, int a, int r, int s, int b,
                                   int groups, int mode) {
  const long long total = (long long)a * r * s * b;
  const int ag = a / groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      dst[i] = from_f<TD>(src[i]);
    } else if (mode == 2) {
      // dst[rr][ss][bi][ai] = src[ai][rr][ss][bi]   (transposed-conv GEMM rows: tap-major, then out channel)
      int ai = (int)(i % a);
      long long t = i / a;
      int bi = (int)(t % b);
      t /= b;
      int ss = (int)(t % s);
      int rr = (int)(t / s);
      dst[i] = from_f<TD>(src[(((long long)ai * r + rr) * s + ss) * b + bi]);
    } else {
      // dst index i -> [row = grp*b + bi][rr][ss][ai]   (row count = groups*b, inner = ag)
      int ai = (int)(i % ag);
      long long t = i / ag;
      int ss = (int)(t % s);
      t /= s;
      int rr = (int)(t % r);
      int row = (int)(t / r);
      int grp = row / b, bi = row - grp * b;
      long long si = (((long long)(grp * ag + ai) * r + (r - 1 - rr)) * s + (s - 1 - ss)) * b + bi;
      dst[i] = from_f<TD>(src[si]);
    }
  }
}

// one launch for every packed copy of the model: blockIdx.y selects the job, blockIdx.x strides over its elements
template <typename TD>
__device__ __forceinline__ void pack_job(const xv2_pack_job& j) {
  const float* __restrict__ src = reinterpret_cast<const float*>(j.src);
  TD* __restrict__ dst = reinterpret_cast<TD*>(j.dst);
  const int a = j.a, r = j.r, s = j.s, b = j.b, mode = j.mode;
  const long long total = (long long)a * r * s * b;
  const int ag = a / j.groups;
  if (mode != 0 && ag % 16 == 0 && b % 32 == 0) {
    // Modes 1 / 2 transpose the (out-channel, in-channel) pair of every tap.  Register transpose, no shared memory: a warp reads
    // 16 source rows (ai) as 16 coalesced 128-byte loads -- lane = bi -- and every lane then owns 16 CONSECUTIVE destination
    // elements (ai0 .. ai0+15 of its row bi), written as full 32-byte (bf16) / 64-byte (fp32) segments.  The element-wise loop
    // below reads with a stride of r*s*b floats: one 32-byte sector per 4-byte element, which made the re-pack L2-bound
    // (3.65 ms per step at BASELINE config 4, 386 M parameters).
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int ta = ag / 16, tb = b / 32;
    const long long nitems = (long long)j.groups * r * s * ta * tb;
    for (long long it = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < nitems; it += nwarps) {
      long long q = it;
      const int bt = (int)(q % tb); q /= tb;
      const int at = (int)(q % ta); q /= ta;
      const int ss = (int)(q % s); q /= s;
      const int rr = (int)(q % r);
      const int grp = (int)(q / r);
      const int bi = bt * 32 + lane, ai0 = at * 16;
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int ai = ai0 + k;
        const long long si = mode == 2 ? (((long long)ai * r + rr) * s + ss) * b + bi
                                       : (((long long)(grp * ag + ai) * r + (r - 1 - rr)) * s + (s - 1 - ss)) * b + bi;
        v[k] = src[si];
      }
      const long long di = mode == 2 ? (((long long)rr * s + ss) * b + bi) * a + ai0
                                     : (((long long)(grp * b + bi) * r + rr) * s + ss) * ag + ai0;
      if constexpr (sizeof(TD) == 2) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
          w[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* o = reinterpret_cast<uint4*>(dst + di);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
      } else {
        float4* o = reinterpret_cast<float4*>(dst + di);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long si;
    if (mode == 0) {
      si = i;
    } else if (mode == 2) {
      int ai = (int)(i % a);
      long long t = i / a;
      int bi = (int)(t % b);
      t /= b;
      int ss = (int)(t % s);
      int rr = (int)(t / s);
      si = (((long long)ai * r + rr) * s + ss) * b + bi;
    } else {
      int ai = (int)(i % ag);
      long long t = i / ag;
      int ss = (int)(t % s);
      t /= s;
      int rr = (int)(t % r);
      int row = (int)(t / r);
      int grp = row / b, bi = row - grp * b;
      si = (((long long)(grp * ag + ai) * r + (r - 1 - rr)) * s + (s - 1 - ss)) * b + bi;
    }
    dst[i] = from_f<TD>(src[si]);
  }
}
__global__ void pack_weights_batched_kernel(const xv2_pack_job* __restrict__ jobs) {
  const xv2_pack_job j = jobs[blockIdx.y];
  if (j.dst_dtype == XV2_BF16) pack_job<__nv_bfloat16>(j);
  else pack_job<float>(j);
}

static int make_geom(const xv2_conv_geom* q, GatherGeom* g) {
  XV2_REQUIRE(q != nullptr, "null geometry");
  XV2_REQUIRE(q->groups >= 1 && q->c % q->groups == 0 && q->k % q->groups == 0, "channels %d/%d not divisible by groups %d",
              q->c, q->k, q->groups);
  XV2_REQUIRE(q->stride >= 1 && q->dil >= 1 && q->ups >= 1 && q->r >= 1 && q->s >= 1, "bad conv parameters");
  XV2_REQUIRE(q->n > 0 && q->h > 0 && q->w > 0 && q->oh > 0 && q->ow > 0 && q->c > 0 && q->k > 0, "empty tensor");
  g->n = q->n; g->h = q->h; g->w = q->w; g->c = q->c; g->oh = q->oh; g->ow = q->ow; g->k = q->k;
  g->r = q->r; g->s = q->s; g->stride = q->stride; g->pad = q->pad; g->dil = q->dil; g->ups = q->ups;
  g->groups = q->groups;
  g->cg = q->c / q->groups;
  g->kg = q->k / q->groups;
  g->ktot = q->r * q->s * g->cg;
  g->pixels = (long long)q->n * q->oh * q->ow;
  return XV2_OK;
}

}  // namespace xv2

using namespace xv2;

extern "C" const char* xv2_last_error(void) { return xv2::g_err; }
extern "C" int xv2_version(void) { return 100; }

extern "C" int xv2_conv_gather_simt(const xv2_conv_geom* q, const void* src, const void* w, const float* bias,
                                    void* out, void* stream) {
  GatherGeom g;
  int rc = make_geom(q, &g);
  if (rc) return rc;
  dim3 grid((unsigned)cdiv(g.pixels, BM), (unsigned)(g.groups * cdiv(g.kg, BN)));
  cudaStream_t st = as_stream(stream);
  if (q->dtype == XV2_F32 && q->out_dtype == XV2_F32)
    conv_gather_kernel<float, float><<<grid, 256, 0, st>>>(g, (const float*)src, (const float*)w, bias, (float*)out);
  else if (q->dtype == XV2_BF16 && q->out_dtype == XV2_BF16)
    conv_gather_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(g, (const __nv_bfloat16*)src,
                                                                          (const __nv_bfloat16*)w, bias,
                                                                          (__nv_bfloat16*)out);
  else if (q->dtype == XV2_BF16 && q->out_dtype == XV2_F32)
    conv_gather_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(g, (const __nv_bfloat16*)src,
                                                                  (const __nv_bfloat16*)w, bias, (float*)out);
  else
    XV2_REQUIRE(false, "unsupported dtype combination %d -> %d", q->dtype, q->out_dtype);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_conv_wgrad_simt(const xv2_conv_geom* q, const void* src, const void* dout, float* dw,
                                   void* stream) {
  GatherGeom g;
  int rc = make_geom(q, &g);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const bool small = g.kg <= 32 && g.ktot <= 64;
  const int wm = small ? 32 : 64, wn = small ? 32 : 64;
  const long long tiles = cdiv(g.ktot, wn) * g.groups * cdiv(g.kg, wm);
  long long splits = cdiv(4LL * kNumSMs, tiles);
  const long long max_splits = cdiv(g.pixels, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long pps = cdiv(cdiv(g.pixels, splits), BK) * BK;
  splits = cdiv(g.pixels, pps);
  dim3 grid((unsigned)cdiv(g.ktot, wn), (unsigned)(g.groups * cdiv(g.kg, wm)), (unsigned)splits);
  if (q->dtype == XV2_F32) {
    if (small) conv_wgrad_kernel<float, 2, 2><<<grid, 256, 0, st>>>(g, (const float*)src, (const float*)dout, dw, pps);
    else conv_wgrad_kernel<float, 4, 4><<<grid, 256, 0, st>>>(g, (const float*)src, (const float*)dout, dw, pps);
  } else if (q->dtype == XV2_BF16) {
    if (small)
      conv_wgrad_kernel<__nv_bfloat16, 2, 2><<<grid, 256, 0, st>>>(g, (const __nv_bfloat16*)src, (const __nv_bfloat16*)dout, dw, pps);
    else
      conv_wgrad_kernel<__nv_bfloat16, 4, 4><<<grid, 256, 0, st>>>(g, (const __nv_bfloat16*)src, (const __nv_bfloat16*)dout, dw, pps);
  } else {
    XV2_REQUIRE(false, "unsupported dtype %d", q->dtype);
  }
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_colsum(const void* x, int64_t pixels, int32_t k, int32_t dtype, float* out, void* stream) {
  XV2_REQUIRE(k >= 1 && k <= 256 && pixels >= 0, "colsum: k=%d out of range", k);
  if (pixels == 0) return XV2_OK;
  const int threads = 256;
  const int lanes = threads / k;
  int blocks = (int)std::min<int64_t>(cdiv(pixels, (int64_t)lanes * 8), 4 * kNumSMs);
  if (blocks < 1) blocks = 1;
  XV2_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<blocks, threads, threads * sizeof(float), as_stream(stream)>>>(
                                   (const T*)x, pixels, k, out)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_pack_weight(const float* src, void* dst, int32_t a, int32_t r, int32_t s, int32_t b,
                               int32_t groups, int32_t mode, int32_t dst_dtype, void* stream) {
  XV2_REQUIRE(a > 0 && r > 0 && s > 0 && b > 0 && groups >= 1 && a % groups == 0, "pack_weight: bad shape");
  XV2_REQUIRE(mode >= 0 && mode <= 2, "pack_weight: bad mode %d", mode);
  XV2_REQUIRE(mode != 2 || groups == 1, "pack_weight: mode 2 needs groups == 1");
  const long long total = (long long)a * r * s * b;
  int blocks = (int)std::min<long long>(cdiv(total, 256), 8 * kNumSMs);
  XV2_DISPATCH_DTYPE(dst_dtype, T, (pack_weight_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(src, (T*)dst, a, r, s, b,
                                                                                               groups, mode)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_pack_weights_batched(const xv2_pack_job* jobs, int32_t njobs, void* stream) {
  XV2_REQUIRE(jobs != nullptr && njobs > 0 && njobs <= 65535, "pack_weights_batched: bad job table");
  pack_weights_batched_kernel<<<dim3(192, njobs), 256, 0, as_stream(stream)>>>(jobs);  // measured: 592 CTAs per job is slower (empty CTAs of the ~340 small jobs)
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
