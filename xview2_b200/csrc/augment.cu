// Train-time augmentation ON THE DEVICE, after the uint8 upload (replaces the albumentations calls of
// data_loading/pytorch_loader.py:57-63,73-92,109-115,124-148):
//
//   RandomScale(p=.2, 1.0-1.3x, image cubic / mask nearest) -> CropNonEmptyMaskIfExists(512) -> H / V flip (p=.33)
//   -> GaussNoise(p=.1, var 10-50, per image) -> RandomBrightnessContrast(p=.2, per image) -> Normalize -> CHW
//
// The whole chain is ONE gather kernel per batch: every output pixel of the 512^2 crop is computed from the decoded 1024^2 tile
// (un-flip, add the crop origin, bicubic sample of the source at 1/scale), then noise, the brightness / contrast line and the
// normalisation are applied in registers and the NHWC bf16 activation is written -- the scaled image, the crop and the float
// image never exist in memory.  The random DECISIONS (probabilities, scale, sigma, alpha, beta, flips) are drawn on the host
// (a dozen scalars per sample); the crop origin needs the mask CONTENT and is chosen on the device (xv2_crop_origin).
//
// Resampling arithmetic = cv2.resize with an explicit dsize: scale = src / dst; cubic: fx = (dx + 0.5) scale - 0.5, taps
// floor(fx) - 1 .. + 2 clamped to the image (replicate), Keys weights with A = -0.75, result rounded to nearest and saturated;
// nearest: sx = min(floor(dx scale), src - 1); positions are evaluated in double like cv2 does.
#include "common.cuh"

namespace xv2 {

// per-sample parameter block (16 floats), written by the host
//  0 scale_w (src_w / scaled_w)   1 scale_h    2 scaled_w   3 scaled_h   4 crop x0   5 crop y0   6 flip_h (mirror x)   7 flip_v
//  8 sigma pre   9 sigma post   10 alpha pre   11 beta pre   12 alpha post   13 beta post   14 noise seed   15 zoom on (0 | 1)
constexpr int kAugParams = 16;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {  // "lowbias32" integer hash
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// standard normal from a counter (Box-Muller on two hashed 24-bit uniforms)
__device__ __forceinline__ float normal_at(uint32_t seed, uint32_t idx) {
  const uint32_t a = hash32(idx * 2u + 0x9e3779b9u * seed), b = hash32(idx * 2u + 1u + 0x85ebca6bu * seed);
  const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);           // [0, 1)
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

// explicit round-to-nearest multiplies / adds (no FMA contraction): the host restatement reproduces these bit for bit
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ void cubic_weights(float t, float* w) {
  const float A = -0.75f;
  const float t1 = add_(t, 1.f), u = add_(1.f, -t);
  w[0] = add_(mul_(add_(mul_(add_(mul_(A, t1), -5.f * A), t1), 8.f * A), t1), -4.f * A);
  w[1] = add_(mul_(mul_(add_(mul_(A + 2.f, t), -(A + 3.f)), t), t), 1.f);
  w[2] = add_(mul_(mul_(add_(mul_(A + 2.f, u), -(A + 3.f)), u), u), 1.f);
  w[3] = add_(add_(add_(1.f, -w[0]), -w[1]), -w[2]);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T>
__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ pre, const uint8_t* __restrict__ post,
                                                      const uint8_t* __restrict__ mask, const float* __restrict__ params,
                                                      const int* __restrict__ origin, T* __restrict__ out,
                                                      uint8_t* __restrict__ out_u8, uint8_t* __restrict__ mask_out, int n, int sh,
                                                      int sw, int oh, int ow) {
  const float mean[3] = {0.485f * 255.f, 0.456f * 255.f, 0.406f * 255.f};
  const float inv[3] = {1.f / (0.229f * 255.f), 1.f / (0.224f * 255.f), 1.f / (0.225f * 255.f)};
  const int oc = post ? 6 : 3;
  const long long total = (long long)n * oh * ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % ow);
    long long t = i / ow;
    const int y = (int)(t % oh);
    const int img = (int)(t / oh);
    const float* P = params + (long long)img * kAugParams;
    const bool zoom = P[15] != 0.f;
    const int x0 = origin ? origin[2 * img] : (int)P[4], y0 = origin ? origin[2 * img + 1] : (int)P[5];
    const int xc = (P[6] != 0.f ? ow - 1 - x : x) + x0, yc = (P[7] != 0.f ? oh - 1 - y : y) + y0;  // position in the scaled image
    const uint8_t* src[2] = {pre + (long long)img * sh * sw * 3, post ? post + (long long)img * sh * sw * 3 : nullptr};
    // cv2.resize computes source positions in DOUBLE from the integer sizes (scale = src / dst)
    const double ifx = (double)sw / (double)(int)P[2], ify = (double)sh / (double)(int)P[3];
    // ---- mask: nearest --------------------------------------------------------------------------------------------------
    if (mask_out) {
      int mx = xc, my = yc;
      if (zoom) {
        mx = min((int)floor((double)xc * ifx), sw - 1);
        my = min((int)floor((double)yc * ify), sh - 1);
      }
      mask_out[i] = mask[((long long)img * sh + my) * sw + mx];
    }
    // ---- image(s): bicubic (or a plain fetch without zoom) ------------------------------------------------------------------
    float wx[4], wy[4];
    int ix[4], iy[4];
    if (zoom) {
      const float fx = (float)__dadd_rn(__dmul_rn((double)xc + 0.5, ifx), -0.5), fy = (float)__dadd_rn(__dmul_rn((double)yc + 0.5, ify), -0.5);
      const int sx = (int)floorf(fx), sy = (int)floorf(fy);
      cubic_weights(fx - (float)sx, wx);
      cubic_weights(fy - (float)sy, wy);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ix[k] = clampi(sx - 1 + k, 0, sw - 1);
        iy[k] = clampi(sy - 1 + k, 0, sh - 1);
      }
    }
    for (int im = 0; im < (post ? 2 : 1); ++im) {
      const float sigma = P[8 + im], alpha = P[10 + 2 * im], beta = P[11 + 2 * im];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float v;
        if (zoom) {
          float acc = 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float row = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) row = __fadd_rn(row, __fmul_rn(wx[k], (float)src[im][((long long)iy[r] * sw + ix[k]) * 3 + ch]));
            acc = __fadd_rn(acc, __fmul_rn(wy[r], row));
          }
          v = fminf(fmaxf(rintf(acc), 0.f), 255.f);  // saturate_cast<uchar>
        } else {
          v = (float)src[im][((long long)yc * sw + xc) * 3 + ch];
        }
        if (sigma > 0.f) {  // GaussNoise: clip(img + N(0, sigma^2)) then the uint8 cast truncates
          const uint32_t idx = (uint32_t)(((i * 2 + im) * 3 + ch));
          v = floorf(fminf(fmaxf(__fadd_rn(v, __fmul_rn(sigma, normal_at((uint32_t)P[14], idx))), 0.f), 255.f));
        }
        if (alpha != 1.f || beta != 0.f)  // RandomBrightnessContrast LUT: clip(x * alpha + beta * 255) truncated to uint8
          v = floorf(fminf(fmaxf(__fadd_rn(__fmul_rn(v, alpha), __fmul_rn(beta, 255.f)), 0.f), 255.f));
        if (out_u8) out_u8[i * oc + im * 3 + ch] = (uint8_t)v;
        if (out) out[i * oc + im * 3 + ch] = from_f<T>(mul_(add_(v, -mean[ch]), inv[ch]));
      }
    }
  }
}

// ---- CropNonEmptyMaskIfExists: origin of the crop in the SCALED mask ----------------------------------------------------
// pass 1: non-zero count of every row of the (nearest-)scaled mask
__global__ void __launch_bounds__(256) mask_rowcount_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ params,
                                                            int* __restrict__ rowcount, int sh, int sw, int max_rows) {
  const int img = blockIdx.y, row = blockIdx.x;
  const float* P = params + (long long)img * kAugParams;
  const int hs = (int)P[3], ws = (int)P[2];
  if (row >= hs) {
    if (threadIdx.x == 0) rowcount[(long long)img * max_rows + row] = 0;
    return;
  }
  const bool zoom = P[15] != 0.f;
  const double ifx = (double)sw / (double)ws, ify = (double)sh / (double)hs;
  const int my = zoom ? min((int)floor((double)row * ify), sh - 1) : row;
  const uint8_t* src = mask + ((long long)img * sh + my) * sw;
  int cnt = 0;
  for (int x = threadIdx.x; x < ws; x += blockDim.x) {
    const int mx = zoom ? min((int)floor((double)x * ifx), sw - 1) : x;
    cnt += src[mx] != 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  __shared__ int sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += sm[k];
    rowcount[(long long)img * max_rows + row] = t;
  }
}
// pass 2 (one thread block per image): k = floor(u0 * count)-th non-zero pixel in row-major order, then
//   x_min = clip(x - floor(u1 * cw), 0, ws - cw), y_min likewise; an all-zero mask gives a uniform origin from (u1, u2).
// uniforms: [n][3] in [0, 1)
__global__ void __launch_bounds__(256) crop_origin_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ params,
                                                          const int* __restrict__ rowcount, const float* __restrict__ uniforms,
                                                          int* __restrict__ origin, int sh, int sw, int max_rows, int ch, int cw) {
  const int img = blockIdx.x;
  const float* P = params + (long long)img * kAugParams;
  const int hs = (int)P[3], ws = (int)P[2];
  const bool zoom = P[15] != 0.f;
  const float u0 = uniforms[3 * img], u1 = uniforms[3 * img + 1], u2 = uniforms[3 * img + 2];
  __shared__ long long total_s;
  __shared__ int row_s, col_s;
  __shared__ long long before_s;
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int r = 0; r < hs; ++r) tot += rowcount[(long long)img * max_rows + r];
    total_s = tot;
    row_s = -1;
    if (tot > 0) {
      long long k = (long long)floorf(u0 * (float)tot);
      if (k >= tot) k = tot - 1;
      long long acc = 0;
      for (int r = 0; r < hs; ++r) {
        const int c = rowcount[(long long)img * max_rows + r];
        if (k < acc + c) {
          row_s = r;
          before_s = k - acc;
          break;
        }
        acc += c;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int px, py;
    if (total_s > 0) {
      const int r = row_s;
      const double ifx = (double)sw / (double)ws, ify = (double)sh / (double)hs;
      const int my = zoom ? min((int)floor((double)r * ify), sh - 1) : r;
      const uint8_t* src = mask + ((long long)img * sh + my) * sw;
      long long left = before_s;
      int col = 0;
      for (int x = 0; x < ws; ++x) {
        const int mx = zoom ? min((int)floor((double)x * ifx), sw - 1) : x;
        if (src[mx] != 0) {
          if (left == 0) {
            col = x;
            break;
          }
          --left;
        }
      }
      px = col - (int)floorf(u1 * (float)cw);
      py = r - (int)floorf(u2 * (float)ch);
      px = clampi(px, 0, ws - cw);
      py = clampi(py, 0, hs - ch);
    } else {
      px = min((int)floorf(u1 * (float)(ws - cw + 1)), ws - cw);
      py = min((int)floorf(u2 * (float)(hs - ch + 1)), hs - ch);
    }
    origin[2 * img] = px;
    origin[2 * img + 1] = py;
    (void)col_s;
  }
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_crop_origin(const uint8_t* mask, const float* params, const float* uniforms, int32_t* rowcount,
                               int32_t* origin, int32_t n, int32_t sh, int32_t sw, int32_t max_rows, int32_t ch, int32_t cw,
                               void* stream) {
  XV2_REQUIRE(mask && params && uniforms && rowcount && origin && n > 0 && sh > 0 && sw > 0 && max_rows >= sh && ch > 0 && cw > 0,
              "crop_origin: bad argument");
  cudaStream_t st = as_stream(stream);
  mask_rowcount_kernel<<<dim3((unsigned)max_rows, (unsigned)n), 256, 0, st>>>(mask, params, rowcount, sh, sw, max_rows);
  crop_origin_kernel<<<n, 256, 0, st>>>(mask, params, rowcount, uniforms, origin, sh, sw, max_rows, ch, cw);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_augment_tiles(const uint8_t* pre, const uint8_t* post, const uint8_t* mask, const float* params,
                                 const int32_t* origin, void* out, uint8_t* out_u8, uint8_t* mask_out, int32_t n, int32_t sh,
                                 int32_t sw, int32_t oh, int32_t ow, int32_t out_dtype, void* stream) {
  XV2_REQUIRE(pre && params && (out || out_u8) && n > 0 && sh > 0 && sw > 0 && oh > 0 && ow > 0, "augment_tiles: bad argument");
  XV2_REQUIRE(!mask_out || mask, "augment_tiles: mask_out needs mask");
  const long long total = (long long)n * oh * ow;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  XV2_DISPATCH_DTYPE(out_dtype, T, (augment_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
                                       pre, post, mask, params, origin, (T*)out, out_u8, mask_out, n, sh, sw, oh, ow)));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
