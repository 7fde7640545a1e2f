// Rest of the submission post-processing and the xView2 scorer ON THE DEVICE, batched over tiles:
//   * connected-component majority vote (utils/post_process.py:39-43): 4-connected components of post > 0 (scipy.ndimage.label's
//     default structure) by union-find label equivalence (every foreground pixel is united with its left / upper neighbour through
//     atomicMin, then flattened), one (component, class) histogram with integer atomics, every pixel takes its component's most
//     frequent class (ties -> smallest class, like np.unique + argmax);
//   * square grey-scale dilation (post_process.py:44-45, skimage dilation(img, square(k)), odd k);
//   * the scorer's per-tile counters (utils/xview2_metrics.py:61-92): TP / FN / FP of the building mask and of damage classes
//     1-4 on target-building pixels, accumulated over tiles as 15 integers -- F1s and the score are scalar host arithmetic.
// Integer work: results are bit-exact against the host formulations (tests/test_postprocess_gpu.py).
#include "common.cuh"

namespace xv2 {

__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&L[a], b);  // hang the larger root under the smaller one
    if (old == a) return;
    a = old;  // somebody re-rooted `a` meanwhile: retry from where it points now
  }
}

__global__ void cc_init_kernel(const uint8_t* __restrict__ post, int* __restrict__ L, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    L[i] = post[i] ? (int)i : -1;
}
__global__ void cc_merge_kernel(const uint8_t* __restrict__ post, int* __restrict__ L, long long total, int h, int w) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!post[i]) continue;
    const int x = (int)(i % w), y = (int)((i / w) % h);
    if (x > 0 && post[i - 1]) uf_union(L, (int)i, (int)i - 1);
    if (y > 0 && post[i - w]) uf_union(L, (int)i, (int)i - w);
  }
}
// flatten to roots and histogram (root, class) votes; votes int32 [total][4] zero-filled by the caller
__global__ void cc_vote_kernel(const uint8_t* __restrict__ post, int* __restrict__ L, int* __restrict__ votes, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = post[i];
    if (!v) continue;
    const int r = uf_find(L, (int)i);
    L[i] = r;  // roots keep L[r] == r, so concurrent finds through this cell stay correct
    atomicAdd(&votes[(long long)r * 4 + (min(v, 4) - 1)], 1);
  }
}
__global__ void cc_assign_kernel(const uint8_t* __restrict__ post, const int* __restrict__ L, const int* __restrict__ votes,
                                 uint8_t* __restrict__ out, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!post[i]) {
      out[i] = 0;
      continue;
    }
    const int r = uf_find(L, (int)i);
    const int4 c = *reinterpret_cast<const int4*>(votes + (long long)r * 4);
    int best = 1, cnt = c.x;
    if (c.y > cnt) { best = 2; cnt = c.y; }
    if (c.z > cnt) { best = 3; cnt = c.z; }
    if (c.w > cnt) { best = 4; cnt = c.w; }
    out[i] = (uint8_t)best;
  }
}

__global__ void dilate_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long total, int h, int w, int k) {
  const int r = k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const long long base = i - (long long)y * w - x;
    int m = 0;
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        m = max(m, (int)in[base + (long long)yy * w + xx]);
      }
    }
    out[i] = (uint8_t)m;
  }
}

// counters[15] = lTP lFN lFP | dTP1 dFN1 dFP1 | ... | dTP4 dFN4 dFP4   (RowPairCalculator.get_row_pair, xview2_metrics.py:77-92)
__global__ void __launch_bounds__(256) score_counts_kernel(const uint8_t* __restrict__ lp, const uint8_t* __restrict__ dp,
                                                           const uint8_t* __restrict__ lt, const uint8_t* __restrict__ dt,
                                                           long long total, unsigned long long* __restrict__ counters) {
  __shared__ unsigned int sm[15];
  if (threadIdx.x < 15) sm[threadIdx.x] = 0;
  __syncthreads();
  unsigned int c[15];
#pragma unroll
  for (int j = 0; j < 15; ++j) c[j] = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int lpb = lp[i] > 0, ltb = lt[i] > 0, dtv = dt[i];
    c[0] += lpb & ltb;
    c[1] += (!lpb) & ltb;
    c[2] += lpb & (!ltb);
    if (dtv > 0) {  // damage is scored on target-building pixels only, with the prediction masked by the predicted buildings
      const int dpv = dp[i] * lpb;
#pragma unroll
      for (int k = 1; k <= 4; ++k) {
        c[3 * k + 0] += (dpv == k) & (dtv == k);
        c[3 * k + 1] += (dpv != k) & (dtv == k);
        c[3 * k + 2] += (dpv == k) & (dtv != k);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 15; ++j) {
    const unsigned int v = __reduce_add_sync(0xffffffffu, c[j]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[j], v);
  }
  __syncthreads();
  if (threadIdx.x < 15 && sm[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)sm[threadIdx.x]);
}

static int pp_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_cc_majority_vote(const uint8_t* post, uint8_t* out, int32_t* labels, int32_t* votes, int32_t n, int32_t h,
                                    int32_t w, void* stream) {
  XV2_REQUIRE(post && out && labels && votes && n > 0 && h > 0 && w > 0, "cc_majority_vote: bad argument");
  const long long total = (long long)n * h * w;
  XV2_REQUIRE(total < (1LL << 31), "cc_majority_vote: at most 2^31 pixels per call (batch fewer tiles)");
  cudaStream_t st = as_stream(stream);
  const int grid = pp_grid(total);
  cudaMemsetAsync(votes, 0, sizeof(int32_t) * 4 * (size_t)total, st);
  cc_init_kernel<<<grid, 256, 0, st>>>(post, labels, total);
  cc_merge_kernel<<<grid, 256, 0, st>>>(post, labels, total, h, w);
  cc_vote_kernel<<<grid, 256, 0, st>>>(post, labels, votes, total);
  cc_assign_kernel<<<grid, 256, 0, st>>>(post, labels, votes, out, total);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_dilate_square(const uint8_t* in, uint8_t* out, int32_t n, int32_t h, int32_t w, int32_t k, void* stream) {
  XV2_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && k >= 1 && (k & 1), "dilate_square: bad argument (odd footprint side expected)");
  const long long total = (long long)n * h * w;
  dilate_kernel<<<pp_grid(total), 256, 0, as_stream(stream)>>>(in, out, total, h, w, k);
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_score_counts(const uint8_t* loc_pred, const uint8_t* dmg_pred, const uint8_t* loc_targ, const uint8_t* dmg_targ,
                                int64_t pixels, int64_t* counters, void* stream) {
  XV2_REQUIRE(loc_pred && dmg_pred && loc_targ && dmg_targ && counters && pixels > 0, "score_counts: bad argument");
  score_counts_kernel<<<pp_grid(pixels), 256, 0, as_stream(stream)>>>(loc_pred, dmg_pred, loc_targ, dmg_targ, pixels,
                                                                       reinterpret_cast<unsigned long long*>(counters));
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
