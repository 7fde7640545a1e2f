// Row-strip weight gradient of 3x3 convolutions on tcgen05:  dW[k][r][s][c] = sum_pixels dY[p][k] * X[p + (r-1, s-1)][c]
//
// The reduction runs over pixels, the slow axis of both NHWC operands, so both UMMA operands are MN-major tiles straight
// from TMA (K = 16 pixel rows per instruction).  A CTA owns one (128 output channels) x (32 input channels) block of dW
// and walks down image columns like conv_strip.cu:
//   * M side: the dY row tile [pixels][128 k] (two 64-channel 128B-swizzled atoms), un-shifted;
//   * N side: the X halo row tile [pixels + 2][32 c] (64B-swizzled, fetched ONCE per row).  The three horizontal taps are the
//     three N atoms of ONE instruction: atom stride (LBO) = one pixel row = 64 B, i.e. N = 3 taps x 32 channels = 96;
//     the three vertical taps are three TMEM accumulators fed from the three live rows of the ring;
//   * per 16 pixels: 3 instructions of 128 x 96 x 16 -> no wasted lanes, operands read once per row instead of once per tap.
// Accumulators stay in TMEM over the CTA's whole pixel range; one epilogue per (k, c) block adds them to dW with
// vector reductions (red.global.add.v4.f32).
// Narrow images (W = 16 / 32 / 64) use the same code with shorter rows (K steps per row = W / 16).
#include "common.cuh"
#include "tc_common.cuh"

namespace xv2 {
using namespace tc;

int strip_encode_act(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int ld, int box_c, int box_w);
int tc_num_sms();

constexpr int kWgXRing = 6;
constexpr int kWgYRing = 4;

struct alignas(64) WgStripParams {
  CUtensorMap map_x0, map_x1, map_dy;
  int n, h, w, wt, wtiles;     // wt = pixels per row tile (min(w, 128)), wtiles = w / wt
  int groups, kchunks, cchunks;  // weight blocks = groups * kchunks * cchunks
  int cg, kg;                  // channels per group
  int kc;                      // valid dY channels per block (32 | 64 | 128)
  int c0;                      // channels of source 0 (chunks at or beyond it come from source 1)
  long long rows_total;        // blocks * n * wtiles * h
  float* dw;                   // [k][9][cg]
};

struct WgPiece {
  int blk, img, wt, ha, hb;
};
__device__ __forceinline__ WgPiece wg_piece_at(long long lo, long long hi, int h, int wtiles, int n) {
  WgPiece p;
  long long col = lo / h;
  p.ha = (int)(lo - col * h);
  const long long len = hi - lo;
  p.hb = (int)((long long)p.ha + len < (long long)h ? p.ha + len : h);
  p.wt = (int)(col % wtiles);
  col /= wtiles;
  p.img = (int)(col % n);
  p.blk = (int)(col / n);
  return p;
}

// KROWB: bytes per pixel row of the dY tile atom (128: 64-channel atoms, 64: one 32-channel atom)
template <int KROWB>
__global__ void __launch_bounds__(192, 1) wgrad_strip_kernel(const __grid_constant__ WgStripParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t XROWB = 64;  // 32 input channels per pixel row
  const uint32_t xpitch = (uint32_t)(p.wt + 8) * XROWB;         // X slot bytes (wt + 2 halo pixels, padded to 8 rows)
  const uint32_t yatom = (uint32_t)p.wt * KROWB;                // one dY atom (wt pixel rows)
  const uint32_t yatoms = p.kc > 64 ? 2u : 1u;
  const uint32_t yslot = yatoms * yatom;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xring0 = base, yring0 = base + ((kWgXRing * xpitch + 1023u) & ~1023u);
  const uint32_t bar_base = yring0 + kWgYRing * yslot;
  const uint32_t xfull0 = bar_base, xempty0 = xfull0 + 8 * kWgXRing, yfull0 = xempty0 + 8 * kWgXRing;
  const uint32_t yempty0 = yfull0 + 8 * kWgYRing, tfull = yempty0 + 8 * kWgYRing, tempty = tfull + 8, tmem_slot = tempty + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.map_x0);
    tma_prefetch_desc(&p.map_dy);
    for (int s = 0; s < kWgXRing; ++s) {
      mbar_init(xfull0 + 8 * s, 1);
      mbar_init(xempty0 + 8 * s, 1);
    }
    for (int s = 0; s < kWgYRing; ++s) {
      mbar_init(yfull0 + 8 * s, 1);
      mbar_init(yempty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  const long long lo0 = p.rows_total * blockIdx.x / gridDim.x;
  const long long hi0 = p.rows_total * (blockIdx.x + 1) / gridDim.x;
  const int blocks_per_group = p.kchunks * p.cchunks;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t xs = 0, xp = 0, ys = 0, yp = 0;
    for (long long lo = lo0; lo < hi0;) {
      const WgPiece pc = wg_piece_at(lo, hi0, p.h, p.wtiles, p.n);
      lo += pc.hb - pc.ha;
      const int g = pc.blk / blocks_per_group, rem = pc.blk - g * blocks_per_group;
      const int kci = rem / p.cchunks, cci = rem - kci * p.cchunks;
      const int cch = cci * 32;  // channel within the group's (concatenated) input
      const bool second = p.groups == 1 && cch >= p.c0;
      const CUtensorMap* mx = second ? &p.map_x1 : &p.map_x0;
      const int xc = second ? cch - p.c0 : g * p.cg + cch;
      const int yc = g * p.kg + kci * 128;
      const int w0 = pc.wt * p.wt;
      for (int row = pc.ha - 1; row <= pc.hb; ++row) {
        mbar_wait(xempty0 + 8 * xs, xp ^ 1);
        mbar_expect_tx(xfull0 + 8 * xs, (uint32_t)(p.wt + 2) * XROWB);
        tma_load_4d(xring0 + xs * xpitch, mx, xfull0 + 8 * xs, xc, w0 - 1, row, pc.img);
        if (++xs == kWgXRing) { xs = 0; xp ^= 1; }
        const int yrow = row - 1;  // dY row whose three X rows are now all in flight
        if (yrow >= pc.ha) {
          mbar_wait(yempty0 + 8 * ys, yp ^ 1);
          mbar_expect_tx(yfull0 + 8 * ys, yslot);
          for (uint32_t a = 0; a < yatoms; ++a)
            tma_load_4d(yring0 + ys * yslot + a * yatom, &p.map_dy, yfull0 + 8 * ys, yc + a * 64, w0, yrow, pc.img);
          if (++ys == kWgYRing) { ys = 0; yp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    const uint32_t idesc = make_idesc_bf16(128, 96, 1, 1);
    // A = dY (MN-major): LBO = atom stride (0 when the tile has a single atom: the upper M rows are then don't-care copies)
    const uint64_t ap = make_smem_desc(0, yatoms > 1 ? yatom : 0, 8 * KROWB, KROWB);
    // B = X (MN-major, 64B swizzle): three N atoms = three horizontal taps, one pixel row (64 B) apart
    const uint64_t bp = make_smem_desc(0, XROWB, 8 * XROWB, XROWB);
    const uint32_t a_hi = (uint32_t)(ap >> 32), b_hi = (uint32_t)(bp >> 32);
    const uint32_t a_lo0 = (uint32_t)ap + (yring0 >> 4), b_lo0 = (uint32_t)bp + (xring0 >> 4);
    const uint32_t ka = (16 * KROWB) >> 4, kb = (16 * XROWB) >> 4;  // 16 pixel rows per K step
    const int ksteps = p.wt >> 4;
    uint32_t xa = 0, xpa = 0, ys = 0, yp = 0, tphase = 0;
    int cur_blk = -1;
    uint32_t fresh = 1;  // accumulators hold nothing yet for the current block
    for (long long lo = lo0; lo < hi0;) {
      const WgPiece pc = wg_piece_at(lo, hi0, p.h, p.wtiles, p.n);
      const int nrows = pc.hb - pc.ha;
      lo += nrows;
      if (pc.blk != cur_blk) {
        if (cur_blk >= 0) {
          if (elect_one()) umma_commit(tfull);  // hand the finished block to the epilogue ...
          __syncwarp();
          mbar_wait(tempty, tphase);            // ... and wait until it has drained TMEM
          tphase ^= 1;
          tc_fence_after();
        }
        cur_blk = pc.blk;
        fresh = 1;
      }
      uint32_t xb = xa + 1, xpb = xpa;
      if (xb == kWgXRing) { xb = 0; xpb ^= 1; }
      uint32_t xc = xb + 1, xpc = xpb;
      if (xc == kWgXRing) { xc = 0; xpc ^= 1; }
      mbar_wait(xfull0 + 8 * xa, xpa);
      mbar_wait(xfull0 + 8 * xb, xpb);
      for (int i = 0; i < nrows; ++i) {
        mbar_wait(xfull0 + 8 * xc, xpc);
        mbar_wait(yfull0 + 8 * ys, yp);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + ((ys * yslot) >> 4);
          const uint32_t xrow[3] = {b_lo0 + ((xa * xpitch) >> 4), b_lo0 + ((xb * xpitch) >> 4), b_lo0 + ((xc * xpitch) >> 4)};
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            for (int j = 0; j < ksteps; ++j)
              umma_bf16_lohi(tmem_base + r * 128, a_lo + j * ka, a_hi, xrow[r] + j * kb, b_hi, idesc,
                             (uint32_t)(!(fresh && j == 0)));
          }
          umma_commit(xempty0 + 8 * xa);
          umma_commit(yempty0 + 8 * ys);
        }
        __syncwarp();
        fresh = 0;
        xa = xb; xpa = xpb;
        xb = xc; xpb = xpc;
        if (++xc == kWgXRing) { xc = 0; xpc ^= 1; }
        if (++ys == kWgYRing) { ys = 0; yp ^= 1; }
      }
      if (elect_one()) {
        umma_commit(xempty0 + 8 * xa);
        umma_commit(xempty0 + 8 * xb);
      }
      __syncwarp();
      xa = xc; xpa = xpc;
    }
    if (cur_blk >= 0) {
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else if (warp >= 2) {
    // ===================== epilogue: TMEM -> red.global.add into dW =====================
    const int q = warp & 3;
    const int kl = q * 32 + lane;  // output channel within the block
    uint32_t tphase = 0;
    int cur_blk = -1;
    for (long long lo = lo0; lo <= hi0;) {
      int blk = -1;
      if (lo < hi0) {
        const WgPiece pc = wg_piece_at(lo, hi0, p.h, p.wtiles, p.n);
        lo += pc.hb - pc.ha;
        blk = pc.blk;
      } else {
        lo = hi0 + 1;
      }
      if (blk == cur_blk) continue;
      if (cur_blk >= 0) {
        const int g = cur_blk / blocks_per_group, rem = cur_blk - g * blocks_per_group;
        const int kci = rem / p.cchunks, cci = rem - kci * p.cchunks;
        mbar_wait(tfull, tphase);
        tphase ^= 1;
        tc_fence_after();
        const bool row_ok = kl < p.kc;
        float* drow = p.dw + ((long long)(g * p.kg + kci * 128 + kl) * 9) * p.cg + cci * 32;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {  // tap t = r * 3 + s lives in accumulator r, columns s * 32 ...
          uint32_t v[32];
          tmem_ld_32x32(trow + (t / 3) * 128 + (t % 3) * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float* d = drow + (long long)t * p.cg;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
      }
      cur_blk = blk;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Returns XV2_EUNSUPPORTED when the shape is not served (caller uses the tile kernel).
int wgrad_strip_launch(const xv2_tc_conv* q, const void* src0, const void* src1, const void* dout, int lddo, float* dw,
                       void* stream) {
  const int groups = q->groups < 1 ? 1 : q->groups;
  const int ld0 = q->ld0 ? q->ld0 : q->c0, ld1 = q->ld1 ? q->ld1 : q->c1;
  const int ldy = lddo ? lddo : q->k;
  const int ctot = q->c0 + q->c1;
  if (q->convt || q->r != 3 || q->s != 3 || q->pad != 1 || q->dil > 1 || ctot % groups || q->k % groups ||
      (groups > 1 && q->c1) || ld0 % 8 || (q->c1 && ld1 % 8) || ldy % 8)
    return XV2_EUNSUPPORTED;
  const int cg = ctot / groups, kg = q->k / groups;
  if (cg % 32 || q->c0 % 32 || q->c1 % 32) return XV2_EUNSUPPORTED;
  if (!(kg == 32 || kg == 64 || kg % 128 == 0)) return XV2_EUNSUPPORTED;
  int wt;
  if (q->w >= 128) {
    if (q->w % 128) return XV2_EUNSUPPORTED;
    wt = 128;
  } else {
    if (!(q->w == 16 || q->w == 32 || q->w == 64)) return XV2_EUNSUPPORTED;
    wt = q->w;
  }
  WgStripParams p;
  memset(&p, 0, sizeof(p));
  const int kc = kg >= 128 ? 128 : kg;
  const int ybox = kc == 32 ? 32 : 64;
  int rc = strip_encode_act(&p.map_x0, src0, q->n, q->h, q->w, q->c0, ld0, 32, wt + 2);
  if (!rc && q->c1) rc = strip_encode_act(&p.map_x1, src1, q->n, q->h, q->w, q->c1, ld1, 32, wt + 2);
  if (!rc) rc = strip_encode_act(&p.map_dy, dout, q->n, q->h, q->w, q->k, ldy, ybox, wt);
  if (rc) return rc;
  p.n = q->n;
  p.h = q->h;
  p.w = q->w;
  p.wt = wt;
  p.wtiles = q->w / wt;
  p.groups = groups;
  p.kchunks = kg >= 128 ? kg / 128 : 1;
  p.cchunks = cg / 32;
  p.cg = cg;
  p.kg = kg;
  p.kc = kc;
  p.c0 = q->c0;
  p.rows_total = (long long)groups * p.kchunks * p.cchunks * q->n * p.wtiles * q->h;
  p.dw = dw;
  const uint32_t krowb = ybox * 2;
  const uint32_t xpitch = (wt + 8) * 64, yslot = (kc > 64 ? 2u : 1u) * wt * krowb;
  const size_t smem = 1024 + ((kWgXRing * xpitch + 1023u) & ~1023u) + (size_t)kWgYRing * yslot + 256;
  long long grid = tc_num_sms();
  if (grid > p.rows_total / 8) grid = p.rows_total / 8 > 0 ? p.rows_total / 8 : 1;
  cudaError_t e;
  if (krowb == 128) {
    e = cudaFuncSetAttribute(wgrad_strip_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = launch_pdl(wgrad_strip_kernel<128>, dim3((unsigned)grid), dim3(192), smem, as_stream(stream), p);
  } else {
    e = cudaFuncSetAttribute(wgrad_strip_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = launch_pdl(wgrad_strip_kernel<64>, dim3((unsigned)grid), dim3(192), smem, as_stream(stream), p);
  }
  if (e != cudaSuccess) {
    set_error("wgrad_strip: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

}  // namespace xv2
