// Row-strip 3x3 convolution for the HBM-bound layers (few channels, large images): forward / data gradient.
//
// Why a second kernel: at 32-128 channels a 3x3 convolution moves ~2 bytes per 300-600 FLOPs, i.e. it is bound by HBM,
// but the tile-per-tap kernel (conv_tc.cu) re-fetches the activation tile once per tap (9x through L2) and the weights
// once per tile.  Here a CTA walks DOWN a 128-pixel-wide column of one image:
//   * every input row (128 + 2 halo pixels, all channels of the group) is fetched ONCE by TMA into a ring of row slots in
//     shared memory (out-of-image halo pixels / rows are zero-filled by the TMA unit = the "same" padding);
//   * the 9 taps are 9 UMMA descriptors into that ring: tap row r picks the slot, tap column s SHIFTS THE DESCRIPTOR START
//     by s pixel rows (the 128B / 64B swizzle is a function of the absolute shared-memory address, so a shifted start
//     reads exactly what TMA wrote; measured on B200: profiles/r01_umma_shift_probe.txt);
//   * the weights of the (group, n-tile) stay resident in shared memory for the whole column;
//   * accumulators live in TMEM (double buffered); 4-8 epilogue warps convert to bf16, store, and reduce the per-channel
//     sum / sum of squares of the ROUNDED outputs (nn.BatchNorm2d statistics, layers.py:93) with a 31-step shuffle
//     transpose so that the statistics pass over the output disappears.
// HBM traffic = input once + output once; L2->SM traffic = the same (+2 halo rows per piece).
//
// Work split: the (weight set, image, column, row) space is linearised and cut into gridDim.x equal contiguous pieces.
#include "common.cuh"
#include "tc_common.cuh"

namespace xv2 {
using namespace tc;

int strip_encode_act(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int ld, int box_c, int box_w);
int strip_encode_weight(CUtensorMap* m, const void* ptr, long long rows, long long kdim, int box_k, int box_rows);
int tc_num_sms();
int strip_encode_out(CUtensorMap* m, void* ptr, long long pixels, int k, int ldo, int box_c);

constexpr int kStripPix = 128;               // output pixels per row tile (UMMA M)
constexpr int kStripHalo = kStripPix + 2;    // input pixels fetched per row
constexpr int kStripPitch = 136;             // row pitch of a slot in pixels (multiple of 8: slots stay swizzle-aligned)
constexpr int kStripMaxRing = 12;

struct alignas(64) StripParams {
  CUtensorMap map_a0, map_a1, map_b, map_out;
  int n, h, w, wtiles;
  int groups, n_tiles;      // weight sets = groups * n_tiles
  int cg, kg, bn;           // in / out channels per group, N tile
  int chunks, chunks0;      // channel chunks (of BK) per input row: total / taken from source 0
  int ring;                 // row slots
  int k_total, ldo;
  long long rows_total;     // wsets * n * wtiles * h
  __nv_bfloat16* out;
  double* stats;            // optional [2 * k_total]
  const float* ep_scale;    // optional inference epilogue: y = act(ep_scale[k] * conv + ep_shift[k])
  const float* ep_shift;
  int ep_act;
};

struct Piece {
  int wset, img, wt, ha, hb;
};
// piece of the linear row space [lo, hi) that starts at lo and ends at the column boundary or hi
__device__ __forceinline__ Piece piece_at(long long lo, long long hi, int h, int wtiles, int n) {
  Piece p;
  long long col = lo / h;
  p.ha = (int)(lo - col * h);
  long long len = hi - lo;
  p.hb = (int)((long long)p.ha + len < (long long)h ? p.ha + len : h);
  p.wt = (int)(col % wtiles);
  col /= wtiles;
  p.img = (int)(col % n);
  p.wset = (int)(col / n);
  return p;
}

template <int BK, int CHUNKS>
__global__ void __launch_bounds__(320, 1) conv_strip_kernel(const __grid_constant__ StripParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t ROWB = BK * 2;                    // bytes per pixel row of a chunk = swizzle span
  constexpr uint32_t CHUNK_BYTES = kStripPitch * ROWB;  // multiple of the swizzle pattern (8 rows)
  const uint32_t tile_bytes = (uint32_t)p.bn * ROWB;   // one (tap, chunk) weight tile
  const uint32_t w_bytes = 9u * CHUNKS * tile_bytes;
  const uint32_t slot_bytes = (uint32_t)CHUNKS * CHUNK_BYTES;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring0 = base + ((w_bytes + 1023u) & ~1023u);
  const uint32_t bar_base = ring0 + p.ring * slot_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * kStripMaxRing, wfull = empty0 + 8 * kStripMaxRing;
  const uint32_t tfull0 = wfull + 8, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
  const uint32_t stage_out0 = (bar_base + 512 + 1023u) & ~1023u;  // 2 x 8 KB output staging (one per column group)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int epi_warps = p.bn >= 64 ? 8 : 4;
  const uint32_t tmem_cols = p.bn <= 16 ? 32u : (p.bn <= 32 ? 64u : (p.bn <= 64 ? 128u : 256u));  // 2 accumulators
  pdl_trigger();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.map_a0);
    tma_prefetch_desc(&p.map_b);
    tma_prefetch_desc(&p.map_out);
    if (CHUNKS > p.chunks0) tma_prefetch_desc(&p.map_a1);
    for (int s = 0; s < p.ring; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(wfull, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, epi_warps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
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
  const uint32_t R = (uint32_t)p.ring;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t slot = 0, phase = 0;
    int cur_wset = -1;
    for (long long lo = lo0; lo < hi0;) {
      const Piece pc = piece_at(lo, hi0, p.h, p.wtiles, p.n);
      lo += pc.hb - pc.ha;
      const int grp = pc.wset / p.n_tiles, nt = pc.wset - grp * p.n_tiles;
      if (pc.wset != cur_wset) {
        // every MMA that reads the resident weights has completed once all row slots have been released
        uint32_t s2 = slot, p2 = phase;
        for (uint32_t j = 0; j < R; ++j) {
          mbar_wait(empty0 + 8 * s2, p2 ^ 1);
          if (++s2 == R) { s2 = 0; p2 ^= 1; }
        }
        mbar_expect_tx(wfull, w_bytes);
        for (int t = 0; t < 9; ++t)
          for (int ch = 0; ch < CHUNKS; ++ch)
            tma_load_2d(base + (uint32_t)(t * CHUNKS + ch) * tile_bytes, &p.map_b, wfull, t * p.cg + ch * BK,
                        grp * p.kg + nt * p.bn);
        cur_wset = pc.wset;
      }
      for (int row = pc.ha - 1; row <= pc.hb; ++row) {
        mbar_wait(empty0 + 8 * slot, phase ^ 1);
        const uint32_t fb = full0 + 8 * slot;
        mbar_expect_tx(fb, (uint32_t)CHUNKS * kStripHalo * ROWB);
        const uint32_t dst = ring0 + slot * slot_bytes;
#pragma unroll
        for (int ch = 0; ch < CHUNKS; ++ch) {
          if (ch < p.chunks0)
            tma_load_4d(dst + ch * CHUNK_BYTES, &p.map_a0, fb, grp * p.cg + ch * BK, pc.wt * kStripPix - 1, row, pc.img);
          else
            tma_load_4d(dst + ch * CHUNK_BYTES, &p.map_a1, fb, (ch - p.chunks0) * BK, pc.wt * kStripPix - 1, row, pc.img);
        }
        if (++slot == R) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop (addresses stay warp-uniform, no per-MMA waterfall); one elected lane issues.
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.bn, 0, 0);
    const uint64_t dproto = make_smem_desc(0, 16, 8 * ROWB, ROWB);
    const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
    const uint32_t a_lo0 = d_lo + (ring0 >> 4), b_lo0 = d_lo + (base >> 4);
    const uint32_t slot16 = slot_bytes >> 4, tile16 = tile_bytes >> 4;
    uint32_t sa = 0, pa = 0;  // ring iterator of the oldest live input row (slot, phase)
    uint32_t acc = 0, acc_phase = 0, wphase = 0;
    int cur_wset = -1;
    for (long long lo = lo0; lo < hi0;) {
      const Piece pc = piece_at(lo, hi0, p.h, p.wtiles, p.n);
      const int nrows = pc.hb - pc.ha;
      lo += nrows;
      if (pc.wset != cur_wset) {
        mbar_wait(wfull, wphase);
        wphase ^= 1;
        cur_wset = pc.wset;
      }
      uint32_t sb = sa + 1, pb = pa;
      if (sb == R) { sb = 0; pb ^= 1; }
      uint32_t sc = sb + 1, pcph = pb;
      if (sc == R) { sc = 0; pcph ^= 1; }
      mbar_wait(full0 + 8 * sa, pa);
      mbar_wait(full0 + 8 * sb, pb);
      for (int i = 0; i < nrows; ++i) {
        mbar_wait(full0 + 8 * sc, pcph);
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * (tmem_cols >> 1);
        if (elect_one()) {
          const uint32_t rows_lo[3] = {a_lo0 + sa * slot16, a_lo0 + sb * slot16, a_lo0 + sc * slot16};
          uint32_t b_lo = b_lo0;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
#pragma unroll
              for (int ch = 0; ch < CHUNKS; ++ch) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_bf16_lohi(d, rows_lo[r] + ((ch * CHUNK_BYTES + s * ROWB + k * 32) >> 4), d_hi, b_lo + 2 * k, d_hi, idesc,
                                 (uint32_t)((r | s | ch | k) != 0));
                b_lo += tile16;
              }
            }
          }
          umma_commit(tfull0 + 8 * acc);
          umma_commit(empty0 + 8 * sa);  // the oldest input row is not needed by later output rows
        }
        __syncwarp();
        sa = sb; pa = pb;
        sb = sc; pb = pcph;
        if (++sc == R) { sc = 0; pcph ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (elect_one()) {
        umma_commit(empty0 + 8 * sa);
        umma_commit(empty0 + 8 * sb);
      }
      __syncwarp();
      sa = sc; pa = pcph;
    }
  } else if (warp >= 2 && warp < 2 + epi_warps) {
    // ===================== epilogue =====================
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int cgrp = (warp - 2) >> 2;          // column group (0 | 1)
    const int cgroups = epi_warps >> 2;
    const int m = q * 32 + lane;               // pixel within the row tile
    const bool issuer = q == 2 && lane == 0;   // lane 0 of the column group's first warp issues the TMA stores
    const uint32_t sbuf = stage_out0 + cgrp * 8192;  // [128 pixels][32 channels] bf16, 64B-swizzled
    const uint32_t row_addr = sbuf + m * 64;
    const uint32_t swz = (uint32_t)((m >> 1) & 3);
    uint32_t acc = 0, acc_phase = 0;
    for (long long lo = lo0; lo < hi0;) {
      const Piece pc = piece_at(lo, hi0, p.h, p.wtiles, p.n);
      const int nrows = pc.hb - pc.ha;
      lo += nrows;
      const int grp = pc.wset / p.n_tiles, nt = pc.wset - grp * p.n_tiles;
      const int co0 = grp * p.kg + nt * p.bn;
      float ssum1 = 0.f, ssq1 = 0.f;  // second column block: lane j holds channel (32 + j) after each row's transpose-reduce
      float rs[32], rq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) rs[j] = rq[j] = 0.f;
      for (int i = 0; i < nrows; ++i) {
        const int pix0 = (pc.img * p.h + (pc.ha + i)) * p.w + pc.wt * kStripPix;
        mbar_wait(tfull0 + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (tmem_cols >> 1);
        int bi = 0;
        for (int c0 = cgrp * 32; c0 < p.bn; c0 += cgroups * 32, ++bi) {
          uint32_t v[32];
          tmem_ld_32x32(trow + c0, v);
          tmem_ld_wait();
          if (issuer) bulk_wait_read0();  // the previous TMA store out of this group's staging buffer has been read
          named_bar_sync(1 + cgrp, 128);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t w4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float a = __uint_as_float(v[j + 2 * t]), b = __uint_as_float(v[j + 2 * t + 1]);
              if (p.ep_scale) {
                const int ch = co0 + c0 + j + 2 * t;
                a = apply_act(fmaf(a, p.ep_scale[ch], p.ep_shift[ch]), p.ep_act);
                b = apply_act(fmaf(b, p.ep_scale[ch + 1], p.ep_shift[ch + 1]), p.ep_act);
              }
              __nv_bfloat162 hp = __floats2bfloat162_rn(a, b);
              w4[t] = *reinterpret_cast<uint32_t*>(&hp);
              f[j + 2 * t] = __uint_as_float(w4[t] << 16);
              f[j + 2 * t + 1] = __uint_as_float(w4[t] & 0xffff0000u);
            }
            st_shared_v4(row_addr + ((((uint32_t)j >> 3) ^ swz) << 4), w4[0], w4[1], w4[2], w4[3]);
          }
          fence_proxy_async();
          named_bar_sync(1 + cgrp, 128);
          if (issuer) {
            tma_store_2d(&p.map_out, sbuf, co0 + c0, pix0);
            bulk_commit();
          }
          if (p.stats) {
            if (bi == 0) {
              // first column block: per-thread running sums over the whole piece, transposed once at its end
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                rs[j] += f[j];
                rq[j] = fmaf(f[j], f[j], rq[j]);
              }
            } else {
              float sq[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
              warp_transpose_sum(f, lane);
              warp_transpose_sum(sq, lane);
              ssum1 += f[0];
              ssq1 += sq[0];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (p.stats) {
        warp_transpose_sum(rs, lane);
        warp_transpose_sum(rq, lane);
        int bi = 0;
        for (int c0 = cgrp * 32; c0 < p.bn; c0 += cgroups * 32, ++bi) {
          atomicAdd(p.stats + co0 + c0 + lane, (double)(bi == 0 ? rs[0] : ssum1));
          atomicAdd(p.stats + p.k_total + co0 + c0 + lane, (double)(bi == 0 ? rq[0] : ssq1));
        }
      }
    }
  }
  if (warp >= 2 && warp < 2 + epi_warps && (warp & 3) == 2 && lane == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// Host side.  Returns XV2_EUNSUPPORTED when the shape is not a strip shape (caller uses the tile-per-tap kernel).
int conv_strip_launch(const xv2_tc_conv* q, const void* src0, const void* src1, const void* w, void* out, double* stats,
                      const float* ep_scale, const float* ep_shift, int ep_act, void* stream) {
  const int groups = q->groups < 1 ? 1 : q->groups;
  const int ld0 = q->ld0 ? q->ld0 : q->c0, ld1 = q->ld1 ? q->ld1 : q->c1;
  const int ctot = q->c0 + q->c1;
  if (q->convt || q->r != 3 || q->s != 3 || q->pad != 1 || (q->dil > 1) || q->out_dtype != XV2_BF16 || q->w % kStripPix ||
      ctot % groups || q->k % groups || (groups > 1 && q->c1) || ld0 % 8 || (q->c1 && ld1 % 8))
    return XV2_EUNSUPPORTED;
  const int cg = ctot / groups, kg = q->k / groups;
  if (!(cg == 32 || cg == 64 || cg == 128) || kg % 32) return XV2_EUNSUPPORTED;
  if (q->c1 && (q->c0 % 64 || q->c1 % 64)) return XV2_EUNSUPPORTED;
  if (cg == 128 && kg > 64) return XV2_EUNSUPPORTED;  // compute-bound: the tile-per-tap kernel with a wide N is the better fit
  const int bk = cg == 32 ? 32 : 64;
  const int chunks = cg / bk;
  const uint32_t rowb = bk * 2, chunk_bytes = kStripPitch * rowb, slot_bytes = chunks * chunk_bytes;
  const uint32_t budget = 232448 - 1024 - 512 - (1024 + 2 * 8192);  // minus the output staging buffers
  int bn = 0, ring = 0;
  for (int cand : {128, 64, 32}) {
    if (kg % cand) continue;
    if (cg == 128 && cand > 32) continue;
    const uint32_t wb = ((9u * chunks * cand * rowb) + 1023u) & ~1023u;
    if (wb + 4 * slot_bytes > budget) continue;
    bn = cand;
    ring = (int)((budget - wb) / slot_bytes);
    break;
  }
  if (!bn) return XV2_EUNSUPPORTED;
  if (ring > kStripMaxRing) ring = kStripMaxRing;
  const int ldo = q->ldo ? q->ldo : q->k;
  if (ldo % 8) return XV2_EUNSUPPORTED;

  StripParams p;
  memset(&p, 0, sizeof(p));
  int rc = strip_encode_act(&p.map_a0, src0, q->n, q->h, q->w, q->c0, ld0, bk, kStripHalo);
  if (!rc && q->c1) rc = strip_encode_act(&p.map_a1, src1, q->n, q->h, q->w, q->c1, ld1, bk, kStripHalo);
  if (!rc) rc = strip_encode_weight(&p.map_b, w, q->k, 9LL * cg, bk, bn);
  if (!rc) rc = strip_encode_out(&p.map_out, out, (long long)q->n * q->h * q->w, q->k, ldo, 32);
  if (rc) return rc;
  p.n = q->n;
  p.h = q->h;
  p.w = q->w;
  p.wtiles = q->w / kStripPix;
  p.groups = groups;
  p.n_tiles = kg / bn;
  p.cg = cg;
  p.kg = kg;
  p.bn = bn;
  p.chunks = chunks;
  p.chunks0 = groups > 1 ? chunks : q->c0 / bk;
  p.ring = ring;
  p.k_total = q->k;
  p.ldo = ldo;
  p.rows_total = (long long)groups * p.n_tiles * q->n * p.wtiles * q->h;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.stats = stats;
  p.ep_scale = ep_scale;
  p.ep_shift = ep_shift;
  p.ep_act = ep_act;
  const uint32_t wb = ((9u * chunks * bn * rowb) + 1023u) & ~1023u;
  const size_t smem = 1024 + wb + (size_t)ring * slot_bytes + 512 + 1024 + 2 * 8192;
  long long grid = tc_num_sms();
  if (grid > p.rows_total / 16) grid = p.rows_total / 16 > 0 ? p.rows_total / 16 : 1;
  cudaError_t e;
#define XV2_STRIP_LAUNCH(BKV, CH)                                                                                        \
  do {                                                                                                                   \
    e = cudaFuncSetAttribute(conv_strip_kernel<BKV, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e == cudaSuccess) e = launch_pdl(conv_strip_kernel<BKV, CH>, dim3((unsigned)grid), dim3(320), smem, as_stream(stream), p); \
  } while (0)
  if (bk == 32) XV2_STRIP_LAUNCH(32, 1);
  else if (chunks == 1) XV2_STRIP_LAUNCH(64, 1);
  else XV2_STRIP_LAUNCH(64, 2);
#undef XV2_STRIP_LAUNCH
  if (e != cudaSuccess) {
    set_error("conv_strip: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

}  // namespace xv2
