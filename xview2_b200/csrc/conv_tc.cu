// tcgen05 implicit-GEMM convolution for sm_100a: forward / data-gradient (one kernel) and weight-gradient.
//
// Forward (also dgrad with swapped + flipped weights, transposed conv, transposed-conv dgrad):
//   D[pixel, k] = sum_{tap, c} A[pixel (+) tap, c] * W[k, tap, c]
//   * M tile = 128 output pixels arranged as a TH x TW spatial patch of one image (TH*TW = 128).
//   * A k-block is (one tap, BK channels).  The activation tile is fetched by ONE TMA tiled load of a
//     {BK, TW, TH, 1} box whose (w, h) coordinates are shifted by the tap offset: out-of-bounds rows/columns
//     (the "same" padding) are zero-filled by the TMA unit, so no im2col buffer and no halo bookkeeping exist.
//   * A second activation tensor can feed the tail of the channel range: the decoder's torch.cat((up, skip)) is
//     never materialised (layers.py:167, layers.py:114).
//   * The 2x2/stride-2 transposed convolution is the same GEMM with N = 4*Cout and a pixel-shuffle store;
//     its data gradient gathers the 2x2 patch with a 5-D tensor map {C, 2, W, 2, H*N}.
//   * 128B (BK=64) or 64B (BK=32) swizzled K-major operand tiles, fp32 accumulators in TMEM (2 x 256 columns,
//     double buffered so the epilogue of tile i overlaps the main loop of tile i+1).
//   * warp 0: TMA producer, warp 1: MMA issuer (one thread) + TMEM owner, warps 2-5: epilogue (TMEM -> regs -> HBM).
//
// Weight gradient:
//   dW[k, tap, c] = sum_pixels dY[pixel, k] * X[pixel (+) tap, c]
//   * the reduction runs over pixels, which is the slow axis of both NHWC operands: both UMMA operands are
//     MN-major tiles [128 pixels][64|32 channels] straight from TMA.  M = (tap, c) atoms, N = k.
//   * split over pixel tiles; fp32 partial tiles are reduced into dW with red.global.add.f32.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

namespace xv2 {
using namespace tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = kNumSMs;

static int ensure_init() {
  if (g_encode) return XV2_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    g_num_sms = sms;
  return XV2_OK;
}

constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr uint32_t kSmemBudget = 232448 - 2048;  // 227 KB minus alignment slack and the barrier block

struct alignas(64) TcParams {
  CUtensorMap map_a0, map_a1, map_b;
  int kb0, kb1;            // k-blocks per tap taken from source 0 / source 1
  int taps, tap_s;         // r*s, s
  int pad, dil;
  int tw_log2, tw, th;     // spatial tile
  int tiles_w, tiles_h;    // per image
  int m_tiles, n_tiles, groups;
  int cg, kg;              // in / out channels per group
  int bn;                  // N tile
  int stages;
  int gather2x2;           // A is the 5-D {C,2,W,2,H*N} view (transposed-conv dgrad): tap = (kh,kw) picks dims 1,3
  int convt;               // epilogue scatters to the (2h+kh, 2w+kw) pixel, n tile lies inside one tap
  int h, w;                // source spatial dims (= tile grid dims)
  int k_total;             // output channels (convt: per tap)
  int ldo;                 // output pixel stride
  int out_f32;
  void* out;
  const float* bias;
  // TMA-store epilogue (bf16 output, N tile a multiple of 32): staged through swizzled shared memory
  CUtensorMap map_out;
  int tma_store;           // 0: direct register -> global stores
  int store_c;             // channels per staged block (64 -> 128B swizzle, 32 -> 64B swizzle)
  int out_bufs;            // staging buffers per epilogue half: 2 for single-tap (1x1 / transposed) convolutions, whose main loop
                           // is one or two k-blocks long -- there the epilogue is the critical path and a second buffer lets the
                           // next block be converted while the TMA store of the previous one is still reading shared memory
  double* stats;           // optional fp64 [2 * k_total]: BN sum / sum of squares of the rounded outputs
  // optional inference epilogue: y = act(ep_scale[k] * conv + ep_shift[k])  (eval-mode BatchNorm + activation folded in)
  const float* ep_scale;
  const float* ep_shift;
  int ep_act;
};

// ---------------------------------------------------------------------------------------------------------------
template <int BK>
__global__ void __launch_bounds__(320, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = 128 * BK * 2;
  constexpr uint32_t SWZ = BK * 2;          // swizzle span = bytes per operand row
  constexpr uint32_t SBO = 8 * SWZ;         // 8-row core-matrix group
  const uint32_t b_bytes = (uint32_t)p.bn * BK * 2;
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 64, tfull0 = bar_base + 128, tempty0 = bar_base + 144;
  const uint32_t tmem_slot = bar_base + 160;
  const uint32_t stage_out0 = (bar_base + 256 + 1023u) & ~1023u;  // 2 x out_bufs x 16 KB output staging (per epilogue half)
  const uint32_t stats_sm = stage_out0 + 2 * p.out_bufs * 16384;   // float [2 * k_total] when p.stats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int epi_warps = p.tma_store ? 8 : 4;
  pdl_trigger();

  if (p.stats) {
    for (int i = threadIdx.x; i < 2 * p.k_total; i += blockDim.x)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(stats_sm + 4 * i), "f"(0.f) : "memory");
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.map_a0);
    tma_prefetch_desc(&p.map_b);
    if (p.kb1 > 0) tma_prefetch_desc(&p.map_a1);
    if (p.tma_store) tma_prefetch_desc(&p.map_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, epi_warps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // everything above touched only shared memory, TMEM and the kernel parameters

  const int total_tiles = p.m_tiles * p.n_tiles * p.groups;
  const int kb_per_tap = p.kb0 + p.kb1;
  const int num_kb = p.taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int t = tile / p.n_tiles;
      const int g = t % p.groups;
      const int m_tile = t / p.groups;
      const int twi = m_tile % p.tiles_w;
      const int t2 = m_tile / p.tiles_w;
      const int thi = t2 % p.tiles_h;
      const int img = t2 / p.tiles_h;
      const int h0 = thi * p.th, w0 = twi * p.tw;
      const int brow = g * p.kg + n_tile * p.bn;
      int bcol = 0;
      const int tap_r = p.taps / p.tap_s;
      for (int tr = 0; tr < tap_r; ++tr) {
        const int hc = h0 + tr * p.dil - p.pad;
        for (int ts = 0; ts < p.tap_s; ++ts) {
          const int wc = w0 + ts * p.dil - p.pad;
          for (int kb = 0; kb < kb_per_tap; ++kb, bcol += BK) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t sa = base + stage * stage_bytes, sb = sa + A_BYTES;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, stage_bytes);
            if (p.gather2x2) {
              tma_load_5d(sa, &p.map_a0, fb, kb * BK, ts, w0, tr, img * p.h + h0);
            } else if (kb < p.kb0) {
              tma_load_4d(sa, &p.map_a0, fb, g * p.cg + kb * BK, wc, hc, img);
            } else {
              tma_load_4d(sa, &p.map_a1, fb, (kb - p.kb0) * BK, wc, hc, img);
            }
            tma_load_2d(sb, &p.map_b, fb, bcol, brow);
            if (++stage == (uint32_t)p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp walks the loop (warp-uniform addresses), one elected lane issues the MMAs and commits
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.bn, 0, 0);
    const uint64_t dproto = make_smem_desc(0, 16, SBO, SWZ);
    const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
    const uint32_t stage16 = stage_bytes >> 4;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    uint32_t a_lo = d_lo + (base >> 4);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 256;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_lo = a_lo + (A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_lohi(d, a_lo + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, (uint32_t)((kb | k) != 0));
          umma_commit(empty0 + 8 * stage);
          if (kb == num_kb - 1) umma_commit(tfull0 + 8 * acc);
        }
        __syncwarp();
        a_lo += stage16;
        if (++stage == (uint32_t)p.stages) {
          stage = 0;
          phase ^= 1;
          a_lo = d_lo + (base >> 4);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2 && p.tma_store) {
    // ===================== epilogue (TMA store): 2 halves x 4 warps; half h owns staged blocks h, h+2, ... =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;       // row of the tile (pixel)
    const bool issuer = (q == ((2 + 4 * half) & 3)) && lane == 0;  // lane 0 of the half's first warp
    const uint32_t sbuf0 = stage_out0 + half * p.out_bufs * 16384;
    const int nblocks = p.bn / p.store_c;
    const uint32_t rowb = (uint32_t)p.store_c * 2;
    const uint32_t swz = p.store_c == 64 ? (uint32_t)(m & 7) : (uint32_t)((m >> 1) & 3);
    uint32_t acc = 0, acc_phase = 0, obuf = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int t = tile / p.n_tiles;
      const int g = t % p.groups;
      const int m_tile = t / p.groups;
      const int twi = m_tile % p.tiles_w;
      const int t2 = m_tile / p.tiles_w;
      const int thi = t2 % p.tiles_h;
      const int img = t2 / p.tiles_h;
      // transposed conv: N index = (kh, kw, k); a STAGED BLOCK lies inside one kh (the tile may span both): co0 = tile's first N index
      const int co0 = p.convt ? n_tile * p.bn : g * p.kg + n_tile * p.bn;
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      for (int sb = half; sb < nblocks; sb += 2) {
        const uint32_t sbuf = sbuf0 + obuf * 16384;
        const uint32_t row_addr = sbuf + m * rowb;
        // Both 32-column TMEM loads of the block are issued BEFORE the buffer hand-over (bulk wait + barrier) and waited for once:
        // with N = 256 and a one-k-block main loop (1x1 convs c -> 4c, transposed convs) the tile time IS the epilogue
        // (ncu: 92 us at 9 % tensor pipe for c64 -> 256 @256^2), so TMEM latency must hide behind the barrier, not add to it.
        uint32_t v[64];
        tmem_ld_32x32(trow + sb * p.store_c, v);
        if (p.store_c == 64) tmem_ld_32x32(trow + sb * p.store_c + 32, v + 32);
        if (issuer) {  // the TMA store that last read THIS buffer has finished reading it
          if (p.out_bufs == 2) bulk_wait_read1();
          else bulk_wait_read0();
        }
        named_bar_sync(1 + half, 128);
        tmem_ld_wait();
#pragma unroll
        for (int pi = 0; pi < 2; ++pi) {
          const int part = pi * 32;
          if (part >= p.store_c) continue;
          const int col = sb * p.store_c + part;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float a = __uint_as_float(v[part + j + 2 * i]), b = __uint_as_float(v[part + j + 2 * i + 1]);
              if (p.bias) {
                a += p.bias[co0 + col + j + 2 * i];
                b += p.bias[co0 + col + j + 2 * i + 1];
              }
              if (p.ep_scale) {
                const int ch = co0 + col + j + 2 * i;
                a = apply_act(fmaf(a, p.ep_scale[ch], p.ep_shift[ch]), p.ep_act);
                b = apply_act(fmaf(b, p.ep_scale[ch + 1], p.ep_shift[ch + 1]), p.ep_act);
              }
              __nv_bfloat162 hp = __floats2bfloat162_rn(a, b);
              w4[i] = *reinterpret_cast<uint32_t*>(&hp);
            }
            const uint32_t chunk = (uint32_t)((part + j) >> 3);  // 16-byte chunk within the staged row
            st_shared_v4(row_addr + ((chunk ^ swz) << 4), w4[0], w4[1], w4[2], w4[3]);
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + half, 128);
        if (issuer) {
          const int cc = co0 + sb * p.store_c;
          if (p.convt) {  // (kw, k) of one kh is a contiguous row segment of the (n, 2h, 2w, k) output
            const int kh = cc / (2 * p.k_total);
            tma_store_4d(&p.map_out, sbuf, cc - kh * 2 * p.k_total, twi * p.tw, kh, img * p.h + thi * p.th);
          } else
            tma_store_2d(&p.map_out, sbuf, cc, ((img * p.h + thi * p.th) * p.w + twi * p.tw));
          bulk_commit();
        }
        if (p.stats) {
          // Column sums of the staged (rounded) tile WITHOUT atomics: warp wq of the half owns 8 (4) column pairs; its lanes
          // split the rows so that one LDS.32 touches every bank once (the row offsets below undo the TMA swizzle), shuffles
          // fold the row groups, and the owning lanes add into the CTA's statistics (each channel has exactly one owner).
          const int wq = (warp - 2) & 3;
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          int cp, nfold;
          if (p.store_c == 64) {
            const int cpl = lane & 7, rg = lane >> 3;  // 4 row groups: rows 2 rg + 8 i (+ 1)
            cp = wq * 8 + cpl;
            nfold = 2;
#pragma unroll
            for (int par = 0; par < 2; ++par) {
              const int r0 = 2 * rg + par;
              const uint32_t a0 = sbuf + r0 * 128 + ((((uint32_t)cp >> 2) ^ (uint32_t)(r0 & 7)) << 4) + ((cp & 3) << 2);
              uint32_t wv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv[i]) : "r"(a0 + i * 1024));
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float a = __uint_as_float(wv[i] << 16), b = __uint_as_float(wv[i] & 0xffff0000u);
                s0 += a;
                s1 += b;
                q0 = fmaf(a, a, q0);
                q1 = fmaf(b, b, q1);
              }
            }
          } else {
            const int cpl = lane & 3, rg = lane >> 2;  // 8 row groups: rows rg + 8 i
            cp = wq * 4 + cpl;
            nfold = 3;
            const uint32_t a0 = sbuf + rg * 64 + ((((uint32_t)cp >> 2) ^ (uint32_t)((rg >> 1) & 3)) << 4) + ((cp & 3) << 2);
            uint32_t wv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv[i]) : "r"(a0 + i * 512));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = __uint_as_float(wv[i] << 16), b = __uint_as_float(wv[i] & 0xffff0000u);
              s0 += a;
              s1 += b;
              q0 = fmaf(a, a, q0);
              q1 = fmaf(b, b, q1);
            }
          }
          for (int f = 0, o = 16; f < nfold; ++f, o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            q0 += __shfl_xor_sync(0xffffffffu, q0, o);
            q1 += __shfl_xor_sync(0xffffffffu, q1, o);
          }
          if (lane < (p.store_c == 64 ? 8 : 4)) {
            const uint32_t sa = stats_sm + 4 * (co0 + sb * p.store_c + 2 * cp);
            float v0, v1, v2, v3;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(sa));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v2), "=f"(v3) : "r"(sa + 4 * p.k_total));
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa), "f"(v0 + s0), "f"(v1 + s1) : "memory");
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa + 4 * p.k_total), "f"(v2 + q0), "f"(v3 + q1) : "memory");
          }
        }
        if (p.out_bufs == 2) obuf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) bulk_wait_all();
  } else if (warp >= 2 && warp < 6) {
    // ===================== epilogue (direct stores: fp32 output / odd N tiles) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int th_i = m >> p.tw_log2, tw_i = m & (p.tw - 1);
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int t = tile / p.n_tiles;
      const int g = t % p.groups;
      const int m_tile = t / p.groups;
      const int twi = m_tile % p.tiles_w;
      const int t2 = m_tile / p.tiles_w;
      const int thi = t2 % p.tiles_h;
      const int img = t2 / p.tiles_h;
      const int hh = thi * p.th + th_i, ww = twi * p.tw + tw_i;
      long long pix;
      int co0;
      if (p.convt) {
        const int nglob = n_tile * p.bn;
        const int tap2 = nglob / p.k_total;
        co0 = nglob - tap2 * p.k_total;
        pix = ((long long)img * (2 * p.h) + (2 * hh + (tap2 >> 1))) * (2 * p.w) + (2 * ww + (tap2 & 1));
      } else {
        co0 = g * p.kg + n_tile * p.bn;
        pix = ((long long)img * p.h + hh) * p.w + ww;
      }
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c0, v);
        tmem_ld_wait();
        const int valid = (p.bn - c0) < 32 ? (p.bn - c0) : 32;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < valid) v[j] = __float_as_uint(__uint_as_float(v[j]) + p.bias[co0 + c0 + j]);
        }
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + pix * p.ldo + co0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < valid) *reinterpret_cast<uint4*>(o + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + co0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < valid) {
              uint32_t w4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 hpair = __floats2bfloat162_rn(__uint_as_float(v[j + 2 * i]), __uint_as_float(v[j + 2 * i + 1]));
                w4[i] = *reinterpret_cast<uint32_t*>(&hpair);
              }
              *reinterpret_cast<uint4*>(o + j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    for (int i = threadIdx.x; i < 2 * p.k_total; i += blockDim.x) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(stats_sm + 4 * i));
      if (v != 0.f) atomicAdd(p.stats + i, (double)v);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient
struct alignas(64) WgParams {
  CUtensorMap map_x0, map_x1, map_dy;  // map_x*: tap-shifted operand (M side); map_dy: un-shifted operand (N side)
  int atom_x, atom_y;       // channels per MN atom (64 -> 128B swizzle, 32 -> 64B swizzle)
  int atoms0, atoms1;       // channel atoms per tap from source 0 / 1 (per group)
  int atoms_per_mtile;      // 128 / atom_x
  int taps, tap_s, pad, dil;
  int tw, th, tiles_w, tiles_h, h, w;
  int pix_tiles;            // n * tiles_h * tiles_w
  int m_tiles, n_tiles, groups, splits;
  int cg, kg, bn;
  int stages;
  int gather2x2;            // tap-shifted operand is the 5-D 2x2 gather view
  int ctot;                 // row length of dW per tap: channels of the shifted operand (all sources, per group)
  int pix;                  // pixels per pipeline stage (th * tw): 128, or 64 so that two CTAs (3 stages each) share an SM
  float* dw;
};

__global__ void __launch_bounds__(192, 2) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ax_bytes = (uint32_t)p.pix * p.atom_x * 2;  // one M atom box
  const uint32_t ay_bytes = (uint32_t)p.pix * p.atom_y * 2;  // one N atom box
  const uint32_t a_bytes = ax_bytes * p.atoms_per_mtile;  // pix * 128 channels
  const uint32_t n_atoms = p.bn / p.atom_y;
  const uint32_t b_bytes = ay_bytes * n_atoms;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 64, tfull0 = bar_base + 128;
  const uint32_t tmem_slot = bar_base + 160;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.map_x0);
    tma_prefetch_desc(&p.map_dy);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull0, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = p.bn <= 32 ? 32u : (p.bn <= 64 ? 64u : (p.bn <= 128 ? 128u : 256u));
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

  // one work unit per CTA: (n_tile, m_tile, group, split)
  int u = blockIdx.x;
  const int n_tile = u % p.n_tiles;
  u /= p.n_tiles;
  const int m_tile = u % p.m_tiles;
  u /= p.m_tiles;
  const int g = u % p.groups;
  const int split = u / p.groups;
  const int per = (p.pix_tiles + p.splits - 1) / p.splits;
  const int pt_begin = split * per;
  const int pt_end = min(p.pix_tiles, pt_begin + per);
  const int atoms_per_tap = p.atoms0 + p.atoms1;
  const int atoms_total = p.taps * atoms_per_tap;
  const int atom_begin = m_tile * p.atoms_per_mtile;
  const int atoms_here = min(p.atoms_per_mtile, atoms_total - atom_begin);

  if (warp == 0 && lane == 0) {
    uint32_t stage = 0, phase = 0;
    const uint32_t tx = ax_bytes * atoms_here + b_bytes;
    for (int pt = pt_begin; pt < pt_end; ++pt) {
      const int twi = pt % p.tiles_w;
      const int t2 = pt / p.tiles_w;
      const int thi = t2 % p.tiles_h;
      const int img = t2 / p.tiles_h;
      const int h0 = thi * p.th, w0 = twi * p.tw;
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
      const uint32_t fb = full0 + 8 * stage;
      mbar_expect_tx(fb, tx);
      for (int a = 0; a < atoms_here; ++a) {
        const int atom = atom_begin + a;
        const int tap = atom / atoms_per_tap, ai = atom - tap * atoms_per_tap;
        const int tr = tap / p.tap_s, ts = tap - tr * p.tap_s;
        if (p.gather2x2) {
          tma_load_5d(sa + a * ax_bytes, &p.map_x0, fb, ai * p.atom_x, ts, w0, tr, img * p.h + h0);
        } else if (ai < p.atoms0) {
          tma_load_4d(sa + a * ax_bytes, &p.map_x0, fb, g * p.cg + ai * p.atom_x, w0 + ts * p.dil - p.pad,
                      h0 + tr * p.dil - p.pad, img);
        } else {
          tma_load_4d(sa + a * ax_bytes, &p.map_x1, fb, (ai - p.atoms0) * p.atom_x, w0 + ts * p.dil - p.pad,
                      h0 + tr * p.dil - p.pad, img);
        }
      }
      for (uint32_t a = 0; a < n_atoms; ++a)
        tma_load_4d(sb + a * ay_bytes, &p.map_dy, fb, g * p.kg + n_tile * p.bn + a * p.atom_y, w0, h0, img);
      if (++stage == (uint32_t)p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // whole warp in the loop (warp-uniform addresses); one elected lane issues
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.bn, 1, 1);
    const uint32_t swz_x = p.atom_x * 2, swz_y = p.atom_y * 2;
    // MN-major: LBO = stride between MN atoms (one TMA box), SBO = stride between 8-pixel (K) groups
    const uint64_t ap = make_smem_desc(0, ax_bytes, 8 * swz_x, swz_x), bp = make_smem_desc(0, ay_bytes, 8 * swz_y, swz_y);
    const uint32_t a_hi = (uint32_t)(ap >> 32), b_hi = (uint32_t)(bp >> 32);
    const uint32_t kx = (16 * swz_x) >> 4, ky = (16 * swz_y) >> 4;  // 16 pixel rows per UMMA K step
    const int ksteps = p.pix >> 4;
    uint32_t stage = 0, phase = 0;
    for (int pt = pt_begin; pt < pt_end; ++pt) {
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = base + stage * stage_bytes;
        const uint32_t a_lo = (uint32_t)ap + (sa >> 4), b_lo = (uint32_t)bp + ((sa + a_bytes) >> 4);
#pragma unroll 4
        for (int k = 0; k < ksteps; ++k)  // pix pixels = pix / 16 UMMA K-steps of 16 rows
          umma_bf16_lohi(tmem_base, a_lo + k * kx, a_hi, b_lo + k * ky, b_hi, idesc, (uint32_t)((pt != pt_begin) | (k != 0)));
        umma_commit(empty0 + 8 * stage);
        if (pt == pt_end - 1) umma_commit(tfull0);
      }
      __syncwarp();
      if (++stage == (uint32_t)p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (pt_end <= pt_begin && elect_one()) mbar_arrive(tfull0);
    __syncwarp();
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int a = m / p.atom_x, cl = m - a * p.atom_x;
    const bool row_ok = (a < atoms_here) && (pt_end > pt_begin);
    const int atom = atom_begin + a;
    const int tap = atom / atoms_per_tap, ai = atom - tap * atoms_per_tap;
    const int c = ai * p.atom_x + cl;  // channel within the group's concatenated input
    mbar_wait(tfull0, 0);
    tc_fence_after();
    if (pt_end > pt_begin) {
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      const long long row_stride = (long long)p.taps * p.ctot;
      float* dst = p.dw + ((long long)(g * p.kg + n_tile * p.bn)) * row_stride + (long long)tap * p.ctot + c;
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c0, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < p.bn) atomicAdd(dst + (long long)(c0 + j) * row_stride, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
static bool spatial_tile(int h, int w, int* th, int* tw, int pix = 128) {
  if (w >= pix) {
    if (w % pix) return false;
    *tw = pix;
    *th = 1;
    return true;
  }
  if (w < 8 || (w & (w - 1))) return false;
  *tw = w;
  *th = pix / w;
  return h % *th == 0;
}
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static int encode_act_map(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int ld, int box_c, int tw,
                          int th) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)tw, (cuuint32_t)th, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d ld %d box %d,%d,%d) failed: %d", n, h, w, c, ld, box_c, tw,
              th, (int)r);
    return XV2_ECUDA;
  }
  return XV2_OK;
}
// 5-D view {C, 2, W, 2, H*N} of a (n, 2h, 2w, c) tensor: element (c, kw, w, kh, hn) = t[n, 2h+kh, 2w+kw, c]
static int encode_gather2x2_map(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int ld, int box_c, int tw,
                                int th) {
  cuuint64_t dims[5] = {(cuuint64_t)c, 2, (cuuint64_t)w, 2, (cuuint64_t)h * n};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)2 * ld * 2, (cuuint64_t)2 * w * ld * 2,
                           (cuuint64_t)4 * w * ld * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, 1, (cuuint32_t)tw, 1, (cuuint32_t)th};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2x2 gather view) failed: %d", (int)r);
    return XV2_ECUDA;
  }
  return XV2_OK;
}
static int encode_weight_map(CUtensorMap* m, const void* ptr, long long rows, long long kdim, int box_k, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights %lld x %lld box %d,%d) failed: %d", rows, kdim, box_k, box_rows, (int)r);
    return XV2_ECUDA;
  }
  return XV2_OK;
}

// shared with conv_strip.cu / wgrad_strip.cu
int strip_encode_act(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int ld, int box_c, int box_w) {
  return encode_act_map(m, ptr, n, h, w, c, ld, box_c, box_w, 1);
}
int strip_encode_weight(CUtensorMap* m, const void* ptr, long long rows, long long kdim, int box_k, int box_rows) {
  return encode_weight_map(m, ptr, rows, kdim, box_k, box_rows);
}
int tc_num_sms() { return g_num_sms; }
int conv_strip_launch(const xv2_tc_conv* q, const void* src0, const void* src1, const void* w, void* out, double* stats,
                      const float* ep_scale, const float* ep_shift, int ep_act, void* stream);
int wgrad_strip_launch(const xv2_tc_conv* q, const void* src0, const void* src1, const void* dout, int lddo, float* dw,
                       void* stream);
static bool strip_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("XV2_NO_STRIP");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// output tile maps for the TMA-store epilogue
static int encode_out_map(CUtensorMap* m, void* ptr, long long pixels, int k, int ldo, int box_c) {
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)pixels};
  cuuint64_t strides[1] = {(cuuint64_t)ldo * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(output %lld x %d ld %d) failed: %d", pixels, k, ldo, (int)r);
    return XV2_ECUDA;
  }
  return XV2_OK;
}
int strip_encode_out(CUtensorMap* m, void* ptr, long long pixels, int k, int ldo, int box_c) {
  return encode_out_map(m, ptr, pixels, k, ldo, box_c);
}
// transposed conv: the (kw, k) pair of an input pixel is CONTIGUOUS in the (n, 2h, 2w, k) output (pixels 2w, 2w+1 of one row),
// so the output is viewed as {2k (kw-major), w, 2 (kh), h*n} and a staged block is a full 64/128-byte row segment
static int encode_out_shuffle_map(CUtensorMap* m, void* ptr, int n, int h, int w, int k, int box_c, int tw, int th) {
  cuuint64_t dims[4] = {(cuuint64_t)2 * k, (cuuint64_t)w, 2, (cuuint64_t)h * n};
  cuuint64_t strides[3] = {(cuuint64_t)2 * k * 2, (cuuint64_t)2 * w * k * 2, (cuuint64_t)4 * w * k * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)tw, 1, (cuuint32_t)th};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(pixel-shuffle output) failed: %d", (int)r);
    return XV2_ECUDA;
  }
  return XV2_OK;
}

static int pick_bn(int kg) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (kg % bn == 0) return bn;
  return 0;
}

}  // namespace xv2

using namespace xv2;

extern "C" int xv2_init(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess || prop.major != 10) {
    set_error("device %d is not sm_100 (compute capability %d.%d)", device, prop.major, prop.minor);
    return XV2_ECUDA;
  }
  return ensure_init();
}

// p->gather flag is carried in `convt` = 2 (transposed-conv data gradient: src0 is (n, 2h, 2w, c0), taps 2x2)
static int conv_tc_impl(const xv2_tc_conv* q, const void* src0, const void* src1, const void* w, const float* bias,
                        void* out, double* stats, const float* ep_scale, const float* ep_shift, int ep_act, void* stream);

extern "C" int xv2_conv_tc(const xv2_tc_conv* q, const void* src0, const void* src1, const void* w,
                           const float* bias, void* out, double* stats, void* stream) {
  return conv_tc_impl(q, src0, src1, w, bias, out, stats, nullptr, nullptr, 0, stream);
}

extern "C" int xv2_conv_tc_bnact(const xv2_tc_conv* q, const void* src0, const void* src1, const void* w, const float* scale,
                                 const float* shift, int32_t act, void* out, void* stream) {
  XV2_REQUIRE(scale && shift, "conv_tc_bnact: null coefficients");
  return conv_tc_impl(q, src0, src1, w, nullptr, out, nullptr, scale, shift, act, stream);
}

static int conv_tc_impl(const xv2_tc_conv* q, const void* src0, const void* src1, const void* w, const float* bias,
                        void* out, double* stats, const float* ep_scale, const float* ep_shift, int ep_act, void* stream) {
  XV2_REQUIRE(q && src0 && w && out, "conv_tc: null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (!bias && strip_enabled()) {
    rc = conv_strip_launch(q, src0, src1, w, out, stats, ep_scale, ep_shift, ep_act, stream);
    if (rc != XV2_EUNSUPPORTED) return rc;
  }
  const int groups = q->groups < 1 ? 1 : q->groups;
  const int ld0 = q->ld0 ? q->ld0 : q->c0, ld1 = q->ld1 ? q->ld1 : q->c1;
  const int ctot = q->c0 + q->c1;
  const bool gather = q->convt == 2;
  const bool convt = q->convt == 1;
  int th, tw;
  if (!spatial_tile(q->h, q->w, &th, &tw) || ctot % groups || q->k % groups || (groups > 1 && q->c1) ||
      (ld0 % 8) || (q->c1 && (ld1 % 8)) || ((convt || gather) && (groups != 1 || q->c1))) {
    set_error("conv_tc: shape not eligible (h %d w %d c %d+%d k %d groups %d)", q->h, q->w, q->c0, q->c1, q->k, groups);
    return XV2_EUNSUPPORTED;
  }
  const int cg = ctot / groups, kg = q->k / groups;
  int bk = 0;
  if (q->c0 % 64 == 0 && q->c1 % 64 == 0 && cg % 64 == 0) bk = 64;
  else if (q->c0 % 32 == 0 && q->c1 % 32 == 0 && cg % 32 == 0) bk = 32;
  // transposed conv with the TMA-store epilogue: an N tile may span both kw taps of one kh (a contiguous output row segment)
  const bool convt_wide = convt && q->out_dtype == XV2_BF16 && !bias && (q->ldo == 0 || q->ldo == q->k);
  // ... or all four taps when they fit one accumulator (k <= 64): the input tile is then fetched once instead of twice
  int convt_n = convt_wide ? 2 * q->k : q->k;
  if (convt_wide && 4 * q->k <= 256 && q->k % 16 == 0) {
    const int sc = (4 * q->k > 64 && (4 * q->k) % 64 == 0) ? 64 : 32;  // staged block (see store_c below) must stay inside one kh
    if ((2 * q->k) % sc == 0) convt_n = 4 * q->k;
  }
  const int bn = pick_bn(convt ? convt_n : kg);
  if (!bk || !bn || (q->out_dtype != XV2_BF16 && q->out_dtype != XV2_F32)) {
    set_error("conv_tc: channels not eligible (c %d+%d k %d groups %d)", q->c0, q->c1, q->k, groups);
    return XV2_EUNSUPPORTED;
  }
  if (convt) XV2_REQUIRE(q->r == 1 && q->s == 1 && q->pad == 0, "conv_tc: convt expects r=s=1 (taps live in N)");
  if (gather) XV2_REQUIRE(q->r == 2 && q->s == 2 && q->pad == 0, "conv_tc: gather expects 2x2 taps");
  const int ldo = q->ldo ? q->ldo : q->k;
  XV2_REQUIRE(ldo % (q->out_dtype == XV2_F32 ? 4 : 8) == 0, "conv_tc: output stride %d not 16-byte aligned", ldo);

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.out_bufs = 1;
  if (gather) {
    rc = encode_gather2x2_map(&p.map_a0, src0, q->n, q->h, q->w, q->c0, ld0, bk, tw, th);
  } else {
    rc = encode_act_map(&p.map_a0, src0, q->n, q->h, q->w, q->c0, ld0, bk, tw, th);
    if (!rc && q->c1) rc = encode_act_map(&p.map_a1, src1, q->n, q->h, q->w, q->c1, ld1, bk, tw, th);
  }
  if (rc) return rc;
  const int taps = q->r * q->s;
  const long long brows = convt ? 4LL * q->k : q->k;
  rc = encode_weight_map(&p.map_b, w, brows, (long long)taps * cg, bk, bn);
  if (rc) return rc;
  p.kb0 = (groups > 1 ? cg : q->c0) / bk;
  p.kb1 = q->c1 / bk;
  p.taps = taps;
  p.tap_s = q->s;
  p.pad = q->pad;
  p.dil = q->dil < 1 ? 1 : q->dil;
  p.tw = tw;
  p.th = th;
  p.tw_log2 = ilog2(tw);
  p.tiles_w = q->w / tw;
  p.tiles_h = q->h / th;
  p.m_tiles = q->n * p.tiles_w * p.tiles_h;
  p.n_tiles = (convt ? 4 * q->k : kg) / bn;
  p.groups = groups;
  p.cg = cg;
  p.kg = kg;
  p.bn = bn;
  const bool tma_store = q->out_dtype == XV2_BF16 && bn % 32 == 0 && !(convt && (ldo != q->k || bias));
  if (stats && (!tma_store || convt || q->k > 2048)) {
    set_error("conv_tc: fused statistics epilogue is not available for this shape");
    return XV2_EUNSUPPORTED;
  }
  uint32_t extra = 0;
  if (tma_store) {
    p.tma_store = 1;
    p.store_c = (bn % 64 == 0 && bn > 64) ? 64 : 32;
    rc = convt ? encode_out_shuffle_map(&p.map_out, out, q->n, q->h, q->w, q->k, p.store_c, tw, th)
               : encode_out_map(&p.map_out, out, (long long)q->n * q->h * q->w, q->k, ldo, p.store_c);
    if (rc) return rc;
    p.out_bufs = (taps == 1) ? 2 : 1;
    extra = 1024 + 2 * p.out_bufs * 16384 + (stats ? 8u * q->k : 0u);
    p.stats = stats;
    p.ep_scale = ep_scale;
    p.ep_shift = ep_shift;
    p.ep_act = ep_act;
  } else if (ep_scale) {
    set_error("conv_tc_bnact: shape not served by the TMA-store epilogue");
    return XV2_EUNSUPPORTED;
  }
  const uint32_t stage_bytes = 128u * bk * 2 + (uint32_t)bn * bk * 2;
  int stages = (int)((kSmemBudget - extra) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  p.gather2x2 = gather ? 1 : 0;
  p.convt = convt ? 1 : 0;
  p.h = q->h;
  p.w = q->w;
  p.k_total = q->k;
  p.ldo = ldo;
  p.out_f32 = q->out_dtype == XV2_F32;
  p.out = out;
  p.bias = bias;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + extra;
  const long long total = (long long)p.m_tiles * p.n_tiles * groups;
  const int grid = (int)(total < g_num_sms ? total : g_num_sms);
  cudaStream_t st = as_stream(stream);
  cudaError_t e;
  if (bk == 64) {
    e = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = launch_pdl(conv_tc_kernel<64>, dim3(grid), dim3(320), smem, st, p);
  } else {
    e = cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = launch_pdl(conv_tc_kernel<32>, dim3(grid), dim3(320), smem, st, p);
  }
  if (e != cudaSuccess) {
    set_error("conv_tc: launch: %s", cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}

extern "C" int xv2_wgrad_tc(const xv2_tc_conv* q, const void* src0, const void* src1, const void* dout, int32_t lddo,
                            float* dw, void* stream) {
  XV2_REQUIRE(q && src0 && dout && dw, "wgrad_tc: null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (strip_enabled()) {
    rc = wgrad_strip_launch(q, src0, src1, dout, lddo, dw, stream);
    if (rc != XV2_EUNSUPPORTED) return rc;
  }
  const int groups = q->groups < 1 ? 1 : q->groups;
  const bool convt = q->convt == 1;
  // Stage geometry knobs, measured on B200 over the 36 1x1 weight gradients of a config-2 step (profiles/r02_wgrad_tc_variants.txt):
  // 128-pixel stages, N <= 128, one CTA per SM: 2.16 ms;  XV2_WG_PIX=64 (3 stages of 32 KB, two CTAs per SM so that one CTA's
  // prologue / atomic epilogue hides behind the other's main loop): 2.35 ms;  XV2_WG_PIX=64 XV2_WG_BN=256: 2.46 ms.
  static const int pix_pref = env_int("XV2_WG_PIX", 128), bn_max = env_int("XV2_WG_BN", 128);
  int th, tw, pix = pix_pref == 128 ? 128 : 64;
  if (!spatial_tile(q->h, q->w, &th, &tw, pix)) pix = 128;
  if (!spatial_tile(q->h, q->w, &th, &tw, pix) || (groups > 1 && q->c1) || (convt && (groups != 1 || q->c1))) {
    set_error("wgrad_tc: shape not eligible (h %d w %d)", q->h, q->w);
    return XV2_EUNSUPPORTED;
  }
  // roles: "shifted" operand (M side) and "plain" operand (N side)
  //   conv : shifted = src (c0 [+c1], tap shifts with zero fill), plain = dout (k)      -> dw[k][tap][c]
  //   convt: shifted = dout (n,2h,2w,k) through the 2x2 gather view,  plain = src0 (c0) -> dw[c0][tap][k]
  const int sh_c0 = convt ? q->k : q->c0, sh_c1 = convt ? 0 : q->c1;
  const int pl_c = convt ? q->c0 : q->k;
  const int sh_ld0 = convt ? (lddo ? lddo : q->k) : (q->ld0 ? q->ld0 : q->c0);
  const int sh_ld1 = q->ld1 ? q->ld1 : q->c1;
  const int pl_ld = convt ? (q->ld0 ? q->ld0 : q->c0) : (lddo ? lddo : q->k);
  const void* sh0 = convt ? dout : src0;
  const void* pl = convt ? src0 : dout;
  if ((sh_c0 + sh_c1) % groups || pl_c % groups || sh_ld0 % 8 || pl_ld % 8 || (sh_c1 && sh_ld1 % 8)) {
    set_error("wgrad_tc: channels not eligible");
    return XV2_EUNSUPPORTED;
  }
  const int cg = (sh_c0 + sh_c1) / groups, kg = pl_c / groups;
  int atom_x = 0, atom_y = 0;
  if (sh_c0 % 64 == 0 && sh_c1 % 64 == 0 && cg % 64 == 0) atom_x = 64;
  else if (sh_c0 % 32 == 0 && sh_c1 % 32 == 0 && cg % 32 == 0) atom_x = 32;
  if (kg % 64 == 0) atom_y = 64;
  else if (kg % 32 == 0) atom_y = 32;
  if (!atom_x || !atom_y) {
    set_error("wgrad_tc: channels not eligible (c %d+%d k %d groups %d)", q->c0, q->c1, q->k, groups);
    return XV2_EUNSUPPORTED;
  }
  int bn = 0;
  for (int cand = (bn_max == 256 ? 256 : 128); cand >= 32; cand -= 32)
    if (kg % cand == 0 && cand % atom_y == 0) {
      bn = cand;
      break;
    }
  if (!bn) {
    set_error("wgrad_tc: no N tile for k/group %d", kg);
    return XV2_EUNSUPPORTED;
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  if (convt) {
    rc = encode_gather2x2_map(&p.map_x0, sh0, q->n, q->h, q->w, sh_c0, sh_ld0, atom_x, tw, th);
  } else {
    rc = encode_act_map(&p.map_x0, sh0, q->n, q->h, q->w, sh_c0, sh_ld0, atom_x, tw, th);
    if (!rc && sh_c1) rc = encode_act_map(&p.map_x1, src1, q->n, q->h, q->w, sh_c1, sh_ld1, atom_x, tw, th);
  }
  if (!rc) rc = encode_act_map(&p.map_dy, pl, q->n, q->h, q->w, pl_c, pl_ld, atom_y, tw, th);
  if (rc) return rc;
  p.atom_x = atom_x;
  p.atom_y = atom_y;
  p.atoms0 = (groups > 1 ? cg : sh_c0) / atom_x;
  p.atoms1 = sh_c1 / atom_x;
  p.atoms_per_mtile = 128 / atom_x;
  p.taps = convt ? 4 : q->r * q->s;
  p.tap_s = convt ? 2 : q->s;
  p.pad = convt ? 0 : q->pad;
  p.dil = q->dil < 1 ? 1 : q->dil;
  p.tw = tw;
  p.th = th;
  p.tiles_w = q->w / tw;
  p.tiles_h = q->h / th;
  p.h = q->h;
  p.w = q->w;
  p.pix_tiles = q->n * p.tiles_w * p.tiles_h;
  const int atoms_total = p.taps * (p.atoms0 + p.atoms1);
  p.m_tiles = (atoms_total + p.atoms_per_mtile - 1) / p.atoms_per_mtile;
  p.n_tiles = kg / bn;
  p.groups = groups;
  p.cg = cg;
  p.kg = kg;
  p.bn = bn;
  p.gather2x2 = convt ? 1 : 0;
  p.ctot = cg;
  p.pix = pix;
  p.dw = dw;
  const long long units = (long long)p.m_tiles * p.n_tiles * groups;
  // split the pixel axis so that the grid fills at most TWO whole waves of one CTA per SM: rounding the split count UP (304 or
  // 320 CTAs for the layer3 / layer4 1x1 shapes, ncu row 80 of profiles/r02_ncu_kernels.md) left a third, nearly empty wave
  long long splits = (2LL * g_num_sms) / units;
  if (splits > p.pix_tiles) splits = p.pix_tiles;
  if (splits < 1) splits = 1;
  // re-balance so that no split is empty
  const int per = (p.pix_tiles + (int)splits - 1) / (int)splits;
  splits = (p.pix_tiles + per - 1) / per;
  p.splits = (int)splits;
  const uint32_t stage_bytes = (uint32_t)pix * 128 * 2 + (uint32_t)pix * bn * 2;
  int stages = (int)(kSmemBudget / stage_bytes);
  if (stages > 4) stages = 4;
  if (pix == 64 && bn <= 128 && stages > 3) stages = 3;  // <= 96 KB + barriers: two CTAs per SM
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  e = launch_pdl(wgrad_tc_kernel, dim3((unsigned)(units * splits)), dim3(192), smem, as_stream(stream), p);
  if (e != cudaSuccess) {
    set_error("wgrad_tc: launch: %s", cudaGetErrorString(e));
    return XV2_ECUDA;
  }
  XV2_LAUNCH_CHECK();
  return XV2_OK;
}
