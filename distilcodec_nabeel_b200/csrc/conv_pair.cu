// Fused ResBlock1 step for the C = 32 decoder stage with an fp32 activation stream in and out:
//
//   x' = conv2( silu( conv1( silu(x) ) + b1 ) ) + b2 + x          (models/convnext_utils.py:109-112)
//
// reads x (fp32, once, with its halo) and writes x' (fp32): 8 bytes per element and step instead of the 12 the
// s = silu(x) side buffer of conv_ws_pair_kernel costs (conv_ws.cu), and both convolutions run at twice the
// tensor-core N of that kernel.
//
// Why N matters here: an M = 128, N = 32, K = 16 tcgen05.mma reads 4 KB of activations + 1 KB of weights from shared
// memory for 65 k MACs; the shared-memory port (128 B/clk) makes it a ~45-cycle instruction where the math needs 16
// (profiles/r1_ncu_summary.md: l1tex data pipe 86 % busy).  The activation bytes per MMA do not depend on N, so the
// MACs per operand byte grow with N.  With C_out = 32 the only way to a larger N is to let one accumulator row stand for
// SEVERAL time steps:
//
//   PHASE FORM (u = 2).  View the (T, 32) bf16 activation tile as (T/2, 64): a 128-byte "super-row" q holds time rows
//   2q and 2q+1.  A K = 16 slice at byte offset 64*pi + 32*h of that 128B-swizzled tile is (rows 2q + pi, channels
//   16h..16h+15) for 128 consecutive q, and a start address m super-rows further down shifts it by 2m time rows: the
//   slice "offset o = 2m + pi" feeds tap o of output row 2v AND tap o-1 of output row 2v+1.  So
//       D[v, (r, co)] += sum_ci  A_o[v, ci] * Wt[(r, co), (o, ci)],      Wt[(r,co),(o,ci)] = W[co, ci, tap o - r]
//   is one M = 128, N = 64 MMA per (o, h): J+1 offsets x 2 slices = 2(J+1) MMAs for 256 output rows where the row
//   form needs 4J N = 32 MMAs.  Wt is the conv's block-Toeplitz weight, packed once at load time (dc_finalize).
//   The accumulator row v = (y[2v], y[2v+1]) is 64 contiguous fp32 of the (T, 32) output tensor, i.e. the epilogue is
//   the ordinary one on the (T/2, 64) view.
//
//   conv2 (dilation 1) always runs in phase form; conv1 does when its dilation is 1, else in row form (two M = 128,
//   N = 32 blocks per tile: odd dilations never pair two outputs on one activation slice).
//
// Roles (one CTA per SM, persistent over tiles of 256 t-rows -> 256 - (J-1) output rows):
//   warp 0        TMA: both weight sets, once
//   warp 1        MMA issuer: conv1(i+1) is issued before conv2(i)
//   8 warps       prep (fp32-stream form only): x rows (fp32, global, one batch of 16-byte loads per tile) -> silu ->
//                 bf16 -> the conv1 operand tile in the UMMA swizzled layout (zero outside the clip = conv1's padding);
//                 in the side-buffer form warp 0 loads the bf16 tile s = silu(x) by TMA instead
//   4 warps       epilogue 1: D1 -> + b1 -> silu -> bf16 -> t tile (128B-swizzled super-rows, zero outside the clip)
//   16 warps      epilogue 2: four groups of 4 warps, one tile in four each: D2 -> + b2 + x -> x' fp32 (+ silu(x') bf16
//                 in the side-buffer form), or for the last conv of the last branch the 3-branch mean -> silu -> bf16
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

#include "epilogue.cuh"

namespace dc {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

static thread_local uint64_t g_launches_px = 0;
uint64_t conv_pairx_launch_count() { return g_launches_px; }

namespace px {
constexpr int C = 32;
constexpr int TR = 256;                       // t rows per tile = 128 super-rows
#ifndef DC_PX_NG2
#define DC_PX_NG2 3   // epilogue-2 groups of the side-buffer form (A/B on one B200: 4 and 6 groups time the same)
#endif
#ifndef DC_PX_CW
// Epilogue-2 transpose chunk width of the side-buffer form.  32 columns = every global access of a warp is whole
// 128-byte lines; with 16 (64-byte half lines) ncu showed the LSU data pipe 82 % busy at 1.3 sectors per wavefront, which
// under the power-capped clock of the full step is what bounds the kernel (it is HBM-bound, 6.15 TB/s, at 1.75 GHz).
// A/B: 40.4-42.1 -> 37.8-38.3 ms per step.  The 4 KB per warp this needs is paid for with 3 groups and a 2-deep tile ring.
#define DC_PX_CW 32
#endif
constexpr int E1_WARPS = 4, G2W = 4;
constexpr int cw(int prep) { return prep ? 16 : DC_PX_CW; }
constexpr int ng2(int prep) { return prep ? 4 : DC_PX_NG2; }      // epilogue-2 groups (one D2 accumulator each)
constexpr int e2_warps(int prep) { return G2W * ng2(prep); }
constexpr int s_stages(int prep) { return (ng2(prep) > 4 || cw(prep) > 16) ? 2 : 3; }  // operand-tile ring depth (smem budget)
// PREP = 8: the operand tile is produced in the kernel from the fp32 stream (prep warps); PREP = 0: it is the bf16
// side buffer s = silu(x) of the previous epilogue, loaded by TMA (warp 0)
template <int PREP>
constexpr int threads() { return 64 + 32 * (PREP + E1_WARPS + e2_warps(PREP)); }
constexpr int S_ROWS = 320;                   // >= TR + (J-1)*dil = 306 (k = 11, dilation 5)
constexpr int S_BYTES = S_ROWS * C * 2;       // 20 KB, 1024-aligned
constexpr int T_BYTES = 136 * 128;            // 128 super-rows + the (J+1)/2 <= 6 the last offsets touch, 1024-aligned
constexpr int W_PH_TILE = 64 * C * 2;         // one offset of a phase-form weight: (2 x 32) rows x 32 ci
constexpr int W_ROW_TILE = C * C * 2;         // one tap of a row-form weight
constexpr int MAX_J = 11;
struct Layout {
  int w1_bytes, w2_bytes, s_off, t_off, stg_off, bias_off, bar_off, total;
};
static Layout layout(int J, bool ph1, int prep) {
  Layout l;
  const int S_STAGES = s_stages(prep), STG_BYTES = e2_warps(prep) * 32 * cw(prep) * 4;
  l.w1_bytes = ph1 ? (J + 1) * W_PH_TILE : J * W_ROW_TILE;
  l.w2_bytes = (J + 1) * W_PH_TILE;
  l.s_off = l.w1_bytes + l.w2_bytes;
  l.t_off = l.s_off + S_STAGES * S_BYTES;
  l.stg_off = l.t_off + 2 * T_BYTES;
  l.bias_off = l.stg_off + STG_BYTES;
  l.bar_off = l.bias_off + 256;
  l.total = l.bar_off + 256 + 1024;
  return l;
}
}  // namespace px

template <int PREP_WARPS>
__global__ void __launch_bounds__(px::threads<PREP_WARPS>(), 1)
conv_pairx_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmS /*PREP_WARPS == 0: the bf16 operand tensor*/,
                  const float* __restrict__ x, int T, int J, int dil, int ph1 /*conv1 in phase form*/,
                  const float* __restrict__ bias1, Epilogue ep, int variant, px::Layout lay, int tiles_per_clip,
                  int total_tiles, int dbg /*timing experiments only (DC_PAIRX_DBG): results are wrong when != 0*/) {
  using namespace px;
  constexpr int NG2 = ng2(PREP_WARPS), S_STAGES = s_stages(PREP_WARPS), CW2 = cw(PREP_WARPS);
  static_assert(threads<PREP_WARPS>() <= 1024 && 128 + NG2 * 64 <= 512, "thread / TMEM budget");
  constexpr uint32_t IDESC64 = ptx::make_idesc_bf16(128, 64), IDESC32 = ptx::make_idesc_bf16(128, 32);
  const int MO = TR - (J - 1), p1 = dil * (J - 1) / 2, p2 = (J - 1) / 2;
  const int rs = TR + (J - 1) * dil;        // rows of x one tile needs
  const int shift_a = -(p1 + p2);           // tile row 0 of x relative to the first output row

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + lay.w1_bytes;
  uint8_t* sS = smem + lay.s_off;
  uint8_t* sT = smem + lay.t_off;
  float* sb1 = reinterpret_cast<float*>(smem + lay.bias_off);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(smem + lay.bar_off);   // [S_STAGES] operand tile written by prep
  uint64_t* sempty = sfull + S_STAGES;                                  // [S_STAGES] ... consumed by conv1's MMAs
  uint64_t* d1full = sempty + S_STAGES;                                 // [2]
  uint64_t* d1empty = d1full + 2;                                       // [2]
  uint64_t* tfull = d1empty + 2;                                        // [2] t tile written by epilogue 1
  uint64_t* tempty = tfull + 2;                                         // [2] ... consumed by conv2's MMAs
  uint64_t* d2full = tempty + 2;                                        // [NG2]
  uint64_t* d2empty = d2full + NG2;                                     // [NG2]
  uint64_t* wbar = d2empty + NG2;                                       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < S_STAGES; ++i) {
        ptx::mbar_init(&sfull[i], PREP_WARPS > 0 ? PREP_WARPS : 1);
        ptx::mbar_init(&sempty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&d1full[i], 1);
        ptx::mbar_init(&d1empty[i], E1_WARPS);
        ptx::mbar_init(&tfull[i], E1_WARPS);
        ptx::mbar_init(&tempty[i], 1);
      }
      for (int i = 0; i < NG2; ++i) {
        ptx::mbar_init(&d2full[i], 1);
        ptx::mbar_init(&d2empty[i], G2W);
      }
      ptx::mbar_init(wbar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);   // D1[2] x 64 + D2[NG2] x 64 <= 512 columns
  }
  if (threadIdx.x < 64) sb1[threadIdx.x] = bias1[threadIdx.x & 31];   // conv1's bias for both rows of a super-row
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_d1 = tmem_base, tm_d2 = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA: the two weight sets, once per CTA
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, (uint32_t)(lay.w1_bytes + lay.w2_bytes));
      if (ph1) {
        for (int o = 0; o <= J; ++o) ptx::tma_load_2d(sW1 + o * W_PH_TILE, &tmW1, wbar, o * C, 0);
      } else {
        for (int j = 0; j < J; ++j) ptx::tma_load_2d(sW1 + j * W_ROW_TILE, &tmW1, wbar, j * C, 0);
      }
      for (int o = 0; o <= J; ++o) ptx::tma_load_2d(sW2 + o * W_PH_TILE, &tmW2, wbar, o * C, 0);
      if constexpr (PREP_WARPS == 0) {
        // the operand tiles: phase form = one box of rs/2 super-rows of the (T/2, 64) view (128B swizzle); row form =
        // two boxes of 160 rows of the (T, 32) tensor (64B swizzle; a box dimension is limited to 256).  Rows outside
        // the clip are zero-filled by TMA = conv1's zero padding.
        ptx::prefetch_tmap(&tmS);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const int clip = tile / tiles_per_clip, g0 = (tile % tiles_per_clip) * MO + shift_a;
          ptx::mbar_wait(&sempty[stage], phase ^ 1);
          uint8_t* dst = sS + stage * S_BYTES;
          if (ph1) {
            ptx::mbar_expect_tx(&sfull[stage], (uint32_t)((rs >> 1) * 128));
            ptx::tma_load_3d(dst, &tmS, &sfull[stage], 0, g0 >> 1, clip);   // g0 is even in phase form (dil == 1)
          } else {
            ptx::mbar_expect_tx(&sfull[stage], (uint32_t)S_BYTES);
            ptx::tma_load_3d(dst, &tmS, &sfull[stage], 0, g0, clip);
            ptx::tma_load_3d(dst + (S_ROWS / 2) * 64, &tmS, &sfull[stage], 0, g0 + S_ROWS / 2, clip);
          }
          if (++stage == S_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: conv1(0), then [conv1(i+1), conv2(i)]...
    if (ptx::elect_one()) {
      ptx::mbar_wait(wbar, 0);
      const uint32_t w1_addr = ptx::smem_u32(sW1), w2_addr = ptx::smem_u32(sW2);
      int n_my = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++n_my;
      const int Jm = (dbg & 32) ? 1 : J;  // timing experiment: one tap only
      int stage = 0;
      uint32_t phase = 0;
      auto conv1 = [&](int i) {
        const int b = i & 1;
        ptx::mbar_wait(&d1empty[b], ((i >> 1) & 1) ^ 1);
        ptx::mbar_wait(&sfull[stage], phase);
        ptx::tc_fence_after();
        const uint32_t s_addr = ptx::smem_u32(sS + stage * S_BYTES);
        if (ph1) {
          // offset o of the 128B-swizzled super-row tile: start = (o >> 1) super-rows + (o & 1) half rows
          const uint64_t da0 = ptx::make_smem_desc<128>(s_addr);
          const uint64_t dw0 = ptx::make_smem_desc<64>(w1_addr);
          uint32_t accum = 0;
          for (int o = 0; o <= Jm; ++o) {
            const uint64_t da = da0 + (uint64_t)(((o >> 1) * 128 + (o & 1) * 64) >> 4);
            const uint64_t dw = dw0 + (uint64_t)((o * W_PH_TILE) >> 4);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              ptx::mma_bf16_ss(tm_d1 + b * 64, da + 2 * h, dw + 2 * h, IDESC64, accum);
              accum = 1;
            }
          }
        } else {
          // row form: t rows [128 mb, 128 mb + 128) of the tile, tap j reads x rows + j*dil (64B-swizzled 64-byte rows)
          const uint64_t dw0 = ptx::make_smem_desc<64>(w1_addr);
          for (int mb = 0; mb < 2; ++mb) {
            uint64_t da = ptx::make_smem_desc<64>(s_addr + mb * 128 * 64);
            uint64_t dw = dw0;
            uint32_t accum = 0;
            for (int j = 0; j < Jm; ++j) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                ptx::mma_bf16_ss(tm_d1 + b * 64 + mb * 32, da + 2 * h, dw + 2 * h, IDESC32, accum);
                accum = 1;
              }
              da += (uint64_t)((dil * 64) >> 4);
              dw += W_ROW_TILE >> 4;
            }
          }
        }
        ptx::mma_commit(&sempty[stage]);
        ptx::mma_commit(&d1full[b]);
        if (++stage == S_STAGES) { stage = 0; phase ^= 1; }
      };
      auto conv2 = [&](int i) {
        const int b = i & 1;                 // t buffer
        const int g = i % NG2;               // D2 accumulator / epilogue-2 group
        ptx::mbar_wait(&d2empty[g], ((uint32_t)(i / NG2) & 1u) ^ 1u);
        ptx::mbar_wait(&tfull[b], (i >> 1) & 1);
        ptx::tc_fence_after();
        const uint64_t dt0 = ptx::make_smem_desc<128>(ptx::smem_u32(sT + b * T_BYTES));
        const uint64_t dw0 = ptx::make_smem_desc<64>(w2_addr);
        uint32_t accum = 0;
        for (int o = 0; o <= Jm; ++o) {
          const uint64_t dt = dt0 + (uint64_t)(((o >> 1) * 128 + (o & 1) * 64) >> 4);
          const uint64_t dw = dw0 + (uint64_t)((o * W_PH_TILE) >> 4);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            ptx::mma_bf16_ss(tm_d2 + g * 64, dt + 2 * h, dw + 2 * h, IDESC64, accum);
            accum = 1;
          }
        }
        ptx::mma_commit(&tempty[b]);
        ptx::mma_commit(&d2full[g]);
      };
      if (n_my > 0) conv1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) conv1(i + 1);
        conv2(i);
      }
    }
  } else if (warp < 2 + PREP_WARPS) {
    if constexpr (PREP_WARPS > 0) {
    // ------------------------------------------------------------ prep: x (fp32, global) -> silu -> bf16 operand tile
    const int tid = (warp - 2) * 32 + lane;            // 0..127
    constexpr int NT = PREP_WARPS * 32, BATCH = (320 * 8 + NT - 1) / NT;   // one batch of loads per tile
    const int n_f4 = rs * (C / 4);                     // float4 items of one tile: row = idx >> 3, channels 4*(idx & 7)..
    int i = 0, stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      const int clip = tile / tiles_per_clip, o0 = (tile % tiles_per_clip) * MO;
      ptx::mbar_wait(&sempty[stage], phase ^ 1);
      uint8_t* dst = sS + stage * S_BYTES;
      const float* xc = x + (size_t)clip * T * C;
      const int g0 = o0 + shift_a;
      for (int base = 0; base < n_f4; base += NT * BATCH) {
        float4 v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const int idx = base + k * NT + tid;
          const int gs = g0 + (idx >> 3);
          v[k] = (idx < n_f4 && gs >= 0 && gs < T && !(dbg & 1))
                     ? __ldg(reinterpret_cast<const float4*>(xc + (size_t)gs * C) + (idx & 7))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const int idx = base + k * NT + tid;
          if (idx < n_f4) {
            const int r = idx >> 3, c4 = idx & 7;
            float2 lo = make_float2(v[k].x, v[k].y), hi = make_float2(v[k].z, v[k].w);
            if (!(dbg & 2)) {
              lo = silu_fast2(lo);
              hi = silu_fast2(hi);
            }
            __nv_bfloat162 a = __floats2bfloat162_rn(lo.x, lo.y), b = __floats2bfloat162_rn(hi.x, hi.y);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&a);
            pk.y = *reinterpret_cast<uint32_t*>(&b);
            const int chunk = c4 >> 1, sub = (c4 & 1) * 8;
            int off;
            if (ph1) {  // 128B swizzle over super-rows: 16-byte chunk index (row parity, chunk) ^ (super-row & 7)
              const int q = r >> 1;
              off = q * 128 + (((((r & 1) << 2) | chunk) ^ (q & 7)) << 4) + sub;
            } else {    // 64B swizzle over 64-byte rows: chunk ^ ((row >> 1) & 3)
              off = r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4) + sub;
            }
            if (!(dbg & 64)) *reinterpret_cast<uint2*>(dst + off) = pk;
          }
        }
      }
      ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sfull[stage]);
      if (++stage == S_STAGES) { stage = 0; phase ^= 1; }
    }
    }
  } else if (warp < 2 + PREP_WARPS + E1_WARPS) {
    // ------------------------------------------------------------ epilogue 1: D1 -> silu(. + b1) -> bf16 t tile in smem
    const int q4 = warp & 3;                            // TMEM lane quarter of this hardware warp
    const int L = q4 * 32 + lane;                       // accumulator row
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      const int o0 = (tile % tiles_per_clip) * MO;
      const int b = i & 1;
      ptx::mbar_wait_sleepy(&d1full[b], (i >> 1) & 1);
      ptx::tc_fence_after();
      ptx::mbar_wait(&tempty[b], ((i >> 1) & 1) ^ 1);   // conv2 of tile i-2 has finished reading this t buffer
      uint8_t* tb = sT + b * T_BYTES;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        // phase form: row L = super-row L, columns 32c.. = t row 2L + c;  row form: block c, t row 128c + L
        const int trow = ph1 ? 2 * L + c : 128 * c + L;
        const int gt = o0 - p2 + trow;                  // sequence position of this t row
        const bool inside = gt >= 0 && gt < T;
        const int q = trow >> 1, par = trow & 1;
        uint32_t acc[32];
        ptx::tmem_ld_32x32(tm_d1 + b * 64 + ((uint32_t)(q4 * 32) << 16) + c * 32, acc);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {                   // 8 channels = one 16-byte chunk
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 bb = *reinterpret_cast<const float2*>(sb1 + g * 8 + 2 * e);
            float2 v = fadd2(make_float2(__uint_as_float(acc[g * 8 + 2 * e]), __uint_as_float(acc[g * 8 + 2 * e + 1])), bb);
            if (!(dbg & 4)) v = silu_fast2(v);
            if (!inside) v = make_float2(0.f, 0.f);
            __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(tb + q * 128 + ((((par << 2) | g) ^ (q & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&tfull[b]);
        ptx::mbar_arrive(&d1empty[b]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue 2 on the (T/2, 64) view: D2 -> global
    const int e2w = warp - (2 + PREP_WARPS + E1_WARPS);   // 0..15
    const int group = e2w / G2W;
    // warp id as epilogue_tile expects: (wg & 3) must be the hardware warp's TMEM lane quarter
    const int wg = 2 + ((warp - 2) & 3);
    float* stg = reinterpret_cast<float*>(smem + lay.stg_off) + e2w * (32 * CW2);
    Epilogue epf = ep;        // prefetch only what prep has not just pulled into L2 (the mean's two other branches)
    if (PREP_WARPS > 0) epf.res = nullptr;
    const int T2 = T >> 1;
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      if (i % NG2 != group) continue;      // group g owns accumulator D2[g]
      const int clip = tile / tiles_per_clip, o0 = (tile % tiles_per_clip) * MO;
      const int v0 = o0 >> 1, vlim = (o0 + MO) >> 1;
      epilogue_prefetch(epf, clip, T2, v0 + (wg & 3) * 32, 0, 64, lane, vlim);
      ptx::mbar_wait_sleepy(&d2full[group], (uint32_t)(i / NG2) & 1u);
      ptx::tc_fence_after();
      if (!(dbg & 8)) {
        epilogue_tile<64, CW2>(ep, variant, stg, tm_d2 + group * 64, clip, v0, 0, T2, wg, lane, vlim);
        epilogue_tile<64, CW2>(ep, variant, stg, tm_d2 + group * 64, clip, v0, 0, T2, wg + 4, lane, vlim);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&d2empty[group]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// s1: conv1 (dilated), s2: conv2 (dilation 1), both 32 -> 32 with the same odd kernel size; T even
bool conv_pairx_supported(const ConvGemmShape& s1, const ConvGemmShape& s2) {
  if (!(s1.C == 32 && s1.N == 32 && s2.C == 32 && s2.N == 32)) return false;
  if (s1.J != s2.J || s2.dil != 1 || s1.J < 3 || s1.J > px::MAX_J || (s1.J & 1) == 0) return false;
  if (s1.shift0 != -s1.dil * (s1.J - 1) / 2 || s2.shift0 != -(s2.J - 1) / 2) return false;
  if (s1.B != s2.B || s1.T != s2.T || (s1.T & 1)) return false;
  if (px::TR + (s1.J - 1) * s1.dil > px::S_ROWS) return false;
  return px::layout(s1.J, s1.dil == 1, 0).total <= 232448 && px::layout(s1.J, s1.dil == 1, 8).total <= 232448;
}

// W1: conv1 weight, phase form [64][(J+1)*32] when s1.dil == 1, else row form [32][J*32]; W2p: conv2 phase form;
// bias2x: conv2's bias repeated twice (64 floats, the two rows of a super-row).  Exactly one of X (fp32 stream form)
// and S (bf16 side-buffer form: S = silu(x), the residual x comes through e2.res) is given.
template <int PREP>
static int launch_px(const float* X, const __nv_bfloat16* S, const __nv_bfloat16* W1, const __nv_bfloat16* W2p,
                     const float* bias1, const float* bias2x, const ConvGemmShape& s1, const Epilogue& e2,
                     cudaStream_t st, int sm_count) {
  using namespace px;
  const int J = s1.J;
  const bool ph1 = s1.dil == 1;
  const Layout lay = layout(J, ph1, PREP);
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_pairx_kernel<PREP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int MO = TR - (J - 1);
  const int tiles_per_clip = (s1.T + MO - 1) / MO;
  const long long total = (long long)s1.B * tiles_per_clip;
  DC_CHECK(total > 0 && total < (1ll << 31), DC_ERR_SHAPE, "conv_pairx: bad tile count");
  CUtensorMap tmW1, tmW2, tmS;
  {
    const uint64_t K = (uint64_t)(ph1 ? J + 1 : J) * C;
    const uint64_t dims[2] = {K, (uint64_t)(ph1 ? 64 : 32)};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)C, (uint32_t)(ph1 ? 64 : 32)};
    DC_TRY(make_tmap_bf16(&tmW1, W1, 2, dims, strides, box, 64));
  }
  {
    const uint64_t K = (uint64_t)(J + 1) * C;
    const uint64_t dims[2] = {K, 64};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)C, 64};
    DC_TRY(make_tmap_bf16(&tmW2, W2p, 2, dims, strides, box, 64));
  }
  if (PREP == 0) {
    const int rs = TR + (J - 1) * s1.dil;
    if (ph1) {  // (B, T/2, 64) view, one box of rs/2 super-rows
      const uint64_t dims[3] = {64, (uint64_t)(s1.T / 2), (uint64_t)s1.B};
      const uint64_t strides[2] = {128, (uint64_t)s1.T * C * 2};
      const uint32_t box[3] = {64, (uint32_t)(rs / 2), 1};
      DC_TRY(make_tmap_bf16(&tmS, S, 3, dims, strides, box, 128));
    } else {    // (B, T, 32), two boxes of S_ROWS / 2 rows
      const uint64_t dims[3] = {(uint64_t)C, (uint64_t)s1.T, (uint64_t)s1.B};
      const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)s1.T * C * 2};
      const uint32_t box[3] = {(uint32_t)C, (uint32_t)(S_ROWS / 2), 1};
      DC_TRY(make_tmap_bf16(&tmS, S, 3, dims, strides, box, 64));
    }
  } else {
    tmS = tmW2;  // unused
  }
  Epilogue e = e2;           // the epilogue runs on the (T/2, 64) view of the (T, 32) tensors
  e.bias = bias2x;
  e.ldo = 64;
  const int grid = (int)(total < sm_count ? total : sm_count);
  {
    const double rows = (double)s1.B * s1.T;
    const double macs = 2.0 * rows * C * J * C;
    const int esig = (e.res ? 4 : 0) | (e.add1 ? 8 : 0) | (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) | (e.out1 ? 32 : 0);
    // algorithmic bytes: the input once (in the fp32 stream form the residual is the same tensor), the outputs, the
    // mean's two other operands
    const double out_bytes = (e.out0 ? (e.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e.out1 ? 2.0 : 0.0) +
                             (e.add1 ? 8.0 : 0.0) + (PREP == 0 && e.res ? 4.0 : 0.0);
    ProfScope ps(PC_CONV_WS, 2.0 * macs, rows * C * (PREP ? 4.0 : 2.0) + 2.0 * J * C * C * 2.0 + rows * C * out_bytes, st,
                 PREP ? "_pairx|C%d N%d J%d d%d e%d" : "_pairs|C%d N%d J%d d%d e%d", C, C, J, s1.dil, esig);
    static const int dbg = getenv("DC_PAIRX_DBG") ? atoi(getenv("DC_PAIRX_DBG")) : 0;
    conv_pairx_kernel<PREP><<<grid, threads<PREP>(), lay.total, st>>>(tmW1, tmW2, tmS, X, s1.T, J, s1.dil, ph1 ? 1 : 0,
                                                                      bias1, e, epilogue_variant(e), lay, tiles_per_clip,
                                                                      (int)total, dbg);
  }
  ++g_launches_px;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int launch_conv_pairx(const float* X, const __nv_bfloat16* S, const __nv_bfloat16* W1, const __nv_bfloat16* W2p,
                      const float* bias1, const float* bias2x, const ConvGemmShape& s1, const ConvGemmShape& s2,
                      const Epilogue& e2, cudaStream_t st, int sm_count) {
  DC_CHECK(conv_pairx_supported(s1, s2), DC_ERR_SHAPE, "conv_pairx: unsupported shapes");
  DC_CHECK((X != nullptr) != (S != nullptr) && W1 && W2p && bias1 && bias2x, DC_ERR_ARG, "conv_pairx: bad operands");
  const void* in = X ? (const void*)X : (const void*)S;
  DC_CHECK(e2.out0 != in && e2.out1 != in, DC_ERR_ARG, "conv_pairx: an output aliases the (halo-read) input");
  if (X) return launch_px<8>(X, nullptr, W1, W2p, bias1, bias2x, s1, e2, st, sm_count);
  return launch_px<0>(nullptr, S, W1, W2p, bias1, bias2x, s1, e2, st, sm_count);
}

}  // namespace dc
