// fp32-accurate shifted-row implicit GEMM on the TENSOR cores (DC_MODE_FP32 with option "fp32_tc"), replacing the CUDA-core
// kernel of gemm_f32.cu for every dense layer of the reference's default numeric mode (enable_bfloat16=False,
// distil_codec.py:545,581; the layers: models/encoders.py:23-37, convnext_utils.py:250-255, grfvq.py:68-96,
// residual_vq.py:61-62, generators.py:50-114, convnext_utils.py:36-102).
//
//   out[b,t,n] = epi( sum_j sum_c A[b, t + shift0 + j*dil, c] * W[n, j*C + c] )          A, W, out: fp32 values
//
// Both operands are split into two bf16 terms, x = hi + mid + r with |r| <= 2^-17 |x| (split_f32_kernel for the
// activations, pack time for the weights), and the three largest cross products are accumulated:
//   A.W ~ Ahi.Whi + Ahi.Wmid + Amid.Whi          (the dropped terms are <= 2^-16 relative per product, random in sign)
// Products of bf16 numbers are exact in the tensor core's fp32 accumulator; what is NOT exact is its accumulation over
// long K: round 1 measured 0.5-2.5e-5 of range per layer growing with K (scripts/split_probe.py), and the end-to-end probe
// of round 2 (scripts/fp32x_e2e_probe.py) 5.6e-5 (3 terms) / 1.1e-4 (6 terms: longer K, worse) on the stress weights —
// against the 1e-4 gate.  So the accumulation is CHUNKED: the MMAs of at most kChunkStages pipeline stages (K <= 256 per
// term) go into a partial accumulator P, and the epilogue warps add every finished P into the tile's total T with
// round-to-nearest fp32 adds (TMEM -> registers -> TMEM), which is the summation the CUDA-core kernel does.
//
// Per stage ONE box of Ahi, Amid, Whi, Wmid is loaded (64 KB) and feeds the three products, i.e. 12 MMAs.
//   warp 0 TMA, warp 1 MMA issuer, warps 2..9 chunk adders + epilogue (shared with gemm_tc.cu: epilogue.cuh)
// TMEM: T (BN columns) + kPBufs partial accumulators (BN columns each).
#include "common.cuh"
#include "ptx.cuh"

#include <string.h>

#include "epilogue.cuh"

namespace dc {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

static thread_local uint64_t g_launches_f32x = 0;
uint64_t gemm_f32x_launch_count() { return g_launches_f32x; }

// ---------------------------------------------------------------- operand split: fp32 (rows, C) -> bf16 (rows, [hi | mid])
__global__ void __launch_bounds__(256) split_f32_kernel(const float4* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                        size_t n4 /*rows * C / 4*/, int C4 /*C / 4*/) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    const size_t row = i / C4;
    const int c4 = (int)(i % C4);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], mid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hi[k] = __float2bfloat16_rn(x[k]);
      mid[k] = __float2bfloat16_rn(x[k] - __bfloat162float(hi[k]));   // the subtraction is exact (Sterbenz)
    }
    __nv_bfloat16* o = out + row * (size_t)(8 * C4) + 4 * c4;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(o + 4 * C4) = *reinterpret_cast<const uint2*>(mid);
  }
}

int launch_split_f32(const float* in, __nv_bfloat16* out, size_t rows, int C, cudaStream_t st) {
  DC_CHECK(C % 4 == 0, DC_ERR_SHAPE, "split_f32: C=%d must be a multiple of 4", C);
  const size_t n4 = rows * (size_t)(C / 4);
  if (n4 == 0) return DC_OK;
  ProfScope ps(PC_CAST, 0, (double)rows * C * 8.0, st, "split");
  const unsigned grid = (unsigned)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  split_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), out, n4, C / 4);
  ++g_launches_f32x;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ---------------------------------------------------------------- the kernel
#ifndef DC_F32X_OCC
#define DC_F32X_OCC 2
#endif
#ifndef DC_F32X_PREFETCH
#define DC_F32X_PREFETCH 1
#endif
// CTA pairs with weight multicast for the 128-wide tiles.  Measured (64 clips, fp32 mode, `--opt cta_pairs=1|2`, two
// alternating runs): 302.0 / 304.4 ms without, 307.2 / 303.4 ms with, identical codes — no gain, so off by default.  The
// 128 x 128 x 16 MMAs of this kernel read 8 KB of shared-memory operands per 64 tensor cycles, the whole 128 B/clk port,
// and the TMA fills of the stages take another 85 B/clk of it: the shared-memory port bounds the kernel, not L2.
#ifndef DC_F32X_PAIRS
#define DC_F32X_PAIRS 0
#endif
#ifndef DC_F32X_WS            // weights-stationary kernel for the C = N = 32 / 64 layers
#define DC_F32X_WS 1
#endif
#ifndef DC_F32X_NARROW_BK32   // N <= 64 layers with C % 64 == 0 on the BK = 32 configuration (two CTAs per SM)
#define DC_F32X_NARROW_BK32 1
#endif
namespace f32x {
constexpr int kChunkStages = 4;   // pipeline stages (of BK = 64: K = 256 per term) per partial accumulation chain
constexpr int kStages = 3;
template <int BN>
struct Cfg {
  static constexpr int PBUFS = BN == 128 ? 3 : 2;                    // partial accumulators: T + PBUFS * BN <= 512 columns
  static constexpr int TMEM_COLS = (1 + PBUFS) * BN <= 64 ? 64 : ((1 + PBUFS) * BN <= 128 ? 128 : ((1 + PBUFS) * BN <= 256 ? 256 : 512));
  static constexpr int CW = BN >= 64 ? 32 : 16;
};
template <int BN, int BK>
struct Smem {
  static constexpr int A_BYTES = 128 * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;       // Ahi, Amid, Whi, Wmid
  static constexpr int STG_OFF = kStages * STAGE_BYTES;
  static constexpr int STG_BYTES = 8 * 32 * Cfg<BN>::CW * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzled tiles need 1024-byte aligned bases");
  static_assert(TOTAL <= 232448, "shared memory budget");
  // narrow tiles (N <= 64, BK = 32) fit twice per SM: their per-tile chain TMA -> MMA chunks -> adds -> epilogue with
  // its DRAM round trips is latency-bound with one CTA per SM (one T accumulator: the epilogue of tile i and the adds
  // of tile i + 1 are the same warps)
  static constexpr int CTAS_PER_SM = (DC_F32X_OCC >= 2 && 2 * (TOTAL + 1024) <= 233472) ? 2 : 1;
};
}  // namespace f32x

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// T[rows of this warp, its column half] (+)= P, 32 columns at a time (16 for the narrowest tiles); tT / tP already
// carry the warp's TMEM lane offset
template <int BN>
__device__ __forceinline__ void f32x_add_partial(uint32_t tT, uint32_t tP, int half, bool first) {
  constexpr int PIECE = (BN / 2) >= 32 ? 32 : 16;
#pragma unroll 1
  for (int c = 0; c < (BN / 2) / PIECE; ++c) {
    const uint32_t col = half * (BN / 2) + c * PIECE;
    if constexpr (PIECE == 32) {
      uint32_t p[32], t[32];
      ptx::tmem_ld_32x32(tP + col, p);
      if (!first) ptx::tmem_ld_32x32(tT + col, t);
      ptx::tmem_ld_wait();
      if (!first) {
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i] = __float_as_uint(__fadd_rn(__uint_as_float(t[i]), __uint_as_float(p[i])));
      }
      tmem_st_32x32(tT + col, p);
    } else {
      uint32_t p[16], t[16];
      ptx::tmem_ld_32x16(tP + col, p);
      if (!first) ptx::tmem_ld_32x16(tT + col, t);
      ptx::tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        o[i] = first ? p[i] : __float_as_uint(__fadd_rn(__uint_as_float(t[i]), __uint_as_float(p[i])));
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
          "%14, %15, %16};" ::"r"(tT + col),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]),
          "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15])
          : "memory");
    }
  }
}

__device__ __forceinline__ uint32_t f32x_zero_taps(const ConvGemmShape& s, int n0, int bn) {
  if (s.zero_taps == 0 || s.phase_cols % bn != 0) return 0;
  return (s.zero_taps >> ((n0 / s.phase_cols) * s.J)) & ((1u << s.J) - 1u);
}

// CL = 2 (BN = 128): thread-block clusters of two CTAs on ADJACENT row tiles of the same N block: each CTA loads one half
// (64 of the 128 rows) of Whi and Wmid and TMA-multicasts it into both CTAs' stages, as in gemm_tc / conv_tsw.  The wide
// fp32 layers re-load a 64 KB stage per 12 MMAs = 85 B per clock and SM (10.9 TB/s of L2 reads on k = 11, C = 256); the pair
// takes 16 of those 64 KB off L2.  See DC_F32X_PAIRS for what it bought.
template <int BN, int BK, int CL>
__global__ void __launch_bounds__(320, CL == 2 ? 1 : f32x::Smem<BN, BK>::CTAS_PER_SM)
gemm_f32x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvGemmShape s,
                 Epilogue ep, int variant, int tiles_per_clip, int m_tiles, int n_tiles) {
  using namespace f32x;
  using L = Smem<BN, BK>;
  constexpr int SW = BK * 2, PBUFS = Cfg<BN>::PBUFS, TMEM_COLS = Cfg<BN>::TMEM_COLS, CW = Cfg<BN>::CW;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, BN);
  static_assert(CL == 1 || BN == 128, "weight multicast splits the 128 weight rows into two boxes of 64");
  const int rank = CL == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int worker = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_items = CL == 2 ? (m_tiles + 1) / 2 : m_tiles;
  auto item_m_blk = [&](int item) { return CL == 2 ? 2 * (item / n_tiles) + rank : item / n_tiles; };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);   // [kStages]
  uint64_t* empty = full + kStages;                                   // [kStages]
  uint64_t* pfull = empty + kStages;                                  // [PBUFS] partial accumulator complete
  uint64_t* pempty = pfull + PBUFS;                                   // [PBUFS] ... added into T by all 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + PBUFS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        ptx::mbar_init(&full[i], 1);
        ptx::mbar_init(&empty[i], CL);   // released by the MMAs of every CTA the stage's weight boxes are multicast to
      }
      for (int i = 0; i < PBUFS; ++i) {
        ptx::mbar_init(&pfull[i], 1);
        ptx::mbar_init(&pempty[i], 8);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();   // the peer's barriers exist before anything is multicast to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_T = tmem_base, tm_P = tmem_base + BN;

  const int total_tiles = m_items * n_tiles;
  const int kchunks = s.C / BK;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: one box of each of the four operands
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int m_blk = item_m_blk(tile), n_blk = tile % n_tiles;
        const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
        const uint32_t skip = f32x_zero_taps(s, n0, BN);
        for (int j = 0; j < s.J; ++j) {
          if ((skip >> j) & 1u) continue;  // all-zero tap of this stride phase (ConvTranspose1d)
          const int trow = t0 + s.shift0 + j * s.dil;
          for (int kc = 0; kc < kchunks; ++kc) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], L::STAGE_BYTES);
            uint8_t* sp = smem + stage * L::STAGE_BYTES;
            ptx::tma_load_3d(sp, &tmA, &full[stage], kc * BK, trow, clip);                                  // A hi
            ptx::tma_load_3d(sp + L::A_BYTES, &tmA, &full[stage], s.C + kc * BK, trow, clip);               // A mid
            if constexpr (CL == 2) {   // this CTA's 64 rows of W hi and W mid, into both CTAs
              ptx::tma_load_2d_multicast(sp + 2 * L::A_BYTES + rank * (L::B_BYTES / 2), &tmW, &full[stage],
                                         j * 2 * s.C + kc * BK, n0 + rank * (BN / 2), (uint16_t)3);
              ptx::tma_load_2d_multicast(sp + 2 * L::A_BYTES + L::B_BYTES + rank * (L::B_BYTES / 2), &tmW, &full[stage],
                                         j * 2 * s.C + s.C + kc * BK, n0 + rank * (BN / 2), (uint16_t)3);
            } else {
              ptx::tma_load_2d(sp + 2 * L::A_BYTES, &tmW, &full[stage], j * 2 * s.C + kc * BK, n0);           // W hi
              ptx::tma_load_2d(sp + 2 * L::A_BYTES + L::B_BYTES, &tmW, &full[stage], j * 2 * s.C + s.C + kc * BK, n0);  // W mid
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: chunks of <= kChunkStages stages -> P[c % PBUFS]
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int chunk = 0;   // running chunk counter over all tiles of this CTA (the adders count the same way)
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int n0m = (tile % n_tiles) * BN;
        const int live = (s.J - __popc(f32x_zero_taps(s, n0m, BN))) * kchunks;  // stages of this tile
        for (int st0 = 0; st0 < live; st0 += kChunkStages, ++chunk) {
          const int pb = chunk % PBUFS;
          ptx::mbar_wait(&pempty[pb], ((uint32_t)(chunk / PBUFS) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tm_P + pb * BN;
          const int st1 = st0 + kChunkStages < live ? st0 + kChunkStages : live;
          for (int sidx = st0; sidx < st1; ++sidx) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_hi = ptx::smem_u32(smem + stage * L::STAGE_BYTES), a_mid = a_hi + L::A_BYTES;
            const uint32_t w_hi = a_hi + 2 * L::A_BYTES, w_mid = w_hi + L::B_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t dah = ptx::make_smem_desc<SW>(a_hi + k * 32), dam = ptx::make_smem_desc<SW>(a_mid + k * 32);
              const uint64_t dwh = ptx::make_smem_desc<SW>(w_hi + k * 32), dwm = ptx::make_smem_desc<SW>(w_mid + k * 32);
              // small terms first, the dominant product last
              ptx::mma_bf16_ss(d_tmem, dam, dwh, IDESC, (sidx > st0 || k != 0) ? 1u : 0u);
              ptx::mma_bf16_ss(d_tmem, dah, dwm, IDESC, 1u);
              ptx::mma_bf16_ss(d_tmem, dah, dwh, IDESC, 1u);
            }
            if constexpr (CL == 2) ptx::mma_commit_multicast(&empty[stage], (uint16_t)3);
            else ptx::mma_commit(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(&pfull[pb]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ chunk adders + epilogue (8 warps)
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFF) + (warp - 2) * (32 * CW);
    const int q = warp & 3;                   // TMEM lane quarter of this hardware warp
    const int half = (warp - 2) >> 2;         // which half of the tile's columns (as epilogue_tile splits them)
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int chunk = 0;
    for (int tile = worker; tile < total_tiles; tile += n_workers) {
      const int m_blk = item_m_blk(tile), n_blk = tile % n_tiles;
      const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
      const int live = (s.J - __popc(f32x_zero_taps(s, n0, BN))) * kchunks;
      const bool ghost = m_blk >= m_tiles;   // CL = 2, odd number of row tiles: computed on zero rows, never stored
      if (!ghost) epilogue_prefetch(ep, clip, s.T, t0 + q * 32, n0 + half * (BN / 2), BN / 2, lane);
      bool first = true;
      for (int st0 = 0; st0 < live; st0 += kChunkStages, ++chunk) {
        const int pb = chunk % PBUFS;
        ptx::mbar_wait_sleepy(&pfull[pb], (uint32_t)(chunk / PBUFS) & 1u);
        ptx::tc_fence_after();
        f32x_add_partial<BN>(tm_T + lane_off, tm_P + pb * BN + lane_off, half, first);
        tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&pempty[pb]);
        first = false;
      }
      // the tile's total is complete in T (this warp's region was written by this warp only): the usual epilogue
      ptx::tc_fence_after();
      if (!ghost) epilogue_tile<BN, CW>(ep, variant, stg, tm_T, clip, t0, n0, s.T, 2 + ((warp - 2) & 7), lane);
      ptx::tc_fence_before();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();   // no CTA exits while its peer may still multicast into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------- launcher
template <int BN, int BK, int CL = 1>
static int launch_f32x_cfg(const __nv_bfloat16* A2, const __nv_bfloat16* W2, const ConvGemmShape& s, const Epilogue& e,
                           cudaStream_t st, int sm_count) {
  using L = f32x::Smem<BN, BK>;
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(gemm_f32x_kernel<BN, BK, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s.T + 127) / 128;
  const long long m_tiles = (long long)s.B * tiles_per_clip;
  const int n_tiles = s.N / BN;
  DC_CHECK(m_tiles * n_tiles > 0 && m_tiles * n_tiles < (1ll << 31), DC_ERR_SHAPE, "gemm_f32x: bad tile count");
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)2 * s.C, (uint64_t)s.T, (uint64_t)s.B};
    const uint64_t strides[2] = {(uint64_t)2 * s.C * 2, (uint64_t)s.T * 2 * s.C * 2};
    const uint32_t box[3] = {(uint32_t)BK, 128, 1};
    DC_TRY(make_tmap_bf16(&tmA, A2, 3, dims, strides, box, BK * 2));
  }
  {
    const uint64_t K = (uint64_t)s.J * 2 * s.C;
    const uint64_t dims[2] = {K, (uint64_t)s.N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN / CL)};   // CL = 2: each CTA of a pair loads half the rows
    DC_TRY(make_tmap_bf16(&tmW, W2, 2, dims, strides, box, BK * 2));
  }
  const long long total = ((m_tiles + CL - 1) / CL) * n_tiles;
  const long long slots = CL == 2 ? sm_count / 2 : (long long)sm_count * L::CTAS_PER_SM;
  const int grid = (int)(total < slots ? total : slots) * CL;
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * s.N * s.J * s.C * s.alg_scale;
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) |
                     (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) | (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? 4.0 : 0.0) + (e.out1 ? 4.0 : 0.0) + (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_GEMM_F32, 2.0 * macs, rows * s.C * 4.0 + (double)s.N * s.J * s.C * 4.0 + rows * s.N * out_bytes, st,
                 CL == 2 ? "x<%d,%d>x2|C%d N%d J%d d%d e%d" : "x<%d,%d>|C%d N%d J%d d%d e%d", BN, BK, s.C, s.N, s.J, s.dil,
                 esig);
    Epilogue eg = e;
    eg.prefetch = DC_F32X_PREFETCH ? e.prefetch : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    DC_CUDA(cudaLaunchKernelEx(&cfg, gemm_f32x_kernel<BN, BK, CL>, tmA, tmW, s, eg, epilogue_variant(e), tiles_per_clip,
                               (int)m_tiles, n_tiles));
  }
  ++g_launches_f32x;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ---------------------------------------------------------------- weights-stationary form for the narrow layers
// C = BK and N = BN in {32, 64} (decoder stages 3 and 4, plain Conv1d): the generic kernel above re-loads the
// activation box once per tap, 54 B of L2 -> SM traffic per output element at C = 32, k = 11 against 16 B of HBM traffic,
// and runs at the L2 bandwidth (10 TB/s) instead of the HBM roof.  Here all J taps of [Whi | Wmid] stay in shared memory
// for the whole launch, the activation tile [hi, mid] is loaded ONCE per row tile with its (J - 1) dil halo, and the taps
// are row-shifted UMMA descriptors into it (as in conv_ws.cu).  Chunking, partial accumulators and the epilogue are the
// generic kernel's (a chunk = the taps that make K = 256 per term), except that the eight consumer warps form two groups
// with one total accumulator each and take alternate tiles.
struct F32xWsLayout {
  int w_bytes, a_half /* one of the two activation boxes, 1024-aligned */, stages, a_off, stg_off, bar_off, total, ra;
};
template <int BN, int BK>
static F32xWsLayout f32x_ws_layout(int J, int dil) {
  F32xWsLayout l;
  l.ra = (128 + (J - 1) * dil + 7) / 8 * 8;
  l.w_bytes = J * 2 * BN * BK * 2;
  l.a_half = (l.ra * BK * 2 + 1023) / 1024 * 1024;
  const int stg_bytes = 8 * 32 * f32x::Cfg<BN>::CW * 4;
  const int budget = (BK == 32 ? 113 * 1024 : 232448) - 1024 - 256 - stg_bytes - l.w_bytes;  // BK = 32: two CTAs per SM
  l.stages = budget / (2 * l.a_half);
  if (l.stages > 3) l.stages = 3;
  l.a_off = l.w_bytes;
  l.stg_off = l.a_off + (l.stages > 0 ? l.stages : 0) * 2 * l.a_half;
  l.bar_off = l.stg_off + stg_bytes;
  l.total = l.bar_off + 256 + 1024;
  return l;
}

template <int BN, int BK>
__global__ void __launch_bounds__(320, BK == 32 ? 2 : 1)
gemm_f32x_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvGemmShape s,
                    Epilogue ep, int variant, int tiles_per_clip, int total_tiles, F32xWsLayout lay) {
  using namespace f32x;
  // two consumer groups of 4 warps on alternate tiles, each with its own total accumulator T and its own ring of two
  // partial accumulators (a group only ever waits for the NEXT phase of its own barriers): T[2] + P[2][2]
  constexpr int SW = BK * 2, PBUFS = 2, TMEM_COLS = 6 * BN <= 256 ? 256 : 512, CW = Cfg<BN>::CW;
  static_assert(6 * BN <= 512, "TMEM columns");
  constexpr int B_BYTES = BN * BK * 2, TAPS_PER_CHUNK = 256 / BK;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                                  // [J][hi, mid][BN][BK]
  uint8_t* sA = smem + lay.a_off;                                      // [stages][hi, mid][ra][BK]
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + lay.bar_off);   // [3]
  uint64_t* aempty = afull + 3;                                        // [3]
  uint64_t* pfull = aempty + 3;                                        // [2 groups][PBUFS]
  uint64_t* pempty = pfull + 2 * PBUFS;                                // [2 groups][PBUFS]
  uint64_t* wbar = pempty + 2 * PBUFS;                                 // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int stages = lay.stages, J = s.J;
  const int chunks_per_tile = (J + TAPS_PER_CHUNK - 1) / TAPS_PER_CHUNK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) {
        ptx::mbar_init(&afull[i], 1);
        ptx::mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < 2 * PBUFS; ++i) {
        ptx::mbar_init(&pfull[i], 1);
        ptx::mbar_init(&pempty[i], 4);
      }
      ptx::mbar_init(wbar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_T = tmem_base, tm_P = tmem_base + 2 * BN;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the weights once, then one tile per stage
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, (uint32_t)lay.w_bytes);
      for (int j = 0; j < J; ++j) {
        ptx::tma_load_2d(sW + (2 * j) * B_BYTES, &tmW, wbar, j * 2 * BK, 0);            // W hi of tap j
        ptx::tma_load_2d(sW + (2 * j + 1) * B_BYTES, &tmW, wbar, j * 2 * BK + BK, 0);   // W mid
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = (uint32_t)(2 * lay.ra * BK * 2);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int clip = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 128;
        ptx::mbar_wait(&aempty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&afull[stage], a_bytes);
        uint8_t* sp = sA + stage * 2 * lay.a_half;
        ptx::tma_load_3d(sp, &tmA, &afull[stage], 0, t0 + s.shift0, clip);                // A hi, rows t0 + shift0 ..
        ptx::tma_load_3d(sp + lay.a_half, &tmA, &afull[stage], BK, t0 + s.shift0, clip);  // A mid
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      ptx::mbar_wait(wbar, 0);
      const uint32_t w_base = ptx::smem_u32(sW);
      const uint32_t tap_step = (uint32_t)(s.dil * SW);   // bytes: tap j reads tile rows r + j * dil
      int stage = 0, it = 0;
      int gchunk[2] = {0, 0};   // chunks issued so far for each consumer group
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int g = it & 1;
        ptx::mbar_wait(&afull[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_hi0 = ptx::smem_u32(sA + stage * 2 * lay.a_half), a_mid0 = a_hi0 + lay.a_half;
        for (int j0 = 0; j0 < J; j0 += TAPS_PER_CHUNK) {
          const int chunk = g ? gchunk[1]++ : gchunk[0]++;
          const int pb = g * PBUFS + chunk % PBUFS;
          ptx::mbar_wait(&pempty[pb], ((uint32_t)(chunk / PBUFS) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tm_P + pb * BN;
          const int j1 = j0 + TAPS_PER_CHUNK < J ? j0 + TAPS_PER_CHUNK : J;
          for (int j = j0; j < j1; ++j) {
            const uint32_t a_hi = a_hi0 + j * tap_step, a_mid = a_mid0 + j * tap_step;
            const uint32_t w_hi = w_base + (2 * j) * B_BYTES, w_mid = w_hi + B_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t dah = ptx::make_smem_desc<SW>(a_hi + k * 32), dam = ptx::make_smem_desc<SW>(a_mid + k * 32);
              const uint64_t dwh = ptx::make_smem_desc<SW>(w_hi + k * 32), dwm = ptx::make_smem_desc<SW>(w_mid + k * 32);
              ptx::mma_bf16_ss(d_tmem, dam, dwh, IDESC, (j > j0 || k != 0) ? 1u : 0u);
              ptx::mma_bf16_ss(d_tmem, dah, dwm, IDESC, 1u);
              ptx::mma_bf16_ss(d_tmem, dah, dwh, IDESC, 1u);
            }
          }
          if (j1 == J) ptx::mma_commit(&aempty[stage]);
          ptx::mma_commit(&pfull[pb]);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ chunk adders + epilogue: 2 groups x 4 warps
    // A warp covers its TMEM lane quarter x all BN columns of the tiles of its group, so that two tiles are in their
    // (DRAM-latency-bound) epilogue at any time.
    float* stg = reinterpret_cast<float*>(smem + lay.stg_off) + (warp - 2) * (32 * CW);
    const int group = (warp - 2) >> 2, q = warp & 3;
    const int wid = 2 + ((warp - 2) & 3);      // epilogue_tile's warp id for column half 0: wid % 4 == warp % 4
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tm_Tg = tm_T + group * BN;
    int chunk = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != group) continue;
      const int clip = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 128;
      epilogue_prefetch(ep, clip, s.T, t0 + q * 32, 0, BN, lane);
      for (int c = 0; c < chunks_per_tile; ++c, ++chunk) {
        const int pb = group * PBUFS + chunk % PBUFS;   // `chunk` counts this group's chunks
        ptx::mbar_wait_sleepy(&pfull[pb], (uint32_t)(chunk / PBUFS) & 1u);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) f32x_add_partial<BN>(tm_Tg + lane_off, tm_P + pb * BN + lane_off, hh, c == 0);
        tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&pempty[pb]);
      }
      ptx::tc_fence_after();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) epilogue_tile<BN, CW>(ep, variant, stg, tm_Tg, clip, t0, 0, s.T, wid + 4 * hh, lane);
      ptx::tc_fence_before();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int BN, int BK>
static bool f32x_ws_fits(const ConvGemmShape& s) {
  if (!DC_F32X_WS || s.C != BK || s.N != BN || s.J < 2 || s.zero_taps != 0 || s.phase_cols != 0) return false;
  const F32xWsLayout l = f32x_ws_layout<BN, BK>(s.J, s.dil);
  return l.ra <= 256 && l.stages >= 2;
}

template <int BN, int BK>
static int launch_f32x_ws(const __nv_bfloat16* A2, const __nv_bfloat16* W2, const ConvGemmShape& s, const Epilogue& e,
                          cudaStream_t st, int sm_count) {
  const F32xWsLayout lay = f32x_ws_layout<BN, BK>(s.J, s.dil);
  static std::atomic<unsigned> attr_dev_mask{0u};
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(gemm_f32x_ws_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 BK == 32 ? 113 * 1024 : 232448));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s.T + 127) / 128;
  const long long total = (long long)s.B * tiles_per_clip;
  DC_CHECK(total > 0 && total < (1ll << 31), DC_ERR_SHAPE, "gemm_f32x_ws: bad tile count");
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)2 * s.C, (uint64_t)s.T, (uint64_t)s.B};
    const uint64_t strides[2] = {(uint64_t)2 * s.C * 2, (uint64_t)s.T * 2 * s.C * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)lay.ra, 1};
    DC_TRY(make_tmap_bf16(&tmA, A2, 3, dims, strides, box, BK * 2));
  }
  {
    const uint64_t K = (uint64_t)s.J * 2 * s.C;
    const uint64_t dims[2] = {K, (uint64_t)s.N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    DC_TRY(make_tmap_bf16(&tmW, W2, 2, dims, strides, box, BK * 2));
  }
  const long long slots = (long long)sm_count * (BK == 32 ? 2 : 1);
  const int grid = (int)(total < slots ? total : slots);
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * s.N * s.J * s.C * s.alg_scale;
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) |
                     (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) | (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? 4.0 : 0.0) + (e.out1 ? 4.0 : 0.0) + (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_GEMM_F32, 2.0 * macs, rows * s.C * 4.0 + (double)s.N * s.J * s.C * 4.0 + rows * s.N * out_bytes, st,
                 "xws<%d,%d>|C%d N%d J%d d%d e%d", BN, BK, s.C, s.N, s.J, s.dil, esig);
    gemm_f32x_ws_kernel<BN, BK><<<grid, 320, lay.total, st>>>(tmA, tmW, s, e, epilogue_variant(e), tiles_per_clip, (int)total,
                                                            lay);
  }
  ++g_launches_f32x;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

bool gemm_f32x_supported(const ConvGemmShape& s) { return s.C % 32 == 0 && s.N % 32 == 0 && s.J >= 1; }

// A2: (B, T, 2C) bf16 [hi | mid] of the fp32 activations (launch_split_f32); W2: [N][J][2C] bf16 [hi | mid] per tap.
int launch_gemm_f32x(const __nv_bfloat16* A2, const __nv_bfloat16* W2, const ConvGemmShape& s_in, const Epilogue& e,
                     cudaStream_t st, int sm_count) {
  ConvGemmShape s = s_in;
  DC_CHECK(gemm_f32x_supported(s), DC_ERR_SHAPE, "gemm_f32x: C=%d, N=%d must be multiples of 32", s.C, s.N);
  if (s.J == 1 && s.shift0 == 0) {  // no halo: flatten clips so every tile is full
    s.T = s.B * s.T;
    s.B = 1;
  }
  if (f32x_ws_fits<32, 32>(s)) return launch_f32x_ws<32, 32>(A2, W2, s, e, st, sm_count);
  if (f32x_ws_fits<64, 64>(s)) return launch_f32x_ws<64, 64>(A2, W2, s, e, st, sm_count);
  if (s.C % 64 == 0 && !(DC_F32X_NARROW_BK32 && s.N % 128 != 0)) {
    if (s.N % 128 == 0) {
      const bool pair = DC_F32X_PAIRS && s.cluster == 2 && (long long)s.B * ((s.T + 127) / 128) >= 2 && sm_count >= 2;
      return pair ? launch_f32x_cfg<128, 64, 2>(A2, W2, s, e, st, sm_count) : launch_f32x_cfg<128, 64>(A2, W2, s, e, st, sm_count);
    }
    if (s.N % 64 == 0) return launch_f32x_cfg<64, 64>(A2, W2, s, e, st, sm_count);
    return launch_f32x_cfg<32, 64>(A2, W2, s, e, st, sm_count);
  }
  if (s.N % 128 == 0) return launch_f32x_cfg<128, 32>(A2, W2, s, e, st, sm_count);
  if (s.N % 64 == 0) return launch_f32x_cfg<64, 32>(A2, W2, s, e, st, sm_count);
  return launch_f32x_cfg<32, 32>(A2, W2, s, e, st, sm_count);
}

}  // namespace dc
