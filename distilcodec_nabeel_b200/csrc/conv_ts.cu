// Tap-shared tcgen05 implicit-GEMM convolution for the C = 128 decoder stage (ResBlock1 convs of HiFiGAN stage 2,
// models/convnext_utils.py:106-113, and the C = 128 ConvTranspose1d, models/generators.py:67-79).
//
//   out[b,t,n] = epi( sum_{j<J} sum_{c<128} A[b, t + shift0 + j*dil, c] * W[n, j*128 + c] )          N = 128
//
// Why a third kernel: with C = N = 128 the generic kernel (gemm_tc.cu) moves 32 KB from L2 per 256 tensor-pipe
// cycles and SM (the activation rows once per tap, the weights once per 128-row tile) = ~25 TB/s chip-wide, more
// than L2 delivers; it ran at 570 TFLOP/s.  Here
//   * a tile is 256 output rows (one weight tile feeds 256 rows), and
//   * per 64-channel chunk the 256 + (J-1)*dil activation rows are loaded ONCE (two TMA boxes of 160 rows) and all J
//     taps read them through row-shifted UMMA descriptors (legal for swizzled tiles: scripts/desc_probe.cu),
// which cuts the L2 traffic per MAC 3.3x (k = 11) and leaves the layer bound by the tensor pipe / HBM.
// Weights are too large to keep resident (J * 32 KB), they stream through a 5-stage ring of 16 KB (tap, chunk) tiles.
// The weight tile is the MMA's A operand (M = 128 channels) and the 256 rows its B operand (N = 256): see the kernel.
// TMEM: 2 tiles x 256 columns (time rows) x 128 lanes (channels) = all 512 columns; 16 epilogue warps.
#include "common.cuh"
#include "ptx.cuh"

#include <string.h>

#include "epilogue.cuh"

namespace dc {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

static thread_local uint64_t g_launches_ts = 0;
uint64_t conv_ts_launch_count() { return g_launches_ts; }

#ifdef DC_TSW_TRACE  // experiment builds: clock64 stamps of CTA 0, tiles 16..79 of its sequence (last launch wins)
__device__ long long g_tsw_trace[12][64];
#define TTRACE(ev, i) do { if (blockIdx.x == 0 && (i) >= 16 && (i) < 80) g_tsw_trace[ev][(i) - 16] = clock64(); } while (0)
extern "C" int dc_debug_tsw_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_tsw_trace, sizeof(g_tsw_trace)); }
#else
#define TTRACE(ev, i) do { } while (0)
#endif
namespace ts {
constexpr int C = 128, N = 128, BK = 64, KCH = C / BK;  // two 64-channel chunks
constexpr int B_BYTES = N * BK * 2;                     // 16 KB per (tap, chunk) weight tile
constexpr int STG_BYTES = 16 * 32 * 32 * 4;             // 16 epilogue warps x (32 rows x 32 fp32)
constexpr int THREADS = 64 + 16 * 32;
// <BOXR, AST, BST>: rows per activation TMA box (two boxes per 64-channel chunk buffer), chunk buffers, weight stages.
//   <160, 2, 5>: any halo up to 64 rows (k = 11, dilation 5).
//   <136, 3, 3>: k = 3 (halo <= 16 rows): a third, smaller activation buffer, loaded one chunk ahead of the weight
//   tiles.  A clock64 trace (scripts/tsw_trace.py) shows the k = 3 layers' MMA thread waiting ~3500 cycles per tile on
//   the activation barrier (7100 cycles per 256-row tile against 3072 of MMA time, at any clock); this variant is
//   worth 2-3 % only, and halving the loaded bytes (a timing experiment) 8 %: the wait is not the load itself.  Open.
template <int BOXR, int AST, int BST>
struct Cfg {
  static constexpr int A_ROWS = 2 * BOXR;
  static constexpr int A_BYTES = A_ROWS * BK * 2;
  static_assert(A_BYTES % 1024 == 0, "swizzled tiles need 1024-byte aligned bases");
  static constexpr int A_OFF = 0, B_OFF = AST * A_BYTES, STG_OFF = B_OFF + BST * B_BYTES, BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};
constexpr int MAX_A_ROWS = 320;
}  // namespace ts

// CL = 2: CTA pairs on adjacent 256-row tiles TMA-multicast the (tap, chunk) weight tiles to each other (each loads 64 of
// the 128 weight rows), see conv_tsw_kernel below; item i of a pair = tiles 2i and 2i + 1, a missing second tile is a ghost.
template <int BOXR, int AST, int BST, int CL>
__global__ void __launch_bounds__(ts::THREADS, 1)
conv_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvGemmShape s,
               Epilogue ep, int variant, int tiles_per_clip, int total_tiles) {
  using namespace ts;
  using L = Cfg<BOXR, AST, BST>;
  const int rank = CL == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int worker = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_items = CL == 2 ? (total_tiles + 1) / 2 : total_tiles;
  auto item_tile = [&](int item) { return CL == 2 ? 2 * item + rank : item; };
  constexpr int BOX_ROWS = BOXR, A_BYTES = L::A_BYTES, B_STAGES = BST, A_STAGES = AST;
  constexpr int A_OFF = L::A_OFF, B_OFF = L::B_OFF, STG_OFF = L::STG_OFF, BAR_OFF = L::BAR_OFF;
  // Operand roles are swapped: the WEIGHT tile (128 output channels x 64) is the MMA's A operand (M = 128) and the
  // 256 activation rows are its B operand (N = 256), so the accumulator is transposed (TMEM lane = channel, column
  // = time row).  An M = 128, N = 128 MMA reads 8 KB of shared memory per 64 tensor cycles = the full 128 B/clk of
  // the shared-memory port; M = 128, N = 256 reads 12 KB per 128 cycles = 96 B/clk and is no longer operand-bound.
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem + A_OFF;
  uint8_t* sB = smem + B_OFF;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [A_STAGES]
  uint64_t* aempty = afull + A_STAGES;                             // [A_STAGES]
  uint64_t* bfull = aempty + A_STAGES;                             // [B_STAGES]
  uint64_t* bempty = bfull + B_STAGES;                             // [B_STAGES]
  uint64_t* tfull = bempty + B_STAGES;                             // [2]
  uint64_t* tempty = tfull + 2;                                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < A_STAGES; ++i) {
        ptx::mbar_init(&afull[i], 1);
        ptx::mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < B_STAGES; ++i) {
        ptx::mbar_init(&bfull[i], 1);
        ptx::mbar_init(&bempty[i], CL);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], 16);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (same order as the MMA issuer consumes)
    if (ptx::elect_one()) {  // one lane, known to the compiler: operands go to uniform registers
      // One thread issues both operand streams.  The weight ring is the back-pressure path (its stages free only as
      // MMAs complete), so an activation load issued AFTER a chunk's weight loads starts a whole chunk late; with a
      // third activation buffer the load of chunk c + 1 is issued BEFORE the weight tiles of chunk c.
      constexpr int AHEAD = A_STAGES >= 3 ? 1 : 0;
      int bs = 0, as = 0;
      uint32_t bphase = 0, aphase = 0;
      const int my_tiles = worker < n_items ? (n_items - worker + n_workers - 1) / n_workers : 0;
      const int n_chunks = my_tiles * KCH;
      auto load_a = [&](int c) {  // chunk c of this CTA's sequence = (tile c / KCH, channel chunk c % KCH)
        const int tile = item_tile(worker + (c / KCH) * n_workers), kc = c % KCH;
        const int clip = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 256;
        ptx::mbar_wait(&aempty[as], aphase ^ 1);
        ptx::mbar_expect_tx(&afull[as], A_BYTES);
        ptx::tma_load_3d(sA + as * A_BYTES, &tmA, &afull[as], kc * BK, t0 + s.shift0, clip);
        ptx::tma_load_3d(sA + as * A_BYTES + BOX_ROWS * BK * 2, &tmA, &afull[as], kc * BK, t0 + s.shift0 + BOX_ROWS, clip);
        if (++as == A_STAGES) { as = 0; aphase ^= 1; }
      };
      for (int c = 0; c < AHEAD && c < n_chunks; ++c) load_a(c);
      for (int c = 0; c < n_chunks; ++c) {
        if (c + AHEAD < n_chunks) load_a(c + AHEAD);
        const int kc = c % KCH;
        for (int j = 0; j < s.J; ++j) {
          ptx::mbar_wait(&bempty[bs], bphase ^ 1);
          ptx::mbar_expect_tx(&bfull[bs], B_BYTES);
          if constexpr (CL == 2)
            ptx::tma_load_2d_multicast(sB + bs * B_BYTES + rank * (B_BYTES / 2), &tmW, &bfull[bs], j * C + kc * BK,
                                       rank * (N / 2), (uint16_t)3);
          else
            ptx::tma_load_2d(sB + bs * B_BYTES, &tmW, &bfull[bs], j * C + kc * BK, 0);
          if (++bs == B_STAGES) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (ptx::elect_one()) {  // one lane, known to the compiler: operands go to uniform registers
      int bs = 0, as = 0, it = 0;
      uint32_t bphase = 0, aphase = 0;
      const uint32_t tap_step = (uint32_t)(s.dil * BK * 2) >> 4;   // descriptor start-address units (16 B) per tap
      for (int item = worker; item < n_items; item += n_workers, ++it) {
        const int p = it & 1;
        TTRACE(0, it);
        ptx::mbar_wait(&tempty[p], ((it >> 1) & 1) ^ 1);
        TTRACE(1, it);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + p * 256;
        long long wa = 0, wb = 0;
        for (int kc = 0; kc < KCH; ++kc) {
#ifdef DC_TSW_TRACE
          long long c0 = clock64();
#endif
          ptx::mbar_wait(&afull[as], aphase);
#ifdef DC_TSW_TRACE
          wa += clock64() - c0;
#endif
          ptx::tc_fence_after();
          uint64_t da = ptx::make_smem_desc<128>(ptx::smem_u32(sA + as * A_BYTES));
          for (int j = 0; j < s.J; ++j) {
#ifdef DC_TSW_TRACE
            long long c1 = clock64();
#endif
            ptx::mbar_wait(&bfull[bs], bphase);
#ifdef DC_TSW_TRACE
            wb += clock64() - c1;
#endif
            ptx::tc_fence_after();
            const uint64_t db = ptx::make_smem_desc<128>(ptx::smem_u32(sB + bs * B_BYTES));
            const uint32_t acc = (kc | j) != 0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)  // D[channel, row] += W[channel, k] * A[row + j*dil, k]
              ptx::mma_bf16_ss(d0, db + 2 * k, da + 2 * k, IDESC, acc | (uint32_t)(k != 0));
            if constexpr (CL == 2) ptx::mma_commit_multicast(&bempty[bs], (uint16_t)3);
            else ptx::mma_commit(&bempty[bs]);
            if (++bs == B_STAGES) { bs = 0; bphase ^= 1; }
            da += tap_step;  // tap j + 1 = the same rows, dil rows further down
          }
          ptx::mma_commit(&aempty[as]);
          if (++as == A_STAGES) { as = 0; aphase ^= 1; }
        }
        ptx::mma_commit(&tfull[p]);
        TTRACE(2, it);
#ifdef DC_TSW_TRACE
        if (blockIdx.x == 0 && it >= 16 && it < 80) { g_tsw_trace[3][it - 16] = wa; g_tsw_trace[4][it - 16] = wb; }
#endif
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: 16 warps = 4 channel quarters x 4 row quarters
    float* stg = reinterpret_cast<float*>(smem + STG_OFF) + (warp - 2) * (32 * 32);
    // epilogue warp index whose low 2 bits equal the hardware warp's TMEM lane quarter (warp % 4)
    const int warp16 = (((warp - 2) >> 2) << 2) | (warp & 3);
    int it = 0;
    for (int item = worker; item < n_items; item += n_workers, ++it) {
      const int tile = item_tile(item);
      const bool ghost = tile >= total_tiles;
      const int clip = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 256;
      const int p = it & 1;
      // this warp's region: 32 channels (one 128-byte line per row) x 64 rows
      if (!ghost) {
        epilogue_prefetch(ep, clip, s.T, t0 + (warp16 >> 2) * 64, (warp16 & 3) * 32, 32, lane);
        epilogue_prefetch(ep, clip, s.T, t0 + (warp16 >> 2) * 64 + 32, (warp16 & 3) * 32, 32, lane);
      }
      if (warp16 == 0 && lane == 0) TTRACE(5, it);
      ptx::mbar_wait_sleepy(&tfull[p], (it >> 1) & 1);
      if (warp16 == 0 && lane == 0) TTRACE(6, it);
      ptx::tc_fence_after();
      if (!ghost) epilogue_tile_transposed(ep, variant, stg, tmem_base + p * 256, clip, t0, s.T, warp16, lane);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[p]);
      if (warp16 == 0 && lane == 0) TTRACE(7, it);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================ wide variant
// Tap-shared activations for the WIDE decoder convs (C = 256 / 512, N a multiple of 256, J > 1): the generic kernel
// (gemm_tc.cu) re-reads the same 128 activation rows from L2 once per tap; here one TMA box per 64-channel chunk
// brings 128 + (J-1)*dil rows and all J taps read it through row-shifted descriptors, so the L2 -> SM operand
// traffic of a k = 11 conv drops 29 % (528 -> 375 KB per tile and chunk).  128 x 256 tiles, two TMEM accumulators,
// weights stream through a 4-stage ring of 32 KB (tap, chunk) tiles, activations through 2 chunk buffers.
namespace tsw {
constexpr int BN = 256, BK = 64;
constexpr int A_ROWS = 184;                       // 128 + max halo (50), multiple of 8
constexpr int A_BYTES = A_ROWS * BK * 2;          // 23,552 B (23 x 1024) per chunk buffer
constexpr int A_STAGES = 2;
constexpr int B_BYTES = BN * BK * 2;              // 32 KB per (tap, chunk) weight tile
constexpr int A_OFF = 0, B_OFF = A_STAGES * A_BYTES;
// <EG, CW, BST>: epilogue groups, transpose chunk width, weight ring depth.  <1,32,4> conv1-type epilogues;
// <2,16,4> residual epilogues with long K; <2,32,3> residual epilogues with short K (epilogue-bound: wider chunks
// matter more than the 4th weight stage)
template <int EG, int CW, int BST>
struct Cfg {
  static constexpr int STG_OFF = B_OFF + BST * B_BYTES;
  static constexpr int STG_BYTES = EG * 8 * 32 * CW * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};
}  // namespace tsw

// CL = 2: the kernel runs as thread-block clusters of two CTAs that work on ADJACENT row tiles of the same N block, i.e. on
// the same weight tiles: each CTA loads one half (128 of the 256 weight rows) of every (tap, chunk) tile and TMA-multicasts
// it into both CTAs' rings, so the L2 -> SM weight traffic per CTA halves (weights are ~94 % of this kernel's operand
// traffic: 44 x 32 KB per tile against 4 x 23 KB of activations at k = 11, C = 256) and a ring of the same depth covers
// twice the L2 round trip.  A weight slot is released to BOTH producers (multicast commit, empty barrier count 2).
// A pair whose second row tile does not exist runs it as a ghost (TMA zero-fills, nothing is stored).
template <int EG, int CW, int BST, int CL>
__global__ void __launch_bounds__(64 + EG * 256, 1)
conv_tsw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvGemmShape s,
                Epilogue ep, int variant, int tiles_per_clip, int m_tiles, int n_tiles) {
  using namespace tsw;
  using L = Cfg<EG, CW, BST>;
  // work items: CL = 1: item = (m_blk, n_blk), N fastest; CL = 2: item = (pair of row tiles, n_blk), this CTA takes
  // row tile 2 * pair + rank
  const int rank = CL == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int worker = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_items = CL == 2 ? (m_tiles + 1) / 2 : m_tiles;
  auto item_m_blk = [&](int item) { return CL == 2 ? 2 * (item / n_tiles) + rank : item / n_tiles; };
  constexpr int B_STAGES = BST, STG_OFF = L::STG_OFF, BAR_OFF = L::BAR_OFF;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem + A_OFF;
  uint8_t* sB = smem + B_OFF;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [A_STAGES]
  uint64_t* aempty = afull + A_STAGES;
  uint64_t* bfull = aempty + A_STAGES;                             // [B_STAGES]
  uint64_t* bempty = bfull + B_STAGES;
  uint64_t* tfull = bempty + B_STAGES;                             // [2]
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int kchunks = s.C / BK;
  const int total_tiles = m_items * n_tiles;
  const uint32_t a_bytes = (uint32_t)(((128 + (s.J - 1) * s.dil + 7) / 8 * 8) * BK * 2);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < A_STAGES; ++i) {
        ptx::mbar_init(&afull[i], 1);
        ptx::mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < B_STAGES; ++i) {
        ptx::mbar_init(&bfull[i], 1);
        ptx::mbar_init(&bempty[i], CL);   // released by the MMAs of every CTA the slot is multicast to
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], 8);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();   // the peer's barriers are initialised before anything is multicast to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int m_blk = item_m_blk(tile), n0 = (tile % n_tiles) * BN;
        const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128;
        for (int kc = 0; kc < kchunks; ++kc) {
          ptx::mbar_wait(&aempty[as], aphase ^ 1);
          ptx::mbar_expect_tx(&afull[as], a_bytes);
          ptx::tma_load_3d(sA + as * A_BYTES, &tmA, &afull[as], kc * BK, t0 + s.shift0, clip);
          if (++as == A_STAGES) { as = 0; aphase ^= 1; }
          for (int j = 0; j < s.J; ++j) {
            ptx::mbar_wait(&bempty[bs], bphase ^ 1);
            ptx::mbar_expect_tx(&bfull[bs], B_BYTES);
            // the weight tile comes as two boxes of 128 rows: both from this CTA, or one from each CTA of the pair
            if constexpr (CL == 2) {
              ptx::tma_load_2d_multicast(sB + bs * B_BYTES + rank * (B_BYTES / 2), &tmW, &bfull[bs], j * s.C + kc * BK,
                                         n0 + rank * (BN / 2), (uint16_t)3);
            } else {
              ptx::tma_load_2d(sB + bs * B_BYTES, &tmW, &bfull[bs], j * s.C + kc * BK, n0);
              ptx::tma_load_2d(sB + bs * B_BYTES + B_BYTES / 2, &tmW, &bfull[bs], j * s.C + kc * BK, n0 + BN / 2);
            }
            if (++bs == B_STAGES) { bs = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      int as = 0, bs = 0, it = 0;
      uint32_t aphase = 0, bphase = 0;
      const uint32_t tap_step = (uint32_t)(s.dil * BK * 2) >> 4;
      for (int tile = worker; tile < total_tiles; tile += n_workers, ++it) {
        const int p = it & 1;
        TTRACE(0, it);
        ptx::mbar_wait(&tempty[p], ((it >> 1) & 1) ^ 1);
        TTRACE(1, it);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + p * BN;
        long long wa = 0, wb = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
#ifdef DC_TSW_TRACE
          long long c0 = clock64();
#endif
          ptx::mbar_wait(&afull[as], aphase);
#ifdef DC_TSW_TRACE
          wa += clock64() - c0;
#endif
          ptx::tc_fence_after();
          uint64_t da = ptx::make_smem_desc<128>(ptx::smem_u32(sA + as * A_BYTES));
          for (int j = 0; j < s.J; ++j) {
#ifdef DC_TSW_TRACE
            long long c1 = clock64();
#endif
            ptx::mbar_wait(&bfull[bs], bphase);
#ifdef DC_TSW_TRACE
            wb += clock64() - c1;
#endif
            ptx::tc_fence_after();
            const uint64_t db = ptx::make_smem_desc<128>(ptx::smem_u32(sB + bs * B_BYTES));
            const uint32_t acc = (kc | j) != 0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) ptx::mma_bf16_ss(d0, da + 2 * k, db + 2 * k, IDESC, acc | (uint32_t)(k != 0));
            if constexpr (CL == 2) ptx::mma_commit_multicast(&bempty[bs], (uint16_t)3);
            else ptx::mma_commit(&bempty[bs]);
            if (++bs == B_STAGES) { bs = 0; bphase ^= 1; }
            da += tap_step;
          }
          ptx::mma_commit(&aempty[as]);
          if (++as == A_STAGES) { as = 0; aphase ^= 1; }
        }
        ptx::mma_commit(&tfull[p]);
        TTRACE(2, it);
#ifdef DC_TSW_TRACE
        if (blockIdx.x == 0 && it >= 16 && it < 80) { g_tsw_trace[3][it - 16] = wa; g_tsw_trace[4][it - 16] = wb; }
#endif
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: EG groups of 8 warps on alternate tiles
    float* stg = reinterpret_cast<float*>(smem + STG_OFF) + (warp - 2) * (32 * CW);
    const int group = (warp - 2) >> 3;
    const int wg = 2 + ((warp - 2) & 7);
    int it = 0;
    for (int tile = worker; tile < total_tiles; tile += n_workers, ++it) {
      if (EG > 1 && (it & 1) != group) continue;
      const int m_blk = item_m_blk(tile), n0 = (tile % n_tiles) * BN;
      const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128;
      const bool ghost = m_blk >= m_tiles;   // CL = 2, odd number of row tiles: the pair's second tile does not exist
      const int p = it & 1;
      if (!ghost) epilogue_prefetch(ep, clip, s.T, t0 + (wg & 3) * 32, n0 + ((wg - 2) >> 2) * (BN / 2), BN / 2, lane);
      if (wg == 2 && lane == 0) TTRACE(5, it);
      ptx::mbar_wait_sleepy(&tfull[p], (it >> 1) & 1);
      if (wg == 2 && lane == 0) TTRACE(6, it);
      ptx::tc_fence_after();
      if (!ghost) epilogue_tile<BN, CW>(ep, variant, stg, tmem_base + p * BN, clip, t0, n0, s.T, wg, lane);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[p]);
      if (wg == 2 && lane == 0) TTRACE(7, it);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();   // no CTA exits while its peer may still multicast into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

bool conv_tsw_supported(const ConvGemmShape& s) {
  return s.J > 1 && s.C % tsw::BK == 0 && s.C >= 256 && s.N % tsw::BN == 0 && s.zero_taps == 0 &&
         128 + (s.J - 1) * s.dil <= tsw::A_ROWS;
}

template <int EG, int CW, int BST, int CL>
static int launch_tsw(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                      cudaStream_t st, int sm_count) {
  using namespace tsw;
  constexpr int TOTAL = Cfg<EG, CW, BST>::TOTAL;
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_tsw_kernel<EG, CW, BST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s.T + 127) / 128;
  const long long m_tiles = (long long)s.B * tiles_per_clip;
  const int n_tiles = s.N / BN;
  DC_CHECK(m_tiles * n_tiles > 0 && m_tiles * n_tiles < (1ll << 31), DC_ERR_SHAPE, "conv_tsw: bad tile count");
  const int RA = (128 + (s.J - 1) * s.dil + 7) / 8 * 8;
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)s.C, (uint64_t)s.T, (uint64_t)s.B};
    const uint64_t strides[2] = {(uint64_t)s.C * 2, (uint64_t)s.T * s.C * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)RA, 1};
    DC_TRY(make_tmap_bf16(&tmA, A, 3, dims, strides, box, 128));
  }
  {
    const uint64_t K = (uint64_t)s.J * s.C;
    const uint64_t dims[2] = {K, (uint64_t)s.N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN / 2)};   // a weight tile = two boxes of 128 rows
    DC_TRY(make_tmap_bf16(&tmW, W, 2, dims, strides, box, 128));
  }
  // CL = 1: one CTA per (row tile, N block); CL = 2: one CTA PAIR per (two row tiles, N block)
  const long long total = ((m_tiles + CL - 1) / CL) * n_tiles;
  const long long workers = total < sm_count / CL ? total : sm_count / CL;
  const int grid = (int)workers * CL;
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * s.N * s.J * s.C * s.alg_scale;
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) |
                     (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) | (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? (e.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e.out1 ? 2.0 : 0.0) +
                             (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_CONV_TS, 2.0 * macs, rows * s.C * 2.0 + (double)s.N * s.J * s.C * 2.0 + rows * s.N * out_bytes, st,
                 CL == 2 ? "w<%d,%d,%d>x2|C%d N%d J%d d%d e%d" : "w<%d,%d,%d>|C%d N%d J%d d%d e%d", EG, CW, BST, s.C, s.N,
                 s.J, s.dil, esig);
    Epilogue eg = e;
    // L2-prefetching the residual tile while the MMAs run pays only where the layer is HBM-latency-bound (A/B on one
    // box: k = 3 at C = 256 -6 % / -11 %; k = 7 +9 %; the tensor-bound k = 11 layers lose ~3 %)
    // option epi_prefetch: 1 = that rule, 2 = also k = 7 at C = 256, 3 = every layer, 0 = never
    eg.prefetch = (e.prefetch >= 3 || (e.prefetch == 2 && s.J * s.C <= 1792) || (e.prefetch == 1 && s.J * s.C <= 768)) ? 1 : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(64 + EG * 256);
    cfg.dynamicSmemBytes = TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    DC_CUDA(cudaLaunchKernelEx(&cfg, conv_tsw_kernel<EG, CW, BST, CL>, tmA, tmW, s, eg, epilogue_variant(e), tiles_per_clip,
                               (int)m_tiles, n_tiles));
  }
  ++g_launches_ts;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int launch_conv_tsw(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                    cudaStream_t st, int sm_count) {
  DC_CHECK(conv_tsw_supported(s), DC_ERR_SHAPE, "conv_tsw: unsupported shape");
  // CTA pairs (weight multicast) need at least one full pair of row tiles and an even share of the SMs
  const bool pair = s.cluster == 2 && (long long)s.B * ((s.T + 127) / 128) >= 2 && sm_count >= 2;
  if (e.res || (e.out0 && e.out1)) {
#ifdef DC_TSW_RES_CW32   // A/B builds: whole-line epilogue chunks + 3-stage ring for every residual layer
    if (pair) return launch_tsw<2, 32, 3, 2>(A, W, s, e, st, sm_count);
#endif
    // short K (k = 3 at C = 512, k <= 7 at C = 256): epilogue-bound, whole-line chunks beat the 4th weight stage
    // (C = 512, k = 3: 3.59 -> 3.21 ms per launch)
    if (s.J * s.C <= 1792)
      return pair ? launch_tsw<2, 32, 3, 2>(A, W, s, e, st, sm_count) : launch_tsw<2, 32, 3, 1>(A, W, s, e, st, sm_count);
    return pair ? launch_tsw<2, 16, 4, 2>(A, W, s, e, st, sm_count) : launch_tsw<2, 16, 4, 1>(A, W, s, e, st, sm_count);
  }
  return pair ? launch_tsw<1, 32, 4, 2>(A, W, s, e, st, sm_count) : launch_tsw<1, 32, 4, 1>(A, W, s, e, st, sm_count);
}

bool conv_ts_supported(const ConvGemmShape& s) {
  return s.C == ts::C && s.N == ts::N && 256 + (s.J - 1) * s.dil <= ts::MAX_A_ROWS && s.J >= 1;
}

template <int BOXR, int AST, int BST, int CL>
static int launch_ts(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                     cudaStream_t st, int sm_count) {
  using namespace ts;
  constexpr int BOX_ROWS = BOXR, TOTAL = Cfg<BOXR, AST, BST>::TOTAL;
  DC_CHECK(256 + (s.J - 1) * s.dil <= 2 * BOXR, DC_ERR_SHAPE, "conv_ts: halo does not fit the activation buffer");
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_ts_kernel<BOXR, AST, BST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s.T + 255) / 256;
  const long long total = (long long)s.B * tiles_per_clip;
  DC_CHECK(total > 0 && total < (1ll << 31), DC_ERR_SHAPE, "conv_ts: bad tile count");
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)s.T, (uint64_t)s.B};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)s.T * C * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BOX_ROWS, 1};
    DC_TRY(make_tmap_bf16(&tmA, A, 3, dims, strides, box, 128));
  }
  {
    const uint64_t K = (uint64_t)s.J * C;
    const uint64_t dims[2] = {K, (uint64_t)N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(N / CL)};   // pairs: each CTA loads half the weight rows
    DC_TRY(make_tmap_bf16(&tmW, W, 2, dims, strides, box, 128));
  }
  const long long items = (total + CL - 1) / CL;
  const int grid = (int)(items < sm_count / CL ? items : sm_count / CL) * CL;
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * N * s.J * C * s.alg_scale;
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) |
                     (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) | (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? (e.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e.out1 ? 2.0 : 0.0) +
                             (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_CONV_TS, 2.0 * macs, rows * C * 2.0 + (double)N * s.J * C * 2.0 + rows * N * out_bytes, st,
                 AST == 2 ? "|C%d N%d J%d d%d e%d" : "<a3>|C%d N%d J%d d%d e%d", C, N, s.J, s.dil, esig);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    DC_CUDA(cudaLaunchKernelEx(&cfg, conv_ts_kernel<BOXR, AST, BST, CL>, tmA, tmW, s, e, epilogue_variant(e), tiles_per_clip,
                               (int)total));
  }
  ++g_launches_ts;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int launch_conv_ts(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                   cudaStream_t st, int sm_count) {
  DC_CHECK(conv_ts_supported(s), DC_ERR_SHAPE, "conv_ts: unsupported shape C=%d N=%d J=%d dil=%d", s.C, s.N, s.J, s.dil);
  const bool pair = s.cluster == 2 && (long long)s.B * ((s.T + 255) / 256) >= 2 && sm_count >= 2;
#ifndef DC_TS_NO_A3
  if (s.J <= 3 && (s.J - 1) * s.dil <= 16)
    return pair ? launch_ts<136, 3, 3, 2>(A, W, s, e, st, sm_count) : launch_ts<136, 3, 3, 1>(A, W, s, e, st, sm_count);
#endif
  return pair ? launch_ts<160, 2, 5, 2>(A, W, s, e, st, sm_count) : launch_ts<160, 2, 5, 1>(A, W, s, e, st, sm_count);
}

}  // namespace dc
