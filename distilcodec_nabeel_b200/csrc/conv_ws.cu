// Weights-stationary, tap-shared tcgen05 implicit-GEMM convolution for the narrow decoder stages (C = 32 / 64):
// the ResBlock1 convs of HiFiGAN stages 3-4 (models/convnext_utils.py:106-113) and the last ConvTranspose1d
// (models/generators.py:67-79), where the generic kernel (gemm_tc.cu) is bound by per-k-block pipeline overhead
// and by re-reading the same activation rows once per tap.
//
//   out[b,t,n] = epi( sum_{j<J} sum_{c<C} A[b, t + shift0 + j*dil, c] * W[n, j*C + c] )      N <= 64, C in {32, 64}
//
// * The whole weight matrix (J taps x N x C bf16, <= 90 KB) is loaded ONCE per CTA and stays in shared memory.
// * Per 128-row output tile ONE TMA box brings the 128 + (J-1)*dil activation rows the taps need (halo included;
//   rows outside the clip are zero-filled by TMA = the conv's zero padding).  Tap j is the same shared-memory tile
//   read j*dil rows further down: the UMMA descriptor's start address is simply advanced by j*dil rows (the
//   128B/64B swizzle is a function of the absolute shared-memory address, so any row offset is legal — verified by
//   scripts/desc_probe.cu on B200).  Activation traffic from L2 drops J-fold, barrier round trips J*C/BK-fold.
// * Four TMEM accumulators decouple the single-thread MMA issuer from the 8 epilogue warps (shared with gemm_tc.cu).
#include "common.cuh"
#include "ptx.cuh"

#include "epilogue.cuh"

namespace dc {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

static thread_local uint64_t g_launches_ws = 0;
uint64_t conv_ws_launch_count() { return g_launches_ws; }

// warp 0 TMA, warp 1 MMA, warps 2..17 epilogue in two groups of 8 that take alternate tiles: the epilogue of these
// HBM-bound layers is one long latency chain per tile (TMEM load -> transpose -> residual loads -> stores), so two
// tiles in flight double the bytes in flight per SM.
constexpr int kWsEpiGroups = 2;
constexpr int kWsThreads = 64 + kWsEpiGroups * 256;
constexpr int kWsAccStages = 4;

struct WsLayout {  // byte offsets inside the 1024-aligned dynamic shared memory
  int w_bytes, a_stage_bytes, stages, a_off, stg_off, bar_off, total;
};
static WsLayout ws_layout(int C, int N, int J, int RA) {
  WsLayout l;
  l.w_bytes = J * N * C * 2;                       // per-tap tiles are multiples of 2 KB -> 1024-aligned
  l.a_stage_bytes = (RA * C * 2 + 1023) / 1024 * 1024;
  const int stg_bytes = kWsEpiGroups * 8 * 32 * (N >= 64 ? 32 : 16) * 4;
  const int fixed = l.w_bytes + stg_bytes + 256 /*barriers*/ + 1024 /*alignment slack*/;
  l.stages = (232448 - fixed) / l.a_stage_bytes;
  if (l.stages > 8) l.stages = 8;
  l.a_off = l.w_bytes;
  l.stg_off = l.a_off + l.stages * l.a_stage_bytes;
  l.bar_off = l.stg_off + stg_bytes;
  l.total = l.bar_off + 256 + 1024;
  return l;
}

template <int C, int N>
__global__ void __launch_bounds__(kWsThreads, 1)
conv_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvGemmShape s,
               Epilogue ep, int variant, WsLayout lay, int tiles_per_clip, int total_tiles) {
  constexpr int SW = C * 2;                        // swizzle span = one activation row (64 or 128 bytes)
  constexpr int TMEM_COLS = kWsAccStages * N;      // 128 or 256
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, N);
  constexpr int W_TAP_BYTES = N * C * 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + lay.a_off;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + lay.bar_off);   // [8]
  uint64_t* empty = full + 8;                                          // [8]
  uint64_t* tfull = empty + 8;                                         // [4]
  uint64_t* tempty = tfull + kWsAccStages;                             // [4]
  uint64_t* wbar = tempty + kWsAccStages;                              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int stages = lay.stages;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 8; ++i) {
        ptx::mbar_init(&full[i], 1);
        ptx::mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < kWsAccStages; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], 8);
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, (uint32_t)lay.w_bytes);
      for (int j = 0; j < s.J; ++j) ptx::tma_load_2d(sW + j * W_TAP_BYTES, &tmW, wbar, j * C, 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = (uint32_t)(((128 + (s.J - 1) * s.dil + 7) / 8 * 8) * C * 2);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int clip = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 128;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full[stage], a_bytes);
        ptx::tma_load_3d(sA + stage * lay.a_stage_bytes, &tmA, &full[stage], 0, t0 + s.shift0, clip);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread; elect.sync tells the
    // compiler that exactly one lane is active, so operands move to uniform registers without a waterfall loop)
    if (ptx::elect_one()) {
      ptx::mbar_wait(wbar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint64_t w_desc0 = ptx::make_smem_desc<SW>(ptx::smem_u32(sW));
      const uint32_t a_step = (uint32_t)(s.dil * C * 2) >> 4;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it % kWsAccStages;
        const uint32_t aphase = (it / kWsAccStages) & 1;
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * N;
        // Descriptors advance by plain adds on the 16-byte-granular start-address field: tap j = the same tile
        // j*dil rows further down (+ j*dil*C*2 bytes), K step = +32 bytes.  The issuing thread is alone on its
        // scheduler, so every extra ALU instruction per MMA costs ~5 cycles of tensor-pipe idle time here.
        uint64_t da = ptx::make_smem_desc<SW>(ptx::smem_u32(sA + stage * lay.a_stage_bytes));
        uint64_t dw = w_desc0;
        uint32_t accum = 0;
        for (int j = 0; j < s.J; ++j) {
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            ptx::mma_bf16_ss(d_tmem, da + 2 * k, dw + 2 * k, IDESC, accum);
            accum = 1;
          }
          da += a_step;
          dw += W_TAP_BYTES >> 4;
        }
        ptx::mma_commit(&empty[stage]);  // the activation tile may be overwritten once these MMAs have read it
        ptx::mma_commit(&tfull[as]);     // accumulator complete
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (group g takes tiles it % groups == g)
    float* stg = reinterpret_cast<float*>(smem + lay.stg_off) + (warp - 2) * (32 * (N >= 64 ? 32 : 16));
    const int group = (warp - 2) >> 3;
    const int wg = 2 + ((warp - 2) & 7);  // warp id within its group, as epilogue_tile expects (2..9)
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if (it % kWsEpiGroups != group) continue;
      const int clip = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * 128;
      const int as = it % kWsAccStages;
      const uint32_t aphase = (it / kWsAccStages) & 1;
      epilogue_prefetch(ep, clip, s.T, t0 + (wg & 3) * 32, ((wg - 2) >> 2) * (N / 2), N / 2, lane);
      ptx::mbar_wait_sleepy(&tfull[as], aphase);
      ptx::tc_fence_after();
      epilogue_tile<N>(ep, variant, stg, tmem_base + as * N, clip, t0, 0, s.T, wg, lane);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ================================================================================================ fused conv pair
// One ResBlock1 step  x' = conv2(silu(conv1(s) + b1)) + b2 + x   (models/convnext_utils.py:109-112, s = silu(x) from
// the previous epilogue) in ONE kernel: conv1's output never goes to HBM.  Unfused, the pair moves 16 B per element
// (conv1: 2 in + 2 out, conv2: 2 + 4 in, 4 + 2 out) and conv1 alone costs as much time as conv2 because its narrow
// MMAs are bound by shared-memory operand reads; fused it moves 12 B and conv1's MMAs hide under conv2's HBM time.
//
// Tile = 128 rows of t = silu(conv1 + b1) -> 128 - (J-1) valid output rows of conv2 (its taps read t rows r .. r+J-1).
//   warp 0        TMA: one box of 128 + (J-1)*dil rows of s per tile (both weight sets are loaded once per CTA)
//   warp 1        MMA: conv1(i+1) is issued before conv2(i), so the tensor pipe works while epilogue 1 of tile i runs
//   warps 2..5    epilogue 1: D1 (TMEM) -> +b1 -> SiLU -> bf16 -> the t tile in shared memory, written directly in
//                 the UMMA K-major swizzled layout (128B/64B swizzle = XOR of the 16-byte chunk index with the row
//                 bits) and zeroed outside [0, T) (conv2's zero padding applies to t); fence.proxy.async; arrive
//   warps 6..21   epilogue 2, four groups of 4 warps, one tile in four each (group g owns accumulator D2[g]): the usual
//                 fused epilogue (residual, dual store / 3-branch mean) -> global; it is the HBM latency chain of
//                 the kernel, so it gets four tiles in flight, epilogue 1 (no global traffic) gets 4 warps
constexpr int kPairE1Warps = 4, kPairE2Warps = 16;
// Epilogue 2 is a memory-latency chain per tile (TMEM -> residual / mean operands from L2 or HBM -> stores; a
// clock64 trace of one CTA showed ~4800 cycles per tile and group with the MMA thread stalling on d2empty): NG2 groups
// of 16 / NG2 warps, one D2 accumulator each, keep NG2 tiles in flight.  A warp covers its TMEM lane quarter x all C
// columns.
#ifndef DC_PAIR_NG2
#define DC_PAIR_NG2 4
#endif
#ifndef DC_PAIR_CW32
#define DC_PAIR_CW32 1
#endif
constexpr int kPairNG2 = DC_PAIR_NG2, kPairG2Warps = kPairE2Warps / kPairNG2;
static_assert(kPairNG2 == 2 || kPairNG2 == 4, "2 groups of 8 warps or 4 groups of 4");
#ifdef DC_PAIR_TRACE  // experiment builds: per-role clock64 stamps of CTA 0, tiles 64..127 of its sequence
__device__ long long g_pair_trace[16][64];
#ifndef DC_PAIR_TRACE_C   // which launches record: -DDC_PAIR_TRACE_C=64 -DDC_PAIR_TRACE_J=7 (dilation 1); the last one wins
#define DC_PAIR_TRACE_C 32
#define DC_PAIR_TRACE_J 11
#endif
#define PTRACE(ev, i) do { if (blockIdx.x == 0 && C == DC_PAIR_TRACE_C && J == DC_PAIR_TRACE_J && (i) >= 64 && (i) < 128) g_pair_trace[ev][(i) - 64] = clock64(); } while (0)
extern "C" int dc_debug_pair_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_pair_trace, sizeof(g_pair_trace));
}
#else
#define PTRACE(ev, i) do { } while (0)
#endif
constexpr int kPairThreads = 64 + 32 * (kPairE1Warps + kPairE2Warps);

struct PairLayout {
  int w_bytes /*one weight set*/, a_stage_bytes, stages, t_bytes, a_off, t_off, stg_off, bar_off, total;
};
static PairLayout pair_layout(int C, int J, int RA) {
  PairLayout l;
  l.w_bytes = J * C * C * 2;
  l.a_stage_bytes = (RA * C * 2 + 1023) / 1024 * 1024;
  l.t_bytes = 144 * C * 2;  // 128 rows + the (J-1) <= 12 rows conv2's last taps touch (never valid outputs), 1024-aligned
  const int stg_bytes = kPairE2Warps * 32 * 16 * 4 + 256;  // 16-column transpose chunks + conv1's bias
  const int fixed = 2 * l.w_bytes + 2 * l.t_bytes + stg_bytes + 256 + 1024;
  l.stages = (232448 - fixed) / l.a_stage_bytes;
  if (l.stages > 6) l.stages = 6;
  l.a_off = 2 * l.w_bytes;
  l.t_off = l.a_off + (l.stages > 0 ? l.stages : 0) * l.a_stage_bytes;
  l.stg_off = l.t_off + 2 * l.t_bytes;
  l.bar_off = l.stg_off + stg_bytes;
  l.total = l.bar_off + 256 + 1024;
  return l;
}

template <int C>
__global__ void __launch_bounds__(kPairThreads, 1)
conv_ws_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, int T, int J, int dil, int shift_a /*tile row 0 of s
                    relative to the first output row*/, const float* __restrict__ bias1, Epilogue ep, int variant,
                    PairLayout lay, int tiles_per_clip, int total_tiles) {
  constexpr int SW = C * 2, PITCH = C * 2;
  constexpr int TMEM_COLS = (2 + kPairNG2) * C <= 128 ? 128 : ((2 + kPairNG2) * C <= 256 ? 256 : 512);  // D1[2], D2[NG2]
  static_assert((2 + kPairNG2) * C <= 512, "TMEM");
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, C);
  constexpr int W_TAP_BYTES = C * C * 2;
  const int MO = 128 - (J - 1), p2 = (J - 1) / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + lay.w_bytes;
  uint8_t* sA = smem + lay.a_off;
  uint8_t* sT = smem + lay.t_off;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + lay.bar_off);  // [6]
  uint64_t* aempty = afull + 6;                                        // [6]
  uint64_t* d1full = aempty + 6;                                       // [2]
  uint64_t* d1empty = d1full + 2;                                      // [2]
  uint64_t* tfull = d1empty + 2;                                       // [2]  t tile written by epilogue 1
  uint64_t* tempty = tfull + 2;                                        // [2]  t tile consumed by conv2's MMAs
  uint64_t* d2full = tempty + 2;                                       // [NG2]
  uint64_t* d2empty = d2full + kPairNG2;                               // [NG2]
  uint64_t* wbar = d2empty + kPairNG2;                                 // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int stages = lay.stages;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 6; ++i) {
        ptx::mbar_init(&afull[i], 1);
        ptx::mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&d1full[i], 1);
        ptx::mbar_init(&d1empty[i], kPairE1Warps);
        ptx::mbar_init(&tfull[i], kPairE1Warps);
        ptx::mbar_init(&tempty[i], 1);
      }
      for (int i = 0; i < kPairNG2; ++i) {
        ptx::mbar_init(&d2full[i], 1);
        ptx::mbar_init(&d2empty[i], kPairG2Warps);
      }
      ptx::mbar_init(wbar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  if (threadIdx.x < C)  // conv1's bias, read by epilogue 1 for every tile
    reinterpret_cast<float*>(smem + lay.stg_off + kPairE2Warps * 32 * 16 * 4)[threadIdx.x] = bias1[threadIdx.x];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_d1 = tmem_base, tm_d2 = tmem_base + 2 * C;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, (uint32_t)(2 * lay.w_bytes));
      for (int j = 0; j < J; ++j) {
        ptx::tma_load_2d(sW1 + j * W_TAP_BYTES, &tmW1, wbar, j * C, 0);
        ptx::tma_load_2d(sW2 + j * W_TAP_BYTES, &tmW2, wbar, j * C, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = (uint32_t)(((128 + (J - 1) * dil + 7) / 8 * 8) * PITCH);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int clip = tile / tiles_per_clip, o0 = (tile % tiles_per_clip) * MO;
        ptx::mbar_wait(&aempty[stage], phase ^ 1);
        PTRACE(14, (tile - (int)blockIdx.x) / (int)gridDim.x);
        ptx::mbar_expect_tx(&afull[stage], a_bytes);
        ptx::tma_load_3d(sA + stage * lay.a_stage_bytes, &tmA, &afull[stage], 0, o0 + shift_a, clip);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: conv1(0), then [conv1(i+1), conv2(i)]...
    if (ptx::elect_one()) {
      ptx::mbar_wait(wbar, 0);
      const uint64_t w1_desc = ptx::make_smem_desc<SW>(ptx::smem_u32(sW1));
      const uint64_t w2_desc = ptx::make_smem_desc<SW>(ptx::smem_u32(sW2));
      const uint32_t a_step = (uint32_t)(dil * PITCH) >> 4, t_step = (uint32_t)PITCH >> 4;
      int n_my = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++n_my;
      int stage = 0;
      uint32_t phase = 0;
      auto conv1 = [&](int i) {
        const int b = i & 1;
        PTRACE(0, i);
        ptx::mbar_wait(&d1empty[b], ((i >> 1) & 1) ^ 1);
        PTRACE(1, i);
        ptx::mbar_wait(&afull[stage], phase);
        PTRACE(2, i);
        ptx::tc_fence_after();
        uint64_t da = ptx::make_smem_desc<SW>(ptx::smem_u32(sA + stage * lay.a_stage_bytes));
        uint64_t dw = w1_desc;
        uint32_t accum = 0;
        for (int j = 0; j < J; ++j) {
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            ptx::mma_bf16_ss(tm_d1 + b * C, da + 2 * k, dw + 2 * k, IDESC, accum);
            accum = 1;
          }
          da += a_step;
          dw += W_TAP_BYTES >> 4;
        }
        ptx::mma_commit(&aempty[stage]);
        ptx::mma_commit(&d1full[b]);
        PTRACE(3, i);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      };
      auto conv2 = [&](int i) {
        const int b = i & 1;                         // t buffer
        const int b2 = i % kPairNG2;                 // D2 accumulator / epilogue-2 group
        const uint32_t ph2 = (uint32_t)(i / kPairNG2) & 1u;
        PTRACE(4, i);
        ptx::mbar_wait(&d2empty[b2], ph2 ^ 1);
        PTRACE(5, i);
        ptx::mbar_wait(&tfull[b], (i >> 1) & 1);
        PTRACE(6, i);
        ptx::tc_fence_after();
        uint64_t dt = ptx::make_smem_desc<SW>(ptx::smem_u32(sT + b * lay.t_bytes));
        uint64_t dw = w2_desc;
        uint32_t accum = 0;
        for (int j = 0; j < J; ++j) {  // conv2 tap j reads t rows r + j
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            ptx::mma_bf16_ss(tm_d2 + b2 * C, dt + 2 * k, dw + 2 * k, IDESC, accum);
            accum = 1;
          }
          dt += t_step;
          dw += W_TAP_BYTES >> 4;
        }
        ptx::mma_commit(&tempty[b]);
        ptx::mma_commit(&d2full[b2]);
        PTRACE(7, i);
      };
      if (n_my > 0) conv1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) conv1(i + 1);
        conv2(i);
      }
    }
  } else if (warp < 2 + kPairE1Warps) {
    // ------------------------------------------------------------ epilogue 1: D1 -> silu(. + b1) -> bf16 t tile in smem
    const int q = warp & 3;
    const int row = q * 32 + lane;                       // t-tile row of this thread (4 warps x 32 rows, all columns)
    const int swz = C >= 64 ? (row & 7) : ((row >> 1) & 3);
    const float* sb1 = reinterpret_cast<const float*>(smem + lay.stg_off + kPairE2Warps * 32 * 16 * 4);
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      const int o0 = (tile % tiles_per_clip) * MO;
      const int b = i & 1;
      const int gt = o0 - p2 + row;                      // sequence position of this t row
      const bool inside = gt >= 0 && gt < T;
      ptx::mbar_wait_sleepy(&d1full[b], (i >> 1) & 1);
      if (warp == 2 && lane == 0) PTRACE(8, i);
      ptx::tc_fence_after();
      ptx::mbar_wait(&tempty[b], ((i >> 1) & 1) ^ 1);    // conv2 of tile i-2 has finished reading this t buffer
      if (warp == 2 && lane == 0) PTRACE(9, i);
      uint8_t* trow = sT + b * lay.t_bytes + row * PITCH;
#pragma unroll
      for (int c = 0; c < C / 32; ++c) {
        uint32_t acc[32];
        ptx::tmem_ld_32x32(tm_d1 + b * C + ((uint32_t)(q * 32) << 16) + c * 32, acc);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {                    // 8 columns = one 16-byte chunk of the bf16 row
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 bb = *reinterpret_cast<const float2*>(sb1 + c * 32 + g * 8 + 2 * e);
            float2 v = silu_fast2(fadd2(make_float2(__uint_as_float(acc[g * 8 + 2 * e]), __uint_as_float(acc[g * 8 + 2 * e + 1])), bb));
            if (!inside) v = make_float2(0.f, 0.f);
            __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int chunk = c * 4 + g;
          *reinterpret_cast<uint4*>(trow + ((chunk ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();                          // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&tfull[b]);
        ptx::mbar_arrive(&d1empty[b]);
        if (warp == 2) PTRACE(10, i);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue 2: D2 -> fused epilogue -> global
    const int e2w = warp - (2 + kPairE1Warps);  // 0..15
    const int group = e2w / kPairG2Warps;
    const int wg = 2 + (e2w % kPairG2Warps);    // warp id as epilogue_tile expects: wg % 4 == warp % 4 (TMEM lane quarter)
    float* stg = reinterpret_cast<float*>(smem + lay.stg_off) + e2w * (32 * 16);
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      if (i % kPairNG2 != group) continue;      // group g owns accumulator D2[g]
      const int clip = tile / tiles_per_clip, o0 = (tile % tiles_per_clip) * MO;
      const uint32_t ph2 = (uint32_t)(i / kPairNG2) & 1u;
      if constexpr (kPairG2Warps == 8)
        epilogue_prefetch(ep, clip, T, o0 + (wg & 3) * 32, ((wg - 2) >> 2) * (C / 2), C / 2, lane, o0 + MO);
      else
        epilogue_prefetch(ep, clip, T, o0 + (wg & 3) * 32, 0, C, lane, o0 + MO);
      if (wg == 2 && lane == 0) PTRACE(11, i);
      ptx::mbar_wait_sleepy(&d2full[group], ph2);
      if (wg == 2 && lane == 0) PTRACE(12, i);
      ptx::tc_fence_after();
      // whole 128-byte lines per access through a 16-row staging tile (same 2 KB per warp as 32 rows x 16 columns)
      if constexpr (DC_PAIR_CW32 && C == 64) {
        epilogue_tile<C, 32, 16>(ep, variant, stg, tm_d2 + group * C, clip, o0, 0, T, wg, lane, o0 + MO);
        if constexpr (kPairG2Warps == 4)        // 4-warp groups: the same warp also takes the other column half
          epilogue_tile<C, 32, 16>(ep, variant, stg, tm_d2 + group * C, clip, o0, 0, T, wg + 4, lane, o0 + MO);
      } else if constexpr (DC_PAIR_CW32 && C == 32 && kPairG2Warps == 4) {
        // all 32 columns are one chunk: "column half 0" of a 64-wide tile
        epilogue_tile<2 * C, 32, 16>(ep, variant, stg, tm_d2 + group * C, clip, o0, 0, T, wg, lane, o0 + MO);
      } else {
        epilogue_tile<C, 16>(ep, variant, stg, tm_d2 + group * C, clip, o0, 0, T, wg, lane, o0 + MO);
        if constexpr (kPairG2Warps == 4)
          epilogue_tile<C, 16>(ep, variant, stg, tm_d2 + group * C, clip, o0, 0, T, wg + 4, lane, o0 + MO);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&d2empty[group]);
      if (wg == 2 && lane == 0) PTRACE(13, i);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// s1: conv1 (dilated), s2: conv2 (dil 1); both C -> C with the same kernel size
bool conv_ws_pair_supported(const ConvGemmShape& s1, const ConvGemmShape& s2) {
  if (!(s1.C == s2.C && s1.N == s1.C && s2.N == s2.C && (s1.C == 32 || s1.C == 64))) return false;
  if (s1.J != s2.J || s2.dil != 1 || s1.J < 2 || s1.J > 13 || (s1.J & 1) == 0) return false;
  if (s1.shift0 != -s1.dil * (s1.J - 1) / 2 || s2.shift0 != -(s2.J - 1) / 2) return false;
  if (s1.B != s2.B || s1.T != s2.T) return false;
  const int RA = (128 + (s1.J - 1) * s1.dil + 7) / 8 * 8;
  if (RA > 256) return false;
  return pair_layout(s1.C, s1.J, RA).stages >= 2;
}

template <int C>
static int launch_pair(const __nv_bfloat16* S, const __nv_bfloat16* W1, const __nv_bfloat16* W2, const float* bias1,
                       const ConvGemmShape& s1, const Epilogue& e2, cudaStream_t st, int sm_count) {
  const int J = s1.J, RA = (128 + (J - 1) * s1.dil + 7) / 8 * 8, MO = 128 - (J - 1);
  const PairLayout lay = pair_layout(C, J, RA);
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_ws_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s1.T + MO - 1) / MO;
  const long long total = (long long)s1.B * tiles_per_clip;
  DC_CHECK(total > 0 && total < (1ll << 31), DC_ERR_SHAPE, "conv_ws_pair: bad tile count");
  CUtensorMap tmA, tmW1, tmW2;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)s1.T, (uint64_t)s1.B};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)s1.T * C * 2};
    const uint32_t box[3] = {(uint32_t)C, (uint32_t)RA, 1};
    DC_TRY(make_tmap_bf16(&tmA, S, 3, dims, strides, box, C * 2));
  }
  for (int w = 0; w < 2; ++w) {
    const uint64_t K = (uint64_t)J * C;
    const uint64_t dims[2] = {K, (uint64_t)C};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)C, (uint32_t)C};
    DC_TRY(make_tmap_bf16(w ? &tmW2 : &tmW1, w ? W2 : W1, 2, dims, strides, box, C * 2));
  }
  const int grid = (int)(total < sm_count ? total : sm_count);
  const int p1 = s1.dil * (J - 1) / 2, p2 = (J - 1) / 2;
  {
    const double rows = (double)s1.B * s1.T;
    const double macs = 2.0 * rows * C * J * C;
    const int esig = (e2.res ? 4 : 0) | (e2.add1 ? 8 : 0) | (e2.out0 ? (e2.out0_dt == DT_F32 ? 16 : 32) : 0) | (e2.out1 ? 32 : 0);
    const double out_bytes = (e2.out0 ? (e2.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e2.out1 ? 2.0 : 0.0) +
                             (e2.res ? 4.0 : 0.0) + (e2.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_CONV_WS, 2.0 * macs, rows * C * 2.0 + 2.0 * J * C * C * 2.0 + rows * C * out_bytes, st,
                 "_pair<%d>|C%d N%d J%d d%d e%d", C, C, C, J, s1.dil, esig);
    conv_ws_pair_kernel<C><<<grid, kPairThreads, lay.total, st>>>(tmA, tmW1, tmW2, s1.T, J, s1.dil, -(p1 + p2), bias1, e2,
                                                                  epilogue_variant(e2), lay, tiles_per_clip, (int)total);
  }
  ++g_launches_ws;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int launch_conv_ws_pair(const __nv_bfloat16* S, const __nv_bfloat16* W1, const __nv_bfloat16* W2, const float* bias1,
                        const ConvGemmShape& s1, const ConvGemmShape& s2, const Epilogue& e2, cudaStream_t st,
                        int sm_count) {
  DC_CHECK(conv_ws_pair_supported(s1, s2), DC_ERR_SHAPE, "conv_ws_pair: unsupported shapes");
  DC_CHECK(bias1 != nullptr, DC_ERR_ARG, "conv_ws_pair: conv1 bias missing");
  if (s1.C == 64) return launch_pair<64>(S, W1, W2, bias1, s1, e2, st, sm_count);
  return launch_pair<32>(S, W1, W2, bias1, s1, e2, st, sm_count);
}

bool conv_ws_supported(const ConvGemmShape& s) {
  if (!((s.C == 32 || s.C == 64) && (s.N == 32 || s.N == 64))) return false;
  const int RA = (128 + (s.J - 1) * s.dil + 7) / 8 * 8;
  if (RA > 256) return false;
  const WsLayout l = ws_layout(s.C, s.N, s.J, RA);
  return l.stages >= 2;
}

template <int C, int N>
static int launch_ws(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                     cudaStream_t st, int sm_count) {
  const int RA = (128 + (s.J - 1) * s.dil + 7) / 8 * 8;
  const WsLayout lay = ws_layout(C, N, s.J, RA);
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_ws_kernel<C, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s.T + 127) / 128;
  const long long total = (long long)s.B * tiles_per_clip;
  DC_CHECK(total > 0 && total < (1ll << 31), DC_ERR_SHAPE, "conv_ws: bad tile count");
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)s.T, (uint64_t)s.B};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)s.T * C * 2};
    const uint32_t box[3] = {(uint32_t)C, (uint32_t)RA, 1};
    DC_TRY(make_tmap_bf16(&tmA, A, 3, dims, strides, box, C * 2));
  }
  {
    const uint64_t K = (uint64_t)s.J * C;
    const uint64_t dims[2] = {K, (uint64_t)N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)C, (uint32_t)N};
    DC_TRY(make_tmap_bf16(&tmW, W, 2, dims, strides, box, C * 2));
  }
  const int grid = (int)(total < sm_count ? total : sm_count);
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * N * s.J * C * s.alg_scale;
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) |
                     (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) | (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? (e.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e.out1 ? 2.0 : 0.0) +
                             (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_CONV_WS, 2.0 * macs, rows * C * 2.0 + (double)N * s.J * C * 2.0 + rows * N * out_bytes, st,
                 "<%d,%d>|C%d N%d J%d d%d e%d", C, N, C, N, s.J, s.dil, esig);
    conv_ws_kernel<C, N><<<grid, kWsThreads, lay.total, st>>>(tmA, tmW, s, e, epilogue_variant(e), lay, tiles_per_clip,
                                                              (int)total);
  }
  ++g_launches_ws;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int launch_conv_ws(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                   cudaStream_t st, int sm_count) {
  DC_CHECK(conv_ws_supported(s), DC_ERR_SHAPE, "conv_ws: unsupported shape C=%d N=%d J=%d dil=%d", s.C, s.N, s.J, s.dil);
  if (s.C == 64 && s.N == 64) return launch_ws<64, 64>(A, W, s, e, st, sm_count);
  if (s.C == 64 && s.N == 32) return launch_ws<64, 32>(A, W, s, e, st, sm_count);
  if (s.C == 32 && s.N == 64) return launch_ws<32, 64>(A, W, s, e, st, sm_count);
  return launch_ws<32, 32>(A, W, s, e, st, sm_count);
}

}  // namespace dc
