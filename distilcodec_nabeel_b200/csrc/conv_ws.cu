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
    if (lane == 0) {
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
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
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
      ptx::mbar_wait(&tfull[as], aphase);
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
  static int attr_dev_mask = 0;
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask & (1 << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_ws_kernel<C, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_dev_mask |= 1 << dev;
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
