// tcgen05 / TMEM / TMA shifted-row implicit GEMM (DC_MODE_BF16) — the generic kernel of the hot path's dense layers:
// pwconv1/2 and 1x1 convs (J = 1), the stem k7, conv_pre k13, the wide dilated ResBlock convs (J = k taps, tap j = the
// same TMA box shifted by shift0 + j*dil frames; TMA out-of-bounds zero fill IS the conv's zero padding) and
// ConvTranspose1d (stride phases stacked along N, union of input shifts along J, all-zero taps of a phase skipped).
// The narrow decoder stages are dispatched to conv_ts.cu (C = N = 128) and conv_ws.cu (C <= 64) from here.
//
//   out[b,t,n] = epi( sum_{j<J} sum_{c<C} A[b, t + shift0 + j*dil, c] * W[n, j*C + c] )
//
// Persistent, warp-specialised, one CTA per SM (64 + EG * 256 threads):
//   warp 0      TMA producer : 3-D box (BK channels x 128 frames x 1 clip) of A + 2-D box (BK x BN) of W per stage
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma (M=128, N=BN, K=16) BK/16 times per stage
//   warps 2..   epilogue     : EG groups of 8 warps (epilogue.cuh): tcgen05.ld -> smem transpose -> fused epilogue
// ACC TMEM accumulators (ACC x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tiles are ordered n-fastest so the CTAs running at any moment share a few A row-blocks (read from HBM once) and
// all of W (<= 27 MB, L2 resident).
#include "common.cuh"

#include <string.h>
#include "ptx.cuh"

#include <mutex>

namespace dc {

static thread_local uint64_t g_launches_tc = 0;
size_t gemm_tc_launch_count() { return g_launches_tc; }

// ---------------------------------------------------------------- tensor maps (driver entry point, no -lcuda)
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

// bf16 tensor of `rank` dims (innermost first), box in elements
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  DC_CHECK(enc != nullptr, DC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_CHECK(r == CUDA_SUCCESS, DC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
           (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return DC_OK;
}

}  // namespace dc

#include "epilogue.cuh"

namespace dc {

// taps (bit j) whose weights are all zero for the N tile [n0, n0 + bn): only when the tile lies inside one stride phase
__device__ __forceinline__ uint32_t tile_zero_taps(const ConvGemmShape& s, int n0, int bn) {
  if (s.zero_taps == 0 || s.phase_cols % bn != 0) return 0;
  return (s.zero_taps >> ((n0 / s.phase_cols) * s.J)) & ((1u << s.J) - 1u);
}

// ---------------------------------------------------------------- the kernel
// warp 0 TMA, warp 1 MMA, then EG groups of 8 epilogue warps; group g takes the tiles with (tile counter % EG) == g.
// EG = 2 (with ACC = 4 TMEM accumulators) is used where the epilogue is an HBM latency chain (C = 128 decoder stage).
template <int BN, int BK, int STAGES, int EG = 1>
struct GemmTcSmem {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  // columns per epilogue chunk (each warp owns BN/2); 16 where two epilogue groups must fit beside a 4-stage ring
  static constexpr int CW = (BN >= 64 && !(BN == 256 && EG == 2 && STAGES == 4)) ? 32 : 16;
  static constexpr int STG_OFF = STAGES * (A_BYTES + B_BYTES);  // 8 warps x (32 rows x CW fp32) transpose buffers
  static constexpr int STG_BYTES = EG * 8 * 32 * CW * 4;
  static constexpr int THREADS = 64 + EG * 256;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 8) * 8 + 16 + 1024 /*alignment slack*/;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// CL = 2 (BN = 256 only): thread-block clusters of two CTAs work on adjacent row tiles of the same N block and each
// TMA-multicasts one half (128 rows) of every weight tile into both CTAs' rings (see conv_tsw_kernel, conv_ts.cu).
template <int BN, int BK, int STAGES, int EG, int ACC, int CL>
__global__ void __launch_bounds__(64 + EG * 256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvGemmShape s,
               Epilogue ep, int variant, int tiles_per_clip, int m_tiles, int n_tiles) {
  using L = GemmTcSmem<BN, BK, STAGES, EG>;
  static_assert(CL == 1 || BN == 256, "weight multicast splits a 256-row tile into two 128-row boxes");
  const int rank = CL == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int worker = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_items = CL == 2 ? (m_tiles + 1) / 2 : m_tiles;
  auto item_m_blk = [&](int item) { return CL == 2 ? 2 * (item / n_tiles) + rank : item / n_tiles; };
  constexpr int SW = BK * 2;  // swizzle span = one K-row of the tile in bytes (128 or 64)
  constexpr int TMEM_COLS = ACC * BN;
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        ptx::mbar_init(&full[i], 1);
        ptx::mbar_init(&empty[i], CL);   // a stage is released by the MMAs of every CTA its weight tile is multicast to
      }
      for (int i = 0; i < ACC; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], 8);
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

  const int total_tiles = m_items * n_tiles;
  const int kchunks = s.C / BK;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {  // one lane, known to the compiler: operands go to uniform registers
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int m_blk = item_m_blk(tile), n_blk = tile % n_tiles;
        const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
        const uint32_t skip = tile_zero_taps(s, n0, BN);
        for (int j = 0; j < s.J; ++j) {
          if ((skip >> j) & 1u) continue;  // all-zero tap of this stride phase (ConvTranspose1d)
          const int trow = t0 + s.shift0 + j * s.dil;
          for (int kc = 0; kc < kchunks; ++kc) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], L::A_BYTES + L::B_BYTES);
            ptx::tma_load_3d(sA + stage * L::A_BYTES, &tmA, &full[stage], kc * BK, trow, clip);
            if constexpr (CL == 2)   // this CTA's half of the weight tile, into both CTAs
              ptx::tma_load_2d_multicast(sB + stage * L::B_BYTES + rank * (L::B_BYTES / 2), &tmB, &full[stage],
                                         j * s.C + kc * BK, n0 + rank * (BN / 2), (uint16_t)3);
            else
              ptx::tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full[stage], j * s.C + kc * BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (ptx::elect_one()) {  // one lane, known to the compiler: operands go to uniform registers
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers, ++it) {
        const int as = it % ACC;
        const uint32_t aphase = (it / ACC) & 1;
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const int n0m = (tile % n_tiles) * BN;
        const int live_kb = (s.J - __popc(tile_zero_taps(s, n0m, BN))) * kchunks;  // the producer skips zero taps
        for (int kb = 0; kb < live_kb; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + stage * L::A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + stage * L::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = ptx::make_smem_desc<SW>(a_addr + k * 32);
            const uint64_t db = ptx::make_smem_desc<SW>(b_addr + k * 32);
            ptx::mma_bf16_ss(d_tmem, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CL == 2) ptx::mma_commit_multicast(&empty[stage], (uint16_t)3);
          else ptx::mma_commit(&empty[stage]);  // frees the smem slot when these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull[as]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFF) + (warp - 2) * (32 * L::CW);
    const int group = (warp - 2) >> 3;
    const int wg = 2 + ((warp - 2) & 7);  // warp id within its group (2..9), keeps warp % 4 = TMEM lane quarter
    int it = 0;
    for (int tile = worker; tile < total_tiles; tile += n_workers, ++it) {
      if (EG > 1 && it % EG != group) continue;
      const int m_blk = item_m_blk(tile), n_blk = tile % n_tiles;
      const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
      const bool ghost = m_blk >= m_tiles;   // CL = 2, odd number of row tiles: computed on zero rows, never stored
      const int as = it % ACC;
      const uint32_t aphase = (it / ACC) & 1;
      if (!ghost) epilogue_prefetch(ep, clip, s.T, t0 + (wg & 3) * 32, n0 + ((wg - 2) >> 2) * (BN / 2), BN / 2, lane);
      ptx::mbar_wait_sleepy(&tfull[as], aphase);
      ptx::tc_fence_after();
      if (!ghost) epilogue_tile<BN, L::CW>(ep, variant, stg, tmem_base + as * BN, clip, t0, n0, s.T, wg, lane);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
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
template <int BN, int BK, int STAGES, int EG, int ACC, int CL>
static int launch_cfg_cl(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                         cudaStream_t st, int sm_count) {
  using L = GemmTcSmem<BN, BK, STAGES, EG>;
  static_assert(ACC * BN <= 512, "TMEM columns");

  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, BK, STAGES, EG, ACC, CL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int tiles_per_clip = (s.T + 127) / 128;
  const long long m_tiles = (long long)s.B * tiles_per_clip;
  const int n_tiles = s.N / BN;
  DC_CHECK(m_tiles * n_tiles < (1ll << 31), DC_ERR_SHAPE, "gemm_tc: too many tiles");

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)s.C, (uint64_t)s.T, (uint64_t)s.B};
    const uint64_t strides[2] = {(uint64_t)s.C * 2, (uint64_t)s.T * s.C * 2};
    const uint32_t box[3] = {(uint32_t)BK, 128, 1};
    DC_TRY(make_tmap_bf16(&tmA, A, 3, dims, strides, box, BK * 2));
  }
  {
    const uint64_t K = (uint64_t)s.J * s.C;
    const uint64_t dims[2] = {K, (uint64_t)s.N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN / CL)};   // CL = 2: each CTA of a pair loads half the rows
    DC_TRY(make_tmap_bf16(&tmB, W, 2, dims, strides, box, BK * 2));
  }
  const long long total = ((m_tiles + CL - 1) / CL) * n_tiles;
  const long long workers = total < sm_count / CL ? total : sm_count / CL;
  const int grid = (int)workers * CL;
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * s.N * s.J * s.C * s.alg_scale;
    // operands once: A rows x C, W, output rows x N (4 B/elt upper bound is not assumed: count bf16 A/W, 4 B out)
    // epilogue signature: 1 act | 2 gamma | 4 residual | 8 mean3 | 16 fp32 out | 32 bf16 out
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) | (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) |
                     (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? (e.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e.out1 ? 2.0 : 0.0) +
                             (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_GEMM_TC, 2.0 * macs, rows * s.C * 2.0 + (double)s.N * s.J * s.C * 2.0 + rows * s.N * out_bytes, st,
                 CL == 2 ? "<%d,%d,%d,%d,%d>x2|C%d N%d J%d d%d e%d" : "<%d,%d,%d,%d,%d>|C%d N%d J%d d%d e%d", BN, BK,
                 STAGES, EG, ACC, s.C, s.N, s.J, s.dil, esig);
    Epilogue eg = e;
    eg.prefetch = 0;  // A/B on one B200: the L2 prefetch helps the HBM-bound narrow kernels (conv_ws/_pair/_ts,
                      // -5 %) but costs the tensor-bound wide ones 3 % (extra L2 traffic next to 17 TB/s of operands)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(L::THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    DC_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, BK, STAGES, EG, ACC, CL>, tmA, tmB, s, eg, epilogue_variant(e),
                               tiles_per_clip, (int)m_tiles, n_tiles));
  }
  ++g_launches_tc;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

template <int BN, int BK, int STAGES, int EG = 1, int ACC = 2>
static int launch_cfg(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                      cudaStream_t st, int sm_count) {
  if constexpr (BN == 256) {   // CTA pairs need one full pair of row tiles
    if (s.cluster == 2 && (long long)s.B * ((s.T + 127) / 128) >= 2 && sm_count >= 2)
      return launch_cfg_cl<BN, BK, STAGES, EG, ACC, 2>(A, W, s, e, st, sm_count);
  }
  return launch_cfg_cl<BN, BK, STAGES, EG, ACC, 1>(A, W, s, e, st, sm_count);
}

int launch_gemm_tc(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s_in, const Epilogue& e,
                   cudaStream_t st, int sm_count) {
  ConvGemmShape s = s_in;
  DC_CHECK(s.N % 32 == 0, DC_ERR_SHAPE, "gemm_tc: N=%d must be a multiple of 32", s.N);
  DC_CHECK(s.C % 32 == 0, DC_ERR_SHAPE, "gemm_tc: C=%d must be a multiple of 32", s.C);
  DC_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, DC_ERR_ARG,
           "gemm_tc: operands must be 16-byte aligned");
  if (s.J == 1 && s.shift0 == 0) {  // no halo: flatten clips so every tile is full
    s.T = s.B * s.T;
    s.B = 1;
  }
  if (conv_ws_supported(s)) return launch_conv_ws(A, W, s, e, st, sm_count);  // narrow decoder stages
  if (conv_ts_supported(s)) return launch_conv_ts(A, W, s, e, st, sm_count);  // C = N = 128 decoder stage
  // wide convs (C >= 256, J > 1): tap-shared activations (conv_ts.cu)
  if (conv_tsw_supported(s)) return launch_conv_tsw(A, W, s, e, st, sm_count);
  if (s.C % 64 != 0) {
    DC_CHECK(s.N % 32 == 0, DC_ERR_SHAPE, "gemm_tc: unsupported N for C=32");
    if (s.N % 64 == 0) return launch_cfg<64, 32, 8>(A, W, s, e, st, sm_count);
    return launch_cfg<32, 32, 8>(A, W, s, e, st, sm_count);
  }
  if (s.N % 256 == 0) {
    // residual / dual-output epilogues (the decoder's conv2) are an HBM latency chain of ~20k cycles per 128 x 256
    // tile; when the tile's MMAs are shorter than that (K <= 1792: C = 256 with k = 3, 7) two epilogue groups, one
    // per TMEM accumulator, pay for the operand stage they cost (measured: 6.7 -> 5.7 ms, 8.3 -> 7.6 ms; with
    // longer K or two N tiles the shallower operand ring loses more than the epilogue gains)
    if ((e.res || (e.out0 && e.out1)) && s.J * s.C <= 1792 && s.N == 256)
      return launch_cfg<256, 64, 3, 2, 2>(A, W, s, e, st, sm_count);
    // longer K with a residual epilogue: keep the 4-stage operand ring and fit the second epilogue group by halving
    // the transpose chunks (16 columns, 2 KB per warp)
    if (e.res || (e.out0 && e.out1)) return launch_cfg<256, 64, 4, 2, 2>(A, W, s, e, st, sm_count);
    // the ConvNeXt MLP's first GEMM: K = C <= 1024 (6-8k MMA cycles per tile) against a 128 x 256 GELU + bf16
    // epilogue of similar length -> two epilogue groups as well
    if (e.act == ACT_GELU && s.J * s.C <= 1024) return launch_cfg<256, 64, 3, 2, 2>(A, W, s, e, st, sm_count);
    return launch_cfg<256, 64, 4>(A, W, s, e, st, sm_count);
  }
  if (s.N % 128 == 0) {
    // short-K layers (the C = 128 decoder stage, K <= 1408) are bound by the epilogue's HBM latency chain: two
    // epilogue groups on alternate tiles; long-K layers are tensor-bound and keep the deeper operand pipeline
    if (s.J * s.C <= 2048) return launch_cfg<128, 64, 5, 2, 4>(A, W, s, e, st, sm_count);
    return launch_cfg<128, 64, 6>(A, W, s, e, st, sm_count);
  }
  if (s.N % 64 == 0) return launch_cfg<64, 64, 8>(A, W, s, e, st, sm_count);
  return launch_cfg<32, 64, 8>(A, W, s, e, st, sm_count);
}

}  // namespace dc
