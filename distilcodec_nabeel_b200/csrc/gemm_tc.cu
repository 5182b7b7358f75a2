// tcgen05 / TMEM / TMA shifted-row implicit GEMM (DC_MODE_BF16) — the kernel every dense layer of the hot path
// runs on: pwconv1/2 and 1x1 convs (J = 1), the stem k7, conv_pre k13, the dilated ResBlock convs (J = k taps,
// tap j = the same TMA box shifted by shift0 + j*dil frames; TMA out-of-bounds zero fill IS the conv's zero
// padding) and ConvTranspose1d (stride phases stacked along N, union of input shifts along J).
//
//   out[b,t,n] = epi( sum_{j<J} sum_{c<C} A[b, t + shift0 + j*dil, c] * W[n, j*C + c] )
//
// Persistent, warp-specialised, one CTA per SM (320 threads):
//   warp 0      TMA producer   : 3-D box (BK channels x 128 frames x 1 clip) of A + 2-D box (BK x BN) of W per stage
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (M=128, N=BN, K=16) BK/16 times per stage; owns TMEM
//   warps 2..9  epilogue       : tcgen05.ld 32x32b -> registers -> bias/act/gamma/residual/mean3 -> global
// Two TMEM accumulators (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tiles are ordered n-fastest so the CTAs running at any moment share a few A row-blocks (read from HBM once) and
// all of W (<= 27 MB, L2 resident).
#include "common.cuh"
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

// ---------------------------------------------------------------- vectorised epilogue: 32 consecutive columns
__device__ __forceinline__ void load32(const void* base, size_t off, int dt, float (&v)[32]) {
  if (dt == DT_F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = p[i];
      v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
  } else {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 x = p[i];
      const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
        v[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
    }
  }
}
__device__ __forceinline__ void store32(void* base, size_t off, int dt, const float (&v)[32]) {
  if (dt == DT_F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
        w[k] = *reinterpret_cast<uint32_t*>(&h);
      }
      p[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__device__ __forceinline__ void epilogue32(const Epilogue& e, size_t row, int n, const uint32_t (&acc)[32]) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  if (e.bias) {
    const float4* b = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = __ldg(b + i);
      v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
    }
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
  } else if (e.act == ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
  }
  if (e.gamma) {
    const float4* g = reinterpret_cast<const float4*>(e.gamma + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = __ldg(g + i);
      v[4 * i] *= x.x; v[4 * i + 1] *= x.y; v[4 * i + 2] *= x.z; v[4 * i + 3] *= x.w;
    }
  }
  const size_t off = row * (size_t)e.ldo + n;
  if (e.res) {
    float r[32];
    load32(e.res, off, e.res_dt, r);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += r[i];
  }
  if (e.add1) {
    float a[32], b[32];
    load32(e.add1, off, e.add_dt, a);
    load32(e.add2, off, e.add_dt, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (v[i] + a[i] + b[i]) * e.scale;
  }
  if (e.out0) store32(e.out0, off, e.out0_dt, v);
  if (e.out1) {
    if (e.out1_silu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
    }
    store32(e.out1, off, e.out1_dt, v);
  }
}

// ---------------------------------------------------------------- the kernel
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

template <int BN, int BK, int STAGES>
struct GemmTcSmem {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int BAR_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024 /*alignment slack*/;
};

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvGemmShape s,
               Epilogue ep, int tiles_per_clip, int m_tiles, int n_tiles) {
  using L = GemmTcSmem<BN, BK, STAGES>;
  constexpr int SW = BK * 2;  // swizzle span = one K-row of the tile in bytes (128 or 64)
  constexpr int TMEM_COLS = 2 * BN;
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

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
        ptx::mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
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
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = m_tiles * n_tiles;
  const int kchunks = s.C / BK;
  const int num_kb = s.J * kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
        for (int j = 0; j < s.J; ++j) {
          const int trow = t0 + s.shift0 + j * s.dil;
          for (int kc = 0; kc < kchunks; ++kc) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], L::A_BYTES + L::B_BYTES);
            ptx::tma_load_3d(sA + stage * L::A_BYTES, &tmA, &full[stage], kc * BK, trow, clip);
            ptx::tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full[stage], j * s.C + kc * BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
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
          ptx::mma_commit(&empty[stage]);  // frees the smem slot when these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull[as]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which interleaved 32-column chunks it takes
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const int t = t0 + q * 32 + lane;
      const bool valid = t < s.T;
      const size_t row = (size_t)clip * s.T + t;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t acc[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c * 32, acc);
        ptx::tmem_ld_wait();
        if (valid) epilogue32(ep, row, n0 + c * 32, acc);
      }
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

// ---------------------------------------------------------------- launcher
template <int BN, int BK, int STAGES>
static int launch_cfg(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                      cudaStream_t st, int sm_count) {
  using L = GemmTcSmem<BN, BK, STAGES>;
  static bool attr_set = false;  // per-process; the attribute is per-function per-device, set it for every device once
  static int attr_dev_mask = 0;
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!attr_set || !(attr_dev_mask & (1 << dev))) {
    DC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 L::TOTAL));
    attr_set = true;
    attr_dev_mask |= 1 << dev;
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
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    DC_TRY(make_tmap_bf16(&tmB, W, 2, dims, strides, box, BK * 2));
  }
  const long long total = m_tiles * n_tiles;
  const int grid = (int)(total < sm_count ? total : sm_count);
  {
    const double rows = (double)s.B * s.T;
    const double macs = rows * s.N * s.J * s.C * s.alg_scale;
    // operands once: A rows x C, W, output rows x N (4 B/elt upper bound is not assumed: count bf16 A/W, 4 B out)
    ProfScope ps(PC_GEMM_TC, 2.0 * macs, rows * s.C * 2.0 + (double)s.N * s.J * s.C * 2.0 + rows * s.N * 4.0, st);
    gemm_tc_kernel<BN, BK, STAGES><<<grid, kGemmThreads, L::TOTAL, st>>>(tmA, tmB, s, e, tiles_per_clip, (int)m_tiles,
                                                                        n_tiles);
  }
  ++g_launches_tc;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
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
  if (s.C % 64 != 0) {
    DC_CHECK(s.N % 32 == 0, DC_ERR_SHAPE, "gemm_tc: unsupported N for C=32");
    if (s.N % 64 == 0) return launch_cfg<64, 32, 8>(A, W, s, e, st, sm_count);
    return launch_cfg<32, 32, 8>(A, W, s, e, st, sm_count);
  }
  if (s.N % 256 == 0) return launch_cfg<256, 64, 4>(A, W, s, e, st, sm_count);
  if (s.N % 128 == 0) return launch_cfg<128, 64, 6>(A, W, s, e, st, sm_count);
  if (s.N % 64 == 0) return launch_cfg<64, 64, 8>(A, W, s, e, st, sm_count);
  return launch_cfg<32, 64, 8>(A, W, s, e, st, sm_count);
}

}  // namespace dc
