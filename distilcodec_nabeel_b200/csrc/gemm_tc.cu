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

// ---------------------------------------------------------------- epilogue helpers (4 consecutive columns per lane)
// After the per-warp shared-memory transpose a lane owns 4 consecutive output columns of one row, so that a
// quarter-warp covers 128 contiguous bytes of an fp32 row (64 of a bf16 row) and every global access is a full
// sector: the residual / mean operands are read and both outputs written fully coalesced.
__device__ __forceinline__ float silu_fast(float x) {  // x * sigmoid(x): 2 MUFU + 3 FP32 ops, ~1e-6 relative
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}
// exact-erf GELU (nn.GELU(), convnext_utils.py:254) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far
// below the bf16 rounding of the stored result): 2 MUFU + 10 FP32 ops instead of erff's ~25 with branches
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.f);          // erf(|x| / sqrt 2)
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float4 ld4(const void* base, size_t off, int dt) {
  if (dt == DT_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
  const uint2 x = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
  return make_float4(__uint_as_float(x.x << 16), __uint_as_float(x.x & 0xffff0000u), __uint_as_float(x.y << 16),
                     __uint_as_float(x.y & 0xffff0000u));
}
__device__ __forceinline__ void st4(void* base, size_t off, int dt, const float4 v) {
  if (dt == DT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
  } else {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = pk;
  }
}
__device__ __forceinline__ float4 ld4_nc(const void* base, size_t off, int dt) {  // read-only for the whole launch
  if (dt == DT_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
  const uint2 x = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
  return make_float4(__uint_as_float(x.x << 16), __uint_as_float(x.x & 0xffff0000u), __uint_as_float(x.y << 16),
                     __uint_as_float(x.y & 0xffff0000u));
}
// v: accumulators of columns n..n+3 of output row `row`; bias/gamma already loaded for these columns; r = the
// residual values (pre-loaded by the caller for the whole chunk: `res` may alias `out0`, so the compiler cannot
// move those loads above earlier stores by itself and the 8 row groups would serialise on memory latency)
__device__ __forceinline__ void epilogue4(const Epilogue& e, size_t row, int n, float4 v, const float4 b,
                                          const float4 g, const float4 r) {
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  if (e.act == ACT_GELU) {
    v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w);
  } else if (e.act == ACT_SILU) {
    v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
  }
  v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
  const size_t off = row * (size_t)e.ldo + n;
  v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  if (e.add1) {  // the other two branch results: never written by this launch
    const float4 a = ld4_nc(e.add1, off, e.add_dt), c = ld4_nc(e.add2, off, e.add_dt);
    v.x = (v.x + a.x + c.x) * e.scale; v.y = (v.y + a.y + c.y) * e.scale;
    v.z = (v.z + a.z + c.z) * e.scale; v.w = (v.w + a.w + c.w) * e.scale;
  }
  if (e.out0) st4(e.out0, off, e.out0_dt, v);
  if (e.out1) {
    if (e.out1_silu) {
      v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
    }
    st4(e.out1, off, e.out1_dt, v);
  }
}

// ---------------------------------------------------------------- specialised epilogues
// The runtime-flag epilogue above costs ~30 instructions per element (flag tests, 64-bit address arithmetic, dtype
// dispatch) and the 8 epilogue warps cannot hide that; the variants below are the epilogues the hot path actually
// uses, with every flag a compile-time constant.  The host picks the variant (epilogue_variant()).
enum { EV_GENERIC = 0, EV_SILU_BF16, EV_RES_F32_BF16S, EV_RES_F32, EV_RES_MEAN_BF16S, EV_F32_BF16S, EV_GELU_BF16,
       EV_GAMMA_RES_F32, EV_F32, EV_BF16 };

static int epilogue_variant(const Epilogue& e) {
  const bool res32 = e.res && e.res_dt == DT_F32, o0f = e.out0 && e.out0_dt == DT_F32,
             o0b = e.out0 && e.out0_dt == DT_BF16, o1s = e.out1 && e.out1_dt == DT_BF16 && e.out1_silu;
  if (e.out1 && !o1s) return EV_GENERIC;
  if (e.res && !res32) return EV_GENERIC;
  if (e.add1) return (e.add_dt == DT_F32 && !e.act && !e.gamma && res32 && !e.out0 && o1s) ? EV_RES_MEAN_BF16S : EV_GENERIC;
  if (e.gamma) return (!e.act && res32 && o0f && !e.out1) ? EV_GAMMA_RES_F32 : EV_GENERIC;
  if (e.act == ACT_SILU) return (!e.res && o0b && !e.out1) ? EV_SILU_BF16 : EV_GENERIC;
  if (e.act == ACT_GELU) return (!e.res && o0b && !e.out1) ? EV_GELU_BF16 : EV_GENERIC;
  if (res32) return o0f ? (o1s ? EV_RES_F32_BF16S : (e.out1 ? EV_GENERIC : EV_RES_F32)) : EV_GENERIC;
  if (o0f) return o1s ? EV_F32_BF16S : (e.out1 ? EV_GENERIC : EV_F32);
  if (o0b && !e.out1) return EV_BF16;
  return EV_GENERIC;
}

__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, const float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

// One 32-row x CW-column chunk of one warp.  stg: the warp's XOR-swizzled transpose tile (already written);
// off0: element offset (row * ldo + n) of this lane's first row group; nvalid: row groups of this lane with t < T.
template <int V, int CW>
__device__ __forceinline__ void epilogue_chunk(const Epilogue& ep, const float* stg, int lane, size_t off0,
                                               int nvalid, const float4 b4, const float4 g4) {
  constexpr int CPR = CW / 4, RPI = 32 / CPR, NG = 32 / RPI;
  constexpr bool kRes = V == EV_RES_F32_BF16S || V == EV_RES_F32 || V == EV_RES_MEAN_BF16S || V == EV_GAMMA_RES_F32;
  constexpr bool kOut0F = V == EV_RES_F32_BF16S || V == EV_RES_F32 || V == EV_F32_BF16S || V == EV_GAMMA_RES_F32 || V == EV_F32;
  constexpr bool kOut0B = V == EV_SILU_BF16 || V == EV_GELU_BF16 || V == EV_BF16;
  constexpr bool kOut1 = V == EV_RES_F32_BF16S || V == EV_RES_MEAN_BF16S || V == EV_F32_BF16S;
  const int cg = lane % CPR, rsub = lane / CPR;
  const size_t stride = (size_t)RPI * ep.ldo;
  float4 r4[NG];
  if constexpr (kRes) {
    const float* rp = reinterpret_cast<const float*>(ep.res) + off0;
#pragma unroll
    for (int k = 0; k < NG; ++k)
      r4[k] = k < nvalid ? *reinterpret_cast<const float4*>(rp + k * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int r = k * RPI + rsub;
    const int rswz = CW == 32 ? (r & 7) : ((r >> 1) & 3);
    float4 v = *reinterpret_cast<const float4*>(stg + r * CW + ((cg ^ rswz) << 2));
    if (k < nvalid) {
      const size_t off = off0 + k * stride;
      v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
      if constexpr (V == EV_SILU_BF16) {
        v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
      }
      if constexpr (V == EV_GELU_BF16) {
        v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w);
      }
      if constexpr (V == EV_GAMMA_RES_F32) {
        v.x *= g4.x; v.y *= g4.y; v.z *= g4.z; v.w *= g4.w;
      }
      if constexpr (kRes) {
        v.x += r4[k].x; v.y += r4[k].y; v.z += r4[k].z; v.w += r4[k].w;
      }
      if constexpr (V == EV_RES_MEAN_BF16S) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.add1) + off));
        const float4 c = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.add2) + off));
        v.x = (v.x + a.x + c.x) * ep.scale; v.y = (v.y + a.y + c.y) * ep.scale;
        v.z = (v.z + a.z + c.z) * ep.scale; v.w = (v.w + a.w + c.w) * ep.scale;
      }
      if constexpr (kOut0F) *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + off) = v;
      if constexpr (kOut0B) st_bf16x4(reinterpret_cast<__nv_bfloat16*>(ep.out0) + off, v);
      if constexpr (kOut1) {
        v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
        st_bf16x4(reinterpret_cast<__nv_bfloat16*>(ep.out1) + off, v);
      }
    }
  }
}

// generic (runtime-flag) chunk
template <int CW>
__device__ __forceinline__ void epilogue_chunk_generic(const Epilogue& ep, const float* stg, int lane, size_t off0,
                                                       int nvalid, int n, const float4 b4, const float4 g4) {
  constexpr int CPR = CW / 4, RPI = 32 / CPR, NG = 32 / RPI;
  const int cg = lane % CPR, rsub = lane / CPR;
  const size_t stride = (size_t)RPI * ep.ldo;
  float4 r4[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k)
    r4[k] = (ep.res && k < nvalid) ? ld4(ep.res, off0 + k * stride, ep.res_dt) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int r = k * RPI + rsub;
    const int rswz = CW == 32 ? (r & 7) : ((r >> 1) & 3);
    const float4 v = *reinterpret_cast<const float4*>(stg + r * CW + ((cg ^ rswz) << 2));
    if (k < nvalid) epilogue4(ep, (off0 + k * stride - n) / ep.ldo, n, v, b4, g4, r4[k]);
  }
}

// ---------------------------------------------------------------- the kernel
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

template <int BN, int BK, int STAGES>
struct GemmTcSmem {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int CW = BN >= 64 ? 32 : 16;               // columns per epilogue chunk (each warp owns BN/2)
  static constexpr int STG_OFF = STAGES * (A_BYTES + B_BYTES);  // 8 warps x (32 rows x CW fp32) transpose buffers
  static constexpr int STG_BYTES = 8 * 32 * CW * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024 /*alignment slack*/;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvGemmShape s,
               Epilogue ep, int variant, int tiles_per_clip, int m_tiles, int n_tiles) {
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
    // TMEM -> registers (thread = row) -> XOR-swizzled per-warp smem tile -> (lane = 4 columns) -> global
    constexpr int CW = L::CW;            // chunk width in columns
    constexpr int CPR = CW / 4;          // 16-byte column groups per row (8 or 4)
    constexpr int RPI = 32 / CPR;        // rows covered by one warp-wide access (4 or 8)
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which half of the tile's columns it takes
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFF) + (warp - 2) * (32 * CW);
    const int cg = lane % CPR, rsub = lane / CPR;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int clip = m_blk / tiles_per_clip, t0 = (m_blk % tiles_per_clip) * 128, n0 = n_blk * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < (BN / 2) / CW; ++c) {
        const int col0 = half * (BN / 2) + c * CW;   // first column of this chunk within the tile
        uint32_t acc[CW];
        if constexpr (CW == 32) ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + col0, acc);
        else ptx::tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + col0, acc);
        // column parameters of this lane's 4 columns (independent of the row)
        const int n = n0 + col0 + cg * 4;
        const float4 b4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g4 = ep.gamma ? __ldg(reinterpret_cast<const float4*>(ep.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
        ptx::tmem_ld_wait();
        // thread = row `lane`: 16-byte group i goes to slot (i ^ swz(lane)); conflict-free for both access phases
        const int wswz = CW == 32 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
        for (int i = 0; i < CPR; ++i)
          *reinterpret_cast<uint4*>(stg + lane * CW + ((i ^ wswz) << 2)) =
              make_uint4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        __syncwarp();
        const int tb = t0 + q * 32 + rsub;                       // first row of this lane
        const int nvalid = min(32 / RPI, max(0, (s.T - tb + RPI - 1) / RPI));
        const size_t off0 = ((size_t)clip * s.T + tb) * (size_t)ep.ldo + n;
        switch (variant) {
          case EV_SILU_BF16: epilogue_chunk<EV_SILU_BF16, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_RES_F32_BF16S: epilogue_chunk<EV_RES_F32_BF16S, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_RES_F32: epilogue_chunk<EV_RES_F32, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_RES_MEAN_BF16S: epilogue_chunk<EV_RES_MEAN_BF16S, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_F32_BF16S: epilogue_chunk<EV_F32_BF16S, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_GELU_BF16: epilogue_chunk<EV_GELU_BF16, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_GAMMA_RES_F32: epilogue_chunk<EV_GAMMA_RES_F32, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_F32: epilogue_chunk<EV_F32, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          case EV_BF16: epilogue_chunk<EV_BF16, CW>(ep, stg, lane, off0, nvalid, b4, g4); break;
          default: epilogue_chunk_generic<CW>(ep, stg, lane, off0, nvalid, n, b4, g4); break;
        }
        __syncwarp();
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
    // epilogue signature: 1 act | 2 gamma | 4 residual | 8 mean3 | 16 fp32 out | 32 bf16 out
    const int esig = (e.act ? 1 : 0) | (e.gamma ? 2 : 0) | (e.res ? 4 : 0) | (e.add1 ? 8 : 0) | (e.out0 ? (e.out0_dt == DT_F32 ? 16 : 32) : 0) |
                     (e.out1 ? 32 : 0);
    const double out_bytes = (e.out0 ? (e.out0_dt == DT_F32 ? 4.0 : 2.0) : 0.0) + (e.out1 ? 2.0 : 0.0) +
                             (e.res ? 4.0 : 0.0) + (e.add1 ? 8.0 : 0.0);
    ProfScope ps(PC_GEMM_TC, 2.0 * macs, rows * s.C * 2.0 + (double)s.N * s.J * s.C * 2.0 + rows * s.N * out_bytes, st,
                 "C%d N%d J%d d%d e%d", s.C, s.N, s.J, s.dil, esig);
    gemm_tc_kernel<BN, BK, STAGES><<<grid, kGemmThreads, L::TOTAL, st>>>(tmA, tmB, s, e, epilogue_variant(e),
                                                                        tiles_per_clip, (int)m_tiles, n_tiles);
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
