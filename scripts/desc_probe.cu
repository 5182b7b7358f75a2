// Probe (development tool, not product): which shared-memory descriptor encodings let tcgen05.mma read an A operand
// that starts at an arbitrary ROW of a larger TMA-written tile?  Needed for tap-shared implicit-GEMM convolutions
// (tap j of a conv = the same smem tile, j*dilation rows further down).
//   layout 0: SWIZZLE_128B tile, 64 bf16 per row (TMA 2-D box 64 x ROWS)
//   layout 1: SWIZZLE_64B  tile, 32 bf16 per row (TMA 2-D box 32 x ROWS)
//   layout 2: SWIZZLE_NONE "plane" tile: [K/8 planes][ROWS][8 bf16] (TMA 3-D box 8 x ROWS x K/8), SBO = 128 B
// For each row shift s and base_offset policy the kernel computes D = A[s : s+128, :] * I and the host checks it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../distilcodec_nabeel_b200/csrc/ptx.cuh"
using namespace dc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_enc get_enc() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return (PFN_enc)p;
}

constexpr int ROWS = 256;

// desc: explicit fields
__device__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}

template <int LAYOUT, int KC /*K elements*/>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, float* out, int shift,
                                                    int bo_policy) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                      // ROWS x KC bf16
  uint8_t* sB = smem + ROWS * KC * 2;      // KC x KC bf16 (N = KC), always swizzled K-major
  uint64_t* bar = (uint64_t*)(sB + 64 * 64 * 2 + 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<64>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(&bar[0], ROWS * KC * 2 + KC * KC * 2);
    if (LAYOUT == 2) ptx::tma_load_3d(sA, &tmA, &bar[0], 0, 0, 0);
    else ptx::tma_load_2d(sA, &tmA, &bar[0], 0, 0);
    ptx::tma_load_2d(sB, &tmB, &bar[0], 0, 0);
    ptx::mbar_wait(&bar[0], 0);
    ptx::tc_fence_after();
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    const uint32_t idesc = ptx::make_idesc_bf16(128, KC);
    for (int k = 0; k < KC / 16; ++k) {
      uint64_t da, db;
      if (LAYOUT == 0) {
        const uint32_t st = a0 + shift * 128 + k * 32;
        const uint32_t bo = bo_policy == 0 ? 0 : ((st >> 7) & 7);
        da = make_desc(st, 16, 1024, bo, 2);
        db = make_desc(b0 + k * 32, 16, 1024, 0, 2);
      } else if (LAYOUT == 1) {
        const uint32_t st = a0 + shift * 64 + k * 32;
        const uint32_t bo = bo_policy == 0 ? 0 : (bo_policy == 1 ? ((st >> 7) & 7) : ((st >> 7) & 3));
        da = make_desc(st, 16, 512, bo, 4);
        db = make_desc(b0 + k * 32, 16, 512, 0, 4);
      } else {
        // planes of ROWS x 16 B; one K=16 step = 2 planes; LBO = plane stride, SBO = 8 rows x 16 B
        const uint32_t st = a0 + (2 * k) * (ROWS * 16) + shift * 16;
        da = bo_policy == 0 ? make_desc(st, ROWS * 16, 128, 0, 0) : make_desc(st, 128, ROWS * 16, 0, 0);
        db = KC == 64 ? make_desc(b0 + k * 32, 16, 1024, 0, 2) : make_desc(b0 + k * 32, 16, 512, 0, 4);
      }
      ptx::mma_bf16_ss(tmem, da, db, idesc, k ? 1u : 0u);
    }
    ptx::mma_commit(&bar[1]);
  }
  __syncwarp();
  ptx::mbar_wait(&bar[1], 0);
  ptx::tc_fence_after();
  uint32_t acc[32];
  for (int c = 0; c < KC / 32; ++c) {
    ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, acc);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * KC + c * 32 + i] = __uint_as_float(acc[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<64>(tmem);
}

template <int LAYOUT, int KC>
static void run(PFN_enc enc, const char* name) {
  std::vector<__nv_bfloat16> hA(ROWS * KC), hB(KC * KC);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < KC; ++c) hA[r * KC + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251) - 125.f);
  for (int n = 0; n < KC; ++n)
    for (int k = 0; k < KC; ++k) hB[n * KC + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dO, 128 * KC * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  {
    CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t gd[2] = {(cuuint64_t)KC, (cuuint64_t)KC}, gs[1] = {(cuuint64_t)KC * 2};
    cuuint32_t bx[2] = {(cuuint32_t)KC, (cuuint32_t)KC}, es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); exit(1); }
  }
  if (LAYOUT == 2) {
    cuuint64_t gd[3] = {8, (cuuint64_t)ROWS, (cuuint64_t)KC / 8}, gs[2] = {(cuuint64_t)KC * 2, 16};
    cuuint32_t bx[3] = {8, (cuuint32_t)ROWS, (cuuint32_t)KC / 8}, es[3] = {1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A (plane) failed %d\n", (int)r); exit(1); }
  } else {
    CUtensorMapSwizzle sw = LAYOUT == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t gd[2] = {(cuuint64_t)KC, (cuuint64_t)ROWS}, gs[1] = {(cuuint64_t)KC * 2};
    cuuint32_t bx[2] = {(cuuint32_t)KC, (cuuint32_t)ROWS}, es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); exit(1); }
  }
  const int smem = ROWS * KC * 2 + 64 * 64 * 2 + 1024 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<LAYOUT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> hO(128 * KC);
  const int shifts[] = {0, 1, 2, 3, 5, 8, 11, 16, 25, 50, 127};
  for (int pol = 0; pol < (LAYOUT == 1 ? 3 : 2); ++pol) {
    printf("%s policy %d:", name, pol);
    for (int s : shifts) {
      CK(cudaMemset(dO, 0, 128 * KC * 4));
      probe_kernel<LAYOUT, KC><<<1, 128, smem>>>(tmA, tmB, dO, s, pol);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" shift %d: CUDA error %s\n", s, cudaGetErrorString(e)); exit(2); }
      CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int i = 0; i < 128; ++i)
        for (int c = 0; c < KC; ++c)
          if (hO[i * KC + c] != __bfloat162float(hA[(s + i) * KC + c])) ++bad;
      printf(" s=%d:%s", s, bad ? "BAD" : "ok");
      if (bad) printf("(%d)", bad);
    }
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main() {
  PFN_enc enc = get_enc();
  run<0, 64>(enc, "SW128 (64 ch rows)");
  run<1, 32>(enc, "SW64  (32 ch rows)");
  run<2, 64>(enc, "NOSWZ planes K=64 ");
  run<2, 32>(enc, "NOSWZ planes K=32 ");
  return 0;
}
