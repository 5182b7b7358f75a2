// conv_post + tanh on the tensor cores (bf16 mode): Conv1d(32 -> 1, k13, pad 6) + tanh, models/generators.py:141-145.
//
// One output channel is no GEMM shape, and on the CUDA cores the layer sits at twice its FMA floor (pointwise.cu,
// 2.15 ms per step against 0.64 ms of HBM time).  Block-Toeplitz form over U = 8 consecutive samples instead: the
// contiguous (L, 32) bf16 input IS a (L/8, 256) matrix of "super-rows", and
//     y[8 q + r] = sum_{o in {0,1,2}} sum_{k < 256} S[q - 1 + o][k] * Wt[r][o * 256 + k],
//     Wt[r][o * 256 + i * 32 + c] = w[j][c],  j = 8 (o - 1) + i - r + 6  (zero outside 0 <= j < 13)
// i.e. M = 128 super-rows (1024 samples) x N x K = 768, with the three o taken as row-shifted UMMA descriptors into ONE
// activation tile of 130 super-rows (as conv_ws.cu does for taps).  N: tcgen05 wants N >= 16 at M = 128, so the second
// eight columns carry the bf16 ROUNDING RESIDUAL of the weights (Wt = hi + lo): the epilogue adds D[r] + D[8 + r] and the
// weights keep 16 mantissa bits for free.  48 MMAs per tile, each bound by its 4 KB activation-operand read (32 clk):
// 1536 clk per 1024 samples = 0.48 ms per step at 1.3 GHz, below the HBM time.  Measured in the step (256 x 10 s):
// 0.85 ms = 0.75 of the copy bandwidth, against 2.15-2.3 ms; agreement with the CUDA-core form < 2e-5 of range
// (tests/test_gpu_e2e.py::test_conv_post_on_the_tensor_cores_matches_the_cuda_core_form).
//   warp 0 TMA (weights once, then one tile per stage), warp 1 MMA issuer, warps 2-5 epilogue (lane = super-row).
#include "common.cuh"
#include "ptx.cuh"

#include <string.h>

#include <atomic>

namespace dc {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

static thread_local uint64_t g_launches_cp = 0;
uint64_t conv_post_launch_count() { return g_launches_cp; }

namespace cpt {
constexpr int U = 8, C = 32, K = 13, KC = U * C;   // 256 operand columns per super-row
constexpr int BOX_ROWS = 130;                      // 128 super-rows + one before and one after
constexpr int BOX_BYTES = BOX_ROWS * 128;          // one 64-column box, 128B-swizzled rows
constexpr int BOX_PITCH = 17 * 1024;               // 1024-aligned box bases
constexpr int A_STAGE = 4 * BOX_PITCH;             // four 64-column boxes = one activation tile
constexpr int STAGES = 2;
constexpr int W_TILE = 16 * 128;                   // one (o, box) weight tile: 16 rows x 64 columns
constexpr int W_BYTES = 12 * W_TILE;
constexpr int A_OFF = W_BYTES, BAR_OFF = A_OFF + STAGES * A_STAGE, TOTAL = BAR_OFF + 256 + 1024;
constexpr int ACCS = 4, THREADS = 64 + 4 * 32;
static_assert(TOTAL <= 232448, "shared memory budget");
}  // namespace cpt

// host: the Toeplitz weight matrix [16][768] (rows 0-7 = bf16(w), rows 8-15 = bf16(w - bf16(w))) from w[13][32]
void conv_post_toeplitz_weights(const float* w /*[13][32]*/, __nv_bfloat16* wt /*[16][768]*/) {
  using namespace cpt;
  for (int r = 0; r < U; ++r)
    for (int o = 0; o < 3; ++o)
      for (int i = 0; i < U; ++i)
        for (int c = 0; c < C; ++c) {
          const int j = U * (o - 1) + i - r + 6;
          const float v = (j >= 0 && j < K) ? w[j * C + c] : 0.f;
          const __nv_bfloat16 hi = __float2bfloat16_rn(v);
          const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
          const int k = o * KC + i * C + c;
          wt[r * 3 * KC + k] = hi;
          wt[(U + r) * 3 * KC + k] = lo;
        }
}

__global__ void __launch_bounds__(cpt::THREADS, 1)
conv_post_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, float bias,
                    float* __restrict__ out, int SR /*super-rows per clip*/, int tiles_per_clip, int total_tiles) {
  using namespace cpt;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, 16);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + A_OFF;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [STAGES]
  uint64_t* aempty = afull + STAGES;                               // [STAGES]
  uint64_t* dfull = aempty + STAGES;                               // [ACCS]
  uint64_t* dempty = dfull + ACCS;                                 // [ACCS]
  uint64_t* wbar = dempty + ACCS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        ptx::mbar_init(&afull[i], 1);
        ptx::mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < ACCS; ++i) {
        ptx::mbar_init(&dfull[i], 1);
        ptx::mbar_init(&dempty[i], 4);
      }
      ptx::mbar_init(wbar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<64>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, W_BYTES);
      for (int t = 0; t < 12; ++t) ptx::tma_load_2d(sW + t * W_TILE, &tmW, wbar, t * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int clip = tile / tiles_per_clip, q0 = (tile % tiles_per_clip) * 128;
        ptx::mbar_wait(&aempty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&afull[stage], 4 * BOX_BYTES);
        for (int b = 0; b < 4; ++b)   // rows q0 - 1 .. q0 + 128; outside [0, SR) the TMA zero-fills = the conv's padding
          ptx::tma_load_3d(sA + stage * A_STAGE + b * BOX_PITCH, &tmA, &afull[stage], b * 64, q0 - 1, clip);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      ptx::mbar_wait(wbar, 0);
      const uint32_t w_base = ptx::smem_u32(sW);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it % ACCS;
        ptx::mbar_wait(&dempty[acc], ((uint32_t)(it / ACCS) & 1u) ^ 1u);
        ptx::mbar_wait(&afull[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_base = ptx::smem_u32(sA + stage * A_STAGE);
        const uint32_t d = tmem_base + acc * 16;
        uint32_t accum = 0;
        for (int o = 0; o < 3; ++o)
          for (int b = 0; b < 4; ++b) {
            const uint32_t a = a_base + b * BOX_PITCH + o * 128;   // tile row p reads box row p + o
            const uint32_t w = w_base + (o * 4 + b) * W_TILE;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ptx::mma_bf16_ss(d, ptx::make_smem_desc<128>(a + k * 32), ptx::make_smem_desc<128>(w + k * 32), IDESC, accum);
              accum = 1;
            }
          }
        ptx::mma_commit(&aempty[stage]);
        ptx::mma_commit(&dfull[acc]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: lane = super-row = 8 consecutive samples
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int clip = tile / tiles_per_clip, q0 = (tile % tiles_per_clip) * 128;
      const int acc = it % ACCS;
      ptx::mbar_wait_sleepy(&dfull[acc], (uint32_t)(it / ACCS) & 1u);
      ptx::tc_fence_after();
      uint32_t v[16];
      ptx::tmem_ld_32x16(tmem_base + acc * 16 + ((uint32_t)(q * 32) << 16), v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&dempty[acc]);
      const int sr = q0 + q * 32 + lane;
      if (sr < SR) {
        float y[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) y[r] = tanhf(__uint_as_float(v[r]) + __uint_as_float(v[8 + r]) + bias);
        float4* o = reinterpret_cast<float4*>(out + ((size_t)clip * SR + sr) * 8);
        o[0] = make_float4(y[0], y[1], y[2], y[3]);
        o[1] = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<64>(tmem_base);
  }
}

bool conv_post_tc_supported(int L) { return L > 0 && L % cpt::U == 0; }

// in: (B, L, 32) bf16 = silu(x) of the last stage; wt: conv_post_toeplitz_weights() on the device
int launch_conv_post_tanh_tc(const __nv_bfloat16* in, const __nv_bfloat16* wt, float bias, float* out, int B, int L,
                             cudaStream_t st, int sm_count) {
  using namespace cpt;
  DC_CHECK(conv_post_tc_supported(L), DC_ERR_SHAPE, "conv_post (tensor-core form): L=%d must be a multiple of 8", L);
  static std::atomic<unsigned> attr_dev_mask{0u};
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_post_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  const int SR = L / U, tiles_per_clip = (SR + 127) / 128;
  const long long total = (long long)B * tiles_per_clip;
  DC_CHECK(total > 0 && total < (1ll << 31), DC_ERR_SHAPE, "conv_post: bad tile count");
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)KC, (uint64_t)SR, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)KC * 2, (uint64_t)SR * KC * 2};
    const uint32_t box[3] = {64, (uint32_t)BOX_ROWS, 1};
    DC_TRY(make_tmap_bf16(&tmA, in, 3, dims, strides, box, 128));
  }
  {
    const uint64_t dims[2] = {(uint64_t)3 * KC, 16};
    const uint64_t strides[1] = {(uint64_t)3 * KC * 2};
    const uint32_t box[2] = {64, 16};
    DC_TRY(make_tmap_bf16(&tmW, wt, 2, dims, strides, box, 128));
  }
  const int grid = (int)(total < sm_count ? total : sm_count);
  ProfScope ps(PC_CONV_POST, 2.0 * B * (double)L * 32 * 13, (double)B * L * (32.0 * 2 + 4.0), st);
  conv_post_tc_kernel<<<grid, THREADS, TOTAL, st>>>(tmA, tmW, bias, out, SR, tiles_per_clip, (int)total);
  ++g_launches_cp;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

}  // namespace dc
