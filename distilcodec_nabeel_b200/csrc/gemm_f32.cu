// fp32 CUDA-core shifted-row implicit GEMM (DC_MODE_FP32): the exact-precision path, parity target 1e-4.
//   out[b,t,n] = epi( sum_j sum_c A[b, t + shift0 + j*dil, c] * W[j*C + c][n] )
// Replaces, in fp32 mode, every Conv1d / ConvTranspose1d / Linear of the reference hot path
// (models/encoders.py:23-37, convnext_utils.py:250-255, grfvq.py:68-96, residual_vq.py:61-62,
//  generators.py:50-114, convnext_utils.py:36-102).
// 128 x BN output tile per 256-thread block, 8 x (BN/16) outputs per thread, BK = 16, register-prefetched
// double buffering through shared memory.  BN = 128 (8 x 8 per thread, packed FFMA2: 32 issue slots per 64 FMAs) where
// N allows; the narrow layers use BN = 64 / 32.  Accumulation order per output is k-ascending in every variant.
#include "common.cuh"

namespace dc {

static thread_local uint64_t g_launches_f32 = 0;
uint64_t gemm_f32_launch_count() { return g_launches_f32; }

template <int BN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       ConvGemmShape s, Epilogue ep, int tiles_per_clip) {
  constexpr int BM = 128, BK = 16, TM = 8, TN = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int clip = blockIdx.x / tiles_per_clip;
  const int t0 = (blockIdx.x % tiles_per_clip) * BM;
  const int n0 = blockIdx.y * BN;

  const float* Ab = A + (size_t)clip * s.T * s.C;
  // A loader: row r = tid/2 (0..127), 8 consecutive k starting at (tid&1)*8
  const int a_r = tid >> 1, a_k = (tid & 1) * 8;
  // B loader: row kk = tid / (BN/4), 4 consecutive n
  constexpr int B_TPR = BN / 4;  // threads per B row
  const int b_k = tid / B_TPR, b_n = (tid % B_TPR) * 4;
  const bool b_active = b_k < BK;

  const int kchunks = s.C / BK;
  const int total = s.J * kchunks;

  float2 acc[TM][TN / 2];  // column pairs: one packed FMA (fma.rn.f32x2) per pair, same IEEE result as two scalar FMAs
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  constexpr bool B2 = BN == 128;  // 32 threads cover a B row: each thread loads rows b_k and b_k + 8
  float4 ra0, ra1, rb, rb2;
  auto load_regs = [&](int it) {
    const int j = it / kchunks, c0 = (it % kchunks) * BK;
    const int t = t0 + a_r + s.shift0 + j * s.dil;
    if (t >= 0 && t < s.T) {
      const float4* p = reinterpret_cast<const float4*>(Ab + (size_t)t * s.C + c0 + a_k);
      ra0 = __ldg(p);
      ra1 = __ldg(p + 1);
    } else {
      ra0 = make_float4(0.f, 0.f, 0.f, 0.f);
      ra1 = ra0;
    }
    if (b_active)
      rb = __ldg(reinterpret_cast<const float4*>(W + (size_t)(j * s.C + c0 + b_k) * s.N + n0 + b_n));
    if constexpr (B2) rb2 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(j * s.C + c0 + b_k + 8) * s.N + n0 + b_n));
  };
  auto store_smem = [&](int buf) {
    As[buf][a_k + 0][a_r] = ra0.x; As[buf][a_k + 1][a_r] = ra0.y; As[buf][a_k + 2][a_r] = ra0.z; As[buf][a_k + 3][a_r] = ra0.w;
    As[buf][a_k + 4][a_r] = ra1.x; As[buf][a_k + 5][a_r] = ra1.y; As[buf][a_k + 6][a_r] = ra1.z; As[buf][a_k + 7][a_r] = ra1.w;
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
    if constexpr (B2) *reinterpret_cast<float4*>(&Bs[buf][b_k + 8][b_n]) = rb2;
  };

  load_regs(0);
  store_smem(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) load_regs(it + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (TN >= 4) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {  // column groups of 4 at a stride of 64: conflict-free LDS.128 (16 B per lane)
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][(j / 4) * 64 + tx * 4]);
          b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
        }
      } else {
        const float2 bv = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * TN]);
        b[0] = bv.x; b[1] = bv.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float2 ai = make_float2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc[i][j] = ffma2(ai, make_float2(b[2 * j], b[2 * j + 1]), acc[i][j]);
      }
    }
    if (it + 1 < total) store_smem(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int t = t0 + ty * TM + i;
    if (t >= s.T) continue;
    const size_t row = (size_t)clip * s.T + t;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = TN >= 4 ? (j / 4) * 64 + tx * 4 + (j & 3) : tx * TN + j;  // the column this accumulator belongs to
      epilogue_store(ep, row, n0 + col, (j & 1) ? acc[i][j / 2].y : acc[i][j / 2].x);
    }
  }
}

int launch_gemm_f32(const float* A, const float* W, const ConvGemmShape& s_in, const Epilogue& e, cudaStream_t st) {
  ConvGemmShape s = s_in;
  DC_CHECK(s.C % 16 == 0, DC_ERR_SHAPE, "gemm_f32: C=%d must be a multiple of 16", s.C);
  DC_CHECK(s.N % 32 == 0, DC_ERR_SHAPE, "gemm_f32: N=%d must be a multiple of 32", s.N);
  if (s.J == 1 && s.shift0 == 0) {  // no halo: flatten clips so tiles stay full
    s.T = s.B * s.T;
    s.B = 1;
  }
  const int tiles_per_clip = (s.T + 127) / 128;
  const long long mt = (long long)s.B * tiles_per_clip;
  DC_CHECK(mt > 0 && mt < (1ll << 31), DC_ERR_SHAPE, "gemm_f32: bad tile count");
  const double rows = (double)s.B * s.T;
  ProfScope ps(PC_GEMM_F32, 2.0 * rows * s.N * s.J * s.C * s.alg_scale,
               rows * s.C * 4.0 + (double)s.N * s.J * s.C * 4.0 + rows * s.N * 4.0, st);
#ifndef DC_F32_NO_BN128
  if (s.N % 128 == 0) {
    dim3 grid((unsigned)mt, s.N / 128);
    gemm_f32_kernel<128><<<grid, 256, 0, st>>>(A, W, s, e, tiles_per_clip);
  } else
#endif
  if (s.N % 64 == 0) {
    dim3 grid((unsigned)mt, s.N / 64);
    gemm_f32_kernel<64><<<grid, 256, 0, st>>>(A, W, s, e, tiles_per_clip);
  } else {
    dim3 grid((unsigned)mt, s.N / 32);
    gemm_f32_kernel<32><<<grid, 256, 0, st>>>(A, W, s, e, tiles_per_clip);
  }
  ++g_launches_f32;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

}  // namespace dc
