// Shared declarations for the DistilCodec B200 hot-path library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/distilcodec_b200.h"

namespace dc {

// ---------------------------------------------------------------- error handling (never throws across the ABI)
void set_error(const char* fmt, ...);
#define DC_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      dc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return DC_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)
#define DC_CHECK(cond, code, ...)                                                              \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      dc::set_error(__VA_ARGS__);                                                              \
      return (code);                                                                           \
    }                                                                                          \
  } while (0)
#define DC_TRY(expr)                                                                           \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r != DC_OK) return _r;                                                                \
  } while (0)

// DT_SPLIT: a (rows, 2C) bf16 tensor holding the two-term split [hi C | mid C] of fp32 values (4 bytes per value like fp32):
// the operand format of the fp32 tensor-core kernel (gemm_f32x.cu), written directly by the producing epilogue
enum { DT_F32 = 0, DT_BF16 = 1, DT_SPLIT = 2 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

// ---------------------------------------------------------------- the one GEMM shape every dense layer maps to
// out[b, t, n] = epi( sum_{j<J} sum_{c<C} A[b, t + shift0 + j*dil, c] * W[n, j*C + c] )
// A is channels-last (B, T, C); rows outside [0, T) read as zero (= the conv's zero padding).
struct ConvGemmShape {
  int B, T, C, J, shift0, dil, N;
  float alg_scale = 1.f;  // algorithmic / executed MACs (< 1 for ConvTranspose1d phase GEMMs with zero-padded taps)
  // ConvTranspose1d phase GEMMs: output columns [ph*phase_cols, (ph+1)*phase_cols) belong to stride phase ph, and
  // bit (ph*J + j) of zero_taps says that tap j's weights are all zero for that phase (the tap is skipped when a
  // whole N tile lies inside one phase)
  int phase_cols = 0;
  uint32_t zero_taps = 0;
  int cluster = 1;  // 2: kernels that support it run as CTA pairs sharing (TMA-multicasting) their weight tiles
};

// Runtime epilogue, applied per output element v = acc:
//   v += bias[n]; v = act(v); v *= gamma[n]; v += res[row, n]; if (add1) v = (v + add1 + add2) * scale;
//   out0[row, n] = v;  out1[row, n] = out1_silu ? silu(v) : v
// row = b*T + t, all row-major with pitch ldo.
struct Epilogue {
  const float* bias = nullptr;
  const float* gamma = nullptr;
  const void* res = nullptr;
  const void* add1 = nullptr;
  const void* add2 = nullptr;
  void* out0 = nullptr;
  void* out1 = nullptr;
  int act = ACT_NONE;
  int res_dt = DT_F32, add_dt = DT_BF16, out0_dt = DT_F32, out1_dt = DT_BF16;
  int ldo = 0;
  float scale = 1.f;
  int out1_silu = 1;
  int prefetch = 1;  // L2-prefetch res / add1 / add2 of the next tile while waiting for its MMAs
  // DT_SPLIT outputs: channels per LOGICAL row of the consumer's tensor (a ConvTranspose1d GEMM row holds `stride` output
  // samples of split_seg = C_out channels each; 0 = ldo)
  int split_seg = 0;
};

// Packed fp32 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, two IEEE fp32 operations per instruction, same rounding as
// the scalar forms) for the CUDA-core kernels whose fp32 ALU work rivals their HBM time on B200.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float ld_as_f32(const void* p, size_t i, int dt) {
  return dt == DT_F32 ? reinterpret_cast<const float*>(p)[i]
                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_from_f32(void* p, size_t i, int dt, float v) {
  if (dt == DT_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// scalar epilogue (used by the fp32 CUDA-core kernel and as the definition the vector paths follow)
__device__ __forceinline__ void epilogue_store(const Epilogue& e, size_t row, int n, float v) {
  if (e.bias) v += e.bias[n];
  if (e.act == ACT_GELU) v = gelu_erf_f(v);
  else if (e.act == ACT_SILU) v = silu_f(v);
  if (e.gamma) v *= e.gamma[n];
  size_t o = row * (size_t)e.ldo + n;
  if (e.res) v += ld_as_f32(e.res, o, e.res_dt);
  if (e.add1) v = (v + ld_as_f32(e.add1, o, e.add_dt) + ld_as_f32(e.add2, o, e.add_dt)) * e.scale;
  if (e.out0) st_from_f32(e.out0, o, e.out0_dt, v);
  if (e.out1) st_from_f32(e.out1, o, e.out1_dt, e.out1_silu ? silu_f(v) : v);
}

// ---------------------------------------------------------------- launchers implemented across the .cu files
// fp32 CUDA-core implicit GEMM.  A fp32 (B,T,C); W fp32 [J*C][N] (N contiguous).
int launch_gemm_f32(const float* A, const float* W, const ConvGemmShape& s, const Epilogue& e, cudaStream_t st);
// fp32-accurate tensor-core implicit GEMM (gemm_f32x.cu): A2 bf16 (B,T,2C) = [hi | mid] split of the fp32 activations
// (launch_split_f32), W2 bf16 [N][J][2C] = the same split of the weights; chunked accumulation, fp32 epilogue.
bool gemm_f32x_supported(const ConvGemmShape& s);
int launch_split_f32(const float* in, __nv_bfloat16* out, size_t rows, int C, cudaStream_t st);
int launch_gemm_f32x(const __nv_bfloat16* A2, const __nv_bfloat16* W2, const ConvGemmShape& s, const Epilogue& e,
                     cudaStream_t st, int sm_count);
uint64_t gemm_f32x_launch_count();
// bf16 tcgen05 implicit GEMM.  A bf16 (B,T,C); W bf16 [N][J*C] (K contiguous).
int launch_gemm_tc(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                   cudaStream_t st, int sm_count);
size_t gemm_tc_launch_count();
// weights-stationary, tap-shared variant for C in {32, 64}, N <= 64 (conv_ws.cu); same operands as launch_gemm_tc
bool conv_ws_supported(const ConvGemmShape& s);
int launch_conv_ws(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                   cudaStream_t st, int sm_count);
uint64_t conv_ws_launch_count();
// fused ResBlock step  out = epi2(conv2(silu(conv1(S) + bias1))): conv1's output stays in shared memory
bool conv_ws_pair_supported(const ConvGemmShape& s1, const ConvGemmShape& s2);
int launch_conv_ws_pair(const __nv_bfloat16* S, const __nv_bfloat16* W1, const __nv_bfloat16* W2, const float* bias1,
                        const ConvGemmShape& s1, const ConvGemmShape& s2, const Epilogue& e2, cudaStream_t st,
                        int sm_count);
// fused ResBlock step for C = 32 with the fp32 activation stream in / out and block-Toeplitz ("phase form") weights
// (conv_pair.cu): input either X fp32 (B,T,32) (the kernel forms silu(x) itself) or S = silu(x) bf16 (B,T,32) by TMA;
// W1 phase form [64][(J+1)*32] if s1.dil == 1 else row form [32][J*32]; W2p phase form
bool conv_pairx_supported(const ConvGemmShape& s1, const ConvGemmShape& s2);
int launch_conv_pairx(const float* X, const __nv_bfloat16* S, const __nv_bfloat16* W1, const __nv_bfloat16* W2p,
                      const float* bias1, const float* bias2x, const ConvGemmShape& s1, const ConvGemmShape& s2,
                      const Epilogue& e2, cudaStream_t st, int sm_count);
uint64_t conv_pairx_launch_count();
// tap-shared 256-row-tile variant for C = N = 128 (conv_ts.cu); same operands as launch_gemm_tc
bool conv_ts_supported(const ConvGemmShape& s);
int launch_conv_ts(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                   cudaStream_t st, int sm_count);
uint64_t conv_ts_launch_count();
// tap-shared variant for the wide convs (C >= 256, N % 256 == 0, J > 1)
bool conv_tsw_supported(const ConvGemmShape& s);
int launch_conv_tsw(const __nv_bfloat16* A, const __nv_bfloat16* W, const ConvGemmShape& s, const Epilogue& e,
                    cudaStream_t st, int sm_count);

// pointwise / bandwidth-bound kernels (pointwise.cu)
int launch_transpose_ncl_to_nlc(const float* in, void* out, int out_dt, int B, int C, int T, cudaStream_t st);
int launch_transpose_nlc_to_ncl(const float* in, float* out, int B, int T, int C, cudaStream_t st);
// dwconv k7 (zero pad 3) + LayerNorm over C, or plain LayerNorm (dw_w == nullptr). in fp32 (B,T,C).
int launch_dwconv_ln(const float* in, const float* dw_w /*[7][C]*/, const float* dw_b, const float* ln_w,
                     const float* ln_b, void* out, int out_dt, int B, int T, int C, cudaStream_t st);
int launch_cast(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t st);
int launch_gather_rows(const float* table, const int64_t* idx, int64_t nrows, int D, int64_t table_rows,
                       float* out_f32 /*nullable*/, __nv_bfloat16* out_bf16 /*nullable*/, cudaStream_t st);
int launch_conv_post_tanh(const void* in, int in_dt, const float* w_host /*[13][32], HOST memory*/, float bias,
                          float* out, int B, int L, cudaStream_t st);
// conv_post.cu: the same layer on the tensor cores (bf16 input; block-Toeplitz over 8 samples)
void conv_post_toeplitz_weights(const float* w /*[13][32], host*/, __nv_bfloat16* wt /*[16][768], host*/);
bool conv_post_tc_supported(int L);
int launch_conv_post_tanh_tc(const __nv_bfloat16* in, const __nv_bfloat16* wt /*device*/, float bias, float* out, int B,
                             int L, cudaStream_t st, int sm_count);
uint64_t conv_post_launch_count();
// weight prepack helpers
int launch_weight_norm_fold(const float* g, const float* v, float* w, int dim0, int inner, cudaStream_t st);
struct PackDesc {
  int N, J, C;            // packed GEMM weight is N x (J*C)
  int phases;             // N = phases * n_inner
  long long s_n, s_c, s_k;  // source strides (elements) for (n_inner, c, kernel tap)
  int kmap[8 * 16];       // kmap[phase*J + j] = source kernel tap or -1 (zero)
  int split;              // 1: the bf16 output is [N][J][2C], channels [0,C) = bf16(w), [C,2C) = bf16(w - bf16(w)) (gemm_f32x.cu)
};
int launch_pack_weight(const float* src, const PackDesc& d, float* out_kn_f32 /*nullable*/,
                       __nv_bfloat16* out_nk_bf16 /*nullable*/, cudaStream_t st);
int launch_row_sqnorm(const float* in, float* out, int64_t rows, int D, cudaStream_t st);

// VQ (vq.cu)
size_t vq_workspace_bytes(int64_t nrows, int D, bool x_is_bf16);
int launch_vq_search(const void* x, int x_dt, const float* x2_opt, int64_t nrows, int D, const float* codebook_f32,
                     const __nv_bfloat16* codebook_bf16, const float* c2, const float* c2max_dev, int K,
                     int64_t* codes, void* ws, size_t ws_bytes, float window_factor, bool use_tc, bool x2_exact,
                     cudaStream_t st, int sm_count, int* stats_host_opt, int cta_pairs = 1);
int launch_vq_resid_max(const float* cb, int64_t rows, int D, float* out /*1 float*/, cudaStream_t st);
int launch_row_sqnorm_torch_order(const float* in, float* out, int64_t rows, int D, cudaStream_t st);
uint64_t vq_launch_count();
uint64_t pointwise_launch_count();
uint64_t gemm_f32_launch_count();

// log-mel front-end (mel.cu): audio (B, Ls) fp32 -> log-mel (B, 128, T), T = (Ls - 256) / 256 + 1
int launch_mel(const float* audio, const float* window /*[1024]*/, const float* twiddle /*[1024][2]*/,
               const float* fb /*[513][128]*/, float* mel, int B, int Ls, cudaStream_t st);
int mel_twiddles_host(float* out /*2 * 1024 floats*/);
uint64_t mel_launch_count();

int sm_count_of_current_device();

// ---------------------------------------------------------------- per-kernel-class timing (bench.py's roofline)
// Thread-local and off by default.  When enabled (dc_profile_enable) every launcher brackets its kernel with a
// CUDA event pair on the launching stream and books the launch's ALGORITHMIC flops / HBM bytes.
enum ProfClass {
  PC_GEMM_TC = 0, PC_GEMM_F32, PC_VQ_SCORE, PC_VQ_PREP, PC_VQ_RESCORE, PC_VQ_EXHAUSTIVE, PC_DWCONV_LN, PC_LAYERNORM,
  PC_CAST, PC_GATHER, PC_TRANSPOSE, PC_CONV_POST, PC_PREPACK, PC_CONV_WS, PC_MEL, PC_CONV_TS, PC_COUNT
};
struct ProfScope {
  int idx = -1;
  cudaStream_t st;
  ProfScope(int cls, double flops, double bytes, cudaStream_t stream, const char* shape_fmt = nullptr, ...);
  ~ProfScope();
};

}  // namespace dc
