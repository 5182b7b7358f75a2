// Fused nearest-code search of the single 32768 x 3584 codebook (the headline kernel of the hot path).
//
// Replaces EuclideanCodebook.forward's eval path (vector_quantization/utils/vector_quantize_pytorch.py:462-538):
//     dist = -sqrt(clamp((x2 + c2_j) + (-2 * x.c_j), 0))   (cdist :41-45, fp32, autocast disabled :462/:473)
//     ind  = argmax_j dist                                   (gumbel_sample eval branch :96; first max wins)
// without ever materialising the N x 32768 distance matrix.
//
// Pass 0  vq_prep_kernel     per row: exact ||x||^2 (fp64 -> fp32), bf16 copy of x when x is fp32, and the candidate
//                            window W = wf * 2e + G where e bounds the bf16-operand error of the tensor-core score
//                            (Cauchy-Schwarz: 2 * u * ||x|| * max||c||) and G = 6 ulp(d^2) covers the reference's own
//                            fp32 rounding grid (ties!).
// Pass 1  vq_score_kernel    persistent tcgen05 GEMM  S = c2_j - 2 * <bf16 x, bf16 c_j>, 256 x 256 tiles, two 128-lane
//                            fp32 accumulators in TMEM (all 512 columns), TMA-fed 3-stage ring of (A 32 KB + B 32 KB).
//                            The epilogue keeps, per row, the running minimum and appends every (j, S_j) with
//                            S_j <= running_min + W to a private candidate list.  Nothing else reaches HBM.
// Pass 2  vq_rescore_kernel  per row: final minimum over the row's lists, survivors S_j <= min + W are re-scored with
//                            the reference's exact fp32 expression (x.c_j accumulated in fp64, rounded once to fp32),
//                            smallest distance wins, lowest index on ties.  Rows whose list overflowed go to pass 3.
// Pass 3  vq_exhaustive_*    exact scan of the whole codebook for overflowed rows (degenerate inputs only).
#include "common.cuh"

#include <string.h>
#include "ptx.cuh"

#include <math.h>

namespace dc {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

static thread_local uint64_t g_launches_vq = 0;
uint64_t vq_launch_count() { return g_launches_vq; }

constexpr int VQ_BM = 256;      // rows per tile (two UMMA M=128 accumulators)
constexpr int VQ_BN = 256;      // codes per tile
constexpr int VQ_BK = 64;       // K elements per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int VQ_STAGES = 3;
constexpr int VQ_CAP = 32;      // candidate slots per (row, codebook split)
constexpr int VQ_NS_MAX = 8;    // max codebook splits
constexpr int VQ_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct VqSmem {
  static constexpr int A_BYTES = VQ_BM * VQ_BK * 2;
  static constexpr int B_BYTES = VQ_BN * VQ_BK * 2;
  static constexpr int BAR_OFF = VQ_STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = BAR_OFF + (2 * VQ_STAGES + 2) * 8 + 16 + 1024;
};

// ---------------------------------------------------------------- workspace layout
struct VqWs {
  __nv_bfloat16* xb;   // rows x D (only when x arrives as fp32)
  float* x2e;          // rows
  float* win;          // rows
  float* best;         // rows x NS_MAX
  int* cnt;            // rows x NS_MAX
  int2* cand;          // rows x NS_MAX x CAP
  int* ovf_rows;       // rows
  unsigned long long* ovf_keys;  // rows
  int* counters;       // [0] overflow rows, [1] rescored candidates, [2] rows with a single survivor, [3] spare
  size_t total;
};
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static VqWs vq_carve(void* base, int64_t rows, int D, bool need_xb) {
  VqWs w;
  size_t off = 0;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.xb = reinterpret_cast<__nv_bfloat16*>(take(need_xb ? (size_t)rows * D * 4 : 0));   // fp32 rows: [hi | lo] bf16 split
  w.x2e = reinterpret_cast<float*>(take((size_t)rows * 4));
  w.win = reinterpret_cast<float*>(take((size_t)rows * 4));
  w.best = reinterpret_cast<float*>(take((size_t)rows * VQ_NS_MAX * 4));
  w.cnt = reinterpret_cast<int*>(take((size_t)rows * VQ_NS_MAX * 4));
  w.cand = reinterpret_cast<int2*>(take((size_t)rows * VQ_NS_MAX * VQ_CAP * 8));
  w.ovf_rows = reinterpret_cast<int*>(take((size_t)rows * 4));
  w.ovf_keys = reinterpret_cast<unsigned long long*>(take((size_t)rows * 8));
  w.counters = reinterpret_cast<int*>(take(64));
  w.total = off;
  return w;
}
size_t vq_workspace_bytes(int64_t nrows, int D, bool x_is_bf16) {
  return vq_carve(nullptr, nrows > 0 ? nrows : 1, D, !x_is_bf16).total;
}

// ---------------------------------------------------------------- pass 0
// ||x||^2 exactly as the reference's CPU path produces it: `(x ** 2).sum(-1)` in cdist
// (vector_quantize_pytorch.py:42) is ATen's cascade_sum over the contiguous row (SumKernel.cpp, the AVX2 build
// that x86 hosts dispatch to): 8 vector lanes x 4 interleaved accumulators = 32 independent chains, chain
// (k, l) owning elements ((i*4 + k)*8 + l); every chain adds 16 squares into a level-0 accumulator, spills it
// into level 1 (and on into levels 2/3 every 16^2 / 16^3 steps), then level sums, then the 4 interleaved
// accumulators, then the 8 lanes are added in order.  With x2 of magnitude ~600 one ulp of it is as large as
// ||c||^2, so reproducing this order is what makes the fp32 argmax (ties included) land on the reference's index:
// 100 % of 8192 golden rows, against 99.5 % with a correctly rounded x2 (tests/test_vq_math.py).
// One warp per row, lane = chain, so the loads are fully coalesced.
template <typename LoadF>
__device__ __forceinline__ float torch_cpu_row_sqsum(LoadF load, int D, int lane) {
  const int size = D >> 5;  // steps per chain
  int lg = 0;
  while ((1 << lg) < size) ++lg;
  const int level_power = max(4, lg / 4);
  const int level_step = 1 << level_power;
  const int level_mask = level_step - 1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int i = 0;
  while (i + level_step <= size) {
    for (int j = 0; j < level_step; ++j, ++i) {
      const float v = load(i * 32 + lane);
      acc[0] = __fadd_rn(acc[0], __fmul_rn(v, v));
    }
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      acc[j] = __fadd_rn(acc[j], acc[j - 1]);
      acc[j - 1] = 0.f;
      if ((i & (level_mask << (j * level_power))) != 0) break;
    }
  }
  for (; i < size; ++i) {
    const float v = load(i * 32 + lane);
    acc[0] = __fadd_rn(acc[0], __fmul_rn(v, v));
  }
#pragma unroll
  for (int j = 1; j < 4; ++j) acc[0] = __fadd_rn(acc[0], acc[j]);
  // interleaved accumulators k = 1..3 live in lanes l + 8k
  float s = acc[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) s = __fadd_rn(s, __shfl_sync(0xffffffffu, acc[0], (lane & 7) + 8 * k));
  // the 8 vector lanes, in order, starting from 0.f
  float f = 0.f;
#pragma unroll
  for (int l = 0; l < 8; ++l) f = __fadd_rn(f, __shfl_sync(0xffffffffu, s, l));
  return f;
}

template <bool X_BF16>
__global__ void __launch_bounds__(256) vq_prep_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                                      float* __restrict__ x2e, float* __restrict__ win,
                                                      const float* __restrict__ c2max_p, int64_t rows, int D,
                                                      float window_factor, int x2_exact,
                                                      int* __restrict__ counters) {
  if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x2, dx2 = 0.f;   // dx2 = ||x - bf16(x)||^2 (0 when x already is bf16)
  if constexpr (X_BF16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)row * D;
    if (x2_exact) {
      double s = 0.0;
      for (int i = lane; i < D; i += 32) {
        const double a = (double)__bfloat162float(p[i]);
        s = fma(a, a, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      x2 = (float)s;
    } else {
      x2 = torch_cpu_row_sqsum([&](int e) { return __bfloat162float(p[e]); }, D, lane);
    }
  } else {
    const float* p = reinterpret_cast<const float*>(x) + (size_t)row * D;
    // Full-precision rows are scored as a two-term bf16 split x = hi + lo + r (K-extended GEMM: [hi | lo] against the
    // codebook tile twice), so that the rounding of x itself (|r| <= 2^-17 |x|) all but vanishes from the candidate
    // window; with hi alone ~1 % of the W0 rows overflowed the candidate lists into the exhaustive pass (11 % of the
    // fp32-mode step).
    __nv_bfloat16* o = xb + (size_t)row * 2 * D;
    for (int i = lane; i < D; i += 32) {
      const float v = __ldg(p + i);
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const float r1 = v - __bfloat162float(hi);            // exact (Sterbenz)
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1);
      o[i] = hi;
      o[D + i] = lo;
      const float d = r1 - __bfloat162float(lo);            // exact: what the split does not represent
      dx2 = fmaf(d, d, dx2);
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) dx2 += __shfl_xor_sync(0xffffffffu, dx2, o2);
    if (x2_exact) {
      double s = 0.0;
      for (int i = lane; i < D; i += 32) {
        const double a = (double)__ldg(p + i);
        s = fma(a, a, s);
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
      x2 = (float)s;
    } else {
      x2 = torch_cpu_row_sqsum([&](int e) { return __ldg(p + e); }, D, lane);
    }
  }
  if (lane == 0) {
    // Candidate window of the bf16 tensor-core scorer.  S~_j = c2_j - 2 <x~, c~_j> (x~, c~ = bf16 roundings, fp32
    // accumulate) differs from the exact s_j = c2_j - 2 <x, c_j> by 2 (<x - x~, c_j> + <x~, c_j - c~_j>) + accumulation,
    // so by Cauchy-Schwarz   |S~_j - s_j| <= e := 2 (||x - x~|| ||c||max + ||x~|| ||c - c~||max) + e_acc
    // with the EXACT rounding-residual norms: ||x - x~|| of this row (above) and max_j ||c_j - c~_j|| of the codebook
    // (dc_finalize).  Those are ~0.3 of the generic 2^-8 |v| bound, which is why a window of 2 e is both rigorous and
    // small enough for the 32-slot candidate lists on every weight set of the tests.  window_factor scales it
    // (option "vq_window", 1.0 = this bound); 6 ulp(d^2) cover the fp32 evaluation of the reference's own expression.
    // Full-precision rows are scored as the two-term split above: x~ = hi + lo, ||x - x~|| <= 2^-17 ||x||.
    const float c2max = c2max_p[0], r2max = c2max_p[1];
    const float xn = sqrtf(x2) * 1.005f /* ||hi|| + ||lo|| */, cn = sqrtf(c2max), dx = sqrtf(dx2) * 1.001f,
                rn = sqrtf(r2max) * 1.001f;
    const float e_acc = 2.f * (float)D * (X_BF16 ? 1.f : 2.f) * 1.1920929e-7f * xn * cn;   // fp32 accumulation over D (2 D) products
    const float e = 2.f * (dx * cn + xn * rn) + e_acc;
    const float dmax2 = x2 + c2max + 2.f * xn * cn;     // upper bound of the reference's d^2
    int ex = 0;
    frexpf(fmaxf(dmax2, 1e-30f), &ex);                  // dmax2 = m * 2^ex, m in [0.5, 1)
    const float ulp = ldexpf(1.f, ex - 24);
    x2e[row] = x2;
    win[row] = window_factor * 2.f * e + 6.f * ulp;
  }
}

// max_j ||c_j - bf16(c_j)||^2 of the codebook (dc_finalize): the second factor of the candidate-window bound above.
// out must be zeroed; non-negative floats order like their bit patterns.
__global__ void __launch_bounds__(256) vq_resid_max_kernel(const float* __restrict__ cb, int64_t rows, int D,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = cb + (size_t)row * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float v = __ldg(p + i);
    const float d = v - __bfloat162float(__float2bfloat16_rn(v));
    s = fmaf(d, d, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(s));
}
int launch_vq_resid_max(const float* cb, int64_t rows, int D, float* out, cudaStream_t st) {
  DC_CUDA(cudaMemsetAsync(out, 0, 4, st));
  if (rows == 0) return DC_OK;
  ProfScope ps(PC_PREPACK, 0, 0, st);
  vq_resid_max_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(cb, rows, D, out);
  ++g_launches_vq;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ||c_j||^2 of the codebook in the same ATen order (`(y ** 2).sum(-1)`, vector_quantize_pytorch.py:43)
__global__ void __launch_bounds__(256) vq_row_sqnorm_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = in + (size_t)row * D;
  const float v = torch_cpu_row_sqsum([&](int e) { return __ldg(p + e); }, D, lane);
  if (lane == 0) out[row] = v;
}
int launch_row_sqnorm_torch_order(const float* in, float* out, int64_t rows, int D, cudaStream_t st) {
  if (rows == 0) return DC_OK;
  DC_CHECK(D % 32 == 0, DC_ERR_SHAPE, "row_sqnorm: D=%d must be a multiple of 32", D);
  ProfScope ps(PC_PREPACK, 0, 0, st);
  vq_row_sqnorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, out, rows, D);
  ++g_launches_vq;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ---------------------------------------------------------------- pass 1: tcgen05 score + windowed candidates
__device__ __noinline__ void vq_append(int2* list, int& cnt, float thr, int idx, float s) {
  if (cnt > VQ_CAP) return;  // already overflowed: the row goes to the exhaustive pass
  if (cnt == VQ_CAP) {       // drop entries that can no longer be within the window (thr only decreases)
    int j = 0;
    for (int e = 0; e < VQ_CAP; ++e) {
      const int2 ent = list[e];
      if (__int_as_float(ent.y) <= thr) list[j++] = ent;
    }
    cnt = j;
    if (cnt == VQ_CAP) {
      cnt = VQ_CAP + 1;
      return;
    }
  }
  list[cnt++] = make_int2(idx, __float_as_int(s));
}

// CL = 2: thread-block clusters of two CTAs score adjacent 256-row blocks against the SAME codebook tiles; each CTA loads
// one half (128 codes) of every codebook tile and TMA-multicasts it into both CTAs' rings, so the codebook traffic from L2
// per CTA halves (operand traffic 64 -> 48 KB per stage).  A stage is released to both producers (multicast commit).
template <int CL>
__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_score_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmC,
                const float* __restrict__ c2, const float* __restrict__ win, float* __restrict__ best_out,
                int* __restrict__ cnt_out, int2* __restrict__ cand_out, int nrows, int n_items, int NS,
                int n_tiles, int tiles_per_item, int kblocks, int kblocks_c /*K blocks of the codebook: kblocks, or
                kblocks / 2 when the rows are a [hi | lo] split*/) {
  using L = VqSmem;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, VQ_BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + VQ_STAGES * L::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + VQ_STAGES;
  uint64_t* tfull = empty + VQ_STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // work items: (256-row block, codebook split); with CL = 2 an item is a PAIR of row blocks and this CTA takes block
  // 2 * (item / NS) + rank (a block past the last row is a ghost: zero rows, nothing recorded)
  const int rank = CL == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int worker = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto item_mblk = [&](int item) { return CL == 2 ? 2 * (item / NS) + rank : item / NS; };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmC);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < VQ_STAGES; ++i) {
        ptx::mbar_init(&full[i], 1);
        ptx::mbar_init(&empty[i], CL);
      }
      ptx::mbar_init(tfull, 1);
      ptx::mbar_init(tempty, 8);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) ptx::cluster_sync();   // the peer's barriers exist before anything is multicast to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {  // one lane, known to the compiler: operands go to uniform registers
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < n_items; item += n_workers) {
        const int mblk = item_mblk(item), ns = item % NS;
        const int t_begin = ns * tiles_per_item;
        const int t_end = min(n_tiles, t_begin + tiles_per_item);
        for (int t = t_begin; t < t_end; ++t) {
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], L::A_BYTES + L::B_BYTES);
            ptx::tma_load_2d(sA + stage * L::A_BYTES, &tmX, &full[stage], kb * VQ_BK, mblk * VQ_BM);
            if constexpr (CL == 2)   // this CTA's half of the codebook tile, into both CTAs
              ptx::tma_load_2d_multicast(sB + stage * L::B_BYTES + rank * (L::B_BYTES / 2), &tmC, &full[stage],
                                         (kb % kblocks_c) * VQ_BK, t * VQ_BN + rank * (VQ_BN / 2), (uint16_t)3);
            else
              ptx::tma_load_2d(sB + stage * L::B_BYTES, &tmC, &full[stage], (kb % kblocks_c) * VQ_BK, t * VQ_BN);
            if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {  // one lane, known to the compiler: operands go to uniform registers
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int item = worker; item < n_items; item += n_workers) {
        const int ns = item % NS;
        const int t_begin = ns * tiles_per_item;
        const int t_end = min(n_tiles, t_begin + tiles_per_item);
        for (int t = t_begin; t < t_end; ++t) {
          ptx::mbar_wait(tempty, tphase ^ 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(sA + stage * L::A_BYTES);
            const uint32_t b_addr = ptx::smem_u32(sB + stage * L::B_BYTES);
#pragma unroll
            for (int k = 0; k < VQ_BK / 16; ++k) {
              const uint64_t db = ptx::make_smem_desc<128>(b_addr + k * 32);
              const uint64_t da0 = ptx::make_smem_desc<128>(a_addr + k * 32);
              const uint64_t da1 = ptx::make_smem_desc<128>(a_addr + 128 * 128 + k * 32);
              const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
              ptx::mma_bf16_ss(tmem_base, da0, db, IDESC, acc);
              ptx::mma_bf16_ss(tmem_base + VQ_BN, da1, db, IDESC, acc);
            }
            if constexpr (CL == 2) ptx::mma_commit_multicast(&empty[stage], (uint16_t)3);
            else ptx::mma_commit(&empty[stage]);
            if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(tfull);
          tphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: running min + windowed candidates
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + half * VQ_BN;
    uint32_t tphase = 0;
    for (int item = worker; item < n_items; item += n_workers) {
      const int mblk = item_mblk(item), ns = item % NS;
      const int t_begin = ns * tiles_per_item;
      const int t_end = min(n_tiles, t_begin + tiles_per_item);
      const int row = mblk * VQ_BM + half * 128 + q * 32 + lane;
      const bool valid = row < nrows;
      const float w = valid ? win[row] : 0.f;
      int2* list = cand_out + ((size_t)(valid ? row : 0) * VQ_NS_MAX + ns) * VQ_CAP;
      float best = INFINITY;
      int cnt = 0;
      for (int t = t_begin; t < t_end; ++t) {
        ptx::mbar_wait_sleepy(tfull, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        const int n0 = t * VQ_BN;
#pragma unroll 1
        for (int c = 0; c < VQ_BN / 32; ++c) {
          uint32_t acc[32];
          ptx::tmem_ld_32x32(t_lane + c * 32, acc);
          float s[32];
          const float4* cp = reinterpret_cast<const float4*>(c2 + n0 + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = __ldg(cp + i);
            s[4 * i] = v.x; s[4 * i + 1] = v.y; s[4 * i + 2] = v.z; s[4 * i + 3] = v.w;
          }
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = fmaf(__uint_as_float(acc[i]), -2.f, s[i]);
          float m = s[0];
#pragma unroll
          for (int i = 1; i < 32; ++i) m = fminf(m, s[i]);
          best = fminf(best, m);
          const float thr = best + w;
          if (valid && m <= thr) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (s[i] <= thr) vq_append(list, cnt, thr, n0 + c * 32 + i, s[i]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty);
      }
      if (valid) {
        best_out[(size_t)row * VQ_NS_MAX + ns] = best;
        cnt_out[(size_t)row * VQ_NS_MAX + ns] = cnt;
      }
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

// CUDA-core scorer with the same outputs as vq_score_kernel (fp32 FMA over bf16-rounded operands).  Used for
// codebooks / dims the tensor-core tiling does not cover (tests with tiny codebooks) and as the kernel-level
// cross-check of the tcgen05 path.  One block per 8 rows; not a performance path.
__global__ void __launch_bounds__(256) vq_score_simt_kernel(const __nv_bfloat16* __restrict__ xb,
                                                            const __nv_bfloat16* __restrict__ cb,
                                                            const float* __restrict__ c2,
                                                            const float* __restrict__ win,
                                                            float* __restrict__ best_out, int* __restrict__ cnt_out,
                                                            int2* __restrict__ cand_out, int nrows, int K, int D,
                                                            int split /*rows are [hi | lo] bf16 splits of fp32 rows*/) {
  extern __shared__ float sx[];  // D floats: the row
  __shared__ float s_scores[256];
  const int row = blockIdx.x;
  const size_t pitch = (size_t)D * (split ? 2 : 1);
  for (int i = threadIdx.x; i < D; i += 256)
    sx[i] = __bfloat162float(xb[row * pitch + i]) + (split ? __bfloat162float(xb[row * pitch + D + i]) : 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float best = INFINITY;
  int cnt = 0;
  const float w = win[row];
  int2* list = cand_out + (size_t)row * VQ_NS_MAX * VQ_CAP;
  for (int j0 = 0; j0 < K; j0 += 256) {
    for (int jj = warp; jj < 256; jj += 8) {
      const int j = j0 + jj;
      float acc = 0.f;
      if (j < K) {
        const __nv_bfloat16* cr = cb + (size_t)j * D;
        for (int i = lane; i < D; i += 32) acc = fmaf(sx[i], __bfloat162float(cr[i]), acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) s_scores[jj] = j < K ? fmaf(acc, -2.f, c2[j]) : INFINITY;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = INFINITY;
      for (int jj = 0; jj < 256; ++jj) m = fminf(m, s_scores[jj]);
      best = fminf(best, m);
      const float thr = best + w;
      if (m <= thr)
        for (int jj = 0; jj < 256; ++jj)
          if (s_scores[jj] <= thr) vq_append(list, cnt, thr, j0 + jj, s_scores[jj]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int ns = 0; ns < VQ_NS_MAX; ++ns) {
      best_out[(size_t)row * VQ_NS_MAX + ns] = ns == 0 ? best : INFINITY;
      cnt_out[(size_t)row * VQ_NS_MAX + ns] = ns == 0 ? cnt : 0;
    }
  }
}

// ---------------------------------------------------------------- exact fp32 distance of the reference
// d = sqrt(max((x2 + c2_j) + (-2 * xy), 0)) with every operation individually rounded (no FMA contraction).
__device__ __forceinline__ float ref_distance(float x2, float c2j, float xy32) {
  const float t1 = __fadd_rn(x2, c2j);
  const float t2 = __fadd_rn(t1, __fmul_rn(xy32, -2.f));
  return fabsf(__fsqrt_rn(fmaxf(t2, 0.f)));  // fabs: fmaxf(-0, 0) may keep the sign, and the key order needs +0
}

template <bool X_BF16>
__device__ __forceinline__ double warp_dot_f64(const void* __restrict__ x, int64_t row, const float* __restrict__ crow,
                                               int D, int lane) {
  double s = 0.0;
  if constexpr (X_BF16) {
    const uint4* xp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)row * D);
    const float4* cp = reinterpret_cast<const float4*>(crow);
    for (int i = lane; i < D / 8; i += 32) {
      const uint4 v = __ldg(xp + i);
      const float4 c0 = __ldg(cp + 2 * i), c1 = __ldg(cp + 2 * i + 1);
      s = fma((double)__uint_as_float(v.x << 16), (double)c0.x, s);
      s = fma((double)__uint_as_float(v.x & 0xffff0000u), (double)c0.y, s);
      s = fma((double)__uint_as_float(v.y << 16), (double)c0.z, s);
      s = fma((double)__uint_as_float(v.y & 0xffff0000u), (double)c0.w, s);
      s = fma((double)__uint_as_float(v.z << 16), (double)c1.x, s);
      s = fma((double)__uint_as_float(v.z & 0xffff0000u), (double)c1.y, s);
      s = fma((double)__uint_as_float(v.w << 16), (double)c1.z, s);
      s = fma((double)__uint_as_float(v.w & 0xffff0000u), (double)c1.w, s);
    }
  } else {
    const float4* xp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + (size_t)row * D);
    const float4* cp = reinterpret_cast<const float4*>(crow);
    for (int i = lane; i < D / 4; i += 32) {
      const float4 v = __ldg(xp + i), c = __ldg(cp + i);
      s = fma((double)v.x, (double)c.x, s);
      s = fma((double)v.y, (double)c.y, s);
      s = fma((double)v.z, (double)c.z, s);
      s = fma((double)v.w, (double)c.w, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// ---------------------------------------------------------------- pass 2
template <bool X_BF16>
__global__ void __launch_bounds__(256) vq_rescore_kernel(const void* __restrict__ x, const float* __restrict__ x2p,
                                                         const float* __restrict__ codebook,
                                                         const float* __restrict__ c2, const float* __restrict__ win,
                                                         const float* __restrict__ best_in,
                                                         const int* __restrict__ cnt_in,
                                                         const int2* __restrict__ cand, int64_t rows, int D, int NS,
                                                         int64_t* __restrict__ codes, int* __restrict__ ovf_rows,
                                                         int* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float b = INFINITY;
  int c = 0;
  if (lane < NS) {
    b = best_in[(size_t)row * VQ_NS_MAX + lane];
    c = cnt_in[(size_t)row * VQ_NS_MAX + lane];
  }
  const bool ovf = __any_sync(0xffffffffu, c > VQ_CAP);
  if (ovf) {
    if (lane == 0) ovf_rows[atomicAdd(&counters[0], 1)] = (int)row;
    return;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
  const float thr = b + win[row];
  const float x2 = x2p[row];
  const int2* list = cand + (size_t)row * VQ_NS_MAX * VQ_CAP;

  // pass A: count survivors, remember the (lowest-index) single one
  int total = 0, only = 0x7fffffff;
  for (int e0 = 0; e0 < NS * VQ_CAP; e0 += 32) {
    const int e = e0 + lane, ns = e / VQ_CAP, k = e % VQ_CAP;
    const int cn = __shfl_sync(0xffffffffu, c, ns);
    bool keep = false;
    int2 ent = make_int2(0, 0);
    if (k < cn) {
      ent = list[(size_t)ns * VQ_CAP + k];
      keep = __int_as_float(ent.y) <= thr;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    total += __popc(m);
    int idx = keep ? ent.x : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
    only = min(only, idx);
  }
  if (total <= 1) {
    if (lane == 0) {
      codes[row] = total == 1 ? only : 0;  // no survivor only for non-finite rows: torch's argmax of all-NaN is 0
      atomicAdd(&counters[2], 1);
    }
    return;
  }
  // pass B: exact re-score of every survivor
  float bd = INFINITY;
  int bi = 0x7fffffff;
  for (int e0 = 0; e0 < NS * VQ_CAP; e0 += 32) {
    const int e = e0 + lane, ns = e / VQ_CAP, k = e % VQ_CAP;
    const int cn = __shfl_sync(0xffffffffu, c, ns);
    bool keep = false;
    int2 ent = make_int2(0, 0);
    if (k < cn) {
      ent = list[(size_t)ns * VQ_CAP + k];
      keep = __int_as_float(ent.y) <= thr;
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int j = __shfl_sync(0xffffffffu, ent.x, src);
      const double xy = warp_dot_f64<X_BF16>(x, row, codebook + (size_t)j * D, D, lane);
      const float d = ref_distance(x2, __ldg(c2 + j), (float)xy);
      if (d < bd || (d == bd && j < bi)) {
        bd = d;
        bi = j;
      }
    }
  }
  if (lane == 0) {
    codes[row] = bi;
    atomicAdd(&counters[1], total);
  }
}

// ---------------------------------------------------------------- pass 3 (overflowed rows only)
__global__ void vq_exhaustive_init_kernel(unsigned long long* keys, const int* counters) {
  const int n = counters[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) keys[i] = ~0ull;
}
template <bool X_BF16>
__global__ void __launch_bounds__(256) vq_exhaustive_kernel(const void* __restrict__ x, const float* __restrict__ x2p,
                                                            const float* __restrict__ codebook,
                                                            const float* __restrict__ c2, int K, int D,
                                                            const int* __restrict__ ovf_rows,
                                                            unsigned long long* __restrict__ keys,
                                                            const int* __restrict__ counters) {
  const int n = counters[0];
  if (n == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (K + gridDim.x - 1) / gridDim.x;
  const int j_begin = blockIdx.x * chunk, j_end = min(K, j_begin + chunk);
  for (int r = 0; r < n; ++r) {
    const int64_t row = ovf_rows[r];
    const float x2 = x2p[row];
    unsigned long long bestk = ~0ull;
    for (int j = j_begin + warp; j < j_end; j += 8) {
      const double xy = warp_dot_f64<X_BF16>(x, row, codebook + (size_t)j * D, D, lane);
      const float d = ref_distance(x2, __ldg(c2 + j), (float)xy);
      // d >= 0 (or NaN): its bit pattern orders like the value; NaN (0x7fc00000) sorts last
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
      bestk = key < bestk ? key : bestk;
    }
    if (lane == 0 && bestk != ~0ull) atomicMin(&keys[r], bestk);
  }
}
__global__ void vq_exhaustive_write_kernel(const int* __restrict__ ovf_rows, const unsigned long long* __restrict__ keys,
                                           const int* __restrict__ counters, int64_t* __restrict__ codes) {
  const int n = counters[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    // all-NaN row: every key carries NaN bits; torch.argmax returns the first NaN -> index 0
    const unsigned db = (unsigned)(k >> 32);
    codes[ovf_rows[i]] = (db & 0x7fffffffu) > 0x7f800000u ? 0 : (int64_t)(k & 0xffffffffu);
  }
}

// ---------------------------------------------------------------- launcher
int launch_vq_search(const void* x, int x_dt, const float* x2_opt, int64_t nrows, int D, const float* codebook_f32,
                     const __nv_bfloat16* codebook_bf16, const float* c2, const float* c2max_dev, int K,
                     int64_t* codes, void* ws, size_t ws_bytes, float window_factor, bool use_tc, bool x2_exact,
                     cudaStream_t st, int sm_count, int* stats_host_opt, int cta_pairs) {
  if (nrows == 0) return DC_OK;
  DC_CHECK(nrows < (1ll << 31) - VQ_BM, DC_ERR_SHAPE, "vq_search: too many rows (%lld)", (long long)nrows);
  DC_CHECK(D % 64 == 0, DC_ERR_SHAPE, "vq_search: D=%d must be a multiple of 64", D);
  const bool xb16 = x_dt == DT_BF16;
  const size_t need = vq_workspace_bytes(nrows, D, xb16);
  DC_CHECK(ws != nullptr && ws_bytes >= need, DC_ERR_WORKSPACE, "vq_search: workspace %zu < %zu bytes", ws_bytes, need);
  DC_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, DC_ERR_ARG, "vq_search: workspace must be 256-byte aligned");
  VqWs w = vq_carve(ws, nrows, D, !xb16);
  const __nv_bfloat16* xb = xb16 ? reinterpret_cast<const __nv_bfloat16*>(x) : w.xb;
  const unsigned row_blocks = (unsigned)((nrows + 7) / 8);

  {
  ProfScope ps(PC_VQ_PREP, 0, (double)nrows * D * (xb16 ? 2.0 : 6.0), st);
  if (xb16)
    vq_prep_kernel<true><<<row_blocks, 256, 0, st>>>(x, w.xb, w.x2e, w.win, c2max_dev, nrows, D, window_factor,
                                                     x2_exact ? 1 : 0, w.counters);
  else
    vq_prep_kernel<false><<<row_blocks, 256, 0, st>>>(x, w.xb, w.x2e, w.win, c2max_dev, nrows, D, window_factor,
                                                      x2_exact ? 1 : 0, w.counters);
  }
  ++g_launches_vq;
  DC_CUDA(cudaGetLastError());

  int NS = 1;
  {
  // algorithmic work of the distance contraction: 2 * N * K * D flops; operands once: x (bf16) + codebook (bf16)
  ProfScope ps(PC_VQ_SCORE, 2.0 * (double)nrows * K * D, ((double)nrows + K) * D * 2.0, st);
  if (use_tc && K % VQ_BN == 0) {
    const int n_mblk = (int)((nrows + VQ_BM - 1) / VQ_BM);
    const int n_tiles = K / VQ_BN;
    // Work item = (256-row block, codebook split); item i -> (block i / NS, split i % NS), so the sm_count items in
    // flight cover sm_count / NS row blocks x NS splits.  Every CTA re-reads its 256 x D x-panel (1.8 MB at D = 3584)
    // from L2 once per codebook tile, and the CTAs of one split stream the same codebook tiles in step, so the set
    // of x-panels in flight must stay L2-resident: (sm_count / NS) * panel <= ~64 MB  ->  NS >= 4 on B200.
    // (With NS = 2 the panels thrash: ncu showed 152 GB of DRAM reads per launch against 1.95 GB algorithmic.)
    const double panel_bytes = (double)VQ_BM * D * 2.0;
    auto panels_in_flight = [&](int ns) {
      const int c = (sm_count + ns - 1) / ns;
      return (double)(n_mblk < c ? n_mblk : c) * panel_bytes;
    };
    int ns_min = 1;
    while (ns_min < VQ_NS_MAX && ns_min * 2 <= n_tiles && panels_in_flight(ns_min) > 64e6) ns_min *= 2;
    // among the admissible split counts pick the one that fills the persistent grid most evenly
    double best_eff = 0.0;
    for (int ns = ns_min; ns <= VQ_NS_MAX && ns <= n_tiles; ns *= 2) {
      const long long items = (long long)n_mblk * ns;
      const long long rounds = (items + sm_count - 1) / sm_count;
      const double eff = (double)items / (double)(rounds * sm_count);
      if (eff > best_eff + 0.02) {
        best_eff = eff;
        NS = ns;
      }
    }
    const int tiles_per_item = (n_tiles + NS - 1) / NS;
    const bool pair = cta_pairs == 2 && n_mblk >= 2 && sm_count >= 2;
    const int CLv = pair ? 2 : 1;
    const int n_items = ((n_mblk + CLv - 1) / CLv) * NS;   // pair items: two row blocks x one codebook split
    CUtensorMap tmX, tmC;
    {
      const uint64_t Dx = (uint64_t)D * (xb16 ? 1 : 2);   // fp32 rows were split into [hi | lo]
      const uint64_t dims[2] = {Dx, (uint64_t)nrows};
      const uint64_t strides[1] = {Dx * 2};
      const uint32_t box[2] = {VQ_BK, VQ_BM};
      DC_TRY(make_tmap_bf16(&tmX, xb, 2, dims, strides, box, 128));
    }
    {
      const uint64_t dims[2] = {(uint64_t)D, (uint64_t)K};
      const uint64_t strides[1] = {(uint64_t)D * 2};
      const uint32_t box[2] = {VQ_BK, (uint32_t)(VQ_BN / CLv)};   // pairs: each CTA loads half the codes of a tile
      DC_TRY(make_tmap_bf16(&tmC, codebook_bf16, 2, dims, strides, box, 128));
    }
    static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
    int dev = 0;
    DC_CUDA(cudaGetDevice(&dev));
    if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
      DC_CUDA(cudaFuncSetAttribute(vq_score_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, VqSmem::TOTAL));
      DC_CUDA(cudaFuncSetAttribute(vq_score_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, VqSmem::TOTAL));
      attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
    }
    const int workers = n_items < sm_count / CLv ? n_items : sm_count / CLv;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(workers * CLv));
    cfg.blockDim = dim3(VQ_THREADS);
    cfg.dynamicSmemBytes = VqSmem::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLv;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 1 : 0;
    const int kblocks_c = D / VQ_BK, kblocks = kblocks_c * (xb16 ? 1 : 2), nr = (int)nrows;
    if (pair)
      DC_CUDA(cudaLaunchKernelEx(&cfg, vq_score_kernel<2>, tmX, tmC, c2, (const float*)w.win, w.best, w.cnt, w.cand, nr, n_items,
                                 NS, n_tiles, tiles_per_item, kblocks, kblocks_c));
    else
      DC_CUDA(cudaLaunchKernelEx(&cfg, vq_score_kernel<1>, tmX, tmC, c2, (const float*)w.win, w.best, w.cnt, w.cand, nr, n_items,
                                 NS, n_tiles, tiles_per_item, kblocks, kblocks_c));
  } else {
    DC_CHECK((size_t)D * 4 <= 48 * 1024, DC_ERR_SHAPE, "vq_search: D=%d too large for the CUDA-core scorer", D);
    vq_score_simt_kernel<<<(unsigned)nrows, 256, (size_t)D * 4, st>>>(xb, codebook_bf16, c2, w.win, w.best, w.cnt,
                                                                       w.cand, (int)nrows, K, D, xb16 ? 0 : 1);
  }
  }
  ++g_launches_vq;
  DC_CUDA(cudaGetLastError());

  const float* x2 = x2_opt ? x2_opt : w.x2e;
  {
  ProfScope ps(PC_VQ_RESCORE, 0, 0, st);
  if (xb16)
    vq_rescore_kernel<true><<<row_blocks, 256, 0, st>>>(x, x2, codebook_f32, c2, w.win, w.best, w.cnt, w.cand, nrows, D,
                                                         NS, codes, w.ovf_rows, w.counters);
  else
    vq_rescore_kernel<false><<<row_blocks, 256, 0, st>>>(x, x2, codebook_f32, c2, w.win, w.best, w.cnt, w.cand, nrows,
                                                          D, NS, codes, w.ovf_rows, w.counters);
  }
  ++g_launches_vq;
  DC_CUDA(cudaGetLastError());

  ProfScope ps_ex(PC_VQ_EXHAUSTIVE, 0, 0, st);
  vq_exhaustive_init_kernel<<<32, 256, 0, st>>>(w.ovf_keys, w.counters);
  const int ex_grid = K >= 128 * 8 ? 128 : (K + 7) / 8;
  if (xb16)
    vq_exhaustive_kernel<true><<<ex_grid, 256, 0, st>>>(x, x2, codebook_f32, c2, K, D, w.ovf_rows, w.ovf_keys, w.counters);
  else
    vq_exhaustive_kernel<false><<<ex_grid, 256, 0, st>>>(x, x2, codebook_f32, c2, K, D, w.ovf_rows, w.ovf_keys, w.counters);
  vq_exhaustive_write_kernel<<<32, 256, 0, st>>>(w.ovf_rows, w.ovf_keys, w.counters, codes);
  g_launches_vq += 3;
  DC_CUDA(cudaGetLastError());

  if (stats_host_opt) {
    int c[4];
    DC_CUDA(cudaMemcpyAsync(c, w.counters, sizeof(c), cudaMemcpyDeviceToHost, st));
    DC_CUDA(cudaStreamSynchronize(st));
    stats_host_opt[0] = (int)nrows;
    stats_host_opt[1] = c[1];
    stats_host_opt[2] = c[0];
    stats_host_opt[3] = c[2];
  }
  return DC_OK;
}

}  // namespace dc
