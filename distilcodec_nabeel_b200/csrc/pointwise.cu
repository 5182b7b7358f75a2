// Bandwidth-bound kernels of the hot path (vectorised, coalesced, one pass over HBM each) and the one-time
// weight prepack kernels.  All activations are channels-last (rows = frames, C contiguous).
#include "common.cuh"
#include "ptx.cuh"

#include <string.h>
#include <algorithm>
#include <type_traits>

namespace dc {

static thread_local uint64_t g_launches_pw = 0;
uint64_t pointwise_launch_count() { return g_launches_pw; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum 8 independent values over the 32 lanes with 9 shuffles instead of 8 x 5: each butterfly step also halves the
// set of values a lane is responsible for.  On return lane L holds the warp-wide total of v[((L >> 2) & 7) bit-
// reversed appropriately]: row index r(L) = ((L >> 4) & 1) * 4 + ((L >> 3) & 1) * 2 + ((L >> 2) & 1).
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? a[i] : a[i + 2], keep = b3 ? a[i + 2] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float c = (b2 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, b2 ? b[0] : b[1], 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}

// ------------------------------------------------------------------------------------------ transposes
// (B, C, T) fp32 -> (B, T, C) fp32|bf16 ; 32x32 tiles through padded shared memory, coalesced both ways.
template <typename TOut>
__global__ void transpose_ncl_to_nlc_kernel(const float* __restrict__ in, TOut* __restrict__ out, int C, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const float* ib = in + (size_t)b * C * T;
  TOut* ob = out + (size_t)b * C * T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? ib[(size_t)c * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      float v = tile[threadIdx.x][i];
      if constexpr (sizeof(TOut) == 4) ob[(size_t)t * C + c] = v;
      else ob[(size_t)t * C + c] = __float2bfloat16_rn(v);
    }
  }
}

int launch_transpose_ncl_to_nlc(const float* in, void* out, int out_dt, int B, int C, int T, cudaStream_t st) {
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  ProfScope ps(PC_TRANSPOSE, 0, (double)B * C * T * (4.0 + (out_dt == DT_F32 ? 4.0 : 2.0)), st);
  if (out_dt == DT_F32) transpose_ncl_to_nlc_kernel<float><<<grid, block, 0, st>>>(in, (float*)out, C, T);
  else transpose_ncl_to_nlc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(in, (__nv_bfloat16*)out, C, T);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}
int launch_transpose_nlc_to_ncl(const float* in, float* out, int B, int T, int C, cudaStream_t st) {
  // (B,T,C) -> (B,C,T) is the same kernel with the roles of C and T swapped
  dim3 grid((C + 31) / 32, (T + 31) / 32, B), block(32, 8);
  ProfScope ps(PC_TRANSPOSE, 0, (double)B * C * T * 8.0, st);
  transpose_ncl_to_nlc_kernel<float><<<grid, block, 0, st>>>(in, out, T, C);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ------------------------------------------------------------------------------------------ dwconv7 + LayerNorm
// One warp per frame.  Lane l owns channels {(i*32 + l)*4 .. +3}.  Depthwise taps come from the 7 neighbouring
// rows (L1/L2 absorb the 7x overlap; DRAM sees each row once).  LayerNorm is two-pass in registers, biased
// variance, eps 1e-6 (F.layer_norm / the manual channels_first form, models/convnext_utils.py:203-213).
// Algorithmic HBM bytes per frame: C*4 read + C*sizeof(TOut) written.
template <int VPL /*float4 groups per lane*/, bool CONV, typename TOut>
__global__ void __launch_bounds__(256) dwconv_ln_kernel(const float* __restrict__ in, const float* __restrict__ dw_w,
                                                        const float* __restrict__ dw_b,
                                                        const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, TOut* __restrict__ out,
                                                        int B, int T) {
  constexpr int C = VPL * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= (long long)B * T) return;
  const int t = (int)(row % T);
  const float* rp = in + (size_t)row * C;

  float4 y[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    if constexpr (CONV) {
      float4 a = __ldg(reinterpret_cast<const float4*>(dw_b + c));
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int tt = t + j - 3;
        if (tt >= 0 && tt < T) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(rp + (long long)(j - 3) * C + c));
          const float4 w = __ldg(reinterpret_cast<const float4*>(dw_w + (size_t)j * C + c));
          a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
        }
      }
      y[i] = a;
    } else {
      y[i] = __ldg(reinterpret_cast<const float4*>(rp + c));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (y[i].x + y[i].y) + (y[i].z + y[i].w);
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = y[i].x - mean, b = y[i].y - mean, c = y[i].z - mean, d = y[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-6f);
  TOut* op = out + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 w = __ldg(reinterpret_cast<const float4*>(ln_w + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(ln_b + c));
    float4 o;
    o.x = (y[i].x - mean) * rstd * w.x + b.x;
    o.y = (y[i].y - mean) * rstd * w.y + b.y;
    o.z = (y[i].z - mean) * rstd * w.z + b.z;
    o.w = (y[i].w - mean) * rstd * w.w + b.w;
    if constexpr (sizeof(TOut) == 4) {
      *reinterpret_cast<float4*>(op + c) = o;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(op + c) = pk;
    }
  }
}

// ---- depthwise k7 + LayerNorm, streaming form (the ConvNeXt blocks' kernel) ------------------------------------------
// Thread i owns channels 4i..4i+3 (its 7 x 4 depthwise weights, bias and LayerNorm affine live in registers) and walks
// down the frames with a 14-row register window, 8 output rows per batch, so every input row is read from HBM once.
// Data movement: the grid is persistent (one CTA per resident slot); each CTA takes a contiguous range of
// (clip, 8-row batch) items.  Clips are back to back in HBM, so the rows it needs form one contiguous range of
// "chunks" (chunk k of a clip = rows 8k-3 .. 8k+4; flat id = clip * (nbT + 1) + k) that the bulk-copy engine
// (cp.async.bulk + mbarrier) streams through an NST-deep shared-memory ring, issued by one thread: no load, address
// or boundary instruction is left in the compute warps' stream, and the ring never drains between clips (no per-run
// start-up bubble, no tail wave).  Rows outside the clip read as zero (= the conv's zero padding) on a slow path that
// only boundary batches take.  LayerNorm statistics are two-pass (mean, then centred sum of squares): an 8-row
// shuffle butterfly per warp, partials through shared memory, lane r of every warp finishes row r and broadcasts.
// Algorithmic HBM bytes per frame: C*4 read + C*sizeof(TOut) written.  Measured (B200, inside the encoder stage at
// the power-capped ~1.3 GHz): 4.3-4.8 TB/s = 0.65-0.73 of the copy peak; alone at 1.9 GHz (ncu) 6.0 TB/s = 0.91.
// History: a warp-per-frame kernel re-reading rows through L1 reached 0.4; register-prefetched global loads 0.60 (a
// third of its issue slots were load addressing and predicates); a single-sync pairwise-merge (Chan) variant needed
// 162 registers and was no faster.
template <int C, typename TOut, int NST, int MINB>
__global__ void __launch_bounds__(C / 4, MINB) dwconv_ln_stream_kernel(const float* __restrict__ in,
                                                                        const float* __restrict__ dw_w /*[7][C]*/,
                                                                        const float* __restrict__ dw_b,
                                                                        const float* __restrict__ ln_w,
                                                                        const float* __restrict__ ln_b,
                                                                        TOut* __restrict__ out, int B, int T) {
  constexpr int R = 8, NW = C / 128;
  constexpr uint32_t ROW_BYTES = C * 4, STAGE_BYTES = R * ROW_BYTES;
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(16) float red[2][R][8];  // [pass][row][warp], padded to 8 warps (unused slots stay 0)
  __shared__ __align__(8) uint64_t full[NST];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbT = (T + R - 1) / R;  // batches per clip; a clip has chunks 0 .. nbT (chunk k = rows 8k-3 .. 8k+4)
  const long long items = (long long)B * nbT;
  const long long i0 = items * blockIdx.x / gridDim.x, i1 = items * (blockIdx.x + 1) / gridDim.x;
  if (i0 >= i1) return;
  int b = (int)(i0 / nbT), m = (int)(i0 - (long long)b * nbT);
  const int b_last = (int)((i1 - 1) / nbT), m_last = (int)((i1 - 1) - (long long)b_last * nbT);
  const int q_last = (int)(((long long)b_last * (nbT + 1) + m_last + 1) - ((long long)b * (nbT + 1) + m));
  for (int i = tid; i < 2 * R * 8; i += C / 4) (&red[0][0][0])[i] = 0.f;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  // producer (thread 0): ring index `issued` <-> chunk (pb, pk) in slot pslot
  int issued = 0, pb = b, pk = m, pslot = 0;
  auto issue_upto = [&](int q_hi) {
    q_hi = min(q_hi, q_last);
    while (issued <= q_hi) {
      const int g = R * pk - 3;
      const int lo = max(0, -g), hi = min(R, T - g);
      const uint32_t bytes = hi > lo ? (uint32_t)(hi - lo) * ROW_BYTES : 0u;
      ptx::mbar_expect_tx(&full[pslot], bytes);
      if (bytes)
        ptx::bulk_load_1d(dsm + (size_t)pslot * STAGE_BYTES + (size_t)lo * ROW_BYTES,
                          in + ((size_t)pb * T + (size_t)(g + lo)) * C, bytes, &full[pslot]);
      ++issued;
      if (++pslot == NST) pslot = 0;
      if (++pk > nbT) {
        pk = 0;
        ++pb;
      }
    }
  };
  if (tid == 0) issue_upto(NST - 1);
  const int c = tid * 4;
  const int my_row = ((tid >> 4) & 1) * 4 + ((tid >> 3) & 1) * 2 + ((tid >> 2) & 1);
  float4 w[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) w[j] = __ldg(reinterpret_cast<const float4*>(dw_w + (size_t)j * C + c));
  const float4 bias = __ldg(reinterpret_cast<const float4*>(dw_b + c));
  const float4 gw = __ldg(reinterpret_cast<const float4*>(ln_w + c));
  const float4 gb = __ldg(reinterpret_cast<const float4*>(ln_b + c));
  const uint8_t* col = dsm + (size_t)c * 4;  // this thread's 4-channel column
  float4 x[14];                              // window: x[i] = row 8m-3+i of the clip
  int q = 0, cs = 0;                         // ring index / slot / phase of the chunk holding rows 8m-3 .. 8m+4
  uint32_t cph = 0;
  bool first = true;

  auto step = [&](const bool fresh) {
    const int g0 = R * m - 3;
    const bool interior = g0 >= 0 && g0 + 14 <= T;  // every window row exists in the clip
    int cs1 = cs + 1;
    uint32_t cph1 = cph;
    if (cs1 == NST) {
      cs1 = 0;
      cph1 ^= 1u;
    }
    const uint8_t* s0 = col + (size_t)cs * STAGE_BYTES;
    const uint8_t* s1 = col + (size_t)cs1 * STAGE_BYTES;
    if (fresh) ptx::mbar_wait(&full[cs], cph);
    if (interior) {
      if (fresh) {
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = *reinterpret_cast<const float4*>(s0 + i * ROW_BYTES);
      }
#pragma unroll
      for (int i = 6; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(s0 + i * ROW_BYTES);
      ptx::mbar_wait(&full[cs1], cph1);
#pragma unroll
      for (int i = 0; i < 6; ++i) x[8 + i] = *reinterpret_cast<const float4*>(s1 + i * ROW_BYTES);
    } else {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fresh) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int t = g0 + i;
          x[i] = (t >= 0 && t < T) ? *reinterpret_cast<const float4*>(s0 + i * ROW_BYTES) : z;
        }
      }
#pragma unroll
      for (int i = 6; i < 8; ++i) {
        const int t = g0 + i;
        x[i] = (t >= 0 && t < T) ? *reinterpret_cast<const float4*>(s0 + i * ROW_BYTES) : z;
      }
      ptx::mbar_wait(&full[cs1], cph1);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int t = g0 + 8 + i;
        x[8 + i] = (t >= 0 && t < T) ? *reinterpret_cast<const float4*>(s1 + i * ROW_BYTES) : z;
      }
    }
    float4 y[R];
    float sv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 lo = make_float2(bias.x, bias.y), hi = make_float2(bias.z, bias.w);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        lo = ffma2(make_float2(x[r + j].x, x[r + j].y), make_float2(w[j].x, w[j].y), lo);
        hi = ffma2(make_float2(x[r + j].z, x[r + j].w), make_float2(w[j].z, w[j].w), hi);
      }
      y[r] = make_float4(lo.x, lo.y, hi.x, hi.y);
      const float2 sm = fadd2(lo, hi);
      sv[r] = sm.x + sm.y;
    }
    {
      const float tot = warp_sum8(sv, lane);
      if ((lane & 3) == 0) red[0][my_row][warp] = tot;
    }
    __syncthreads();
    const bool clip_end = m == nbT - 1;
    // every thread has read chunk q (and, at the end of a clip, what it needs of chunk q + 1): refill their slots
    if (tid == 0) issue_upto((clip_end ? q + 1 : q) + NST);
    {
      const float4 p0 = *reinterpret_cast<const float4*>(&red[0][lane & 7][0]), p1 = *reinterpret_cast<const float4*>(&red[0][lane & 7][4]);
      const float mine = (((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w))) * (-1.f / C);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float nmean = __shfl_sync(0xffffffffu, mine, r);
        const float2 nm = make_float2(nmean, nmean);
        const float2 dl = fadd2(make_float2(y[r].x, y[r].y), nm), dh = fadd2(make_float2(y[r].z, y[r].w), nm);
        y[r] = make_float4(dl.x, dl.y, dh.x, dh.y);  // keep the centred values
        const float2 sq = ffma2(dh, dh, fmul2(dl, dl));
        sv[r] = sq.x + sq.y;
      }
    }
    {
      const float tot = warp_sum8(sv, lane);
      if ((lane & 3) == 0) red[1][my_row][warp] = tot;
    }
    __syncthreads();
    {
      const float4 p0 = *reinterpret_cast<const float4*>(&red[1][lane & 7][0]), p1 = *reinterpret_cast<const float4*>(&red[1][lane & 7][4]);
      const float v = ((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w));
      const float mine = rsqrtf(v * (1.f / C) + 1e-6f);
      TOut* ob = out + ((size_t)b * T + (size_t)R * m) * C + c;
      auto emit = [&](int r, float rstd) {
        const float2 rs = make_float2(rstd, rstd);
        const float2 ol = ffma2(fmul2(make_float2(y[r].x, y[r].y), rs), make_float2(gw.x, gw.y), make_float2(gb.x, gb.y));
        const float2 oh = ffma2(fmul2(make_float2(y[r].z, y[r].w), rs), make_float2(gw.z, gw.w), make_float2(gb.z, gb.w));
        if constexpr (sizeof(TOut) == 4) {
          *reinterpret_cast<float4*>(ob + (size_t)r * C) = make_float4(ol.x, ol.y, oh.x, oh.y);
        } else {
          __nv_bfloat162 lo = __floats2bfloat162_rn(ol.x, ol.y), hi = __floats2bfloat162_rn(oh.x, oh.y);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(ob + (size_t)r * C) = pk;
        }
      };
      if (R * m + R <= T) {
#pragma unroll
        for (int r = 0; r < R; ++r) emit(r, __shfl_sync(0xffffffffu, mine, r));
      } else {
        const int rows_here = T - R * m;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float rstd = __shfl_sync(0xffffffffu, mine, r);
          if (r < rows_here) emit(r, rstd);
        }
      }
    }
    // slide the window by R rows (red[0] of the next batch is written after this batch's second __syncthreads,
    // i.e. after every read of red[0] above; its red[1] writes come after its own first __syncthreads)
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = x[i + R];
    // advance to the next item: at the end of a clip two chunks are consumed (q and the mostly-padding q + 1)
    const int adv = clip_end ? 2 : 1;
    q += adv;
    for (int a = 0; a < adv; ++a) {
      if (++cs == NST) {
        cs = 0;
        cph ^= 1u;
      }
    }
    if (clip_end) {
      m = 0;
      ++b;
    } else {
      ++m;
    }
  };

  for (long long left = i1 - i0; left > 0; --left) {
    const bool fresh = first || m == 0;
    first = false;
    step(fresh);
  }
}

template <int C, int NST, int MINB>
static int dwconv_ln_stream_dispatch(const float* in, const float* dw_w, const float* dw_b, const float* ln_w,
                                      const float* ln_b, void* out, int out_dt, int B, int T, cudaStream_t st) {
  const long long items = (long long)B * ((T + 7) / 8);
  int dev = 0, sms = 0;
  DC_CUDA(cudaGetDevice(&dev));
  DC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)std::min<long long>(items, (long long)sms * MINB);
  const size_t smem = (size_t)NST * 8 * C * 4;
  ProfScope ps(PC_DWCONV_LN, 0, (double)B * T * C * (4.0 + (out_dt == DT_F32 ? 4.0 : 2.0)), st, "C%d", C);
  static std::atomic<unsigned> attr_dev_mask[2];  // the opt-in shared-memory size is a per-device function attribute
  const int which = out_dt == DT_F32 ? 0 : 1;
  if (!(attr_dev_mask[which].load(std::memory_order_acquire) & (1u << dev))) {
    if (which == 0)
      DC_CUDA(cudaFuncSetAttribute(dwconv_ln_stream_kernel<C, float, NST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      DC_CUDA(cudaFuncSetAttribute(dwconv_ln_stream_kernel<C, __nv_bfloat16, NST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev_mask[which].fetch_or(1u << dev, std::memory_order_release);
  }
  if (which == 0)
    dwconv_ln_stream_kernel<C, float, NST, MINB><<<grid, C / 4, smem, st>>>(in, dw_w, dw_b, ln_w, ln_b, (float*)out, B, T);
  else
    dwconv_ln_stream_kernel<C, __nv_bfloat16, NST, MINB><<<grid, C / 4, smem, st>>>(in, dw_w, dw_b, ln_w, ln_b, (__nv_bfloat16*)out, B, T);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

template <int VPL>
static int dwconv_ln_dispatch(const float* in, const float* dw_w, const float* dw_b, const float* ln_w,
                              const float* ln_b, void* out, int out_dt, int B, int T, cudaStream_t st) {
  const long long rows = (long long)B * T;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  constexpr int C = VPL * 128;
  ProfScope ps(dw_w ? PC_DWCONV_LN : PC_LAYERNORM, 0, (double)rows * C * (4.0 + (out_dt == DT_F32 ? 4.0 : 2.0)), st,
               "C%d", C);
  // LayerNorm only (the stem / inter-stage / final norms); depthwise conv + LN goes to dwconv_ln_stream_kernel
  if (out_dt == DT_F32) dwconv_ln_kernel<VPL, false, float><<<grid, 256, 0, st>>>(in, dw_w, dw_b, ln_w, ln_b, (float*)out, B, T);
  else dwconv_ln_kernel<VPL, false, __nv_bfloat16><<<grid, 256, 0, st>>>(in, dw_w, dw_b, ln_w, ln_b, (__nv_bfloat16*)out, B, T);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int launch_dwconv_ln(const float* in, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b,
                     void* out, int out_dt, int B, int T, int C, cudaStream_t st) {
  if (dw_w) {  // depthwise conv + LN: streaming kernel; <C, ring depth, CTAs per SM> (128 registers x C/4 threads)
    switch (C) {
      case 256: return dwconv_ln_stream_dispatch<256, 3, 8>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
      case 512: return dwconv_ln_stream_dispatch<512, 3, 4>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
      case 768: return dwconv_ln_stream_dispatch<768, 3, 2>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
      case 1024: return dwconv_ln_stream_dispatch<1024, 3, 2>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
      default: set_error("dwconv_ln: unsupported channel count %d (256/512/768/1024)", C); return DC_ERR_SHAPE;
    }
  }
  switch (C) {
    case 256: return dwconv_ln_dispatch<2>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
    case 512: return dwconv_ln_dispatch<4>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
    case 768: return dwconv_ln_dispatch<6>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
    case 1024: return dwconv_ln_dispatch<8>(in, dw_w, dw_b, ln_w, ln_b, out, out_dt, B, T, st);
    default: set_error("dwconv_ln: unsupported channel count %d (256/512/768/1024)", C); return DC_ERR_SHAPE;
  }
}

// ------------------------------------------------------------------------------------------ cast
__global__ void cast_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(in + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    out[i] = pk;
  }
}
int launch_cast(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t st) {
  DC_CHECK(n % 4 == 0, DC_ERR_SHAPE, "cast: element count must be a multiple of 4");
  const size_t n4 = n / 4;
  if (n4 == 0) return DC_OK;
  unsigned grid = (unsigned)((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  ProfScope ps(PC_CAST, 0, (double)n * 6.0, st);
  cast_kernel<<<grid, 256, 0, st>>>((const float4*)in, (uint2*)out, n4);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ------------------------------------------------------------------------------------------ codebook gather
// batched_embedding / einx.get_at (vector_quantize_pytorch.py:243-247, residual_vq.py:123): out[r,:] = table[idx[r],:]
// one warp per row, float4 loads; writes the fp32 row (API output quantized_fup) and/or a bf16 copy (GEMM operand).
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table,
                                                          const int64_t* __restrict__ idx, int64_t nrows, int D,
                                                          int64_t table_rows, float* __restrict__ o32,
                                                          __nv_bfloat16* __restrict__ o16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= nrows) return;
  int64_t k = idx[row];
  k = k < 0 ? 0 : (k >= table_rows ? table_rows - 1 : k);  // the reference would raise; never read out of bounds
  const float4* src = reinterpret_cast<const float4*>(table + (size_t)k * D);
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = __ldg(src + i);
    if (o32) reinterpret_cast<float4*>(o32 + (size_t)row * D)[i] = v;
    if (o16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(o16 + (size_t)row * D)[i] = pk;
    }
  }
}
int launch_gather_rows(const float* table, const int64_t* idx, int64_t nrows, int D, int64_t table_rows, float* o32,
                       __nv_bfloat16* o16, cudaStream_t st) {
  if (nrows == 0) return DC_OK;
  // compulsory bytes: a table row is read from HBM at most once per launch (repeats hit L2), outputs written once
  ProfScope ps(PC_GATHER, 0,
               (double)D * (4.0 * (double)(nrows < table_rows ? nrows : table_rows) +
                            (double)nrows * ((o32 ? 4.0 : 0.0) + (o16 ? 2.0 : 0.0))),
               st);
  gather_rows_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(table, idx, nrows, D, table_rows, o32, o16);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ------------------------------------------------------------------------------------------ conv_post + tanh
// Conv1d(32 -> 1, k13, pad 6) + tanh (models/generators.py:141-145) on the already SiLU'd channels-last input.
// 416 FMAs per output sample and a single output channel: CUDA-core work, and bound by instruction issue, not by HBM
// (round 2: 2.33 ms per step at 0.27 of the copy bandwidth with one FMA pair per (tap, 2 samples): the packed operands
// (x[j], x[j+1]) and (w, w) had to be formed with moves, 72 instructions per channel and 4 samples).
// Now the two lanes of a packed FMA are two CHANNELS: the shared-memory tile holds (channel pair) float2 elements, a
// weight pair is one 64-bit constant-bank operand, and a thread keeps 8 consecutive samples x 2 partial sums: per channel
// pair 20 LDS.64 + 104 FFMA2, i.e. 15.5 instructions per channel and 4 samples.  Tile layout [pair][row % 8][row / 8]:
// the fill (lane = row) and the compute reads (lane = 8-row group) are both conflict-free.
struct ConvPostW {
  float w[13 * 32];  // [tap][channel]
};
// SPLIT: the input is the (rows, 2 * 32) bf16 two-term split [hi | mid] of fp32 values (DT_SPLIT): value = hi + mid
template <typename TIn, bool SPLIT = false>
__global__ void __launch_bounds__(128) conv_post_tanh_kernel(const TIn* __restrict__ in, const ConvPostW W, float bias,
                                                            float* __restrict__ out, int L) {
  constexpr int C = 32, K = 13, TILE = 512, ROWS = TILE + 16;   // 12 halo rows + 4 padding
  // R8 = 5 (mod 16) and PL = 1 (mod 4): the fill's 64-bit stores (half-warp = 4 rows x 4 column groups of one
  // coalesced 512-byte global access) land in 16 distinct bank pairs
  constexpr int R8 = 69, PL = 8 * R8 + 1;
  static_assert(R8 >= ROWS / 8 && R8 % 16 == 5 && PL % 4 == 1, "tile pitches");
  constexpr int PITCH = SPLIT ? 2 * C : C;                      // input elements per row
  extern __shared__ float2 sx2[];                               // [C / 2][PL]: pair plane, element (row % 8) * R8 + row / 8
  const int b = blockIdx.y, l0 = blockIdx.x * TILE;
  const TIn* ib = in + (size_t)b * L * PITCH;
  // fill: every global access of a warp is one contiguous run of whole rows (lane = (row, 16-byte group)); with lane =
  // row each 128-byte line was touched by 4 separate instructions and, with 3 x 68 KB of the SM's L1 carved out as
  // shared memory, re-fetched from L2: the kernel was bound by that, not by its FMAs (2.3 ms before and after the
  // instruction count was halved)
  if constexpr (sizeof(TIn) == 2) {
    const int g = threadIdx.x & 3;
#pragma unroll 4
    for (int r = threadIdx.x >> 2; r < ROWS; r += 32) {
      const int l = l0 + r - 6;
      const bool ok = l >= 0 && l < L;
      uint4 v = make_uint4(0u, 0u, 0u, 0u), m = make_uint4(0u, 0u, 0u, 0u);
      if (ok) v = __ldg(reinterpret_cast<const uint4*>(ib + (size_t)l * PITCH) + g);
      if (SPLIT && ok) m = __ldg(reinterpret_cast<const uint4*>(ib + (size_t)l * PITCH + C) + g);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w}, w[4] = {m.x, m.y, m.z, m.w};
      float2* dst = sx2 + (g * 4) * PL + (r & 7) * R8 + (r >> 3);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        dst[k * PL] = make_float2(__uint_as_float(u[k] << 16) + (SPLIT ? __uint_as_float(w[k] << 16) : 0.f),
                                  __uint_as_float(u[k] & 0xffff0000u) + (SPLIT ? __uint_as_float(w[k] & 0xffff0000u) : 0.f));
    }
  } else {
    const int g = threadIdx.x & 7;
#pragma unroll 4
    for (int r = threadIdx.x >> 3; r < ROWS; r += 16) {
      const int l = l0 + r - 6;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (l >= 0 && l < L) v = __ldg(reinterpret_cast<const float4*>(ib + (size_t)l * C) + g);
      float2* dst = sx2 + (g * 2) * PL + (r & 7) * R8 + (r >> 3);
      dst[0] = make_float2(v.x, v.y);
      dst[PL] = make_float2(v.z, v.w);
    }
  }
  __syncthreads();
  // thread (t, h): samples l0 + 8 t .. + 7 over the channel pairs [8 h, 8 h + 8); they need tile rows 8 t .. 8 t + 19 =
  // (row % 8, row / 8) = (k & 7, t + (k >> 3)).  The two halves are added through shared memory (12 warps per SM
  // instead of 6: with 64-thread blocks the LDS -> FFMA2 latency was exposed and the kernel ran at 3.1 ms).
  const int t = threadIdx.x & 63, h = threadIdx.x >> 6;
  float2 acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
  for (int ci = 0; ci < C / 4; ++ci) {
    float2 x[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) x[k] = sx2[(h * (C / 4) + ci) * PL + (k & 7) * R8 + t + (k >> 3)];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      // warp-uniform h: both candidates are constant-bank operands
      const float2 w2 = h ? make_float2(W.w[j * C + C / 2 + 2 * ci], W.w[j * C + C / 2 + 2 * ci + 1])
                          : make_float2(W.w[j * C + 2 * ci], W.w[j * C + 2 * ci + 1]);
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = ffma2(x[o + j], w2, acc[o]);
    }
  }
  __syncthreads();                                   // the tile is dead: reuse its first 2 KB for the partial sums
  float* red = reinterpret_cast<float*>(sx2);        // [8][64]
  if (h == 1) {
#pragma unroll
    for (int o = 0; o < 8; ++o) red[o * 64 + t] = acc[o].x + acc[o].y;
  }
  __syncthreads();
  if (h == 1) return;
  const int l = l0 + t * 8;
  float* ob = out + (size_t)b * L;
  float y[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) y[o] = tanhf((acc[o].x + acc[o].y) + red[o * 64 + t] + bias);
  if (l + 7 < L) {
    *reinterpret_cast<float4*>(ob + l) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(ob + l + 4) = make_float4(y[4], y[5], y[6], y[7]);
  } else {
    for (int o = 0; o < 8; ++o)
      if (l + o < L) ob[l + o] = y[o];
  }
}
int launch_conv_post_tanh(const void* in, int in_dt, const float* w_host /*[13][32], host*/, float bias, float* out,
                          int B, int L, cudaStream_t st) {
  DC_CHECK(L % 4 == 0, DC_ERR_SHAPE, "conv_post: output length must be a multiple of 4");
  ConvPostW W;
  memcpy(W.w, w_host, sizeof(W.w));
  constexpr int SMEM = 16 * (8 * 69 + 1) * 8;  // [C / 2][PL] float2, see the kernel
  static std::atomic<unsigned> attr_dev_mask{0u};  // once per (function, device); atomic because host threads driving different devices meet here
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load(std::memory_order_acquire) & (1u << dev))) {
    DC_CUDA(cudaFuncSetAttribute(conv_post_tanh_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    DC_CUDA(cudaFuncSetAttribute(conv_post_tanh_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    DC_CUDA(cudaFuncSetAttribute(conv_post_tanh_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_dev_mask.fetch_or(1u << dev, std::memory_order_release);
  }
  dim3 grid((L + 511) / 512, B);
  ProfScope ps(PC_CONV_POST, 2.0 * B * (double)L * 32 * 13, (double)B * L * (32.0 * (in_dt == DT_BF16 ? 2 : 4) + 4.0), st);
  if (in_dt == DT_SPLIT)
    conv_post_tanh_kernel<__nv_bfloat16, true><<<grid, 128, SMEM, st>>>((const __nv_bfloat16*)in, W, bias, out, L);
  else if (in_dt == DT_F32) conv_post_tanh_kernel<float><<<grid, 128, SMEM, st>>>((const float*)in, W, bias, out, L);
  else conv_post_tanh_kernel<__nv_bfloat16><<<grid, 128, SMEM, st>>>((const __nv_bfloat16*)in, W, bias, out, L);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ------------------------------------------------------------------------------------------ prepack (one-time)
// weight_norm fold: w[i,:] = g[i] * v[i,:] / ||v[i,:]||_2   (torch._weight_norm, dim = 0)
__global__ void weight_norm_fold_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                        float* __restrict__ w, int inner) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  const float* vi = v + (size_t)i * inner;
  float s = 0.f;
  for (int k = threadIdx.x; k < inner; k += blockDim.x) s = fmaf(vi[k], vi[k], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float scale = g[i] / sqrtf(red[0]);
  for (int k = threadIdx.x; k < inner; k += blockDim.x) w[(size_t)i * inner + k] = vi[k] * scale;
}
int launch_weight_norm_fold(const float* g, const float* v, float* w, int dim0, int inner, cudaStream_t st) {
  ProfScope ps(PC_PREPACK, 0, 0, st);
  weight_norm_fold_kernel<<<dim0, 256, 0, st>>>(g, v, w, inner);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// generic repack of a Conv1d (O,I,k) / ConvTranspose1d (I,O,k) / Linear (O,I) weight into the GEMM operand
//   Wp[n][j*C + c],  n = phase * n_inner + ni,  source = src[ni*s_n + c*s_c + kmap[phase][j]*s_k] (0 if kmap < 0)
__global__ void pack_weight_kernel(const float* __restrict__ src, PackDesc d, float* __restrict__ o_kn,
                                   __nv_bfloat16* __restrict__ o_nk) {
  const int Cx = d.split ? 2 * d.C : d.C;   // packed channels per tap
  const long long K = (long long)d.J * Cx, total = K * d.N;
  const int n_inner = d.N / d.phases;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / K);
    const int k = (int)(i % K);
    const int j = k / Cx, cx = k % Cx, c = cx % d.C;
    const int ph = n / n_inner, ni = n % n_inner;
    const int kk = d.kmap[ph * d.J + j];
    const float v = kk < 0 ? 0.f : src[ni * d.s_n + c * d.s_c + kk * d.s_k];
    if (d.split) {  // two-term bf16 split of the fp32 weight: hi, then what hi leaves
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      o_nk[(size_t)n * K + k] = cx < d.C ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
      continue;
    }
    if (o_kn) o_kn[(size_t)k * d.N + n] = v;
    if (o_nk) o_nk[(size_t)n * K + k] = __float2bfloat16_rn(v);
  }
}
int launch_pack_weight(const float* src, const PackDesc& d, float* o_kn, __nv_bfloat16* o_nk, cudaStream_t st) {
  DC_CHECK(d.phases * d.J <= 8 * 16, DC_ERR_SHAPE, "pack_weight: tap table too large");
  ProfScope ps(PC_PREPACK, 0, 0, st);
  pack_weight_kernel<<<148 * 8, 256, 0, st>>>(src, d, o_kn, o_nk);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// row squared norms, fp64 accumulate, correctly rounded to fp32 (||c||^2 of the codebook; default ||x||^2)
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(in + (size_t)row * D);
  double s = 0.0;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = __ldg(p + i);
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = (float)s;
}
int launch_row_sqnorm(const float* in, float* out, int64_t rows, int D, cudaStream_t st) {
  if (rows == 0) return DC_OK;
  ProfScope ps(PC_PREPACK, 0, 0, st);
  row_sqnorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, out, rows, D);
  ++g_launches_pw;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

}  // namespace dc
