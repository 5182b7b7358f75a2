// Fused epilogue of the tcgen05 implicit-GEMM kernels (gemm_tc.cu, conv_ws.cu): bias / activation / LayerScale /
// residual / 3-branch mean / dual fp32 + bf16(SiLU) stores, applied to 128 x BN accumulator tiles read from TMEM.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace dc {

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
// two elements per call with packed fp32 arithmetic (FMUL2 / FADD2; bit-identical to the scalar form): the epilogue
// warps are issue-bound, so 3 packed + 4 MUFU instructions per pair instead of 10 matter.
// DC_SILU_NR=1 (A/B, rejected): reciprocal by integer seed + two packed Newton steps instead of MUFU.RCP (relative
// error < 7e-6).  ncu shows the XU pipe at 65 % on the short-K launches run alone, but in the step the extra 6 issue
// slots per pair cost more than the MUFU results they save: 535 -> 550 ms per step, conv_ws_pairs 46.7 -> 58.5 ms,
// no launch faster (profiles/r2_ncu_summary.md).
#ifndef DC_SILU_NR
#define DC_SILU_NR 0
#endif
// DC_SILU_TANH=1 (default): x * sigmoid(x) = h + h * tanh(h), h = x / 2, with MUFU.TANH: ONE MUFU result and 2 packed
// FP32 instructions per pair instead of two MUFU results and 3.  tanh.approx.f32 has a relative error of 2^-11, i.e. an
// absolute error of the result below 2.5e-4 |x| — every caller rounds to bf16 right after (relative 2^-9).  ncu had
// shown the short-K conv1 launches stalled on mio_throttle with the XU pipe at 65 % and the issue slots at 55 %; this
// form relieves both.  A/B on one box, alternating: 494.1 / 492.8 -> 484.6 / 487.4 ms per step; error of the whole
// chain vs the oracle 5.03e-3 -> 4.65e-3 (smoke), every parity test unchanged.
#ifndef DC_SILU_TANH
#define DC_SILU_TANH 1
#endif
__device__ __forceinline__ float2 silu_fast2(float2 x) {
#if DC_SILU_TANH
  const float2 h = fmul2(x, make_float2(0.5f, 0.5f));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  return ffma2(h, t, h);
#else
  float2 t = fmul2(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float2 e;
#if DC_SILU_NR
  t.x = fminf(t.x, 126.f);
  t.y = fminf(t.y, 126.f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const float2 nd = ffma2(e, make_float2(-1.f, -1.f), make_float2(-1.f, -1.f));  // -(1 + e)
  // 0x7EF311C7 - bits(d), with the sign bit of -d folded into the constant
  float2 r = make_float2(__uint_as_float(0xFEF311C7u - __float_as_uint(nd.x)),
                         __uint_as_float(0xFEF311C7u - __float_as_uint(nd.y)));
  const float2 two = make_float2(2.f, 2.f);
  r = fmul2(r, ffma2(nd, r, two));
  r = fmul2(r, ffma2(nd, r, two));
  return fmul2(x, r);
#else
  float2 r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const float2 d = fadd2(e, make_float2(1.f, 1.f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return fmul2(x, r);
#endif
#endif
}
// exact-erf GELU (nn.GELU(), convnext_utils.py:254) in 12 instructions with ONE MUFU op:
//   gelu(x) = max(x, 0) - a * Phi(-a),  a = |x|,  Phi(-a) = 0.5 erfc(a / sqrt 2) = 2^q(a)
// with q a degree-7 minimax fit of log2 Phi(-a) on [0, 6.5] (a is clamped there: a * Phi(-a) < 3e-10 beyond).
// Measured against the double-precision definition: relative error of Phi(-a) < 9e-6 (so the small negative-side
// outputs keep 5 digits), absolute error of gelu < 8e-7 — far below the bf16 rounding of the stored result and the
// fp32-mode tolerance.  The previous Abramowitz-Stegun form needed rcp + ex2 (2 MUFU, 16 instructions): at 16 MUFU
// results per clock and SM the C -> 4C GEMMs of the encoder (K = 768/1024 only) were bound by their own epilogue.
__device__ __forceinline__ float gelu_fast(float x) {
  const float a = fminf(fabsf(x), 6.5f);
  float q = fmaf(-1.757783398e-06f, a, 5.993675039e-05f);
  q = fmaf(q, a, -9.163646306e-04f);
  q = fmaf(q, a, 8.447394132e-03f);
  q = fmaf(q, a, -5.382452560e-02f);
  q = fmaf(q, a, -4.586156732e-01f);
  q = fmaf(q, a, -1.151180187e+00f);
  q = fmaf(q, a, -1.000003870e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return fmaf(x + fabsf(x), 0.5f, -a * e);  // max(x, 0) written so that a NaN input stays NaN
}
// two elements at once with packed fp32 arithmetic (FFMA2): the Horner chain costs 7 issue slots per PAIR
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 a = make_float2(fminf(ax.x, 6.5f), fminf(ax.y, 6.5f));
  auto k2 = [](float c) { return make_float2(c, c); };
  float2 q = ffma2(k2(-1.757783398e-06f), a, k2(5.993675039e-05f));
  q = ffma2(q, a, k2(-9.163646306e-04f));
  q = ffma2(q, a, k2(8.447394132e-03f));
  q = ffma2(q, a, k2(-5.382452560e-02f));
  q = ffma2(q, a, k2(-4.586156732e-01f));
  q = ffma2(q, a, k2(-1.151180187e+00f));
  q = ffma2(q, a, k2(-1.000003870e+00f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  const float2 na = make_float2(-a.x, -a.y);
  return ffma2(fadd2(x, ax), k2(0.5f), fmul2(na, e));
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
// four consecutive values of GEMM row `row`, columns n..n+3, as the two-term bf16 split [hi seg | mid seg] of the logical
// row they belong to: logical rows have `seg` channels, a GEMM row holds ldo / seg of them (ConvTranspose1d phases)
__device__ __forceinline__ void st4_split(void* base, size_t row, int n, int ldo, int seg, const float4 v) {
  const float x[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], mid[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    hi[k] = __float2bfloat16_rn(x[k]);
    mid[k] = __float2bfloat16_rn(x[k] - __bfloat162float(hi[k]));
  }
  if (seg <= 0) seg = ldo;
  const size_t lrow = row * (size_t)(ldo / seg) + (size_t)(n / seg);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(base) + lrow * (size_t)(2 * seg) + (n % seg);
  *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(o + seg) = *reinterpret_cast<const uint2*>(mid);
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
  if (e.out0) {
    if (e.out0_dt == DT_SPLIT) st4_split(e.out0, row, n, e.ldo, e.split_seg, v);
    else st4(e.out0, off, e.out0_dt, v);
  }
  if (e.out1) {
    if (e.out1_silu) {
      v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
    }
    if (e.out1_dt == DT_SPLIT) st4_split(e.out1, row, n, e.ldo, e.split_seg, v);
    else st4(e.out1, off, e.out1_dt, v);
  }
}

// ---------------------------------------------------------------- specialised epilogues
// The runtime-flag epilogue above costs ~30 instructions per element (flag tests, 64-bit address arithmetic, dtype
// dispatch) and the 8 epilogue warps cannot hide that; the variants below are the epilogues the hot path actually
// uses, with every flag a compile-time constant.  The host picks the variant (epilogue_variant()).
enum { EV_GENERIC = 0, EV_SILU_BF16, EV_RES_F32_BF16S, EV_RES_F32, EV_RES_MEAN_BF16S, EV_F32_BF16S, EV_GELU_BF16,
       EV_GAMMA_RES_F32, EV_F32, EV_BF16 };

static inline int epilogue_variant(const Epilogue& e) {
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
template <int V, int CW, bool LINEAR = false /* staging tile written un-swizzled (transposed accumulators) */,
          int ROWS = 32 /* rows in the staging tile */>
__device__ __forceinline__ void epilogue_chunk(const Epilogue& ep, const float* stg, int lane, size_t off0,
                                               int nvalid, const float4 b4, const float4 g4) {
  constexpr int CPR = CW / 4, RPI = 32 / CPR, NG = ROWS / RPI;
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
    const int rswz = LINEAR ? 0 : (CW == 32 ? (r & 7) : ((r >> 1) & 3));
    float4 v = *reinterpret_cast<const float4*>(stg + r * CW + ((cg ^ rswz) << 2));
    if (k < nvalid) {
      const size_t off = off0 + k * stride;
      // packed fp32 arithmetic on the two column pairs (same IEEE results as the scalar forms)
      float2 lo = fadd2(make_float2(v.x, v.y), make_float2(b4.x, b4.y));
      float2 hi = fadd2(make_float2(v.z, v.w), make_float2(b4.z, b4.w));
      if constexpr (V == EV_SILU_BF16) {
        lo = silu_fast2(lo);
        hi = silu_fast2(hi);
      }
      if constexpr (V == EV_GELU_BF16) {
        lo = gelu_fast2(lo);
        hi = gelu_fast2(hi);
      }
      if constexpr (V == EV_GAMMA_RES_F32) {
        lo = fmul2(lo, make_float2(g4.x, g4.y));
        hi = fmul2(hi, make_float2(g4.z, g4.w));
      }
      if constexpr (kRes) {
        lo = fadd2(lo, make_float2(r4[k].x, r4[k].y));
        hi = fadd2(hi, make_float2(r4[k].z, r4[k].w));
      }
      if constexpr (V == EV_RES_MEAN_BF16S) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.add1) + off));
        const float4 c = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.add2) + off));
        const float2 sc = make_float2(ep.scale, ep.scale);
        lo = fmul2(fadd2(fadd2(lo, make_float2(a.x, a.y)), make_float2(c.x, c.y)), sc);
        hi = fmul2(fadd2(fadd2(hi, make_float2(a.z, a.w)), make_float2(c.z, c.w)), sc);
      }
      v = make_float4(lo.x, lo.y, hi.x, hi.y);
      if constexpr (kOut0F) *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + off) = v;
      if constexpr (kOut0B) st_bf16x4(reinterpret_cast<__nv_bfloat16*>(ep.out0) + off, v);
      if constexpr (kOut1) {
        lo = silu_fast2(lo);
        hi = silu_fast2(hi);
        st_bf16x4(reinterpret_cast<__nv_bfloat16*>(ep.out1) + off, make_float4(lo.x, lo.y, hi.x, hi.y));
      }
    }
  }
}

// generic (runtime-flag) chunk
template <int CW, bool LINEAR = false, int ROWS = 32>
__device__ __forceinline__ void epilogue_chunk_generic(const Epilogue& ep, const float* stg, int lane, size_t off0,
                                                       int nvalid, int n, const float4 b4, const float4 g4) {
  constexpr int CPR = CW / 4, RPI = 32 / CPR, NG = ROWS / RPI;
  const int cg = lane % CPR, rsub = lane / CPR;
  const size_t stride = (size_t)RPI * ep.ldo;
  float4 r4[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k)
    r4[k] = (ep.res && k < nvalid) ? ld4(ep.res, off0 + k * stride, ep.res_dt) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int r = k * RPI + rsub;
    const int rswz = LINEAR ? 0 : (CW == 32 ? (r & 7) : ((r >> 1) & 3));
    const float4 v = *reinterpret_cast<const float4*>(stg + r * CW + ((cg ^ rswz) << 2));
    if (k < nvalid) epilogue4(ep, (off0 + k * stride - n) / ep.ldo, n, v, b4, g4, r4[k]);
  }
}

template <int CW, bool LINEAR, int ROWS = 32>
__device__ __forceinline__ void epilogue_dispatch(const Epilogue& ep, int variant, const float* stg, int lane,
                                                  size_t off0, int nvalid, int n, const float4 b4, const float4 g4) {
  switch (variant) {
    case EV_SILU_BF16: epilogue_chunk<EV_SILU_BF16, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_RES_F32_BF16S: epilogue_chunk<EV_RES_F32_BF16S, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_RES_F32: epilogue_chunk<EV_RES_F32, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_RES_MEAN_BF16S: epilogue_chunk<EV_RES_MEAN_BF16S, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_F32_BF16S: epilogue_chunk<EV_F32_BF16S, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_GELU_BF16: epilogue_chunk<EV_GELU_BF16, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_GAMMA_RES_F32: epilogue_chunk<EV_GAMMA_RES_F32, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_F32: epilogue_chunk<EV_F32, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    case EV_BF16: epilogue_chunk<EV_BF16, CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, b4, g4); break;
    default: epilogue_chunk_generic<CW, LINEAR, ROWS>(ep, stg, lane, off0, nvalid, n, b4, g4); break;
  }
}

// L2 prefetch of the fp32 operands an epilogue will read for its next tile (residual; the two other branch results
// for the 3-branch mean), issued while the warp would otherwise just wait for the tile's MMAs: the loads in
// epilogue_chunk then hit L2 (~300 cycles) instead of exposing one DRAM round trip (~1 us) per chunk.
// Region: rows [t_first, t_first + 32) x columns [n_first, n_first + ncols) of the (clip, T, ldo) tensor.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void epilogue_prefetch(const Epilogue& ep, int clip, int T, int t_first, int n_first,
                                                  int ncols, int lane, int t_lim = 0x7fffffff) {
  if (!ep.prefetch || (!ep.res && !ep.add1)) return;
  const int t = t_first + lane;
  if (t >= T || t >= t_lim) return;
  const size_t off = ((size_t)clip * T + t) * (size_t)ep.ldo + n_first;
  for (int c = 0; c < ncols; c += 32) {  // 32 fp32 = one 128-byte line
    if (ep.res && ep.res_dt == DT_F32) prefetch_l2(reinterpret_cast<const float*>(ep.res) + off + c);
    if (ep.add1 && ep.add_dt == DT_F32) {
      prefetch_l2(reinterpret_cast<const float*>(ep.add1) + off + c);
      prefetch_l2(reinterpret_cast<const float*>(ep.add2) + off + c);
    }
  }
}

// One 128 x BN output tile: the 8 epilogue warps (CTA warps 2..9) each take a TMEM lane quarter (rows) and one half of
// the columns.  TMEM -> registers (thread = row) -> XOR-swizzled per-warp smem tile -> (lane = 4 columns) -> global.
// tmem_acc: TMEM address (lane 0) of the tile's accumulator; stg: this warp's 32 x CW fp32 transpose buffer.
// SROWS = 16: the staging tile holds 16 rows only (half the shared memory at the same chunk width): the two lane
// halves of the warp stage and store their rows one after the other.  conv_ws_pair<64> has 2 KB of staging per warp;
// with 16-column chunks every global access of its epilogue touched half a 128-byte line and the L1 data pipe ran at
// 82-86 % of its wavefront peak at 1.75 GHz (profiles/r2_ncu_summary.md), i.e. saturated at the in-step clock.
template <int BN, int CW = (BN >= 64 ? 32 : 16) /* chunk width in columns */, int SROWS = 32>
__device__ __forceinline__ void epilogue_tile(const Epilogue& ep, int variant, float* stg, uint32_t tmem_acc, int clip,
                                              int t0, int n0, int T, int warp, int lane, int t_lim = 0x7fffffff) {
  // t_lim: rows t >= min(T, t_lim) are not stored (tiles whose last accumulator rows are not valid outputs)
  static_assert(SROWS == 32 || SROWS == 16, "staging rows");
  constexpr int CPR = CW / 4;             // 16-byte column groups per row (8 or 4)
  constexpr int RPI = 32 / CPR;           // rows covered by one warp-wide access (4 or 8)
  const int q = warp & 3;                 // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;       // which half of the tile's columns it takes
  const int cg = lane % CPR, rsub = lane / CPR;
#pragma unroll 1
  for (int c = 0; c < (BN / 2) / CW; ++c) {
    const int col0 = half * (BN / 2) + c * CW;  // first column of this chunk within the tile
    // column parameters of this lane's 4 columns (independent of the row)
    const int n = n0 + col0 + cg * 4;
    const float4 b4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g4 = ep.gamma ? __ldg(reinterpret_cast<const float4*>(ep.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
    // thread = row `lane`: 16-byte group i goes to slot (i ^ swz(lane)); conflict-free for both access phases
    const int wswz = CW == 32 ? (lane & 7) : ((lane >> 1) & 3);
    if constexpr (SROWS == 32) {
      uint32_t acc[CW];
      if constexpr (CW == 32) ptx::tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + col0, acc);
      else ptx::tmem_ld_32x16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0, acc);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < CPR; ++i)
        *reinterpret_cast<uint4*>(stg + lane * CW + ((i ^ wswz) << 2)) =
            make_uint4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
      __syncwarp();
      const int tb = t0 + q * 32 + rsub;  // first row of this lane
      const int nvalid = min(32 / RPI, max(0, (min(T, t_lim) - tb + RPI - 1) / RPI));
      const size_t off0 = ((size_t)clip * T + tb) * (size_t)ep.ldo + n;
      epilogue_dispatch<CW, false, 32>(ep, variant, stg, lane, off0, nvalid, n, b4, g4);
      __syncwarp();
    } else {
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        // TMEM loads are warp-wide: both lane halves read their rows in either pass (16 columns at a time, so that
        // no more than 16 accumulator registers are live), the half whose rows this pass stages writes them
#pragma unroll
        for (int hc = 0; hc < CW / 16; ++hc) {
          uint32_t acc[16];
          ptx::tmem_ld_32x16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0 + hc * 16, acc);
          ptx::tmem_ld_wait();
          if ((lane >> 4) == pass) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(stg + (lane & 15) * CW + (((hc * 4 + i) ^ wswz) << 2)) =
                  make_uint4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
          }
        }
        __syncwarp();
        const int tb = t0 + q * 32 + pass * 16 + rsub;
        const int nvalid = min(16 / RPI, max(0, (min(T, t_lim) - tb + RPI - 1) / RPI));
        const size_t off0 = ((size_t)clip * T + tb) * (size_t)ep.ldo + n;
        epilogue_dispatch<CW, false, 16>(ep, variant, stg, lane, off0, nvalid, n, b4, g4);
        __syncwarp();
      }
    }
  }
}
// ---------------------------------------------------------------- direct epilogue for TRANSPOSED accumulators
// conv_ts computes D^T (TMEM lane = output channel, column = time row): after tcgen05.ld a lane holds 32 consecutive
// ROWS of ONE channel, and the 32 lanes of the warp are 32 consecutive channels = the contiguous dimension of the
// channels-last tensors.  A warp-wide 4-byte access per row is therefore one full 128-byte line (fp32) / one 64-byte
// segment (bf16): coalesced with no shared-memory transpose at all.  ncu (prof_r1_ts3) showed the shared-memory data
// pipe 63 % busy on the short-K C = 128 layers with a third of it the epilogue's transpose traffic.
// acc: 32 rows (t_first ..) of this lane's channel; off0: element offset of (t_first, channel); nvalid: rows with t < T.
template <int V>
__device__ __forceinline__ void epilogue_rows_direct(const Epilogue& ep, const uint32_t (&acc)[32], size_t off0, int nvalid,
                                                     float bias) {
  constexpr bool kRes = V == EV_RES_F32_BF16S || V == EV_RES_F32 || V == EV_RES_MEAN_BF16S;
  constexpr bool kOut0F = V == EV_RES_F32_BF16S || V == EV_RES_F32 || V == EV_F32_BF16S || V == EV_F32;
  constexpr bool kOut0B = V == EV_SILU_BF16 || V == EV_BF16;
  constexpr bool kOut1 = V == EV_RES_F32_BF16S || V == EV_RES_MEAN_BF16S || V == EV_F32_BF16S;
  constexpr int G = 8;  // rows per batch: G residual (+ 2G mean operand) loads in flight per lane
  const size_t ld = (size_t)ep.ldo;
  const float2 b2 = make_float2(bias, bias);
#pragma unroll
  for (int g0 = 0; g0 < 32; g0 += G) {
    if (g0 >= nvalid) break;  // warp-uniform
    float r[G], a1[G], a2[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const bool ok = g0 + i < nvalid;
      const size_t off = off0 + (size_t)(g0 + i) * ld;
      if constexpr (kRes) r[i] = ok ? reinterpret_cast<const float*>(ep.res)[off] : 0.f;
      if constexpr (V == EV_RES_MEAN_BF16S) {
        a1[i] = ok ? __ldg(reinterpret_cast<const float*>(ep.add1) + off) : 0.f;
        a2[i] = ok ? __ldg(reinterpret_cast<const float*>(ep.add2) + off) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < G; i += 2) {
      float2 v = fadd2(make_float2(__uint_as_float(acc[g0 + i]), __uint_as_float(acc[g0 + i + 1])), b2);
      if constexpr (V == EV_SILU_BF16) v = silu_fast2(v);
      if constexpr (kRes) v = fadd2(v, make_float2(r[i], r[i + 1]));
      if constexpr (V == EV_RES_MEAN_BF16S)
        v = fmul2(fadd2(fadd2(v, make_float2(a1[i], a1[i + 1])), make_float2(a2[i], a2[i + 1])), make_float2(ep.scale, ep.scale));
      float2 sv = v;
      if constexpr (kOut1) sv = silu_fast2(v);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (g0 + i + h < nvalid) {  // warp-uniform
          const size_t off = off0 + (size_t)(g0 + i + h) * ld;
          const float x = h ? v.y : v.x, sx = h ? sv.y : sv.x;
          if constexpr (kOut0F) reinterpret_cast<float*>(ep.out0)[off] = x;
          if constexpr (kOut0B) reinterpret_cast<__nv_bfloat16*>(ep.out0)[off] = __float2bfloat16_rn(x);
          if constexpr (kOut1) reinterpret_cast<__nv_bfloat16*>(ep.out1)[off] = __float2bfloat16_rn(sx);
        }
      }
    }
  }
}
// A/B on one B200 (C = 128 stage, 256 clips): activation-only epilogues -3..-4 %; the residual variants lose 15-25 %
// (4-byte accesses keep a quarter of the bytes in flight per load instruction), so those stay on the transpose path.
__device__ __forceinline__ bool epilogue_direct_supported(int variant) {
  return variant == EV_SILU_BF16 || variant == EV_BF16;
}


// Transposed accumulator (conv_ts.cu): TMEM lanes = output channels, TMEM columns = time rows of a 256-row tile.
// The 16 epilogue warps split the tile as (4 channel quarters) x (4 row quarters of 64); a warp's chunk is 32
// channels x 32 rows: thread = channel writes its 32 row values as scalar stores stg[row][channel] (consecutive
// lanes -> consecutive words, conflict-free), after which the tile has the same [row][column] form as above.
// tmem_acc: TMEM address (lane 0, first column) of the tile's accumulator; warp16: epilogue warp index 0..15 whose
// (warp16 % 4) equals the hardware warp's TMEM lane quarter.
__device__ __forceinline__ void epilogue_tile_transposed(const Epilogue& ep, int variant, float* stg, uint32_t tmem_acc,
                                                         int clip, int t0, int T, int warp16, int lane) {
  constexpr int CW = 32, RPI = 4;
  const int q = warp16 & 3;       // channel quarter = TMEM lane quarter
  const int rq = warp16 >> 2;     // row quarter: rows [64 rq, 64 rq + 64)
  const int cg = lane & 7, rsub = lane >> 3;
  const int n = q * 32 + cg * 4;
  const float4 b4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 g4 = ep.gamma ? __ldg(reinterpret_cast<const float4*>(ep.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int r0 = rq * 64 + c * 32;  // first time row of this chunk within the tile
    uint32_t acc[32];
    ptx::tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + r0, acc);
    ptx::tmem_ld_wait();
    if (epilogue_direct_supported(variant)) {  // warp-uniform: no shared memory on this path
      const int tfirst = t0 + r0;
      const int nv = min(32, max(0, T - tfirst));
      const int ch = q * 32 + lane;
      const size_t o = ((size_t)clip * T + tfirst) * (size_t)ep.ldo + ch;
      const float bch = ep.bias ? __ldg(ep.bias + ch) : 0.f;
      if (variant == EV_SILU_BF16) epilogue_rows_direct<EV_SILU_BF16>(ep, acc, o, nv, bch);
      else epilogue_rows_direct<EV_BF16>(ep, acc, o, nv, bch);
      continue;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) stg[i * CW + lane] = __uint_as_float(acc[i]);
    __syncwarp();
    const int tb = t0 + r0 + rsub;
    const int nvalid = min(32 / RPI, max(0, (T - tb + RPI - 1) / RPI));
    const size_t off0 = ((size_t)clip * T + tb) * (size_t)ep.ldo + n;
    epilogue_dispatch<CW, true>(ep, variant, stg, lane, off0, nvalid, n, b4, g4);
    __syncwarp();
  }
}

}  // namespace dc
