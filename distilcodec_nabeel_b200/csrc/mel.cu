// Log-mel front-end on the GPU: replaces the reference's CPU `spec_transform`
//   LinearSpectrogram.forward  models/mel_spec.py:26-57   reflect pad (384, 384) -> torch.stft(n_fft 1024, hop 256,
//                                                         periodic hann, center=False) on the CPU (:39) -> sqrt(re^2+im^2+1e-6)
//   LogMelSpectrogram.forward  models/mel_spec.py:109-122 spec^T @ fb (513 x 128 slaney filterbank) -> log(clamp(., 1e-5))
// fused into one kernel: audio (B, Ls) fp32 -> log-mel (B, 128, T) fp32, T = (Ls - 256) / 256 + 1.
//
// One CTA = 8 consecutive frames of one clip, 256 threads.  Per frame: windowed samples (reflect indexing) -> a
// 1024-point radix-4 Stockham FFT in shared memory (5 passes, one radix-4 butterfly per thread per pass, twiddles from
// a table computed in double precision on the host) -> 513 magnitudes.  Then every thread m < 128 accumulates its mel
// bin for the 8 frames, so each filterbank row is read once per 8 frames.  HBM-trivial (1 KB in, 0.5 KB out per frame).
#include "common.cuh"

#include <math.h>

#include <vector>

namespace dc {

static thread_local uint64_t g_launches_mel = 0;
uint64_t mel_launch_count() { return g_launches_mel; }

constexpr int MEL_NFFT = 1024, MEL_HOP = 256, MEL_BINS = 513, MEL_N = 128, MEL_FPB = 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(256) mel_kernel(const float* __restrict__ audio, const float* __restrict__ window,
                                                  const float2* __restrict__ twiddle /*[1024]: exp(-2 pi i m / 1024)*/,
                                                  const float* __restrict__ fb /*[513][128]*/, float* __restrict__ mel,
                                                  int Ls, int T) {
  __shared__ float2 buf[2][MEL_NFFT];
  __shared__ float2 tw[MEL_NFFT];
  __shared__ float mag[MEL_FPB][MEL_BINS + 3];
  const int tid = threadIdx.x, b = blockIdx.y, t_first = blockIdx.x * MEL_FPB;
  const float* a = audio + (size_t)b * Ls;
  for (int i = tid; i < MEL_NFFT; i += 256) tw[i] = twiddle[i];
  const int pad = (MEL_NFFT - MEL_HOP) / 2;  // 384 on both sides (reflect, edge sample not repeated)

  for (int f = 0; f < MEL_FPB; ++f) {
    const int t = t_first + f;
    if (t >= T) {  // block-uniform
      for (int i = tid; i < MEL_BINS; i += 256) mag[f][i] = 0.f;
      continue;
    }
    __syncthreads();  // buf reuse across frames; tw visible
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = tid + r * 256;
      int i = t * MEL_HOP + k - pad;
      if (i < 0) i = -i;
      if (i >= Ls) i = 2 * (Ls - 1) - i;
      buf[0][k] = make_float2(__ldg(a + i) * __ldg(window + k), 0.f);
    }
    __syncthreads();
    int src = 0;
#pragma unroll
    for (int p = 1; p < MEL_NFFT; p *= 4) {  // Stockham autosort, radix 4, forward transform
      const int k = tid & (p - 1);
      const int j = ((tid - k) << 2) + k;
      const int step = 256 / p;              // twiddle index of exp(-2 pi i k / (4p)) in the 1024-entry table
      const float2 u0 = buf[src][tid];
      const float2 u1 = cmul(buf[src][tid + 256], tw[(k * step) & 1023]);
      const float2 u2 = cmul(buf[src][tid + 512], tw[(2 * k * step) & 1023]);
      const float2 u3 = cmul(buf[src][tid + 768], tw[(3 * k * step) & 1023]);
      const float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y), v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
      const float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
      const float2 d = make_float2(u1.x - u3.x, u1.y - u3.y);
      const float2 v3 = make_float2(d.y, -d.x);  // (-i) * (u1 - u3)
      float2* o = buf[src ^ 1];
      o[j] = make_float2(v0.x + v2.x, v0.y + v2.y);
      o[j + p] = make_float2(v1.x + v3.x, v1.y + v3.y);
      o[j + 2 * p] = make_float2(v0.x - v2.x, v0.y - v2.y);
      o[j + 3 * p] = make_float2(v1.x - v3.x, v1.y - v3.y);
      src ^= 1;
      __syncthreads();
    }
    for (int i = tid; i < MEL_BINS; i += 256) {
      const float2 x = buf[src][i];
      mag[f][i] = sqrtf(x.x * x.x + x.y * x.y + 1e-6f);  // models/mel_spec.py:55
    }
  }
  __syncthreads();
  if (tid < MEL_N) {
    float acc[MEL_FPB];
#pragma unroll
    for (int f = 0; f < MEL_FPB; ++f) acc[f] = 0.f;
    for (int i = 0; i < MEL_BINS; ++i) {
      const float w = __ldg(fb + (size_t)i * MEL_N + tid);
#pragma unroll
      for (int f = 0; f < MEL_FPB; ++f) acc[f] = fmaf(mag[f][i], w, acc[f]);
    }
    float* o = mel + ((size_t)b * MEL_N + tid) * T;
#pragma unroll
    for (int f = 0; f < MEL_FPB; ++f)
      if (t_first + f < T) o[t_first + f] = logf(fmaxf(acc[f], 1e-5f));  // models/mel_spec.py:101
  }
}

int mel_twiddles_host(float* out /*2 * 1024*/) {
  for (int m = 0; m < MEL_NFFT; ++m) {
    const double ang = -2.0 * M_PI * (double)m / (double)MEL_NFFT;
    out[2 * m] = (float)cos(ang);
    out[2 * m + 1] = (float)sin(ang);
  }
  return DC_OK;
}

int launch_mel(const float* audio, const float* window, const float* twiddle, const float* fb, float* mel, int B,
               int Ls, cudaStream_t st) {
  DC_CHECK(Ls >= MEL_NFFT - 2 * ((MEL_NFFT - MEL_HOP) / 2) && Ls > (MEL_NFFT - MEL_HOP) / 2, DC_ERR_SHAPE,
           "mel: clip of %d samples is shorter than the reflect padding", Ls);
  const int T = (Ls - MEL_HOP) / MEL_HOP + 1;
  DC_CHECK(T >= 1, DC_ERR_SHAPE, "mel: clip too short");
  dim3 grid((T + MEL_FPB - 1) / MEL_FPB, B);
  // ~5 N log2 N flops per FFT + the 513 x 128 filterbank product per frame
  ProfScope ps(PC_MEL, (double)B * T * (5.0 * 1024 * 10 + 2.0 * 513 * 128), (double)B * ((double)Ls * 4 + (double)T * 128 * 4), st);
  mel_kernel<<<grid, 256, 0, st>>>(audio, window, reinterpret_cast<const float2*>(twiddle), fb, mel, Ls, T);
  ++g_launches_mel;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

}  // namespace dc
