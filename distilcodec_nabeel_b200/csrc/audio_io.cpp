// Host side of SURVEY.md section 8f row f-4: audio file decode + resample into caller-owned (pinned) float buffers,
// multi-threaded over files, replacing the reference's single-threaded librosa path
//   load_and_resample_audio   distilcodec/distil_codec.py:657-684   (librosa.load(sr=None, mono=False) + librosa.resample)
//   load_wav                  distilcodec/models/meldataset.py:18-20 (librosa.load(path, sr=sr): mono mean + resample)
//   save_wav                  distilcodec/distil_codec.py:640-654   (soundfile.write of float32 -> 16-bit PCM WAV)
// so that the GPUs (~4.8 k audio-seconds per second each) are fed from files on disk.  Pure host C++: no CUDA here.
//
// Formats: RIFF/WAVE, PCM 8/16/24/32-bit, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE, any channel count.  (Compressed
// formats — the reference's test.mp3 — need a decoder that is not in this image)
//
// Resampler: polyphase FIR, Kaiser-windowed sinc, the design of scipy.signal.resample_poly (the published algorithm
// the tests pin it against: firwin(2*half+1, 1/max(up,down), window=('kaiser', beta)) * up, half = zeros*max(up,down),
// output length ceil(n*up/down), zero padding at the ends).  librosa's default is soxr_hq, a different (longer) filter:
// the sample values are not bit-equal to the reference's, which is outside the numeric-parity surface (both paths of
// every parity test consume the same array).  zeros = 10, beta = 5.0 are scipy's defaults; zeros = 32,
// beta = 14.77 ("hq") has a stop band below -140 dB like soxr's high-quality setting.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/distilcodec_b200.h"

namespace dc {
void set_error(const char* fmt, ...);
}

namespace {

struct WavFmt {
  int format = 0;  // 1 PCM, 3 IEEE float
  int channels = 0, sample_rate = 0, bits = 0, block_align = 0;
  int64_t data_off = 0, data_bytes = 0;
};

static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// Walks the RIFF chunks of `f`; tolerant of odd-sized chunks (pad byte) and of a data chunk whose size field is 0 or
// larger than the file (streamed writers): the size is then what the file holds.
static int parse_wav(FILE* f, WavFmt* w, const char* path) {
  uint8_t h[12];
  if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
    dc::set_error("%s: not a RIFF/WAVE file", path);
    return DC_ERR_ARG;
  }
  fseek(f, 0, SEEK_END);
  const int64_t file_size = ftell(f);
  int64_t pos = 12;
  bool have_fmt = false;
  while (pos + 8 <= file_size) {
    uint8_t ch[8];
    fseek(f, (long)pos, SEEK_SET);
    if (fread(ch, 1, 8, f) != 8) break;
    const int64_t sz = rd32(ch + 4);
    if (!memcmp(ch, "fmt ", 4)) {
      uint8_t b[40] = {0};
      const size_t n = (size_t)std::min<int64_t>(sz, 40);
      if (sz < 16 || fread(b, 1, n, f) != n) {
        dc::set_error("%s: truncated fmt chunk", path);
        return DC_ERR_ARG;
      }
      w->format = rd16(b);
      w->channels = rd16(b + 2);
      w->sample_rate = (int)rd32(b + 4);
      w->block_align = rd16(b + 12);
      w->bits = rd16(b + 14);
      if (w->format == 0xFFFE && sz >= 26) w->format = rd16(b + 24);  // WAVE_FORMAT_EXTENSIBLE: first 2 bytes of the GUID
      have_fmt = true;
    } else if (!memcmp(ch, "data", 4)) {
      if (!have_fmt) {
        dc::set_error("%s: data chunk before fmt chunk", path);
        return DC_ERR_ARG;
      }
      w->data_off = pos + 8;
      w->data_bytes = (sz == 0 || pos + 8 + sz > file_size) ? file_size - (pos + 8) : sz;
      break;
    }
    pos += 8 + sz + (sz & 1);
  }
  if (!have_fmt || w->data_off == 0) {
    dc::set_error("%s: no fmt / data chunk", path);
    return DC_ERR_ARG;
  }
  const bool ok_fmt = (w->format == 1 && (w->bits == 8 || w->bits == 16 || w->bits == 24 || w->bits == 32)) ||
                      (w->format == 3 && (w->bits == 32 || w->bits == 64));
  if (!ok_fmt || w->channels < 1 || w->sample_rate < 1) {
    dc::set_error("%s: unsupported WAVE encoding (format tag %d, %d bits, %d channels)", path, w->format, w->bits,
                  w->channels);
    return DC_ERR_SHAPE;
  }
  if (w->block_align < w->channels * (w->bits / 8)) w->block_align = w->channels * (w->bits / 8);
  return DC_OK;
}

// one sample -> float in [-1, 1) with the scaling libsndfile (soundfile / librosa's reader) uses
static inline float sample_at(const uint8_t* p, int format, int bits) {
  if (format == 3) {
    if (bits == 32) {
      float v;
      memcpy(&v, p, 4);
      return v;
    }
    double v;
    memcpy(&v, p, 8);
    return (float)v;
  }
  switch (bits) {
    case 8: return ((int)p[0] - 128) * (1.f / 128.f);
    case 16: return (int16_t)rd16(p) * (1.f / 32768.f);
    case 24: {
      int32_t v = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24);
      return (v >> 8) * (1.f / 8388608.f);
    }
    default: return (int32_t)rd32(p) * (1.f / 2147483648.f);
  }
}

// frames [f0, f0 + n) of the data chunk -> mono (mean over channels, what librosa's to_mono does) or channel-major
static int read_frames(FILE* f, const WavFmt& w, int64_t f0, int64_t n, bool mono, float* out, int64_t ch_stride,
                       const char* path) {
  const int bps = w.bits / 8;
  const int64_t CH = 1 << 15;  // frames per read
  std::vector<uint8_t> buf((size_t)CH * w.block_align);
  fseek(f, (long)(w.data_off + f0 * w.block_align), SEEK_SET);
  const float inv = 1.f / (float)w.channels;
  for (int64_t done = 0; done < n;) {
    const int64_t m = std::min(CH, n - done);
    if ((int64_t)fread(buf.data(), (size_t)w.block_align, (size_t)m, f) != m) {
      dc::set_error("%s: short read", path);
      return DC_ERR_ARG;
    }
    if (mono && w.format == 1 && w.bits == 16 && w.block_align == 2 * w.channels) {  // the common case, kept tight
      const int16_t* q = reinterpret_cast<const int16_t*>(buf.data());
      if (w.channels == 1) {
        for (int64_t i = 0; i < m; ++i) out[done + i] = q[i] * (1.f / 32768.f);
      } else if (w.channels == 2) {
        for (int64_t i = 0; i < m; ++i)
          out[done + i] = (q[2 * i] * (1.f / 32768.f) + q[2 * i + 1] * (1.f / 32768.f)) * 0.5f;
      } else {
        for (int64_t i = 0; i < m; ++i) {
          float sacc = 0.f;
          for (int c = 0; c < w.channels; ++c) sacc += q[i * w.channels + c] * (1.f / 32768.f);
          out[done + i] = sacc * inv;
        }
      }
      done += m;
      continue;
    }
    for (int64_t i = 0; i < m; ++i) {
      const uint8_t* p = buf.data() + i * w.block_align;
      if (mono) {
        float s = 0.f;
        for (int c = 0; c < w.channels; ++c) s += sample_at(p + c * bps, w.format, w.bits);
        out[done + i] = w.channels == 1 ? s : s * inv;
      } else {
        for (int c = 0; c < w.channels; ++c) out[c * ch_stride + done + i] = sample_at(p + c * bps, w.format, w.bits);
      }
    }
    done += m;
  }
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------- resampler
static double bessel_i0(double x) {  // power series, converges fast for the beta values used here
  double s = 1.0, t = 1.0;
  const double q = x * x * 0.25;
  for (int k = 1; k < 200; ++k) {
    t *= q / ((double)k * k);
    s += t;
    if (t < 1e-18 * s) break;
  }
  return s;
}

// dot product of n contiguous floats, double accumulation in four independent chains (vectorises to 4-wide FMA where
// the CPU has AVX2; the clones are chosen at load time, so the library still runs on any x86-64)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2,fma", "default")))
#endif
static double dot_f32(const float* a, const float* b, int64_t n) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t i = 0;
  for (; i + 4 <= n; i += 4) {
    s0 += (double)a[i] * (double)b[i];
    s1 += (double)a[i + 1] * (double)b[i + 1];
    s2 += (double)a[i + 2] * (double)b[i + 2];
    s3 += (double)a[i + 3] * (double)b[i + 3];
  }
  for (; i < n; ++i) s0 += (double)a[i] * (double)b[i];
  return (s0 + s1) + (s2 + s3);
}

struct Resampler {
  int up = 1, down = 1, half = 0, taps = 0;
  // polyphase bank: for output m, phase p = (m*down) mod up and i0 = ceil((m*down - half) / up):
  //   y[m] = sum_{k < taps} bank[p][k] * x[i0 + k],   bank[p][k] = hc[m*down - (i0 + k)*up]  (0 outside the filter)
  std::vector<float> bank;
  void design(int sr_in, int sr_out, int zeros, double beta) {
    int a = sr_out, b = sr_in;
    while (b) {
      const int t = a % b;
      a = b;
      b = t;
    }
    up = sr_out / a;
    down = sr_in / a;
    const int mx = std::max(up, down);
    half = zeros * mx;
    const int nt = 2 * half + 1;
    std::vector<double> h((size_t)nt);
    const double fc = 1.0 / mx, i0b = bessel_i0(beta);
    double sum = 0.0;
    for (int n = 0; n < nt; ++n) {
      const double m = n - half;
      const double xs = M_PI * fc * m;
      const double sinc = m == 0 ? 1.0 : sin(xs) / xs;
      const double r = m / half;
      const double win = bessel_i0(beta * sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
      h[n] = fc * sinc * win;
      sum += h[n];
    }
    // unity gain at DC (firwin scale=True), then * up; centred: hc(c) = h[c + half], c in [-half, half]
    taps = 2 * half / up + 2;
    bank.assign((size_t)up * taps, 0.f);
    for (int p = 0; p < up; ++p) {
      // t = q*up + p  =>  i0 = q + ceil((p - half) / up); tap k multiplies x[i0 + k] by hc(t - (i0 + k)*up)
      const int c0 = p - ceil_div(p - half, up) * up;  // offset into the centred filter of tap 0: in (half - up, half]
      for (int k = 0; k < taps; ++k) {
        const int c = c0 - k * up;
        if (c >= -half && c <= half) bank[(size_t)p * taps + k] = (float)(h[c + half] / sum * up);
      }
    }
  }
  static int ceil_div(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }
  int64_t out_len(int64_t n_in) const { return (n_in * up + down - 1) / down; }
  // y[m] = sum_i hc(m*down - i*up) x[i] over |m*down - i*up| <= half  (upfirdn with the filter centred, zero padding)
  void run(const float* x, int64_t n_in, float* y, int64_t m0, int64_t m1) const {
    for (int64_t m = m0; m < m1; ++m) {
      const int64_t t = m * down;
      const int p = (int)(t % up);
      const int64_t i0 = t / up + ceil_div(p - half, up);
      const float* bp = bank.data() + (size_t)p * taps;
      int64_t k0 = 0, k1 = taps;
      if (i0 < 0) k0 = -i0;
      if (i0 + k1 > n_in) k1 = n_in - i0;
      y[m] = k1 > k0 ? (float)dot_f32(bp + k0, x + i0 + k0, k1 - k0) : 0.f;
    }
  }
};

static void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
  if (threads <= 1 || n < 4096) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t a = t * per, b = std::min(n, a + per);
    if (a >= b) break;
    th.emplace_back(fn, a, b);
  }
  for (auto& t : th) t.join();
}

static int host_threads(int requested) {
  if (requested > 0) return requested;
  const unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}
}  // namespace

extern "C" {

int dc_audio_probe(const char* path, dc_audio_info* info) {
  if (!path || !info) {
    dc::set_error("dc_audio_probe: null argument");
    return DC_ERR_ARG;
  }
  FILE* f = fopen(path, "rb");
  if (!f) {
    dc::set_error("%s: cannot open", path);
    return DC_ERR_ARG;
  }
  WavFmt w;
  const int rc = parse_wav(f, &w, path);
  fclose(f);
  if (rc) return rc;
  info->sample_rate = w.sample_rate;
  info->channels = w.channels;
  info->bits_per_sample = w.bits;
  info->is_float = w.format == 3;
  info->frames = w.data_bytes / w.block_align;
  return DC_OK;
}

int dc_audio_resampled_length(int64_t n_in, int sr_in, int sr_out, int64_t* n_out) {
  if (!n_out || n_in < 0 || sr_in <= 0 || sr_out <= 0) {
    dc::set_error("dc_audio_resampled_length: bad argument");
    return DC_ERR_ARG;
  }
  Resampler r;
  int a = sr_out, b = sr_in;
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  r.up = sr_out / a;
  r.down = sr_in / a;
  *n_out = r.out_len(n_in);
  return DC_OK;
}

int dc_audio_resample(const float* in, int64_t n_in, int sr_in, int sr_out, int zeros, double beta, float* out,
                      int64_t cap, int64_t* n_out, int threads) {
  if (!in || !out || !n_out || n_in < 0 || sr_in <= 0 || sr_out <= 0 || zeros < 1 || zeros > 256 || beta < 0) {
    dc::set_error("dc_audio_resample: bad argument");
    return DC_ERR_ARG;
  }
  if (sr_in == sr_out) {
    if (cap < n_in) {
      dc::set_error("dc_audio_resample: output buffer too small (%lld < %lld)", (long long)cap, (long long)n_in);
      return DC_ERR_WORKSPACE;
    }
    memcpy(out, in, (size_t)n_in * 4);
    *n_out = n_in;
    return DC_OK;
  }
  Resampler r;
  r.design(sr_in, sr_out, zeros, beta);
  const int64_t n = r.out_len(n_in);
  if (cap < n) {
    dc::set_error("dc_audio_resample: output buffer too small (%lld < %lld)", (long long)cap, (long long)n);
    return DC_ERR_WORKSPACE;
  }
  parallel_for(n, host_threads(threads), [&](int64_t a, int64_t b) { r.run(in, n_in, out, a, b); });
  *n_out = n;
  return DC_OK;
}

// One file -> mono float32 at target_sr (target_sr <= 0: the file's own rate).  `frame_offset` / `max_frames` select a
// window of the file before resampling (load_and_resample_audio's `limited` crop; max_frames <= 0: to the end).
static int load_one(const char* path, int target_sr, int zeros, double beta, int64_t frame_offset, int64_t max_frames,
                    float* out, int64_t cap, int64_t* n_out, int* sr_out, int threads) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    dc::set_error("%s: cannot open", path);
    return DC_ERR_ARG;
  }
  WavFmt w;
  int rc = parse_wav(f, &w, path);
  if (rc) {
    fclose(f);
    return rc;
  }
  int64_t frames = w.data_bytes / w.block_align;
  if (frame_offset < 0) frame_offset = 0;
  if (frame_offset > frames) frame_offset = frames;
  frames -= frame_offset;
  if (max_frames > 0 && frames > max_frames) frames = max_frames;
  const int sr = target_sr > 0 ? target_sr : w.sample_rate;
  if (sr_out) *sr_out = sr;
  if (sr == w.sample_rate) {
    if (cap < frames) {
      fclose(f);
      dc::set_error("%s: output buffer too small (%lld < %lld samples)", path, (long long)cap, (long long)frames);
      return DC_ERR_WORKSPACE;
    }
    rc = read_frames(f, w, frame_offset, frames, true, out, 0, path);
    fclose(f);
    if (rc) return rc;
    *n_out = frames;
    return DC_OK;
  }
  std::vector<float> tmp((size_t)std::max<int64_t>(frames, 1));
  rc = read_frames(f, w, frame_offset, frames, true, tmp.data(), 0, path);
  fclose(f);
  if (rc) return rc;
  return dc_audio_resample(tmp.data(), frames, w.sample_rate, sr, zeros, beta, out, cap, n_out, threads);
}

int dc_audio_load(const char* path, int target_sr, int zeros, double beta, int64_t frame_offset, int64_t max_frames,
                  float* out, int64_t cap, int64_t* n_out, int* sr_out) {
  if (!path || !out || !n_out) {
    dc::set_error("dc_audio_load: null argument");
    return DC_ERR_ARG;
  }
  return load_one(path, target_sr, zeros, beta, frame_offset, max_frames, out, cap, n_out, sr_out, 0);
}

int dc_audio_load_batch(const char* const* paths, int n, int target_sr, int zeros, double beta, float* out,
                        int64_t row_stride, int64_t left_pad, int64_t* lengths, int* status, int threads) {
  if (!paths || !out || !lengths || n < 0 || row_stride <= 0 || left_pad < 0 || left_pad > row_stride) {
    dc::set_error("dc_audio_load_batch: bad argument");
    return DC_ERR_ARG;
  }
  const int nt = std::min(host_threads(threads), std::max(n, 1));
  std::atomic<int> next(0), failed(0);
  std::string first_error;
  std::atomic<bool> have_error(false);
  auto worker = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      float* row = out + (int64_t)i * row_stride;
      for (int64_t k = 0; k < left_pad; ++k) row[k] = 0.f;
      int64_t got = 0;
      const int rc = load_one(paths[i], target_sr, zeros, beta, 0, 0, row + left_pad, row_stride - left_pad, &got,
                              nullptr, 1);
      if (rc != DC_OK) {
        got = 0;
        failed.fetch_add(1);
        bool expected = false;
        if (have_error.compare_exchange_strong(expected, true)) first_error = dc_last_error();
      }
      for (int64_t k = left_pad + got; k < row_stride; ++k) row[k] = 0.f;  // right pad to the batch length (:133-137)
      lengths[i] = got;
      if (status) status[i] = rc;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  if (failed.load() > 0) {
    dc::set_error("dc_audio_load_batch: %d of %d files failed; first: %s", failed.load(), n, first_error.c_str());
    return status ? DC_OK : DC_ERR_ARG;  // with a status array the caller decides per file (the reference substitutes noise, :157-160)
  }
  return DC_OK;
}

int dc_audio_write_wav(const char* path, const float* data, int64_t n, int sample_rate) {
  if (!path || (!data && n > 0) || n < 0 || sample_rate <= 0) {
    dc::set_error("dc_audio_write_wav: bad argument");
    return DC_ERR_ARG;
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    dc::set_error("%s: cannot create", path);
    return DC_ERR_ARG;
  }
  const uint32_t bytes = (uint32_t)(n * 2);
  uint8_t h[44];
  memcpy(h, "RIFF", 4);
  const uint32_t riff = 36 + bytes;
  memcpy(h + 4, &riff, 4);
  memcpy(h + 8, "WAVEfmt ", 8);
  const uint32_t fmt_sz = 16, sr = (uint32_t)sample_rate, byte_rate = sr * 2;
  const uint16_t tag = 1, ch = 1, align = 2, bits = 16;
  memcpy(h + 16, &fmt_sz, 4);
  memcpy(h + 20, &tag, 2);
  memcpy(h + 22, &ch, 2);
  memcpy(h + 24, &sr, 4);
  memcpy(h + 28, &byte_rate, 4);
  memcpy(h + 32, &align, 2);
  memcpy(h + 34, &bits, 2);
  memcpy(h + 36, "data", 4);
  memcpy(h + 40, &bytes, 4);
  bool ok = fwrite(h, 1, 44, f) == 44;
  std::vector<int16_t> buf(1 << 15);
  for (int64_t done = 0; ok && done < n;) {
    const int64_t m = std::min<int64_t>((int64_t)buf.size(), n - done);
    for (int64_t i = 0; i < m; ++i) {
      // libsndfile's float -> PCM_16 conversion (what soundfile.write does for float32 input): scale by 0x8000,
      // round to nearest (lrintf), clip
      float v = data[done + i] * 32768.f;
      v = v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v);
      buf[(size_t)i] = (int16_t)lrintf(v);
    }
    ok = fwrite(buf.data(), 2, (size_t)m, f) == (size_t)m;
    done += m;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    dc::set_error("%s: write failed", path);
    return DC_ERR_ARG;
  }
  return DC_OK;
}

}  // extern "C"
