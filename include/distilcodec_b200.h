/* distilcodec_b200.h — C ABI of libdistilcodec_b200.so
 *
 * B200-native (sm_100a) implementation of the DistilCodec inference hot path
 *   mel -> ConvNeXt encoder -> single-codebook Euclidean VQ (32768 x 3584) -> HiFiGAN-style decoder -> wav.
 *
 * The reference (nabeelscicom/DistilCodec_nabeel) has no FFI; its boundary is the attribute triple
 * `self.encoder / self.quantizer / self.generator` of `DistilCodec` (distilcodec/distil_codec.py:52-54).
 * Each entry point below replaces one call the reference's API makes on those attributes; the file:line of the
 * replaced interface is cited on every function.  INTEGRATION.md shows the Python (ctypes) binding.
 *
 * Conventions
 *   - every function returns 0 (DC_OK) or a negative dc_status; dc_last_error() gives the message (thread-local)
 *   - all pointers named *_dev are device pointers on the handle's device; the caller (PyTorch's caching
 *     allocator in the shipped host side) owns every input, output and workspace buffer
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it, no hidden
 *     synchronisation or allocation after dc_finalize()
 *   - activations cross the ABI time-major / channels-last ("NLC": (B, T, C) row-major) unless stated otherwise;
 *     the reference's channels-first (B, C, T) tensors are strided views of the same memory
 *   - a handle is bound to one device and one numeric mode and is not re-entrant
 */
#ifndef DISTILCODEC_B200_H_
#define DISTILCODEC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dc_handle_s* dc_handle;

typedef enum {
  DC_OK = 0,
  DC_ERR_ARG = -1,       /* null pointer, bad enum, unknown tensor name                     */
  DC_ERR_SHAPE = -2,     /* shape not supported by the kernels                              */
  DC_ERR_CUDA = -3,      /* a CUDA runtime / driver call failed                             */
  DC_ERR_ARCH = -4,      /* device is not compute capability 10.x                           */
  DC_ERR_STATE = -5,     /* call order violated (e.g. forward before finalize)              */
  DC_ERR_WORKSPACE = -6  /* workspace too small (see dc_workspace_bytes)                    */
} dc_status;

/* numeric mode: what `enable_bfloat16` selects in the reference API (distil_codec.py:545,550,581,590) */
typedef enum {
  DC_MODE_FP32 = 0, /* fp32-accurate: tensor cores on split-bf16 operands with chunked fp32 accumulation
                       (option "fp32_tc", default) or CUDA-core fp32 kernels; parity target 1e-4          */
  DC_MODE_BF16 = 1  /* tcgen05 tensor-core kernels, bf16 operands / fp32 accumulate; target 1e-2   */
} dc_mode;

typedef enum { DC_STAGE_ENCODER = 0, DC_STAGE_QUANTIZER = 1, DC_STAGE_DECODE_CODES = 2, DC_STAGE_GENERATOR = 3 } dc_stage;

int dc_version(void);
const char* dc_last_error(void);

/* Architecture hyper-parameters: the fields of configs/model_config.json the hot path depends on
 * (`encoder`, `quantizer`, `decoder` sections; consumed at distil_codec.py:46-54). */
typedef struct {
  int n_mels;             /* encoder.input_channels                128 */
  int enc_depths[4];      /* encoder.depths                        3,3,9,3 */
  int enc_dims[4];        /* encoder.dims                          256,512,768,1024 */
  int codebook_size;      /* quantizer.codebook_size               32768 */
  int codebook_dim;       /* quantizer.codebook_dim                3584 */
  int n_ups;              /* len(decoder.upsample_rates)           5 */
  int up_rates[8];        /* decoder.upsample_rates                8,4,2,2,2 */
  int up_kernels[8];      /* decoder.upsample_kernel_sizes         16,12,4,4,4 */
  int up_initial_channel; /* decoder.upsample_initial_channel      1024 */
  int rb_kernels[3];      /* decoder.resblock_kernel_sizes         3,7,11 */
  int rb_dilations[3];    /* decoder.resblock_dilation_sizes[*]    1,3,5 */
  int pre_kernel;         /* decoder.pre_conv_kernel_size          13 */
  int post_kernel;        /* decoder.post_conv_kernel_size         13 */
} dc_config;
int dc_default_config(dc_config* cfg);

/* Lifetime.  Replaces module construction + `.to(device)` (distil_codec.py:52-54, 72-75).  cfg NULL = defaults. */
int dc_create(int device, int mode, const dc_config* cfg, dc_handle* out);
int dc_destroy(dc_handle h);
/* Tunables: "vq_window" (scale of the candidate window of the bf16 tensor-core scorer; 1.0 [default] = the rigorous
 * bound built from the exact rounding-residual norms of the row and of the codebook, so the exact winner always
 * survives to the fp32 re-score), "vq_tensor_core" (1 = tcgen05 scorer [default], 0 = CUDA-core scorer), "vq_x2_exact"
 * (0 [default] = ||x||^2 summed in the order of the reference's CPU path, ATen cascade_sum; 1 = correctly rounded),
 * "fuse_pairs" (1 [default] = the narrow decoder stages run each conv1 -> SiLU -> conv2 step as one kernel),
 * "pairx" (which fused kernel the C = 32 stage uses: 0 conv_ws_pair, 1 conv_pair on the fp32 stream, 2 [default]
 * conv_pair with the bf16 side buffer), "cta_pairs" (2 [default] = the tensor-bound kernels run as thread-block clusters of
 * two CTAs that TMA-multicast the weight / codebook tiles they share; 1 = single CTAs; results are bit-identical),
 * "post_tc" (bf16 mode: 1 [default] = conv_post + tanh as a block-Toeplitz tensor-core GEMM, 0 = the CUDA-core kernel),
 * "epi_prefetch" (L2 prefetch of the residual / mean operands of the next tile: 1 [default] = where the layer is
 * HBM-latency-bound, 2 = also k = 7 at C = 256, 3 = every layer, 0 = never),
 * "fp32_tc" (DC_MODE_FP32 only: 1 [default] = dense layers on the tensor cores with split-bf16 operands and chunked fp32
 * accumulation, 0 = the CUDA-core fp32 kernel). */
int dc_set_option(dc_handle h, const char* key, double value);

/* Weight ingestion.  Replaces `load_state_dict` on the three modules (distil_codec.py:91-94): call once per
 * state_dict entry with the reference's key (prefixed `encoder.` / `quantizer.` / `generator.`), an fp32 device
 * pointer and the tensor's shape; then dc_finalize() folds weight_norm (models/generators.py:50,70,106),
 * repacks every matrix into the kernels' layouts and precomputes ||c||^2.  The library copies what it needs;
 * only `quantizer.grvq.rvqs.0.layers.0._codebook.embed` (470 MB fp32) is referenced in place and must stay
 * alive and unmodified until the next dc_finalize() or dc_destroy(). */
int dc_set_tensor(dc_handle h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dc_finalize(dc_handle h, void* stream);

/* Scratch requirement of one stage for a batch of B clips x T frames. */
int dc_workspace_bytes(dc_handle h, int stage, int B, int T, size_t* bytes);

/* encoder(mel): ConvNeXtEncoder.forward, models/encoders.py:68-76 (call sites distil_codec.py:520,551).
 *   mel_ncl_dev : fp32 (B, 128, T) channels-first, exactly what LogMelSpectrogram emits
 *   enc_nlc_dev : fp32 (B, T, 1024) */
int dc_encoder_forward(dc_handle h, const float* mel_ncl_dev, int B, int T, float* enc_nlc_dev, void* ws_dev,
                       size_t ws_bytes, void* stream);

/* quantizer(enc): DownsampleGRVQ.forward, vector_quantization/grfvq.py:105-132 (call sites distil_codec.py:525,555).
 *   enc_nlc_dev       : fp32 (B, T, 1024)
 *   codes_dev         : int64 (B, T)            == GRVQResult.codes[0, :, :, 0]
 *   x_pjt_in_dev      : (B, T, 3584) bf16 in DC_MODE_BF16, fp32 in DC_MODE_FP32   == GRVQResult.x_pjt_in
 *   fup_dev           : fp32 (B, T, 3584)       == GRVQResult.quantized_fup (codebook rows); nullable
 *   quantized_nlc_dev : fp32 (B, T, 1024)       == GRVQResult.quantized viewed channels-last */
int dc_quantizer_forward(dc_handle h, const float* enc_nlc_dev, int B, int T, int64_t* codes_dev, void* x_pjt_in_dev,
                         float* fup_dev, float* quantized_nlc_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* quantizer.encode(enc): DownsampleGRVQ.encode, grfvq.py:134-139 — the codes alone (what a tokenisation job keeps of
 * `DistilCodec.encode`, distil_codec.py:545-563): downsample + project_in + search; no codebook gather, no project_out /
 * upsample tail, no x_pjt_in / quantized_fup stores.  Same workspace requirement as dc_quantizer_forward.
 *   enc_nlc_dev : fp32 (B, T, 1024);  codes_dev : int64 (B, T) */
int dc_quantizer_encode(dc_handle h, const float* enc_nlc_dev, int B, int T, int64_t* codes_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);

/* Nearest-code search only: EuclideanCodebook.forward eval path, vector_quantization/utils/
 * vector_quantize_pytorch.py:462-538 (cdist :41-45, argmax :96).  Returns exactly
 *   argmax_j -sqrt(max((x2 + c2_j) + (-2 * x.c_j), 0))   with the lowest index on ties,
 * all in fp32 like the reference, with x.c_j evaluated exactly (fp64) then rounded to fp32.
 *   x_dev    : (N, 3584) rows, dtype x_is_bf16 ? bf16 : fp32
 *   x2_dev   : optional fp32 (N) row square-norms as the caller's reference computes them (strict parity with
 *              a particular reduction order, e.g. a CUDA reference); NULL = computed here in the summation order
 *              of the reference's CPU path (`(x ** 2).sum(-1)`, ATen cascade_sum), see option "vq_x2_exact"
 *   stats_host : optional int[4] {rows, exactly re-scored candidates, rows sent to the exhaustive pass, rows decided
 *                without re-scoring}; forces a stream synchronisation */
int dc_vq_search(dc_handle h, const void* x_dev, int x_is_bf16, const float* x2_dev, int64_t N, int64_t* codes_dev,
                 void* ws_dev, size_t ws_bytes, void* stream, int* stats_host);
int dc_vq_workspace_bytes(dc_handle h, int64_t N, int x_is_bf16, size_t* bytes);

/* quantizer.decode(indices): DownsampleGRVQ.decode, grfvq.py:141-146 -> get_output_from_indices,
 * utils/residual_vq.py:301-303,135-138 (call sites distil_codec.py:591,630).
 *   codes_dev : int64 (B, T);  z_nlc_dev : fp32 (B, T, 1024) */
int dc_quantizer_decode(dc_handle h, const int64_t* codes_dev, int B, int T, float* z_nlc_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);

/* generator(z): HiFiGANGenerator.forward, models/generators.py:118-147 (call sites distil_codec.py:528,577,592,631).
 *   z_nlc_dev : fp32 (B, T, 1024);  wav_dev : fp32 (B, 256*T) */
int dc_generator_forward(dc_handle h, const float* z_nlc_dev, int B, int T, float* wav_dev, void* ws_dev,
                         size_t ws_bytes, void* stream);

/* spec_transform(audios): LogMelSpectrogram.forward, models/mel_spec.py:109-122 -> LinearSpectrogram.forward :26-57
 * (call sites distil_codec.py:138,191; the reference forces this stage onto the CPU, mel_spec.py:39).  Needs the two
 * non-persistent buffers of that module, handed over with dc_set_tensor under the names `spec_transform.fb`
 * (513, 128) and `spec_transform.spectrogram.window` (1024); n_fft 1024 / hop 256 / 128 mels only.
 *   audio_dev   : fp32 (B, Ls) mono samples exactly as the reference passes them (already left-padded by one zero)
 *   mel_ncl_dev : fp32 (B, 128, T), T = (Ls - 256) / 256 + 1 (integer division) */
int dc_mel_forward(dc_handle h, const float* audio_dev, int B, int Ls, float* mel_ncl_dev, void* stream);

/* Strided host<->device (or device<->device) copy of `rows` rows of `width_bytes`, enqueued on `stream` — the time
 * tiles and per-clip crops of the host-buffer legs (mel[:, :, lo:hi] in, codes[:, s:e] / wav[:, s*256:e*256] out:
 * the slicing `save_wav` does, distil_codec.py:640-654, and the `[:hop_len]` crops of `encode`, :556-570) move
 * straight between PINNED host tensors and the device by DMA, with no pageable staging copy.  Pointers may be host
 * (pinned, for the copy to be asynchronous) or device; the current device must be the one that owns the device side. */
int dc_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                    void* stream);

/* Host utility (no CUDA call): the block-Toeplitz operand the tensor-core form of conv_post + tanh
 * (models/generators.py:141-145) multiplies by.  w: the weight-norm-folded conv_post weight as [13 taps][32 channels]
 * fp32; wt: [16][768] bf16 bit patterns, rows 0-7 = bf16(W), rows 8-15 = bf16(W - bf16(W)), with
 * W[r][o*256 + i*32 + c] = w[8*(o-1) + i - r + 6][c] (0 outside the 13 taps), so that
 * y[8q + r] = sum_o sum_k S[q - 1 + o][k] * W[r][o*256 + k] for the (L/8, 256) view S of the (L, 32) input.
 * dc_finalize() calls it; exported so that the layout can be checked against the plain convolution without a GPU. */
int dc_conv_post_toeplitz_weights(const float* w, uint16_t* wt);

/* ---- audio file I/O + resample on the host -------------------------------------------
 * Replaces the reference's single-threaded librosa path that feeds the hot path:
 *   load_and_resample_audio  distil_codec.py:657-684  (librosa.load(sr=None, mono=False) + librosa.resample + mean)
 *   load_wav                 models/meldataset.py:18-20 (librosa.load(path, sr=sr)), called per clip at distil_codec.py:155
 *   save_wav                 distil_codec.py:640-654  (soundfile.write of float32 -> 16-bit PCM)
 * All buffers are caller-owned HOST memory (pin them to upload asynchronously); no CUDA call is made.
 * Formats: RIFF/WAVE PCM 8/16/24/32, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE, any channel count (mono = mean).
 * Resampler: polyphase Kaiser-windowed sinc, scipy.signal.resample_poly's design: half-length `zeros`*max(up,down)
 * taps, Kaiser `beta`; (10, 5.0) are scipy's defaults, (32, 14.77) a high-quality setting.  Output length
 * ceil(n*up/down).  Not bit-equal to librosa's soxr_hq: sample values only, outside the parity surface. */
typedef struct {
  int sample_rate, channels, bits_per_sample, is_float;
  int64_t frames;
} dc_audio_info;
int dc_audio_probe(const char* path, dc_audio_info* info);
int dc_audio_resampled_length(int64_t n_in, int sr_in, int sr_out, int64_t* n_out);
/* threads <= 0: all host threads */
int dc_audio_resample(const float* in, int64_t n_in, int sr_in, int sr_out, int zeros, double beta, float* out,
                      int64_t cap, int64_t* n_out, int threads);
/* one file -> mono float32 at target_sr (<= 0: the file's rate) of frames [frame_offset, frame_offset + max_frames)
 * (max_frames <= 0: to the end; the `limited` crop of load_and_resample_audio); *sr_out = the rate of `out` */
int dc_audio_load(const char* path, int target_sr, int zeros, double beta, int64_t frame_offset, int64_t max_frames,
                  float* out, int64_t cap, int64_t* n_out, int* sr_out);
/* n files -> rows of a (n, row_stride) float32 batch in parallel: row i = `left_pad` zeros (the reference's one-sample
 * left pad, distil_codec.py:134), the clip, zeros to the end of the row (:133-137); lengths[i] = samples of clip i.
 * status (nullable) receives the per-file dc_status; without it any failure fails the call. */
int dc_audio_load_batch(const char* const* paths, int n, int target_sr, int zeros, double beta, float* out,
                        int64_t row_stride, int64_t left_pad, int64_t* lengths, int* status, int threads);
/* mono float32 in [-1, 1] -> 16-bit PCM WAV with libsndfile's float conversion (scale 0x8000, round, clip) */
int dc_audio_write_wav(const char* path, const float* data, int64_t n, int sample_rate);

/* Layout helpers between the reference's channels-first tensors and the ABI's channels-last ones. */
int dc_ncl_to_nlc(const float* in_dev, float* out_dev, int B, int C, int T, void* stream);
int dc_nlc_to_ncl(const float* in_dev, float* out_dev, int B, int T, int C, void* stream);

/* ---- op-level entry points (used by the parity tests and micro-benchmarks; same kernels as the stages) ---- */

/* Shifted-row implicit-GEMM convolution, the one shape every dense layer of the path maps to:
 *   out[b,t,n] = act( sum_{j<J} sum_{c<C} a[b, t + shift0 + j*dil, c] * w[n, j*C + c] + bias[n] ) (+ res[b,t,n])
 * a (B,T,C), w (N, J*C), bias (N) or NULL, res/out (B,T,N); all fp32 device arrays (bf16 mode rounds a and w to
 * bf16 and runs the tcgen05 kernel, fp32 mode runs the CUDA-core kernel).  act: 0 none, 1 exact GELU, 2 SiLU. */
int dc_op_conv_gemm(dc_handle h, const float* a_dev, const float* w_dev, const float* bias_dev, const float* res_dev,
                    float* out_dev, int B, int T, int C, int J, int shift0, int dil, int N, int act, void* stream);
/* depthwise k7 conv + LayerNorm(eps 1e-6) over C of a channels-last fp32 tensor (models/convnext_utils.py:265-268);
 * dw_w_dev (C,1,7) reference layout or NULL for LayerNorm only (convnext_utils.py:203-213). */
int dc_op_dwconv_ln(dc_handle h, const float* in_dev, const float* dw_w_dev, const float* dw_b_dev,
                    const float* ln_w_dev, const float* ln_b_dev, float* out_dev, int B, int T, int C, void* stream);

/* Per-kernel-class device timing for roofline reports (thread-local, off by default).  While enabled, every kernel
 * launch is bracketed by a CUDA event pair on its stream and booked with its ALGORITHMIC flops and HBM bytes.
 * dc_profile_collect synchronises the recorded events, returns one row per kernel class and layer shape that
 * launched (the part of `name` before '[' is the class) and clears the records. */
typedef struct {
  char name[64];     /* kernel class [+ shape], e.g. "gemm_tc[C256 N256 J3 d1 e5]", "vq_score", "dwconv_ln[C768]" */
  uint64_t launches;
  double ms;         /* sum of event-pair durations */
  double flops;      /* algorithmic FLOPs (2 x MACs of the reference's layer), summed over launches */
  double bytes;      /* algorithmic HBM bytes (each operand once), summed over launches */
} dc_profile_row;
int dc_profile_enable(int on);
int dc_profile_collect(dc_profile_row* rows, int cap, int* n);

/* number of kernels this library has launched on the calling thread since load (bench.py's gpu_launches) */
uint64_t dc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DISTILCODEC_B200_H_ */
