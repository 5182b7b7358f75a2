"""Stage-level and end-to-end parity of the CUDA path (through the C ABI) with the reference's golden outputs and
the oracle.  fp32 mode: 1e-4; bf16 mode (what enable_bfloat16 selects): 1e-2; max-abs relative to max-abs.
W0 = random-init weights (the BASELINE gate, nearly blind end to end: SURVEY finding 4);
W1 = stress weights where every term is visible and the codes steer the waveform."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests.conftest import engine, golden, rel_err, state_dict
from tests.golden.inputs import make_latents, make_mel

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 1e-2}
# The BASELINE gate (1e-2 in bf16) is defined on random-init weights (W0).  The W1 stress set re-draws LayerScale at
# O(1) (gamma 0.3..0.9 instead of 1e-6), so the bf16 operand rounding of all 36 encoder GEMMs reaches the output
# undamped: max-abs error is ~1e-2 of the output range there, measured 0.8e-2 .. 1.0e-2 depending on summation order.
# That is the precision, not this implementation: the reference's own encoder under its own bf16 autocast on the same
# B200 is 1.17e-2 from its fp32 result on these weights, this path 0.77e-2
# (tests/test_gpu_dropin_joined.py::test_bf16_encoder_error_on_the_stress_weights_is_the_precision_floor gates the ratio).
TOL_ENC_W1_BF16 = 2e-2


@pytest.mark.parametrize("variant", ["W0", "W1"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_encoder_matches_reference(variant, mode):
    g = golden(f"e2e_{variant}.npz")
    eng = engine(variant, mode)
    enc = eng.encoder(torch.from_numpy(g["mel"]).to(eng.device))          # (B, T, 1024)
    tol = TOL_ENC_W1_BF16 if (variant, mode) == ("W1", "bf16") else TOL[mode]
    assert rel_err(enc.transpose(1, 2), torch.from_numpy(g["enc"])) < tol


@pytest.mark.parametrize("variant", ["W0", "W1"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_quantizer_matches_reference_on_identical_latents(variant, mode):
    """quantizer(enc) on the reference's own encoder output: x_pjt_in, codes, quantized_fup, quantized."""
    g = golden(f"e2e_{variant}.npz")
    sd = state_dict(variant)
    eng = engine(variant, mode)
    enc = torch.from_numpy(g["enc"]).transpose(1, 2).contiguous().to(eng.device)
    codes, xin, fup, quant = eng.quantizer(enc)
    assert rel_err(xin.float(), torch.from_numpy(g["x_pjt_in"])) < TOL[mode]
    ref_codes = torch.from_numpy(g["codes"].astype(np.int64))[0, :, :, 0]
    E = sd["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
    assert torch.equal(fup.cpu(), E[codes.cpu()])                          # gather is bit-exact
    if variant == "W1":
        agree = (codes.cpu() == ref_codes).float().mean().item()
        assert agree >= (1.0 if mode == "fp32" else 0.95), agree
    # the search itself is exact for the x the kernel produced: compare with the oracle on that same x
    x_gpu = xin.float().cpu().reshape(-1, xin.shape[-1])
    assert torch.equal(codes.cpu().reshape(-1), R.vq_search(x_gpu, E))
    # everything downstream of the codes, evaluated on the SAME codes (BASELINE: "evaluated on the same codes")
    z_ref = R.quantizer_decode(sd, codes.cpu()[None, :, :, None])
    assert rel_err(quant.transpose(1, 2), z_ref) < TOL[mode]


@pytest.mark.parametrize("variant", ["W0", "W1"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_decode_and_generator_match_reference(variant, mode):
    g = golden(f"e2e_{variant}.npz")
    eng = engine(variant, mode)
    codes = torch.from_numpy(g["codes"].astype(np.int64))[0, :, :, 0].contiguous().to(eng.device)
    z = eng.decode_codes(codes)
    assert rel_err(z.transpose(1, 2), torch.from_numpy(g["z_dec"])) < TOL[mode]
    zq = torch.from_numpy(g["quantized"]).transpose(1, 2).contiguous().to(eng.device)
    wav = eng.generator(zq)
    ref = torch.from_numpy(g["wav"])[:, 0]
    assert wav.shape == ref.shape
    assert rel_err(wav, ref) < TOL[mode]
    assert float((wav.cpu() - ref).abs().max()) < TOL[mode]               # BASELINE's absolute max-abs gate


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_generator_stage_sizes_and_ragged_T(mode):
    """T not a multiple of the 128-row tile, batch 3: every decoder stage against the oracle (W1)."""
    sd = state_dict("W1")
    eng = engine("W1", mode)
    z = make_latents(3, 37, seed=31) * 0.5
    ref = R.generator_forward(sd, z)[:, 0]
    wav = eng.generator(z.transpose(1, 2).contiguous().to(eng.device))
    assert wav.shape == (3, 37 * 256)
    assert rel_err(wav, ref) < TOL[mode]


def test_end_to_end_w1_fp32_codes_and_waveform():
    """mel -> codes -> wav through the three stage calls; W1 so that codes matter (fp32: identical codes)."""
    g = golden("e2e_W1.npz")
    eng = engine("W1", "fp32")
    mel = torch.from_numpy(g["mel"]).to(eng.device)
    enc = eng.encoder(mel)
    codes, _, _, quant = eng.quantizer(enc, want_fup=False)
    wav = eng.generator(quant)
    assert np.array_equal(codes.cpu().numpy(), g["codes"][0, :, :, 0])
    assert rel_err(wav, torch.from_numpy(g["wav"])[:, 0]) < 1e-4


def test_clip_sharding_is_bit_identical_to_unsharded():
    """no cross-clip op -> running logical shards of the batch gives the unsharded result bit for bit."""
    from distilcodec_nabeel_b200.sharding import Pipeline, shard_clips
    eng = engine("W1", "bf16")
    mel = make_mel(5, 64, seed=41)
    pipe = Pipeline(eng)
    codes, wav = pipe.reconstruct(mel.pin_memory())
    for ws in (2, 4):
        parts = [Pipeline(eng).reconstruct(mel[list(shard_clips(5, ws, r))].pin_memory())
                 for r in range(ws) if len(shard_clips(5, ws, r))]
        assert torch.equal(torch.cat([p[0] for p in parts]), codes)
        assert torch.equal(torch.cat([p[1] for p in parts]), wav)
    # chunked passes (workspace-limited) equal one pass
    c2, w2 = Pipeline(eng, chunk=2).reconstruct(mel.pin_memory())
    assert torch.equal(c2, codes) and torch.equal(w2, wav)
    # decode leg from host codes reproduces the waveform of the same codes
    w3 = pipe.decode(codes)
    assert rel_err(w3, wav) < 1e-2


def test_workspace_too_small_is_an_error_not_a_crash():
    import ctypes as C
    from distilcodec_nabeel_b200 import _abi
    eng = engine("W1", "bf16")
    mel = make_mel(1, 16).to(eng.device)
    out = torch.empty(1, 16, 1024, device=eng.device)
    ws = torch.empty(1024, dtype=torch.uint8, device=eng.device)
    rc = eng.lib.dc_encoder_forward(eng.h, mel.data_ptr(), 1, 16, out.data_ptr(), ws.data_ptr(), ws.numel(), 0)
    assert rc == -6 and b"workspace" in eng.lib.dc_last_error()
    with pytest.raises(RuntimeError):
        _abi.check(rc)


@pytest.mark.parametrize("variant,mode", [("W1", "fp32"), ("W1", "bf16"), ("W0", "fp32"), ("W0", "bf16")])
def test_real_audio_through_reference_api_golden(variant, mode):
    """BASELINE configs[0] stand-in (no MP3 decoder in the image): 3 s of the reference's bundled 24 kHz clip
    data/org_audios/0001.wav through the reference's own DistilCodec.encode + decode on CPU
    (tests/golden/make_golden_audio.py) vs the CUDA path on the same log-mel, batch 1, T = 281 frames."""
    g = golden(f"audio_{variant}.npz")
    eng = engine(variant, mode)
    mel = torch.from_numpy(np.ascontiguousarray(g["mel"])).to(eng.device)   # the reference's mel is a transposed view
    enc = eng.encoder(mel)
    codes, xin, _, quant = eng.quantizer(enc, want_fup=False)
    ref_codes = torch.from_numpy(g["codes"].astype(np.int64))[0, :, :, 0]
    agree = (codes.cpu() == ref_codes).float().mean().item()
    if variant == "W1":
        assert agree >= (0.999 if mode == "fp32" else 0.97), agree     # bf16: upstream rounding may flip near-ties
    # decode leg on the REFERENCE's codes (what decode_from_codes receives): waveform within tolerance
    wav = eng.generator(eng.decode_codes(ref_codes.contiguous().to(eng.device)))
    ref_wav = torch.from_numpy(g["wav"])[:, 0]
    assert wav.shape == ref_wav.shape == (1, 281 * 256)
    assert rel_err(wav, ref_wav) < TOL[mode]
    assert float((wav.cpu() - ref_wav).abs().max()) < TOL[mode] * max(1.0, float(ref_wav.abs().max()))
    # the search is exact for the latents the kernels produced
    E = state_dict(variant)["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
    assert torch.equal(codes.cpu().reshape(-1), R.vq_search(xin.float().cpu().reshape(-1, xin.shape[-1]), E))


def test_long_clip_interior_is_shift_invariant():
    """Size-independent property for long inputs: the network has a finite receptive field (encoder +-60 frames,
    decoder ~+-20: SURVEY section 5) and no global-in-time op, so the codes / waveform of a window cut out of a long
    clip equal the long clip's in the window's interior.  Long clip: T = 4100 frames (43.7 s, ragged vs the
    128-frame tiles), window = frames [1031, 3131)."""
    eng = engine("W1", "bf16")
    mel = make_mel(1, 4100, seed=77)
    a0, a1, margin = 1031, 3131, 160
    pipe_codes = []
    wavs = []
    for m in (mel, mel[:, :, a0:a1].contiguous()):
        enc = eng.encoder(m.to(eng.device))
        codes, _, _, quant = eng.quantizer(enc, want_fup=False)
        pipe_codes.append(codes.cpu())
        wavs.append(eng.generator(quant).cpu())
    full_c, win_c = pipe_codes
    full_w, win_w = wavs
    assert full_w.shape == (1, 4100 * 256)
    ci = slice(margin, (a1 - a0) - margin)
    agree = (full_c[:, a0 + margin:a1 - margin] == win_c[:, ci]).float().mean().item()
    assert agree == 1.0, agree
    w_full = full_w[:, (a0 + margin) * 256:(a1 - margin) * 256]
    w_win = win_w[:, margin * 256:((a1 - a0) - margin) * 256]
    assert rel_err(w_win, w_full) < 1e-5


@pytest.mark.parametrize("T", [1, 3])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_tiny_clips_all_stages(T, mode):
    """Clips far shorter than any tile (T = 1 frame = 256 samples): every TMA box is mostly out of bounds."""
    sd = state_dict("W1")
    eng = engine("W1", mode)
    mel = make_mel(2, T, seed=90 + T)
    enc = eng.encoder(mel.to(eng.device))
    tol_enc = TOL_ENC_W1_BF16 if mode == "bf16" else TOL[mode]
    assert rel_err(enc.transpose(1, 2), R.encoder_forward(sd, mel)) < tol_enc
    codes, xin, _, quant = eng.quantizer(enc, want_fup=False)
    E = sd["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
    assert torch.equal(codes.cpu().reshape(-1), R.vq_search(xin.float().cpu().reshape(-1, xin.shape[-1]), E))
    wav = eng.generator(quant)
    ref = R.generator_forward(sd, quant.transpose(1, 2).cpu())[:, 0]
    assert wav.shape == (2, 256 * T) and rel_err(wav, ref) < TOL[mode]


def test_non_default_stream_and_wav_tokenisation():
    """All work is enqueued on the caller's current stream (C-ABI contract); wav -> codes on the device (GPU mel)."""
    from distilcodec_nabeel_b200.sharding import Pipeline
    from tests.golden.inputs import make_wav
    eng = engine("W1", "bf16")
    mel = make_mel(2, 50, seed=61).to(eng.device)
    ref_codes, ref_wav = Pipeline(eng).reconstruct_device(mel)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(eng.device)
    with torch.cuda.stream(side):
        codes, wav = Pipeline(eng).reconstruct_device(mel)
    side.synchronize()
    assert torch.equal(codes, ref_codes) and torch.equal(wav, ref_wav)
    wavs = make_wav(3, 256 * 40 + 100, seed=62)
    c_all = Pipeline(eng).tokenize_wav(wavs.pin_memory())
    c_one = torch.cat([Pipeline(eng).tokenize_wav(wavs[i:i + 1].pin_memory()) for i in range(3)])
    assert c_all.shape == (3, 40) and torch.equal(c_all, c_one)       # batch-independent


def test_large_batch_is_deterministic_and_batch_invariant():
    """Race / aliasing detector at a size where every persistent CTA processes hundreds of tiles: the same input
    twice gives bit-identical codes and waveform, and a clip's result does not depend on its batch (a kernel that
    read activations another CTA had already overwritten would break both)."""
    eng = engine("W1", "bf16")
    mel = make_mel(8, 937, seed=123).repeat(8, 1, 1).to(eng.device)             # 64 clips x 10 s
    from distilcodec_nabeel_b200.sharding import Pipeline
    pipe = Pipeline(eng)
    c1, w1 = pipe.reconstruct_device(mel)
    c2, w2 = pipe.reconstruct_device(mel)
    assert torch.equal(c1, c2) and torch.equal(w1, w2)
    assert torch.equal(c1[:8], c1[8:16]) and torch.equal(w1[:8], w1[56:64])   # repeated clips, different tiles/CTAs
    c3, w3 = pipe.reconstruct_device(mel[:3].contiguous())
    assert torch.equal(c3, c1[:3]) and torch.equal(w3, w1[:3])


def test_ten_minute_clips_tokenize_then_decode_host_buffers():
    """BASELINE configs[4] at its real clip length: 10-minute clips (T = 56,250 frames, 14.4 M samples each) through the
    host-buffer legs (tokenize = wav->codes leg from log-mel, decode = codes->wav leg).  The oracle cannot run this size
    in seconds, so the check is the size-independent window property against the same engine on a short cut (which the
    golden tests tie to the reference): identical codes and waveform in the window's interior."""
    from distilcodec_nabeel_b200.sharding import Pipeline
    eng = engine("W1", "bf16")
    T = 56250
    mel = make_mel(2, 1500, seed=5).repeat(1, 1, 38)[:, :, :T].contiguous().pin_memory()   # periodic, ragged vs every tile
    pipe = Pipeline(eng)
    codes = pipe.tokenize(mel)
    assert codes.shape == (2, T) and codes.dtype == torch.int64
    wav = pipe.decode(codes)
    assert wav.shape == (2, T * 256) and bool(torch.isfinite(wav).all())
    a0, a1, margin = 40000, 41200, 200
    m = mel[:, :, a0:a1].contiguous().to(eng.device)
    c_win, _ = pipe.encode_device(m)
    assert torch.equal(c_win[:, margin:-margin].cpu(), codes[:, a0 + margin:a1 - margin])
    w_win = pipe.decode_device(codes[:, a0:a1].contiguous().to(eng.device)).cpu()
    w_full = wav[:, (a0 + margin) * 256:(a1 - margin) * 256]
    assert rel_err(w_win[:, margin * 256:-margin * 256], w_full) < 1e-5
    # the input is periodic with period 1500 frames: so are the codes away from the clip ends
    assert torch.equal(codes[:, 3000:4500], codes[:, 33000:34500])


def test_time_tiled_legs_are_bit_identical_to_whole_clip():
    """Row f-3: tokenize_long / decode_long (time tiles with real-context halos) against the whole-clip legs."""
    from distilcodec_nabeel_b200.sharding import Pipeline, decode_long, tokenize_long
    eng = engine("W1", "bf16")
    pipe = Pipeline(eng)
    mel = make_mel(2, 4100, seed=9).pin_memory()
    codes = pipe.tokenize(mel)
    for tile in (1000, 1537, 4100, 9000):
        assert torch.equal(tokenize_long(pipe, mel, tile=tile), codes), tile
    wav = pipe.decode(codes)
    for tile in (1000, 1537):
        assert torch.equal(decode_long(pipe, codes, tile=tile), wav), tile


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process_agree():
    """Engines on cuda:0 and cuda:1 of the same process (per-device function attributes, tensor maps, workspaces):
    identical results."""
    from distilcodec_nabeel_b200 import Engine
    from distilcodec_nabeel_b200.sharding import Pipeline
    sd = state_dict("W1")
    mel = make_mel(3, 300, seed=31)
    outs = []
    for d in (0, 1):
        eng = Engine(sd, d, "bf16")
        c, w = Pipeline(eng).reconstruct_device(mel.to(eng.device))
        outs.append((c.cpu(), w.cpu()))
        eng.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("variant", ["W0", "W1"])
def test_fp32_stream_pair_kernel_matches_oracle_and_the_two_kernel_form(variant):
    """Narrowest decoder stage (C = 32): the fused ResBlock-step kernel with the fp32 activation stream and
    block-Toeplitz weights (conv_pair.cu) against (a) the oracle and (b) the same engine with the kernel switched off
    (conv_ws_pair / two-kernel form: same bf16 operands, different fp32 summation order)."""
    sd = state_dict(variant)
    eng = engine(variant, "bf16")
    for B, T, seed in ((3, 37, 31), (2, 301, 32), (1, 1, 33)):
        z = make_latents(B, T, seed=seed) * 0.5
        ref = R.generator_forward(sd, z)[:, 0]
        zd = z.transpose(1, 2).contiguous().to(eng.device)
        outs, counts = {}, {}
        try:
            for form in (0, 1, 2):      # conv_ws_pair | conv_pair on the fp32 stream | conv_pair on the bf16 side buffer
                eng.set_option("pairx", form)
                n0 = eng.launch_count()
                outs[form] = eng.generator(zd)
                counts[form] = eng.launch_count() - n0
        finally:
            eng.set_option("pairx", 2)
        assert counts[0] == counts[1] == counts[2]              # one launch per ResBlock step in every form
        for form in (1, 2):
            assert rel_err(outs[form], ref) < TOL["bf16"], (B, T, form)
            assert rel_err(outs[form], outs[0]) < 2e-4, (B, T, form)   # only the summation order differs


@pytest.mark.parametrize("variant", ["W0", "W1"])
def test_conv_post_on_the_tensor_cores_matches_the_cuda_core_form(variant):
    """conv_post + tanh as a block-Toeplitz GEMM over 8 samples (conv_post.cu, weights split hi + lo) against the
    CUDA-core kernel with fp32 weights (pointwise.cu): same bf16 input, fp32 accumulation in both."""
    eng = engine(variant, "bf16")
    for B, T, seed in ((3, 37, 41), (2, 301, 42), (1, 1, 43)):
        zd = (make_latents(B, T, seed=seed) * 0.5).transpose(1, 2).contiguous().to(eng.device)
        try:
            eng.set_option("post_tc", 0)
            n0 = eng.launch_count()
            y0 = eng.generator(zd)
            n1 = eng.launch_count()
            eng.set_option("post_tc", 1)
            y1 = eng.generator(zd)
            n2 = eng.launch_count()
        finally:
            eng.set_option("post_tc", 1)
        assert n1 - n0 == n2 - n1                       # one launch either way
        assert y0.shape == y1.shape == (B, 256 * T)
        assert rel_err(y1, y0) < 2e-5, (B, T)


@pytest.mark.parametrize("variant", ["W0", "W1"])
def test_fp32_mode_on_tensor_cores_keeps_the_fp32_gate(variant):
    """DC_MODE_FP32 (the reference API's default, enable_bfloat16=False) with its dense layers on the tensor cores
    (option "fp32_tc": two-term bf16 split of both operands, three cross products, accumulation chunked to K <= 256 per
    term and summed with round-to-nearest fp32 adds, gemm_f32x.cu) against the oracle, next to the CUDA-core fp32 kernel:
    both inside 1e-4, the tensor-core form within a small factor of the CUDA-core one."""
    sd = state_dict(variant)
    eng = engine(variant, "fp32")
    mel = make_mel(2, 130, seed=71)
    z = make_latents(2, 61, seed=72) * 0.5
    ref_enc = R.encoder_forward(sd, mel)
    ref_wav = R.generator_forward(sd, z)[:, 0]
    err = {}
    try:
        for tc in (1, 0):
            eng.set_option("fp32_tc", tc)
            enc = eng.encoder(mel.to(eng.device))
            wav = eng.generator(z.transpose(1, 2).contiguous().to(eng.device))
            err[tc] = (rel_err(enc.transpose(1, 2), ref_enc), rel_err(wav, ref_wav))
    finally:
        eng.set_option("fp32_tc", 1)
    print(f"{variant} fp32 mode rel err (encoder, waveform): tensor cores {err[1][0]:.2e} {err[1][1]:.2e}; "
          f"CUDA cores {err[0][0]:.2e} {err[0][1]:.2e}")
    assert max(err[0]) < 1e-4 and max(err[1]) < 1e-4
    assert max(err[1]) < 3e-5, err[1]                 # a 3x margin to the gate on both weight sets
