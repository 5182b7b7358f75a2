"""Row f-4: native audio file decode + resample (csrc/audio_io.cpp behind dc_audio_*), host only.
The resampler is pinned against scipy.signal.resample_poly (the published algorithm it restates), the decoder against
files written by the stdlib `wave` module, the batch layout against what preprocess_audio_batch builds
(distil_codec.py:146-195), and the helper signatures against load_and_resample_audio / load_wav of the reference."""
import os
import struct
import wave
from math import gcd

import numpy as np
import pytest
import scipy.signal as ss
import torch

from distilcodec_nabeel_b200 import audio
from tests.conftest import golden


def _write_pcm(path, data, sr, width):
    """data: float (n, ch) in [-1, 1) -> PCM WAV of `width` bytes per sample via the stdlib writer"""
    n, ch = data.shape
    scale = {1: 128, 2: 32768, 3: 8388608, 4: 2147483648}[width]
    q = np.clip(np.round(data * scale), -scale, scale - 1).astype(np.int64)
    w = wave.open(path, "wb")
    w.setnchannels(ch)
    w.setsampwidth(width)
    w.setframerate(sr)
    if width == 1:
        raw = (q + 128).astype(np.uint8).tobytes()
    elif width == 3:
        b = q.astype("<i4").reshape(-1).view(np.uint8).reshape(-1, 4)[:, :3]
        raw = b.tobytes()
    else:
        raw = q.astype({2: "<i2", 4: "<i4"}[width]).tobytes()
    w.writeframes(raw)
    w.close()
    return q.astype(np.float64) / scale


def _write_float(path, data, sr, bits=32, extensible=False):
    n, ch = data.shape
    payload = data.astype("<f4" if bits == 32 else "<f8").tobytes()
    fmt_tag = 0xFFFE if extensible else 3
    fmt = struct.pack("<HHIIHH", fmt_tag, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits)
    if extensible:
        fmt += struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", 3) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    junk = b"LIST" + struct.pack("<I", 5) + b"abcde\x00"                 # odd-sized chunk + pad byte before data
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(payload)) + payload
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


@pytest.mark.parametrize("width", [1, 2, 3, 4])
@pytest.mark.parametrize("ch", [1, 2])
def test_pcm_decode_matches_stdlib_writer(tmp_path, width, ch):
    rng = np.random.default_rng(width * 10 + ch)
    x = rng.uniform(-0.9, 0.9, size=(5000, ch))
    p = str(tmp_path / "a.wav")
    q = _write_pcm(p, x, 24000, width)
    info = audio.probe(p)
    assert info == {"sample_rate": 24000, "channels": ch, "bits_per_sample": 8 * width, "is_float": False, "frames": 5000}
    y, sr = audio.load_wav(p, None)
    assert sr == 24000 and y.dtype == np.float32 and y.shape == (5000,)
    np.testing.assert_allclose(y, q.mean(axis=1), atol=2e-7)                # mono = channel mean (librosa.to_mono)


@pytest.mark.parametrize("bits,ext", [(32, False), (64, False), (32, True)])
def test_float_and_extensible_headers(tmp_path, bits, ext):
    x = np.random.default_rng(3).standard_normal((777, 3)).astype(np.float32) * 0.1
    p = str(tmp_path / "f.wav")
    _write_float(p, x, 16000, bits, ext)
    y, sr = audio.load_wav(p, None)
    assert sr == 16000
    np.testing.assert_allclose(y, x.mean(axis=1), atol=1e-7)


@pytest.mark.parametrize("sr_in", [44100, 16000, 48000, 22050, 8000, 32000])
def test_resampler_equals_scipy_resample_poly(sr_in):
    x = np.random.default_rng(sr_in).standard_normal(sr_in + 137).astype(np.float32)
    g = gcd(sr_in, 24000)
    ref = ss.resample_poly(x.astype(np.float64), 24000 // g, sr_in // g)       # scipy's defaults: zeros 10, kaiser 5.0
    y = audio.resample(x, sr_in, 24000, audio.SCIPY)
    assert y.shape == ref.shape and audio.resampled_length(x.shape[0], sr_in, 24000) == ref.shape[0]
    assert float(np.abs(y - ref).max()) < 2e-6
    y1 = audio.resample(x, sr_in, 24000, audio.SCIPY, threads=1)
    assert np.array_equal(y, y1)                                                # thread count does not change a bit


def test_hq_resampler_preserves_a_tone_and_rejects_images():
    sr_in, sr_out, f0 = 44100, 24000, 3000.0
    t = np.arange(sr_in) / sr_in
    x = np.sin(2 * np.pi * f0 * t).astype(np.float32) + np.sin(2 * np.pi * 19000.0 * t).astype(np.float32)  # 19 kHz > 12 kHz Nyquist
    y = audio.resample(x, sr_in, sr_out, audio.HQ)
    to = np.arange(y.shape[0]) / sr_out
    want = np.sin(2 * np.pi * f0 * to)
    mid = slice(2000, -2000)
    assert float(np.abs(y[mid] - want[mid]).max()) < 1e-4                        # 19 kHz gone (< -80 dB), tone intact


def test_reference_helper_signatures(tmp_path):
    rng = np.random.default_rng(5)
    x = rng.uniform(-0.5, 0.5, size=(44100 * 2, 1))
    p = str(tmp_path / "m.wav")
    q = _write_pcm(p, x, 44100, 2)
    y, sr, dur = audio.load_and_resample_audio(p, 24000)                        # distil_codec.py:657-684
    assert sr == 24000 and y.dtype == np.float32 and y.shape == (1, 48000) and dur == pytest.approx(2.0)
    g = gcd(44100, 24000)
    ref = ss.resample_poly(q[:, 0], 24000 // g, 44100 // g, window=("kaiser", audio.HQ[1]))
    assert y.shape[1] == ref.shape[0]
    w, sr2 = audio.load_wav(p, 24000)                                           # meldataset.py:18-20
    assert sr2 == 24000 and np.array_equal(w, y[0])
    y2, _, _ = audio.load_and_resample_audio(p, 24000, limited=0.5, rng=np.random.default_rng(1))
    assert y2.shape == (1, 12000)                                               # the random `limited` window (:669-672)
    class M:            # stands in for the reference's modules: install() only rebinds module-level names
        load_wav = None
        load_and_resample_audio = None
    audio.install(M, M)
    assert np.array_equal(M.load_wav(p, 24000)[0], w) and M.load_and_resample_audio(p, 24000)[0].shape == (1, 48000)


def test_batch_loader_layout_and_errors(tmp_path):
    paths, want = [], []
    for i, (n, sr, ch) in enumerate([(24000, 24000, 1), (30000, 48000, 2), (5000, 24000, 1), (22050, 22050, 1)]):
        x = np.random.default_rng(i).uniform(-0.8, 0.8, size=(n, ch))
        p = str(tmp_path / f"{i}.wav")
        _write_pcm(p, x, sr, 2)
        paths.append(p)
        want.append(audio.load_wav(p, 24000)[0])
    batch, lengths, status = audio.load_batch(paths, 24000, threads=3, pin=False)
    assert status == [0, 0, 0, 0] and lengths.tolist() == [len(w) for w in want]
    assert batch.shape == (4, 1 + max(len(w) for w in want)) and batch.dtype == torch.float32
    for i, w in enumerate(want):                    # preprocess_audio_batch: F.pad(audio, (1, max_len - n)) (:186-189)
        assert float(batch[i, 0]) == 0.0
        assert np.array_equal(batch[i, 1:1 + len(w)].numpy(), w)
        assert float(batch[i, 1 + len(w):].abs().sum()) == 0.0
    b1, l1, _ = audio.load_batch(paths, 24000, threads=1, pin=False)
    assert torch.equal(b1, batch) and torch.equal(l1, lengths)
    bad = paths[:2] + [str(tmp_path / "missing.wav")]
    with pytest.raises(RuntimeError):
        audio.load_batch(bad, 24000, pin=False)
    b2, l2, st2 = audio.load_batch(bad, 24000, pin=False, on_error="noise")     # the reference's substitute (:157-160)
    assert st2[2] != 0 and int(l2[2]) == 24000 and 0.02 < float(b2[2, 1:24001].std()) < 0.08
    with pytest.raises(RuntimeError, match="RIFF"):
        open(tmp_path / "x.wav", "wb").write(b"not a wave file at all")
        audio.probe(str(tmp_path / "x.wav"))


def test_write_wav_round_trip_and_bundled_clip(tmp_path):
    pcm = golden("audio_W0.npz")["pcm"]                                         # 3 s of the reference's 0001.wav, s16 / 32768
    p = str(tmp_path / "out.wav")
    audio.write_wav(p, torch.from_numpy(pcm), 24000)                            # save_wav's soundfile.write (:651)
    w = wave.open(p)
    assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (24000, 1, 2, pcm.shape[0])
    back, sr = audio.load_wav(p, 24000)
    assert sr == 24000 and np.array_equal(back, pcm)                            # s16 -> float -> s16 is lossless
    audio.write_wav(p, np.array([2.0, -2.0, 0.5], dtype=np.float32), 8000)      # clips like libsndfile
    assert np.array_equal(audio.load_wav(p, None)[0], np.array([32767 / 32768, -1.0, 0.5], dtype=np.float32))
    ref_wav = "/root/reference/data/org_audios/0001.wav"
    if os.path.isfile(ref_wav):                                                 # dev container only
        y, _ = audio.load_wav(ref_wav, 24000)
        assert np.array_equal(y[:pcm.shape[0]], pcm)
