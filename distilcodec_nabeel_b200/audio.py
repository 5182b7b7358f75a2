"""Audio file I/O + resample on the host, bound to the C ABI's dc_audio_* entry points
(csrc/audio_io.cpp: native C++, one host thread per file), mirroring the reference's helpers:

    load_and_resample_audio(file_path, target_sr, mono=True, limited=None)    distilcodec/distil_codec.py:657-684
    load_wav(full_path, sr)                                                    distilcodec/models/meldataset.py:18-20
    DistilCodec.save_wav(...)'s soundfile.write                                distil_codec.py:640-654

plus `load_batch`, which decodes and resamples a list of files in parallel straight into ONE pinned (B, 1 + max_n)
float32 batch laid out exactly like `preprocess_audio_batch` builds it (distil_codec.py:186-191: one zero on the left,
zeros to the longest clip on the right) — the input `Pipeline.tokenize_wav_padded` uploads by DMA.

The resampler is a polyphase Kaiser-windowed sinc (scipy.signal.resample_poly's design, against which the tests pin it);
librosa's soxr_hq differs in sample values only, outside the numeric-parity surface of the hot path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi

SCIPY = (10, 5.0)       # scipy.signal.resample_poly defaults
HQ = (32, 14.769656459379492)   # Kaiser beta for ~145 dB stop band


def probe(path: str) -> dict:
    """-> {sample_rate, channels, bits_per_sample, is_float, frames} of a RIFF/WAVE file."""
    info = _abi.DcAudioInfo()
    _abi.check(_abi.load().dc_audio_probe(str(path).encode(), C.byref(info)), f"dc_audio_probe({path})")
    return {"sample_rate": info.sample_rate, "channels": info.channels, "bits_per_sample": info.bits_per_sample,
            "is_float": bool(info.is_float), "frames": int(info.frames)}


def resampled_length(n: int, sr_in: int, sr_out: int) -> int:
    out = C.c_int64()
    _abi.check(_abi.load().dc_audio_resampled_length(n, sr_in, sr_out, C.byref(out)), "dc_audio_resampled_length")
    return int(out.value)


def resample(x: np.ndarray, orig_sr: int, target_sr: int, quality: Tuple[int, float] = HQ, threads: int = 0) -> np.ndarray:
    """`librosa.resample(y, orig_sr=, target_sr=)` for a 1-D or (channels, n) float array."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 2:
        return np.stack([resample(c, orig_sr, target_sr, quality, threads) for c in x])
    n = resampled_length(x.shape[0], orig_sr, target_sr)
    out = np.empty(n, dtype=np.float32)
    got = C.c_int64()
    _abi.check(_abi.load().dc_audio_resample(x.ctypes.data, x.shape[0], orig_sr, target_sr, quality[0], quality[1],
                                             out.ctypes.data, n, C.byref(got), threads), "dc_audio_resample")
    return out[:got.value]


def load_wav(full_path: str, sr: Optional[int], quality: Tuple[int, float] = HQ) -> Tuple[np.ndarray, int]:
    """models/meldataset.py:18-20 `librosa.load(full_path, sr=sr)`: mono (channel mean) float32 at `sr` (None: the
    file's own rate) -> (wav, sr)."""
    info = probe(full_path)
    target = int(sr) if sr else info["sample_rate"]
    cap = max(1, resampled_length(info["frames"], info["sample_rate"], target))
    out = np.empty(cap, dtype=np.float32)
    got, sr_out = C.c_int64(), C.c_int()
    _abi.check(_abi.load().dc_audio_load(str(full_path).encode(), target, quality[0], quality[1], 0, 0,
                                         out.ctypes.data, cap, C.byref(got), C.byref(sr_out)), f"dc_audio_load({full_path})")
    return out[:got.value], int(sr_out.value)


def load_and_resample_audio(file_path: str, target_sr: int, mono: bool = True, limited: Optional[float] = None,
                            quality: Tuple[int, float] = HQ, rng: Optional[np.random.Generator] = None):
    """distil_codec.py:657-684, same return triple `(y_resampled (1, n) float32, target_sr, audio_duration)`.
    `limited` (seconds): a random window of that length when the clip is longer (the reference draws the start with
    np.random.randint, :670-672).  `audio_duration` follows the reference literally: `len(y) / orig_sr` with `y` as
    librosa.load(mono=False) returns it, i.e. the number of CHANNELS for multi-channel files (:668)."""
    info = probe(file_path)
    sr0, frames, ch = info["sample_rate"], info["frames"], info["channels"]
    len_y = frames if ch == 1 else ch
    audio_duration = len_y / sr0
    offset, count = 0, 0
    if limited is not None and audio_duration > limited and len_y - int(sr0 * limited) > 1000 and ch == 1:
        hi = len_y - int(sr0 * limited)
        offset = int(rng.integers(0, hi)) if rng is not None else int(np.random.randint(0, hi))
        count = int(sr0 * limited)
    if not mono and ch > 1:
        raise NotImplementedError("multi-channel output (mono=False) is not on the hot path; the codec is mono")
    n_in = count or (frames - offset)
    cap = max(1, resampled_length(n_in, sr0, int(target_sr)))
    out = np.empty(cap, dtype=np.float32)
    got, sr_out = C.c_int64(), C.c_int()
    _abi.check(_abi.load().dc_audio_load(str(file_path).encode(), int(target_sr), quality[0], quality[1], offset, count,
                                         out.ctypes.data, cap, C.byref(got), C.byref(sr_out)), f"dc_audio_load({file_path})")
    return out[None, :got.value], target_sr, audio_duration


def load_batch(paths: Sequence[str], target_sr: int, quality: Tuple[int, float] = HQ, threads: int = 0,
               left_pad: int = 1, pin: Optional[bool] = None, on_error: str = "raise"):
    """Decode + resample `paths` in parallel into ONE float32 batch (B, left_pad + max_n) laid out like
    `preprocess_audio_batch` (distil_codec.py:186-191).  -> (batch, lengths (B,) int64, status list).
    on_error="noise" substitutes 1 s of N(0, 0.05) like the reference does for unreadable files (:157-160)."""
    lib = _abi.load()
    B = len(paths)
    infos = []
    for p in paths:
        try:
            infos.append(probe(p))
        except RuntimeError:
            if on_error == "raise":
                raise
            infos.append(None)
    lens = [resampled_length(i["frames"], i["sample_rate"], target_sr) if i else target_sr for i in infos]
    stride = left_pad + max(lens + [1])
    if pin is None:
        pin = torch.cuda.is_available()
    batch = torch.empty(B, stride, dtype=torch.float32, pin_memory=bool(pin))
    lengths = torch.zeros(B, dtype=torch.int64)
    status = (C.c_int * max(B, 1))()
    arr = (C.c_char_p * max(B, 1))(*[str(p).encode() for p in paths])
    rc = lib.dc_audio_load_batch(arr, B, int(target_sr), quality[0], quality[1], batch.data_ptr(), stride, left_pad,
                                 lengths.data_ptr(), status, threads)
    _abi.check(rc, "dc_audio_load_batch")
    st = [int(status[i]) for i in range(B)]
    bad = [i for i, s in enumerate(st) if s != 0]
    if bad and on_error == "raise":
        raise RuntimeError(f"audio.load_batch: {len(bad)} file(s) failed, first: {paths[bad[0]]}: "
                           f"{lib.dc_last_error().decode('utf-8', 'replace')}")
    for i in bad:   # the reference's substitute for an unreadable clip (distil_codec.py:157-160)
        noise = torch.from_numpy((np.random.normal(size=(target_sr,)) * 0.05).astype(np.float32))
        batch[i].zero_()
        batch[i, left_pad:left_pad + target_sr] = noise
        lengths[i] = target_sr
    return batch, lengths, st


def write_wav(path: str, audio, sample_rate: int) -> None:
    """`soundfile.write(path, audio_float32, sr)` as save_wav uses it (distil_codec.py:651): 16-bit PCM mono."""
    a = np.ascontiguousarray(torch.as_tensor(audio).detach().float().cpu().numpy().reshape(-1), dtype=np.float32)
    _abi.check(_abi.load().dc_audio_write_wav(str(path).encode(), a.ctypes.data, a.shape[0], int(sample_rate)),
               f"dc_audio_write_wav({path})")


def install(distil_codec_module, meldataset_module=None, quality: Tuple[int, float] = HQ) -> None:
    """Point the reference's module-level helpers at the native loader without touching its files:
    `distilcodec.distil_codec.load_and_resample_audio`, `distilcodec.distil_codec.load_wav` (imported there from
    models/meldataset.py) and, if given, `meldataset.load_wav`."""
    def _lw(full_path, sr):
        return load_wav(full_path, sr, quality)

    def _lra(file_path, target_sr, mono=True, limited=None):
        return load_and_resample_audio(file_path, target_sr, mono, limited, quality)

    distil_codec_module.load_and_resample_audio = _lra
    if hasattr(distil_codec_module, "load_wav"):
        distil_codec_module.load_wav = _lw
    if meldataset_module is not None:
        meldataset_module.load_wav = _lw
