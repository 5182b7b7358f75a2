"""GPU log-mel front-end (dc_mel_forward) vs the reference's CPU `spec_transform`
(models/mel_spec.py): committed output of the real reference on bundled audio, and the oracle on synthetic audio.
Floating point: tolerance 1e-4 max-abs on the log-mel (fp32 FFT vs pocketfft)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restatement as R
from tests.conftest import engine, golden
from tests.golden.inputs import make_wav

pytestmark = pytest.mark.gpu


def test_real_audio_matches_reference_mel():
    """3 s of data/org_audios/0001.wav: mel produced by the reference's own preprocess_raw_audio_batch
    (distil_codec.py:99-145: left-pad by one zero sample :134, then spec_transform :138)."""
    g = golden("audio_W1.npz")
    eng = engine("W1", "fp32", 1024)
    audio = F.pad(torch.from_numpy(g["pcm"])[None], (1, 0))                  # (1, n + 1)
    mel = eng.mel(audio.to(eng.device).contiguous())
    ref = torch.from_numpy(np.ascontiguousarray(g["mel"]))
    assert mel.shape == ref.shape == (1, 128, 281)
    assert float((mel.cpu() - ref).abs().max()) < 1e-4


@pytest.mark.parametrize("n", [1024, 24000 + 255, 24000 * 3 + 7, 256 * 8 * 5])
def test_synthetic_audio_matches_oracle(n):
    """Ragged lengths (frame count not a multiple of the 8-frame block, n = 255 mod 256 gives the extra frame,
    SURVEY appendix C), batch 3, incl. a silent clip (every bin at the 1e-5 clamp / 1e-3 magnitude floor)."""
    eng = engine("W1", "fp32", 1024)
    wav = make_wav(3, n, seed=9)
    wav[2] = 0.0
    audio = F.pad(wav, (1, 0))
    ref = R.log_mel(audio[:, None, :])
    mel = eng.mel(audio.to(eng.device).contiguous())
    assert mel.shape == ref.shape
    assert float((mel.cpu() - ref).abs().max()) < 1e-4


def test_mel_feeds_the_encoder_like_the_reference_pipeline():
    """wav -> mel -> encoder -> codes entirely on the device equals codes from the reference's CPU mel."""
    g = golden("audio_W1.npz")
    eng = engine("W1", "fp32")
    audio = F.pad(torch.from_numpy(g["pcm"])[None], (1, 0)).to(eng.device).contiguous()
    codes_gpu_mel, _, _, _ = eng.quantizer(eng.encoder(eng.mel(audio)), want_fup=False)
    ref_mel = torch.from_numpy(np.ascontiguousarray(g["mel"])).to(eng.device)
    codes_ref_mel, _, _, _ = eng.quantizer(eng.encoder(ref_mel), want_fup=False)
    assert (codes_gpu_mel == codes_ref_mel).float().mean().item() >= 0.995
    assert np.mean(codes_gpu_mel.cpu().numpy() == g["codes"][0, :, :, 0]) >= 0.995
