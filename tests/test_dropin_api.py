"""The drop-in boundary against the REAL reference API (needs the reference: /root/reference or baseline/_ref).

`patch(codec)` replaces the three attributes of a real `DistilCodec` (distil_codec.py:52-54); its public methods then
run unchanged on the shim modules.  There is no GPU here, so the shims' engine is replaced by an oracle-backed stand-in
(test infrastructure; the product has no CPU path) — what is under test is the CONTRACT between the unchanged reference
methods and the shim modules: argument layouts, GRVQResult fields, per-clip post-processing, token bookkeeping."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import restatement as R
from tests.conftest import state_dict

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not importable (neither /root/reference nor baseline/_ref)")


class OracleEngine:
    """Same method surface as distilcodec_nabeel_b200.Engine, computed by the oracle on CPU."""

    def __init__(self, sd):
        self.sd, self.device = sd, torch.device("cpu")

    def ncl_to_nlc(self, x):
        return x.transpose(1, 2).contiguous()

    def encoder(self, mel):
        return R.encoder_forward(self.sd, mel).transpose(1, 2).contiguous()

    def quantizer(self, enc_nlc, want_fup=True):
        q = R.quantizer_forward(self.sd, enc_nlc.transpose(1, 2))
        return q["codes"][0, :, :, 0], q["x_pjt_in"], q["quantized_fup"], q["quantized"].transpose(1, 2).contiguous()

    def quantizer_encode(self, enc_nlc):
        return self.quantizer(enc_nlc, want_fup=False)[0]

    def decode_codes(self, codes):
        return R.quantizer_decode(self.sd, codes[None, :, :, None]).transpose(1, 2).contiguous()

    def generator(self, z_nlc):
        return R.generator_forward(self.sd, z_nlc.transpose(1, 2))[:, 0]


@pytest.fixture(scope="module")
def codecs():
    from distilcodec_nabeel_b200 import patch
    sd = state_dict("W1", 1024)
    ref = ref_loader.build_reference_codec(sd, codebook_size=1024)
    patched = patch(ref_loader.build_reference_codec(sd, codebook_size=1024), device="cpu")
    eng = OracleEngine(sd)
    patched.encoder._engines.get = lambda mode=None: eng       # the only substitution: where the FLOPs happen
    return ref, patched


def _pcm(seconds=1.0):
    """the first seconds of the reference's data/org_audios/0001.wav (committed in tests/golden/audio_W0.npz, so the
    test also runs where only the installed package baseline/_ref is present)"""
    from tests.conftest import golden
    return golden("audio_W0.npz")["pcm"][:int(seconds * 24000)].copy()


def test_patch_replaces_exactly_the_three_attributes(codecs):
    from distilcodec_nabeel_b200 import B200Encoder, B200Generator, B200Quantizer
    ref, patched = codecs
    assert isinstance(patched.encoder, B200Encoder) and isinstance(patched.quantizer, B200Quantizer)
    assert isinstance(patched.generator, B200Generator)
    assert type(patched) is type(ref)                              # the class and its methods are untouched
    for name in ("encoder", "quantizer", "generator"):
        a, b = getattr(ref, name).state_dict(), getattr(patched, name).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(patched.quantizer.grvq.codebooks, ref.quantizer.grvq.codebooks)


def test_encode_runs_unchanged_on_the_shims(codecs):
    """DistilCodec.encode (distil_codec.py:545-573) incl. its per-clip post-processing and token lookups."""
    ref, patched = codecs
    clips = [[_pcm(1.0), 24000], [_pcm(0.6), 24000]]               # ragged batch: padded to the longest
    with torch.no_grad():
        r0, gen0, hop0 = ref.encode(copy.deepcopy(clips), enable_bfloat16=False, raw_audio=True)
        r1, gen1, hop1 = patched.encode(copy.deepcopy(clips), enable_bfloat16=False, raw_audio=True)
    assert gen0 == gen1 and hop0 == hop1
    assert [f.name for f in r0.__dataclass_fields__.values()] == [f.name for f in r1.__dataclass_fields__.values()]
    assert torch.equal(r0.codes, r1.codes) and r1.codes.dtype == torch.int64 and r1.codes.shape == r0.codes.shape
    assert r0.codes_list == r1.codes_list and len(r1.codes_list) == 2
    for a, b in zip(r0.x_pjt_in_list + r0.quantized_fup_list, r1.x_pjt_in_list + r1.quantized_fup_list):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-5)
    assert r1.quantized.shape == r0.quantized.shape and torch.allclose(r0.quantized, r1.quantized, atol=1e-5)
    assert float(r1.total_loss) == float(r0.total_loss) == 0.0


def test_forward_pieces_and_decode_layouts(codecs):
    ref, patched = codecs
    with torch.no_grad():
        _, mel, _, _ = ref.preprocess_raw_audio_batch([[_pcm(0.5), 24000]])
        enc0, enc1 = ref.encoder(mel), patched.encoder(mel)
        assert enc1.shape == enc0.shape and torch.allclose(enc0, enc1, atol=1e-5)
        q0, q1 = ref.quantizer(enc0), patched.quantizer(enc0)
        assert torch.equal(q0.codes, q1.codes)
        z0, z1 = ref.quantizer.decode(q0.codes), patched.quantizer.decode(q0.codes)
        assert z1.shape == z0.shape and torch.allclose(z0, z1, atol=1e-5)
        y0, y1 = ref.generator(z0), patched.generator(z0)
        assert y1.shape == y0.shape and torch.allclose(y0, y1, atol=1e-5)
        assert patched.quantizer.encode(enc0).shape == ref.quantizer.encode(enc0).shape
        # decode_from_codes_batch's (B,1,T,1) layout: every clip is decoded (the reference decodes clip 0 only)
        cb = torch.cat([q0.codes, q0.codes.flip(2)], 0)                       # (2,1,T,1)
        zb = patched.quantizer.decode(cb)
        assert zb.shape == (2, 1024, q0.codes.shape[2]) and torch.allclose(zb[0:1], z0, atol=1e-5)


def test_mel_frontend_patch_and_buffers(codecs):
    """patch(..., mel_frontend=True) swaps spec_transform too; mel_buffers() rebuilds the reference's two
    non-persistent buffers (models/mel_spec.py:24,85-98) bit-for-bit."""
    from distilcodec_nabeel_b200 import B200MelSpectrogram, load_config, mel_buffers, patch
    ref, _ = codecs
    bufs = mel_buffers(load_config())
    assert torch.equal(bufs["spec_transform.fb"], ref.spec_transform.fb)
    assert torch.equal(bufs["spec_transform.spectrogram.window"], ref.spec_transform.spectrogram.window)
    sd = state_dict("W1", 1024)
    c = patch(ref_loader.build_reference_codec(sd, codebook_size=1024), device="cpu", mel_frontend=True)
    assert isinstance(c.spec_transform, B200MelSpectrogram)
    assert "spec_transform.fb" in c.encoder._engines.state_dict


def test_bulk_encode_is_result_identical_to_reference_encode(codecs):
    """distilcodec_nabeel_b200.bulk.encode against DistilCodec.encode (distil_codec.py:545-573):
    same GRVQResult, the SAME token dict objects, same per-clip feature tensors, same lengths; features=False only
    empties the two feature lists."""
    from distilcodec_nabeel_b200 import bulk
    ref, patched = codecs
    clips = [[_pcm(1.0), 24000], [_pcm(0.6), 24000], [_pcm(0.3), 24000]]
    with torch.no_grad():
        r0, gen0, hop0 = ref.encode(copy.deepcopy(clips), enable_bfloat16=False, raw_audio=True)
    r1, gen1, hop1 = bulk.encode(patched, copy.deepcopy(clips), enable_bfloat16=False, raw_audio=True)
    assert gen0 == gen1 and hop0 == hop1
    assert torch.equal(r0.codes, r1.codes)
    assert r0.codes_list == r1.codes_list and [len(c) for c in r1.codes_list] == hop1
    table = patched.gr_audio_code2token["g0r0"]["audio_code_token"]
    assert all(tok is table[str(tok["in_codebook_id"])] for tok in r1.codes_list[1])
    assert len(r1.x_pjt_in_list) == len(r1.quantized_fup_list) == 3
    for a, b in zip(r0.x_pjt_in_list + r0.quantized_fup_list, r1.x_pjt_in_list + r1.quantized_fup_list):
        assert a.shape == b.shape and a.dtype == b.dtype and torch.allclose(a, b, atol=1e-5)
    r2, _, _ = bulk.encode(patched, copy.deepcopy(clips), raw_audio=True, features=False)
    assert r2.codes_list == r0.codes_list and r2.x_pjt_in_list == [] and r2.quantized_fup_list == []
    # generic (groups, residual levels) ordering of audio_tokenize: frame-major, then group, then level
    fake = copy.copy(patched)
    fake.gr_audio_code2token = {f"g{g}r{r}": {"audio_code_token": {str(n): (g, r, n) for n in range(5)}}
                                for g in range(2) for r in range(3)}
    codes = np.arange(2 * 4 * 3).reshape(2, 4, 3) % 5
    flat = torch.from_numpy(codes).transpose(1, 0).reshape(4, 6).flatten().tolist()      # as distil_codec.py:563
    want = ref.audio_tokenize.__func__(fake, codes=flat, n_groups=2, n_residual=3)
    assert bulk.tokenize_codes(fake, codes) == want
