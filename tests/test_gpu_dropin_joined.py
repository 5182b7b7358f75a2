"""The joined drop-in path on hardware: a REAL reference `DistilCodec` (the reference's own package, imported from
baseline/_ref on the GPU box or /root/reference in the dev container) whose three hot-path attributes are replaced by
`patch()` and backed by libdistilcodec_b200.so, driven through the reference's unchanged public methods

    DistilCodec.encode(raw_audio=True, enable_bfloat16=...)     distil_codec.py:545-573
    DistilCodec.decode_from_codes(...)                          distil_codec.py:581-594
    DistilCodec.decode_from_codes_batch(...)                    distil_codec.py:598-639

against the UNPATCHED reference codec with the same weights (fp32, TF32 off, on the same device)."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_loader
from tests.conftest import golden, rel_err, state_dict

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="reference package not importable "
                                 "(neither /root/reference nor baseline/_ref)")]


def _pcm(seconds):
    return golden("audio_W0.npz")["pcm"][:int(seconds * 24000)].copy()


@pytest.fixture(scope="module")
def codecs():
    from distilcodec_nabeel_b200 import patch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = state_dict("W1")
    dev = torch.device("cuda", 0)
    ref = ref_loader.build_reference_codec(sd)
    ref.device = dev                       # what from_pretrained does (distil_codec.py:84, 95)
    ref.move_to_cuda()
    patched = ref_loader.build_reference_codec(sd)
    patched.device = dev
    patch(patched, device=dev, mel_frontend=False)
    yield ref, patched
    patched.encoder._engines.invalidate()


def test_encode_on_the_cuda_library_matches_the_reference(codecs):
    from distilcodec_nabeel_b200 import B200Encoder
    ref, patched = codecs
    assert isinstance(patched.encoder, B200Encoder) and type(patched) is type(ref)
    clips = [[_pcm(2.0), 24000], [_pcm(1.3), 24000]]           # ragged: the reference pads to the longest
    n0 = patched.encoder._engines.get("fp32").launch_count()
    with torch.no_grad():
        r0, gen0, hop0 = ref.encode(copy.deepcopy(clips), enable_bfloat16=False, raw_audio=True)
        r1, gen1, hop1 = patched.encode(copy.deepcopy(clips), enable_bfloat16=False, raw_audio=True)
    assert patched.encoder._engines.get("fp32").launch_count() > n0      # the CUDA library did the work
    assert gen0 == gen1 and hop0 == hop1
    assert r1.codes.shape == r0.codes.shape and r1.codes.dtype == torch.int64
    agree = float((r0.codes == r1.codes).float().mean())
    assert agree >= 0.999, agree                                          # fp32 mode, W1: identical codes
    assert rel_err(r1.x_pjt_in.float(), r0.x_pjt_in.float()) < 1e-4
    assert rel_err(r1.quantized, r0.quantized) < 1e-4 or agree < 1.0
    assert [len(c) for c in r1.codes_list] == hop1
    if agree == 1.0:
        assert r0.codes_list == r1.codes_list
    for a, b in zip(r0.x_pjt_in_list, r1.x_pjt_in_list):
        assert a.shape == b.shape and rel_err(b.float(), a.float()) < 1e-4
    # enable_bfloat16=True: the tcgen05 engine under the reference's own autocast context
    with torch.no_grad():
        r2, gen2, hop2 = patched.encode(copy.deepcopy(clips), enable_bfloat16=True, raw_audio=True)
    assert patched.encoder._engines.get("bf16").launch_count() > 0
    assert gen2 == gen0 and hop2 == hop0
    agree_bf16 = float((r0.codes == r2.codes).float().mean())
    assert agree_bf16 >= 0.95, agree_bf16                                  # upstream bf16 rounding may flip near-ties
    assert rel_err(r2.x_pjt_in.float(), r0.x_pjt_in.float()) < 2e-2


@pytest.mark.parametrize("bf16", [False, True])
def test_decode_from_codes_on_the_cuda_library_matches_the_reference(codecs, bf16):
    ref, patched = codecs
    with torch.no_grad():
        r0, _, hop = ref.encode([[_pcm(1.5), 24000]], enable_bfloat16=False, raw_audio=True)
        codes = r0.codes[0, 0, :hop[0], 0].tolist()
        y0 = ref.decode_from_codes(list(codes), minus_token_offset=False, enable_bfloat16=False)
        y1 = patched.decode_from_codes(list(codes), minus_token_offset=False, enable_bfloat16=bf16)
    assert tuple(y1.shape) == tuple(y0.shape) == (1, 1, 256 * hop[0])
    assert rel_err(y1.float(), y0.float()) < (1e-2 if bf16 else 1e-4)


def test_decode_from_codes_batch_decodes_every_clip(codecs):
    """Deliberate deviation: the reference's batch method passes (B,1,T,1) and therefore decodes clip
    0 only; the shim decodes all B.  Parity target = per-clip decode_from_codes of the RIGHT-PADDED code sequence."""
    ref, patched = codecs
    g = torch.Generator().manual_seed(3)
    lens = [70, 41, 57]
    seqs = [torch.randint(0, 32768, (n,), generator=g).tolist() for n in lens]
    with torch.no_grad():
        outs = patched.decode_from_codes_batch(copy.deepcopy(seqs), minus_token_offset=False, enable_bfloat16=False)
        assert len(outs) == 3
        for i, n in enumerate(lens):
            padded = seqs[i] + [0] * (max(lens) - n)
            y_ref = ref.decode_from_codes(padded, minus_token_offset=False, enable_bfloat16=False)
            assert tuple(outs[i].shape) == tuple(y_ref.shape) == (1, 1, 256 * max(lens))
            assert rel_err(outs[i].float(), y_ref.float()) < 1e-4, i


def test_bf16_encoder_error_on_the_stress_weights_is_the_precision_floor(codecs):
    """tests/test_gpu_e2e.py gates the W1 encoder at 2e-2 in bf16 mode (1e-2 is defined on W0).  The reason is the
    precision, not this implementation: the reference's OWN encoder under its own bf16 autocast (distil_codec.py:560-563)
    on the same GPU is as far from its fp32 result.  Gate: this path's error stays within 1.25x of the reference's."""
    ref, patched = codecs
    mel = torch.from_numpy(golden("e2e_W1.npz")["mel"]).to(ref.device)
    with torch.no_grad():
        y32 = ref.encoder(mel).float()
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            y_ref16 = ref.encoder(mel).float()
            y_b200 = patched.encoder(mel).float()
    assert patched.encoder._engines.get("bf16").launch_count() > 0
    e_ref, e_b200 = rel_err(y_ref16, y32), rel_err(y_b200, y32)
    print(f"W1 encoder, bf16: reference autocast {e_ref:.3e}, this path {e_b200:.3e}")
    assert e_ref > 2e-3                       # the stress weights do expose the bf16 rounding
    assert e_b200 < max(1e-2, 1.25 * e_ref), (e_b200, e_ref)
