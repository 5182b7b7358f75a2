"""The oracle against the reference itself, imported from /root/reference behind oracle/shims (dev container only;
on the GPU box the reference does not exist and these tests skip — the golden vectors cover that side)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import restatement as R
from tests.conftest import rel_err, state_dict
from tests.golden.inputs import make_mel

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not importable (neither /root/reference nor baseline/_ref)")


@pytest.fixture(scope="module")
def ref_small():
    """Reference DistilCodec with a 1024-entry codebook (fast) and W1 weights."""
    sd = state_dict("W1", 1024)
    return ref_loader.build_reference_codec(sd, codebook_size=1024), sd


def test_state_dict_keys_and_shapes_match_reference(ref_small):
    codec, sd = ref_small
    ref_sd = {k: v for k, v in codec.state_dict().items() if not k.startswith("spec_transform")}
    assert set(ref_sd) == set(sd)
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), k


def test_stage_by_stage(ref_small):
    codec, sd = ref_small
    mel = make_mel(2, 24, seed=21)
    ref = ref_loader.run_reference(codec, mel)
    out = R.codec_forward(sd, mel)
    assert rel_err(out["enc"], ref["enc"]) < 1e-6
    assert rel_err(out["x_pjt_in"], ref["x_pjt_in"]) < 1e-6
    assert torch.equal(out["codes"], ref["codes"])
    assert rel_err(out["quantized_fup"], ref["quantized_fup"]) == 0.0
    assert rel_err(out["quantized"], ref["quantized"]) < 1e-6
    assert rel_err(out["wav"], ref["wav"]) < 1e-5
    assert rel_err(R.quantizer_decode(sd, ref["codes"]), ref["z_dec"]) < 1e-6


def test_decode_from_codes_batch_layout_bug_is_real(ref_small):
    """the reference's quantizer.decode consumes only indices[0] -> (B,1,T,1) decodes clip 0 only."""
    codec, _ = ref_small
    codes = torch.randint(0, 1024, (3, 1, 10, 1), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        z = codec.quantizer.decode(codes)
        z0 = codec.quantizer.decode(codes[:1])
    assert z.shape == (1, 1024, 10) and torch.equal(z, z0)
