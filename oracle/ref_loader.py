"""Import the REAL reference (`/root/reference/distilcodec`) on CPU behind five stub modules.
TEST INFRASTRUCTURE (oracle/): used only by tests/ and tests/golden/make_golden.py, and only in the dev
container — `/root/reference` does not exist on the GPU box, where `available()` is False.

Stub recipe: SURVEY.md Appendix B (soundfile, librosa, matplotlib, vector_quantize_pytorch, einx.get_at).
"""
from __future__ import annotations

import copy
import os
import sys
import warnings

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("DISTILCODEC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "distilcodec", "distil_codec.py"))


def import_reference():
    """-> the reference's `distilcodec` package (imported once)."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    shims = os.path.join(_HERE, "shims")
    for p in (REFERENCE_ROOT, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import distilcodec  # noqa: F401  (the reference package)
    return distilcodec


def build_reference_codec(state_dict=None, cfg=None, codebook_size=None):
    """Construct the reference `DistilCodec` (distilcodec/distil_codec.py:29-70) in eval mode on CPU and load
    `state_dict` (keys `encoder.* / quantizer.* / generator.*`, oracle/weights.py) into it."""
    from .weights import load_config
    ref = import_reference()
    cfg = copy.deepcopy(cfg or load_config())
    if codebook_size is not None:
        cfg["quantizer"]["codebook_size"] = codebook_size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        codec = ref.DistilCodec(cfg).eval()
    if state_dict is not None:
        missing, unexpected = codec.load_state_dict(state_dict, strict=False)
        missing = [k for k in missing if not k.startswith("spec_transform")]
        assert not missing and not unexpected, (missing[:5], unexpected[:5])
    return codec


@torch.no_grad()
def run_reference(codec, mel: torch.Tensor):
    """mel (B,128,T) fp32 -> dict of stage-boundary tensors, calling the reference modules exactly as
    DistilCodec.forward does (distil_codec.py:518-530) plus the decode path (:591-592)."""
    enc = codec.encoder(mel)
    r = codec.quantizer(enc)
    wav = codec.generator(r.quantized)
    z_dec = codec.quantizer.decode(r.codes)
    return {"enc": enc, "quantized": r.quantized, "codes": r.codes, "quantized_fup": r.quantized_fup,
            "x_pjt_in": r.x_pjt_in, "wav": wav, "z_dec": z_dec,
            "total_loss": r.total_loss, "commitment_loss": r.commitment_loss,
            "codebook_diversity_loss": r.codebook_diversity_loss}
