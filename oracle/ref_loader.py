"""Import the REAL reference (`distilcodec` package) behind five stub modules.
TEST INFRASTRUCTURE (oracle/): used only by tests/, tests/golden/make_golden.py and bench.py's CPU legs.

Where the reference is looked for, in order: `$DISTILCODEC_REFERENCE`, `/root/reference` (dev container only — it
does not exist on the GPU box), `baseline/_ref/` (the reference `pip install --target`-ed by `__graft_entry__.build()`;
git-ignored, but it travels to the GPU box with the repo snapshot, so the reference-backed tests and the
`--impl reference` arm run the reference's own modules there too).

Stub recipe: SURVEY.md Appendix B (soundfile, librosa, matplotlib, vector_quantize_pytorch, einx.get_at).
"""
from __future__ import annotations

import copy
import os
import sys
import warnings

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_CANDIDATES = [os.environ.get("DISTILCODEC_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]


def _find_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "distilcodec", "distil_codec.py")):
            return c
    return None


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return REFERENCE_ROOT is not None


def import_reference():
    """-> the reference's `distilcodec` package (imported once)."""
    if not available():
        raise RuntimeError(f"reference not found (looked in {[c for c in _CANDIDATES if c]})")
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
