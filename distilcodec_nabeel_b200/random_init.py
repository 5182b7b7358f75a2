"""Deterministic random-init weight sets of the reference architecture (inputs for the benchmark, smoke and tests;
not on the compute path: the product loads whatever state_dict the caller hands to `patch()` / `Engine`).

The HF checkpoint is not available offline, so parity is evaluated on random-init weights of the
architecture in configs/model_config.json.  The reference's own constructors draw from torch's CPU RNG, whose
`normal_` stream depends on the host ISA dispatch, and the reference sources are not present on the GPU box.  To get
bit-identical weights on every machine we therefore draw from numpy's counter-based Philox generator and
follow the reference's *init distributions* and its exact `state_dict()` key set / shapes:

  * encoder  Conv1d/Linear ~ N(0, 0.02), bias 0, LN (1, 0), gamma 1e-6
        (distilcodec/models/encoders.py:63-66, convnext_utils.py:196-197, 256-260)
  * quantizer Conv1d/Linear ~ N(0, 0.02), bias 0 (vector_quantization/grfvq.py:100-103); ConvTranspose1d keeps
        torch's default U(+-1/sqrt(fan_in)); codebook ~ kaiming_uniform on (1, 32768, 3584) = U(+-sqrt(6/(32768*3584)))
        (vector_quantization/utils/vector_quantize_pytorch.py:71-74, 293, 327)
  * generator: every conv is weight_norm-parametrised (models/generators.py:50, 70, 106; convnext_utils.py:36-102)
        so `init_weights` (applied after parametrisation, generators.py:115-116, convnext_utils.py:67, 104) is a
        no-op on original0/original1 and the convs keep torch's default U(+-1/sqrt(fan_in)); g = ||v|| per dim-0 slice.

Two sets:
  W0  "init"   — the distributions above (the BASELINE gate; numerically almost blind end to end).
  W1  "stress" — same keys; LayerScale gamma, biases, LN affine, weight-norm g and the codebook re-drawn O(1) so
                 every term of every kernel is numerically visible and the codes steer the waveform.
The same state_dict is loaded into the reference modules (oracle/ref_loader.py, tests only) and into the CUDA path.
"""
from __future__ import annotations

import json
import math
import os
from collections import OrderedDict

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CONFIG = os.path.join(_HERE, "model_config.json")


def load_config(path: str | None = None) -> dict:
    with open(path or DEFAULT_CONFIG) as f:
        return json.load(f)


# ----------------------------------------------------------------------------------------------------------
# manifest: (key, shape, kind, meta) in the reference's state_dict order
# ----------------------------------------------------------------------------------------------------------

def _convnext_block(prefix: str, dim: int):
    """Keys of ConvNeXtBlock (models/convnext_utils.py:232-261) in registration order."""
    return [
        (prefix + "gamma", (dim,), "gamma", {}),
        (prefix + "dwconv.weight", (dim, 1, 7), "normal02", {"fan_in": 7}),
        (prefix + "dwconv.bias", (dim,), "zero_bias", {"fan_in": 7}),
        (prefix + "norm.weight", (dim,), "ln_w", {}),
        (prefix + "norm.bias", (dim,), "ln_b", {}),
        (prefix + "pwconv1.weight", (4 * dim, dim), "normal02", {"fan_in": dim}),
        (prefix + "pwconv1.bias", (4 * dim,), "zero_bias", {"fan_in": dim}),
        (prefix + "pwconv2.weight", (dim, 4 * dim), "normal02", {"fan_in": 4 * dim}),
        (prefix + "pwconv2.bias", (dim,), "zero_bias", {"fan_in": 4 * dim}),
    ]


def _wn_conv(prefix: str, shape: tuple, bias_len: int, fan_in: int):
    """Keys of a weight_norm-parametrised conv: bias, original0 (g), original1 (v)."""
    return [
        (prefix + "bias", (bias_len,), "default_bias", {"fan_in": fan_in}),
        (prefix + "parametrizations.weight.original0", (shape[0], 1, 1), "wn_g", {"fan_in": fan_in}),
        (prefix + "parametrizations.weight.original1", shape, "default_w", {"fan_in": fan_in}),
    ]


def manifest(cfg: dict):
    enc, dec, q = cfg["encoder"], cfg["decoder"], cfg["quantizer"]
    dims, depths, ks = enc["dims"], enc["depths"], enc["kernel_size"]
    m = []
    # ---- encoder (models/encoders.py:8-61) ----
    p = "encoder."
    cin = enc["input_channels"]
    m += [(p + "downsample_layers.0.0.weight", (dims[0], cin, ks), "normal02", {"fan_in": cin * ks}),
          (p + "downsample_layers.0.0.bias", (dims[0],), "zero_bias", {"fan_in": cin * ks}),
          (p + "downsample_layers.0.1.weight", (dims[0],), "ln_w", {}),
          (p + "downsample_layers.0.1.bias", (dims[0],), "ln_b", {})]
    for i in range(len(dims) - 1):
        m += [(p + f"downsample_layers.{i+1}.0.weight", (dims[i],), "ln_w", {}),
              (p + f"downsample_layers.{i+1}.0.bias", (dims[i],), "ln_b", {}),
              (p + f"downsample_layers.{i+1}.1.weight", (dims[i + 1], dims[i], 1), "normal02", {"fan_in": dims[i]}),
              (p + f"downsample_layers.{i+1}.1.bias", (dims[i + 1],), "zero_bias", {"fan_in": dims[i]})]
    for s, (d, n) in enumerate(zip(dims, depths)):
        for j in range(n):
            m += _convnext_block(p + f"stages.{s}.{j}.", d)
    m += [(p + "norm.weight", (dims[-1],), "ln_w", {}), (p + "norm.bias", (dims[-1],), "ln_b", {})]
    # ---- quantizer (vector_quantization/grfvq.py:28-98) ----
    p = "quantizer."
    D, cd, cs = q["input_dim"], q["codebook_dim"], q["codebook_size"]
    r = p + "grvq.rvqs.0."
    m += [(r + "project_in.weight", (cd, D), "normal02", {"fan_in": D}),
          (r + "project_in.bias", (cd,), "zero_bias", {"fan_in": D}),
          (r + "project_out.weight", (D, cd), "normal02", {"fan_in": cd}),
          (r + "project_out.bias", (D,), "zero_bias", {"fan_in": cd}),
          (r + "layers.0._codebook.initted", (1,), "one", {}),
          (r + "layers.0._codebook.cluster_size", (1, cs), "one", {}),
          (r + "layers.0._codebook.embed_avg", (1, cs, cd), "codebook_copy", {}),
          (r + "layers.0._codebook.embed", (1, cs, cd), "codebook", {})]
    m += [(p + "downsample.0.0.weight", (D, D, 1), "normal02", {"fan_in": D}),
          (p + "downsample.0.0.bias", (D,), "zero_bias", {"fan_in": D})]
    m += _convnext_block(p + "downsample.0.1.", D)
    m += [(p + "upsample.0.0.weight", (D, D, 1), "default_w", {"fan_in": D}),
          (p + "upsample.0.0.bias", (D,), "default_bias", {"fan_in": D})]
    m += _convnext_block(p + "upsample.0.1.", D)
    # ---- generator (models/generators.py:29-116) ----
    p = "generator."
    C0, kpre, kpost = dec["upsample_initial_channel"], dec["pre_conv_kernel_size"], dec["post_conv_kernel_size"]
    m += _wn_conv(p + "conv_pre.", (C0, dec["num_mels"], kpre), C0, dec["num_mels"] * kpre)
    ch = C0
    for i, (u, k) in enumerate(zip(dec["upsample_rates"], dec["upsample_kernel_sizes"])):
        # ConvTranspose1d weight is (in, out, k); torch's fan_in for it is size(1) * k
        m += _wn_conv(p + f"ups.{i}.", (ch, ch // 2, k), ch // 2, (ch // 2) * k)
        ch //= 2
    ch = C0
    for i in range(len(dec["upsample_rates"])):
        ch //= 2
        for b, k in enumerate(dec["resblock_kernel_sizes"]):
            for grp in ("convs1", "convs2"):
                for n in range(3):
                    m += _wn_conv(p + f"resblocks.{i}.blocks.{b}.{grp}.{n}.", (ch, ch, k), ch, ch * k)
    m += _wn_conv(p + "conv_post.", (1, ch, kpost), 1, ch * kpost)
    return m


# ----------------------------------------------------------------------------------------------------------
# generation
# ----------------------------------------------------------------------------------------------------------

def _u(rng, shape, bound):
    # exact arithmetic only: uint32 -> float in [0,1) -> affine map; bit-reproducible everywhere
    x = rng.random(size=shape, dtype=np.float32)
    return (x * np.float32(2.0) - np.float32(1.0)) * np.float32(bound)


def _n(rng, shape, std):
    return rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)


def make_state_dict(variant: str = "W0", seed: int = 1234, cfg: dict | None = None,
                    include_codebook: bool = True, codebook_size: int | None = None) -> "OrderedDict[str, torch.Tensor]":
    """Return an OrderedDict keyed exactly like `DistilCodec(cfg).state_dict()` restricted to
    encoder.* / quantizer.* / generator.*  (spec_transform has no state).

    `codebook_size` lets tests shrink the 32768-entry codebook (the reference modules are constructed with the
    same shrunken config); None = the configured size."""
    assert variant in ("W0", "W1")
    cfg = cfg or load_config()
    if codebook_size is not None:
        cfg = json.loads(json.dumps(cfg))
        cfg["quantizer"]["codebook_size"] = codebook_size
    rng = np.random.Generator(np.random.Philox(key=seed + (0 if variant == "W0" else 7919)))
    stress = variant == "W1"
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    v_cache = {}
    for key, shape, kind, meta in manifest(cfg):
        fan_in = meta.get("fan_in", 1)
        if kind == "normal02":
            a = _n(rng, shape, 0.02)
            if stress and key.endswith("dwconv.weight"):
                a = _n(rng, shape, 0.3)
        elif kind == "zero_bias":
            a = _u(rng, shape, 0.1) if stress else np.zeros(shape, np.float32)
        elif kind == "ln_w":
            a = (np.float32(1.0) + _u(rng, shape, 0.5)) if stress else np.ones(shape, np.float32)
        elif kind == "ln_b":
            a = _u(rng, shape, 0.2) if stress else np.zeros(shape, np.float32)
        elif kind == "gamma":
            a = (np.float32(0.6) + _u(rng, shape, 0.3)) if stress else np.full(shape, 1e-6, np.float32)
        elif kind == "one":
            a = np.ones(shape, np.float32)
        elif kind == "default_w":
            a = _u(rng, shape, 1.0 / math.sqrt(fan_in))
        elif kind == "default_bias":
            a = _u(rng, shape, (0.1 if stress else 1.0 / math.sqrt(fan_in)))
        elif kind == "wn_g":
            a = None  # filled after original1 is drawn (g = ||v||, optionally rescaled for W1)
        elif kind == "codebook":
            if not include_codebook:
                continue
            # W1: codebook at the scale of x = project_in(z) (sigma ~ 0.43)
            a = _n(rng, shape, 0.43) if stress else _u(rng, shape, math.sqrt(6.0 / (shape[1] * shape[2])))
        elif kind == "codebook_copy":
            if not include_codebook:
                continue
            a = None
        else:
            raise KeyError(kind)
        sd[key] = None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        v_cache[key] = (kind, meta)
    # dependent entries
    for key in list(sd.keys()):
        kind, meta = v_cache[key]
        if kind == "wn_g":
            v = sd[key.replace("original0", "original1")]
            g = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            if stress:
                # make every conv roughly variance-preserving so the waveform is O(0.1..1) instead of 0.009:
                # U(+-b) has std b/sqrt(3); default b = 1/sqrt(fan_in) -> layer gain 1/sqrt(3); rescale to ~1.3
                g = g * 2.25
                if "conv_post" in key:
                    g = g * 0.5
            sd[key] = g.contiguous()
        elif kind == "codebook_copy":
            sd[key] = sd[key.replace("embed_avg", "embed")].clone()
    return sd


def split_state_dict(sd):
    """-> dict(encoder=..., quantizer=..., generator=...) with the prefixes stripped (the layout of the
    reference checkpoint, distilcodec/distil_codec.py:91-94)."""
    out = {"encoder": OrderedDict(), "quantizer": OrderedDict(), "generator": OrderedDict()}
    for k, v in sd.items():
        top, rest = k.split(".", 1)
        out[top][rest] = v
    return out


def checksum(sd) -> str:
    import hashlib
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
