"""Drop-in replacements for the three hot-path attributes of the reference's `DistilCodec`
(distilcodec/distil_codec.py:52-54): `encoder`, `quantizer`, `generator`.

They mirror the reference modules' interface (same call signatures, argument meaning, result fields and
state_dict keys) and run every FLOP in libdistilcodec_b200.so:

    ConvNeXtEncoder.forward      models/encoders.py:68-76            -> B200Encoder.forward
    DownsampleGRVQ.forward       vector_quantization/grfvq.py:105    -> B200Quantizer.forward  (returns GRVQResult)
    DownsampleGRVQ.decode/encode vector_quantization/grfvq.py:134-146-> B200Quantizer.decode / .encode
    HiFiGANGenerator.forward     models/generators.py:118-147        -> B200Generator.forward

`patch(codec)` installs them on an existing reference `DistilCodec` instance; `distil_codec.py` stays unchanged.
Numeric mode follows the caller exactly like the reference does: inside `torch.autocast('cuda', bfloat16)`
(what `enable_bfloat16=True` sets up, distil_codec.py:550,590) the bf16 tensor-core engine runs, otherwise the
fp32 engine.  Engines are created lazily per mode from the same state_dict.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn

from .engine import Engine, load_config


@dataclass
class GRVQResult:
    """Field-for-field mirror of vector_quantization/grfvq.py:13-24 (same names, same order).  The three `*_list`
    fields start empty; DistilCodec.encode appends the per-clip tokens / features to them (distil_codec.py:563-570)."""
    quantized: torch.Tensor
    codes: torch.Tensor
    codes_list: list
    total_loss: torch.Tensor
    commitment_loss: torch.Tensor
    codebook_diversity_loss: torch.Tensor
    quantized_fup: torch.Tensor
    quantized_fup_list: list
    x_pjt_in: torch.Tensor
    x_pjt_in_list: list


class EngineSet:
    """Lazily built engines (one per numeric mode) over one flat state_dict with `encoder.`/`quantizer.`/
    `generator.` prefixes.  Shared by the three shim modules of one codec."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, config: Optional[dict] = None,
                 force_mode: Optional[str] = None, **engine_kwargs):
        self.state_dict = state_dict
        self.device = torch.device(device)
        self.config = config or load_config()
        self.force_mode = force_mode
        self.engine_kwargs = engine_kwargs
        self._engines: Dict[str, Engine] = {}

    def mode_now(self) -> str:
        if self.force_mode:
            return self.force_mode
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return "bf16"
        return "fp32"

    def get(self, mode: Optional[str] = None) -> Engine:
        mode = mode or self.mode_now()
        if mode not in self._engines:
            self._engines[mode] = Engine(self.state_dict, self.device, mode, self.config, **self.engine_kwargs)
        return self._engines[mode]

    def invalidate(self):
        for e in self._engines.values():
            e.close()
        self._engines.clear()


class _Shim(nn.Module):
    """Common state_dict plumbing: the shim exposes exactly the reference module's keys."""
    prefix = ""

    def __init__(self, engines: EngineSet):
        super().__init__()
        self._engines = engines

    def state_dict(self, *args, **kwargs):  # same keys as the reference module's state_dict()
        p = self.prefix
        return OrderedDict((k[len(p):], v) for k, v in self._engines.state_dict.items() if k.startswith(p))

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        p = self.prefix
        mine = {k for k in self._engines.state_dict if k.startswith(p)}
        missing = [k[len(p):] for k in mine if k[len(p):] not in state_dict]
        unexpected = [k for k in state_dict if p + k not in mine]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:4]}, unexpected {unexpected[:4]}")
        for k, v in state_dict.items():
            if p + k in mine:
                self._engines.state_dict[p + k] = v.detach().clone()
        self._engines.invalidate()
        return nn.modules.module._IncompatibleKeys(missing, unexpected)

    def to(self, *args, **kwargs):  # weights are staged by the engine; `.to(device)` only re-targets it
        device = kwargs.get("device", args[0] if args else None)
        if isinstance(device, (str, torch.device, int)) and device is not None:
            dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
            if dev.type != "cuda":
                raise RuntimeError("the B200 hot path has no CPU implementation")
            if dev != self._engines.device:
                self._engines.device = dev
                self._engines.invalidate()
        return self

    def _dev(self, x: torch.Tensor) -> torch.Tensor:
        eng_dev = self._engines.device
        if eng_dev.type == "cuda" and eng_dev.index is None:
            eng_dev = torch.device("cuda", torch.cuda.current_device())
            self._engines.device = eng_dev
        if x.device != eng_dev:
            raise RuntimeError(f"input on {x.device} but the codec was moved to {eng_dev}")
        return x


class B200Encoder(_Shim):
    """models/encoders.py ConvNeXtEncoder: forward(x: f32[B,128,T]) -> f32[B,1024,T]."""
    prefix = "encoder."

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._dev(x).float().contiguous()
        enc = self._engines.get().encoder(x)          # (B, T, 1024)
        return enc.transpose(1, 2)                    # (B, 1024, T) view, like the reference's layout


class _CodebookView:
    """`quantizer.grvq.codebooks` (residual_vq.py:289-291): tensor (G=1, R=1, K, D), read at distil_codec.py:201,268."""

    def __init__(self, engines: EngineSet):
        self._engines = engines

    @property
    def codebooks(self) -> torch.Tensor:
        e = self._engines.state_dict["quantizer.grvq.rvqs.0.layers.0._codebook.embed"]
        return e.reshape(1, 1, *e.shape[-2:])


class B200Quantizer(_Shim):
    """vector_quantization/grfvq.py DownsampleGRVQ (groups = 1, one codebook, downsample_factor [1])."""
    prefix = "quantizer."

    def __init__(self, engines: EngineSet):
        super().__init__(engines)
        self.downsample_factor = [1]
        self.grvq = _CodebookView(engines)

    @torch.no_grad()
    def forward(self, z: torch.Tensor) -> GRVQResult:
        eng = self._engines.get()
        z_nlc = eng.ncl_to_nlc(self._dev(z).float())
        codes, xin, fup, quant = eng.quantizer(z_nlc, want_fup=True)
        zero = torch.zeros((), dtype=torch.float32, device=z.device)  # eval-mode losses are 0 (vq :974,:1056)
        B, T = codes.shape
        return GRVQResult(quantized=quant.transpose(1, 2), codes=codes.view(1, B, T, 1), codes_list=[],
                          total_loss=zero, commitment_loss=zero, codebook_diversity_loss=zero, quantized_fup=fup,
                          quantized_fup_list=[], x_pjt_in=xin, x_pjt_in_list=[])

    @torch.no_grad()
    def encode(self, z: torch.Tensor) -> torch.Tensor:
        """grfvq.py:134-139: (B,1024,T) -> codes (B, G*R = 1, T)."""
        eng = self._engines.get()
        return eng.quantizer_encode(eng.ncl_to_nlc(self._dev(z).float())).unsqueeze(1)

    @torch.no_grad()
    def decode(self, indices: torch.Tensor) -> torch.Tensor:
        """grfvq.py:141-146: indices (G=1, B, T, R=1) int64 -> (B, 1024, T).

        Deliberate deviation: `decode_from_codes_batch` (distil_codec.py:598-639) passes
        (B, 1, T, 1); the reference then silently decodes clip 0 only.  With one group and one codebook that layout
        is unambiguous whenever shape[1] == 1, so it is treated as a batch of B clips here."""
        if indices.dim() != 4 or indices.shape[-1] != 1:
            raise RuntimeError(f"indices must be (1, B, T, 1), got {tuple(indices.shape)}")
        if indices.shape[0] == 1:
            idx = indices[0, :, :, 0]
        elif indices.shape[1] == 1:
            idx = indices[:, 0, :, 0]
        else:
            raise RuntimeError("only one codebook group is configured (n_groups = 1)")
        eng = self._engines.get()
        z = eng.decode_codes(self._dev(idx).long().contiguous())
        return z.transpose(1, 2)


class B200Generator(_Shim):
    """models/generators.py HiFiGANGenerator: forward(x[B,1024,T], template=None, is_debug=False) -> [B,1,256T]."""
    prefix = "generator."

    @torch.no_grad()
    def forward(self, x: torch.Tensor, template=None, is_debug: bool = False) -> torch.Tensor:
        if template is not None:
            raise RuntimeError("use_template is false in configs/model_config.json; templates are not supported")
        eng = self._engines.get()
        wav = eng.generator(eng.ncl_to_nlc(self._dev(x).float()))
        return wav.unsqueeze(1)

    def remove_parametrizations(self):  # generators.py:149-155 — weight_norm is already folded at load time
        return None


def mel_buffers(cfg: dict) -> Dict[str, torch.Tensor]:
    """The two non-persistent buffers of the reference's LogMelSpectrogram (models/mel_spec.py:24,85-98), built the
    way the reference builds them, under the names the C ABI expects."""
    import torchaudio.functional as AF
    sc = cfg["spec_transform"]
    fb = AF.melscale_fbanks(n_freqs=sc["n_fft"] // 2 + 1, f_min=float(sc["fmin"]),
                            f_max=float(sc["fmax"] or sc["sampling_rate"] // 2), n_mels=sc["num_mels"],
                            sample_rate=sc["sampling_rate"], norm="slaney", mel_scale="slaney")
    return {"spec_transform.fb": fb, "spec_transform.spectrogram.window": torch.hann_window(sc["win_size"])}


class B200MelSpectrogram(_Shim):
    """models/mel_spec.py LogMelSpectrogram: forward(x[B,1,Ls] or [B,Ls]) -> log-mel [B,128,T], on the GPU (the
    reference forces the STFT onto the CPU, mel_spec.py:39).  SURVEY.md section 8f, row f-1."""
    prefix = "spec_transform."

    @torch.no_grad()
    def forward(self, x: torch.Tensor, return_linear: bool = False, sample_rate: Optional[int] = None):
        if return_linear:
            raise RuntimeError("return_linear is not on the inference path (only the log-mel is produced)")
        sr = self._engines.config["spec_transform"]["sampling_rate"]
        if sample_rate is not None and sample_rate != sr:
            raise RuntimeError(f"resampling is not part of the hot path; pass {sr} Hz audio")
        eng = self._engines.get()
        x = x.to(eng.device, non_blocking=True).float()
        return eng.mel(x.squeeze(1).contiguous() if x.dim() == 3 else x.contiguous())


def build_modules(state_dict: Dict[str, torch.Tensor], device="cuda", config: Optional[dict] = None,
                  force_mode: Optional[str] = None, **engine_kwargs):
    """-> (encoder, quantizer, generator) shims sharing one EngineSet."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None and torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
    es = EngineSet(dict(state_dict), dev, config, force_mode, **engine_kwargs)
    return B200Encoder(es), B200Quantizer(es), B200Generator(es)


def patch(codec, device=None, force_mode: Optional[str] = None, mel_frontend: bool = False, **engine_kwargs):
    """Install the B200 hot path on a reference `DistilCodec` instance (after construction / `from_pretrained`):
    replaces `codec.encoder`, `codec.quantizer`, `codec.generator` (distil_codec.py:52-54) and leaves everything
    else — the class, its methods, the CPU mel front-end — untouched.  `mel_frontend=True` additionally replaces
    `codec.spec_transform` (distil_codec.py:56-63) by the GPU log-mel kernel.  Returns the codec."""
    sd = OrderedDict()
    for name in ("encoder", "quantizer", "generator"):
        mod = getattr(codec, name, None)
        if mod is None:
            continue
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach()
    if device is None:
        device = getattr(codec, "device", None) or "cuda"
    cfg = getattr(codec, "codec_config", None)  # the dict DistilCodec was constructed from (distil_codec.py:41)
    st = getattr(codec, "spec_transform", None)
    if mel_frontend and st is not None:  # the module's own (non-persistent) buffers
        sd["spec_transform.fb"] = st.fb.detach()
        sd["spec_transform.spectrogram.window"] = st.spectrogram.window.detach()
    enc, q, gen = build_modules(sd, device, cfg, force_mode, **engine_kwargs)
    if mel_frontend and st is not None:
        codec.spec_transform = B200MelSpectrogram(enc._engines)
    codec.encoder = enc
    codec.quantizer = q
    if getattr(codec, "generator", None) is not None:
        codec.generator = gen
    return codec
