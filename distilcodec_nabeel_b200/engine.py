"""Host side of the B200 hot path: one `Engine` = one C-ABI handle (one device, one numeric mode).

PyTorch is used for device memory (the caching allocator owns every input, output and workspace buffer) and for the
current CUDA stream; all arithmetic happens in libdistilcodec_b200.so.  Activations cross the C ABI channels-last
(B, T, C); the reference's channels-first tensors are returned as strided views of the same memory.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Optional

import torch

from . import _abi

DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_config.json")
_MODES = {"fp32": _abi.MODE_FP32, "bf16": _abi.MODE_BF16}


def load_config(path: Optional[str] = None) -> dict:
    with open(path or DEFAULT_CONFIG) as f:
        return json.load(f)


def _dc_config(cfg: dict, codebook_size: Optional[int] = None) -> _abi.DcConfig:
    """configs/model_config.json sections -> dc_config (only what the hot path depends on)."""
    enc, dec, q = cfg["encoder"], cfg["decoder"], cfg["quantizer"]
    c = _abi.DcConfig()
    c.n_mels = enc["input_channels"]
    assert len(enc["depths"]) == 4 and len(enc["dims"]) == 4, "the encoder has four stages"
    for i in range(4):
        c.enc_depths[i] = enc["depths"][i]
        c.enc_dims[i] = enc["dims"][i]
    c.codebook_size = codebook_size or q["codebook_size"]
    c.codebook_dim = q["codebook_dim"]
    rates, ks = dec["upsample_rates"], dec["upsample_kernel_sizes"]
    c.n_ups = len(rates)
    for i, (r, k) in enumerate(zip(rates, ks)):
        c.up_rates[i] = r
        c.up_kernels[i] = k
    c.up_initial_channel = dec["upsample_initial_channel"]
    rk = dec["resblock_kernel_sizes"]
    rd = dec["resblock_dilation_sizes"][0]
    assert len(rk) == 3 and len(rd) == 3 and all(list(d) == list(rd) for d in dec["resblock_dilation_sizes"])
    for i in range(3):
        c.rb_kernels[i] = rk[i]
        c.rb_dilations[i] = rd[i]
    c.pre_kernel = dec["pre_conv_kernel_size"]
    c.post_kernel = dec["post_conv_kernel_size"]
    return c


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """Owns a dc_handle.  `state_dict` uses the reference's keys prefixed `encoder.` / `quantizer.` / `generator.`
    (any subset of the three modules); tensors may live on any device, they are staged to `device` as fp32."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: "torch.device | int | str" = 0,
                 mode: str = "bf16", config: Optional[dict] = None, workspace_limit_bytes: int = 24 << 30):
        if not torch.cuda.is_available():
            raise RuntimeError("distilcodec_nabeel_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _abi.load()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mode = mode
        self.cfg = config or load_config()
        self.hop = 1
        for r in self.cfg["decoder"]["upsample_rates"]:
            self.hop *= r
        self.latent_dim = self.cfg["encoder"]["dims"][-1]
        self.code_dim = self.cfg["quantizer"]["codebook_dim"]
        self.workspace_limit = workspace_limit_bytes
        self._ws: Optional[torch.Tensor] = None
        self._keep = []  # tensors the library references in place (the fp32 codebook)
        emb = [v for k, v in state_dict.items() if k.endswith("_codebook.embed")]
        dcc = _dc_config(self.cfg, emb[0].shape[1] if emb else None)
        h = C.c_void_p()
        _abi.check(self.lib.dc_create(self.device.index, _MODES[mode], C.byref(dcc), C.byref(h)), "dc_create")
        self.h = h
        self.codebook: Optional[torch.Tensor] = None
        with torch.cuda.device(self.device):
            for name, t in state_dict.items():
                if name.endswith(("_codebook.embed_avg", "_codebook.cluster_size", "_codebook.initted")):
                    continue  # EMA training state, unused at inference
                d = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                if name.endswith("_codebook.embed"):
                    self._keep.append(d)
                    self.codebook = d
                shape = (C.c_int64 * max(d.dim(), 1))(*d.shape)
                _abi.check(self.lib.dc_set_tensor(self.h, name.encode(), d.data_ptr(), shape, d.dim()),
                           f"dc_set_tensor({name})")
            _abi.check(self.lib.dc_finalize(self.h, self._stream()), "dc_finalize")
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            self.lib.dc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_option(self, key: str, value: float) -> None:
        _abi.check(self.lib.dc_set_option(self.h, key.encode(), float(value)), f"dc_set_option({key})")

    def workspace_bytes(self, stage: int, B: int, T: int) -> int:
        n = C.c_size_t()
        _abi.check(self.lib.dc_workspace_bytes(self.h, stage, B, T, C.byref(n)), "dc_workspace_bytes")
        return n.value

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def clips_per_call(self, stage: int, B: int, T: int) -> int:
        """Largest clip count whose workspace fits `workspace_limit` (workspace is linear in B)."""
        per = self.workspace_bytes(stage, 1, T)
        cap = max(1, self.workspace_limit // max(per, 1))
        cap = min(cap, max(1, ((1 << 31) - 1) // (T * self.hop) - 1))
        return max(1, min(B, cap))

    @property
    def act_dtype(self) -> torch.dtype:
        return torch.bfloat16 if self.mode == "bf16" else torch.float32

    def _check_in(self, t: torch.Tensor, dtype, shape_tail=None):
        if t.device != self.device:
            raise RuntimeError(f"tensor on {t.device}, engine on {self.device}")
        if t.dtype != dtype:
            raise RuntimeError(f"expected {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("C-ABI inputs must be contiguous")
        if shape_tail is not None and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
            raise RuntimeError(f"expected trailing shape {shape_tail}, got {tuple(t.shape)}")

    # ------------------------------------------------------------------------------------------ stages
    def encoder(self, mel_ncl: torch.Tensor) -> torch.Tensor:
        """mel (B, n_mels, T) fp32 channels-first -> latents (B, T, 1024) fp32.  models/encoders.py:68-76."""
        self._check_in(mel_ncl, torch.float32)
        B, _, T = mel_ncl.shape
        out = torch.empty(B, T, self.latent_dim, dtype=torch.float32, device=self.device)
        step = self.clips_per_call(_abi.STAGE_ENCODER, B, T)
        with torch.cuda.device(self.device):
            for b0 in range(0, B, step):
                b = min(step, B - b0)
                ws = self._workspace(self.workspace_bytes(_abi.STAGE_ENCODER, b, T))
                _abi.check(self.lib.dc_encoder_forward(self.h, mel_ncl[b0:b0 + b].data_ptr(), b, T,
                                                       out[b0:b0 + b].data_ptr(), ws.data_ptr(), ws.numel(),
                                                       self._stream()), "dc_encoder_forward")
        return out

    def quantizer(self, enc_nlc: torch.Tensor, want_fup: bool = True):
        """latents (B, T, 1024) fp32 -> (codes (B,T) int64, x_pjt_in (B,T,3584), quantized_fup (B,T,3584) fp32 or
        None, quantized (B,T,1024) fp32).  vector_quantization/grfvq.py:105-132."""
        self._check_in(enc_nlc, torch.float32, (self.latent_dim,))
        B, T, _ = enc_nlc.shape
        dev = self.device
        codes = torch.empty(B, T, dtype=torch.int64, device=dev)
        xin = torch.empty(B, T, self.code_dim, dtype=self.act_dtype, device=dev)
        fup = torch.empty(B, T, self.code_dim, dtype=torch.float32, device=dev) if want_fup else None
        quant = torch.empty(B, T, self.latent_dim, dtype=torch.float32, device=dev)
        step = self.clips_per_call(_abi.STAGE_QUANTIZER, B, T)
        with torch.cuda.device(dev):
            for b0 in range(0, B, step):
                b = min(step, B - b0)
                ws = self._workspace(self.workspace_bytes(_abi.STAGE_QUANTIZER, b, T))
                _abi.check(self.lib.dc_quantizer_forward(
                    self.h, enc_nlc[b0:b0 + b].data_ptr(), b, T, codes[b0:b0 + b].data_ptr(),
                    xin[b0:b0 + b].data_ptr(), _ptr(fup[b0:b0 + b]) if want_fup else None,
                    quant[b0:b0 + b].data_ptr(), ws.data_ptr(), ws.numel(), self._stream()), "dc_quantizer_forward")
        return codes, xin, fup, quant

    def quantizer_encode(self, enc_nlc: torch.Tensor) -> torch.Tensor:
        """latents (B, T, 1024) fp32 -> codes (B, T) int64 only (DownsampleGRVQ.encode, grfvq.py:134-139): skips the
        codebook gather and the project_out / upsample tail of the full forward."""
        self._check_in(enc_nlc, torch.float32, (self.latent_dim,))
        B, T, _ = enc_nlc.shape
        codes = torch.empty(B, T, dtype=torch.int64, device=self.device)
        step = self.clips_per_call(_abi.STAGE_QUANTIZER, B, T)
        with torch.cuda.device(self.device):
            for b0 in range(0, B, step):
                b = min(step, B - b0)
                ws = self._workspace(self.workspace_bytes(_abi.STAGE_QUANTIZER, b, T))
                _abi.check(self.lib.dc_quantizer_encode(self.h, enc_nlc[b0:b0 + b].data_ptr(), b, T,
                                                        codes[b0:b0 + b].data_ptr(), ws.data_ptr(), ws.numel(),
                                                        self._stream()), "dc_quantizer_encode")
        return codes

    def decode_codes(self, codes: torch.Tensor) -> torch.Tensor:
        """codes (B, T) int64 -> z (B, T, 1024) fp32.  grfvq.py:141-146."""
        self._check_in(codes, torch.int64)
        B, T = codes.shape
        z = torch.empty(B, T, self.latent_dim, dtype=torch.float32, device=self.device)
        step = self.clips_per_call(_abi.STAGE_DECODE_CODES, B, T)
        with torch.cuda.device(self.device):
            for b0 in range(0, B, step):
                b = min(step, B - b0)
                ws = self._workspace(self.workspace_bytes(_abi.STAGE_DECODE_CODES, b, T))
                _abi.check(self.lib.dc_quantizer_decode(self.h, codes[b0:b0 + b].data_ptr(), b, T,
                                                        z[b0:b0 + b].data_ptr(), ws.data_ptr(), ws.numel(),
                                                        self._stream()), "dc_quantizer_decode")
        return z

    def generator(self, z_nlc: torch.Tensor) -> torch.Tensor:
        """z (B, T, 1024) fp32 -> waveform (B, hop*T) fp32 in (-1, 1).  models/generators.py:118-147."""
        self._check_in(z_nlc, torch.float32, (self.latent_dim,))
        B, T, _ = z_nlc.shape
        wav = torch.empty(B, T * self.hop, dtype=torch.float32, device=self.device)
        step = self.clips_per_call(_abi.STAGE_GENERATOR, B, T)
        with torch.cuda.device(self.device):
            for b0 in range(0, B, step):
                b = min(step, B - b0)
                ws = self._workspace(self.workspace_bytes(_abi.STAGE_GENERATOR, b, T))
                _abi.check(self.lib.dc_generator_forward(self.h, z_nlc[b0:b0 + b].data_ptr(), b, T,
                                                         wav[b0:b0 + b].data_ptr(), ws.data_ptr(), ws.numel(),
                                                         self._stream()), "dc_generator_forward")
        return wav

    def mel(self, audio: torch.Tensor) -> torch.Tensor:
        """audio (B, Ls) or (B, 1, Ls) fp32 on the device -> log-mel (B, 128, T) fp32 (models/mel_spec.py:109-122)."""
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        self._check_in(audio, torch.float32)
        B, Ls = audio.shape
        T = (Ls - 256) // 256 + 1
        out = torch.empty(B, 128, T, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.dc_mel_forward(self.h, audio.data_ptr(), B, Ls, out.data_ptr(), self._stream()),
                       "dc_mel_forward")
        return out

    # ------------------------------------------------------------------------------------------ ops
    def vq_search(self, x: torch.Tensor, x2: Optional[torch.Tensor] = None, stats: bool = False):
        """Nearest-code search only.  x (N, 3584) bf16 or fp32; x2 optional fp32 (N,) row square-norms as the
        caller's reference computes them.  -> codes (N,) int64 [, stats dict]."""
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("vq_search: x must be fp32 or bf16")
        self._check_in(x, x.dtype, (self.code_dim,))
        N = x.shape[0]
        if x2 is not None:
            self._check_in(x2, torch.float32)
            assert x2.numel() == N
        codes = torch.empty(N, dtype=torch.int64, device=self.device)
        n = C.c_size_t()
        is_bf16 = int(x.dtype == torch.bfloat16)
        _abi.check(self.lib.dc_vq_workspace_bytes(self.h, N, is_bf16, C.byref(n)), "dc_vq_workspace_bytes")
        st = (C.c_int * 4)() if stats else None
        with torch.cuda.device(self.device):
            ws = self._workspace(n.value)
            _abi.check(self.lib.dc_vq_search(self.h, x.data_ptr(), is_bf16, _ptr(x2), N, codes.data_ptr(),
                                             ws.data_ptr(), ws.numel(), self._stream(), st), "dc_vq_search")
        if stats:
            return codes, {"rows": st[0], "rescored": st[1], "exhaustive_rows": st[2], "single": st[3]}
        return codes

    def op_conv_gemm(self, a, w, bias, res, shift0, dil, act):
        """Test hook for the implicit-GEMM kernel: a (B,T,C) fp32, w (N, J*C) fp32 -> (B,T,N) fp32."""
        B, T, Cc = a.shape
        N, JC = w.shape
        J = JC // Cc
        out = torch.empty(B, T, N, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.dc_op_conv_gemm(self.h, a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(res),
                                                out.data_ptr(), B, T, Cc, J, shift0, dil, N, act, self._stream()),
                       "dc_op_conv_gemm")
        return out

    def op_dwconv_ln(self, x, dw_w, dw_b, ln_w, ln_b):
        """Test hook: x (B,T,C) fp32; dw_w (C,1,7) or None -> LayerNorm(dwconv7(x)) (B,T,C) fp32."""
        B, T, Cc = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.dc_op_dwconv_ln(self.h, x.data_ptr(), _ptr(dw_w), _ptr(dw_b), ln_w.data_ptr(),
                                                ln_b.data_ptr(), out.data_ptr(), B, T, Cc, self._stream()),
                       "dc_op_dwconv_ln")
        return out

    def ncl_to_nlc(self, x: torch.Tensor) -> torch.Tensor:
        """(B, C, T) fp32 -> contiguous (B, T, C); zero-copy when x already is a transposed view of NLC memory."""
        xt = x.transpose(1, 2)
        if xt.is_contiguous():
            return xt
        x = x.contiguous()
        B, Cc, T = x.shape
        out = torch.empty(B, T, Cc, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.dc_ncl_to_nlc(x.data_ptr(), out.data_ptr(), B, Cc, T, self._stream()), "dc_ncl_to_nlc")
        return out

    def profile(self, on: bool) -> None:
        """Bracket every kernel launch of this thread with CUDA events (roofline reports); see `profile_rows`."""
        _abi.check(self.lib.dc_profile_enable(int(on)), "dc_profile_enable")

    def profile_rows(self):
        """-> [{name, launches, ms, flops, bytes}] per kernel class since `profile(True)`; clears the records."""
        cap = 512
        rows = (_abi.DcProfileRow * cap)()
        n = C.c_int()
        _abi.check(self.lib.dc_profile_collect(rows, cap, C.byref(n)), "dc_profile_collect")
        return [{"name": rows[i].name.decode(), "launches": int(rows[i].launches), "ms": rows[i].ms,
                 "flops": rows[i].flops, "bytes": rows[i].bytes} for i in range(min(n.value, cap))]

    def launch_count(self) -> int:
        return int(self.lib.dc_launch_count())
