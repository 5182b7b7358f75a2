"""Clip sharding and the host-buffer pipeline around the three stage calls.

The hot path has no cross-clip operation (the batch dimension is never reduced in
models/encoders.py:68-76, the eval VQ path or models/generators.py:118-147), so multi-GPU = every rank runs the
same single-GPU path on its own contiguous slice of the clip list and results are concatenated in input order.
No collective touches the data path; `gather_by_clip` is the only exchange (codes / waveforms / per-rank
counters, a few bytes per frame) and goes through `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).

`Pipeline` is the call a user of the reference's `DistilCodec.forward` / `encode` + `decode_from_codes`
(distil_codec.py:518-530, 545-594) makes with HOST buffers: log-mel (B, 128, T) in pinned host memory in,
codes (B, T) and waveform (B, 256 T) in pinned host memory out, with the host<->device copies issued on a side
stream so that chunk i+1 uploads while chunk i computes.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_clips(n_clips: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced partition of clip indices: the first `n_clips % world_size` ranks get one extra clip.
    Concatenating the shards of rank 0..world_size-1 restores input order."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if n_clips < 0:
        raise ValueError("n_clips must be >= 0")
    base, extra = divmod(n_clips, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_sizes(n_clips: int, world_size: int) -> List[int]:
    return [len(shard_clips(n_clips, world_size, r)) for r in range(world_size)]


def gather_by_clip(local: torch.Tensor, n_clips: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Concatenate per-rank results (first dim = this rank's clips, in `shard_clips` order) back into input order.

    Works on the backend of the default process group (NCCL: `local` on the rank's GPU; gloo: CPU tensors).
    Without an initialised process group it is the identity.  `dst=None` -> every rank gets the full tensor
    (all_gather); `dst=r` -> only rank r does, the others return None."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_clips:
            raise ValueError("single process: local result must cover every clip")
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_clips, ws)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} clips, expected {sizes[rank]}")
    # pad every shard to the largest so that one all_gather of equal-sized buffers suffices
    mx = max(sizes) if sizes else 0
    tail = tuple(local.shape[1:])
    buf = local.new_zeros((mx,) + tail)
    buf[: sizes[rank]] = local
    parts = [local.new_empty((mx,) + tail) for _ in range(ws)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    if dst is not None and rank != dst:
        return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


class Pipeline:
    """mel (host) -> codes, waveform (host) on one GPU, in chunks of clips that fit the workspace budget.

    engine : distilcodec_nabeel_b200.Engine (one device, one numeric mode)
    chunk  : clips per device pass; None = as many as `engine.workspace_limit` allows for the generator stage
    """

    def __init__(self, engine, chunk: Optional[int] = None):
        self.eng = engine
        self.chunk = chunk
        self.dev = engine.device
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _chunk(self, B: int, T: int) -> int:
        if self.chunk:
            return max(1, min(B, self.chunk))
        from . import _abi
        return self.eng.clips_per_call(_abi.STAGE_GENERATOR, B, T)

    # ---- device-resident legs (used by bench.py's kernel-only timing) --------------------------------------
    def encode_device(self, mel_dev: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """mel (B,128,T) on device -> (codes (B,T) int64, quantized (B,T,1024) fp32) on device."""
        enc = self.eng.encoder(mel_dev)
        codes, _, _, quant = self.eng.quantizer(enc, want_fup=False)
        return codes, quant

    def reconstruct_device(self, mel_dev: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """DistilCodec.forward minus the CPU front-end (distil_codec.py:518-530): -> (codes, wav (B, 256T))."""
        codes, quant = self.encode_device(mel_dev)
        return codes, self.eng.generator(quant)

    def decode_device(self, codes_dev: torch.Tensor) -> torch.Tensor:
        """decode_from_codes (distil_codec.py:581-594) for a batch: codes (B,T) -> wav (B, 256T)."""
        return self.eng.generator(self.eng.decode_codes(codes_dev))

    # ---- host-buffer legs ---------------------------------------------------------------------------------
    def _run_host(self, mel_host: torch.Tensor, want_wav: bool, codes_out: torch.Tensor,
                  wav_out: Optional[torch.Tensor]):
        B, _, T = mel_host.shape
        step = self._chunk(B, T)
        main = torch.cuda.current_stream(self.dev)
        pending = None  # (event, device tensors kept alive until their D2H copy has been issued)
        nxt = None
        with torch.cuda.device(self.dev):
            for b0 in range(0, B, step):
                b1 = min(B, b0 + step)
                if nxt is None:
                    with torch.cuda.stream(self.copy_stream):
                        cur = mel_host[b0:b1].to(self.dev, non_blocking=True)
                        ev_up = torch.cuda.Event()
                        ev_up.record(self.copy_stream)
                else:
                    cur, ev_up = nxt
                self.h2d_bytes += cur.numel() * cur.element_size()
                # prefetch the next chunk while this one computes
                if b1 < B:
                    b2 = min(B, b1 + step)
                    with torch.cuda.stream(self.copy_stream):
                        n_t = mel_host[b1:b2].to(self.dev, non_blocking=True)
                        n_ev = torch.cuda.Event()
                        n_ev.record(self.copy_stream)
                    nxt = (n_t, n_ev)
                else:
                    nxt = None
                main.wait_event(ev_up)
                cur.record_stream(main)
                if want_wav:
                    codes, wav = self.reconstruct_device(cur)
                else:
                    codes, _ = self.encode_device(cur)
                    wav = None
                ev_done = torch.cuda.Event()
                ev_done.record(main)
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(ev_done)
                    codes_out[b0:b1].copy_(codes, non_blocking=True)
                    codes.record_stream(self.copy_stream)
                    self.d2h_bytes += codes.numel() * 8
                    if wav is not None:
                        wav_out[b0:b1].copy_(wav, non_blocking=True)
                        wav.record_stream(self.copy_stream)
                        self.d2h_bytes += wav.numel() * 4
                pending = (codes, wav)
        self.copy_stream.synchronize()
        del pending

    def reconstruct(self, mel_host: torch.Tensor, codes_out: Optional[torch.Tensor] = None,
                    wav_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full encode -> quantize -> decode from HOST log-mel (B,128,T) fp32.  Returns host (codes, wav)."""
        B, _, T = mel_host.shape
        if codes_out is None:
            codes_out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)
        if wav_out is None:
            wav_out = torch.empty(B, T * self.eng.hop, dtype=torch.float32, pin_memory=True)
        self._run_host(mel_host, True, codes_out, wav_out)
        return codes_out, wav_out

    def tokenize(self, mel_host: torch.Tensor, codes_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """wav->codes leg (DistilCodec.encode, distil_codec.py:545-573) from HOST log-mel.  Returns host codes."""
        B, _, T = mel_host.shape
        if codes_out is None:
            codes_out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)
        self._run_host(mel_host, False, codes_out, None)
        return codes_out

    def tokenize_wav(self, wav_host: torch.Tensor, codes_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """wav -> codes entirely on the device (DistilCodec.encode with raw_audio=True, distil_codec.py:545-573 +
        :99-145) from HOST audio (B, n) fp32 at the model rate, equal lengths: left-pad by one zero sample (:134), GPU
        log-mel (the reference runs this stage on the CPU), encoder, VQ.  Needs the engine's mel buffers."""
        B, n = wav_host.shape
        T = (n + 1 - 256) // 256 + 1
        if codes_out is None:
            codes_out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)
        step = self._chunk(B, T)
        with torch.cuda.device(self.dev):
            for b0 in range(0, B, step):
                b1 = min(B, b0 + step)
                w = wav_host[b0:b1].to(self.dev, non_blocking=True)
                self.h2d_bytes += w.numel() * 4
                mel = self.eng.mel(torch.nn.functional.pad(w, (1, 0)).contiguous())
                codes, _ = self.encode_device(mel)
                codes_out[b0:b1].copy_(codes, non_blocking=True)
                self.d2h_bytes += codes.numel() * 8
            torch.cuda.current_stream(self.dev).synchronize()
        return codes_out

    def decode(self, codes_host: torch.Tensor, wav_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """codes->wav leg (decode_from_codes, distil_codec.py:581-594) from HOST codes (B,T) int64."""
        B, T = codes_host.shape
        if wav_out is None:
            wav_out = torch.empty(B, T * self.eng.hop, dtype=torch.float32, pin_memory=True)
        step = self._chunk(B, T)
        with torch.cuda.device(self.dev):
            for b0 in range(0, B, step):
                b1 = min(B, b0 + step)
                c = codes_host[b0:b1].to(self.dev, non_blocking=True)
                self.h2d_bytes += c.numel() * 8
                wav = self.decode_device(c)
                wav_out[b0:b1].copy_(wav, non_blocking=True)
                self.d2h_bytes += wav.numel() * 4
            torch.cuda.current_stream(self.dev).synchronize()
        return wav_out


# ---- time tiling: clips longer than one workspace -------------------------------------
# The path has a finite receptive field and no global-in-time operation: encoder = stem k7 + 18 ConvNeXt depthwise k7
# (models/encoders.py:8-76) -> +-57 frames, quantizer pre/post blocks +-6 each (grfvq.py:28-103), decoder conv_pre k7 +
# 5 ConvTranspose/ResBlock stages (models/generators.py:29-147) -> < +-25 frames.  A time tile computed with HALO real
# frames of context on each side (clip ends keep the convs' zero padding) therefore reproduces the whole-clip result
# on its interior; every output element is produced by the same instruction sequence wherever its tile starts, so the
# match is bit-exact (tests/test_gpu_e2e.py).
ENCODE_HALO = 96
DECODE_HALO = 48


def time_tiles(T: int, tile: int, halo: int):
    """[(lo, hi, s, e)]: compute frames [lo, hi) to obtain frames [s, e); tiles cover [0, T) in order."""
    if tile <= 0:
        raise ValueError("tile must be positive")
    out = []
    for s in range(0, T, tile):
        e = min(T, s + tile)
        out.append((max(0, s - halo), min(T, e + halo), s, e))
    return out


def tokenize_long(pipe: "Pipeline", mel_host: torch.Tensor, tile: int = 8192,
                  codes_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wav->codes leg for clips of any length with device memory bounded by `tile` frames: HOST log-mel (B,128,T) ->
    host codes (B,T), identical to `Pipeline.tokenize` on the whole clip."""
    B, _, T = mel_host.shape
    if codes_out is None:
        codes_out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)
    for lo, hi, s, e in time_tiles(T, tile, ENCODE_HALO):
        c = pipe.tokenize(mel_host[:, :, lo:hi].contiguous())
        codes_out[:, s:e] = c[:, s - lo:e - lo]
    return codes_out


def decode_long(pipe: "Pipeline", codes_host: torch.Tensor, tile: int = 8192,
                wav_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """codes->wav leg for clips of any length: HOST codes (B,T) -> host waveform (B, hop*T), identical to
    `Pipeline.decode` on the whole clip."""
    B, T = codes_host.shape
    hop = pipe.eng.hop
    if wav_out is None:
        wav_out = torch.empty(B, T * hop, dtype=torch.float32, pin_memory=True)
    for lo, hi, s, e in time_tiles(T, tile, DECODE_HALO):
        w = pipe.decode(codes_host[:, lo:hi].contiguous())
        wav_out[:, s * hop:e * hop] = w[:, (s - lo) * hop:(e - lo) * hop]
    return wav_out


def run_sharded(fn, items: Sequence, world_size: int, rank: int):
    """Apply `fn` to this rank's shard of `items` (a list of clips) and return (indices, results)."""
    idx = shard_clips(len(items), world_size, rank)
    return list(idx), [fn(items[i]) for i in idx]
