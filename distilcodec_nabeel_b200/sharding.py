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


def _rows_pitch(t: torch.Tensor) -> Tuple[int, int, int]:
    """View with a contiguous last dimension whose leading dimensions collapse into one pitch -> (rows, width_bytes,
    pitch_bytes) for dc_copy2d_async."""
    es = t.element_size()
    if t.dim() == 0 or t.stride(-1) != 1 and t.shape[-1] > 1:
        raise ValueError("copy2d: last dimension must be contiguous")
    width = t.shape[-1]
    rows = 1
    pitch = width
    lead = [(n, st) for n, st in zip(t.shape[:-1], t.stride()[:-1]) if n > 1]
    if lead:
        pitch = lead[-1][1]
        expect = pitch
        for n, st in reversed(lead):
            if st != expect:
                raise ValueError("copy2d: leading dimensions do not collapse into one pitch")
            expect *= n
            rows *= n
    return rows, width * es, pitch * es


class Pipeline:
    """mel (host) -> codes, waveform (host) on one GPU, in chunks of clips that fit the workspace budget.

    engine : distilcodec_nabeel_b200.Engine (one device, one numeric mode)
    chunk  : clips per device pass; None = as many as `engine.workspace_limit` allows for the generator stage

    Every host-buffer leg is the same three-stage software pipeline over work items (a clip range x a time tile):
    upload (copy stream) -> compute (caller's stream) -> download (copy stream), so item i+1 uploads and item i-1
    downloads while item i computes.  Strided slices of PINNED host tensors (time tiles, per-clip crops) move by
    dc_copy2d_async, i.e. straight by DMA without a pageable staging copy.
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
    def encode_device(self, mel_dev: torch.Tensor, codes_only: bool = False):
        """mel (B,128,T) on device -> (codes (B,T) int64, quantized (B,T,1024) fp32) on device.  codes_only: the
        tokenisation legs stop at the search (quantizer.encode, grfvq.py:134-139) and return (codes, None)."""
        enc = self.eng.encoder(mel_dev)
        if codes_only:
            return self.eng.quantizer_encode(enc), None
        codes, _, _, quant = self.eng.quantizer(enc, want_fup=False)
        return codes, quant

    def reconstruct_device(self, mel_dev: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """DistilCodec.forward minus the CPU front-end (distil_codec.py:518-530): -> (codes, wav (B, 256T))."""
        codes, quant = self.encode_device(mel_dev)
        return codes, self.eng.generator(quant)

    def decode_device(self, codes_dev: torch.Tensor) -> torch.Tensor:
        """decode_from_codes (distil_codec.py:581-594) for a batch: codes (B,T) -> wav (B, 256T)."""
        return self.eng.generator(self.eng.decode_codes(codes_dev))

    def tokenize_wav_device(self, wav_padded_dev: torch.Tensor) -> torch.Tensor:
        """audio (B, n+1) on device, already left-padded by one zero sample (distil_codec.py:134) -> codes (B,T)."""
        return self.encode_device(self.eng.mel(wav_padded_dev), codes_only=True)[0]

    # ---- the copy / compute pipeline ----------------------------------------------------------------------
    def _copy2d(self, dst: torch.Tensor, src: torch.Tensor, stream: torch.cuda.Stream) -> int:
        """dst <- src (same shape; host views must be of pinned tensors) on `stream`; returns the bytes moved."""
        if tuple(dst.shape) != tuple(src.shape) or dst.dtype != src.dtype:
            raise ValueError(f"copy2d: {tuple(src.shape)} {src.dtype} -> {tuple(dst.shape)} {dst.dtype}")
        if src.numel() == 0:
            return 0
        for t in (dst, src):
            if not t.is_cuda and not t.is_pinned():
                raise ValueError("copy2d: host tensors must be pinned (pageable memory would make the copy synchronous)")
        rows, width, sp = _rows_pitch(src)
        rows_d, width_d, dp = _rows_pitch(dst)
        if (rows, width) != (rows_d, width_d):   # different collapses of the same shape: fall back to row = last dim
            raise ValueError("copy2d: source and destination do not collapse to the same rows")
        from . import _abi
        _abi.check(self.eng.lib.dc_copy2d_async(dst.data_ptr(), dp, src.data_ptr(), sp, width, rows,
                                                stream.cuda_stream), "dc_copy2d_async")
        return rows * width

    def _run_items(self, items: Sequence, upload, compute, download) -> None:
        """upload(item) -> device inputs [copy stream]; compute(item, inputs) -> device outputs [current stream];
        download(item, outputs) [copy stream, after the compute].  Item i+1 is uploaded while item i computes."""
        main = torch.cuda.current_stream(self.dev)
        cs = self.copy_stream

        def issue(item):
            with torch.cuda.stream(cs):
                ins = upload(item)
                ev = torch.cuda.Event()
                ev.record(cs)
            return ins, ev

        keep = None
        with torch.cuda.device(self.dev):
            cs.wait_stream(main)                       # inputs the caller produced on its stream
            nxt = issue(items[0]) if items else None
            for k, item in enumerate(items):
                ins, ev_up = nxt
                nxt = issue(items[k + 1]) if k + 1 < len(items) else None
                main.wait_event(ev_up)
                for t in ins:
                    t.record_stream(main)
                outs = compute(item, ins)
                ev_done = torch.cuda.Event()
                ev_done.record(main)
                with torch.cuda.stream(cs):
                    cs.wait_event(ev_done)
                    download(item, outs)
                    for t in outs:
                        t.record_stream(cs)
                keep = (ins, outs)
            cs.synchronize()
        del keep

    def _clip_items(self, B: int, T: int):
        step = self._chunk(B, T)
        return [(b0, min(B, b0 + step)) for b0 in range(0, B, step)]

    def _up(self, dst_dev: torch.Tensor, src_host: torch.Tensor) -> None:
        self.h2d_bytes += self._copy2d(dst_dev, src_host, self.copy_stream)

    def _down(self, dst_host: torch.Tensor, src_dev: torch.Tensor) -> None:
        self.d2h_bytes += self._copy2d(dst_host, src_dev, self.copy_stream)

    @staticmethod
    def _pinned(t: torch.Tensor) -> torch.Tensor:
        return t if t.is_pinned() else t.contiguous().pin_memory()

    # ---- host-buffer legs ---------------------------------------------------------------------------------
    def _run_host(self, mel_host: torch.Tensor, want_wav: bool, codes_out: torch.Tensor,
                  wav_out: Optional[torch.Tensor], tiles=None, crop_hop: int = 0):
        """mel_host (B,128,T) pinned.  `tiles` = [(lo, hi, s, e)] time tiles (default: the whole clip)."""
        mel_host = self._pinned(mel_host)
        B, M, T = mel_host.shape
        hop = self.eng.hop
        tiles = tiles or [(0, T, 0, T)]
        items = [(b0, b1, *t) for t in tiles for (b0, b1) in self._clip_items(B, t[1] - t[0])]

        def upload(it):
            b0, b1, lo, hi, _, _ = it
            cur = torch.empty(b1 - b0, M, hi - lo, dtype=torch.float32, device=self.dev)
            self._up(cur, mel_host[b0:b1, :, lo:hi])
            return (cur,)

        def compute(it, ins):
            if want_wav:
                return self.reconstruct_device(ins[0])
            return (self.encode_device(ins[0], codes_only=True)[0],)

        def download(it, outs):
            b0, b1, lo, hi, s, e = it
            self._down(codes_out[b0:b1, s:e], outs[0][:, s - lo:e - lo])
            if want_wav:
                self._down(wav_out[b0:b1, s * hop:e * hop], outs[1][:, (s - lo) * hop:(e - lo) * hop])

        self._run_items(items, upload, compute, download)

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

    def tokenize(self, mel_host: torch.Tensor, codes_out: Optional[torch.Tensor] = None, tiles=None) -> torch.Tensor:
        """wav->codes leg (DistilCodec.encode, distil_codec.py:545-573) from HOST log-mel.  Returns host codes."""
        B, _, T = mel_host.shape
        if codes_out is None:
            codes_out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)
        self._run_host(mel_host, False, codes_out, None, tiles)
        return codes_out

    def tokenize_wav(self, wav_host: torch.Tensor, codes_out: Optional[torch.Tensor] = None,
                     left_padded: bool = False) -> torch.Tensor:
        """wav -> codes entirely on the device (DistilCodec.encode with raw_audio=True, distil_codec.py:545-573 +
        :99-145) from HOST audio (B, n) fp32 at the model rate, equal lengths: left-pad by one zero sample (:134), GPU
        log-mel (the reference runs this stage on the CPU), encoder, VQ.  Needs the engine's mel buffers.
        left_padded=True: `wav_host` is (B, 1 + n) and already carries the pad (what `audio.load_batch` builds)."""
        wav_host = self._pinned(wav_host)
        B, n = wav_host.shape
        if left_padded:
            n -= 1
        T = (n + 1 - 256) // 256 + 1
        if codes_out is None:
            codes_out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)

        def upload(it):
            b0, b1 = it
            w = torch.empty(b1 - b0, n + 1, dtype=torch.float32, device=self.dev)
            if left_padded:
                self._up(w, wav_host[b0:b1])
            else:
                w[:, :1].zero_()                         # the reference's one-sample left pad (distil_codec.py:134)
                self._up(w[:, 1:], wav_host[b0:b1])
            return (w,)

        def compute(it, ins):
            return (self.tokenize_wav_device(ins[0]),)

        def download(it, outs):
            self._down(codes_out[it[0]:it[1]], outs[0])

        self._run_items(self._clip_items(B, T), upload, compute, download)
        return codes_out

    def tokenize_files(self, paths: Sequence[str], threads: int = 0) -> List[torch.Tensor]:
        """Files on disk -> codes: `preprocess_audio_batch` + `encode` (distil_codec.py:146-195, 545-563) with the
        native loader (audio.load_batch: parallel decode + resample into one pinned batch) and the GPU front-end.
        Clip i keeps its first `n_i // 256` codes.  -> list of host int64 tensors."""
        from . import audio
        sr = self.eng.cfg["spec_transform"]["sampling_rate"]
        batch, lengths, _ = audio.load_batch(list(paths), sr, threads=threads, left_pad=1, pin=True)
        codes = self.tokenize_wav(batch, left_padded=True)
        return [codes[i, :int(lengths[i]) // self.eng.hop].clone() for i in range(len(paths))]

    def tokenize_wav_ragged(self, wavs: Sequence, lengths_out: Optional[list] = None) -> List[torch.Tensor]:
        """Clips of different lengths exactly as `preprocess_raw_audio_batch` + `encode` treat them
        (distil_codec.py:99-145, 556-563): every clip is right-padded with zero AUDIO to the longest of the call, the
        whole batch is tokenised, and clip i keeps its first `n_i // 256` codes.  -> list of host int64 code tensors."""
        ns = [int(w.shape[-1]) for w in wavs]
        if not ns:
            return []
        mx = max(ns)
        batch = torch.zeros(len(ns), mx, dtype=torch.float32).pin_memory()
        for i, w in enumerate(wavs):
            batch[i, :ns[i]] = torch.as_tensor(w, dtype=torch.float32).reshape(-1)
        codes = self.tokenize_wav(batch)
        if lengths_out is not None:
            lengths_out.extend(n // self.eng.hop for n in ns)
        return [codes[i, :ns[i] // self.eng.hop].clone() for i in range(len(ns))]

    def decode(self, codes_host: torch.Tensor, wav_out: Optional[torch.Tensor] = None, tiles=None) -> torch.Tensor:
        """codes->wav leg (decode_from_codes, distil_codec.py:581-594) from HOST codes (B,T) int64."""
        codes_host = self._pinned(codes_host)
        B, T = codes_host.shape
        hop = self.eng.hop
        if wav_out is None:
            wav_out = torch.empty(B, T * hop, dtype=torch.float32, pin_memory=True)
        tiles = tiles or [(0, T, 0, T)]
        items = [(b0, b1, *t) for t in tiles for (b0, b1) in self._clip_items(B, t[1] - t[0])]

        def upload(it):
            b0, b1, lo, hi, _, _ = it
            c = torch.empty(b1 - b0, hi - lo, dtype=torch.int64, device=self.dev)
            self._up(c, codes_host[b0:b1, lo:hi])
            return (c,)

        def compute(it, ins):
            return (self.decode_device(ins[0]),)

        def download(it, outs):
            b0, b1, lo, hi, s, e = it
            self._down(wav_out[b0:b1, s * hop:e * hop], outs[0][:, (s - lo) * hop:(e - lo) * hop])

        self._run_items(items, upload, compute, download)
        return wav_out

    def vq_search(self, x_host: torch.Tensor, codes_out: Optional[torch.Tensor] = None, rows_per_pass: int = 8192
                  ) -> torch.Tensor:
        """Nearest-code search alone (EuclideanCodebook.forward eval path, vector_quantize_pytorch.py:462-538) from HOST
        rows x (N, 3584) bf16 or fp32 to host codes (N,), in row blocks that upload while the previous block is scored."""
        x_host = self._pinned(x_host)
        N = x_host.shape[0]
        if codes_out is None:
            codes_out = torch.empty(N, dtype=torch.int64, pin_memory=True)
        items = [(r0, min(N, r0 + rows_per_pass)) for r0 in range(0, N, rows_per_pass)]

        def upload(it):
            xb = torch.empty(it[1] - it[0], x_host.shape[1], dtype=x_host.dtype, device=self.dev)
            self._up(xb, x_host[it[0]:it[1]])
            return (xb,)

        def compute(it, ins):
            return (self.eng.vq_search(ins[0]),)

        def download(it, outs):
            self._down(codes_out[it[0]:it[1]], outs[0])

        self._run_items(items, upload, compute, download)
        return codes_out

    def decode_ragged(self, codes_list: Sequence, tails: str = "exact") -> List[torch.Tensor]:
        """Batched decode of code sequences of different lengths — what `decode_from_codes_batch` is meant to do
        (distil_codec.py:598-639: right-pad with code 0 to the longest, decode the batch, split per clip; the
        reference's own layout bug makes it decode clip 0 only) plus `save_wav`'s per-clip crop (:640-654).
        -> list of host fp32 waveforms, clip i of length `256 * len(codes_i)`.

        tails="padded": clip i is the batch row cropped, i.e. identical to `decode_from_codes` of the RIGHT-PADDED
                        sequence (its last ~25 frames see the code-0 padding as right context).
        tails="exact" : identical to `decode_from_codes(codes_i)` of the clip alone (length mask): the last
                        DECODE_HALO frames of every shorter clip are re-decoded from a tile that ends at the clip's own
                        end, where the convolutions see their zero padding (bit-exact, same argument as time tiling)."""
        if tails not in ("exact", "padded"):
            raise ValueError("tails must be 'exact' or 'padded'")
        lens = [int(len(c)) for c in codes_list]
        if not lens:
            return []
        hop = self.eng.hop
        mx = max(lens)
        batch = torch.zeros(len(lens), max(mx, 1), dtype=torch.int64).pin_memory()
        for i, c in enumerate(codes_list):
            batch[i, :lens[i]] = torch.as_tensor(c, dtype=torch.int64).reshape(-1)
        wav = self.decode(batch)
        outs = [wav[i, :lens[i] * hop].clone() for i in range(len(lens))]
        if tails == "exact":
            H = DECODE_HALO
            short = [i for i, n in enumerate(lens) if 0 < n < mx]
            # tail tiles of 2H frames ending at the clip end (whole clip when it is shorter than that), grouped by size
            by_len = {}
            for i in short:
                by_len.setdefault(min(lens[i], 2 * H), []).append(i)
            for L, idx in by_len.items():
                tail_codes = torch.stack([batch[i, lens[i] - L:lens[i]] for i in idx]).pin_memory()
                tw = self.decode(tail_codes)
                keep = min(L, H)
                for r, i in enumerate(idx):
                    outs[i][(lens[i] - keep) * hop:] = tw[r, (L - keep) * hop:]
        return outs


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
    host codes (B,T), identical to `Pipeline.tokenize` on the whole clip.  Tiles move between the pinned host tensors
    and the device by strided DMA and are double-buffered like every host-buffer leg."""
    T = mel_host.shape[2]
    return pipe.tokenize(mel_host, codes_out, tiles=time_tiles(T, tile, ENCODE_HALO))


def decode_long(pipe: "Pipeline", codes_host: torch.Tensor, tile: int = 8192,
                wav_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """codes->wav leg for clips of any length: HOST codes (B,T) -> host waveform (B, hop*T), identical to
    `Pipeline.decode` on the whole clip."""
    T = codes_host.shape[1]
    return pipe.decode(codes_host, wav_out, tiles=time_tiles(T, tile, DECODE_HALO))


def tokenize_long_device(pipe: "Pipeline", mel_dev: torch.Tensor, tile: int = 8192) -> torch.Tensor:
    """`tokenize_long` with the log-mel already in HBM: -> device codes (B,T)."""
    B, _, T = mel_dev.shape
    codes = torch.empty(B, T, dtype=torch.int64, device=mel_dev.device)
    step = pipe._chunk(B, min(T, tile + 2 * ENCODE_HALO))
    for lo, hi, s, e in time_tiles(T, tile, ENCODE_HALO):
        for b0 in range(0, B, step):
            c, _ = pipe.encode_device(mel_dev[b0:b0 + step, :, lo:hi].contiguous(), codes_only=True)
            codes[b0:b0 + step, s:e] = c[:, s - lo:e - lo]
    return codes


def decode_long_device(pipe: "Pipeline", codes_dev: torch.Tensor, tile: int = 8192) -> torch.Tensor:
    """`decode_long` with the codes already in HBM: -> device waveform (B, hop*T)."""
    B, T = codes_dev.shape
    hop = pipe.eng.hop
    wav = torch.empty(B, T * hop, dtype=torch.float32, device=codes_dev.device)
    step = pipe._chunk(B, min(T, tile + 2 * DECODE_HALO))
    for lo, hi, s, e in time_tiles(T, tile, DECODE_HALO):
        for b0 in range(0, B, step):
            w = pipe.decode_device(codes_dev[b0:b0 + step, lo:hi].contiguous())
            wav[b0:b0 + step, s * hop:e * hop] = w[:, (s - lo) * hop:(e - lo) * hop]
    return wav


def run_sharded(fn, items: Sequence, world_size: int, rank: int):
    """Apply `fn` to this rank's shard of `items` (a list of clips) and return (indices, results)."""
    idx = shard_clips(len(items), world_size, rank)
    return list(idx), [fn(items[i]) for i in idx]
