"""distilcodec_nabeel_b200 — B200-native (sm_100a) DistilCodec inference hot path.

ConvNeXt encoder -> single-codebook Euclidean VQ (32768 x 3584) -> HiFiGAN-style decoder, as hand-written CUDA
kernels behind a C ABI (include/distilcodec_b200.h, libdistilcodec_b200.so), plus the host-side mirror of the
reference's module interface.  There is no CPU or PyTorch fallback: without the built library and a CUDA device
every entry point raises.
"""
from . import _abi
from .engine import Engine, load_config
from .modules import (B200Encoder, B200Generator, B200MelSpectrogram, B200Quantizer, EngineSet, GRVQResult,
                      build_modules, mel_buffers, patch)
from .sharding import (Pipeline, decode_long, decode_long_device, gather_by_clip, shard_clips, time_tiles, tokenize_long,
                       tokenize_long_device)
from . import audio, bulk

__all__ = ["Engine", "EngineSet", "B200Encoder", "B200Quantizer", "B200Generator", "GRVQResult", "Pipeline",
           "build_modules", "patch", "B200MelSpectrogram", "mel_buffers", "load_config", "shard_clips", "gather_by_clip", "tokenize_long", "decode_long",
           "tokenize_long_device", "decode_long_device", "time_tiles", "audio", "bulk", "_abi"]
