"""Deterministic synthetic inputs (numpy Philox; bit-identical on every machine) shared by the golden
generator, the tests and bench.py."""
import numpy as np
import torch


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=seed))


def make_mel(B, T, seed=11):
    """Log-mel-like input (B,128,T): the reference clamps at log(1e-5) = -11.5 (models/mel_spec.py:101)."""
    r = _rng(seed)
    m = r.standard_normal(size=(B, 128, T), dtype=np.float32) * np.float32(2.0) - np.float32(4.0)
    return torch.from_numpy(np.maximum(m, np.float32(-11.5)))


def make_wav(B, n, seed=3):
    """0.1 * N(0,1) clipped to +-1."""
    r = _rng(seed)
    w = r.standard_normal(size=(B, n), dtype=np.float32) * np.float32(0.1)
    return torch.from_numpy(np.clip(w, -1.0, 1.0))


def make_vq_rows(N, D=3584, kind="bf16", scale=0.42, seed=2):
    """Synthetic project_in outputs: N(0, scale) rows; kind='bf16' rounds to bf16-representable fp32 values
    (what the reference sees under enable_bfloat16: x is bf16 then .float(), vector_quantize_pytorch.py:473)."""
    r = _rng(seed)
    x = torch.from_numpy(r.standard_normal(size=(N, D), dtype=np.float32) * np.float32(scale))
    if kind == "bf16":
        x = x.to(torch.bfloat16).float()
    return x


def make_latents(B, T, C=1024, seed=5):
    """Unit-variance latents like the encoder's final LayerNorm emits."""
    r = _rng(seed)
    return torch.from_numpy(r.standard_normal(size=(B, C, T), dtype=np.float32))
