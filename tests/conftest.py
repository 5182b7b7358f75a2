"""pytest configuration: the `gpu` marker and session-scoped weight sets / engines.

`-m "not gpu"`: oracle vs committed golden vectors of the reference, host logic, C-ABI load/exports (no GPU needed).
`-m gpu`      : parity of the CUDA path (through the C ABI) with the oracle and the golden vectors, on a B200.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_SD_CACHE = {}


def state_dict(variant: str, codebook_size=None):
    """oracle/weights.py W0 / W1, cached per session (the 32768 x 3584 codebook takes ~6 s to draw)."""
    from oracle import weights
    key = (variant, codebook_size)
    if key not in _SD_CACHE:
        _SD_CACHE[key] = weights.make_state_dict(variant, codebook_size=codebook_size)
    return _SD_CACHE[key]


@pytest.fixture(scope="session")
def sd_W0():
    return state_dict("W0")


@pytest.fixture(scope="session")
def sd_W1():
    return state_dict("W1")


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


_ENGINES = {}


def engine(variant: str, mode: str, codebook_size=None):
    """Session-cached Engine (C-ABI handle) on cuda:0.  Raises (never falls back) if the library is missing."""
    from distilcodec_nabeel_b200 import Engine, load_config, mel_buffers
    key = (variant, mode, codebook_size)
    if key not in _ENGINES:
        sd = dict(state_dict(variant, codebook_size))
        sd.update(mel_buffers(load_config()))          # the front-end's two non-persistent buffers (fb, window)
        _ENGINES[key] = Engine(sd, 0, mode)
    return _ENGINES[key]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to max |b| (the tolerance convention of SURVEY.md section 4)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
