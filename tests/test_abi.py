"""The C-ABI boundary without a GPU: the shared library loads, exports every symbol include/distilcodec_b200.h
declares, and fails loudly (status + message, never a fallback) when no usable device exists."""
import ctypes as C
import os
import re

import pytest
import torch

from distilcodec_nabeel_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "distilcodec_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(_abi.SIGNATURES), "ctypes signature table out of sync with the header"


def test_version_and_default_config():
    lib = _abi.load()
    assert lib.dc_version() >= 100
    cfg = _abi.DcConfig()
    assert lib.dc_default_config(C.byref(cfg)) == 0
    assert cfg.n_mels == 128 and list(cfg.enc_dims) == [256, 512, 768, 1024] and list(cfg.enc_depths) == [3, 3, 9, 3]
    assert cfg.codebook_size == 32768 and cfg.codebook_dim == 3584
    assert list(cfg.up_rates)[:5] == [8, 4, 2, 2, 2] and list(cfg.up_kernels)[:5] == [16, 12, 4, 4, 4]


def test_packaged_config_equals_default_config():
    from distilcodec_nabeel_b200.engine import _dc_config, load_config
    lib = _abi.load()
    a, b = _abi.DcConfig(), _dc_config(load_config())
    lib.dc_default_config(C.byref(a))
    assert bytes(a) == bytes(b)


def test_error_convention_null_arguments():
    lib = _abi.load()
    assert lib.dc_default_config(None) == -1            # DC_ERR_ARG
    assert b"null" in lib.dc_last_error()
    assert lib.dc_set_option(None, b"vq_window", 1.0) == -1
    assert lib.dc_destroy(None) == 0
    n = C.c_size_t()
    assert lib.dc_workspace_bytes(None, 0, 1, 1, C.byref(n)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _abi.load()
    h = C.c_void_p()
    rc = lib.dc_create(0, _abi.MODE_BF16, None, C.byref(h))
    assert rc < 0 and not h.value
    assert len(lib.dc_last_error()) > 0
    with pytest.raises(RuntimeError):
        _abi.check(rc, "dc_create")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_has_no_cpu_fallback():
    from distilcodec_nabeel_b200 import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine({}, 0, "bf16")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distilcodec_nabeel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_conv_post_toeplitz_operand_reproduces_the_convolution():
    """conv_post + tanh runs as a GEMM over 8 consecutive samples on the tensor cores (csrc/conv_post.cu).  The operand
    layout is host code: check it against Conv1d(32 -> 1, k13, pad 6) (models/generators.py:141-145) in numpy, with the
    kernel's own indexing  y[8q + r] = sum_o S[q - 1 + o] . Wt[r][o*256 : (o+1)*256]  and zero padding at both ends."""
    import ctypes

    import numpy as np
    import torch

    lib = _abi.load()
    rng = np.random.default_rng(5)
    w = rng.standard_normal((13, 32)).astype(np.float32)          # [tap][channel]
    wt = np.zeros((16, 768), dtype=np.uint16)
    assert lib.dc_conv_post_toeplitz_weights(w.ctypes.data_as(ctypes.c_void_p), wt.ctypes.data_as(ctypes.c_void_p)) == 0
    W = torch.from_numpy(wt.view(np.int16)).view(torch.bfloat16).float().numpy()          # bf16 bit patterns -> fp32
    hi, lo = W[:8], W[8:]
    np.testing.assert_array_equal(hi, torch.from_numpy(hi + lo).bfloat16().float().numpy())  # rows 0-7 = bf16(w)
    Wsum = hi + lo                                                                           # = w to 16 mantissa bits
    L = 8 * 37
    s = rng.standard_normal((L, 32)).astype(np.float32)
    ref = torch.nn.functional.conv1d(torch.from_numpy(s.T[None]), torch.from_numpy(w.T[None].copy()), padding=6)[0, 0].numpy()
    S = np.zeros((L // 8 + 2, 256), dtype=np.float32)            # one zero super-row before and after = the padding
    S[1:-1] = s.reshape(L // 8, 256)
    y = np.zeros(L, dtype=np.float64)
    for q in range(L // 8):
        for o in range(3):
            y[8 * q:8 * q + 8] += Wsum[:, o * 256:(o + 1) * 256].astype(np.float64) @ S[q + o].astype(np.float64)
    assert np.abs(y - ref).max() < 2e-4 * np.abs(ref).max()      # weights rounded to 16 mantissa bits, nothing else
    # every weight appears exactly 8 times (once per output phase r), the rest of the 16 x 768 operand is zero
    assert np.count_nonzero(hi) == 8 * 13 * 32
