"""Per-kernel parity on the GPU, through the C ABI (dc_op_*), against the matching torch fp32 op on CPU.
fp32 mode tolerance 1e-4, bf16 (tcgen05) mode 1e-2, both max-abs relative to the reference's max-abs."""
import pytest
import torch
import torch.nn.functional as F

from tests.conftest import engine, rel_err

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _conv_ref(a_nlc, w_oik, bias, dil, pad, act, res):
    y = F.conv1d(a_nlc.transpose(1, 2), w_oik, bias, padding=pad, dilation=dil).transpose(1, 2)
    if act == 1:
        y = F.gelu(y)
    elif act == 2:
        y = F.silu(y)
    if res is not None:
        y = y + res
    return y


# (B, T, C, N, k, dil, act, with_res): the layer shapes of the hot path (and ragged T / tiny batches)
CONV_CASES = [
    (2, 200, 128, 256, 7, 1, 0, False),    # encoder stem
    (3, 97, 256, 1024, 1, 1, 1, False),    # pwconv1 + GELU
    (3, 97, 1024, 256, 1, 1, 0, True),     # pwconv2 + residual
    (1, 130, 768, 1024, 1, 1, 0, False),   # inter-stage 1x1
    (2, 140, 1024, 1024, 13, 1, 0, False), # conv_pre k13
    (2, 300, 512, 512, 3, 5, 2, False),    # ResBlock conv k3 d5 + SiLU
    (2, 300, 256, 256, 11, 5, 0, True),    # ResBlock conv k11 d5 (+-25 halo) + residual
    (1, 515, 128, 128, 7, 3, 2, False),    # C = N = 128: tap-shared 256-row tiles (conv_ts)
    (3, 300, 128, 128, 11, 5, 0, True),    # ... largest halo (+-25), residual, several clips, ragged second tile
    (2, 256, 128, 128, 3, 1, 2, False),
    (2, 700, 64, 64, 11, 1, 0, True),
    (2, 1000, 32, 32, 3, 1, 2, False),     # C = 32 stage (64-byte swizzle tiles)
    (1, 1, 1024, 3584, 1, 1, 0, False),    # project_in, a single frame
    (1, 5, 3584, 1024, 1, 1, 0, False),    # project_out
]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "B%dT%dC%dN%dk%dd%da%dr%d" % tuple(int(v) for v in c))
def test_conv_gemm(mode, case):
    B, T, C, N, k, dil, act, with_res = case
    eng = engine("W1", mode, 1024)
    a = _rand(B, T, C, seed=1)
    w = _rand(N, C, k, seed=2, scale=(C * k) ** -0.5)
    bias = _rand(N, seed=3, scale=0.1)
    res = _rand(B, T, N, seed=4) if with_res else None
    pad = dil * (k - 1) // 2
    ref = _conv_ref(a, w, bias, dil, pad, act, res)
    w_pack = w.permute(0, 2, 1).reshape(N, k * C).contiguous()   # [N][j*C + c]
    dev = eng.device
    out = eng.op_conv_gemm(a.to(dev), w_pack.to(dev), bias.to(dev), None if res is None else res.to(dev),
                           -pad, dil, act)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("k,s", [(16, 8), (12, 4), (4, 2)])
def test_conv_transpose_as_phase_gemm(mode, k, s):
    """ConvTranspose1d(k, s, p=(k-s)//2) (models/generators.py:67-79) = one shifted-row GEMM whose N stacks the s
    output phases; the packing rule lives in api.cu pack_dense and is restated here independently."""
    B, T, C, O = 2, 50, 128, 64
    p = (k - s) // 2
    eng = engine("W1", mode, 1024)
    a = _rand(B, T, C, seed=5)
    w = _rand(C, O, k, seed=6, scale=(C * k / s) ** -0.5)        # (in, out, k)
    bias = _rand(O, seed=7, scale=0.1)
    ref = F.conv_transpose1d(a.transpose(1, 2), w, bias, stride=s, padding=p).transpose(1, 2)   # (B, T*s, O)
    shifts = sorted({(ph + p - kk) // s for ph in range(s) for kk in range(k) if (ph + p - kk) % s == 0})
    sh0, J = shifts[0], shifts[-1] - shifts[0] + 1
    wp = torch.zeros(s * O, J * C)
    for ph in range(s):
        for j in range(J):
            kk = ph + p - (sh0 + j) * s
            if 0 <= kk < k:
                wp[ph * O:(ph + 1) * O, j * C:(j + 1) * C] = w[:, :, kk].T
    dev = eng.device
    out = eng.op_conv_gemm(a.to(dev), wp.to(dev), bias.repeat(s).to(dev), None, sh0, 1, 0)     # (B, T, s*O)
    torch.cuda.synchronize()
    assert rel_err(out.reshape(B, T * s, O), ref) < TOL[mode]


@pytest.mark.parametrize("C", [256, 512, 768, 1024])
@pytest.mark.parametrize("conv", [True, False])
def test_dwconv_layernorm(C, conv):
    """dwconv k7 + LN over C (convnext_utils.py:265-268) and the plain channels_first LN (:208-213)."""
    eng = engine("W1", "fp32", 1024)
    B, T = 3, 61
    x = _rand(B, T, C, seed=8)
    ln_w, ln_b = 1 + _rand(C, seed=9, scale=0.2), _rand(C, seed=10, scale=0.2)
    dev = eng.device
    if conv:
        dw_w, dw_b = _rand(C, 1, 7, seed=11, scale=0.3), _rand(C, seed=12, scale=0.1)
        y = F.conv1d(x.transpose(1, 2), dw_w, dw_b, padding=3, groups=C).transpose(1, 2)
        out = eng.op_dwconv_ln(x.to(dev), dw_w.to(dev), dw_b.to(dev), ln_w.to(dev), ln_b.to(dev))
    else:
        y = x
        out = eng.op_dwconv_ln(x.to(dev), None, None, ln_w.to(dev), ln_b.to(dev))
    ref = F.layer_norm(y, (C,), ln_w, ln_b, 1e-6)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-5


def test_layout_helpers_round_trip():
    eng = engine("W1", "fp32", 1024)
    x = _rand(3, 130, 77, seed=13).to(eng.device)            # (B, C, T) ragged sizes
    nlc = eng.ncl_to_nlc(x)
    assert torch.equal(nlc, x.transpose(1, 2).contiguous())
    assert eng.ncl_to_nlc(nlc.transpose(1, 2)).data_ptr() == nlc.data_ptr()   # zero-copy for NLC-backed views


# (B, T, C, N, k, dil, act, with_res): wide convs (conv_tsw) with odd / even / single row-tile counts
PAIR_CASES = [
    (1, 300, 256, 256, 7, 1, 0, True),      # 3 row tiles: the last pair runs a ghost tile
    (3, 130, 512, 512, 3, 5, 2, False),     # 6 row tiles, two N blocks
    (2, 300, 256, 256, 11, 5, 0, True),     # largest halo
    (1, 641, 512, 1024, 3, 1, 0, False),    # 6 row tiles (ragged last), four N blocks
    (5, 128, 256, 512, 7, 3, 2, False),     # 5 single-tile clips: pairs straddle clips
    (1, 100, 256, 256, 3, 1, 0, False),     # one row tile: falls back to single CTAs
    (3, 97, 256, 1024, 1, 1, 1, False),     # J = 1 (gemm_tc): pwconv1 + GELU, clips flattened, 3 row tiles (ghost)
    (3, 97, 1024, 256, 1, 1, 0, True),      # pwconv2 + residual
    (2, 140, 1024, 1024, 13, 1, 0, False),  # conv_pre k13
    (1, 1, 1024, 3584, 1, 1, 0, False),     # a single frame: single CTAs
    (1, 515, 128, 128, 7, 3, 2, False),     # C = N = 128 (conv_ts): 3 tiles of 256 rows (ghost)
    (3, 300, 128, 128, 11, 5, 0, True),     # ... 6 tiles, residual
    (2, 256, 128, 128, 3, 1, 2, False),     # ... the k = 3 variant (third activation buffer)
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=lambda c: "B%dT%dC%dN%dk%dd%da%dr%d" % tuple(int(v) for v in c))
def test_wide_conv_cta_pairs_with_weight_multicast(case):
    """conv_tsw / gemm_tc / conv_ts as clusters of two CTAs that TMA-multicast the weight tiles (option "cta_pairs" = 2) against
    the same kernels as single CTAs and against torch: the MMA sequence per tile is unchanged, so the results are
    bit-identical."""
    B, T, C, N, k, dil, act, with_res = case
    eng = engine("W1", "bf16", 1024)
    a = _rand(B, T, C, seed=11)
    w = _rand(N, C, k, seed=12, scale=(C * k) ** -0.5)
    bias = _rand(N, seed=13, scale=0.1)
    res = _rand(B, T, N, seed=14) if with_res else None
    pad = dil * (k - 1) // 2
    ref = _conv_ref(a, w, bias, dil, pad, act, res)
    w_pack = w.permute(0, 2, 1).reshape(N, k * C).contiguous()
    dev = eng.device
    args = (a.to(dev), w_pack.to(dev), bias.to(dev), None if res is None else res.to(dev), -pad, dil, act)
    outs = {}
    try:
        for cl in (1, 2):
            eng.set_option("cta_pairs", cl)
            outs[cl] = eng.op_conv_gemm(*args)
            torch.cuda.synchronize()
    finally:
        eng.set_option("cta_pairs", 2)
    assert rel_err(outs[2], ref) < TOL["bf16"]
    assert torch.equal(outs[1], outs[2])
