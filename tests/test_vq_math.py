"""CPU model of the arithmetic the CUDA nearest-code search performs (csrc/vq.cu), checked against the reference's
golden indices: (1) ||x||^2 summed in ATen's cascade order reproduces torch's CPU sum bit-for-bit, (2) the windowed
bf16 candidate pass + exact fp32 re-score returns the reference's index on every golden row.  This pins the
DESIGN of the kernel on CPU; tests/test_gpu_vq.py pins the kernel itself."""
import numpy as np
import pytest
import torch

from tests.conftest import golden, state_dict
from tests.golden.inputs import make_vq_rows


def aten_cascade_sqsum(x: torch.Tensor, W: int = 8) -> torch.Tensor:
    """Model of torch_cpu_row_sqsum (csrc/vq.cu): (x*x).sum(-1) in ATen's cascade_sum order (AVX2 build)."""
    n, D = x.shape
    v = (x * x).reshape(n, D // (4 * W), 4, W)
    size = v.shape[1]
    lp = max(4, int(np.ceil(np.log2(size))) // 4)
    step = 1 << lp
    z = lambda: torch.zeros(n, 4, W)
    acc = [z(), z(), z(), z()]
    i = 0
    while i + step <= size:
        for _ in range(step):
            acc[0] = acc[0] + v[:, i]
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = z()
            if i & ((step - 1) << (j * lp)):
                break
    while i < size:
        acc[0] = acc[0] + v[:, i]
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    s = acc[0][:, 0]
    for k in range(1, 4):
        s = s + acc[0][:, k]
    f = torch.zeros(n)
    for l in range(W):
        f = f + s[:, l]
    return f


@pytest.mark.parametrize("D", [3584, 1024, 8192 * 4])
def test_cascade_order_equals_torch_cpu_sum(D):
    x = make_vq_rows(64, D=D, kind="fp32", seed=9)
    assert torch.equal(aten_cascade_sqsum(x), (x ** 2).sum(-1))


@pytest.mark.parametrize("variant", ["W0", "W1"])
def test_windowed_candidates_plus_exact_rescore_reproduce_reference(variant):
    n = 512
    g = golden(f"vq_{variant}.npz")
    E = state_dict(variant)["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
    x = make_vq_rows(4096, kind="bf16")[:n]
    ref = torch.from_numpy(g["codes_bf16"][:n].astype(np.int64))
    c2 = aten_cascade_sqsum(E)
    x2 = aten_cascade_sqsum(x)
    # pass 1: bf16-operand score (x is bf16-exact here, the codebook is rounded)
    S = c2[None] - 2.0 * (x @ E.to(torch.bfloat16).float().T)
    u = 0.00390625 + 3584 * 1.1920929e-7
    e = 2 * u * x2.sqrt() * c2.max().sqrt()
    win = 0.25 * 2 * e + 6 * torch.ldexp(torch.ones(n), torch.frexp(x2 + c2.max() + 2 * x2.sqrt() * c2.max().sqrt())[1] - 24)
    cand = S <= (S.min(-1, keepdim=True).values + win[:, None])
    assert cand.sum(-1).max() <= 32 * 8, "candidate list capacity of the kernel"
    # pass 2: exact re-score of the candidates with the reference's fp32 expression, lowest index on ties
    xy = (x.double() @ E.double().T).float()
    d = ((x2[:, None] + c2[None]) + xy * -2).clamp(min=0).sqrt()
    d = torch.where(cand, d, torch.full_like(d, float("inf")))
    idx = (-d).argmax(-1)
    assert torch.equal(idx, ref)
    # a correctly rounded x2 would NOT reproduce the reference on W0 (the reason the kernel follows ATen's order)
    if variant == "W0":
        x2e = (x.double() ** 2).sum(-1).float()
        de = ((x2e[:, None] + c2[None]) + xy * -2).clamp(min=0).sqrt()
        assert ((-de).argmax(-1) == ref).float().mean() < 1.0
