"""Accuracy probe for an fp32-accurate tensor-core GEMM: 3-way bf16 split of both operands, the 6 largest cross terms,
K-extended operands through the existing bf16 kernels (dc_op_conv_gemm).  Compares with a float64 reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distilcodec_nabeel_b200 import Engine
from distilcodec_nabeel_b200 import random_init as weights


def split3(x):
    h = x.bfloat16().float()
    r = x - h
    m = r.bfloat16().float()
    l = (r - m).bfloat16().float()
    return h, m, l


sd = weights.make_state_dict("W0")
e16, e32 = Engine(sd, 0, "bf16"), Engine(sd, 0, "fp32")
torch.manual_seed(0)
for (B, T, C, N, J, dil) in ((4, 1024, 512, 256, 1, 1), (2, 2048, 256, 256, 11, 3), (2, 4096, 64, 64, 7, 1), (1, 4096, 1024, 1024, 1, 1)):
    a = torch.randn(B, T, C, device="cuda")
    w = torch.randn(N, J * C, device="cuda") * 0.05
    shift0 = -dil * (J - 1) // 2
    # float64 reference
    a64 = torch.nn.functional.pad(a.double(), (0, 0, -shift0, dil * (J - 1) + shift0))
    ref = torch.zeros(B, T, N, dtype=torch.float64, device="cuda")
    for j in range(J):
        ref += a64[:, j * dil:j * dil + T] @ w[:, j * C:(j + 1) * C].double().T
    ah, am, al = split3(a)
    wh, wm, wl = split3(w.reshape(N, J, C))
    A6 = torch.cat([ah, ah, am, ah, al, am], -1).contiguous()                  # (B,T,6C)
    W6 = torch.cat([wh, wm, wh, wl, wh, wm], -1).reshape(N, J * 6 * C).contiguous()
    A3 = torch.cat([ah, ah, am], -1).contiguous()
    W3 = torch.cat([wh, wm, wh], -1).reshape(N, J * 3 * C).contiguous()
    outs = {
        "bf16 1-term": e16.op_conv_gemm(a, w, None, None, shift0, dil, 0),
        "split 3-term": e16.op_conv_gemm(A3, W3, None, None, shift0, dil, 0),
        "split 6-term": e16.op_conv_gemm(A6, W6, None, None, shift0, dil, 0),
        "cuda-core fp32": e32.op_conv_gemm(a, w, None, None, shift0, dil, 0),
    }
    print(f"B{B} T{T} C{C} N{N} J{J} d{dil}: " + "  ".join(
        f"{k}: {float((v.double() - ref).abs().max() / ref.abs().max()):.2e}" for k, v in outs.items()))
