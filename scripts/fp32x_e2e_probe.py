"""End-to-end accuracy probe for an fp32-accurate TENSOR-CORE mode: every dense layer of the generator and the encoder as
a K-extended bf16 GEMM over 3-way bf16 splits of both operands (the 6 largest cross terms, or the 3 largest), run through
the existing tcgen05 kernels (dc_op_conv_gemm with 6C / 3C channels); everything else (LayerNorm, depthwise conv, SiLU,
ConvTranspose) in torch fp32 on the device.  Compared with the fp32 oracle on the CPU: does the chain keep the 1e-4 gate?
usage: python scripts/fp32x_e2e_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from distilcodec_nabeel_b200 import Engine
from oracle import restatement as R
from oracle import weights
from tests.golden.inputs import make_latents, make_mel

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)


def split3(x):
    h = x.bfloat16().float()
    r = x - h
    m = r.bfloat16().float()
    l = (r - m).bfloat16().float()
    return h, m, l


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())


class X:
    """conv / linear through the bf16 tensor-core kernel on split operands"""

    def __init__(self, eng, terms):
        self.eng, self.terms = eng, terms

    def conv(self, x_ncl, w_oik, bias, dil=1, pad=0):
        a = x_ncl.transpose(1, 2).contiguous()                     # (B,T,C)
        O, I, k = w_oik.shape
        ah, am, al = split3(a)
        wh, wm, wl = split3(w_oik.permute(0, 2, 1).contiguous())   # (O,k,I)
        if self.terms == 6:
            A = torch.cat([ah, ah, am, ah, al, am], -1).contiguous()
            W = torch.cat([wh, wm, wh, wl, wh, wm], -1)
        elif self.terms == 3:
            A = torch.cat([ah, ah, am], -1).contiguous()
            W = torch.cat([wh, wm, wh], -1)
        else:
            A, W = ah.contiguous(), wh
        W = W.reshape(O, -1).contiguous()
        y = self.eng.op_conv_gemm(A, W, bias, None, -pad, dil, 0)
        return y.transpose(1, 2)

    def linear(self, x_nlc, w, b):
        return self.conv(x_nlc.transpose(1, 2), w[:, :, None], b).transpose(1, 2)


def generator(sd, z, X_):
    p = "generator."
    wn = lambda q: R.weight_norm_weight(sd, q)
    x = X_.conv(z, wn(p + "conv_pre."), sd[p + "conv_pre.bias"], 1, 6)
    for i, (u, k) in enumerate(zip((8, 4, 2, 2, 2), (16, 12, 4, 4, 4))):
        x = F.conv_transpose1d(F.silu(x), wn(p + f"ups.{i}."), sd[p + f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        outs = []
        for b, kk in enumerate((3, 7, 11)):
            xb = x
            for n, d in enumerate((1, 3, 5)):
                q = p + f"resblocks.{i}.blocks.{b}."
                xt = X_.conv(F.silu(xb), wn(q + f"convs1.{n}."), sd[q + f"convs1.{n}.bias"], d, (kk * d - d) // 2)
                xt = X_.conv(F.silu(xt), wn(q + f"convs2.{n}."), sd[q + f"convs2.{n}.bias"], 1, (kk - 1) // 2)
                xb = xt + xb
            outs.append(xb)
        x = torch.stack(outs, 0).mean(0)
    x = F.conv1d(F.silu(x), wn(p + "conv_post."), sd[p + "conv_post.bias"], padding=6)
    return torch.tanh(x)


def encoder(sd, mel, X_):
    p = "encoder."
    x = X_.conv(mel, sd[p + "downsample_layers.0.0.weight"], sd[p + "downsample_layers.0.0.bias"], 1, 3)
    x = R.layer_norm_channels_first(x, sd[p + "downsample_layers.0.1.weight"], sd[p + "downsample_layers.0.1.bias"])
    for s, n in enumerate((3, 3, 9, 3)):
        if s > 0:
            x = R.layer_norm_channels_first(x, sd[p + f"downsample_layers.{s}.0.weight"], sd[p + f"downsample_layers.{s}.0.bias"])
            x = X_.conv(x, sd[p + f"downsample_layers.{s}.1.weight"], sd[p + f"downsample_layers.{s}.1.bias"])
        for j in range(n):
            q = p + f"stages.{s}.{j}."
            C = x.shape[1]
            h = F.conv1d(x, sd[q + "dwconv.weight"], sd[q + "dwconv.bias"], padding=3, groups=C).permute(0, 2, 1)
            h = F.layer_norm(h, (C,), sd[q + "norm.weight"], sd[q + "norm.bias"], 1e-6)
            h = F.gelu(X_.linear(h, sd[q + "pwconv1.weight"], sd[q + "pwconv1.bias"]))
            h = X_.linear(h, sd[q + "pwconv2.weight"], sd[q + "pwconv2.bias"])
            x = x + (sd[q + "gamma"] * h).permute(0, 2, 1)
    return R.layer_norm_channels_first(x, sd[p + "norm.weight"], sd[p + "norm.bias"])


with torch.no_grad():
    for variant in ("W0", "W1"):
        sd_cpu = weights.make_state_dict(variant, codebook_size=1024)
        sd = {k: v.to(dev) for k, v in sd_cpu.items()}
        eng = Engine(sd_cpu, 0, "bf16")
        z = make_latents(2, 48, seed=31) * 0.5
        mel = make_mel(2, 64, seed=5)
        ref_w = R.generator_forward(sd_cpu, z)
        ref_e = R.encoder_forward(sd_cpu, mel)
        for terms in (1, 3, 6):
            X_ = X(eng, terms)
            print(variant, f"{terms}-term split:  generator rel err {rel(generator(sd, z.to(dev), X_), ref_w):.3e}   "
                           f"encoder rel err {rel(encoder(sd, mel.to(dev), X_), ref_e):.3e}", flush=True)
        eng.close()
