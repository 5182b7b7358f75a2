"""CPU restatement (the ORACLE) of the DistilCodec inference hot path in plain torch fp32 functional ops.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (distilcodec_nabeel_b200) never does and has no CPU fallback.

Every function follows one piece of the reference (paths relative to /root/reference/distilcodec) and is pinned
against the real reference modules by tests/test_oracle.py (in the dev container, where the reference imports)
and against committed outputs of the reference in tests/golden/ (everywhere).

All functions take the flat state_dict of oracle/weights.py (keys `encoder.* / quantizer.* / generator.*`).
Layout is the reference's: activations (B, C, T) channels-first.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ blocks

def layer_norm_channels_first(x, w, b, eps=1e-6):
    """models/convnext_utils.py:208-213 — manual LN over dim 1 of (B,C,T), biased variance."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    xn = (x - u) / torch.sqrt(s + eps)
    return w[:, None] * xn + b[:, None]


def convnext_block(sd, p, x):
    """models/convnext_utils.py:263-282 — dwconv k7 -> (N,L,C) -> LN -> Linear(C,4C) -> exact GELU ->
    Linear(4C,C) -> gamma* -> (N,C,L) -> + input.  DropPath is identity in eval (:176-183)."""
    C = x.shape[1]
    h = F.conv1d(x, sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=C)
    h = h.permute(0, 2, 1)
    h = F.layer_norm(h, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    h = F.linear(h, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"])
    h = sd[p + "gamma"] * h
    return x + h.permute(0, 2, 1)


def weight_norm_weight(sd, p):
    """torch.nn.utils.parametrizations.weight_norm (dim=0): w = g * v / ||v||, norm over all dims but 0.
    Keys `<p>parametrizations.weight.original0/1` (models/generators.py:50,70,106; convnext_utils.py:36-102)."""
    g = sd[p + "parametrizations.weight.original0"]
    v = sd[p + "parametrizations.weight.original1"]
    return torch._weight_norm(v, g, 0)


# ------------------------------------------------------------------------------------------------ encoder

def encoder_forward(sd, mel, depths=(3, 3, 9, 3), prefix="encoder."):
    """models/encoders.py:68-76.  mel (B,128,T) -> (B,1024,T)."""
    p = prefix
    x = F.conv1d(mel, sd[p + "downsample_layers.0.0.weight"], sd[p + "downsample_layers.0.0.bias"], padding=3)
    x = layer_norm_channels_first(x, sd[p + "downsample_layers.0.1.weight"], sd[p + "downsample_layers.0.1.bias"])
    for s, n in enumerate(depths):
        if s > 0:
            x = layer_norm_channels_first(x, sd[p + f"downsample_layers.{s}.0.weight"],
                                          sd[p + f"downsample_layers.{s}.0.bias"])
            x = F.conv1d(x, sd[p + f"downsample_layers.{s}.1.weight"], sd[p + f"downsample_layers.{s}.1.bias"])
        for j in range(n):
            x = convnext_block(sd, p + f"stages.{s}.{j}.", x)
    return layer_norm_channels_first(x, sd[p + "norm.weight"], sd[p + "norm.bias"])


# ------------------------------------------------------------------------------------------------ VQ

def cdist_neg(x, embed):
    """vector_quantization/utils/vector_quantize_pytorch.py:41-45 and :496 — the exact fp32 expression
    dist = -sqrt(clamp((x2 + y2) + (-2 * x.y), 0)).   x (N,D) fp32, embed (K,D) fp32 -> (N,K)."""
    x2 = (x ** 2).sum(-1)
    y2 = (embed ** 2).sum(-1)
    xy = torch.einsum("id,jd->ij", x, embed) * -2
    return -(x2[:, None] + y2[None, :] + xy).clamp(min=0).sqrt()


def vq_search(x, embed, chunk=2048):
    """EuclideanCodebook.forward eval path (vector_quantize_pytorch.py:462-538): x.float() :473, dist :496,
    argmax (first max wins) via gumbel_sample eval branch :96.  x (N,D) -> int64 (N,)."""
    x = x.float()
    out = []
    for i in range(0, x.shape[0], chunk):
        out.append(cdist_neg(x[i:i + chunk], embed).argmax(dim=-1))
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.long)


def quantizer_pre(sd, enc, prefix="quantizer."):
    """grfvq.py:107 `downsample` = Conv1d(k=1,s=1) + ConvNeXtBlock, then `.mT` -> z (B,T,1024)."""
    p = prefix
    h = F.conv1d(enc, sd[p + "downsample.0.0.weight"], sd[p + "downsample.0.0.bias"])
    h = convnext_block(sd, p + "downsample.0.1.", h)
    return h.mT


def quantizer_post(sd, q_down, prefix="quantizer."):
    """grfvq.py:109 / :144 `upsample` = ConvTranspose1d(k=1,s=1) + ConvNeXtBlock on (B,1024,T)."""
    p = prefix
    h = F.conv_transpose1d(q_down, sd[p + "upsample.0.0.weight"], sd[p + "upsample.0.0.bias"])
    return convnext_block(sd, p + "upsample.0.1.", h)


def quantizer_forward(sd, enc, prefix="quantizer."):
    """DownsampleGRVQ.forward (grfvq.py:105-132) -> GroupedResidualVQ.forward (utils/residual_vq.py:305-356)
    -> ResidualVQ.forward (:140-259) with groups=1, num_quantizers=1, eval mode.
    Returns dict(quantized (B,1024,T), codes (1,B,T,1) int64, quantized_fup (B,T,3584), x_pjt_in (B,T,3584))."""
    p = prefix + "grvq.rvqs.0."
    z = quantizer_pre(sd, enc, prefix)                                               # (B,T,1024)
    x = F.linear(z, sd[p + "project_in.weight"], sd[p + "project_in.bias"])          # residual_vq.py:152
    embed = sd[p + "layers.0._codebook.embed"][0]
    B, T, D = x.shape
    idx = vq_search(x.reshape(B * T, D), embed).reshape(B, T)
    fup = embed[idx]                                                                 # batched_embedding :243-247
    q_down = F.linear(fup, sd[p + "project_out.weight"], sd[p + "project_out.bias"])  # residual_vq.py:241
    quantized = quantizer_post(sd, q_down.mT, prefix)
    return {"quantized": quantized, "codes": idx.reshape(1, B, T, 1), "quantized_fup": fup, "x_pjt_in": x, "z": z}


def quantizer_decode(sd, codes, prefix="quantizer."):
    """DownsampleGRVQ.decode (grfvq.py:141-146): codes (G=1,B,T,R=1) -> gather (residual_vq.py:103-133, einx
    get_at :123), sum over q, project_out (:135-138), upsample."""
    p = prefix + "grvq.rvqs.0."
    embed = sd[p + "layers.0._codebook.embed"][0]
    idx = codes[0, :, :, 0]
    q_down = F.linear(embed[idx], sd[p + "project_out.weight"], sd[p + "project_out.bias"])
    return quantizer_post(sd, q_down.mT, prefix)


# ------------------------------------------------------------------------------------------------ decoder

def resblock1(sd, p, x, k, dilations=(1, 3, 5)):
    """models/convnext_utils.py:106-113: 3 x [silu -> conv(d) -> silu -> conv(1) -> + x]."""
    for n, d in enumerate(dilations):
        w1 = weight_norm_weight(sd, p + f"convs1.{n}.")
        w2 = weight_norm_weight(sd, p + f"convs2.{n}.")
        xt = F.silu(x)
        xt = F.conv1d(xt, w1, sd[p + f"convs1.{n}.bias"], dilation=d, padding=(k * d - d) // 2)
        xt = F.silu(xt)
        xt = F.conv1d(xt, w2, sd[p + f"convs2.{n}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def parallel_block(sd, p, x, kernel_sizes=(3, 7, 11)):
    """models/convnext_utils.py:137-138: mean over the 3 ResBlock1 branches."""
    return torch.stack([resblock1(sd, p + f"blocks.{b}.", x, k) for b, k in enumerate(kernel_sizes)], 0).mean(0)


def generator_forward(sd, z, rates=(8, 4, 2, 2, 2), ksizes=(16, 12, 4, 4, 4), prefix="generator.",
                      return_stages=False):
    """HiFiGANGenerator.forward (models/generators.py:118-147), use_template=False.  (B,1024,T) -> (B,1,256T)."""
    p = prefix
    x = F.conv1d(z, weight_norm_weight(sd, p + "conv_pre."), sd[p + "conv_pre.bias"], padding=6)
    stages = [x]
    for i, (u, k) in enumerate(zip(rates, ksizes)):
        x = F.silu(x)
        x = F.conv_transpose1d(x, weight_norm_weight(sd, p + f"ups.{i}."), sd[p + f"ups.{i}.bias"],
                               stride=u, padding=(k - u) // 2)
        x = parallel_block(sd, p + f"resblocks.{i}.", x)
        stages.append(x)
    x = F.silu(x)
    x = F.conv1d(x, weight_norm_weight(sd, p + "conv_post."), sd[p + "conv_post.bias"], padding=6)
    x = torch.tanh(x)
    return (x, stages) if return_stages else x


# ------------------------------------------------------------------------------------------------ whole path

@torch.no_grad()
def codec_forward(sd, mel):
    """DistilCodec.forward minus the CPU front-end (distil_codec.py:518-530): mel -> enc -> VQ -> wav."""
    enc = encoder_forward(sd, mel)
    q = quantizer_forward(sd, enc)
    wav = generator_forward(sd, q["quantized"])
    return {"enc": enc, **q, "wav": wav}


# ------------------------------------------------------------------------------------------------ front-end

def log_mel(audio, n_fft=1024, hop=256, win=1024, n_mels=128, sr=24000, f_min=0.0, f_max=12000.0):
    """models/mel_spec.py:26-57,100-122 + distil_codec.py:134-138: audio (B,1,n+1) (already left-padded by one
    zero sample) -> log-mel (B,128,T).  Uses torchaudio's slaney filterbank like the reference (:85-93)."""
    import torchaudio
    y = audio.squeeze(1)
    y = F.pad(y.unsqueeze(1), ((win - hop) // 2, (win - hop + 1) // 2), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec)
    lin = torch.sqrt(spec.pow(2).sum(-1) + 1e-6)
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sr, "slaney", "slaney")
    mel = (lin.transpose(-1, -2) @ fb).transpose(-1, -2)
    return torch.log(torch.clamp(mel, min=1e-5))


def top2_gap(x, embed, chunk=1024):
    """Relative gap between the two smallest distances per row, for BASELINE's 'mismatch only where the top-2
    gap is < 1e-3 relative' clause.  fp64 to be independent of the fp32 rounding under test."""
    gaps = []
    e = embed.double()
    e2 = (e ** 2).sum(-1)
    for i in range(0, x.shape[0], chunk):
        xx = x[i:i + chunk].double()
        d2 = ((xx ** 2).sum(-1)[:, None] + e2[None] - 2 * xx @ e.T).clamp(min=0).sqrt()
        t = d2.topk(2, dim=-1, largest=False).values
        gaps.append((t[:, 1] - t[:, 0]) / t[:, 1].clamp(min=1e-30))
    return torch.cat(gaps)
