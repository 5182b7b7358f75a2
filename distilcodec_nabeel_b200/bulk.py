"""Host-side post-processing around the hot path for bulk work.

`DistilCodec.encode` (distil_codec.py:545-573) follows the quantizer with, per clip: a Python loop over every token
that formats `str(code)` and walks two dict levels (`audio_tokenize`, :532-543), and two synchronous pageable
device->host copies of (T, 3584) fp32 features (806 MB each for a 10-minute clip).  Once the three modules run on
the B200 kernels this tail costs more than the model.  `encode` below returns EXACTLY what the reference method returns
(same GRVQResult fields, the same token dict objects in `codes_list`, the same CPU tensors in the two feature lists,
same `gen_time_lengths` / `n_hop_lengths`), but
  * maps codes to token entries through a per-(group, residual) list built once per codec (no str(), no f-string,
    one list index per token);
  * moves the codes to the host once for the whole batch and the features through pinned buffers with asynchronous
    copies and ONE synchronisation;
  * can skip the feature lists (`features=False`) when the caller only wants tokens: the lists are then empty,
    everything else is unchanged.
It calls the codec's own `preprocess_*`, `encoder` and `quantizer`, so it works on a patched or an unpatched codec.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

_LUT_ATTR = "_b200_token_luts"


def token_luts(codec) -> dict:
    """{(g, r): [entry for code 0, entry for code 1, ...]} with the reference's own dict objects
    (`gr_audio_code2token[f'g{g}r{r}']['audio_code_token'][str(code)]`, distil_codec.py:200-221,540)."""
    luts = getattr(codec, _LUT_ATTR, None)
    if luts is not None and luts[0] is codec.gr_audio_code2token:
        return luts[1]
    out = {}
    for key, val in codec.gr_audio_code2token.items():
        if not (isinstance(val, dict) and "audio_code_token" in val):
            continue
        g, r = key[1:].split("r")
        table = val["audio_code_token"]
        out[(int(g), int(r))] = [table[str(n)] for n in range(len(table))]
    setattr(codec, _LUT_ATTR, (codec.gr_audio_code2token, out))
    return out


def tokenize_codes(codec, codes_gtr) -> list:
    """`audio_tokenize` (distil_codec.py:532-543) for one clip.  codes_gtr: integer array-like (G, T, R) of codebook
    indices.  Order of the result: frame-major, then group, then residual level, as the reference flattens them."""
    luts = token_luts(codec)
    G, T, R = codes_gtr.shape
    if G == 1 and R == 1:
        lut = luts[(0, 0)]
        return [lut[c] for c in codes_gtr.reshape(-1).tolist()]
    cols = [[luts[(g, r)][c] for c in codes_gtr[g, :, r].tolist()] for g in range(G) for r in range(R)]
    out: list = []
    for frame in zip(*cols):
        out.extend(frame)
    return out


def _to_host_async(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        return t.contiguous().cpu()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


def encode(codec, audio_pathes: list, enable_bfloat16: bool = False, raw_audio: bool = False,
           features: bool = True) -> Tuple[object, List[float], List[int]]:
    """Result-identical, faster `DistilCodec.encode`.  Returns (GRVQResult, gen_time_lengths, n_hop_lengths)."""
    if raw_audio:
        _, mel_specs, gen_time_lengths, n_hop_lengths = codec.preprocess_raw_audio_batch(audio_pathes)
    else:
        _, mel_specs, gen_time_lengths, n_hop_lengths = codec.preprocess_audio_batch(audio_pathes=audio_pathes)
    with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=enable_bfloat16):  # as :550
        ret = codec.quantizer(codec.encoder(mel_specs))
    codes_host = _to_host_async(ret.codes)                      # (G, B, T, R) int64, one copy for the whole batch
    pending = []
    if features:
        for b, hop in enumerate(n_hop_lengths):
            # the reference's reshape(hop, 2, -1).reshape(hop * 2, -1) of a contiguous (hop, D) block is a view
            pj = ret.x_pjt_in[b, :hop, :].reshape(hop * 2, -1)
            fu = ret.quantized_fup[b, :hop, :].reshape(hop * 2, -1)
            pending.append((_to_host_async(pj), _to_host_async(fu)))
    if ret.codes.is_cuda:
        torch.cuda.current_stream(ret.codes.device).synchronize()
    codes_np = codes_host.numpy()
    for b, hop in enumerate(n_hop_lengths):
        ret.codes_list.append(tokenize_codes(codec, codes_np[:, b, :hop, :]))
        if features:
            ret.x_pjt_in_list.append(pending[b][0])
            ret.quantized_fup_list.append(pending[b][1])
    return ret, gen_time_lengths, n_hop_lengths
