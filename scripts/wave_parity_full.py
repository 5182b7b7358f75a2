"""Latent / waveform parity at the real clip length (10 s = 937 frames), B200 path vs the oracle on the host CPU,
"evaluated on the same codes" as BASELINE states it.  Prints one JSON object per (weights, mode).
usage: python scripts/wave_parity_full.py [clips]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distilcodec_nabeel_b200 import Engine
from distilcodec_nabeel_b200 import random_init
from oracle import restatement as R
from tests.golden.inputs import make_mel

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.set_num_threads(os.cpu_count() or 1)


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())


def mabs(a, b):
    return float((a.double().cpu() - b.double()).abs().max())


mel = make_mel(clips, 937, seed=77)
for variant in ("W0", "W1"):
    sd = random_init.make_state_dict(variant)
    with torch.no_grad():
        ref_enc = R.encoder_forward(sd, mel)                                   # (B,1024,T)
    for mode in ("bf16", "fp32"):
        eng = Engine(sd, 0, mode)
        enc = eng.encoder(mel.cuda())
        codes, xin, fup, quant = eng.quantizer(enc)
        wav = eng.generator(quant)
        torch.cuda.synchronize()
        with torch.no_grad():
            z_ref = R.quantizer_decode(sd, codes.cpu()[None, :, :, None])      # the same codes
            w_ref = R.generator_forward(sd, z_ref)[:, 0]
        print(json.dumps({"weights": variant, "mode": mode, "clips": clips, "frames_per_clip": 937,
                          "encoder_latent_rel_to_max": rel(enc.transpose(1, 2), ref_enc),
                          "quantized_latent_rel_to_max": rel(quant.transpose(1, 2), z_ref),
                          "waveform_rel_to_max": rel(wav, w_ref), "waveform_max_abs": mabs(wav, w_ref),
                          "waveform_peak": float(w_ref.abs().max())}))
        eng.close()
