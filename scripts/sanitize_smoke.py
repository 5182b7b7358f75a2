"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck), tiny shapes, small codebook."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distilcodec_nabeel_b200 import Engine, load_config, mel_buffers
from distilcodec_nabeel_b200 import random_init as weights
from tests.golden.inputs import make_mel, make_wav

sd = dict(weights.make_state_dict("W1", codebook_size=1024))
sd.update(mel_buffers(load_config()))
for mode in ("bf16", "fp32"):
    eng = Engine(sd, 0, mode)
    wav = torch.nn.functional.pad(make_wav(2, 256 * 5 + 17), (1, 0)).cuda().contiguous()
    mel = eng.mel(wav)
    enc = eng.encoder(mel)
    codes, xin, fup, quant = eng.quantizer(enc)
    z = eng.decode_codes(codes)
    wavo = eng.generator(quant)
    torch.cuda.synchronize()
    print(mode, "ok", tuple(mel.shape), tuple(wavo.shape), int(codes.sum()), float(wavo.abs().max()))
    eng.close()
