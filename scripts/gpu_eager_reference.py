"""The "library kernel to beat" (SURVEY section 2b / BASELINE.md section 3): the reference's OWN PyTorch modules on the
same B200 through PyTorch eager (cuDNN / cuBLAS / ATen), per stage, in the two numeric settings its API offers:
fp32 (TF32 off, as the parity oracle is defined) and bf16 autocast (what enable_bfloat16=True does,
distil_codec.py:550,590).  Falls back to the oracle restatement when the reference package is not importable.
Prints one JSON object.   usage: python scripts/gpu_eager_reference.py [clips] [seconds]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ref_loader
from oracle import restatement as R
from oracle import weights
from tests.golden.inputs import make_mel

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
T = int(secs * 24000) // 256
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
sd = weights.make_state_dict("W0")
mel = make_mel(clips, T, seed=17).to(dev)
audio_s = clips * T * 256 / 24000

if ref_loader.available():
    codec = ref_loader.build_reference_codec(sd)
    codec.device = dev
    codec.move_to_cuda()
    enc_f, q_f, gen_f = codec.encoder, (lambda e: codec.quantizer(e).quantized), codec.generator
    kind = "reference modules (baseline/_ref)"
else:
    sdd = {k: v.to(dev) for k, v in sd.items()}
    enc_f = lambda m: R.encoder_forward(sdd, m)
    q_f = lambda e: R.quantizer_forward(sdd, e)["quantized"]
    gen_f = lambda z: R.generator_forward(sdd, z)
    kind = "oracle restatement (same ATen ops)"


def timed(fn, x, reps=3):
    with torch.no_grad():
        y = fn(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            y = fn(x)
        torch.cuda.synchronize()
    return y, (time.perf_counter() - t0) / reps * 1e3


out = {"what": "PyTorch eager on the same GPU: " + kind, "clips": clips, "seconds_each": secs, "frames": clips * T,
       "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "tf32": False}
for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
    with ctx:
        enc, t_enc = timed(enc_f, mel)
        q, t_q = timed(q_f, enc)
        wav, t_gen = timed(gen_f, q)
    tot = t_enc + t_q + t_gen
    out[name] = {"encoder_ms": round(t_enc, 2), "quantizer_ms": round(t_q, 2), "generator_ms": round(t_gen, 2),
                 "total_ms": round(tot, 2), "audio_s_per_s": round(audio_s / (tot / 1e3), 1),
                 "wav_dtype": str(wav.dtype), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
print(json.dumps(out))
