"""Index parity at BASELINE configs[1] size: 64 clips x 10 s (59,968 frames) through encoder + quantizer on the B200,
the nearest-code search re-done by the oracle (CPU, fp32, the reference's exact expression) on the SAME project_in
rows.  Prints one JSON object; usage: python scripts/vq_parity_full.py W0|W1 [clips]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distilcodec_nabeel_b200 import Engine
from distilcodec_nabeel_b200 import random_init
from oracle import restatement as R
from tests.golden.inputs import make_mel

variant = sys.argv[1] if len(sys.argv) > 1 else "W0"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.set_num_threads(os.cpu_count() or 1)
sd = random_init.make_state_dict(variant)
eng = Engine(sd, 0, "bf16")
mel = make_mel(clips, 937, seed=2024)
enc = eng.encoder(mel.cuda())
codes, xin, _, _ = eng.quantizer(enc)
torch.cuda.synchronize()
x = xin.float().cpu().reshape(-1, xin.shape[-1])
E = sd["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
t = time.time()
ref = R.vq_search(x, E)
cpu_s = time.time() - t
got = codes.cpu().reshape(-1)
bad = (got != ref).nonzero().reshape(-1)
out = {"weights": variant, "frames": int(got.numel()), "identical": int((got == ref).sum()),
       "fraction": float((got == ref).float().mean()), "oracle_cpu_seconds": round(cpu_s, 1), "mismatches": []}
if bad.numel():
    gaps = R.top2_gap(x[bad], E)
    out["max_top2_gap_of_mismatches"] = float(gaps.max())
    out["mismatches"] = [{"row": int(i), "b200": int(got[i]), "oracle": int(ref[i]), "top2_rel_gap": float(g)}
                         for i, g in list(zip(bad.tolist(), gaps.tolist()))[:20]]
print(json.dumps(out))
