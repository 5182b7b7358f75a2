"""Per-kernel timing of one stage at the bench size (256 x 10 s): python scripts/bench_stage.py [encoder|generator] [filter]
Uses the engine's event profiler; prints every kernel row whose name contains `filter`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distilcodec_nabeel_b200 import Engine
from distilcodec_nabeel_b200 import random_init as weights
from tests.golden.inputs import make_mel

stage = sys.argv[1] if len(sys.argv) > 1 else "encoder"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
B, T = int(os.environ.get("CLIPS", 256)), 937
eng = Engine(weights.make_state_dict("W0"), 0, "bf16")
mel = make_mel(8, T, seed=1).repeat(B // 8, 1, 1).cuda()
x = mel if stage == "encoder" else torch.randn(B, T, 1024, device="cuda")
fn = eng.encoder if stage == "encoder" else eng.generator
for _ in range(3):
    fn(x)
torch.cuda.synchronize()
eng.profile(True)
for _ in range(3):
    fn(x)
torch.cuda.synchronize()
rows = eng.profile_rows()
eng.profile(False)
tot = 0.0
for r in sorted(rows, key=lambda r: -r["ms"]):
    if flt in r["name"]:
        ms = r["ms"] / 3
        tot += ms
        print(f"{r['name']:48s} n={r['launches'] // 3:3d} {ms:8.3f} ms  {r['flops'] / 3 / ms / 1e9 if ms else 0:8.1f} TF/s  "
              f"{r['bytes'] / 3 / ms / 1e6 if ms else 0:8.1f} GB/s")
print(f"total {tot:.3f} ms")
