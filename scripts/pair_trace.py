"""Per-role timeline of the fused conv pair kernel (experiment build with -DDC_PAIR_TRACE, DC_LIB pointing at it):
runs ONE layer through dc_op-level engine hooks is not available for pairs, so the whole generator runs and the trace of
the LAST pair launch (C = 32, k = 11, d = 5 — the buffer is overwritten by every launch) is printed."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from distilcodec_nabeel_b200 import Engine, _abi
from distilcodec_nabeel_b200 import random_init as weights

eng = Engine(weights.make_state_dict("W0"), 0, "bf16")
x = torch.randn(64, 937, 1024, device="cuda")
for _ in range(2):
    eng.generator(x)
torch.cuda.synchronize()
lib = C.CDLL(os.environ["DC_LIB"])
buf = (C.c_longlong * (16 * 64))()
assert lib.dc_debug_pair_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(16, 64)
names = ["mma:c1 start", "mma:d1empty ok", "mma:afull ok", "mma:c1 issued", "mma:c2 start", "mma:d2empty ok", "mma:tfull ok",
         "mma:c2 issued", "e1:d1full ok", "e1:tempty ok", "e1:done", "e2:wait", "e2:d2full ok", "e2:done", "tma:aempty ok"]
base = t[0, 8]
print("tile period (mma c1 start):", np.diff(t[0, 8:40]).tolist())
for i in range(8, 14):
    print(f"--- tile {i} (relative to conv1({8}) start)")
    for ev in range(15):
        if ev in (11, 12, 13) and t[ev, i] == 0:
            continue
        print(f"   {names[ev]:16s} {int(t[ev, i] - base):8d}")
