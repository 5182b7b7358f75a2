"""Per-role timeline of conv_tsw_kernel (experiment build with -DDC_TSW_TRACE, DC_LIB pointing at it) for ONE layer run
through the dc_op_conv_gemm hook: python scripts/tsw_trace.py C N J dil res(0/1)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from distilcodec_nabeel_b200 import Engine
from distilcodec_nabeel_b200 import random_init as weights

Cc, N, J, dil, res = (int(v) for v in sys.argv[1:6])
act = int(sys.argv[6]) if len(sys.argv) > 6 else 0      # 2 = SiLU (conv1-type epilogue when res = 0)
eng = Engine(weights.make_state_dict("W0"), 0, "bf16")
B, T = 64, 937 * {128: 64, 256: 32}.get(Cc, 8)
a = torch.randn(B, T, Cc, device="cuda")
w = torch.randn(N, J * Cc, device="cuda") * 0.02
r = torch.randn(B, T, N, device="cuda") if res else None
for _ in range(2):
    eng.op_conv_gemm(a, w, None, r, -dil * (J - 1) // 2, dil, act)
torch.cuda.synchronize()
lib = C.CDLL(os.environ["DC_LIB"])
buf = (C.c_longlong * (12 * 64))()
assert lib.dc_debug_tsw_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(12, 64)
print("tile period:", np.diff(t[0, 8:28]).tolist())
print("mma: wait tempty", (t[1] - t[0])[8:28].tolist())
print("mma: issue span ", (t[2] - t[1])[8:28].tolist())
print("mma: of which wait afull", t[3, 8:28].tolist())
print("mma: of which wait bfull", t[4, 8:28].tolist())
for g in (0, 1):
    sel = [i for i in range(8, 28) if t[6, i] != 0]
    break
print("epi (group of tile): wait tfull", [(int(t[6, i] - t[5, i])) for i in range(8, 28) if t[6, i]])
print("epi: work            ", [(int(t[7, i] - t[6, i])) for i in range(8, 28) if t[6, i]])
