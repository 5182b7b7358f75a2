"""GPU diagnostic: where do dc_vq_search's indices differ from the reference's golden indices?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distilcodec_nabeel_b200 import Engine
from distilcodec_nabeel_b200 import random_init as weights
from tests.golden.inputs import make_vq_rows

sd = weights.make_state_dict("W0")
E = sd["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
g = np.load("tests/golden/vq_W0.npz")
eng = Engine(sd, 0, "bf16")
for kind in ("bf16", "fp32"):
    x = make_vq_rows(4096, kind=kind)
    ref = g[f"codes_{kind}"].astype(np.int64)
    xd = x.cuda().to(torch.bfloat16).contiguous() if kind == "bf16" else x.cuda()
    x2 = (x ** 2).sum(-1)
    def run(name, x2t=None, **opts):
        for k, v in opts.items(): eng.set_option(k, v)
        c, st = eng.vq_search(xd, None if x2t is None else x2t.cuda(), stats=True)
        for k in opts: eng.set_option(k, {"vq_window": 0.25, "vq_x2_exact": 0, "vq_tensor_core": 1}[k])
        c = c.cpu().numpy(); bad = np.nonzero(c != ref)[0]
        print(kind, name, "mismatch", len(bad), st, bad[:10].tolist(), flush=True)
        return c, bad
    c0, bad = run("default")
    run("x2 from torch cpu", x2)
    run("x2 exact", vq_x2_exact=1)
    run("window 1.0", vq_window=1.0)
    run("window 0.05", vq_window=0.05)
    run("simt scorer", vq_tensor_core=0)
    c2 = (E ** 2).sum(-1)
    for r in bad[:6]:
        xr = x[r]
        xy = (xr.double()[None] @ E.double().T).float()[0]
        d = ((x2[r] + c2) + xy * -2).clamp(min=0).sqrt()
        S = c2 - 2 * (xr.to(torch.bfloat16).float()[None] @ E.to(torch.bfloat16).float().T)[0]
        print(" row", r, "ref", ref[r], "got", c0[r], "d_ref", d[ref[r]].item(), "d_got", d[c0[r]].item(), "dmin", d.min().item(),
              "S_ref-Smin", (S[ref[r]] - S.min()).item(), "S_got-Smin", (S[c0[r]] - S.min()).item(), "x2", x2[r].item())
