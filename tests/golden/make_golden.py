"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on CPU behind oracle/shims.

Run once in the dev container (the GPU box has no reference):   python tests/golden/make_golden.py
Inputs are regenerated from numpy Philox streams (tests/golden/inputs.py) so only reference OUTPUTS and the
small mel input are stored.  Weights: oracle/weights.py W0 / W1 (loaded into the reference with load_state_dict).

Files
  e2e_{W0,W1}.npz : mel (2,128,40) -> reference encoder / quantizer / generator outputs at every stage boundary
  vq_{W0,W1}.npz  : reference EuclideanCodebook.forward indices for N=4096 synthetic rows (bf16-exact x and fp32 x)
                    against the full 32768 x 3584 codebook, plus the fp64 top-2 relative gap per row
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, weights  # noqa: E402
from oracle.restatement import top2_gap  # noqa: E402
from tests.golden.inputs import make_mel, make_vq_rows  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(os.cpu_count())
    for variant in ("W0", "W1"):
        t0 = time.time()
        sd = weights.make_state_dict(variant)
        codec = ref_loader.build_reference_codec(sd)
        print(variant, "reference built", round(time.time() - t0, 1), "s", "checksum", weights.checksum(sd)[:16])
        mel = make_mel(2, 40)
        out = ref_loader.run_reference(codec, mel)
        np.savez_compressed(
            os.path.join(OUT, f"e2e_{variant}.npz"),
            mel=mel.numpy(), enc=out["enc"].numpy(), x_pjt_in=out["x_pjt_in"].numpy(),
            codes=out["codes"].numpy().astype(np.int32), quantized=out["quantized"].numpy(),
            z_dec=out["z_dec"].numpy(), wav=out["wav"].numpy(),
            losses=np.array([float(out["total_loss"]), float(out["commitment_loss"]),
                             float(out["codebook_diversity_loss"])], np.float32),
            weights_checksum=np.array(weights.checksum(sd)))
        # VQ-only: call the reference's EuclideanCodebook.forward (vector_quantize_pytorch.py:462-538) directly
        cb = codec.quantizer.grvq.rvqs[0].layers[0]._codebook
        res = {}
        for kind in ("bf16", "fp32"):
            x = make_vq_rows(4096, kind=kind, scale=0.42)
            with torch.no_grad():
                idx = torch.cat([cb(x[i:i + 1024][None])[1][0] for i in range(0, x.shape[0], 1024)])
            gap = top2_gap(x, cb.embed[0])
            res[f"codes_{kind}"] = idx.numpy().astype(np.int32)
            res[f"gap_{kind}"] = gap.float().numpy()
            print(variant, kind, "vq rows done", round(time.time() - t0, 1), "s; median gap", float(gap.median()))
        np.savez_compressed(os.path.join(OUT, f"vq_{variant}.npz"), **res)


if __name__ == "__main__":
    main()
