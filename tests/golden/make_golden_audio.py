"""Generate tests/golden/audio_{W0,W1}.npz: the REAL reference's public API on bundled real audio.

BASELINE configs[0] asks for `test.mp3`; no MP3 decoder exists in this image, so the bundled
24 kHz clip data/org_audios/0001.wav (first 3 s) stands in.  Run once in the dev container:
    python tests/golden/make_golden_audio.py
Calls, unchanged: DistilCodec.encode(raw_audio=True) (distil_codec.py:545-573; CPU mel front-end :99-145) and the
decode path quantizer.decode + generator (:591-592; decode_from_codes itself hard-codes .cuda()).
"""
import os
import sys
import wave

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    w = wave.open(os.path.join(ref_loader.REFERENCE_ROOT, "data", "org_audios", "0001.wav"))
    assert w.getframerate() == 24000 and w.getnchannels() == 1 and w.getsampwidth() == 2
    pcm = np.frombuffer(w.readframes(3 * 24000), dtype=np.int16).astype(np.float32) / 32768.0
    for variant in ("W0", "W1"):
        sd = weights.make_state_dict(variant)
        codec = ref_loader.build_reference_codec(sd)
        with torch.no_grad():
            _, mel, _, n_hop = codec.preprocess_raw_audio_batch([[pcm, 24000]])
            r = codec.encode([[pcm, 24000]], enable_bfloat16=False, raw_audio=True)
            codes = r["quantized_ret"].codes if isinstance(r, dict) else r[0].codes
            wav = codec.generator(codec.quantizer.decode(codes))
        np.savez_compressed(os.path.join(OUT, f"audio_{variant}.npz"), pcm=pcm, mel=mel.numpy(),
                            codes=codes.numpy().astype(np.int32), wav=wav.numpy(), n_hop=np.array(n_hop))
        print(variant, "mel", tuple(mel.shape), "codes", tuple(codes.shape), "wav", tuple(wav.shape), "n_hop", n_hop)


if __name__ == "__main__":
    main()
