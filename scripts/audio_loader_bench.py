"""Host-side throughput of the native audio loader (row f-4): N synthetic WAV files (44.1 kHz stereo s16) decoded,
down-mixed and resampled to 24 kHz in parallel into one pinned batch.  Prints one JSON object.
usage: python scripts/audio_loader_bench.py [files] [seconds] [threads]"""
import json
import os
import sys
import tempfile
import time
import wave

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from distilcodec_nabeel_b200 import audio

files = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with tempfile.TemporaryDirectory() as d:
    rng = np.random.default_rng(0)
    base = (rng.standard_normal((int(44100 * secs), 2)) * 3000).astype("<i2")
    paths = []
    for i in range(files):
        p = os.path.join(d, f"{i}.wav")
        w = wave.open(p, "wb")
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(np.roll(base, i * 1000, 0).tobytes())
        w.close()
        paths.append(p)
    out = {}
    for name, q in (("scipy_default", audio.SCIPY), ("hq", audio.HQ)):
        audio.load_batch(paths[:2], 24000, quality=q, threads=threads, pin=False)
        t0 = time.perf_counter()
        batch, lengths, st = audio.load_batch(paths, 24000, quality=q, threads=threads, pin=False)
        dt = time.perf_counter() - t0
        out[name] = {"seconds": round(dt, 3), "audio_s_per_s": round(files * secs / dt, 1)}
    t0 = time.perf_counter()
    audio.load_batch(paths, 24000, quality=audio.HQ, threads=1, pin=False)
    out["hq_one_thread"] = {"audio_s_per_s": round(files * secs / (time.perf_counter() - t0), 1)}
print(json.dumps({"files": files, "seconds_each": secs, "format": "44.1 kHz stereo s16 -> 24 kHz mono f32",
                  "host_threads": threads or os.cpu_count(), **out}))
