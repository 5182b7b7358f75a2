#!/usr/bin/env python
"""bench.py — throughput of the DistilCodec hot path on B200 (BASELINE.json metric: audio-seconds processed per
second, encode+decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--detail-out PATH]

Default workload `recon` = BASELINE configs[3]: one step is one pass of the hot path (log-mel -> ConvNeXt encoder ->
32768x3584 VQ -> HiFiGAN decoder -> wav) over 256 synthetic 10 s clips PER GPU (weak scaling: clips shard across
ranks, no data-path collective).  Weights are random-init of the architecture in configs/model_config.json
(distilcodec_nabeel_b200/random_init.py, W0 — the HF checkpoint is not available offline).

The other BASELINE configs are side workloads with the same line shape (their numbers go to profiles/):
    vq_only        configs[1]  nearest-code search only, 64 x 10 s of synthetic project_in rows (59,968 x 3584)
    wav2codes_30s  configs[2]  wav -> codes of 64 synthetic 30 s clips from HOST wav (GPU mel + encoder + VQ)
    bulk_10min     configs[4]  tokenise + decode of 10-minute clips through the time-tiled legs (per GPU: --clips)

Prints ONE compact JSON line on stdout (rank 0; < 4 kB).  The per-kernel-class and per-layer tables go to a side
file (`--detail-out`, default profiles/bench_detail_<workload>_n<N>.json) whose path is in the line.
  value        whole-job audio-s/s with the input batch already resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-buffer call (pinned host input in, host codes / wav out, copies inside
               the timed region)
  roofline     the dominant kernel class by device time (event pairs recorded inside the library around every
               launch of the timed steps): algorithmic FLOPs or bytes / summed launch time vs MEASURED_PEAKS.json
  per_rank_ms  every rank's own ms/step (value uses the max)
  cpu_baseline the reference's CPU path on a bounded sample (N=1, rank 0): the real reference modules when
               baseline/_ref (or /root/reference) is importable — kind "reference" — else the oracle port
`--impl reference` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "audio_seconds_per_second_encode_decode"
UNIT = "audio-s/s"
HOP, SR = 256, 24000
FPS = SR / HOP
# SURVEY.md section 8d, MFLOP per frame: encoder 154.25 + quantizer dense 52.46 + distance 234.88 + generator 1603.49
MFLOP_PER_FRAME = {"recon": 2045.07, "vq_only": 234.88, "wav2codes_30s": 415.36, "bulk_10min": 415.36 + 1629.7}
WORKLOADS = {
    "recon": "BASELINE configs[3]: full encode->decode reconstruction, {clips} synthetic {secs:g} s clips per GPU",
    "vq_only": "BASELINE configs[1]: VQ-only nearest-code search, {clips} x {secs:g} s of synthetic project_in rows "
               "vs the 32768x3584 codebook",
    "wav2codes_30s": "BASELINE configs[2]: wav->codes (GPU log-mel + encoder + VQ) of {clips} synthetic {secs:g} s "
                     "clips per GPU",
    "bulk_10min": "BASELINE configs[4]: tokenise + codes->wav decode of {clips} synthetic {secs:g} s clips per GPU "
                  "through the time-tiled legs",
}
DEFAULT_SHAPE = {"recon": (256, 10.0), "vq_only": (64, 10.0), "wav2codes_30s": (64, 30.0), "bulk_10min": (4, 600.0)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            # "under load": samples at or above the median power draw
            med_p = sorted(pw)[len(pw) // 2]
            load = sorted(s for s, p in zip(sm, pw) if p >= med_p) or sorted(sm)
            out.update(sm_mhz=load[len(load) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


# ---------------------------------------------------------------------------------------------- the CPU arm
def cpu_reference_rate(workload: str, seconds_of_audio: float, clips: int, steps: int, warmup: int):
    """The reference's CPU implementation of the workload on `clips` x `seconds_of_audio`, fp32, all host threads.
    The real reference modules (baseline/_ref or /root/reference behind oracle/shims) when importable, else the
    oracle port (oracle/restatement.py, pinned to the reference's outputs by tests/).
    -> (audio-s/s, ms/step, cores, frames per clip, kind)"""
    from oracle import ref_loader
    from oracle import restatement as R
    from oracle import weights
    from tests.golden.inputs import make_mel, make_vq_rows
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = weights.make_state_dict("W0")
    T = int(seconds_of_audio * SR) // HOP
    kind = "port"
    codec = None
    if ref_loader.available():
        try:
            codec = ref_loader.build_reference_codec(sd)
            kind = "reference"
        except Exception as e:  # noqa: BLE001 — an unimportable reference falls back to the port, and says so
            print(f"bench.py: reference modules not usable ({e!r}); timing the oracle port", file=sys.stderr)
            codec = None
    E = sd["quantizer.grvq.rvqs.0.layers.0._codebook.embed"][0]
    if workload == "vq_only":
        x = make_vq_rows(clips * T, seed=2)

        def step():
            if codec is not None:   # EuclideanCodebook.forward eval path (vector_quantize_pytorch.py:462-538)
                cb = codec.quantizer.grvq.rvqs[0].layers[0]._codebook
                return cb(x[None])[1]
            return R.vq_search(x, E)
    else:
        mel = make_mel(clips, T, seed=17)

        def step():
            if codec is not None:
                enc = codec.encoder(mel)
                r = codec.quantizer(enc)
                if workload == "wav2codes_30s":
                    return r.codes
                return codec.generator(r.quantized)
            if workload == "wav2codes_30s":
                return R.quantizer_forward(sd, R.encoder_forward(sd, mel))["codes"]
            return R.codec_forward(sd, mel)
    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    audio_s = clips * T * HOP / SR * steps
    return audio_s / dt, dt / steps * 1e3, cores, T, kind


def run_reference_arm(args, rank: int, world: int, out):
    if rank != 0:
        return
    clips = args.cpu_clips or 1
    secs = min(args.seconds, args.cpu_seconds)
    rate, ms, cores, T, kind = cpu_reference_rate(args.workload, secs, clips, args.steps, args.warmup)
    what = "the reference's own PyTorch modules" if kind == "reference" else "oracle port (torch restatement)"
    sample = (f"{clips} clip(s) x {secs:g} s (T={T} frames) per step of the same workload on the host CPU, fp32, "
              f"{cores} threads, {what}")
    cfg = workload_config(args, world)
    cfg["cpu_sample"] = {"clips_per_step": clips, "seconds_per_clip": secs}
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


def workload_config(args, world: int) -> dict:
    T = int(args.seconds * SR) // HOP
    return {"workload": WORKLOADS[args.workload].format(clips=args.clips, secs=args.seconds),
            "clips_per_gpu": args.clips, "frames_per_clip": T, "weights": "random-init W0 of configs/model_config.json",
            "parallelism": f"clip-sharded x{world}, no collective on the data path",
            "l2": "inputs and per-step working set >> 126 MB L2; two alternating input batches"}


def _claim_stdout():
    """Route everything that libraries print on fd 1 (e.g. NCCL's version banner) to stderr and return a file object
    on the real stdout, so that the JSON line is the only thing printed there."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


# ---------------------------------------------------------------------------------------------- workloads (GPU arm)
class Workload:
    """device_step(i) runs one step on HBM-resident inputs; host_step(i) the same from pinned host buffers to pinned
    host results.  `result_checksum()` is a 32-bit digest of the last device step's integer output."""

    def __init__(self, args, eng, pipe, dev, rank):
        self.args, self.eng, self.pipe, self.dev, self.rank = args, eng, pipe, dev, rank
        self.T = int(args.seconds * SR) // HOP
        self.B = args.clips
        self.codes = None

    def result_checksum(self) -> int:
        return int(self.codes.sum().item()) & 0xFFFFFFFF


class Recon(Workload):
    def __init__(self, *a):
        super().__init__(*a)
        from tests.golden.inputs import make_mel
        B, T = self.B, self.T
        base = make_mel(8, T, seed=100 + self.rank)
        reps = (B + 7) // 8
        self.mel_host = [base.roll(s, 0).repeat(reps, 1, 1)[:B].contiguous().pin_memory() for s in (0, 3)]
        self.mel_dev = [m.to(self.dev) for m in self.mel_host]
        self.codes_host = torch.empty(B, T, dtype=torch.int64).pin_memory()
        self.wav_host = torch.empty(B, T * HOP, dtype=torch.float32).pin_memory()

    def device_step(self, i):
        self.codes, self.wav = self.pipe.reconstruct_device(self.mel_dev[i & 1])

    def host_step(self, i):
        self.pipe.reconstruct(self.mel_host[i & 1], self.codes_host, self.wav_host)


class VqOnly(Workload):
    def __init__(self, *a):
        super().__init__(*a)
        from tests.golden.inputs import make_vq_rows
        N = self.B * self.T
        base = make_vq_rows(4096, seed=2 + self.rank).to(torch.bfloat16)
        reps = (N + 4095) // 4096
        self.x_host = [base.roll(s, 0).repeat(reps, 1)[:N].contiguous().pin_memory() for s in (0, 1531)]
        self.x_dev = [x.to(self.dev) for x in self.x_host]
        self.codes_host = torch.empty(N, dtype=torch.int64).pin_memory()

    def device_step(self, i):
        self.codes = self.eng.vq_search(self.x_dev[i & 1])

    def host_step(self, i):
        self.pipe.vq_search(self.x_host[i & 1], self.codes_host)


class Wav2Codes(Workload):
    def __init__(self, *a):
        super().__init__(*a)
        from tests.golden.inputs import make_wav
        n = int(self.args.seconds * SR)
        B = self.B
        base = make_wav(8, n, seed=3 + self.rank)
        reps = (B + 7) // 8
        self.wav_host = [base.roll(s, 0).repeat(reps, 1)[:B].contiguous().pin_memory() for s in (0, 3)]
        self.wav_dev = [w.to(self.dev) for w in self.wav_host]
        self.T = (n + 1 - 256) // 256 + 1
        self.codes_host = torch.empty(B, self.T, dtype=torch.int64).pin_memory()

    def device_step(self, i):
        self.codes = self.pipe.tokenize_wav_device(self.wav_dev[i & 1])

    def host_step(self, i):
        self.pipe.tokenize_wav(self.wav_host[i & 1], self.codes_host)


class Bulk(Workload):
    def __init__(self, *a):
        super().__init__(*a)
        from tests.golden.inputs import make_mel
        B, T = self.B, self.T
        base = make_mel(1, T, seed=200 + self.rank)
        self.mel_host = [base.roll(s, 2).repeat(B, 1, 1).contiguous().pin_memory() for s in (0, 977)]
        self.mel_dev = [m.to(self.dev) for m in self.mel_host]
        self.codes_host = torch.empty(B, T, dtype=torch.int64).pin_memory()
        self.wav_host = torch.empty(B, T * HOP, dtype=torch.float32).pin_memory()
        self.tile = self.args.tile

    def device_step(self, i):
        from distilcodec_nabeel_b200 import decode_long_device, tokenize_long_device
        self.codes = tokenize_long_device(self.pipe, self.mel_dev[i & 1], self.tile)
        self.wav = decode_long_device(self.pipe, self.codes, self.tile)

    def host_step(self, i):
        from distilcodec_nabeel_b200 import decode_long, tokenize_long
        tokenize_long(self.pipe, self.mel_host[i & 1], self.tile, self.codes_host)
        decode_long(self.pipe, self.codes_host, self.tile, self.wav_host)


WORKLOAD_CLASSES = {"recon": Recon, "vq_only": VqOnly, "wav2codes_30s": Wav2Codes, "bulk_10min": Bulk}


def kernel_tables(rows, K, peaks):
    """profile rows -> (per-class table sorted by time, per-layer table)."""
    step_ms = sum(r["ms"] for r in rows) or 1.0
    by_class = {}
    for r in rows:
        name = r["name"].split("[")[0]
        c = by_class.setdefault(name, {"name": name, "launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for f in ("launches", "ms", "flops", "bytes"):
            c[f] += r[f]
    layers = []
    for r in sorted(rows, key=lambda r: -r["ms"]):
        e = {"name": r["name"], "launches_per_step": r["launches"] / K, "ms_per_step": round(r["ms"] / K, 3)}
        if r["flops"] > 0 and r["ms"] > 0:
            e["tflops"] = round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 1)
        if r["bytes"] > 0 and r["ms"] > 0:
            e["gbs"] = round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1)
        layers.append(e)
    kernels = []
    for r in sorted(by_class.values(), key=lambda r: -r["ms"]):
        k = {"name": r["name"], "launches_per_step": r["launches"] / K, "ms_per_step": r["ms"] / K,
             "share": r["ms"] / step_ms}
        # both roofs from the algorithmic FLOPs / bytes the library books per launch; the binding one (larger fraction)
        # is reported as `bound` (a narrow decoder conv is an HBM kernel that happens to use the tensor core)
        tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["flops"] > 0 and r["ms"] > 0 else 0.0
        gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["bytes"] > 0 and r["ms"] > 0 else 0.0
        f_t, f_h = tf / peaks["bf16_tflops_sustained"], gb / peaks["hbm_gbs"]
        if tf > 0 or gb > 0:
            if f_t >= f_h:
                k.update(bound="tensor", achieved=tf, unit="TFLOP/s", peak=peaks["bf16_tflops_sustained"], frac=f_t)
            else:
                k.update(bound="hbm", achieved=gb, unit="GB/s", peak=peaks["hbm_gbs"], frac=f_h)
            k.update(frac_tensor=round(f_t, 4), frac_hbm=round(f_h, 4))
        kernels.append(k)
    return kernels, layers


def _short(k):
    """compact per-kernel summary for the JSON line"""
    if k is None:
        return None
    return {"kernel": k["name"], "bound": k.get("bound"), "achieved": round(k.get("achieved", 0.0), 1),
            "peak": k.get("peak"), "unit": k.get("unit"), "frac": round(k.get("frac", 0.0), 4),
            "ms_per_step": round(k["ms_per_step"], 3), "share_of_step": round(k["share"], 4)}


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="recon", choices=sorted(WORKLOADS))
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU per step (default: the BASELINE config's)")
    ap.add_argument("--seconds", type=float, default=0.0, help="clip length (default: the BASELINE config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine tunable key=value (dc_set_option), repeatable")
    ap.add_argument("--chunk", type=int, default=0, help="clips per device pass of the host-buffer leg; 0 = default")
    ap.add_argument("--tile", type=int, default=16384, help="frames per time tile (bulk_10min)")
    ap.add_argument("--cpu-clips", type=int, default=0, help="bounded CPU sample: clips per step (default 1; 2 in "
                                                             "the GPU arm's cpu_baseline)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--detail-out", default=None, help="side file for the per-kernel / per-layer tables")
    args = ap.parse_args()
    dc, ds = DEFAULT_SHAPE[args.workload]
    args.clips = args.clips or dc
    args.seconds = args.seconds or ds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world, out)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the "
                         "CPU baseline)")
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from distilcodec_nabeel_b200 import Engine, Pipeline, load_config, mel_buffers
    from distilcodec_nabeel_b200 import random_init as weights   # synthetic W0 weights (no oracle/ import in this arm)

    K, W = args.steps, max(args.warmup, 3)
    sd = weights.make_state_dict("W0")
    sd.update(mel_buffers(load_config()))
    eng = Engine(sd, local, args.mode, workspace_limit_bytes=64 << 30)
    del sd
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    pipe = Pipeline(eng, chunk=args.chunk or None)
    wl = WORKLOAD_CLASSES[args.workload](args, eng, pipe, dev, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather_ranks(x: float):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]

    # ---------------------------------------------------------------- device-resident: `value`
    for i in range(W):
        wl.device_step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.profile(True)
    n0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        wl.device_step(i)
    ev1.record()
    barrier()
    per_rank_total = gather_ranks(ev0.elapsed_time(ev1))
    ms_total = max(per_rank_total)
    launches = eng.launch_count() - n0
    rows = eng.profile_rows()
    eng.profile(False)
    clocks = sampler.stop() if sampler else {}
    audio_s_per_step = world * wl.B * wl.T * HOP / SR
    value = audio_s_per_step * K / (ms_total / 1e3)
    checksum = wl.result_checksum()

    # ---------------------------------------------------------------- host buffers: `e2e`
    for i in range(2):
        wl.host_step(i)
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        wl.host_step(i)
    torch.cuda.synchronize(dev)
    e2e_s = max(gather_ranks(time.perf_counter() - t0))
    barrier()
    e2e_value = audio_s_per_step * K / e2e_s
    h2d, d2h = pipe.h2d_bytes // K, pipe.d2h_bytes // K
    total_launches = int(sum(gather_ranks(float(launches))))

    # ---------------------------------------------------------------- per-kernel roofline (rank 0's records)
    peaks = load_peaks()
    kernels, layers = kernel_tables(rows, K, peaks)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.isfile(tp):
        traffic = json.load(open(tp))
    roofline = None
    if kernels:
        roofline = _short(kernels[0])
        roofline["traffic"] = (traffic or {}).get(kernels[0]["name"])
        roofline["launches_per_step"] = kernels[0]["launches_per_step"]
        roofline["peak_source"] = f"{peaks['source']}, sustained bf16 figure (kernel timed inside a long step)"
    by_name = {k["name"]: k for k in kernels}
    vq = _short(by_name.get("vq_score"))
    dw = _short(by_name.get("dwconv_ln"))
    per_rank_ms = [round(t / K, 2) for t in per_rank_total]

    frames_per_s = audio_s_per_step * FPS * K / (ms_total / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.mode, "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": total_launches, "clocks": clocks, "roofline": roofline, "roofline_vq_score": vq,
            "roofline_dwconv_ln": dw, "per_rank_ms": per_rank_ms,
            "slow_rank": int(max(range(world), key=lambda r: per_rank_total[r])),
            "model_tflops": round(frames_per_s * MFLOP_PER_FRAME[args.workload] * 1e6 / 1e12, 1),
            "codes_checksum": checksum}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only, bounded sample)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cclips = args.cpu_clips or 2
        csecs = min(args.seconds, args.cpu_seconds)
        rate, ms, cores, Tc, kind = cpu_reference_rate(args.workload, csecs, cclips, 1, 1)
        what = "the reference's own PyTorch modules" if kind == "reference" else "oracle port (torch restatement)"
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{cclips} clips x {csecs:g} s (T={Tc}) of the same workload, one timed pass "
                                          f"after one warm-up, fp32, {what}, {ms:.0f} ms"}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        detail = args.detail_out or os.path.join(ROOT, "profiles", f"bench_detail_{args.workload}_n{world}.json")
        try:
            os.makedirs(os.path.dirname(os.path.abspath(detail)), exist_ok=True)
            with open(detail, "w") as f:
                json.dump({"line": line, "kernels": kernels, "layers": layers}, f, indent=1)
            line["detail"] = os.path.relpath(detail, ROOT)
        except OSError as e:
            line["detail"] = f"not written: {e}"
        print(json.dumps(line), file=out, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
