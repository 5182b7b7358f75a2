#!/usr/bin/env python
"""bench.py — throughput of the DistilCodec hot path on B200 (BASELINE.json metric: audio-seconds processed per
second, encode+decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips C] [--seconds S] [--impl reference]

A step = one pass of the hot path (mel -> ConvNeXt encoder -> 32768x3584 VQ -> HiFiGAN decoder -> wav) over one
batch of synthetic clips: BASELINE configs[3], "full encode->decode reconstruction, 256 synthetic 10 s clips", per
GPU (weak scaling: clips shard across ranks, no data-path collective).  Weights are random-init of the architecture
in configs/model_config.json (distilcodec_nabeel_b200/random_init.py, W0 — the HF checkpoint is not available offline).

Prints ONE JSON line (rank 0):
  value        whole-job audio-s/s with the mel batch already resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-buffer call (Pipeline.reconstruct: pinned host mel in, host codes+wav
               out, copies inside the timed region)
  roofline     the dominant kernel class by device time (event pairs recorded inside the library around every
               launch of the timed steps): algorithmic FLOPs / summed launch time vs MEASURED_PEAKS.json
  kernels      the same for every kernel class (share of the step, achieved TFLOP/s or GB/s)
  cpu_baseline the oracle (CPU restatement of the reference, torch fp32, all host threads) on a bounded sample
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
MFLOP_PER_FRAME = 2045.07          # encoder 154.25 + quantizer 52.46 + distance 234.88 + generator 1603.49


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


def cpu_reference_rate(seconds_of_audio: float, clips: int, steps: int, warmup: int):
    """The reference's CPU path (oracle port: torch fp32 restatement of encoder / VQ / generator, pinned to the
    reference's outputs by tests/) on `clips` x `seconds_of_audio`, all host threads.  -> (audio-s/s, ms/step, cores)"""
    from oracle import restatement as R
    from oracle import weights
    from tests.golden.inputs import make_mel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = weights.make_state_dict("W0")
    T = int(seconds_of_audio * SR) // HOP
    mel = make_mel(clips, T, seed=17)
    with torch.no_grad():
        for _ in range(warmup):
            R.codec_forward(sd, mel)
        t0 = time.perf_counter()
        for _ in range(steps):
            R.codec_forward(sd, mel)
        dt = time.perf_counter() - t0
    audio_s = clips * T * HOP / SR * steps
    return audio_s / dt, dt / steps * 1e3, cores, T


def run_reference_arm(args, rank: int, world: int, out):
    if rank != 0:
        return
    clips, secs = 1, 10.0
    rate, ms, cores, T = cpu_reference_rate(secs, clips, args.steps, min(args.warmup, 1))
    sample = f"{clips} clip x {secs:.0f} s (T={T} frames) per step, full encode->quantize->decode, fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


def workload_config(args, world: int) -> dict:
    T = int(args.seconds * SR) // HOP
    return {"workload": f"BASELINE configs[3]: full encode->decode reconstruction, {args.clips} synthetic "
                        f"{args.seconds:g} s clips per GPU (log-mel (B,128,{T}) -> codes + 24 kHz waveform)",
            "clips_per_gpu": args.clips, "frames_per_clip": T, "weights": "random-init W0 of configs/model_config.json",
            "parallelism": f"clip-sharded x{world}, no collective on the data path",
            "l2": "per-step working set (tens of GB of activations, 705 MB codebook) >> 126 MB L2; two alternating "
                  "input batches"}


def _claim_stdout():
    """Route everything that libraries print on fd 1 (e.g. NCCL's version banner) to stderr and return a file object
    on the real stdout, so that the JSON line is the only thing printed there."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU per step (BASELINE configs[3]: 256)")
    ap.add_argument("--seconds", type=float, default=10.0, help="clip length")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine tunable key=value (dc_set_option), repeatable")
    ap.add_argument("--chunk", type=int, default=0, help="clips per device pass of the host-buffer (e2e) leg; 0 = Pipeline default")
    ap.add_argument("--cpu-clips", type=int, default=2, help="bounded CPU-baseline sample: clips of --cpu-seconds")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()

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

    from distilcodec_nabeel_b200 import Engine, Pipeline
    from distilcodec_nabeel_b200 import random_init as weights   # synthetic W0 weights (no oracle/ import in this arm)
    from tests.golden.inputs import make_mel

    K, W = args.steps, max(args.warmup, 3)
    T = int(args.seconds * SR) // HOP
    B = args.clips
    sd = weights.make_state_dict("W0")
    eng = Engine(sd, local, args.mode, workspace_limit_bytes=64 << 30)
    del sd
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    pipe = Pipeline(eng, chunk=args.chunk or None)
    # two alternating synthetic batches (bit-identical on every machine: numpy Philox), kept on host (pinned) and in HBM
    base = make_mel(8, T, seed=100 + rank)
    reps = (B + 7) // 8
    mel_host = [base.roll(s, 0).repeat(reps, 1, 1)[:B].contiguous().pin_memory() for s in (0, 3)]
    mel_dev = [m.to(dev) for m in mel_host]
    codes_host = torch.empty(B, T, dtype=torch.int64).pin_memory()
    wav_host = torch.empty(B, T * HOP, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------------------------------------------------------- device-resident: `value`
    for i in range(W):
        pipe.reconstruct_device(mel_dev[i & 1])
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
        codes, wav = pipe.reconstruct_device(mel_dev[i & 1])
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count() - n0
    rows = eng.profile_rows()
    eng.profile(False)
    clocks = sampler.stop() if sampler else {}
    audio_s_per_step = world * B * T * HOP / SR
    value = audio_s_per_step * K / (ms_total / 1e3)
    checksum = int(codes.sum().item()) & 0xFFFFFFFF

    # ---------------------------------------------------------------- host buffers: `e2e`
    for i in range(2):
        pipe.reconstruct(mel_host[i & 1], codes_host, wav_host)
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        pipe.reconstruct(mel_host[i & 1], codes_host, wav_host)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = audio_s_per_step * K / e2e_s
    h2d, d2h = pipe.h2d_bytes // K, pipe.d2h_bytes // K
    total_launches = int(sum_over_ranks(float(launches)))

    # ---------------------------------------------------------------- per-kernel roofline (rank 0's records)
    peaks = load_peaks()
    step_ms = sum(r["ms"] for r in rows) or 1.0
    layer_rows = rows
    by_class = {}
    for r in rows:
        c = by_class.setdefault(r["name"].split("[")[0], {"name": r["name"].split("[")[0], "launches": 0, "ms": 0.0,
                                                          "flops": 0.0, "bytes": 0.0})
        for f in ("launches", "ms", "flops", "bytes"):
            c[f] += r[f]
    rows = list(by_class.values())
    layers = []
    for r in sorted(layer_rows, key=lambda r: -r["ms"])[:120]:
        e = {"name": r["name"], "launches_per_step": r["launches"] / K, "ms_per_step": round(r["ms"] / K, 3)}
        if r["flops"] > 0 and r["ms"] > 0:
            e["tflops"] = round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 1)
        if r["bytes"] > 0 and r["ms"] > 0:
            e["gbs"] = round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1)
        layers.append(e)
    kernels = []
    for r in sorted(rows, key=lambda r: -r["ms"]):
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
    roofline = None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.isfile(tp):
        traffic = json.load(open(tp))
    if kernels:
        d = kernels[0]
        roofline = {"kernel": d["name"], "bound": d.get("bound"), "achieved": d.get("achieved"), "peak": d.get("peak"),
                    "unit": d.get("unit"), "frac": d.get("frac"),
                    "traffic": (traffic or {}).get(d["name"]),
                    "peak_source": f"{peaks['source']} (sustained bf16 figure: kernel timed inside a long step)",
                    "launches_per_step": d["launches_per_step"], "share_of_step": d["share"]}
    vq = next((k for k in kernels if k["name"] == "vq_score"), None)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.mode, "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": total_launches, "clocks": clocks, "roofline": roofline, "roofline_vq_score": vq,
            "kernels": kernels, "layers": layers,
            "model_tflops": audio_s_per_step * (SR / HOP) * MFLOP_PER_FRAME * 1e6 * K / (ms_total / 1e3) / 1e12,
            "codes_checksum": checksum}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only, bounded sample)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, ms, cores, Tc = cpu_reference_rate(args.cpu_seconds, args.cpu_clips, 1, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_clips} clips x {args.cpu_seconds:g} s (T={Tc}) of the same workload, "
                                          f"one timed pass after one warm-up, oracle (torch fp32 restatement), {ms:.0f} ms"}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), file=out, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
