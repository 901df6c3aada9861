#!/usr/bin/env python
"""bench.py — channel-samples/s of the effect-node path on N B200s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port)

A "step" is one dspb_process call over one batch of synthetic input: C channels x n samples per
channel (n = blocks_per_step device blocks), state carried across steps.  `value` is measured with
inputs and outputs resident in HBM; `e2e` goes through the host-buffer C-ABI call (pinned host
memory, H2D and D2H inside the timed region).  Channels shard across ranks with no data-path
collective ("weak" scaling: every GPU runs `--channels` channels).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "channel_samples_per_sec"
UNIT = "channel-samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DSPB_BENCH_WORKLOAD", "target"))
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU (default: the workload's)")
    ap.add_argument("--block", type=int, default=1024, help="device block (samples)")
    ap.add_argument("--blocks-per-step", type=int, default=16)
    ap.add_argument("--fir-mode", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", action="store_true", help="also time an NCCL all_gather of the outputs")
    return ap.parse_args()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` captures (profiles/): keyed by (workload, kernel kind); None = not captured yet.
# (workload, kind, channels, samples per step) -> bytes
TRAFFIC = {
    ("target", "fir", 4096, 16384): 335779840 + 229628672,     # profiles/r01s3_target_fir_fft_kernel.txt
    ("target", "fused", 4096, 16384): 539587328 + 487191552,   # profiles/r01s4_target_fused_chain_kernel.txt
}

DEFAULT_CHANNELS = {"config1": 2, "config2": 256, "config2_one_pole": 256, "config3": 1024, "config4": 4096, "config5": 1024, "target": 4096}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    """Dense bf16 TFLOP/s: the sustained figure (a kernel timed inside a long step), else the recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_oracle(spec, channels, n, threads, steps=1, warmup=0):
    """Times the CPU oracle (the reference's arithmetic and 128-sample per-node block structure)."""
    from dsp_stuff_b200 import signals as S
    from oracle import oracle

    o = oracle.Oracle(channels, threads=threads)
    spec.apply(o)
    x = S.noise(channels, n)
    for _ in range(warmup):
        o.process(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.process(x)
    dt = time.perf_counter() - t0
    return channels * n * steps / dt, dt


def cpu_baseline(spec, budget_s):
    cores = len(os.sched_getaffinity(0))
    # calibrate on a small sample, then size the real one to the budget
    n0 = 1024
    rate, dt = time_oracle(spec, cores, n0, cores)
    n = int(max(1024, min(48000 * 120, rate * budget_s / cores)) // 128 * 128)
    rate, dt = time_oracle(spec, cores, n, cores)
    # SURVEY §8d also asks for the single-thread figure: one channel, ~2 s
    r1, _ = time_oracle(spec, 1, n0, 1)
    n1 = int(max(1024, min(48000 * 60, r1 * 2.0)) // 128 * 128)
    r1, dt1 = time_oracle(spec, 1, n1, 1)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} channels x {n} samples of the same graph and noise input, {cores} threads (one channel each), {dt:.1f} s",
            "single_thread": {"value": r1, "unit": UNIT, "cores": 1, "sample": f"1 channel x {n1} samples, {dt1:.1f} s"}}


def run_reference(args, spec, alg_bytes, C, n):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    rate0, _ = time_oracle(spec, cores, 1024, cores)
    # each step: a bounded sample (cores channels), sized so the whole run stays within ~2 minutes
    per_step = max(1.0, min(8.0, 100.0 / max(1, args.steps + args.warmup)))
    ns = int(max(1024, min(n, rate0 * per_step / cores)) // 128 * 128)
    rate, dt = time_oracle(spec, cores, ns, cores, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "channels_per_gpu": C, "samples_per_step": n, "block": args.block},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = {cores} channels x {ns} samples of the workload graph (oracle/dsp_oracle.cpp, the "
                                   f"reference's arithmetic; the Rust reference cannot be built here), {cores} threads"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The JSON line is the ONLY thing on stdout: libraries (NCCL prints its version banner there) were
    redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    from dsp_stuff_b200 import signals as S

    if args.workload not in S.WORKLOADS:
        raise SystemExit(f"unknown workload {args.workload}; have {sorted(S.WORKLOADS)}")
    factory, alg_bytes = S.WORKLOADS[args.workload]
    spec = factory()
    C = args.channels or DEFAULT_CHANNELS[args.workload]
    n = args.block * args.blocks_per_step

    if args.impl == "reference":
        run_reference(args, spec, alg_bytes, C, n)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from dsp_stuff_b200.engine import Engine

    eng = Engine(C, block=args.block, max_samples=n, device=local, fir_mode=args.fir_mode)
    spec.apply(eng)
    n_in, n_out = eng._n_in, eng._n_out

    # synthetic input, identical on host and device; this rank's channels are [rank*C, (rank+1)*C)
    x_host = [torch.from_numpy(S.noise(C, n, seed=42 + i, channel_offset=rank * C)).pin_memory() for i in range(n_in)]
    y_host = [torch.empty((C, n), dtype=torch.float32).pin_memory() for _ in range(n_out)]
    x_dev = [t.cuda(non_blocking=True) for t in x_host]
    y_dev = [torch.empty((C, n), dtype=torch.float32, device="cuda") for _ in range(n_out)]
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        eng.process_device(x_dev, y_dev, n)
    launches_per_step = eng.kernel_launches
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile(True)  # CUDA events around every kernel step, on the launching stream, inside the timed region
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.process_device(x_dev, y_dev, n)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    step_times = eng.profile_read()
    eng.profile(False)
    plan = eng.plan_steps()

    # end to end through the host-buffer C-ABI call (pinned memory; H2D + kernels + D2H per step)
    e2e_ms = 0.0
    if not args.no_e2e:
        for _ in range(2):
            eng.process_host(x_host, y_host, n)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.process_host(x_host, y_host, n)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3

    gather_ms = None
    if args.gather and world > 1:
        outs = [torch.empty_like(y_dev[0]) for _ in range(world)] if rank == 0 else None
        dist.gather(y_dev[0], outs, dst=0)  # untimed: the first call sets up the NVLink connections
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        dist.gather(y_dev[0], outs, dst=0)
        g1.record(stream)
        barrier()
        gather_ms = g0.elapsed_time(g1)

    t = torch.tensor([ms, e2e_ms, gather_ms or 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, gather_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        total = float(world) * C * n * args.steps
        value = total / (ms * 1e-3)
        peak, peak_src = load_peaks()
        # dominant kernel = the schedule step with the largest summed device time; its achieved bandwidth is
        # its own ALGORITHMIC bytes per launch (per channel-sample figure x C x n) / its average launch time
        kernels = []
        for i, (tot_ms, rounds) in enumerate(step_times):
            if rounds:
                avg = tot_ms / rounds
                kernels.append({"step": i, "kind": plan[i]["kind"], "alg_bytes_per_channel_sample": plan[i]["alg_bytes"],
                                "avg_ms": avg, "share_of_step": tot_ms / ms,
                                "achieved_gbs": plan[i]["alg_bytes"] * C * n / (avg * 1e-3) / 1e9})
        dom = max(kernels, key=lambda k: k["avg_ms"])
        per_gpu_rate = C * n * args.steps / (ms * 1e-3)
        achieved = dom["achieved_gbs"]
        step_achieved = alg_bytes * per_gpu_rate / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "channels_per_gpu": C, "samples_per_step": n, "block": args.block,
                       "blocks_per_step": args.blocks_per_step, "fir_mode": {0: "fft", 1: "direct_f64", 2: "toeplitz_tcgen05_bf16x2", 3: "fft_packed_f32x2"}[args.fir_mode],
                       "l2_policy": f"inputs+outputs {2 * C * n * 4 / 2**20:.0f} MiB per step exceed the 126 MB L2",
                       "x_realtime_per_channel": value / world / C / 48000.0},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC.get((args.workload, dom["kind"], C, n)), "peak_source": peak_src,
                         "kernel": f"step {dom['step']} ({dom['kind']})", "kernel_avg_ms": dom["avg_ms"],
                         "kernel_alg_bytes_per_channel_sample": dom["alg_bytes_per_channel_sample"],
                         "kernel_share_of_step": dom["share_of_step"], "kernels": kernels,
                         "whole_step": {"alg_bytes_per_channel_sample": alg_bytes, "achieved": step_achieved,
                                        "frac": step_achieved / peak, "frac_of_8000_nominal": step_achieved / 8000.0}},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
        }
        if args.fir_mode == 2 and dom["kind"] == "fir":
            # Toeplitz tensor-core FIR: the bounding roofline is the tensor pipe.  Algorithmic flops of the kernel =
            # 3 split-bf16 products x 2 N flop per channel-sample (N taps), all launches of the step (split pre-pass included)
            n_taps = max([len(nd.taps) for nd in spec.nodes if nd.typename == "fir" and nd.taps is not None] or [1])
            tf = 3 * 2 * n_taps * C * n / (dom["avg_ms"] * 1e-3) / 1e12
            tpeak, tsrc = load_tensor_peak()
            line["roofline"].update({"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                                     "peak_source": tsrc, "traffic": None,
                                     "flops_per_channel_sample": 3 * 2 * n_taps,
                                     "hbm_view": {"achieved_gbs": achieved, "peak_gbs": peak}})
        if not args.no_e2e:
            line["e2e"] = {"value": float(world) * C * n * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": n_in * C * n * 4, "d2h_bytes_per_step": n_out * C * n * 4}
        if gather_ms:
            line["gather_to_rank0_ms"] = gather_ms
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(spec, args.cpu_seconds)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
