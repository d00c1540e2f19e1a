#!/usr/bin/env python
"""bench.py -- Picard iterations/sec on BASELINE.json's headline configuration.

Workload (config.workload): c3 = N=128 components, T=1e7 samples, f64, Picard-O extended tanh, mixed Laplace /
uniform sources mixed by a seed-42 N(0,1) matrix, explicit seed-43 orthogonal w_init.  With
--gpus N the T samples are sharded over N ranks (strong scaling, one NCCL allreduce of the packed N x N
moments per pass).

A "step" is one outer Picard iteration (core.rs:211-391): the stored-Y gradient pass + the LOSS passes of the
line search over the sample matrix + the N x N work between them.
  value  : iterations/sec of the core loop with the preprocessed data resident in HBM (picard_core_run),
           device time from CUDA events recorded on the library's stream, max over ranks.
  e2e    : iterations/sec through the reference-facing call Picard.fit_with_config on HOST buffers (pinned):
           H2D of X, centering, whitening, the whole fit to convergence, D2H of the sources -- all timed.
  roofline: the pass kernel with the largest share of the step (LOSS pass of a line-search try: 2 N^2 T_local f64
           flops per launch; stored-Y gradient pass: 2 N^2 T_local, 4 N^2 T_local with the Hessian moments) against
           the FP64 tensor (DMMA) peak measured by a microbenchmark at the start of THIS run (picard_fp64_peak_probe;
           MEASURED_PEAKS.json has no FP64 entry); launch durations from CUDA events inside the library around each
           pass launch.  For the INT8 tensor-core engines the `engine` sub-object gives the integer-op view.
  parity : run beside the measurement -- the INT8 engines against the FP64 kernels on the full device-resident data
           (LOSS moments + stored-Y gradient at one W), and a 1e5-sample slice against the CPU oracle.
  cpu_baseline: the CPU oracle (a restatement of the reference's cost structure, oracle/picard_oracle.cpp)
           timed on this box's host cores on a bounded sample (N=128, T=T_cpu) and scaled linearly in T.

`--impl reference` times the CPU oracle as the reference arm (the Rust reference cannot be built in this
image: no cargo/rustc): a step is ONE outer iteration of the oracle's core loop on T = 1e6 samples
(1/10 of the workload), iterations/s scaled linearly in T.
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "picard_iters_per_sec"
UNIT = "iterations/s"
WORKLOADS = {
    # name: (N, T, ortho, extended, density kind, alpha, source kind)
    "c3": dict(n=128, t=10_000_000, ortho=True, extended=True, kind=0, alpha=1.0, n_laplace=64,
               desc="N=128,T=1e7 f64 Picard-O extended tanh, mixed Laplace/uniform sources"),
    "c2": dict(n=64, t=1_000_000, ortho=True, extended=True, kind=0, alpha=1.0, n_laplace=32,
               desc="N=64,T=1e6 f64 Picard-O extended tanh, mixed Laplace/uniform sources"),
    "tiny": dict(n=16, t=200_000, ortho=True, extended=True, kind=0, alpha=1.0, n_laplace=8, desc="N=16,T=2e5 (debug)"),
    # the other BASELINE configs: parity-test cases, runnable here for the record (profiles/), not the headline bench line
    "c1": dict(n=3, t=10_000, ortho=False, extended=False, kind=0, alpha=1.0, n_laplace=3,
               desc="N=3,T=1e4 f64 Picard tanh non-ortho, Laplace sources (reference bench case)"),
    "c4": dict(n=256, t=50_000_000, ortho=False, extended=False, kind=1, alpha=0.1, n_laplace=256, per_gpu_t=6_250_000,
               desc="N=256,T=5e7 f64 Picard non-ortho exp(alpha=0.1), Laplace sources, 8xB200 sharded "
                    "(with fewer GPUs the per-GPU shard of 6.25e6 samples is kept and T shrinks)"),
    "c5": dict(n=32, t=1_000_000, ortho=True, extended=True, kind=0, alpha=1.0, n_laplace=16, jade_it=50,
               desc="N=32,T=1e6 f64 jade_it=50 warm start then Picard-O extended (JADE runs inside the e2e fit)"),
}
FP64_PEAK_FALLBACK_TFLOPS = 37.19  # round-1 microbenchmark (profiles/microbench/fp64_pipes_r01.jsonl); used only if the in-run probe fails
# dram__bytes_read.sum + dram__bytes_write.sum per launch at c3 on one GPU, from the committed ncu --set full captures (None = not
# captured for the current kernel).  Scaled to the rank's share of the samples; reported with its source, never silently.
NCU_TRAFFIC_C3 = {"loss": (7.721367e9 + 10.192575e9, "profiles/summary_r02y.txt (ncu --set full of loss_i8_kernel inside this bench command, one GPU, T = 1e7)"),
                  "grady": (10.327482e9 + 0.091629e9, "profiles/summary_r02z.txt (ncu --set full of grad_i8_kernel inside this bench command, one GPU, T = 1e7)")}


def _hbm_peak():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the SURVEY.md §8(d) figure."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            v = float(json.load(fh)["hbm_gbs"])
        if v > 0:
            return v, "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, ValueError, KeyError, TypeError):
        pass
    return 6462.4, "SURVEY.md 8(d): 6 462 GB/s measured on this pool (MEASURED_PEAKS.json not found)"


HBM_PEAK_GBS = _hbm_peak()


class ClockSampler(threading.Thread):
    """Samples SM clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line).  NVML in-process
    (nvidia_ml_py) when available: spawning nvidia-smi every 100 ms was measured to stall the solver's kernel launches
    (driver locks) by up to 1 ms per line-search try; falls back to nvidia-smi at a 1 s period."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.source = "nvml"
        self._halt = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
            self.source = "nvidia-smi"

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if i < len(ids) and ids[i].isdigit():
                return int(ids[i])
        return i

    def _sample_nvml(self):
        nv = self._nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for nm, v in zip(names, out[2:]):
            if v.strip().lower() == "active":
                self.reasons.add(nm)

    def run(self):
        while not self._halt.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.05 if self._h is not None else 1.0)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": self.source}


def host_inputs(n):
    import _data
    import numpy as np
    a = np.random.default_rng(42).standard_normal((n, n))
    w0 = _data.orthogonal(n, 43)
    return a, w0


def config_dict(args, wl, world, t_total, t_local):
    """The `config` object: identical in the b200 arm and the reference arm of the same launch."""
    return {"workload": wl["desc"], "name": args.workload, "n": wl["n"], "t_total": int(t_total), "t_per_gpu": int(t_local),
            "parallelism": f"sample-sharded x{world}",
            "l2": "inputs (8*N*T_local bytes per pass) far exceed the 126 MB L2",
            "timing": "b200 arm: CUDA events on the library stream around picard_core_run, max over ranks (wall clock cross-check in "
                      "wall_ms_per_step); reference arm: wall clock of the CPU core loop on the bounded sample named in "
                      "cpu_baseline.sample, iterations/s scaled linearly in T to t_total, rank 0 only"}


def shard_sizes(wl, world):
    from picard_ica_b200.dist import shard_range
    t_total = wl["t"]
    if "per_gpu_t" in wl and world * wl["per_gpu_t"] < t_total:
        t_total = world * wl["per_gpu_t"]  # c4 is specified for 8 GPUs: keep its per-GPU shard when fewer are available
    return t_total, shard_range(t_total, 0, world)


def cpu_oracle_rate(wl, t_cpu, warm_iters, timed_iters, threads=None):
    """Iterations/sec of the CPU oracle core loop (oracle/picard_oracle.cpp, the reference's cost structure on the same BLAS) on a
    bounded sample: warm_iters untimed outer iterations, then timed_iters timed ones (both from W = I on the same whitened data);
    scaled linearly in T to the workload."""
    import numpy as np
    import _data
    from oracle import oracle as orc
    if threads:
        orc.set_threads(threads)
    n = wl["n"]
    kind = "mixed" if 0 < wl["n_laplace"] < n else ("laplace" if wl["n_laplace"] else "uniform")
    xw = _data.whitened(n, t_cpu, seed=42, kind=kind)
    kw = dict(covariance=np.eye(n) if wl["extended"] else None, want_y=False)
    if warm_iters > 0:
        orc.core_run(xw, wl["kind"], wl["alpha"], wl["ortho"], wl["extended"], max_iter=warm_iters, **kw)
    t0 = time.perf_counter()
    r = orc.core_run(xw, wl["kind"], wl["alpha"], wl["ortho"], wl["extended"], max_iter=timed_iters, **kw)
    wall = time.perf_counter() - t0
    secs = r.seconds if r.seconds > 0 else wall
    it_s_sample = r.n_iterations / secs
    scaled = it_s_sample * (t_cpu / wl["t"])
    return dict(value=scaled, iters=r.n_iterations, seconds=secs, t_cpu=t_cpu, it_s_sample=it_s_sample,
                cores=orc.get_threads(), loss_evals=r.loss_evals)


def host_description():
    try:
        import psutil
        ram = psutil.virtual_memory().total / 2 ** 30
    except Exception:
        ram = float("nan")
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"cpu": model, "logical_cores": os.cpu_count(), "ram_gib": round(ram, 1)}


def run_reference(args, wl, rank, world, out):
    """The reference arm: the reference's own CPU path (here: the C++ port of it, see the module docstring) on the host cores.
    A step is ONE outer iteration of the core loop on T = --ref-t samples (default 1e6 = T/10)."""
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    t_total, (b0, e0) = shard_sizes(wl, world)
    res = cpu_oracle_rate(wl, args.ref_t, args.warmup, args.steps, ncores)
    value = res["value"]
    sample = (f"oracle core loop (C++ port of core.rs on OpenBLAS, {res['cores']} threads), N={wl['n']}, T={args.ref_t} "
              f"(1/{wl['t'] / args.ref_t:g} of the workload): {args.warmup} warm-up + {res['iters']} timed outer iterations in "
              f"{res['seconds']:.1f} s ({res['it_s_sample']:.4f} it/s on the sample), iterations/s scaled linearly in T to T={wl['t']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(res["iters"], 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, wl, world, t_total, e0 - b0),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample, "host": host_description()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


def main():
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-t", type=int, default=1_000_000, help="samples of the cpu_baseline sample of the b200 arm (T/10)")
    ap.add_argument("--cpu-iters", type=int, default=2, help="timed outer iterations of the cpu_baseline sample (after one warm-up iteration)")
    ap.add_argument("--ref-t", type=int, default=1_000_000, help="samples of the reference arm's bounded sample (T/10, SURVEY.md §8d)")
    ap.add_argument("--flags", type=int, default=0, help="PICARD_FLAG_* bits (ablation)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world, out)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import picard_ica_b200 as P
    from picard_ica_b200 import _ffi
    from picard_ica_b200.dist import Communicator, shard_range
    import ctypes as C

    if not torch.cuda.is_available() or _ffi.lib().picard_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libpicard_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # BENCH_AFFINITY=1: one slice of the host cores per rank.  Off by default: at 8 GPUs it made no difference within the noise
    # (397.9 against 392.1 it/s) and one of three runs lost 24 ms in a single pass, the mark of a preempted spinning thread on a
    # crowded slice (profiles/README.md, call w3).
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity") and os.environ.get("BENCH_AFFINITY"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if per >= 1:
                mine = cores[local_rank * per:(local_rank + 1) * per]
                os.sched_setaffinity(0, mine)
                affinity = mine
        except OSError:
            affinity = None
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = Communicator.from_torch_distributed(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _ffi.lib()
    n, t_total = wl["n"], wl["t"]
    if "per_gpu_t" in wl and world * wl["per_gpu_t"] < t_total:
        t_total = world * wl["per_gpu_t"]  # c4 is specified for 8 GPUs: keep its per-GPU shard when fewer are available
    t0, t1 = shard_range(t_total, rank, world)
    t_local = t1 - t0
    ld = (t_local + 15) // 16 * 16
    a_mix, w0 = host_inputs(n)
    dp = _ffi.dp

    def hp(a):
        return a.ctypes.data_as(dp)

    # ---- synthetic data on the device: S (counter-based, identical for any sharding), X = A S
    s_dev = torch.empty((n, ld), dtype=torch.float64, device=dev)
    x_dev = torch.empty((n, ld), dtype=torch.float64, device=dev)
    st = lib.picard_synth_sources(C.c_void_p(s_dev.data_ptr()), C.c_int64(n), C.c_int64(t_local), C.c_int64(ld), C.c_int64(t0),
                                  C.c_int64(wl["n_laplace"]), C.c_uint64(42), C.c_int32(local_rank), None)
    assert st == 0
    st = lib.picard_apply_device(hp(a_mix), None, C.c_int64(n), C.c_int64(n), C.c_void_p(s_dev.data_ptr()), C.c_int64(ld),
                                 C.c_void_p(x_dev.data_ptr()), C.c_int64(ld), C.c_int64(t_local), C.c_int32(local_rank), None)
    assert st == 0
    del s_dev
    # ---- preprocessing for the core loop: x1 = w_init K (x - mean)  (solver.rs:77-140), on the device
    mean = np.zeros(n); k = np.zeros((n, n))
    err = C.create_string_buffer(1024)
    st = lib.picard_center_whiten_device(C.c_void_p(x_dev.data_ptr()), C.c_int64(n), C.c_int64(t_local), C.c_int64(ld), C.c_int64(n),
                                         C.c_int32(1), comm.handle if comm else None, C.c_int32(local_rank), hp(mean), hp(k), err,
                                         C.c_size_t(1024))
    assert st == 0, err.value
    a_tot = np.ascontiguousarray(w0 @ k)
    x1_dev = torch.empty((n, ld), dtype=torch.float64, device=dev)
    st = lib.picard_apply_device(hp(a_tot), hp(mean), C.c_int64(n), C.c_int64(n), C.c_void_p(x_dev.data_ptr()), C.c_int64(ld),
                                 C.c_void_p(x1_dev.data_ptr()), C.c_int64(ld), C.c_int64(t_local), C.c_int32(local_rank), None)
    assert st == 0
    torch.cuda.synchronize()

    cfg = P.PicardConfig(density=P.Tanh(wl["alpha"]) if wl["kind"] == 0 else (P.Exp(wl["alpha"]) if wl["kind"] == 1 else P.Cube()),
                         ortho=wl["ortho"], extended=wl["extended"], comm=comm, device=local_rank, flags=args.flags)
    core = P.CoreLoop(x1_dev[:, :t_local], cfg, covariance_identity=wl["extended"])

    def run_iters(k_iters):
        """Exactly k_iters outer iterations of the real trajectory (restarting from W = I if it converges first)."""
        done = 0
        restarts = 0
        while done < k_iters:
            d, conv = core.run(k_iters - done)
            done += d
            if done < k_iters and (conv or d == 0):
                core.reset(); restarts += 1
                if restarts > k_iters:
                    raise RuntimeError("core loop makes no progress")
        return restarts

    run_iters(args.warmup)
    # FP64 tensor (DMMA) peak of THIS device at its current clocks: the roofline denominator (~60 ms microbenchmark, warm GPU)
    peak = C.c_double(0.0)
    if lib.picard_fp64_peak_probe(C.c_int32(local_rank), C.c_double(60.0), C.byref(peak)) != 0 or not (peak.value > 1.0):
        peak_tflops, peak_source = FP64_PEAK_FALLBACK_TFLOPS, "round-1 DMMA microbenchmark (profiles/microbench/fp64_pipes_r01.jsonl): the in-run probe failed"
    else:
        peak_tflops = peak.value
        peak_source = ("FP64 DMMA m8n8k4 microbenchmark run by this process right before the timed region (picard_fp64_peak_probe, "
                       "148 SMs x 8 warps x 8 chains); MEASURED_PEAKS.json has no FP64 entry")
    s0 = core.stats()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    w_t0 = time.perf_counter()
    restarts = run_iters(args.steps)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w_t0)
    clocks = sampler.stop()
    s1 = core.stats()
    dev_ms = s1["core_ms"] - s0["core_ms"]  # CUDA events on the library's stream, recorded around every run() call
    tm = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(tm[0]), float(tm[1])
    d = {k_: s1[k_] - s0[k_] for k_ in s1}
    value = args.steps / (dev_ms / 1e3)

    # ---- roofline of the dominant kernel: the pass kind with the largest share of the timed region
    n2t = float(n) * n * t_local
    cand = [("fused", 4.0 * n2t, d["fused_passes"], d["pass_ms_fused"]), ("grad", 4.0 * n2t, d["grad_passes"], d["pass_ms_grad"]),
            ("loss", 2.0 * n2t, d["loss_passes"], d["pass_ms_loss"]), ("grady", (2.0 if wl["ortho"] else 4.0) * n2t, d["grady_passes"], d["pass_ms_grady"])]
    cand = [c for c in cand if c[2] > 0 and c[3] > 0]
    roof = None
    if cand:
        name, flops, cnt, ms = max(cand, key=lambda c: c[3])
        avg_ms = ms / cnt
        ach = flops / (avg_ms * 1e-3) / 1e12
        i8 = (name == "loss" and d["i8_loss_passes"] == d["loss_passes"]) or (name == "grady" and d["i8_grad_passes"] == d["grady_passes"])
        kname = {"loss": "loss_i8_kernel (LOSS pass of a line-search try + Y' store; tcgen05.mma kind::i8, 21 digit products)" if i8
                 else "rb_loss_kernel (LOSS + Y store, FP64 DMMA)",
                 "grady": "grad_i8_kernel (stored-Y gradient; tcgen05.mma kind::i8, 21 digit products)" if i8
                 else "rb_grady_kernel (stored-Y gradient, FP64 DMMA)"}.get(name, f"pass_kernel<{name}>")
        traffic, traffic_src = NCU_TRAFFIC_C3.get(name, (None, None)) if (n == 128 and i8) else (None, None)
        # LOSS: read the digit image of x1 (6.03 B / element) and write Y' (8 B); FP64 LOSS: read x1 (8 B), write Y'; gradient: read Y'
        # (the INT8 gradient kernel's two column halves each read Y': the second read is an L2 hit)
        alg_bytes = ((14.03125 if i8 else 16.0) if name == "loss" else 8.0) * n * t_local
        roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                "traffic": traffic * (t_local / 1e7) if traffic else None, "traffic_source": traffic_src,
                "algorithmic_bytes": alg_bytes, "avg_launch_ms": avg_ms, "launches": cnt, "flops_per_launch": flops,
                "peak_source": peak_source, "share_of_step": ms / dev_ms, "hbm_gbs": alg_bytes / (avg_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": HBM_PEAK_GBS[0], "hbm_peak_source": HBM_PEAK_GBS[1],
                "hbm_frac": alg_bytes / (avg_ms * 1e-3) / 1e9 / HBM_PEAK_GBS[0],
                "what": "achieved = ALGORITHMIC f64 flops of the pass (2 N^2 T_local) / mean launch duration; peak = FP64 tensor peak "
                        "(BASELINE.md's roofline).  The INT8 engines reach the f64 result through exact digit splitting, so frac can "
                        "exceed 1: the `engine` object gives the integer-op view; hbm_frac = algorithmic bytes / launch duration against the "
                        "measured copy bandwidth (the INT8 LOSS pass is within 2x of that bound too)"}
        if i8:
            ops = 21.0 * 2.0 * 128 * 128 * t_local
            roof["engine"] = {"what": "error-free balanced radix-256 splitting, 6 digits, 21 digit products, s32 accumulators in TMEM "
                                      "(tests/host/i8_split_check.cpp: within 3e-13 of f64)",
                              "int8_tops": ops / (avg_ms * 1e-3) / 1e12, "int8_peak_nominal_tops": 4500.0}
        roof["all_passes"] = {nm: {"launches": c, "avg_ms": m / c, "tflops_f64_equivalent": f / (m / c * 1e-3) / 1e12}
                              for nm, f, c, m in cand}
    pass_mix = {"fused": d["fused_passes"], "grad": d["grad_passes"], "loss": d["loss_passes"], "ls_tries": d["ls_tries"],
                "fallbacks": d["fallbacks"], "sign_changes": d["sign_changes"], "restarts": restarts, "grady": d["grady_passes"],
                "i8_loss": d["i8_loss_passes"], "i8_grad": d["i8_grad_passes"],
                "pass_ms": {"fused": d["pass_ms_fused"], "grad": d["pass_ms_grad"], "loss": d["pass_ms_loss"], "grady": d["pass_ms_grady"]}}
    pass_ms_total = d["pass_ms_fused"] + d["pass_ms_grad"] + d["pass_ms_loss"] + d["pass_ms_grady"]
    non_pass_ms_per_step = (dev_ms - pass_ms_total) / args.steps  # this rank's own split (pass times are per rank)
    state = core.state()
    core.close()

    # ---- parity beside the measurement: on the data that was just timed
    parity = None
    if not args.no_parity:
        import _data
        from oracle import oracle as orc
        wp = np.ascontiguousarray(_data.orthogonal(n, 7) + 0.01 * np.random.default_rng(11).standard_normal((n, n)))
        kind_c, alpha_c, want_h = wl["kind"], wl["alpha"], not wl["ortho"]

        def moments_dev(flags):
            gr = np.zeros((n, n)); sd = np.zeros(n); hr = np.zeros((n, n)); sq = np.zeros(n); lrow = np.zeros(n)
            stt = _ffi.Stats(); e2 = C.create_string_buffer(1024)
            rc = lib.picard_eval_moments_device_ex(C.c_void_p(x1_dev.data_ptr()), C.c_int64(n), C.c_int64(t_local), C.c_int64(ld), hp(wp),
                                                   C.c_int32(kind_c), C.c_double(alpha_c), C.c_int32(3), C.c_int32(int(want_h)),
                                                   C.c_int32(local_rank), C.c_uint32(flags), C.c_int32(int(wl["extended"])), C.c_int32(0),
                                                   None, hp(gr), hp(sd), hp(hr), hp(sq), hp(lrow), C.byref(stt), e2, C.c_size_t(1024))
            assert rc == 0, e2.value
            return dict(gr=gr, sd=sd, hr=hr, sq=sq, lrow=lrow), stt.as_dict()

        parity = {"tolerance": 1e-10, "definition": "max|a - b| / max|b| per quantity (BASELINE.json north_star)", "w": "orthogonal(seed 7) + 0.01 N(0,1)"}
        keys = ("gr", "sd", "lrow") + (("hr", "sq") if want_h else ())
        default, st_def = moments_dev(0)
        fp64, st_fp = moments_dev(P.FLAG_NO_INT8)
        parity["default_engine_vs_fp64_kernels_full_T"] = {
            "t_local": int(t_local), "engine": {"i8_loss_passes": st_def["i8_loss_passes"], "i8_grad_passes": st_def["i8_grad_passes"],
                                                "i8_fallbacks": st_def["i8_fallbacks"], "i8_range": st_def["i8_range"]},
            **{k_: _data.rel_err(default[k_], fp64[k_]) for k_ in keys}}
        # the same point twice more: every pass kernel reduces in a fixed order, so repeated launches must agree to the last bit
        again = [moments_dev(0)[0] for _ in range(2)]
        parity["default_engine_run_to_run"] = {"launches": 3, **{k_: max(_data.rel_err(a_[k_], default[k_]) for a_ in again) for k_ in keys}}
        if rank == 0:
            ts = int(min(t_local, 100_000))
            xs = x1_dev[:, :ts].cpu().numpy()
            ref = orc.eval_point(xs, wp, kind_c, alpha_c, ortho=False, extended=False)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import _gpu
            got, st_s = _gpu.eval_moments_ex(xs, wp, kind_c, alpha_c, mode=3, want_h=want_h, whitened=bool(wl["extended"]), device=local_rank)
            parity["default_engine_vs_oracle_slice"] = {"t": ts, "i8_loss_passes": st_s["i8_loss_passes"], "i8_grad_passes": st_s["i8_grad_passes"],
                                                        **{k_: _data.rel_err(got[k_], getattr(ref, k_)) for k_ in keys}}
        worst = max([v for blk in parity.values() if isinstance(blk, dict) for k_, v in blk.items() if k_ in keys] or [0.0])
        parity["worst"] = worst
        parity["ok"] = bool(worst <= 1e-10)
    del x1_dev

    # ---- e2e: the reference-facing call on host buffers (rank-local shard), H2D + whole fit + D2H timed
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty((n, t_local), dtype=torch.float64, pin_memory=True)
        x_host.copy_(x_dev[:, :t_local])
        torch.cuda.synchronize()
        del x_dev
        torch.cuda.empty_cache()
        xh = x_host.numpy()
        cfg2 = P.PicardConfig(density=cfg.density, ortho=wl["ortho"], extended=wl["extended"], w_init=None if "jade_it" in wl else w0,
                              jade_it=wl.get("jade_it"), comm=comm, device=local_rank)
        os.environ.setdefault("PICARD_TRACE", "1")  # stage timings of every e2e call on stderr (a few extra stream syncs)
        runs = []
        first_call_s = None
        for _rep in range(4):  # one untimed warm-up call, then three complete timed calls; the median is reported, all are listed
            res = None
            barrier()
            e_t0 = time.perf_counter()
            res = P.Picard.fit_with_config(xh, cfg2)
            gn = float(res.gradient_norm)  # result read on the host
            chk = float(res.sources[0, 0]) + float(res.sources[-1, -1])  # the sources are on the host
            barrier()
            e_s = time.perf_counter() - e_t0
            te = torch.tensor([e_s], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            if _rep == 0:  # the process's first fit also page-locks the arena the `sources` result lives in (seconds, once per process)
                first_call_s = float(te[0])
                continue
            runs.append((float(te[0]), res.stats, res.n_iterations, bool(res.converged), gn))
        runs.sort(key=lambda r: r[0])
        e_s, st_e, n_it_e, conv_e, gn = runs[1]
        e2e = {"value": n_it_e / e_s, "unit": UNIT, "h2d_bytes_per_step": st_e["h2d_bytes"] / max(n_it_e, 1),
               "d2h_bytes_per_step": st_e["d2h_bytes"] / max(n_it_e, 1), "iterations": n_it_e,
               "converged": conv_e, "gradient_norm": gn, "seconds": e_s, "seconds_all_runs": [r[0] for r in runs],
               "warmup_calls": 1, "first_call_seconds": first_call_s,
               "core_ms": st_e["core_ms"], "preprocess_ms": st_e["preprocess_ms"], "h2d_ms": st_e["h2d_ms"], "d2h_ms": st_e["d2h_ms"],
               "what": "Picard.fit_with_config on pinned host X (this rank's shard): H2D, centering, whitening, fit to convergence, "
                       "D2H of sources into the library's pinned result arena; iterations / wall seconds (median of 3 calls after one warm-up "
                       "call that also page-locks the arena: first_call_seconds)"}
        del res, runs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_oracle_rate(wl, args.cpu_t, 1, args.cpu_iters, os.cpu_count() or 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "host": host_description(),
               "sample": f"oracle core loop (C++ port of core.rs on OpenBLAS) N={n}, T={r['t_cpu']}: 1 warm-up + {r['iters']} timed outer "
                         f"iterations in {r['seconds']:.2f} s ({r['it_s_sample']:.4f} it/s on the sample), scaled linearly in T to T={t_total}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict(args, wl, world, t_total, shard_range(t_total, 0, world)[1] - shard_range(t_total, 0, world)[0]),
            "wall_ms_per_step": wall_ms / args.steps, "clocks": clocks, "e2e": e2e, "gpu_launches": int(d["kernel_launches"]),
            "roofline": roof, "cpu_baseline": cpu, "passes": pass_mix, "parity": parity,
            "non_pass_ms_per_step": non_pass_ms_per_step, "host_affinity": affinity,
            "state": {"n_iterations": state["n_iterations"], "gradient_norm": state["gradient_norm"], "loss": state["loss"]},
        }
        print(json.dumps(line), file=out, flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
