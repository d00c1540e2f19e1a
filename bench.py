#!/usr/bin/env python
"""bench.py -- Picard iterations/sec on BASELINE.json's headline configuration.

Workload (config.workload): c3 = N=128 components, T=1e7 samples, f64, Picard-O extended tanh, mixed Laplace /
uniform sources mixed by a seed-42 N(0,1) matrix, explicit seed-43 orthogonal w_init.  With
--gpus N the T samples are sharded over N ranks (strong scaling, one NCCL allreduce of the packed N x N
moments per pass).

A "step" is one outer Picard iteration (core.rs:211-391): gradient pass + line-search passes over the
sample matrix + the N x N work between them.
  value  : iterations/sec of the core loop with the preprocessed data resident in HBM (picard_core_run),
           device time from CUDA events recorded on the library's stream, max over ranks.
  e2e    : iterations/sec through the reference-facing call Picard.fit_with_config on HOST buffers (pinned):
           H2D of X, centering, whitening, the whole fit to convergence, D2H of the sources -- all timed.
  roofline: the fused pass kernel (K1): algorithmic FP64 flops (4 N^2 T_local per launch)
           / its mean launch duration (CUDA events inside the library around each pass launch).
  cpu_baseline: the CPU oracle (a restatement of the reference's cost structure, oracle/picard_oracle.cpp)
           timed on this box's host cores on a bounded sample (N=128, T=T_cpu) and scaled linearly in T.

`--impl reference` times the CPU oracle as the reference arm (the Rust reference cannot be built in this
image: no cargo/rustc).
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
I8_LOSS_TRAFFIC = 19.233e9  # dram bytes per launch of loss_i8_kernel at c3 from the ncu --set full capture (profiles/summary_r01i.txt)
FP64_PEAK_TFLOPS = 37.19  # measured by us on this pool's B200 (profiles/microbench/fp64_pipes_r01.jsonl, DMMA m8n8k4);
#                           MEASURED_PEAKS.json has no FP64 figure


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


def cpu_oracle_rate(wl, t_cpu, max_iter, threads=None):
    """Iterations/sec of the CPU oracle core loop on a bounded sample, scaled linearly in T to the workload."""
    import numpy as np
    import _data
    from oracle import oracle as orc
    if threads:
        orc.set_threads(threads)
    n = wl["n"]
    kind = "mixed" if 0 < wl["n_laplace"] < n else ("laplace" if wl["n_laplace"] else "uniform")
    xw = _data.whitened(n, t_cpu, seed=42, kind=kind)
    t0 = time.perf_counter()
    r = orc.core_run(xw, wl["kind"], wl["alpha"], wl["ortho"], wl["extended"], max_iter=max_iter,
                     covariance=np.eye(n) if wl["extended"] else None, want_y=False)
    wall = time.perf_counter() - t0
    secs = r.seconds if r.seconds > 0 else wall
    it_s_sample = r.n_iterations / secs
    scaled = it_s_sample * (t_cpu / wl["t"])
    return dict(value=scaled, iters=r.n_iterations, seconds=secs, t_cpu=t_cpu, it_s_sample=it_s_sample,
                cores=orc.get_threads(), loss_evals=r.loss_evals)


def run_reference(args, wl, rank, out):
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    per_step = []
    total_it, total_s = 0, 0.0
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_oracle_rate(wl, args.cpu_t, args.cpu_iters, ncores)
        if i >= args.warmup:
            total_it += res["iters"]; total_s += res["seconds"]
            per_step.append(res["seconds"])
    value = (total_it / total_s) * (args.cpu_t / wl["t"])
    sample = (f"oracle core loop, N={wl['n']}, T={args.cpu_t} (1/{wl['t'] // args.cpu_t} of the workload), {args.cpu_iters} outer "
              f"iterations per step, iterations/s scaled linearly in T to T={wl['t']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(len(per_step), 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "n": wl["n"], "t_total": wl["t"], "t_per_gpu": wl["t"],
                   "parallelism": "host CPU, all cores (OpenBLAS threads); reference arm runs on rank 0 only",
                   "l2": "n/a (CPU)", "timing": "wall clock of the oracle core loop on the bounded sample, scaled linearly in T"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
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
    ap.add_argument("--cpu-t", type=int, default=100_000, help="samples of the CPU-baseline sample")
    ap.add_argument("--cpu-iters", type=int, default=6, help="outer iterations of the CPU-baseline sample")
    ap.add_argument("--flags", type=int, default=0, help="PICARD_FLAG_* bits (ablation)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, out)
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

    # ---- roofline of the dominant kernel: the fused pass (falls back to grad / loss variants if none ran)
    n2t = float(n) * n * t_local
    cand = [("fused", 4.0 * n2t, d["fused_passes"], d["pass_ms_fused"]), ("grad", 4.0 * n2t, d["grad_passes"], d["pass_ms_grad"]),
            ("loss", 2.0 * n2t, d["loss_passes"], d["pass_ms_loss"]), ("grady", (2.0 if wl["ortho"] else 4.0) * n2t, d["grady_passes"], d["pass_ms_grady"])]
    cand = [c for c in cand if c[2] > 0 and c[3] > 0]
    roof = None
    if cand:
        name, flops, cnt, ms = max(cand, key=lambda c: c[3])
        avg_ms = ms / cnt
        ach = flops / (avg_ms * 1e-3) / 1e12
        # the LOSS pass of a whitened 64 < N <= 128 problem runs on the INT8 tensor cores (i8_loss.cu) unless PICARD_I8=0
        i8 = name == "loss" and 64 < n <= 128 and os.environ.get("PICARD_I8", "") != "0"
        kname = {"loss": "loss_i8_kernel (LOSS + Y store on tcgen05.mma kind::i8, 28 slice products)" if i8 else "rb_loss_kernel (LOSS + Y store)",
                 "grady": "rb_grady_kernel (stored-Y gradient)"}.get(name, f"pass_kernel<{name}>")
        roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": ach / FP64_PEAK_TFLOPS,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture at c3 on one GPU
                # (profiles/summary_r01f.txt), scaled to this rank's share of the samples; null for kernels not captured
                "traffic": {"loss": I8_LOSS_TRAFFIC if i8 else 20.436e9, "grady": 10.248e9}.get(name, None) and
                {"loss": I8_LOSS_TRAFFIC if i8 else 20.436e9, "grady": 10.248e9}[name] * (t_local / 1e7) * (n / 128.0)
                if (n == 128) else None,
                # LOSS: read x1 (8 B / element; its 7-slice INT8 image is 7.06 B / element) and write Y' (8 B); gradient: read Y'
                "algorithmic_bytes": ((15.0625 if i8 else 16.0) if name == "loss" else 8.0) * n * t_local, "avg_launch_ms": avg_ms, "launches": cnt,
                "flops_per_launch": flops, "peak_source": "FP64 DMMA m8n8k4 microbenchmark measured by us "
                "(profiles/microbench/fp64_pipes_r01.jsonl); MEASURED_PEAKS.json has no FP64 entry",
                "share_of_step": ms / dev_ms, "hbm_gbs": 8.0 * n * t_local / (avg_ms * 1e-3) / 1e9}
        if i8:
            # achieved / peak above stay on the ALGORITHMIC f64 flops of the pass (2 N^2 T) against the FP64 tensor peak -- the
            # roofline BASELINE.md defines; what the tensor cores actually execute is 28 INT8 products of that shape:
            ops = 28.0 * 2.0 * 128 * 128 * t_local
            roof["engine"] = {"what": "error-free 7-slice INT8 splitting, s32 accumulators in TMEM, result within 2e-13 of f64 (tools/ozaki_numerics.py)",
                              "int8_tops": ops / (avg_ms * 1e-3) / 1e12, "int8_peak_nominal_tops": 4500.0,
                              "int8_ceiling_for_128x32x32_mma_tops": 1484.8,
                              "ceiling_source": "profiles/microbench/umma_i8_probe_r01.jsonl: 51 cycles per MMA whatever N <= 64; TMEM (512 columns) "
                                                "holds 7 level accumulators + the W' slices only for N = 32"}
    pass_mix = {"fused": d["fused_passes"], "grad": d["grad_passes"], "loss": d["loss_passes"], "ls_tries": d["ls_tries"],
                "fallbacks": d["fallbacks"], "sign_changes": d["sign_changes"], "restarts": restarts, "grady": d["grady_passes"],
                "pass_ms": {"fused": d["pass_ms_fused"], "grad": d["pass_ms_grad"], "loss": d["pass_ms_loss"], "grady": d["pass_ms_grady"]}}
    state = core.state()
    core.close()
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
        for _rep in range(3):  # three complete calls; the median is reported, all three are listed
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
            runs.append((float(te[0]), res.stats, res.n_iterations, bool(res.converged), gn))
        runs.sort(key=lambda r: r[0])
        e_s, st_e, n_it_e, conv_e, gn = runs[1]
        e2e = {"value": n_it_e / e_s, "unit": UNIT, "h2d_bytes_per_step": st_e["h2d_bytes"] / max(n_it_e, 1),
               "d2h_bytes_per_step": st_e["d2h_bytes"] / max(n_it_e, 1), "iterations": n_it_e,
               "converged": conv_e, "gradient_norm": gn, "seconds": e_s, "seconds_all_runs": [r[0] for r in runs],
               "core_ms": st_e["core_ms"], "preprocess_ms": st_e["preprocess_ms"], "h2d_ms": st_e["h2d_ms"], "d2h_ms": st_e["d2h_ms"],
               "what": "Picard.fit_with_config on pinned host X (this rank's shard): H2D, centering, whitening, fit to convergence, "
                       "D2H of sources; iterations / wall seconds (median of 3 calls)"}
        del res, runs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_oracle_rate(wl, args.cpu_t, args.cpu_iters, os.cpu_count() or 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"oracle core loop N={n}, T={r['t_cpu']}, {r['iters']} outer iterations in {r['seconds']:.2f} s "
                         f"({r['it_s_sample']:.3f} it/s), scaled linearly in T to T={t_total}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "n": n, "t_total": t_total, "t_per_gpu": t_local,
                       "parallelism": f"sample-sharded x{world}", "l2": "inputs (8*N*T_local bytes per pass) far exceed the 126 MB L2",
                       "timing": "CUDA events on the library stream around picard_core_run; wall clock cross-check in wall_ms_per_step"},
            "wall_ms_per_step": wall_ms / args.steps, "clocks": clocks, "e2e": e2e, "gpu_launches": int(d["kernel_launches"]),
            "roofline": roof, "cpu_baseline": cpu, "passes": pass_mix,
            "state": {"n_iterations": state["n_iterations"], "gradient_norm": state["gradient_norm"], "loss": state["loss"]},
        }
        print(json.dumps(line), file=out, flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
