#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 proximal-gradient hot path.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # restated CPU baseline (the Julia reference cannot run here)

Workload (BASELINE.json metric: "prox-grad iterations/sec and fused-step HBM GB/s, Lasso n=10^8"):
FISTA + NormL1(lambda=1), n = 10^8 Float32 in total, gamma = 0.1, beta = 0.5, x, grad, z_prev ~ synthetic, the gradient
supplied as a resident buffer ("fused step only").  One step = one iteration of the inner loop: the fused
grad-step + prox + extrapolation kernel (pb_ffb_step, 5 vector streams = 20 B/element) followed by the per-iteration
scalar exchange and the host-side stop test of the driver loop.  The loop is the product's own (pb_solve, reached through
pa.FastForwardBackward(maxit=K, tol<0)(x0=..., f=LinearFunction(c), g=NormL1(1), gamma=0.1, extrapolation_sequence=
repeat(0.5)) on device-resident tensors); `--loop python` spells the same iteration out in this file instead.  With N GPUs
the iterate is row-sharded over the ranks (strong scaling: n is fixed).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TOTAL = 100_000_000
GAMMA, BETA, LAMBDA = 0.1, 0.5, 1.0
BYTES_PER_ELT = 20  # read x, grad, z_prev; write z, x_next (float32)
METRIC = "prox-grad iterations/sec (FISTA+NormL1 fused step, Lasso n=1e8 fp32)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", type=int, default=N_TOTAL)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--loop", default="native", choices=["native", "python"],
                   help="native = the library's driver loop pb_solve through pa.FastForwardBackward (default); "
                        "python = the same iteration spelled out in this file (pb_ffb_step + exchange per step)")
    p.add_argument("--exchange", default="device", choices=["device", "nccl", "memcpy"],
                   help="per-iteration scalar exchange: device = in-kernel NVLink push + pinned-flag poll (default); "
                        "nccl = all_gather + D2H copy; memcpy = cudaMemcpy read-back (N=1 only)")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    p.add_argument("--cpu-threads", type=int, default=0,
                   help="OpenMP threads of the CPU port (0 = all host cores this process may use; set explicitly because "
                        "torch.distributed.run exports OMP_NUM_THREADS=1 to its workers)")
    return p.parse_args()


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(n_per_gpu):
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch at `n` elements).  The kernel streams every element once, so
    a launch on n_per_gpu elements moves n_per_gpu / n of the captured bytes; null when there is no capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_step_ffb_l1_f32"]
        return float(t["dram_bytes_per_launch_at_n"]) * (n_per_gpu / float(t["n"]))
    except Exception:
        return None


def host_threads(args):
    if args.cpu_threads > 0:
        return args.cpu_threads
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, None
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.fh = open(self.path, "w")
        try:
            self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the C port of the oracle (unfused passes, all host threads)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_run(n, steps, warmup, seconds_budget=None, threads=0):
    """Time `steps` unfused FISTA iterations of the port on n elements (after `warmup`).  If seconds_budget is given the
    step count is reduced to fit it (at least 2).  Returns (seconds_per_step, steps_done, threads)."""
    from oracle import fb_port

    if threads > 0:
        fb_port.set_threads(threads)
    T = np.float32
    z, zp, grad = (np.empty(n, T) for _ in range(3))
    fb_port.fill(z, 3)
    fb_port.fill(zp, 4)
    fb_port.fill(grad, 5)
    port = fb_port.FistaPort(z, zp, grad, fb_port.PROX_L1, LAMBDA)
    for w in (port.x, port.grad_f_x, port.y, port.res, port.z_new):
        fb_port.fill(w, 7)          # first touch in parallel
    t0 = time.perf_counter()
    port.step(GAMMA, BETA)
    t_first = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        port.step(GAMMA, BETA)
    if seconds_budget is not None:
        steps = int(max(2, min(steps, seconds_budget / max(t_first, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        port.step(GAMMA, BETA)
    dt = (time.perf_counter() - t0) / steps
    return dt, steps, fb_port.num_threads()


def run_reference(args):
    """`--impl reference`: the restated CPU path (C port of the oracle; Julia is not installed, DESIGN.md says why),
    all host threads, same workload/config/metric; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    # bound the run: probe one step on the full size, shrink the sample if K+W steps would exceed ~150 s
    nthr = host_threads(args)
    dt_probe, _, threads = cpu_port_run(min(n, 10_000_000), 2, 1, threads=nthr)
    est = dt_probe * (n / min(n, 10_000_000)) * (args.steps + args.warmup)
    n_s = n if est <= 150 else max(1_000_000, int(n * 150 / est))
    dt, steps, threads = cpu_port_run(n_s, args.steps, args.warmup, threads=nthr)
    dt_full = dt * (n / n_s)
    val = 1.0 / dt_full
    sample = f"{steps} unfused FISTA iterations (14 vector passes each) on n={n_s} fp32" + ("" if n_s == n else f", time scaled by {n / n_s:.3g} to n={n}")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "iterations/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt_full * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "M3 Lasso FISTA fused-step-only, n=1e8 fp32, NormL1(1), gamma=0.1, beta=0.5, gradient supplied", "n": n},
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "restated CPU baseline (C port of oracle, OpenMP); the Julia reference cannot run in this image"},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# the CUDA path
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import proxb200 as pa
    from proxb200 import _lib as L
    from proxb200.host import Context, LocalComm, TorchDistComm, ptr, shard_bounds

    ctx = Context.get()
    from proxb200.host import DeviceExchangeComm

    if args.exchange == "device":
        comm = DeviceExchangeComm(ctx)
    elif world > 1:
        comm = TorchDistComm()
    else:
        comm = LocalComm()
    lo, hi = shard_bounds(args.n, world)[rank]
    n = hi - lo
    # synthetic data as a pure function of the GLOBAL element index (counter-based: csrc/util_kernels.cu, the generator of the CPU
    # port's fill): every N works on the same n-vector, so the per-iteration scalars can be compared bit for bit across N
    x, grad, z_prev = (torch.empty(n, device="cuda", dtype=torch.float32) for _ in range(3))
    for buf, seed in ((x, 3), (z_prev, 4), (grad, 5)):
        L.check(ctx.lib.pb_fill_counter(ctx.h, L.PB_F32, n, lo, seed, 1.0, ptr(buf)))
    z = torch.empty_like(x)
    x_next = torch.empty_like(x)
    desc = L.pb_prox(L.PB_PROX_L1, 0, LAMBDA, 0.0, None, None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"x": x, "z": z, "z_prev": z_prev, "x_next": x_next}

    def step(ev=None):
        s = state
        if ev:
            ev[0].record()
        L.check(ctx.lib.pb_ffb_step(ctx.h, L.PB_F32, n, ptr(s["x"]), ptr(grad), ptr(s["z_prev"]), GAMMA, BETA, C.byref(desc),
                                    None, ptr(s["z"]), None, ptr(s["x_next"])))
        if ev:
            ev[1].record()
        sc = comm.exchange(ctx)                       # per-iteration scalar read-back (+ all-gather at N > 1)
        s["x"], s["x_next"] = s["x_next"], s["x"]     # fast_forward_backward.jl:135 (already computed) / :136
        s["z_prev"], s["z"] = s["z"], s["z_prev"]
        return sc

    K = args.steps
    sampler = None
    multi_iter = False
    if args.loop == "native":
        # The product's own driver loop (pb_solve, csrc/solve.cu) through the public solver API, on device-resident tensors:
        # f = <c, x> makes c the "supplied gradient buffer", beta = 0.5 is a constant extrapolation sequence, tol < 0 never
        # stops, so K iterations = K launches of K2, each followed by the scalar exchange and the host stop test.
        import itertools

        f_lin = pa.LinearFunction(grad)
        kw = dict(x0=x, f=f_lin, g=pa.NormL1(LAMBDA), gamma=GAMMA, extrapolation_sequence=itertools.repeat(np.float32(BETA)),
                  comm=comm, n_global=args.n)
        warm = pa.FastForwardBackward(maxit=max(3, args.warmup), tol=-1.0, driver="native")
        warm(**kw)
        del warm
        solver = pa.FastForwardBackward(maxit=K, tol=-1.0, driver="native")
        solver.profile = True            # CUDA events inside pb_solve: around the K iterations and around every K2 launch
        sampler = ClockSampler(local) if rank == 0 else None
        launches0 = ctx.launches()
        barrier()
        zsol, its = solver(**kw)
        barrier()
        assert its == K
        launches = ctx.launches() - launches0
        tm = solver.last_timing
        ms_total, kern_ms = tm["loop_ms"], tm["step_kernel_ms"] / max(1, tm["step_kernel_launches"])
        last_res_inf = float(solver.last_state.res_norm_inf)
        parity = dict(solver.last_parity)
        parity["g_z"] = float(solver.last_state.g_z)
        multi_iter = bool(getattr(solver, "last_multi_iter_kernel", False))
        del zsol
    else:
        for _ in range(max(3, args.warmup)):
            step()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local) if rank == 0 else None
        launches0 = ctx.launches()
        barrier()
        e0.record()
        for k in range(K):
            sc = step(evs[k])
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = ctx.launches() - launches0
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        last_res_inf = sc.res_inf
        parity = {"res_inf": sc.res_inf, "res_sq": sc.res_sq, "gdr": sc.gdr, "gsum": sc.gsum, "g_z": float(np.float32(LAMBDA) * np.float32(sc.gsum))}
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total, kern_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms_max = float(t[0]), float(t[1])
    ms_per_step = ms_total / K
    value = 1e3 / ms_per_step

    # kernel-only duration (no in-kernel exchange, no read-back between launches): separates the HBM pass from C1
    fused_was_on = args.exchange == "device"
    if fused_was_on:
        L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_FUSED_EXCHANGE, 0))
    s = state
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize()
    k0.record()
    for _ in range(reps):
        L.check(ctx.lib.pb_ffb_step(ctx.h, L.PB_F32, n, ptr(s["x"]), ptr(grad), ptr(s["z_prev"]), GAMMA, BETA, C.byref(desc),
                                    None, ptr(s["z"]), None, ptr(s["x_next"])))
    k1.record()
    torch.cuda.synchronize()
    kern_only_ms = k0.elapsed_time(k1) / reps
    if fused_was_on:
        L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_FUSED_EXCHANGE, 1))
    tk = torch.tensor([kern_only_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    kern_only_ms = float(tk[0])

    # ---- e2e: the user-facing solver call on HOST buffers (upload, K iterations with per-iteration read-back, download) ----
    e2e = None
    if not args.no_e2e:
        del x, z, z_prev, x_next, state
        torch.cuda.empty_cache()
        x0_h = torch.empty(n, dtype=torch.float32).pin_memory()
        b_h = torch.empty(n, dtype=torch.float32).pin_memory()
        x0_h.normal_(generator=torch.Generator().manual_seed(10 + rank))
        b_h.normal_(generator=torch.Generator().manual_seed(20 + rank))
        solver = pa.FastForwardBackward(maxit=K, tol=-1.0)   # negative tol: never stop before maxit

        def solve():
            f = pa.SquaredDistance(b_h)                        # H2D of b inside the timed region
            return solver(x0=x0_h, f=f, g=pa.NormL1(LAMBDA), gamma=1.0, comm=comm, n_global=args.n)

        solve()       # warm-up: the same call once, untimed (the solver keeps its device work vectors for the next call of this size)
        barrier()
        t0 = time.perf_counter()
        zsol, its = solve()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        assert its == K and tuple(zsol.shape) == (n,)
        e2e = {
            "value": K / dt, "unit": "iterations/s",
            "h2d_bytes_per_step": int(2 * 4 * n * world / K), "d2h_bytes_per_step": int((4 * n * world + 128 * K * world) / K),
            "what": f"FastForwardBackward(maxit={K}, tol<0)(x0=host, f=SquaredDistance(host b), g=NormL1(1), gamma=1): pinned host x0,b "
                    f"-> device, {K} iterations (sqdist gradient kernel + fused step + scalar read-back each), solution -> host; wall clock, max over ranks",
            "seconds": dt, "driver": solver.last_driver, "phases": getattr(solver, "last_timing", None),
        }
        # the literal single-call form: one fused step through the C ABI with HOST buffers (N=1 only; PCIe bound)
        if world == 1:
            zp_h = torch.empty(n, dtype=torch.float32).pin_memory().normal_()
            zo_h = torch.empty(n, dtype=torch.float32).pin_memory()
            xo_h = torch.empty(n, dtype=torch.float32).pin_memory()
            sc_h = (C.c_double * L.PB_NSCALARS)()
            reps = 3

            def host_step():
                L.check(ctx.lib.pb_ffb_step_host(ctx.h, L.PB_F32, n, C.c_void_p(x0_h.data_ptr()), C.c_void_p(b_h.data_ptr()),
                                                 C.c_void_p(zp_h.data_ptr()), GAMMA, BETA, C.byref(desc), C.c_void_p(zo_h.data_ptr()),
                                                 C.c_void_p(xo_h.data_ptr()), sc_h))

            host_step()
            t0 = time.perf_counter()
            for _ in range(reps):
                host_step()
            dth = (time.perf_counter() - t0) / reps
            e2e["step_host_call"] = {"value": 1.0 / dth, "unit": "iterations/s", "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": 8 * n + 128,
                                     "what": "pb_ffb_step_host: x, grad, z_prev uploaded and z, x_next downloaded EVERY step (PCIe bound)"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = args.n      # full size: a smaller sample would sit in the host's last-level cache and flatter the CPU
        dt_s, steps_s, threads = cpu_port_run(n_s, 1000, 2, seconds_budget=args.cpu_seconds, threads=host_threads(args))
        cpu = {"value": 1.0 / (dt_s * args.n / n_s), "unit": "iterations/s", "cores": threads, "kind": "port",
               "sample": f"{steps_s} unfused iterations (14 vector passes) of the C port on n={n_s} fp32, time scaled x{args.n / n_s:.3g} to n={args.n}"}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_ELT * n / (kern_ms_max * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "M3 Lasso FISTA fused-step-only: n=1e8 fp32 total, NormL1(1), gamma=0.1, beta=0.5, gradient supplied as a resident buffer; "
                                   "step = one iteration of the inner loop = pb_ffb_step (K2) + per-iteration scalar exchange + host stop test" +
                                   (", driven by the library's loop pb_solve via pa.FastForwardBackward(maxit=K, tol<0, driver=native)" if args.loop == "native" else ", driven by a Python loop in bench.py"),
                       "loop": args.loop, "iterations_in_one_persistent_kernel": multi_iter,
                       "n": args.n, "n_per_gpu": n, "parallelism": f"row-shard x{world}", "exchange": args.exchange, "l2": "inputs exceed L2 (5 x %.0f MB streams per GPU)" % (4 * n / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(n),
                         "kernel": ("k_step_multi<float, L1> (csrc/step_multi.cu: ONE persistent launch loops over the K iterations; kernel_ms = its "
                                    "CUDA-event duration / K, i.e. per iteration including the in-kernel fold, exchange and stop test)") if multi_iter
                                   else "k_step<float, L1, EXTRAP> (pb_ffb_step)",
                         "kernel_ms": kern_ms_max,
                         "algorithmic_bytes_per_launch": BYTES_PER_ELT * n, "peak_source": peak_src,
                         "note": (("kernel_ms = CUDA-event duration of the ONE persistent launch that runs all K iterations (k_step_multi: streaming "
                                   "CTAs + a service CTA that folds the partials, exchanges the scalar block with the other GPUs inside the kernel and "
                                   "takes the stop decision), divided by K: every iteration's fold, exchange and stop test are inside it. ")
                                  if multi_iter else
                                  ("kernel_ms = CUDA-event duration of the K2 launches inside the timed region (pipelined driver: K2 writes per-CTA "
                                   "partials, a 1-CTA kernel on a side stream folds them and does the scalar exchange while the next K2 runs; "
                                   "--loop python / other exchanges: K2 includes the last-CTA fold (+ in-kernel exchange)). ")) +
                                 "kernel_only_* = the one-iteration K2 kernel (pb_ffb_step) launched back to back without exchange and read-back",
                         "kernel_only_ms": kern_only_ms, "kernel_only_achieved": BYTES_PER_ELT * n / (kern_only_ms * 1e-3) / 1e9,
                         "kernel_only_frac": BYTES_PER_ELT * n / (kern_only_ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "last_residual_inf": last_res_inf,
            # rank-combined reductions of the LAST iteration (norm(res, Inf), norm(res)^2, dot(grad, res), sum|z|, g(z)): the data
            # is a function of the global index and the sums are double-double, so these are bit-identical for N = 1, 2, 4, 8
            "parity": parity,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
