#!/usr/bin/env python
"""Benchmark of the wealth-consumption operator hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): SSY long-run-risk model on the (18,18,18,18)
grid, N = 104 976 states, dense fp64 transition matrix P (88.2 GB) resident in HBM.
A "step" is one evaluation of  T w = 1 + beta (a_row . P (a_col . w^theta))^(1/theta)
(one pass over P); the steps chain, w <- T w.  metric = T-operator evaluations/s.

N > 1 (torchrun, one rank per GPU): the same grid with P row-sharded over the ranks
(strong scaling); every application ends with an all-gather of the result slices.
torch.distributed (gloo) is used only to exchange the NCCL id / IPC handles and to
take the max over ranks; the product path is ctypes -> libsdfs_b200.so.

--impl reference: the reference itself (JAX) cannot be installed offline, so this
arm times the oracle's restatement of the reference's own arithmetic (the N^2-term
broadcast-and-sum of ssy_wc_ratio.py:116-148) on the host cores, on a bounded sample
of (l,k) slabs of the same workload, scaled to evaluations/s.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line.  NCCL prints its version banner to stdout when NCCL_DEBUG=VERSION (set on
# the GPU boxes) and honours NCCL_DEBUG_FILE only above that level: raise it to WARN and send the log to stderr.
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ssy_T_operator_evals_per_s"
UNIT = "evals/s"
# BASELINE.json's metric string, quoted verbatim next to the machine-readable name: its three parts are carried by
# `value` (evals/s), `time_to_fixed_point` (tol 1e-8) and `roofline` (HBM GB/s vs peak)
BASELINE_METRIC = "SSY T-operator evals/s & time-to-fixed-point (tol 1e-8); HBM GB/s vs peak"
DEFAULT_SHAPES = (18, 18, 18, 18)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shapes", default=",".join(map(str, DEFAULT_SHAPES)))
    ap.add_argument("--no-solve", action="store_true", help="skip the time-to-fixed-point solve")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-threads", type=int, default=0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# --------------------------------------------------------------------------- CPU legs
def cpu_reference_form(shapes, steps, warmup, threads):
    """Oracle restatement of the reference's own N^2 broadcast-and-sum, on a bounded sample:
    each step evaluates `threads` (l,k) slabs concurrently (NumPy releases the GIL)."""
    import oracle as O
    from oracle.operators import ssy_broadcast_slab
    from concurrent.futures import ThreadPoolExecutor
    ssy = O.SSY()
    op = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    w = np.full(shapes, 800.0)
    L, K = shapes[0], shapes[1]
    slabs = [(l, k) for l in range(L) for k in range(K)]
    threads = max(1, min(threads, len(slabs)))
    pool = ThreadPoolExecutor(threads)
    times = []
    pos = 0
    for it in range(warmup + steps):
        todo = [slabs[(pos + i) % len(slabs)] for i in range(threads)]
        pos += threads
        t0 = time.perf_counter()
        list(pool.map(lambda lk: ssy_broadcast_slab(op, w, *lk), todo))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    pool.shutdown()
    per_eval = np.mean(times) / threads * len(slabs)           # seconds per full T evaluation
    # factored (sum-factorised) CPU form of the same operator, for the other side of the comparison
    op.T(w)
    t0 = time.perf_counter()
    for _ in range(5):
        op.T(w)
    fact = (time.perf_counter() - t0) / 5
    return dict(value=1.0 / per_eval, unit=UNIT, cores=threads, kind="port",
                sample=f"{threads} of {len(slabs)} (l,k) slabs per step x {steps} steps of the reference's "
                       f"N^2 broadcast-sum (ssy_wc_ratio.py:116-148), scaled to a full evaluation",
                ms_per_step=float(np.mean(times) * 1e3),
                factored_evals_per_s=1.0 / fact,
                factored_note="same oracle, sum-factorised over the Kronecker factors (80 N bytes instead of 8 N^2)",
                host_cpus=os.cpu_count())


def run_reference(args, shapes):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.cpu_threads or min(os.cpu_count() or 1, 16)
    # bound the run: each step is `threads` slabs (~0.5 s each single-threaded)
    cb = cpu_reference_form(shapes, args.steps, args.warmup, threads)
    N = int(np.prod(shapes))
    line = {"impl": "reference", "metric": METRIC, "baseline_metric": BASELINE_METRIC, "value": cb["value"], "unit": UNIT,
            "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(shapes, args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "JAX/jaxopt/quantecon are not installable offline: this is the oracle port of the "
                    "reference arithmetic on host cores (kind=port), N=%d" % N}
    print(json.dumps(line))


def workload_config(shapes, gpus):
    N = int(np.prod(shapes))
    tag = "BASELINE configs[1]" if tuple(shapes) == tuple(DEFAULT_SHAPES) else "non-default grid via --shapes"
    return {"workload": f"SSY {tuple(shapes)} grid, N={N}, dense fp64 P ({8 * N * N / 1e9:.1f} GB) resident in HBM, "
                        f"w <- T w chained, w0=800 ({tag})",
            "shapes": list(shapes), "N": N,
            "parallelism": f"row-shard x{gpus}" if gpus > 1 else "single GPU",
            "l2": "inputs larger than L2 (P >> 126 MB), no flush needed",
            "solver": "newton + on-device BiCGSTAB (reference stopping rule), analytic JVP"}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args, shapes):
    import sdfs_via_autodiff_b200 as S
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    ctx = S.Context(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method="env://")
        from sdfs_via_autodiff_b200 import dist as sd
        sd.init_comm(ctx, rank, world, dist, max_N=int(np.prod(shapes)))
    S.Context._default = ctx
    N = int(np.prod(shapes))

    def barrier():
        ctx.sync()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    t0 = time.perf_counter()
    op = S.make_T_ssy(S.SSY(), shapes, storage="dense", ctx=ctx)      # device discretiser + dense expansion
    ctx.sync()
    build_s = time.perf_counter() - t0
    nloc = op.row_end - op.row_begin

    w = ctx.full(shapes, 800.0)
    for _ in range(args.warmup):
        w = op(w)
    # ---- timed region: K chained applications, inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.prof_enable(args.steps)
    n0 = ctx.launch_count
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        w = op(w)
    ms = ctx.timer_stop_ms()
    barrier()
    launches = ctx.launch_count - n0
    kern_ms, kern_n = ctx.prof_read()
    ctx.prof_enable(0)
    ms = max_over_ranks(ms)
    value = args.steps / (ms / 1e3)

    # ---- e2e through the public API with HOST buffers: h2d(w) -> T -> d2h(Tw) every step
    w_host = ctx.pinned_empty(shapes)            # page-locked host buffers (cudaMallocHost)
    out_host = ctx.pinned_empty(shapes)
    w_host[...] = 800.0
    for _ in range(2):
        op(w_host).numpy(out=out_host)
        w_host, out_host = out_host, w_host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        op(w_host).numpy(out=out_host)           # h2d(w) -> prologue + dense pass (+ all-gather) -> d2h(Tw)
        w_host, out_host = out_host, w_host
    ctx.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e = {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
           "ms_per_step": e2e_s / args.steps * 1e3}

    # ---- roofline of the dominant kernel (dense row-stream pass), per launch on this rank
    peak, peak_src = peaks()
    alg_bytes = 8.0 * nloc * N + 8.0 * N + 24.0 * nloc      # P rows + x + (a_row, w_out, ...) per row
    kern_avg_ms = kern_ms / max(1, kern_n)
    achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("shapes") == list(shapes) and tj.get("n_gpus") == world:
                traffic = tj.get("k_dense_apply_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_dense_apply<1> (TMA ring, fused T epilogue)",
                "kernel_ms": kern_avg_ms, "kernel_launches_timed": kern_n, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                "kernel_share_of_step": kern_avg_ms / (ms / args.steps)}

    # ---- time to fixed point (reported beside the throughput; not part of the timed K steps)
    solve = None
    if not args.no_solve:
        w0 = ctx.full(shapes, 800.0)
        # one untimed outer iteration: first launch of the cooperative loop kernel (module load, function
        # attributes, first peer stores into freshly mapped IPC memory) is a one-off 0.1-0.5 s
        with contextlib.redirect_stdout(sys.stderr):       # the reference's 'hit maximum iteration' warning is unconditional
            S.newton_solver(op, w0, tol=1e-8, max_iter=1, verbose=False)
        barrier()
        t0 = time.perf_counter()
        ws, k, info = S.newton_solver(op, w0, tol=1e-8, verbose=False, return_info=True)
        ctx.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        res = float(np.max(np.abs(np.asarray(op(ws)) - np.asarray(ws))))
        solve = {"algo": "newton+bicgstab(device), reference stopping rule (inner atol 1e-4)", "tol": 1e-8, "seconds": dt,
                 "outer_iters": int(k), "inner_iters": [int(x) for x in info["inner_iters"]],
                 "operator_applications": int(info["matvecs"]), "max_abs_Tw_minus_w": res,
                 "apps_per_s": info["matvecs"] / dt,
                 "warmup": "one untimed outer iteration (first launch of the loop kernel)"}
        # the reference's rule stops when BiCGSTAB returns its zero start (||Tw-w||_2 <= 1e-4); a solve that
        # really reaches max|Tw-w| <= 1e-8 needs a tighter inner tolerance:
        barrier()
        t0 = time.perf_counter()
        wt, kt, it_ = S.newton_solver(op, w0, tol=1e-8, bicgstab_atol=1e-9, krylov_rtol=1e-10, verbose=False,
                                      return_info=True)
        ctx.sync()
        dtt = max_over_ranks(time.perf_counter() - t0)
        rest = float(np.max(np.abs(np.asarray(op(wt)) - np.asarray(wt))))
        solve["tight"] = {"inner_atol": 1e-9, "inner_rtol": 1e-10, "seconds": dtt, "outer_iters": int(kt),
                          "operator_applications": int(it_["matvecs"]), "max_abs_Tw_minus_w": rest}

    extra = {}
    if world == 1:
        # factor-form operator on the same grid (same T, 80 N algorithmic bytes)
        kop = S.make_T_ssy(S.SSY(), shapes, storage="kron", ctx=ctx)
        wk = ctx.full(shapes, 800.0)
        for _ in range(3):
            wk = kop(wk)
        ctx.sync()
        ctx.timer_start()
        for _ in range(50):
            wk = kop(wk)
        kms = ctx.timer_stop_ms() / 50
        extra["factor_form"] = {"evals_per_s": 1e3 / kms, "ms": kms,
                                "note": "sum-factorised Kronecker apply of the same T (never materialises P)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = args.cpu_threads or min(os.cpu_count() or 1, 16)
        cpu = cpu_reference_form(shapes, 3, 1, threads)

    if rank == 0:
        line = {"metric": METRIC, "baseline_metric": BASELINE_METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(shapes, world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "time_to_fixed_point": solve, "operator_build_s": build_s, **extra}
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    shapes = tuple(int(s) for s in args.shapes.split(","))
    if args.impl == "reference":
        run_reference(args, shapes)
    else:
        run_ours(args, shapes)


if __name__ == "__main__":
    main()
