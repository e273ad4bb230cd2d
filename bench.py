#!/usr/bin/env python
"""bench.py -- the PQP hot path on B200, one JSON line (contract in the task statement, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|c1|c2|c3|c4|c5] [--iters I]

Default (--workload all): the line every N of the driver's 1/2/4/8-GPU sweep shares.
  top level  BASELINE config 5: 2^20 independent MPC problems IN TOTAL (horizon 30, 12 states, 4 inputs: N = 480 duals, M = 120),
             one Hessian, sharded contiguously over the ranks (strong scaling), I = 1000 PQP updates per solve, no collective
             in the loop, one NCCL all-gather of U per step inside the timed region (and inside e2e).  metric = QP solves/s.
  "single"   BASELINE config 3: one condensed QP from the testing/test_generator.c distribution, N = 8192 duals, M = 2048,
             seed 12346, I = 1000 updates per solve: the HBM-bound single-problem regime the north star's roofline target is
             stated on.  A single problem does not shard ("replicas only"): every rank solves its own replica.
  "batched_c4"  BASELINE config 4: 4096 states per GPU (weak scaling), the launch the committed ncu capture is of.
  "onchip"   BASELINE config 2: N = 1024, M = 512, seed 12345 (Q lives on chip; exchange-latency bound).
  "one_mpc_sized_problem"  ONE problem of config 4's size (N = 480, M = 120: what a single controller solves per period) on the
             one-cluster kernel (16 SMs, y exchanged through distributed shared memory): latency only.
  "setup"    the x-independent dual construction (convertToDual's two GEMMs on tcgen05 3xTF32) of the "single" instance.
--workload c1|c2|c3|c4|c5 makes that config the top level (ncu captures, sweeps); c1 = the shipped example, 312 updates.

value     whole-job throughput with inputs resident in HBM.
e2e       the same through the C ABI with HOST buffers: pinned X / Fd -> device, solve, Y, U -> host, every step; with more than
          one rank the all-gather of U and the read-back of the gathered U are inside it.
roofline  single problem: algorithmic bytes / the loop kernel's own CUDA-event time against MEASURED_PEAKS.json hbm_gbs.
          batched: the int8 digit-plane operations the algorithm needs / the loop kernel's time against the MEASURED dense
          int8 tcgen05 rate of this GPU (profiles/tensor_peaks_r2.json, tools/tensor_peak_probe.cu); the fp32-equivalent
          4*N^2*B flop per update and its ratio to the measured bf16 peak are reported beside it.
parity    errors of the timed instance against the CPU oracle (float) and its float64 twin, outside the timed region.
cpu_baseline  the reference's own CPU code (oracle/_ref = PQP_CPU.c compiled where it lay; the oracle port if that library did
          not travel) on a bounded sample, 1 thread (the reference has no threading).
--impl reference: that CPU code alone, on ALL host threads (one independent problem per thread); never loads the product library.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE  # noqa: E402  (pure numpy; loads no native library)

WORKLOADS = {
    # the shipped example (a correctness config, latency only): 312 updates = where PQP_CPU.c itself stops
    "c1": dict(kind="single", example=True, N=28, M=7, seed=0, iters=312),
    "c2": dict(kind="single", N=1024, M=512, seed=12345),
    "c3": dict(kind="single", N=8192, M=2048, seed=12346),
    "c4": dict(kind="batched", pH=30, nS=12, nI=4, B=4096, seed=2024),
    # ONE problem of config 4's size (what a single controller solves per period): the one-cluster kernel, latency only
    "c4s": dict(kind="single", N=480, M=120, seed=12347),
    # BASELINE config 5: 2^20 independent MPC problems in total, sharded contiguously over the ranks (strong scaling)
    "c5": dict(kind="batched", pH=30, nS=12, nI=4, B_total=1 << 20, seed=2025),
}
PARITY_UPDATES_SINGLE = 100   # oracle: 0.13 s per update at N = 8192
PARITY_STATES_BATCHED = 8


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_json(*parts):
    try:
        return json.load(open(os.path.join(ROOT, *parts)))
    except (OSError, ValueError):
        return {}


def ncu_traffic(kernel):
    """dram bytes read+written per launch of `kernel`, from the committed `ncu --set full` summary (profiles/ncu_traffic.json,
    written by tools/ncu_summary.py); None if that kernel has not been captured."""
    return load_json("profiles", "ncu_traffic.json").get(kernel)


def measured_peaks():
    d = load_json("MEASURED_PEAKS.json")
    if d:
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


def tensor_peaks():
    """Measured dense tcgen05 rates of this GPU class (tools/tensor_peak_probe.cu -> profiles/tensor_peaks_r2.json): int8 for the
    batched loop, tf32 for the setup GEMMs.  MEASURED_PEAKS.json only carries bf16."""
    t = load_json("profiles", "tensor_peaks_r2.json")
    m = load_json("MEASURED_PEAKS.json")
    bf16 = float(m.get("bf16_tflops_sustained", m.get("bf16_tflops", t.get("bf16_tops", 1500.0))))
    return {"i8": float(t.get("i8_tops", 2.0 * bf16)), "tf32": float(t.get("tf32_tops", 0.5 * bf16)), "bf16": bf16,
            "source": ("measured (profiles/tensor_peaks_r2.json: tcgen05 cta_group::1 M=128 N=256, operands resident in shared memory, all "
                       "SMs, >= 0.25 s per shape)" if t else "fallback (2x / 0.5x the bf16 figure)"),
            "bf16_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if m else "fallback"}


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def active_set(y, rel=1e-6):
    y = np.asarray(y, np.float64)
    return y > rel * np.abs(y).max()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
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
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# the CPU side: reference arm, cpu_baseline, parity.  Uses oracle/ (the checker) and numpy only.
# ---------------------------------------------------------------------------------------------------------
def cpu_engine():
    from oracle.oracle import Oracle, Reference
    if Reference.available():
        return Reference(np.float32), "reference"
    return Oracle(np.float32), "port"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def single_problem_host(w):
    """Host-side instance + its dual, formed with numpy (input preparation for the CPU arm, never timed).  Does not load the
    product library: the generator is its numpy port (bench_problems.generator_instance, bit-identical, tests/test_io_abi.py),
    the example is read by the oracle's own loader."""
    if w.get("example"):
        from oracle.oracle import Oracle
        o = Oracle(np.float32)
        p = o.load_example(os.path.join(ROOT, "tests", "golden", "example"))
        Fp = o.compute_fp(p["Fp1"], p["Fp2"], p["Fp3"], p["D"], p["x"])
        Qd, Fd, _, _ = o.convert_to_dual(p["Qp_inv"], p["Gp"], p["Kp"], Fp, 0.0)
        return np.ascontiguousarray(Qd, np.float32), np.ascontiguousarray(Fd, np.float32)
    from bench_problems import generator_instance
    prob, d = generator_instance(w["seed"], w["M"], w["N"])
    q = np.diag(prob["Qp_inv"]).astype(np.float32)
    GQ = prob["Gp"] * q[None, :]
    Qd = GQ @ prob["Gp"].T
    Fd = GQ @ prob["Fp"] + prob["Kp"]
    return np.ascontiguousarray(Qd, np.float32), np.ascontiguousarray(Fd, np.float32)


def batched_problem_host(wb, n_states):
    from bench_problems import condensed_mpc
    from oracle.oracle import Oracle
    prob, d, X = condensed_mpc(wb["seed"], wb["pH"], wb["nS"], wb["nI"], n_states=n_states, x_scale=BENCH_X_SCALE,
                               min_violated=BENCH_MIN_VIOLATED)
    o = Oracle(np.float32)
    Fp = o.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[0])
    Qd, Fd, _, _ = o.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
    return prob, d, X, Qd, Fd


def cpu_single_rate(engine, Qd, Fd, updates):
    """PQP updates/s of the reference loop (updateY2 + copyMatrix, PQP_CPU_test.c:717-744) on one core."""
    y = np.full(Fd.size, 1000.0, np.float32)
    t0 = time.perf_counter()
    engine.iterate(y, Qd, Fd, updates)
    dt = time.perf_counter() - t0
    return updates / dt, dt


def cpu_updates_for(N, seconds=6.0):
    per_it = 2.0e-9 * N * N  # ~0.13 s at N=8192 on one core
    setup = 6.0e-9 * N * N   # the split matrices are rebuilt per call
    return max(3, int(max(seconds - setup, 1.0) / per_it))


def cpu_parallel_rate(engine, Qd, Fd, updates, threads):
    """`threads` independent replicas of the reference loop at once (the reference itself is single-threaded, PQP_CPU.c has no
    OpenMP/pthreads: the only way it uses more cores is more problems).  ctypes releases the GIL during the C call.
    Returns (aggregate updates/s, wall seconds)."""
    if threads == 1:
        return cpu_single_rate(engine, Qd, Fd, updates)
    ys = [np.full(Fd.size, 1000.0, np.float32) for _ in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(lambda y: engine.iterate(y, Qd, Fd, updates), ys))
        dt = time.perf_counter() - t0
    return threads * updates / dt, dt


def run_reference(args, wname, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = PQP_CPU.c compiled where it lay; the
    oracle port if that library did not travel) on the box's host cores, all of them, one independent problem per thread.
    Same metric / unit / config as the product arm's top level; each step is a bounded sample of that workload."""
    if rank != 0:
        return
    w = WORKLOADS[wname]
    engine, kind = cpu_engine()
    threads = host_threads()
    out = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
           "scaling": "strong" if wname == "c5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if w["kind"] == "single":
        Qd, Fd = single_problem_host(w)
        # memory: every replica builds its own two dense split matrices (8 N^2 bytes), as the reference does per call
        threads = max(1, min(threads, int(24e9 // (8.0 * w["N"] * w["N"]))))
        upd = cpu_updates_for(w["N"], 4.0)
        for _ in range(args.warmup):
            cpu_parallel_rate(engine, Qd, Fd, max(1, upd // 4), threads)
        t = 0.0
        for _ in range(args.steps):
            _, dt = cpu_parallel_rate(engine, Qd, Fd, upd, threads)
            t += dt
        value = threads * upd * args.steps / t
        sample = (f"{threads} independent replicas of the N={w['N']} instance, {upd} PQP updates each per step (a full step is "
                  f"{args.iters}); includes the reference's per-call split setup")
        out.update(metric="pqp_iters_per_sec", unit="iterations/s", value=value, ms_per_step=1e3 * t / args.steps,
                   config={"workload": wname, "N": w["N"], "M": w["M"], "seed": w["seed"], "iters_per_step": upd,
                           "replicas": threads, "l2": "inputs larger than L2" if w["N"] >= 8192 else "n/a (CPU)"})
    else:
        prob, d, X, Qd, Fd = batched_problem_host(w, 8)
        upd = 2000
        for _ in range(args.warmup):
            cpu_parallel_rate(engine, Qd, Fd, 200, threads)
        t = 0.0
        for _ in range(args.steps):
            _, dt = cpu_parallel_rate(engine, Qd, Fd, upd, threads)
            t += dt
        value = (threads * upd * args.steps / t) / args.iters  # solves/s at `iters` updates per solve
        sample = (f"{threads} threads x {upd} updates of one N={d.N} problem per step, scaled to {args.iters} updates per solve")
        total = w.get("B_total", w.get("B", 0) * max(world, 1))
        out.update(metric="qp_solves_per_sec", unit="solves/s", value=value, ms_per_step=1e3 * t / args.steps,
                   config={"workload": wname, "N": d.N, "M": d.M, "B_total": total, "iters_per_solve": args.iters})
    out["cpu_baseline"] = {"value": out["value"], "unit": out["unit"], "cores": threads, "kind": kind, "sample": sample}
    out["e2e"] = {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out["gpu_launches"] = 0
    print(json.dumps(out), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class Ctx:
    """torch / distributed plumbing shared by the legs"""

    def __init__(self, args, rank, world, local_rank):
        import torch
        self.torch, self.args, self.rank, self.world, self.local_rank = torch, args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step_fn, stream_ptr, sampler=None, steps=None):
        """W warm-ups, then exactly K steps bracketed by barrier+synchronize, CUDA events on the library's stream (which waits
        for torch's stream at the end, so a trailing collective is inside), max over ranks."""
        torch = self.torch
        steps = steps or self.args.steps
        ext = torch.cuda.ExternalStream(stream_ptr, device=torch.device("cuda", self.local_rank))
        for _ in range(self.args.warmup):
            step_fn()
        self.barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            step_fn()
        ext.wait_stream(torch.cuda.current_stream())
        e1.record(ext)
        e1.synchronize()
        self.barrier()
        clocks = sampler.stop() if sampler else None
        return self.max_over_ranks(e0.elapsed_time(e1)), clocks


def single_leg(ctx, wname, *, sampler, cpu, parity, setup):
    """One single-problem workload (c1/c2/c3): device-resident value, e2e, HBM roofline, optional cpu_baseline / parity / setup."""
    import pqp_for_mpc_b200 as pqp
    torch, args, world = ctx.torch, ctx.args, ctx.world
    w = WORKLOADS[wname]
    N, M = w["N"], w["M"]
    iters = w.get("iters", args.iters)
    peak, peak_src, _ = measured_peaks()
    if w.get("example"):
        prob, d = pqp.load_example(os.path.join(ROOT, "tests", "golden", "example"))
    else:
        prob, d = pqp.generate_testproblem(w["seed"], M, N)
    s = pqp.Solver(d, prob, device=ctx.local_rank)
    ldq = (N + 31) // 32 * 32
    s.solve(prob["x"][None] if w.get("example") else None, iters=1, status=False)
    Fd_host, _ = s.linear_terms(1)
    Fd_dev = torch.from_numpy(Fd_host[0]).cuda()
    Y_dev = torch.empty(N, dtype=torch.float32, device="cuda")
    kern_ms = []

    def step_dev():
        rc = pqp.lib().pqp_solve_dual(s.handle, pqp._as_ptr(Fd_dev.data_ptr()), 1, iters, None, pqp._as_ptr(Y_dev.data_ptr()), None)
        if rc:
            raise pqp.PQPError(rc, "pqp_solve_dual")
        kern_ms.append(s.last_solve_ms)

    l0 = s.launch_count
    ms, clocks = ctx.timed(step_dev, s.stream, sampler)
    launches = (s.launch_count - l0) * args.steps // (args.steps + args.warmup)
    k_ms = statistics.mean(kern_ms[args.warmup:])
    value = world * iters * args.steps / (ms * 1e-3)
    kernel = s.last_kernel
    # the upper-triangle loop (symmetric Qd) is charged with the strictly upper triangle only: 2N^2 bytes, not 4N^2
    sym = kernel.startswith("gemv_sym")
    sym_mb = ((N + 127) // 128) * ((N + 127) // 128 + 1) // 2 * 65536 / 1e6  # the unit array: 64 KB per tile of the triangle
    bytes_iter = (2.0 * N * (N - 1) + 16.0 * N) if sym else (4.0 * N * ldq + 16.0 * N)
    achieved = bytes_iter * iters / (k_ms * 1e-3) / 1e9

    # end-to-end leg: pinned host Fd in, Y + status out, through the C ABI every step
    Fd_pin = torch.from_numpy(Fd_host[0].copy()).pin_memory()
    Y_pin = torch.empty(N, dtype=torch.float32).pin_memory()
    st = np.zeros(1, pqp.STATUS_DTYPE)

    def step_e2e():
        rc = pqp.lib().pqp_solve_dual(s.handle, pqp._as_ptr(Fd_pin.data_ptr()), 1, iters, None, pqp._as_ptr(Y_pin.data_ptr()), pqp._as_ptr(st))
        if rc:
            raise pqp.PQPError(rc, "pqp_solve_dual")

    ms_e2e, _ = ctx.timed(step_e2e, s.stream)
    block = dict(
        metric="pqp_iters_per_sec", unit="iterations/s", value=value, ms_per_step=ms / args.steps,
        config={"workload": wname, "N": N, "M": M, "seed": w["seed"], "iters_per_step": iters,
                "generator": ("example/*.txt as shipped" if w.get("example") else "testing/test_generator.c distribution, splitmix64-seeded"),
                "kernel": kernel, "us_per_solve": 1e3 * ms / args.steps,
                "parallelism": "replicas only (a single problem does not shard)" if world > 1 else "1 GPU",
                "l2": (f"inputs larger than L2, no flush: the loop streams the upper triangle of the symmetric Qd as {sym_mb:.0f} MB of 128x128 "
                       "tiles (L2: 126 MB) on every one of the iters_per_step updates of a step (one step = one launch); what stays in L2 / "
                       "shared memory between updates is the loop's own working set, by design" if sym and sym_mb > 126 else
                       "upper triangle of Qd smaller than L2: on-chip/L2 resident by design" if sym else
                       "inputs larger than L2: Q (268 MB)" if N >= 8192 else "Q smaller than L2: on-chip/L2 resident by design")},
        e2e={"value": world * iters * args.steps / (ms_e2e * 1e-3), "unit": "iterations/s", "h2d_bytes_per_step": 4 * N,
             "d2h_bytes_per_step": 4 * N + st.itemsize, "ms_per_step": ms_e2e / args.steps, "call": "pqp_solve_dual(host Fd -> host Y, status)"},
        roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(kernel),
                  "peak_source": peak_src, "kernel": kernel, "kernel_ms_per_step": k_ms, "bytes_per_iteration": bytes_iter,
                  "frac_of_8TBs_nominal": achieved / 8000.0,
                  "algorithmic_bytes": ("strictly upper triangle of the symmetric Qd, 2N(N-1) B, + 16N B of vectors; frac > 1 means the triangle "
                                        "is served from L2/shared memory, not HBM" if sym else "4*N*ldq + 16N B"),
                  "full_matrix_equivalent_gbs": (4.0 * N * ldq + 16.0 * N) * iters / (k_ms * 1e-3) / 1e9,
                  **({"note": "HBM is the roofline the contract names for this path; measured, the upper-triangle loop is issue/latency-bound at "
                              "this size (DESIGN.md 8.2: ncu DRAM traffic is about half the algorithmic bytes)"} if sym else {})},
        gpu_launches=int(launches))
    if clocks is not None:
        block["clocks"] = clocks
    if setup and not w.get("example"):
        tp = tensor_peaks()
        flop = 2.0 * N * M * M + 2.0 * N * N * M
        g_ms = s.setup_gemm_ms
        # a symmetric Qp_inv: only the 128 x 192 tiles that touch the upper triangle of Qd are multiplied (and mirrored)
        mirrored = s.setup_mirrored
        tiles_all = -(-N // 128) * -(-N // 192)
        tiles = sum(1 for ti in range(-(-N // 128)) for tj in range(-(-N // 192)) if 192 * tj + 191 >= 128 * ti) if mirrored else tiles_all
        done = 2.0 * N * M * M + 2.0 * N * N * M * tiles / tiles_all
        block["setup"] = {"what": "convertToDual's GEMMs on the device: GQ = Gp*Qp_inv, Qd = GQ*Gp' (PQP_CPU.c:492, :442), tcgen05 3xTF32"
                                  + ("; Qp_inv is symmetric, so Qd is built from the tiles of its upper triangle and mirrored" if mirrored else ""),
                          "ms": g_ms, "flop": flop, "tflops_fp32_equivalent": flop / (g_ms * 1e-3) / 1e12, "mirrored": bool(mirrored),
                          "tiles_multiplied": tiles, "tiles_of_the_full_product": tiles_all,
                          "roofline": {"bound": "tensor", "achieved": 3.0 * done / (g_ms * 1e-3) / 1e12, "peak": tp["tf32"], "unit": "TFLOP/s",
                                       "frac": 3.0 * done / (g_ms * 1e-3) / 1e12 / tp["tf32"], "peak_source": tp["source"],
                                       "note": "achieved counts the three tf32 products (hi*hi, hi*lo, lo*hi) per fp32 product of the tiles "
                                               "that were multiplied; tflops_fp32_equivalent counts the whole product"}}
    if ctx.rank == 0 and (cpu or parity):
        engine, kind = cpu_engine()
        Qd_host, _, _ = s.dual(want_gq=False)
        if parity:
            # the timed instance, fixed count, against the oracle (float) and its float64 twin -- outside every timed region.  The
            # float run doubles as the cpu_baseline sample (it is the reference's loop, split setup included).
            from oracle.oracle import Oracle
            K = min(PARITY_UPDATES_SINGLE, iters)
            Yk, _, _ = s.solve(Fd=Fd_host[0], iters=K, status=False)
            t0 = time.perf_counter()
            y32 = engine.solve_fixed(Qd_host, Fd_host[0], K)
            y32 = y32[0] if isinstance(y32, tuple) else y32
            dt = time.perf_counter() - t0
            y64, _ = Oracle(np.float64).solve_fixed(Qd_host, Fd_host[0], K)
            e_gf, e_gd, e_fd = relerr(Yk[0], y32), relerr(Yk[0], y64), relerr(y32, y64)
            block["parity"] = {"updates": K, "err_gpu_vs_f32_oracle": e_gf, "err_gpu_vs_f64_twin": e_gd, "err_f32_oracle_vs_f64_twin": e_fd,
                               "active_set_identical": bool(np.array_equal(active_set(Yk[0]), active_set(y32))),
                               "active": int(active_set(y32).sum()), "metric": "normwise ||a-b||inf/||b||inf on y",
                               "passes": bool(e_gf <= 1e-5 or (e_gf <= 2 * e_fd and e_gd <= max(e_fd, 1e-5))),
                               "rule": "err(gpu,f32) <= 1e-5, or err(gpu,f32) <= 2 err(f32,f64) and err(gpu,f64) <= max(err(f32,f64), 1e-5) (DESIGN.md 4)",
                               "oracle": kind}
            if cpu:
                block["cpu_baseline"] = {"value": K / dt, "unit": "iterations/s", "cores": 1, "kind": kind,
                                         "sample": f"{K} updates of the same N={N} instance in {dt:.1f} s (per-call split setup included)"}
        elif cpu:
            upd = cpu_updates_for(N, 12.0)
            rate, dt = cpu_single_rate(engine, Qd_host, Fd_host[0], upd)
            block["cpu_baseline"] = {"value": rate, "unit": "iterations/s", "cores": 1, "kind": kind,
                                     "sample": f"{upd} updates of the same N={N} instance in {dt:.1f} s (per-call split setup included)"}
    s.close()
    return block


def batched_leg(ctx, wname, *, sampler, cpu, parity):
    """One batched workload (c4 weak / c5 strong): problems sharded contiguously over the ranks, no collective in the loop, one
    all-gather of U per step inside the timed region and inside e2e."""
    import pqp_for_mpc_b200 as pqp
    from bench_problems import condensed_mpc, shard_range
    torch, args, world, rank, dist = ctx.torch, ctx.args, ctx.world, ctx.rank, ctx.dist
    wb = WORKLOADS[wname]
    iters = args.iters
    strong = "B_total" in wb
    total = wb["B_total"] if strong else wb["B"] * world
    if strong and total % world:
        raise SystemExit("bench.py: c5 needs a rank count that divides 2^20")
    prob, d, Xall = condensed_mpc(wb["seed"], wb["pH"], wb["nS"], wb["nI"], n_states=total, x_scale=BENCH_X_SCALE,
                                  min_violated=BENCH_MIN_VIOLATED)
    lo, hi = shard_range(total, world, rank)
    B = hi - lo
    X = Xall[lo:hi]
    del Xall
    s = pqp.Solver(d, prob, device=ctx.local_rank, batch_capacity=B)
    X_dev = torch.from_numpy(X).cuda()
    Y_dev = torch.empty((B, d.N), dtype=torch.float32, device="cuda")
    U_dev = torch.empty((B, d.M), dtype=torch.float32, device="cuda")
    U_all = torch.empty((total, d.M), dtype=torch.float32, device="cuda") if world > 1 else None
    kern_ms = []

    def step_dev():
        rc = pqp.lib().pqp_solve_batch_primal(s.handle, pqp._as_ptr(X_dev.data_ptr()), None, B, iters, None,
                                              pqp._as_ptr(Y_dev.data_ptr()), pqp._as_ptr(U_dev.data_ptr()), None)
        if rc:
            raise pqp.PQPError(rc, "pqp_solve_batch_primal")
        kern_ms.append(s.last_solve_ms)
        if dist is not None:  # the one collective of the path: final gather of U over NVLink
            dist.all_gather_into_tensor(U_all, U_dev)

    l0 = s.launch_count
    ms, clocks = ctx.timed(step_dev, s.stream, sampler)
    launches = (s.launch_count - l0) * args.steps // (args.steps + args.warmup)
    k_ms = statistics.mean(kern_ms[args.warmup:])
    solves = world * B * args.steps / (ms * 1e-3)
    kernel = s.last_kernel

    # end to end: pinned host X in; Y (this rank's duals) and U out.  With more than one rank U goes to the device buffer the
    # all-gather reads, and rank 0 reads the GATHERED U back to the host -- the result a caller of the sharded solve wants.
    X_pin = torch.from_numpy(X.copy()).pin_memory()
    Y_pin = torch.empty((B, d.N), dtype=torch.float32).pin_memory()
    U_pin = torch.empty((total if (world > 1 and rank == 0) else B, d.M), dtype=torch.float32).pin_memory()

    def step_e2e():
        u_out = U_dev.data_ptr() if dist is not None else U_pin.data_ptr()
        rc = pqp.lib().pqp_solve_batch_primal(s.handle, pqp._as_ptr(X_pin.data_ptr()), None, B, iters, None,
                                              pqp._as_ptr(Y_pin.data_ptr()), pqp._as_ptr(u_out), None)
        if rc:
            raise pqp.PQPError(rc, "pqp_solve_batch_primal")
        if dist is not None:
            dist.all_gather_into_tensor(U_all, U_dev)
            if rank == 0:
                U_pin.copy_(U_all, non_blocking=True)

    ms_e2e, _ = ctx.timed(step_e2e, s.stream)
    d2h = int(Y_pin.numel() * 4 + U_pin.numel() * 4)

    # roofline: the int8 digit-plane operations the algorithm needs (no tile padding) against the measured dense int8 rate
    tp = tensor_peaks()
    flop_iter = 4.0 * d.N * d.N * B                       # fp32-equivalent: the two N x N x B contractions of PQP_CPU.c:608-609
    structured = kernel.endswith("_paired")                # [[A,-A],[-A,A]] duals: half the products
    digit_ops_iter = 6.0 * flop_iter * (0.5 if structured else 1.0)
    tops = digit_ops_iter * iters / (k_ms * 1e-3) / 1e12
    tfl = flop_iter * iters / (k_ms * 1e-3) / 1e12
    block = dict(
        metric="qp_solves_per_sec", unit="solves/s", value=solves, ms_per_step=ms / args.steps, problem_iterations_per_sec=solves * iters,
        config={"workload": wname, "N": d.N, "M": d.M, "B_per_gpu": B, "B_total": total, "iters_per_solve": iters, "seed": wb["seed"],
                "states": f"x ~ N(0, {BENCH_X_SCALE:g}^2), at least {BENCH_MIN_VIOLATED} clearly violated rows each (bench_problems.py): no degenerate problem",
                "kernel": kernel, "parallelism": f"problems sharded contiguously over {world} GPU(s), no collective in the loop, all-gather of U per step",
                "l2": "operands L2-resident by design (1.5 MB of digit planes shared by all CTAs); X, Y, U, Fd of the batch stream through HBM once per solve"},
        e2e={"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": d2h,
             "ms_per_step": ms_e2e / args.steps,
             "call": "pqp_solve_batch_primal(host X -> host Y, U)" + (" + all_gather_into_tensor(U) + gathered U -> host (rank 0)" if world > 1 else "")},
        roofline={"bound": "tensor", "achieved": tops, "peak": tp["i8"], "unit": "TOP/s", "frac": tops / tp["i8"],
                  "traffic": ncu_traffic(kernel) if not strong else None, "peak_source": tp["source"], "kernel": kernel, "kernel_ms_per_step": k_ms,
                  "ops_per_iteration": digit_ops_iter,
                  "note": ("achieved = int8 digit-plane operations the algorithm needs per update (6 byte-plane products per fp32 product of the two "
                           "N x N x B contractions" + (", halved by the [[A,-A],[-A,A]] structure of the MPC dual" if structured else "") +
                           "; exact int32 accumulation in TMEM; no tile padding counted) / the loop kernel's own time"),
                  "fp32_equivalent_tflops": tfl, "fp32_equivalent_flop_per_iteration": flop_iter,
                  "fp32_equivalent_frac_of_bf16_peak": tfl / tp["bf16"], "bf16_peak": tp["bf16"], "bf16_peak_source": tp["bf16_source"]},
        gpu_launches=int(launches))
    if clocks is not None:
        block["clocks"] = clocks
    if rank == 0 and parity:
        # a seeded sample of the timed states against the oracle (float) and its float64 twin, and the whole shard's health
        from oracle.oracle import Oracle
        o32, o64 = Oracle(np.float32), Oracle(np.float64)
        Y = Y_pin.numpy()
        Qd, _, _ = s.dual()
        pool = min(B, 65536)                                # keep the read-back of Fd small: sample from the first 64 Ki states
        idx = np.sort(np.random.default_rng(7).choice(pool, min(PARITY_STATES_BATCHED, pool), replace=False))
        Fd, _ = s.linear_terms(int(idx.max()) + 1, want_fp=False)
        with ThreadPoolExecutor(min(16, host_threads())) as ex:
            r32 = list(ex.map(lambda b: o32.solve_fixed(Qd, Fd[b], iters)[0], idx))
            r64 = list(ex.map(lambda b: o64.solve_fixed(Qd, Fd[b], iters)[0], idx))
        e = np.array([[relerr(Y[b], a), relerr(Y[b], c), relerr(a, c)] for b, a, c in zip(idx, r32, r64)])
        same = [bool(np.array_equal(active_set(Y[b], 1e-5), active_set(a, 1e-5))) for b, a in zip(idx, r32)]
        act = (Y > 1e-5 * np.maximum(Y.max(axis=1, keepdims=True), 1e-30)).sum(axis=1)
        block["parity"] = {"updates": iters, "states_checked": [int(b) for b in idx],
                           "err_gpu_vs_f32_oracle": {"worst": float(e[:, 0].max()), "median": float(np.median(e[:, 0]))},
                           "err_gpu_vs_f64_twin": {"worst": float(e[:, 1].max()), "median": float(np.median(e[:, 1]))},
                           "err_f32_oracle_vs_f64_twin": {"worst": float(e[:, 2].max()), "median": float(np.median(e[:, 2]))},
                           "active_sets_identical": bool(all(same)), "metric": "normwise ||a-b||inf/||b||inf on y, per state",
                           "passes": bool(all(a <= 1e-5 or b <= 3 * max(c, 1e-5) for a, b, c in e)),
                           "rule": "per state: err(gpu,f32) <= 1e-5 or err(gpu,f64) <= 3 max(err(f32,f64), 1e-5) (tests/test_headline_parity_gpu.py)",
                           "shard": {"states": int(B), "non_finite": int((~np.isfinite(Y).all(axis=1)).sum()), "all_zero": int((Y == 0).all(axis=1).sum()),
                                     "active_constraints_median": float(np.median(act)), "active_fraction_median": float(np.median(act) / d.N)}}
    if rank == 0 and cpu:
        engine, kind = cpu_engine()
        from oracle.oracle import Oracle
        o = Oracle(np.float32)
        Fp = o.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[0])
        Qd0, Fd0, _, _ = o.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
        rate, dt = cpu_single_rate(engine, Qd0, Fd0, 5000)
        block["cpu_baseline"] = {"value": rate / iters, "unit": "solves/s", "cores": 1, "kind": kind,
                                 "sample": f"5000 updates of one N={d.N} problem in {dt:.1f} s, scaled to {iters} updates/solve"}
    s.close()
    del X_dev, Y_dev, U_dev, U_all, X_pin, Y_pin, U_pin
    torch.cuda.empty_cache()
    return block


def run_ours(args, rank, world, local_rank):
    import torch
    import pqp_for_mpc_b200 as pqp

    if not torch.cuda.is_available() or pqp.device_count() == 0:
        raise SystemExit("bench.py: no B200 visible -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    ctx = Ctx(args, rank, world, local_rank)
    cpu = not args.no_cpu
    top_name = "c5" if args.workload == "all" else args.workload
    smp = ClockSampler(local_rank) if rank == 0 else None
    if WORKLOADS[top_name]["kind"] == "batched":
        top = batched_leg(ctx, top_name, sampler=smp, cpu=cpu, parity=cpu)
    else:
        top = single_leg(ctx, top_name, sampler=smp, cpu=cpu, parity=cpu and top_name != "c1", setup=True)
    if args.workload == "all":
        top["single"] = single_leg(ctx, "c3", sampler=None, cpu=cpu, parity=cpu, setup=True)
        top["setup"] = top["single"].pop("setup", None)
        top["batched_c4"] = batched_leg(ctx, "c4", sampler=None, cpu=False, parity=cpu)
        top["onchip"] = single_leg(ctx, "c2", sampler=None, cpu=False, parity=False, setup=False)
        top["one_mpc_sized_problem"] = single_leg(ctx, "c4s", sampler=None, cpu=False, parity=False, setup=False)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if rank == 0:
        line = {"metric": top.pop("metric"), "value": top.pop("value"), "unit": top.pop("unit"), "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": top.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "strong" if top_name == "c5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update(top)
        print(json.dumps(line), file=_OUT, flush=True)


_OUT = sys.stdout


def own_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that whatever a library
    prints there (NCCL's version banner under NCCL_DEBUG=WARN/VERSION, INFO traces) cannot get in front of the line."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=None, help="PQP updates per solve (default 1000 = NUM_ITER, PQP_CPU.c:24; c1: 312)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and parity legs (they run the CPU oracle)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...); "
              "running 1 rank", file=sys.stderr)
    top_name = "c5" if args.workload == "all" else args.workload
    if args.iters is None:
        args.iters = WORKLOADS[top_name].get("iters", 1000)
    if args.impl == "reference":
        run_reference(args, top_name, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
