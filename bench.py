#!/usr/bin/env python
"""bench.py -- the PQP hot path on B200, one JSON line (contract in the task statement, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4] [--iters I]

Workloads (BASELINE.json configs):
  c3 (default)  single condensed QP from the testing/test_generator.c distribution, N=8192 duals, M=2048, seed 12346,
                I=1000 fixed PQP updates per solve: the HBM-bound persistent-GEMV regime the north star's roofline
                target is stated on.  A *step* is one solve (I updates).  A single problem does not shard
                ("replicas only"): with N GPUs every rank solves its own replica.
  c2            same distribution, N=1024, M=512, seed 12345 (Q on-chip resident; barrier-latency bound).
  c1            the shipped example (N=28), 312 updates per solve = where PQP_CPU.c stops; latency only (config.us_per_solve).
  c4            batched MPC: 4096 states per GPU sharing one Hessian (horizon 30, 12 states, 4 inputs, N=480);
                a step is one batch solve of I updates; problems shard over ranks with no collective in the loop
                and one NCCL all-gather of U at the end of every step (inside the timed region).
  c5            the same shapes, 2^20 states IN TOTAL sharded contiguously over the ranks (strong scaling; BASELINE config 5).
The c3 line also carries the c4 numbers under "batched" and the c2 numbers under "onchip", so one default run reports every
single-GPU config of BASELINE.json.

value    whole-job throughput with inputs resident in HBM (PQP iterations/s for c2/c3, QP solves/s for c4).
e2e      the same through the C ABI with host buffers: pinned-host Fd/X -> device, solve, Y/U -> host, every step.
roofline algorithmic bytes (4*N*ldq + 16*N per iteration, or 2*N*(N-1) + 16*N when the symmetric Qd is iterated from its upper
         triangle; c4: 4*N^2*B flop per update) / the iteration kernel's own
         CUDA-event time (recorded by the library on its stream), against MEASURED_PEAKS.json (hbm_gbs; c4: the sustained
         dense-bf16 figure, the only measured tensor number); traffic = DRAM bytes per launch from the committed ncu capture.
cpu_baseline: the reference's own CPU code (oracle/_ref = PQP_CPU.c compiled where it lay) or, if that library did not
         travel, the oracle port; on a bounded sample; 1 thread (the reference has no threading).
--impl reference: the same CPU code alone, on ALL host threads (one independent problem per thread).
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # the shipped example (a correctness config, latency only): 312 updates = where PQP_CPU.c itself stops
    "c1": dict(kind="single", example=True, N=28, M=7, seed=0, iters=312),
    "c3": dict(kind="single", N=8192, M=2048, seed=12346),
    "c2": dict(kind="single", N=1024, M=512, seed=12345),
    "c4": dict(kind="batched", pH=30, nS=12, nI=4, B=4096, seed=2024),
    # BASELINE config 5: 2^20 independent MPC problems in total, sharded contiguously over the ranks (strong scaling)
    "c5": dict(kind="batched", pH=30, nS=12, nI=4, B_total=1 << 20, seed=2025),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def ncu_traffic(kernel):
    """dram bytes read+written per launch of `kernel`, from the committed `ncu --set full` summary (profiles/ncu_traffic.json,
    written by tools/ncu_summary.py); None if that kernel has not been captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except (OSError, ValueError):
        return None


def batched_roofline(kernel, tfl, k_ms, flop_iter, N, B, iters):
    """Tensor-bound roofline of the batched loop.  achieved = ALGORITHMIC flop (4*N^2*B per update: the two N x N x B
    contractions of PQP_CPU.c:608-609 in fp32 terms) / the kernel's own time; peak = the measured dense bf16 tensor throughput
    (MEASURED_PEAKS.json, sustained: the launch runs for tens of ms) -- the only measured tensor figure there is.  The int8
    kernel spends 6 digit-plane products per algorithmic product on padded 128-row tiles; `executed_int8_tops` is what the
    tensor pipe actually ran, against the nominal 4500 dense int8 TOP/s of a B200."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1500.0)))
    src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md ~1.5 PFLOP/s dense bf16)"
    out = {"bound": "tensor", "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak, "traffic": ncu_traffic(kernel),
           "peak_source": src, "kernel": kernel, "kernel_ms_per_step": k_ms, "flop_per_iteration": flop_iter,
           "note": "fp32-equivalent flop 4*N^2*B per update"}
    if kernel.startswith("batched_imma"):
        mt, nks = (N + 127) // 128, (N + 31) // 32
        nks = (nks + 2) // 3 * 3 if nks >= 3 else nks
        group = 64 if kernel.endswith("pair") else 32  # problems sharing one pass over the digit planes of Q
        groups = (B + group - 1) // group
        # per update and group: (matrix, M tile, K step) x three MMAs of N = 3g, 2g, g columns, M = 128, K = 32
        macs = 2 * mt * nks * 128 * (6 * group) * 32 * groups
        tops = 2.0 * macs * iters / (k_ms * 1e-3) / 1e12
        out.update(executed_int8_tops=tops, executed_frac_of_nominal_int8=tops / 4500.0,
                   note=out["note"] + "; 6 int8 digit-plane products per fp32 product, exact int32 accumulation in TMEM")
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


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
# reference arm / cpu baseline: the reference's CPU code on the box's host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_engine():
    from oracle.oracle import Oracle, Reference
    if Reference.available():
        return Reference(np.float32), "reference"
    return Oracle(np.float32), "port"


def single_problem_host(w):
    """Host-side instance + its dual, formed with numpy BLAS (input preparation for the CPU arm, never timed)."""
    import pqp_for_mpc_b200 as pqp
    if w.get("example"):
        prob, d = pqp.load_example(os.path.join(ROOT, "tests", "golden", "example"))
        nO = d.M
        Fp = prob["Fp1"].reshape(d.M, -1) @ prob["D"] + prob["Fp2"].reshape(d.M, -1) @ prob["x"] - prob["Fp3"]
        GQ = prob["Gp"].reshape(d.N, d.M) @ prob["Qp_inv"].reshape(d.M, d.M)
        Qd = GQ @ prob["Gp"].reshape(d.N, d.M).T
        Fd = GQ @ Fp + prob["Kp"]
        return prob, d, np.ascontiguousarray(Qd, np.float32), np.ascontiguousarray(Fd, np.float32)
    prob, d = pqp.generate_testproblem(w["seed"], w["M"], w["N"])
    q = np.diag(prob["Qp_inv"]).astype(np.float32)
    GQ = prob["Gp"] * q[None, :]
    Qd = GQ @ prob["Gp"].T
    Fd = GQ @ prob["Fp"] + prob["Kp"]
    return prob, d, np.ascontiguousarray(Qd, np.float32), np.ascontiguousarray(Fd, np.float32)


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


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_parallel_rate(engine, Qd, Fd, updates, threads):
    """`threads` independent replicas of the reference loop at once (the reference itself is single-threaded, PQP_CPU.c has no
    OpenMP/pthreads: the only way it uses more cores is more problems).  ctypes releases the GIL during the C call.
    Returns (aggregate updates/s, wall seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    if threads == 1:
        return cpu_single_rate(engine, Qd, Fd, updates)
    ys = [np.full(Fd.size, 1000.0, np.float32) for _ in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(lambda y: engine.iterate(y, Qd, Fd, updates), ys))
        dt = time.perf_counter() - t0
    return threads * updates / dt, dt


def run_reference(args, w, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = PQP_CPU.c compiled where it lay; the
    oracle port if that library did not travel) on the box's host cores, all of them, one independent problem per thread."""
    if rank != 0:
        return
    engine, kind = cpu_engine()
    threads = host_threads()
    out = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
           "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if w["kind"] == "single":
        prob, d, Qd, Fd = single_problem_host(w)
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
                   config={"workload": args.workload, "N": w["N"], "M": w["M"], "seed": w["seed"], "iters_per_step": upd,
                           "replicas": threads, "l2": "inputs larger than L2" if w["N"] >= 8192 else "n/a (CPU)"})
    else:
        from bench_problems import condensed_mpc
        prob, d, X = condensed_mpc(w["seed"], w["pH"], w["nS"], w["nI"], n_states=8)
        from oracle.oracle import Oracle
        o = Oracle(np.float32)
        Fp = o.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[0])
        Qd, Fd, _, _ = o.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
        upd = 2000
        for _ in range(args.warmup):
            cpu_parallel_rate(engine, Qd, Fd, 200, threads)
        t = 0.0
        for _ in range(args.steps):
            _, dt = cpu_parallel_rate(engine, Qd, Fd, upd, threads)
            t += dt
        value = (threads * upd * args.steps / t) / args.iters  # solves/s at `iters` updates per solve
        sample = (f"{threads} threads x {upd} updates of one N={d.N} problem per step, scaled to {args.iters} updates per solve")
        out.update(metric="qp_solves_per_sec", unit="solves/s", value=value, ms_per_step=1e3 * t / args.steps,
                   config={"workload": args.workload, "N": d.N, "M": d.M, "B": w.get("B", w.get("B_total")), "iters_per_solve": args.iters})
    out["cpu_baseline"] = {"value": out["value"], "unit": out["unit"], "cores": threads, "kind": kind, "sample": sample}
    out["e2e"] = {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out["gpu_launches"] = 0
    print(json.dumps(out), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, w, rank, world, local_rank):
    import torch
    import pqp_for_mpc_b200 as pqp

    if not torch.cuda.is_available() or pqp.device_count() == 0:
        raise SystemExit("bench.py: no B200 visible -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src, _ = measured_peaks()
    result = {}

    def timed(step_fn, stream_ptr, sampler=None):
        """W warm-ups, then exactly K steps bracketed by barrier+synchronize, CUDA events on the library's stream."""
        ext = torch.cuda.ExternalStream(stream_ptr, device=torch.device("cuda", local_rank))
        for _ in range(args.warmup):
            step_fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(args.steps):
            step_fn()
        e1.record(ext)
        e1.synchronize()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = max_over_ranks(e0.elapsed_time(e1))
        return ms, clocks

    if w["kind"] == "single":
        N, M = w["N"], w["M"]
        if w.get("example"):
            prob, d = pqp.load_example(os.path.join(ROOT, "tests", "golden", "example"))
        else:
            prob, d = pqp.generate_testproblem(w["seed"], M, N)
        s = pqp.Solver(d, prob, device=local_rank)
        ldq = (N + 31) // 32 * 32
        # device-resident leg: Fd and Y stay on the GPU
        s.solve(prob["x"][None] if w.get("example") else None, iters=1, status=False)
        Fd_host, _ = s.linear_terms(1)
        Fd_dev = torch.from_numpy(Fd_host[0]).cuda()
        Y_dev = torch.empty(N, dtype=torch.float32, device="cuda")
        kern_ms = []

        def step_dev():
            rc = pqp.lib().pqp_solve_dual(s.handle, pqp._as_ptr(Fd_dev.data_ptr()), 1, args.iters, None,
                                          pqp._as_ptr(Y_dev.data_ptr()), None)
            if rc:
                raise pqp.PQPError(rc, "pqp_solve_dual")
            kern_ms.append(s.last_solve_ms)

        l0 = s.launch_count
        ms, clocks = timed(step_dev, s.stream, ClockSampler(local_rank) if rank == 0 else None)
        launches = (s.launch_count - l0) * args.steps // (args.steps + args.warmup)
        k_ms = statistics.mean(kern_ms[args.warmup:])
        value = world * args.iters * args.steps / (ms * 1e-3)
        kernel = s.last_kernel
        # the upper-triangle loop (symmetric Qd) is charged with the strictly upper triangle only: 2N^2 bytes, not 4N^2
        sym = kernel.startswith("gemv_sym")
        sym_mb = ((N + 127) // 128) * ((N + 127) // 128 + 1) // 2 * 65536 / 1e6  # the unit array: 64 KB per tile of the triangle
        bytes_iter = (2.0 * N * (N - 1) + 16.0 * N) if sym else (4.0 * N * ldq + 16.0 * N)
        achieved = bytes_iter * args.iters / (k_ms * 1e-3) / 1e9

        # end-to-end leg: pinned host Fd in, Y out, through the C ABI every step
        Fd_pin = torch.from_numpy(Fd_host[0].copy()).pin_memory()
        Y_pin = torch.empty(N, dtype=torch.float32).pin_memory()
        st = np.zeros(1, pqp.STATUS_DTYPE)

        def step_e2e():
            rc = pqp.lib().pqp_solve_dual(s.handle, pqp._as_ptr(Fd_pin.data_ptr()), 1, args.iters, None,
                                          pqp._as_ptr(Y_pin.data_ptr()), pqp._as_ptr(st))
            if rc:
                raise pqp.PQPError(rc, "pqp_solve_dual")

        barrier()
        t0 = time.perf_counter()
        ms_e2e, _ = timed(step_e2e, s.stream)
        e2e_value = world * args.iters * args.steps / (ms_e2e * 1e-3)

        result.update(metric="pqp_iters_per_sec", unit="iterations/s", value=value, ms_per_step=ms / args.steps,
                      config={"workload": args.workload, "N": N, "M": M, "seed": w["seed"], "iters_per_step": args.iters,
                              "generator": ("example/*.txt as shipped" if w.get("example") else
                                            "testing/test_generator.c distribution, splitmix64-seeded"), "kernel": kernel,
                              "us_per_solve": 1e3 * ms / args.steps,
                              "parallelism": "replicas only (a single problem does not shard)" if world > 1 else "1 GPU",
                              "l2": (f"inputs larger than L2, no flush: the loop streams the upper triangle of the symmetric Qd as "
                                     f"{sym_mb:.0f} MB of 128x128 tiles (L2: 126 MB) on every one of the iters_per_step updates of a step "
                                     "(one step = one launch); what stays in L2 / shared memory between updates is the loop's own "
                                     "working set, by design" if sym and sym_mb > 126 else
                                     "upper triangle of Qd smaller than L2: on-chip/L2 resident by design" if sym else
                                     "inputs larger than L2: Q (268 MB)" if N >= 8192 else "Q smaller than L2: on-chip/L2 resident by design")},
                      e2e={"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": 4 * N,
                           "d2h_bytes_per_step": 4 * N + st.itemsize, "ms_per_step": ms_e2e / args.steps,
                           "call": "pqp_solve_dual(host Fd -> host Y, status)"},
                      roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "traffic": ncu_traffic(kernel), "peak_source": peak_src, "kernel": kernel, "kernel_ms_per_step": k_ms,
                                "bytes_per_iteration": bytes_iter, "frac_of_8TBs_nominal": achieved / 8000.0,
                                "algorithmic_bytes": ("strictly upper triangle of the symmetric Qd, 2N(N-1) B, + 16N B of vectors; frac > 1 means "
                                                      "the triangle is served from L2/shared memory, not HBM" if sym else "4*N*ldq + 16N B"),
                                "full_matrix_equivalent_gbs": (4.0 * N * ldq + 16.0 * N) * args.iters / (k_ms * 1e-3) / 1e9,
                                **({"note": "HBM is the roofline the contract names for this path; measured, the upper-triangle loop is "
                                            "issue/latency-bound at this size (DESIGN.md 8.2: ncu DRAM traffic is about half the algorithmic "
                                            "bytes, and removing 41% of it did not change the time)"} if sym else {})},
                      gpu_launches=int(launches), clocks=clocks)
        if rank == 0 and not args.no_cpu:
            engine, kind = cpu_engine()
            Qd_host, _, _ = s.dual(want_gq=False)
            upd = cpu_updates_for(N, 12.0)
            rate, dt = cpu_single_rate(engine, Qd_host, Fd_host[0], upd)
            result["cpu_baseline"] = {"value": rate, "unit": "iterations/s", "cores": 1, "kind": kind,
                                      "sample": f"{upd} updates of the same N={N} instance in {dt:.1f} s (per-call split setup included)"}
        s.close()

    if w["kind"] == "batched" or (args.workload == "c3" and not args.no_batched):
        wb = w if w["kind"] == "batched" else WORKLOADS["c4"]
        from bench_problems import condensed_mpc, shard_range
        strong = "B_total" in wb
        total = wb["B_total"] if strong else wb["B"] * world
        prob, d, Xall = condensed_mpc(wb["seed"], wb["pH"], wb["nS"], wb["nI"], n_states=total)
        lo, hi = shard_range(total, world, rank)
        if strong and total % world:
            raise SystemExit("bench.py: c5 needs a rank count that divides 2^20")
        B = hi - lo
        X = Xall[lo:hi]
        s = pqp.Solver(d, prob, device=local_rank, batch_capacity=B)
        X_dev = torch.from_numpy(X).cuda()
        Y_dev = torch.empty((B, d.N), dtype=torch.float32, device="cuda")
        U_dev = torch.empty((B, d.M), dtype=torch.float32, device="cuda")
        U_all = torch.empty((total, d.M), dtype=torch.float32, device="cuda") if world > 1 else None
        kern_ms = []

        def step_dev():
            rc = pqp.lib().pqp_solve_batch_primal(s.handle, pqp._as_ptr(X_dev.data_ptr()), None, B, args.iters, None,
                                                  pqp._as_ptr(Y_dev.data_ptr()), pqp._as_ptr(U_dev.data_ptr()), None)
            if rc:
                raise pqp.PQPError(rc, "pqp_solve_batch_primal")
            kern_ms.append(s.last_solve_ms)
            if dist is not None:  # the one collective of the path: final gather of U over NVLink
                dist.all_gather_into_tensor(U_all, U_dev)

        l0 = s.launch_count
        ms, clocks_b = timed(step_dev, s.stream, ClockSampler(local_rank) if (rank == 0 and w["kind"] == "batched") else None)
        launches = (s.launch_count - l0) * args.steps // (args.steps + args.warmup)
        k_ms = statistics.mean(kern_ms[args.warmup:])
        solves = world * B * args.steps / (ms * 1e-3)
        flop_iter = 4.0 * d.N * d.N * B
        tfl = flop_iter * args.iters / (k_ms * 1e-3) / 1e12

        X_pin = torch.from_numpy(X.copy()).pin_memory()
        Y_pin = torch.empty((B, d.N), dtype=torch.float32).pin_memory()
        U_pin = torch.empty((B, d.M), dtype=torch.float32).pin_memory()

        def step_e2e():
            rc = pqp.lib().pqp_solve_batch_primal(s.handle, pqp._as_ptr(X_pin.data_ptr()), None, B, args.iters, None,
                                                  pqp._as_ptr(Y_pin.data_ptr()), pqp._as_ptr(U_pin.data_ptr()), None)
            if rc:
                raise pqp.PQPError(rc, "pqp_solve_batch_primal")

        ms_e2e, _ = timed(step_e2e, s.stream)
        batched = {"metric": "qp_solves_per_sec", "unit": "solves/s", "value": solves, "ms_per_step": ms / args.steps,
                   "problem_iterations_per_sec": solves * args.iters,
                   "config": {"workload": "c5" if strong else "c4", "N": d.N, "M": d.M, "B_per_gpu": B, "B_total": total,
                              "iters_per_solve": args.iters,
                              "kernel": s.last_kernel, "parallelism": f"problems sharded over {world} GPU(s), all-gather of U per step"},
                   "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "solves/s",
                           "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": int(Y_pin.numel() * 4 + U_pin.numel() * 4),
                           "call": "pqp_solve_batch_primal(host X -> host Y, U)"},
                   "roofline": batched_roofline(s.last_kernel, tfl, k_ms, flop_iter, d.N, B, args.iters) | (
                       {"traffic": None} if strong else {}),  # the committed ncu capture is of the c4 launch
                   "gpu_launches": int(launches)}
        s.close()
        if w["kind"] == "batched":
            result.update(metric=batched["metric"], unit=batched["unit"], value=batched["value"], ms_per_step=batched["ms_per_step"],
                          config=batched["config"], e2e=batched["e2e"], roofline=batched["roofline"],
                          gpu_launches=batched["gpu_launches"], clocks=clocks_b)
            if rank == 0 and not args.no_cpu:
                engine, kind = cpu_engine()
                from oracle.oracle import Oracle
                o = Oracle(np.float32)
                Fp = o.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[0])
                Qd, Fd, _, _ = o.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
                rate, dt = cpu_single_rate(engine, Qd, Fd, 5000)
                result["cpu_baseline"] = {"value": rate / args.iters, "unit": "solves/s", "cores": 1, "kind": kind,
                                          "sample": f"5000 updates of one N={d.N} problem in {dt:.1f} s, scaled to {args.iters} updates/solve"}
        else:
            result["batched"] = batched

    if args.workload == "c3" and not args.no_batched:
        # the n=1024 single-problem config (BASELINE config 2) rides along: on-chip resident, exchange-latency bound, a few ms
        w2 = WORKLOADS["c2"]
        prob2, d2 = pqp.generate_testproblem(w2["seed"], w2["M"], w2["N"])
        s2 = pqp.Solver(d2, prob2, device=local_rank)
        s2.solve(iters=1, status=False)
        Fd2, _ = s2.linear_terms(1)
        Fd2_dev = torch.from_numpy(Fd2[0]).cuda()
        Y2_dev = torch.empty(w2["N"], dtype=torch.float32, device="cuda")
        Fd2_pin = torch.from_numpy(Fd2[0].copy()).pin_memory()
        Y2_pin = torch.empty(w2["N"], dtype=torch.float32).pin_memory()
        st2 = np.zeros(1, pqp.STATUS_DTYPE)

        def step2_dev():
            rc = pqp.lib().pqp_solve_dual(s2.handle, pqp._as_ptr(Fd2_dev.data_ptr()), 1, args.iters, None, pqp._as_ptr(Y2_dev.data_ptr()), None)
            if rc:
                raise pqp.PQPError(rc, "pqp_solve_dual")

        def step2_e2e():
            rc = pqp.lib().pqp_solve_dual(s2.handle, pqp._as_ptr(Fd2_pin.data_ptr()), 1, args.iters, None, pqp._as_ptr(Y2_pin.data_ptr()),
                                          pqp._as_ptr(st2))
            if rc:
                raise pqp.PQPError(rc, "pqp_solve_dual")

        ms2, _ = timed(step2_dev, s2.stream)
        ms2e, _ = timed(step2_e2e, s2.stream)
        result["onchip"] = {"metric": "pqp_iters_per_sec", "unit": "iterations/s", "value": world * args.iters * args.steps / (ms2 * 1e-3),
                            "ms_per_step": ms2 / args.steps,
                            "e2e": {"value": world * args.iters * args.steps / (ms2e * 1e-3), "unit": "iterations/s",
                                    "h2d_bytes_per_step": 4 * w2["N"], "d2h_bytes_per_step": 4 * w2["N"] + st2.itemsize},
                            "config": {"workload": "c2", "N": w2["N"], "M": w2["M"], "seed": w2["seed"], "iters_per_step": args.iters,
                                       "kernel": s2.last_kernel, "note": "Q (4 MB) lives in registers across 64 SMs: exchange-latency bound, no HBM roofline"}}
        s2.close()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {"metric": result.pop("metric"), "value": result.pop("value"), "unit": result.pop("unit"), "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": result.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update(result)
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
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=None, help="PQP updates per solve (default 1000 = NUM_ITER, PQP_CPU.c:24; c1: 312)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-batched", action="store_true", help="c3 only: skip the appended c4 leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...); "
              "running 1 rank", file=sys.stderr)
    w = WORKLOADS[args.workload]
    if args.iters is None:
        args.iters = w.get("iters", 1000)
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
