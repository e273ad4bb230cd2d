"""Experiment driver (GPU box): device time of pqp_setup's two GEMMs (convertToDual) at the config-C3 shape, warp-specialised pipeline
against the round-1 kernel.  usage: python tools/setup_probe.py [N] [M]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
prob, d = pqp.generate_testproblem(12346, M, N)
flop = 2.0 * N * M * M + 2.0 * N * N * M
ref = None
for ws, sym in (("1", "1"), ("1", "0"), ("0", "0")):
    os.environ["PQP_GEMM_WS"] = ws
    os.environ["PQP_GEMM_SYM"] = sym
    ms = []
    for rep in range(3):
        with pqp.Solver(d, prob) as s:
            ms.append(s.setup_gemm_ms)
            if rep == 2:
                Qd, _, _ = s.dual(want_gq=False)
    if ref is None:
        ref = Qd
    print(f"PQP_GEMM_WS={ws} PQP_GEMM_SYM={sym}: setup GEMMs {min(ms):.3f} ms = {flop / (min(ms) * 1e-3) / 1e12:.1f} TFLOP/s fp32-equivalent "
          f"({3 * flop / (min(ms) * 1e-3) / 1e12:.0f} TFLOP/s of tf32 MMAs); max |Qd - Qd(first)| / max|Qd| = {np.abs(Qd - ref).max() / np.abs(ref).max():.2e}", flush=True)
