"""CPU test of bench.py's reference arm (`--impl reference`): it times the reference's own CPU code (oracle/_ref when it was
built here, else the oracle port) and must print ONE JSON line carrying the contract's keys.  The smallest workload (c1, the
shipped example) keeps it to a few seconds; no GPU is touched."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    # exactly one line on stdout, and nothing else: bench.py owns stdout and sends whatever else writes to fd 1 to stderr
    assert len(out.stdout.splitlines()) == 1, out.stdout[:500]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pqp_iters_per_sec" and d["unit"] == "iterations/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] == "c1" and d["config"]["N"] == 28
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_stdout_carries_only_the_json_line_even_if_a_library_prints():
    """A child writing to fd 1 behind Python's back (what NCCL's version banner does) must end up on stderr."""
    code = ("import os, sys; sys.argv = ['bench.py']; sys.path.insert(0, %r); import bench; bench.own_stdout(); "
            "os.write(1, b'NCCL version x.y\\n'); print('library noise'); print('{\"ok\": 1}', file=bench._OUT, flush=True)") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout == '{"ok": 1}\n', out.stdout
    assert "NCCL version x.y" in out.stderr and "library noise" in out.stderr


def test_reference_arm_never_loads_the_product_library():
    """`--impl reference` must time the reference's CPU code with none of the product in the process: the instance comes from the
    numpy port of the generator / the oracle's own loader, so libpqp_b200.so is never mapped (round-1 review, item 9)."""
    code = ("import sys, numpy as np; sys.argv = ['bench.py']; sys.path.insert(0, %r); import bench\n"
            "class A: gpus = 1; steps = 1; warmup = 0; iters = 1000\n"
            "bench._OUT = open('/dev/null', 'w')\n"
            "bench.cpu_updates_for = lambda N, seconds=6.0: 2\n"
            "for w in ('c1', 'c2', 'c4'):\n"
            "    a = A(); a.iters = bench.WORKLOADS[w].get('iters', 1000); bench.run_reference(a, w, 0, 1)\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libpqp_b200' not in maps and 'libpqp_compat' not in maps, 'product library mapped'\n"
            "assert 'libpqp_ref' in maps or 'libpqp_oracle' in maps\n"
            "print('ok')") % ROOT
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_default_reference_line_is_the_c5_metric():
    """With no --workload the reference arm reports the top-level metric of the product arm: batched QP solves/s on config 5."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    code = ("import sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; sys.path.insert(0, %r); import bench\n"
            "bench.cpu_parallel_rate = lambda engine, Qd, Fd, updates, threads: (100.0 * threads, updates / 100.0)\n"
            "bench.main()") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip())
    assert d["impl"] == "reference" and d["metric"] == "qp_solves_per_sec" and d["unit"] == "solves/s" and d["scaling"] == "strong"
    assert d["config"]["workload"] == "c5" and d["config"]["B_total"] == 1 << 20 and d["config"]["N"] == 480
