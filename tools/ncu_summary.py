"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py <rep> <out.txt> [title] [bench_kernel_name]
With bench_kernel_name, also records dram bytes (read+write) per launch of the first captured kernel under that name in
profiles/ncu_traffic.json, which bench.py reports as roofline.traffic."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_tensor_cycles_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__warp_issue_stalled", "smsp__average_warp"]
with open(out, "w") as f:
    f.write(f"# {title}\n")
    for r in rows[2:]:
        f.write("-" * 100 + "\n")
        for h, u, v in zip(hdr, units, r):
            if any(h == w or h.startswith(w) for w in WANT) and not h.endswith((".min", ".max.pct", ".sum.pct")):
                if "pct_of_peak_sustained_elapsed" in h and not h.startswith(("gpu__dram", "lts__throughput", "l1tex__throughput", "sm__throughput")):
                    continue
                f.write(f"{h:75s} {v:>22s} {u}\n")
print(open(out).read()[:6000])

if len(sys.argv) > 4:
    import json, os
    name = sys.argv[4]
    r = rows[2]
    def col(metric):
        i = hdr.index(metric)
        v, u = float(r[i].replace(",", "")), units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[name] = col("dram__bytes_read.sum") + col("dram__bytes_write.sum")
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print("traffic", name, d[name])
