"""Summarise ncu outputs into small text files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/x_launches.csv   -> per-kernel totals and share of the step
  python tools/ncu_summary.py raw gpurun_out/x.ncu-rep             -> selected metrics of a --set full capture"""
import collections
import csv
import subprocess
import sys


def launches(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        agg[r["Kernel Name"][:100]][0] += 1
        agg[r["Kernel Name"][:100]][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (ncu-serialised, cold cache: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e6:12.3f} ms  {100 * v[1] / tot:6.2f}%  n={v[0]:4d}  {k}")


KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers",
        "launch__grid_size", "launch__block_size", "launch__shared_mem", "sm__throughput.avg.pct", "sm__pipe_fma_cycles_active",
        "sm__inst_executed_pipe_fma.sum.pct", "sm__pipe_tensor", "sm__inst_executed_pipe_lsu.sum.pct", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp", "smsp__average_warps_issue_stalled", "gpu__dram_throughput", "sm__inst_executed_pipe_uniform",
        "smsp__cycles_active.avg")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(f"# kernel: {d.get('Kernel Name', '?')[:120]}")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) for k in KEYS):
                print(f"{h:90s} {v:>20s} {u}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
