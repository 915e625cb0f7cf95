"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed here.
usage: python profiles/summarize.py launches <csv>      -> per-kernel totals / shares
       python profiles/summarize.py rep <file.ncu-rep>   -> key metrics of each captured launch"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel device time (ncu gpu__time_duration.sum, cold-cache, serialised) — %s" % path)
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-62s n=%5d total=%11.1f us avg=%9.1f us share=%5.1f%%" % (k[:62], n, t, t / n, 100 * t / tot))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for D in rows[2:]:
        print("## %s  (%s)" % (D[H.index("Kernel Name")], path))
        for i, h in enumerate(H):
            if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h):
                print("%-78s %-14s %s" % (h, U[i], D[i]))


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2])
