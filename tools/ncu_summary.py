"""Summarise ncu outputs into profiles/: launch-list shares and key metrics of a --set full capture."""
import csv
import subprocess
import sys
from collections import defaultdict


def launch_shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[hdr + 1:]:
        if r[mi] == "gpu__time_duration.sum":
            name = r[ki].split("(")[0][:60]
            tot[name] += float(r[vi].replace(",", ""))
            cnt[name] += 1
    s = sum(tot.values())
    out = ["%-62s %5s %12s %7s %10s" % ("kernel", "n", "total ms", "share", "avg ms")]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append("%-62s %5d %12.3f %6.1f%% %10.3f" % (k, cnt[k], v / 1e6, 100 * v / s, v / 1e6 / cnt[k]))
    return "\n".join(out)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full_metrics(rep):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    out = []
    for v in rows[2:]:            # one block per captured launch
        out.append("-" * 100)
        out.append("kernel: " + v[h.index("Kernel Name")])
        for i, n in enumerate(h):
            if n in KEYS or ("issue_stalled" in n and n.endswith("per_issue_active.ratio")):
                out.append("  %-88s %-14s %s" % (n, u[i], v[i]))
    return "\n".join(out)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        print("=" * 100)
        print(a)
        print(launch_shares(a) if a.endswith(".csv") else full_metrics(a))
