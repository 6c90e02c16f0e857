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


def gather_json(rep, source, particles):
    """profiles/gather_ncu.json: the figures of the gather launches that bench.py copies into its roofline object."""
    import json
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    g = [r for r in rows[2:] if "k_gather_sticks" in r[h.index("Kernel Name")]]
    col = lambda r, n: float(r[h.index(n)].replace(",", ""))
    avg = lambda n: sum(col(r, n) for r in g) / len(g)
    tot = lambda n: sum(col(r, n) for r in g)
    # time-weighted fractions (a longer launch counts more)
    tw = lambda n: sum(col(r, n) * col(r, "gpu__time_duration.sum") for r in g) / tot("gpu__time_duration.sum")
    unit_ms = 1e-6 if rows[1][h.index("gpu__time_duration.sum")] in ("ns", "nsecond") else (1e-3 if rows[1][h.index("gpu__time_duration.sum")].startswith("us") else 1.0)
    byte_unit = lambda n: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[rows[1][h.index(n)]]
    d = {"source": source, "kernel": "k_gather_sticks<4,cls,flags>", "launches": len(g),
         "dram_bytes_per_launch": (tot("dram__bytes_read.sum") * byte_unit("dram__bytes_read.sum") + tot("dram__bytes_write.sum") * byte_unit("dram__bytes_write.sum")) / len(g),
         "l1_data_pipe_frac": tw("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100,
         "issue_active_frac": tw("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
         "dram_frac": tw("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") / 100,
         "warp_instructions_per_particle": tot("smsp__inst_executed.sum") / particles,
         "gather_ms_per_%d_particles_under_ncu" % particles: tot("gpu__time_duration.sum") * unit_ms,
         "registers_per_thread": int(avg("launch__registers_per_thread")),
         "l1_hit_rate": avg("l1tex__t_sector_hit_rate.pct") / 100, "l2_hit_rate": avg("lts__t_sector_hit_rate.pct") / 100}
    return json.dumps(d, indent=1)


def hot_loop(rep, launch_skip):
    """Per-instruction view of the hottest loop of one captured launch (source page): executions, active lanes, L1 tag
    requests of the global loads and shared-memory wavefronts per execution, and their sum per loop iteration."""
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch_skip), "--launch-count", "1"],
                                  stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(txt.splitlines()))
    name = rows[0][1] if len(rows[0]) > 1 else ""
    h = rows[1]
    data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
    if len(data) % 2 == 0 and data[0][0] == data[len(data) // 2][0]:
        data = data[:len(data) // 2]          # ncu prints the listing twice when source correlation is on
    c = lambda n: h.index(n)
    ex = [int(r[c("Instructions Executed")]) for r in data]
    # the hot loop: the longest run of instructions executed at least 60 % as often as the most executed LDS/LDG
    mem = [i for i, r in enumerate(data) if ("LDG" in r[c("Source")] or "LDS" in r[c("Source")])]
    top = max(ex[i] for i in mem)
    hot = [i for i in range(len(data)) if ex[i] >= 0.6 * top]
    lo, hi = hot[0], hot[-1]
    out = ["kernel: %s" % name, "hot loop: SASS %d..%d, %d instructions, executed %.3f M times per launch" % (lo, hi, hi - lo + 1, top / 1e6),
           "%-58s %9s %8s %9s %9s" % ("instruction", "lanes on", "L1 tags", "sh. wavef", "ideal")]
    tg = ws = wi = 0.0
    for i in range(lo, hi + 1):
        r = data[i]
        src = r[c("Source")].strip()
        if not any(k in src for k in ("LDG", "LDS", "STS", "STG", "RED", "ATOM")) or ex[i] == 0:
            continue
        e = ex[i]
        t, w, wid = int(r[c("L1 Tag Requests Global")]) / e, int(r[c("L1 Wavefronts Shared")]) / e, int(r[c("L1 Wavefronts Shared Ideal")]) / e
        tg, ws, wi = tg + t * e / top, ws + w * e / top, wi + wid * e / top
        out.append("%-58s %9.1f %8.2f %9.2f %9.2f" % (src[:58], int(r[c("Predicated-On Thread Instructions Executed")]) / e, t, w, wid))
    n_inst = sum(ex[lo:hi + 1]) / top
    out.append("per iteration: %.0f warp instructions (= %.0f issue cycles on 4 schedulers), %.1f L1 tag requests + %.1f shared wavefronts (ideal %.1f) = %.1f L1 data-pipe wavefronts"
               % (n_inst, n_inst / 4, tg, ws, wi, tg + ws))
    return "\n".join(out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hot-loop":         # hot-loop <rep> <launch index>
        print(hot_loop(sys.argv[2], int(sys.argv[3])))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gather-json":      # gather-json <rep> <particles> <source text>
        print(gather_json(sys.argv[2], sys.argv[4], int(sys.argv[3])))
        sys.exit(0)
    for a in sys.argv[1:]:
        print("=" * 100)
        print(a)
        print(launch_shares(a) if a.endswith(".csv") else full_metrics(a))
