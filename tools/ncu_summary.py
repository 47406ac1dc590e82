"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: the handful of counters DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py raw.csv [more.csv ...]"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print("kernel:", d["Kernel Name"], " grid", d["Grid Size"], " block", d["Block Size"])
        for k in KEYS:
            hit = [n for n in names if n == k or n.endswith("." + k)]
            if hit and d[hit[0]] != "":
                print(f"  {k:82s} {d[hit[0]]:>16s} {u[hit[0]]}")
        st = []
        for n in names:
            if STALL in n and n.endswith("_per_warp_active.pct") is False and n.endswith(".ratio"):
                try:
                    st.append((float(d[n]), n.split(STALL)[1].split("_per_")[0]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        if st:
            print("  stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))


for p in sys.argv[1:]:
    print("==", p)
    main(p)
