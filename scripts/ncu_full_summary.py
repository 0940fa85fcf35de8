"""Summarise `ncu -i <rep> --page raw --csv` (a --set full capture) into a small JSON list: one entry per profiled
launch with duration, DRAM traffic, cache hit rates, SM utilisation, occupancy, registers and instruction count.
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv && python scripts/ncu_full_summary.py /tmp/raw.csv"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_inst",
}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        e = {"kernel": r[names.index("Kernel Name")][:70]}
        for col, key in WANT.items():
            if col in names:
                i = names.index(col)
                e[key] = f"{r[i]} {units[i]}".strip()
        # the three largest warp-stall reasons (cycles a warp waits per issued instruction, by reason)
        stalls = []
        for i, nm in enumerate(names):
            if nm.startswith("smsp__average_warps_issue_stalled_") and nm.endswith("_per_issue_active.ratio") and "not_issued" not in nm:
                try:
                    stalls.append((float(r[i].replace(",", "")), nm[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        e["top_stalls"] = ", ".join(f"{n} {v:.2f}" for v, n in stalls[:4])
        for col, key in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                         ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                         ("smsp__inst_executed_pipe_fp64.sum", "fp64_warp_insts"),
                         ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
                         ("launch__occupancy_limit_shared_mem", "occ_limit_smem"), ("launch__occupancy_limit_registers", "occ_limit_regs"),
                         ("sm__warps_active.avg.per_cycle_active", "warps_per_cycle")):
            if col in names:
                e[key] = r[names.index(col)]
        out.append(e)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
