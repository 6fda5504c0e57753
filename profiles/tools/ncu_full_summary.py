"""Per-kernel summary of an `ncu --set full` report (raw page): duration, launch shape, issue / pipe activity, DRAM
bytes, top stall reasons per issued instruction.

usage: ncu_full_summary.py report.ncu-rep "<header line(s)>" > summary.txt"""
import csv
import io
import subprocess
import sys

NAMES = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
         "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
         "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
         "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
         "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    print(sys.argv[2] if len(sys.argv) > 2 else rep)
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        d = dict(zip(hdr, zip(units, r)))
        print("\n== " + r[ki][:72])
        for n in NAMES:
            if n in d:
                print("  %-70s %s %s" % (n, d[n][1], d[n][0]))
        st = []
        for k, (u, v) in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("  stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
        for v, k in sorted(st, reverse=True)[:10]:
            print("    %-22s %.3f" % (k, v))


if __name__ == "__main__":
    main()
