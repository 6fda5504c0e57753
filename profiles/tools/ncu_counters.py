"""Executed-instruction counters of the solver kernels from an ncu report -> JSON (profiles/counters.json).

usage: ncu_counters.py report.ncu-rep [more.ncu-rep ...] > counters.json

Per kernel (first launch found in each report): executed warp instructions per SASS opcode
(`sass__inst_executed_per_opcode`, which -- unlike `smsp__sass_thread_inst_executed_op_ffma_pred_on` -- sees the
packed FFMA2 / FMUL2 / FADD2 forms), and from them
    fp32 flop  = 32 lanes x (4 FFMA2 + 2 FMUL2 + 2 FADD2 + 2 FFMA + FMUL + FADD)   [thread-level predication ignored:
                 an upper bound by the divergent share, ~12 % in the reverse sweep]
    mufu ops   = 32 lanes x MUFU
plus duration, registers, issue slots used, FMA / XU pipe activity and DRAM bytes of that (profiled) launch.
bench.py divides the per-launch counts by its own live-measured launch duration.
"""
import json
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__issue_active.avg.per_cycle_active",
       "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    out = {}
    for rep in sys.argv[1:]:
        txt = ncu(rep, "--page", "raw", "--print-metric-instances", "details", "--metrics", "sass__inst_executed_per_opcode")
        # one block per profiled launch: a kernel header line followed by the metric table
        blocks = re.split(r"\n  (?=\S.*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\))", txt)
        import csv, io
        rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
        hdr = rows[0] if rows else []
        for n, blk in enumerate(blocks[1:]):
            name = re.match(r"(?:void )?([\w:]+<[^>]*>|[\w:]+)", blk.strip()).group(1)
            body = blk[blk.index("sass__inst_executed_per_opcode"):]
            total = int(re.search(r"sass__inst_executed_per_opcode\s+(\d+)", body).group(1))
            flat = re.sub(r"\s+", " ", body)
            ops = {m.group(1): int(m.group(2)) for m in re.finditer(r"([A-Z][A-Z0-9_]*): ?(\d+)", flat)}
            if name in out:
                continue
            g = lambda k: ops.get(k, 0)
            rec = {"report": rep.split("/")[-1], "inst_executed": total, "per_opcode": dict(sorted(ops.items(), key=lambda kv: -kv[1])),
                   "fp32_flop_executed": 32 * (4 * g("FFMA2") + 2 * g("FMUL2") + 2 * g("FADD2") + 2 * g("FFMA") + g("FMUL") + g("FADD")),
                   "mufu_lane_ops_executed": 32 * g("MUFU")}
            if len(rows) > 2 + n:
                r = rows[2 + n]
                for k in RAW:
                    if k in hdr:
                        try:
                            rec[k] = float(r[hdr.index(k)].replace(",", ""))
                        except ValueError:
                            pass
            out[name] = rec
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
