"""Summarise `ncu -i X.ncu-rep --page raw --csv`: one column per captured kernel, the metrics profiles/ quotes.
usage: python scripts/ncu_raw_summary.py raw.csv "header line 1" "header line 2" """
import csv
import sys

METRICS = ["Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "math_pipe_throttle", "wait", "not_selected", "branch_resolving",
          "lg_throttle", "membar", "dispatch_stall", "no_instruction", "sleeping"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
for line in sys.argv[2:]:
    print("# " + line)
names = [r[hdr.index("Kernel Name")] for r in data]
print(f"{'Kernel Name':88s} {'':16s} " + " | ".join(n.split('::')[-1][:40] for n in names))
for m in METRICS:
    if m in hdr:
        j = hdr.index(m)
        print(f"{m:88s} {units[j]:16s} " + " | ".join(r[j] for r in data))
for s in STALLS:
    m = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
    if m in hdr:
        j = hdr.index(m)
        print(f"{'stall ' + s + ' (warps per issue-active cycle)':88s} {'':16s} " + " | ".join(r[j] for r in data))
