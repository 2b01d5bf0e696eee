"""Summarise an .ncu-rep (raw page + SASS source page) into a short text: key metrics, instruction mix, stall mix."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in KEYS:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k} [{units[i]}]: {[r[i] for r in data]}")
print("-- warp stall reasons (samples)")
st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): [r[i] for r in data] for i, h in enumerate(hdr)
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
for k, v in sorted(st.items(), key=lambda kv: -int(kv[1][0] or 0)):
    print(f"   {k}: {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if "Warp Stall Sampling (All Samples)" in r]
if hi:
    byop, execs, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[hi[0] + 1:]:
        if len(r) < 6:
            continue
        try:
            s, ex = int(r[2]), int(r[5])
        except ValueError:
            continue
        toks = r[1].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
        op = op.split(".")[0]
        byop[op] += s
        execs[op] += ex
        tot += s
    print("-- SASS opcode mix (all profiled launches): samples%, warp-level instructions executed")
    allx = sum(execs.values())
    for op, c in execs.most_common(16):
        print(f"   {op:10s} exec {c:12d} ({100 * c / allx:5.1f}%)  stall-samples {100 * byop[op] / max(tot, 1):5.1f}%")
