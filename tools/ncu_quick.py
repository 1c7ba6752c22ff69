"""Compact digest of one `ncu --set full` report: key counters, per-opcode stall samples.  usage: ncu_quick.py rep rods"""
import collections, csv, io, subprocess, sys
rep, rods = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
get = lambda k: float(vals[hdr.index(k)].replace(",", "")) if k in hdr else float("nan")
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]:
    print(f"{k} = {get(k)}")
print("inst/rod", get("smsp__inst_executed.sum") / rods, "shared wavefronts/rod", get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / rods)
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = get(k)
        if v > 0.05: print(f"  {k[34:-23]:28s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
h = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
tot = sum(int(r[isamp] or 0) for r in data)
byop, exop = collections.Counter(), collections.Counter()
for r in data:
    t = r[ia].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    byop[op] += int(r[isamp] or 0); exop[op] += int(r[iex] or 0)
print("total samples", tot)
for op, c in byop.most_common(14):
    print(f"  {op:10s} {100 * c / tot:5.1f}%  exec/rod {exop[op] / rods:7.1f}  samples/exec {c / max(exop[op] / rods, 1e-9):6.1f}")
