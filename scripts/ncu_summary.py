"""Summarise an .ncu-rep (raw page CSV) into one line per kernel launch: duration, DRAM bytes / throughput, tensor-pipe and
issue utilisation, registers, achieved occupancy.  usage: ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
want = {
    "Kernel Name": "kernel", "gpu__time_duration.sum": "dur", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__inst_executed_pipe_tensor.sum": "tc_inst", "sm__issue_active.avg.pct_of_peak_sustained_active": "issue%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%", "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid", "launch__block_size": "block", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conf", "smsp__cycles_active.avg": "cyc",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma%",
}
idx = {h: i for i, h in enumerate(hdr)}
cols = [(k, v) for k, v in want.items() if k in idx]
print(" | ".join(v for _, v in cols))
def scale(val, unit, target):
    try:
        x = float(val.replace(",", ""))
    except ValueError:
        return val
    f = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit)
    if target == "us" and f:
        return f"{x * f:.1f}us"
    g = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit)
    if target == "MB" and g:
        return f"{x * g:.1f}MB"
    return f"{x:.1f}" if "." in val else val
for r in rows[2:]:
    out = []
    for k, v in cols:
        i = idx[k]
        val, unit = r[i], units[i]
        if v == "dur": val = scale(val, unit, "us")
        elif v in ("rd", "wr"): val = scale(val, unit, "MB")
        elif v == "kernel": val = val[:70]
        else: val = scale(val, unit, None)
        out.append(val)
    print(" | ".join(out))
