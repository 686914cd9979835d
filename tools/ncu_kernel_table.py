"""One line per kernel from the `ncu --page raw --csv` exports in profiles/: duration, achieved DRAM GB/s against the
measured HBM peak, issue-slot utilisation, pipes, registers.   python tools/ncu_kernel_table.py profiles/r2e_*_ncu_raw.csv"""
import csv
import json
import os
import sys

peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) else 6554.2
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
print(f"| kernel | duration | DRAM read + write | DRAM GB/s (% of measured {peak:.0f}) | warp instr | issue slots busy | ALU / FMA / tensor pipe | smem LSU wavefronts | regs | grid x block | warps/SM |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def val(k, to=None):
        v = float(r[col[k]].replace(",", "")) if k in col and r[col[k]] else float("nan")
        return v * scale.get(units[col[k]], 1) if to else v
    name = r[col["Kernel Name"]] if "Kernel Name" in col else os.path.basename(path)
    t = val("gpu__time_duration.sum", True)
    rd, wr = val("dram__bytes_read.sum", True), val("dram__bytes_write.sum", True)
    gbs = (rd + wr) / t / 1e9
    print(f"| `{name[:44]}` | {t * 1e3:.3f} ms | {rd / 1e6:.1f} + {wr / 1e6:.1f} MB | {gbs:.0f} ({100 * gbs / peak:.1f} %) | "
          f"{val('smsp__inst_executed.sum'):.3g} | {val('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} % | "
          f"{val('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.0f} / {val('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.0f} / "
          f"{val('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} % | "
          f"{val('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.0f} % | {val('launch__registers_per_thread'):.0f} | "
          f"{val('launch__grid_size'):.0f} x {val('launch__block_size'):.0f} | {val('sm__warps_active.avg.per_cycle_active'):.1f} |")
