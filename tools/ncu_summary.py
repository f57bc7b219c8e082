"""Text summary of an `ncu --set full` report: one block per launch with time, grid, registers, DRAM bytes, pipe utilisations and the
top stall reasons (warps per issue-active cycle).
    ncu -i gpurun_out/step.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/rNN_ncu_full_step_kernels.txt"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))

    def f(key, scale=1.0, default=0.0):
        try:
            return float(d.get(key, default)) * scale
        except ValueError:
            return default

    stalls = []
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print(d.get("Kernel Name", "")[:110])
    print("    us=%.1f  grid=%sx%s  regs=%s  rd MB=%.1f  wr MB=%.1f  dram%%=%.1f  tensor%%=%.1f  xu%%=%.1f  fma%%=%.1f  issue%%=%.1f  warps%%=%.1f  inst=%.1fM" % (
        f("gpu__time_duration.sum"), d.get("launch__grid_size", "?"), d.get("launch__block_size", "?"), d.get("launch__registers_per_thread", "?"),
        f("dram__bytes_read.sum"), f("dram__bytes_write.sum"), f("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed") or f("dram__throughput.avg.pct_of_peak_sustained_elapsed") or f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        f("sm__warps_active.avg.pct_of_peak_sustained_active"), f("smsp__inst_executed.sum", 1e-6)))
    print("    stalls: " + ", ".join("%s=%.2f" % (n, v) for v, n in stalls[:4]))
    print()
