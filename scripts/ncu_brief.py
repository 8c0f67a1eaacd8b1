#!/usr/bin/env python
"""Print a compact per-launch table of the metrics that matter for the LUT kernels from an .ncu-rep (no GPU needed).
Usage: python scripts/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; data = rows[2:]
ki = hdr.index("Kernel Name")
M = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
     ("smsp__inst_executed.sum", "winst"),
     ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1data%"),
     ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ldreq"),
     ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ldsect"),
     ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "twave"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "dwave_lg"),
     ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
     ("lts__t_sectors.sum", "L2sect"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
     ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
     ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW")]
seen = {}
for r in data:
    seen[r[ki]] = r  # last instance of each distinct kernel
for name, r in seen.items():
    print(name[:110])
    print("   " + "  ".join("%s=%s" % (lab, r[hdr.index(m)][:12]) for m, lab in M if m in hdr))
