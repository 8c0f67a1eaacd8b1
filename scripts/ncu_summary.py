#!/usr/bin/env python
"""Read an .ncu-rep (ncu --set full) here, without a GPU, and write a per-kernel summary (markdown) for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_ncu_summary.md ["title"]"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global load t-stage wavefronts"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate % (LUT gathers)"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "L1 LSU writeback stage %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> L1 read bytes"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by L1"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared load wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "shared load bank conflicts"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes.sum.per_second", "DRAM GB/s"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    seen, picked = set(), []
    for r in data:  # last instance of each kernel (warm)
        pass
    for r in reversed(data):
        name = r[ki].split("(")[0]
        if name not in seen:
            seen.add(name)
            picked.append(r)
    picked.reverse()
    stall_cols = [(i, h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on`; read with "
                "`scripts/ncu_summary.py`. One launch per kernel (the last captured instance).\n\n" % title)
        for r in picked:
            f.write("## `%s`\n\n| metric | value |\n|---|---|\n" % r[ki].split("(")[0].replace("void ", ""))
            for m, label in METRICS:
                if m in hdr and r[hdr.index(m)] != "":
                    i = hdr.index(m)
                    f.write("| %s (`%s`) | %s %s |\n" % (label, m, r[i], units[i]))
            st = sorted(((float(r[i]), h) for i, h in stall_cols if r[i]), reverse=True)[:6]
            f.write("| top stall reasons (warps per issue) | %s |\n\n" % ", ".join(
                "%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                for v, h in st))
    print("wrote", out)


if __name__ == "__main__":
    main()
