#!/usr/bin/env python
"""Write profiles/traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the bench's
kernels, from one `ncu --set full` capture taken at the bench's launch size.  bench.py copies the dominant kernel's
figure into roofline.traffic.  Usage: python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys

MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def bench_name(kernel):
    if "lut_stage_cell_kernel<1" in kernel or "lut_stage_kernel<1" in kernel:
        return "lut_stage1"
    if "lut_stage" in kernel:
        return "lut_stage2"
    if "resize_sr" in kernel:
        return "resize_sr"
    return kernel.split("(")[0]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in data:
        name = r[col["Kernel Name"]]
        rd = float(r[col["dram__bytes_read.sum"]]) * MULT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * MULT[units[col["dram__bytes_write.sum"]]]
        dur = float(r[col["gpu__time_duration.sum"]]) * TIME.get(units[col["gpu__time_duration.sum"]], 1.0)
        res[bench_name(name)] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "kernel": name.split("(")[0],
                                 "duration_us_under_ncu": dur, "capture": rep.split("/")[-1]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
