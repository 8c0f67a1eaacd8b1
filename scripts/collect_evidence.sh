#!/bin/bash
# After `gpurun -- 'bash scripts/gpu_run.sh final'`: copy the judged files from gpurun_out/ into profiles/ and rebuild the summaries.
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/bench_r2_n1.json gpurun_out/bench_r2_reference.json gpurun_out/r2_configs.jsonl gpurun_out/r2_ncu_launches.csv profiles/
cp gpurun_out/r2_ncu_summary.md gpurun_out/r2_ncu_uniform_summary.md gpurun_out/r2_ncu_u8_summary.md gpurun_out/traffic.json profiles/  # written on the box
python - <<'PY'
import json, subprocess
d = json.load(open("profiles/bench_r2_n1.json"))
ks = d["roofline"]["kernel_share"]
tab = subprocess.run(["python", "scripts/ncu_launch_share.py", "profiles/r2_ncu_launches.csv", "1000"], capture_output=True, text=True).stdout
open("profiles/r2_ncu_launch_share.md", "w").write("""# r2: ncu launch list of `bench.py --steps 2 --warmup 3` (gpu__time_duration.sum, --clock-control none)

Only the 8-frame launches of the float32 step (>= 1 ms each: `python scripts/ncu_launch_share.py profiles/r2_ncu_launches.csv 1000`;
the e2e arm launches per frame and row band, see r2_ncu_launches.csv for everything).  Cold-cache and serialised under ncu: compare
SHARES with bench.py's `roofline.kernel_share` (CUDA events, profiles/bench_r2_n1.json: %.3f / %.3f / %.3f for stage 1 / stage 2 /
resampler), not absolutes.

%s""" % (ks["lut_stage1"], ks["lut_stage2"], ks["resize_sr"], tab))
r = d["roofline"]
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "kernel_ms", {k: round(v, 3) for k, v in r["kernel_ms"].items()})
print("share", {k: round(v, 3) for k, v in ks.items()}, "path frac", round(r["path"]["frac"], 4), "GB/s", round(r["path"]["achieved"]))
print("uniform", round(d["value_uniform"]), {k: round(v, 3) for k, v in d["other_input"]["kernel_ms"].items()})
print("u8", {k: (round(v["value"]), round(v["kernel_ms"]["resize_sr"], 3)) for k, v in r["path_u8"].items()})
print("e2e", round(d["e2e"]["value"]), round(d["e2e"]["d2h_GBps_achieved"], 1), round(d["e2e"]["d2h_link_GBps_plain_copy"], 1))
print("parity", d["parity"]["f32_max_abs_err"], "cfg5", round(d["rowband_cfg5"]["value"]), round(d["rowband_cfg5"]["ms_per_frame"], 3))
print("cpu", round(d["cpu_baseline"]["value"], 2), round(d["cpu_baseline"]["port"]["value"], 2))
print(tab)
PY
