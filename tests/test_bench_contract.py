"""CPU: bench.py's reference arm prints exactly one JSON line with the contract's keys (the GPU arm needs a GPU; its line
is checked on the box by scripts/gpu_*.sh and by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports; the arm must still use every allowed core
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "1", "--sample-rows", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "MPix/s" and d["higher_is_better"] is True and d["value"] > 0
    # the unmodified reference (baseline/_ref, staged by __graft_entry__.build() where /root/reference exists) when it is
    # there, with the C/OpenMP port timed beside it; the port alone otherwise
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["cpu_baseline"]["port"]["kind"] == "port" and d["cpu_baseline"]["port"]["value"] > 0
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    # both arms print the same config, so the driver's same_config check holds
    import bench
    assert d["config"] == bench.config_block(8, "natural")
    assert d["e2e"] == {"value": d["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
