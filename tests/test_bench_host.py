"""Host-side pieces of bench.py that run without a GPU: the reference arm's JSON line (the CPU port of the
reference path) and the clock sampler's bookkeeping."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scenes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_reports_only_samples_of_the_timed_region():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)

    class _Proc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    now = time.perf_counter()
    s.proc, s.t0 = _Proc(), now - 0.05
    s.lines = [(now - 1.0, "1200, 1965, 200.0, Not Active, Not Active, Not Active, Not Active\n"),      # before the region
               (now - 0.04, "1965, 1965, 310.5, Not Active, Not Active, Not Active, Active\n"),
               (now - 0.02, "1950, 1965, 305.0, Not Active, Not Active, Not Active, Not Active\n"),
               (now - 0.01, "garbage line\n")]
    r = s.stop()
    assert r["samples_in_timed_region"] == 3 and r["samples"] == 2          # the garbage line does not parse
    assert r["sm_mhz"] == 1957.5 and r["sm_max_mhz"] == 1965.0 and r["power_w_max"] == 310.5
    assert r["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_nvidia_smi():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.proc, s.lines, s.t0 = None, [], None
    assert s.stop()["sm_mhz"] is None
