"""CPU: bench.py's line contract that does not need a GPU -- every arm (--impl ours / reference / upstream_style) describes the
workload with the SAME `config` object at the same N, and the reference arm runs (on a bounded sample) and prints the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_is_identical_across_arms_and_depends_only_on_the_workload():
    import bench
    argv = sys.argv
    try:
        sys.argv = ["bench.py"]
        a = bench.parse()
    finally:
        sys.argv = argv
    c1 = bench.config_dict(a, 1)
    assert c1 == bench.config_dict(a, 1)
    assert set(c1) == {"workload", "views", "width", "height", "gaussians", "sh_degree", "l2", "views_per_rank", "views_per_launch", "parallelism"}
    assert (c1["views"], c1["width"], c1["height"], c1["gaussians"], c1["sh_degree"]) == (24, 1920, 1080, 60000, 3)      # BASELINE config 2
    c8 = bench.config_dict(a, 8)
    assert c8["views_per_rank"] == 3 and c8["views_per_launch"] == 3 and "x8" in c8["parallelism"]
    assert a.ref_views_per_step == 24                                   # the reference arm times the whole workload per step


def test_reference_arm_prints_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-views-per-step", "1", "--width", "480", "--height", "270", "--gaussians", "4000"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-1000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert set(line["config"]) == {"workload", "views", "width", "height", "gaussians", "sh_degree", "l2", "views_per_rank", "views_per_launch",
                                   "parallelism"}


def test_clock_sampler_picks_the_samples_of_the_timed_region():
    """The sampler runs from process start (idle-time rows included); only rows sampled in or next to the timed region may be reported."""
    import bench
    rows = [(100.00, "idle-a"), (100.05, "idle-b"), (103.00, "warm"), (103.05, "load-1"), (103.10, "load-2"), (105.0, "e2e")]
    assert bench.ClockSampler.select_rows(rows, 103.04, 103.11) == ["load-1", "load-2"]
    assert bench.ClockSampler.select_rows(rows, 103.06, 103.09) == ["warm", "load-1", "load-2"]  # region shorter than the period: +- one period
    assert sorted(bench.ClockSampler.select_rows(rows, 104.0, 104.01)) == ["e2e", "load-1", "load-2"]   # nothing near: the three nearest, never the idle rows
    assert bench.ClockSampler.select_rows([], 1.0, 2.0) == []
