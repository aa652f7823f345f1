"""bench.py's reference arm runs here (CPU only): one JSON line with the contract's keys, the CPU path timed on a bounded
sample, nothing from the GPU arm touched.  (The GPU arm's line is checked on the B200 box by test_gpu_parity / the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-pairs", "1", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys():
    d = _run()
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "c2_1280x720_orb5000" and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # same config as the GPU arm (the bounded sample is stated separately) and --warmup honoured as given
    assert d["config"]["pairs_per_gpu_per_step"] == 96 and cb["sample_pairs"] == 1 and d["warmup"] == 0 and d["steps"] == 1


def test_reference_arm_sharded_c5_workload_is_selectable():
    d = _run("--workload", "c5_1024_sharded")
    assert d["config"]["workload"] == "c5_1024_sharded" and d["config"]["global_pairs_per_step"] == 1024
    assert d["config"]["width"] == 1920 and d["config"]["n_features"] == 10000


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_surf_workload_is_bounded_and_says_what_it_scaled():
    """BASELINE config 3 on the CPU: the numpy SURF restatement is timed on a keypoint sample and scaled (a full pair would take
    ~25 s); the line says so."""
    d = _run("--workload", "c3_1280x720_surf128")
    cb = d["cpu_baseline"]
    assert d["config"]["workload"] == "c3_1280x720_surf128" and cb["extrapolated"] is True and "SCALED" in cb["sample"]
    assert 0 < d["value"] < 5 and cb["kind"] == "port"
