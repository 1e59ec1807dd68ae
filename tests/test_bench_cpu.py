"""bench.py on the CPU: the pieces that do not need a GPU -- the reference arm end to end (it is the one leg that runs on
host cores only), the clock-sample reduction and the workload description -- so that a typo there cannot cost the round's
benchmark line."""
import importlib.util
import json
import os
import subprocess
import sys

from util import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--l1", "64", "--l2", "4", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "G tri-box tests/s" and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cessna Level1 64 + Level2 4^3")
    # the workload description carries the same keys and numbers as the GPU arm's (derived there from the reference's own structures)
    assert set(line["config"]) == {"workload", "l1", "l2", "mesh", "cache", "tri_box_tests_per_model", "triangles", "grid"}
    if line["cpu_baseline"]["kind"] == "reference":
        assert line["config"]["tri_box_tests_per_model"] == 38439 + 1668608 and line["config"]["triangles"] == 7446 and line["config"]["grid"] == [60, 16, 64]
        assert "whole CPU path" in line["cpu_baseline"]["sample"] and line["ray_tests_per_step"] > 0
        c1 = line["c1"]                                    # BASELINE.json configs[0]: cessna Level-1 64, brute-force fill + ClassifyTessellation
        assert c1["fill_ray_tests"] == 60 * 16 * 64 * 7446 and c1["tri_box_tests"] == 38439 and c1["ms_per_model"] > 0


def test_batch_reference_arm_prints_models_per_second():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "models/s" and line["value"] > 0 and line["scaling"] == "weak"
    assert line["config"]["l1"] == 64 and line["config"]["l2"] == 4 and "configs[4]" in line["config"]["workload"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_samples_are_reduced_like_the_recipe_says():
    b = _bench()
    s = b.ClockSampler(0)
    row = lambda sm, reasons: ["0", str(sm), "1965", "250.5", "0x0"] + reasons
    na = ["Not Active"] * 4
    s.rows = [(0.5, row(1000, na)), (1.1, row(1950, na)), (1.2, row(1965, ["Not Active", "Not Active", "Not Active", "Active"])), (1.3, row(1965, na)),
              (1.35, ["garbage"]), (2.5, row(300, ["Active"] * 4))]
    out = s.stop(1.0, 2.0, 0)
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3 and out["reasons"] == ["sw_power_cap"]
    assert out["power_w_max"] == 250.5 and out["window"] == "timed steps"
    assert b.ClockSampler(0).stop(0, 1, 140)["sm_mhz"] is None                 # no nvidia-smi: the keys are still there


def test_workload_description_and_peaks():
    b = _bench()

    class A:
        mesh, l1, l2 = "cessna", 256, 16
    cfg = b.workload_config(A)
    assert "configs[1]" in cfg["workload"] and cfg["l1"] == 256 and cfg["l2"] == 16 and "L2 flushed" in cfg["cache"]
    peak, src = b.measured_peaks()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
    assert b.FLOPS_PER_TRIBOX == 124


def test_cpu_baseline_reports_all_cores_and_one_thread(tmp_path):
    """the reference's TriBoxOverlap loop nest on every host core and on one thread (its actual execution model)."""
    b = _bench()
    path = b.make_mesh_file("cessna", str(tmp_path))
    out = b.cpu_baseline(path, 64, 4, os.cpu_count() or 1, target_seconds=0.5)
    assert out["kind"] in ("reference", "port") and out["value"] > 0 and out["cores"] == (os.cpu_count() or 1) and "boundary cells" in out["sample"]
    one = out["single_thread"]
    assert one["cores"] == 1 and one["value"] > 0 and one["unit"] == out["unit"]
