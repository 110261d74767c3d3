"""Host-side pieces of bench.py that do not need a GPU: the algorithmic-work table, the
peak lookup, the sha-checked ncu traffic record, the hang guard and the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_algorithmic_flops_per_image_match_the_survey_table():
    # 14 F_D + 3 F_G with ch = 128
    want = {8: 0.747e9, 16: 3.32e9, 32: 13.60e9, 64: 41.71e9, 128: 69.83e9}
    for res, w in want.items():
        assert abs(bench.step_flops(res) - w) <= 5e-3 * w, (res, bench.step_flops(res), w)


def test_peaks_come_from_the_driver_file_when_present():
    burst, sustained, hbm, where = bench.peaks()
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        assert (burst, sustained, hbm) == (p["bf16_tflops"], p["bf16_tflops_sustained"], p["hbm_gbs"])
        assert where.startswith("measured")
    else:
        assert where.startswith("fallback")
    assert burst >= sustained > 0 and hbm > 0


def test_ncu_traffic_is_only_quoted_for_the_kernel_source_it_was_taken_from(tmp_path, monkeypatch):
    traffic, note = bench.ncu_traffic()
    with open(os.path.join(ROOT, "profiles", "r2", "conv4_ncu_summary.json")) as f:
        rec = json.load(f)
    src = os.path.join(ROOT, "progressive-gan-pytorch_b200", "csrc", "conv4_tc.cu")
    if rec["conv4_tc_cu_sha16"] == bench.file_sha16(src):
        assert traffic == rec["dram_bytes_per_launch"] and traffic > 0
    else:
        assert traffic is None and "another build" in note
    # a different kernel source must void the record
    monkeypatch.setattr(bench, "file_sha16", lambda path: "0" * 16)
    traffic, note = bench.ncu_traffic()
    assert traffic is None and "another build" in note


def test_hang_guard_ends_the_process_with_a_stack_dump():
    code = "import bench, time; bench.arm_deadline(0.5); time.sleep(30)"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "Timeout" in r.stderr and "most recent call first" in r.stderr


def test_hang_guard_can_be_disabled():
    code = "import bench; assert bench.arm_deadline(0) == 0; print('alive')"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "alive" in r.stdout


def test_reference_arm_prints_the_contract_line_at_a_small_size():
    """`--impl reference` end to end on the host cores (8 px, batch 4 keeps it to seconds): one JSON
    line with the contract's keys; kind "reference" when the unmodified modules are staged."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--res", "8",
                        "--batch", "4", "--steps", "2", "--warmup", "1"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "img/s" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "img/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "progan_modules.py"))
    assert cb["kind"] == ("reference" if staged or os.path.isdir("/root/reference") else "port")


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
