"""CPU: the committed bench line (profiles/r02_bench_n1.json, written by `python bench.py` on a B200) carries every key of the
bench contract and its derived figures are consistent with the raw ones next to them."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = os.path.join(ROOT, "profiles", "r02_bench_n1.json")


@pytest.fixture(scope="module")
def line():
    with open(LINE) as f:
        return json.load(f)


def test_contract_keys(line):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    # BASELINE.json's metric is a sentence ("decoded images/sec at 1/2/4/8 B200 ...; % HBM roofline; vs CPU"): the line names its
    # first clause, the roofline and the CPU figure ride in `roofline` / `cpu_baseline`
    assert json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"].startswith("decoded images/sec")
    assert line["metric"] == "decoded_images_per_sec" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["gpu_launches"] > 0 and line["n_gpus"] == 1 and line["warmup"] >= 3
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in line["clocks"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["pipelined_check"] == "ok"
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in line["cpu_baseline"], k
    assert line["cpu_baseline"]["kind"] in ("reference", "port")


def test_value_follows_from_the_step_time(line):
    images = 64 * line["n_gpus"]
    assert line["value"] == pytest.approx(images / (line["ms_per_step"] / 1e3), rel=1e-6)
    e = line["e2e"]
    assert e["value"] == pytest.approx(images / (e["ms_per_step"] / 1e3), rel=1e-6)
    assert e["serial"]["value"] <= e["value"] * 1.02          # calls in flight never lose to synchronous calls
    assert e["value"] < line["value"]                          # host buffers in and out cannot beat device-resident inputs


def test_roofline_fields_are_consistent(line):
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] / 1e3) / 1e9, rel=1e-9)
    # 32 feature rows of C*4 bytes per (centre, joint) in the sampling phase of config #2 (B=64, K=10, J=15, C=256)
    assert r["algorithmic_bytes_per_launch"] == 64 * 10 * 15 * 32 * 256 * 4
    assert r["gathered_bytes_per_launch"] == r["dedup"]["distinct_rows"] * 256 * 4 <= r["algorithmic_bytes_per_launch"]
    assert r["gathered_frac"] <= r["frac"]
    if r["traffic"]:
        assert r["dram_frac"] == pytest.approx(r["traffic"] / (r["kernel_ms"] / 1e3) / 1e9 / r["peak"], rel=1e-9)
    assert r["kernel_ms"] == pytest.approx(r["stage_ms"][r["stage"]], rel=1e-9) == max(r["stage_ms"].values())
    p = r["path"]
    assert p["ms"] == pytest.approx(sum(r["stage_ms"].values()), rel=1e-6)
    assert p["serial_replay"]["ms"] <= p["ms"] and p["pipelined"]["ms"] == pytest.approx(line["ms_per_step"], rel=1e-9)


def test_extra_workloads_cover_the_other_configs(line):
    x = line["extra_workloads"]
    for k in ("panoptic_spread", "single", "crowded", "mupots"):
        assert k in x and "error" not in x[k], k
        assert x[k]["value"] > 0 and x[k]["roofline"]["stage_ms"]
    # config #3 runs its dense layers; the others have none
    assert x["mupots"]["roofline"]["stage"] == "dense_layers"
    assert x["single"]["serial_replay_ms"] < 0.05
