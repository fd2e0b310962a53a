"""bench.py's JSON contract, checked on the arm that runs without a GPU (`--impl reference`: the reference's own CPU implementation
through oracle/_ref when it is built, else the C / torch port), plus the static shape of the B200 arm's line builder."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not (REPO / "oracle" / "_ref" / "ref_xfeat").exists(), reason="oracle/_ref is built by __graft_entry__.build() where the reference tree is mounted")
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, check=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("vga_640x480")


def test_b200_arm_declares_every_contract_key():
    src = (REPO / "bench.py").read_text()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"', '"scaling"', '"vs_baseline"',
                '"dtype"', '"data"', '"config"', '"e2e"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"gpu_launches"', '"roofline"', '"bound"',
                '"achieved"', '"peak"', '"frac"', '"traffic"', '"cpu_baseline"', '"cores"', '"kind"', '"sample"', '"clocks"', '"workload"'):
        assert key in src, key


def test_conv_layer_roofline_arithmetic():
    """bench.conv_layer_roofline on the committed per-kernel times reproduces profiles/r01_conv_roofline_per_layer.md (0.27)."""
    sys.path.insert(0, str(REPO))
    import bench
    line = json.loads((REPO / "profiles" / "r01_bench_n1.json").read_text())
    ms = line["roofline"]["kernel_ms_per_step"]                     # one launch of every layer per step
    r = bench.conv_layer_roofline(ms, 480, 640, 32, {"bf16_tflops": 1604.2, "hbm_gbs": 6521.1, "source": "measured"}, split_cost=6.0, split_name="3xTF32")
    assert 1.5 < r["measured_ms"] < 1.9 and 0.44 < r["roof_ms"] < 0.48 and 0.24 < r["frac"] < 0.30


def test_committed_round2_line_carries_the_contract():
    """profiles/r02_bench_n1.json is the line `python bench.py` printed on the B200: every key of the contract, a sustained timed
    region, clean clocks, the roofline / cpu_baseline / e2e objects."""
    line = json.loads((REPO / "profiles" / "r02_bench_n1.json").read_text())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("vga_640x480") and "model" not in line["config"]
    assert line["run"]["timed_region_s"] >= 2.0
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and r["traffic"] > 0
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != line["value"]
    assert line["gpu_launches"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    c = line["clocks"]
    assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"] and not any(x in c["reasons"] for x in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))
