"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own modules -- or the oracle port when they are not
vendored -- on the host cores) prints ONE JSON line with the keys the driver reads.  The GPU arm needs a B200 and is exercised by
the driver itself; this test only pins the contract of the line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=580)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "videos/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("penn_mvf") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-9
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and e2e["unit"] == d["unit"]
    assert abs(e2e["value"] - d["value"]) < 1e-9


def test_workloads_are_the_baseline_configs():
    """bench.py's workloads are the per-GPU shards of BASELINE.json configs[1..4]."""
    sys.path.insert(0, ROOT)
    import bench
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    w = bench.WORKLOADS
    assert "32 videos" in cfgs[1] and "20 frames" in cfgs[1] and "2 views" in cfgs[1]
    assert (w["penn_cfg2"]["videos_per_gpu"], w["penn_cfg2"]["T"], w["penn_cfg2"]["head"]["entities"]) == (32, 20, 3)
    assert "global batch 256" in cfgs[2] and 256 // 8 == w["penn_cfg2"]["videos_per_gpu"]      # --global-videos 256 at N = 2 / 4 / 8
    assert "64 videos" in cfgs[3] and "80 frames" in cfgs[3] and "8 B200" in cfgs[3]
    assert (w["finegym_cfg4"]["videos_per_gpu"] * 8, w["finegym_cfg4"]["T"], w["finegym_cfg4"]["head"]["entities"]) == (64, 80, 6)
    assert "32 videos" in cfgs[4] and "240 frames" in cfgs[4] and "16 entity" in cfgs[4]
    assert (w["long_cfg5"]["videos_per_gpu"] * 8, w["long_cfg5"]["T"], w["long_cfg5"]["head"]["entities"]) == (32, 240, 16)
    for name, wl in w.items():          # ViT-B/16 x 3 feature layers: 196 patch tokens of 3 x 768 channels
        assert (wl["P"], wl["c_in"]) == (196, 2304), name
    assert bench.METRIC.startswith("training videos/sec") and bench.UNIT == "videos/s"
