"""CPU suite, part 4: the bench.py JSON contract of the reference arm (runs the reference's CPU kernel,
no GPU needed) and the static pieces of the product arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "RoIs/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "roi_align" in d["config"]["workload"] and d["steps"] == 1


def test_algorithmic_bytes_match_design():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.ALG_BYTES == 1 * 256 * 200 * 272 * 4 + 1000 * 5 * 4 + 1000 * 256 * 7 * 7 * 4 == 105_901_600
    peak, src = bench.peaks()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
