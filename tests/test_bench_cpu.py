"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the keys the
driver reads, on a bounded sample, and the workload table covers every BASELINE.json config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "multilingual_v3_b8x32", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"] == "multilingual_v3_b8x32"


def test_workloads_cover_every_baseline_config():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    covered = {w["baseline_config"] for w in bench.WORKLOADS.values()}
    assert covered == set(range(len(base["configs"])))
    assert bench.WORKLOADS[bench.DEFAULT_WORKLOAD]["baseline_config"] == 2     # the config the metric is quoted on
