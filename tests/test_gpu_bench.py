"""bench.py at a small size on the GPU: the JSON contract (roofline, e2e with the D2H ceiling, cpu_baseline) and the
in-run parity check -- every stream of the run hashed and compared with the reference -- on the headline workload, on
the short-frame CELT workload and on the true configs[4] mix (CELT FB / SILK WB / hybrid FB / CELT 10 ms)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--frames", "10", *extra],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,streams", [("celt", 256), ("celt10", 128), ("mix", 256), ("silk", 256)])
def test_bench_line_and_in_run_parity(workload, streams):
    d = run_bench("--workload", workload, "--streams", str(streams))
    assert d["unit"] == "audio-s/s" and d["value"] > 0 and d["gpu_launches"] > 0
    p = d["parity"]
    assert p["streams_checked"] == streams and p["mismatches"] == 0 and p["bad_status_frames"] == 0
    assert d["e2e"]["matches_device_leg"] and d["e2e"]["d2h_ceiling_gbs"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["roofline"]["bound"] == "issue" and 0 < d["roofline"]["frac"] < 1
    assert d["roofline"]["hbm"]["bound"] == "hbm" and d["roofline"]["hbm"]["achieved"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0


def test_reference_arm_has_the_same_config():
    a = run_bench("--workload", "celt", "--streams", "64", "--no-e2e", "--no-parity", "--no-cpu-baseline")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "10", "--streams", "64"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    b = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert b["impl"] == "reference" and b["metric"] == a["metric"] and b["unit"] == a["unit"]
    for k in ("workload", "streams_per_gpu", "frames_per_step", "sharding"):
        assert a["config"][k] == b["config"][k]
