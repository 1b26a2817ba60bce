"""bench.py pieces that need no GPU: the reference arm (the reference's own CPU path, oracle/_ref) prints the contract's
JSON line, and the clock sampler degrades to "unavailable" instead of failing when there is no NVML / nvidia-smi."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(built):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmzref.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-mib", "16"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "deflate_l1_crc32_input_throughput" and line["unit"] == "GiB/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0.3 < line["ratio"] < 0.6  # zlib level 1 on the synthetic text


def test_clock_sampler_without_a_gpu_reports_unavailable():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    sys.path.insert(0, ROOT)
    import bench
    c = bench.ClockSampler(0)
    c.start()
    out = c.stop()
    assert out["sm_mhz"] is None and "reasons" in out
