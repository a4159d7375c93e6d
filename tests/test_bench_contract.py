"""bench.py's contract on the CPU: the reference arm prints ONE JSON line with the keys the driver reads, on the GPU arm's own
config string; the helpers behind `roofline` are consistent with DESIGN.md section 4."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "shape", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "200"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "nuclei/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "200 nuclei" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "nuclei/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("shape: geometry feature set(s), 10000 nuclei per GPU")


def test_roofline_helpers():
    import bench
    assert bench.kernel_bytes("k_hue_batch", 64, 4) == 12288 + 512 + 16 + 32
    assert bench.kernel_bytes("k_color_warp", 64, 4) == 12884 and bench.kernel_bytes("k_glcm", 64, 4) == 13712
    assert bench.algorithmic_bytes("color", 64, 18) == 3 * 64 * 64 + 8 * 31 + 8 + 72    # 12 616
    h = bench.source_hash()
    assert len(h) == 16 and h == bench.source_hash()
    t = bench.measured_traffic("color", "k_hue_batch", 100000, 64)
    tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    hc = bench.source_hash(bench.TRAFFIC_SOURCES["color"])
    assert (t is None) == (tj["color"]["src_sha16"] != hc)  # stale captures read as null, never as a number
    a = bench.glcm_atomics(100000, 7.3)
    assert 0.0 < a["frac_of_random_address_peak"] < 1.0
    assert 0.0 < a["smem_pipe"]["frac"] < 1.0 and abs(a["smem_pipe"]["peak_per_s"] - 148 * 1.965e9) < 1e6
