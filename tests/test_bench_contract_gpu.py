"""GPU: bench.py prints one JSON line that honours the driver's contract (small batch)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_line_contract():
    j = _run(["--blocks", "96", "--steps", "3", "--warmup", "3", "--cpu-blocks", "16", "--e2e-steps", "1"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in j, k
    assert j["unit"] == "Msamples/s" and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert j["n_gpus"] == 1 and j["steps"] == 3
    # PSD + display render + (iq-correction coefficients, forcing, scan) kernels of the demodulator, every step
    assert j["gpu_launches"] == 3 * 5
    assert "workload" in j["config"] and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert j["cpu_baseline"]["kind"] in ("live", "port") and j["cpu_baseline"]["cores"] >= 1
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 96 * 32768 * 8 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    assert j["parity"]["ok"] is True
    assert e["h2d_ceiling_gbs_per_gpu"] > 1.0 and 0.0 < e["e2e_frac_of_ceiling"] < 1.5
    # the other BASELINE configurations ride along as sub-records, each with its own parity
    c = j["configs"]
    assert set(c) == {"C1", "C3", "C4", "C5"}
    assert c["C1"]["parity"]["psd_max_db_err"] <= 1e-4
    assert all(m["audio_rms_err"] <= 1e-5 for m in c["C3"]["modes"].values()) and set(c["C3"]["modes"]) == {"AM", "USB", "LSB"}
    assert c["C4"]["bitwise_equal"] is True and c["C4"]["parity"]["count_mismatches"] == 0
    assert c["C4"]["parity"]["peak_max_db_err"] <= 1e-4 and c["C4"]["parity"]["rows_max_db_err"] <= 1e-4
    assert c["C5"]["parity"]["psd_max_db_err"] <= 1e-4 and c["C5"]["parity"]["of_which_not_on_a_quantisation_boundary"] == 0
    assert sum(c["C5"]["streams_per_rank"]) == 8
    test_bench_line_contract.config = j["config"]


def test_reference_arm_contract():
    j = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-blocks", "16"])
    assert j["impl"] == "reference" and j["unit"] == "Msamples/s" and j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["kind"] in ("live", "port")
    ours = getattr(test_bench_line_contract, "config", None)
    if ours is not None:
        assert j["config"] == ours          # both arms describe the same workload with the same dict
