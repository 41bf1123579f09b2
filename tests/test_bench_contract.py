"""The bench.py output contract, checked on the committed record of the last GPU run (the
bench itself needs a B200) and on the reference arm, which runs on the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_common(d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "audio_seconds_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k


def test_committed_bench_record_follows_the_contract():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r1_bench_full_final.json")).read().strip().splitlines()[-1])
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                        "sw_thermal_slowdown"}
    assert d["ber_percent"]["clean"] == 0.0
    assert d["cpu_baseline"]["kind"] in ("port", "reference")


def test_reference_arm_runs_on_the_cpu_and_follows_the_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--iters", "2", "--seconds", "1", "--no-attacks"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(d)
    assert d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


def test_last_round2_record_carries_parity_phases_and_rooflines():
    """The last full `python bench.py` record of round 2 (profiles/r2_bench_final_v6.json): the base contract,
    the oracle parity block at the bench configuration for both loop precisions, SURVEY 8(d)'s phases (i)-(iv),
    BASELINE configs[2] sweeps, and a roofline whose fraction follows from its own achieved / peak."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_final_v6.json")).read().strip().splitlines()[-1])
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "profiling hooks off" in d["timed_region"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    assert all(abs(x["frac"] - x["achieved"] / x["peak"]) < 1e-9 for x in d["rooflines_all"])
    p = d["parity"]
    assert p["all_ok"] is True and p["clips"] >= 4
    for loop in ("fp16_loop", "tf32_loop"):
        assert p[loop]["gpu_bits_equal_oracle_bits_on_gpu_audio"] is True and p[loop]["flips_total"] == 0
    assert p["cross_detect"]["ok"] is True
    ph = d["phases"]
    for k in ("detect_only", "attack_suite_to_detect", "embed", "full_step", "config3_attack_sweep_1024_clips",
              "config3_extensions_1024_clips", "attack_kernels"):
        assert k in ph, k
    assert "error" not in ph["config3_extensions_1024_clips"]
    assert abs(ph["full_step"]["value"] - d["value"]) < 1e-6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["alt_precision"]["embed_precision"] == "tf32" and d["nonfinite_gradient_clips"] == 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
