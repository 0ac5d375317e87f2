"""CPU checks of bench.py's host-side arithmetic (no GPU): the whole-step roofline model of SURVEY 8d."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_step_roofline_matches_survey_numbers():
    b = _bench()
    # ML-20M shape, B = 500: SURVEY 8d quotes "G step ~0.87 GB -> ~132 us HBM vs 27 us tensor -> HBM-bound" and 24.9 / 74.7 MFLOP/user
    r = b.step_roofline(20108, 500, 73.0, 30.0, 18000, 18000, 0.08, 0.16, 0.32, 6551.0, 1641.2)
    assert r["G"]["bound"] == "hbm" and 0.85e9 < r["G"]["bytes"] < 0.89e9
    assert 0.125 < r["G"]["t_hbm_ms"] < 0.140
    assert abs(r["A"]["flops"] / 500 - 24.9e6) < 0.3e6
    assert r["D"]["bound"] == "tensor"
    s = r["step"]
    assert abs(s["t_roof_ms"] - (r["A"]["t_roof_ms"] + r["D"]["t_roof_ms"] + r["G"]["t_roof_ms"])) < 1e-12
    assert abs(s["frac"] - s["t_roof_ms"] / 0.56) < 1e-9
    for ph in "ADG":
        assert 0.0 < r[ph]["frac"] < 1.0


def test_measured_peaks_fallbacks():
    b = _bench()
    v, src = b.measured_peaks("hbm_gbs")
    assert v > 1000 and ("measured" in src or "fallback" in src)
    v, src = b.measured_peaks("bf16_tflops")
    assert v > 100
