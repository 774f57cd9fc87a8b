"""The bench line contract, checked on the committed line of the round's final state (profiles/r01n_bench_n1.json, the
stdout of `python bench.py` on a B200): keys and types the driver and the judge read.  CPU only; a regression in
bench.py's output shape shows up the next time the line is regenerated and committed."""
import json
import os

LINE = json.loads(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                    "r01n_bench_n1.json")).read().strip().splitlines()[-1])


def test_base_contract_keys():
    for key, typ in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                     ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str),
                     ("config", dict), ("gpu_launches", int), ("clocks", dict), ("e2e", dict), ("roofline", dict),
                     ("cpu_baseline", dict)):
        assert isinstance(LINE[key], typ), key
    assert LINE["vs_baseline"] is None                              # BASELINE.json publishes no number
    assert LINE["scaling"] == "weak" and LINE["higher_is_better"] and LINE["data"] == "synthetic"
    assert LINE["warmup"] >= 3 and LINE["n_gpus"] == 1 and "workload" in LINE["config"] and "model" not in LINE["config"]
    assert LINE["gpu_launches"] > 0 and LINE["value"] > 0
    assert abs(LINE["ms_per_step"] * LINE["value"] / 1e3 - LINE["config"]["batch"]) / LINE["config"]["batch"] < 0.02


def test_roofline_cpu_baseline_e2e_and_clocks():
    r = LINE["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    c = LINE["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == LINE["unit"] and c["sample"]
    e = LINE["e2e"]
    assert e["unit"] == LINE["unit"] and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != LINE["value"]                              # a separate measurement, not the device number repeated
    k = LINE["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] * 0.9 and isinstance(k["reasons"], list)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(k["reasons"])


def test_extras_carry_their_cpu_legs():
    x = LINE["extras"]
    for key in ("neumf_train", "neumf_train_reference_batch", "twotower_train", "topk_ml1m", "svd_fit"):
        assert x[key]["value"] > 0 and x[key]["cpu_baseline"]["value"] > 0, key
        assert x[key]["value"] > x[key]["cpu_baseline"]["value"], key
    assert x["twotower_train_tc_graph"]["value"] > x["twotower_train_tc"]["value"]
