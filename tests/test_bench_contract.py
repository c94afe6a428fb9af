"""The bench.py JSON line: the committed round-1 records (profiles/, written by bench.py on the B200 box) carry every key
of the driver's contract, and the reference arm's line has its own required shape.  CPU only."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def load(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["r1_bench_h2_v10.json", "r1_bench_qem_v10.json", "r1_bench_hea10_v8.json", "r1_bench_lih12_v8.json",
                                  "r1_bench_h2_n8_v8.json"])
def test_record_has_contract_keys(name):
    d = load(name)
    assert [k for k in BASE if k not in d] == []
    assert d["metric"] == "noisy_circuit_energy_evals_per_sec" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"] and d["value"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["value"] != d["value"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d.get("cpu_baseline"):            # rank 0 at N = 1, unless the run was started with --no-cpu-baseline
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert name != "r1_bench_qem_v10.json" or (d["dedup"]["unique_rows"] < d["dedup"]["rows"] and d["dedup"]["value"] > d["value"])
    assert name != "r1_bench_h2_v10.json" or d["cpu_baseline"]["by_threads"].keys() >= {"1", "6"}


def test_reference_arm_record():
    d = load("r1_bench_h2_reference_arm_v8.json")
    assert d["impl"] == "reference" and d["metric"] == "noisy_circuit_energy_evals_per_sec"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    h2 = load("r1_bench_h2_v10.json")
    assert d["config"]["workload"] == h2["config"]["workload"]            # same workload on both arms


def test_every_record_parses():
    for path in glob.glob(os.path.join(ROOT, "profiles", "r1_bench_*.json")):
        with open(path) as fh:
            assert "value" in json.load(fh), path


def test_round2_default_record_carries_every_config_and_parity():
    """The round-2 default line: headline contract keys, a parity figure per workload (all within the 1e-10 bar), the `also`
    block with BASELINE configs[2..4] + the batch-axes and 7-qubit riders, each with roofline and end-to-end numbers."""
    d = load("r2_bench_default_v3.json")
    assert [k for k in BASE if k not in d] == []
    assert d["parity_max_abs_err"] is not None and d["parity_max_abs_err"] < 1e-10
    assert d["e2e"]["value"] != d["value"] and d["e2e"]["in_flight"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 65536 * 8
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["by_threads"].keys() >= {"1", "6"}
    also = d["also"]
    assert set(also) >= {"h2_qem", "hea10", "lih12", "h2_bonds", "swap7_qem"}
    for key, rec in also.items():
        assert "error" not in rec, (key, rec)
        assert rec["value"] > 0 and rec["ms_per_step"] > 0 and rec["gpu_launches"] >= rec["steps"]
        assert rec["parity_max_abs_err"] is not None and rec["parity_max_abs_err"] < 1e-10, key
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rec["roofline"])
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(rec["e2e"])
    for key in ("hea10", "lih12"):
        rf = also[key]["roofline"]
        assert rf["bound"] == "hbm" and 0 < rf["frac_executed"] < rf["frac"]       # canonical bytes vs the bytes really moved
    assert also["hea10"]["roofline"]["frac"] >= 0.70 and also["lih12"]["roofline"]["frac"] >= 0.70
    assert also["h2_qem"]["estimator"]["value"] > 0 and also["h2_qem"]["e2e"]["unique_rows"] < also["h2_qem"]["e2e"]["rows_represented"]


def test_round2_streamed_protocol_record():
    """Second half of round 2: microsecond-scale workloads time their K steps in ONE event region (4 streams, distinct buffers
    cycled past the L2); the per-step-event figure stays as `isolated` and must be the smaller one."""
    d = load("r2_bench_default_v6.json")
    assert [k for k in BASE if k not in d] == []
    assert d["parity_max_abs_err"] < 1e-10 and d["gpu_launches"] == d["steps"]
    assert "ONE pair of CUDA events" in d["config"]["timing"] and "L2" in d["config"]["l2"]
    assert 0 < d["isolated"]["value"] < d["value"] and d["isolated"]["steps"] >= 1
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - 65536) < 1.0          # value = batch / step time
    assert d["e2e"]["in_flight"] >= 8 and d["e2e"]["streams"] == 4 and d["e2e"]["value"] > d["e2e"]["synchronous"]["value"]
    for key in ("h2_qem", "h2_bonds", "swap7_qem"):
        rec = d["also"][key]
        assert rec["isolated"]["value"] <= rec["value"] * 1.05 and rec["parity_max_abs_err"] < 1e-10
    for key in ("hea10", "lih12"):
        assert "isolated" not in d["also"][key]                              # streaming steps keep per-step events


def test_round2_thread_per_circuit_record():
    """Final state of round 2: the sweep and the QP variant batch run on the thread-per-circuit kernels; the record names them,
    keeps the contract keys and the parity field, and the roofline is stated on the flops those kernels execute."""
    d = load("r2_bench_default_v7.json")
    assert [k for k in BASE if k not in d] == []
    assert d["parity_max_abs_err"] < 1e-10 and d["gpu_launches"] == d["steps"]
    assert d["roofline"]["kernel"].startswith("dmx_frame16t_kernel") and d["roofline"]["executed_flops_per_circuit"] == 8 * 64 + 32
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - 65536) < 1.0 and d["value"] > 2.0 * 7.2e9       # warp kernels: 7.2e9
    assert 0 < d["isolated"]["value"] < d["value"] and d["e2e"]["value"] > d["e2e"]["synchronous"]["value"]
    qem = d["also"]["h2_qem"]
    assert qem["roofline"]["kernel"].startswith("dmx_frame16v_kernel") and qem["parity_max_abs_err"] < 1e-10
    assert qem["value"] > 2.0 * 2.9e9 and qem["estimator"]["value"] > 0
    assert d["also"]["h2_bonds"]["value"] > 2.0 * 6.5e9 and d["also"]["h2_bonds"]["parity_max_abs_err"] < 1e-10
    for key in ("hea10", "lih12", "swap7_qem"):
        assert d["also"][key]["parity_max_abs_err"] < 1e-10


def test_round2_final_record_with_fused_estimator():
    d, before = load("r2_bench_default_v8.json"), load("r2_bench_default_v7.json")
    assert [k for k in BASE if k not in d] == [] and d["parity_max_abs_err"] < 1e-10
    assert d["roofline"]["kernel"].startswith("dmx_frame16t_kernel") and d["value"] > 2.0e10 and d["e2e"]["value"] > 4.0e9
    est, est0 = d["also"]["h2_qem"]["estimator"], before["also"]["h2_qem"]["estimator"]
    assert est["value"] > 1.3 * est0["value"] and est["variants_per_estimate"] == 125000         # draw + evaluate in one kernel
    assert "dmx_qp_sampled_energies" in d["also"]["h2_qem"]["e2e"]["device_sampling"]["api"]
    for key in ("h2_qem", "hea10", "lih12", "h2_bonds", "swap7_qem"):
        assert d["also"][key]["parity_max_abs_err"] < 1e-10
    ref = load("r2_bench_reference_arm_v8.json")
    assert ref["impl"] == "reference" and "libdmx.so is not loaded" in ref["native"] and ref["value"] < 1e5


def test_round2_final_multi_gpu_records():
    for name, n in (("r2_bench_default_n2_v3.json", 2), ("r2_bench_default_n8_v4.json", 8)):
        d = load(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["value"] > 0.85 * n * 2.3e10
        assert d["roofline"]["kernel"].startswith("dmx_frame16t_kernel") and d["parity_max_abs_err"] < 1e-10
        est = d["also"]["h2_qem"]["estimator"]
        assert "NCCL all-reduce" in est["contains"] and est["variants_per_estimate"] == 125000 * n


def test_round2_multi_gpu_records():
    for name, n in (("r2_bench_default_n2_v1.json", 2), ("r2_bench_default_n8_v2.json", 8)):
        d = load(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["value"] > 0.9 * n * 3.3e9
        est = d["also"]["h2_qem"]["estimator"]
        assert "NCCL all-reduce" in est["contains"] and est["variants_per_estimate"] == 125000 * n


def test_round2_reference_arm_record():
    d = load("r2_bench_reference_arm_v3.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and "libdmx.so is not loaded" in d["native"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

