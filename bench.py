#!/usr/bin/env python
"""Headline benchmark: noisy-circuit energy evaluations per second (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload h2_sweep|h2_qem|hea10|lih12] [--impl reference]

Default workload = BASELINE.json configs[1]: H2 STO-3G 4-qubit UCCSD ansatz, depolarize(1e-3) after every 1-qubit
gate and TwoQubitDepolarizingChannel(1e-3) after every CNOT, parameter sweep of 65,536 circuits per launch.  A
"step" is one launch of the hot path over one batch; inputs are resident in HBM for `value`, in host memory for
`e2e` (dmx_energy_batch_host: H2D + kernels + D2H inside the timed region).  Multi-GPU: one process per GPU
(torchrun), the batch is weak-scaled (each rank owns its own 65,536 circuits), no data-path collective
(independent circuits) — time is max over ranks.

`--impl reference` times the reference's CPU path for the same circuit: the repo's C restatement of cirq's
density-matrix algorithm (oracle/dm_oracle.c, kind "port" — cirq/openfermion themselves are not installable
here) on all host cores, on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "noisy_circuit_energy_evals_per_sec"
UNIT = "circuits/s"


# ---------------------------------------------------------------------------------------------------------
# workloads (built through the repo's public API: HydrogenAnsatz + INoiseWrapper + QPU, or the op builder)
# ---------------------------------------------------------------------------------------------------------

PLAN_OPTIONS = {}


def workload_h2_sweep():
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
    from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
    from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w = HydrogenAnsatz()
    noise = INoiseModel([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "depolarize p=1e-3")
    n_w = INoiseWrapper(w, noise)
    prog = QPU.compile_energy_program(n_w, n_w.get_noisy_circuit(), **PLAN_OPTIONS)
    B = 65536

    def make_inputs(rank: int, seed_offset: int = 0):
        # params[b] = -0.5 + b/65535, shifted per rank so ranks do distinct work
        p = -0.5 + (np.arange(B, dtype=np.float64) + 0.25 * rank) / (B - 1)
        return p.reshape(B, 1), None

    return dict(name="H2 STO-3G 4q UCCSD, depolarize(1e-3) on every gate, 65536-circuit parameter sweep (BASELINE configs[1])",
                prog=prog, batch=B, make_inputs=make_inputs, terms=QPU.get_hamiltonian_terms(w))


def workload_h2_qem():
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
    from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
    from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager, TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w = HydrogenAnsatz()
    w.operator_parameters["alpha0"] = 0.014133516193216759
    noise = INoiseModel([cirq.bit_flip(1e-3), cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "bitflip+depolarize p=1e-3")
    n_w = INoiseWrapper(w, noise)
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    mgr = IErrorMitigationManager(clean, noise, QPU.get_hamiltonian_objective_operator(w))
    prog = mgr._variant_program()
    B = 1_000_000 // 8  # per-GPU shard of the 1e6-variant table (BASELINE configs[2] is quoted on 8 GPUs)

    def make_inputs(rank: int, seed_offset: int = 0):
        state = np.random.get_state()
        np.random.seed(20260102 + rank)
        qps = mgr._qp_matrix()
        rows = np.zeros((B, len(qps)), np.uint8)
        for s, qp in enumerate(qps):
            rows[:, s] = np.random.choice(len(qp), size=B, p=np.abs(qp) / np.sum(np.abs(qp)))
        np.random.set_state(state)
        return None, rows

    return dict(name="H2 QP error mitigation: 125000 sampled variants per GPU (1e6 over 8), bit_flip+depolarize(1e-3) (BASELINE configs[2])",
                prog=prog, batch=B, make_inputs=make_inputs, terms=mgr._observable_terms(4))


def _hea_builder(n: int, depth: int, p: float = 1e-3):
    from vqe_errormitigation_b200 import cirq_compat as cirq, engine
    from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
    b = engine.ProgramBuilder(n)
    dep1 = cirq.channel(cirq.depolarize(p))
    dep2 = cirq.channel(TwoQubitDepolarizingChannel(p))
    cnot = cirq.unitary(cirq.CNOT)
    for layer in range(depth):
        for q in range(n):
            b.add_rotation(q, 1, f"t{layer}_{q}_0", 1.0)
            b.add_kraus([q], dep1)
            b.add_rotation(q, 2, f"t{layer}_{q}_1", 1.0)
            b.add_kraus([q], dep1)
        for q in range(n - 1):
            b.add_unitary([q, q + 1], cnot)
            b.add_kraus([q, q + 1], dep2)
    return b


def workload_hea10():
    from vqe_errormitigation_b200 import engine
    n, depth, B = 10, 20, 256
    b = _hea_builder(n, depth)
    rng = np.random.default_rng(20260104)
    n_terms = 100
    xs, zs = np.zeros(n_terms, np.uint32), np.zeros(n_terms, np.uint32)
    for t in range(n_terms):
        wgt = rng.integers(1, 5)
        for q in rng.choice(n, wgt, replace=False):
            code = rng.integers(1, 4)
            bit = 1 << (n - 1 - int(q))
            if code in (1, 2):
                xs[t] |= bit
            if code in (2, 3):
                zs[t] |= bit
    terms = (xs, zs, rng.normal(size=n_terms))
    prog = engine.CircuitProgram(b, hamiltonian=terms, options=dict(PLAN_OPTIONS))

    def make_inputs(rank: int, seed_offset: int = 0):
        r = np.random.default_rng(20260103 + rank)
        return r.uniform(-np.pi, np.pi, size=(B, prog.n_params)), None

    return dict(name=f"synthetic 10-qubit hardware-efficient ansatz depth 20, noise on every gate, batch {B} per step (BASELINE configs[3], sub-batch of 4096)",
                prog=prog, batch=B, make_inputs=make_inputs, terms=terms)


WORKLOADS = {"h2_sweep": workload_h2_sweep, "h2_qem": workload_h2_qem, "hea10": workload_hea10}


# ---------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons = [], set()
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.002)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_port_rate(wl, sample: int, threads: int = 0, rank: int = 0):
    """Time the C restatement of the reference algorithm on `sample` circuits of the workload."""
    from oracle import c_oracle
    prog = wl["prog"]
    params, variants = wl["make_inputs"](rank)
    p = params[:sample] if params is not None else None
    v = variants[:sample] if variants is not None else None
    t0 = time.perf_counter()
    c_oracle.energy_batch(prog.n_qubits, prog.ops, prog.pool, wl["terms"], params=p, variants=v, batch=sample,
                          n_params=prog.n_params, n_slots=prog.n_slots, threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, dt


def run_reference(args, wl):
    """--impl reference: CPU port on all host threads, bounded sample per step."""
    from oracle import c_oracle
    cores = c_oracle.max_threads()
    # calibrate so the whole run stays within ~2 minutes
    rate0, _ = cpu_port_rate(wl, min(64, wl["batch"]), threads=cores)
    budget_s = 90.0
    sample = int(max(16, min(wl["batch"], rate0 * budget_s / max(args.steps + args.warmup, 1))))
    for _ in range(args.warmup):
        cpu_port_rate(wl, sample, threads=cores)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_port_rate(wl, sample, threads=cores)
        t_total += dt
    value = sample * args.steps / t_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "sample_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} circuits of the workload per step, {args.steps} steps (oracle/dm_oracle.c, pthreads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="h2_sweep", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="planner option key=value (dmx_plan_set_option)")
    args = ap.parse_args()
    for kv in args.opt:
        k, v = kv.split("=")
        PLAN_OPTIONS[k] = int(v)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args, WORKLOADS[args.workload]())
        return

    import torch
    import torch.distributed as dist
    from vqe_errormitigation_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU execution path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS[args.workload]()
    prog, B = wl["prog"], wl["batch"]
    info = prog.info

    # ---- CPU baseline first (GPU idle), rank 0 at N=1 only -----------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        cores = c_oracle.max_threads()
        r1, _ = cpu_port_rate(wl, min(32, B), threads=cores)
        sample = int(max(32, min(B, r1 * 12.0)))
        rate, dt = cpu_port_rate(wl, sample, threads=cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {sample} circuits of the workload, {dt:.1f} s, C restatement of cirq's per-op "
                                  f"density-matrix algorithm (oracle/dm_oracle.c) on {cores} threads"}

    params_h, variants_h = wl["make_inputs"](rank)
    params_d = torch.tensor(params_h, device="cuda") if params_h is not None else None
    variants_d = torch.tensor(variants_h, device="cuda") if variants_h is not None else None
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    fp64_peak = engine.measure_fp64_peak()
    hbm_peak, hbm_src = measured_peaks()

    def step():
        prog.energies(params_d, variants_d, batch=B, out=out)

    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = engine.launch_count()
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                      # L2 flush between timed iterations (not inside the events)
            starts[i].record()
            step()
            ends[i].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
    launches = engine.launch_count() - launches0
    if world > 1:
        dist.barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * B * args.steps / (dev_ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H + sync inside the timed region) -----------------
    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            prog.energies_host(params_h, variants_h, batch=B)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = prog.energies_host(params_h, variants_h, batch=B)
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = (params_h.nbytes if params_h is not None else 0) + (variants_h.nbytes if variants_h is not None else 0)
        e2e = {"value": world * B * args.steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(B * 8), "api": "dmx_energy_batch_host (CircuitProgram.energies_host)",
               "checksum": float(np.sum(res))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
    ms_per_step = dev_ms_max / args.steps
    kernel_s = ms_per_step * 1e-3
    path = {0: "dmx_tile_kernel", 1: "dmx_segment_kernel", 2: "dmx_tile_kernel (streaming passes)"}[info["path"]]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(args.workload)
    if info["path"] == 2:
        ach = info["hbm_bytes_per_circuit"] * B / kernel_s / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "kernel": path, "peak_source": hbm_src,
                    "algorithmic_bytes_per_circuit": info["hbm_bytes_per_circuit"]}
    else:
        ach = info["flops_per_circuit"] * B / kernel_s / 1e12
        sm_clock = (clocks.summary()["sm_mhz"] or 1965.0) * 1e6
        smem_peak = 148 * 128 * sm_clock / 1e9
        smem_ach = info["smem_bytes_per_circuit"] * B / kernel_s / 1e9
        roofline = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                    "traffic": traffic, "kernel": path,
                    "peak_source": "FP64 FMA micro-benchmark run live (dmx_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                    "executed_flops_per_circuit": info["flops_per_circuit"],
                    "canonical_flops_per_circuit": info["canonical_flops_per_circuit"],
                    "canonical_equiv_tflops": info["canonical_flops_per_circuit"] * B / kernel_s / 1e12,
                    "smem": {"achieved_gbs": smem_ach, "peak_gbs": smem_peak, "frac": smem_ach / smem_peak,
                             "note": "shared-memory exchange bytes per circuit x rate vs 148 SM x 128 B/clk at the sampled SM clock"},
                    "hbm": {"achieved_gbs": info["hbm_bytes_per_circuit"] * B / kernel_s / 1e9, "peak_gbs": hbm_peak}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "l2": "flushed (192 MiB memset) between timed steps",
                       "timing": "CUDA events per step on the launch stream, summed; max over ranks",
                       "plan": {k: info[k] for k in ("path", "n_blocks", "n_segments", "n_passes", "tile_qubits", "n_terms")},
                       "wall_ms_incl_flush": 1e3 * wall},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
