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

    qp_bytes = int(sum(len(q) for q in mgr._qp_matrix()) * 8)

    def sampled_step(step: int):
        # whole mitigated estimate on the device: QP tables in, (mean, stderr, N) out
        return mgr.get_mu_effective_sampled(B * max(int(os.environ.get("WORLD_SIZE", "1")), 1), seed=20260102 + step)

    return dict(name="H2 QP error mitigation: 125000 sampled variants per GPU (1e6 over 8), bit_flip+depolarize(1e-3) (BASELINE configs[2])",
                prog=prog, batch=B, make_inputs=make_inputs, terms=mgr._observable_terms(4), sampled_step=sampled_step,
                sampled_h2d=qp_bytes)


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
    n, depth, B = 10, 20, 4096
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
    prog = engine.CircuitProgram(b, hamiltonian=terms, options={"tile_qubits": 6, **PLAN_OPTIONS})

    def make_inputs(rank: int, seed_offset: int = 0):
        r = np.random.default_rng(20260103 + rank)
        return r.uniform(-np.pi, np.pi, size=(B, prog.n_params)), None

    # SURVEY.md §8d canonical traffic: one read + one write of complex128 rho per layer = depth * 32 B * 4^n
    return dict(name=f"synthetic 10-qubit hardware-efficient ansatz depth 20, noise on every gate, batch {B} per step (BASELINE configs[3])",
                prog=prog, batch=B, make_inputs=make_inputs, terms=terms, canonical_bytes=depth * 32.0 * 4.0 ** n,
                cpu_sample=16)


def _lih_like(n: int = 12, n_occ: int = 4, p: float = 1e-3, n_terms: int = 631, max_excitations=None):
    """SURVEY.md §8d config 5 (the reference has no LiH data): 12 spin orbitals, HF state rx(pi) on q0..q3, every
    spin-conserving single (16) and double (76) excitation as Pauli-exponential blocks emitted by the rule of
    i_wave_function.py:158-189 (basis change, CNOT ladder, rz(2*sign*theta), un-ladder, inverse basis change), noise
    after every gate; one parameter per excitation.  Hamiltonian: 631 seeded number-conserving JW-shaped terms."""
    from vqe_errormitigation_b200 import cirq_compat as cirq, engine
    from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
    rng = np.random.default_rng(20260105)
    b = engine.ProgramBuilder(n)
    dep1 = cirq.channel(cirq.depolarize(p))
    dep2 = cirq.channel(TwoQubitDepolarizingChannel(p))
    cnot = cirq.unitary(cirq.CNOT)
    had = cirq.unitary(cirq.H)
    rxm, rxp = cirq.unitary(cirq.rx(-np.pi / 2)), cirq.unitary(cirq.rx(np.pi / 2))

    def gate(qs, m):
        b.add_unitary(list(qs), m)
        b.add_kraus(list(qs), dep1 if len(qs) == 1 else dep2)

    for q in range(n_occ):
        gate((q,), cirq.unitary(cirq.rx(np.pi)))
    occ, virt = list(range(n_occ)), list(range(n_occ, n))
    singles = [(i, a) for i in occ for a in virt if (i - a) % 2 == 0]
    doubles = [(i, j, a, c) for x, i in enumerate(occ) for j in occ[x + 1:] for y, a in enumerate(virt) for c in virt[y + 1:]
               if (i % 2) + (j % 2) == (a % 2) + (c % 2)]
    excitations = [("s", e) for e in singles] + [("d", e) for e in doubles]
    order = rng.permutation(len(excitations))
    if max_excitations is not None:
        order = order[:max_excitations]
    n_blocks = 0

    def exponential(string, sign, name):
        nonlocal n_blocks
        qs = [q for q, _ in string]
        for q, pl in string:
            if pl == "X":
                gate((q,), had)
            elif pl == "Y":
                gate((q,), rxm)
        for a, c in zip(qs[:-1], qs[1:]):
            gate((a, c), cnot)
        b.add_rotation(qs[-1], 2, name, 2.0 * sign)
        b.add_kraus([qs[-1]], dep1)
        for a, c in reversed(list(zip(qs[:-1], qs[1:]))):
            gate((a, c), cnot)
        for q, pl in string:
            if pl == "X":
                gate((q,), had)
            elif pl == "Y":
                gate((q,), rxp)
        n_blocks += 1

    for k in order:
        kind, e = excitations[k]
        name = f"t{k}"
        if kind == "s":
            i, a = e
            zs = [(q, "Z") for q in range(i + 1, a)]
            exponential([(i, "X")] + zs + [(a, "Y")], +1.0, name)
            exponential([(i, "Y")] + zs + [(a, "X")], -1.0, name)
        else:
            i, j, a, c = e
            z1 = [(q, "Z") for q in range(i + 1, j)]
            z2 = [(q, "Z") for q in range(a + 1, c)]
            for pat, sg in (("XXXY", 1), ("XXYX", 1), ("XYXX", -1), ("YXXX", -1), ("YYYX", -1), ("YYXY", -1), ("YXYY", 1), ("XYYY", 1)):
                exponential([(i, pat[0])] + z1 + [(j, pat[1]), (a, pat[2])] + z2 + [(c, pat[3])], float(sg), name)
    # Hamiltonian terms (x mask / z mask, qubit 0 = most significant bit)
    bit = lambda q: 1 << (n - 1 - q)
    terms = {(0, 0)}
    for q in range(n):
        terms.add((0, bit(q)))
    for q in range(n):
        for r in range(q + 1, n):
            terms.add((0, bit(q) | bit(r)))
    hop = [(q, r) for q in range(n) for r in range(q + 2, n, 2)]
    for q, r in hop:
        zz = sum(bit(t) for t in range(q + 1, r))
        terms.add((bit(q) | bit(r), zz))                       # X Z..Z X
        terms.add((bit(q) | bit(r), zz | bit(q) | bit(r)))     # Y Z..Z Y
    quads = [(q0, q1, q2, q3) for q0 in range(n) for q1 in range(q0 + 1, n) for q2 in range(q1 + 1, n) for q3 in range(q2 + 1, n)]
    for qi in rng.permutation(len(quads)):
        if len(terms) >= n_terms:
            break
        q0, q1, q2, q3 = quads[qi]
        zz = sum(bit(t) for t in range(q0 + 1, q1)) | sum(bit(t) for t in range(q2 + 1, q3))
        xm = bit(q0) | bit(q1) | bit(q2) | bit(q3)
        for ys in ((), (q0, q1), (q0, q2), (q0, q3), (q1, q2), (q1, q3), (q2, q3), (q0, q1, q2, q3)):
            if len(terms) < n_terms:
                terms.add((xm, zz | sum(bit(t) for t in ys)))
    tl = sorted(terms)
    ham = (np.array([t[0] for t in tl], np.uint32), np.array([t[1] for t in tl], np.uint32), rng.normal(scale=0.1, size=len(tl)))
    return b, ham, n_blocks


def workload_lih12():
    from vqe_errormitigation_b200 import engine
    n, B = 12, 8
    b, terms, n_blocks = _lih_like(n)
    prog = engine.CircuitProgram(b, hamiltonian=terms, options={"tile_qubits": 6, **PLAN_OPTIONS})

    def make_inputs(rank: int, seed_offset: int = 0):
        r = np.random.default_rng(20260106 + rank)
        return r.normal(scale=0.1, size=(B, prog.n_params)), None

    # SURVEY.md §8d canonical traffic: one pass (read + write of complex128 rho) per Pauli-exponential block
    return dict(name=f"synthetic LiH-like 12-qubit UCCSD ansatz ({n_blocks} Pauli-exponential blocks, noise on every gate, "
                     f"{len(terms[2])} Pauli terms), {B} optimiser chains per GPU per step (BASELINE configs[4]; no LiH data in the reference)",
                prog=prog, batch=B, make_inputs=make_inputs, terms=terms, canonical_bytes=n_blocks * 32.0 * 4.0 ** n,
                cpu_sample=0, cpu_ops_fraction=0.002)


WORKLOADS = {"h2_sweep": workload_h2_sweep, "h2_qem": workload_h2_qem, "hea10": workload_hea10, "lih12": workload_lih12}


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
    """Time the C restatement of the reference algorithm on `sample` circuits of the workload.  Workloads whose single
    circuit takes the CPU many minutes (12 qubits: rho is 256 MiB, 25k ops) set `cpu_ops_fraction`: only that leading
    fraction of the op list is evolved (one identity term as the observable) and the rate is scaled to the whole list."""
    from oracle import c_oracle
    prog = wl["prog"]
    params, variants = wl["make_inputs"](rank)
    ops, terms, scale = prog.ops, wl["terms"], 1.0
    frac = wl.get("cpu_ops_fraction")
    if frac:
        m = max(8, int(len(ops) * frac))
        ops, scale = ops[:m], m / len(ops)
        terms = (np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.ones(1))
        reps = -(-sample // len(params))
        params = np.tile(params, (reps, 1))
    p = params[:sample] if params is not None else None
    v = variants[:sample] if variants is not None else None
    t0 = time.perf_counter()
    c_oracle.energy_batch(prog.n_qubits, ops, prog.pool, terms, params=p, variants=v, batch=sample,
                          n_params=prog.n_params, n_slots=prog.n_slots, threads=threads)
    dt = time.perf_counter() - t0
    return sample * scale / dt, dt


def cpu_sample_note(wl, sample):
    if wl.get("cpu_ops_fraction"):
        return (f"{sample} circuits x the first {wl['cpu_ops_fraction']:.2%} of the op list, rate scaled to the whole list "
                f"(a full 12-qubit circuit takes the CPU port tens of minutes)")
    return f"{sample} circuits of the workload"


def run_reference(args, wl):
    """--impl reference: CPU port on all host threads, bounded sample per step."""
    from oracle import c_oracle
    cores = c_oracle.max_threads()
    # calibrate so the whole run stays within ~2 minutes
    probe = min(64, wl["batch"]) if not wl.get("cpu_ops_fraction") else cores
    rate0, dt0 = cpu_port_rate(wl, probe, threads=cores)
    budget_s = 90.0
    if wl.get("cpu_ops_fraction"):
        sample = cores
    else:
        sample = int(max(min(16, wl["batch"]), min(wl["batch"], probe / dt0 * budget_s / max(args.steps + args.warmup, 1))))
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
                             "sample": f"{cpu_sample_note(wl, sample)} per step, {args.steps} steps (oracle/dm_oracle.c, pthreads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------------------

_REAL_STDOUT = None


def _capture_stdout():
    """Route fd 1 to stderr for the whole run (library banners such as NCCL's version line print to stdout) and keep the
    original stdout for the single JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _capture_stdout()
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
        if wl.get("cpu_ops_fraction"):
            sample = cores
        else:
            probe = min(32, B, max(cores, 1))
            _, dt1 = cpu_port_rate(wl, probe, threads=cores)
            sample = int(max(probe, min(B, probe / dt1 * 12.0)))
        rate, dt = cpu_port_rate(wl, sample, threads=cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{cpu_sample_note(wl, sample)}, {dt:.1f} s, C restatement of cirq's per-op "
                                  f"density-matrix algorithm (oracle/dm_oracle.c) on {cores} threads"}
        if not wl.get("cpu_ops_fraction"):
            # the reference's own shapes: one process (CPU.get_custom_optimized_state) and its Pool(6)
            # (MAX_MULTI_PROCESSING_COUNT, src/error_mitigation.py:14), ~2 s of work each
            by = {str(cores): rate}
            for th in (1, 6):
                if th < cores:
                    n_s = int(max(8, min(B, rate / cores * th * 2.0)))
                    by[str(th)] = cpu_port_rate(wl, n_s, threads=th)[0]
            cpu_baseline["by_threads"] = by

    params_h, variants_h = wl["make_inputs"](rank)
    params_d = torch.tensor(params_h, device="cuda") if params_h is not None else None
    variants_d = torch.tensor(variants_h, device="cuda") if variants_h is not None else None
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    fp64_peak = engine.measure_fp64_peak()
    hbm_peak, hbm_src = measured_peaks()

    def step():
        prog.energies(params_d, variants_d, batch=B, out=out)

    short_steps = prog.info["path"] != 2

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
            if short_steps:
                # a ~150 us device-side spin ahead of the flush lets the host run ahead, so that the start event, the
                # launch and the end event of a microsecond-scale step are queued back to back (otherwise host jitter -
                # eight ranks and their clock samplers share the CPU - leaks into the event interval)
                torch.cuda._sleep(300_000)
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
        # inputs and result live in page-locked host memory (the contract's "pinned host memory")
        def pinned(a):
            if a is None:
                return None
            b = engine.pinned_empty(a.shape, a.dtype)
            b[...] = a
            return b
        params_p, variants_p, res = pinned(params_h), pinned(variants_h), engine.pinned_empty(B)
        for _ in range(3):
            prog.energies_host(params_p, variants_p, batch=B, out=res)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prog.energies_host(params_p, variants_p, batch=B, out=res)
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = (params_h.nbytes if params_h is not None else 0) + (variants_h.nbytes if variants_h is not None else 0)
        e2e = {"value": world * B * args.steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(B * 8), "api": "dmx_energy_batch_host (CircuitProgram.energies_host)",
               "checksum": float(np.sum(res)),
               "transfer": ("zero-copy: the kernel reads the parameters from and writes the energies to the mapped page-locked host "
                            "buffers over PCIe (one launch + one synchronise per step)"
                            if prog.info["path"] == 1 and variants_h is None and not any(o.startswith("host_zero_copy=0") for o in args.opt)
                            else "cudaMemcpyAsync H2D / D2H from and to the page-locked host buffers, pipelined in chunks over three streams")}

    # ---- QEM only: the same estimate with the variants drawn on the device (no variant table crosses PCIe) ----------
    if e2e is not None and wl.get("sampled_step"):
        for i in range(3):
            wl["sampled_step"](1000 + i)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            est = wl["sampled_step"](i)
        ts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e["device_sampling"] = {"value": world * B * args.steps / float(ts.item()), "unit": UNIT,
                                  "h2d_bytes_per_step": int(wl["sampled_h2d"]), "d2h_bytes_per_step": 24,
                                  "api": "IErrorMitigationManager.get_mu_effective_sampled (dmx_qp_sample_variants + "
                                         "dmx_energy_batch + dmx_qp_reduce [+ 24-byte all-reduce])",
                                  "estimate": [float(est[0]), float(est[1]), int(est[2])]}

    # ---- QEM only: deduplicated rate - the reference evaluates each distinct identifier once and weights it
    # by its count (error_mitigation.py:492-497); here: unique rows of this rank's table, found on the host outside the timing
    dedup = None
    if variants_h is not None and world == 1 and not args.no_e2e:
        uniq = np.unique(variants_h, axis=0)
        uniq_d = torch.tensor(uniq, device="cuda")
        uout = torch.empty(len(uniq), dtype=torch.float64, device="cuda")
        for _ in range(3):
            prog.energies(None, uniq_d, batch=len(uniq), out=uout)
        torch.cuda.synchronize()
        us, ue = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_ms = 0.0
        for _ in range(args.steps):
            torch.cuda._sleep(300_000)
            flush.zero_()
            us.record()
            prog.energies(None, uniq_d, batch=len(uniq), out=uout)
            ue.record()
            torch.cuda.synchronize()
            t_ms += us.elapsed_time(ue)
        dedup = {"unique_rows": int(len(uniq)), "rows": int(B), "value": B * args.steps / (t_ms * 1e-3), "unit": UNIT,
                 "note": "variants represented per second when only the distinct rows are evaluated (device time of the unique-row "
                         "batch; np.unique on the host not included); the roofline fraction is defined on the raw rate"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
    ms_per_step = dev_ms_max / args.steps
    kernel_s = ms_per_step * 1e-3
    path = {0: "dmx_tile_kernel", 1: "dmx_frame32_kernel (dmx_frame_kernel when more than 32 coefficients are live)",
            2: "dmx_stream3_kernel (streaming passes)"}[info["path"]]
    traffic = ncu_util = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            recorded = json.load(fh)
        traffic = recorded.get(args.workload)
        ncu_util = recorded.get("_ncu", {}).get(args.workload)
    if info["path"] == 2:
        # algorithmic bytes = SURVEY.md §8d canonical traffic (complex128 rho, one read + write per layer / block);
        # the engine's own traffic (real Pauli vector, fused passes) is reported beside it
        canon = wl.get("canonical_bytes") or info["canonical_hbm_bytes_per_circuit"]
        ach = canon * B / kernel_s / 1e9
        sm_clock = (clocks.summary()["sm_mhz"] or 1965.0) * 1e6
        smem_peak = 148 * 128 * sm_clock / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "kernel": path, "peak_source": hbm_src,
                    "algorithmic_bytes_per_circuit": canon,
                    "algorithmic_model": "SURVEY.md 8d: 32 B x 4^n per layer (HEA) / per Pauli-exponential block (UCCSD)",
                    "executed": {"hbm_bytes_per_circuit": info["hbm_bytes_per_circuit"],
                                 "hbm_gbs": info["hbm_bytes_per_circuit"] * B / kernel_s / 1e9,
                                 "smem_bytes_per_circuit": info["smem_bytes_per_circuit"],
                                 "smem_gbs": info["smem_bytes_per_circuit"] * B / kernel_s / 1e9,
                                 "smem_frac": info["smem_bytes_per_circuit"] * B / kernel_s / 1e9 / smem_peak,
                                 "fp64_tflops": info["flops_per_circuit"] * B / kernel_s / 1e12,
                                 "fp64_frac": info["flops_per_circuit"] * B / kernel_s / 1e12 / fp64_peak,
                                 "note": "the binding resource is the shared-memory exchange between sweeps "
                                         "(148 SM x 128 B/clk at the sampled clock), then HBM"}}
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
    if ncu_util:
        roofline["ncu"] = ncu_util      # recorded capture (profiles/), not re-measured by this run
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "l2": "flushed (192 MiB memset) between timed steps",
                       "timing": "CUDA events per step on the launch stream, summed; max over ranks" + ("; a device-side spin precedes each flush so the host queues ahead" if short_steps else ""),
                       "plan": {k: info[k] for k in ("path", "n_blocks", "n_segments", "n_passes", "tile_qubits", "n_terms")},
                       "wall_ms_incl_flush": 1e3 * wall},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            **({"dedup": dedup} if dedup else {}),
            "cpu_baseline": cpu_baseline}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
