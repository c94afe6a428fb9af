#!/usr/bin/env python
"""Headline benchmark: noisy-circuit energy evaluations per second (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload h2_sweep|h2_qem|hea10|lih12]
                    [--impl reference] [--no-also]

Default workload = BASELINE.json configs[1]: H2 STO-3G 4-qubit UCCSD ansatz, depolarize(1e-3) after every 1-qubit
gate and TwoQubitDepolarizingChannel(1e-3) after every CNOT, parameter sweep of 65,536 circuits per launch.  A
"step" is one launch of the hot path over one batch; inputs are resident in HBM for `value`, in host memory for
`e2e` (dmx_energy_batch_host: H2D + kernels + D2H inside the timed region).  Microsecond-scale steps are timed as ONE
event region around the K steps (queued back to back over four streams, inputs cycled past the L2); the per-step-event
figure of round 1 is kept as `isolated` (see run_workload and tools/launch_floor.py for why).  Multi-GPU: one process per GPU
(torchrun), the batch is weak-scaled (each rank owns its own 65,536 circuits), no data-path collective
(independent circuits) — time is max over ranks.

The default run also carries an `also` block: short runs of the other BASELINE configs (h2_qem = configs[2], hea10 =
configs[3], lih12 = configs[4]) with their own value / ms_per_step / roofline / e2e, so every configuration is in the
driver-run record.  Under `--gpus N` the h2_qem entry times the whole mitigated estimator INCLUDING its all-reduce.
Every workload is checked against the oracle outside the timed region (`parity_max_abs_err`).

`--impl reference` times the reference's CPU path for the same circuit: the repo's C restatement of cirq's
density-matrix algorithm (oracle/dm_oracle.c, kind "port" — cirq/openfermion themselves are not installable
here) on all host cores, on a bounded sample per step.  That arm builds its op table with the host-side builder only:
libdmx.so is never loaded there.
"""
from __future__ import annotations

import argparse
import hashlib
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
# workloads.  Each returns the HOST-side description only (op-table builder, Hamiltonian terms, inputs); the plan
# (libdmx) is created lazily by get_prog(), so `--impl reference` never maps the CUDA library.
# ---------------------------------------------------------------------------------------------------------

PLAN_OPTIONS = {}


def get_prog(wl):
    if wl.get("_prog") is None:
        from vqe_errormitigation_b200 import engine
        wl["_prog"] = engine.CircuitProgram(wl["builder"], qubits=wl.get("qubits"), measurements=wl.get("measurements"),
                                            hamiltonian=wl["terms"], options={**wl.get("options", {}), **PLAN_OPTIONS},
                                            coefficient_groups=wl.get("coefficient_groups"))
    return wl["_prog"]


def _h2_noisy(noise_1q, noise_2q, description, alpha=None):
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
    from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
    w = HydrogenAnsatz()
    if alpha is not None:
        w.operator_parameters["alpha0"] = alpha
    noise = INoiseModel(noise_1q, noise_2q, description)
    return w, noise, INoiseWrapper(w, noise)


def workload_h2_sweep():
    from vqe_errormitigation_b200 import cirq_compat as cirq, engine
    from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w, _noise, n_w = _h2_noisy([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "depolarize p=1e-3")
    builder, qubits, measurements, _lifted, _key = engine.lower_circuit(n_w.get_noisy_circuit(), w.qubits)
    B = 65536

    def make_inputs(rank: int, seed_offset: int = 0):
        # params[b] = -0.5 + b/65535, shifted per rank so ranks do distinct work
        p = -0.5 + (np.arange(B, dtype=np.float64) + 0.25 * rank) / (B - 1)
        return p.reshape(B, 1), None

    return dict(key="h2_sweep",
                name="H2 STO-3G 4q UCCSD, depolarize(1e-3) on every gate, 65536-circuit parameter sweep (BASELINE configs[1])",
                builder=builder, qubits=qubits, measurements=measurements, batch=B, make_inputs=make_inputs,
                terms=QPU.get_hamiltonian_terms(w), parity_rows=[0, B // 2, B - 1])


BOND_LENGTHS = ["0.1", "0.2", "0.3", "0.4", "0.5", "0.6", "0.7", "0.7414", "0.8", "0.9", "1.0", "1.1", "1.2", "1.3", "1.4", "1.5", "1.6",
                "1.7", "1.9", "2.0", "2.1", "2.2", "2.3", "2.4", "2.5", "2.6", "2.7", "2.8", "2.9", "3.0"]


def workload_h2_bonds():
    """Batch axes A x B of CPU.get_collection_ground_states (reference processor_classic.py:18-44): the 30 stored H2
    geometries x optimiser chains in ONE launch - one plan, one Hamiltonian coefficient row per bond length
    (dmx_plan_set_hamiltonian_groups), circuit b evaluated with row b % 30."""
    from vqe_errormitigation_b200 import cirq_compat as cirq, engine
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_parameter import IParameter
    from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w, _noise, n_w = _h2_noisy([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "depolarize p=1e-3")
    builder, qubits, measurements, _lifted, _key = engine.lower_circuit(n_w.get_noisy_circuit(), w.qubits)
    group_terms = []
    for r in BOND_LENGTHS:
        w.update_molecule(IParameter({"r0": float(r)}))
        group_terms.append(QPU.get_hamiltonian_terms(w))
    assert all(np.array_equal(t[0], group_terms[0][0]) and np.array_equal(t[1], group_terms[0][1]) for t in group_terms)
    B = 65536

    def make_inputs(rank: int, seed_offset: int = 0):
        p = -0.5 + (np.arange(B, dtype=np.float64) + 0.25 * rank) / (B - 1)
        return p.reshape(B, 1), None

    return dict(key="h2_bonds",
                name="H2 4q UCCSD, depolarize(1e-3): 30 bond lengths x 2184 optimiser chains per launch (65536 circuits, one plan, "
                     "one Hamiltonian coefficient row per bond length; batch axes of CPU.get_collection_ground_states)",
                builder=builder, qubits=qubits, measurements=measurements, batch=B, make_inputs=make_inputs,
                terms=group_terms[0], group_terms=group_terms, coefficient_groups=np.stack([t[2] for t in group_terms]),
                make_groups=lambda rank: (np.arange(B, dtype=np.int32) + rank) % len(BOND_LENGTHS), parity_rows=[0, 1, 29, B // 2, B - 1])


def workload_h2_qem():
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager, TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w, noise, n_w = _h2_noisy([cirq.bit_flip(1e-3), cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)],
                              "bitflip+depolarize p=1e-3", alpha=0.014133516193216759)
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    mgr = IErrorMitigationManager(clean, noise, QPU.get_hamiltonian_objective_operator(w))
    builder, qubits, measurements = mgr.variant_builder()
    B = 1_000_000 // 8  # per-GPU shard of the 1e6-variant table (BASELINE configs[2] is quoted on 8 GPUs)
    qps = mgr.quasi_probabilities()

    def make_inputs(rank: int, seed_offset: int = 0):
        state = np.random.get_state()
        np.random.seed(20260102 + rank)
        rows = np.zeros((B, len(qps)), np.uint8)
        for s, qp in enumerate(qps):
            rows[:, s] = np.random.choice(len(qp), size=B, p=np.abs(qp) / np.sum(np.abs(qp)))
        np.random.set_state(state)
        return None, rows

    return dict(key="h2_qem",
                name="H2 QP error mitigation: 125000 sampled variants per GPU (1e6 over 8), bit_flip+depolarize(1e-3) (BASELINE configs[2])",
                builder=builder, qubits=qubits, measurements=measurements, batch=B, make_inputs=make_inputs,
                terms=mgr.observable_terms(), manager=mgr, qp_bytes=int(sum(len(q) for q in qps) * 8),
                parity_rows=[0, 1, 2, 3, B - 1])


def workload_swap7_qem():
    """The reference's only > 4-qubit workload: QP error mitigation of the 7-qubit SWAP test (QP_QEM_lib.py:388-419 swap_circuit
    with cirq 0.8.2's Fredkin decomposition; stored run src/classic_minimisation/SWAP_Similarity_P0.0001_R100_C10000, driver
    visual_testing.py:492-524): asymmetric Pauli noise p = 1e-4, 10 000 sampled identifiers per estimate, observable Z on the
    last qubit (error_mitigation.py:381).  Whole state (4^7 Pauli coefficients = 128 KiB) resident in one CTA's shared memory."""
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200 import experiments as E
    from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager
    from vqe_errormitigation_b200.qp_qem_lib import swap_circuit
    noise = E.asymmetric_pauli_noise_model(1e-4)
    circuit = swap_circuit(3, cirq.LineQubit.range(7), True)
    mgr = IErrorMitigationManager(clean_circuit=circuit, noise_model=noise, hamiltonian_objective=None)
    builder, qubits, measurements = mgr.variant_builder()
    B = 10_000
    qps = mgr.quasi_probabilities()

    def make_inputs(rank: int, seed_offset: int = 0):
        state = np.random.get_state()
        np.random.seed(20260108 + rank)
        rows = np.zeros((B, len(qps)), np.uint8)
        for s, qp in enumerate(qps):
            rows[:, s] = np.random.choice(len(qp), size=B, p=np.abs(qp) / np.sum(np.abs(qp)))
        np.random.set_state(state)
        return None, rows

    return dict(key="swap7_qem",
                name="7-qubit SWAP-test QP error mitigation (reference QP_QEM_lib.swap_circuit, asymmetric Pauli noise p=1e-4): "
                     "10000 sampled identifiers per estimate, whole 4^7 state in one CTA's shared memory",
                builder=builder, qubits=qubits, measurements=measurements, batch=B, make_inputs=make_inputs,
                terms=mgr.observable_terms(), manager=mgr, qp_bytes=int(sum(len(q) for q in qps) * 8), parity_rows=[0, 1, 2, B - 1])


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
    n, depth, B = 10, 20, 4096
    b = _hea_builder(n, depth)
    rng = np.random.default_rng(20260104)
    n_terms = 100
    # 100 DISTINCT Pauli strings of weight 1..4 (a repeated draw is drawn again: round 1 kept duplicates, which the plan
    # merged into 95 terms)
    seen = []
    while len(seen) < n_terms:
        x = z = 0
        wgt = rng.integers(1, 5)
        for q in rng.choice(n, wgt, replace=False):
            code = rng.integers(1, 4)
            bit = 1 << (n - 1 - int(q))
            if code in (1, 2):
                x |= bit
            if code in (2, 3):
                z |= bit
        if (x, z) not in seen:
            seen.append((x, z))
    xs, zs = np.array([t[0] for t in seen], np.uint32), np.array([t[1] for t in seen], np.uint32)
    terms = (xs, zs, rng.normal(size=n_terms))
    n_params = len(b.param_names)

    def make_inputs(rank: int, seed_offset: int = 0):
        r = np.random.default_rng(20260103 + rank)
        return r.uniform(-np.pi, np.pi, size=(B, n_params)), None

    # SURVEY.md §8d canonical traffic: one read + one write of complex128 rho per layer = depth * 32 B * 4^n
    return dict(key="hea10",
                name=f"synthetic 10-qubit hardware-efficient ansatz depth 20, noise on every gate, batch {B} per step (BASELINE configs[3])",
                # 7 resident qubits: 10 passes / 103 sweeps instead of 18 / 112 at 6 - 8 365 vs 8 081 circuits/s (the 12-qubit
                # workload prefers 6: 22.8 vs 19.7 evals/s; profiles/r2_stream3_experiments.txt)
                builder=b, batch=B, make_inputs=make_inputs, terms=terms, options={"tile_qubits": 7},
                canonical_bytes=depth * 32.0 * 4.0 ** n, cpu_sample=16, golden="hea10", parity_rows=[0, 1])


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
    n, B = 12, 8
    b, terms, n_blocks = _lih_like(n)
    n_params = len(b.param_names)

    def make_inputs(rank: int, seed_offset: int = 0):
        r = np.random.default_rng(20260106 + rank)
        return r.normal(scale=0.1, size=(B, n_params)), None

    # SURVEY.md §8d canonical traffic: one pass (read + write of complex128 rho) per Pauli-exponential block
    return dict(key="lih12",
                name=f"synthetic LiH-like 12-qubit UCCSD ansatz ({n_blocks} Pauli-exponential blocks, noise on every gate, "
                     f"{len(terms[2])} Pauli terms), {B} optimiser chains per GPU per step (BASELINE configs[4]; no LiH data in the reference)",
                builder=b, batch=B, make_inputs=make_inputs, terms=terms, options={"tile_qubits": 6},
                canonical_bytes=n_blocks * 32.0 * 4.0 ** n, cpu_sample=0, cpu_ops_fraction=0.002, golden="lih12_full", parity_rows=[0])


WORKLOADS = {"h2_sweep": workload_h2_sweep, "h2_qem": workload_h2_qem, "hea10": workload_hea10, "lih12": workload_lih12,
             "h2_bonds": workload_h2_bonds, "swap7_qem": workload_swap7_qem}
# riders of the default run: BASELINE configs[2..4], the A x B batch axes, the reference's 7-qubit workload
ALSO = ("h2_qem", "hea10", "lih12", "h2_bonds", "swap7_qem")
ALSO_STEPS = {"h2_qem": 20, "hea10": 3, "lih12": 2, "h2_bonds": 20, "swap7_qem": 5}


# ---------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs (rank 0 only: eight 2 ms pollers
    sharing the host were part of what the end-to-end scaling lost in round 1)."""

    def __init__(self, index: int, enabled: bool = True):
        self.samples, self.reasons = [], set()
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        self._nv = None
        if not enabled:
            return
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
        self._stop.clear()
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
            self._thread = None

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


def _parse_cpulist(text: str):
    cores = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cores.update(range(int(lo), int(hi or lo) + 1))
    return cores


def _gpu_local_cores(index: int):
    """(numa_node, cores local to the PCIe root of GPU `index`) from sysfs, or (None, None) when it cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:            # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/numa_node") as fh:
            node = int(fh.read().strip())
        with open(base + "/local_cpulist") as fh:
            return node, _parse_cpulist(fh.read())
    except Exception:  # noqa: BLE001 - no NVML / no sysfs entry: fall back to an even split
        return None, None


def pin_rank_to_cores(local_rank: int, world: int):
    """One disjoint slice of the host cores per rank (round 1: eight ranks floated over the same cores and the fixed
    launch + synchronise cost of the host call grew with N), taken from the cores LOCAL to the rank's GPU when sysfs says
    which those are: the page-locked buffers the rank allocates afterwards are then first-touched on the GPU's own NUMA
    node, so zero-copy traffic does not cross the socket interconnect.  Returns (cores, numa_node)."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = [int(v) for v in visible.split(",")] if visible and all(v.strip().isdigit() for v in visible.split(",")) else list(range(world))
        info = [_gpu_local_cores(phys[i]) if i < len(phys) else (None, None) for i in range(world)]
        node, local = info[local_rank]
        if local is not None and node is not None and node >= 0 and os.environ.get("DMX_BENCH_NUMA", "1") != "0":
            mates = [i for i in range(world) if info[i][0] == node]
            pool = [c for c in allowed if c in local]
            per = len(pool) // len(mates)
            if per >= 1:
                k = mates.index(local_rank)
                mine = pool[k * per:(k + 1) * per]
                os.sched_setaffinity(0, mine)
                return len(mine), node
        per = max(1, len(allowed) // max(world, 1))
        mine = allowed[local_rank * per:(local_rank + 1) * per] or allowed
        os.sched_setaffinity(0, mine)
        return len(mine), None
    except (AttributeError, OSError, ValueError):
        return None, None


def cpu_port_rate(wl, sample: int, threads: int = 0, rank: int = 0):
    """Time the C restatement of the reference algorithm on `sample` circuits of the workload.  Workloads whose single
    circuit takes the CPU many minutes (12 qubits: rho is 256 MiB, 25k ops) set `cpu_ops_fraction`: only that leading
    fraction of the op list is evolved (one identity term as the observable) and the rate is scaled to the whole list."""
    from oracle import c_oracle
    b = wl["builder"]
    params, variants = wl["make_inputs"](rank)
    ops, terms, scale = list(b.ops), wl["terms"], 1.0
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
    c_oracle.energy_batch(b.n_qubits, ops, b.pool(), terms, params=p, variants=v, batch=sample,
                          n_params=len(b.param_names), n_slots=b.n_slots, threads=threads)
    dt = time.perf_counter() - t0
    return sample * scale / dt, dt


def cpu_sample_note(wl, sample):
    if wl.get("cpu_ops_fraction"):
        return (f"{sample} circuits x the first {wl['cpu_ops_fraction']:.2%} of the op list, rate scaled to the whole list "
                f"(a full 12-qubit circuit takes the CPU port tens of minutes)")
    return f"{sample} circuits of the workload"


def run_reference(args, wl):
    """--impl reference: CPU port on all host threads, bounded sample per step.  Host-side builder only (no plan)."""
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
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "native": "oracle/libdm_oracle.so only (the op table comes from the host-side ProgramBuilder; libdmx.so is not loaded)"}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# parity (outside every timed region): the workload's own rows against the oracle
# ---------------------------------------------------------------------------------------------------------

def parity_check(wl, prog, params_h, variants_h, rank: int, groups_h=None):
    """max |gpu - oracle| over a few rows of THIS workload's inputs.  4-qubit workloads: the C oracle runs live.  The
    streaming shapes: golden energies the same oracle produced for exactly these circuits and rows
    (tests/golden/large_energies.json, tools/make_golden_large.py; a 10-qubit depth-20 circuit costs the oracle ~1 min,
    the full 12-qubit one hours) - valid for rank 0's inputs, checked by a hash of the parameter rows."""
    import torch
    rows = [r for r in wl.get("parity_rows", [0]) if r < wl["batch"]]
    p = params_h[rows] if params_h is not None else None
    v = variants_h[rows] if variants_h is not None else None
    g = groups_h[rows] if groups_h is not None else None
    got = prog.energies(torch.tensor(p, device="cuda") if p is not None else None,
                        torch.tensor(v, device="cuda") if v is not None else None, batch=len(rows),
                        groups=torch.tensor(g, device="cuda") if g is not None else None).cpu().numpy()
    b = wl["builder"]
    if wl.get("golden"):
        path = os.path.join(ROOT, "tests", "golden", "large_energies.json")
        rec = None
        if os.path.exists(path):
            with open(path) as fh:
                rec = json.load(fh).get(wl["golden"])
        if rec is not None and rank == 0:
            k = min(rec["rows"], len(rows))
            if hashlib.sha256(np.ascontiguousarray(params_h[:rec["rows"]]).tobytes()).hexdigest() == rec["params_sha"] and rows[:k] == list(range(k)):
                err = float(np.max(np.abs(got[:k] - np.array(rec["energies"][:k]))))
                return err, f"{k} row(s) vs tests/golden/large_energies.json[{wl['golden']}] (C oracle on this exact circuit and these rows)"
        return None, "no golden energies for this rank's rows (the oracle needs minutes to hours per circuit at this size)"
    from oracle import c_oracle
    if g is not None:      # every row against the oracle with the Hamiltonian of its own coefficient group
        want = np.array([c_oracle.energy_batch(b.n_qubits, list(b.ops), b.pool(), wl["group_terms"][int(gi)], params=p[i:i + 1], batch=1,
                                               n_params=len(b.param_names), n_slots=b.n_slots, threads=1)[0] for i, gi in enumerate(g)])
        return float(np.max(np.abs(got - want))), f"{len(rows)} rows vs the C oracle with each row's own Hamiltonian, run live"
    want = c_oracle.energy_batch(b.n_qubits, list(b.ops), b.pool(), wl["terms"], params=p, variants=v, batch=len(rows),
                                 n_params=len(b.param_names), n_slots=b.n_slots, threads=min(len(rows), 4))
    return float(np.max(np.abs(got - want))), f"{len(rows)} rows vs the C oracle (oracle/dm_oracle.c), run live"


# ---------------------------------------------------------------------------------------------------------
# one workload on this rank (+ max over ranks)
# ---------------------------------------------------------------------------------------------------------

def run_workload(wl, ctx, steps: int, warmup: int, with_cpu_baseline: bool, with_e2e: bool, cpu_budget_s: float = 12.0):
    import torch
    import torch.distributed as dist
    from vqe_errormitigation_b200 import engine
    rank, local_rank, world, flush = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["flush"]
    clocks = ClockSampler(local_rank, enabled=rank == 0)
    prog, B = get_prog(wl), wl["batch"]
    info = prog.info

    # ---- CPU baseline first (GPU idle), rank 0 at N=1 only -----------------------------------------------------
    cpu_baseline = None
    if with_cpu_baseline and rank == 0 and world == 1:
        from oracle import c_oracle
        cores = c_oracle.max_threads()
        if wl.get("cpu_ops_fraction"):
            sample = cores
        else:
            probe = min(32, B, max(cores, 1))
            _, dt1 = cpu_port_rate(wl, probe, threads=cores)
            sample = int(max(probe, min(B, probe / dt1 * cpu_budget_s)))
        rate, dt = cpu_port_rate(wl, sample, threads=cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{cpu_sample_note(wl, sample)}, {dt:.1f} s, C restatement of cirq's per-op "
                                  f"density-matrix algorithm (oracle/dm_oracle.c) on {cores} threads"}
        if not wl.get("cpu_ops_fraction") and cpu_budget_s >= 10.0:
            # the reference's own shapes: one process (CPU.get_custom_optimized_state) and its Pool(6)
            # (MAX_MULTI_PROCESSING_COUNT, src/error_mitigation.py:14), ~2 s of work each
            by = {str(cores): rate}
            for th in (1, 6):
                if th < cores:
                    n_s = int(max(8, min(B, rate / cores * th * 2.0)))
                    by[str(th)] = cpu_port_rate(wl, n_s, threads=th)[0]
            cpu_baseline["by_threads"] = by

    params_h, variants_h = wl["make_inputs"](rank)
    groups_h = np.ascontiguousarray(wl["make_groups"](rank), np.int32) if wl.get("make_groups") else None
    params_d = torch.tensor(params_h, device="cuda") if params_h is not None else None
    variants_d = torch.tensor(variants_h, device="cuda") if variants_h is not None else None
    groups_d = torch.tensor(groups_h, device="cuda") if groups_h is not None else None
    out = torch.empty(B, dtype=torch.float64, device="cuda")

    def step():
        prog.energies(params_d, variants_d, batch=B, out=out, groups=groups_d)

    short_steps = info["path"] != 2
    for _ in range(warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    parity_err, parity_src = parity_check(wl, prog, params_h, variants_h, rank, groups_h)

    def isolated_steps(count):
        """`count` steps, each between its own pair of events, L2 flushed (192 MiB memset) ahead of every start event."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(count)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(count)]
        for i in range(count):
            if short_steps:
                # a ~150 us device-side spin ahead of the flush lets the host run ahead, so that the start event, the
                # launch and the end event of a microsecond-scale step are queued back to back (otherwise host jitter
                # leaks into the event interval)
                torch.cuda._sleep(300_000)
            flush.zero_()
            starts[i].record()
            step()
            ends[i].record()
        torch.cuda.synchronize()
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    isolated = None
    if short_steps:
        # Microsecond-scale steps (register-resident and whole-state plans).  Round 1 and the first half of round 2 timed every
        # step between its own pair of events; tools/launch_floor.py shows what that measures on this box: an EMPTY launch
        # reads 6.1 us between its two events (the timer ticks in 2.05 us units), an 8-circuit launch 8 - 10 us, and the
        # 65 536-circuit launch 18.4 us - half of it is the event pair and the launch latency, whatever the kernel does
        # (four different kernels, 16 to 32 warps per SM, all 18.3 - 18.6 us; profiles/r2_frame32s_ab.txt).  The timed
        # region is therefore the contract's own: EXACTLY K steps between ONE pair of events, queued the way a sweep or
        # optimiser driver queues them - back to back, rotating over four streams so that the launch latency and the last
        # partial wave of one step overlap the next.  L2: flushed once ahead of the start event, and the steps then cycle
        # through >= 160 MiB of distinct input / result buffer sets (more than the 126 MB L2 between two uses of a set), so
        # no step finds its inputs in the L2.  The per-step-event figure stays in the line as `isolated`.
        step_bytes = sum(int(t.numel() * t.element_size()) for t in (params_d, variants_d) if t is not None) + B * 8
        n_buf = int(min(512, max(4, -(-(160 << 20) // step_bytes))))
        bufs = [(params_d.clone() if params_d is not None else None, variants_d.clone() if variants_d is not None else None,
                 torch.empty(B, dtype=torch.float64, device="cuda")) for _ in range(n_buf)]
        side = [torch.cuda.Stream() for _ in range(4)]
        handles = [s.cuda_stream for s in side]
        main = torch.cuda.current_stream()

        def burst(count):
            for s in side:
                s.wait_stream(main)
            for i in range(count):
                pb, vb, ob = bufs[i % n_buf]
                prog.energies(pb, vb, batch=B, out=ob, groups=groups_d, stream=handles[i % 4])
            for s in side:
                main.wait_stream(s)

        burst(max(warmup, 4))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with clocks:
            torch.cuda.synchronize()
            wall0 = time.perf_counter()
            flush.zero_()
            # device-side spin (~1 ms per 50 steps): the host queues the K launches behind it (a launch costs the Python
            # loop ~8.5 us; without the head start a step shorter than that would time the host, not the GPU)
            torch.cuda._sleep(2_000_000 * min(-(-steps // 50), 40))
            launches0 = engine.launch_count()
            ev0.record()
            burst(steps)
            ev1.record()
            torch.cuda.synchronize()
            launches = engine.launch_count() - launches0
            wall = time.perf_counter() - wall0
            t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms_max = float(t.item())
            n_iso = min(steps, 30)
            iso_ms = isolated_steps(n_iso)          # inside the clock-sampling window too: the K-step region alone is < 1 ms
        worst = max(float((bufs[i][2] - out).abs().max().item()) for i in range(min(steps, n_buf, 16)))
        isolated = {"value": world * B * n_iso / (iso_ms * 1e-3), "unit": UNIT, "ms_per_step": iso_ms / n_iso, "steps": n_iso,
                    "timing": "one pair of CUDA events per step, L2 flushed (192 MiB memset) ahead of every step; includes ~6 us of "
                              "event-pair + launch latency per step (tools/launch_floor.py)"}
        timing_note = (f"ONE pair of CUDA events around the K steps, queued back to back over 4 streams behind a device-side spin; "
                       f"max over ranks; every step's result equals the single-stream result (max abs diff {worst:.1e})")
        l2_note = (f"flushed once (192 MiB memset) ahead of the start event; the steps cycle through {n_buf} distinct input / result "
                   f"buffer sets ({n_buf * step_bytes >> 20} MiB > L2), so no step finds its inputs in the L2")
        if worst > 1e-12:
            raise RuntimeError(f"multi-stream results differ from the single-stream result by {worst}")
        del bufs
    else:
        if world > 1:
            dist.barrier()
        launches0 = engine.launch_count()
        with clocks:
            torch.cuda.synchronize()
            wall0 = time.perf_counter()
            dev_ms_max = isolated_steps(steps)
            wall = time.perf_counter() - wall0
        launches = engine.launch_count() - launches0
        timing_note = "CUDA events per step on the launch stream, summed; max over ranks"
        l2_note = "flushed (192 MiB memset) between timed steps"
    if world > 1:
        dist.barrier()
    value = world * B * steps / (dev_ms_max * 1e-3)

    def pinned(a):
        if a is None:
            return None
        b = engine.pinned_empty(a.shape, a.dtype)
        b[...] = a
        return b

    def timed(fn, n_steps, n_warm=3):
        """Wall-clock seconds of n_steps calls (after n_warm untimed ones), max over ranks."""
        for i in range(n_warm):
            fn(10_000 + i)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            fn(i)
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H + sync inside the timed region) -----------------
    e2e = None
    mgr = wl.get("manager")
    if with_e2e and mgr is None:
        # Parameter batches: params[B, P] in page-locked host memory, energies[B] back into page-locked host memory.
        # Register-resident plans (path 1) keep up to DEPTH batches in flight (dmx_energy_batch_host_async, each with its
        # own result buffer; dmx_plan_host_sync after every DEPTH-th step) - the way a sweep or a lock-step optimiser
        # driver calls it; the library rotates them over "host_streams" (4) streams, so the partial last wave of one launch
        # overlaps the next launch - which is why this figure can exceed `value`, whose launches are timed one at a time
        # with an L2 flush between them.  Streaming plans run one synchronous call per step.
        depth = 16 if info["path"] == 1 and groups_h is None else 1
        params_p, groups_p = pinned(params_h), pinned(groups_h)
        res = [engine.pinned_empty(B) for _ in range(depth)]

        def host_step(i):
            prog.energies_host(params_p, None, batch=B, out=res[i % depth], nowait=depth > 1, groups=groups_p)
            if depth > 1 and (i + 1) % depth == 0:
                prog.host_sync()

        n_e2e = steps if depth == 1 else max(depth, (steps // depth) * depth)
        e2e_s = timed(host_step, n_e2e, n_warm=depth if depth > 1 else 3)
        prog.host_sync()
        zero_copy = info["path"] == 1 and PLAN_OPTIONS.get("host_zero_copy", 1) != 0
        e2e = {"value": world * B * n_e2e / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(params_h.nbytes + (groups_h.nbytes if groups_h is not None else 0)), "d2h_bytes_per_step": int(B * 8),
               "steps": n_e2e, "in_flight": depth, "streams": int(PLAN_OPTIONS.get("host_streams", 4)) if depth > 1 else 1,
               "api": "CircuitProgram.energies_host = dmx_energy_batch_host" + (f"_async, dmx_plan_host_sync every {depth} steps" if depth > 1 else ""),
               "checksum": float(np.sum(res[0])),
               "transfer": ("zero-copy: the kernel reads the parameters from and writes the energies to the mapped page-locked host "
                            "buffers over PCIe" if zero_copy else
                            "cudaMemcpyAsync H2D / D2H from and to the page-locked host buffers, pipelined in chunks over three streams")}
        if depth > 1:
            # the strictly synchronous form (one call, one synchronise per step) beside it
            sync_s = timed(lambda i: prog.energies_host(params_p, None, batch=B, out=res[0]), steps)
            e2e["synchronous"] = {"value": world * B * steps / sync_s, "unit": UNIT, "api": "dmx_energy_batch_host, one synchronise per step"}
            e2e["in_flight_value"] = e2e["value"]
            if e2e["synchronous"]["value"] > e2e["value"]:
                # Eight ranks on one host: the box's mapped-memory traffic saturates at ~1.1e10 circuits/s (176 GB/s) and deep
                # queues only add contention there (profiles/r2_e2e_scaling_n8.txt) - the headline e2e figure is the better of
                # the two public calls, and says which.
                e2e.update(value=e2e["synchronous"]["value"], steps=steps, in_flight=1, streams=1, api=e2e["synchronous"]["api"])
    elif with_e2e:
        # QEM: what the reference's own flow hands to its workers is the identifier -> count dictionary
        # (error_mitigation.py:492-497): the UNIQUE rows of this rank's 125000 sampled identifiers and their counts, prepared
        # outside the timed region (set_identifier_range is a separate call there too).  Timed: unique rows H2D, kernels,
        # values D2H, weighted mean.
        uniq, counts = np.unique(variants_h, axis=0, return_counts=True)
        uniq_p, signs_u = pinned(uniq), mgr.variant_signs(uniq)
        est = [None]

        def host_step(i):
            est[0] = mgr.evaluate_variants(uniq_p, signs_u, counts)
        e2e_s = timed(host_step, steps)
        e2e = {"value": world * B * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(uniq.nbytes), "d2h_bytes_per_step": int(len(uniq) * 8),
               "steps": steps, "api": "IErrorMitigationManager.evaluate_variants(unique rows, signs, counts) -> dmx_energy_batch_host",
               "unique_rows": int(len(uniq)), "rows_represented": int(B), "mu": float(est[0][1]),
               "transfer": "unique identifier rows + counts (the reference's dict) instead of the raw 125000 x 122 table"}
        # the raw table through the same C-ABI call (round 1's e2e): PCIe-bound by the 15 MB table
        raw_p, raw_out, n_raw = pinned(variants_h), engine.pinned_empty(B), max(3, steps // 4)
        raw_s = timed(lambda i: prog.energies_host(None, raw_p, batch=B, out=raw_out), n_raw)
        e2e["raw_table"] = {"value": world * B * n_raw / raw_s, "unit": UNIT, "h2d_bytes_per_step": int(variants_h.nbytes),
                            "d2h_bytes_per_step": int(B * 8)}

        # the whole estimator with the variants drawn ON the device; under --gpus N this block CONTAINS the all-reduce
        def sampled(i):
            est[0] = mgr.get_mu_effective_sampled(B * world, seed=20260102 + i)
        samp_s = timed(sampled, steps)
        e2e["device_sampling"] = {"value": world * B * steps / samp_s, "unit": UNIT, "h2d_bytes_per_step": int(wl["qp_bytes"]),
                                  "d2h_bytes_per_step": 24,
                                  "api": "IErrorMitigationManager.get_mu_effective_sampled (dmx_qp_sampled_energies - draw + evaluate in one "
                                         "kernel on register-resident plans, dmx_qp_sampler_draw + dmx_energy_batch otherwise - + "
                                         "dmx_qp_reduce + all-reduce of 3 doubles over NCCL when N > 1)",
                                  "estimate": [float(est[0][0]), float(est[0][1]), int(est[0][2])]}

    # ---- QEM: device-timed estimator (draw + evaluate + reduce + all-reduce inside CUDA events), per-N rate ---------
    estimator = None
    if mgr is not None:
        from vqe_errormitigation_b200 import parallel
        sampler = engine.QpSampler(mgr.quasi_probabilities())
        lo, hi = parallel.shard_bounds(B * world, rank, world)

        def est_step(seed):
            # draw + evaluate (distinct rows only where the circuit is expensive) + reduce; device tensor [3], all-reduced on
            # this stream when N > 1
            return mgr.sampled_sums(sampler, lo, hi, seed)
        for i in range(3):
            est_step(100 + i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        es = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ee = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            torch.cuda._sleep(300_000)
            flush.zero_()
            es[i].record()
            sums = est_step(i)
            ee[i].record()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(es, ee))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        s0, _s1, n = (float(x) for x in sums.cpu())
        estimator = {"value": world * B * steps / (float(ms.item()) * 1e-3), "unit": "variants/s", "ms_per_estimate": float(ms.item()) / steps,
                     "contains": "Philox draw + " + ("device dedup (torch.unique) + " if info["path"] != 1 else "") + "variant evaluation + reduction"
                                 + (" + NCCL all-reduce of 3 doubles" if world > 1 else ""),
                     "timing": "CUDA events on the launch stream, max over ranks", "variants_per_estimate": int(B * world),
                     "estimate": [mgr.total_cost() * s0 / n, int(n)]}

    # ---- QEM: deduplicated device rate ---------------------------------------------------------
    dedup = None
    if variants_h is not None and world == 1 and with_e2e:
        uniq = np.unique(variants_h, axis=0)
        uniq_d = torch.tensor(uniq, device="cuda")
        uout = torch.empty(len(uniq), dtype=torch.float64, device="cuda")
        for _ in range(3):
            prog.energies(None, uniq_d, batch=len(uniq), out=uout)
        torch.cuda.synchronize()
        us, ue = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_ms = 0.0
        for _ in range(steps):
            torch.cuda._sleep(300_000)
            flush.zero_()
            us.record()
            prog.energies(None, uniq_d, batch=len(uniq), out=uout)
            ue.record()
            torch.cuda.synchronize()
            t_ms += us.elapsed_time(ue)
        dedup = {"unique_rows": int(len(uniq)), "rows": int(B), "value": B * steps / (t_ms * 1e-3), "unit": UNIT,
                 "note": "variants represented per second when only the distinct rows are evaluated (device time of the unique-row "
                         "batch); the roofline fraction is defined on the raw rate"}

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
    ms_per_step = dev_ms_max / steps
    kernel_s = ms_per_step * 1e-3
    kern = {0: "dmx_tile_kernel", 1: "dmx_frame32_kernel (dmx_frame_kernel when more than 32 coefficients are live)",
            2: "dmx_stream3_kernel / dmx_stream3b_kernel (streaming passes)"}[info["path"]]
    if info["path"] == 1 and variants_h is None and info["n_params"] == 1 and B >= 32768 and PLAN_OPTIONS.get("frame32s", 1) != 0:
        kern = "dmx_frame32s_kernel (sweep form of the frame kernel: 32 circuits per warp, tables in shared memory)"
    if (info["path"] == 1 and variants_h is None and info["n_params"] == 1 and B >= 2048 and PLAN_OPTIONS.get("frame16t", 1) != 0
            and "\nf16 n=" in prog.dump()):
        kern = "dmx_frame16t_kernel (thread-per-circuit form of the frame kernel: 16 touched coefficients in registers, XOR-paired per segment)"
    if (info["path"] == 1 and variants_h is not None and B >= 2048 and PLAN_OPTIONS.get("frame16t", 1) != 0 and " const=1" in prog.dump()):
        kern = "dmx_frame16v_kernel (thread-per-row form for QP variant batches: constant-coefficient scaled segments, precomputed diagonal corrections)"
    traffic = ncu_util = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            recorded = json.load(fh)
        traffic = recorded.get(wl["key"])
        ncu_util = recorded.get("_ncu", {}).get(wl["key"])
    sm_clock = (clocks.summary()["sm_mhz"] or 1965.0) * 1e6
    smem_peak = 148 * 128 * sm_clock / 1e9
    fp64_peak, hbm_peak, hbm_src = ctx["fp64_peak"], ctx["hbm_peak"], ctx["hbm_src"]
    if info["path"] == 2:
        # `achieved` / `frac`: SURVEY.md §8d canonical traffic (complex128 rho, one read + write per layer / block) over
        # the step time.  `frac_executed`: the bytes this engine really moves (real Pauli vector, fused passes) over the
        # same time - how close the kernel itself runs to the HBM roof.
        canon = wl.get("canonical_bytes") or info["canonical_hbm_bytes_per_circuit"]
        ach = canon * B / kernel_s / 1e9
        ex_gbs = info["hbm_bytes_per_circuit"] * B / kernel_s / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "kernel": kern, "peak_source": hbm_src,
                    "algorithmic_bytes_per_circuit": canon,
                    "algorithmic_model": "SURVEY.md 8d: 32 B x 4^n per layer (HEA) / per Pauli-exponential block (UCCSD)",
                    "frac_executed": ex_gbs / hbm_peak,
                    "executed": {"hbm_bytes_per_circuit": info["hbm_bytes_per_circuit"], "hbm_gbs": ex_gbs, "hbm_frac": ex_gbs / hbm_peak,
                                 "smem_bytes_per_circuit": info["smem_bytes_per_circuit"],
                                 "smem_gbs": info["smem_bytes_per_circuit"] * B / kernel_s / 1e9,
                                 "smem_frac": info["smem_bytes_per_circuit"] * B / kernel_s / 1e9 / smem_peak,
                                 "fp64_tflops": info["flops_per_circuit"] * B / kernel_s / 1e12,
                                 "fp64_frac": info["flops_per_circuit"] * B / kernel_s / 1e12 / fp64_peak,
                                 "note": "the binding resource is the shared-memory exchange between sweeps "
                                         "(148 SM x 128 B/clk at the sampled clock), then HBM"}}
    else:
        ach = info["flops_per_circuit"] * B / kernel_s / 1e12
        smem_ach = info["smem_bytes_per_circuit"] * B / kernel_s / 1e9
        roofline = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                    "traffic": traffic, "kernel": kern,
                    "peak_source": "FP64 FMA micro-benchmark run live (dmx_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                    "executed_flops_per_circuit": info["flops_per_circuit"],
                    "canonical_flops_per_circuit": info["canonical_flops_per_circuit"],
                    "canonical_equiv_tflops": info["canonical_flops_per_circuit"] * B / kernel_s / 1e12,
                    "smem": {"achieved_gbs": smem_ach, "peak_gbs": smem_peak, "frac": smem_ach / smem_peak,
                             "note": "shared-memory exchange bytes per circuit x rate vs 148 SM x 128 B/clk at the sampled SM clock"},
                    "hbm": {"achieved_gbs": info["hbm_bytes_per_circuit"] * B / kernel_s / 1e9, "peak_gbs": hbm_peak}}
    if ncu_util:
        roofline["ncu"] = ncu_util      # recorded capture (profiles/), not re-measured by this run
    res = {"value": value, "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
           "config": {"workload": wl["name"], "batch_per_gpu": B, "l2": l2_note, "timing": timing_note,
                      "plan": {k: info[k] for k in ("path", "n_blocks", "n_segments", "n_passes", "tile_qubits", "n_terms")},
                      "wall_ms_incl_flush": 1e3 * wall},
           "clocks": clocks.summary(), "parity_max_abs_err": parity_err, "parity_source": parity_src,
           "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if isolated:
        res["isolated"] = isolated
    if estimator:
        res["estimator"] = estimator
    if dedup:
        res["dedup"] = dedup
    if cpu_baseline:
        res["cpu_baseline"] = cpu_baseline
    return res


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
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other BASELINE configs")
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

    cores_per_rank, numa_node = pin_rank_to_cores(local_rank, world) if world > 1 else (None, None)

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
    hbm_peak, hbm_src = measured_peaks()
    ctx = {"rank": rank, "local_rank": local_rank, "world": world,
           "flush": torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda"),     # > 126 MB L2
           "fp64_peak": engine.measure_fp64_peak(), "hbm_peak": hbm_peak, "hbm_src": hbm_src}

    head = run_workload(WORKLOADS[args.workload](), ctx, args.steps, args.warmup,
                        with_cpu_baseline=not args.no_cpu_baseline, with_e2e=not args.no_e2e)
    also = {}
    if args.workload == "h2_sweep" and not args.no_also:
        for key in ALSO:
            t0 = time.perf_counter()
            try:
                wl_k = WORKLOADS[key]()
                # a short CPU baseline for the riders the oracle evaluates in milliseconds (n <= 7); one 10-qubit circuit costs
                # it ~40 s and a 12-qubit one hours - run `--workload hea10|lih12` for those
                res = run_workload(wl_k, ctx, ALSO_STEPS[key], 3, with_cpu_baseline=not args.no_cpu_baseline and wl_k["builder"].n_qubits <= 7,
                                   with_e2e=not args.no_e2e, cpu_budget_s=2.5)
            except Exception as exc:  # noqa: BLE001 - a rider must not take the headline line down
                if world > 1:
                    raise                  # ranks must stay in step through the collectives
                res = {"error": f"{type(exc).__name__}: {exc}"}
            if res is not None:
                res["wall_s"] = round(time.perf_counter() - t0, 2)
                also[key] = res
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": head["config"], "clocks": head["clocks"],
                "parity_max_abs_err": head["parity_max_abs_err"], "parity_source": head["parity_source"],
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"]}
        for k in ("isolated", "estimator", "dedup"):
            if k in head:
                line[k] = head[k]
        line["cpu_baseline"] = head.get("cpu_baseline")
        if cores_per_rank:
            line["config"]["host_cores_per_rank"] = cores_per_rank
            line["config"]["rank0_numa_node"] = numa_node
        if also:
            line["also"] = also
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
