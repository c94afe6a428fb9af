"""``python tools/check_sampled_rows.py first_seed n_seeds`` - IErrorMitigationManager._sampled_energy_rows (IWaveFunction objective) on the CPU double against the exact
sum_t w_t <P_t> of each variant circuit (large shot count, 6-sigma band)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np, pytest
import oracle_backend
from oracle import dm_oracle as O
from vqe_errormitigation_b200 import cirq_compat as cirq
from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager, TwoQubitDepolarizingChannel, SingleQubitPauliChannel, TwoQubitPauliChannel
from vqe_errormitigation_b200.openfermion_compat import jordan_wigner
from vqe_errormitigation_b200.processors.processor_quantum import QPU
mp = pytest.MonkeyPatch(); oracle_backend.install(mp)
fails = 0
for seed in range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    r0 = float(rng.choice([0.7414, 0.5, 1.2]))
    w = HydrogenAnsatz(mol_param=r0)
    p = float(rng.choice([1e-3, 5e-3]))
    model = [INoiseModel([cirq.depolarize(p)], [TwoQubitDepolarizingChannel(p)], "d"),
             INoiseModel([SingleQubitPauliChannel(p, p, 6 * p)], [TwoQubitPauliChannel(p, p, 6 * p)], "a"),
             INoiseModel([cirq.bit_flip(p), cirq.phase_flip(p)], [], "f")][int(rng.integers(3))]
    n_w = INoiseWrapper(w, model)
    w.operator_parameters["alpha0"] = float(rng.normal() * 0.2)
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    objective = n_w if rng.random() < 0.5 else w
    mgr = IErrorMitigationManager(clean, model, objective)
    gate_ops = list(clean.all_operations())
    rows = np.zeros((3, len(gate_ops)), np.uint8)
    for r in range(1, 3):
        for s in rng.choice(len(gate_ops), size=int(rng.integers(1, 4)), replace=False):
            rows[r, s] = int(rng.integers(1, 4)) if len(gate_ops[s].qubits) == 1 else 16 * int(rng.integers(0, 4)) + int(rng.integers(0, 4))
    shots = 200000
    mgr._density_representation, mgr._meas_reps = False, 1
    got = mgr._sampled_energy_rows(rows, np.full(3, shots))
    ham = jordan_wigner(w.molecule.get_molecular_hamiltonian())
    basis = w.pauli_basis_map()
    for r in range(3):
        base = IErrorMitigationManager.get_updated_circuit(clean, model.get_operators, list(rows[r]))
        exact = 0.0
        for term, coeff in ham.terms.items():
            if len(term) == 0:
                exact += complex(coeff).real
                continue
            c = base.copy()
            if len(term) == 4:
                c.append([(basis[pp](w.qubits[q], 1), model.get_operators(w.qubits[q])) for q, pp in term])
            else:
                for q in range(4):
                    c.append([cirq.I(w.qubits[q]), model.get_operators(w.qubits[q])])
            rho = O.simulate(4, O.from_circuit(c, w.qubits)[1])
            zmask = sum(1 << (3 - q) for q, _ in term)
            par = np.array([bin(i & zmask).count("1") % 2 for i in range(16)])
            exact += complex(coeff).real * float(np.sum(np.where(par == 0, 1, -1) * O.probabilities(rho)))
        tol = 6 * 1.0 / np.sqrt(shots)
        if abs(got[r] - exact) > tol:
            fails += 1
            print("FAIL seed", seed, "row", r, got[r], exact, tol, flush=True)
    print("seed", seed, "ok", np.round(got, 4), flush=True)
print("failures", fails)
