"""H2 / STO-3G UCCSD ansatz (mirror of ``src/data_containers/model_hydrogen.py``).

``observable_measurement`` returns the sampled-energy objective the reference's COBYLA loop minimises
(``src/processors/processor_classic.py:77-92``): five measurement circuits (one computational-basis circuit
for the Z / ZZ terms, one per XXYY-type term), shot histograms, parity probabilities, and the weighted sum
over Hamiltonian terms (reference lines 83-177).  The five density-matrix evolutions run as one batch on the
GPU; only the multinomial shot sampling stays on the host (it is defined by numpy's global RNG stream).
"""
from __future__ import annotations

import itertools
import os
from typing import Callable, List, Sequence, Tuple

import numpy as np

from .. import cirq_compat as cirq
from ..openfermion_compat import MolecularData, QubitOperator, jordan_wigner
from .helper_interfaces.i_parameter import IParameter
from .helper_interfaces.i_wave_function import IGeneralizedUCCSD


class HydrogenAnsatz(IGeneralizedUCCSD):
    def __init__(self, mol_param: float = .7414):
        super().__init__(IParameter({"r0": mol_param}))

    def _generate_molecule(self, p: IParameter) -> MolecularData:
        """``cwd/mol_data/H2_<r>.hdf5`` if present (the reference's layout, model_hydrogen.py:31-57), else the
        packaged table decoded from the same files.  The reference would run psi4 for unknown bond lengths
        (:59-65); psi4 is not available here, so that case raises FileNotFoundError."""
        description = str(p["r0"])
        geometry = [("H", (0., 0., 0.)), ("H", (0., 0., p["r0"]))]
        filename = os.path.join(os.getcwd(), "mol_data", "H2_" + description)
        molecule = MolecularData(geometry, "sto-3g", 1, description=description, filename=filename)
        molecule.load()
        return molecule

    def initial_state(self, qubits: Sequence[cirq.Qid]):
        """Hartree-Fock |1100>: rx(pi) on the first two qubits (reference :70-81)."""
        yield [cirq.rx(np.pi).on(qubits[0]), cirq.rx(np.pi).on(qubits[1])]

    def observable_measurement(self) -> Tuple[Callable[[cirq.Circuit, Callable, int], float], int]:
        hamiltonian = jordan_wigner(self.molecule.get_molecular_hamiltonian())
        qubits = self.qubits
        by_size = {}
        weight = {}
        for term, coefficient in hamiltonian.terms.items():
            by_size.setdefault(len(term), []).append(term)
            weight[term] = coefficient
        to_z_basis = IGeneralizedUCCSD.pauli_basis_map()
        n = len(qubits)

        def parity_probability(bits: np.ndarray, columns: Sequence[int]) -> float:
            """Fraction of shots with even parity on ``columns`` (the (0,0)+(1,1) / eight even-parity
            histogram entries the reference adds up at :136-165)."""
            return float(np.mean(bits[:, columns].sum(axis=1) % 2 == 0))

        def measure_observable(base_circuit: cirq.Circuit, noise_wrapper: Callable, meas_reps: int) -> float:
            # circuit for the one- and two-Z terms: identity "basis change" (+ its noise) on every Z_i
            circuits = []
            c01 = base_circuit.copy()
            for term in by_size.get(1, []):
                c01.append([(to_z_basis[p](qubits[q], 1), noise_wrapper(qubits[q])) for q, p in term])
            c01.append(cirq.measure(q, key=f"Z{i}") for i, q in enumerate(qubits))
            circuits.append(c01)
            for term in by_size.get(4, []):
                c02 = base_circuit.copy()
                c02.append([(to_z_basis[p](qubits[q], 1), noise_wrapper(qubits[q])) for q, p in term])
                c02.append(cirq.measure(q, key=f"{term[i][1]}{term[i][0]}") for i, q in enumerate(qubits))
                circuits.append(c02)
            shots = sample_circuits(circuits, meas_reps)

            expectation = {(): 1.0}
            z_bits = shots[0]
            for i in range(n):
                expectation[((i, "Z"),)] = 2 * parity_probability(z_bits, [i]) - 1
            for i, j in itertools.combinations(range(n), 2):
                expectation[((i, "Z"), (j, "Z"))] = 2 * parity_probability(z_bits, [i, j]) - 1
            for term, bits in zip(by_size.get(4, []), shots[1:]):
                expectation[term] = 2 * parity_probability(bits, list(range(n))) - 1
            result = 0
            for term in itertools.chain.from_iterable(by_size.values()):
                if term in expectation:
                    result += weight[term] * expectation[term]
            return result

        return measure_observable, 5


def sample_circuits(circuits: List[cirq.Circuit], repetitions: int) -> List[np.ndarray]:
    """Evolve the circuits on the GPU, then draw ``repetitions`` computational-basis shots from each
    (cirq's run(): evolve once, sample |diag rho| renormalised).  Returns one [reps, n_qubits] bit array per
    circuit, columns in qubit order."""
    from .. import engine
    out = []
    for circuit in circuits:
        prog, params = engine.program_for_circuit(circuit)
        probs = prog.probabilities(params, renormalise=True)[0]
        outcomes = np.random.choice(len(probs), size=repetitions, p=probs)
        n = prog.n_qubits
        bits = ((outcomes[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1).astype(np.int8)
        out.append(bits)
    return out
