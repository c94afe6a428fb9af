"""Stand-in for the slice of ``openfermion==0.11.0`` the reference touches.

openfermion is a pinned, un-vendored dependency of the reference (``requirements.txt:17``) and is
not installable here, so the algorithms it contributes to the hot path's *inputs* are restated
from their published definitions:

* ``FermionOperator`` / ``QubitOperator`` term dictionaries with **insertion-ordered** products
  (the order of the 8 ansatz Pauli strings at ``src/data_containers/helper_interfaces/
  i_wave_function.py:137-140`` depends on it, and with noise after every gate that order changes
  the density matrix at O(p));
* ``uccsd_generator``, ``normal_ordered``, ``hermitian_conjugated``, ``jordan_wigner``,
  ``count_qubits``, ``get_sparse_operator`` (call sites: ``i_wave_function.py:98-107,128,139``,
  ``src/processors/processor_quantum.py:24,34,78``, ``src/data_containers/model_hydrogen.py:91``);
* ``MolecularData.load`` / ``get_molecular_hamiltonian`` (``model_hydrogen.py:48-57``).

Validation (tests/test_molecule.py): the 15-term H2 Hamiltonian reproduces the stored ``hf_energy``
and ``fci_energy`` of all 30 complete ``H2_r.hdf5`` saves to < 1e-12.
"""
from __future__ import annotations

import itertools
import json
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse

EQ_TOLERANCE = 1e-8

# --------------------------------------------------------------------------------------------
# symbolic operators
# --------------------------------------------------------------------------------------------


class _SymbolicOperator:
    """Sum of coefficient * (ordered tuple of factors); ``terms`` keeps insertion order."""

    def __init__(self, term=None, coefficient=1.0):
        self.terms: Dict[tuple, complex] = {}
        if term is None:
            return
        if isinstance(term, str):
            term = self._parse_string(term)
        coefficient, term = self._simplify(tuple(term), coefficient)
        self.terms[term] = coefficient

    # subclasses override
    @staticmethod
    def _parse_string(text: str) -> tuple:
        raise NotImplementedError

    @staticmethod
    def _simplify(term: tuple, coefficient):
        return coefficient, term

    # -- algebra (loop order follows openfermion: left terms outer, right terms inner) ------------
    def __imul__(self, other):
        if isinstance(other, (int, float, complex, np.number)):
            for t in self.terms:
                self.terms[t] *= other
            return self
        if not isinstance(other, type(self)):
            raise TypeError("cannot multiply operators of different kinds")
        product: Dict[tuple, complex] = {}
        for lt, lc in self.terms.items():
            for rt, rc in other.terms.items():
                c, t = self._simplify(lt + rt, lc * rc)
                product[t] = product.get(t, 0.0) + c
        self.terms = product
        return self

    def __mul__(self, other):
        out = self.copy()
        out *= other
        return out

    def __rmul__(self, other):
        if isinstance(other, (int, float, complex, np.number)):
            return self * other
        return NotImplemented

    def __iadd__(self, other):
        if not isinstance(other, type(self)):
            raise TypeError("cannot add operators of different kinds")
        for t, c in other.terms.items():
            if t in self.terms:
                self.terms[t] = self.terms[t] + c
                if abs(self.terms[t]) < EQ_TOLERANCE:
                    del self.terms[t]
            else:
                self.terms[t] = c
        return self

    def __add__(self, other):
        out = self.copy()
        out += other
        return out

    def __isub__(self, other):
        neg = other.copy()
        neg *= -1.0
        self += neg
        return self

    def __sub__(self, other):
        out = self.copy()
        out -= other
        return out

    def __neg__(self):
        return self * -1.0

    def __truediv__(self, divisor):
        return self * (1.0 / divisor)

    def copy(self):
        out = type(self)()
        out.terms = dict(self.terms)
        return out

    def __iter__(self):
        for t, c in self.terms.items():
            single = type(self)()
            single.terms[t] = c
            yield single

    def __len__(self):
        return len(self.terms)

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        keys = set(self.terms) | set(other.terms)
        return all(abs(self.terms.get(k, 0.0) - other.terms.get(k, 0.0)) < EQ_TOLERANCE for k in keys)

    def compress(self, abs_tol: float = EQ_TOLERANCE):
        for t in list(self.terms):
            c = self.terms[t]
            if abs(c) < abs_tol:
                del self.terms[t]
                continue
            if abs(complex(c).imag) < abs_tol:
                c = complex(c).real
            elif abs(complex(c).real) < abs_tol:
                c = 1j * complex(c).imag
            self.terms[t] = c

    def __repr__(self):
        if not self.terms:
            return "0"
        return " +\n".join(f"{c} [{self._term_str(t)}]" for t, c in self.terms.items())

    @staticmethod
    def _term_str(term):
        return " ".join(map(str, term))


class FermionOperator(_SymbolicOperator):
    """Terms are tuples of ``(mode, 1=creation|0=annihilation)``; no implicit reordering."""

    @staticmethod
    def _parse_string(text: str) -> tuple:
        out = []
        for tok in text.split():
            if tok.endswith("^"):
                out.append((int(tok[:-1]), 1))
            else:
                out.append((int(tok), 0))
        return tuple(out)

    @staticmethod
    def _term_str(term):
        return " ".join(f"{i}{'^' if d else ''}" for i, d in term)


_PAULI_PRODUCT = {
    ("I", "I"): (1.0, "I"), ("I", "X"): (1.0, "X"), ("I", "Y"): (1.0, "Y"), ("I", "Z"): (1.0, "Z"),
    ("X", "I"): (1.0, "X"), ("Y", "I"): (1.0, "Y"), ("Z", "I"): (1.0, "Z"),
    ("X", "X"): (1.0, "I"), ("Y", "Y"): (1.0, "I"), ("Z", "Z"): (1.0, "I"),
    ("X", "Y"): (1.0j, "Z"), ("Y", "X"): (-1.0j, "Z"),
    ("Y", "Z"): (1.0j, "X"), ("Z", "Y"): (-1.0j, "X"),
    ("Z", "X"): (1.0j, "Y"), ("X", "Z"): (-1.0j, "Y"),
}


class QubitOperator(_SymbolicOperator):
    """Terms are tuples of ``(qubit, 'X'|'Y'|'Z')`` sorted by qubit; same-qubit factors multiply out."""

    @staticmethod
    def _parse_string(text: str) -> tuple:
        out = []
        for tok in text.split():
            out.append((int(tok[1:]), tok[0].upper()))
        return tuple(out)

    @staticmethod
    def _simplify(term: tuple, coefficient):
        if not term:
            return coefficient, term
        ordered = sorted(term, key=lambda f: f[0])  # stable: keeps left-to-right order per qubit
        out: List[Tuple[int, str]] = []
        cur_q, cur_p = ordered[0]
        for q, p in ordered[1:]:
            if q == cur_q:
                phase, cur_p = _PAULI_PRODUCT[(cur_p, p)]
                coefficient = coefficient * phase
            else:
                if cur_p != "I":
                    out.append((cur_q, cur_p))
                cur_q, cur_p = q, p
        if cur_p != "I":
            out.append((cur_q, cur_p))
        return coefficient, tuple(out)

    @staticmethod
    def _term_str(term):
        return " ".join(f"{p}{q}" for q, p in term)


# --------------------------------------------------------------------------------------------
# fermionic utilities
# --------------------------------------------------------------------------------------------


def hermitian_conjugated(op: FermionOperator) -> FermionOperator:
    """Reverse each product and flip creation/annihilation; coefficients conjugated."""
    out = FermionOperator()
    for term, c in op.terms.items():
        conj_term = tuple((i, 1 - d) for i, d in reversed(term))
        out.terms[conj_term] = np.conj(c)
    return out


def _normal_ordered_term(term: Sequence[Tuple[int, int]], coefficient) -> FermionOperator:
    """Bubble a ladder product into creation-first / descending-index order (anticommutation rules)."""
    term = list(term)
    result = FermionOperator()
    for i in range(1, len(term)):
        for j in range(i, 0, -1):
            right, left = term[j], term[j - 1]
            if right[1] and not left[1]:
                # a_p a†_q = delta_pq - a†_q a_p
                term[j - 1], term[j] = right, left
                coefficient = -coefficient
                if right[0] == left[0]:
                    contracted = term[:j - 1] + term[j + 1:]
                    result += _normal_ordered_term(contracted, -coefficient)
            elif right[1] == left[1]:
                if right[0] == left[0]:
                    return result  # a a = a† a† = 0
                if right[0] > left[0]:
                    term[j - 1], term[j] = right, left
                    coefficient = -coefficient
    result += FermionOperator(tuple(term), coefficient)
    return result


def normal_ordered(op: FermionOperator) -> FermionOperator:
    out = FermionOperator()
    for term, c in op.terms.items():
        out += _normal_ordered_term(term, c)
    return out


def uccsd_generator(single_amplitudes, double_amplitudes, anti_hermitian: bool = True) -> FermionOperator:
    """t_ij a†_i a_j + t_ijkl a†_i a_j a†_k a_l, minus the Hermitian conjugates (App. B.3)."""
    gen = FermionOperator()
    singles = np.asarray(single_amplitudes)
    doubles = np.asarray(double_amplitudes)
    for i, j in zip(*singles.nonzero()):
        i, j = int(i), int(j)
        t = singles[i, j]
        gen += FermionOperator(((i, 1), (j, 0)), t)
        if anti_hermitian:
            gen += FermionOperator(((j, 1), (i, 0)), -t)
    for i, j, k, l in zip(*doubles.nonzero()):
        i, j, k, l = int(i), int(j), int(k), int(l)
        t = doubles[i, j, k, l]
        gen += FermionOperator(((i, 1), (j, 0), (k, 1), (l, 0)), t)
        if anti_hermitian:
            gen += FermionOperator(((l, 1), (k, 0), (j, 1), (i, 0)), -t)
    return gen


class InteractionOperator:
    """H = constant + sum h[p,q] a†_p a_q + sum g[p,q,r,s] a†_p a†_q a_r a_s (spin-orbital tensors)."""

    def __init__(self, constant: float, one_body_tensor: np.ndarray, two_body_tensor: np.ndarray):
        self.constant = constant
        self.one_body_tensor = one_body_tensor
        self.two_body_tensor = two_body_tensor
        self.n_qubits = one_body_tensor.shape[0]

    def to_fermion_operator(self) -> FermionOperator:
        op = FermionOperator((), self.constant)
        n = self.n_qubits
        for p, q in itertools.product(range(n), repeat=2):
            c = self.one_body_tensor[p, q]
            if c != 0.0:
                op += FermionOperator(((p, 1), (q, 0)), c)
        for p, q, r, s in itertools.product(range(n), repeat=4):
            c = self.two_body_tensor[p, q, r, s]
            if c != 0.0:
                op += FermionOperator(((p, 1), (q, 1), (r, 0), (s, 0)), c)
        return op


def count_qubits(op) -> int:
    if isinstance(op, InteractionOperator):
        return op.n_qubits
    n = 0
    for term in op.terms:
        for f in term:
            n = max(n, f[0] + 1)
    return n


def jordan_wigner(op) -> QubitOperator:
    """a†_j -> (X_j - iY_j)/2 * Z_{j-1}...Z_0 ; products taken left to right in dict order."""
    if isinstance(op, InteractionOperator):
        qo = jordan_wigner(op.to_fermion_operator())
        qo.compress()
        return qo
    out = QubitOperator()
    for term, c in op.terms.items():
        acc = QubitOperator((), c)
        for mode, dagger in term:
            z_string = tuple((k, "Z") for k in range(mode))
            x_part = QubitOperator(z_string + ((mode, "X"),), 0.5)
            y_part = QubitOperator(z_string + ((mode, "Y"),), -0.5j if dagger else 0.5j)
            acc *= x_part + y_part
        out += acc
    return out


_PAULI_MATRIX = {
    "I": np.eye(2, dtype=complex),
    "X": np.array([[0, 1], [1, 0]], dtype=complex),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "Z": np.array([[1, 0], [0, -1]], dtype=complex),
}


def get_sparse_operator(op: QubitOperator, n_qubits: Optional[int] = None) -> scipy.sparse.csc_matrix:
    """Kronecker expansion with qubit 0 as the most significant factor (csc, complex)."""
    if n_qubits is None:
        n_qubits = count_qubits(op)
    dim = 1 << n_qubits
    total = scipy.sparse.csc_matrix((dim, dim), dtype=complex)
    for term, c in op.terms.items():
        labels = ["I"] * n_qubits
        for q, p in term:
            labels[q] = p
        mat = scipy.sparse.identity(1, dtype=complex, format="csc")
        for lab in labels:
            mat = scipy.sparse.kron(mat, _PAULI_MATRIX[lab], format="csc")
        total = total + c * mat
    return total.tocsc()


def pauli_masks(op: QubitOperator, n_qubits: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(xmask, zmask, coeff) per term; qubit q <-> bit (n-1-q); Y sets both masks (App. E.7)."""
    xs, zs, cs = [], [], []
    for term, c in op.terms.items():
        x = z = 0
        for q, p in term:
            bit = 1 << (n_qubits - 1 - q)
            if p in ("X", "Y"):
                x |= bit
            if p in ("Z", "Y"):
                z |= bit
        xs.append(x)
        zs.append(z)
        cs.append(complex(c).real)
        if abs(complex(c).imag) > 1e-12:
            raise ValueError("Hamiltonian term with complex coefficient")
    return np.array(xs, np.uint32), np.array(zs, np.uint32), np.array(cs, np.float64)


# --------------------------------------------------------------------------------------------
# molecular data
# --------------------------------------------------------------------------------------------

_PACKAGED_TABLE = os.path.join(os.path.dirname(__file__), "moldata", "h2_sto3g.json")
_packaged_cache: Optional[dict] = None


def _packaged() -> dict:
    global _packaged_cache
    if _packaged_cache is None:
        with open(_PACKAGED_TABLE) as fh:
            _packaged_cache = json.load(fh)
    return _packaged_cache


class MolecularData:
    """Holds what ``MolecularData.load()`` would give the reference (integrals, amplitudes, energies)."""

    _FIELDS = ("nuclear_repulsion", "hf_energy", "mp2_energy", "cisd_energy", "ccsd_energy", "fci_energy",
               "n_qubits", "n_orbitals", "n_electrons")

    def __init__(self, geometry=None, basis=None, multiplicity=None, charge=0, description="", filename=""):
        self.geometry = geometry
        self.basis = basis
        self.multiplicity = multiplicity
        self.charge = charge
        self.description = description
        self.filename = filename[:-5] if filename.endswith(".hdf5") else filename
        self.nuclear_repulsion = None
        self.hf_energy = None
        self.mp2_energy = None
        self.cisd_energy = None
        self.ccsd_energy = None
        self.fci_energy = None
        self.n_qubits = None
        self.n_orbitals = None
        self.n_electrons = None
        self.one_body_integrals = None
        self.two_body_integrals = None
        self.ccsd_single_amps = None
        self.ccsd_double_amps = None

    # -- loading --------------------------------------------------------------------------------
    def load(self) -> "MolecularData":
        path = f"{self.filename}.hdf5"
        if os.path.exists(path):
            self._load_hdf5(path)
        else:
            self._load_packaged(self.description)
        return self

    def _load_hdf5(self, path: str):
        from .moldata.hdf5_lite import Hdf5File
        f = Hdf5File(path)
        for name in self._FIELDS:
            if name in f:
                value = f[name]
                setattr(self, name, value if not isinstance(value, np.ndarray) else value)
        for name in ("one_body_integrals", "two_body_integrals", "ccsd_single_amps", "ccsd_double_amps"):
            if name in f:
                value = f[name]
                setattr(self, name, np.asarray(value, dtype=float) if isinstance(value, np.ndarray) else None)

    def _load_packaged(self, description: str):
        table = _packaged()["molecules"]
        if description not in table:
            raise FileNotFoundError(
                f"no molecular data for H2 r={description!r}: neither {self.filename}.hdf5 nor the packaged table has it "
                f"(psi4 is not available; reference falls back to run_psi4 at model_hydrogen.py:59-65)")
        rec = table[description]
        for name in self._FIELDS:
            setattr(self, name, rec.get(name))
        self.one_body_integrals = np.array(rec["one_body_integrals"], float)
        self.two_body_integrals = np.array(rec["two_body_integrals"], float).reshape((self.n_orbitals,) * 4)
        nq = self.n_qubits
        self.ccsd_single_amps = np.zeros((nq, nq))
        for (i, j), v in rec.get("ccsd_single_amps", []):
            self.ccsd_single_amps[i, j] = v
        self.ccsd_double_amps = np.zeros((nq,) * 4)
        for (i, j, k, l), v in rec.get("ccsd_double_amps", []):
            self.ccsd_double_amps[i, j, k, l] = v

    # -- Hamiltonian ----------------------------------------------------------------------------
    def get_molecular_hamiltonian(self) -> InteractionOperator:
        """spinorb_from_spatial + InteractionOperator(nuc, one, two/2); |x| < 1e-8 zeroed (App. B.3)."""
        h = np.asarray(self.one_body_integrals)
        g = np.asarray(self.two_body_integrals)
        m = h.shape[0]
        n = 2 * m
        one = np.zeros((n, n))
        two = np.zeros((n, n, n, n))
        for p, q in itertools.product(range(m), repeat=2):
            one[2 * p, 2 * q] = h[p, q]
            one[2 * p + 1, 2 * q + 1] = h[p, q]
            for r, s in itertools.product(range(m), repeat=2):
                v = g[p, q, r, s]
                two[2 * p, 2 * q + 1, 2 * r + 1, 2 * s] = v
                two[2 * p + 1, 2 * q, 2 * r, 2 * s + 1] = v
                two[2 * p, 2 * q, 2 * r, 2 * s] = v
                two[2 * p + 1, 2 * q + 1, 2 * r + 1, 2 * s + 1] = v
        one[np.abs(one) < EQ_TOLERANCE] = 0.0
        two[np.abs(two) < EQ_TOLERANCE] = 0.0
        return InteractionOperator(self.nuclear_repulsion, one, 0.5 * two)
