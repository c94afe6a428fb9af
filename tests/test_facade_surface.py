"""The cirq names and QP_QEM_lib helpers the reference reaches through ``cirq.Simulator`` / ``cirq.sample``
(processor_quantum.py:48-56, QP_QEM_lib.py:520-590).  Every test runs twice, like tests/test_trajectories.py: on the CPU test
double (host logic; evolutions go through the C oracle) and, marked ``gpu``, on the real engine through the C ABI."""
import numpy as np
import pytest

import oracle_backend
from oracle import dm_oracle as O
from vqe_errormitigation_b200 import cirq_compat as cirq
from vqe_errormitigation_b200 import engine, trajectories
from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
from vqe_errormitigation_b200.processors.processor_quantum import QPU


@pytest.fixture(params=["oracle_double", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "oracle_double":
        oracle_backend.install(monkeypatch)
    else:
        import torch
        assert torch.cuda.is_available()
    trajectories._CACHE.clear()
    engine._PROGRAM_CACHE.clear()
    yield request.param
    trajectories._CACHE.clear()
    engine._PROGRAM_CACHE.clear()


def test_simulator_and_sample_facades(backend):
    """``cirq.Simulator`` / ``cirq.sample`` as ``QPU.get_trial_results`` / ``get_noisy_trial_result`` call them."""
    q = cirq.LineQubit.range(2)
    body = [cirq.ry(1.1).on(q[0]), cirq.rx(0.4).on(q[1]), cirq.CNOT(q[0], q[1])]
    # simulate: the state vector of the pure final state, up to the one global phase
    psi = cirq.Simulator().simulate(cirq.Circuit(body)).state_vector()
    rho = O.simulate(2, O.from_circuit(cirq.Circuit(body))[1])
    assert np.allclose(np.outer(psi, psi.conj()), rho, atol=1e-12)
    assert abs(np.imag(psi[np.argmax(abs(psi))])) < 1e-15
    with pytest.raises(ValueError):
        cirq.Simulator().simulate(cirq.Circuit(body + [cirq.depolarize(0.1).on(q[0])]))
    # run / sample: terminal measurement records with the distribution of diag(rho)
    measured = cirq.Circuit(body + [cirq.measure(q[0], q[1], key="m")])
    np.random.seed(11)
    res = QPU.get_trial_results(measured, None, 20000)
    freq = np.array([res.histogram(key="m")[k] for k in range(4)]) / 20000
    assert np.allclose(freq, np.real(np.diag(rho)), atol=4 * 0.5 / np.sqrt(20000))
    model = INoiseModel([cirq.depolarize(0.2)], [TwoQubitDepolarizingChannel(0.2)], "d")
    wrapper = INoiseWrapper(HydrogenAnsatz(), model)
    noisy = QPU.get_noisy_trial_result(measured, None, 20000, noise_model=wrapper)
    noisy_rho = O.simulate(2, O.from_circuit(cirq._with_noise(cirq.Circuit(body), wrapper))[1])
    nfreq = np.array([noisy.histogram(key="m")[k] for k in range(4)]) / 20000
    assert np.allclose(nfreq, np.real(np.diag(noisy_rho)), atol=4 * 0.5 / np.sqrt(20000))
    assert not np.allclose(np.diag(noisy_rho), np.diag(rho), atol=0.02)       # the noise model did act
    # module paths the reference's annotations spell out
    assert cirq.circuits.circuit is cirq.Circuit and cirq.study.resolver is cirq.ParamResolver
    assert cirq.ops.gate_features.SingleQubitGate is cirq.SingleQubitGate


def test_gate_set_tomography_estimate(backend):
    """``QP_QEM_lib.o_tilde_1q_tomo`` (:520-590): sampled entries against Tr(Q_j . G(rho_k)) for a unitary, a noisy unitary and a
    post-selected projector tuple of ``basis_ops_sim``."""
    from vqe_errormitigation_b200 import qp_qem_lib as L
    q = cirq.LineQubit(0)
    reps = 20000
    tol = 5 / np.sqrt(reps)
    preps = [np.eye(2), O.X, cirq.unitary(cirq.ry(np.pi / 2)), cirq.unitary(cirq.rx(-np.pi / 2))]
    rho0 = np.diag([1.0, 0.0]).astype(complex)
    states = [u @ rho0 @ u.conj().T for u in preps]
    paulis = [O.I2, O.X, O.Y, O.Z]

    def exact(channel):
        return np.array([[np.real(np.trace(pj @ channel(rk))) for rk in states] for pj in paulis])

    np.random.seed(3)
    u = cirq.unitary(cirq.rx(0.7))
    got = L.o_tilde_1q_tomo(cirq.rx(0.7).on(q), reps, noisy=False, error=0.0)
    assert np.allclose(got.real, exact(lambda r: u @ r @ u.conj().T), atol=tol) and np.all(got[0] == 1)

    def dep(r, p=0.1):
        return (1 - 4 * p / 3) * r + (4 * p / 3) * np.trace(r) * np.eye(2) / 2

    got = L.o_tilde_1q_tomo(cirq.rx(0.7).on(q), reps, noisy=True, error=0.1)
    assert np.allclose(got.real, exact(lambda r: dep(dep(u @ dep(r) @ u.conj().T))), atol=tol)
    # pi_x-type entry of basis_ops_sim: (rx^3, measure 'M', rx) - post-selected on M = 0, unnormalised
    ops = L.apply_to_qubit(L.basis_ops_sim()[11], q)
    mats = [cirq.unitary(g) for g in L.basis_ops()[11]]
    full = mats[2] @ mats[1] @ mats[0]
    got = L.o_tilde_1q_tomo(ops, reps, noisy=False, error=0.0)
    assert np.allclose(got.real, exact(lambda r: full @ r @ full.conj().T), atol=tol)
