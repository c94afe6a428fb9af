"""Noisy-VQE minimisation driver (mirror of ``src/calculate_minimisation.py``), running on the B200 engine.

Same module constants, same functions and the same on-disk result format (``jsonpickle_compat``) as the
reference (``calculate_minimisation.py:8-71``).  Plotting (``read_and_plot``, reference ``:89-90``) is out of
scope; ``python -m vqe_errormitigation_b200.calculate_minimisation`` runs the reference's ``__main__``
experiment (``:74-88``) and prints the stored containers instead.
"""
from __future__ import annotations

import os

from . import jsonpickle_compat as jsonpickle
from .data_containers.helper_interfaces.i_collection import IMeasurementCollection
from .data_containers.model_hydrogen import HydrogenAnsatz
from .processors.processor_classic import CPU

DATA_DIR = os.getcwd() + '/classic_minimisation'
QPU_ITER = 10
CPU_ITER = 10


def check_dir(rel_dir: str) -> bool:
    return os.path.exists(rel_dir)


def create_dir(rel_dir: str) -> bool:
    """Returns whether the directory already existed; creates it otherwise (reference :18-23)."""
    existed = check_dir(rel_dir)
    if not existed:
        os.mkdir(rel_dir)
    return existed


def write_object(obj: object, filename: str):
    create_dir(DATA_DIR)
    with open(f'{DATA_DIR}/{filename}', 'w') as wf:
        wf.write(jsonpickle.encode(value=obj, indent=4))


def read_object(filename: str) -> object:
    file_path = f'{DATA_DIR}/{filename}'
    if not check_dir(file_path):
        raise FileNotFoundError(f'Indicated file path does not exist: {file_path}')
    with open(file_path, 'r') as rf:
        return jsonpickle.decode(rf.read())


def calculate_and_write_collection(collection: IMeasurementCollection, filename: str):
    """noise models x bond lengths x CPU_ITER restarts, each a COBYLA run of QPU_ITER-shot energies
    (reference :38-48); the list of ``IContainer`` is stored under ``DATA_DIR/filename``."""
    collection.container = CPU.get_collection_ground_states(collection=collection, cpu_iter=CPU_ITER, qpu_iter=QPU_ITER)
    write_object(collection.container, filename)


if __name__ == '__main__':
    from .data_containers.helper_interfaces.i_noise_wrapper import INoiseModel
    from .error_mitigation import SingleQubitPauliChannel, TwoQubitPauliChannel

    clean_ansatz = HydrogenAnsatz()
    filename = 'H2_temptest6'
    parameter_space = [.7414]
    noise_space = [INoiseModel(noise_gates_1q=[SingleQubitPauliChannel(p_x=p, p_y=p, p_z=6 * p)],
                               noise_gates_2q=[TwoQubitPauliChannel(p_x=p, p_y=p, p_z=6 * p)],
                               description=f'asymmetric depolarization (p_tot={16 * p})') for p in [0.000, 1e-4]]
    measure_collection = IMeasurementCollection(w=clean_ansatz, p_space=parameter_space, n_space=noise_space)
    calculate_and_write_collection(collection=measure_collection, filename=filename)
    for container in read_object(filename):
        print(container.label, container.molecule_param, container.measured_value, container.measured_std, container.fci_value)
