"""B200-native noisy density-matrix engine with the call surface of MiniSean/VQE_ErrorMitigation's hot path.

Layout (mirrors the reference's ``src/`` tree for the path in scope):
    cirq_compat            circuit IR with cirq's names (cirq itself is not installable here)
    openfermion_compat     FermionOperator / QubitOperator / jordan_wigner / MolecularData stand-ins
    processors.processor_quantum.QPU, processors.processor_classic.CPU
    circuit_noise_extension.Noisify
    data_containers.model_hydrogen.HydrogenAnsatz, data_containers.helper_interfaces.*
    error_mitigation       BasisUnitarySet, QuasiCircuitIdentifier, IErrorMitigationManager, channel classes
    qp_qem_lib             the collaborator dialect (QP_QEM_lib.py) incl. the 7-qubit SWAP test
    engine                 circuit -> dmx_op table -> libdmx.so (CUDA), batched evaluation
    parallel               batch sharding over torch.distributed
    csrc/                  CUDA kernels + C ABI (include/dmx.h)
"""
__version__ = "0.1.0"
