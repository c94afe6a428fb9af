/*
 * dmx.h — C ABI of the B200 noisy density-matrix engine (libdmx.so).
 *
 * The reference (MiniSean/VQE_ErrorMitigation) has no FFI of its own: its hot path is the handful of
 * calls into cirq's DensityMatrixSimulator followed by Tr(rho H).  This header is the explicit form
 * of that boundary.  Each entry point names the reference call it stands in for
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C, no C++/torch types; all device buffers are caller-owned (e.g. torch tensors), the
 *     plan owns only its own constant tables.
 *   - qubit 0 is the most significant bit of a basis-state index (cirq's big-endian order); density
 *     matrices are row-major complex128 (re,im interleaved), element (r,c) at rho[(r<<n)+c].
 *   - every function returns 0 on success or a negative dmx_status; dmx_last_error() gives the
 *     message of the last failure on the calling thread.  No exceptions cross the ABI.
 *   - there is no CPU execution path: without a CUDA device the execution calls fail with
 *     DMX_ERR_CUDA.  Plan construction/inspection is host-only work and needs no device.
 *   - a plan is immutable after dmx_plan_finalize(); it may be used from several host threads /
 *     streams concurrently (workspace is per call).  Its device tables are uploaded to the device that is
 *     current at the first execution call; executing it with another current device fails with DMX_ERR_STATE.
 */
#ifndef DMX_H_
#define DMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMX_VERSION 100 /* 0.1.0 */
#define DMX_MAX_QUBITS 14

typedef enum dmx_status {
    DMX_OK = 0,
    DMX_ERR_INVALID = -1,     /* bad argument / malformed op table */
    DMX_ERR_UNSUPPORTED = -2, /* valid request the engine cannot run (e.g. > DMX_MAX_QUBITS) */
    DMX_ERR_CUDA = -3,        /* CUDA runtime error or no device */
    DMX_ERR_NOMEM = -4,
    DMX_ERR_STATE = -5        /* call order (e.g. executing a plan that was not finalized) */
} dmx_status;

/* Operation kinds.  Matrices live in one pool of doubles (complex128 interleaved, row-major); an op
 * refers to its first matrix by offset (in doubles).  A 1-qubit matrix is 8 doubles, a 2-qubit matrix
 * 32 doubles with q0 the more significant qubit (np.kron(A_q0, B_q1) ordering). */
typedef enum dmx_op_kind {
    /* rho <- U rho U^dagger.  U is NOT checked for unitarity: the reference wraps projectors as gates
     * (IBasisGateSingle/Two, src/error_mitigation.py:277-300) and cirq applies them as-is. */
    DMX_OP_UNITARY = 1,
    /* rho <- sum_k K_k rho K_k^dagger, n_mats Kraus operators back to back (a cirq mixture
     * (p_k, U_k) is passed as K_k = sqrt(p_k) U_k; channels of src/error_mitigation.py:743-910,
     * cirq.depolarize / bit_flip / phase_flip / amplitude_damp / phase_damp). */
    DMX_OP_KRAUS = 2,
    /* U = exp(-i theta sigma_axis / 2), theta = scale * params[b, param] + offset: the resolved form of
     * cirq.rx/ry/rz(2*sign*alpha_k) (src/data_containers/helper_interfaces/i_wave_function.py:183,
     * src/processors/processor_quantum.py:63-65).  axis: 0 = X, 1 = Y, 2 = Z.  param = -1: constant angle
     * theta = offset. */
    DMX_OP_ROTATION = 3,
    /* Per-circuit basis operation of quasi-probability error mitigation: circuit b applies
     * B[variants[b, slot]] where B is a table of 16 one-qubit matrices starting at `mat`
     * (BasisUnitarySet order, src/error_mitigation.py:119-136).  On two qubits the index is
     * 16*a + b and the operation is kron(B[a], B[b]) (src/error_mitigation.py:150-153,556-565). */
    DMX_OP_VARIANT = 4,
    /* Terminal measurement marker: ignored by the evolution (cirq evolves once and samples the
     * diagonal); kept so op indices line up with the caller's circuit. */
    DMX_OP_MEASURE = 5
} dmx_op_kind;

/* flag: apply this UNITARY/KRAUS op only in circuits whose variant index at `slot` is non-zero
 * ("noise after correction", src/error_mitigation.py:570-571).  Must directly follow the
 * DMX_OP_VARIANT op of the same slot (other conditional ops of that slot may sit in between). */
#define DMX_OPF_IF_VARIANT_NONZERO 1u

typedef struct dmx_op {
    int32_t kind;     /* dmx_op_kind */
    int32_t n_qubits; /* 1 or 2 */
    int32_t q0, q1;   /* qubit indices (q1 ignored for 1-qubit ops) */
    int32_t mat;      /* offset (in doubles) of the first matrix in the pool; -1 if none */
    int32_t n_mats;   /* UNITARY: 1; KRAUS: number of Kraus operators; VARIANT: 16 */
    int32_t param;    /* ROTATION: parameter column; otherwise -1 */
    int32_t axis;     /* ROTATION: 0/1/2 */
    int32_t slot;     /* VARIANT and conditional ops: column of the variant table; otherwise -1 */
    uint32_t flags;
    double scale;     /* ROTATION */
    double offset;    /* ROTATION */
} dmx_op;

typedef struct dmx_plan dmx_plan;

/* What the planner decided; filled by dmx_plan_get_info. */
typedef struct dmx_plan_info {
    int32_t n_qubits, n_params, n_slots, n_ops_in;
    int32_t n_blocks;       /* fused micro-ops after gate+channel fusion */
    int32_t n_segments;     /* full-vector monomial segments (segment kernel); 0 if not eligible */
    int32_t n_passes;       /* tile passes (1 when the whole state fits one CTA's shared memory) */
    int32_t tile_qubits;    /* qubits resident per tile */
    int32_t path;           /* 0 = tile kernel (whole state in smem), 1 = segment kernel, 2 = streaming */
    int32_t n_terms;        /* Hamiltonian terms */
    int64_t state_bytes;    /* bytes of one state vector in the engine's representation (8 * 4^n) */
    double flops_per_circuit;     /* FP64 flops the planned kernels execute per circuit */
    double hbm_bytes_per_circuit; /* algorithmic HBM bytes per circuit (params/variants/energy + state passes) */
    double smem_bytes_per_circuit;/* shared-memory bytes moved per circuit */
    double canonical_flops_per_circuit; /* SURVEY.md §8d cost model (complex rho) for the same circuit */
    double canonical_hbm_bytes_per_circuit;
    int64_t workspace_bytes_per_state; /* streaming path: workspace the engine needs per circuit in flight
                                          (state_bytes, or 2 * state_bytes when passes re-order the qubits and
                                          therefore ping-pong between two buffers); 0 otherwise */
} dmx_plan_info;

int dmx_version(void);
const char* dmx_last_error(void);

/* ---- plan construction (host only) ---------------------------------------------------------------
 * Stands in for building + resolving the cirq.Circuit that the reference hands to
 * DensityMatrixSimulator.simulate (src/processors/processor_quantum.py:30-31,
 * src/error_mitigation.py:365-374): the op list is the circuit, params are the ParamResolver values,
 * variants are the QuasiCircuitIdentifier ids. */
int dmx_plan_create(int n_qubits, int n_params, int n_slots, const dmx_op* ops, int n_ops,
                    const double* matrix_pool, size_t pool_len, dmx_plan** out);

/* Observable H = sum_t coeff[t] * P_t, P_t given by bit masks over qubits (bit n-1-q <-> qubit q):
 * X on xmask&~zmask, Z on zmask&~xmask, Y on xmask&zmask.  Replaces get_sparse_operator(jordan_wigner(...))
 * + (csc(rho) * H).diagonal().sum().real (src/processors/processor_quantum.py:33-39,
 * src/error_mitigation.py:377-383). */
int dmx_plan_set_hamiltonian(dmx_plan* plan, int n_terms, const uint32_t* xmask, const uint32_t* zmask,
                             const double* coeff);

/* Coefficient groups (optional, before finalize, after dmx_plan_set_hamiltonian): the same n_terms Pauli strings with
 * n_groups different weight rows, coeff[n_groups][n_terms] in the term order of dmx_plan_set_hamiltonian.  The grouped
 * execution calls let circuit b use row group[b].  This is the batch axis "bond lengths" of
 * CPU.get_collection_ground_states (src/processors/processor_classic.py:18-44 loops noise models x bond lengths x
 * restarts serially; IMeasurementCollection.get_wave_functions, src/data_containers/helper_interfaces/i_collection.py:56-78):
 * every H2 geometry has the same 15 Jordan-Wigner strings with different coefficients, so 30 molecules x R restarts
 * are ONE plan and ONE launch.  The un-grouped calls on such a plan use row 0. */
int dmx_plan_set_hamiltonian_groups(dmx_plan* plan, int n_groups, const double* coeff);

/* Options (before finalize).  key/value pairs; unknown keys -> DMX_ERR_INVALID.
 * Per plan:
 *   "fuse"         0/1  gate+channel block fusion (default 1)
 *   "segments"     0/1  allow the segment / fixed-frame kernels when eligible (default 1)
 *   "tile_qubits"  k    resident qubits per streaming tile, 4..7 (default 6)
 *   "remap"        0/1  streaming: re-order the qubits in HBM between passes (default 1)
 *   "triples"      0/1  streaming: three resident codes per thread instead of two (default 1)
 * Execution switches (also per plan - they are read from the plan at launch time, so one plan's setting never changes
 * how another plan runs):
 *   "stream_kernel"   0/1  0 = generic tile kernel for every streaming pass (default 1)
 *   "bulk"            0-2  streaming tile loads: 1 = one cp.async.bulk + mbarrier per contiguous tile (default), 0 = register
 *                          staging, 2 = direct global loads in the first sweep
 *   "direct_store"    0/1  streaming: the last sweep stores its registers straight to HBM where the layout allows (default 1)
 *   "wide"            0/1  streaming: tile-wide Clifford x Pauli-noise sweeps (default 0: measured slower)
 *   "scale_smem"      0/1  streaming: scale tables in shared memory instead of the constant bank (default 0: measured slower)
 *   "stream_chunk"    n    circuits per streaming pass launch, 0 = automatic (default 0)
 *   "frame32"         0/1  compact warp-per-circuit frame kernel when <= 32 coefficients are live (default 1)
 *   "frame32s"        v    sweep form of the frame kernel (one rotation class, no variant table): 0 never, 1 automatic
 *                          (default: from 2^15 circuits a launch, the fewest circuits per warp - 8 / 16 / 24 / 32 - whose
 *                          warps are all resident at once), 8 / 16 = 32 circuits per warp advanced 8 / 16 at a time,
 *                          40 / 48 / 56 / 64 = always 8 / 16 / 24 / 32 circuits per warp
 *   "frame16t"        0-2  thread-per-circuit forms of the frame kernel for plans whose rotations touch at most 16 coefficients
 *                          in an XOR pattern (noisy H2 UCCSD: 16 of 23 live ones): dmx_frame16t_kernel for one-parameter sweeps,
 *                          dmx_frame16v_kernel for QP variant batches of a resolved circuit.  0 = never, 1 = device-resident
 *                          batches from 2048 circuits (default), 2 = also mapped host buffers (measured slower end to end)
 *   "frame16t_scaled" 0/1  sweeps: carry the own-coefficient weights as host-tracked scales, 3 instead of 4 FP64 instructions
 *                          per coefficient and segment, when every rotation moves every touched coefficient (default 1)
 *   "frame16t_threads" n   CTA size of dmx_frame16t_kernel: 64, 128 (default) or 256
 *   "tile_threads"    n    whole-state 7-qubit tiles: 512 threads x 2 coefficient groups (default) or 1024 x 1 (measured slower)
 *   "segment_cpb"     4|8  circuits per warp / thread of the frame kernels (default 8)
 *   "host_zero_copy"  0/1  dmx_energy_batch_host works in place on mapped page-locked buffers (default 1)
 *   "host_streams"    n    dmx_energy_batch_host_async: consecutive in-flight batches rotate over n streams, 1..8 (default 4)
 *   "host_chunk_rows" n    rows per chunk of the dmx_energy_batch_host DMA pipeline, >= 1024 (default 32768)
 */
int dmx_plan_set_option(dmx_plan* plan, const char* key, int value);

/* Lowers, fuses and schedules the op list (host only); uploads tables lazily on first execution. */
int dmx_plan_finalize(dmx_plan* plan);
int dmx_plan_get_info(const dmx_plan* plan, dmx_plan_info* info);
/* Text dump of the lowered program (for tests / debugging).  Returns bytes needed incl. NUL. */
int64_t dmx_plan_dump(const dmx_plan* plan, char* buf, int64_t buf_len);
void dmx_plan_destroy(dmx_plan* plan);

/* ---- execution (device pointers, caller-owned; stream is a cudaStream_t passed as void*) ------------
 * d_params   [B, n_params] float64 row-major (may be NULL when n_params == 0)
 * d_variants [B, n_slots]  uint8  row-major (NULL = all-identity variants)
 * d_workspace: dmx_workspace_bytes(plan, B) bytes (may be NULL when that is 0). */
size_t dmx_workspace_bytes(const dmx_plan* plan, int64_t batch);

/* energies[b] = Re Tr(rho_b H): QPU.get_simulated_noisy_expectation_value
 * (src/processors/processor_quantum.py:28-39) and the density branch of
 * IErrorMitigationManager.process_circuit (src/error_mitigation.py:372-383), batched.
 * rho is NOT renormalised (projector basis ops leave Tr rho < 1). */
int dmx_energy_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants,
                     double* d_energy, void* d_workspace, size_t workspace_bytes, void* stream);

/* As dmx_energy_batch with a coefficient row per circuit: d_group [B] int32 in [0, n_groups) (NULL: row 0). */
int dmx_energy_batch_grouped(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants,
                             const int32_t* d_group, double* d_energy, void* d_workspace, size_t workspace_bytes, void* stream);

/* d_rho [B, 2^n, 2^n] complex128: simulate(...).final_density_matrix. */
int dmx_density_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants,
                      double* d_rho, void* d_workspace, size_t workspace_bytes, void* stream);

/* d_diag [B, 2^n]: |diag rho|, divided by its sum when renormalise != 0 — the probability vector
 * DensityMatrixSimulator.run samples from (src/data_containers/model_hydrogen.py:132,155,
 * src/error_mitigation.py:391). */
int dmx_diag_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants,
                   double* d_diag, int renormalise, void* d_workspace, size_t workspace_bytes, void* stream);

/* d_pauli [B, 4^n]: the engine's native state, r[P] = Tr(rho P) (index: 2 bits per qubit,
 * qubit 0 most significant, code 0=I 1=Z 2=X 3=Y). */
int dmx_pauli_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants,
                    double* d_pauli, void* d_workspace, size_t workspace_bytes, void* stream);

/* d_out[3] = { sum_b w[b]*v[b], sum_b (w[b]*v[b])^2, sum_b count[b] } with
 * w[b] = d_weights[b] (sign * prod(weights) * count) — the per-rank partial of np.mean at
 * src/error_mitigation.py:399,451; the 24-byte cross-GPU sum is the caller's all-reduce.
 * d_counts may be NULL (every row counts 1).  Deterministic (fixed reduction tree). */
int dmx_qp_reduce(const double* d_values, const double* d_weights, const int32_t* d_counts, int64_t batch,
                  double* d_out3, void* stream);

/* Device-side draw of QP circuit variants — replaces the host loop of QuasiCircuitIdentifier._get_identifiers
 * (src/error_mitigation.py:309-335: per gate np.random.choice(len(qp), p=|qp|/sum|qp|); sign = sign prod qp[idx],
 * :321-323) and IErrorMitigationManager.set_identifier_range (:484-489) when 1e5..1e6 samples are wanted.
 * qp [n_slots][stride] (host): the quasi-probabilities of every gate slot, n_choices[s] <= min(stride, 256) valid
 * entries per row.  RNG: Philox4x32-10, counter (sample lo, sample hi, slot / 4, 0), key = seed; slot s takes word
 * s % 4 of the call as u = (word + 0.5) / 2^32 and draws the first index whose cumulative |qp| / sum|qp| exceeds u.
 * Sample b of a call is global sample first_sample + b, so ranks drawing disjoint ranges of one stream need no
 * communication.  d_variants [batch][n_slots] uint8, d_signs [batch] in {-1, 0, +1}.  Statistical (not bitwise)
 * parity with numpy's generator; bit-exact against the numpy Philox restatement in tests/philox_ref.py.
 * dmx_qp_sampler_create uploads the tables once (current device); dmx_qp_sampler_draw is one kernel launch on
 * `stream`; dmx_qp_sample_variants = create + draw + synchronise + destroy. */
typedef struct dmx_qp_sampler dmx_qp_sampler;
int dmx_qp_sampler_create(int n_slots, const int32_t* n_choices, const double* qp, int stride, dmx_qp_sampler** out);
int dmx_qp_sampler_draw(const dmx_qp_sampler* sampler, uint64_t seed, uint64_t first_sample, int64_t batch,
                        uint8_t* d_variants, double* d_signs, void* stream);
void dmx_qp_sampler_destroy(dmx_qp_sampler* sampler);
int dmx_qp_sample_variants(int n_slots, const int32_t* n_choices, const double* qp, int stride, uint64_t seed,
                           uint64_t first_sample, int64_t batch, uint8_t* d_variants, double* d_signs, void* stream);

/* Draw and evaluate in one call: d_energy[b] = energy of the variant row that dmx_qp_sampler_draw(sampler, seed, first_sample, ...)
 * would write for sample b, d_signs[b] = its sign - the inner loop of IErrorMitigationManager.get_mu_effective
 * (src/error_mitigation.py:401-452: identifiers -> process_circuit -> sign * weight * value) without the identifier table: on
 * plans with the thread-per-row form (resolved Clifford + Pauli-noise circuits such as noisy H2, see "frame16t") one thread
 * runs a sample's Philox stream, keeps its few non-identity draws and evolves the row (dmx_frame16q_kernel); other whole-state
 * plans draw into stream-ordered scratch and run dmx_energy_batch.  Bit-identical to the two-step path.  The plan must have no
 * circuit parameters (a resolved circuit) and as many variant slots as the sampler; streaming plans: DMX_ERR_UNSUPPORTED. */
int dmx_qp_sampled_energies(dmx_plan* plan, const dmx_qp_sampler* sampler, uint64_t seed, uint64_t first_sample, int64_t batch,
                            double* d_energy, double* d_signs, void* stream);

/* Shot sampling on the device - the sampling half of DensityMatrixSimulator.run as the reference uses it
 * (src/data_containers/model_hydrogen.py:132,155: evolve once, draw `repetitions` outcomes from |diag rho|, histogram) and the
 * parity read-out of measure_observable (:103-112,136-166: expectation = 2 P(even parity on the term's qubits) - 1).
 * d_probs [rows, dim] probability vectors (dmx_diag_batch output; dim = 2^n <= 128), `shots` draws per row (or
 * d_row_shots[r] when that device array is given: meas_reps x count of a QP identifier, src/error_mitigation.py:386-388) from
 * Philox4x32-10 (counter (shot / 4 lo, shot / 4 hi, row lo, row hi) with row = first_row + r, key = seed, word shot % 4 as
 * u = (word + 0.5) / 2^32, outcome = first cumulative entry > u).  h_masks [n_masks] (host): bit n-1-q <-> qubit q; every
 * mask is evaluated on the SAME shots.  d_expect [rows, n_masks].  Statistical parity with numpy's generator, bit-exact
 * against the numpy Philox restatement in tests/philox_ref.py. */
int dmx_shot_expectations(const double* d_probs, int64_t rows, int dim, int64_t shots, const int64_t* d_row_shots, const uint32_t* h_masks,
                          int n_masks, uint64_t seed, uint64_t first_row, double* d_expect, void* stream);

/* Same as dmx_energy_batch but with HOST buffers: copies params/variants in, runs, copies energies
 * out, and waits.  This is the call a non-CUDA-aware host (the reference's Python) binds. */
int dmx_energy_batch_host(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants,
                          double* h_energy);

/* Host-buffer call with a coefficient row per circuit (h_group [B] int32, NULL: row 0). */
int dmx_energy_batch_host_grouped(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants,
                                  const int32_t* h_group, double* h_energy);

/* Asynchronous form for callers that keep several batches in flight (a lock-step optimiser driver, a sweep): queues the
 * batch and returns; dmx_plan_host_sync waits for everything queued on the plan.  h_energy / h_params must stay valid and
 * untouched until then.  Only page-locked, mapped buffers on register-resident plans run asynchronously (the kernel then
 * reads / writes host memory itself); every other case executes synchronously inside the call, so the pair
 * async + sync is always correct. */
int dmx_energy_batch_host_async(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants,
                                double* h_energy);
int dmx_plan_host_sync(dmx_plan* plan);

/* ---- measurement helpers ---------------------------------------------------------------------------- */
/* FP64 FMA micro-benchmark on the current device: returns achieved TFLOP/s (roofline denominator the
 * driver's MEASURED_PEAKS.json does not carry).  iters FMAs per thread, 148*k CTAs. */
int dmx_measure_fp64_peak(double* tflops_out, void* stream);
/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t dmx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DMX_H_ */
