/*
 * dm_oracle.c — CPU oracle, C restatement.  TEST INFRASTRUCTURE ONLY: linked/loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg, never by the product.
 *
 * Same algorithm as oracle/dm_oracle.py (which restates cirq 0.8.2's DensityMatrixSimulator, the
 * reference's un-vendored dependency; call sites src/processors/processor_quantum.py:30-39,
 * src/error_mitigation.py:370-383): rho is a dense 2^n x 2^n complex128 matrix, unitary-like ops act as
 * U rho U^dagger, mixtures/channels as sum_k K rho K^dagger, the energy is Re Tr(rho H) summed over Pauli
 * strings.  Nothing is fused or re-expressed: this is the per-operation work the reference's CPU path does,
 * minus Python dispatch.  Parity pinning: validated against dm_oracle.py (tests/test_oracle.py), which is
 * itself pinned on the reference's HF/FCI known answers and on its stored noisy SWAP-test run (statistically);
 * deterministic noisy energies are otherwise unpinned by the reference.
 *
 * The op table uses the same layout as include/dmx.h's dmx_op so the same circuit description can be handed
 * to both implementations; the struct is re-declared here on purpose (no shared code with the product).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef double complex cplx;

/* > 1: the row loops of one circuit are split over this many OpenMP threads (orc_energy_batch_wide: a few LARGE
 * circuits, e.g. 10-12 qubits in the parity tests); 1: serial inside a circuit (orc_energy_batch: one circuit per
 * pthread).  Same arithmetic per element either way - every (row, column) group is written by exactly one thread. */
static int g_wide = 1;

typedef struct {
    int32_t kind, n_qubits, q0, q1, mat, n_mats, param, axis, slot;
    uint32_t flags;
    double scale, offset;
} orc_op;

enum { OP_UNITARY = 1, OP_KRAUS = 2, OP_ROTATION = 3, OP_VARIANT = 4, OP_MEASURE = 5 };
#define OPF_IF_VARIANT_NONZERO 1u

/* out = M applied to the row index bits (left) and conj(M) to the column bits (right): out = M rho M^dagger */
static void conjugate_1q(int n, const cplx* rho, cplx* out, const cplx m[4], int q, int accumulate) {
    const int dim = 1 << n, bit = 1 << (n - 1 - q);
#pragma omp parallel for schedule(static) num_threads(g_wide) if (g_wide > 1)
    for (int r = 0; r < dim; ++r) {
        if (r & bit) continue;
        for (int c = 0; c < dim; ++c) {
            if (c & bit) continue;
            const cplx a00 = rho[(size_t)r * dim + c], a01 = rho[(size_t)r * dim + (c | bit)];
            const cplx a10 = rho[(size_t)(r | bit) * dim + c], a11 = rho[(size_t)(r | bit) * dim + (c | bit)];
            /* t = M a */
            const cplx t00 = m[0] * a00 + m[1] * a10, t01 = m[0] * a01 + m[1] * a11;
            const cplx t10 = m[2] * a00 + m[3] * a10, t11 = m[2] * a01 + m[3] * a11;
            /* o = t M^dagger */
            const cplx o00 = t00 * conj(m[0]) + t01 * conj(m[1]), o01 = t00 * conj(m[2]) + t01 * conj(m[3]);
            const cplx o10 = t10 * conj(m[0]) + t11 * conj(m[1]), o11 = t10 * conj(m[2]) + t11 * conj(m[3]);
            if (accumulate) {
                out[(size_t)r * dim + c] += o00; out[(size_t)r * dim + (c | bit)] += o01;
                out[(size_t)(r | bit) * dim + c] += o10; out[(size_t)(r | bit) * dim + (c | bit)] += o11;
            } else {
                out[(size_t)r * dim + c] = o00; out[(size_t)r * dim + (c | bit)] = o01;
                out[(size_t)(r | bit) * dim + c] = o10; out[(size_t)(r | bit) * dim + (c | bit)] = o11;
            }
        }
    }
}

static void conjugate_2q(int n, const cplx* rho, cplx* out, const cplx m[16], int q0, int q1, int accumulate) {
    const int dim = 1 << n, b0 = 1 << (n - 1 - q0), b1 = 1 << (n - 1 - q1);
    const int off[4] = {0, b1, b0, b0 | b1}; /* local index = 2*bit(q0) + bit(q1) */
#pragma omp parallel for schedule(static) num_threads(g_wide) if (g_wide > 1)
    for (int r = 0; r < dim; ++r) {
        if (r & (b0 | b1)) continue;
        for (int c = 0; c < dim; ++c) {
            if (c & (b0 | b1)) continue;
            cplx a[4][4], t[4][4];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) a[i][j] = rho[(size_t)(r | off[i]) * dim + (c | off[j])];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) {
                    cplx s = 0;
                    for (int l = 0; l < 4; ++l) s += m[i * 4 + l] * a[l][j];
                    t[i][j] = s;
                }
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) {
                    cplx s = 0;
                    for (int l = 0; l < 4; ++l) s += t[i][l] * conj(m[j * 4 + l]);
                    if (accumulate) out[(size_t)(r | off[i]) * dim + (c | off[j])] += s;
                    else out[(size_t)(r | off[i]) * dim + (c | off[j])] = s;
                }
        }
    }
}

static void apply_kraus(int n, cplx** rho, cplx** tmp, const double* pool, int mat, int n_mats, int nq, int q0, int q1) {
    const size_t elems = (size_t)1 << (2 * n);
    const int len = nq == 1 ? 8 : 32;
    if (n_mats == 1) {
        const cplx* m = (const cplx*)(pool + mat);
        if (nq == 1) conjugate_1q(n, *rho, *tmp, m, q0, 0); else conjugate_2q(n, *rho, *tmp, m, q0, q1, 0);
    } else {
        memset(*tmp, 0, elems * sizeof(cplx));
        for (int k = 0; k < n_mats; ++k) {
            const cplx* m = (const cplx*)(pool + mat + (size_t)k * len);
            if (nq == 1) conjugate_1q(n, *rho, *tmp, m, q0, 1); else conjugate_2q(n, *rho, *tmp, m, q0, q1, 1);
        }
    }
    cplx* s = *rho; *rho = *tmp; *tmp = s;
}

static void rotation(int axis, double theta, cplx m[4]) {
    const double c = cos(theta / 2), s = sin(theta / 2);
    if (axis == 0) { m[0] = c; m[1] = -I * s; m[2] = -I * s; m[3] = c; }
    else if (axis == 1) { m[0] = c; m[1] = -s; m[2] = s; m[3] = c; }
    else { m[0] = cexp(-I * (theta / 2)); m[1] = 0; m[2] = 0; m[3] = cexp(I * (theta / 2)); }
}

/* Evolve one circuit; rho/tmp are caller scratch of 4^n cplx each; result in *rho. */
static void evolve(int n, const orc_op* ops, int n_ops, const double* pool, const double* params, const uint8_t* variants,
                   cplx** rho, cplx** tmp) {
    const size_t elems = (size_t)1 << (2 * n);
    memset(*rho, 0, elems * sizeof(cplx));
    (*rho)[0] = 1.0;
    for (int i = 0; i < n_ops; ++i) {
        const orc_op* op = &ops[i];
        if (op->kind == OP_MEASURE) continue;
        if (op->flags & OPF_IF_VARIANT_NONZERO) {
            if (!variants || variants[op->slot] == 0) continue;
        }
        if (op->kind == OP_UNITARY) {
            apply_kraus(n, rho, tmp, pool, op->mat, 1, op->n_qubits, op->q0, op->q1);
        } else if (op->kind == OP_KRAUS) {
            apply_kraus(n, rho, tmp, pool, op->mat, op->n_mats, op->n_qubits, op->q0, op->q1);
        } else if (op->kind == OP_ROTATION) {
            cplx m[4];
            rotation(op->axis, op->scale * params[op->param] + op->offset, m);
            conjugate_1q(n, *rho, *tmp, m, op->q0, 0);
            cplx* s = *rho; *rho = *tmp; *tmp = s;
        } else if (op->kind == OP_VARIANT) {
            /* the basis op is always applied, identity included (src/error_mitigation.py:568-569) */
            const int idx = variants ? variants[op->slot] : 0;
            if (op->n_qubits == 1) {
                const cplx* m = (const cplx*)(pool + op->mat + (size_t)idx * 8);
                conjugate_1q(n, *rho, *tmp, m, op->q0, 0);
            } else {
                const cplx* a = (const cplx*)(pool + op->mat + (size_t)(idx >> 4) * 8);
                const cplx* b = (const cplx*)(pool + op->mat + (size_t)(idx & 15) * 8);
                cplx m[16];
                for (int i0 = 0; i0 < 2; ++i0) for (int i1 = 0; i1 < 2; ++i1)
                    for (int j0 = 0; j0 < 2; ++j0) for (int j1 = 0; j1 < 2; ++j1)
                        m[(i0 * 2 + i1) * 4 + (j0 * 2 + j1)] = a[i0 * 2 + j0] * b[i1 * 2 + j1];
                conjugate_2q(n, *rho, *tmp, m, op->q0, op->q1, 0);
            }
            cplx* s = *rho; *rho = *tmp; *tmp = s;
        }
    }
}

/* Re Tr(rho P) for P given by masks (bit n-1-q <-> qubit q; Y where both set). */
static double pauli_expectation(int n, const cplx* rho, uint32_t x, uint32_t z) {
    const int dim = 1 << n;
    cplx acc = 0;
    const int ny = __builtin_popcount(x & z);
    static const cplx ipow[4] = {1, I, -1, -I};
    for (int i = 0; i < dim; ++i) {
        /* Tr(rho P) = sum_i rho[i, i^x] P[i^x, i];  P[i^x,i] = i^{ny} (-1)^{popc(i & z)} */
        const int sign = __builtin_popcount((uint32_t)i & z) & 1;
        cplx v = rho[(size_t)i * dim + (i ^ (int)x)];
        acc += sign ? -v : v;
    }
    return creal(acc * ipow[ny & 3]);
}

typedef struct {
    int n, n_ops, n_terms, n_params, n_slots;
    const orc_op* ops;
    const double* pool;
    const uint32_t *xm, *zm;
    const double *coeff, *params;
    const uint8_t* variants;
    int64_t batch;
    double* out;
    atomic_llong next;
    atomic_int failed;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    const size_t elems = (size_t)1 << (2 * j->n);
    cplx* rho = (cplx*)malloc(elems * sizeof(cplx));
    cplx* tmp = (cplx*)malloc(elems * sizeof(cplx));
    if (!rho || !tmp) {
        atomic_store(&j->failed, 1);
    } else {
        for (;;) {
            long long b = atomic_fetch_add(&j->next, 1);
            if (b >= j->batch) break;
            evolve(j->n, j->ops, j->n_ops, j->pool, j->params ? j->params + b * j->n_params : NULL,
                   j->variants ? j->variants + b * j->n_slots : NULL, &rho, &tmp);
            double e = 0;
            for (int t = 0; t < j->n_terms; ++t) e += j->coeff[t] * pauli_expectation(j->n, rho, j->xm[t], j->zm[t]);
            j->out[b] = e;
        }
    }
    free(rho); free(tmp);
    return NULL;
}

int orc_max_threads(void) {
    long v = sysconf(_SC_NPROCESSORS_ONLN);
    return v > 0 ? (int)v : 1;
}

/* One circuit per work item, pthread pool with a shared atomic cursor (the reference's own parallelism is a
 * multiprocessing.Pool over circuit variants, src/error_mitigation.py:426-427). */
int orc_energy_batch(int n, const orc_op* ops, int n_ops, const double* pool, int n_terms, const uint32_t* xm, const uint32_t* zm,
                     const double* coeff, const double* params, int n_params, const uint8_t* variants, int n_slots, int64_t batch,
                     double* out, int n_threads) {
    job_t j;
    j.n = n; j.n_ops = n_ops; j.n_terms = n_terms; j.n_params = n_params; j.n_slots = n_slots; j.ops = ops; j.pool = pool;
    j.xm = xm; j.zm = zm; j.coeff = coeff; j.params = params; j.variants = variants; j.batch = batch; j.out = out;
    atomic_init(&j.next, 0);
    atomic_init(&j.failed, 0);
    if (n_threads <= 0) n_threads = orc_max_threads();
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > batch) n_threads = batch > 0 ? (int)batch : 1;
    if (n_threads == 1) {
        worker(&j);
    } else {
        pthread_t th[256];
        int started = 0;
        for (int t = 0; t < n_threads; ++t)
            if (pthread_create(&th[started], NULL, worker, &j) == 0) ++started;
        if (started == 0) worker(&j);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    }
    return atomic_load(&j.failed) ? -1 : 0;
}

/* A few large circuits: evolved one after the other, each with its row loops split over n_threads OpenMP threads. */
int orc_energy_batch_wide(int n, const orc_op* ops, int n_ops, const double* pool, int n_terms, const uint32_t* xm, const uint32_t* zm,
                          const double* coeff, const double* params, int n_params, const uint8_t* variants, int n_slots, int64_t batch,
                          double* out, int n_threads) {
    if (n_threads <= 0) n_threads = orc_max_threads();
    g_wide = n_threads;
    const int rc = orc_energy_batch(n, ops, n_ops, pool, n_terms, xm, zm, coeff, params, n_params, variants, n_slots, batch, out, 1);
    g_wide = 1;
    return rc;
}

/* Single circuit, full density matrix out (row-major complex128). */
int orc_density(int n, const orc_op* ops, int n_ops, const double* pool, const double* params, const uint8_t* variants, double* rho_out) {
    const size_t elems = (size_t)1 << (2 * n);
    cplx* rho = (cplx*)malloc(elems * sizeof(cplx));
    cplx* tmp = (cplx*)malloc(elems * sizeof(cplx));
    if (!rho || !tmp) { free(rho); free(tmp); return -1; }
    evolve(n, ops, n_ops, pool, params, variants, &rho, &tmp);
    memcpy(rho_out, rho, elems * sizeof(cplx));
    free(rho); free(tmp);
    return 0;
}

