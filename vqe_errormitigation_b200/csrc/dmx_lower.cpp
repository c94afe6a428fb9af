// dmx_lower.cpp — host-only part of libdmx: lowers a circuit of gates / Kraus channels / rotations /
// QP-variant slots to real Pauli-transfer blocks, fuses gate+channel blocks, and schedules them either as
//   path 0: one tile pass with the whole state resident in shared memory (n <= 7),
//   path 1: the segment kernel (n <= 4, all fixed blocks monomial: Clifford gates + Pauli channels),
//   path 2: streaming tile passes over a state kept in HBM/L2 (n > 7).
// No CUDA in this file; it is exercised on CPU by tests/test_lowering.py through dmx_plan_dump().
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>

#include "dmx_plan.hpp"

namespace dmx {

using cplx = std::complex<double>;
static const double kTol = 1e-13;

uint32_t pauli_index(int n, uint32_t xmask, uint32_t zmask) {
    uint32_t idx = 0;
    for (int b = 0; b < n; ++b) {
        uint32_t x = (xmask >> b) & 1u, z = (zmask >> b) & 1u;
        idx |= ((x << 1) | z) << (2 * b);
    }
    return idx;
}

// Pauli matrices by code: 0=I 1=Z 2=X 3=Y
static void pauli1(int code, cplx out[4]) {
    static const cplx I(0, 1);
    switch (code) {
        case 0: out[0] = 1; out[1] = 0; out[2] = 0; out[3] = 1; break;
        case 1: out[0] = 1; out[1] = 0; out[2] = 0; out[3] = -1; break;
        case 2: out[0] = 0; out[1] = 1; out[2] = 1; out[3] = 0; break;
        default: out[0] = 0; out[1] = -I; out[2] = I; out[3] = 0; break;
    }
}

static std::vector<cplx> pauli_matrix(int nq, int code) {
    if (nq == 1) {
        std::vector<cplx> m(4);
        pauli1(code, m.data());
        return m;
    }
    cplx a[4], b[4];
    pauli1(code >> 2, a);
    pauli1(code & 3, b);
    std::vector<cplx> m(16);
    for (int i0 = 0; i0 < 2; ++i0)
        for (int i1 = 0; i1 < 2; ++i1)
            for (int j0 = 0; j0 < 2; ++j0)
                for (int j1 = 0; j1 < 2; ++j1)
                    m[(i0 * 2 + i1) * 4 + (j0 * 2 + j1)] = a[i0 * 2 + j0] * b[i1 * 2 + j1];
    return m;
}

// R[a][b] = (1/d) sum_k Tr(P_a K_k P_b K_k^dagger), d = 2^nq.  Real for Hermiticity-preserving maps.
static std::vector<double> ptm_from_kraus(int nq, const std::vector<std::vector<cplx>>& kraus) {
    const int d = 1 << nq, dim = d * d;
    std::vector<double> R(dim * dim, 0.0);
    std::vector<std::vector<cplx>> P(dim);
    for (int a = 0; a < dim; ++a) P[a] = pauli_matrix(nq, a);
    std::vector<cplx> KP(d * d), KPK(d * d);
    for (const auto& K : kraus) {
        for (int b = 0; b < dim; ++b) {
            // KP = K * P_b ; KPK = KP * K^dagger
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    cplx s = 0;
                    for (int l = 0; l < d; ++l) s += K[i * d + l] * P[b][l * d + j];
                    KP[i * d + j] = s;
                }
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    cplx s = 0;
                    for (int l = 0; l < d; ++l) s += KP[i * d + l] * std::conj(K[j * d + l]);
                    KPK[i * d + j] = s;
                }
            for (int a = 0; a < dim; ++a) {
                cplx tr = 0;
                for (int i = 0; i < d; ++i)
                    for (int l = 0; l < d; ++l) tr += P[a][i * d + l] * KPK[l * d + i];
                R[a * dim + b] += tr.real() / d;
            }
        }
    }
    return R;
}

static std::vector<double> identity(int dim) {
    std::vector<double> m(dim * dim, 0.0);
    for (int i = 0; i < dim; ++i) m[i * dim + i] = 1.0;
    return m;
}

static std::vector<double> matmul(const std::vector<double>& a, const std::vector<double>& b, int dim) {
    std::vector<double> c(dim * dim, 0.0);
    for (int i = 0; i < dim; ++i)
        for (int l = 0; l < dim; ++l) {
            double av = a[i * dim + l];
            if (av == 0.0) continue;
            for (int j = 0; j < dim; ++j) c[i * dim + j] += av * b[l * dim + j];
        }
    return c;
}

static std::vector<double> kron4(const std::vector<double>& a, const std::vector<double>& b) {
    std::vector<double> c(256);
    for (int a0 = 0; a0 < 4; ++a0)
        for (int a1 = 0; a1 < 4; ++a1)
            for (int b0 = 0; b0 < 4; ++b0)
                for (int b1 = 0; b1 < 4; ++b1)
                    c[(a0 * 4 + a1) * 16 + (b0 * 4 + b1)] = a[a0 * 4 + b0] * b[a1 * 4 + b1];
    return c;
}

// swap the two qubits of a 16x16 PTM
static std::vector<double> swap_qubits(const std::vector<double>& m) {
    std::vector<double> c(256);
    auto sw = [](int i) { return ((i & 3) << 2) | (i >> 2); };
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 16; ++j) c[sw(i) * 16 + sw(j)] = m[i * 16 + j];
    return c;
}

// classify and snap: entries below tolerance become exact zeros (error <= 1e-13 * |r|, far below the
// 1e-10 parity bar); entries within tolerance of +-1 in monomial matrices become exact.
static int classify(std::vector<double>& m, int dim) {
    bool diag = true, mono = true;
    for (double& v : m)
        if (std::fabs(v) < kTol) v = 0.0;
    std::vector<int> colcount(dim, 0);
    for (int i = 0; i < dim; ++i) {
        int nz = 0;
        for (int j = 0; j < dim; ++j)
            if (m[i * dim + j] != 0.0) {
                ++nz;
                ++colcount[j];
                if (i != j) diag = false;
            }
        if (nz != 1) mono = false;
    }
    for (int j = 0; j < dim; ++j)
        if (colcount[j] != 1) mono = false;
    if (mono)
        for (double& v : m)
            if (v != 0.0 && std::fabs(std::fabs(v) - 1.0) < kTol) v = v > 0 ? 1.0 : -1.0;
    if (diag && mono) {
        bool ident = true;
        for (int i = 0; i < dim; ++i)
            if (m[i * dim + i] != 1.0) ident = false;
        if (ident) return CLS_IDENT;
    }
    if (diag) {
        // a diagonal with zero entries (not a permutation) is still "diag" for execution purposes
        return CLS_DIAG;
    }
    return mono ? CLS_MONO : CLS_DENSE;
}

static bool is_diag(const std::vector<double>& m, int dim) {
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
            if (i != j && std::fabs(m[i * dim + j]) >= kTol) return false;
    return true;
}

// 16x16 form of block-like matrix `m` (on qubits bq[0..nq)) in the ordering (q0,q1)
static std::vector<double> embed16(const std::vector<double>& m, int nq, const int* bq, int q0, int q1) {
    if (nq == 2) {
        if (bq[0] == q0 && bq[1] == q1) return m;
        if (bq[0] == q1 && bq[1] == q0) return swap_qubits(m);
        throw std::runtime_error("embed16: qubit mismatch");
    }
    if (bq[0] == q0) return kron4(m, identity(4));
    if (bq[0] == q1) return kron4(identity(4), m);
    throw std::runtime_error("embed16: qubit mismatch");
}

static std::vector<std::vector<cplx>> read_mats(const Plan& p, const dmx_op& op, int count) {
    const int d = 1 << op.n_qubits, len = 2 * d * d;
    if (op.mat < 0 || (size_t)op.mat + (size_t)count * len > p.pool.size())
        throw std::runtime_error("matrix offset out of range");
    std::vector<std::vector<cplx>> out(count, std::vector<cplx>(d * d));
    for (int k = 0; k < count; ++k)
        for (int i = 0; i < d * d; ++i)
            out[k][i] = cplx(p.pool[op.mat + k * len + 2 * i], p.pool[op.mat + k * len + 2 * i + 1]);
    return out;
}

static std::vector<cplx> rotation_matrix(int axis, double theta) {
    const double c = std::cos(theta / 2), s = std::sin(theta / 2);
    const cplx I(0, 1);
    switch (axis) {
        case 0: return {c, -I * s, -I * s, c};
        case 1: return {c, -s, s, c};
        default: return {std::exp(-I * (theta / 2)), 0, 0, std::exp(I * (theta / 2))};
    }
}

// ---- canonical (complex-rho) cost model of SURVEY.md §8d, evaluated on the source ops ---------------
static bool is_exact_permutation(const std::vector<cplx>& m, int d) {
    for (int i = 0; i < d; ++i) {
        int ones = 0;
        for (int j = 0; j < d; ++j) {
            cplx v = m[i * d + j];
            if (v == cplx(1, 0)) ++ones;
            else if (v != cplx(0, 0)) return false;
        }
        if (ones != 1) return false;
    }
    return true;
}

static bool is_scaled_pauli(const std::vector<cplx>& m, int nq) {
    const int d = 1 << nq, dim = d * d;
    for (int a = 0; a < dim; ++a) {
        auto P = pauli_matrix(nq, a);
        cplx ratio = 0;
        bool ok = true, have = false;
        for (int i = 0; i < d * d && ok; ++i) {
            if (std::abs(P[i]) > 0.5) {
                cplx r = m[i] / P[i];
                if (!have) { ratio = r; have = true; }
                else if (std::abs(r - ratio) > 1e-12) ok = false;
            } else if (std::abs(m[i]) > 1e-12) ok = false;
        }
        if (ok) return true;
    }
    return false;
}

static double canonical_flops_of_op(const Plan& p, const dmx_op& op) {
    const double N = std::pow(4.0, p.n);
    switch (op.kind) {
        case DMX_OP_UNITARY: {
            auto m = read_mats(p, op, 1)[0];
            if (is_exact_permutation(m, 1 << op.n_qubits)) return 0.0;
            return (op.n_qubits == 1 ? 28.0 : 56.0) * N;
        }
        case DMX_OP_ROTATION: return 28.0 * N;
        case DMX_OP_KRAUS: {
            auto ms = read_mats(p, op, op.n_mats);
            bool pauli = true;
            for (auto& m : ms) pauli = pauli && is_scaled_pauli(m, op.n_qubits);
            if (pauli) {
                if (op.n_qubits == 1) return 6.0 * N;
                // 2q: depolarizing (uniform weight on the 15 non-identity Paulis) = 6N, product form = 12N
                std::vector<double> w;
                for (auto& m : ms) {
                    double s = 0;
                    for (auto& v : m) s += std::norm(v);
                    w.push_back(s / 4.0);
                }
                std::sort(w.begin(), w.end());
                bool uniform = ms.size() >= 16 && std::fabs(w[0] - w[14]) < 1e-15;
                return (uniform ? 6.0 : 12.0) * N;
            }
            return ms.size() * (op.n_qubits == 1 ? 30.0 : 62.0) * N;
        }
        case DMX_OP_VARIANT: return 0.0;  // identity correction in the typical variant
        default: return 0.0;
    }
}

// ---- lowering -----------------------------------------------------------------------------------------

static void lower_ops(Plan& p) {
    p.blocks.clear();
    const int n = p.n;
    double canon = 0.0;
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const dmx_op& op = p.ops[i];
        if (op.kind == DMX_OP_MEASURE) continue;
        if (op.n_qubits != 1 && op.n_qubits != 2) throw std::runtime_error("op acts on neither 1 nor 2 qubits");
        if (op.q0 < 0 || op.q0 >= n || (op.n_qubits == 2 && (op.q1 < 0 || op.q1 >= n || op.q1 == op.q0)))
            throw std::runtime_error("qubit index out of range");
        canon += canonical_flops_of_op(p, op);
        Block b;
        b.nq = op.n_qubits;
        b.q[0] = op.q0;
        b.q[1] = op.n_qubits == 2 ? op.q1 : op.q0;
        b.src_op = (int)i;
        const int dim = b.dim();
        if (op.flags & DMX_OPF_IF_VARIANT_NONZERO) {
            if (op.kind != DMX_OP_UNITARY && op.kind != DMX_OP_KRAUS)
                throw std::runtime_error("only UNITARY/KRAUS ops can be conditional");
            if (p.blocks.empty() || p.blocks.back().kind != BK_VAR || p.blocks.back().slot != op.slot)
                throw std::runtime_error("conditional op does not follow the VARIANT op of its slot");
            Block& v = p.blocks.back();
            auto R = ptm_from_kraus(op.n_qubits, read_mats(p, op, op.kind == DMX_OP_UNITARY ? 1 : op.n_mats));
            if (v.nq != b.nq) throw std::runtime_error("conditional op arity differs from its VARIANT op");
            if (v.nq == 2) R = embed16(R, 2, b.q, v.q[0], v.q[1]);
            else if (b.q[0] != v.q[0]) throw std::runtime_error("conditional op on a different qubit");
            v.cond = matmul(R, v.cond, dim);
            continue;
        }
        switch (op.kind) {
            case DMX_OP_UNITARY:
                b.kind = BK_FIXED;
                b.M = ptm_from_kraus(op.n_qubits, read_mats(p, op, 1));
                break;
            case DMX_OP_KRAUS:
                if (op.n_mats < 1) throw std::runtime_error("KRAUS op without matrices");
                b.kind = BK_FIXED;
                b.M = ptm_from_kraus(op.n_qubits, read_mats(p, op, op.n_mats));
                break;
            case DMX_OP_ROTATION: {
                if (op.n_qubits != 1) throw std::runtime_error("ROTATION must be a 1-qubit op");
                if (op.axis < 0 || op.axis > 2) throw std::runtime_error("ROTATION axis must be 0,1,2");
                if (op.param >= p.n_params) throw std::runtime_error("ROTATION param column out of range");
                if (op.param < 0) {  // constant angle: theta = offset
                    b.kind = BK_FIXED;
                    b.M = ptm_from_kraus(1, {rotation_matrix(op.axis, op.offset)});
                    break;
                }
                b.kind = BK_ROT;
                auto P0 = ptm_from_kraus(1, {rotation_matrix(op.axis, 0.0)});
                auto P1 = ptm_from_kraus(1, {rotation_matrix(op.axis, M_PI / 2)});
                auto P2 = ptm_from_kraus(1, {rotation_matrix(op.axis, M_PI)});
                b.A.resize(16); b.C.resize(16); b.S.resize(16);
                for (int e = 0; e < 16; ++e) {
                    b.A[e] = 0.5 * (P0[e] + P2[e]);
                    b.C[e] = 0.5 * (P0[e] - P2[e]);
                    b.S[e] = P1[e] - b.A[e];
                    // the three are exact 0/+-1 patterns; snap away the 1e-17 dust of cos(pi/2)
                    for (std::vector<double>* v : {&b.A, &b.C, &b.S})
                        (*v)[e] = std::fabs((*v)[e]) < kTol ? 0.0 : ((*v)[e] > 0 ? 1.0 : -1.0);
                }
                b.param = op.param;
                b.scale = op.scale;
                b.offset = op.offset;
                break;
            }
            case DMX_OP_VARIANT: {
                if (op.slot < 0 || op.slot >= p.n_slots) throw std::runtime_error("VARIANT slot out of range");
                if (op.n_mats != 16) throw std::runtime_error("VARIANT needs a table of 16 one-qubit matrices");
                b.kind = BK_VAR;
                b.slot = op.slot;
                dmx_op one = op;
                one.n_qubits = 1;
                auto mats = read_mats(p, one, 16);
                b.tab.resize(16 * 16);
                for (int k = 0; k < 16; ++k) {
                    auto R = ptm_from_kraus(1, {mats[k]});
                    for (double& v : R) if (std::fabs(v) < kTol) v = 0.0;
                    std::copy(R.begin(), R.end(), b.tab.begin() + 16 * k);
                }
                b.cond = identity(dim);
                break;
            }
            default: throw std::runtime_error("unknown op kind");
        }
        if (b.kind == BK_FIXED) b.cls = classify(b.M, dim);
        p.blocks.push_back(std::move(b));
    }
    p.info.canonical_flops_per_circuit = canon;
}

static int max_row_nnz(const std::vector<double>& m, int dim) {
    int worst = 0;
    for (int i = 0; i < dim; ++i) {
        int nz = 0;
        for (int j = 0; j < dim; ++j) nz += std::fabs(m[i * dim + j]) >= kTol;
        worst = std::max(worst, nz);
    }
    return worst;
}

// gate+channel fusion: a fixed block is composed into the latest block on its qubits when that block
// covers them; 1-qubit blocks are absorbed into the 2-qubit block that follows; fixed 1-qubit blocks
// around a rotation are folded into its A/C/S matrices.
static void fuse_blocks(Plan& p) {
    const int n = p.n;
    std::vector<int> frontier(n, -1);
    std::vector<Block>& B = p.blocks;
    // Small registers are candidates for the segment kernel, which needs every block to stay "2-sparse" (at most
    // two non-zeros per row: Clifford x Pauli-noise x one Pauli rotation).  Fusing two such blocks can produce a
    // denser matrix (rz . H . rz), so a fusion that would leave the 2-sparse class is skipped there.
    const bool keep_sparse = p.opt_segments && n <= 4;
    auto allowed = [&](const std::vector<double>& fused, int dim, const std::vector<double>& x, int dx, const std::vector<double>& y, int dy) {
        if (!keep_sparse) return true;
        if (max_row_nnz(x, dx) > 2 || max_row_nnz(y, dy) > 2) return true;  // already dense: nothing to protect
        return max_row_nnz(fused, dim) <= 2;
    };
    for (size_t i = 0; i < B.size(); ++i) {
        Block& b = B[i];
        const int dim = b.dim();
        if (b.kind == BK_FIXED) {
            if (b.cls == CLS_IDENT) { b.dead = true; continue; }
            int f0 = frontier[b.q[0]], f1 = b.nq == 2 ? frontier[b.q[1]] : f0;
            if (f0 >= 0 && f0 == f1) {
                Block& f = B[f0];
                bool covers = f.nq == 2 || b.nq == 1;
                if (covers && f.kind == BK_FIXED) {
                    std::vector<double> fused = f.nq == 2 ? matmul(embed16(b.M, b.nq, b.q, f.q[0], f.q[1]), f.M, 16) : matmul(b.M, f.M, 4);
                    if (allowed(fused, f.dim(), f.M, f.dim(), b.M, dim)) {
                        f.M = fused;
                        f.cls = classify(f.M, f.dim());
                        b.dead = true;
                        continue;
                    }
                }
                if (covers && f.kind == BK_ROT && b.nq == 1 && (!keep_sparse || b.cls != CLS_DENSE)) {
                    f.A = matmul(b.M, f.A, 4); f.C = matmul(b.M, f.C, 4); f.S = matmul(b.M, f.S, 4);
                    b.dead = true;
                    continue;
                }
            }
            if (b.nq == 2) {
                // absorb fixed 1-qubit frontier blocks of either qubit
                for (int s = 0; s < 2; ++s) {
                    int f = frontier[b.q[s]];
                    if (f >= 0 && B[f].kind == BK_FIXED && B[f].nq == 1 && !B[f].dead) {
                        std::vector<double> fused = matmul(b.M, embed16(B[f].M, 1, B[f].q, b.q[0], b.q[1]), 16);
                        if (!allowed(fused, 16, b.M, 16, B[f].M, 4)) continue;
                        b.M = fused;
                        b.cls = classify(b.M, 16);
                        B[f].dead = true;
                    }
                }
            }
        } else if (b.kind == BK_ROT) {
            int f = frontier[b.q[0]];
            if (f >= 0 && B[f].kind == BK_FIXED && B[f].nq == 1 && !B[f].dead && (!keep_sparse || B[f].cls != CLS_DENSE)) {
                b.A = matmul(b.A, B[f].M, 4); b.C = matmul(b.C, B[f].M, 4); b.S = matmul(b.S, B[f].M, 4);
                B[f].dead = true;
            }
        }
        (void)dim;
        frontier[b.q[0]] = (int)i;
        if (b.nq == 2) frontier[b.q[1]] = (int)i;
    }
    std::vector<Block> live;
    for (auto& b : B)
        if (!b.dead && !(b.kind == BK_FIXED && b.cls == CLS_IDENT)) live.push_back(std::move(b));
    B.swap(live);
}

static void drop_identities(Plan& p) {
    std::vector<Block> live;
    for (auto& b : p.blocks)
        if (!(b.kind == BK_FIXED && b.cls == CLS_IDENT)) live.push_back(std::move(b));
    p.blocks.swap(live);
}

// ---- device op emission -------------------------------------------------------------------------------

static int push_coefs(Plan& p, const std::vector<double>& v) {
    int off = (int)p.coefs.size();
    p.coefs.insert(p.coefs.end(), v.begin(), v.end());
    return off;
}

// identical tables are stored once (gate+noise blocks repeat many times in ansatz circuits)
struct CoefCache {
    std::map<std::vector<double>, int> seen;
    int get(Plan& p, const std::vector<double>& v) {
        auto it = seen.find(v);
        if (it != seen.end()) return it->second;
        int off = push_coefs(p, v);
        seen.emplace(v, off);
        return off;
    }
};

static void mono_tables(const std::vector<double>& M, int dim, std::vector<double>& scale, uint64_t& packed) {
    scale.assign(dim, 0.0);
    packed = 0;
    const int bits = dim == 4 ? 2 : 4;
    for (int a = 0; a < dim; ++a)
        for (int b = 0; b < dim; ++b)
            if (M[a * dim + b] != 0.0) {
                scale[a] = M[a * dim + b];
                packed |= (uint64_t)b << (bits * a);
            }
}

static bool is_unital3(const std::vector<double>& m) {
    if (m[0] != 1.0) return false;
    for (int j = 1; j < 4; ++j) if (m[j] != 0.0 || m[4 * j] != 0.0) return false;
    return true;
}

static int plan_cpb(int n) {
    if (n >= 6) return 1;
    int cpb = 1;
    while ((cpb << (2 * n)) < 4096 && cpb < 16) cpb <<= 1;
    return cpb;
}

// Builds the sweeps of one pass from its blocks (`order`: block ids in a valid order).  Blocks are re-ordered
// within the pass (they only need to respect the per-qubit program order) so that a sweep over a qubit pair
// collects as many blocks as possible.
struct SweepBuilder {
    Plan& p;
    CoefCache& cc;
    const std::vector<int>& local_pos;
    const std::vector<int>& tile_q;
    int chain_cap;
    Phase phase{};
    double flops = 0.0, sweeps_elems = 0.0;

    SweepBuilder(Plan& plan, CoefCache& cache, const std::vector<int>& lp, const std::vector<int>& tq)
        : p(plan), cc(cache), local_pos(lp), tile_q(tq) {
        chain_cap = std::max(8, 128 / p.cpb);
        phase.first_sweep = (int)p.sweeps.size();
        phase.first_chain = (int)p.chains.size();
    }

    // ---- one open sweep ----
    int qa = -1, qb = -1;
    std::vector<MicroOp> mus;
    std::vector<ChainDev> new_chains;
    std::vector<ChainFactor> new_factors;
    std::vector<int> pendA, pendB;   // pending 1-qubit blocks per side
    SweepDev sw{};
    bool open = false;

    void start(int a, int b) {
        qa = a; qb = b; mus.clear(); new_chains.clear(); new_factors.clear(); pendA.clear(); pendB.clear();
        sw = SweepDev{};
        open = true;
    }

    static MicroOp mu(int kind, int a0 = 0, int a1 = 0, int a2 = 0, int a3 = 0) {
        MicroOp m{};
        m.kind = kind; m.a0 = a0; m.a1 = a1; m.a2 = a2; m.a3 = a3;
        return m;
    }

    void flush_side(std::vector<int>& pend, bool sideA) {
        if (pend.empty()) return;
        // split at variant blocks; runs of FIXED/ROT blocks become one chain (or a fixed matrix)
        size_t i = 0;
        while (i < pend.size()) {
            const Block& b0 = p.blocks[pend[i]];
            if (b0.kind == BK_VAR) {
                std::vector<double> T(16 * 16);
                for (int k = 0; k < 16; ++k) {
                    std::vector<double> Bk(b0.tab.begin() + 16 * k, b0.tab.begin() + 16 * (k + 1));
                    auto t = matmul(b0.cond, Bk, 4);
                    std::copy(t.begin(), t.end(), T.begin() + 16 * k);
                }
                mus.push_back(mu(sideA ? MU_VAR_A : MU_VAR_B, b0.slot, cc.get(p, T)));
                ++i;
                continue;
            }
            size_t j = i;
            bool any_rot = false;
            while (j < pend.size() && p.blocks[pend[j]].kind != BK_VAR) { any_rot |= p.blocks[pend[j]].kind == BK_ROT; ++j; }
            if (!any_rot) {
                std::vector<double> M = identity(4);
                for (size_t t = i; t < j; ++t) M = matmul(p.blocks[pend[t]].M, M, 4);
                int cls = classify(M, 4);
                if (cls == CLS_DIAG || cls == CLS_IDENT) {
                    std::vector<double> dg = {M[0], M[5], M[10], M[15]};
                    if (cls != CLS_IDENT) mus.push_back(mu(sideA ? MU_DIAG_A : MU_DIAG_B, cc.get(p, dg)));
                } else {
                    mus.push_back(mu(sideA ? MU_FIX_A : MU_FIX_B, cc.get(p, M), 0, 0, is_unital3(M)));
                }
            } else {
                ChainDev ch{(int)new_factors.size(), 0};
                bool unital = true;
                for (size_t t = i; t < j; ++t) {
                    const Block& b = p.blocks[pend[t]];
                    ChainFactor f{};
                    if (b.kind == BK_ROT) {
                        std::vector<double> acs;
                        acs.insert(acs.end(), b.A.begin(), b.A.end());
                        acs.insert(acs.end(), b.C.begin(), b.C.end());
                        acs.insert(acs.end(), b.S.begin(), b.S.end());
                        f.kind = 1; f.coef = cc.get(p, acs); f.rot_id = b.rot_id;
                        unital = unital && is_unital3(b.A);
                        for (int e = 0; e < 4; ++e)
                            if (b.C[e] != 0.0 || b.S[e] != 0.0 || b.C[4 * e] != 0.0 || b.S[4 * e] != 0.0) unital = false;
                    } else {
                        f.kind = 0; f.coef = cc.get(p, b.M); f.rot_id = -1;
                        unital = unital && is_unital3(b.M);
                    }
                    new_factors.push_back(f);
                    ++ch.n_factors;
                }
                mus.push_back(mu(sideA ? MU_CHAIN_A : MU_CHAIN_B, (int)new_chains.size(), 0, 0, unital));
                new_chains.push_back(ch);
            }
            i = j;
        }
        pend.clear();
    }

    std::vector<double> oriented(const Block& b) const {   // 16x16 of a 2q block in (qa,qb) order
        return (b.q[0] == qa) ? b.M : swap_qubits(b.M);
    }

    void add_two_qubit(const Block& b) {
        flush_side(pendA, true);
        flush_side(pendB, false);
        if (b.kind == BK_VAR) {
            // variant tables are in the block's own (q0,q1) order: the sweep is oriented accordingly by the caller
            std::vector<double> cond = b.cond;
            int cls = classify(cond, 16);
            MicroOp m = mu(MU_VAR2, b.slot, cc.get(p, b.tab));
            if (cls == CLS_DIAG || cls == CLS_IDENT) {
                std::vector<double> dg(16);
                for (int i = 0; i < 16; ++i) dg[i] = cond[i * 16 + i];
                m.a2 = cc.get(p, dg); m.a3 = CLS_DIAG;
            } else {
                m.a2 = cc.get(p, cond); m.a3 = CLS_DENSE;
            }
            mus.push_back(m);
            return;
        }
        std::vector<double> M = oriented(b);
        int cls = classify(M, 16);
        if (cls == CLS_DIAG || cls == CLS_IDENT) {
            std::vector<double> dg(16);
            for (int i = 0; i < 16; ++i) dg[i] = M[i * 16 + i];
            mus.push_back(mu(MU_DIAG2, cc.get(p, dg)));
        } else {
            mus.push_back(mu(MU_MAT2, cc.get(p, M)));
        }
    }

    static uint32_t code_offset(int code, int p0, int p1) { return ((uint32_t)(code >> 2) << p0) | ((uint32_t)(code & 3) << p1); }

    // monomial 16x16: out[a] = sc[a] * in[src[a]]; usable as an addressing permutation only when src is GF(2)-linear
    static bool linear_perm(const std::vector<double>& M, int src[16], double sc[16]) {
        for (int a = 0; a < 16; ++a) {
            src[a] = -1;
            for (int c = 0; c < 16; ++c)
                if (M[a * 16 + c] != 0.0) { src[a] = c; sc[a] = M[a * 16 + c]; }
            if (src[a] < 0) return false;
        }
        for (int a = 0; a < 16; ++a)
            for (int c = 0; c < 16; ++c)
                if (src[a ^ c] != (src[a] ^ src[c])) return false;
        return true;
    }

    void set_perm(const std::vector<double>& M, bool at_load) {
        int src[16];
        double sc[16];
        linear_perm(M, src, sc);
        const int p0 = local_pos[qa], p1 = local_pos[qb];
        if (at_load) {
            // v[e] = sc[e] * tile[code src[e]]
            sw.load_perm = 1;
            sw.lcoef = cc.get(p, std::vector<double>(sc, sc + 16));
            for (int j = 0; j < 4; ++j) sw.lmask[j] = code_offset(src[1 << j], p0, p1);
        } else {
            // out[a] = sc[a] * v[src[a]]  <=>  register e goes to code dst[e] = src^-1[e] with scale sc[dst[e]]
            int dst[16];
            for (int a = 0; a < 16; ++a) dst[src[a]] = a;
            std::vector<double> se(16);
            for (int e = 0; e < 16; ++e) se[e] = sc[dst[e]];
            sw.store_perm = 1;
            sw.scoef = cc.get(p, se);
            for (int j = 0; j < 4; ++j) sw.smask[j] = code_offset(dst[1 << j], p0, p1);
        }
    }

    void close() {
        if (!open) return;
        flush_side(pendA, true);
        flush_side(pendB, false);
        open = false;
        if (mus.empty() && !sw.load_perm && !sw.store_perm) return;
        if ((int)(p.chains.size() - phase.first_chain + new_chains.size()) > chain_cap && p.sweeps.size() > (size_t)phase.first_sweep) end_phase();
        const int chain_base = (int)p.chains.size() - phase.first_chain;
        const int factor_base = (int)p.factors.size();
        for (ChainDev ch : new_chains) { ch.first_factor += factor_base; p.chains.push_back(ch); }
        p.factors.insert(p.factors.end(), new_factors.begin(), new_factors.end());
        for (MicroOp& m : mus)
            if (m.kind == MU_CHAIN_A || m.kind == MU_CHAIN_B) m.a0 += chain_base;
        sw.p0 = local_pos[qa]; sw.p1 = local_pos[qb];
        for (int j = 0; j < 4; ++j) {
            const uint32_t unit = code_offset(1 << j, sw.p0, sw.p1);
            if (!sw.load_perm) sw.lmask[j] = unit;
            if (!sw.store_perm) sw.smask[j] = unit;
        }
        sw.first_mu = (int)p.mus.size(); sw.n_mu = (int)mus.size();
        p.mus.insert(p.mus.end(), mus.begin(), mus.end());
        p.sweeps.push_back(sw);
        // cost model (per element of the state): one smem read + write; flops by micro-op
        sweeps_elems += 1.0;
        for (const MicroOp& m : mus) {
            switch (m.kind) {
                case MU_CHAIN_A: case MU_CHAIN_B: case MU_FIX_A: case MU_FIX_B: flops += m.a3 ? 15.0 / 4 : 28.0 / 4; break;
                case MU_DIAG_A: case MU_DIAG_B: case MU_DIAG2: flops += 1.0; break;
                case MU_MAT2: flops += 31.0; break;
                default: break;
            }
        }
        flops += (sw.load_perm ? 1.0 : 0.0) + (sw.store_perm ? 1.0 : 0.0);
    }

    void end_phase() {
        phase.n_sweeps = (int)p.sweeps.size() - phase.first_sweep;
        phase.n_chains = (int)p.chains.size() - phase.first_chain;
        if (phase.n_sweeps > 0) p.phases.push_back(phase);
        phase.first_sweep = (int)p.sweeps.size();
        phase.first_chain = (int)p.chains.size();
    }

    void run(const std::vector<int>& order) {
        const int n = std::max(p.n, 2);
        std::vector<std::vector<int>> qq(n);
        std::vector<size_t> head(n, 0);
        for (int bi : order) {
            qq[p.blocks[bi].q[0]].push_back(bi);
            if (p.blocks[bi].nq == 2) qq[p.blocks[bi].q[1]].push_back(bi);
        }
        auto front = [&](int q) { return head[q] < qq[q].size() ? qq[q][head[q]] : -1; };
        auto ready = [&](int bi) {
            const Block& b = p.blocks[bi];
            for (int s = 0; s < b.nq; ++s) if (front(b.q[s]) != bi) return false;
            return true;
        };
        auto pop = [&](int bi) {
            const Block& b = p.blocks[bi];
            for (int s = 0; s < b.nq; ++s) ++head[b.q[s]];
        };
        size_t remaining = order.size();
        while (remaining > 0) {
            // earliest ready block opens the sweep
            int b0 = -1;
            for (int q = 0; q < n; ++q) { int f = front(q); if (f >= 0 && ready(f) && (b0 < 0 || f < b0)) b0 = f; }
            if (b0 < 0) throw std::runtime_error("sweep builder: no ready block");
            const Block& first = p.blocks[b0];
            int a = first.q[0], b = first.nq == 2 ? first.q[1] : -1;
            if (b < 0) {
                // partner: the other qubit of the next two-qubit block on this qubit's queue (if resident), else a neighbour
                for (size_t t = head[a]; t < qq[a].size() && b < 0; ++t) {
                    const Block& nb = p.blocks[qq[a][t]];
                    if (nb.nq == 2) b = nb.q[0] == a ? nb.q[1] : nb.q[0];
                }
                if (b < 0)
                    for (int q : tile_q) if (q != a) { b = q; if (q > a) break; }
                if (b < 0) throw std::runtime_error("sweep builder: single-qubit tile");
            }
            start(a, b);
            bool closed = false;
            while (!closed) {
                bool progress = false;
                // 1) every ready one-qubit block on either side
                for (int side = 0; side < 2; ++side) {
                    int q = side == 0 ? qa : qb;
                    while (true) {
                        int f = front(q);
                        if (f < 0 || p.blocks[f].nq != 1) break;
                        (side == 0 ? pendA : pendB).push_back(f);
                        pop(f); --remaining; progress = true;
                    }
                }
                // 2) a ready two-qubit block on exactly this pair
                int f = front(qa);
                if (f >= 0 && p.blocks[f].nq == 2 && ready(f)) {
                    const Block& blk = p.blocks[f];
                    bool same_pair = (blk.q[0] == qa && blk.q[1] == qb) || (blk.q[0] == qb && blk.q[1] == qa);
                    if (same_pair) {
                        bool mono = blk.kind == BK_FIXED && blk.cls == CLS_MONO;
                        std::vector<double> Mo;
                        if (mono) {
                            Mo = oriented(blk);
                            classify(Mo, 16);
                            int src_[16];
                            double sc_[16];
                            mono = linear_perm(Mo, src_, sc_);   // non-Clifford monomials fall back to a dense micro-op
                        }
                        if (blk.kind == BK_VAR && blk.q[0] != qa) {
                            // variant tables are oriented (q0,q1): only usable when the sweep has the same orientation
                            if (mus.empty() && pendA.empty() && pendB.empty() && !sw.load_perm) { std::swap(qa, qb); }
                            else { close(); closed = true; continue; }
                        }
                        if (mono) {
                            bool empty = mus.empty() && pendA.empty() && pendB.empty() && !sw.load_perm;
                            if (empty) { set_perm(Mo, true); pop(f); --remaining; progress = true; }
                            else { flush_side(pendA, true); flush_side(pendB, false); set_perm(Mo, false); pop(f); --remaining; close(); closed = true; continue; }
                        } else {
                            add_two_qubit(blk);
                            pop(f); --remaining; progress = true;
                        }
                    }
                }
                if (!progress) { close(); closed = true; }
            }
        }
        close();
        end_phase();
    }
};

// Triple-sweep builder (see Sweep3 in dmx_plan.hpp).  A sweep is  [monomial blocks]* [general blocks]* [monomial blocks]*
// on three tile qubits: the leading / trailing monomials are composed into the load / store permutation, everything
// between runs on the 64 register-resident coefficients.  Only FIXED and ROT blocks (no variant slots).
struct TripleBuilder {
    Plan& p;
    CoefCache& cc;
    const std::vector<int>& local_pos;
    const std::vector<int>& tile_q;
    Phase phase{};
    double flops = 0.0, sweeps_elems = 0.0;

    TripleBuilder(Plan& plan, CoefCache& cache, const std::vector<int>& lp, const std::vector<int>& tq)
        : p(plan), cc(cache), local_pos(lp), tile_q(tq) {
        phase.first_sweep = (int)p.sweeps3.size();
        phase.first_chain = (int)p.chains.size();
    }

    struct Mono {            // out[a] = sc[a] * in[src[a]] on the 64 codes of the triple
        int src[64];
        double sc[64];
        bool trivial = true;
        Mono() { for (int a = 0; a < 64; ++a) { src[a] = a; sc[a] = 1.0; } }
        void then(const Mono& m) {      // apply m after this
            int ns[64];
            double nc[64];
            for (int a = 0; a < 64; ++a) { ns[a] = src[m.src[a]]; nc[a] = m.sc[a] * sc[m.src[a]]; }
            for (int a = 0; a < 64; ++a) { src[a] = ns[a]; sc[a] = nc[a]; }
            trivial = false;
        }
    };

    int T[3] = {-1, -1, -1};
    std::vector<std::array<int, 3>>* record = nullptr;   // dry runs: the qubit triple of every emitted sweep
    int slot_of(int q) const { for (int s = 0; s < 3; ++s) if (T[s] == q) return s; return -1; }
    static int shift(int slot) { return 4 - 2 * slot; }

    // block -> monomial on the 64 codes (false when the block is not a GF(2)-linear signed permutation x scales)
    bool as_mono(const Block& b, Mono& out) const {
        if (b.kind != BK_FIXED || (b.cls != CLS_MONO && b.cls != CLS_DIAG && b.cls != CLS_IDENT)) return false;
        const int dim = b.dim();
        std::vector<int> src(dim, -1);
        std::vector<double> sc(dim, 0.0);
        for (int a = 0; a < dim; ++a) {
            for (int c = 0; c < dim; ++c)
                if (b.M[a * dim + c] != 0.0) { if (src[a] >= 0) return false; src[a] = c; sc[a] = b.M[a * dim + c]; }
            if (src[a] < 0) return false;
        }
        for (int a = 0; a < dim; ++a)
            for (int c = 0; c < dim; ++c)
                if (src[a ^ c] != (src[a] ^ src[c])) return false;
        out = Mono();
        out.trivial = false;
        if (b.nq == 1) {
            const int sh = shift(slot_of(b.q[0]));
            for (int a = 0; a < 64; ++a) {
                const int c = (a >> sh) & 3;
                out.src[a] = (a & ~(3 << sh)) | (src[c] << sh);
                out.sc[a] = sc[c];
            }
        } else {
            const int s0 = shift(slot_of(b.q[0])), s1 = shift(slot_of(b.q[1]));
            for (int a = 0; a < 64; ++a) {
                const int c = (((a >> s0) & 3) << 2) | ((a >> s1) & 3);
                const int sv = src[c];
                out.src[a] = (a & ~(3 << s0) & ~(3 << s1)) | ((sv >> 2) << s0) | ((sv & 3) << s1);
                out.sc[a] = sc[c];
            }
        }
        return true;
    }

    static MicroOp mu(int kind, int a0 = 0, int a1 = 0, int a2 = 0, int a3 = 0) {
        MicroOp m{};
        m.kind = kind; m.a0 = a0; m.a1 = a1; m.a2 = a2; m.a3 = a3;
        return m;
    }

    std::vector<MicroOp> mus;
    std::vector<ChainDev> new_chains;
    std::vector<ChainFactor> new_factors;
    std::vector<int> pend[3];

    void flush_slot(int slot) {
        std::vector<int>& pd = pend[slot];
        if (pd.empty()) return;
        bool any_rot = false;
        for (int bi : pd) any_rot |= p.blocks[bi].kind == BK_ROT;
        if (!any_rot) {
            std::vector<double> M = identity(4);
            for (int bi : pd) M = matmul(p.blocks[bi].M, M, 4);
            const int cls = classify(M, 4);
            if (cls == CLS_DIAG || cls == CLS_IDENT) {
                if (cls != CLS_IDENT) { mus.push_back(mu(MU3_DIAG, cc.get(p, {M[0], M[5], M[10], M[15]}), slot)); flops += 1.0; }
            } else {
                const bool un = is_unital3(M);
                mus.push_back(mu(MU3_MAT, cc.get(p, M), slot, 0, un));
                flops += un ? 15.0 / 4 : 28.0 / 4;
            }
        } else {
            ChainDev ch{(int)new_factors.size(), 0};
            bool unital = true;
            for (int bi : pd) {
                const Block& b = p.blocks[bi];
                ChainFactor f{};
                if (b.kind == BK_ROT) {
                    std::vector<double> acs;
                    acs.insert(acs.end(), b.A.begin(), b.A.end());
                    acs.insert(acs.end(), b.C.begin(), b.C.end());
                    acs.insert(acs.end(), b.S.begin(), b.S.end());
                    f.kind = 1; f.coef = cc.get(p, acs); f.rot_id = b.rot_id;
                    unital = unital && is_unital3(b.A);
                    for (int e = 0; e < 4; ++e)
                        if (b.C[e] != 0.0 || b.S[e] != 0.0 || b.C[4 * e] != 0.0 || b.S[4 * e] != 0.0) unital = false;
                } else {
                    f.kind = 0; f.coef = cc.get(p, b.M); f.rot_id = -1;
                    unital = unital && is_unital3(b.M);
                }
                new_factors.push_back(f);
                ++ch.n_factors;
            }
            mus.push_back(mu(MU3_MAT, (int)new_chains.size(), slot, 1, unital));
            new_chains.push_back(ch);
            flops += unital ? 15.0 / 4 : 28.0 / 4;
        }
        pd.clear();
    }

    uint32_t code_offset(int code) const {
        return ((uint32_t)((code >> 4) & 3) << local_pos[T[0]]) | ((uint32_t)((code >> 2) & 3) << local_pos[T[1]]) |
               ((uint32_t)(code & 3) << local_pos[T[2]]);
    }

    void emit(const Mono& lp, const Mono& sp) {
        for (int s = 0; s < 3; ++s) flush_slot(s);
        if (mus.empty() && lp.trivial && sp.trivial) return;
        Sweep3 sw{};
        for (int s = 0; s < 3; ++s) sw.p[s] = local_pos[T[s]];
        const int chain_base = (int)p.chains.size() - phase.first_chain;
        const int factor_base = (int)p.factors.size();
        for (ChainDev ch : new_chains) { ch.first_factor += factor_base; p.chains.push_back(ch); }
        p.factors.insert(p.factors.end(), new_factors.begin(), new_factors.end());
        for (MicroOp& m : mus)
            if (m.kind == MU3_MAT && m.a2) m.a0 += chain_base;
        for (int j = 0; j < 6; ++j) { sw.lmask[j] = code_offset(1 << j); sw.smask[j] = code_offset(1 << j); }
        if (!lp.trivial) {
            sw.load_perm = 1;
            sw.lcoef = cc.get(p, std::vector<double>(lp.sc, lp.sc + 64));
            for (int j = 0; j < 6; ++j) sw.lmask[j] = code_offset(lp.src[1 << j]);
            flops += 1.0;
        }
        if (!sp.trivial) {
            // out[a] = sc[a] * v[src[a]]  <=>  register e goes to code dst[e] = src^-1[e] with scale sc[dst[e]]
            int dst[64];
            for (int a = 0; a < 64; ++a) dst[sp.src[a]] = a;
            std::vector<double> se(64);
            for (int e = 0; e < 64; ++e) se[e] = sp.sc[dst[e]];
            sw.store_perm = 1;
            sw.scoef = cc.get(p, se);
            for (int j = 0; j < 6; ++j) sw.smask[j] = code_offset(dst[1 << j]);
            flops += 1.0;
        }
        sw.first_mu = (int)p.mus.size(); sw.n_mu = (int)mus.size();
        p.mus.insert(p.mus.end(), mus.begin(), mus.end());
        p.sweeps3.push_back(sw);
        if (record) record->push_back({T[0], T[1], T[2]});
        sweeps_elems += 1.0;
        mus.clear(); new_chains.clear(); new_factors.clear();
    }

    // ---- tile-wide monomials (plan option "wide") --------------------------------------------------------------------
    struct MonoT {           // out[a] = sc[a] * in[src[a]] over the 4^k tile-local indices
        std::vector<uint32_t> src;
        std::vector<double> sc;
        bool trivial = true;
        explicit MonoT(int E = 0) : src(E), sc(E, 1.0) { for (int a = 0; a < E; ++a) src[a] = (uint32_t)a; }
        void then(const MonoT& m) {      // apply m after this
            const size_t E = src.size();
            std::vector<uint32_t> ns(E);
            std::vector<double> nc(E);
            for (size_t a = 0; a < E; ++a) { ns[a] = src[m.src[a]]; nc[a] = m.sc[a] * sc[m.src[a]]; }
            src.swap(ns); sc.swap(nc);
            trivial = false;
        }
        void swap_with(MonoT& o) { src.swap(o.src); sc.swap(o.sc); std::swap(trivial, o.trivial); }
    };

    // block -> monomial over the whole tile (any resident qubits)
    bool as_mono_tile(const Block& b, MonoT& out) const {
        if (b.kind != BK_FIXED || (b.cls != CLS_MONO && b.cls != CLS_DIAG && b.cls != CLS_IDENT)) return false;
        const int dim = b.dim();
        std::vector<int> src(dim, -1);
        std::vector<double> sc(dim, 0.0);
        for (int a = 0; a < dim; ++a) {
            for (int c = 0; c < dim; ++c)
                if (b.M[a * dim + c] != 0.0) { if (src[a] >= 0) return false; src[a] = c; sc[a] = b.M[a * dim + c]; }
            if (src[a] < 0) return false;
        }
        for (int a = 0; a < dim; ++a)
            for (int c = 0; c < dim; ++c)
                if (src[a ^ c] != (src[a] ^ src[c])) return false;
        const int E = 1 << (2 * (int)tile_q.size());
        out = MonoT(E);
        out.trivial = false;
        if (b.nq == 1) {
            const int sh = local_pos[b.q[0]];
            for (int a = 0; a < E; ++a) {
                const int c = (a >> sh) & 3;
                out.src[a] = (uint32_t)((a & ~(3 << sh)) | (src[c] << sh));
                out.sc[a] = sc[c];
            }
        } else {
            const int s0 = local_pos[b.q[0]], s1 = local_pos[b.q[1]];
            for (int a = 0; a < E; ++a) {
                const int c = (((a >> s0) & 3) << 2) | ((a >> s1) & 3);
                const int sv = src[c];
                out.src[a] = (uint32_t)((a & ~(3 << s0) & ~(3 << s1)) | ((sv >> 2) << s0) | ((sv & 3) << s1));
                out.sc[a] = sc[c];
            }
        }
        return true;
    }

    // canonical tile index of (thread index t, register code e) for the current triple
    uint32_t canon(uint32_t t, int e) const {
        int ps[3] = {local_pos[T[0]], local_pos[T[1]], local_pos[T[2]]};
        std::sort(ps, ps + 3);
        uint32_t a = t;
        for (int j = 0; j < 3; ++j) a = ((a >> ps[j]) << (ps[j] + 2)) | (a & ((1u << ps[j]) - 1u));
        return a | code_offset(e);
    }

    struct WideSide {
        bool wide = false;        // needs the tmask / variant form (else the legacy triple-local form is exact)
        int nvar = 0;
        uint32_t tmask[8] = {}, rmask[6] = {}, vmask[4] = {};
        std::vector<double> table;    // [2^nvar][64]
        bool all_one = true;          // every scale is exactly 1
    };

    // Analyse one side of a sweep.  load: v[e] = sc[a] * tile[src[a]] with a = canon(t, e); store: the element of a goes to
    // dst[a] = src^-1[a] with scale sc[dst[a]].  False when the scales need more than 16 thread-dependent variants or the
    // thread classes are not the cosets of a subspace.
    bool analyse(const MonoT& m, bool store, WideSide& out) const {
        const int k = (int)tile_q.size(), E = 1 << (2 * k), NTB = 2 * (k - 3), NT = 1 << NTB;
        std::vector<uint32_t> dst;
        if (store) { dst.resize(E); for (int a = 0; a < E; ++a) dst[m.src[a]] = (uint32_t)a; }
        auto image = [&](uint32_t a) { return store ? dst[a] : m.src[a]; };
        auto scale = [&](uint32_t a) { return store ? m.sc[dst[a]] : m.sc[a]; };
        out = WideSide();
        for (int j = 0; j < NTB; ++j) out.tmask[j] = image(canon(1u << j, 0));
        for (int j = 0; j < 6; ++j) out.rmask[j] = image(canon(0, 1 << j));
        uint32_t triple_bits = 0;
        for (int s2 = 0; s2 < 3; ++s2) triple_bits |= 3u << local_pos[T[s2]];
        bool identity_threads = true;
        for (int j = 0; j < NTB; ++j) identity_threads = identity_threads && out.tmask[j] == canon(1u << j, 0);
        bool local_regs = true;
        for (int j = 0; j < 6; ++j) local_regs = local_regs && !(out.rmask[j] & ~triple_bits);
        // thread classes by scale vector
        std::map<std::vector<double>, int> cls;
        std::vector<int> cls_of(NT);
        std::vector<std::vector<double>> vecs;
        for (int t = 0; t < NT; ++t) {
            std::vector<double> v(64);
            for (int e = 0; e < 64; ++e) { v[e] = scale(canon((uint32_t)t, e)); out.all_one = out.all_one && v[e] == 1.0; }
            auto it = cls.find(v);
            if (it == cls.end()) { it = cls.emplace(v, (int)vecs.size()).first; vecs.push_back(v); }
            cls_of[t] = it->second;
        }
        const int V = (int)vecs.size();
        if (V > 16 || (V & (V - 1))) return false;
        int nvar = 0;
        while ((1 << nvar) < V) ++nvar;
        if (V > 1) {
            // the threads of class(0) must form a subspace K and the classes its cosets; variant bits = functionals vanishing on K
            std::vector<uint32_t> K;
            for (int t = 0; t < NT; ++t) if (cls_of[t] == cls_of[0]) K.push_back((uint32_t)t);
            if ((int)K.size() * V != NT) return false;
            for (int t = 0; t < NT; ++t)
                for (uint32_t kk : K) if (cls_of[t ^ kk] != cls_of[t]) return false;
            std::vector<unsigned> chosen;
            for (uint32_t f = 1; f < (uint32_t)NT && (int)chosen.size() < nvar; ++f) {
                bool ok = true;
                for (uint32_t kk : K) if (__builtin_popcount(f & kk) & 1) { ok = false; break; }
                if (!ok) continue;
                std::vector<unsigned> trial = chosen;
                trial.push_back(f);
                if (gf2_rank_small(trial) == (int)trial.size()) chosen.push_back(f);
            }
            if ((int)chosen.size() != nvar) return false;
            for (int j = 0; j < nvar; ++j) out.vmask[j] = chosen[j];
        }
        out.nvar = nvar;
        out.table.assign((size_t)V * 64, 1.0);
        std::vector<char> filled(V, 0);
        for (int t = 0; t < NT; ++t) {
            int var = 0;
            for (int j = 0; j < nvar; ++j) var |= (__builtin_popcount((uint32_t)t & out.vmask[j]) & 1) << j;
            if (!filled[var]) { std::copy(vecs[cls_of[t]].begin(), vecs[cls_of[t]].end(), out.table.begin() + (size_t)var * 64); filled[var] = 1; }
            else for (int e = 0; e < 64; ++e) if (out.table[(size_t)var * 64 + e] != vecs[cls_of[t]][e]) return false;   // functionals do not separate
        }
        out.wide = !(identity_threads && local_regs && V == 1);
        return true;
    }

    static int gf2_rank_small(std::vector<unsigned> v) {
        int rank = 0;
        for (int bit = 0; bit < 16; ++bit) {
            int piv = -1;
            for (size_t i = rank; i < v.size(); ++i) if ((v[i] >> bit) & 1u) { piv = (int)i; break; }
            if (piv < 0) continue;
            std::swap(v[rank], v[piv]);
            for (size_t i = 0; i < v.size(); ++i) if ((int)i != rank && ((v[i] >> bit) & 1u)) v[i] ^= v[rank];
            ++rank;
        }
        return rank;
    }

    size_t wide_doubles = 0;      // variant tables + descriptors of this pass (bounded: they live in shared memory)

    // shared-memory bytes the pass needs so far beside the tile (chain matrices + every coefficient table its sweeps and
    // micro-ops name), counted like pass_stream_eligible does
    size_t pass_bytes() const {
        std::map<int, int> tables;
        size_t extra = 0;
        auto mu_table = [&](const MicroOp& m) {
            switch (m.kind) {
                case MU3_MAT: if (!m.a2) tables[m.a0] = 16; break;
                case MU3_DIAG: tables[m.a0] = 4; break;
                case MU3_DIAG2: tables[m.a0] = 16; break;
                case MU3_MAT2: tables[m.a0] = 256; break;
                default: break;
            }
        };
        for (size_t si = (size_t)phase.first_sweep; si < p.sweeps3.size(); ++si) {
            const Sweep3& sw = p.sweeps3[si];
            if (sw.load_perm) tables[sw.lcoef] = std::max(tables[sw.lcoef], 64 << ((sw.wide & 1) ? sw.l_nvar : 0));
            if (sw.store_perm) tables[sw.scoef] = std::max(tables[sw.scoef], 64 << ((sw.wide & 2) ? sw.s_nvar : 0));
            if (sw.wide) extra += 16;
            for (int mi = 0; mi < sw.n_mu; ++mi) mu_table(p.mus[sw.first_mu + mi]);
        }
        for (const MicroOp& m : mus) mu_table(m);
        size_t coefs = extra;
        for (const auto& t : tables) coefs += (size_t)((t.second + 1) & ~1);
        const size_t chains = p.chains.size() - (size_t)phase.first_chain + new_chains.size() + 3;     // + the slots still pending
        return chains * 128 + coefs * 8;
    }

    void emit_wide(const WideSide& ls, const WideSide& ss, bool lead_trivial, bool trail_trivial) {
        for (int s2 = 0; s2 < 3; ++s2) flush_slot(s2);
        if (mus.empty() && lead_trivial && trail_trivial) return;
        Sweep3 sw{};
        for (int s2 = 0; s2 < 3; ++s2) sw.p[s2] = local_pos[T[s2]];
        const int chain_base = (int)p.chains.size() - phase.first_chain;
        const int factor_base = (int)p.factors.size();
        for (ChainDev ch : new_chains) { ch.first_factor += factor_base; p.chains.push_back(ch); }
        p.factors.insert(p.factors.end(), new_factors.begin(), new_factors.end());
        for (MicroOp& m : mus)
            if (m.kind == MU3_MAT && m.a2) m.a0 += chain_base;
        for (int j = 0; j < 6; ++j) { sw.lmask[j] = code_offset(1 << j); sw.smask[j] = code_offset(1 << j); }
        const int NTB = 2 * ((int)tile_q.size() - 3);
        if (!lead_trivial) {
            for (int j = 0; j < 6; ++j) sw.lmask[j] = ls.rmask[j];
            if (!ls.all_one || ls.wide) {
                sw.load_perm = 1;
                sw.lcoef = cc.get(p, ls.table);
                flops += 1.0;
            }
            if (ls.wide) {
                sw.wide |= 1 | 4;
                sw.l_nvar = ls.nvar;
                for (int j = 0; j < NTB; ++j) sw.ltmask[j] = ls.tmask[j];
                for (int j = 0; j < ls.nvar; ++j) sw.lvmask[j] = ls.vmask[j];
                wide_doubles += ls.table.size();
            } else if (!sw.load_perm) {
                sw.load_perm = 1;      // pure permutation: the kernel takes the mask path only when load_perm is set
                sw.lcoef = cc.get(p, ls.table);
            }
        }
        if (!trail_trivial) {
            for (int j = 0; j < 6; ++j) sw.smask[j] = ss.rmask[j];
            sw.store_perm = 1;
            sw.scoef = cc.get(p, ss.table);
            flops += 1.0;
            if (ss.wide) {
                sw.wide |= 2 | 4;
                sw.s_nvar = ss.nvar;
                for (int j = 0; j < NTB; ++j) sw.stmask[j] = ss.tmask[j];
                for (int j = 0; j < ss.nvar; ++j) sw.svmask[j] = ss.vmask[j];
                wide_doubles += ss.table.size();
            }
        }
        if (sw.wide) wide_doubles += 16;
        sw.first_mu = (int)p.mus.size(); sw.n_mu = (int)mus.size();
        p.mus.insert(p.mus.end(), mus.begin(), mus.end());
        p.sweeps3.push_back(sw);
        if (record) record->push_back({T[0], T[1], T[2]});
        sweeps_elems += 1.0;
        mus.clear(); new_chains.clear(); new_factors.clear();
    }

    // Sweep builder with tile-wide leading / trailing monomials.  The triple is chosen for the DENSE work: the builder looks
    // through the monomial blocks that are ready (they cost no register residency any more) and takes the qubits of the
    // earliest dense blocks behind them.
    void run_wide(const std::vector<int>& order) {
        const int n = p.n, k = (int)tile_q.size(), E = 1 << (2 * k);
        // the pass's tables share the CTA's shared memory with the tile (pass_stream_eligible: 20 KB for 6-qubit tiles,
        // 90 KB for 7); a wide side is only taken while its variant table still fits, with room left for the rest of the pass
        const size_t PASS_LIMIT = (k == 7 ? 90u : 20u) * 1024 - 3072;
        std::vector<std::vector<int>> qq(n);
        std::vector<size_t> head(n, 0);
        std::map<int, MonoT> mono_of;              // blocks of this pass that are tile-wide monomials
        for (int bi : order) {
            const Block& b = p.blocks[bi];
            if (b.kind == BK_VAR) throw std::runtime_error("triple sweeps do not take variant blocks");
            qq[b.q[0]].push_back(bi);
            if (b.nq == 2) qq[b.q[1]].push_back(bi);
            MonoT m;
            if (as_mono_tile(b, m)) mono_of.emplace(bi, std::move(m));
        }
        auto front_of = [&](const std::vector<size_t>& h, int q) { return h[q] < qq[q].size() ? qq[q][h[q]] : -1; };
        auto ready_in = [&](const std::vector<size_t>& h, int bi) {
            const Block& b = p.blocks[bi];
            for (int s2 = 0; s2 < b.nq; ++s2) if (front_of(h, b.q[s2]) != bi) return false;
            return true;
        };
        auto pop_in = [&](std::vector<size_t>& h, int bi) {
            const Block& b = p.blocks[bi];
            for (int s2 = 0; s2 < b.nq; ++s2) ++h[b.q[s2]];
        };
        auto earliest_ready = [&](const std::vector<size_t>& h, bool want_mono, bool want_dense, const std::vector<char>* skip) {
            int best = -1;
            for (int q : tile_q) {
                const int f = front_of(h, q);
                if (f < 0 || !ready_in(h, f) || (best >= 0 && f >= best)) continue;
                (void)skip;
                const bool mono = mono_of.count(f) != 0;
                if ((mono && !want_mono) || (!mono && !want_dense)) continue;
                best = f;
            }
            return best;
        };
        size_t remaining = order.size();
        while (remaining > 0) {
            // ---- triple: look through the ready monomials, take the qubits of the earliest dense blocks ----
            for (int attempt = 0; attempt < 2; ++attempt) {
                int nt = 0;
                T[0] = T[1] = T[2] = -1;
                if (attempt == 0) {
                    std::vector<size_t> h2 = head;
                    while (true) { const int bi = earliest_ready(h2, true, false, nullptr); if (bi < 0) break; pop_in(h2, bi); }
                    std::vector<int> dense;
                    for (int q : tile_q) { const int f = front_of(h2, q); if (f >= 0 && ready_in(h2, f)) dense.push_back(f); }
                    std::sort(dense.begin(), dense.end());
                    dense.erase(std::unique(dense.begin(), dense.end()), dense.end());
                    for (int bi : dense) {
                        const Block& b = p.blocks[bi];
                        int need = 0;
                        for (int s2 = 0; s2 < b.nq; ++s2) need += slot_of(b.q[s2]) < 0;
                        if (nt + need > 3) continue;
                        for (int s2 = 0; s2 < b.nq; ++s2) if (slot_of(b.q[s2]) < 0) T[nt++] = b.q[s2];
                    }
                    // further dense work on the chosen qubits' neighbours: partners along pending two-qubit dense blocks
                    while (nt > 0 && nt < 3) {
                        int best_q = -1, best_i = 1 << 30;
                        for (int s2 = 0; s2 < nt; ++s2)
                            for (size_t t = h2[T[s2]]; t < qq[T[s2]].size(); ++t) {
                                const Block& nb = p.blocks[qq[T[s2]][t]];
                                if (nb.nq != 2 || mono_of.count(qq[T[s2]][t])) continue;
                                const int partner = nb.q[0] == T[s2] ? nb.q[1] : nb.q[0];
                                if (slot_of(partner) >= 0) continue;
                                if (qq[T[s2]][t] < best_i) { best_i = qq[T[s2]][t]; best_q = partner; }
                                break;
                            }
                        if (best_q < 0) break;
                        T[nt++] = best_q;
                    }
                    if (nt == 0) {       // only monomials are left in this pass: any triple will do
                        const int b0 = earliest_ready(head, true, true, nullptr);
                        if (b0 < 0) throw std::runtime_error("triple builder: no ready block");
                        T[nt++] = p.blocks[b0].q[0];
                        if (p.blocks[b0].nq == 2) T[nt++] = p.blocks[b0].q[1];
                    }
                    // spectators: the qubits whose dense work comes next
                    while (nt < 3) {
                        int best_q = -1, best_i = 1 << 30;
                        for (int q : tile_q) {
                            if (slot_of(q) >= 0) continue;
                            int f = (1 << 30) - 1;
                            for (size_t t = h2[q]; t < qq[q].size(); ++t) if (!mono_of.count(qq[q][t])) { f = qq[q][t]; break; }
                            if (f < best_i) { best_i = f; best_q = q; }
                        }
                        if (best_q < 0) throw std::runtime_error("triple builder: tile has fewer than three qubits");
                        T[nt++] = best_q;
                    }
                } else {
                    // fallback (guarantees progress): the earliest ready block's own qubits, as the local builder does
                    const int b0 = earliest_ready(head, true, true, nullptr);
                    if (b0 < 0) throw std::runtime_error("triple builder: no ready block");
                    T[nt++] = p.blocks[b0].q[0];
                    if (p.blocks[b0].nq == 2) T[nt++] = p.blocks[b0].q[1];
                    while (nt < 3) {
                        int best_q = -1, best_i = 1 << 30;
                        for (int q : tile_q) {
                            if (slot_of(q) >= 0) continue;
                            const int f = front_of(head, q) >= 0 ? front_of(head, q) : (1 << 30) - 1;
                            if (f < best_i) { best_i = f; best_q = q; }
                        }
                        if (best_q < 0) throw std::runtime_error("triple builder: tile has fewer than three qubits");
                        T[nt++] = best_q;
                    }
                }
                // ---- absorb ----
                std::vector<size_t> h = head;
                size_t popped = 0;
                MonoT lead(E), trail(E);
                WideSide ls, ss;
                analyse(lead, false, ls);
                analyse(trail, true, ss);
                auto budget_ok = [&](const WideSide& cand, const WideSide& other) {
                    if (!cand.wide) return true;
                    return pass_bytes() + (cand.table.size() + (other.wide ? other.table.size() : 0) + 64 + 16) * 8 <= PASS_LIMIT;
                };
                // stage 0: leading monomials, anywhere in the tile
                while (true) {
                    const int bi = earliest_ready(h, true, false, nullptr);
                    if (bi < 0) break;
                    // a monomial inside the triple may also wait for the body (one-qubit) - keep the local builder's order: lead
                    MonoT cand = lead;
                    cand.then(mono_of.at(bi));
                    WideSide cs;
                    if (!analyse(cand, false, cs) || !budget_ok(cs, ss)) break;
                    lead.swap_with(cand); ls = cs;
                    pop_in(h, bi); ++popped;
                }
                // stage 1: dense work inside the triple (one-qubit blocks of any class queue per slot)
                bool to_trailing = false;
                while (!to_trailing) {
                    int bi = -1;
                    for (int s2 = 0; s2 < 3; ++s2) {
                        const int f = front_of(h, T[s2]);
                        if (f < 0 || !ready_in(h, f) || (bi >= 0 && f >= bi)) continue;
                        const Block& b = p.blocks[f];
                        bool inside = true;
                        for (int t = 0; t < b.nq; ++t) inside = inside && slot_of(b.q[t]) >= 0;
                        if (inside) bi = f;
                    }
                    if (bi < 0) break;
                    const Block& b = p.blocks[bi];
                    if (b.nq == 1) { pend[slot_of(b.q[0])].push_back(bi); pop_in(h, bi); ++popped; continue; }
                    if (mono_of.count(bi) && b.cls == CLS_MONO) { to_trailing = true; break; }
                    int s0 = slot_of(b.q[0]), s1 = slot_of(b.q[1]);
                    std::vector<double> M = b.M;
                    if (s0 > s1) { M = swap_qubits(M); std::swap(s0, s1); }
                    flush_slot(s0); flush_slot(s1);
                    const int pair = s0 == 0 ? (s1 == 1 ? 0 : 1) : 2;
                    const int cls = classify(M, 16);
                    if (cls == CLS_DIAG || cls == CLS_IDENT) {
                        std::vector<double> dg(16);
                        for (int i = 0; i < 16; ++i) dg[i] = M[i * 16 + i];
                        if (cls != CLS_IDENT) { mus.push_back(mu(MU3_DIAG2, cc.get(p, dg), pair)); flops += 1.0; }
                    } else {
                        mus.push_back(mu(MU3_MAT2, cc.get(p, M), pair));
                        flops += 31.0;
                    }
                    pop_in(h, bi); ++popped;
                }
                // stage 2: trailing monomials, anywhere in the tile - but never past a pending one-qubit block of the body
                while (true) {
                    const int bi = earliest_ready(h, true, false, nullptr);
                    if (bi < 0) break;
                    MonoT cand = trail;
                    cand.then(mono_of.at(bi));
                    WideSide cs;
                    if (!analyse(cand, true, cs) || !budget_ok(cs, ls)) break;
                    trail.swap_with(cand); ss = cs;
                    pop_in(h, bi); ++popped;
                }
                if (popped == 0) {
                    for (int s2 = 0; s2 < 3; ++s2) pend[s2].clear();
                    if (attempt == 1) throw std::runtime_error("triple builder: no progress");
                    continue;
                }
                head = h;
                remaining -= popped;
                emit_wide(ls, ss, lead.trivial, trail.trivial);
                break;
            }
        }
        phase.n_sweeps = (int)p.sweeps3.size() - phase.first_sweep;
        phase.n_chains = (int)p.chains.size() - phase.first_chain;
        if (phase.n_sweeps > 0) p.phases.push_back(phase);
    }

    void run(const std::vector<int>& order) {
        if (p.opt_wide && tile_q.size() >= 4) { run_wide(order); return; }
        const int n = p.n;
        std::vector<std::vector<int>> qq(n);
        std::vector<size_t> head(n, 0);
        std::vector<char> resident(n, 0);
        for (int q : tile_q) resident[q] = 1;
        for (int bi : order) {
            const Block& b = p.blocks[bi];
            if (b.kind == BK_VAR) throw std::runtime_error("triple sweeps do not take variant blocks");
            qq[b.q[0]].push_back(bi);
            if (b.nq == 2) qq[b.q[1]].push_back(bi);
        }
        auto front = [&](int q) { return head[q] < qq[q].size() ? qq[q][head[q]] : -1; };
        auto ready = [&](int bi) {
            const Block& b = p.blocks[bi];
            for (int s = 0; s < b.nq; ++s) if (front(b.q[s]) != bi) return false;
            return true;
        };
        auto pop = [&](int bi) {
            const Block& b = p.blocks[bi];
            for (int s = 0; s < b.nq; ++s) ++head[b.q[s]];
        };
        size_t remaining = order.size();
        while (remaining > 0) {
            int b0 = -1;
            for (int q = 0; q < n; ++q) { int f = front(q); if (f >= 0 && ready(f) && (b0 < 0 || f < b0)) b0 = f; }
            if (b0 < 0) throw std::runtime_error("triple builder: no ready block");
            // ---- choose the triple: the opening block's qubits, then partners along the pending two-qubit blocks ----
            int nt = 0;
            T[0] = T[1] = T[2] = -1;
            T[nt++] = p.blocks[b0].q[0];
            if (p.blocks[b0].nq == 2) T[nt++] = p.blocks[b0].q[1];
            bool grew = true;
            while (nt < 3 && grew) {
                grew = false;
                int best_q = -1, best_i = 1 << 30;
                for (int s = 0; s < nt; ++s)
                    for (size_t t = head[T[s]]; t < qq[T[s]].size(); ++t) {
                        const Block& nb = p.blocks[qq[T[s]][t]];
                        if (nb.nq != 2) continue;
                        const int partner = nb.q[0] == T[s] ? nb.q[1] : nb.q[0];
                        if (slot_of(partner) >= 0) continue;
                        if (qq[T[s]][t] < best_i) { best_i = qq[T[s]][t]; best_q = partner; }
                        break;
                    }
                if (best_q >= 0) { T[nt++] = best_q; grew = true; }
            }
            while (nt < 3) {     // spectators: resident qubits with the most urgent pending work, else any
                int best_q = -1, best_i = 1 << 30;
                for (int q : tile_q) {
                    if (slot_of(q) >= 0) continue;
                    const int f = front(q) >= 0 ? front(q) : (1 << 30) - 1;
                    if (f < best_i) { best_i = f; best_q = q; }
                }
                if (best_q < 0) throw std::runtime_error("triple builder: tile has fewer than three qubits");
                T[nt++] = best_q;
            }
            // ---- absorb ----
            Mono lp, sp;
            int stage = 0;     // 0: load permutation, 1: body, 2: store permutation
            while (true) {
                int bi = -1;
                for (int s = 0; s < 3; ++s) {
                    const int f = front(T[s]);
                    if (f < 0 || !ready(f) || (bi >= 0 && f >= bi)) continue;
                    const Block& b = p.blocks[f];
                    bool inside = true;
                    for (int t = 0; t < b.nq; ++t) inside = inside && slot_of(b.q[t]) >= 0;
                    if (!inside) continue;
                    Mono m;
                    const bool mono = as_mono(b, m);
                    if (stage == 2 && !mono) continue;          // only monomials ride on the store permutation
                    bi = f;
                }
                if (bi < 0) break;
                const Block& b = p.blocks[bi];
                Mono m;
                const bool mono = as_mono(b, m);
                if (stage == 0) {
                    if (mono) { lp.then(m); pop(bi); --remaining; continue; }
                    stage = 1;
                }
                if (stage == 1) {
                    if (b.nq == 1) { pend[slot_of(b.q[0])].push_back(bi); pop(bi); --remaining; continue; }
                    if (mono && b.cls == CLS_MONO) { stage = 2; }
                    else {
                        int s0 = slot_of(b.q[0]), s1 = slot_of(b.q[1]);
                        std::vector<double> M = b.M;
                        if (s0 > s1) { M = swap_qubits(M); std::swap(s0, s1); }
                        flush_slot(s0); flush_slot(s1);
                        const int pair = s0 == 0 ? (s1 == 1 ? 0 : 1) : 2;
                        const int cls = classify(M, 16);
                        if (cls == CLS_DIAG || cls == CLS_IDENT) {
                            std::vector<double> dg(16);
                            for (int i = 0; i < 16; ++i) dg[i] = M[i * 16 + i];
                            if (cls != CLS_IDENT) { mus.push_back(mu(MU3_DIAG2, cc.get(p, dg), pair)); flops += 1.0; }
                        } else {
                            mus.push_back(mu(MU3_MAT2, cc.get(p, M), pair));
                            flops += 31.0;
                        }
                        pop(bi); --remaining;
                        continue;
                    }
                }
                // stage 2
                sp.then(m); pop(bi); --remaining;
            }
            emit(lp, sp);
        }
        phase.n_sweeps = (int)p.sweeps3.size() - phase.first_sweep;
        phase.n_chains = (int)p.chains.size() - phase.first_chain;
        if (phase.n_sweeps > 0) p.phases.push_back(phase);
    }
};

static void schedule_single_pass(Plan& p) {
    CoefCache cc;
    std::vector<int> local_pos(p.n);
    Pass pass;
    std::vector<int> order;
    for (int q = 0; q < p.n; ++q) { pass.tile_q.push_back(q); local_pos[q] = code_pos(p.n, q); }
    for (size_t i = 0; i < p.blocks.size(); ++i) order.push_back((int)i);
    pass.first_phase = (int)p.phases.size();
    if (p.n == 1) {
        // a single qubit has no pair: the tile gets a virtual idle qubit below it (tile index = code << 2);
        // the execution layer runs such plans with k = 2 and reads / writes the coefficients at stride 4
        local_pos[0] = 2;
        local_pos.push_back(0);
        pass.tile_q.push_back(1);
    }
    SweepBuilder sb(p, cc, local_pos, pass.tile_q);
    sb.run(order);
    pass.n_phases = (int)p.phases.size() - pass.first_phase;
    p.passes.push_back(pass);
    p.tile_k = p.n;
    p.info.flops_per_circuit = sb.flops * std::pow(4.0, p.n);
    p.info.smem_bytes_per_circuit = sb.sweeps_elems * 16.0 * std::pow(4.0, p.n);
}

// DAG-greedy pass builder for n > tile_k: the two lowest qubits stay resident (128-byte runs in HBM),
// blocks join the open pass when all their predecessors are scheduled and their qubits still fit.
static void schedule_streaming(Plan& p) {
    const int n = p.n, k = std::min(p.opt_tile_qubits, n);
    p.tile_k = k;
    CoefCache cc;
    const size_t nb = p.blocks.size();
    std::vector<std::vector<int>> chain(n);
    for (size_t i = 0; i < nb; ++i) {
        chain[p.blocks[i].q[0]].push_back((int)i);
        if (p.blocks[i].nq == 2) chain[p.blocks[i].q[1]].push_back((int)i);
    }
    std::vector<size_t> head(n, 0);
    size_t remaining = nb;
    double flops = 0.0, sweeps = 0.0;
    while (remaining > 0) {
        std::vector<char> in_tile(n, 0);
        int count = 0;
        for (int q = n - 1; q >= n - 2 && q >= 0; --q) { in_tile[q] = 1; ++count; }
        std::vector<int> order;
        while (true) {
            // candidates are the heads of the per-qubit chains; take the earliest one that is ready and fits
            int best = -1;
            for (int q0 = 0; q0 < n; ++q0) {
                if (head[q0] >= chain[q0].size()) continue;
                int i = chain[q0][head[q0]];
                if (best >= 0 && i >= best) continue;
                const Block& b = p.blocks[i];
                bool ready = true;
                int need = 0;
                for (int s = 0; s < b.nq; ++s) {
                    int q = b.q[s];
                    if (head[q] >= chain[q].size() || chain[q][head[q]] != i) ready = false;
                    if (!in_tile[q]) ++need;
                }
                if (!ready || count + need > k) continue;
                best = i;
            }
            if (best < 0) break;
            const Block& b = p.blocks[best];
            for (int s = 0; s < b.nq; ++s) {
                if (!in_tile[b.q[s]]) { in_tile[b.q[s]] = 1; ++count; }
                ++head[b.q[s]];
            }
            --remaining;
            order.push_back(best);
        }
        if (order.empty()) throw std::runtime_error("streaming scheduler made no progress");
        // fill the tile up to k qubits (highest free qubits first: longer contiguous runs)
        for (int q = n - 1; q >= 0 && count < k; --q) if (!in_tile[q]) { in_tile[q] = 1; ++count; }
        Pass pass;
        std::vector<int> local_pos(n, -1);
        for (int q = 0; q < n; ++q) if (in_tile[q]) pass.tile_q.push_back(q);
        for (size_t i = 0; i < pass.tile_q.size(); ++i) local_pos[pass.tile_q[i]] = 2 * ((int)pass.tile_q.size() - 1 - (int)i);
        pass.first_phase = (int)p.phases.size();
        SweepBuilder sb(p, cc, local_pos, pass.tile_q);
        sb.run(order);
        pass.n_phases = (int)p.phases.size() - pass.first_phase;
        flops += sb.flops;
        sweeps += sb.sweeps_elems;
        p.passes.push_back(std::move(pass));
    }
    p.info.flops_per_circuit = flops * std::pow(4.0, n);
    p.info.smem_bytes_per_circuit = sweeps * 16.0 * std::pow(4.0, n);
}

// image of tile index bit `bit` in the 4-bit bank-pair index under the swizzle of dmx_stream.cuh / dmx_exec.cu
static unsigned swz_col(int bit) {
    unsigned v = bit < 4 ? 1u << bit : 0u;
    for (int i = 0; i < 4; ++i) if ((SWZ_M[i] >> bit) & 1u) v ^= 1u << i;
    return v;
}

static int gf2_rank4(unsigned v[4]) {
    int rank = 0;
    for (int bit = 0; bit < 4; ++bit) {
        int piv = -1;
        for (int i = rank; i < 4; ++i) if ((v[i] >> bit) & 1u) { piv = i; break; }
        if (piv < 0) continue;
        std::swap(v[rank], v[piv]);
        for (int i = 0; i < 4; ++i) if (i != rank && ((v[i] >> bit) & 1u)) v[i] ^= v[rank];
        ++rank;
    }
    return rank;
}

bool pass_stream_eligible(const Plan& p, const Pass& ps) {
    const int k = (int)ps.tile_q.size();
    if (k != 6 && k != 7) return false;
    int n_sweeps = 0, n_mus = 0, n_chains = 0;
    std::map<int, int> tables;      // pass-local fixed coefficient tables: offset -> length
    for (int ph = 0; ph < ps.n_phases; ++ph) {
        const Phase& phase = p.phases[ps.first_phase + ph];
        n_chains += phase.n_chains;
        for (int si = 0; si < phase.n_sweeps; ++si) {
            ++n_sweeps;
            int first_mu, n_mu;
            if (p.triples) {
                const Sweep3& sw = p.sweeps3[phase.first_sweep + si];
                first_mu = sw.first_mu; n_mu = sw.n_mu;
                if (sw.load_perm) tables[sw.lcoef] = std::max(tables[sw.lcoef], 64 << ((sw.wide & 1) ? sw.l_nvar : 0));
                if (sw.store_perm) tables[sw.scoef] = std::max(tables[sw.scoef], 64 << ((sw.wide & 2) ? sw.s_nvar : 0));
                if (sw.wide) tables[-1 - (phase.first_sweep + si)] = 16;      // descriptor
            } else {
                const SweepDev& sw = p.sweeps[phase.first_sweep + si];
                first_mu = sw.first_mu; n_mu = sw.n_mu;
                if (sw.load_perm) tables[sw.lcoef] = 16;
                if (sw.store_perm) tables[sw.scoef] = 16;
            }
            n_mus += n_mu;
            for (int mi = 0; mi < n_mu; ++mi) {
                const MicroOp& m = p.mus[first_mu + mi];
                switch (m.kind) {
                    case MU_CHAIN_A: case MU_CHAIN_B: break;
                    case MU_FIX_A: case MU_FIX_B: case MU_DIAG2: tables[m.a0] = 16; break;
                    case MU_DIAG_A: case MU_DIAG_B: tables[m.a0] = 4; break;
                    case MU_MAT2: tables[m.a0] = 256; break;
                    case MU3_MAT: if (!m.a2) tables[m.a0] = 16; break;
                    case MU3_DIAG: tables[m.a0] = 4; break;
                    case MU3_DIAG2: tables[m.a0] = 16; break;
                    case MU3_MAT2: tables[m.a0] = 256; break;
                    default: return false;      // variant micro-ops
                }
            }
        }
    }
    size_t coefs = 0;
    for (const auto& t : tables) coefs += (size_t)((t.second + 1) & ~1);
    const bool ok = n_sweeps <= STREAM_MAX_SWEEPS && n_mus <= STREAM_MAX_MUS && (size_t)n_chains * 128 + coefs * 8 <= (k == 7 ? 90u : 20u) * 1024;
    if (!ok && std::getenv("DMX_VERBOSE"))
        std::fprintf(stderr, "[dmx] pass not eligible: %d sweeps, %d micro-ops, %d chains, %zu table doubles\n", n_sweeps, n_mus, n_chains, coefs);
    return ok;
}

// Remapping pass builder (see Pass::pos_in / pos_out).  Same DAG-greedy block selection as schedule_streaming, but no
// qubit is pinned: a pass only has to share two qubits with its predecessor; they become the two lowest positions of
// the layout the predecessor stores (and this pass loads), so every tile is made of >= 128-byte runs on both sides.
static void schedule_streaming_remap(Plan& p) {
    const int n = p.n, k = std::min(p.opt_tile_qubits, n);
    p.tile_k = k;
    const size_t nb = p.blocks.size();
    std::vector<std::vector<int>> chain(n);
    for (size_t i = 0; i < nb; ++i) {
        chain[p.blocks[i].q[0]].push_back((int)i);
        if (p.blocks[i].nq == 2) chain[p.blocks[i].q[1]].push_back((int)i);
    }
    std::vector<size_t> head(n, 0);
    size_t remaining = nb;
    std::vector<std::vector<char>> tiles;
    std::vector<std::vector<int>> orders;
    std::vector<char> prev(n, 0);
    bool have_prev = false;
    const int BIG = 1 << 30;
    auto pending = [&](const std::vector<size_t>& h, int q) { return h[q] < chain[q].size() ? chain[q][h[q]] : BIG; };

    // earliest-first growth of a tile from `seeds` (the original greedy), then fill up to k qubits
    auto grow = [&](std::vector<char> in_tile, int first_block) {
        std::vector<size_t> h = head;
        int count = 0;
        for (int q = 0; q < n; ++q) count += in_tile[q];
        auto take = [&](int i) {
            const Block& b = p.blocks[i];
            for (int s = 0; s < b.nq; ++s) {
                if (!in_tile[b.q[s]]) { in_tile[b.q[s]] = 1; ++count; }
                ++h[b.q[s]];
            }
        };
        if (first_block >= 0) take(first_block);
        while (true) {
            int best = -1;
            for (int q0 = 0; q0 < n; ++q0) {
                if (h[q0] >= chain[q0].size()) continue;
                int i = chain[q0][h[q0]];
                if (best >= 0 && i >= best) continue;
                const Block& b = p.blocks[i];
                bool ready = true;
                int need = 0;
                for (int s = 0; s < b.nq; ++s) {
                    int q = b.q[s];
                    if (h[q] >= chain[q].size() || chain[q][h[q]] != i) ready = false;
                    if (!in_tile[q]) ++need;
                }
                if (!ready || count + need > k) continue;
                best = i;
            }
            if (best < 0) break;
            take(best);
        }
        // fill: qubits of the previous tile first (they can be the shared low pair), then the most urgent ones
        std::vector<int> cand;
        for (int q = 0; q < n; ++q) if (!in_tile[q]) cand.push_back(q);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
            const int pa = have_prev && prev[a] ? 0 : 1, pb = have_prev && prev[b] ? 0 : 1;
            if (pa != pb) return pa < pb;
            return pending(h, a) < pending(h, b);
        });
        for (int q : cand) { if (count >= k) break; in_tile[q] = 1; ++count; }
        return in_tile;
    };
    // everything a fixed tile can execute: ready blocks whose qubits are all resident, to closure
    auto absorb = [&](const std::vector<char>& in_tile, std::vector<size_t>& h, std::vector<int>& order) {
        h = head; order.clear();
        while (true) {
            int best = -1;
            for (int q0 = 0; q0 < n; ++q0) {
                if (!in_tile[q0] || h[q0] >= chain[q0].size()) continue;
                int i = chain[q0][h[q0]];
                if (best >= 0 && i >= best) continue;
                const Block& b = p.blocks[i];
                bool ok = true;
                for (int s = 0; s < b.nq; ++s) {
                    int q = b.q[s];
                    if (!in_tile[q] || h[q] >= chain[q].size() || chain[q][h[q]] != i) ok = false;
                }
                if (ok) best = i;
            }
            if (best < 0 || (int)order.size() >= STREAM_MAX_BLOCKS) break;   // the kernel's parameter tables are bounded
            const Block& b = p.blocks[best];
            for (int s = 0; s < b.nq; ++s) ++h[b.q[s]];
            order.push_back(best);
        }
    };
    while (remaining > 0) {
        // candidate tiles: the earliest-first greedy, the same forced to start at every qubit's pending block, every
        // window of k adjacent qubits; the one that executes the most work wins (it must share two qubits with the
        // previous tile)
        std::vector<std::vector<char>> cands;
        cands.push_back(grow(std::vector<char>(n, 0), -1));
        for (int q = 0; q < n; ++q) {
            if (head[q] >= chain[q].size()) continue;
            const int i = chain[q][head[q]];
            const Block& b = p.blocks[i];
            bool ready = true;
            for (int s = 0; s < b.nq; ++s) if (chain[b.q[s]][head[b.q[s]]] != i) ready = false;
            if (ready) cands.push_back(grow(std::vector<char>(n, 0), i));
        }
        for (int a = 0; a + k <= n; ++a) {
            std::vector<char> t(n, 0);
            for (int q = a; q < a + k; ++q) t[q] = 1;
            cands.push_back(t);
        }
        if (have_prev) {   // tiles seeded with two qubits of the previous tile always qualify
            std::vector<int> pq;
            for (int q = 0; q < n; ++q) if (prev[q]) pq.push_back(q);
            std::stable_sort(pq.begin(), pq.end(), [&](int a, int b) { return pending(head, a) < pending(head, b); });
            std::vector<char> t(n, 0);
            t[pq[0]] = t[pq[1]] = 1;
            cands.push_back(grow(t, -1));
        }
        size_t best_work = 0;
        int best_shared = -1;
        std::vector<char> best_tile;
        std::vector<size_t> best_h;
        std::vector<int> best_order;
        for (const auto& t : cands) {
            int shared = 0;
            for (int q = 0; q < n; ++q) shared += t[q] && prev[q];
            if (have_prev && shared < 2) continue;
            std::vector<size_t> h;
            std::vector<int> order;
            absorb(t, h, order);
            if (order.size() > best_work || (order.size() == best_work && shared > best_shared)) {
                best_work = order.size(); best_shared = shared; best_tile = t; best_h = h; best_order = order;
            }
        }
        if (best_work == 0) throw std::runtime_error("streaming scheduler made no progress");
        head = best_h; remaining -= best_work;
        tiles.push_back(best_tile); orders.push_back(best_order);
        prev = best_tile; have_prev = true;
    }
    if (n >= 2 && !(tiles.back()[n - 1] && tiles.back()[n - 2])) {
        // no room in the last tile: one transpose-only pass restores the standard order
        std::vector<char> t(n, 0);
        int count = 0;
        t[n - 1] = t[n - 2] = 1; count = 2;
        for (int q = 0; q < n && count < k; ++q) if (!t[q] && tiles.back()[q]) { t[q] = 1; ++count; }
        for (int q = n - 1; q >= 0 && count < k; --q) if (!t[q]) { t[q] = 1; ++count; }
        tiles.push_back(t); orders.push_back({});
    }
    // layouts
    const size_t L = tiles.size();
    // Triple sweeps: which qubits the FIRST and the LAST sweep of every pass hold in registers (a dry run of the sweep
    // builder; the choice of triples does not depend on the layout).  The execution layer lets the first sweep read the
    // tile where the bulk copy put it (global order) and the last sweep write its registers straight to global memory
    // when neither of the two lowest-positioned qubits of that side is register-resident in that sweep: the 16 lanes of
    // a half-warp then walk the 16 doubles of a 128-byte run (conflict-free shared-memory reads, full-line stores).
    std::vector<std::array<int, 3>> first3(L, std::array<int, 3>{-1, -1, -1}), last3(L, std::array<int, 3>{-1, -1, -1});
    // Wide sweeps are decided per pass: a pass whose tables no longer fit beside the tile (pass_stream_eligible), or that
    // would need more sweeps than the triple-local builder, is built by the local builder instead.
    std::vector<char> pass_wide(L, 0);
    const int want_wide = p.opt_wide;
    if (p.opt_triples && k >= 3) {
        for (size_t pi = 0; pi < L; ++pi) {
            if (orders[pi].empty()) continue;
            std::vector<int> tq, lp(n, -1);
            for (int q = 0; q < n; ++q) if (tiles[pi][q]) tq.push_back(q);
            for (size_t i = 0; i < tq.size(); ++i) lp[tq[i]] = 2 * (int)i;
            size_t local_sweeps = 0;
            for (int mode = 0; mode <= (want_wide ? 1 : 0); ++mode) {
                const size_t n_sw = p.sweeps3.size(), n_mu = p.mus.size(), n_ch = p.chains.size(), n_fa = p.factors.size(), n_ph = p.phases.size(),
                             n_co = p.coefs.size();
                CoefCache scratch;
                std::vector<std::array<int, 3>> rec;
                p.opt_wide = mode;
                bool ok = true;
                {
                    TripleBuilder tb(p, scratch, lp, tq);
                    tb.record = &rec;
                    tb.run(orders[pi]);
                }
                if (mode == 1) {
                    Pass probe;
                    probe.tile_q = tq;
                    probe.first_phase = (int)n_ph; probe.n_phases = (int)p.phases.size() - (int)n_ph;
                    ok = pass_stream_eligible(p, probe) && rec.size() < local_sweeps;
                } else {
                    local_sweeps = rec.size();
                }
                p.sweeps3.resize(n_sw); p.mus.resize(n_mu); p.chains.resize(n_ch); p.factors.resize(n_fa); p.phases.resize(n_ph); p.coefs.resize(n_co);
                if (ok && !rec.empty()) { first3[pi] = rec.front(); last3[pi] = rec.back(); pass_wide[pi] = (char)mode; }
            }
        }
    }
    p.opt_wide = want_wide;
    auto in3 = [](const std::array<int, 3>& t, int q) { return t[0] == q || t[1] == q || t[2] == q; };
    std::vector<std::vector<int>> pos(L + 1, std::vector<int>(n));
    for (int q = 0; q < n; ++q) pos[L][q] = n - 1 - q;
    pos[0] = pos[L];
    for (size_t pi = 1; pi < L; ++pi) {
        std::vector<int> seq;   // qubits by ascending position
        std::vector<int> shared;
        for (int q = 0; q < n; ++q) if (tiles[pi - 1][q] && tiles[pi][q]) shared.push_back(q);
        if (shared.size() < 2) throw std::runtime_error("remap scheduler: consecutive tiles share fewer than two qubits");
        // the low pair: two shared qubits, preferably resident neither in the last sweep of the storing pass (2 points:
        // direct stores save a shared-memory round trip) nor in the first sweep of the loading pass (1 point each side of
        // the bulk-staged load); ties keep the lowest qubit numbers
        int best_a = shared[0], best_b = shared[1], best_score = -1;
        for (size_t i = 0; i < shared.size(); ++i)
            for (size_t j = i + 1; j < shared.size(); ++j) {
                const int a = shared[i], b = shared[j];
                const int score = 2 * (!in3(last3[pi - 1], a) && !in3(last3[pi - 1], b)) + (!in3(first3[pi], a) && !in3(first3[pi], b));
                if (score > best_score) { best_score = score; best_a = a; best_b = b; }
            }
        seq.push_back(best_a); seq.push_back(best_b);
        std::vector<char> used(n, 0);
        for (int q : seq) used[q] = 1;
        for (int q = n - 1; q >= 0; --q) if (!used[q] && tiles[pi][q]) { seq.push_back(q); used[q] = 1; }   // this pass's tile is contiguous
        for (int q = n - 1; q >= 0; --q) if (!used[q]) { seq.push_back(q); used[q] = 1; }
        for (int i = 0; i < n; ++i) pos[pi][seq[i]] = i;
    }
    CoefCache cc;
    double flops = 0.0, sweeps = 0.0;
    for (size_t pi = 0; pi < L; ++pi) {
        Pass pass;
        pass.pos_in = pos[pi]; pass.pos_out = pos[pi + 1];
        std::vector<int> tq;
        for (int q = 0; q < n; ++q) if (tiles[pi][q]) tq.push_back(q);
        // Tile-local order.  The staging loops walk the tile in the order of the GLOBAL addresses of each side, so the
        // 16 lanes of a half-warp differ in four tile bits that depend on where the three lowest-positioned tile qubits
        // of that side sit inside the tile.  Shared-memory accesses are conflict-free iff those four bits map to
        // linearly independent bank-pair indices under the swizzle: search the orders for one that does so on both sides
        // (default: the order of the stored layout).
        std::sort(tq.begin(), tq.end(), [&](int a, int b) { return pass.pos_out[a] > pass.pos_out[b]; });
        {
            const int kk = (int)tq.size();
            auto side_rank = [&](const std::vector<int>& order, const std::vector<int>& ps) {
                std::vector<int> tl(n, -1);
                for (int i = 0; i < kk; ++i) tl[order[i]] = kk - 1 - i;
                std::vector<int> byp = order;
                std::sort(byp.begin(), byp.end(), [&](int a, int b) { return ps[a] < ps[b]; });
                unsigned v[4] = {swz_col(2 * tl[byp[0]] + 1), swz_col(2 * tl[byp[1]]), swz_col(2 * tl[byp[1]] + 1), swz_col(2 * tl[byp[2]])};
                return gf2_rank4(v);
            };
            // What an order is worth (see dmx_exec.cu::build_stream_tables for the matching run-time checks):
            //  * direct store (the last sweep writes global memory from registers): the two lowest tile-local qubits outside
            //    its triple must be the low pair of the stored layout - then the lanes of a half-warp walk one 128-byte run;
            //  * bulk-staged load (first sweep reads the linear image): its threads hold the load-side low pair in their low
            //    bits, so that sweep's store into the swizzled tile walks the tile-local bits of that pair - they must map
            //    to four independent bank-pair bits;
            //  * a side that stays on the register-staged path wants the old staging-loop criterion.
            auto lowest_two_outside = [&](const std::vector<int>& order, const std::array<int, 3>& t3, const std::vector<int>& ps) {
                int found = 0;
                for (int i = kk - 1; i >= 0 && found < 2; --i) {       // order[i] sits at tile-local position kk-1-i
                    const int q = order[i];
                    if (in3(t3, q)) continue;
                    if (ps[q] > 1) return false;
                    ++found;
                }
                return found == 2;
            };
            auto low_pair_clear = [&](const std::array<int, 3>& t3, const std::vector<int>& ps) {
                if (t3[0] < 0) return false;
                int cnt = 0;
                for (int q : tq) if (ps[q] <= 1) { if (in3(t3, q)) return false; ++cnt; }
                return cnt == 2;
            };
            const bool store_ok = p.opt_triples && low_pair_clear(last3[pi], pass.pos_out);
            bool load_ok = p.opt_triples && pi > 0 && low_pair_clear(first3[pi], pass.pos_in);
            if (load_ok) for (int q : tq) if (pass.pos_in[q] >= kk) load_ok = false;      // contiguous tile only
            auto bulk_rank = [&](const std::vector<int>& order) {
                std::vector<int> tl(n, -1);
                for (int i = 0; i < kk; ++i) tl[order[i]] = kk - 1 - i;
                unsigned v[4];
                int j = 0;
                for (int q : order) if (pass.pos_in[q] <= 1) { v[j++] = swz_col(2 * tl[q]); v[j++] = swz_col(2 * tl[q] + 1); }
                return gf2_rank4(v);
            };
            auto score_of = [&](const std::vector<int>& order) {
                int sc = 0;
                if (store_ok && lowest_two_outside(order, last3[pi], pass.pos_out)) sc += 100;
                else sc += 5 * side_rank(order, pass.pos_out);
                if (load_ok) sc += 10 * bulk_rank(order);
                else sc += 5 * (pi > 0 ? side_rank(order, pass.pos_in) : 4);
                return sc;
            };
            const int full_score = (store_ok ? 100 : 20) + (load_ok ? 40 : 20);
            std::vector<int> best = tq, cand = tq;
            int best_score = score_of(tq);
            std::sort(cand.begin(), cand.end());
            if (best_score < full_score && kk >= 3) {
                do {
                    const int sc = score_of(cand);
                    if (sc > best_score) { best_score = sc; best = cand; if (sc == full_score) break; }
                } while (std::next_permutation(cand.begin(), cand.end()));
            }
            tq = best;
        }
        pass.tile_q = tq;
        std::vector<int> local_pos(n, -1);
        for (size_t i = 0; i < tq.size(); ++i) local_pos[tq[i]] = 2 * ((int)tq.size() - 1 - (int)i);
        pass.first_phase = (int)p.phases.size();
        if (!orders[pi].empty()) {
            if (p.triples) {
                p.opt_wide = pass_wide[pi];
                TripleBuilder sb(p, cc, local_pos, pass.tile_q);
                sb.run(orders[pi]);
                p.opt_wide = want_wide;
                flops += sb.flops;
                sweeps += sb.sweeps_elems;
            } else {
                SweepBuilder sb(p, cc, local_pos, pass.tile_q);
                sb.run(orders[pi]);
                flops += sb.flops;
                sweeps += sb.sweeps_elems;
            }
        }
        pass.n_phases = (int)p.phases.size() - pass.first_phase;
        p.passes.push_back(std::move(pass));
    }
    p.info.flops_per_circuit = flops * std::pow(4.0, n);
    p.info.smem_bytes_per_circuit = sweeps * 16.0 * std::pow(4.0, n);
}

// ---- segment tables (path 1) --------------------------------------------------------------------------

static const int SEG_N = 4, SEG_E = 256;

struct Mono {  // r_out[j] = w[j] * r_in[src[j]]
    int src[SEG_E];
    double w[SEG_E];
    void reset() { for (int j = 0; j < SEG_E; ++j) { src[j] = j; w[j] = 1.0; } }
};

static inline int get_code(int idx, int pos) { return (idx >> pos) & 3; }
static inline int set_code(int idx, int pos, int c) { return (idx & ~(3 << pos)) | (c << pos); }

// compose a monomial/diagonal fixed block onto the running monomial
static void apply_mono_block(Mono& m, const Block& b) {
    const int dim = b.dim();
    const int p0 = code_pos(SEG_N, b.q[0]), p1 = code_pos(SEG_N, b.q[1]);
    int srcl[16];
    double wl[16];
    for (int a = 0; a < dim; ++a) {
        srcl[a] = a; wl[a] = 0.0;
        for (int c = 0; c < dim; ++c)
            if (b.M[a * dim + c] != 0.0) { srcl[a] = c; wl[a] = b.M[a * dim + c]; }
    }
    Mono out;
    for (int j = 0; j < SEG_E; ++j) {
        int a = b.nq == 1 ? get_code(j, p0) : get_code(j, p0) * 4 + get_code(j, p1);
        int c = srcl[a];
        int jin = b.nq == 1 ? set_code(j, p0, c) : set_code(set_code(j, p0, c >> 2), p1, c & 3);
        out.src[j] = m.src[jin];
        out.w[j] = wl[a] * m.w[jin];
    }
    m = out;
}

static bool build_segments(Plan& p, std::string& why) {
    if (p.n > SEG_N) { why = "n > 4"; return false; }
    if (p.term_idx.empty()) { why = "no Hamiltonian"; return false; }
    // qubit q of the circuit sits at the same index as in a 4-qubit register whose extra qubits idle
    p.segs.clear();
    p.slots.assign(p.n_slots, SlotInfo());
    std::vector<char> slot_seen(p.n_slots, 0);
    Mono cur;
    cur.reset();
    CoefCache cc;
    for (const Block& b : p.blocks) {
        if (b.kind == BK_FIXED) {
            if (b.cls == CLS_DENSE) {
                // A fixed block whose rows hold at most two non-zeros (a resolved rz/rx/ry, possibly fused with
                // Clifford gates and Pauli noise) closes a segment like a rotation with constant coefficients.
                const int dim = b.dim();
                const int p0 = code_pos(SEG_N, b.q[0]), p1 = code_pos(SEG_N, b.q[1]);
                Segment sg;
                sg.src0.resize(SEG_E); sg.src1.resize(SEG_E); sg.active.resize(SEG_E);
                sg.w0.resize(SEG_E); sg.w1.resize(SEG_E);
                sg.rot_id = -2;
                // Primary column per row.  The block is (monomial) x (one Pauli rotation) x (monomial): rows with one
                // non-zero ("passive") fix an affine GF(2) map on the local codes; the rows with two non-zeros must be
                // matched by the same affine map so that the primary sources form a group automorphism (needed for
                // the constant-XOR partner structure of the fixed frame).  Two extensions exist (cos- or sin-term
                // as primary); the first consistent one is taken.
                int primary[16], nz1[16], nz2[16], cnt[16];
                for (int a = 0; a < dim; ++a) {
                    cnt[a] = 0; nz1[a] = nz2[a] = -1;
                    for (int c = 0; c < dim; ++c)
                        if (b.M[a * dim + c] != 0.0) { if (cnt[a] == 0) nz1[a] = c; else nz2[a] = c; ++cnt[a]; }
                    if (cnt[a] > 2) { why = "dense fixed block (non-Clifford gate or non-Pauli channel)"; return false; }
                }
                bool ok_primary = false;
                int a0 = -1;
                for (int a = 0; a < dim; ++a) if (cnt[a] == 2) { a0 = a; break; }
                if (a0 >= 0 && cnt[0] == 1) {
                    for (int cand = 0; cand < 2 && !ok_primary; ++cand) {
                        ok_primary = true;
                        for (int a = 0; a < dim && ok_primary; ++a) {
                            if (cnt[a] == 1) { primary[a] = nz1[a]; continue; }
                            if (cnt[a] == 0) { ok_primary = false; break; }
                            int d = a ^ a0;
                            if (cnt[d] != 1) { ok_primary = false; break; }
                            int pa = (cand == 0 ? nz1[a0] : nz2[a0]) ^ nz1[d] ^ nz1[0];
                            if (pa != nz1[a] && pa != nz2[a]) { ok_primary = false; break; }
                            primary[a] = pa;
                        }
                        if (ok_primary) {   // must be a permutation
                            int seen = 0;
                            for (int a = 0; a < dim; ++a) seen |= 1 << primary[a];
                            ok_primary = seen == (1 << dim) - 1;
                        }
                    }
                }
                if (!ok_primary) { why = "2-sparse fixed block without a Pauli-rotation structure"; return false; }
                for (int j = 0; j < SEG_E; ++j) {
                    int a = b.nq == 1 ? get_code(j, p0) : get_code(j, p0) * 4 + get_code(j, p1);
                    int cols[2] = {primary[a], primary[a]}, nnz = 0;
                    double vals[2] = {b.M[a * dim + primary[a]], 0.0};
                    for (int c = 0; c < dim; ++c)
                        if (b.M[a * dim + c] != 0.0) {
                            ++nnz;
                            if (nnz > 2) { why = "dense fixed block (non-Clifford gate or non-Pauli channel)"; return false; }
                            if (c != primary[a]) { cols[1] = c; vals[1] = b.M[a * dim + c]; }
                        }
                    int jj[2];
                    for (int t = 0; t < 2; ++t)
                        jj[t] = b.nq == 1 ? set_code(j, p0, cols[t]) : set_code(set_code(j, p0, cols[t] >> 2), p1, cols[t] & 3);
                    sg.src0[j] = (uint8_t)cur.src[jj[0]]; sg.w0[j] = vals[0] * cur.w[jj[0]];
                    sg.src1[j] = (uint8_t)cur.src[jj[1]]; sg.w1[j] = vals[1] * cur.w[jj[1]];
                    sg.active[j] = vals[1] != 0.0;
                }
                p.segs.push_back(std::move(sg));
                cur.reset();
                continue;
            }
            apply_mono_block(cur, b);
        } else if (b.kind == BK_VAR) {
            if (slot_seen[b.slot]) { why = "variant slot used twice"; return false; }
            slot_seen[b.slot] = 1;
            SlotInfo& s = p.slots[b.slot];
            s.nq = b.nq;
            s.seg = (int)p.segs.size();
            s.label.resize(SEG_E);
            // label of INPUT element i = local code of the position it currently occupies
            const int p0 = code_pos(SEG_N, b.q[0]), p1 = code_pos(SEG_N, b.q[1]);
            std::vector<int> where(SEG_E, -1);  // input element -> current position (cur.src is position -> input)
            for (int j = 0; j < SEG_E; ++j) where[cur.src[j]] = j;
            for (int i = 0; i < SEG_E; ++i) {
                int j = where[i];
                if (j < 0) { why = "non-invertible monomial before a variant slot"; return false; }
                s.label[i] = (uint8_t)(b.nq == 1 ? get_code(j, p0) : get_code(j, p0) * 4 + get_code(j, p1));
            }
            if (b.nq == 1) {
                std::vector<double> dt(16 * 4, 1.0);
                for (int k = 0; k < 16; ++k) {
                    std::vector<double> Bk(b.tab.begin() + 16 * k, b.tab.begin() + 16 * (k + 1));
                    auto t = k == 0 ? identity(4) : matmul(b.cond, Bk, 4);
                    if (is_diag(t, 4)) {
                        s.diag_ok[k >> 5] |= 1u << (k & 31);
                        for (int c = 0; c < 4; ++c) dt[k * 4 + c] = t[c * 4 + c];
                    }
                }
                s.tab_off = cc.get(p, dt);
            } else {
                std::vector<double> dt(16 * 4 + 16, 1.0);
                std::vector<char> ok1(16, 0);
                for (int k = 0; k < 16; ++k) {
                    std::vector<double> Bk(b.tab.begin() + 16 * k, b.tab.begin() + 16 * (k + 1));
                    if (is_diag(Bk, 4)) {
                        ok1[k] = 1;
                        for (int c = 0; c < 4; ++c) dt[k * 4 + c] = Bk[c * 4 + c];
                    }
                }
                bool cond_diag = is_diag(b.cond, 16);
                for (int c = 0; c < 16; ++c) dt[64 + c] = cond_diag ? b.cond[c * 16 + c] : 1.0;
                for (int k = 0; k < 256; ++k) {
                    bool ok = k == 0 || (cond_diag && ok1[k >> 4] && ok1[k & 15]);
                    if (ok) s.diag_ok[k >> 5] |= 1u << (k & 31);
                }
                s.tab_off = cc.get(p, dt);
            }
        } else {  // BK_ROT: close the segment
            Segment sg;
            sg.src0.resize(SEG_E); sg.src1.resize(SEG_E); sg.active.resize(SEG_E);
            sg.w0.resize(SEG_E); sg.w1.resize(SEG_E);
            sg.rot_id = b.rot_id;
            const int p0 = code_pos(SEG_N, b.q[0]);
            for (int j = 0; j < SEG_E; ++j) {
                int a = get_code(j, p0);
                int nA = 0, nC = 0, nS = 0, bA = 0, bC = 0, bS = 0;
                for (int c = 0; c < 4; ++c) {
                    if (b.A[a * 4 + c] != 0.0) { ++nA; bA = c; }
                    if (b.C[a * 4 + c] != 0.0) { ++nC; bC = c; }
                    if (b.S[a * 4 + c] != 0.0) { ++nS; bS = c; }
                }
                if (nA > 1 || nC > 1 || nS > 1 || (nA == 1 && (nC || nS)) || (nC != nS)) {
                    why = "rotation block is not a 2-sparse row pattern";
                    return false;
                }
                if (nC == 1) {
                    int j0 = set_code(j, p0, bC), j1 = set_code(j, p0, bS);
                    sg.active[j] = 1;
                    sg.src0[j] = (uint8_t)cur.src[j0]; sg.w0[j] = b.C[a * 4 + bC] * cur.w[j0];
                    sg.src1[j] = (uint8_t)cur.src[j1]; sg.w1[j] = b.S[a * 4 + bS] * cur.w[j1];
                } else {
                    int j0 = set_code(j, p0, bA);
                    sg.active[j] = 0;
                    sg.src0[j] = (uint8_t)cur.src[j0]; sg.w0[j] = nA ? b.A[a * 4 + bA] * cur.w[j0] : 0.0;
                    sg.src1[j] = sg.src0[j]; sg.w1[j] = 0.0;
                }
            }
            p.segs.push_back(std::move(sg));
            cur.reset();
        }
    }
    for (int s = 0; s < p.n_slots; ++s)
        if (!slot_seen[s]) { p.slots[s].seg = -1; p.slots[s].label.assign(SEG_E, 0); }
    // energy stage: E = sum_t c_t * w[idx_t] * r_in[src[idx_t]]
    const int shift = 2 * (SEG_N - p.n);
    p.e_src.clear(); p.e_w.clear(); p.e_fac.clear();
    for (size_t t = 0; t < p.term_idx.size(); ++t) {
        int j = (int)(p.term_idx[t] << shift);
        p.e_src.push_back((uint8_t)cur.src[j]);
        p.e_w.push_back(p.term_coef[t] * cur.w[j]);
        p.e_fac.push_back(cur.w[j]);
    }
    p.nseg = (int)p.segs.size();
    p.coefs.reserve(p.coefs.size());
    return true;
}

// ---- fixed-frame conversion of the segment tables ----------------------------------------------------------

static inline int parity8(unsigned v) { return __builtin_popcount(v) & 1; }

// rows[r] = GF(2) functional giving bit r of the physical index; returns false if not invertible
static bool gf2_invertible(const unsigned rows[8]) {
    unsigned m[8];
    for (int i = 0; i < 8; ++i) m[i] = rows[i];
    int rank = 0;
    for (int bit = 0; bit < 8 && rank < 8; ++bit) {
        int piv = -1;
        for (int i = rank; i < 8; ++i) if ((m[i] >> bit) & 1u) { piv = i; break; }
        if (piv < 0) continue;
        std::swap(m[rank], m[piv]);
        for (int i = 0; i < 8; ++i) if (i != rank && ((m[i] >> bit) & 1u)) m[i] ^= m[rank];
        ++rank;
    }
    return rank == 8;
}

static int gf2_rank(const std::vector<unsigned>& vecs) {
    std::vector<unsigned> basis;
    for (unsigned v : vecs) {
        for (unsigned b : basis) v = std::min(v, v ^ b);
        if (v) basis.push_back(v);
    }
    return (int)basis.size();
}

// Liveness of the fixed-frame program: forward (reachable from the initial state) and backward (able to reach an energy
// term) over the segments; a coefficient outside both can be dropped without changing Tr(rho H).  Variant slots only scale
// coefficients (diagonal corrections), so the analysis holds for every variant row the frame kernel accepts.
static void build_frame32(Plan& p) {
    p.f32_n = 0;
    const int K = p.nseg;
    std::vector<std::vector<char>> fwd(K + 1, std::vector<char>(SEG_E, 0)), need(K + 1, std::vector<char>(SEG_E, 0));
    for (int t = 0; t < SEG_E; ++t) { fwd[0][t] = p.f_init[t] != 0.0; need[K][t] = p.f_eweight[t] != 0.0; }
    auto w0 = [&](int k, int t) { return p.f_w[((size_t)k * SEG_E + t) * 2]; };
    auto w1 = [&](int k, int t) { return p.f_w[((size_t)k * SEG_E + t) * 2 + 1]; };
    for (int k = 0; k < K; ++k)
        for (int t = 0; t < SEG_E; ++t) {
            const int u = t ^ p.fsegs[k].xor_mask;
            fwd[k + 1][t] = (w0(k, t) != 0.0 && fwd[k][t]) || (w1(k, t) != 0.0 && fwd[k][u]);
        }
    for (int k = K - 1; k >= 0; --k)
        for (int t = 0; t < SEG_E; ++t) {
            if (!need[k + 1][t]) continue;
            if (w0(k, t) != 0.0) need[k][t] = 1;
            if (w1(k, t) != 0.0) need[k][t ^ p.fsegs[k].xor_mask] = 1;
        }
    std::vector<int> lane_of(SEG_E, -1), orig;
    for (int t = 0; t < SEG_E; ++t) {
        bool live = false;
        for (int k = 0; k <= K; ++k) live = live || (fwd[k][t] && need[k][t]);
        if (live) { lane_of[t] = (int)orig.size(); orig.push_back(t); }
    }
    if (orig.empty() || orig.size() > 32) return;
    p.f32_n = (int)orig.size();
    orig.resize(32, -1);
    p.f32_orig = orig;
    p.f32_partner.assign((size_t)std::max(K, 1) * 32, 0);
    p.f32_w.assign((size_t)std::max(K, 1) * 64, 0.0);
    p.f32_init.assign(32, 0.0);
    p.f32_eweight.assign(32, 0.0);
    for (int l = 0; l < 32; ++l) {
        const int t = orig[l];
        for (int k = 0; k < K; ++k) p.f32_partner[(size_t)k * 32 + l] = (uint8_t)l;
        if (t < 0) continue;
        p.f32_init[l] = p.f_init[t];
        p.f32_eweight[l] = p.f_eweight[t];
        for (int k = 0; k < K; ++k) {
            const int u = t ^ p.fsegs[k].xor_mask;
            p.f32_w[((size_t)k * 32 + l) * 2] = w0(k, t);
            // bit 7: the rotation moves this coefficient (it takes the cos factor) even when its partner is outside the
            // live set — such a partner is either zero at this stage or feeds a coefficient nothing needs
            if (w1(k, t) != 0.0) p.f32_partner[(size_t)k * 32 + l] |= 0x80;
            if (w1(k, t) != 0.0 && lane_of[u] >= 0) {
                p.f32_w[((size_t)k * 32 + l) * 2 + 1] = w1(k, t);
                p.f32_partner[(size_t)k * 32 + l] = (uint8_t)(lane_of[u] | 0x80);
            }
        }
    }
}

// Thread-per-circuit tables (Plan::f16_*) from the compact frame.  Needs the per-group compact energy weights.
static void build_frame16t(Plan& p) {
    p.f16_n = 0;
    const int K = p.nseg;
    p.f16_const = 0;
    if (p.f32_n <= 0 || K < 1 || K > 48) return;
    const bool rot_mode = p.f_classes.size() == 1 && p.n_slots == 0;           // parameter sweeps: one rotation class, no variant slots
    const bool const_mode = p.f_classes.empty() && p.n_slots > 0;              // QP variant batches of a resolved circuit
    if (!rot_mode && !const_mode) return;
    for (int k = 0; k < K; ++k) if (rot_mode ? p.fsegs[k].cs_sel != 0 : p.fsegs[k].cs_sel >= 0) return;
    auto flagged = [&](int k, int l) { return (p.f32_partner[(size_t)k * 32 + l] & 0x80u) != 0; };
    std::vector<int> touched;
    std::vector<unsigned> used_masks;
    for (int l = 0; l < 32; ++l) {
        if (p.f32_orig[l] < 0) continue;
        bool any = false;
        for (int k = 0; k < K; ++k) any = any || flagged(k, l);
        if (any) touched.push_back(l);
    }
    if (touched.empty() || touched.size() > 16) return;
    for (int k = 0; k < K; ++k) {
        bool any = false;
        for (int l : touched) any = any || flagged(k, l);
        if (any && p.fsegs[k].xor_mask) used_masks.push_back((unsigned)p.fsegs[k].xor_mask);
    }
    // reduced echelon basis of span{masks}: every pivot bit is set in exactly one basis vector
    std::vector<unsigned> basis;
    std::vector<int> pivot;
    for (unsigned v : used_masks) {
        for (size_t j = 0; j < basis.size(); ++j) if ((v >> pivot[j]) & 1u) v ^= basis[j];
        if (!v) continue;
        const int pb = 31 - __builtin_clz(v);
        for (size_t j = 0; j < basis.size(); ++j) if ((basis[j] >> pb) & 1u) basis[j] ^= v;
        basis.push_back(v); pivot.push_back(pb);
    }
    const int rho = (int)basis.size();
    if (rho > 4) return;
    auto coords = [&](unsigned t) { unsigned c = 0; for (int j = 0; j < rho; ++j) c |= ((t >> pivot[j]) & 1u) << j; return c; };
    auto rep = [&](unsigned t) { for (int j = 0; j < rho; ++j) if ((t >> pivot[j]) & 1u) t ^= basis[j]; return t; };
    std::vector<unsigned> reps;
    std::vector<int> slot_of(32, -1), lane_of_slot(16, -1);
    for (int l : touched) {
        const unsigned t = (unsigned)p.f32_orig[l], r = rep(t);
        size_t c = 0;
        while (c < reps.size() && reps[c] != r) ++c;
        if (c == reps.size()) reps.push_back(r);
        if (((c + 1) << rho) > 16) return;
        const int sl = (int)((c << rho) | coords(t));
        if (lane_of_slot[sl] >= 0) return;                                       // cannot happen for distinct t; defensive
        slot_of[l] = sl; lane_of_slot[sl] = l;
    }
    p.f16_mask.assign(K, 0);
    p.f16_tab.assign((size_t)K * 64, 0.0);
    p.f16_init.assign(16, 0.0);
    for (int k = 0; k < K; ++k) {
        bool any = false;
        for (int l : touched) any = any || flagged(k, l);
        const unsigned m = any ? coords((unsigned)p.fsegs[k].xor_mask) : 0u;
        p.f16_mask[k] = (uint8_t)m;
        for (int i = 0; i < 16; ++i) {
            double* e = &p.f16_tab[((size_t)k * 16 + i) * 4];
            const int l = lane_of_slot[i];
            if (l < 0) { e[1] = 1.0; continue; }
            const double w0 = p.f32_w[((size_t)k * 32 + l) * 2], w1 = p.f32_w[((size_t)k * 32 + l) * 2 + 1];
            if (flagged(k, l) && rot_mode) { e[0] = w0; e[2] = w1; } else { e[1] = w0; e[2] = w1; }
            if (w1 != 0.0) {
                const int lp = p.f32_partner[(size_t)k * 32 + l] & 31;
                if (slot_of[lp] != (i ^ (int)m)) return;                          // partner outside the XOR pattern: keep frame32
            }
        }
    }
    for (int i = 0; i < 16; ++i) if (lane_of_slot[i] >= 0) p.f16_init[i] = p.f32_init[lane_of_slot[i]];
    const size_t G = (size_t)std::max(p.n_groups, 1);
    p.f16_eweight_g.assign(G * 17, 0.0);
    for (size_t g = 0; g < G; ++g) {
        for (int i = 0; i < 16; ++i) if (lane_of_slot[i] >= 0) p.f16_eweight_g[g * 17 + i] = p.f32_eweight_g[g * 32 + lane_of_slot[i]];
        double c = 0.0;                                                           // coefficients no rotation touches: constants of the plan
        for (int l = 0; l < 32; ++l) {
            if (p.f32_orig[l] < 0 || slot_of[l] >= 0) continue;
            double v = p.f32_init[l];
            for (int k = 0; k < K; ++k) v *= p.f32_w[((size_t)k * 32 + l) * 2];
            c += p.f32_eweight_g[g * 32 + l] * v;
        }
        p.f16_eweight_g[g * 17 + 16] = c;
    }
    p.f16_lane = lane_of_slot;
    p.f16_n = (int)touched.size();

    if (const_mode) {
        // scaled constant form + correction tables
        p.f16_scaled = 0;
        std::vector<int> stat;
        for (int l = 0; l < 32; ++l) if (p.f32_orig[l] >= 0 && slot_of[l] < 0) stat.push_back(l);
        if (stat.size() > 8) { p.f16_n = 0; return; }
        std::vector<double> lam(16, 1.0), nl(16);
        p.f16_kappa.assign((size_t)K * 16, 0.0);
        for (int k = 0; k < K; ++k) {
            const int m = p.f16_mask[k];
            nl = lam;
            for (int i = 0; i < 16; ++i) {
                const int l = lane_of_slot[i];
                if (l < 0) continue;
                const double w0 = p.f32_w[((size_t)k * 32 + l) * 2], w1 = p.f32_w[((size_t)k * 32 + l) * 2 + 1];
                if (w0 == 0.0) { p.f16_n = 0; return; }
                nl[i] = w0 * lam[i];
                if (m == 0 || w1 == 0.0) continue;
                const int j = i ^ m, lo = i < j ? i : j;
                int pair = 0;
                for (int q = 0; q < lo; ++q) pair += ((q ^ m) > q) ? 1 : 0;
                p.f16_kappa[(size_t)k * 16 + 2 * pair + (i < j ? 0 : 1)] = w1 * lam[j] / (w0 * lam[i]);
            }
            lam = nl;
        }
        p.f16_sweight_g = p.f16_eweight_g;
        for (size_t g = 0; g < G; ++g)
            for (int i = 0; i < 16; ++i) p.f16_sweight_g[g * 17 + i] *= lam[i];
        p.f16_static_lane = stat;
        p.f16_cstat_g.assign(G * 8, 0.0);
        for (size_t g = 0; g < G; ++g)
            for (size_t j = 0; j < stat.size(); ++j) {
                double v = p.f32_init[stat[j]];
                for (int k = 0; k < K; ++k) v *= p.f32_w[((size_t)k * 32 + stat[j]) * 2];
                p.f16_cstat_g[g * 8 + j] = p.f32_eweight_g[g * 32 + stat[j]] * v;
            }
        p.f16_fac_row.assign(p.n_slots, -1);
        size_t rows = 0;
        for (int sl = 0; sl < p.n_slots; ++sl) {
            if (p.slots[sl].seg < 0) continue;
            p.f16_fac_row[sl] = (int)rows;
            rows += p.slots[sl].nq == 1 ? 16 : 256;
        }
        p.f16_fac.assign(rows * 24, 1.0);
        for (int sl = 0; sl < p.n_slots; ++sl) {
            if (p.f16_fac_row[sl] < 0) continue;
            const SlotInfo& si = p.slots[sl];
            const double* t = p.coefs.data() + si.tab_off;
            const int nidx = si.nq == 1 ? 16 : 256;
            for (int v = 1; v < nidx; ++v) {
                if (!((si.diag_ok[v >> 5] >> (v & 31)) & 1u)) continue;
                double* row = &p.f16_fac[((size_t)p.f16_fac_row[sl] + v) * 24];
                for (int c = 0; c < 24; ++c) {
                    const int l = c < 16 ? lane_of_slot[c] : (c - 16 < (int)stat.size() ? stat[c - 16] : -1);
                    if (l < 0) continue;
                    const unsigned lab = p.f_labels[(size_t)sl * SEG_E + p.f32_orig[l]];
                    row[c] = si.nq == 1 ? t[v * 4 + lab] : t[(v >> 4) * 4 + (lab >> 2)] * t[(v & 15) * 4 + (lab & 3u)] * t[64 + lab];
                }
            }
        }
        p.f16_const = 1;
        return;
    }

    // scaled form
    p.f16_scaled = 0;
    bool all_moved = true;
    for (int k = 0; k < K && all_moved; ++k) {
        if (p.f16_mask[k] == 0) all_moved = false;
        for (int i = 0; i < 16 && all_moved; ++i) {
            const int l = lane_of_slot[i];
            if (l >= 0 && (!flagged(k, l) || p.f32_w[((size_t)k * 32 + l) * 2] == 0.0)) all_moved = false;
        }
    }
    if (!all_moved) return;
    std::vector<double> lam(16, 1.0), nl(16);
    p.f16_kappa.assign((size_t)K * 16, 0.0);
    for (int k = 0; k < K; ++k) {
        const int m = p.f16_mask[k];
        int pair = 0;
        for (int i = 0; i < 16; ++i) {
            const int j = i ^ m;
            if (j < i) continue;
            const int ij[2] = {i, j};
            for (int h = 0; h < 2; ++h) {
                const int a = ij[h], b = ij[1 - h], l = lane_of_slot[a];
                nl[a] = lam[a];
                if (l < 0) continue;
                const double wc = p.f32_w[((size_t)k * 32 + l) * 2], w1 = p.f32_w[((size_t)k * 32 + l) * 2 + 1];
                if (w1 != 0.0) p.f16_kappa[(size_t)k * 16 + 2 * pair + h] = w1 * lam[b] / (wc * lam[a]);
                nl[a] = wc * lam[a];
            }
            ++pair;
        }
        lam = nl;
    }
    p.f16_sweight_g = p.f16_eweight_g;
    for (size_t g = 0; g < G; ++g)
        for (int i = 0; i < 16; ++i) p.f16_sweight_g[g * 17 + i] *= lam[i];
    p.f16_scaled = 1;
}

static bool build_frame(Plan& p, std::string& why) {
    const int K = p.nseg;
    std::vector<int> holder(SEG_E), nh(SEG_E);
    for (int i = 0; i < SEG_E; ++i) holder[i] = i;
    std::vector<double> w((size_t)K * SEG_E * 2, 0.0);
    std::vector<unsigned> masks(K, 0u);
    std::vector<uint8_t> labels((size_t)std::max(p.n_slots, 1) * SEG_E, 0);
    auto place_labels = [&](int stage) {
        for (int s = 0; s < p.n_slots; ++s)
            if (p.slots[s].seg == stage)
                for (int i = 0; i < SEG_E; ++i) labels[(size_t)s * SEG_E + holder[i]] = p.slots[s].label[i];
    };
    for (int k = 0; k < K; ++k) {
        const Segment& sg = p.segs[k];
        place_labels(k);
        std::vector<char> used(SEG_E, 0);
        int mask = -1;
        for (int j = 0; j < SEG_E; ++j) {
            int t = holder[sg.src0[j]];
            if (used[t]) { why = "segment primary sources are not a permutation"; return false; }
            used[t] = 1;
            nh[j] = t;
            w[((size_t)k * SEG_E + t) * 2] = sg.w0[j];
            if (sg.active[j] && sg.w1[j] != 0.0) {
                int partner = holder[sg.src1[j]];
                int m = t ^ partner;
                if (mask < 0) mask = m;
                else if (mask != m) { why = "rotation partners are not related by a constant XOR mask"; return false; }
                w[((size_t)k * SEG_E + t) * 2 + 1] = sg.w1[j];
            }
        }
        masks[k] = mask < 0 ? 0u : (unsigned)mask;
        holder = nh;
    }
    place_labels(K);
    // energy weights per thread
    std::vector<double> ew(SEG_E, 0.0);
    for (size_t t = 0; t < p.e_w.size(); ++t) ew[holder[p.e_src[t]]] += p.e_w[t];

    // GF(2) relabelling: three functionals (warp bits) that annihilate as many xor masks as possible
    std::vector<unsigned> chosen;
    {
        std::map<unsigned, int> freq;
        for (unsigned m : masks) if (m) ++freq[m];
        std::vector<std::pair<int, unsigned>> order;
        for (auto& kv : freq) order.push_back({-kv.second, kv.first});
        std::sort(order.begin(), order.end());
        for (auto& o : order) {
            std::vector<unsigned> trial = chosen;
            trial.push_back(o.second);
            if (gf2_rank(trial) <= 5) chosen = trial;
        }
    }
    unsigned rows[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    {
        // annihilator of span(chosen): all f with parity(f & m) == 0 for every chosen m
        std::vector<unsigned> ann;
        for (unsigned f = 1; f < 256; ++f) {
            bool ok = true;
            for (unsigned m : chosen) if (parity8(f & m)) { ok = false; break; }
            if (ok) ann.push_back(f);
        }
        std::vector<unsigned> picked;
        for (unsigned f : ann) {
            std::vector<unsigned> trial = picked;
            trial.push_back(f);
            if (gf2_rank(trial) == (int)trial.size()) picked = trial;
            if (picked.size() == 3) break;
        }
        if (picked.size() < 3) { why = "internal: annihilator too small"; return false; }
        rows[5] = picked[0]; rows[6] = picked[1]; rows[7] = picked[2];
        int filled = 0;
        for (unsigned f = 1; f < 256 && filled < 5; ++f) {
            std::vector<unsigned> trial = {rows[5], rows[6], rows[7]};
            for (int i = 0; i < filled; ++i) trial.push_back(rows[i]);
            trial.push_back(f);
            if (gf2_rank(trial) == (int)trial.size()) rows[filled++] = f;
        }
        if (filled < 5 || !gf2_invertible(rows)) { why = "internal: relabelling not invertible"; return false; }
    }
    auto phys = [&](unsigned t) {
        unsigned pidx = 0;
        for (int r = 0; r < 8; ++r) pidx |= (unsigned)parity8(rows[r] & t) << r;
        return pidx;
    };
    p.fsegs.assign(K, FrameSeg());
    p.f_w.assign((size_t)K * SEG_E * 2, 0.0);
    p.f_init.assign(SEG_E, 0.0);
    p.f_eweight.assign(SEG_E, 0.0);
    p.e_frame.assign(p.e_src.size(), -1);
    p.f_labels.assign((size_t)std::max(p.n_slots, 1) * SEG_E, 0);
    for (unsigned t = 0; t < (unsigned)SEG_E; ++t) {
        unsigned q = phys(t);
        p.f_init[q] = (t & 0xAAu) ? 0.0 : 1.0;
        p.f_eweight[q] = ew[t];
        for (size_t te = 0; te < p.e_src.size(); ++te) if ((unsigned)holder[p.e_src[te]] == t) p.e_frame[te] = (int)q;
        for (int k = 0; k < K; ++k) {
            p.f_w[((size_t)k * SEG_E + q) * 2] = w[((size_t)k * SEG_E + t) * 2];
            p.f_w[((size_t)k * SEG_E + q) * 2 + 1] = w[((size_t)k * SEG_E + t) * 2 + 1];
        }
        for (int s = 0; s < p.n_slots; ++s) p.f_labels[(size_t)s * SEG_E + q] = labels[(size_t)s * SEG_E + t];
    }
    // rotation classes: (param, scale, offset); with offset == 0 the sign of scale folds into w1 (sin is odd)
    p.f_classes.clear();
    for (int k = 0; k < K; ++k) {
        FrameSeg& fs = p.fsegs[k];
        fs.xor_mask = (int)phys(masks[k]);     // linear map: phys(t ^ m) = phys(t) ^ phys(m)
        fs.use_shfl = (fs.xor_mask >> 5) == 0;
        const int rid = p.segs[k].rot_id;
        if (rid < 0) { fs.cs_sel = -2; continue; }
        RotInfo ri = p.rots[rid];
        double sign = 1.0;
        if (ri.offset == 0.0 && ri.scale < 0.0) { sign = -1.0; ri.scale = -ri.scale; }
        int cls = -1;
        for (size_t c = 0; c < p.f_classes.size(); ++c)
            if (p.f_classes[c].param == ri.param && p.f_classes[c].scale == ri.scale && p.f_classes[c].offset == ri.offset) cls = (int)c;
        if (cls < 0) { cls = (int)p.f_classes.size(); p.f_classes.push_back(ri); }
        fs.cs_sel = cls;
        if (sign < 0)
            for (int q = 0; q < SEG_E; ++q) p.f_w[((size_t)k * SEG_E + q) * 2 + 1] = -p.f_w[((size_t)k * SEG_E + q) * 2 + 1];
    }
    p.f_single_class = p.f_classes.size() <= 1;
    build_frame32(p);
    {   // per-group energy weights: linear in the coefficients, same routing as the base weights
        const size_t T = p.term_idx.size(), G = (size_t)std::max(p.n_groups, 1);
        p.f_eweight_g.assign(G * SEG_E, 0.0);
        p.f32_eweight_g.assign(G * 32, 0.0);
        for (size_t g = 0; g < G; ++g) {
            for (size_t t = 0; t < T; ++t)
                if (p.e_frame[t] >= 0) p.f_eweight_g[g * SEG_E + p.e_frame[t]] += p.term_coef_g[g * T + t] * p.e_fac[t];
            if (p.f32_n > 0)
                for (int l = 0; l < 32; ++l)
                    if (p.f32_orig[l] >= 0) p.f32_eweight_g[g * 32 + l] = p.f_eweight_g[g * SEG_E + p.f32_orig[l]];
        }
    }
    build_frame16t(p);
    return true;
}

// ---- top level ------------------------------------------------------------------------------------------

void lower_plan(Plan& p) {
    if (p.n < 1 || p.n > DMX_MAX_QUBITS) throw std::runtime_error("n_qubits out of range");
    if (const char* env = std::getenv("DMX_WIDE")) p.opt_wide = std::atoi(env) != 0;      // test hook: whole-suite A/B of the wide builder
    p.blocks.clear(); p.rots.clear(); p.mus.clear(); p.sweeps.clear(); p.factors.clear(); p.chains.clear(); p.phases.clear();
    p.coefs.clear(); p.passes.clear();
    p.remapped = false; p.triples = false; p.sweeps3.clear(); p.f32_n = 0; p.f16_n = 0;
    p.cpb = plan_cpb(p.n);
    p.segs.clear(); p.slots.clear(); p.term_idx.clear(); p.term_coef.clear();
    std::memset(&p.info, 0, sizeof(p.info));
    lower_ops(p);
    if (p.opt_fuse) fuse_blocks(p); else drop_identities(p);
    int rid = 0;
    for (Block& b : p.blocks)
        if (b.kind == BK_ROT) {
            b.rot_id = rid++;
            p.rots.push_back(RotInfo{b.param, 0, b.scale, b.offset});
        }
    // merge Hamiltonian terms onto engine indices
    {
        std::map<uint32_t, double> acc;
        std::vector<uint32_t> order;
        for (size_t t = 0; t < p.hc.size(); ++t) {
            uint32_t idx = pauli_index(p.n, p.hx[t], p.hz[t]);
            if (!acc.count(idx)) order.push_back(idx);
            acc[idx] += p.hc[t];
        }
        for (uint32_t idx : order) { p.term_idx.push_back(idx); p.term_coef.push_back(acc[idx]); }
        // coefficient groups: the same merge per row.  For a grouped plan the base coefficients only decide which Pauli
        // coefficients are live (frame32 compaction), so they become sum_g |c_g| - non-zero wherever any row is.
        const size_t T = p.term_idx.size(), G = (size_t)std::max(p.n_groups, 1);
        p.term_coef_g.assign(G * T, 0.0);
        if (p.n_groups > 0) {
            std::map<uint32_t, size_t> at;
            for (size_t t = 0; t < T; ++t) at[p.term_idx[t]] = t;
            for (size_t g = 0; g < G; ++g)
                for (size_t t = 0; t < p.hc.size(); ++t)
                    p.term_coef_g[g * T + at[pauli_index(p.n, p.hx[t], p.hz[t])]] += p.hgroups[g * p.hc.size() + t];
            for (size_t t = 0; t < T; ++t) {
                double a = 0.0;
                for (size_t g = 0; g < G; ++g) a += std::fabs(p.term_coef_g[g * T + t]);
                p.term_coef[t] = a;
            }
        } else {
            p.term_coef_g = p.term_coef;
        }
    }
    if (p.n <= 7) schedule_single_pass(p);
    else {
        bool remapped = false;
        if (p.opt_remap && p.opt_tile_qubits >= 6) {
            bool has_var = false;
            for (const Block& b : p.blocks) has_var = has_var || b.kind == BK_VAR;
            p.triples = p.opt_triples && !has_var;
            try {
                schedule_streaming_remap(p);
                remapped = true;
            } catch (const std::runtime_error& e) {
                remapped = false;      // e.g. a builder limit: fall back to the pinned-qubit scheduler below
                if (std::getenv("DMX_VERBOSE")) std::fprintf(stderr, "[dmx] remapping scheduler gave up: %s\n", e.what());
            }
            for (size_t pi = 0; pi < p.passes.size(); ++pi) {
                const Pass& ps = p.passes[pi];
                const bool ok = ps.n_phases == 0 || pass_stream_eligible(p, ps);
                if (!ok && remapped && std::getenv("DMX_VERBOSE")) std::fprintf(stderr, "[dmx] pass %zu is not eligible for the stream kernel\n", pi);
                remapped = remapped && ok;
            }
            if (!remapped) {   // the generic tile kernel only knows the standard layout: re-plan with pinned low qubits
                p.mus.clear(); p.sweeps.clear(); p.sweeps3.clear(); p.factors.clear(); p.chains.clear(); p.phases.clear(); p.coefs.clear(); p.passes.clear();
                p.triples = false;
            }
        }
        if (!remapped) schedule_streaming(p);
        p.remapped = remapped;
    }
    const double tile_flops = p.info.flops_per_circuit, tile_smem = p.info.smem_bytes_per_circuit;
    p.path = p.n <= 7 ? 0 : 2;
    p.seg_eligible = false;
    if (p.opt_segments && p.n <= SEG_N) {
        p.seg_eligible = build_segments(p, p.seg_reason);
        if (p.seg_eligible) p.seg_eligible = build_frame(p, p.seg_reason);
        if (p.seg_eligible) p.path = 1;
    } else if (p.n <= SEG_N) {
        p.seg_reason = "disabled by option";
    }
    // cost model
    const double N = std::pow(4.0, p.n);
    dmx_plan_info& in = p.info;
    in.n_qubits = p.n; in.n_params = p.n_params; in.n_slots = p.n_slots; in.n_ops_in = (int)p.ops.size();
    in.n_blocks = (int)p.blocks.size();
    in.n_segments = p.seg_eligible ? p.nseg : 0;
    in.n_passes = (int)p.passes.size();
    in.tile_qubits = p.tile_k;
    in.path = p.path;
    in.n_terms = (int)p.term_idx.size();
    in.state_bytes = (int64_t)(8.0 * N);
    in.workspace_bytes_per_state = p.path == 2 ? in.state_bytes * (p.remapped ? 2 : 1) : 0;
    const double io = 8.0 * p.n_params + (double)p.n_slots + 8.0;
    if (p.path == 1) {
        int smem_segs = 0;
        for (const FrameSeg& fs : p.fsegs) smem_segs += !fs.use_shfl;
        // the compact frame kernel executes the segments on 32 lanes per circuit instead of 256 threads
        in.flops_per_circuit = (p.nseg * 4.0 + 1.0) * (p.f32_n > 0 ? 32.0 : (double)SEG_E) + 2.0 * p.e_w.size();
        // thread-per-circuit form: 16 slots x (fma + 2 mul + fma = 6 flops) per segment, 16 fma for the energy
        if (p.f16_n > 0 && p.opt_frame16t) in.flops_per_circuit = p.nseg * (p.f16_scaled && p.opt_f16_scaled ? 64.0 : 96.0) + 32.0;
        // thread-per-row variant form: one fma per slot and segment, 24 fma for the energy (corrections: a few rows in a hundred)
        if (p.f16_n > 0 && p.opt_frame16t && p.f16_const) in.flops_per_circuit = p.nseg * 32.0 + 48.0;
        // per circuit: smem-exchange segments move 8 B out + 8 B in per coefficient; the energy stage a few terms
        in.smem_bytes_per_circuit = smem_segs * SEG_E * 16.0 + 16.0 * p.e_w.size();
        in.hbm_bytes_per_circuit = io;
    } else {
        in.flops_per_circuit = tile_flops + 2.0 * p.term_idx.size();
        in.smem_bytes_per_circuit = tile_smem;
        // streaming: every pass reads and writes the state once, except the first (init in smem, no read)
        in.hbm_bytes_per_circuit = io + (p.path == 2 ? (2.0 * p.passes.size() - 1.0) * 8.0 * N : 0.0);
    }
    in.canonical_flops_per_circuit += 0.0;
    {
        // canonical energy epilogue: 8*G*2^n + 2*T*2^n with G distinct x-masks
        std::map<uint32_t, int> xs;
        for (uint32_t x : p.hx) xs[x] = 1;
        in.canonical_flops_per_circuit += (8.0 * xs.size() + 2.0 * p.hc.size()) * std::pow(2.0, p.n);
        in.canonical_hbm_bytes_per_circuit = io + (p.n > 7 ? 32.0 * N * p.passes.size() : 0.0);
    }
    p.finalized = true;
}

static void dump_vec(std::ostringstream& os, const char* name, const std::vector<double>& v) {
    os << " " << name << "=";
    char buf[40];
    for (size_t i = 0; i < v.size(); ++i) {
        std::snprintf(buf, sizeof buf, "%.17g", v[i]);
        os << (i ? "," : "") << buf;
    }
}

template <class T>
static void dump_ints(std::ostringstream& os, const char* name, const std::vector<T>& v) {
    os << " " << name << "=";
    for (size_t i = 0; i < v.size(); ++i) os << (i ? "," : "") << (long long)v[i];
}

std::string dump_plan(const Plan& p) {
    std::ostringstream os;
    os << "plan n=" << p.n << " params=" << p.n_params << " slots=" << p.n_slots << " path=" << p.path
       << " tile_k=" << p.tile_k << " cpb=" << p.cpb << " blocks=" << p.blocks.size() << " passes=" << p.passes.size()
       << " nseg=" << (p.seg_eligible ? p.nseg : 0) << " seg_reason=\"" << p.seg_reason << "\"\n";
    for (size_t i = 0; i < p.blocks.size(); ++i) {
        const Block& b = p.blocks[i];
        os << "block " << i << " kind=" << b.kind << " nq=" << b.nq << " q0=" << b.q[0] << " q1=" << b.q[1]
           << " cls=" << b.cls;
        if (b.kind == BK_FIXED) dump_vec(os, "M", b.M);
        if (b.kind == BK_ROT) {
            os << " rot=" << b.rot_id << " param=" << b.param;
            dump_vec(os, "scale", {b.scale}); dump_vec(os, "offset", {b.offset});
            dump_vec(os, "A", b.A); dump_vec(os, "C", b.C); dump_vec(os, "S", b.S);
        }
        if (b.kind == BK_VAR) { os << " slot=" << b.slot; dump_vec(os, "tab", b.tab); dump_vec(os, "cond", b.cond); }
        os << "\n";
    }
    for (size_t i = 0; i < p.passes.size(); ++i) {
        const Pass& ps = p.passes[i];
        os << "pass " << i << " first_phase=" << ps.first_phase << " nphases=" << ps.n_phases;
        dump_ints(os, "tile", ps.tile_q);
        if (!ps.pos_in.empty()) { dump_ints(os, "pos_in", ps.pos_in); dump_ints(os, "pos_out", ps.pos_out); }
        os << "\n";
    }
    for (size_t i = 0; i < p.phases.size(); ++i)
        os << "phase " << i << " first_sweep=" << p.phases[i].first_sweep << " nsweeps=" << p.phases[i].n_sweeps
           << " first_chain=" << p.phases[i].first_chain << " nchains=" << p.phases[i].n_chains << "\n";
    for (size_t i = 0; i < p.sweeps.size(); ++i) {
        const SweepDev& w = p.sweeps[i];
        os << "sweep " << i << " p0=" << w.p0 << " p1=" << w.p1 << " first_mu=" << w.first_mu << " nmu=" << w.n_mu
           << " lperm=" << w.load_perm << " lcoef=" << w.lcoef << " sperm=" << w.store_perm << " scoef=" << w.scoef
           << " lmask=" << w.lmask[0] << "," << w.lmask[1] << "," << w.lmask[2] << "," << w.lmask[3]
           << " smask=" << w.smask[0] << "," << w.smask[1] << "," << w.smask[2] << "," << w.smask[3] << "\n";
    }
    for (size_t i = 0; i < p.sweeps3.size(); ++i) {
        const Sweep3& w = p.sweeps3[i];
        os << "sweep3 " << i << " p0=" << w.p[0] << " p1=" << w.p[1] << " p2=" << w.p[2] << " first_mu=" << w.first_mu << " nmu=" << w.n_mu
           << " lperm=" << w.load_perm << " lcoef=" << w.lcoef << " sperm=" << w.store_perm << " scoef=" << w.scoef << " lmask=";
        for (int j = 0; j < 6; ++j) os << (j ? "," : "") << w.lmask[j];
        os << " smask=";
        for (int j = 0; j < 6; ++j) os << (j ? "," : "") << w.smask[j];
        if (w.wide) {
            os << " wide=" << w.wide << " lnvar=" << w.l_nvar << " snvar=" << w.s_nvar << " ltmask=";
            for (int j = 0; j < 8; ++j) os << (j ? "," : "") << w.ltmask[j];
            os << " stmask=";
            for (int j = 0; j < 8; ++j) os << (j ? "," : "") << w.stmask[j];
            os << " lvmask=";
            for (int j = 0; j < 4; ++j) os << (j ? "," : "") << w.lvmask[j];
            os << " svmask=";
            for (int j = 0; j < 4; ++j) os << (j ? "," : "") << w.svmask[j];
        }
        os << "\n";
    }
    for (size_t i = 0; i < p.mus.size(); ++i)
        os << "mu " << i << " kind=" << p.mus[i].kind << " a0=" << p.mus[i].a0 << " a1=" << p.mus[i].a1 << " a2=" << p.mus[i].a2
           << " a3=" << p.mus[i].a3 << "\n";
    for (size_t i = 0; i < p.chains.size(); ++i)
        os << "chain " << i << " first=" << p.chains[i].first_factor << " n=" << p.chains[i].n_factors << "\n";
    for (size_t i = 0; i < p.factors.size(); ++i)
        os << "factor " << i << " kind=" << p.factors[i].kind << " coef=" << p.factors[i].coef << " rot=" << p.factors[i].rot_id << "\n";
    dump_vec(os, "coefs", p.coefs);
    os << "\n";
    {
        std::vector<long long> ti(p.term_idx.begin(), p.term_idx.end());
        os << "terms";
        dump_ints(os, "idx", ti);
        dump_vec(os, "coef", p.term_coef);
        os << "\n";
    }
    if (p.seg_eligible) {
        for (int k = 0; k < p.nseg; ++k) {
            const Segment& s = p.segs[k];
            os << "seg " << k << " rot=" << s.rot_id;
            dump_ints(os, "src0", s.src0); dump_ints(os, "src1", s.src1); dump_ints(os, "active", s.active);
            dump_vec(os, "w0", s.w0); dump_vec(os, "w1", s.w1);
            os << "\n";
        }
        for (int k = 0; k < p.nseg; ++k) {
            os << "fseg " << k << " mask=" << p.fsegs[k].xor_mask << " shfl=" << p.fsegs[k].use_shfl << " cs=" << p.fsegs[k].cs_sel;
            std::vector<double> wk(p.f_w.begin() + (size_t)k * 512, p.f_w.begin() + (size_t)(k + 1) * 512);
            dump_vec(os, "w", wk);
            os << "\n";
        }
        os << "frame single_class=" << p.f_single_class;
        dump_vec(os, "init", p.f_init);
        dump_vec(os, "eweight", p.f_eweight);
        os << "\n";
        if (p.f32_n > 0) {
            os << "f32 n=" << p.f32_n;
            dump_ints(os, "orig", p.f32_orig);
            dump_vec(os, "init", p.f32_init);
            dump_vec(os, "eweight", p.f32_eweight);
            os << "\n";
            for (int k = 0; k < p.nseg; ++k) {
                os << "f32seg " << k;
                dump_ints(os, "partner", std::vector<int>(p.f32_partner.begin() + (size_t)k * 32, p.f32_partner.begin() + (size_t)(k + 1) * 32));
                dump_vec(os, "w", std::vector<double>(p.f32_w.begin() + (size_t)k * 64, p.f32_w.begin() + (size_t)(k + 1) * 64));
                os << "\n";
            }
        }
        if (p.f16_n > 0) {
            os << "f16 n=" << p.f16_n << " scaled=" << p.f16_scaled << " const=" << p.f16_const;
            dump_ints(os, "lane", p.f16_lane);
            dump_ints(os, "mask", std::vector<int>(p.f16_mask.begin(), p.f16_mask.end()));
            dump_vec(os, "init", p.f16_init);
            dump_vec(os, "eweight", std::vector<double>(p.f16_eweight_g.begin(), p.f16_eweight_g.begin() + 17));
            if (p.f16_scaled || p.f16_const) dump_vec(os, "sweight", std::vector<double>(p.f16_sweight_g.begin(), p.f16_sweight_g.begin() + 17));
            if (p.f16_const) {
                dump_ints(os, "static", p.f16_static_lane);
                dump_vec(os, "cstat", std::vector<double>(p.f16_cstat_g.begin(), p.f16_cstat_g.begin() + 8));
                dump_ints(os, "facrow", p.f16_fac_row);
            }
            os << "\n";
            for (int k = 0; k < p.nseg; ++k) {
                os << "f16seg " << k;
                dump_vec(os, "tab", std::vector<double>(p.f16_tab.begin() + (size_t)k * 64, p.f16_tab.begin() + (size_t)(k + 1) * 64));
                if (p.f16_scaled || p.f16_const) dump_vec(os, "kappa", std::vector<double>(p.f16_kappa.begin() + (size_t)k * 16, p.f16_kappa.begin() + (size_t)(k + 1) * 16));
                os << "\n";
            }
        }
        for (size_t c = 0; c < p.f_classes.size(); ++c) {
            os << "fclass " << c << " param=" << p.f_classes[c].param;
            dump_vec(os, "scale", {p.f_classes[c].scale}); dump_vec(os, "offset", {p.f_classes[c].offset});
            os << "\n";
        }
        for (int sl = 0; sl < p.n_slots; ++sl) {
            os << "flabel " << sl;
            std::vector<uint8_t> lb(p.f_labels.begin() + (size_t)sl * 256, p.f_labels.begin() + (size_t)(sl + 1) * 256);
            dump_ints(os, "label", lb);
            os << "\n";
        }
        os << "estage";
        dump_ints(os, "src", p.e_src);
        dump_vec(os, "w", p.e_w);
        os << "\n";
        for (int s = 0; s < p.n_slots; ++s) {
            const SlotInfo& si = p.slots[s];
            os << "slot " << s << " nq=" << si.nq << " seg=" << si.seg << " tab=" << si.tab_off;
            std::vector<long long> ok(si.diag_ok, si.diag_ok + 8);
            dump_ints(os, "diag_ok", ok);
            dump_ints(os, "label", si.label);
            os << "\n";
        }
    }
    for (size_t i = 0; i < p.rots.size(); ++i) {
        os << "rot " << i << " param=" << p.rots[i].param;
        dump_vec(os, "scale", {p.rots[i].scale}); dump_vec(os, "offset", {p.rots[i].offset});
        os << "\n";
    }
    return os.str();
}

}  // namespace dmx
