// dmx_plan.hpp — internal plan structures shared by the host-side lowering (dmx_lower.cpp) and the
// CUDA execution layer (dmx_exec.cu).  Not part of the public ABI (include/dmx.h is).
//
// Representation.  The engine does not store rho as a 2^n x 2^n complex matrix.  It stores the
// Pauli-coefficient vector r[P] = Tr(rho P) over all 4^n Pauli strings: 4^n REAL doubles — exactly the
// degrees of freedom of a Hermitian rho, i.e. half the bytes of complex128 rho (SURVEY.md App. E.9 notes
// that every map on the path preserves Hermiticity).  In this basis
//   * U rho U^dagger and sum_k K rho K^dagger on one/two qubits are real 4x4 / 16x16 matrices acting on
//     the 2-bit Pauli code(s) of those qubits (the Pauli transfer matrix),
//   * every Pauli-type channel of the reference (depolarize, bit/phase flip, asymmetric Pauli, two-qubit
//     depolarizing, product Pauli) is DIAGONAL, every Clifford gate (H, CNOT, X, rx(+-pi/2), rx(pi)) is a
//     signed permutation, so gate+noise blocks are "monomial" (one non-zero per row),
//   * Tr(rho H) = sum_t c_t r[P_t] is a gather-dot over the Hamiltonian's terms.
// Index layout: qubit q owns bits [2(n-1-q)+1 : 2(n-1-q)] = (x,z); code = 2x+z: 0=I 1=Z 2=X 3=Y.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/dmx.h"

namespace dmx {

enum BlockKind : int { BK_FIXED = 0, BK_ROT = 1, BK_VAR = 2 };
enum MatClass : int { CLS_IDENT = 0, CLS_DIAG = 1, CLS_MONO = 2, CLS_DENSE = 3 };

// One lowered micro-op on 1 or 2 qubits.  dim = 4 (1q) or 16 (2q); local index = code(q[0])*4 + code(q[1]).
struct Block {
    int kind = BK_FIXED;
    int nq = 1;
    int q[2] = {0, 0};
    int dim() const { return nq == 1 ? 4 : 16; }
    // FIXED: M (dim*dim row-major).  ROT: M(theta) = A + cos(theta) C + sin(theta) S.
    std::vector<double> M, A, C, S;
    int cls = CLS_DENSE;
    // ROT
    int param = -1;
    double scale = 0.0, offset = 0.0;
    int rot_id = -1;
    // VAR: slot, tab = 16 one-qubit PTMs (16*16 doubles) of the pure basis ops, cond = PTM (dim*dim) of the
    // ops applied only when idx != 0 (identity when there are none).
    int slot = -1;
    std::vector<double> tab, cond;
    int cond_cls = CLS_IDENT;
    bool dead = false;
    int src_op = -1;  // index of the first source op (for diagnostics)
};

// ---- device-facing tables -------------------------------------------------------------------------

// A tile pass is executed as a list of SWEEPS.  One sweep walks the tile once: every thread loads a group of 16
// coefficients (all codes of a qubit pair (A,B), the other codes fixed) into registers, applies a list of
// micro-ops to it and stores it back — so several fused blocks cost ONE shared-memory round trip.  A monomial
// two-qubit block (Clifford gate x Pauli noise) at the start / end of a sweep is folded into the load / store
// addressing (permutation) plus 16 scales.
enum MuKind : int32_t {
    MU_CHAIN_A = 0, MU_CHAIN_B = 1,  // per-circuit 4x4 from the chain table (product of rotations / fixed 1q blocks)
    MU_FIX_A = 2, MU_FIX_B = 3,      // fixed 4x4 from the coefficient table
    MU_DIAG_A = 4, MU_DIAG_B = 5,    // 4 scales
    MU_DIAG2 = 6,                    // 16 scales
    MU_MAT2 = 7,                     // dense 16x16
    MU_VAR_A = 8, MU_VAR_B = 9,      // variant slot on one qubit: 16 precomposed 4x4 (noise-after-correction included)
    MU_VAR2 = 10                     // variant slot on the pair: kron(B[a], B[b]) then the conditional block
};

struct MicroOp {      // 32 bytes
    int32_t kind;
    int32_t a0;       // CHAIN: slot in the phase's chain table; FIX/DIAG/MAT2: coefficient offset; VAR: variant slot
    int32_t a1;       // VAR: coefficient offset of the matrix table
    int32_t a2;       // VAR2: coefficient offset of the conditional block
    int32_t a3;       // VAR2: class of the conditional block (CLS_DIAG / CLS_DENSE); CHAIN/FIX: 1 = unital 3x3 form
    int32_t pad[3];
};

// Register e of a group holds local code e = 4*code(A) + code(B); its bits are (zB, xB, zA, xA).  Clifford
// permutations are GF(2)-linear on those 4 bits, so the tile index of register e is base ^ XOR_{j in bits(e)} mask[j]:
// identity addressing and any monomial block use the same code path (4 masks), one LOP3 per element.
struct SweepDev {     // 64 bytes
    int32_t p0, p1;           // tile-local bit positions of the codes of qubit A and qubit B
    int32_t first_mu, n_mu;
    int32_t load_perm, lcoef; // load:  v[e] = coef[lcoef + e] * tile[base ^ X_l(e)]      (scale skipped when !load_perm)
    int32_t store_perm, scoef;// store: tile[base ^ X_s(e)] = coef[scoef + e] * v[e]
    uint32_t lmask[4], smask[4];
};

// Triple sweeps (streaming plans executed by dmx_stream3_kernel): a thread holds all 64 codes of THREE tile qubits
// (slots A, B, C; register e = 16*code(A) + 4*code(B) + code(C), bits zC,xC,zB,xB,zA,xA), so a CNOT chain advances two
// links per shared-memory round trip.  Leading / trailing monomial blocks (Clifford gate x Pauli noise, one- or
// two-qubit) are composed into ONE load / store permutation (6 XOR masks + 64 scales).
enum Mu3Kind : int32_t {
    MU3_MAT = 16,    // 4x4 on slot a1: a2 = 1 -> per-circuit chain a0 (phase-relative), else fixed coefficients at a0; a3 = unital
    MU3_DIAG = 17,   // 4 scales at a0 on slot a1
    MU3_DIAG2 = 18,  // 16 scales at a0 on slot pair a1 (0: AB, 1: AC, 2: BC), index = 4*code(first) + code(second)
    MU3_MAT2 = 19    // dense 16x16 at a0 on slot pair a1
};

struct Sweep3 {
    int32_t p[3];             // tile-local bit positions of the codes of slots A, B, C
    int32_t first_mu, n_mu;
    int32_t load_perm, lcoef; // load:  v[e] = coef[lcoef + e] * tile[base ^ X_l(e)]   (64 scales; skipped when !load_perm)
    int32_t store_perm, scoef;// store: tile[base ^ X_s(e)] = coef[scoef + e] * v[e]
    uint32_t lmask[6], smask[6];
    // "Wide" monomials (plan option "wide"): leading / trailing Clifford x Pauli-noise blocks on ANY qubits of the tile, not
    // only on the register triple.  Their GF(2)-linear permutation rides on the exchange addressing - the thread's base
    // becomes a linear image of its index (tmask) and the register masks may reach outside the triple - and their diagonal
    // part becomes a 64-entry scale table that depends on the thread only through a few parities of its index (variants).
    //   load : v[e] = coef[lcoef + 64 * var_l(t) + e] * tile[T_l(t) ^ X_l(e)],  T_l(t) = XOR_j t_j * ltmask[j]
    //   store: tile[T_s(t) ^ X_s(e)] = coef[scoef + 64 * var_s(t) + e] * v[e],  var(t) = sum_j parity(t & vmask[j]) << j
    // t = the thread's values of the non-triple tile bits in ascending position.  A wide sweep reads and writes different
    // address sets, so the kernel puts a block barrier between its loads and its stores (bit 2).
    int32_t wide = 0;         // bit 0: load side, bit 1: store side, bit 2: barrier between loads and stores
    int32_t l_nvar = 0, s_nvar = 0;
    uint32_t ltmask[8] = {}, stmask[8] = {};
    uint32_t lvmask[4] = {}, svmask[4] = {};
};

struct ChainFactor {  // one factor of a per-circuit 1-qubit matrix product (applied in order)
    int32_t kind;     // 0: fixed (16 coefficients), 1: rotation (48 coefficients A, C, S)
    int32_t coef;
    int32_t rot_id;   // global rotation id (kind 1)
    int32_t pad;
};

struct ChainDev {
    int32_t first_factor, n_factors;
};

struct Phase {        // sweeps that share one chain table (bounded shared memory)
    int32_t first_sweep, n_sweeps, first_chain, n_chains;
};

struct RotInfo {      // per global rotation id
    int32_t param;
    int32_t pad;
    double scale, offset;
};

struct Pass {         // one pass over the state with `tile_q` resident
    std::vector<int> tile_q;          // resident qubits by DESCENDING tile-local position (tile_q[i] sits at bits 2*(k-1-i))
    int first_phase = 0, n_phases = 0;
    // Global layout of the state in HBM before (pos_in) and after (pos_out) this pass: pos[q] = position of qubit q
    // (its code occupies index bits 2*pos, 2*pos+1).  Empty = standard order pos[q] = n-1-q.  The remapping scheduler
    // lets every pass store the state in the order the NEXT pass wants to load it: only two qubits shared by
    // consecutive passes have to sit at the low positions (128-byte runs), instead of pinning qubits n-1, n-2 forever.
    std::vector<int> pos_in, pos_out;
};

// Segment kernel tables (n padded to 4 qubits: 256 elements).
struct Segment {
    std::vector<uint8_t> src0, src1, active;  // 256 each
    std::vector<double> w0, w1;               // 256 each
    int rot_id = -1;                          // -1: pure monomial (no local op)
};

// Fixed-frame form of the segment tables (what the kernel executes).  Thread t keeps "its" Pauli coefficient
// for the whole circuit (the permutations of the Clifford gates are folded into a relabelling), so a segment is
//   r[t] <- cm * w0[t] * r[t] + sm * w1[t] * r[t ^ xor_mask]
// with (cm, sm) = (cos, sin) of the segment's rotation class, (1, 1) for constant blocks.  Threads are relabelled
// by a GF(2)-linear map so that as many xor masks as possible stay inside a warp (register shuffle, no smem).
struct FrameSeg {
    int xor_mask = 0;   // physical thread index mask
    int use_shfl = 1;   // mask has no warp bits -> __shfl_xor_sync; else shared-memory exchange
    int cs_sel = -2;    // >= 0: rotation class index; -2: constant coefficients (cm = sm = 1)
};

struct SlotInfo {
    int nq = 1;
    int seg = 0;                  // segment whose INPUT index space the labels refer to (n_segments = energy stage)
    std::vector<uint8_t> label;   // 256: local code (1q) or code0*4+code1 (2q) per input element
    int tab_off = 0;              // coefficient offset: 1q -> dtab[16][4] (diag of T[idx]); 2q -> d1[16][4] then dcond[16]
    uint32_t diag_ok[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // bit idx set if T[idx] is diagonal (256 bits)
};

struct Plan {
    // inputs
    int n = 0, n_params = 0, n_slots = 0;
    std::vector<dmx_op> ops;
    std::vector<double> pool;
    std::vector<uint32_t> hx, hz;
    std::vector<double> hc;
    // Coefficient groups (dmx_plan_set_hamiltonian_groups): the same Pauli strings with n_groups different weight rows, e.g.
    // one row per bond length; circuit b of a grouped call takes row group[b].  Input order [n_groups][hc.size()].
    int n_groups = 0;
    std::vector<double> hgroups;
    int opt_fuse = 1, opt_segments = 1, opt_tile_qubits = 6, opt_remap = 1, opt_triples = 1;
    // execution switches (read at launch time from THIS plan; they used to be process-wide statics)
    int opt_stream_kernel = 1;          // 0 -> generic dmx_tile_kernel for every streaming pass
    long long opt_stream_chunk = 0;     // circuits per streaming pass launch (0: automatic)
    int opt_host_zero_copy = 1;         // dmx_energy_batch_host works in place on mapped page-locked buffers
    int opt_host_streams = 4;           // dmx_energy_batch_host_async: streams consecutive in-flight batches rotate over
    long long opt_host_chunk_rows = 32768;
    int opt_frame32 = 1;                // compact warp-per-circuit frame kernel when <= 32 coefficients are live
    int opt_frame32s = 1;               // one-rotation-class parameter sweeps: 32 circuits per warp (dmx_frame32s_kernel)
    int opt_seg_cpb = 8;                // circuits per warp / thread of the frame kernels
    int opt_bulk = 1;                   // streaming tile load: 0 register staging, 1 cp.async.bulk + mbarrier, 2 direct global loads
    int opt_wide = 0;                   // streaming triple sweeps: absorb tile-wide leading / trailing monomial blocks (Sweep3::wide)
    int opt_scale_smem = 0;             // streaming: permutation scale tables in shared memory (LDS.128) instead of the constant bank
    int opt_tile_threads = 512;         // whole-state 7-qubit tiles: 512 threads x 2 groups or 1024 threads x 1 group per sweep
    int opt_direct_store = 1;           // streaming: the last sweep of a pass writes its registers straight to global memory
    bool finalized = false;

    // lowered program
    std::vector<Block> blocks;
    std::vector<RotInfo> rots;
    std::vector<MicroOp> mus;
    std::vector<SweepDev> sweeps;
    std::vector<Sweep3> sweeps3;    // used instead of `sweeps` when `triples` (Phase::first_sweep indexes this table)
    bool triples = false;
    std::vector<ChainFactor> factors;
    std::vector<ChainDev> chains;
    std::vector<Phase> phases;
    int cpb = 1;            // circuits per CTA when the whole state is resident (n <= 5)
    std::vector<double> coefs;
    std::vector<Pass> passes;
    int path = 0;           // 0 tile (single pass, state in smem), 1 segment, 2 streaming
    bool remapped = false;  // streaming passes re-order the qubits in HBM (Pass::pos_in/pos_out): two state buffers
    int tile_k = 0;
    // energy terms in the engine's index space
    std::vector<uint32_t> term_idx;
    std::vector<double> term_coef;
    std::vector<double> term_coef_g;  // [max(n_groups, 1)][term_idx.size()]: what the tile / energy kernels read
    std::vector<double> e_fac;        // segment path: e_w[t] = term_coef[t] * e_fac[t]
    std::vector<int> e_frame;         // segment path: frame index (physical thread order) that holds term t at the energy stage
    std::vector<double> f_eweight_g;  // [max(n_groups, 1)][256]
    std::vector<double> f32_eweight_g;// [max(n_groups, 1)][32]
    // segment path
    int nseg = 0;
    std::vector<Segment> segs;
    std::vector<uint8_t> e_src;   // energy stage: E = sum_t e_w[t] * r[e_src[t]]
    std::vector<double> e_w;
    std::vector<SlotInfo> slots;
    bool seg_eligible = false;
    std::string seg_reason;
    // fixed-frame tables (physical thread order)
    std::vector<FrameSeg> fsegs;
    std::vector<double> f_w;          // [nseg][256][2] = (w0, w1)
    std::vector<double> f_init;       // [256]
    std::vector<double> f_eweight;    // [256] energy weight of the coefficient held by thread t
    std::vector<uint8_t> f_labels;    // [n_slots][256]
    std::vector<RotInfo> f_classes;   // distinct (param, scale, offset) rotation classes
    int f_single_class = 0;           // 1: every rotation segment uses class 0 (kept in registers)
    // Compact frame ("frame32"): only the Pauli coefficients that are both reachable from |0..0> and able to influence a
    // Hamiltonian term are ever non-trivial (noisy H2 UCCSD: 23 of 256).  When at most 32 are, one WARP carries a circuit:
    // lane l owns coefficient f32_orig[l]; the partner of a segment comes from __shfl_sync(r, f32_partner[k][l]).
    int f32_n = 0;                    // live coefficients (0: compact form not available)
    std::vector<int> f32_orig;        // [32] frame index of the lane's coefficient (-1: idle lane)
    std::vector<uint8_t> f32_partner; // [nseg][32]: partner lane | 0x80 when the segment's rotation moves the coefficient
    std::vector<double> f32_w;        // [nseg][32][2]
    std::vector<double> f32_init, f32_eweight;   // [32]

    // Thread-per-circuit form of the compact frame ("frame16t", dmx_frame16t_kernel): one rotation class, at most 16
    // coefficients that any rotation ever touches.  A rotation pairs coefficient t with t ^ xor_mask in the frame index
    // space, so the touched coefficients sit in a few cosets of span{xor masks} (GF(2) rank <= 4) and can be numbered
    // 0..15 such that every segment pairs slot i with slot i ^ f16_mask[k]: a compile-time register pattern per mask value,
    // no shuffles, no lane idles.  Coefficients no rotation touches are constants of the plan and fold into the energy.
    int f16_n = 0;                       // touched coefficients (0: form not available)
    std::vector<int> f16_lane;           // [16] compact-frame lane (f32 index) held by slot i, -1 = unused slot
    std::vector<uint8_t> f16_mask;       // [nseg]
    std::vector<double> f16_tab;         // [nseg][16][4] = (wc, wk, w1, 0): own factor fma(wc, cos, wk), partner factor w1 * sin
    std::vector<double> f16_init;        // [16]
    std::vector<double> f16_eweight_g;   // [max(n_groups, 1)][17]: 16 energy weights + the constant part of that group
    // Scaled form (every rotation moves every used slot, which is what a chain of Pauli exponentials on one orbit does): the
    // own-coefficient weights wc are carried as a per-slot scale lambda that only the host tracks (lambda' = wc * lambda), the
    // kernel evolves rho = r / lambda by rho_i' = cos * rho_i + kappa_i * (sin * rho_j), kappa = w1 lambda_j / (wc lambda_i) -
    // three FP64 instructions and 8 bytes of table per slot and segment instead of four and 32 - and lambda folds into the
    // energy weights.  f16_kappa is stored in the order the unrolled pair loop of mask f16_mask[k] consumes it.
    int f16_scaled = 0;
    std::vector<double> f16_kappa;       // [nseg][16]: (kappa_i, kappa_j) of the p-th pair (i < j = i ^ mask) at [2p], [2p + 1]
    std::vector<double> f16_sweight_g;   // [max(n_groups, 1)][17]: energy weights times the final lambda + the constant part
    // Constant-coefficient form with variant slots (QP error-mitigation batches of a resolved circuit, dmx_frame16v_kernel):
    // r_i' = w0 r_i + w1 r_j per segment, scaled like above to rho_i' = rho_i + kappa_i rho_j (one DFMA per slot and segment);
    // the coefficients no rotation touches are constants times the product of the row's diagonal corrections.
    int f16_const = 0;
    std::vector<int> f16_static_lane;    // compact-frame lanes of the untouched live coefficients (at most 8)
    std::vector<double> f16_cstat_g;     // [max(n_groups, 1)][8]: energy weight x constant value of those coefficients
    std::vector<int> f16_fac_row;        // [n_slots] first row of the slot's correction table, -1: slot absent from the circuit
    std::vector<double> f16_fac;         // [rows][24]: diagonal correction of (slot, index) on the 16 slots + 8 static coefficients
    int opt_f16_scaled = 1;              // use the scaled form of dmx_frame16t_kernel when the plan offers it
    int opt_f16_threads = 128;           // CTA size of dmx_frame16t_kernel
    int opt_frame16t = 1;                // 0 never, 1 device-resident batches from 2048 circuits, 2 also mapped host buffers

    dmx_plan_info info{};

    // device mirror (owned by dmx_exec.cu)
    void* dev = nullptr;
};

// host-only lowering: ops -> blocks -> dev_ops/passes (+ segment tables).  Throws std::runtime_error.
void lower_plan(Plan& p);
// limits of dmx_stream_kernel's parameter tables (dmx_stream.cuh) and the check both layers use
constexpr int STREAM_MAX_SWEEPS = 128, STREAM_MAX_MUS = 128, STREAM_MAX_BLOCKS = 120;   // sweeps, micro-ops <= blocks of a pass
bool pass_stream_eligible(const Plan& p, const Pass& ps);
std::string dump_plan(const Plan& p);

// Bank swizzle of a tile in shared memory: position = index with its low 4 bits (the 8-byte bank pair inside a 128-byte
// line) XORed by a GF(2)-linear hash of the higher bits; bit i of the hash = parity(index & SWZ_M[i]).  The masks were
// searched offline (tools/search_swizzle.py) so that for EVERY choice of two or three resident-code positions in tiles
// of 5..7 qubits the 16 lanes of a half-warp — which walk the four lowest remaining index bits — hit 16 distinct bank
// pairs, and so do the linear staging loops: conflict-free 8-byte accesses.  Linear => swz(a ^ b) = swz(a) ^ swz(b).
constexpr unsigned SWZ_M[4] = {0x1750u, 0x2aa0u, 0x1160u, 0x0ad0u};

// helpers shared with exec
inline int code_pos(int n, int q) { return 2 * (n - 1 - q); }
uint32_t pauli_index(int n, uint32_t xmask, uint32_t zmask);

}  // namespace dmx

struct dmx_plan {
    dmx::Plan p;
};
