// dmx_stream.cuh — streaming pass kernels for n > tile_k (BASELINE configs 4-5: 10..14 qubits).
//
// One CTA owns one tile (4^k Pauli coefficients, k = 6 or 7 resident qubits) of one circuit for one pass:
//     global -> shared (16-byte loads, whole tile in flight) -> sweeps -> global.
// dmx_stream3_kernel (default): every thread keeps the 64 coefficients of THREE resident codes in registers during a
// sweep (THREADS = 4^k / 64); dmx_stream_kernel: 2 groups x 16 coefficients of a code pair (THREADS = 4^k / 32).  Either
// way the whole tile is register-resident while the micro-ops run; shared memory is only the exchange medium between
// sweeps.
// Compared with the generic dmx_tile_kernel this kernel
//   * reads per-circuit chain matrices (products of rotations / fixed one-qubit blocks) that dmx_chain_kernel
//     computed ONCE per (circuit, pass) instead of once per tile,
//   * takes the sweep / micro-op tables of the pass as kernel parameters (constant bank) and keeps its coefficients
//     in shared memory,
//   * addresses the tile with one 3-input XOR per access (per-sweep masks precombined on the host).
// k = 6 tiles are 32 KiB: four CTAs share an SM, so the HBM phases of one tile overlap the sweeps of the others.
#pragma once
#include <cstdint>

#include "dmx_plan.hpp"

namespace dmx {

struct SweepX {               // 96 bytes, host-built (one per sweep of a pass)
    int32_t pa, pb;           // tile bit positions of the two resident codes, pa < pb
    int32_t first_mu, n_mu;   // pass-local micro-op range
    int32_t lscale, sscale;   // offsets into the pass coefficient area (16 scales) or -1
    uint32_t gtop;            // swizzled byte offset separating a thread's two groups
    int32_t pad;
    uint32_t lA[4], lB[4];    // load : byte offset of register e = lA[e >> 2] ^ lB[e & 3]   (swizzled, pre-shifted)
    uint32_t sA[4], sB[4];    // store: same
};

struct MuX {                  // 16 bytes
    int32_t kind;             // MuKind (CHAIN_* and FIX_* share a code path: `off` points at a 4x4 either way)
    int32_t off;              // offset (doubles) into the pass coefficient area [chain matrices | fixed coefficients]
    int32_t unital;           // 4x4: row 0 = column 0 = e0 -> only the 3x3 block acts
    int32_t pad;
};

struct StreamArgs {
    // the sweep and micro-op tables of the pass travel in the kernel parameters (constant bank, uniform loads)
    SweepX sweeps[STREAM_MAX_SWEEPS];
    MuX mus[STREAM_MAX_MUS];
    const double* coefs;          // pass-local fixed coefficient table
    const double* chainmats;      // [chunk][total_chains][16]
    int n_sweeps, n_mus, n_coefs;
    int total_chains, first_chain, n_chains;
    int n, k;
    int load, pad0;
    const double* src;            // [chunk][4^n] state before the pass (layout pos_in)
    double* dst;                  // state after the pass (layout pos_out); == src unless the pass re-orders the qubits
    // Staging maps, one set per side (a pass may store the state in a different qubit order than it loaded it).
    // A side enumerates the tile's elements in the order of their GLOBAL addresses: u = 2 * tid + j * 2 * THREADS;
    // gruns deposit the bits of u (and oruns those of the tile id) into the global index, uruns into the tile index.
    int n_gruns_l, n_oruns_l, n_gruns_s, n_oruns_s;
    int gruns_l[8][3], uruns_l[8][3], oruns_l[8][3];   // (source shift, destination shift, width)
    int gruns_s[8][3], uruns_s[8][3], oruns_s[8][3];
    uint32_t giter_l[16], giter_s[16];   // global index contribution of staging iteration j
    uint32_t siter_l[16], siter_s[16];   // swizzled tile index of the same
    uint32_t pbit_l, pbit_s;             // swizzled tile offset between the two elements of a 16-byte global chunk
    uint32_t init_mask;                  // first pass: coefficient is 1 when (global index & init_mask) == 0
};

// ---- triple sweeps (dmx_stream3_kernel) -------------------------------------------------------------------
struct SweepX3 {              // 128 bytes
    int32_t p[3];             // tile bit positions of the three resident codes, ascending (for the group base)
    int32_t first_mu, n_mu;
    int32_t lscale, sscale;   // 64 scales: >= 0 offset into StreamArgs3::cfix (constant bank), <= -2 offset -(v+2) into the
                              // shared-memory coefficient area, -1 none
    int32_t pad;              // wide sweeps (Sweep3::wide): 1 + offset (doubles, in the shared-memory coefficient area) of a 32-word
                              // descriptor {flags, l_nvar, s_nvar, -, ltm[8], stm[8], lvm[4], svm[4]}; the masks ltm / stm are
                              // swizzled byte offsets, the variant tables [2^nvar][64] sit at lscale / sscale (shared memory)
    uint32_t lA[4], lB[4], lC[4];   // load : byte offset of register e = lA[e >> 4] ^ lB[(e >> 2) & 3] ^ lC[e & 3]
    uint32_t sA[4], sB[4], sC[4];   // store: same
};

constexpr int STREAM3_CFIX = 1536;   // doubles of permutation scale tables carried in the kernel parameters

struct StreamArgs3 {
    SweepX3 sweeps[STREAM_MAX_SWEEPS];
    MuX mus[STREAM_MAX_MUS];      // MuX::pad carries the slot (0..2) or slot pair (0: AB, 1: AC, 2: BC)
    double cfix[STREAM3_CFIX];    // scale tables of the load / store permutations: read through the constant cache, so they
                                  // do not compete with the tile exchange for the shared-memory pipe
    const double* coefs;
    const double* chainmats;
    int n_sweeps, n_mus, n_coefs;
    int total_chains, first_chain, n_chains;
    int n, k;
    int load, pad0;
    const double* src;
    double* dst;
    int n_gruns_l, n_oruns_l, n_gruns_s, n_oruns_s;
    int gruns_l[8][3], uruns_l[8][3], oruns_l[8][3];
    int gruns_s[8][3], uruns_s[8][3], oruns_s[8][3];
    uint32_t giter_l[32], giter_s[32];
    uint32_t siter_l[32], siter_s[32];
    uint32_t pbit_l, pbit_s;
    uint32_t init_mask;
    // ---- Blackwell staging modes (dmx_stream3_kernel) ----
    // load_mode  LOAD_STAGED: 16-byte global loads -> registers -> swizzled tile (the portable path above)
    //            LOAD_BULK:   ONE cp.async.bulk of the whole tile (contiguous in the load-side layout) into shared memory,
    //                         completion on an mbarrier; the first sweep reads that LINEAR (global-order) image directly
    //            LOAD_DIRECT: the first sweep reads its registers straight from global memory (same linear addressing)
    // store_mode STORE_DIRECT: the last sweep writes its registers straight to global memory in the store-side layout
    // Both need the two lowest-positioned qubits of that side outside the sweep's register triple (the half-warp then
    // walks a 128-byte run: conflict-free / fully coalesced); dmx_lower.cpp picks the layouts accordingly.
    int load_mode, store_mode;
    int first_lp[3];                     // first sweep: bit positions of its three codes in the LINEAR load image, ascending
    int first_tl[4];                     // ... and where the thread's own bit pairs (non-register qubits in load-linear order)
                                         // sit in the tile-local index: base of that sweep's store into the swizzled tile
    int last_g[4];                       // last sweep: global bit position of the thread's j-th bit pair (non-register qubits
                                         // in tile-local order) in the store-side layout
    uint32_t first_A[4], first_B[4], first_C[4];   // linear BYTE offsets of register e = A[e >> 4] ^ B[(e >> 2) & 3] ^ C[e & 3]
    uint32_t last_A[4], last_B[4], last_C[4];      // store side: GLOBAL element-index offsets of register e
    uint32_t mbar_off;                   // byte offset of the mbarrier inside dynamic shared memory
    uint32_t tile_bytes;
};

enum StreamLoadMode : int { LOAD_INIT = 0, LOAD_STAGED = 1, LOAD_BULK = 2, LOAD_DIRECT = 3 };
enum StreamStoreMode : int { STORE_STAGED = 0, STORE_DIRECT = 1 };

__host__ __device__ __forceinline__ unsigned parity32(unsigned v) {
#ifdef __CUDA_ARCH__
    return __popc(v) & 1u;
#else
    return (unsigned)__builtin_popcount(v) & 1u;
#endif
}

// bank swizzle (masks: dmx_plan.hpp)
__host__ __device__ __forceinline__ unsigned swz_hd(unsigned i) {
    const unsigned h = parity32(i & 0x1750u) | (parity32(i & 0x2aa0u) << 1) | (parity32(i & 0x1160u) << 2) | (parity32(i & 0x0ad0u) << 3);
    return i ^ h;
}
static_assert(SWZ_M[0] == 0x1750u && SWZ_M[1] == 0x2aa0u && SWZ_M[2] == 0x1160u && SWZ_M[3] == 0x0ad0u, "swizzle masks out of sync");

__host__ __device__ __forceinline__ unsigned insert2_hd(unsigned g, int p) { return ((g >> p) << (p + 2)) | (g & ((1u << p) - 1u)); }

}  // namespace dmx

#ifdef __CUDACC__

namespace dmx {

// bit pair j of v -> bit position pos[j] (four pairs: the thread index of a k <= 7 tile with three register-resident codes)
__device__ __forceinline__ unsigned deposit_pairs(unsigned v, const int (&pos)[4]) {
    return ((v & 3u) << pos[0]) | (((v >> 2) & 3u) << pos[1]) | (((v >> 4) & 3u) << pos[2]) | (((v >> 6) & 3u) << pos[3]);
}

__device__ __forceinline__ unsigned map_runs(unsigned v, const int (*runs)[3], int n) {
    unsigned g = 0;
    for (int i = 0; i < n; ++i) g |= ((v >> runs[i][0]) & ((1u << runs[i][2]) - 1u)) << runs[i][1];
    return g;
}

// ---- mbarrier + bulk async copy (sm_90+ PTX; SASS: SYNCS.* / UBLKCP) -----------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // make the init visible to the async proxy
}

__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// global -> shared bulk copy (16-byte aligned, size a multiple of 16); completes `bytes` transactions on the mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---- per-circuit chain matrices, once per (circuit, chain) ---------------------------------------------------
__global__ void dmx_chain_kernel(const ChainDev* __restrict__ chains, const ChainFactor* __restrict__ factors,
                                 const double* __restrict__ coefs, const RotInfo* __restrict__ rots,
                                 const double* __restrict__ params, int n_params, int total_chains, long long chunk,
                                 long long circuit_offset, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= chunk * total_chains) return;
    const long long c = i / total_chains;
    const int chn = (int)(i - c * total_chains);
    const long long b = c + circuit_offset;
    double M[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) M[e] = (e % 5 == 0) ? 1.0 : 0.0;
    const ChainDev ch = chains[chn];
    for (int f = 0; f < ch.n_factors; ++f) {
        const ChainFactor fa = factors[ch.first_factor + f];
        const double* __restrict__ cf = coefs + fa.coef;
        double F[16];
        if (fa.kind == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) F[e] = cf[e];
        } else {
            const RotInfo ri = rots[fa.rot_id];
            double sv, cv;
            sincos(ri.scale * params[b * n_params + ri.param] + ri.offset, &sv, &cv);
#pragma unroll
            for (int e = 0; e < 16; ++e) F[e] = cf[e] + cv * cf[16 + e] + sv * cf[32 + e];
        }
        double T[16];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
                T[r * 4 + cc] = F[r * 4] * M[cc] + F[r * 4 + 1] * M[4 + cc] + F[r * 4 + 2] * M[8 + cc] + F[r * 4 + 3] * M[12 + cc];
#pragma unroll
        for (int e = 0; e < 16; ++e) M[e] = T[e];
    }
    double2* dst = reinterpret_cast<double2*>(out + i * 16);
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[e] = make_double2(M[2 * e], M[2 * e + 1]);
}

// ---- register-level micro-ops on the two groups of a thread --------------------------------------------------
// group element e = 4 * code(A) + code(B)
// SA / SB: strides of the acted-on code and of the spectator code in e
template <int SA, int SB>
__device__ __forceinline__ void apply4(double (&v)[16], const double* __restrict__ M) {
    double m[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) m[e] = M[e];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double a = v[s * SB], b = v[SA + s * SB], c = v[2 * SA + s * SB], d = v[3 * SA + s * SB];
        v[s * SB] = m[0] * a + m[1] * b + m[2] * c + m[3] * d;
        v[SA + s * SB] = m[4] * a + m[5] * b + m[6] * c + m[7] * d;
        v[2 * SA + s * SB] = m[8] * a + m[9] * b + m[10] * c + m[11] * d;
        v[3 * SA + s * SB] = m[12] * a + m[13] * b + m[14] * c + m[15] * d;
    }
}

// Both groups get the same matrix: the coefficients are loaded from shared memory once.
template <int SA, int SB>
__device__ __forceinline__ void apply3x2(double (&v0)[16], double (&v1)[16], const double* __restrict__ M) {
    const double2* __restrict__ M2 = reinterpret_cast<const double2*>(M);    // 4x4 tables are 16-byte aligned
    const double2 r1a = M2[2], r1b = M2[3], r2a = M2[4], r2b = M2[5], r3a = M2[6], r3b = M2[7];
    const double m5 = r1a.y, m6 = r1b.x, m7 = r1b.y, m9 = r2a.y, m10 = r2b.x, m11 = r2b.y, m13 = r3a.y, m14 = r3b.x, m15 = r3b.y;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        {
            const double a = v0[SA + s * SB], b = v0[2 * SA + s * SB], c = v0[3 * SA + s * SB];
            v0[SA + s * SB] = m5 * a + m6 * b + m7 * c;
            v0[2 * SA + s * SB] = m9 * a + m10 * b + m11 * c;
            v0[3 * SA + s * SB] = m13 * a + m14 * b + m15 * c;
        }
        {
            const double a = v1[SA + s * SB], b = v1[2 * SA + s * SB], c = v1[3 * SA + s * SB];
            v1[SA + s * SB] = m5 * a + m6 * b + m7 * c;
            v1[2 * SA + s * SB] = m9 * a + m10 * b + m11 * c;
            v1[3 * SA + s * SB] = m13 * a + m14 * b + m15 * c;
        }
    }
}

__device__ __forceinline__ void dense16x(double (&v)[16], const double* __restrict__ M) {
    double nv[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 16; ++c) s += M[r * 16 + c] * v[c];
        nv[r] = s;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = nv[e];
}

extern __shared__ __align__(16) double dmx_stream_smem[];

// GENERAL = false: unital one-qubit maps (3x3) and diagonal blocks only — every circuit built from rotations, Clifford
// gates and Pauli-type noise; GENERAL = true adds non-unital 4x4 maps (amplitude damping) and dense two-qubit blocks.
template <int THREADS, bool GENERAL>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 1 : 4) dmx_stream_kernel(const __grid_constant__ StreamArgs a) {
    constexpr int NITER = 16;                               // 4^k / (2 * THREADS) staging iterations (k = 6: 128 thr, k = 7: 512 thr)
    const unsigned E = 32u * THREADS;                       // tile elements
    double* tile = dmx_stream_smem;
    double* cf = tile + E;                                  // [n_chains*16 chain matrices | n_coefs fixed coefficients]
    char* const tb = reinterpret_cast<char*>(tile);
    const int tid = threadIdx.x;

    const unsigned tiles_per_circuit = 1u << (2 * (a.n - a.k));
    const long long circuit = blockIdx.x / tiles_per_circuit;
    const unsigned tile_id = blockIdx.x % tiles_per_circuit;
    const double* const st_in = a.src + (circuit << (2 * a.n));
    double* const st = a.dst + (circuit << (2 * a.n));

    // ---- pass tables -> shared memory ----
    {
        const double* cm = a.chainmats + ((size_t)circuit * a.total_chains + a.first_chain) * 16;
        for (int i = tid; i < a.n_chains * 16; i += THREADS) cf[i] = cm[i];
        for (int i = tid; i < a.n_coefs; i += THREADS) cf[a.n_chains * 16 + i] = a.coefs[i];
    }

    // ---- tile load / init ----
    if (a.load) {
        const unsigned gbase = map_runs(2u * tid, a.gruns_l, a.n_gruns_l) | map_runs(tile_id, a.oruns_l, a.n_oruns_l);
        const unsigned s_thr = swz_hd(map_runs(2u * tid, a.uruns_l, a.n_gruns_l));
        double2 x[NITER];
#pragma unroll
        for (int j = 0; j < NITER; ++j) x[j] = *reinterpret_cast<const double2*>(st_in + (gbase | a.giter_l[j]));
#pragma unroll
        for (int j = 0; j < NITER; ++j) {
            const unsigned sp = s_thr ^ a.siter_l[j];
            tile[sp] = x[j].x;
            tile[sp ^ a.pbit_l] = x[j].y;
        }
    } else {
        const unsigned gbase = map_runs(2u * tid, a.gruns_s, a.n_gruns_s) | map_runs(tile_id, a.oruns_s, a.n_oruns_s);
        const unsigned s_thr = swz_hd(map_runs(2u * tid, a.uruns_s, a.n_gruns_s));
#pragma unroll
        for (int j = 0; j < NITER; ++j) {
            const unsigned g = gbase | a.giter_s[j];     // |0..0><0..0| is symmetric under qubit relabelling: any layout works
            const unsigned sp = s_thr ^ a.siter_s[j];
            tile[sp] = (g & a.init_mask) ? 0.0 : 1.0;
            tile[sp ^ a.pbit_s] = ((g | 1u) & a.init_mask) ? 0.0 : 1.0;
        }
    }
    __syncthreads();

    // ---- sweeps ----
    for (int si = 0; si < a.n_sweeps; ++si) {
        const SweepX& sw = a.sweeps[si];
        const unsigned sb0 = swz_hd(insert2_hd(insert2_hd((unsigned)tid, sw.pa), sw.pb)) << 3;
        const unsigned sb1 = sb0 ^ sw.gtop;
        double v0[16], v1[16];
        {
            unsigned mA[4], mB[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { mA[j] = sw.lA[j]; mB[j] = sw.lB[j]; }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const unsigned off = mA[e >> 2] ^ mB[e & 3];
                v0[e] = *reinterpret_cast<const double*>(tb + (sb0 ^ off));
                v1[e] = *reinterpret_cast<const double*>(tb + (sb1 ^ off));
            }
            if (sw.lscale >= 0) {
                const double2* __restrict__ lc = reinterpret_cast<const double2*>(cf + sw.lscale);
#pragma unroll
                for (int e = 0; e < 8; ++e) { const double2 s = lc[e]; v0[2 * e] *= s.x; v1[2 * e] *= s.x; v0[2 * e + 1] *= s.y; v1[2 * e + 1] *= s.y; }
            }
        }
        for (int mi = 0; mi < sw.n_mu; ++mi) {
            const MuX m = a.mus[sw.first_mu + mi];
            const double* __restrict__ M = cf + m.off;
            switch (m.kind) {
                case MU_CHAIN_A: case MU_FIX_A:
                    if (!GENERAL || m.unital) apply3x2<4, 1>(v0, v1, M);
                    else { apply4<4, 1>(v0, M); apply4<4, 1>(v1, M); }
                    break;
                case MU_CHAIN_B: case MU_FIX_B:
                    if (!GENERAL || m.unital) apply3x2<1, 4>(v0, v1, M);
                    else { apply4<1, 4>(v0, M); apply4<1, 4>(v1, M); }
                    break;
                case MU_DIAG_A: {
                    const double d0 = M[0], d1 = M[1], d2 = M[2], d3 = M[3];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        v0[s] *= d0; v0[4 + s] *= d1; v0[8 + s] *= d2; v0[12 + s] *= d3;
                        v1[s] *= d0; v1[4 + s] *= d1; v1[8 + s] *= d2; v1[12 + s] *= d3;
                    }
                    break;
                }
                case MU_DIAG_B: {
                    const double d0 = M[0], d1 = M[1], d2 = M[2], d3 = M[3];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        v0[4 * s] *= d0; v0[4 * s + 1] *= d1; v0[4 * s + 2] *= d2; v0[4 * s + 3] *= d3;
                        v1[4 * s] *= d0; v1[4 * s + 1] *= d1; v1[4 * s + 2] *= d2; v1[4 * s + 3] *= d3;
                    }
                    break;
                }
                case MU_DIAG2: {
                    const double2* __restrict__ M2 = reinterpret_cast<const double2*>(M);
#pragma unroll
                    for (int e = 0; e < 8; ++e) { const double2 s = M2[e]; v0[2 * e] *= s.x; v1[2 * e] *= s.x; v0[2 * e + 1] *= s.y; v1[2 * e + 1] *= s.y; }
                    break;
                }
                default:   // MU_MAT2
                    if (GENERAL) { dense16x(v0, M); dense16x(v1, M); }
                    break;
            }
        }
        {
            unsigned mA[4], mB[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { mA[j] = sw.sA[j]; mB[j] = sw.sB[j]; }
            if (sw.sscale >= 0) {
                const double2* __restrict__ sc = reinterpret_cast<const double2*>(cf + sw.sscale);
#pragma unroll
                for (int e = 0; e < 8; ++e) { const double2 s = sc[e]; v0[2 * e] *= s.x; v1[2 * e] *= s.x; v0[2 * e + 1] *= s.y; v1[2 * e + 1] *= s.y; }
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const unsigned off = mA[e >> 2] ^ mB[e & 3];
                *reinterpret_cast<double*>(tb + (sb0 ^ off)) = v0[e];
                *reinterpret_cast<double*>(tb + (sb1 ^ off)) = v1[e];
            }
        }
        __syncthreads();
    }

    // ---- tile store ----
    {
        const unsigned gbase = map_runs(2u * tid, a.gruns_s, a.n_gruns_s) | map_runs(tile_id, a.oruns_s, a.n_oruns_s);
        const unsigned s_thr = swz_hd(map_runs(2u * tid, a.uruns_s, a.n_gruns_s));
#pragma unroll
        for (int j = 0; j < NITER; ++j) {
            const unsigned sp = s_thr ^ a.siter_s[j];
            *reinterpret_cast<double2*>(st + (gbase | a.giter_s[j])) = make_double2(tile[sp], tile[sp ^ a.pbit_s]);
        }
    }
}

// ---- triple-sweep kernel: 64 coefficients (three resident codes) per thread -----------------------------------
// THREADS = 4^k / 64 (k = 6: 64 threads, four CTAs per SM; k = 7: 256 threads).  Register e = 16*code(A) + 4*code(B) + code(C).
template <int S>   // stride of the acted-on code in e: 16 (slot A), 4 (slot B), 1 (slot C)
__device__ __forceinline__ int spectator_offset(int t) {     // t = 0..15 enumerates the other two codes
    return S == 16 ? t : (S == 4 ? ((t >> 2) * 16 + (t & 3)) : t * 4);
}

template <int S>
__device__ __forceinline__ void apply3_64(double (&v)[64], const double* __restrict__ M) {
    const double2* __restrict__ M2 = reinterpret_cast<const double2*>(M);
    const double2 r1a = M2[2], r1b = M2[3], r2a = M2[4], r2b = M2[5], r3a = M2[6], r3b = M2[7];
    const double m5 = r1a.y, m6 = r1b.x, m7 = r1b.y, m9 = r2a.y, m10 = r2b.x, m11 = r2b.y, m13 = r3a.y, m14 = r3b.x, m15 = r3b.y;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int o = spectator_offset<S>(t);
        const double x = v[o + S], y = v[o + 2 * S], z = v[o + 3 * S];
        v[o + S] = m5 * x + m6 * y + m7 * z;
        v[o + 2 * S] = m9 * x + m10 * y + m11 * z;
        v[o + 3 * S] = m13 * x + m14 * y + m15 * z;
    }
}

template <int S>
__device__ __forceinline__ void apply4_64(double (&v)[64], const double* __restrict__ M) {
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int o = spectator_offset<S>(t);
        const double w = v[o], x = v[o + S], y = v[o + 2 * S], z = v[o + 3 * S];
        v[o] = M[0] * w + M[1] * x + M[2] * y + M[3] * z;
        v[o + S] = M[4] * w + M[5] * x + M[6] * y + M[7] * z;
        v[o + 2 * S] = M[8] * w + M[9] * x + M[10] * y + M[11] * z;
        v[o + 3 * S] = M[12] * w + M[13] * x + M[14] * y + M[15] * z;
    }
}

template <int S>
__device__ __forceinline__ void diag1_64(double (&v)[64], const double* __restrict__ d) {
    const double d0 = d[0], d1 = d[1], d2 = d[2], d3 = d[3];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int o = spectator_offset<S>(t);
        v[o] *= d0; v[o + S] *= d1; v[o + 2 * S] *= d2; v[o + 3 * S] *= d3;
    }
}

// pair 0: (A,B), 1: (A,C), 2: (B,C); local index = 4 * code(first) + code(second)
template <int PAIR>
__device__ __forceinline__ int pair_index(int e) { return PAIR == 0 ? (e >> 2) : (PAIR == 1 ? ((e >> 4) * 4 + (e & 3)) : (e & 15)); }
template <int PAIR>
__device__ __forceinline__ int pair_element(int l, int spect) {   // inverse of pair_index for spectator code `spect`
    return PAIR == 0 ? (l * 4 + spect) : (PAIR == 1 ? ((l >> 2) * 16 + spect * 4 + (l & 3)) : (spect * 16 + l));
}

template <int PAIR>
__device__ __forceinline__ void diag2_64(double (&v)[64], const double* __restrict__ d) {
    const double2* __restrict__ d2 = reinterpret_cast<const double2*>(d);
    double c[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 t = d2[i]; c[2 * i] = t.x; c[2 * i + 1] = t.y; }
#pragma unroll
    for (int e = 0; e < 64; ++e) v[e] *= c[pair_index<PAIR>(e)];
}

template <int PAIR>
__device__ __forceinline__ void dense2_64(double (&v)[64], const double* __restrict__ M) {
#pragma unroll
    for (int sp = 0; sp < 4; ++sp) {
        double nv[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc += M[r * 16 + c] * v[pair_element<PAIR>(c, sp)];
            nv[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) v[pair_element<PAIR>(r, sp)] = nv[r];
    }
}

// (Measured and rejected, profiles/r2_stream3_experiments.txt: __maxnreg__(200) - five CTAs per SM, 80 B of spills outside the
// sweep loop - runs HEA-10 at 7 577 circuits/s against 8 235 with four CTAs at 238 registers.)
template <int THREADS, bool GENERAL>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 1 : 4) dmx_stream3_kernel(const __grid_constant__ StreamArgs3 a) {
    constexpr int NITER = 32;                               // 4^k / (2 * THREADS)
    const unsigned E = 64u * THREADS;
    double* tile = dmx_stream_smem;
    double* cf = tile + E;                                  // [n_chains*16 chain matrices | n_coefs fixed coefficients]
    char* const tb = reinterpret_cast<char*>(tile);
    const int tid = threadIdx.x;

    const unsigned tiles_per_circuit = 1u << (2 * (a.n - a.k));
    const long long circuit = blockIdx.x / tiles_per_circuit;
    const unsigned tile_id = blockIdx.x % tiles_per_circuit;
    const double* const st_in = a.src + (circuit << (2 * a.n));
    double* const st = a.dst + (circuit << (2 * a.n));
    // (Measured and rejected, profiles/r2_stream3_experiments.txt: a per-thread index on the permutation scale tables turns
    // ptxas' LDCU.64 + 2 IMAD.MOV per scale into one vector LDC.64, but the kernel then needs 255 registers and HEA-10 drops
    // from 8 235 to 7 335 circuits/s; scale tables in shared memory - LDS.128 - lose 6 %.)
    // the tile of the load side is contiguous for LOAD_BULK / LOAD_DIRECT: tile_id supplies the upper index bits
    const double* const tile_src = st_in + map_runs(tile_id, a.oruns_l, a.n_oruns_l);
    const unsigned mbar = smem_u32(reinterpret_cast<char*>(dmx_stream_smem) + a.mbar_off);
    if (a.load_mode == LOAD_BULK && tid == 0) {
        // issued first: the 32 KB are in flight while the CTA stages its tables
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, a.tile_bytes);
        bulk_g2s(smem_u32(tile), tile_src, a.tile_bytes, mbar);     // UBLKCP: the whole tile, one instruction
    }
    {
        const double* cm = a.chainmats + ((size_t)circuit * a.total_chains + a.first_chain) * 16;
        for (int i = tid; i < a.n_chains * 16; i += THREADS) cf[i] = cm[i];
        for (int i = tid; i < a.n_coefs; i += THREADS) cf[a.n_chains * 16 + i] = a.coefs[i];
    }

    if (a.load_mode == LOAD_STAGED) {
        const unsigned gbase = map_runs(2u * tid, a.gruns_l, a.n_gruns_l) | map_runs(tile_id, a.oruns_l, a.n_oruns_l);
        const unsigned s_thr = swz_hd(map_runs(2u * tid, a.uruns_l, a.n_gruns_l));
        double2 x[NITER];
#pragma unroll
        for (int j = 0; j < NITER; ++j) x[j] = *reinterpret_cast<const double2*>(st_in + (gbase | a.giter_l[j]));
#pragma unroll
        for (int j = 0; j < NITER; ++j) {
            const unsigned sp = s_thr ^ a.siter_l[j];
            tile[sp] = x[j].x;
            tile[sp ^ a.pbit_l] = x[j].y;
        }
    } else if (a.load_mode == LOAD_INIT) {
        const unsigned gbase = map_runs(2u * tid, a.gruns_s, a.n_gruns_s) | map_runs(tile_id, a.oruns_s, a.n_oruns_s);
        const unsigned s_thr = swz_hd(map_runs(2u * tid, a.uruns_s, a.n_gruns_s));
#pragma unroll
        for (int j = 0; j < NITER; ++j) {
            const unsigned g = gbase | a.giter_s[j];
            const unsigned sp = s_thr ^ a.siter_s[j];
            tile[sp] = (g & a.init_mask) ? 0.0 : 1.0;
            tile[sp ^ a.pbit_s] = ((g | 1u) & a.init_mask) ? 0.0 : 1.0;
        }
    }
    __syncthreads();
    if (a.load_mode == LOAD_BULK) mbar_wait(mbar, 0);

    for (int si = 0; si < a.n_sweeps; ++si) {
        const SweepX3& sw = a.sweeps[si];
        unsigned sb = swz_hd(insert2_hd(insert2_hd(insert2_hd((unsigned)tid, sw.p[0]), sw.p[1]), sw.p[2])) << 3;
        // wide sweep: the thread's base is a GF(2)-linear image of its index (per side), its scale table one of 2^nvar variants
        unsigned sbl = sb, sbs = sb, wflags = 0;
        int lvar = 0, svar = 0;
        if (sw.pad) {
            constexpr int NTB = THREADS == 64 ? 6 : 8;
            const unsigned* __restrict__ W = reinterpret_cast<const unsigned*>(cf + (sw.pad - 1));
            wflags = W[0];
            if (wflags & 1u) {
                sbl = 0;
#pragma unroll
                for (int j = 0; j < NTB; ++j) sbl ^= ((tid >> j) & 1) ? W[4 + j] : 0u;
                for (unsigned j = 0; j < W[1]; ++j) lvar |= (__popc((unsigned)tid & W[20 + j]) & 1) << j;
            }
            if (wflags & 2u) {
                sbs = 0;
#pragma unroll
                for (int j = 0; j < NTB; ++j) sbs ^= ((tid >> j) & 1) ? W[12 + j] : 0u;
                for (unsigned j = 0; j < W[2]; ++j) svar |= (__popc((unsigned)tid & W[24 + j]) & 1) << j;
            }
        }
        double v[64];
        if (si == 0 && a.load_mode >= LOAD_BULK) {
            // this thread holds the elements whose non-register qubits, taken in LOAD-LINEAR order, spell tid: that is the
            // identity its store into the swizzled tile-local layout must use
            sb = swz_hd(deposit_pairs((unsigned)tid, a.first_tl)) << 3;
            sbs = sb;
            // first sweep on the LINEAR image of the tile (what the bulk copy delivered, or global memory itself): the
            // thread's non-register bits fill the remaining index bits in ascending order, so a half-warp walks the 16
            // doubles of a 128-byte run
            const unsigned lb = insert2_hd(insert2_hd(insert2_hd((unsigned)tid, a.first_lp[0]), a.first_lp[1]), a.first_lp[2]) << 3;
            unsigned ab[16], sc[4];
#pragma unroll
            for (int i = 0; i < 16; ++i) ab[i] = a.first_A[i >> 2] ^ a.first_B[i & 3];
#pragma unroll
            for (int c = 0; c < 4; ++c) sc[c] = lb ^ a.first_C[c];
            if (a.load_mode == LOAD_BULK) {
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] = *reinterpret_cast<const double*>(tb + (sc[e & 3] ^ ab[e >> 2]));
            } else {
                const char* const gsrc = reinterpret_cast<const char*>(tile_src);
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] = __ldcs(reinterpret_cast<const double*>(gsrc + (sc[e & 3] ^ ab[e >> 2])));
            }
            if (sw.lscale >= 0) {
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] *= a.cfix[sw.lscale + e];
            } else if (sw.lscale <= -2) {
                const double2* __restrict__ lc = reinterpret_cast<const double2*>(cf - (sw.lscale + 2));
#pragma unroll
                for (int e = 0; e < 32; ++e) { const double2 s = lc[e]; v[2 * e] *= s.x; v[2 * e + 1] *= s.y; }
            }
            // the sweep stores in the swizzled tile-local layout: every thread must have read the linear image first
            if (a.load_mode == LOAD_BULK) __syncthreads();
        } else {
            // address of register e = (sb ^ lC[c]) ^ (lA[a] ^ lB[b]): 4 per-thread bases x 16 uniform offsets, one XOR per access
            unsigned ab[16], sc[4];
#pragma unroll
            for (int i = 0; i < 16; ++i) ab[i] = sw.lA[i >> 2] ^ sw.lB[i & 3];
#pragma unroll
            for (int c = 0; c < 4; ++c) sc[c] = sbl ^ sw.lC[c];
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] = *reinterpret_cast<const double*>(tb + (sc[e & 3] ^ ab[e >> 2]));
            if (sw.lscale >= 0) {
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] *= a.cfix[sw.lscale + e];
            } else if (sw.lscale <= -2) {
                const double2* __restrict__ lc = reinterpret_cast<const double2*>(cf - (sw.lscale + 2) + lvar * 64);
#pragma unroll
                for (int e = 0; e < 32; ++e) { const double2 s = lc[e]; v[2 * e] *= s.x; v[2 * e + 1] *= s.y; }
            }
            // a wide sweep writes other addresses than it reads: every thread must have loaded before anyone stores
            if (wflags & 4u) __syncthreads();
        }
        for (int mi = 0; mi < sw.n_mu; ++mi) {
            const MuX m = a.mus[sw.first_mu + mi];
            const double* __restrict__ M = cf + m.off;
            const int which = m.pad;
            switch (m.kind) {
                case MU3_MAT:
                    if (!GENERAL || m.unital) {
                        if (which == 0) apply3_64<16>(v, M); else if (which == 1) apply3_64<4>(v, M); else apply3_64<1>(v, M);
                    } else {
                        if (which == 0) apply4_64<16>(v, M); else if (which == 1) apply4_64<4>(v, M); else apply4_64<1>(v, M);
                    }
                    break;
                case MU3_DIAG:
                    if (which == 0) diag1_64<16>(v, M); else if (which == 1) diag1_64<4>(v, M); else diag1_64<1>(v, M);
                    break;
                case MU3_DIAG2:
                    if (which == 0) diag2_64<0>(v, M); else if (which == 1) diag2_64<1>(v, M); else diag2_64<2>(v, M);
                    break;
                default:   // MU3_MAT2
                    if (GENERAL) {
                        if (which == 0) dense2_64<0>(v, M); else if (which == 1) dense2_64<1>(v, M); else dense2_64<2>(v, M);
                    }
                    break;
            }
        }
        {
            if (sw.sscale >= 0) {
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] *= a.cfix[sw.sscale + e];
            } else if (sw.sscale <= -2) {
                const double2* __restrict__ sc2 = reinterpret_cast<const double2*>(cf - (sw.sscale + 2) + svar * 64);
#pragma unroll
                for (int e = 0; e < 32; ++e) { const double2 s = sc2[e]; v[2 * e] *= s.x; v[2 * e + 1] *= s.y; }
            }
            if (si == a.n_sweeps - 1 && a.store_mode == STORE_DIRECT) {
                // last sweep: registers straight to global memory in the store-side layout (128-byte runs per half-warp)
                const unsigned gb = deposit_pairs((unsigned)tid, a.last_g) | map_runs(tile_id, a.oruns_s, a.n_oruns_s);
                unsigned ab[16], sc[4];
#pragma unroll
                for (int i = 0; i < 16; ++i) ab[i] = a.last_A[i >> 2] ^ a.last_B[i & 3];
#pragma unroll
                for (int c = 0; c < 4; ++c) sc[c] = gb ^ a.last_C[c];
#pragma unroll
                for (int e = 0; e < 64; ++e) st[sc[e & 3] ^ ab[e >> 2]] = v[e];
                return;
            }
            unsigned ab[16], sc[4];
#pragma unroll
            for (int i = 0; i < 16; ++i) ab[i] = sw.sA[i >> 2] ^ sw.sB[i & 3];
#pragma unroll
            for (int c = 0; c < 4; ++c) sc[c] = sbs ^ sw.sC[c];
#pragma unroll
            for (int e = 0; e < 64; ++e) *reinterpret_cast<double*>(tb + (sc[e & 3] ^ ab[e >> 2])) = v[e];
        }
        __syncthreads();
    }

    {
        const unsigned gbase = map_runs(2u * tid, a.gruns_s, a.n_gruns_s) | map_runs(tile_id, a.oruns_s, a.n_oruns_s);
        const unsigned s_thr = swz_hd(map_runs(2u * tid, a.uruns_s, a.n_gruns_s));
#pragma unroll
        for (int j = 0; j < NITER; ++j) {
            const unsigned sp = s_thr ^ a.siter_s[j];
            *reinterpret_cast<double2*>(st + (gbase | a.giter_s[j])) = make_double2(tile[sp], tile[sp ^ a.pbit_s]);
        }
    }
}

}  // namespace dmx

#endif  // __CUDACC__
