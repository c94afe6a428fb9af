// dmx_exec.cu — sm_100a kernels and the C ABI of libdmx (include/dmx.h).
//
// Kernels (all FP64, no tensor cores — SURVEY.md §2.1):
//   dmx_tile_kernel     Pauli-vector tile resident in shared memory; interprets the fused block list of one
//                       pass (DIAG/MONO/MAT/ROT/VAR on 1 or 2 qubits).  n <= 7: whole circuit in one launch
//                       (several circuits per CTA for small n), fused Tr(rho H) epilogue.  n > 7: fallback for
//                       streaming passes the stream kernels do not take (variant slots, tiles below 6 qubits).
//   dmx_stream3_kernel / dmx_stream_kernel / dmx_chain_kernel (dmx_stream.cuh)
//                       n > 7: one tile of a streaming pass, triple (64 coefficients per thread) or pair sweeps.
//   dmx_frame32_kernel  n <= 4, Clifford + Pauli-noise (+ rotations) circuits whose live coefficient set fits a warp
//                       (noisy H2: 23 of 256): warp-per-circuit, partner by indexed shuffle, no barriers.
//   dmx_frame_kernel    same circuit class, any live set: thread-per-Pauli-coefficient in a fixed frame; a segment
//                       is r <- cm*w0*r + sm*w1*r[t^mask] with the partner fetched by a warp shuffle (or one
//                       shared-memory exchange when the mask crosses warps).
//   dmx_qp_sample_kernel  Philox4x32-10 draw of QP circuit variants.
//   dmx_energy_kernel   gather-dot Tr(rho H) on a state in global memory (streaming path).
//   dmx_pauli_to_rho / dmx_pauli_to_diag   output converters (final_density_matrix, run() probabilities).
//   dmx_qp_reduce_*     deterministic two-stage reduction of the QP-weighted estimator.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "dmx_plan.hpp"
#include "dmx_stream.cuh"

using namespace dmx;

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// NVTX range for the host-side phases (plan lowering, launch sequences, reductions): visible in nsys / ncu --nvtx
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(DMX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// device structures
// ---------------------------------------------------------------------------------------------------------
struct BitRun {
    int lshift, gshift, width;
};

struct TileArgs {
    const SweepDev* sweeps;
    const MicroOp* mus;
    const ChainDev* chains;
    const ChainFactor* factors;
    const Phase* phases;
    int first_phase, n_phases, max_chains;
    const double* coefs;
    const RotInfo* rots;
    const double* params;
    int n_params;
    const uint8_t* variants;
    int n_slots;
    int n, k, cpb, pad;         // pad: extra low index bits of the tile (2 when n == 1: virtual idle qubit)
    long long batch;
    const int* worklist;        // optional indirection: circuit id = worklist[i], count in *work_count
    const int* work_count;
    double* state;              // streaming: [chunk, 4^n]
    long long circuit_offset;   // streaming: first circuit of this chunk (for params/variants rows)
    int load, store;
    int n_truns, n_oruns;
    BitRun truns[8], oruns[8];
    int epilogue;               // 0 none, 1 energy, 2 copy state out
    const uint32_t* term_idx;
    const double* term_coef;    // [groups][n_terms]
    int n_terms;
    const int* group;           // optional: circuit b uses coefficient row group[b] (row 0 when NULL)
    double* out;
    int variants_compact;       // worklist runs: the variant row of worklist entry i is row i of `variants` (not row worklist[i])
};

__device__ __forceinline__ unsigned map_bits(unsigned v, const BitRun* runs, int n) {
    unsigned g = 0;
    for (int i = 0; i < n; ++i) g |= ((v >> runs[i].lshift) & ((1u << runs[i].width) - 1u)) << runs[i].gshift;
    return g;
}

__device__ __forceinline__ unsigned insert2(unsigned g, int p) {  // open a 2-bit hole at bit p
    return ((g >> p) << (p + 2)) | (g & ((1u << p) - 1u));
}

// bank swizzle of the tile in shared memory (dmx_plan.hpp: SWZ_M)
__device__ __forceinline__ unsigned swz(unsigned i) { return swz_hd(i); }

// tile index of register e (bits zB, xB, zA, xA) = base ^ XOR of the masks of its set bits
#define DMX_XADDR(base, m, e) ((base) ^ (((e) & 1) ? (m)[0] : 0u) ^ (((e) & 2) ? (m)[1] : 0u) ^ (((e) & 4) ? (m)[2] : 0u) ^ (((e) & 8) ? (m)[3] : 0u))
// same with BYTE offsets into the tile (masks pre-shifted by 3): one LOP3 per access, no index scaling
#define DMX_TILE_AT(tile_bytes, off) (*reinterpret_cast<double*>((tile_bytes) + (off)))

// ---------------------------------------------------------------------------------------------------------
// tile kernel: the Pauli vector (or one tile of it) lives in shared memory; a pass is a list of sweeps, each
// sweep = load a 16-coefficient group (all codes of a qubit pair) into registers, apply the fused micro-ops,
// store.  Monomial two-qubit blocks ride on the load / store addressing.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat4_apply(const double* __restrict__ M, double& v0, double& v1, double& v2, double& v3) {
    double n0 = M[0] * v0 + M[1] * v1 + M[2] * v2 + M[3] * v3;
    double n1 = M[4] * v0 + M[5] * v1 + M[6] * v2 + M[7] * v3;
    double n2 = M[8] * v0 + M[9] * v1 + M[10] * v2 + M[11] * v3;
    double n3 = M[12] * v0 + M[13] * v1 + M[14] * v2 + M[15] * v3;
    v0 = n0; v1 = n1; v2 = n2; v3 = n3;
}

// unital, trace-preserving 1-qubit map: row 0 = column 0 = e0, only the 3x3 block acts
__device__ __forceinline__ void mat3_apply(const double* __restrict__ M, double& v1, double& v2, double& v3) {
    double n1 = M[5] * v1 + M[6] * v2 + M[7] * v3;
    double n2 = M[9] * v1 + M[10] * v2 + M[11] * v3;
    double n3 = M[13] * v1 + M[14] * v2 + M[15] * v3;
    v1 = n1; v2 = n2; v3 = n3;
}

__device__ __forceinline__ void side_a(double (&v)[16], const double* __restrict__ M, bool unital) {
    if (unital) {
#pragma unroll
        for (int c1 = 0; c1 < 4; ++c1) mat3_apply(M, v[4 + c1], v[8 + c1], v[12 + c1]);
    } else {
#pragma unroll
        for (int c1 = 0; c1 < 4; ++c1) mat4_apply(M, v[c1], v[4 + c1], v[8 + c1], v[12 + c1]);
    }
}

__device__ __forceinline__ void side_b(double (&v)[16], const double* __restrict__ M, bool unital) {
    if (unital) {
#pragma unroll
        for (int c0 = 0; c0 < 4; ++c0) mat3_apply(M, v[4 * c0 + 1], v[4 * c0 + 2], v[4 * c0 + 3]);
    } else {
#pragma unroll
        for (int c0 = 0; c0 < 4; ++c0) mat4_apply(M, v[4 * c0], v[4 * c0 + 1], v[4 * c0 + 2], v[4 * c0 + 3]);
    }
}

__device__ __forceinline__ void dense16(double (&v)[16], const double* __restrict__ M) {
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

extern __shared__ __align__(16) double dmx_smem[];

// G = 16-coefficient groups a thread carries per sweep; MAXT = CTA size bound.  A 7-qubit tile (1024 groups) runs either as
// 512 threads x 2 groups or as 1024 threads x 1 group (plan option "tile_threads"): the tile fills an SM's shared memory, so
// the second form is the only way to put 32 warps instead of 16 behind the barrier-separated sweeps.
template <int G, int MAXT = 512>
__global__ void __launch_bounds__(MAXT) dmx_tile_kernel(const TileArgs a) {
    double* tile = dmx_smem;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tbits = 2 * a.k + a.pad;               // index bits of one circuit's tile
    const unsigned tile_elems = 1u << tbits;
    const unsigned E = (unsigned)a.cpb << tbits;
    const unsigned NG = E >> 4;
    double* mats = tile + E;                         // [cpb][max_chains][16]
    const bool streaming = a.k < a.n;

    long long total = a.batch;
    if (a.work_count) {
        asm volatile("griddepcontrol.wait;" ::: "memory");   // launched programmatically behind the frame kernel (launch_tile)
        total = *a.work_count;
    }
    long long ngroups;
    unsigned tiles_per_circuit = 1;
    if (streaming) {
        tiles_per_circuit = 1u << (2 * (a.n - a.k));
        ngroups = total * tiles_per_circuit;
    } else {
        ngroups = (total + a.cpb - 1) / a.cpb;
    }

    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        long long c_first;
        unsigned tile_base = 0;
        if (streaming) {
            c_first = grp / tiles_per_circuit;
            tile_base = map_bits((unsigned)(grp % tiles_per_circuit), a.oruns, a.n_oruns);
        } else {
            c_first = grp * a.cpb;
        }
        // ---- load / init ----
        // Element pairs t = 2*tid + j*2*nthr: the per-thread part (2*tid) and the per-iteration part (warp-uniform)
        // occupy disjoint bits, and both the tile->global bit map and the swizzle are bitwise-linear, so the address
        // arithmetic per element is one OR / XOR with hoisted values.
        const unsigned pair_stride = 2u * nthr;                 // power of two
        const unsigned g_thr = streaming ? map_bits(2u * tid, a.truns, a.n_truns) : 0u;
        const unsigned s_thr = swz(2u * tid);
        if (streaming) {
            double* st = a.state + (c_first << (2 * a.n));
            if (a.load) {
                // batches of 8 independent 16-byte loads per thread keep enough bytes in flight to cover HBM latency
                for (unsigned tj0 = 0; tj0 < tile_elems; tj0 += 8u * pair_stride) {
                    double2 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const unsigned tj = tj0 + u * pair_stride;
                        if (2u * tid + tj < tile_elems) {
                            const unsigned g = g_thr | map_bits(tj, a.truns, a.n_truns) | tile_base;
                            v[u] = __ldcs(reinterpret_cast<const double2*>(st + g));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const unsigned tj = tj0 + u * pair_stride;
                        if (2u * tid + tj < tile_elems) {
                            const unsigned sp = s_thr ^ swz(tj);
                            tile[sp] = v[u].x;
                            tile[sp ^ 1u] = v[u].y;
                        }
                    }
                }
            } else {
                for (unsigned tj = 0; tj < tile_elems; tj += pair_stride) {
                    if (2u * tid + tj >= tile_elems) break;
                    const unsigned g = g_thr | map_bits(tj, a.truns, a.n_truns) | tile_base;
                    const unsigned sp = s_thr ^ swz(tj);
                    tile[sp] = (g & 0xAAAAAAAAu) ? 0.0 : 1.0;
                    tile[sp ^ 1u] = ((g | 1u) & 0xAAAAAAAAu) ? 0.0 : 1.0;
                }
            }
        } else {
            const unsigned padmask = (1u << a.pad) - 1u;
            for (unsigned e = tid; e < E; e += nthr) {
                unsigned loc = e & (tile_elems - 1);
                tile[swz(e)] = ((loc & padmask) || ((loc >> a.pad) & 0xAAAAAAAAu)) ? 0.0 : 1.0;
            }
        }

        for (int ph = 0; ph < a.n_phases; ++ph) {
            const Phase phase = a.phases[a.first_phase + ph];
            // ---- per-circuit chain matrices (products of rotations / fixed 1-qubit blocks) ----
            __syncthreads();   // previous phase's sweeps are done with the chain table
            for (int i = tid; i < a.cpb * phase.n_chains; i += nthr) {
                const int c = i / phase.n_chains, chn = i - c * phase.n_chains;
                const long long ci = c_first + c;
                double M[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) M[e] = (e % 5 == 0) ? 1.0 : 0.0;
                if (ci < total) {
                    const long long b = a.worklist ? a.worklist[ci] : ci + a.circuit_offset;
                    const ChainDev ch = a.chains[phase.first_chain + chn];
                    for (int f = 0; f < ch.n_factors; ++f) {
                        const ChainFactor fa = a.factors[ch.first_factor + f];
                        const double* __restrict__ cf = a.coefs + fa.coef;
                        double F[16];
                        if (fa.kind == 0) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) F[e] = cf[e];
                        } else {
                            const RotInfo ri = a.rots[fa.rot_id];
                            double sv, cv;
                            sincos(ri.scale * a.params[b * a.n_params + ri.param] + ri.offset, &sv, &cv);
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
                }
                double* dst = mats + (size_t)(c * a.max_chains + chn) * 16;
#pragma unroll
                for (int e = 0; e < 16; ++e) dst[e] = M[e];
            }
            __syncthreads();

            for (int si = 0; si < phase.n_sweeps; ++si) {
                const SweepDev sw = a.sweeps[phase.first_sweep + si];
                const int p0 = sw.p0, p1 = sw.p1;
                const int pa = p0 < p1 ? p0 : p1, pb = p0 < p1 ? p1 : p0;
                double v[G][16];
                unsigned base[G], sbase[G];
                unsigned lm[4], sm[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { lm[j] = swz(sw.lmask[j]) << 3; sm[j] = swz(sw.smask[j]) << 3; }
                char* const tile_bytes = reinterpret_cast<char*>(tile);
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const unsigned g = tid + i * nthr;
                    base[i] = insert2(insert2(g < NG ? g : 0u, pa), pb);
                    sbase[i] = swz(base[i]) << 3;
                    if (sw.load_perm) {
                        const double* __restrict__ lc = a.coefs + sw.lcoef;
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[i][e] = lc[e] * DMX_TILE_AT(tile_bytes, DMX_XADDR(sbase[i], lm, e));
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[i][e] = DMX_TILE_AT(tile_bytes, DMX_XADDR(sbase[i], lm, e));
                    }
                }
                for (int mi = 0; mi < sw.n_mu; ++mi) {
                    const MicroOp m = a.mus[sw.first_mu + mi];
#pragma unroll
                    for (int i = 0; i < G; ++i) {
                        const int c = a.cpb == 1 ? 0 : (int)(base[i] >> tbits);
                        switch (m.kind) {
                            case MU_CHAIN_A: side_a(v[i], mats + (size_t)(c * a.max_chains + m.a0) * 16, m.a3 != 0); break;
                            case MU_CHAIN_B: side_b(v[i], mats + (size_t)(c * a.max_chains + m.a0) * 16, m.a3 != 0); break;
                            case MU_FIX_A: side_a(v[i], a.coefs + m.a0, m.a3 != 0); break;
                            case MU_FIX_B: side_b(v[i], a.coefs + m.a0, m.a3 != 0); break;
                            case MU_DIAG_A: {
                                const double* __restrict__ d = a.coefs + m.a0;
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[i][e] *= d[e >> 2];
                                break;
                            }
                            case MU_DIAG_B: {
                                const double* __restrict__ d = a.coefs + m.a0;
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[i][e] *= d[e & 3];
                                break;
                            }
                            case MU_DIAG2: {
                                const double* __restrict__ d = a.coefs + m.a0;
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[i][e] *= d[e];
                                break;
                            }
                            case MU_MAT2: dense16(v[i], a.coefs + m.a0); break;
                            default: {   // variant micro-ops
                                const long long ci = c_first + c;
                                unsigned idx = 0;
                                if (a.variants && ci < total) {
                                    const long long b = a.variants_compact ? ci : (a.worklist ? a.worklist[ci] : ci + a.circuit_offset);
                                    idx = a.variants[b * a.n_slots + m.a0];
                                }
                                if (!idx) break;
                                if (m.kind == MU_VAR_A) side_a(v[i], a.coefs + m.a1 + 16 * (idx & 15u), false);
                                else if (m.kind == MU_VAR_B) side_b(v[i], a.coefs + m.a1 + 16 * (idx & 15u), false);
                                else {
                                    side_a(v[i], a.coefs + m.a1 + 16 * (idx >> 4), false);
                                    side_b(v[i], a.coefs + m.a1 + 16 * (idx & 15u), false);
                                    const double* __restrict__ cc = a.coefs + m.a2;
                                    if (m.a3 == CLS_DIAG) {
#pragma unroll
                                        for (int e = 0; e < 16; ++e) v[i][e] *= cc[e];
                                    } else {
                                        dense16(v[i], cc);
                                    }
                                }
                                break;
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    if (tid + i * nthr >= NG) continue;
                    if (sw.store_perm) {
                        const double* __restrict__ sc = a.coefs + sw.scoef;
#pragma unroll
                        for (int e = 0; e < 16; ++e) DMX_TILE_AT(tile_bytes, DMX_XADDR(sbase[i], sm, e)) = sc[e] * v[i][e];
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) DMX_TILE_AT(tile_bytes, DMX_XADDR(sbase[i], sm, e)) = v[i][e];
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();

        // ---- epilogue ----
        if (streaming) {
            if (a.store) {
                double* st = a.state + (c_first << (2 * a.n));
                for (unsigned tj = 0; tj < tile_elems; tj += pair_stride) {
                    if (2u * tid + tj >= tile_elems) break;
                    const unsigned g = g_thr | map_bits(tj, a.truns, a.n_truns) | tile_base;
                    const unsigned sp = s_thr ^ swz(tj);
                    *reinterpret_cast<double2*>(st + g) = make_double2(tile[sp], tile[sp ^ 1u]);
                }
            }
        } else if (a.epilogue == 1) {
            const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
            for (int c = warp; c < a.cpb; c += nwarps) {
                long long ci = c_first + c;
                if (ci >= total) continue;
                const unsigned cbase = (unsigned)c << tbits;
                const long long b = a.worklist ? a.worklist[ci] : ci;
                const double* __restrict__ tc = a.term_coef + (a.group ? (size_t)a.group[b] * a.n_terms : 0);
                double e = 0.0;
                for (int t = lane; t < a.n_terms; t += 32) e += tc[t] * tile[swz(cbase | (a.term_idx[t] << a.pad))];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
                if (lane == 0) a.out[b] = e;
            }
        } else if (a.epilogue == 2) {
            const unsigned padmask = (1u << a.pad) - 1u;
            for (unsigned e = tid; e < E; e += nthr) {
                long long ci = c_first + (e >> tbits);
                unsigned loc = e & (tile_elems - 1);
                if (ci < total && !(loc & padmask)) a.out[(ci << (2 * a.n)) + (loc >> a.pad)] = tile[swz(e)];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// segment kernel, fixed frame (n <= 4 padded to 4 qubits: 256 Pauli coefficients).
// Thread t owns one Pauli coefficient of CPB circuits for the whole circuit.  Per segment
//     r <- cm * w0[t] * r + sm * w1[t] * r[t ^ mask]
// where the partner value comes from a warp shuffle when the mask stays inside the warp (the GF(2) relabelling
// in dmx_lower.cpp arranges that for as many segments as possible) or from one shared-memory exchange.
// ---------------------------------------------------------------------------------------------------------
struct SlotDev {
    int seg, nq, tab_off, pad;
    unsigned diag_ok[8];
};

struct FrameArgs {
    int nseg, n_params, n_slots, n_classes, n_eterms;
    long long batch;
    const double2* w;           // [nseg][256] (w0, w1)
    const int4* segs;           // [nseg] (xor_mask, use_shfl, cs_sel, 0)
    const RotInfo* classes;     // [n_classes]
    const double* init;         // [256]
    const double* eweight;      // [groups][256]
    const int* group;           // optional coefficient row per circuit
    const int* eslot;           // [256] compact index of the thread's energy term or -1
    const double* params;
    const uint8_t* variants;
    const SlotDev* slots;
    const uint8_t* labels;      // [n_slots][256]
    const double* coefs;
    double* out;
    int* worklist;              // circuits that need the general path
    int* work_count;
};

constexpr int SEG_MAXACT = 24;

// CSMODE 0: no rotation classes (all segments constant), 1: one class, (cos, sin) kept in registers,
// 2: general, (cos, sin) per (circuit, class) read from shared memory each segment.
template <int CPB, int CSMODE>
__global__ void __launch_bounds__(256) dmx_frame_kernel(const FrameArgs a) {
    const int NC = a.n_classes > 0 ? a.n_classes : 1;
    double* ex = dmx_smem;                                        // [2][CPB][256]
    double2* cs = reinterpret_cast<double2*>(ex + 2 * CPB * 256);  // [CPB][NC]
    double* esm = reinterpret_cast<double*>(cs + CPB * NC);        // [CPB][32*ceil(n_eterms/32)]
    const int epad = ((a.n_eterms + 31) / 32) * 32;
    int* nact = reinterpret_cast<int*>(esm + CPB * (epad > 0 ? epad : 32));   // [CPB] (-1: general path)
    int* any_act = nact + CPB;
    unsigned* act = reinterpret_cast<unsigned*>(nact + CPB + 1);   // [CPB][SEG_MAXACT]  slot<<8 | idx
    const int p = threadIdx.x, warp = p >> 5, lane = p & 31;
    const long long ngroups = (a.batch + CPB - 1) / CPB;
    const double init = a.init[p];
    const int eslot = a.eslot[p];

    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long b0 = grp * CPB;
        if (CSMODE != 0) {
            for (int i = p; i < CPB * NC; i += 256) {
                int c = i / NC, cl = i - c * NC;
                double sv = 0.0, cv = 1.0;
                if (b0 + c < a.batch) {
                    RotInfo ri = a.classes[cl];
                    sincos(ri.scale * a.params[(b0 + c) * a.n_params + ri.param] + ri.offset, &sv, &cv);
                }
                cs[i] = make_double2(cv, sv);
            }
        }
        if (p == 0) *any_act = 0;
        __syncthreads();
        if (a.variants) {
            for (int c = warp; c < CPB; c += 8) {
                int count = 0;
                bool general = false;
                if (b0 + c < a.batch) {
                    const uint8_t* row = a.variants + (b0 + c) * a.n_slots;
                    for (int s0 = 0; s0 < a.n_slots; s0 += 32) {
                        int s = s0 + lane;
                        unsigned idx = s < a.n_slots ? row[s] : 0u;
                        if (idx && a.slots[s].seg < 0) idx = 0;  // slot not present in the circuit
                        bool nz = idx != 0;
                        if (nz && !((a.slots[s].diag_ok[idx >> 5] >> (idx & 31)) & 1u)) general = true;
                        unsigned m = __ballot_sync(0xffffffffu, nz);
                        int pos = count + __popc(m & ((1u << lane) - 1u));
                        if (nz && pos < SEG_MAXACT) act[c * SEG_MAXACT + pos] = ((unsigned)s << 8) | idx;
                        count += __popc(m);
                    }
                    general = __any_sync(0xffffffffu, general) || count > SEG_MAXACT;
                }
                if (lane == 0) {
                    nact[c] = general ? -1 : count;
                    if (general || count) atomicOr(any_act, 1);
                    if (general) a.worklist[atomicAdd(a.work_count, 1)] = (int)(b0 + c);
                }
            }
            __syncthreads();
        }
        const bool have_act = a.variants && *any_act;

        double r[CPB];
        double2 csr[CPB];
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            r[c] = init;
            csr[c] = CSMODE == 1 ? cs[c * NC] : make_double2(1.0, 1.0);
        }
        int parity = 0;

        for (int k = 0; k <= a.nseg; ++k) {
            if (have_act) {
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    int na = nact[c];
                    for (int e = 0; e < na; ++e) {
                        unsigned ent = act[c * SEG_MAXACT + e];
                        int s = ent >> 8;
                        unsigned idx = ent & 255u;
                        SlotDev sd = a.slots[s];
                        if (sd.seg != k) continue;
                        unsigned lab = a.labels[s * 256 + p];
                        const double* t = a.coefs + sd.tab_off;
                        double f = sd.nq == 1 ? t[idx * 4 + lab]
                                              : t[(idx >> 4) * 4 + (lab >> 2)] * t[(idx & 15u) * 4 + (lab & 3u)] * t[64 + lab];
                        r[c] *= f;
                    }
                }
            }
            if (k == a.nseg) break;
            const int4 sg = a.segs[k];
            const double2 wk = a.w[k * 256 + p];
            const bool active = wk.y != 0.0;
            double other[CPB];
            if (sg.y) {
#pragma unroll
                for (int c = 0; c < CPB; ++c) other[c] = __shfl_xor_sync(0xffffffffu, r[c], sg.x);
            } else {
                double* buf = ex + parity * CPB * 256;
                parity ^= 1;
#pragma unroll
                for (int c = 0; c < CPB; ++c) buf[c * 256 + p] = r[c];
                __syncthreads();
#pragma unroll
                for (int c = 0; c < CPB; ++c) other[c] = active ? buf[c * 256 + (p ^ sg.x)] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < CPB; ++c) {
                double2 v;
                if (sg.z < 0) v = make_double2(1.0, 1.0);
                else if (CSMODE == 1) v = csr[c];
                else v = cs[c * NC + sg.z];
                const double cm = active ? v.x : 1.0;
                r[c] = cm * (wk.x * r[c]) + v.y * (wk.y * other[c]);
            }
        }
        // energy stage: threads that hold a Hamiltonian term drop weight * r into a compact row, one warp per
        // circuit sums it in a fixed order
        if (eslot >= 0) {
#pragma unroll
            for (int c = 0; c < CPB; ++c) {
                const int g = (a.group && b0 + c < a.batch) ? a.group[b0 + c] : 0;
                esm[c * epad + eslot] = a.eweight[g * 256 + p] * r[c];
            }
        }
        __syncthreads();
        for (int c = warp; c < CPB; c += 8) {
            if (b0 + c >= a.batch) continue;
            double e = 0.0;
            for (int t = lane; t < a.n_eterms; t += 32) e += esm[c * epad + t];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
            if (lane == 0 && !(a.variants && nact[c] < 0)) a.out[b0 + c] = e;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// compact frame kernel ("frame32"): liveness analysis on the host (dmx_lower.cpp::build_frame32) shows that only a few of
// the 256 Pauli coefficients can both be reached from |0..0> and reach a Hamiltonian term (noisy H2 UCCSD: 23).  When at
// most 32 can, one WARP carries CPB circuits: lane l owns coefficient orig[l], the partner of a segment comes from
// __shfl_sync(r, partner[k][l]), the energy is a warp butterfly.  No shared-memory exchange, no block barrier; warps of a
// CTA run independent circuit groups.
// ---------------------------------------------------------------------------------------------------------
struct Frame32Args {
    int nseg, n_params, n_slots, n_classes;
    int any_const;                // some segment has cs_sel < 0
    int stage_io;                 // parameters / energies are mapped HOST memory: move them as one block per CTA (see the kernel)
    long long batch;
    const double2* w;             // [nseg][32] (w0, w1)
    const unsigned char* partner; // [nseg][32] partner lane | 0x80 (rotation moves this coefficient)
    const int* cs_sel;            // [nseg] rotation class or < 0 (constant coefficients)
    const RotInfo* classes;
    RotInfo cls0;                 // classes[0] by value (CSMODE 1 reads the angle without a dependent table load)
    const double* init;           // [32]
    const double* eweight;        // [groups][32]
    const int* group;             // optional coefficient row per circuit
    const int* orig;              // [32] frame index (labels are indexed by it), -1 idle
    const double* params;
    const uint8_t* variants;
    const SlotDev* slots;
    const uint8_t* labels;        // [n_slots][256]
    const double* coefs;
    double* out;
    int* worklist;
    int* work_count;
};

// diagonal factor of variant entry `ent` for the coefficient whose frame index is p.  Inlined ONCE: the entry loop of the
// kernel runs over the circuits of a group with a run-time index (a call in the segment loop pinned the eight circuit
// registers to the callee-saved set and cost ~30 register moves per segment, 13 % of the QEM kernel's instructions).
__device__ __forceinline__ double frame32_variant_factor(const SlotDev* __restrict__ slots, const uint8_t* __restrict__ labels,
                                                      const double* __restrict__ coefs, unsigned ent, int p) {
    const int s = (ent >> 8) & 4095u;
    const unsigned idx = ent & 255u;
    const SlotDev sd = slots[s];
    const unsigned lab = labels[s * 256 + p];
    const double* t = coefs + sd.tab_off;
    return sd.nq == 1 ? t[idx * 4 + lab] : t[(idx >> 4) * 4 + (lab >> 2)] * t[(idx & 15u) * 4 + (lab & 3u)] * t[64 + lab];
}

// Scan of one group's identifier rows (kept out of line: its registers must not count against the segment loop's 64).
// Returns stage_any; fills act[] / nact[] (per-warp shared memory) and pushes rows that need the general path.
template <int CPB>
__device__ __noinline__ unsigned frame32_scan_group(const uint8_t* __restrict__ variants, int n_slots, long long batch, const SlotDev* __restrict__ slots,
                                                    int* worklist, int* work_count, long long b0, int lane, unsigned* act, int* nact) {
    unsigned stage_any = 0u;
    __syncwarp();
    if (lane < CPB) nact[lane] = 0;
    __syncwarp();
    unsigned general = 0u;                                  // bit c: row c needs the general path (uniform)
    const long long rows_here = batch - b0 < CPB ? batch - b0 : CPB;
    const unsigned nsl = (unsigned)n_slots;
    const unsigned valid = rows_here > 0 ? (unsigned)rows_here * nsl : 0u;      // bytes of this group inside the table
    const uint8_t* const gbase = variants + b0 * n_slots;
    const bool word_ok = (reinterpret_cast<unsigned long long>(variants) & 3ull) == 0ull;
    // Eight 128-byte steps are loaded ahead (independent loads: one memory round trip for the 976 bytes of an H2 group
    // instead of one per step - the serial loads were 27 % of the kernel's stall samples), and a group whose rows are all
    // identity leaves after ONE vote.
    constexpr int AHEAD = 8;
    for (unsigned s0 = 0; s0 < valid; s0 += 128u * AHEAD) {
        unsigned words[AHEAD];
        unsigned any = 0u;
#pragma unroll
        for (int i = 0; i < AHEAD; ++i) {
            const unsigned fi = s0 + 128u * i + 4u * lane;
            words[i] = 0u;
            if (word_ok && fi + 4u <= valid) {
                words[i] = *reinterpret_cast<const unsigned*>(gbase + fi);
            } else if (fi < valid) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (fi + j < valid) words[i] |= (unsigned)gbase[fi + j] << (8 * j);
            }
            any |= words[i];
        }
        if (!__any_sync(0xffffffffu, any != 0u)) continue;          // identity entries only: nothing to record
#pragma unroll 1
      for (int i = 0; i < AHEAD; ++i) {
        unsigned word = words[0];
#pragma unroll
        for (int q = 1; q < AHEAD; ++q) word = i == q ? words[q] : word;       // selects: `words` stays in registers
        if (!__any_sync(0xffffffffu, word != 0u)) continue;
        const unsigned f0 = s0 + 128u * i;
        const unsigned f = f0 + 4u * lane;
        const int c_lo = (int)(f0 / nsl);
        const int c_hi = (int)min((f0 + 127u) / nsl, (unsigned)(CPB - 1));
        unsigned ent[4];                // (stage << 20 | slot << 8 | idx) of this lane's four bytes, 0 = identity
        int row[4];
        bool gen[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned idx = (word >> (8 * j)) & 255u;
            int sl = 0, seg = -1;
            row[j] = -1;
            if (idx) {
                int rj = c_lo;                                       // the step touches rows c_lo .. c_hi (two or three of them)
                for (int cc = c_lo + 1; cc <= c_hi; ++cc) rj += (f + j >= (unsigned)cc * nsl) ? 1 : 0;
                row[j] = rj;
                sl = (int)(f + j - (unsigned)rj * nsl);
                seg = slots[sl].seg;
                if (seg < 0) idx = 0u;                              // slot not present in the circuit
            }
            const bool nz = idx != 0u;
            gen[j] = nz && !((slots[sl].diag_ok[idx >> 5] >> (idx & 31u)) & 1u);
            ent[j] = nz ? ((unsigned)seg << 20) | ((unsigned)sl << 8) | idx : 0u;
            stage_any |= __reduce_or_sync(0xffffffffu, nz ? (seg < 31 ? 1u << seg : 0x80000000u) : 0u);
        }
        const unsigned below = (1u << lane) - 1u;
        for (int cc = c_lo; cc <= c_hi; ++cc) {               // rows this 128-byte step touches (uniform bounds)
            unsigned m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = __ballot_sync(0xffffffffu, ent[j] != 0u && row[j] == cc);
            if (!(m[0] | m[1] | m[2] | m[3])) continue;
            if (__any_sync(0xffffffffu, (gen[0] && row[0] == cc) || (gen[1] && row[1] == cc) || (gen[2] && row[2] == cc) || (gen[3] && row[3] == cc)))
                general |= 1u << cc;
            // entries are recorded in ascending slot order (flat index 4 * lane + j), whatever the row's alignment in the
            // step: equal rows give bit-identical energies wherever they sit in the batch
            const int base = nact[cc];
            int pos = base + __popc(m[0] & below) + __popc(m[1] & below) + __popc(m[2] & below) + __popc(m[3] & below);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((m[j] >> lane) & 1u) {
                    if (pos < SEG_MAXACT) act[cc * SEG_MAXACT + pos] = ent[j];
                    ++pos;
                }
            }
            __syncwarp();
            if (lane == 0) nact[cc] = base + __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
            __syncwarp();
        }
      }
    }
    if (lane < CPB) {
        const int cnt = nact[lane];
        const bool gen = ((general >> lane) & 1u) || cnt > SEG_MAXACT;
        if (gen) {
            nact[lane] = -1;
            worklist[atomicAdd(work_count, 1)] = (int)(b0 + lane);
        }
    }
    __syncwarp();
    return stage_any;
}

// VAR = false: plain parameter sweeps (no variant code, fewer registers); VAR = true: QP variant rows.
#ifndef DMX_F32_MINB
#define DMX_F32_MINB 4       // 64 registers (no spills): four 256-thread CTAs per SM
#endif
template <int CPB, int CSMODE, bool VAR>
__global__ void __launch_bounds__(256, DMX_F32_MINB) dmx_frame32_kernel(const Frame32Args a) {
    const int NC = a.n_classes > 0 ? a.n_classes : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Shared memory, carved by byte offsets from the one __shared__ base (no integer round trip on the pointers: the
    // compiler must keep the address space, or every table read in the segment loop becomes a generic LD):
    //   [0, A)  per-warp (cos, sin) table for CSMODE 2;  [A, B)  per-warp active variant entries + counts;
    //   [B', ..) segment tables, staged once per CTA (every warp walks them for every circuit group)
    unsigned char* const smem_base = reinterpret_cast<unsigned char*>(dmx_smem);
    const unsigned off_act = 8u * CPB * NC * 16u;
    const unsigned off_sw = (off_act + 8u * CPB * (SEG_MAXACT + 1) * 4u + 15u) & ~15u;
    double2* cs = reinterpret_cast<double2*>(smem_base) + (size_t)warp * CPB * NC;
    unsigned* act = reinterpret_cast<unsigned*>(smem_base + off_act) + (size_t)warp * CPB * (SEG_MAXACT + 1);
    int* nact = reinterpret_cast<int*>(act + CPB * SEG_MAXACT);
    // Staged form of the weights: the own-coefficient factor of a segment is A = w0 * (moved by the rotation ? cos : 1).
    // Kept as (wc, wk) with A = fma(wc, cos, wk) it needs no select in the segment loop (a select on a double is two
    // FSELs per circuit and segment; ptxas also turns a predicated DMUL into DMUL + 2 FSEL).
    double2* sw = reinterpret_cast<double2*>(smem_base + off_sw);                                   // [nseg][32] (wc, wk)
    double* sw1 = reinterpret_cast<double*>(smem_base + off_sw + (unsigned)a.nseg * 32u * 16u);     // [nseg][32] w1
    unsigned char* spart = smem_base + off_sw + (unsigned)a.nseg * 32u * 24u;
    int* ssel = reinterpret_cast<int*>(smem_base + off_sw + (unsigned)a.nseg * 32u * 25u);
    // Every first-touch load of a launch misses to HBM (the tables are a few KB, but a sweep step starts cold), so the
    // lane constants and the first round's angle are requested BEFORE the staging loop and its barrier: one DRAM round
    // trip for all of them instead of one per dependent level (tables -> barrier -> constants -> class -> angle).
    const double init = a.init[lane];
    const double ew0 = a.eweight[lane];
    const int orig_l = a.orig[lane];
    double x_first = 0.0;
    if (CSMODE == 1 && !VAR && !a.stage_io) {
        const long long bf = ((long long)blockIdx.x * 8 + warp) * CPB + lane;
        if (lane < CPB && bf < a.batch) x_first = a.params[bf * a.n_params + a.cls0.param];
    }
    for (int i = threadIdx.x; i < a.nseg * 32; i += 256) {
        const double2 w = a.w[i];
        const unsigned char pbyte = a.partner[i];
        const bool moved = (pbyte & 0x80u) && a.cs_sel[i >> 5] >= 0;
        sw[i] = moved ? make_double2(w.x, 0.0) : make_double2(0.0, w.x);
        sw1[i] = w.y;
        spart[i] = pbyte;
    }
    for (int i = threadIdx.x; i < a.nseg; i += 256) ssel[i] = a.cs_sel[i];
    // Plain single-class sweeps on mapped host buffers (STAGE_IO): the CTA's 8 * CPB circuits of the first round are consecutive rows, so their
    // parameters come in - and their energies go out - as one contiguous 8 * CPB * 8-byte block per CTA through shared memory
    // (the `cs` area, unused in this mode) instead of one 64-byte request per warp.  It matters when the buffers are mapped
    // host memory (dmx_energy_batch_host zero-copy): eight GPUs issuing 64-byte PCIe transactions saturate the host's
    // transaction rate at ~1.1e10 circuits/s (profiles/r2_e2e_scaling_n8.txt).  Device-resident buffers keep the per-warp
    // accesses: the extra block barrier costs 5 % there (18.6 -> 19.5 us per 65 536 circuits).
    const bool STAGE_IO = CSMODE == 1 && !VAR && a.stage_io;
    double* const io = reinterpret_cast<double*>(smem_base);                 // [8 * CPB] angles in, energies out
    if (STAGE_IO && threadIdx.x < 8 * CPB) {
        const long long b = (long long)blockIdx.x * 8 * CPB + threadIdx.x;
        io[threadIdx.x] = b < a.batch ? a.params[b * a.n_params + a.cls0.param] : 0.0;
    }
    __syncthreads();
    const long long ngroups = (a.batch + CPB - 1) / CPB;
    const long long gwarp = (long long)blockIdx.x * 8 + warp, nwarps = (long long)gridDim.x * 8;
    const int p = orig_l < 0 ? 0 : orig_l;

    // Every warp runs the same number of rounds (rows past the batch are predicated off at the loads and stores): with a
    // trip count that depends only on kernel parameters the compiler can prove the warp converged and emits bare SHFLs
    // (a warp-index-dependent bound costs a WARPSYNC + ENDCOLLECTIVE around every 64-bit shuffle).
    const long long rounds = (ngroups + nwarps - 1) / nwarps;
    for (long long it = 0; it < rounds; ++it) {
        const long long grp = gwarp + it * nwarps;
        const long long b0 = grp * CPB;
        // Variant rows: the non-identity entries of each row are compacted into act[c][] as (stage << 20 | slot << 8 | idx);
        // stage_any has a bit per stage that holds at least one entry of some circuit of the group.
        // The CPB identifier rows of a group are ONE contiguous run of CPB * n_slots bytes (8-byte aligned relative to the
        // table), and almost all of it is zero (identity entries): the warp walks the run as 32-bit words, 128 bytes per step,
        // and only a step that holds a non-zero byte does any per-entry work.  (Round 2, first half: one byte load per slot
        // and row with its own 64-bit address arithmetic - 41 % of the kernel's instructions,
        // profiles/r2_ncu_full_frame32var_kernel_v1.txt.)
        unsigned stage_any = 0u;
        if (VAR) stage_any = frame32_scan_group<CPB>(a.variants, a.n_slots, a.batch, a.slots, a.worklist, a.work_count, b0, lane, act, nact);

        // (cos, sin) after the variant scan: the scan needs the registers
        double2 csr[CPB];
        if (CSMODE == 1) {
            // (Measured and rejected, profiles/r2_frame32s_ab.txt: computing the CTA's 64 (cos, sin) pairs once by 64 fully
            // populated threads ahead of the staging barrier - 18.6 -> 26.7 us per 65 536 circuits: the other six warps wait
            // at the barrier behind a DRAM round trip plus a sincos, and the kernel spills at 64 registers.)
            double sv = 0.0, cv = 1.0;
            if (lane < CPB && b0 + lane < a.batch) {
                const RotInfo ri = a.cls0;
                const double x = (it == 0 && !VAR) ? (STAGE_IO ? io[warp * CPB + lane] : x_first) : a.params[(b0 + lane) * a.n_params + ri.param];
                sincos(ri.scale * x + ri.offset, &sv, &cv);
            }
#pragma unroll
            for (int c = 0; c < CPB; ++c) csr[c] = make_double2(__shfl_sync(0xffffffffu, cv, c), __shfl_sync(0xffffffffu, sv, c));
        } else {
#pragma unroll
            for (int c = 0; c < CPB; ++c) csr[c] = make_double2(1.0, 1.0);
            if (CSMODE == 2) {
                __syncwarp();
                for (int i = lane; i < CPB * NC; i += 32) {
                    const int c = i / NC, cl = i - c * NC;
                    double sv = 0.0, cv = 1.0;
                    if (b0 + c < a.batch) {
                        const RotInfo ri = a.classes[cl];
                        sincos(ri.scale * a.params[(b0 + c) * a.n_params + ri.param] + ri.offset, &sv, &cv);
                    }
                    cs[i] = make_double2(cv, sv);
                }
                __syncwarp();
            }
        }
        double r[CPB];
#pragma unroll
        for (int c = 0; c < CPB; ++c) r[c] = init;

        int sel_n = -1;
        double2 wk_n = make_double2(0.0, 1.0);
        double w1_n = 0.0;
        unsigned pb_n = 0u;
        if (a.nseg > 0) { sel_n = ssel[0]; wk_n = sw[lane]; w1_n = sw1[lane]; pb_n = spart[lane]; }
        for (int k = 0; k <= a.nseg; ++k) {
            if (VAR && ((stage_any >> (k < 31 ? k : 31)) & 1u)) {
#pragma unroll 1
                for (int c = 0; c < CPB; ++c) {
                    const int na = nact[c];
                    for (int e = 0; e < na; ++e) {
                        const unsigned ent = act[c * SEG_MAXACT + e];
                        if ((int)(ent >> 20) == k) {
                            const double fac = frame32_variant_factor(a.slots, a.labels, a.coefs, ent, p);
#pragma unroll
                            for (int cc = 0; cc < CPB; ++cc) r[cc] *= cc == c ? fac : 1.0;
                        }
                    }
                }
            }
            if (k == a.nseg) break;
            const int sel = sel_n;
            const double2 wk = wk_n;
            const double w1 = w1_n;
            const unsigned pb = pb_n;
            {   // tables of the next segment: in flight while this one computes
                const int kn = k + 1 < a.nseg ? k + 1 : k;
                sel_n = ssel[kn]; wk_n = sw[kn * 32 + lane]; w1_n = sw1[kn * 32 + lane]; pb_n = spart[kn * 32 + lane];
            }
            const int src = pb & 31u;
            // sel < 0: constant-coefficient segment (resolved rotation) - the weights carry cos / sin already.  The test
            // on `sel` (shared-memory data) must not become a branch around the shuffles, see `rounds` above: plans
            // without such segments (a.any_const == 0, kernel parameter) skip it, the others use selects.
            if (CSMODE == 0) {          // no parameterised rotation at all (resolved circuits, QP variant batches)
                const double w0 = wk.x + wk.y;
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    const double other = __shfl_sync(0xffffffffu, r[c], src);
                    r[c] = fma(w1, other, w0 * r[c]);
                }
            } else if (!a.any_const) {
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    const double other = __shfl_sync(0xffffffffu, r[c], src);
                    const double2 v = CSMODE == 1 ? csr[c] : cs[c * NC + sel];
                    r[c] = fma(v.y, w1 * other, fma(wk.x, v.x, wk.y) * r[c]);
                }
            } else {
                const bool rot = sel >= 0;
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    const double other = __shfl_sync(0xffffffffu, r[c], src);
                    const double2 v = CSMODE == 1 ? csr[c] : cs[c * NC + (rot ? sel : 0)];
                    const double sm = rot ? v.y : 1.0;          // wk.x = 0 on constant-coefficient segments: cos drops out
                    r[c] = fma(sm, w1 * other, fma(wk.x, v.x, wk.y) * r[c]);
                }
            }
        }
        // energy: transposed butterfly in a fixed order (deterministic).  Each step halves the number of circuits a lane
        // carries: after offsets 16 / 8 / 4 the lane holds the partial sum of circuit (lane >> 2) (CPB = 8) over its 8-lane
        // class, offsets 2 / 1 finish it; CPB - 1 + 2 shuffles instead of 5 CPB.
        {
            double e[CPB];
            if (a.group) {      // coefficient groups: the circuit's own row of energy weights (uniform branch on a kernel parameter)
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    const int g = b0 + c < a.batch ? a.group[b0 + c] : 0;
                    e[c] = a.eweight[g * 32 + lane] * r[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPB; ++c) e[c] = ew0 * r[c];
            }
            int off = 16;
#pragma unroll
            for (int m = CPB / 2; m >= 1; m >>= 1, off >>= 1) {
                const bool hi = lane & off;
#pragma unroll
                for (int j = 0; j < m; ++j) {
                    const double keep = hi ? e[j + m] : e[j], send = hi ? e[j] : e[j + m];
                    e[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
#pragma unroll
            for (; off > 0; off >>= 1) e[0] += __shfl_xor_sync(0xffffffffu, e[0], off);
            constexpr int LPC = 32 / CPB;           // lanes per circuit after the halving steps
            const int c = lane / LPC;
            if (STAGE_IO && it == 0) {
                if ((lane & (LPC - 1)) == 0) io[warp * CPB + c] = e[0];       // every lane of the warp has read its angle by now
            } else if ((lane & (LPC - 1)) == 0 && b0 + c < a.batch && !(VAR && nact[c] < 0)) {
                a.out[b0 + c] = e[0];
            }
        }
    }
    if (STAGE_IO) {
        __syncthreads();
        const long long b = (long long)blockIdx.x * 8 * CPB + threadIdx.x;
        if (threadIdx.x < 8 * CPB && b < a.batch) a.out[b] = io[threadIdx.x];
    }
}

// ---------------------------------------------------------------------------------------------------------
// frame32, sweep form ("frame32s"): parameter sweeps of a plan with ONE rotation class and no variant table (the noisy
// H2 UCCSD energy sweep, BASELINE configs[1]).  A warp owns cw = 32 (or 16 / 8) consecutive circuits: one coalesced
// parameter load, ONE sincos for cw lanes (the 8-circuit kernel above spends a third of its instructions on a sincos that
// only 8 lanes use), then cw / 8 rounds of 8 circuits through the segment tables (read through L1: no staging, no block
// barrier), and one coalesced store of the cw energies.  Table entry per (segment, lane): (wc, wk, w1, partner).
// ---------------------------------------------------------------------------------------------------------
struct Frame32sArgs {
    int nseg, n_params;
    int cw;                       // circuits per warp: 32 (large batches), or 24 / 16 / 8 so that a mid-size batch still fills the SMs
    int stage_io;                 // parameters / energies are mapped HOST memory: one contiguous block per CTA through shared memory
    long long batch;
    const double4* tab;           // [nseg][32]: x = wc, y = wk (own-coefficient factor = fma(wc, cos, wk)), z = w1, w = partner lane (bits)
    RotInfo cls;                  // the single rotation class
    const double* init;           // [32]
    const double* eweight;        // [groups][32]
    const int* group;             // optional coefficient row per circuit
    const double* params;
    double* out;
};

// Four CTAs per SM (126 registers): measured against six (80 registers) and eight (64, spilling) CTAs per SM on the H2 sweep,
// isolated launch 18.4 / 18.6 / 18.9 us, streamed 7.66e9 / 7.00e9 / 6.52e9 circuits/s (profiles/r2_frame32s_ab.txt).
template <int CPB>      // circuits advanced together by one warp: 8 or 16 (twice the ILP, 128 registers)
__global__ void __launch_bounds__(128, 4) dmx_frame32s_kernel(const Frame32sArgs a) {
    constexpr int LPC = 32 / CPB;
    const int lane = threadIdx.x & 31;
    const int nsub = a.cw / CPB;      // kernel parameter: uniform, the shuffles below stay bare

    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long b = w * a.cw + lane;
    const bool mine_row = lane < a.cw && b < a.batch;
    // Segment table -> shared memory, as two planes of double2 ((wc, wk) and (w1, partner)) so that each LDS.128 of the loop
    // below is conflict-free.  All of a thread's loads are independent: one DRAM round trip on a cold launch.  (Reading the
    // table through L1 instead made every segment of every round wait for a load - 20.5 us per 65 536 circuits whatever
    // the circuits per warp; prefetch.global.L1 ahead of the loop was worse still, 24.8 - 30.5 us: profiles/r2_frame32s_ab.txt.)
    double2* const tabA = reinterpret_cast<double2*>(dmx_smem);
    double2* const tabB = tabA + a.nseg * 32;
    // Mapped host buffers (zero-copy end-to-end calls): the CTA's 4 cw circuits are consecutive rows, so their angles come
    // in - and their energies go out - as ONE contiguous block of up to 1 KB per CTA through `io` instead of one partial
    // line per warp (PCIe transaction rate, see dmx_frame32_kernel); the table barrier below covers the inbound block.
    double* const io = reinterpret_cast<double*>(tabB + a.nseg * 32);          // [128]
    const long long row_t = (long long)blockIdx.x * 4 * a.cw + threadIdx.x;
    const bool mine_t = (int)threadIdx.x < 4 * a.cw && row_t < a.batch;
    double x = 0.0;
    if (a.stage_io) { if (mine_t) io[threadIdx.x] = a.params[row_t * a.n_params + a.cls.param]; }
    else if (mine_row) x = a.params[b * a.n_params + a.cls.param];
    {
        const double2* const src = reinterpret_cast<const double2*>(a.tab);
        for (int i = threadIdx.x; i < a.nseg * 32; i += blockDim.x) {
            tabA[i] = __ldg(src + 2 * i);
            tabB[i] = __ldg(src + 2 * i + 1);
        }
    }
    // rows past the batch are predicated off at the load and the store: every warp runs the same instruction stream, so
    // the shuffles need no convergence barrier
    const double init = __ldg(a.init + lane);
    const double ew0 = __ldg(a.eweight + lane);
    const int g_lane = (a.group && mine_row) ? a.group[b] : 0;
    if (a.stage_io) {
        __syncthreads();
        if (mine_row) x = io[(threadIdx.x >> 5) * a.cw + lane];
    }
    double sv, cv;
    sincos(fma(a.cls.scale, x, a.cls.offset), &sv, &cv);
    double eout = 0.0;
    if (!a.stage_io) __syncthreads();       // table staged (the angle load and the sincos overlap the staging loads of the other warps)
#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
        double2 csr[CPB];
#pragma unroll
        for (int c = 0; c < CPB; ++c) csr[c] = make_double2(__shfl_sync(0xffffffffu, cv, sub * CPB + c), __shfl_sync(0xffffffffu, sv, sub * CPB + c));
        double r[CPB];
#pragma unroll
        for (int c = 0; c < CPB; ++c) r[c] = init;
        auto entry = [&](int k) {
            const double2 lo = tabA[k * 32 + lane], hi = tabB[k * 32 + lane];
            return make_double4(lo.x, lo.y, hi.x, hi.y);
        };
        double4 t_n = a.nseg > 0 ? entry(0) : make_double4(0.0, 1.0, 0.0, 0.0);
        for (int k = 0; k < a.nseg; ++k) {
            const double4 t = t_n;
            {   // next segment's entry: in flight while this one computes
                const int kn = k + 1 < a.nseg ? k + 1 : k;
                t_n = entry(kn);
            }
            const int src = __double2loint(t.w) & 31;
#pragma unroll
            for (int c = 0; c < CPB; ++c) {
                const double other = __shfl_sync(0xffffffffu, r[c], src);
                r[c] = fma(csr[c].y, t.z * other, fma(t.x, csr[c].x, t.y) * r[c]);
            }
        }
        // energy: transposed butterfly (fixed order); afterwards lane 4 c holds the energy of circuit c of this round
        double e[CPB];
        if (a.group) {
#pragma unroll
            for (int c = 0; c < CPB; ++c) e[c] = __ldg(a.eweight + __shfl_sync(0xffffffffu, g_lane, sub * CPB + c) * 32 + lane) * r[c];
        } else {
#pragma unroll
            for (int c = 0; c < CPB; ++c) e[c] = ew0 * r[c];
        }
        int off = 16;
#pragma unroll
        for (int m = CPB / 2; m >= 1; m >>= 1, off >>= 1) {
            const bool hi = lane & off;
#pragma unroll
            for (int j = 0; j < m; ++j) {
                const double keep = hi ? e[j + m] : e[j], send = hi ? e[j] : e[j + m];
                e[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
#pragma unroll
        for (; off > 0; off >>= 1) e[0] += __shfl_xor_sync(0xffffffffu, e[0], off);
        const double mine = __shfl_sync(0xffffffffu, e[0], (lane & (CPB - 1)) * LPC);
        if (lane / CPB == sub) eout = mine;
    }
    if (a.stage_io) {
        if (lane < a.cw) io[(threadIdx.x >> 5) * a.cw + lane] = eout;      // each warp overwrites only the angles it read itself
        __syncthreads();
        if (mine_t) a.out[row_t] = io[threadIdx.x];
    } else if (mine_row) {
        a.out[b] = eout;
    }
}

// ---------------------------------------------------------------------------------------------------------
// frame32, thread-per-circuit form ("frame16t"): parameter sweeps of a plan with ONE rotation class, no variant slots and
// at most 16 coefficients that any rotation touches (Plan::f16_*; noisy H2 UCCSD: 16 of the 23 live ones, the other 7 are
// constants of the plan).  One THREAD carries one circuit: its 16 coefficients stay in registers from the initial state
// to the energy, a segment pairs register i with register i ^ mask - the mask is a property of the plan, so each of the
// 16 possible values is its own unrolled code path chosen by a warp-uniform switch - and the segment weights come from
// shared memory as broadcast LDS.128.  No shuffles (the warp-per-circuit kernels spend 2 SHFL.32 per circuit and
// segment and idle 9 of 32 lanes), no per-warp sincos shared by 8 lanes: 4 FP64 instructions per coefficient and segment.
// ---------------------------------------------------------------------------------------------------------
struct Frame16tArgs {
    int nseg, n_params;
    long long batch;
    const void* tab;              // general form: double4 [nseg][16], x = wc, y = wk (own factor = fma(wc, cos, wk)), z = w1 (partner
                                  // factor = w1 * sin); scaled form: double2 [nseg][8], (kappa_i, kappa_j) per pair
    const double* eweight;        // [groups][17]: 16 weights + the group's constant
    const int* group;             // optional coefficient row per circuit
    const double* params;
    double* out;
    RotInfo cls;
    double init[16];
    unsigned char mask[48];
};

template <int M>
__device__ __forceinline__ void f16_segment(double (&r)[16], const double4* __restrict__ T, double c, double s) {
    if (M == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { const double4 t = T[i]; r[i] *= fma(t.x, c, t.y); }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = i ^ M;
            if (j > i) {
                const double4 ti = T[i], tj = T[j];
                const double ri = r[i], rj = r[j];
                // same rounding sequence as the warp-per-circuit kernels: sin * (w1 * partner) + (wc * cos + wk) * own
                r[i] = fma(s, ti.z * rj, fma(ti.x, c, ti.y) * ri);
                r[j] = fma(s, tj.z * ri, fma(tj.x, c, tj.y) * rj);
            }
        }
    }
}

// index of the pair (i, i ^ M), i < i ^ M, in ascending order of i
__host__ __device__ constexpr int f16_pair_index(int i, int M) {
    int n = 0;
    for (int q = 0; q < i; ++q) n += ((q ^ M) > q) ? 1 : 0;
    return n;
}

// Scaled form (Plan::f16_scaled): rho_i' = cos * rho_i + kappa_i * (sin * rho_j); the table holds (kappa_i, kappa_j) per pair in
// the order of this loop, one LDS.128 each.
template <int M>
__device__ __forceinline__ void f16_segment_scaled(double (&r)[16], const double2* __restrict__ Kp, double c, double s) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = i ^ M;
        if (j > i) {
            const double2 kp = Kp[f16_pair_index(i, M)];
            const double ri = r[i], rj = r[j];
            r[i] = fma(kp.x, s * rj, c * ri);
            r[j] = fma(kp.y, s * ri, c * rj);
        }
    }
}

#define DMX_F16_SWITCH(FN, ...)                                                                   \
    switch (a.mask[k]) {                   /* kernel parameter: uniform */                        \
        case 1: FN<1>(__VA_ARGS__); break;   case 2: FN<2>(__VA_ARGS__); break;                   \
        case 3: FN<3>(__VA_ARGS__); break;   case 4: FN<4>(__VA_ARGS__); break;                   \
        case 5: FN<5>(__VA_ARGS__); break;   case 6: FN<6>(__VA_ARGS__); break;                   \
        case 7: FN<7>(__VA_ARGS__); break;   case 8: FN<8>(__VA_ARGS__); break;                   \
        case 9: FN<9>(__VA_ARGS__); break;   case 10: FN<10>(__VA_ARGS__); break;                 \
        case 11: FN<11>(__VA_ARGS__); break; case 12: FN<12>(__VA_ARGS__); break;                 \
        case 13: FN<13>(__VA_ARGS__); break; case 14: FN<14>(__VA_ARGS__); break;                 \
        case 15: FN<15>(__VA_ARGS__); break; default: FN<0>(__VA_ARGS__); break;                  \
    }

template <bool SCALED>
__global__ void __launch_bounds__(256) dmx_frame16t_kernel(const __grid_constant__ Frame16tArgs a) {
    // table: [nseg][16] double4 (general form) or [nseg][8] double2 (scaled form); then the 17 energy weights of group 0
    const int tab_d2 = a.nseg * (SCALED ? 8 : 32);
    double2* const stab = reinterpret_cast<double2*>(dmx_smem);
    double* const sew = reinterpret_cast<double*>(stab + tab_d2);
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = b < a.batch;
    // the angle first: its DRAM round trip overlaps the table staging
    const double x = mine ? a.params[b * a.n_params + a.cls.param] : 0.0;
    const int g = (a.group && mine) ? a.group[b] : 0;
    {
        const double2* const src = reinterpret_cast<const double2*>(a.tab);
        for (int i = threadIdx.x; i < tab_d2; i += blockDim.x) stab[i] = __ldg(src + i);
    }
    if (threadIdx.x < 17) sew[threadIdx.x] = __ldg(a.eweight + threadIdx.x);
    double sv, cv;
    sincos(fma(a.cls.scale, x, a.cls.offset), &sv, &cv);
    double r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = a.init[i];
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < a.nseg; ++k) {
        if (SCALED) {
            const double2* __restrict__ Kp = stab + k * 8;
            DMX_F16_SWITCH(f16_segment_scaled, r, Kp, cv, sv)
        } else {
            const double4* __restrict__ T = reinterpret_cast<const double4*>(stab) + k * 16;
            DMX_F16_SWITCH(f16_segment, r, T, cv, sv)
        }
    }
    double e;
    if (a.group) {
        const double* __restrict__ ew = a.eweight + (size_t)g * 17;
        e = __ldg(ew + 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) e = fma(__ldg(ew + i), r[i], e);
    } else {
        e = sew[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) e = fma(sew[i], r[i], e);
    }
    if (mine) a.out[b] = e;
}

// ---------------------------------------------------------------------------------------------------------
// frame16v: thread-per-ROW form for QP variant batches of a resolved circuit (Plan::f16_const).  The segments have constant
// coefficients, so in the scaled form a segment is ONE DFMA per slot (rho_i' = rho_i + kappa_i rho_j); a row's non-identity
// entries (a handful in 122 bytes) are found by a warp-wide word scan of the 32 contiguous rows the warp owns and handed to
// the owning thread through shared memory, sorted by (stage, slot) so that equal rows give bit-identical energies wherever
// they sit; each entry multiplies the 16 slots and the (otherwise constant) untouched coefficients by one 192-byte row of
// precomputed diagonal factors.  Rows with a non-diagonal correction (R-type / projector basis operations) or more than
// F16V_MAXACT entries go to the worklist of the general tile path, exactly like dmx_frame32_kernel's.
// ---------------------------------------------------------------------------------------------------------
constexpr int F16V_MAXACT = 6;

struct Frame16vArgs {
    int nseg, n_slots;
    long long batch;
    const double2* kappa;         // [nseg][8]: (kappa_i, kappa_j) per pair in loop order
    const double* sweight;        // [groups][17]: energy weights x final scale, + unused
    const double* cstat;          // [groups][8]: energy weight x constant value of the untouched coefficients
    const int* group;
    const uint8_t* variants;
    const SlotDev* slots;
    const int* fac_row;           // [n_slots]
    const double* fac;            // [rows][24]
    double* out;
    int* worklist;
    int* work_count;
    double init[16];
    unsigned char mask[48];
};

template <int M>
__device__ __forceinline__ void f16_segment_const(double (&r)[16], const double2* __restrict__ Kp) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = i ^ M;
        if (j > i) {
            const double2 kp = Kp[f16_pair_index(i, M)];
            const double ri = r[i], rj = r[j];
            r[i] = fma(kp.x, rj, ri);
            r[j] = fma(kp.y, ri, rj);
        }
    }
}

// one non-identity byte of a row: slot `sl`, basis-operation index `idx` -> entry in the thread's own list.  An entry is
// (stage << 12 | slot, row of the correction table): the sort key and everything the segment loop needs, so applying a
// correction costs no dependent table look-up.
__device__ __forceinline__ void f16v_add_entry(const SlotDev* __restrict__ slots, const int2* __restrict__ smeta, unsigned sl, unsigned idx,
                                               uint2* mya, int& cnt, bool& general) {
    const int2 meta = smeta[sl];                            // (stage, first table row) from shared memory
    if (meta.x < 0) return;                                 // slot not present in the circuit
    if (!((slots[sl].diag_ok[idx >> 5] >> (idx & 31u)) & 1u)) { general = true; return; }
    if (cnt < F16V_MAXACT) mya[cnt] = make_uint2(((unsigned)meta.x << 12) | sl, (unsigned)meta.y + idx);
    else general = true;
    ++cnt;
}

// the part of a thread-per-row kernel after the row's entries are known: sort them, evolve, energy
__device__ __forceinline__ void f16v_finish(const Frame16vArgs& a, const double2* __restrict__ skap, const double* __restrict__ sew, uint2* mya, int cnt,
                                            bool general, bool mine, long long b) {
    if (general) {
        if (mine) a.worklist[atomicAdd(a.work_count, 1)] = (int)b;
        return;                                             // no block barrier and no shuffle below: the thread may leave
    }
    // (stage, slot) order (slots are found in ascending order; a later slot may belong to an earlier stage)
    for (int i = 1; i < cnt; ++i) {
        const uint2 v = mya[i];
        int j = i - 1;
        while (j >= 0 && mya[j].x > v.x) { mya[j + 1] = mya[j]; --j; }
        mya[j + 1] = v;
    }

    double r[16], F[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = a.init[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) F[j] = 1.0;
    int next = 0;
    uint2 ent = cnt > 0 ? mya[0] : make_uint2(0xffffffffu, 0u);
#pragma unroll 1
    for (int k = 0; k <= a.nseg; ++k) {
        while ((ent.x >> 12) == (unsigned)k) {
            const double2* __restrict__ row = reinterpret_cast<const double2*>(a.fac + (size_t)ent.y * 24);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 f2 = __ldg(row + i); r[2 * i] *= f2.x; r[2 * i + 1] *= f2.y; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { const double2 f2 = __ldg(row + 8 + j); F[2 * j] *= f2.x; F[2 * j + 1] *= f2.y; }
            ++next;
            ent = next < cnt ? mya[next] : make_uint2(0xffffffffu, 0u);
        }
        if (k == a.nseg) break;
        const double2* __restrict__ Kp = skap + k * 8;
        DMX_F16_SWITCH(f16_segment_const, r, Kp)
    }
    double e = 0.0;
    if (a.group) {
        const int g = mine ? a.group[b] : 0;
        const double* __restrict__ ew = a.sweight + (size_t)g * 17;
        const double* __restrict__ cs = a.cstat + (size_t)g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) e = fma(__ldg(cs + j), F[j], e);
#pragma unroll
        for (int i = 0; i < 16; ++i) e = fma(__ldg(ew + i), r[i], e);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) e = fma(sew[16 + j], F[j], e);
#pragma unroll
        for (int i = 0; i < 16; ++i) e = fma(sew[i], r[i], e);
    }
    if (mine) a.out[b] = e;
}

__global__ void __launch_bounds__(128, 7) dmx_frame16v_kernel(const __grid_constant__ Frame16vArgs a) {
    double2* const skap = reinterpret_cast<double2*>(dmx_smem);                 // [nseg][8]
    double* const sew = reinterpret_cast<double*>(skap + a.nseg * 8);           // [16] weights + [8] static terms of group 0 (+ pad)
    uint2* const act = reinterpret_cast<uint2*>(sew + 26);                      // [128][F16V_MAXACT]
    int2* const smeta = reinterpret_cast<int2*>(act + 128 * F16V_MAXACT);       // [n_slots] (stage, first correction-table row)
    unsigned* const srow = reinterpret_cast<unsigned*>(smeta + ((a.n_slots + 1) & ~1)); // the CTA's 128 rows, as loaded (+ slack)
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * 128;
    const long long b = row0 + tid;
    const bool mine = b < a.batch;
    const unsigned S = (unsigned)a.n_slots;
    {
        // ---- the CTA's rows are ONE contiguous run of 128 * n_slots bytes: coalesced 16-byte loads into shared memory ----
        const long long rows_here = a.batch - row0 < 128 ? a.batch - row0 : 128;
        const unsigned nbytes = (unsigned)rows_here * S;
        const uint8_t* const gbase = a.variants + row0 * (long long)S;
        if ((reinterpret_cast<unsigned long long>(gbase) & 15ull) == 0ull) {
            const unsigned n16 = nbytes >> 4;
            const uint4* const src = reinterpret_cast<const uint4*>(gbase);
            uint4* const dst = reinterpret_cast<uint4*>(srow);
            for (unsigned i = tid; i < n16; i += 128) dst[i] = __ldcs(src + i);
            uint8_t* const sb = reinterpret_cast<uint8_t*>(srow);
            for (unsigned i = (n16 << 4) + tid; i < nbytes; i += 128) sb[i] = gbase[i];
        } else {
            uint8_t* const sb = reinterpret_cast<uint8_t*>(srow);
            for (unsigned i = tid; i < nbytes; i += 128) sb[i] = gbase[i];
        }
    }
    for (int i = tid; i < a.nseg * 8; i += 128) skap[i] = __ldg(a.kappa + i);
    for (int i = tid; i < a.n_slots; i += 128) smeta[i] = make_int2(a.slots[i].seg, __ldg(a.fac_row + i));
    if (tid < 16) sew[tid] = __ldg(a.sweight + tid);
    else if (tid < 24) sew[tid] = __ldg(a.cstat + (tid - 16));
    __syncthreads();

    // ---- the thread's own row: n_slots bytes at byte offset tid * n_slots, read as aligned words and funnel-shifted; almost
    // all of it is zero, so the loop only remembers the last non-zero word (predicated, no branch) and the entries are
    // decoded after it, by all the lanes that have one at the same time ----
    uint2* const mya = act + tid * F16V_MAXACT;
    int cnt = 0;
    bool general = false;
    {
        const unsigned o = (unsigned)tid * S, sh = (o & 3u) * 8u;
        const unsigned* const wbase = srow + (o >> 2);
        const unsigned nwords = (S + 3u) >> 2;
        const unsigned tail = S & 3u;
        const unsigned tail_mask = tail ? (1u << (8u * tail)) - 1u : 0xffffffffu;
        unsigned lo = mine ? wbase[0] : 0u, nzw = 0u, nzi = 0u, nnz = 0u;
#pragma unroll 4
        for (unsigned i = 0; i < nwords; ++i) {
            const unsigned hi = mine ? wbase[i + 1] : 0u;
            unsigned w = __funnelshift_r(lo, hi, sh);
            if (i + 1 == nwords) w &= tail_mask;
            lo = hi;
            const bool nz = w != 0u;
            nzw = nz ? w : nzw;
            nzi = nz ? i : nzi;
            nnz += nz ? 1u : 0u;
        }
        if (nnz == 1u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned idx = (nzw >> (8 * j)) & 255u;
                if (idx) f16v_add_entry(a.slots, smeta, 4u * nzi + j, idx, mya, cnt, general);
            }
        } else if (nnz > 1u) {
            lo = wbase[0];
#pragma unroll 1
            for (unsigned i = 0; i < nwords; ++i) {
                const unsigned hi = wbase[i + 1];
                unsigned w = __funnelshift_r(lo, hi, sh);
                if (i + 1 == nwords) w &= tail_mask;
                lo = hi;
                if (!w) continue;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    const unsigned idx = (w >> (8 * j)) & 255u;
                    if (idx) f16v_add_entry(a.slots, smeta, 4u * i + j, idx, mya, cnt, general);
                }
            }
        }
    }
    f16v_finish(a, skap, sew, mya, cnt, general, mine, b);
}

// ---------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void dmx_energy_kernel(const double* __restrict__ state, int n, long long batch, const uint32_t* __restrict__ term_idx,
                                  const double* __restrict__ term_coef, int n_terms, const int* __restrict__ group,
                                  double* __restrict__ out, long long out_offset) {
    const int lane = threadIdx.x & 31;
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= batch) return;
    const double* r = state + (w << (2 * n));
    const double* __restrict__ tc = term_coef + (group ? (size_t)group[out_offset + w] * n_terms : 0);
    double e = 0.0;
    for (int t = lane; t < n_terms; t += 32) e += tc[t] * r[term_idx[t]];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if (lane == 0) out[out_offset + w] = e;
}

__device__ __forceinline__ unsigned spread_bits(unsigned v) {  // bit b -> bit 2b
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// rho[i][j] = 2^-n sum_z r[idx(x=i^j, z)] * i^{ny} (-1)^{popc(i&z&~x) + popc(~i&x&z)}
__global__ void dmx_pauli_to_rho(const double* __restrict__ pauli, int n, long long batch, double2* __restrict__ rho) {
    const unsigned dim = 1u << n;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)dim * dim;
    if (e >= batch * per) return;
    long long b = e / per;
    unsigned ij = (unsigned)(e - b * per);
    unsigned i = ij >> n, jc = ij & (dim - 1), x = i ^ jc;
    const double* r = pauli + (b << (2 * n));
    const unsigned xs = spread_bits(x) << 1;
    double re = 0.0, im = 0.0;
    for (unsigned z = 0; z < dim; ++z) {
        double v = r[xs | spread_bits(z)];
        int ny = __popc(x & z);
        int sg = (__popc(i & z & ~x) + __popc(~i & x & z)) & 1;
        if (sg) v = -v;
        switch (ny & 3) {
            case 0: re += v; break;
            case 1: im += v; break;
            case 2: re -= v; break;
            default: im -= v; break;
        }
    }
    const double sc = 1.0 / (double)dim;
    rho[e] = make_double2(re * sc, im * sc);
}

// diag[i] = | 2^-n sum_z r[idx(0,z)] (-1)^{popc(i&z)} |, optionally divided by the sum over i
__global__ void dmx_pauli_to_diag(const double* __restrict__ pauli, int n, long long batch, double* __restrict__ diag, int renorm) {
    __shared__ double red[32];
    __shared__ double total;
    const unsigned dim = 1u << n;
    const long long b = blockIdx.x;
    const double* r = pauli + (b << (2 * n));
    double* d = diag + b * dim;
    double part = 0.0;
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) {
        double s = 0.0;
        for (unsigned z = 0; z < dim; ++z) {
            double v = r[spread_bits(z)];
            s += (__popc(i & z) & 1) ? -v : v;
        }
        s = fabs(s / (double)dim);
        d[i] = s;
        part += s;
    }
    if (!renorm) return;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
        total = t;
    }
    __syncthreads();
    const double inv = 1.0 / total;
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) d[i] *= inv;
}

constexpr int QP_BLOCKS = 148, QP_THREADS = 256;

__global__ void dmx_qp_reduce_stage1(const double* __restrict__ v, const double* __restrict__ w, const int* __restrict__ cnt,
                                     long long batch, double* __restrict__ partial) {
    __shared__ double red[3][QP_THREADS / 32];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    // fixed assignment of elements to threads -> bit-reproducible result for a given batch
    for (long long i = (long long)blockIdx.x * QP_THREADS + threadIdx.x; i < batch; i += (long long)QP_BLOCKS * QP_THREADS) {
        double x = w[i] * v[i];
        s0 += x;
        s1 += x * x;
        s2 += cnt ? (double)cnt[i] : 1.0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; red[2][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < QP_THREADS / 32; ++k) t += red[threadIdx.x][k];
        partial[blockIdx.x * 3 + threadIdx.x] = t;
    }
}

__global__ void dmx_qp_reduce_stage2(const double* __restrict__ partial, double* __restrict__ out3) {
    // one warp, fixed order: lane l sums blocks l, l + 32, ...; then a butterfly
    const int lane = threadIdx.x & 31;
    double t[3] = {0.0, 0.0, 0.0};
    for (int k = lane; k < QP_BLOCKS; k += 32) {
#pragma unroll
        for (int j = 0; j < 3; ++j) t[j] += partial[k * 3 + j];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j) t[j] += __shfl_xor_sync(0xffffffffu, t[j], off);
    }
    if (lane < 3) out3[lane] = lane == 0 ? t[0] : (lane == 1 ? t[1] : t[2]);
}

__global__ void dmx_fp64_fma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------------------------------------------------
// QP variant sampler: Philox4x32-10 counter RNG, one warp per sample, lanes over gate slots.
// counter = (sample index lo, sample index hi, slot / 4, 0), key = seed  ->  one 32-bit word per slot, u = (word + 0.5) / 2^32
// (entries whose |qp| / sum|qp| is below ~2^-32 are therefore quantised away);
// idx = first entry of the slot's cumulative |qp| table with cdf > u (np.random.choice(p=|qp|/sum|qp|),
// src/error_mitigation.py:326-335); sign = sign prod qp[idx] (:321-323).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned (&c)[4], unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// One warp per sample.  One Philox call serves FOUR consecutive gate slots: counter (sample lo, sample hi, slot / 4, 0), key = seed,
// slot s takes word s % 4 as u = (word + 0.5) / 2^32 and idx = first cdf entry > u (cdf[last] = 1).  Entry 0 (the identity
// correction, probability ~ 1 - O(p)) is tested first; only the rare other draws run the binary search.  The identity test
// runs on integers: cdf[0] > (word + 0.5) / 2^32  <=>  word < ceil(cdf[0] * 2^32 - 0.5) =: thr0 (every step exact in double),
// which takes four int -> double conversions and FP64 compares per lane off the FP64 pipe.
__global__ void __launch_bounds__(256) dmx_qp_sample_kernel(int n_slots, const int* __restrict__ n_choices, const double* __restrict__ cdf,
                                     const unsigned char* __restrict__ sgn /* 0: +, 1: -, 2: zero */, int stride,
                                     const ulonglong2* __restrict__ thr0 /* [ceil(S/4)*2]: identity threshold of every slot, see below (padding: 2^32) */,
                                     const unsigned* __restrict__ sgn0 /* [ceil(S/4)]: sign code of entry 0, 4 slots per word */,
                                     const uint4* __restrict__ thr32 /* [ceil(S/4)]: thr0 - 1 as 32-bit words (fast test: word <= thr32) */,
                                     const unsigned* __restrict__ summ /* [ceil(S/4)]: bit 0 sign parity, bit 1 zero weight of the four identity
                                                                          entries, bit 2: some threshold is 0 (no fast test) */,
                                     unsigned long long seed, unsigned long long first_sample, long long batch,
                                     unsigned char* __restrict__ variants, double* __restrict__ signs,
                                     const int* __restrict__ sample_list = nullptr /* list mode: row b is sample first_sample + sample_list[b] ... */,
                                     const int* __restrict__ list_count = nullptr /* ... for b < *list_count (device-side count) */) {
    const int lane = threadIdx.x & 31;
    if (list_count) batch = *list_count;
    // Persistent warps walk the samples with a grid stride and keep the tables of their first slot group (every group when
    // the circuit has <= 128 slots) in registers.  ncu (profiles/r2_ncu_full_sampler_kernel_v1.txt): 260 instructions per
    // sample, ALU pipe 56 % - the kernel is instruction bound, so the common case is kept short: ONE chained 32-bit test
    // says that all four slots of a lane drew the identity (99.4 % of the lane-groups at p = 1e-3), and then the signs
    // come from a per-group summary word; only the other lanes look at slots one by one.
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    uint4 q0 = make_uint4(~0u, ~0u, ~0u, ~0u);
    unsigned sum0 = 0u;
    if (4 * lane < n_slots) { q0 = thr32[lane]; sum0 = summ[lane]; }
    for (long long b = warp0; b < batch; b += nwarps) {
        const unsigned long long gidx = first_sample + (unsigned long long)(sample_list ? (long long)sample_list[b] : b);
        unsigned char* const out = variants + b * n_slots;
        const unsigned misalign = (unsigned)(reinterpret_cast<unsigned long long>(out) & 3ull);      // 0: word stores, 2: half-word stores
        unsigned neg = 0u, zero = 0u;
        for (int g = lane; 4 * g < n_slots; g += 32) {
            unsigned c[4] = {(unsigned)gidx, (unsigned)(gidx >> 32), (unsigned)g, 0u};
            philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
            const uint4 q = g == lane ? q0 : thr32[g];
            const unsigned sm = g == lane ? sum0 : summ[g];
            unsigned packed = 0u;
            if (c[0] <= q.x && c[1] <= q.y && c[2] <= q.z && c[3] <= q.w && !(sm & 4u)) {
                neg ^= sm & 1u;
                zero |= (sm >> 1) & 1u;
            } else {
                // slot by slot: entry 0 from the compact tables, the full rows (stride apart) only for the other draws
                const ulonglong2 fa = thr0[2 * g], fb = thr0[2 * g + 1];
                const unsigned long long first[4] = {fa.x, fa.y, fb.x, fb.y};
                const unsigned s0 = sgn0[g];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int s = 4 * g + i;
                    unsigned sg = (s0 >> (8 * i)) & 255u;
                    if (!((unsigned long long)c[i] < first[i])) {             // never for the padding slots (threshold 2^32)
                        const double u = ((double)c[i] + 0.5) * (1.0 / 4294967296.0);
                        const double* __restrict__ row = cdf + (size_t)s * stride;
                        int hi = n_choices[s] - 1, lo = hi > 0 ? 1 : 0;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (row[mid] > u) hi = mid; else lo = mid + 1;
                        }
                        packed |= (unsigned)lo << (8 * i);
                        sg = sgn[(size_t)s * stride + lo];
                    }
                    neg ^= sg == 1u;
                    zero |= sg == 2u;
                }
            }
            if (4 * g + 3 < n_slots && misalign == 0u) {
                *reinterpret_cast<unsigned*>(out + 4 * g) = packed;
            } else if (4 * g + 3 < n_slots && misalign == 2u) {
                *reinterpret_cast<unsigned short*>(out + 4 * g) = (unsigned short)packed;
                *reinterpret_cast<unsigned short*>(out + 4 * g + 2) = (unsigned short)(packed >> 16);
            } else {
                for (int i = 0; i < 4 && 4 * g + i < n_slots; ++i) out[4 * g + i] = (unsigned char)(packed >> (8 * i));
            }
        }
        neg = __reduce_xor_sync(0xffffffffu, neg);
        zero = __reduce_or_sync(0xffffffffu, zero);
        if (lane == 0 && signs) signs[b] = zero ? 0.0 : (neg ? -1.0 : 1.0);
    }
}

// ---------------------------------------------------------------------------------------------------------
// frame16q: draw AND evaluate - one thread per SAMPLE of the quasi-probability estimator.  The thread runs the sampler's own
// Philox stream (same counters, thresholds, cumulative tables and sign rules as dmx_qp_sample_kernel, so sample b is the row
// dmx_qp_sampler_draw would write), keeps the few non-identity draws as entries and finishes like dmx_frame16v_kernel; the
// identifier table (122 bytes per sample) is never written or read.  Samples the thread-per-row form cannot take are listed;
// dmx_qp_sampled_energies re-draws exactly those into a compact table for the tile kernel.
// ---------------------------------------------------------------------------------------------------------
struct Frame16qArgs {
    Frame16vArgs v;
    const int* n_choices;
    const double* cdf;
    const unsigned char* sgn;
    int stride;
    const ulonglong2* thr0;
    const unsigned* sgn0;
    const uint4* thr32;
    const unsigned* summ;
    unsigned long long seed, first_sample;
    double* signs;
};

constexpr int F16Q_SPILL = 320;      // entries of the overflow list (local memory): one per slot of the largest plan taken without a worklist

// DEFER = true: samples with a non-diagonal draw or more than F16V_MAXACT entries are listed for the general path.
// DEFER = false (the host has checked that no drawable basis operation of the sampler is non-diagonal in this plan): nothing
// is listed; the one-in-10^10 row with more than F16V_MAXACT non-identity draws continues in a local-memory list.
template <bool DEFER>
__global__ void __launch_bounds__(128, 7) dmx_frame16q_kernel(const __grid_constant__ Frame16qArgs q) {
    const Frame16vArgs& a = q.v;
    const int ng = (a.n_slots + 3) >> 2;
    double2* const skap = reinterpret_cast<double2*>(dmx_smem);                 // [nseg][8]
    double* const sew = reinterpret_cast<double*>(skap + a.nseg * 8);           // [16] + [8] (+ pad)
    uint2* const act = reinterpret_cast<uint2*>(sew + 26);                      // [128][F16V_MAXACT]
    int2* const smeta = reinterpret_cast<int2*>(act + 128 * F16V_MAXACT);       // [n_slots]
    uint4* const sthr = reinterpret_cast<uint4*>(smeta + ((a.n_slots + 1) & ~1)); // [ng] identity thresholds of four slots
    unsigned* const ssum = reinterpret_cast<unsigned*>(sthr + ng);              // [ng]
    const int tid = threadIdx.x;
    const long long b = (long long)blockIdx.x * 128 + tid;
    const bool mine = b < a.batch;
    for (int i = tid; i < a.nseg * 8; i += 128) skap[i] = __ldg(a.kappa + i);
    for (int i = tid; i < a.n_slots; i += 128) smeta[i] = make_int2(a.slots[i].seg, __ldg(a.fac_row + i));
    for (int i = tid; i < ng; i += 128) { sthr[i] = __ldg(q.thr32 + i); ssum[i] = __ldg(q.summ + i); }
    if (tid < 16) sew[tid] = __ldg(a.sweight + tid);
    else if (tid < 24) sew[tid] = __ldg(a.cstat + (tid - 16));
    __syncthreads();
    if (!mine) return;

    uint2* mya = act + tid * F16V_MAXACT;
    uint2 spill[DEFER ? 1 : F16Q_SPILL];
    int cnt = 0;
    bool general = false;
    unsigned neg = 0u, zero = 0u;
    const unsigned long long gidx = q.first_sample + (unsigned long long)b;
    // a lane-group (four slots, one Philox call) that is not all-identity - 0.6 % of them at p = 1e-3 - needs a search of the
    // cumulative tables in global memory: a chain of dependent loads.  Taken inside the loop, the warp's 32 samples would
    // pay those chains one after the other; the loop only notes the group (up to four per sample, packed in a word) and
    // the chains run after it, for all lanes that have one at the same time.
    auto slow_group = [&](int g, const unsigned (&c)[4]) {
        const ulonglong2 fa = q.thr0[2 * g], fb = q.thr0[2 * g + 1];
        const unsigned long long first[4] = {fa.x, fa.y, fb.x, fb.y};
        const unsigned s0 = q.sgn0[g];
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int s = 4 * g + i;
            unsigned sg = (s0 >> (8 * i)) & 255u;
            const unsigned ci = i == 0 ? c[0] : (i == 1 ? c[1] : (i == 2 ? c[2] : c[3]));
            const unsigned long long fi = i == 0 ? first[0] : (i == 1 ? first[1] : (i == 2 ? first[2] : first[3]));
            if (!((unsigned long long)ci < fi)) {             // never for the padding slots (threshold 2^32)
                const double u = ((double)ci + 0.5) * (1.0 / 4294967296.0);
                const double* __restrict__ row = q.cdf + (size_t)s * q.stride;
                int hi = q.n_choices[s] - 1, lo = hi > 0 ? 1 : 0;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (row[mid] > u) hi = mid; else lo = mid + 1;
                }
                sg = q.sgn[(size_t)s * q.stride + lo];
                if (lo) {
                    if (!DEFER && cnt == F16V_MAXACT && mya != spill) {          // the shared-memory list is full: continue in local memory
                        for (int e = 0; e < F16V_MAXACT; ++e) spill[e] = mya[e];
                        mya = spill;
                    }
                    if (!DEFER && mya == spill) {
                        const int2 meta = smeta[s];
                        if (meta.x >= 0) {
                            if (!((a.slots[s].diag_ok[lo >> 5] >> (lo & 31)) & 1u)) general = true;
                            else if (cnt < F16Q_SPILL) mya[cnt++] = make_uint2(((unsigned)meta.x << 12) | (unsigned)s, (unsigned)meta.y + (unsigned)lo);
                            else general = true;
                        }
                    } else {
                        f16v_add_entry(a.slots, smeta, (unsigned)s, (unsigned)lo, mya, cnt, general);
                    }
                }
            }
            neg ^= sg == 1u;
            zero |= sg == 2u;
        }
    };
    unsigned pend = 0u;
    int npend = 0;
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
        unsigned c[4] = {(unsigned)gidx, (unsigned)(gidx >> 32), (unsigned)g, 0u};
        philox4x32_10(c, (unsigned)q.seed, (unsigned)(q.seed >> 32));
        const uint4 t = sthr[g];
        const unsigned sm = ssum[g];
        if (c[0] <= t.x && c[1] <= t.y && c[2] <= t.z && c[3] <= t.w && !(sm & 4u)) {
            neg ^= sm & 1u;
            zero |= (sm >> 1) & 1u;
        } else if (npend < 4 && g < 256) {
            pend |= (unsigned)g << (8 * npend);
            ++npend;
        } else {
            slow_group(g, c);
        }
    }
#pragma unroll 1
    for (int e = 0; e < npend; ++e) {
        const int g = (int)((pend >> (8 * e)) & 255u);
        unsigned c[4] = {(unsigned)gidx, (unsigned)(gidx >> 32), (unsigned)g, 0u};
        philox4x32_10(c, (unsigned)q.seed, (unsigned)(q.seed >> 32));
        slow_group(g, c);
    }
    q.signs[b] = zero ? 0.0 : (neg ? -1.0 : 1.0);
    if (!DEFER && general) {            // cannot happen with tables the host has checked; never pass a wrong number on silently
        a.out[b] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    f16v_finish(a, skap, sew, mya, cnt, general, mine, b);
}

// ---------------------------------------------------------------------------------------------------------
// Shot sampler (K4, sampled mode): DensityMatrixSimulator.run evolves once and draws `shots` computational-basis outcomes
// from |diag rho| renormalised; measure_observable (src/data_containers/model_hydrogen.py:114-175) turns the histogram into
// parity expectations 2 P(even) - 1.  One warp per row of probabilities: the cumulative distribution is built once in
// shared memory (sequential double sum, like numpy.cumsum), the lanes draw the shots - Philox4x32-10, counter (shot / 4 lo,
// shot / 4 hi, row lo, row hi), key = seed, word (shot % 4) as u = (word + 0.5) / 2^32, outcome = first cdf entry > u - into
// a shared-memory histogram, and every requested parity mask is evaluated on that one histogram (the reference derives
// all its Z / ZZ expectations from the same shots too).  Bit-exact against tests/philox_ref.py::shot_expectations.
// ---------------------------------------------------------------------------------------------------------
constexpr int SHOT_MAX_DIM = 128;      // 2^7 outcomes: every circuit the whole-state kernels handle

__global__ void __launch_bounds__(128) dmx_shot_expect_kernel(const double* __restrict__ probs, long long rows, int dim, long long shots_all,
                                                              const long long* __restrict__ row_shots, const unsigned* __restrict__ masks, int n_masks, unsigned long long seed,
                                                              unsigned long long first_row, double* __restrict__ expect) {
    __shared__ double cdf[4][SHOT_MAX_DIM];
    __shared__ int hist[4][SHOT_MAX_DIM];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= rows) return;
    const double* __restrict__ p = probs + r * dim;
    const long long shots = row_shots ? row_shots[r] : shots_all;
    for (int i = lane; i < dim; i += 32) hist[warp][i] = 0;
    if (lane == 0) {
        double acc = 0.0;
        for (int i = 0; i < dim; ++i) { acc += p[i]; cdf[warp][i] = acc; }
        const double total = acc;
        for (int i = 0; i < dim; ++i) cdf[warp][i] = cdf[warp][i] / total;
        cdf[warp][dim - 1] = 1.0;
    }
    __syncwarp();
    const unsigned long long row = first_row + (unsigned long long)r;
    const long long calls = (shots + 3) / 4;
    for (long long c = lane; c < calls; c += 32) {
        unsigned w[4] = {(unsigned)c, (unsigned)((unsigned long long)c >> 32), (unsigned)row, (unsigned)(row >> 32)};
        philox4x32_10(w, (unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (4 * c + j >= shots) break;
            const double u = ((double)w[j] + 0.5) * (1.0 / 4294967296.0);
            int lo = 0, hi = dim - 1;            // first index with cdf > u (cdf[dim - 1] = 1 > u)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cdf[warp][mid] > u) hi = mid; else lo = mid + 1;
            }
            atomicAdd(&hist[warp][lo], 1);
        }
    }
    __syncwarp();
    for (int m = lane; m < n_masks; m += 32) {
        const unsigned mask = masks[m];
        long long even = 0;
        for (int i = 0; i < dim; ++i) if (!(__popc((unsigned)i & mask) & 1)) even += hist[warp][i];
        expect[r * n_masks + m] = 2.0 * ((double)even / (double)shots) - 1.0;
    }
}

// device tables of a QP sampler (dmx_qp_sampler_create): [cdf | first cdf entries | n_choices | first sign codes | sign codes]
struct dmx_qp_sampler {
    int n_slots = 0, stride = 0, device = -1, sm_count = 0;
    char* d = nullptr;
    size_t cb = 0, fb = 0, nb = 0, gb = 0, sb = 0, qb = 0;      // byte sizes of the sub-tables (each 16-byte aligned)
    std::vector<uint32_t> drawable;                             // host: [n_slots][8] bit idx = some 32-bit word draws idx
};

// ---------------------------------------------------------------------------------------------------------
// device mirror of a plan
// ---------------------------------------------------------------------------------------------------------
struct DevPlan {
    int device = -1;
    SweepDev* sweeps = nullptr;
    MicroOp* mus = nullptr;
    ChainDev* chains = nullptr;
    ChainFactor* factors = nullptr;
    Phase* phases = nullptr;
    double* coefs = nullptr;
    RotInfo* rots = nullptr;
    uint32_t* term_idx = nullptr;
    double* term_coef = nullptr;
    // segment (fixed frame)
    double2* f_w = nullptr;
    int4* f_segs = nullptr;
    RotInfo* f_classes = nullptr;
    double* f_init = nullptr;
    double* f_eweight = nullptr;
    int* f_eslot = nullptr;
    int n_eterms = 0, n_smem_segs = 0;
    // compact frame (frame32)
    double4* g_tab4 = nullptr;     // frame32s: (wc, wk, w1, partner) per (segment, lane)
    double4* h_tab16 = nullptr;    // frame16t: (wc, wk, w1, -) per (segment, slot)
    double* h_ew16 = nullptr;      // frame16t: [groups][17]
    double* h_kap16 = nullptr;     // frame16t, scaled form: [nseg][16]
    double* h_sw16 = nullptr;      // frame16t, scaled form: [groups][17]
    double* h_cstat = nullptr;     // frame16v: [groups][8]
    int* h_facrow = nullptr;       // frame16v: [n_slots]
    double* h_fac = nullptr;       // frame16v: [rows][24]
    double2* g_w = nullptr; unsigned char* g_partner = nullptr; int* g_cs_sel = nullptr; double* g_init = nullptr; double* g_ew = nullptr; int* g_orig = nullptr;
    SlotDev* slots = nullptr;
    uint8_t* labels = nullptr;
    // host-buffer API staging
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;   // host-buffer API: compute / copy-in / copy-out
    static constexpr int MAX_CHUNKS = 8;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {};
    int f32s_warps_per_sm = 0;           // resident warps of dmx_frame32s_kernel with this plan's table (occupancy query, once)
    static constexpr int MAX_XS = 8;     // streams the asynchronous host calls rotate over
    cudaStream_t xs[MAX_XS] = {};
    unsigned async_calls = 0;
    void* h_stage = nullptr; size_t h_cap = 0;
    void* d_stage = nullptr; size_t d_cap = 0;
    std::mutex mu;       // host-buffer API staging (held for the whole dmx_energy_batch_host call)
    int sm_count = 148;
    // streaming passes in the form dmx_stream_kernel executes (dmx_stream.cuh); !ok -> generic dmx_tile_kernel
    struct PassX {
        bool ok = false;
        std::vector<SweepX> sweeps;  // travel in the kernel parameters
        std::vector<SweepX3> sweeps3; // triple-sweep plans
        std::vector<double> cfix;     // triple-sweep plans: permutation scale tables for the kernel parameters
        std::vector<MuX> mus;
        bool general = false;        // needs the 4x4 / dense micro-ops
        double* coefs = nullptr;     // device: pass-local fixed coefficients
        int n_sweeps = 0, n_mus = 0, n_coefs = 0, first_chain = 0, n_chains = 0;
        size_t smem = 0;
        // Blackwell staging modes of dmx_stream3_kernel (dmx_stream.cuh: StreamLoadMode / StreamStoreMode)
        int load_mode = LOAD_STAGED, store_mode = STORE_STAGED;
        int first_lp[3] = {0, 0, 0}, first_tl[4] = {0, 0, 0, 0}, last_g[4] = {0, 0, 0, 0};
        uint32_t first_A[4] = {}, first_B[4] = {}, first_C[4] = {}, last_A[4] = {}, last_B[4] = {}, last_C[4] = {};
        uint32_t mbar_off = 0;
    };
    std::vector<PassX> passx;
};

template <class T>
static cudaError_t upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    if (src.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, src.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static std::mutex g_upload_mu;

// Streaming passes -> tables of dmx_stream_kernel.  A pass qualifies when its tile has 6 or 7 qubits, it holds no
// variant micro-ops and its tables fit shared memory beside the tile.
static cudaError_t build_stream_tables(const Plan& p, DevPlan* d) {
    d->passx.assign(p.passes.size(), DevPlan::PassX{});
    if (p.path != 2) return cudaSuccess;
    for (size_t pi = 0; pi < p.passes.size(); ++pi) {
        const Pass& ps = p.passes[pi];
        DevPlan::PassX& px = d->passx[pi];
        const int k = (int)ps.tile_q.size();
        if (k != 6 && k != 7) continue;
        const int threads = k == 7 ? 512 : 128;
        if (ps.n_phases > 0 && !pass_stream_eligible(p, ps)) continue;
        px.first_chain = ps.n_phases > 0 ? p.phases[ps.first_phase].first_chain : 0;
        px.n_chains = 0;
        for (int ph = 0; ph < ps.n_phases; ++ph) px.n_chains += p.phases[ps.first_phase + ph].n_chains;
        std::vector<SweepX> sx;
        std::vector<MuX> mx;
        std::vector<double> lc;                       // pass-local coefficient table
        std::map<std::pair<int, int>, int> where;     // (global offset, length) -> local offset
        auto local = [&](int off, int len) {
            auto key = std::make_pair(off, len);
            auto it = where.find(key);
            if (it != where.end()) return it->second;
            if (lc.size() & 1) lc.push_back(0.0);     // 16-byte alignment of every table
            int at = (int)lc.size();
            lc.insert(lc.end(), p.coefs.begin() + off, p.coefs.begin() + off + len);
            where.emplace(key, at);
            return at;
        };
        bool ok = true;
        std::vector<SweepX3> sx3;
        std::vector<double> cfix;
        std::map<int, int> cfix_where;
        auto scale_table = [&](int off) {      // 64 scales: constant bank while it has room, else shared memory
            auto it = cfix_where.find(off);
            if (it != cfix_where.end()) return it->second;
            int at;
            if (!p.opt_scale_smem && cfix.size() + 64 <= (size_t)STREAM3_CFIX) {
                at = (int)cfix.size();
                cfix.insert(cfix.end(), p.coefs.begin() + off, p.coefs.begin() + off + 64);
            } else {
                at = -(px.n_chains * 16 + local(off, 64)) - 2;
            }
            cfix_where.emplace(off, at);
            return at;
        };
        for (int ph = 0; ph < ps.n_phases && ok && p.triples; ++ph) {
            const Phase& phase = p.phases[ps.first_phase + ph];
            for (int si = 0; si < phase.n_sweeps && ok; ++si) {
                const Sweep3& sw = p.sweeps3[phase.first_sweep + si];
                SweepX3 x{};
                int sorted[3] = {sw.p[0], sw.p[1], sw.p[2]};
                std::sort(sorted, sorted + 3);
                for (int j = 0; j < 3; ++j) x.p[j] = sorted[j];
                x.first_mu = (int)mx.size(); x.n_mu = sw.n_mu;
                auto S = [](uint32_t m) { return swz_hd(m) << 3; };
                // variant tables of a wide side are indexed per thread: shared memory, never the constant bank
                auto variant_table = [&](int off, int nvar) { return -(px.n_chains * 16 + local(off, 64 << nvar)) - 2; };
                x.lscale = sw.load_perm ? ((sw.wide & 1) ? variant_table(sw.lcoef, sw.l_nvar) : scale_table(sw.lcoef)) : -1;
                x.sscale = sw.store_perm ? ((sw.wide & 2) ? variant_table(sw.scoef, sw.s_nvar) : scale_table(sw.scoef)) : -1;
                if (sw.wide) {
                    // Both sides in mask form, over a re-labelled thread index: thread `tid` holds the canonical thread
                    // coordinate t = XOR_j tid_j * u[j].  The lane vectors u[0..3] are chosen so that the four low tid bits map
                    // to four independent bank-pair bits on BOTH sides (conflict-free 8-byte accesses per half-warp); the rest
                    // completes the basis.  Canonical image of thread bit j on a side that is not wide: the j-th non-triple bit.
                    const int NTB = 2 * (k - 3);
                    uint32_t canon[8] = {}, TLm[8], TSm[8];
                    {
                        int j = 0;
                        for (int pos = 0; pos < 2 * k; pos += 2)
                            if (pos != sw.p[0] && pos != sw.p[1] && pos != sw.p[2]) { canon[j++] = 1u << pos; canon[j++] = 1u << (pos + 1); }
                    }
                    for (int j = 0; j < NTB; ++j) {
                        TLm[j] = S((sw.wide & 1) ? sw.ltmask[j] : canon[j]);
                        TSm[j] = S((sw.wide & 2) ? sw.stmask[j] : canon[j]);
                    }
                    auto image = [&](const uint32_t* Tm, unsigned u) { uint32_t v = 0; for (int j = 0; j < NTB; ++j) if ((u >> j) & 1u) v ^= Tm[j]; return v; };
                    auto rank_of = [](std::vector<unsigned> v, unsigned mask) {
                        for (auto& w : v) w &= mask;
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
                    };
                    std::vector<unsigned> u, bl, bs;       // chosen thread vectors and their bank-pair images per side
                    for (int pass2 = 0; pass2 < 2 && u.size() < 4; ++pass2)
                        for (unsigned c = 1; c < (1u << NTB) && u.size() < 4; ++c) {
                            std::vector<unsigned> tu = u, tl = bl, ts = bs;
                            tu.push_back(c); tl.push_back((image(TLm, c) >> 3) & 15u); ts.push_back((image(TSm, c) >> 3) & 15u);
                            if (rank_of(tu, ~0u) != (int)tu.size()) continue;
                            const bool full = rank_of(tl, 15u) == (int)tl.size() && rank_of(ts, 15u) == (int)ts.size();
                            const bool half = rank_of(tl, 15u) == (int)tl.size() || rank_of(ts, 15u) == (int)ts.size();
                            if (full || (pass2 == 1 && half)) { u = tu; bl = tl; bs = ts; }
                        }
                    for (unsigned c = 1; c < (1u << NTB) && (int)u.size() < NTB; ++c) {
                        std::vector<unsigned> tu = u;
                        tu.push_back(c);
                        if (rank_of(tu, ~0u) == (int)tu.size()) u = tu;
                    }
                    uint32_t W[32] = {};
                    W[0] = (uint32_t)sw.wide | 3u; W[1] = (sw.wide & 1) ? (uint32_t)sw.l_nvar : 0u; W[2] = (sw.wide & 2) ? (uint32_t)sw.s_nvar : 0u;
                    for (int j = 0; j < NTB; ++j) { W[4 + j] = image(TLm, u[j]); W[12 + j] = image(TSm, u[j]); }
                    for (int j = 0; j < 4; ++j)
                        for (int b = 0; b < NTB; ++b) {
                            if (__builtin_popcount(u[b] & sw.lvmask[j]) & 1) W[20 + j] |= 1u << b;
                            if (__builtin_popcount(u[b] & sw.svmask[j]) & 1) W[24 + j] |= 1u << b;
                        }
                    if (lc.size() & 1) lc.push_back(0.0);
                    const int at = (int)lc.size();
                    lc.resize(lc.size() + 16);
                    std::memcpy(&lc[at], W, sizeof(W));
                    x.pad = px.n_chains * 16 + at + 1;
                }
                for (int c = 0; c < 4; ++c) {       // register bits (zC, xC, zB, xB, zA, xA) = masks 0..5
                    x.lC[c] = ((c & 1) ? S(sw.lmask[0]) : 0u) ^ ((c & 2) ? S(sw.lmask[1]) : 0u);
                    x.lB[c] = ((c & 1) ? S(sw.lmask[2]) : 0u) ^ ((c & 2) ? S(sw.lmask[3]) : 0u);
                    x.lA[c] = ((c & 1) ? S(sw.lmask[4]) : 0u) ^ ((c & 2) ? S(sw.lmask[5]) : 0u);
                    x.sC[c] = ((c & 1) ? S(sw.smask[0]) : 0u) ^ ((c & 2) ? S(sw.smask[1]) : 0u);
                    x.sB[c] = ((c & 1) ? S(sw.smask[2]) : 0u) ^ ((c & 2) ? S(sw.smask[3]) : 0u);
                    x.sA[c] = ((c & 1) ? S(sw.smask[4]) : 0u) ^ ((c & 2) ? S(sw.smask[5]) : 0u);
                }
                for (int mi = 0; mi < sw.n_mu && ok; ++mi) {
                    const MicroOp& m = p.mus[sw.first_mu + mi];
                    MuX y{};
                    y.kind = m.kind; y.pad = m.a1;
                    switch (m.kind) {
                        case MU3_MAT:
                            y.off = m.a2 ? (phase.first_chain + m.a0 - px.first_chain) * 16 : px.n_chains * 16 + local(m.a0, 16);
                            y.unital = m.a3; px.general |= !m.a3; break;
                        case MU3_DIAG: y.off = px.n_chains * 16 + local(m.a0, 4); break;
                        case MU3_DIAG2: y.off = px.n_chains * 16 + local(m.a0, 16); break;
                        case MU3_MAT2: y.off = px.n_chains * 16 + local(m.a0, 256); px.general = true; break;
                        default: ok = false; break;
                    }
                    mx.push_back(y);
                }
                sx3.push_back(x);
            }
        }
        for (int ph = 0; ph < ps.n_phases && ok && !p.triples; ++ph) {
            const Phase& phase = p.phases[ps.first_phase + ph];
            for (int si = 0; si < phase.n_sweeps && ok; ++si) {
                const SweepDev& sw = p.sweeps[phase.first_sweep + si];
                SweepX x{};
                x.pa = std::min(sw.p0, sw.p1); x.pb = std::max(sw.p0, sw.p1);
                x.first_mu = (int)mx.size(); x.n_mu = sw.n_mu;
                x.lscale = sw.load_perm ? px.n_chains * 16 + local(sw.lcoef, 16) : -1;
                x.sscale = sw.store_perm ? px.n_chains * 16 + local(sw.scoef, 16) : -1;
                x.gtop = swz_hd(insert2_hd(insert2_hd((unsigned)threads, x.pa), x.pb)) << 3;
                auto S = [](uint32_t m) { return swz_hd(m) << 3; };
                for (int c = 0; c < 4; ++c) {
                    x.lB[c] = ((c & 1) ? S(sw.lmask[0]) : 0u) ^ ((c & 2) ? S(sw.lmask[1]) : 0u);
                    x.lA[c] = ((c & 1) ? S(sw.lmask[2]) : 0u) ^ ((c & 2) ? S(sw.lmask[3]) : 0u);
                    x.sB[c] = ((c & 1) ? S(sw.smask[0]) : 0u) ^ ((c & 2) ? S(sw.smask[1]) : 0u);
                    x.sA[c] = ((c & 1) ? S(sw.smask[2]) : 0u) ^ ((c & 2) ? S(sw.smask[3]) : 0u);
                }
                for (int mi = 0; mi < sw.n_mu && ok; ++mi) {
                    const MicroOp& m = p.mus[sw.first_mu + mi];
                    MuX y{};
                    y.kind = m.kind;
                    switch (m.kind) {
                        case MU_CHAIN_A: case MU_CHAIN_B:
                            y.off = (phase.first_chain + m.a0 - px.first_chain) * 16; y.unital = m.a3; px.general |= !m.a3; break;
                        case MU_FIX_A: case MU_FIX_B:
                            y.off = px.n_chains * 16 + local(m.a0, 16); y.unital = m.a3; px.general |= !m.a3; break;
                        case MU_DIAG_A: case MU_DIAG_B: y.off = px.n_chains * 16 + local(m.a0, 4); break;
                        case MU_DIAG2: y.off = px.n_chains * 16 + local(m.a0, 16); break;
                        case MU_MAT2: y.off = px.n_chains * 16 + local(m.a0, 256); px.general = true; break;
                        default: ok = false; break;     // variant micro-ops: generic kernel
                    }
                    mx.push_back(y);
                }
                sx.push_back(x);
            }
        }
        if (!ok || sx.size() + sx3.size() > (size_t)STREAM_MAX_SWEEPS || mx.size() > (size_t)STREAM_MAX_MUS) continue;
        px.n_sweeps = (int)(sx.size() + sx3.size()); px.n_mus = (int)mx.size(); px.n_coefs = (int)lc.size();
        px.smem = ((size_t)32 * threads + (size_t)px.n_chains * 16 + lc.size() + 2) * 8;
        px.mbar_off = (uint32_t)px.smem;      // 8-byte aligned by construction
        px.smem += 16;
        if (p.triples && !sx3.empty()) {
            // First / last sweep of the pass in the LINEAR order of each side (tile qubits sorted by their position in
            // pos_in / pos_out).  Register bits of a sweep = masks over tile-local index bits; both sides only permute bits.
            std::vector<int> std_pos(p.n);
            for (int q = 0; q < p.n; ++q) std_pos[q] = p.n - 1 - q;
            const std::vector<int>& pin = ps.pos_in.empty() ? std_pos : ps.pos_in;
            const std::vector<int>& pout = ps.pos_out.empty() ? std_pos : ps.pos_out;
            // tile-local position (bit index of the code) of every tile qubit
            std::vector<int> tl(p.n, -1);
            for (int i = 0; i < k; ++i) tl[ps.tile_q[i]] = 2 * (k - 1 - i);
            auto side = [&](const std::vector<int>& pos, const Sweep3& sw, const uint32_t* masks, bool store, uint32_t* A, uint32_t* B,
                            uint32_t* C) -> bool {
                std::vector<int> byp = ps.tile_q;                 // tile qubits by ascending position on this side
                std::sort(byp.begin(), byp.end(), [&](int x, int y) { return pos[x] < pos[y]; });
                std::vector<int> rank(p.n, -1);
                for (int i = 0; i < k; ++i) rank[byp[i]] = i;
                if (pos[byp[0]] != 0 || pos[byp[1]] != 1) return false;      // the side has no 128-byte runs
                if (!store) for (int i = 0; i < k; ++i) if (pos[byp[i]] != i) return false;     // bulk / direct load: contiguous tile
                int reg_q[3];
                for (int j = 0; j < 3; ++j) {
                    reg_q[j] = ps.tile_q[k - 1 - sw.p[j] / 2];
                    if (rank[reg_q[j]] < 2) return false;   // a register-resident qubit sits in the low pair
                }
                auto is_reg = [&](int q) { return q == reg_q[0] || q == reg_q[1] || q == reg_q[2]; };
                if (!store) {
                    int r3[3] = {rank[reg_q[0]], rank[reg_q[1]], rank[reg_q[2]]};
                    std::sort(r3, r3 + 3);
                    for (int j = 0; j < 3; ++j) px.first_lp[j] = 2 * r3[j];
                    int j = 0;
                    for (int i = 0; i < k; ++i) if (!is_reg(byp[i])) px.first_tl[j++] = tl[byp[i]];   // thread bit pairs: load-linear order
                } else {
                    // thread identity of the last sweep = non-register qubits in tile-local order; coalescing needs its two
                    // lowest ones to be the low pair of the store side
                    std::vector<int> byt = ps.tile_q;
                    std::sort(byt.begin(), byt.end(), [&](int x, int y) { return tl[x] < tl[y]; });
                    int j = 0;
                    for (int i = 0; i < k; ++i) if (!is_reg(byt[i])) { if (j < 2 && pos[byt[i]] > 1) return false; px.last_g[j++] = 2 * pos[byt[i]]; }
                }
                auto map_mask = [&](uint32_t m) {
                    uint32_t out = 0;
                    for (int t = 0; t < k; ++t)
                        for (int b = 0; b < 2; ++b)
                            if ((m >> (2 * t + b)) & 1u) {
                                const int q = ps.tile_q[k - 1 - t];
                                out |= 1u << (2 * (store ? pos[q] : rank[q]) + b);
                            }
                    return store ? out : out << 3;         // global element index / linear byte offset
                };
                uint32_t mm[6];
                for (int j = 0; j < 6; ++j) mm[j] = map_mask(masks[j]);
                for (int c = 0; c < 4; ++c) {       // register bits (zC, xC, zB, xB, zA, xA) = masks 0..5
                    C[c] = ((c & 1) ? mm[0] : 0u) ^ ((c & 2) ? mm[1] : 0u);
                    B[c] = ((c & 1) ? mm[2] : 0u) ^ ((c & 2) ? mm[3] : 0u);
                    A[c] = ((c & 1) ? mm[4] : 0u) ^ ((c & 2) ? mm[5] : 0u);
                }
                return true;
            };
            const Sweep3& sw_first = p.sweeps3[p.phases[ps.first_phase].first_sweep];
            const Phase& ph_last = p.phases[ps.first_phase + ps.n_phases - 1];
            const Sweep3& sw_last = p.sweeps3[ph_last.first_sweep + ph_last.n_sweeps - 1];
            if (pi > 0 && p.opt_bulk && !sw_first.wide && side(pin, sw_first, sw_first.lmask, false, px.first_A, px.first_B, px.first_C))
                px.load_mode = p.opt_bulk == 2 ? LOAD_DIRECT : LOAD_BULK;
            if (p.opt_direct_store && !sw_last.wide && side(pout, sw_last, sw_last.smask, true, px.last_A, px.last_B, px.last_C))
                px.store_mode = STORE_DIRECT;
        }
        if (px.smem > (k == 7 ? 227u : 56u) * 1024) continue;       // k = 6: keep four CTAs per SM
        px.sweeps = sx; px.sweeps3 = sx3; px.mus = mx; px.cfix = cfix;
        if (lc.empty()) lc.push_back(0.0);
        cudaError_t e = upload(&px.coefs, lc);
        if (e != cudaSuccess) return e;
        px.ok = true;
    }
    if (std::getenv("DMX_VERBOSE")) {
        int n_ok = 0, n_bulk = 0, n_direct_l = 0, n_direct_s = 0;
        for (const auto& px : d->passx) {
            n_ok += px.ok; n_bulk += px.ok && px.load_mode == LOAD_BULK; n_direct_l += px.ok && px.load_mode == LOAD_DIRECT;
            n_direct_s += px.ok && px.store_mode == STORE_DIRECT;
        }
        std::fprintf(stderr, "[dmx] streaming plan: %zu passes, %d on the stream kernel, tile load: %d bulk-async, %d direct; %d direct stores\n",
                     d->passx.size(), n_ok, n_bulk, n_direct_l, n_direct_s);
    }
    return cudaSuccess;
}

static void free_devplan(DevPlan* d);

static int upload_plan(Plan& p, DevPlan* d, int dev);

// The plan's tables live on the device that was current at its first execution; a later call from another device would
// hand kernels pointers of a different GPU, so it is refused (DMX_ERR_STATE).  A failed upload frees what it allocated.
static int ensure_device(Plan& p, DevPlan** out) {
    std::lock_guard<std::mutex> lock(g_upload_mu);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (p.dev) {
        DevPlan* d = (DevPlan*)p.dev;
        if (d->device != dev)
            return fail(DMX_ERR_STATE, "plan tables live on device " + std::to_string(d->device) + " but the current device is " +
                                           std::to_string(dev) + " (a plan is bound to the device of its first execution)");
        *out = d;
        return DMX_OK;
    }
    DevPlan* d = new DevPlan();
    d->device = dev;
    const int rc = upload_plan(p, d, dev);
    if (rc != DMX_OK) { free_devplan(d); return rc; }
    p.dev = d;
    *out = d;
    return DMX_OK;
}

static int upload_plan(Plan& p, DevPlan* d, int dev) {
    CUDA_TRY(cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, dev));
    {
        auto nonempty = [](auto v) { if (v.empty()) v.resize(1); return v; };
        CUDA_TRY(upload(&d->sweeps, nonempty(p.sweeps)));
        CUDA_TRY(upload(&d->mus, nonempty(p.mus)));
        CUDA_TRY(upload(&d->chains, nonempty(p.chains)));
        CUDA_TRY(upload(&d->factors, nonempty(p.factors)));
        CUDA_TRY(upload(&d->phases, nonempty(p.phases)));
    }
    {
        std::vector<double> c = p.coefs;
        if (c.empty()) c.push_back(0.0);
        CUDA_TRY(upload(&d->coefs, c));
    }
    {
        std::vector<RotInfo> r = p.rots;
        if (r.empty()) r.push_back(RotInfo{0, 0, 0.0, 0.0});
        CUDA_TRY(upload(&d->rots, r));
    }
    CUDA_TRY(upload(&d->term_idx, p.term_idx));
    CUDA_TRY(upload(&d->term_coef, p.term_coef_g));
    if (p.seg_eligible) {
        const int K = std::max(p.nseg, 1);
        std::vector<double2> w((size_t)K * 256, make_double2(0, 0));
        std::vector<int4> sg(K, make_int4(0, 1, -2, 0));
        for (int k = 0; k < p.nseg; ++k) {
            sg[k] = make_int4(p.fsegs[k].xor_mask, p.fsegs[k].use_shfl, p.fsegs[k].cs_sel, 0);
            d->n_smem_segs += !p.fsegs[k].use_shfl;
            for (int j = 0; j < 256; ++j) w[(size_t)k * 256 + j] = make_double2(p.f_w[((size_t)k * 256 + j) * 2], p.f_w[((size_t)k * 256 + j) * 2 + 1]);
        }
        CUDA_TRY(upload(&d->f_w, w));
        CUDA_TRY(upload(&d->f_segs, sg));
        std::vector<RotInfo> cl = p.f_classes;
        if (cl.empty()) cl.push_back(RotInfo{0, 0, 0.0, 0.0});
        CUDA_TRY(upload(&d->f_classes, cl));
        CUDA_TRY(upload(&d->f_init, p.f_init));
        CUDA_TRY(upload(&d->f_eweight, p.f_eweight_g));
        std::vector<int> eslot(256, -1);
        int ne = 0;
        for (int j = 0; j < 256; ++j) if (p.f_eweight[j] != 0.0) eslot[j] = ne++;
        d->n_eterms = ne;
        CUDA_TRY(upload(&d->f_eslot, eslot));
        std::vector<SlotDev> sl(std::max(p.n_slots, 1));
        std::vector<uint8_t> lab = p.f_labels;
        if (lab.empty()) lab.assign(256, 0);
        for (int s2 = 0; s2 < p.n_slots; ++s2) {
            sl[s2].seg = p.slots[s2].seg; sl[s2].nq = p.slots[s2].nq; sl[s2].tab_off = p.slots[s2].tab_off; sl[s2].pad = 0;
            std::memcpy(sl[s2].diag_ok, p.slots[s2].diag_ok, sizeof(sl[s2].diag_ok));
        }
        CUDA_TRY(upload(&d->slots, sl));
        CUDA_TRY(upload(&d->labels, lab));
        if (p.f32_n > 0) {
            std::vector<double2> gw((size_t)K * 32, make_double2(0, 0));
            std::vector<int> sel(K, -2);
            for (int k = 0; k < p.nseg; ++k) {
                sel[k] = p.fsegs[k].cs_sel;
                for (int l = 0; l < 32; ++l) gw[(size_t)k * 32 + l] = make_double2(p.f32_w[((size_t)k * 32 + l) * 2], p.f32_w[((size_t)k * 32 + l) * 2 + 1]);
            }
            CUDA_TRY(upload(&d->g_w, gw));
            {
                std::vector<double4> t4((size_t)K * 32, make_double4(0.0, 1.0, 0.0, 0.0));
                for (int k = 0; k < p.nseg; ++k)
                    for (int l = 0; l < 32; ++l) {
                        const unsigned char pb = p.f32_partner[(size_t)k * 32 + l];
                        const bool moved = (pb & 0x80u) && p.fsegs[k].cs_sel >= 0;
                        const double w0 = p.f32_w[((size_t)k * 32 + l) * 2], w1 = p.f32_w[((size_t)k * 32 + l) * 2 + 1];
                        double pw;
                        const long long bits = (long long)(pb & 31u);
                        std::memcpy(&pw, &bits, sizeof(pw));
                        t4[(size_t)k * 32 + l] = moved ? make_double4(w0, 0.0, w1, pw) : make_double4(0.0, w0, w1, pw);
                    }
                CUDA_TRY(upload(&d->g_tab4, t4));
            }
            if (p.f16_n > 0) {
                std::vector<double4> t16((size_t)K * 16);
                for (size_t i = 0; i < t16.size(); ++i)
                    t16[i] = make_double4(p.f16_tab[i * 4], p.f16_tab[i * 4 + 1], p.f16_tab[i * 4 + 2], p.f16_tab[i * 4 + 3]);
                CUDA_TRY(upload(&d->h_tab16, t16));
                CUDA_TRY(upload(&d->h_ew16, p.f16_eweight_g));
                if (p.f16_scaled || p.f16_const) {
                    CUDA_TRY(upload(&d->h_kap16, p.f16_kappa));
                    CUDA_TRY(upload(&d->h_sw16, p.f16_sweight_g));
                }
                if (p.f16_const) {
                    CUDA_TRY(upload(&d->h_cstat, p.f16_cstat_g));
                    CUDA_TRY(upload(&d->h_facrow, p.f16_fac_row));
                    CUDA_TRY(upload(&d->h_fac, p.f16_fac));
                }
            }
            CUDA_TRY(upload(&d->g_partner, p.f32_partner));
            CUDA_TRY(upload(&d->g_cs_sel, sel));
            CUDA_TRY(upload(&d->g_init, p.f32_init));
            CUDA_TRY(upload(&d->g_ew, p.f32_eweight_g));
            CUDA_TRY(upload(&d->g_orig, p.f32_orig));
        }
    }
    CUDA_TRY(build_stream_tables(p, d));
    {
        // stream-ordered scratch (chain matrices, sampler tables) comes from the device's default pool: keep freed blocks
        // cached instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
    }
    return DMX_OK;
}

static void free_device(Plan& p) {
    free_devplan((DevPlan*)p.dev);
    p.dev = nullptr;
}

static void free_devplan(DevPlan* d) {
    if (!d) return;
    cudaFree(d->sweeps); cudaFree(d->mus); cudaFree(d->chains); cudaFree(d->factors); cudaFree(d->phases);
    cudaFree(d->coefs); cudaFree(d->rots); cudaFree(d->term_idx); cudaFree(d->term_coef);
    cudaFree(d->f_w); cudaFree(d->f_segs); cudaFree(d->f_classes); cudaFree(d->f_init); cudaFree(d->f_eweight); cudaFree(d->f_eslot);
    cudaFree(d->slots); cudaFree(d->labels);
    cudaFree(d->h_tab16); cudaFree(d->h_ew16); cudaFree(d->h_kap16); cudaFree(d->h_sw16);
    cudaFree(d->h_cstat); cudaFree(d->h_facrow); cudaFree(d->h_fac);
    cudaFree(d->g_tab4); cudaFree(d->g_w); cudaFree(d->g_partner); cudaFree(d->g_cs_sel); cudaFree(d->g_init); cudaFree(d->g_ew); cudaFree(d->g_orig);
    for (auto& px : d->passx) cudaFree(px.coefs);
    for (int i = 0; i < DevPlan::MAX_XS; ++i) if (d->xs[i]) cudaStreamDestroy(d->xs[i]);
    if (d->h_stage) cudaFreeHost(d->h_stage);
    if (d->d_stage) cudaFree(d->d_stage);
    if (d->stream) {
        cudaStreamDestroy(d->stream); cudaStreamDestroy(d->s_in); cudaStreamDestroy(d->s_out);
        for (int i = 0; i < DevPlan::MAX_CHUNKS; ++i) { cudaEventDestroy(d->ev_in[i]); cudaEventDestroy(d->ev_k[i]); }
    }
    delete d;
}

// ---------------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------------
static void make_runs(const std::vector<int>& qubits, int n, BitRun* runs, int* n_runs) {
    // qubits ascending; local position of qubits[i] = 2*(m-1-i); merge neighbours into runs
    const int m = (int)qubits.size();
    int count = 0;
    int i = m - 1;
    while (i >= 0) {
        int jv = i;
        while (jv - 1 >= 0 && qubits[jv - 1] == qubits[jv] - 1) --jv;
        BitRun r;
        r.lshift = 2 * (m - 1 - i);
        r.gshift = 2 * (n - 1 - qubits[i]);
        r.width = 2 * (i - jv + 1);
        if (count >= 8) throw std::runtime_error("tile qubits form more than 8 runs");
        runs[count++] = r;
        i = jv - 1;
    }
    *n_runs = count;
}

static int pass_max_chains(const Plan& p, const Pass& ps) {
    int m = 1;
    for (int i = 0; i < ps.n_phases; ++i) m = std::max(m, p.phases[ps.first_phase + i].n_chains);
    return m;
}

static int launch_tile(const TileArgs& a, int grid, int threads, int groups_per_thread, size_t smem, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured[3] = {0, 0, 0};
    const int gi = threads == 1024 ? 2 : (groups_per_thread == 2 ? 1 : 0);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured[gi]) {
            const int want = (int)std::max(smem, (size_t)48 * 1024);
            if (gi == 2) CUDA_TRY(cudaFuncSetAttribute(dmx_tile_kernel<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            else if (gi) CUDA_TRY(cudaFuncSetAttribute(dmx_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            else CUDA_TRY(cudaFuncSetAttribute(dmx_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            configured[gi] = want;
        }
    }
    if (a.work_count) {
        // Deferred-row launch behind a frame kernel: almost always the worklist is empty (Pauli-type noise has no R-type /
        // projector corrections).  Programmatic dependent launch lets this grid start while the frame kernel still runs;
        // the kernel waits (griddepcontrol.wait) before it reads the work count, so the ~4 us launch is off the critical path.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (gi == 2) CUDA_TRY(cudaLaunchKernelEx(&cfg, dmx_tile_kernel<1, 1024>, a));
        else if (gi) CUDA_TRY(cudaLaunchKernelEx(&cfg, dmx_tile_kernel<2>, a));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, dmx_tile_kernel<1>, a));
    } else if (gi == 2) dmx_tile_kernel<1, 1024><<<grid, threads, smem, st>>>(a);
    else if (gi) dmx_tile_kernel<2><<<grid, threads, smem, st>>>(a);
    else dmx_tile_kernel<1><<<grid, threads, smem, st>>>(a);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

enum OutMode { OUT_ENERGY = 1, OUT_PAULI = 2 };

static void tile_geometry(unsigned elems, int* threads, int* gpt, int wide_cta = 0) {
    const unsigned ng = elems >> 4;
    if (ng > 512 && wide_cta) { *threads = 1024; *gpt = 1; }
    else if (ng > 512) { *threads = 512; *gpt = 2; }
    else if (ng > 256) { *threads = 512; *gpt = 1; }
    else { *threads = 256; *gpt = 1; }
}

static void fill_common(TileArgs& a, Plan& p, DevPlan* d, const Pass& ps, const double* params, const uint8_t* variants) {
    a.sweeps = d->sweeps; a.mus = d->mus; a.chains = d->chains; a.factors = d->factors; a.phases = d->phases;
    a.first_phase = ps.first_phase; a.n_phases = ps.n_phases; a.max_chains = pass_max_chains(p, ps);
    a.coefs = d->coefs; a.rots = d->rots; a.params = params; a.n_params = p.n_params; a.variants = variants; a.n_slots = p.n_slots;
    a.n = p.n;
}

// Whole-state-in-smem execution (n <= 7).  worklist != nullptr: run only the listed circuits (energy only).
static int run_tile_single(Plan& p, DevPlan* d, long long batch, const double* params, const uint8_t* variants, double* out,
                           int mode, const int* worklist, const int* work_count, cudaStream_t st, const int* group = nullptr,
                           bool variants_compact = false) {
    const Pass& ps = p.passes[0];
    TileArgs a{};
    fill_common(a, p, d, ps, params, variants);
    a.pad = p.n == 1 ? 2 : 0;
    a.k = p.n; a.cpb = p.cpb; a.batch = batch; a.worklist = worklist; a.work_count = work_count;
    a.epilogue = mode; a.term_idx = d->term_idx; a.term_coef = d->term_coef; a.n_terms = (int)p.term_idx.size(); a.out = out;
    a.group = group;
    a.variants_compact = variants_compact ? 1 : 0;
    const unsigned E = (unsigned)a.cpb << (2 * p.n + a.pad);
    int threads, gpt;
    tile_geometry(E, &threads, &gpt, p.opt_tile_threads == 1024);
    const size_t smem = (size_t)E * 8 + (size_t)a.cpb * a.max_chains * 128;
    if (smem > 227 * 1024) return fail(DMX_ERR_UNSUPPORTED, "tile does not fit shared memory");
    long long groups = (batch + a.cpb - 1) / a.cpb;
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / threads, (227 * 1024) / std::max<size_t>(smem, 1)));
    long long cap = (long long)d->sm_count * per_sm;
    int grid = (int)std::max<long long>(1, std::min(groups, cap));
    return launch_tile(a, grid, threads, gpt, smem, st);
}


template <int THREADS, bool GENERAL>
static int launch_stream(const StreamArgs& a, long long grid, size_t smem, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured) {
            const int want = (int)std::max(smem, (size_t)48 * 1024);
            CUDA_TRY(cudaFuncSetAttribute(dmx_stream_kernel<THREADS, GENERAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            CUDA_TRY(cudaFuncSetAttribute(dmx_stream_kernel<THREADS, GENERAL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            configured = want;
        }
    }
    dmx_stream_kernel<THREADS, GENERAL><<<(unsigned)grid, THREADS, smem, st>>>(a);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

template <int THREADS, bool GENERAL>
static int launch_stream3(const StreamArgs3& a, long long grid, size_t smem, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured) {
            const int want = (int)std::max(smem, (size_t)48 * 1024);
            CUDA_TRY(cudaFuncSetAttribute(dmx_stream3_kernel<THREADS, GENERAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            CUDA_TRY(cudaFuncSetAttribute(dmx_stream3_kernel<THREADS, GENERAL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            configured = want;
        }
    }
    dmx_stream3_kernel<THREADS, GENERAL><<<(unsigned)grid, THREADS, smem, st>>>(a);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

// Streaming execution (n > tile_k): state[chunk][4^n] in `state`; all passes over one chunk of circuits.
// Remapped plans ping-pong between `state` and `state2` (a tile's load and store address sets differ, so a pass cannot
// run in place); the last pass always lands in `state`.
static int run_streaming(Plan& p, DevPlan* d, long long chunk, long long circuit_offset, const double* params, const uint8_t* variants,
                         double* state, double* state2, cudaStream_t st, double* chain_scratch = nullptr) {
    NvtxRange nvtx("dmx:streaming_passes");
    if (p.remapped && !state2) return fail(DMX_ERR_INVALID, "remapped plan needs a second state buffer");
    const int total_chains = (int)p.chains.size();
    bool any_x = false;
    for (const auto& px : d->passx) any_x |= px.ok && p.opt_stream_kernel;
    double* chainmats = chain_scratch;
    if (any_x && total_chains > 0) {
        // per-circuit chain matrices of every pass, computed once (stream-ordered scratch: plans may run on several streams)
        if (!chainmats) CUDA_TRY(cudaMallocAsync((void**)&chainmats, (size_t)chunk * total_chains * 128, st));
        const long long items = chunk * total_chains;
        dmx_chain_kernel<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(d->chains, d->factors, d->coefs, d->rots, params, p.n_params,
                                                                          total_chains, chunk, circuit_offset, chainmats);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    int rc = DMX_OK;
    for (size_t pi = 0; pi < p.passes.size() && rc == DMX_OK; ++pi) {
        const Pass& ps = p.passes[pi];
        const int k = (int)ps.tile_q.size();
        std::vector<int> others;
        {
            std::vector<char> in(p.n, 0);
            for (int q : ps.tile_q) in[q] = 1;
            for (int q = 0; q < p.n; ++q) if (!in[q]) others.push_back(q);
        }
        BitRun truns[8], oruns[8];
        int n_truns = 0, n_oruns = 0;
        const long long groups = chunk << (2 * (p.n - k));
        if (groups >= (1LL << 31)) { rc = fail(DMX_ERR_UNSUPPORTED, "streaming chunk too large for one launch"); break; }
        const DevPlan::PassX& px = d->passx[pi];
        if (!ps.pos_in.empty() && !(px.ok && p.opt_stream_kernel)) { rc = fail(DMX_ERR_STATE, "remapped plan needs the stream kernel"); break; }
        if (px.ok && p.opt_stream_kernel) {
            if ((int)others.size() > 8) { rc = fail(DMX_ERR_UNSUPPORTED, "more than 8 non-resident qubits"); break; }
            std::vector<int> std_pos(p.n);
            for (int q = 0; q < p.n; ++q) std_pos[q] = p.n - 1 - q;
            // fields shared by StreamArgs (pair sweeps) and StreamArgs3 (triple sweeps)
            auto fill = [&](auto& a, unsigned threads, unsigned niter) {
                std::copy(px.mus.begin(), px.mus.end(), a.mus);
                a.coefs = px.coefs; a.chainmats = chainmats;
                a.n_sweeps = px.n_sweeps; a.n_mus = px.n_mus; a.n_coefs = px.n_coefs;
                a.total_chains = total_chains; a.first_chain = px.first_chain; a.n_chains = px.n_chains;
                a.n = p.n; a.k = k; a.load = pi > 0;
                if constexpr (std::is_same<std::decay_t<decltype(a)>, StreamArgs3>::value) {
                    a.load_mode = pi > 0 ? px.load_mode : LOAD_INIT;
                    a.store_mode = px.store_mode;
                    for (int j = 0; j < 3; ++j) a.first_lp[j] = px.first_lp[j];
                    for (int j = 0; j < 4; ++j) {
                        a.first_tl[j] = px.first_tl[j]; a.last_g[j] = px.last_g[j];
                        a.first_A[j] = px.first_A[j]; a.first_B[j] = px.first_B[j]; a.first_C[j] = px.first_C[j];
                        a.last_A[j] = px.last_A[j]; a.last_B[j] = px.last_B[j]; a.last_C[j] = px.last_C[j];
                    }
                    a.mbar_off = px.mbar_off;
                    a.tile_bytes = 8u << (2 * k);
                }
                const size_t last = p.passes.size() - 1;
                double* bufs[2] = {state, p.remapped ? state2 : state};
                a.dst = bufs[(last - pi) & 1];
                a.src = bufs[(last - pi + 1) & 1];
                auto side = [&](const std::vector<int>& pos, int& n_gruns, int& n_oruns, int (*gruns)[3], int (*uruns)[3], int (*oruns)[3],
                                uint32_t* giter, uint32_t* siter, uint32_t& pbit) {
                    std::vector<int> tq = ps.tile_q;                  // tile_q[i] sits at tile bits 2*(k-1-i)
                    std::vector<int> tl(p.n, -1);
                    for (int i = 0; i < k; ++i) tl[tq[i]] = k - 1 - i;
                    std::sort(tq.begin(), tq.end(), [&](int x, int y) { return pos[x] < pos[y]; });
                    n_gruns = k;
                    for (int i = 0; i < k; ++i) {
                        gruns[i][0] = 2 * i; gruns[i][1] = 2 * pos[tq[i]]; gruns[i][2] = 2;
                        uruns[i][0] = 2 * i; uruns[i][1] = 2 * tl[tq[i]]; uruns[i][2] = 2;
                    }
                    n_oruns = (int)others.size();
                    for (int i = 0; i < n_oruns; ++i) { oruns[i][0] = 2 * i; oruns[i][1] = 2 * pos[others[i]]; oruns[i][2] = 2; }
                    auto dep = [&](unsigned v, int (*runs)[3]) {
                        unsigned g = 0;
                        for (int i = 0; i < k; ++i) g |= ((v >> runs[i][0]) & 3u) << runs[i][1];
                        return g;
                    };
                    for (unsigned jj = 0; jj < niter; ++jj) { giter[jj] = dep(jj * 2u * threads, gruns); siter[jj] = swz_hd(dep(jj * 2u * threads, uruns)); }
                    pbit = swz_hd(dep(1u, uruns));
                };
                side(ps.pos_in.empty() ? std_pos : ps.pos_in, a.n_gruns_l, a.n_oruns_l, a.gruns_l, a.uruns_l, a.oruns_l, a.giter_l, a.siter_l, a.pbit_l);
                side(ps.pos_out.empty() ? std_pos : ps.pos_out, a.n_gruns_s, a.n_oruns_s, a.gruns_s, a.uruns_s, a.oruns_s, a.giter_s, a.siter_s, a.pbit_s);
                a.init_mask = 0xAAAAAAAAu;
            };
            if (p.triples) {
                static thread_local StreamArgs3 a3;      // ~19 KB of kernel parameters: keep it off the stack
                a3 = StreamArgs3{};
                std::copy(px.sweeps3.begin(), px.sweeps3.end(), a3.sweeps);
                std::copy(px.cfix.begin(), px.cfix.end(), a3.cfix);
                fill(a3, k == 7 ? 256u : 64u, 32u);
                if (k == 7) rc = px.general ? launch_stream3<256, true>(a3, groups, px.smem, st) : launch_stream3<256, false>(a3, groups, px.smem, st);
                else rc = px.general ? launch_stream3<64, true>(a3, groups, px.smem, st) : launch_stream3<64, false>(a3, groups, px.smem, st);
            } else {
                static thread_local StreamArgs a2;
                a2 = StreamArgs{};
                std::copy(px.sweeps.begin(), px.sweeps.end(), a2.sweeps);
                fill(a2, k == 7 ? 512u : 128u, 16u);
                if (k == 7) rc = px.general ? launch_stream<512, true>(a2, groups, px.smem, st) : launch_stream<512, false>(a2, groups, px.smem, st);
                else rc = px.general ? launch_stream<128, true>(a2, groups, px.smem, st) : launch_stream<128, false>(a2, groups, px.smem, st);
            }
            continue;
        }
        try {
            make_runs(ps.tile_q, p.n, truns, &n_truns);
            make_runs(others, p.n, oruns, &n_oruns);
        } catch (const std::exception& e) { rc = fail(DMX_ERR_UNSUPPORTED, e.what()); break; }
        TileArgs a{};
        fill_common(a, p, d, ps, params, variants);
        a.k = k; a.cpb = 1; a.pad = 0; a.batch = chunk; a.state = state; a.circuit_offset = circuit_offset;
        a.load = pi > 0; a.store = 1;
        a.n_truns = n_truns; a.n_oruns = n_oruns;
        for (int i = 0; i < 8; ++i) { a.truns[i] = truns[i]; a.oruns[i] = oruns[i]; }
        const unsigned E = 1u << (2 * a.k);
        int threads, gpt;
        tile_geometry(E, &threads, &gpt);
        const size_t smem = (size_t)E * 8 + (size_t)a.max_chains * 128;
        if (smem > 227 * 1024) { rc = fail(DMX_ERR_UNSUPPORTED, "tile does not fit shared memory"); break; }
        rc = launch_tile(a, (int)groups, threads, gpt, smem, st);
    }
    if (chainmats && !chain_scratch) {
        cudaError_t e = cudaFreeAsync(chainmats, st);
        if (rc == DMX_OK && e != cudaSuccess) rc = fail(DMX_ERR_CUDA, std::string("cudaFreeAsync: ") + cudaGetErrorString(e));
    }
    return rc;
}

template <int CPB, int CSMODE>
static int launch_frame(const FrameArgs& a, int sm_count, cudaStream_t st) {
    const int NC = a.n_classes > 0 ? a.n_classes : 1;
    const int epad = std::max(((a.n_eterms + 31) / 32) * 32, 32);
    size_t smem = (size_t)2 * CPB * 256 * 8 + (size_t)CPB * NC * 16 + (size_t)CPB * epad * 8 + (CPB + 1) * 4 + (size_t)CPB * SEG_MAXACT * 4 + 16;
    if (smem > 227 * 1024) return fail(DMX_ERR_UNSUPPORTED, "segment kernel: tables do not fit shared memory");
    static std::mutex mu;
    static size_t configured = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured) {
            CUDA_TRY(cudaFuncSetAttribute(dmx_frame_kernel<CPB, CSMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
            configured = std::max(smem, (size_t)48 * 1024);
        }
    }
    long long groups = (a.batch + CPB - 1) / CPB;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dmx_frame_kernel<CPB, CSMODE>, 256, smem));
    per_sm = std::max(per_sm, 1);
    int grid = (int)std::max<long long>(1, std::min<long long>(groups, (long long)sm_count * per_sm));
    dmx_frame_kernel<CPB, CSMODE><<<grid, 256, smem, st>>>(a);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

template <int CPB, int CSMODE, bool VAR>
static int launch_frame32_impl(const Frame32Args& a, int sm_count, cudaStream_t st) {
    const int NC = a.n_classes > 0 ? a.n_classes : 1;
    const size_t smem = (size_t)8 * CPB * NC * 16 + (size_t)8 * CPB * (SEG_MAXACT + 1) * 4 + 48 + (size_t)a.nseg * (32 * 24 + 32 + 4);
    if (smem > 200 * 1024) return fail(DMX_ERR_UNSUPPORTED, "frame32 kernel: tables do not fit shared memory");
    static std::mutex mu;
    static size_t configured = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured) {
            CUDA_TRY(cudaFuncSetAttribute(dmx_frame32_kernel<CPB, CSMODE, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
            configured = std::max(smem, (size_t)48 * 1024);
        }
    }
    // one CTA per 8 circuit groups, no occupancy cap: measured faster than a persistent grid (H2 sweep 18.7 vs 20.6 us per
    // 65 536 circuits) - the hardware scheduler balances the tail, and re-staging the 4 KB segment tables per CTA is cheap
    const long long groups = (a.batch + CPB - 1) / CPB;
    const int grid = (int)std::max<long long>(1, std::min<long long>((groups + 7) / 8, 0x7fffffffLL));
    dmx_frame32_kernel<CPB, CSMODE, VAR><<<grid, 256, smem, st>>>(a);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

template <int CPB, int CSMODE>
static int launch_frame32(const Frame32Args& a, int sm_count, cudaStream_t st) {
    // variant rows with one rotation class take the general-class kernel ((cos, sin) in shared memory): holding them in
    // registers next to the variant bookkeeping spills 300 B at 64 registers
    return a.variants ? launch_frame32_impl<CPB, CSMODE == 1 ? 2 : CSMODE, true>(a, sm_count, st) : launch_frame32_impl<CPB, CSMODE, false>(a, sm_count, st);
}


static int run_segment(Plan& p, DevPlan* d, long long batch, const double* params, const uint8_t* variants, double* out, cudaStream_t st,
                       const int* group = nullptr, bool host_mapped = false) {
    FrameArgs a{};
    a.nseg = p.nseg; a.n_params = p.n_params; a.n_slots = p.n_slots; a.n_classes = (int)p.f_classes.size(); a.n_eterms = d->n_eterms;
    a.batch = batch; a.w = d->f_w; a.segs = d->f_segs; a.classes = d->f_classes; a.init = d->f_init; a.eweight = d->f_eweight;
    a.eslot = d->f_eslot; a.params = params; a.variants = variants; a.slots = d->slots; a.labels = d->labels; a.coefs = d->coefs; a.out = out;
    a.group = group;
    // worklist of rows that need the general path: per call and stream-ordered (a plan may run on several streams at once)
    int* wl = nullptr;
    if (variants) {
        CUDA_TRY(cudaMallocAsync((void**)&wl, (size_t)(batch + 1) * sizeof(int), st));
        a.worklist = wl + 1;
        a.work_count = wl;
        CUDA_TRY(cudaMemsetAsync(wl, 0, sizeof(int), st));
    }
    auto done = [&](int rc) {
        if (wl) { const cudaError_t e = cudaFreeAsync(wl, st); if (rc == DMX_OK && e != cudaSuccess) rc = fail(DMX_ERR_CUDA, cudaGetErrorString(e)); }
        return rc;
    };
    const int mode = p.f_classes.empty() ? 0 : (p.f_single_class ? 1 : 2);
    int rc;
    if (p.f32_n > 0 && p.opt_frame32 && p.n_slots < 4096 && p.nseg < 4095) {
        Frame32Args g{};
        g.nseg = p.nseg; g.n_params = p.n_params; g.n_slots = p.n_slots; g.n_classes = (int)p.f_classes.size(); g.batch = batch;
        g.w = d->g_w; g.partner = d->g_partner; g.cs_sel = d->g_cs_sel; g.classes = d->f_classes; g.init = d->g_init; g.eweight = d->g_ew;
        g.orig = d->g_orig; g.params = params; g.variants = variants; g.slots = d->slots; g.labels = d->labels; g.coefs = d->coefs;
        g.out = out; g.worklist = a.worklist; g.work_count = a.work_count; g.group = group;
        g.any_const = 0;
        if (!p.f_classes.empty()) g.cls0 = p.f_classes[0];
        for (int k = 0; k < p.nseg; ++k) g.any_const |= p.fsegs[k].cs_sel < 0;
        g.stage_io = host_mapped ? 1 : 0;
        // Sweep form (dmx_frame32s_kernel) from 2^15 circuits a launch.  Measured on B200 (profiles/r2_frame32s_ab.txt, tables
        // staged in shared memory): the H2 sweep streams at 7.7e9 circuits/s against 6.2e9 for the 8-circuit kernel (whose warps
        // each pay a sincos that 8 lanes use); an isolated 65 536-circuit launch reads 18.4 us with either - that figure is
        // the event pair, the launch latency and one warp's critical path (tools/launch_floor.py), not the kernel.
        // opt_frame32s: 0 never, 1 automatic, 8 / 16 always 32 circuits per warp with that many advanced together,
        // 32 + cw: always cw circuits per warp (cw = 8 / 16 / 24 / 32).
        // QP variant rows of a resolved circuit: one thread per row (dmx_frame16v_kernel); rows it cannot take land on the
        // worklist of the general tile path like the warp kernel's
        const size_t f16v_smem = (size_t)p.nseg * 128 + 26 * 8 + 128 * F16V_MAXACT * 8 + (size_t)((p.n_slots + 1) & ~1) * 8 +
                                 (((size_t)128 * p.n_slots + 15) & ~(size_t)15) + 16;          // tables + entry lists + the CTA's 128 rows
        if (p.f16_n > 0 && p.f16_const && p.opt_frame16t && mode == 0 && variants && batch >= 2048 && !host_mapped && f16v_smem <= 48 * 1024) {
            Frame16vArgs f{};
            f.nseg = p.nseg; f.n_slots = p.n_slots; f.batch = batch; f.kappa = reinterpret_cast<const double2*>(d->h_kap16);
            f.sweight = d->h_sw16; f.cstat = d->h_cstat; f.group = group; f.variants = variants; f.slots = d->slots;
            f.fac_row = d->h_facrow; f.fac = d->h_fac; f.out = out; f.worklist = a.worklist; f.work_count = a.work_count;
            for (int i = 0; i < 16; ++i) f.init[i] = p.f16_init[i];
            for (int k = 0; k < p.nseg; ++k) f.mask[k] = p.f16_mask[k];
            const long long grid = (batch + 127) / 128;
            if (grid >= (1LL << 31)) return done(fail(DMX_ERR_UNSUPPORTED, "batch too large for one launch"));
            dmx_frame16v_kernel<<<(unsigned)grid, 128, f16v_smem, st>>>(f);
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
            return done(run_tile_single(p, d, batch, params, variants, out, OUT_ENERGY, wl + 1, wl, st, group));
        }
        // Thread-per-circuit form (dmx_frame16t_kernel) for batches that give every SM a few warps of circuits; single
        // evaluations and small batches keep the warp-per-circuit kernels (one circuit's latency spread over 32 lanes).
        if (p.f16_n > 0 && p.opt_frame16t && mode == 1 && !variants && !g.any_const && batch >= 2048 && (!host_mapped || p.opt_frame16t == 2)) {
            Frame16tArgs f{};
            const bool scaled = p.f16_scaled && p.opt_f16_scaled;
            f.nseg = p.nseg; f.n_params = p.n_params; f.batch = batch; f.group = group;
            f.tab = scaled ? (const void*)d->h_kap16 : (const void*)d->h_tab16; f.eweight = scaled ? d->h_sw16 : d->h_ew16;
            f.params = params; f.out = out; f.cls = p.f_classes[0];
            for (int i = 0; i < 16; ++i) f.init[i] = p.f16_init[i];
            for (int k = 0; k < p.nseg; ++k) f.mask[k] = p.f16_mask[k];
            const int threads = p.opt_f16_threads;
            const long long grid = (batch + threads - 1) / threads;
            if (grid >= (1LL << 31)) return done(fail(DMX_ERR_UNSUPPORTED, "batch too large for one launch"));
            const size_t smem = (size_t)p.nseg * (scaled ? 128 : 512) + 256;
            if (scaled) dmx_frame16t_kernel<true><<<(unsigned)grid, threads, smem, st>>>(f);
            else dmx_frame16t_kernel<false><<<(unsigned)grid, threads, smem, st>>>(f);
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
            return done(DMX_OK);
        }
        int cw = 0, cpb = 8;
        const int f32s_opt = p.opt_frame32s;
        // (mapped host buffers stay on the 8-circuit kernel: its 1024 CTAs each moving a 512-byte block keep more PCIe reads in
        // flight than 512 CTAs moving 1 KB - 4.38e9 against 3.96e9 circuits/s end to end)
        if (f32s_opt == 1 && batch >= (1 << 15) && !host_mapped) {
            // fewest circuits per warp (most parallelism) whose warps are all resident at once
            if (!d->f32s_warps_per_sm) {
                int ctas = 0;
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dmx_frame32s_kernel<8>, 128, (size_t)p.nseg * 1024 + 1024));
                d->f32s_warps_per_sm = std::max(4, 4 * ctas);
            }
            const long long resident = (long long)d->sm_count * d->f32s_warps_per_sm;
            cw = 32;
            for (int c = 8; c < 32; c += 8)
                if ((batch + c - 1) / c <= resident) { cw = c; break; }
        }
        else if (f32s_opt == 8 || f32s_opt == 16) { cw = 32; cpb = f32s_opt; }
        else if (f32s_opt > 32) cw = f32s_opt - 32;
        if (cw && mode == 1 && !variants && !g.any_const && p.nseg <= 44) {
            // parameter sweep with one rotation class (dmx_frame32s_kernel)
            Frame32sArgs f{};
            f.nseg = p.nseg; f.n_params = p.n_params; f.cw = cw; f.batch = batch; f.tab = d->g_tab4; f.cls = p.f_classes[0];
            f.init = d->g_init; f.eweight = d->g_ew; f.group = group; f.params = params; f.out = out;
            f.stage_io = host_mapped ? 1 : 0;
            const long long warps = (batch + cw - 1) / cw;
            const long long grid = (warps + 3) / 4;
            const size_t smem = (size_t)p.nseg * 1024 + 1024;   // tables + io block, <= 45 KB: inside the default dynamic limit
            if (grid >= (1LL << 31)) return done(fail(DMX_ERR_UNSUPPORTED, "batch too large for one launch"));
            if (cpb == 16) dmx_frame32s_kernel<16><<<(unsigned)grid, 128, smem, st>>>(f);
            else dmx_frame32s_kernel<8><<<(unsigned)grid, 128, smem, st>>>(f);
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
            return done(DMX_OK);
        }
        if (p.opt_seg_cpb == 4)
            rc = mode == 0 ? launch_frame32<4, 0>(g, d->sm_count, st) : mode == 1 ? launch_frame32<4, 1>(g, d->sm_count, st) : launch_frame32<4, 2>(g, d->sm_count, st);
        else
            rc = mode == 0 ? launch_frame32<8, 0>(g, d->sm_count, st) : mode == 1 ? launch_frame32<8, 1>(g, d->sm_count, st) : launch_frame32<8, 2>(g, d->sm_count, st);
        if (rc == DMX_OK && variants) rc = run_tile_single(p, d, batch, params, variants, out, OUT_ENERGY, wl + 1, wl, st, group);
        return done(rc);
    }
    if (p.opt_seg_cpb == 4) {
        rc = mode == 0 ? launch_frame<4, 0>(a, d->sm_count, st) : mode == 1 ? launch_frame<4, 1>(a, d->sm_count, st) : launch_frame<4, 2>(a, d->sm_count, st);
    } else {
        rc = mode == 0 ? launch_frame<8, 0>(a, d->sm_count, st) : mode == 1 ? launch_frame<8, 1>(a, d->sm_count, st) : launch_frame<8, 2>(a, d->sm_count, st);
    }
    if (rc == DMX_OK && variants) {
        // circuits with non-diagonal correction ops (R-type / projector basis ops) take the general tile path
        rc = run_tile_single(p, d, batch, params, variants, out, OUT_ENERGY, wl + 1, wl, st, group);
    }
    return done(rc);
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int dmx_version(void) { return DMX_VERSION; }
const char* dmx_last_error(void) { return g_last_error.c_str(); }
int64_t dmx_launch_count(void) { return g_launches.load(); }

int dmx_plan_create(int n_qubits, int n_params, int n_slots, const dmx_op* ops, int n_ops, const double* matrix_pool,
                    size_t pool_len, dmx_plan** out) {
    if (!out) return fail(DMX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_qubits < 1) return fail(DMX_ERR_INVALID, "n_qubits must be >= 1");
    if (n_qubits > DMX_MAX_QUBITS) return fail(DMX_ERR_UNSUPPORTED, "n_qubits exceeds DMX_MAX_QUBITS");
    if (n_params < 0 || n_slots < 0 || n_ops < 0 || (n_ops > 0 && !ops)) return fail(DMX_ERR_INVALID, "bad sizes");
    try {
        dmx_plan* pl = new dmx_plan();
        pl->p.n = n_qubits; pl->p.n_params = n_params; pl->p.n_slots = n_slots;
        pl->p.ops.assign(ops, ops + n_ops);
        if (pool_len) pl->p.pool.assign(matrix_pool, matrix_pool + pool_len);
        *out = pl;
    } catch (const std::bad_alloc&) { return fail(DMX_ERR_NOMEM, "out of host memory"); }
    return DMX_OK;
}

int dmx_plan_set_hamiltonian(dmx_plan* plan, int n_terms, const uint32_t* xmask, const uint32_t* zmask, const double* coeff) {
    if (!plan || n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff))) return fail(DMX_ERR_INVALID, "bad Hamiltonian arguments");
    if (plan->p.finalized) return fail(DMX_ERR_STATE, "plan already finalized");
    const uint32_t lim = plan->p.n >= 32 ? 0xffffffffu : ((1u << plan->p.n) - 1u);
    for (int t = 0; t < n_terms; ++t)
        if ((xmask[t] | zmask[t]) & ~lim) return fail(DMX_ERR_INVALID, "Hamiltonian mask has bits beyond n_qubits");
    plan->p.hx.assign(xmask, xmask + n_terms);
    plan->p.hz.assign(zmask, zmask + n_terms);
    plan->p.hc.assign(coeff, coeff + n_terms);
    return DMX_OK;
}

int dmx_plan_set_hamiltonian_groups(dmx_plan* plan, int n_groups, const double* coeff) {
    if (!plan || n_groups < 1 || !coeff) return fail(DMX_ERR_INVALID, "bad coefficient-group arguments");
    if (plan->p.finalized) return fail(DMX_ERR_STATE, "plan already finalized");
    if (plan->p.hc.empty()) return fail(DMX_ERR_STATE, "set the Hamiltonian's Pauli strings first (dmx_plan_set_hamiltonian)");
    if (n_groups > (1 << 20)) return fail(DMX_ERR_UNSUPPORTED, "too many coefficient groups");
    plan->p.n_groups = n_groups;
    plan->p.hgroups.assign(coeff, coeff + (size_t)n_groups * plan->p.hc.size());
    return DMX_OK;
}

int dmx_plan_set_option(dmx_plan* plan, const char* key, int value) {
    if (!plan || !key) return fail(DMX_ERR_INVALID, "NULL argument");
    if (plan->p.finalized) return fail(DMX_ERR_STATE, "plan already finalized");
    std::string k(key);
    if (k == "fuse") plan->p.opt_fuse = value != 0;
    else if (k == "segments") plan->p.opt_segments = value != 0;
    else if (k == "tile_qubits") {
        if (value < 4 || value > 7) return fail(DMX_ERR_INVALID, "tile_qubits must be in 4..7");
        plan->p.opt_tile_qubits = value;
    } else if (k == "remap") {
        plan->p.opt_remap = value != 0;
    } else if (k == "triples") {
        plan->p.opt_triples = value != 0;
    } else if (k == "stream_kernel") {
        plan->p.opt_stream_kernel = value != 0;
    } else if (k == "stream_chunk") {
        if (value < 0) return fail(DMX_ERR_INVALID, "stream_chunk must be >= 0");
        plan->p.opt_stream_chunk = value;
    } else if (k == "host_zero_copy") {
        plan->p.opt_host_zero_copy = value != 0;
    } else if (k == "host_streams") {
        if (value < 1 || value > DevPlan::MAX_XS) return fail(DMX_ERR_INVALID, "host_streams must be in 1..8");
        plan->p.opt_host_streams = value;
    } else if (k == "host_chunk_rows") {
        if (value < 1024) return fail(DMX_ERR_INVALID, "host_chunk_rows must be >= 1024");
        plan->p.opt_host_chunk_rows = value;
    } else if (k == "frame32") {
        plan->p.opt_frame32 = value != 0;
    } else if (k == "segment_cpb") {
        if (value != 4 && value != 8) return fail(DMX_ERR_INVALID, "segment_cpb must be 4 or 8");
        plan->p.opt_seg_cpb = value;
    } else if (k == "bulk") {
        if (value < 0 || value > 2) return fail(DMX_ERR_INVALID, "bulk must be 0 (register staging), 1 (cp.async.bulk) or 2 (direct global loads)");
        plan->p.opt_bulk = value;
    } else if (k == "frame32s") {
        if (value != 0 && value != 1 && value != 8 && value != 16 && value != 40 && value != 48 && value != 56 && value != 64)
            return fail(DMX_ERR_INVALID, "frame32s must be 0, 1 (automatic), 8, 16 or 32 + circuits per warp (40, 48, 56, 64)");
        plan->p.opt_frame32s = value;
    } else if (k == "wide") {
        plan->p.opt_wide = value != 0;
    } else if (k == "scale_smem") {
        plan->p.opt_scale_smem = value != 0;
    } else if (k == "direct_store") {
        plan->p.opt_direct_store = value != 0;
    } else if (k == "frame16t") {
        if (value < 0 || value > 2) return fail(DMX_ERR_INVALID, "frame16t must be 0 (never), 1 (device-resident batches) or 2 (also mapped host buffers)");
        plan->p.opt_frame16t = value;
    } else if (k == "frame16t_scaled") {
        plan->p.opt_f16_scaled = value != 0;
    } else if (k == "frame16t_threads") {
        if (value != 64 && value != 128 && value != 256) return fail(DMX_ERR_INVALID, "frame16t_threads must be 64, 128 or 256");
        plan->p.opt_f16_threads = value;
    } else if (k == "tile_threads") {
        if (value != 512 && value != 1024) return fail(DMX_ERR_INVALID, "tile_threads must be 512 or 1024");
        plan->p.opt_tile_threads = value;
    } else return fail(DMX_ERR_INVALID, "unknown option: " + k);
    return DMX_OK;
}

int dmx_plan_finalize(dmx_plan* plan) {
    if (!plan) return fail(DMX_ERR_INVALID, "plan is NULL");
    if (plan->p.finalized) return fail(DMX_ERR_STATE, "plan already finalized");
    NvtxRange nvtx("dmx:plan_finalize");
    try {
        lower_plan(plan->p);
    } catch (const std::bad_alloc&) {
        return fail(DMX_ERR_NOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(DMX_ERR_INVALID, e.what());
    }
    return DMX_OK;
}

int dmx_plan_get_info(const dmx_plan* plan, dmx_plan_info* info) {
    if (!plan || !info) return fail(DMX_ERR_INVALID, "NULL argument");
    if (!plan->p.finalized) return fail(DMX_ERR_STATE, "plan not finalized");
    *info = plan->p.info;
    return DMX_OK;
}

int64_t dmx_plan_dump(const dmx_plan* plan, char* buf, int64_t buf_len) {
    if (!plan || !plan->p.finalized) { fail(DMX_ERR_STATE, "plan not finalized"); return -1; }
    std::string s = dump_plan(plan->p);
    if (buf && buf_len > 0) {
        size_t ncopy = std::min<size_t>(s.size(), (size_t)buf_len - 1);
        std::memcpy(buf, s.data(), ncopy);
        buf[ncopy] = 0;
    }
    return (int64_t)s.size() + 1;
}

void dmx_plan_destroy(dmx_plan* plan) {
    if (!plan) return;
    free_device(plan->p);
    delete plan;
}

size_t dmx_workspace_bytes(const dmx_plan* plan, int64_t batch) {
    if (!plan || !plan->p.finalized || batch <= 0) return 0;
    if (plan->p.path != 2) return 0;
    // at least one state; more lets several circuits share a launch (caller may pass any multiple)
    return (size_t)plan->p.info.workspace_bytes_per_state * (size_t)std::min<int64_t>(batch, 64);
}

static int check_exec(dmx_plan* plan, int64_t batch, const double* d_params, const void* out) {
    if (!plan || !out) return fail(DMX_ERR_INVALID, "NULL argument");
    if (!plan->p.finalized) return fail(DMX_ERR_STATE, "plan not finalized");
    if (batch < 0) return fail(DMX_ERR_INVALID, "negative batch");
    if (plan->p.n_params > 0 && !plan->p.rots.empty() && !d_params && batch > 0) return fail(DMX_ERR_INVALID, "d_params is NULL but the circuit has parameters");
    return DMX_OK;
}

static int energy_impl(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants, double* d_energy,
                       void* d_workspace, size_t workspace_bytes, cudaStream_t st, const int* d_group = nullptr, bool host_mapped = false) {
    NvtxRange nvtx("dmx:energy_batch");
    Plan& p = plan->p;
    if (p.term_idx.empty()) return fail(DMX_ERR_STATE, "no Hamiltonian set");
    DevPlan* d;
    int rc = ensure_device(p, &d);
    if (rc) return rc;
    if (batch == 0) return DMX_OK;
    if (p.path == 1) return run_segment(p, d, batch, d_params, d_variants, d_energy, st, d_group, host_mapped);
    if (p.path == 0) return run_tile_single(p, d, batch, d_params, d_variants, d_energy, OUT_ENERGY, nullptr, nullptr, st, d_group);
    const size_t sb = (size_t)p.info.state_bytes;
    const long long cap = (long long)(workspace_bytes / (size_t)p.info.workspace_bytes_per_state);
    if (!d_workspace || cap < 1) return fail(DMX_ERR_INVALID, "workspace too small: need dmx_workspace_bytes()");
    double* const state2 = p.remapped ? (double*)((char*)d_workspace + (size_t)cap * sb) : nullptr;
    // circuits per pass launch: large enough to fill the GPU for several waves ("stream_chunk" overrides)
    long long chunk = std::min<long long>(cap, p.opt_stream_chunk > 0 ? p.opt_stream_chunk : 64);
    double* chain_scratch = nullptr;
    if (!p.chains.empty()) CUDA_TRY(cudaMallocAsync((void**)&chain_scratch, (size_t)chunk * p.chains.size() * 128, st));
    struct Release { double* ptr; cudaStream_t st; ~Release() { if (ptr) cudaFreeAsync(ptr, st); } } release{chain_scratch, st};
    for (long long off = 0; off < batch; off += chunk) {
        long long cur = std::min<long long>(chunk, batch - off);
        rc = run_streaming(p, d, cur, off, d_params, d_variants, (double*)d_workspace, state2, st, chain_scratch);
        if (rc) return rc;
        int threads = 128;
        long long warps = cur;
        int grid = (int)((warps * 32 + threads - 1) / threads);
        dmx_energy_kernel<<<grid, threads, 0, st>>>((const double*)d_workspace, p.n, cur, d->term_idx, d->term_coef, (int)p.term_idx.size(), d_group, d_energy, off);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    return DMX_OK;
}

int dmx_energy_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants, double* d_energy,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
    int rc = check_exec(plan, batch, d_params, d_energy);
    if (rc) return rc;
    return energy_impl(plan, batch, d_params, d_variants, d_energy, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int dmx_energy_batch_grouped(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants, const int32_t* d_group,
                             double* d_energy, void* d_workspace, size_t workspace_bytes, void* stream) {
    int rc = check_exec(plan, batch, d_params, d_energy);
    if (rc) return rc;
    if (d_group && plan->p.n_groups < 1) return fail(DMX_ERR_STATE, "plan has no coefficient groups (dmx_plan_set_hamiltonian_groups)");
    return energy_impl(plan, batch, d_params, d_variants, d_energy, d_workspace, workspace_bytes, (cudaStream_t)stream, d_group);
}

int dmx_pauli_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants, double* d_pauli,
                    void* d_workspace, size_t workspace_bytes, void* stream) {
    (void)d_workspace; (void)workspace_bytes;
    int rc = check_exec(plan, batch, d_params, d_pauli);
    if (rc) return rc;
    Plan& p = plan->p;
    DevPlan* d;
    rc = ensure_device(p, &d);
    if (rc) return rc;
    if (batch == 0) return DMX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.n <= 7) return run_tile_single(p, d, batch, d_params, d_variants, d_pauli, OUT_PAULI, nullptr, nullptr, st);
    // streaming: the output buffer is the state buffer (remapped plans need a scratch twin for the ping-pong)
    double* twin = nullptr;
    if (p.remapped) CUDA_TRY(cudaMallocAsync((void**)&twin, (size_t)batch * (size_t)p.info.state_bytes, st));
    rc = run_streaming(p, d, batch, 0, d_params, d_variants, d_pauli, twin, st);
    if (twin) cudaFreeAsync(twin, st);
    return rc;
}

int dmx_density_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants, double* d_rho,
                      void* d_workspace, size_t workspace_bytes, void* stream) {
    int rc = check_exec(plan, batch, d_params, d_rho);
    if (rc) return rc;
    Plan& p = plan->p;
    const size_t need = (size_t)p.info.state_bytes * (size_t)batch;
    if (batch > 0 && (!d_workspace || workspace_bytes < need)) return fail(DMX_ERR_INVALID, "density output needs batch * 8 * 4^n bytes of workspace");
    rc = dmx_pauli_batch(plan, batch, d_params, d_variants, (double*)d_workspace, nullptr, 0, stream);
    if (rc || batch == 0) return rc;
    long long total = (long long)batch << (2 * p.n);
    int threads = 256;
    long long grid = (total + threads - 1) / threads;
    dmx_pauli_to_rho<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>((const double*)d_workspace, p.n, batch, (double2*)d_rho);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

int dmx_diag_batch(dmx_plan* plan, int64_t batch, const double* d_params, const uint8_t* d_variants, double* d_diag, int renormalise,
                   void* d_workspace, size_t workspace_bytes, void* stream) {
    int rc = check_exec(plan, batch, d_params, d_diag);
    if (rc) return rc;
    Plan& p = plan->p;
    const size_t need = (size_t)p.info.state_bytes * (size_t)batch;
    if (batch > 0 && (!d_workspace || workspace_bytes < need)) return fail(DMX_ERR_INVALID, "diag output needs batch * 8 * 4^n bytes of workspace");
    rc = dmx_pauli_batch(plan, batch, d_params, d_variants, (double*)d_workspace, nullptr, 0, stream);
    if (rc || batch == 0) return rc;
    dmx_pauli_to_diag<<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>((const double*)d_workspace, p.n, batch, d_diag, renormalise);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

int dmx_qp_reduce(const double* d_values, const double* d_weights, const int32_t* d_counts, int64_t batch, double* d_out3, void* stream) {
    if (!d_values || !d_weights || !d_out3 || batch < 0) return fail(DMX_ERR_INVALID, "bad arguments");
    NvtxRange nvtx("dmx:qp_reduce");
    cudaStream_t st = (cudaStream_t)stream;
    double* partial = nullptr;       // per call, stream-ordered: concurrent reductions on different streams do not share it
    CUDA_TRY(cudaMallocAsync((void**)&partial, QP_BLOCKS * 3 * sizeof(double), st));
    dmx_qp_reduce_stage1<<<QP_BLOCKS, QP_THREADS, 0, st>>>(d_values, d_weights, d_counts, batch, partial);
    dmx_qp_reduce_stage2<<<1, 32, 0, st>>>(partial, d_out3);
    CUDA_TRY(cudaFreeAsync(partial, st));
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

int dmx_qp_sampler_create(int n_slots, const int32_t* n_choices, const double* qp, int stride, dmx_qp_sampler** out) {
    if (n_slots < 1 || !n_choices || !qp || stride < 1 || !out) return fail(DMX_ERR_INVALID, "bad arguments");
    std::vector<double> cdf((size_t)n_slots * stride, 1.0);
    std::vector<unsigned char> sg((size_t)n_slots * stride, 0);
    for (int s = 0; s < n_slots; ++s) {
        const int K = n_choices[s];
        if (K < 1 || K > stride || K > 256) return fail(DMX_ERR_INVALID, "n_choices must be in 1..min(stride, 256)");
        const double* row = qp + (size_t)s * stride;
        double total = 0.0;
        for (int i = 0; i < K; ++i) total += std::fabs(row[i]);
        if (!(total > 0.0)) return fail(DMX_ERR_INVALID, "quasi-probability row sums to zero");
        double acc = 0.0;
        for (int i = 0; i < K; ++i) {
            acc += std::fabs(row[i]) / total;
            cdf[(size_t)s * stride + i] = i == K - 1 ? 1.0 : acc;
            sg[(size_t)s * stride + i] = row[i] == 0.0 ? 2 : (row[i] < 0.0 ? 1 : 0);
        }
    }
    // idx is drawable iff some 32-bit word w has cdf[idx - 1] <= (w + 0.5) / 2^32 < cdf[idx]: the number of words below a
    // cumulative value is ceil(cdf * 2^32 - 0.5) clamped to [0, 2^32] (the identity-threshold formula below)
    std::vector<uint32_t> drawable((size_t)n_slots * 8, 0u);
    for (int s = 0; s < n_slots; ++s) {
        unsigned long long below = 0ull;
        for (int i = 0; i < n_choices[s]; ++i) {
            const double x = cdf[(size_t)s * stride + i] * 4294967296.0 - 0.5;
            const unsigned long long t = x <= 0.0 ? 0ull : (x >= 4294967296.0 ? 1ull << 32 : (unsigned long long)std::ceil(x));
            if (t > below) drawable[(size_t)s * 8 + (i >> 5)] |= 1u << (i & 31);
            below = std::max(below, t);
        }
    }
    const int groups = (n_slots + 3) / 4;
    // identity thresholds: cdf[0] > (word + 0.5) / 2^32  <=>  word < ceil(cdf[0] * 2^32 - 0.5); both products are exact in double
    std::vector<unsigned long long> first((size_t)groups * 4, 1ull << 32);
    std::vector<unsigned char> sg0((size_t)groups * 4, 0);
    for (int s = 0; s < n_slots; ++s) {
        const double x = cdf[(size_t)s * stride] * 4294967296.0 - 0.5;
        first[s] = x <= 0.0 ? 0ull : (x >= 4294967296.0 ? 1ull << 32 : (unsigned long long)std::ceil(x));
        sg0[s] = sg[(size_t)s * stride];
    }
    // fast identity test of a whole lane-group: word <= threshold - 1 on 32 bits, plus the sign summary of the four identity entries
    std::vector<unsigned> fast((size_t)groups * 5, 0u);           // [groups] uint4 thresholds, then [groups] summary words
    for (int g = 0; g < groups; ++g) {
        unsigned summary = 0u;
        for (int i = 0; i < 4; ++i) {
            const unsigned long long t = first[(size_t)g * 4 + i];
            fast[(size_t)g * 4 + i] = t ? (unsigned)(t - 1) : 0u;
            if (!t) summary |= 4u;
            if (g * 4 + i < n_slots) {
                if (sg0[(size_t)g * 4 + i] == 1) summary ^= 1u;
                if (sg0[(size_t)g * 4 + i] == 2) summary |= 2u;
            }
        }
        fast[(size_t)groups * 4 + g] = summary;
    }
    auto* sm = new dmx_qp_sampler();
    sm->n_slots = n_slots; sm->stride = stride; sm->drawable = drawable;
    // every sub-table starts 16-byte aligned (thr0 is read as ulonglong2)
    sm->cb = (cdf.size() * 8 + 15) & ~(size_t)15; sm->fb = first.size() * 8; sm->nb = ((size_t)n_slots * 4 + 15) & ~(size_t)15; sm->gb = (sg0.size() + 15) & ~(size_t)15; sm->qb = fast.size() * 4; sm->sb = ((sg.size() + 15) & ~(size_t)15);
    if (cudaGetDevice(&sm->device) != cudaSuccess || cudaDeviceGetAttribute(&sm->sm_count, cudaDevAttrMultiProcessorCount, sm->device) != cudaSuccess ||
        sm->sm_count < 1 || cudaMalloc((void**)&sm->d, sm->cb + sm->fb + sm->nb + sm->gb + sm->sb + sm->qb) != cudaSuccess ||
        cudaMemcpy(sm->d, cdf.data(), cdf.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sm->d + sm->cb, first.data(), sm->fb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sm->d + sm->cb + sm->fb, n_choices, (size_t)n_slots * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sm->d + sm->cb + sm->fb + sm->nb, sg0.data(), sg0.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sm->d + sm->cb + sm->fb + sm->nb + sm->gb, sg.data(), sg.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sm->d + sm->cb + sm->fb + sm->nb + sm->gb + sm->sb, fast.data(), sm->qb, cudaMemcpyHostToDevice) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        if (sm->d) cudaFree(sm->d);
        delete sm;
        return fail(DMX_ERR_CUDA, std::string("dmx_qp_sampler_create: ") + cudaGetErrorString(e));
    }
    *out = sm;
    return DMX_OK;
}

void dmx_qp_sampler_destroy(dmx_qp_sampler* sampler) {
    if (!sampler) return;
    if (sampler->d) cudaFree(sampler->d);
    delete sampler;
}

int dmx_qp_sampler_draw(const dmx_qp_sampler* sm, uint64_t seed, uint64_t first_sample, int64_t batch, uint8_t* d_variants,
                        double* d_signs, void* stream) {
    if (!sm || batch < 0 || (batch > 0 && (!d_variants || !d_signs))) return fail(DMX_ERR_INVALID, "bad arguments");
    if (batch == 0) return DMX_OK;
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != sm->device) return fail(DMX_ERR_STATE, "sampler tables live on another device");
    const int threads = 256;
    const long long grid = std::min<long long>((batch * 32 + threads - 1) / threads, (long long)sm->sm_count * 8);      // persistent: 64 warps per SM
    const char* base = sm->d;
    dmx_qp_sample_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(
        sm->n_slots, (const int*)(base + sm->cb + sm->fb), (const double*)base, (const unsigned char*)(base + sm->cb + sm->fb + sm->nb + sm->gb),
        sm->stride, (const ulonglong2*)(base + sm->cb), (const unsigned*)(base + sm->cb + sm->fb + sm->nb),
        (const uint4*)(base + sm->cb + sm->fb + sm->nb + sm->gb + sm->sb),
        (const unsigned*)(base + sm->cb + sm->fb + sm->nb + sm->gb + sm->sb + (size_t)((sm->n_slots + 3) / 4) * 16), seed, first_sample, batch, d_variants,
        d_signs);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DMX_OK;
}

int dmx_qp_sampled_energies(dmx_plan* plan, const dmx_qp_sampler* sm, uint64_t seed, uint64_t first_sample, int64_t batch,
                            double* d_energy, double* d_signs, void* stream) {
    int rc = check_exec(plan, batch, nullptr, d_energy);
    if (rc) return rc;
    if (!sm || (batch > 0 && !d_signs)) return fail(DMX_ERR_INVALID, "NULL argument");
    Plan& p = plan->p;
    if (p.term_idx.empty()) return fail(DMX_ERR_STATE, "no Hamiltonian set");
    if (sm->n_slots != p.n_slots) return fail(DMX_ERR_INVALID, "sampler and plan have different numbers of variant slots");
    if (p.path == 2) return fail(DMX_ERR_UNSUPPORTED, "streaming plans take dmx_qp_sampler_draw + dmx_energy_batch (caller-owned workspace)");
    DevPlan* d;
    rc = ensure_device(p, &d);
    if (rc) return rc;
    if (sm->device != d->device) return fail(DMX_ERR_STATE, "sampler tables live on another device");
    if (batch == 0) return DMX_OK;
    NvtxRange nvtx("dmx:qp_sampled_energies");
    cudaStream_t st = (cudaStream_t)stream;
    const int ng = (p.n_slots + 3) / 4;
    const char* base = sm->d;
    const int* n_choices = (const int*)(base + sm->cb + sm->fb);
    const double* cdf = (const double*)base;
    const unsigned char* sgn = (const unsigned char*)(base + sm->cb + sm->fb + sm->nb + sm->gb);
    const ulonglong2* thr0 = (const ulonglong2*)(base + sm->cb);
    const unsigned* sgn0 = (const unsigned*)(base + sm->cb + sm->fb + sm->nb);
    const uint4* thr32 = (const uint4*)(base + sm->cb + sm->fb + sm->nb + sm->gb + sm->sb);
    const unsigned* summ = (const unsigned*)(base + sm->cb + sm->fb + sm->nb + sm->gb + sm->sb + (size_t)ng * 16);
    const size_t smem = (size_t)p.nseg * 128 + 26 * 8 + 128 * F16V_MAXACT * 8 + (size_t)((p.n_slots + 1) & ~1) * 8 + (size_t)ng * 20 + 16;
    const bool fused = p.path == 1 && p.f16_n > 0 && p.f16_const && p.opt_frame16t && p.f_classes.empty() && smem <= 48 * 1024 && batch < (1LL << 31);
    // scratch, stream-ordered: [count | sample list] and the identifier rows (all of them on the two-step path, the listed ones otherwise)
    int* wl = nullptr;
    uint8_t* rows = nullptr;
    struct Release { void* a; void* b; cudaStream_t st; ~Release() { if (a) cudaFreeAsync(a, st); if (b) cudaFreeAsync(b, st); } } release{nullptr, nullptr, st};
    if (!fused) {
        CUDA_TRY(cudaMallocAsync((void**)&rows, (size_t)batch * p.n_slots, st));
        release.a = rows;
        rc = dmx_qp_sampler_draw(sm, seed, first_sample, batch, rows, d_signs, stream);
        if (rc) return rc;
        return energy_impl(plan, batch, nullptr, rows, d_energy, nullptr, 0, st);
    }
    // can a draw of this sampler be a non-diagonal basis operation of this plan (R-type / projector corrections)?
    bool may_defer = p.n_slots > F16Q_SPILL;
    for (int sl = 0; sl < p.n_slots && !may_defer; ++sl) {
        if (p.slots[sl].seg < 0) continue;
        for (int w = 0; w < 8; ++w) {
            const uint32_t nonid = sm->drawable[(size_t)sl * 8 + w] & (w == 0 ? ~1u : ~0u);
            if (nonid & ~p.slots[sl].diag_ok[w]) { may_defer = true; break; }
        }
    }
    if (may_defer) {
        CUDA_TRY(cudaMallocAsync((void**)&rows, (size_t)batch * p.n_slots, st));
        release.a = rows;
        CUDA_TRY(cudaMallocAsync((void**)&wl, (size_t)(batch + 1) * sizeof(int), st));
        release.b = wl;
        CUDA_TRY(cudaMemsetAsync(wl, 0, sizeof(int), st));
    }
    Frame16qArgs q{};
    Frame16vArgs& f = q.v;
    f.nseg = p.nseg; f.n_slots = p.n_slots; f.batch = batch; f.kappa = reinterpret_cast<const double2*>(d->h_kap16);
    f.sweight = d->h_sw16; f.cstat = d->h_cstat; f.group = nullptr; f.variants = nullptr; f.slots = d->slots;
    f.fac_row = d->h_facrow; f.fac = d->h_fac; f.out = d_energy; f.worklist = wl ? wl + 1 : nullptr; f.work_count = wl;
    for (int i = 0; i < 16; ++i) f.init[i] = p.f16_init[i];
    for (int k = 0; k < p.nseg; ++k) f.mask[k] = p.f16_mask[k];
    q.n_choices = n_choices; q.cdf = cdf; q.sgn = sgn; q.stride = sm->stride; q.thr0 = thr0; q.sgn0 = sgn0; q.thr32 = thr32; q.summ = summ;
    q.seed = seed; q.first_sample = first_sample; q.signs = d_signs;
    if (may_defer) dmx_frame16q_kernel<true><<<(unsigned)((batch + 127) / 128), 128, smem, st>>>(q);
    else dmx_frame16q_kernel<false><<<(unsigned)((batch + 127) / 128), 128, smem, st>>>(q);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    if (!may_defer) return DMX_OK;                    // one launch: nothing can need the general path
    // listed samples (a non-diagonal basis operation, or more entries than a thread keeps): drawn again, row i = list entry i,
    // and evaluated by the general tile path
    dmx_qp_sample_kernel<<<(unsigned)std::max(1, sm->sm_count), 256, 0, st>>>(p.n_slots, n_choices, cdf, sgn, sm->stride, thr0, sgn0, thr32, summ,
                                                                             seed, first_sample, batch, rows, nullptr, wl + 1, wl);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return run_tile_single(p, d, batch, nullptr, rows, d_energy, OUT_ENERGY, wl + 1, wl, st, nullptr, true);
}

int dmx_qp_sample_variants(int n_slots, const int32_t* n_choices, const double* qp, int stride, uint64_t seed, uint64_t first_sample,
                           int64_t batch, uint8_t* d_variants, double* d_signs, void* stream) {
    if (batch < 0 || !d_variants || !d_signs) return fail(DMX_ERR_INVALID, "bad arguments");
    dmx_qp_sampler* sm = nullptr;
    int rc = dmx_qp_sampler_create(n_slots, n_choices, qp, stride, &sm);
    if (rc) return rc;
    rc = dmx_qp_sampler_draw(sm, seed, first_sample, batch, d_variants, d_signs, stream);
    if (rc == DMX_OK && batch > 0 && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = fail(DMX_ERR_CUDA, "sampler launch failed");
    dmx_qp_sampler_destroy(sm);
    return rc;
}

int dmx_shot_expectations(const double* d_probs, int64_t rows, int dim, int64_t shots, const int64_t* d_row_shots, const uint32_t* h_masks,
                          int n_masks, uint64_t seed, uint64_t first_row, double* d_expect, void* stream) {
    if (!d_probs || !d_expect || !h_masks || rows < 0 || (!d_row_shots && shots < 1) || n_masks < 1 || n_masks > 4096)
        return fail(DMX_ERR_INVALID, "bad arguments");
    if (dim < 1 || dim > SHOT_MAX_DIM || (dim & (dim - 1))) return fail(DMX_ERR_UNSUPPORTED, "dim must be a power of two up to 128");
    if (rows == 0) return DMX_OK;
    NvtxRange nvtx("dmx:shot_expectations");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* d_masks = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d_masks, (size_t)n_masks * 4, st));
    CUDA_TRY(cudaMemcpyAsync(d_masks, h_masks, (size_t)n_masks * 4, cudaMemcpyHostToDevice, st));
    const long long grid = (rows + 3) / 4;
    if (grid >= (1LL << 31)) { cudaFreeAsync(d_masks, st); return fail(DMX_ERR_UNSUPPORTED, "too many rows for one launch"); }
    dmx_shot_expect_kernel<<<(unsigned)grid, 128, 0, st>>>(d_probs, rows, dim, shots, (const long long*)d_row_shots, d_masks, n_masks, seed, first_row, d_expect);
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    CUDA_TRY(cudaFreeAsync(d_masks, st));
    CUDA_TRY(e);
    return DMX_OK;
}

static bool is_pinned_host(const void* ptr) {
    if (!ptr) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// device-visible alias of a page-locked host buffer (nullptr when the buffer is pageable or not mapped)
static void* mapped_alias(const void* ptr) {
    if (!ptr) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}


// Host-buffer call.  The batch is cut into chunks that flow through three streams — H2D of chunk i+1, the kernels of chunk i
// and the D2H of chunk i-1 overlap — so the call costs about max(copy in, compute, copy out) instead of their sum.  Buffers
// that are already page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors) are used in place; pageable
// ones go through the plan's pinned staging buffer.
static int energy_batch_host_impl(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants, double* h_energy,
                                  bool wait, const int32_t* h_group = nullptr);

int dmx_energy_batch_host_grouped(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants,
                                  const int32_t* h_group, double* h_energy) {
    if (h_group && plan && plan->p.n_groups < 1) return fail(DMX_ERR_STATE, "plan has no coefficient groups (dmx_plan_set_hamiltonian_groups)");
    return energy_batch_host_impl(plan, batch, h_params, h_variants, h_energy, true, h_group);
}

int dmx_energy_batch_host(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants, double* h_energy) {
    return energy_batch_host_impl(plan, batch, h_params, h_variants, h_energy, true);
}

int dmx_energy_batch_host_async(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants, double* h_energy) {
    return energy_batch_host_impl(plan, batch, h_params, h_variants, h_energy, false);
}

int dmx_plan_host_sync(dmx_plan* plan) {
    if (!plan) return fail(DMX_ERR_INVALID, "plan is NULL");
    DevPlan* d = (DevPlan*)plan->p.dev;
    if (!d || !d->stream) return DMX_OK;       // nothing was ever queued
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    for (int i = 0; i < DevPlan::MAX_XS; ++i) if (d->xs[i]) CUDA_TRY(cudaStreamSynchronize(d->xs[i]));
    return DMX_OK;
}

static int energy_batch_host_impl(dmx_plan* plan, int64_t batch, const double* h_params, const uint8_t* h_variants, double* h_energy,
                                  bool wait, const int32_t* h_group) {
    int rc = check_exec(plan, batch, h_params, h_energy);
    if (rc) return rc;
    Plan& p = plan->p;
    DevPlan* d;
    rc = ensure_device(p, &d);
    if (rc) return rc;
    if (batch == 0) return DMX_OK;
    NvtxRange nvtx("dmx:energy_batch_host");
    std::lock_guard<std::mutex> lock(d->mu);
    if (!d->stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < DevPlan::MAX_CHUNKS; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_k[i], cudaEventDisableTiming));
        }
    }
    // Zero-copy: a parameter sweep on a register-resident plan moves 8 B in and 8 B out per circuit, so two DMA copies
    // and their stream hand-offs cost more than the kernel.  With mapped page-locked buffers the kernel reads the parameters
    // and writes the energies over PCIe itself: one launch, one synchronise.  (Variant tables - 1 byte per gate and row,
    // read lane-strided - stay on the DMA pipeline below.)
    if (p.opt_host_zero_copy && p.path == 1 && !h_variants) {
        double* const e_dev = (double*)mapped_alias(h_energy);
        const double* const p_dev = h_params ? (const double*)mapped_alias(h_params) : nullptr;
        const int* const g_dev = h_group ? (const int*)mapped_alias(h_group) : nullptr;
        if (e_dev && (p_dev || !h_params) && (g_dev || !h_group)) {
            if (!wait) {
                // dmx_energy_batch_host_async: consecutive batches rotate over a few streams, so the 1.7-wave tail of one
                // launch overlaps the head of the next (they are independent: every call has its own result buffer); the
                // caller collects with dmx_plan_host_sync
                cudaStream_t xs = d->stream;
                if (p.opt_host_streams > 1) {
                    const int slot = (int)(d->async_calls++ % (unsigned)std::min(p.opt_host_streams, DevPlan::MAX_XS));
                    if (!d->xs[slot]) CUDA_TRY(cudaStreamCreateWithFlags(&d->xs[slot], cudaStreamNonBlocking));
                    xs = d->xs[slot];
                }
                return energy_impl(plan, batch, p_dev, nullptr, e_dev, nullptr, 0, xs, g_dev, true);
            }
            rc = energy_impl(plan, batch, p_dev, nullptr, e_dev, nullptr, 0, d->stream, g_dev, true);
            const cudaError_t e = cudaStreamSynchronize(d->stream);
            if (rc) return rc;
            CUDA_TRY(e);
            return DMX_OK;
        }
    }
    const size_t prow = h_params ? (size_t)p.n_params * 8 : 0, vrow = h_variants ? (size_t)p.n_slots : 0;
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t pb = align((size_t)batch * prow), vb = align((size_t)batch * vrow), eb = align((size_t)batch * 8);
    const size_t gb = h_group ? align((size_t)batch * 4) : 0;
    size_t ws = 0;
    if (p.path == 2) ws = (size_t)p.info.workspace_bytes_per_state * (size_t)std::min<int64_t>(batch, 64);
    const bool pin_in = is_pinned_host(h_params) && is_pinned_host(h_variants), pin_out = is_pinned_host(h_energy);
    const size_t hneed = (pin_in ? 0 : pb + vb) + (pin_out ? 0 : eb), dneed = pb + vb + eb + ws + gb;
    if (d->h_cap < hneed) {
        if (d->h_stage) CUDA_TRY(cudaFreeHost(d->h_stage));
        d->h_stage = nullptr; d->h_cap = 0;
        CUDA_TRY(cudaMallocHost(&d->h_stage, hneed));
        d->h_cap = hneed;
    }
    if (d->d_cap < dneed) {
        if (d->d_stage) CUDA_TRY(cudaFree(d->d_stage));
        d->d_stage = nullptr; d->d_cap = 0;
        CUDA_TRY(cudaMalloc(&d->d_stage, dneed));
        d->d_cap = dneed;
    }
    char* hs = (char*)d->h_stage;
    char* ds = (char*)d->d_stage;
    const char* src_p = pin_in ? (const char*)h_params : hs;
    const char* src_v = pin_in ? (const char*)h_variants : hs + pb;
    char* dst_e = pin_out ? (char*)h_energy : hs + (pin_in ? 0 : pb + vb);
    // coefficient rows: one small (pageable or pinned) copy ahead of the pipeline, ordered before every chunk on s_in
    int* const d_grp = h_group ? (int*)(ds + pb + vb + eb + ws) : nullptr;
    if (h_group) CUDA_TRY(cudaMemcpyAsync(d_grp, h_group, (size_t)batch * 4, cudaMemcpyHostToDevice, d->s_in));
    // small-state plans: pipeline in up to MAX_CHUNKS chunks; streaming plans are long-running, one chunk
    int n_chunks = 1;
    if (p.path != 2) n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(DevPlan::MAX_CHUNKS, batch / p.opt_host_chunk_rows));
    const int64_t per = ((batch + n_chunks - 1) / n_chunks + 7) & ~(int64_t)7;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t lo = (int64_t)c * per, cnt = std::min<int64_t>(per, batch - lo);
        if (cnt <= 0) { n_chunks = c; break; }
        if (!pin_in) {
            if (prow) std::memcpy(hs + lo * prow, (const char*)h_params + lo * prow, (size_t)cnt * prow);
            if (vrow) std::memcpy(hs + pb + lo * vrow, (const char*)h_variants + lo * vrow, (size_t)cnt * vrow);
        }
        if (prow) CUDA_TRY(cudaMemcpyAsync(ds + lo * prow, src_p + lo * prow, (size_t)cnt * prow, cudaMemcpyHostToDevice, d->s_in));
        if (vrow) CUDA_TRY(cudaMemcpyAsync(ds + pb + lo * vrow, src_v + lo * vrow, (size_t)cnt * vrow, cudaMemcpyHostToDevice, d->s_in));
        CUDA_TRY(cudaEventRecord(d->ev_in[c], d->s_in));
        CUDA_TRY(cudaStreamWaitEvent(d->stream, d->ev_in[c], 0));
        rc = energy_impl(plan, cnt, prow ? (const double*)(ds + lo * prow) : nullptr, vrow ? (const uint8_t*)(ds + pb + lo * vrow) : nullptr,
                         (double*)(ds + pb + vb) + lo, ws ? ds + pb + vb + eb : nullptr, ws, d->stream, d_grp ? d_grp + lo : nullptr);
        if (rc) { cudaStreamSynchronize(d->stream); cudaStreamSynchronize(d->s_in); cudaStreamSynchronize(d->s_out); return rc; }
        CUDA_TRY(cudaEventRecord(d->ev_k[c], d->stream));
        CUDA_TRY(cudaStreamWaitEvent(d->s_out, d->ev_k[c], 0));
        CUDA_TRY(cudaMemcpyAsync(dst_e + lo * 8, ds + pb + vb + lo * 8, (size_t)cnt * 8, cudaMemcpyDeviceToHost, d->s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(d->s_out));
    if (!pin_out) std::memcpy(h_energy, dst_e, (size_t)batch * 8);
    return DMX_OK;
}

int dmx_measure_fp64_peak(double* tflops_out, void* stream) {
    if (!tflops_out) return fail(DMX_ERR_INVALID, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    CUDA_TRY(cudaMalloc((void**)&out, 8));
    const int iters = 1 << 15, threads = 512, grid = sms * 4;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    dmx_fp64_fma_kernel<<<grid, threads, 0, st>>>(out, iters);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, st));
        dmx_fp64_fma_kernel<<<grid, threads, 0, st>>>(out, iters);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    g_launches += 4;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)grid;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    return DMX_OK;
}

}  // extern "C"
