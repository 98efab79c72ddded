// The hot loop: per-element K_e and its scatter into the global CSR values.
//
// Reference semantics being reproduced (file:line in InteractiveComputerGraphics/fenris @ 7181b15):
//   assemble_element_elliptic_matrix   src/assembly/local/elliptic.rs:361-439
//     per quadrature point: J = X G_ref^T (element.rs / hexahedron.rs:101-107), det J, J^{-1}
//     (Err "Singular element Jacobian encountered" when det == 0, elliptic.rs:401-404),
//     grad phi_I = J^{-T} grad_ref phi_I (elliptic.rs:415-418), scale = w |det J| (elliptic.rs:422),
//     C_IJ += scale * contraction(grad phi_I, grad phi_J)   (operators.rs:176-188, fenris-solid lib.rs:381-391)
//   LaplaceOperator::contract = a.b                          src/assembly/operators/laplace.rs:60-68
//   LinearElasticMaterial: mu[(a.b) I + b a^T] + lambda a b^T   fenris-solid/src/materials.rs:108-122
//   symmetric fill (upper -> lower)                          src/util.rs:38-50
//   scatter K_e rows into the CSR rows of the element's nodes   src/assembly/global.rs:155-178, 504-537
//
// B200 design: the tabulated reference gradients / weights / Lame data are staged once per CTA in
// shared memory; an element is owned by a group of G lanes (G = 8 for 4-node elements, 32 otherwise): the lanes gather
// connectivity + coordinates with coalesced loads, compute the Jacobians and physical gradients (one quadrature
// point per lane) into shared memory, and then each lane forms whole s x s node blocks
//     K_ab = mu [tr(S_ab) I + S_ab^T] + lambda S_ab,   S_ab = sum_q w_q |det J_q| grad phi_a (x) grad phi_b
// (one 3x3 FMA accumulation per pair instead of the reference's per-point 3x3 temporaries) and adds them to the CSR
// through a precomputed node-block map (position of node b in the block row of node a) - no column search.
// Both triangles are computed directly; K_ba^T == K_ab holds bit-for-bit because the products commute and the
// quadrature weights enter as sqrt(scale) on both factors, so the mirror step of the reference is implicit.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "fb200_internal.h"

namespace fb200 {

enum { MODE_ATOMIC = 0, MODE_COLORED = 1, MODE_DUMP = 3, MODE_COLORED_LIST = 4 /* dispatch only: coloured launch over p.elem_list */ };

struct AssembleParams {
    const double* vertices;
    const int32_t* conn;
    const int64_t* blk_off;
    const uint16_t* blockmap;
    double* values;
    const double* tab;          // device tables, layout in DeviceTables
    int nq;
    int uniform;                // all quadrature points carry the same operator parameters
    double mu, lam;             // the uniform parameters
    const int32_t* elem_list;   // optional indirection (colour lists) - generic kernel
    // Hex8 warp kernel: connectivity / map rows already arranged in processing order (row = position)
    const int32_t* conn_pos;
    const uint16_t* map_pos;
    const int32_t* elem_ids;    // element id of each position (nullptr: first_elem + position); error reports + dump
    uint64_t first_elem;
    unsigned int* ticket32;      // dynamic position counter of the Hex8 warp kernel
    // fused zero-fill (values = contributions without a memset pass)
    int zfuse;
    const int64_t* zero_off;     // per chunk of CHUNK positions: range in zero_nodes
    const int32_t* zero_nodes;   // nodes whose first contribution comes from that chunk
    const int64_t* zero_base;    // ... index of their first value
    const int32_t* zero_len;     // ... number of values in their s rows
    uint32_t* row_epoch;         // per node: epoch of the last clear
    uint32_t epoch;
    uint32_t num_chunks;
    uint32_t zero_look;          // how many chunks the clearing warps may run ahead of the ticket counter
    uint64_t count;             // elements to process
    unsigned long long* errword;
    double* dump;               // MODE_DUMP: count * (S N)^2 doubles, column-major per element
    uint64_t dump_first;        // MODE_DUMP: first element of the contiguous range
    // gather mode
    const int64_t* adj_off;
    const int32_t* adj_inc;
    uint64_t num_nodes;
    uint64_t num_owned;
    int accumulate;
    int debug;                  // measurement knobs of the Hex8 DMMA kernel (FB200_DEBUG)
    int generic_only;           // elem_list is an arbitrary subset (quadrature-table groups): generic element kernel only
    // chunk-local scatter lists (chunks.cpp); num_chunks above
    const int64_t* slot_off;
    const uint16_t* contrib;
    const int32_t* slot_node;
    const uint16_t* slot_k;
    const uint16_t* slot_cbeg;
    const uint8_t* slot_flags;
    const int64_t* slot_dst;     // first value of the slot's block / row length (precomputed: no dependent loads in the slot loop)
    const int32_t* slot_rl;
    // tile lists of the Hex8 tile kernel (tiles.cpp)
    uint32_t num_tiles;
    const uint32_t* tile_hdr;
    const int32_t* tile_nodes;
    const uint32_t* tile_flush;
    const uint8_t* tile_lnodes;
    const uint16_t* tile_emap;
    const int32_t* tile_elem;
    const uint32_t* tile_list;   // launch over a subset of the tiles (one tile colour): ticket t -> tile tile_list[t]; NULL = all tiles
    const uint32_t* tile_wait;   // owner stores: tiles whose published stores a tile's reductions wait for
    uint32_t* tile_flag;         // ... per tile: epoch of the last launch whose stores of the tile are published
    uint32_t tile_epoch;         // ... this launch's epoch; 0 = no ownership (the values were zero-filled, complete rows are stored)
    int tile_static;             // tiles round-robin over the CTAs instead of an atomic ticket
    // fused interface exchange (comm.cu): per node 0, or (block-row offset on the neighbouring rank + 1) | neighbour slot << 31
    const uint32_t* peer_row;
    double* peer_values[2];
};

// an overwriting assembly ("values = contributions", global.rs:124-131) was requested and nothing has cleared the values yet: every
// launcher calls this before its first kernel - except the Hex8 tile kernel with owner lists, which stores every row itself
static fb200_status clear_values_if_pending(fb200_ctx* ctx) {
    if (!ctx->pending_zero) return FB200_OK;
    ctx->pending_zero = false;
    if (ctx->nnz) FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_values, 0, ctx->nnz * sizeof(double), ctx->stream));
    return FB200_OK;
}

__device__ __forceinline__ void flag_error(unsigned long long* errword, uint64_t elem, int code) {
    atomicMin(errword, ((unsigned long long)elem << 8) | (unsigned long long)code);
}

// Jacobian, determinant and inverse at one quadrature point; closed forms as in nalgebra (det by first-row cofactors).
template <int NG, int D>
__device__ __forceinline__ bool jacobian_inverse(const double* __restrict__ X /* [NG][D] */, const double* __restrict__ gg /* [NG][D] */,
                                                 double (&Jinv)[D][D], double* det_out) {
    double J[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) J[i][j] = 0.0;
#pragma unroll
    for (int a = 0; a < NG; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) J[i][j] = fma(X[a * D + i], gg[a * D + j], J[i][j]);
    double det;
    if constexpr (D == 2) {
        det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        if (det == 0.0) { *det_out = 0.0; return false; }
        const double r = 1.0 / det;
        Jinv[0][0] = J[1][1] * r;
        Jinv[0][1] = -J[0][1] * r;
        Jinv[1][0] = -J[1][0] * r;
        Jinv[1][1] = J[0][0] * r;
    } else {
        const double c00 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
        const double c01 = J[1][0] * J[2][2] - J[2][0] * J[1][2];
        const double c02 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
        det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02;
        if (det == 0.0) { *det_out = 0.0; return false; }
        const double r = 1.0 / det;
        Jinv[0][0] = c00 * r;
        Jinv[0][1] = (J[0][2] * J[2][1] - J[2][2] * J[0][1]) * r;
        Jinv[0][2] = (J[0][1] * J[1][2] - J[1][1] * J[0][2]) * r;
        Jinv[1][0] = -c01 * r;
        Jinv[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * r;
        Jinv[1][2] = (J[0][2] * J[1][0] - J[1][2] * J[0][0]) * r;
        Jinv[2][0] = c02 * r;
        Jinv[2][1] = (J[0][1] * J[2][0] - J[2][1] * J[0][0]) * r;
        Jinv[2][2] = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) * r;
    }
    *det_out = det;
    return true;
}

// Geometry of one element at quadrature points q = lane, lane+G, ...: writes the (scaled) physical gradients
// s_g[(q*N + a)*D + i] and, for non-uniform parameters, the per-point scales.
template <int N, int NG, int D, int G>
__device__ __forceinline__ void element_geometry(int lane, int nq, bool uniform, const double* __restrict__ s_w, const double* __restrict__ s_mu,
                                                 const double* __restrict__ s_lam, const double* __restrict__ s_ggeo,
                                                 const double* __restrict__ s_gref, const double* __restrict__ s_X, double* __restrict__ s_sc,
                                                 double* __restrict__ s_g, unsigned long long* errword, uint64_t elem) {
    for (int q = lane; q < nq; q += G) {
        double Jinv[D][D], det;
        const bool ok = jacobian_inverse<NG, D>(s_X, s_ggeo + q * NG * D, Jinv, &det);
        if (!ok) {
            flag_error(errword, elem, FB200_ERR_SINGULAR_JACOBIAN);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) Jinv[i][j] = 0.0;
        }
        const double alpha = s_w[q] * fabs(det);
        double gscale = 1.0;
        if (uniform) {
            gscale = sqrt(alpha);
        } else {
            s_sc[q] = alpha * s_mu[q];
            s_sc[nq + q] = alpha * s_lam[q];
        }
        const double* gr = s_gref + q * N * D;
        double* go = s_g + q * N * D;
#pragma unroll 4
        for (int a = 0; a < N; ++a) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) acc = fma(Jinv[j][i], gr[a * D + j], acc);  // (J^{-T} g)_i
                go[a * D + i] = acc * gscale;
            }
        }
    }
}

// s x s block K_ab of one node pair from the staged gradients.
template <int N, int D, int OP>
__device__ __forceinline__ void node_block(int a, int b, int nq, bool uniform, double mu, double lam, const double* __restrict__ s_sc,
                                           const double* __restrict__ s_g, double (&K)[D][D]) {
    double M[D][D], L[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) { M[i][j] = 0.0; L[i][j] = 0.0; }
    if (uniform) {
        for (int q = 0; q < nq; ++q) {
            const double* ga = s_g + (q * N + a) * D;
            const double* gb = s_g + (q * N + b) * D;
            double va[D], vb[D];
#pragma unroll
            for (int i = 0; i < D; ++i) { va[i] = ga[i]; vb[i] = gb[i]; }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) M[i][j] = fma(va[i], vb[j], M[i][j]);
        }
        if (OP == FB200_LAPLACE) {
            double tr = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) tr += M[i][i];
            K[0][0] = tr;
        } else {
            double tr = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) tr += M[i][i];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) K[i][j] = mu * ((i == j ? tr : 0.0) + M[j][i]) + lam * M[i][j];
        }
    } else {
        // per-point parameters (UniformQuadratureTable::from_points_weights_and_data): M = sum a mu g g^T, L = sum a lam g g^T
        for (int q = 0; q < nq; ++q) {
            const double* ga = s_g + (q * N + a) * D;
            const double* gb = s_g + (q * N + b) * D;
            const double am = s_sc[q], al = s_sc[nq + q];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const double pr = ga[i] * gb[j];
                    M[i][j] = fma(am, pr, M[i][j]);
                    L[i][j] = fma(al, pr, L[i][j]);
                }
        }
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) tr += M[i][i];
        if (OP == FB200_LAPLACE) {
            K[0][0] = tr;
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) K[i][j] = (i == j ? tr : 0.0) + M[j][i] + L[i][j];
        }
    }
}

template <int N, int NG, int D>
__host__ __device__ constexpr int table_doubles(int nq) { return nq * (3 + NG * D + N * D); }
template <int N, int NG, int D>
__host__ __device__ constexpr int slot_doubles(int nq) { return NG * D + 2 * nq + nq * N * D; }

// ------------------------------------------------------------------------------------------------ element-parallel kernel
template <int N, int NG, int D, int OP, int MODE, int G, int THREADS>
__global__ void __launch_bounds__(THREADS) assemble_elements_kernel(const AssembleParams p) {
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    constexpr int SLOTS = THREADS / G;
    constexpr int SN = S * N;
    extern __shared__ double smem[];
    const int nq = p.nq;
    const bool uniform = p.uniform != 0;
    const int tab_len = table_doubles<N, NG, D>(nq);
    for (int i = threadIdx.x; i < tab_len; i += THREADS) smem[i] = p.tab[i];
    const double* s_w = smem;
    const double* s_mu = smem + nq;
    const double* s_lam = smem + 2 * nq;
    const double* s_ggeo = smem + 3 * nq;
    const double* s_gref = s_ggeo + nq * NG * D;
    const int sd = slot_doubles<N, NG, D>(nq);
    const int slot = threadIdx.x / G, lane = threadIdx.x % G;
    double* s_X = smem + tab_len + slot * sd;
    double* s_sc = s_X + NG * D;
    double* s_g = s_sc + 2 * nq;
    long long* s_base = reinterpret_cast<long long*>(smem + tab_len + SLOTS * sd) + slot * N;
    int* s_rowlen = reinterpret_cast<int*>(reinterpret_cast<long long*>(smem + tab_len + SLOTS * sd) + SLOTS * N) + slot * N;
    __syncthreads();

    const uint64_t per_sweep = (uint64_t)gridDim.x * SLOTS;
    const uint64_t iters = (p.count + per_sweep - 1) / per_sweep;
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t idx = (it * gridDim.x + blockIdx.x) * SLOTS + slot;
        const bool valid = idx < p.count;
        const uint64_t e = valid ? (p.elem_list ? (uint64_t)p.elem_list[idx] : idx) : 0;
        if (valid) {
            const int32_t* en = p.conn + e * N;
            if (MODE != MODE_DUMP) {  // element matrices alone need no pattern
                for (int a = lane; a < N; a += G) {
                    const int32_t node = en[a];
                    const long long b0 = p.blk_off[node], b1 = p.blk_off[node + 1];
                    s_base[a] = (long long)(S * S) * b0;
                    s_rowlen[a] = (int)(b1 - b0) * S;
                }
            }
            for (int t = lane; t < NG * D; t += G) {
                const int a = t / D, i = t - a * D;
                s_X[t] = p.vertices[(uint64_t)en[a] * D + i];
            }
        }
        __syncwarp();
        if (valid) element_geometry<N, NG, D, G>(lane, nq, uniform, s_w, s_mu, s_lam, s_ggeo, s_gref, s_X, s_sc, s_g, p.errword, e);
        __syncwarp();
        if (valid) {
            const uint16_t* map = MODE == MODE_DUMP ? nullptr : p.blockmap + e * (uint64_t)(N * N);
            for (int t = lane; t < N * N; t += G) {
                const int a = t / N, b = t - a * N;
                double K[D][D];
                node_block<N, D, OP>(a, b, nq, uniform, p.mu, p.lam, s_sc, s_g, K);
                if (MODE == MODE_DUMP) {
                    double* out = p.dump + idx * (uint64_t)(SN * SN);
#pragma unroll
                    for (int i = 0; i < S; ++i)
#pragma unroll
                        for (int j = 0; j < S; ++j) out[(S * b + j) * SN + (S * a + i)] = K[i][j];
                } else {
                    const long long idx0 = s_base[a] + (long long)S * (long long)map[t];
                    const int rl = s_rowlen[a];
#pragma unroll
                    for (int i = 0; i < S; ++i)
#pragma unroll
                        for (int j = 0; j < S; ++j) {
                            double* dst = p.values + idx0 + (long long)i * rl + j;
                            if (MODE == MODE_ATOMIC) atomicAdd(dst, K[i][j]);
                            else *dst += K[i][j];
                        }
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ Hex8 warp-per-element kernel
#include "hex8_kernel.cuh"
#include "hex8_mma_kernel.cuh"
#include "hex8_tile_kernel.cuh"
#include "hex27_mma_kernel.cuh"
#include "tet4_chunk_kernel.cuh"

// ------------------------------------------------------------------------------------------------ row-owner (gather) kernel
// One warp per node row-block.  The warp walks the node's incident elements in groups of 32/GE elements, each group of GE
// lanes recomputes that element's geometry, forms the N blocks K_{a,b} of the node's local row a, and the groups then add
// them one after the other into a shared-memory image of the node's s CSR rows; the image is finally written with
// coalesced stores (values = image, or += when accumulating).  Every CSR value is produced by exactly one warp in a
// fixed order: deterministic, no atomics, no zero-fill pass.
constexpr int kGatherMaxRowBlocks = 128;

template <int N, int NG, int D, int OP, int GE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) assemble_gather_kernel(const AssembleParams p) {
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    constexpr int GROUPS = 32 / GE;
    extern __shared__ double smem[];
    const int nq = p.nq;
    const bool uniform = p.uniform != 0;
    const int tab_len = table_doubles<N, NG, D>(nq);
    for (int i = threadIdx.x; i < tab_len; i += WARPS * 32) smem[i] = p.tab[i];
    const double* s_w = smem;
    const double* s_mu = smem + nq;
    const double* s_lam = smem + 2 * nq;
    const double* s_ggeo = smem + 3 * nq;
    const double* s_gref = s_ggeo + nq * NG * D;
    const int sd = slot_doubles<N, NG, D>(nq);
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const int group = wl / GE, lane = wl % GE;
    double* s_warp = smem + tab_len + warp * (GROUPS * sd + S * S * kGatherMaxRowBlocks);
    double* s_X = s_warp + group * sd;
    double* s_sc = s_X + NG * D;
    double* s_g = s_sc + 2 * nq;
    double* s_row = s_warp + GROUPS * sd;  // [S][S*cnt]
    __syncthreads();

    const uint64_t nwarps = (uint64_t)gridDim.x * WARPS;
    for (uint64_t node = (uint64_t)blockIdx.x * WARPS + warp; node < p.num_nodes; node += nwarps) {
        const long long b0 = p.blk_off[node], b1 = p.blk_off[node + 1];
        const int cnt = (int)(b1 - b0);
        const int rl = cnt * S;
        const int total = rl * S;
        for (int t = wl; t < total; t += 32) s_row[t] = 0.0;
        const long long a0 = p.adj_off[node], a1 = p.adj_off[node + 1];
        for (long long base = a0; base < a1; base += GROUPS) {
            const long long ai = base + group;
            int32_t inc = ai < a1 ? p.adj_inc[ai] : -1;
            uint64_t e = inc >= 0 ? (uint64_t)(inc / N) : 0;
            const int a = inc >= 0 ? inc - (int)e * N : 0;
            const bool valid = inc >= 0 && e < p.num_owned;
            __syncwarp();
            if (valid) {
                const int32_t* en = p.conn + e * N;
                for (int t = lane; t < NG * D; t += GE) {
                    const int an = t / D, i = t - an * D;
                    s_X[t] = p.vertices[(uint64_t)en[an] * D + i];
                }
            }
            __syncwarp();
            if (valid) element_geometry<N, NG, D, GE>(lane, nq, uniform, s_w, s_mu, s_lam, s_ggeo, s_gref, s_X, s_sc, s_g, p.errword, e);
            __syncwarp();
            // blocks (a, b), b = lane, lane+GE, ...  kept in registers until this group's turn to add
            constexpr int TB = (N + GE - 1) / GE;
            double K[TB][D][D];
            int pos[TB];
#pragma unroll
            for (int tb = 0; tb < TB; ++tb) {
                const int b = lane + tb * GE;
                pos[tb] = -1;
                if (valid && b < N) {
                    node_block<N, D, OP>(a, b, nq, uniform, p.mu, p.lam, s_sc, s_g, K[tb]);
                    pos[tb] = (int)p.blockmap[e * (uint64_t)(N * N) + a * N + b];
                }
            }
#pragma unroll 1
            for (int gsel = 0; gsel < GROUPS; ++gsel) {
                if (group == gsel) {
#pragma unroll
                    for (int tb = 0; tb < TB; ++tb) {
                        if (pos[tb] >= 0) {
#pragma unroll
                            for (int i = 0; i < S; ++i)
#pragma unroll
                                for (int j = 0; j < S; ++j) s_row[i * rl + S * pos[tb] + j] += K[tb][i][j];
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        double* dst = p.values + (long long)(S * S) * b0;
        if (p.accumulate) {
            for (int t = wl; t < total; t += 32) dst[t] += s_row[t];
        } else {
            for (int t = wl; t < total; t += 32) dst[t] = s_row[t];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ host side
static fb200_status validate(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q) {
    if (!ctx) return FB200_ERR_STATE;
    if (!op || !q) return fail(ctx, FB200_ERR_SHAPE, "null operator / quadrature");
    if (!ctx->has_space) return fail(ctx, FB200_ERR_STATE, "assembly needs fb200_space_upload (a finite element space)");
    if (op->kind != FB200_LAPLACE && op->kind != FB200_LINEAR_ELASTIC)
        return fail(ctx, FB200_ERR_UNSUPPORTED, "operator has no device specialisation (no CPU fallback)");
    if (q->dim != ctx->ei.d) return fail(ctx, FB200_ERR_SHAPE, "quadrature dimension != element reference dimension");
    if (q->num_points < 1 || q->num_points > 64) return fail(ctx, FB200_ERR_UNSUPPORTED, "1..64 quadrature points supported");
    if (!q->weights || !q->points) return fail(ctx, FB200_ERR_SHAPE, "null quadrature arrays");
    if (op->kind == FB200_LINEAR_ELASTIC && !q->data) return fail(ctx, FB200_ERR_SHAPE, "linear elasticity needs Lame data per point");
    return FB200_OK;
}

fb200_status upload_tables(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q) {
    const int nq = q->num_points, n = ctx->ei.n, ng = ctx->ei.ng, d = ctx->ei.d;
    const size_t len = (size_t)nq * (3 + ng * d + n * d);
    std::vector<double> h(len, 0.0);
    double* w = h.data();
    double* mu = w + nq;
    double* lam = mu + nq;
    double* ggeo = lam + nq;
    double* gref = ggeo + (size_t)nq * ng * d;
    bool uniform = true;
    for (int k = 0; k < nq; ++k) {
        w[k] = q->weights[k];
        if (op->kind == FB200_LINEAR_ELASTIC) {
            mu[k] = q->data[2 * k];
            lam[k] = q->data[2 * k + 1];
            if (mu[k] != mu[0] || lam[k] != lam[0]) uniform = false;
        }
        // sqrt(w |det J|) is used to keep K_e exactly symmetric; negative weights need the general path
        if (!(w[k] > 0.0)) uniform = false;
        reference_gradients(geometry_type(ctx->elem_type), q->points + (size_t)k * d, ggeo + (size_t)k * ng * d);
        reference_gradients(ctx->elem_type, q->points + (size_t)k * d, gref + (size_t)k * n * d);
    }
    if (!uniform && op->kind == FB200_LAPLACE) {
        // Laplace has no parameters: express it through the general path with mu = 1, lambda = 0
        for (int k = 0; k < nq; ++k) { mu[k] = 1.0; lam[k] = 0.0; }
    }
    // unchanged tables stay resident (no sync per call).  The derived fields depend on the operator kind as well - a Laplace table and an
    // elastic table with mu = lambda = 0 have identical contents - so they are refreshed on a hit too.
    const bool hit = ctx->tab.d_data && ctx->tab.host == h;
    ctx->tab.nq = nq;
    ctx->tab.uniform_params = uniform;
    ctx->tab.mu0 = op->kind == FB200_LINEAR_ELASTIC ? mu[0] : 1.0;
    ctx->tab.lam0 = op->kind == FB200_LINEAR_ELASTIC ? lam[0] : 0.0;
    if (hit) return FB200_OK;
    ctx->tab.host.clear();  // (a failed upload below must not leave a stale key)
    if (ctx->tab.capacity < len) {
        dev_free(ctx->tab.d_data);
        FB200_TRY(dev_alloc(ctx, &ctx->tab.d_data, len));
        ctx->tab.capacity = len;
    }
    FB200_CUDA(ctx, cudaMemcpyAsync(ctx->tab.d_data, h.data(), len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // h is pageable
    ctx->tab.host = h;
    return FB200_OK;
}

template <int N, int NG, int D, int OP, int MODE>
static fb200_status launch_elements(fb200_ctx* ctx, AssembleParams& p) {
    FB200_TRY(clear_values_if_pending(ctx));
    constexpr int G = (N <= 4) ? 8 : 32;
    constexpr int THREADS = 128;
    constexpr int SLOTS = THREADS / G;
    if (p.count == 0) return FB200_OK;
    const size_t smem = sizeof(double) * (table_doubles<N, NG, D>(p.nq) + SLOTS * slot_doubles<N, NG, D>(p.nq)) + SLOTS * N * (sizeof(long long) + sizeof(int));
    auto kernel = assemble_elements_kernel<N, NG, D, OP, MODE, G, THREADS>;
    if (smem > 48 * 1024) FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_UNSUPPORTED, "quadrature rule too large for shared memory");
    const uint64_t want = (p.count + SLOTS - 1) / SLOTS;
    const int blocks = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * per_sm);
    kernel<<<blocks, THREADS, smem, ctx->stream>>>(p);
    return check_launch(ctx, "assemble_elements_kernel");
}

constexpr int kHex8Chunk = 8;  // positions per ticket = one 2x2x2 Morton cell

// ---- fused zero-fill bookkeeping: which chunk of the processing order touches a node first
__global__ void first_pos_kernel(const int32_t* __restrict__ conn_pos, uint64_t count, int n, int* first_pos) {
    const uint64_t total = count * (uint64_t)n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) atomicMin(&first_pos[conn_pos[t]], (int)(t / n));
}
__global__ void zero_count_kernel(const int* __restrict__ first_pos, uint64_t num_nodes, int chunk, unsigned long long* cnt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_nodes; i += stride) {
        const int fp = first_pos[i];
        atomicAdd(&cnt[fp >= 0x7f000000 ? 0 : fp / chunk], 1ull);  // rows no owned element touches are cleared by chunk 0
    }
}
__global__ void zero_fill_kernel(const int* __restrict__ first_pos, uint64_t num_nodes, int chunk, unsigned long long* cursor, int32_t* nodes,
                                 const int64_t* __restrict__ blk_off, int ss, int64_t* base, int32_t* len) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_nodes; i += stride) {
        const int fp = first_pos[i];
        const unsigned long long slot = atomicAdd(&cursor[fp >= 0x7f000000 ? 0 : fp / chunk], 1ull);
        nodes[slot] = (int32_t)i;
        base[slot] = (int64_t)ss * blk_off[i];
        len[slot] = (int32_t)((blk_off[i + 1] - blk_off[i]) * ss);
    }
}

static fb200_status ensure_zero_lists(fb200_ctx* ctx, const OrderedCopy& oc) {
    if (ctx->zero_valid && ctx->zero_count == oc.count) return FB200_OK;
    dev_free(ctx->d_zero_off);
    dev_free(ctx->d_zero_nodes);
    dev_free(ctx->d_zero_base);
    dev_free(ctx->d_zero_len);
    ctx->zero_valid = false;
    const uint64_t N = ctx->N, chunks = (oc.count + kHex8Chunk - 1) / kHex8Chunk;
    int* d_first = nullptr;
    unsigned long long* d_cursor = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_first, N));
    fb200_status st = dev_alloc(ctx, &ctx->d_zero_off, chunks + 2);
    if (st == FB200_OK) st = dev_alloc(ctx, &ctx->d_zero_nodes, N);
    if (st == FB200_OK) st = dev_alloc(ctx, &ctx->d_zero_base, N);
    if (st == FB200_OK) st = dev_alloc(ctx, &ctx->d_zero_len, N);
    if (st == FB200_OK) st = dev_alloc(ctx, &d_cursor, chunks + 2);
    if (st == FB200_OK && !ctx->d_row_epoch) {
        st = dev_alloc(ctx, &ctx->d_row_epoch, N);
        if (st == FB200_OK) cudaMemsetAsync(ctx->d_row_epoch, 0, std::max<uint64_t>(N, 1) * sizeof(uint32_t), ctx->stream);
    }
    if (st == FB200_OK) {
        cudaMemsetAsync(d_first, 0x7f, N * sizeof(int), ctx->stream);  // 0x7f7f7f7f > any position; normalised below
        cudaMemsetAsync(ctx->d_zero_off, 0, (chunks + 2) * sizeof(int64_t), ctx->stream);
        const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(oc.count * ctx->ei.n, 256), (uint64_t)ctx->sm_count * 16));
        if (oc.count) {
            first_pos_kernel<<<blocks, 256, 0, ctx->stream>>>(oc.conn, oc.count, ctx->ei.n, d_first);
            st = check_launch(ctx, "first_pos_kernel");
        }
    }
    const int nb = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(N, 256), (uint64_t)ctx->sm_count * 16));
    if (st == FB200_OK && N) {
        zero_count_kernel<<<nb, 256, 0, ctx->stream>>>(d_first, N, kHex8Chunk, reinterpret_cast<unsigned long long*>(ctx->d_zero_off));
        st = check_launch(ctx, "zero_count_kernel");
    }
    if (st == FB200_OK) st = exclusive_scan_i64(ctx, ctx->d_zero_off, chunks + 1);
    if (st == FB200_OK && N) {
        cudaMemcpyAsync(d_cursor, ctx->d_zero_off, (chunks + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream);
        zero_fill_kernel<<<nb, 256, 0, ctx->stream>>>(d_first, N, kHex8Chunk, d_cursor, ctx->d_zero_nodes, ctx->d_blk_off, ctx->sdim * ctx->sdim,
                                                      ctx->d_zero_base, ctx->d_zero_len);
        st = check_launch(ctx, "zero_fill_kernel");
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_first);
    if (d_cursor) cudaFree(d_cursor);
    FB200_TRY(st);
    ctx->zero_valid = true;
    ctx->zero_count = oc.count;
    return FB200_OK;
}

// rows of a per-element array gathered into processing order: dst[pos] = src[ids[pos]]
__global__ void permute_rows_kernel(const uint32_t* __restrict__ src, const int32_t* __restrict__ ids, uint64_t count, int words_per_row,
                                    uint32_t* __restrict__ dst) {
    const uint64_t total = count * (uint64_t)words_per_row;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t pos = t / words_per_row;
        const int k = (int)(t - pos * words_per_row);
        dst[t] = src[(uint64_t)ids[pos] * words_per_row + k];
    }
}

// Connectivity and scatter map re-laid in a processing order (Morton order, or the concatenated colour lists), so that the
// element loop reads them with unit stride and without a dependent index load.
static fb200_status ensure_ordered(fb200_ctx* ctx, OrderedCopy& oc, const int32_t* d_ids, uint64_t count) {
    if (oc.valid && oc.count == count && oc.ids == d_ids) return FB200_OK;
    dev_free(oc.conn);
    dev_free(oc.map);
    oc.valid = false;
    const int n = ctx->ei.n;
    FB200_TRY(dev_alloc(ctx, &oc.conn, count * n));
    FB200_TRY(dev_alloc(ctx, &oc.map, count * (uint64_t)(n * n)));
    if (count) {
        const int blocks = (int)std::min<uint64_t>(div_up(count * n, 256), (uint64_t)ctx->sm_count * 16);
        permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(ctx->d_conn), d_ids, count, n,
                                                             reinterpret_cast<uint32_t*>(oc.conn));
        FB200_TRY(check_launch(ctx, "permute_rows_kernel<conn>"));
        permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(ctx->d_blockmap), d_ids, count, n * n / 2,
                                                             reinterpret_cast<uint32_t*>(oc.map));
        FB200_TRY(check_launch(ctx, "permute_rows_kernel<map>"));
    }
    oc.ids = d_ids;
    oc.count = count;
    oc.valid = true;
    return FB200_OK;
}

template <int OP, int MODE, int MINB, bool DYN, bool HINT, int CHUNK = 8, bool ZFUSE = false>
static fb200_status launch_hex8(fb200_ctx* ctx, AssembleParams& p) {
    if (!ZFUSE) FB200_TRY(clear_values_if_pending(ctx));
    constexpr int THREADS = 128, WARPS = THREADS / 32;
    constexpr int S = OP == FB200_LAPLACE ? 1 : 3, SN = S * 8, TS = 33;
    constexpr int KLEN = S == 1 ? SN * (SN + 1) : SN * SN + 4;
    if (p.count == 0) return FB200_OK;
    const int tab_len = (p.nq * (1 + 2 * TS) + 1) & ~1;
    const int warp_doubles = 24 + p.nq * TS + KLEN + (KLEN & 1) + 28;
    const size_t smem = sizeof(double) * (size_t)(tab_len + WARPS * warp_doubles);
    auto kernel = assemble_hex8_kernel<OP, MODE, THREADS, MINB, DYN, HINT, CHUNK, ZFUSE>;
    if (smem > 48 * 1024) FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    constexpr int BLOCK = THREADS + (ZFUSE ? 32 : 0);
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_UNSUPPORTED, "quadrature rule too large for shared memory");
    const uint64_t want = (p.count + WARPS - 1) / WARPS;
    // Resident CTAs per SM.  More warps hide latency better, but every element in flight widens the front of CSR rows that must
    // stay in L2 between a node's first and last contribution and adds same-address pressure on the L2 reduction units; 3 CTAs
    // (12 warps) per SM measured best for the atomic scatter (profiles/r01).
    static const int grid_cap = std::getenv("FB200_GRID_CAP") ? std::atoi(std::getenv("FB200_GRID_CAP")) : 3;
    if (MODE == MODE_ATOMIC && grid_cap > 0) per_sm = std::min(per_sm, grid_cap);
    const int blocks = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * per_sm);
    p.ticket32 = reinterpret_cast<unsigned int*>(ctx->d_ticket);
    if (DYN) FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned long long), ctx->stream));
    if (ZFUSE) {
        // element warps wait on flags written by the clearing warps of OTHER CTAs: all CTAs must be co-resident
        void* args[] = {(void*)&p};
        FB200_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kernel, dim3(blocks), dim3(BLOCK), args, smem, ctx->stream));
        return check_launch(ctx, "assemble_hex8_kernel<zfuse>");
    }
    kernel<<<blocks, BLOCK, smem, ctx->stream>>>(p);
    return check_launch(ctx, "assemble_hex8_kernel");
}

// FP64 tensor-core variant (hex8_mma_kernel.cuh): reference gradients in registers, S = G G^T by DMMA
template <int OP, int MODE, bool DYN, bool HINT, int CHUNK = 8>
static fb200_status launch_hex8_mma(fb200_ctx* ctx, AssembleParams& p) {
    FB200_TRY(clear_values_if_pending(ctx));
    constexpr int THREADS = 128, WARPS = THREADS / 32, MINB = 3;
    constexpr int S = OP == FB200_LAPLACE ? 1 : 3, SN = S * 8;
    constexpr int KLEN = S == 1 ? SN * (SN + 1) : SN * SN + 4;
    if (p.count == 0) return FB200_OK;
    constexpr int warp_doubles = 24 + 8 * 28 + KLEN + (KLEN & 1) + 28;
    const size_t smem = sizeof(double) * (size_t)(WARPS * warp_doubles);
    auto kernel = assemble_hex8_mma_kernel<OP, MODE, THREADS, MINB, DYN, HINT, CHUNK>;
    int per_sm = 1;
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_CUDA, "assemble_hex8_mma_kernel does not fit on an SM");
    const uint64_t want = (p.count + WARPS - 1) / WARPS;
    // resident CTAs per SM: see launch_hex8 (the front of CSR rows under accumulation must stay in L2)
    static const int grid_cap = std::getenv("FB200_GRID_CAP") ? std::atoi(std::getenv("FB200_GRID_CAP")) : 3;
    if (grid_cap > 0) per_sm = std::min(per_sm, grid_cap);
    const int blocks = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * per_sm);
    p.ticket32 = reinterpret_cast<unsigned int*>(ctx->d_ticket);
    static const int debug = std::getenv("FB200_DEBUG") ? std::atoi(std::getenv("FB200_DEBUG")) : 0;
    p.debug = debug;
    if (DYN) FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned long long), ctx->stream));
    kernel<<<blocks, THREADS, smem, ctx->stream>>>(p);
    return check_launch(ctx, "assemble_hex8_mma_kernel");
}

// ---- chunk-local scatter lists (host build, see chunks.cpp), cached per (order, pattern)
static void free_chunks(ChunkLists& cl) {
    dev_free(cl.d_slot_off);
    dev_free(cl.d_contrib);
    dev_free(cl.d_slot_node);
    dev_free(cl.d_slot_k);
    dev_free(cl.d_slot_cbeg);
    dev_free(cl.d_slot_flags);
    dev_free(cl.d_slot_dst);
    dev_free(cl.d_slot_rl);
    dev_free(cl.d_conn_pos);
    cl.valid = false;
    cl.count = 0;
    cl.ids = nullptr;
}

template <class T, class A>
static fb200_status upload_vec(fb200_ctx* ctx, T** d, const std::vector<T, A>& h) {
    FB200_TRY(dev_alloc(ctx, d, h.size()));
    if (!h.empty()) FB200_CUDA(ctx, h2d_copy(ctx, *d, h.data(), h.size() * sizeof(T)));
    return FB200_OK;
}

static fb200_status ensure_chunks(fb200_ctx* ctx, const int32_t* d_ids, uint64_t count, int chunk_elems) {
    ChunkLists& cl = ctx->chunks;
    if (cl.valid && cl.count == count && cl.ids == d_ids && cl.chunk_elems == chunk_elems && cl.sdim == ctx->sdim) return FB200_OK;
    free_chunks(cl);
    const int n = ctx->ei.n;
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> conn(ctx->E * n), ids(count);
    std::vector<int64_t> blk_off(ctx->N + 1);
    std::vector<uint16_t> map(ctx->E * (uint64_t)(n * n));
    FB200_CUDA(ctx, d2h_copy(ctx, conn.data(), ctx->d_conn, conn.size() * sizeof(int32_t)));
    FB200_CUDA(ctx, d2h_copy(ctx, ids.data(), d_ids, ids.size() * sizeof(int32_t)));
    FB200_CUDA(ctx, d2h_copy(ctx, blk_off.data(), ctx->d_blk_off, blk_off.size() * sizeof(int64_t)));
    FB200_CUDA(ctx, d2h_copy(ctx, map.data(), ctx->d_blockmap, map.size() * sizeof(uint16_t)));
    HostChunks hc;
    build_chunk_lists(n, ctx->sdim, count, chunk_elems, ids.data(), conn.data(), ctx->E, ctx->N, blk_off.data(), map.data(), hc);
    cl.sdim = ctx->sdim;
    cl.num_chunks = (uint32_t)(hc.slot_off.size() - 1);
    cl.total_slots = hc.slot_node.size();
    FB200_TRY(upload_vec(ctx, &cl.d_slot_off, hc.slot_off));
    FB200_TRY(upload_vec(ctx, &cl.d_contrib, hc.contrib));
    FB200_TRY(upload_vec(ctx, &cl.d_slot_node, hc.slot_node));
    FB200_TRY(upload_vec(ctx, &cl.d_slot_k, hc.slot_k));
    FB200_TRY(upload_vec(ctx, &cl.d_slot_cbeg, hc.slot_cbeg));
    FB200_TRY(upload_vec(ctx, &cl.d_slot_flags, hc.slot_flags));
    FB200_TRY(upload_vec(ctx, &cl.d_slot_dst, hc.slot_dst));
    FB200_TRY(upload_vec(ctx, &cl.d_slot_rl, hc.slot_rl));
    FB200_TRY(dev_alloc(ctx, &cl.d_conn_pos, count * n));
    if (count) {
        const int blocks = (int)std::min<uint64_t>(div_up(count * n, 256), (uint64_t)ctx->sm_count * 16);
        permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(ctx->d_conn), d_ids, count, n,
                                                             reinterpret_cast<uint32_t*>(cl.d_conn_pos));
        FB200_TRY(check_launch(ctx, "permute_rows_kernel<chunk conn>"));
    }
    cl.count = count;
    cl.ids = d_ids;
    cl.chunk_elems = chunk_elems;
    cl.valid = true;
    return FB200_OK;
}

template <int OP, int C, int T>
static fb200_status launch_tet4_chunks_t(fb200_ctx* ctx, AssembleParams& p) {
    FB200_TRY(clear_values_if_pending(ctx));
    if (p.count == 0) {  // (a rank without owned elements still takes part in the neighbour barriers of the fused exchange)
        if (ctx->p2p.enabled && ctx->p2p.num_peers > 0) {
            if (!p.accumulate) FB200_TRY(p2p_neighbour_barrier(ctx));
            ctx->p2p.pending = true;
        }
        return FB200_OK;
    }
    FB200_TRY(ensure_chunks(ctx, ctx->d_order, ctx->order_count, C));
    const ChunkLists& cl = ctx->chunks;
    p.conn_pos = cl.d_conn_pos;
    p.elem_ids = ctx->d_order;
    p.num_chunks = cl.num_chunks;
    p.slot_off = cl.d_slot_off;
    p.contrib = cl.d_contrib;
    p.slot_node = cl.d_slot_node;
    p.slot_k = cl.d_slot_k;
    p.slot_cbeg = cl.d_slot_cbeg;
    p.slot_flags = cl.d_slot_flags;
    p.slot_dst = cl.d_slot_dst;
    p.slot_rl = cl.d_slot_rl;
    // fused interface exchange (comm.cu): interface slots are also reduced into the neighbouring rank's rows; the values were just
    // cleared when the call overwrites, so no neighbour may add to them before that (neighbour barrier)
    const bool peer = ctx->p2p.enabled && ctx->p2p.num_peers > 0;
    if (peer) {
        p.peer_row = ctx->p2p.d_peer_row;
        p.peer_values[0] = ctx->p2p.values[0];
        p.peer_values[1] = ctx->p2p.values[1];
        if (!p.accumulate) FB200_TRY(p2p_neighbour_barrier(ctx));
        ctx->p2p.pending = true;
    }
    const size_t smem = sizeof(double) * 12 * C;
    auto kernel = peer ? assemble_tet4_chunk_kernel<OP, T, C, true> : assemble_tet4_chunk_kernel<OP, T, C, false>;
    FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, T, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_CUDA, "assemble_tet4_chunk_kernel does not fit on an SM");
    const int blocks = (int)std::min<uint64_t>(cl.num_chunks, (uint64_t)ctx->sm_count * per_sm);
    p.ticket32 = reinterpret_cast<unsigned int*>(ctx->d_ticket);
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned long long), ctx->stream));
    kernel<<<blocks, T, smem, ctx->stream>>>(p);
    return check_launch(ctx, "assemble_tet4_chunk_kernel");
}

// chunk size: larger chunks merge more contributions per reduction (BCC tets: 28 / 23 / 17 reductions per element at 512 / 1024 /
// 2048 elements per chunk instead of 144), smaller ones leave more CTAs per SM to hide the list-load latency
template <int OP>
static fb200_status launch_tet4_chunks(fb200_ctx* ctx, AssembleParams& p) {
    static const int chunk = std::getenv("FB200_TET4_CHUNK") ? std::atoi(std::getenv("FB200_TET4_CHUNK")) : 1024;
    if (chunk == 512) return launch_tet4_chunks_t<OP, 512, 256>(ctx, p);
    if (chunk == 2048) return launch_tet4_chunks_t<OP, 2048, 1024>(ctx, p);
    return launch_tet4_chunks_t<OP, 1024, 512>(ctx, p);
}

// ---- tile lists of the Hex8 tile kernel (host build, see tiles.cpp), cached per (order, pattern, tile shape)
static void free_tiles(TileLists& tl) {
    dev_free(tl.d_hdr);
    dev_free(tl.d_nodes);
    dev_free(tl.d_flush);
    dev_free(tl.d_wait);
    dev_free(tl.d_colour_tiles);
    tl.colour_off.clear();
    dev_free(tl.d_flag);
    dev_free(tl.d_zero_nodes);
    dev_free(tl.d_lnodes);
    dev_free(tl.d_emap);
    dev_free(tl.d_elem);
    tl.valid = false;
    tl.unusable = false;
    tl.owner = false;
    tl.count = 0;
    tl.ids = nullptr;
    tl.num_tiles = 0;
    tl.zero_node_count = 0;
}

static fb200_status ensure_tiles(fb200_ctx* ctx, const TileShape& shape) {
    TileLists& tl = ctx->tiles;
    const int32_t* d_ids = ctx->d_order;
    const uint64_t count = ctx->order_count;
    if ((tl.valid || tl.unusable) && tl.count == count && tl.ids == d_ids && tl.tile_bits == shape.tile_bits && tl.owner_stores == shape.owner_stores)
        return FB200_OK;
    free_tiles(tl);
    tl.count = count;
    tl.ids = d_ids;
    tl.tile_bits = shape.tile_bits;
    tl.owner_stores = shape.owner_stores;
    if (ctx->h_order_codes.size() != count) {
        tl.unusable = true;
        return FB200_OK;
    }
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SetupTimer tm;
    BigVec<int32_t> conn(ctx->E * 8), ids(count);
    BigVec<uint16_t> map(ctx->E * (uint64_t)64);
    BigVec<int64_t> blk_off(ctx->N + 1);
    FB200_CUDA(ctx, d2h_staged(ctx, conn.data(), ctx->d_conn, conn.size() * sizeof(int32_t)));
    FB200_CUDA(ctx, d2h_staged(ctx, ids.data(), d_ids, ids.size() * sizeof(int32_t)));
    FB200_CUDA(ctx, d2h_staged(ctx, map.data(), ctx->d_blockmap, map.size() * sizeof(uint16_t)));
    FB200_CUDA(ctx, d2h_staged(ctx, blk_off.data(), ctx->d_blk_off, blk_off.size() * sizeof(int64_t)));
    tm.lap("tiles: download conn / map / offsets");
    HostTiles ht;
    build_tile_lists(shape, count, ids.data(), ctx->h_order_codes.data(), conn.data(), ctx->E, ctx->E_owned, ctx->N, map.data(), blk_off.data(), ht);
    tm.lap("tiles: build_tile_lists");
    if (ht.bank_conflict_share < 0.0 || ht.flush.size() >= (1ull << 32) || ht.nodes.size() >= (1ull << 32)) {
        tl.unusable = true;  // degenerate elements (repeated nodes) or lists beyond 32-bit offsets: keep the per-element kernel
        return FB200_OK;
    }
    tl.num_tiles = (uint32_t)(ht.hdr.size() / kTileHdrWords);
    FB200_TRY(upload_vec(ctx, &tl.d_hdr, ht.hdr));
    FB200_TRY(upload_vec(ctx, &tl.d_nodes, ht.nodes));
    FB200_TRY(upload_vec(ctx, &tl.d_flush, ht.flush));
    FB200_TRY(upload_vec(ctx, &tl.d_wait, ht.wait));
    FB200_TRY(upload_vec(ctx, &tl.d_colour_tiles, ht.colour_tiles));
    tl.colour_off = ht.colour_off;
    FB200_TRY(upload_vec(ctx, &tl.d_zero_nodes, ht.zero_nodes));
    FB200_TRY(upload_vec(ctx, &tl.d_lnodes, ht.lnodes));
    FB200_TRY(upload_vec(ctx, &tl.d_emap, ht.emap));
    FB200_TRY(upload_vec(ctx, &tl.d_elem, ht.elem));
    FB200_TRY(dev_alloc(ctx, &tl.d_flag, tl.num_tiles));
    FB200_CUDA(ctx, cudaMemsetAsync(tl.d_flag, 0, std::max<size_t>(tl.num_tiles, 1) * sizeof(uint32_t), ctx->stream));
    tm.lap("tiles: upload lists");
    tl.zero_node_count = ht.zero_nodes.size();
    tl.owner = ht.owner_stores;
    tl.valid = true;
    return FB200_OK;
}

static const TileShape kHex8TileShape{6, 64, kTileGroupWarps, 128, 1216};
static int hex8_tile_setting(const fb200_ctx* ctx) {
    static const int env_tile = std::getenv("FB200_HEX8_TILE") ? std::atoi(std::getenv("FB200_HEX8_TILE")) : 64;
    return ctx->tune_hex8_tile >= 0 ? ctx->tune_hex8_tile : env_tile;
}
static int hex8_owner_setting(const fb200_ctx* ctx) {
    static const int env_owner = std::getenv("FB200_HEX8_OWNER") ? std::atoi(std::getenv("FB200_HEX8_OWNER")) : 1;
    return (ctx->tune_owner >= 0 ? ctx->tune_owner : env_owner) ? 1 : 0;
}

// the s rows of every listed node = 0 (one warp per node)
__global__ void zero_node_rows_kernel(const int32_t* __restrict__ nodes, uint64_t count, const int64_t* __restrict__ blk_off, int ss,
                                      double* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t k = gwarp; k < count; k += nwarps) {
        const int32_t node = nodes[k];
        const int64_t b = blk_off[node];
        const int64_t len = (blk_off[node + 1] - b) * ss;
        double* v = values + b * ss;
        for (int64_t t = lane; t < len; t += 32) v[t] = 0.0;
    }
}

// Hex8 tile kernel (hex8_tile_kernel.cuh): a CTA accumulates a tile of the Morton order in shared memory and updates every CSR
// node block of the tile once.  *used = false when the mesh has no usable tile lists (the caller falls back to the element kernel).
template <int OP, int MAXN, int MAXP>
static fb200_status launch_hex8_tile_t(fb200_ctx* ctx, AssembleParams& p, const TileShape& shape, bool* used) {
    *used = false;
    FB200_TRY(ensure_tiles(ctx, shape));
    const TileLists& tl = ctx->tiles;
    if (!tl.valid) return FB200_OK;
    *used = true;
    p.tile_epoch = 0;
    if (ctx->pending_zero) {
        if (tl.owner) {
            // owner lists: every row has a storing tile, except the rows ghost elements touch (and rows of nodes without owned
            // elements) - those few are cleared here
            ctx->pending_zero = false;
            if (++ctx->tile_epoch == 0) {  // the 32-bit epoch wrapped: restart the flags
                FB200_CUDA(ctx, cudaMemsetAsync(tl.d_flag, 0, std::max<size_t>(tl.num_tiles, 1) * sizeof(uint32_t), ctx->stream));
                ctx->tile_epoch = 1;
            }
            p.tile_epoch = ctx->tile_epoch;
            if (tl.zero_node_count) {
                const int ss = ctx->sdim * ctx->sdim;
                const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(tl.zero_node_count * 32, 256), (uint64_t)ctx->sm_count * 8));
                zero_node_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(tl.d_zero_nodes, tl.zero_node_count, ctx->d_blk_off, ss, ctx->d_values);
                FB200_TRY(check_launch(ctx, "zero_node_rows_kernel"));
            }
        } else {
            FB200_TRY(clear_values_if_pending(ctx));
        }
    }
    p.num_tiles = tl.num_tiles;
    p.tile_hdr = tl.d_hdr;
    p.tile_nodes = tl.d_nodes;
    p.tile_flush = tl.d_flush;
    p.tile_lnodes = tl.d_lnodes;
    p.tile_emap = tl.d_emap;
    p.tile_elem = tl.d_elem;
    p.tile_wait = tl.d_wait;
    p.tile_flag = tl.d_flag;
    // fused interface exchange (comm.cu): the flush also reduces interface rows into the neighbouring ranks' values.  The rows were
    // just cleared when the call overwrites: no neighbour may add to them before that, hence the neighbour barrier
    const bool peer = ctx->p2p.enabled && ctx->p2p.num_peers > 0;
    if (peer) {
        p.peer_row = ctx->p2p.d_peer_row;
        p.peer_values[0] = ctx->p2p.values[0];
        p.peer_values[1] = ctx->p2p.values[1];
        if (!p.accumulate) FB200_TRY(p2p_neighbour_barrier(ctx));
        ctx->p2p.pending = true;
    }
    if (tl.num_tiles == 0) return FB200_OK;  // (a rank without owned elements still takes part in the neighbour barriers)
    const size_t smem = Hex8TileSmem<OP, MAXN, MAXP>::bytes;
    auto kernel = peer ? assemble_hex8_tile_kernel<OP, MAXN, MAXP, true> : assemble_hex8_tile_kernel<OP, MAXN, MAXP, false>;
    FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    constexpr int THREADS = (2 * kTileGroupWarps + kTileHelperWarps) * 32;
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_CUDA, "assemble_hex8_tile_kernel does not fit on an SM");
    const int blocks = (int)std::min<uint64_t>(tl.num_tiles, (uint64_t)ctx->sm_count);  // one persistent CTA per SM
    p.ticket32 = reinterpret_cast<unsigned int*>(ctx->d_ticket);
    static const int debug = std::getenv("FB200_DEBUG") ? std::atoi(std::getenv("FB200_DEBUG")) : 0;
    p.debug = debug;
    // owner stores make a tile wait for lower-numbered tiles.  With the atomic ticket a CTA holds the six tiles it fetches tables for; a
    // CTA that waits lets them age while the others take newer tiles that depend on exactly those: measured 5.3 - 6.2 ms per C3
    // assembly instead of 2.45.  Round-robin keeps Morton neighbours in the same step of neighbouring CTAs: waits of ~3 polls, 2.58 ms
    // (profiles/r02/README.md).  Without waits (accumulating call, zero-fill lists) the ticket balances the SMs better: 2.45 vs 2.52 ms.
    static const int env_static = std::getenv("FB200_TILE_STATIC") ? std::atoi(std::getenv("FB200_TILE_STATIC")) : -1;
    p.tile_static = env_static >= 0 ? env_static : (p.tile_epoch != 0 ? 1 : 0);
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned long long), ctx->stream));
    if (p.tile_static && p.tile_epoch) {
        // round-robin tiles + waits between tiles: every CTA must be resident (with the ticket a waited-for tile is always held by a
        // running CTA; here tile c + k CTAs is held by CTA c whether it runs or not).  A cooperative launch guarantees that or fails;
        // then the ticket schedule is used - slower with waits, but free of that requirement.
        void* args[] = {(void*)&p};
        const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)kernel, dim3(blocks), dim3(THREADS), args, smem, ctx->stream);
        if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorNotSupported || ce == cudaErrorLaunchOutOfResources) {
            cudaGetLastError();
            p.tile_static = 0;
            kernel<<<blocks, THREADS, smem, ctx->stream>>>(p);
        } else if (ce != cudaSuccess) {
            return cuda_fail(ctx, ce, "cudaLaunchCooperativeKernel(assemble_hex8_tile_kernel)");
        }
    } else {
        kernel<<<blocks, THREADS, smem, ctx->stream>>>(p);
    }
    if (debug & 64) {  // diagnostics of the owner-store waits (hex8_tile_kernel.cuh)
        unsigned long long dc[16];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(dc, ctx->d_ticket, sizeof(dc), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "[tile waits] blocked %llu spins %llu max %llu | owner distance: mean %.1f max %llu | blocked by distance <8 %llu <64 %llu <512 %llu more %llu | spins %llu %llu %llu %llu\n",
                     dc[2], dc[3], dc[4], dc[2] ? (double)dc[5] / (double)dc[2] : 0.0, dc[6], dc[7], dc[8], dc[9], dc[10], dc[11], dc[12], dc[13], dc[14]);
        cudaMemset(ctx->d_ticket, 0, sizeof(dc));
    }
    return check_launch(ctx, "assemble_hex8_tile_kernel");
}

// Deterministic coloured scatter of a Hex8 space (FB200_SCATTER_COLORED; the reference's CsrParAssembler, global.rs:314-376, at tile
// granularity): one launch of the tile kernel per TILE colour.  Tiles of a colour share no node, so a launch adds at most once to every
// CSR value; sums inside a tile are formed in a fixed order and the launches run in colour order: the result is bitwise reproducible.
template <int OP, int MAXN, int MAXP>
static fb200_status launch_hex8_tile_colored_t(fb200_ctx* ctx, AssembleParams& p, const TileShape& shape, bool* used) {
    *used = false;
    FB200_TRY(ensure_tiles(ctx, shape));
    const TileLists& tl = ctx->tiles;
    if (!tl.valid || tl.colour_off.size() < 2) return FB200_OK;
    *used = true;
    FB200_TRY(clear_values_if_pending(ctx));
    AssembleParams q = p;
    q.tile_epoch = 0;
    q.tile_static = 0;
    q.accumulate = 1;  // every flush entry is a reduction onto the cleared (or the caller's) values
    q.tile_hdr = tl.d_hdr;
    q.tile_nodes = tl.d_nodes;
    q.tile_flush = tl.d_flush;
    q.tile_lnodes = tl.d_lnodes;
    q.tile_emap = tl.d_emap;
    q.tile_elem = tl.d_elem;
    q.tile_wait = tl.d_wait;
    q.tile_flag = tl.d_flag;
    q.ticket32 = reinterpret_cast<unsigned int*>(ctx->d_ticket);
    const size_t smem = Hex8TileSmem<OP, MAXN, MAXP>::bytes;
    auto kernel = assemble_hex8_tile_kernel<OP, MAXN, MAXP, false>;
    FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int THREADS = (2 * kTileGroupWarps + kTileHelperWarps) * 32;
    for (size_t c = 0; c + 1 < tl.colour_off.size(); ++c) {
        const uint64_t n = tl.colour_off[c + 1] - tl.colour_off[c];
        if (n == 0) continue;
        q.tile_list = tl.d_colour_tiles + tl.colour_off[c];
        q.num_tiles = (uint32_t)n;
        const int blocks = (int)std::min<uint64_t>(n, (uint64_t)ctx->sm_count);
        FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned long long), ctx->stream));
        kernel<<<blocks, THREADS, smem, ctx->stream>>>(q);
        FB200_TRY(check_launch(ctx, "assemble_hex8_tile_kernel<colour>"));
    }
    return FB200_OK;
}

// tile shape: 4 x 4 x 4 elements, one CTA of 16 compute + 8 helper warps per SM (double-buffered accumulators: 2 x 85.5 KB);
// FB200_HEX8_TILE = 64 | 0 (off), overridden by fb200_set_tuning("hex8_tile"); FB200_HEX8_OWNER / "hex8_owner_stores": first-writer stores
template <int OP>
static fb200_status launch_hex8_tile(fb200_ctx* ctx, AssembleParams& p, bool* used) {
    *used = false;
    if (hex8_tile_setting(ctx) != 64) return FB200_OK;
    TileShape shape = kHex8TileShape;
    shape.owner_stores = hex8_owner_setting(ctx);
    return launch_hex8_tile_t<OP, 128, 1216>(ctx, p, shape, used);
}

// Hex27 / Hex20 / Tet10: one CTA per element, DMMA node-block contraction (hex27_mma_kernel.cuh, templated on the node counts)
template <int N, int NG, int OP, int MODE>
static fb200_status launch_hex27_mma(fb200_ctx* ctx, AssembleParams& p) {
    FB200_TRY(clear_values_if_pending(ctx));
    if (p.count == 0) return FB200_OK;
    const size_t smem = hex27_smem_bytes<OP, N>(p.nq);
    auto kernel = assemble_hex27_mma_kernel<OP, MODE, N, NG>;
    FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kH27Threads, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_CUDA, "assemble_hex27_mma_kernel does not fit on an SM");
    const int blocks = (int)std::min<uint64_t>(p.count, (uint64_t)ctx->sm_count * per_sm);
    p.ticket32 = reinterpret_cast<unsigned int*>(ctx->d_ticket);
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned long long), ctx->stream));
    kernel<<<blocks, kH27Threads, smem, ctx->stream>>>(p);
    return check_launch(ctx, "assemble_hex27_mma_kernel");
}

// element-parallel launch: the Hex8 warp kernel when it applies, the generic kernel otherwise
template <int N, int NG, int D, int OP, int MODE>
static fb200_status launch_element_parallel(fb200_ctx* ctx, AssembleParams& p) {
    if (p.generic_only) return launch_elements<N, NG, D, OP, MODE>(ctx, p);
    if constexpr (N == 8 && NG == 8 && D == 3) {
        static const bool force_v1 = std::getenv("FB200_HEX8_V1") != nullptr;
        if (p.uniform && !force_v1) {
            // hand the kernel position-indexed connectivity / map rows (no dependent index load in the element loop)
            p.first_elem = 0;
            if (MODE == MODE_DUMP) {
                p.conn_pos = p.conn + p.dump_first * 8;
                p.map_pos = nullptr;
                p.elem_ids = nullptr;
                p.first_elem = p.dump_first;
            } else if (p.elem_list == nullptr) {
                p.conn_pos = p.conn;
                p.map_pos = p.blockmap;
                p.elem_ids = nullptr;
            } else if (MODE == MODE_ATOMIC) {
                if (p.nq <= 8 && !p.zfuse && p.elem_list == ctx->d_order && p.count == ctx->order_count) {
                    bool used = false;
                    FB200_TRY(launch_hex8_tile<OP>(ctx, p, &used));
                    if (used) return FB200_OK;
                }
                FB200_TRY(ensure_ordered(ctx, ctx->ord_morton, ctx->d_order, ctx->order_count));
                p.conn_pos = ctx->ord_morton.conn;
                p.map_pos = ctx->ord_morton.map;
                p.elem_ids = ctx->d_order;
                if (p.zfuse) {
                    FB200_TRY(ensure_zero_lists(ctx, ctx->ord_morton));
                    p.zero_off = ctx->d_zero_off;
                    p.zero_nodes = ctx->d_zero_nodes;
                    p.zero_base = ctx->d_zero_base;
                    p.zero_len = ctx->d_zero_len;
                    p.row_epoch = ctx->d_row_epoch;
                    p.epoch = ++ctx->epoch;
                    p.num_chunks = (uint32_t)((ctx->ord_morton.count + kHex8Chunk - 1) / kHex8Chunk);
                    static const int look = std::getenv("FB200_ZERO_LOOK") ? std::atoi(std::getenv("FB200_ZERO_LOOK")) : 2048;
                    p.zero_look = (uint32_t)look;
                }
            } else {
                FB200_TRY(ensure_ordered(ctx, ctx->ord_colors, ctx->d_color_elems, ctx->h_color_elems.size()));
                const uint64_t off = (uint64_t)(p.elem_list - ctx->d_color_elems);
                p.conn_pos = ctx->ord_colors.conn + off * 8;
                p.map_pos = ctx->ord_colors.map + off * 64;
                p.elem_ids = p.elem_list;
            }
            // tuning knobs (environment, read once); defaults = best measured on B200 (profiles/r01)
            static const bool dyn = std::getenv("FB200_STATIC_SCHED") == nullptr;
            static const bool hint = std::getenv("FB200_NO_L2_HINTS") == nullptr;
            static const bool use_mma = std::getenv("FB200_HEX8_DFMA") == nullptr;
            if (use_mma && p.nq <= 8 && !p.zfuse) {
                if constexpr (MODE == MODE_ATOMIC) {
                    if (dyn) return hint ? launch_hex8_mma<OP, MODE, true, true, kHex8Chunk>(ctx, p) : launch_hex8_mma<OP, MODE, true, false, kHex8Chunk>(ctx, p);
                    return hint ? launch_hex8_mma<OP, MODE, false, true>(ctx, p) : launch_hex8_mma<OP, MODE, false, false>(ctx, p);
                } else {
                    return launch_hex8_mma<OP, MODE, false, false>(ctx, p);
                }
            }
            if constexpr (MODE == MODE_ATOMIC) {
                if (p.zfuse) return launch_hex8<OP, MODE, 4, true, true, kHex8Chunk, true>(ctx, p);  // 160-thread CTAs: 4 per SM keep ~96 registers
                if (dyn) return hint ? launch_hex8<OP, MODE, 6, true, true, kHex8Chunk>(ctx, p) : launch_hex8<OP, MODE, 6, true, false, kHex8Chunk>(ctx, p);
                return hint ? launch_hex8<OP, MODE, 6, false, true>(ctx, p) : launch_hex8<OP, MODE, 6, false, false>(ctx, p);
            } else {
                return launch_hex8<OP, MODE, 6, false, false>(ctx, p);
            }
        }
    }
    if constexpr ((N == 27 || N == 20) && NG == 8 && D == 3) {
        // the dense high-order contraction on the FP64 tensor pipe: Hex27 (81 x 81) and Hex20 (60 x 60)
        static const bool force_v1 = std::getenv("FB200_HEX27_V1") != nullptr;
        if (p.uniform && p.nq <= kH27GQ && !force_v1) return launch_hex27_mma<N, NG, OP, MODE>(ctx, p);
    }
    if constexpr (N == 10 && NG == 4 && D == 3) {
        // Tet10 (30 x 30 K_e, 4 points = one k-step) also instantiates the tensor-pipe kernel, but a CTA of 10 warps per element is too much
        // machinery for it: measured 2.86 ms against 1.78 ms for the generic element kernel on 324 000 elements (profiles/r02/README.md),
        // so the generic kernel stays the default - north_star reserves the tensor path for dense >= 60 x 60 contractions anyway.
        // FB200_TET10_MMA=1 selects the DMMA instantiation (parity-tested the same way).
        static const bool use_mma = std::getenv("FB200_TET10_MMA") != nullptr;
        if (p.uniform && p.nq <= kH27GQ && use_mma) return launch_hex27_mma<N, NG, OP, MODE>(ctx, p);
    }
    return launch_elements<N, NG, D, OP, MODE>(ctx, p);
}

template <int N, int NG, int D, int OP>
static fb200_status launch_gather(fb200_ctx* ctx, AssembleParams& p) {
    constexpr int GE = (N <= 8) ? 8 : (N <= 16 ? 16 : 32);
    constexpr int WARPS = 4;
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    if (p.num_nodes == 0) return FB200_OK;
    if (ctx->max_row_blocks > kGatherMaxRowBlocks)
        return fail(ctx, FB200_ERR_UNSUPPORTED, "gather scatter supports at most 128 coupled nodes per node; use ATOMIC or COLORED");
    const size_t smem = sizeof(double) * (table_doubles<N, NG, D>(p.nq) + WARPS * ((32 / GE) * slot_doubles<N, NG, D>(p.nq) + S * S * kGatherMaxRowBlocks));
    auto kernel = assemble_gather_kernel<N, NG, D, OP, GE, WARPS>;
    if (smem > 48 * 1024) FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    FB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    if (per_sm < 1) return fail(ctx, FB200_ERR_UNSUPPORTED, "quadrature rule too large for shared memory");
    const uint64_t want = (p.num_nodes + WARPS - 1) / WARPS;
    const int blocks = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * per_sm * 4);
    kernel<<<blocks, WARPS * 32, smem, ctx->stream>>>(p);
    return check_launch(ctx, "assemble_gather_kernel");
}

template <int N, int NG, int D, int OP>
static fb200_status dispatch_mode(fb200_ctx* ctx, AssembleParams& p, int mode) {
    switch (mode) {
        case FB200_SCATTER_ATOMIC: {
            // elements are visited in a locality-preserving (Morton) order so that all contributions to a CSR row
            // arrive while the row is still resident in L2; the sum does not depend on the order (up to fp rounding)
            static const bool no_order = std::getenv("FB200_NO_ORDER") != nullptr;
            if (p.generic_only) return launch_element_parallel<N, NG, D, OP, MODE_ATOMIC>(ctx, p);
            if (ctx->d_order && !no_order && ctx->order_count == p.count) p.elem_list = ctx->d_order;
            if constexpr (N == 4 && NG == 4 && D == 3) {
                static const bool tet_v1 = std::getenv("FB200_TET4_V1") != nullptr;
                if (p.uniform && p.nq == 1 && p.elem_list != nullptr && !tet_v1) return launch_tet4_chunks<OP>(ctx, p);
            }
            return launch_element_parallel<N, NG, D, OP, MODE_ATOMIC>(ctx, p);
        }
        case FB200_SCATTER_GATHER: return launch_gather<N, NG, D, OP>(ctx, p);
        case MODE_DUMP: return launch_element_parallel<N, NG, D, OP, MODE_DUMP>(ctx, p);
        case MODE_COLORED_LIST: return launch_element_parallel<N, NG, D, OP, MODE_COLORED>(ctx, p);
        case FB200_SCATTER_COLORED: {
            if constexpr (N == 8 && NG == 8 && D == 3) {
                // Hex8, uniform operator data: colours of tiles instead of colours of elements (deterministic, ~5x faster)
                static const int env_ct = std::getenv("FB200_HEX8_COLORED_TILES") ? std::atoi(std::getenv("FB200_HEX8_COLORED_TILES")) : 1;
                const bool want = ctx->tune_colored_tiles >= 0 ? ctx->tune_colored_tiles != 0 : env_ct != 0;
                if (want && !p.generic_only && p.uniform && p.nq <= 8 && hex8_tile_setting(ctx) == 64 && ctx->d_order && ctx->order_count == p.count &&
                    std::getenv("FB200_HEX8_V1") == nullptr) {
                    bool used = false;
                    TileShape shape = kHex8TileShape;
                    shape.owner_stores = hex8_owner_setting(ctx);
                    FB200_TRY((launch_hex8_tile_colored_t<OP, 128, 1216>(ctx, p, shape, &used)));
                    if (used) return FB200_OK;
                }
            }
            const uint64_t ncol = ctx->h_color_off.size() - 1;
            for (uint64_t c = 0; c < ncol; ++c) {
                AssembleParams pc = p;
                pc.elem_list = ctx->d_color_elems + ctx->h_color_off[c];
                pc.count = ctx->h_color_off[c + 1] - ctx->h_color_off[c];
                FB200_TRY((launch_element_parallel<N, NG, D, OP, MODE_COLORED>(ctx, pc)));
            }
            return FB200_OK;
        }
        default: return fail(ctx, FB200_ERR_UNSUPPORTED, "unknown scatter mode");
    }
}

template <int N, int NG, int D>
static fb200_status dispatch_op(fb200_ctx* ctx, AssembleParams& p, int op, int mode) {
    if (op == FB200_LAPLACE) return dispatch_mode<N, NG, D, FB200_LAPLACE>(ctx, p, mode);
    return dispatch_mode<N, NG, D, FB200_LINEAR_ELASTIC>(ctx, p, mode);
}

static fb200_status dispatch(fb200_ctx* ctx, AssembleParams& p, int op, int mode) {
    switch (ctx->elem_type) {
        case FB200_QUAD4: return dispatch_op<4, 4, 2>(ctx, p, op, mode);
        case FB200_TET4: return dispatch_op<4, 4, 3>(ctx, p, op, mode);
        case FB200_HEX8: return dispatch_op<8, 8, 3>(ctx, p, op, mode);
        case FB200_HEX27: return dispatch_op<27, 8, 3>(ctx, p, op, mode);
        case FB200_TET10: return dispatch_op<10, 4, 3>(ctx, p, op, mode);
        case FB200_HEX20: return dispatch_op<20, 8, 3>(ctx, p, op, mode);
        default: return fail(ctx, FB200_ERR_UNSUPPORTED, "element type has no device specialisation (no CPU fallback)");
    }
}

void free_ordered(fb200_ctx* ctx) {
    free_chunks(ctx->chunks);
    free_tiles(ctx->tiles);
    dev_free(ctx->d_zero_off);
    dev_free(ctx->d_zero_nodes);
    dev_free(ctx->d_zero_base);
    dev_free(ctx->d_zero_len);
    ctx->zero_valid = false;
    for (OrderedCopy* oc : {&ctx->ord_morton, &ctx->ord_colors}) {
        dev_free(oc->conn);
        dev_free(oc->map);
        oc->valid = false;
        oc->count = 0;
        oc->ids = nullptr;
    }
}

static void fill_params(fb200_ctx* ctx, AssembleParams& p) {
    std::memset(&p, 0, sizeof(p));
    p.vertices = ctx->d_vertices;
    p.conn = ctx->d_conn;
    p.blk_off = ctx->d_blk_off;
    p.blockmap = ctx->d_blockmap;
    p.values = ctx->d_values;
    p.tab = ctx->tab.d_data;
    p.nq = ctx->tab.nq;
    p.uniform = ctx->tab.uniform_params ? 1 : 0;
    p.mu = ctx->tab.mu0;
    p.lam = ctx->tab.lam0;
    p.elem_list = nullptr;
    p.count = ctx->E_owned;
    p.errword = ctx->d_errword;
    p.adj_off = ctx->d_adj_off;
    p.adj_inc = ctx->d_adj_inc;
    p.num_nodes = ctx->N;
    p.num_owned = ctx->E_owned;
}

}  // namespace fb200

// element lists per (colour, rule) of a quadrature table with a rule per element - one pseudo colour holding all owned elements for the
// ATOMIC scatter; group (c, r) is flat[off[c * num_rules + r] .. off[c * num_rules + r + 1])
fb200_status fb200::group_elements_by_rule(fb200_ctx* ctx, uint32_t num_rules, const uint32_t* element_rule, bool colored,
                                           std::vector<uint64_t>& col_off, std::vector<int32_t>& flat, std::vector<uint64_t>& off) {
    col_off = {0, ctx->E_owned};
    if (colored) col_off = ctx->h_color_off;
    std::vector<std::vector<int32_t>> lists((col_off.size() - 1) * (size_t)num_rules);
    for (size_t c = 0; c + 1 < col_off.size(); ++c)
        for (uint64_t k = col_off[c]; k < col_off[c + 1]; ++k) {
            const uint64_t e = colored ? ctx->h_color_elems[k] : k;
            if (e >= ctx->E_owned) continue;  // ghost elements of a partition are not assembled
            const uint32_t r = element_rule[e];
            if (r >= num_rules) return fail(ctx, FB200_ERR_INDEX_OOB, "element_to_rule_map entry out of bounds (quadrature_table.rs:361-366)", (int64_t)e);
            lists[c * num_rules + r].push_back((int32_t)e);
        }
    flat.clear();
    off.assign(lists.size() + 1, 0);
    for (size_t i = 0; i < lists.size(); ++i) {
        off[i + 1] = off[i] + lists[i].size();
        flat.insert(flat.end(), lists[i].begin(), lists[i].end());
    }
    return FB200_OK;
}

using namespace fb200;

extern "C" {

fb200_status fb200_assemble_into_csr_device(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                            int32_t scatter_mode, int32_t accumulate) {
    // linear operators: K does not depend on u (laplace.rs:62, materials.rs:110); StVK / NeoHookean: tangent stiffness at u (mass_source.cu)
    if (ctx && op && (op->kind == FB200_STVK || op->kind == FB200_NEO_HOOKEAN)) return assemble_state_dependent(ctx, op, q, u, scatter_mode, accumulate);
    FB200_TRY(validate(ctx, op, q));
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern: call fb200_assemble_pattern or fb200_pattern_adopt first");
    const int s = op->kind == FB200_LAPLACE ? 1 : ctx->ei.d;
    if (s != ctx->sdim) return fail(ctx, FB200_ERR_SHAPE, "pattern solution_dim does not match the operator");
    if (scatter_mode == FB200_SCATTER_COLORED && !ctx->has_colors)
        return fail(ctx, FB200_ERR_STATE, "coloured scatter needs fb200_color_nodes or fb200_colors_adopt");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->p2p.pending) {  // a fused assembly was never closed by fb200_interface_allreduce: neighbours may still be adding to our rows
        ctx->p2p.pending = false;
        FB200_TRY(p2p_neighbour_barrier(ctx));
    }
    FB200_TRY(upload_tables(ctx, op, q));
    if (scatter_mode == FB200_SCATTER_GATHER) FB200_TRY(build_adjacency(ctx));
    AssembleParams p;
    fill_params(ctx, p);
    p.accumulate = accumulate ? 1 : 0;
    // values = contributions: the Hex8 atomic kernel clears rows itself just before their first contribution (fused zero-fill);
    // every other path starts from an explicit memset
    // (measured slower than memset + kernel so far - the clearing warps stall on the fence that publishes the rows - so it is
    //  opt-in: FB200_ZFUSE=1; see profiles/r01/README.md)
    static const bool no_zfuse = std::getenv("FB200_ZFUSE") == nullptr || std::getenv("FB200_STATIC_SCHED") != nullptr ||
                                 std::getenv("FB200_NO_ORDER") != nullptr || std::getenv("FB200_HEX8_V1") != nullptr;
    p.zfuse = (!accumulate && scatter_mode == FB200_SCATTER_ATOMIC && ctx->elem_type == FB200_HEX8 && p.uniform && !no_zfuse && ctx->d_order &&
               ctx->order_count == p.count && p.count > 0)
                  ? 1
                  : 0;
    // "values = contributions": whichever kernel runs first clears the values (clear_values_if_pending) - except the Hex8 tile kernel with
    // owner lists, which stores every row itself and needs no zero-fill pass (hex8_tile_kernel.cuh "ownership")
    ctx->pending_zero = !accumulate && scatter_mode != FB200_SCATTER_GATHER && !p.zfuse;
    const fb200_status st = dispatch(ctx, p, op->kind, scatter_mode);
    if (st == FB200_OK) FB200_TRY(clear_values_if_pending(ctx));  // (nothing was launched: no owned elements)
    ctx->pending_zero = false;
    return st;
}

fb200_status fb200_assemble_into_csr(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                     int32_t scatter_mode, int32_t accumulate, double* values) {
    if (!ctx) return FB200_ERR_STATE;
    if (!values) return fail(ctx, FB200_ERR_SHAPE, "null values");
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern: call fb200_assemble_pattern or fb200_pattern_adopt first");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (accumulate && ctx->nnz)
        FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_values, values, ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_TRY(fb200_assemble_into_csr_device(ctx, op, q, u, scatter_mode, accumulate));
    if (ctx->nnz) FB200_CUDA(ctx, cudaMemcpyAsync(values, ctx->d_values, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_errword(ctx);
}

// CompactQuadratureTable / GeneralQuadratureTable (quadrature_table.rs:57-210, 312-439): a rule per element.  Elements are grouped by
// rule on the host; every group runs the uniform-table element kernel over its own element list with accumulate semantics.
fb200_status fb200_assemble_into_csr_table_device(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                                  const uint32_t* element_rule, const double* u, int32_t scatter_mode, int32_t accumulate) {
    if (!ctx) return FB200_ERR_STATE;
    if (!op || !rules || !element_rule || num_rules == 0) return fail(ctx, FB200_ERR_SHAPE, "null operator / rules / element map");
    // state-dependent operators (StVK, NeoHookean) run the kernel of mass_source.cu per group; the linear ones ignore u
    const bool nonlinear = op->kind == FB200_STVK || op->kind == FB200_NEO_HOOKEAN;
    for (uint32_t r = 0; r < num_rules && !nonlinear; ++r) FB200_TRY(validate(ctx, op, &rules[r]));
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern: call fb200_assemble_pattern or fb200_pattern_adopt first");
    const int s = op->kind == FB200_LAPLACE ? 1 : ctx->ei.d;
    if (s != ctx->sdim) return fail(ctx, FB200_ERR_SHAPE, "pattern solution_dim does not match the operator");
    if (scatter_mode != FB200_SCATTER_ATOMIC && scatter_mode != FB200_SCATTER_COLORED)
        return fail(ctx, FB200_ERR_UNSUPPORTED, "quadrature tables with a rule per element support the ATOMIC and COLORED scatter");
    if (scatter_mode == FB200_SCATTER_COLORED && !ctx->has_colors)
        return fail(ctx, FB200_ERR_STATE, "coloured scatter needs fb200_color_nodes or fb200_colors_adopt");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool colored = scatter_mode == FB200_SCATTER_COLORED;
    std::vector<uint64_t> col_off, off;
    std::vector<int32_t> flat;
    FB200_TRY(group_elements_by_rule(ctx, num_rules, element_rule, colored, col_off, flat, off));
    int32_t* d_lists = nullptr;
    FB200_TRY(upload_vec(ctx, &d_lists, flat));
    fb200_status st = FB200_OK;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(ctx->d_values, 0, ctx->nnz * sizeof(double), ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "memset values");
    }
    std::vector<double> zeros;
    const double* state = u;  // uploaded with the first group, NULL afterwards
    if (nonlinear && !state) {
        zeros.assign((size_t)ctx->sdim * ctx->N + 1, 0.0);
        state = zeros.data();
    }
    for (uint32_t r = 0; r < num_rules && st == FB200_OK; ++r) {
        bool any = false;
        for (size_t c = 0; c + 1 < col_off.size(); ++c) any |= off[c * num_rules + r + 1] > off[c * num_rules + r];
        if (!any) continue;
        if (!nonlinear) st = upload_tables(ctx, op, &rules[r]);
        for (size_t c = 0; c + 1 < col_off.size() && st == FB200_OK; ++c) {
            const size_t i = c * num_rules + r;
            if (off[i + 1] == off[i]) continue;
            if (nonlinear) {
                st = assemble_state_dependent_list(ctx, op, &rules[r], state, d_lists + off[i], off[i + 1] - off[i], colored ? 1 : 0);
                if (state) cudaStreamSynchronize(ctx->stream);  // the copy of the state reads pageable memory
                state = nullptr;
                continue;
            }
            AssembleParams p;
            fill_params(ctx, p);
            p.accumulate = 1;
            p.generic_only = 1;
            p.elem_list = d_lists + off[i];
            p.count = off[i + 1] - off[i];
            // COLORED: one launch per (colour, rule), plain read-modify-write inside a colour
            st = dispatch(ctx, p, op->kind, colored ? (int)MODE_COLORED_LIST : (int)FB200_SCATTER_ATOMIC);
        }
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_lists);
    return st;
}

fb200_status fb200_values_device(fb200_ctx* ctx, double** device_ptr, uint64_t* nnz) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern");
    if (device_ptr) *device_ptr = ctx->d_values;
    if (nnz) *nnz = ctx->nnz;
    return FB200_OK;
}

fb200_status fb200_values_download(fb200_ctx* ctx, double* values) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nnz) FB200_CUDA(ctx, cudaMemcpyAsync(values, ctx->d_values, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_errword(ctx);
}

fb200_status fb200_values_upload(fb200_ctx* ctx, const double* values) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nnz) FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_values, values, ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FB200_OK;
}

fb200_status fb200_element_matrices(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, uint64_t first, uint64_t count,
                                    double* out) {
    return fb200_element_matrices_u(ctx, op, q, nullptr, first, count, out);
}

fb200_status fb200_element_matrices_u(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u, uint64_t first,
                                      uint64_t count, double* out) {
    const bool nonlinear = ctx && op && (op->kind == FB200_STVK || op->kind == FB200_NEO_HOOKEAN);
    if (!nonlinear) FB200_TRY(validate(ctx, op, q));
    if (!ctx) return FB200_ERR_STATE;
    if (nonlinear && (!ctx->has_space || ctx->ragged)) return fail(ctx, FB200_ERR_STATE, "element matrices need a space (fb200_space_upload)");
    if (first + count > ctx->E) return fail(ctx, FB200_ERR_SHAPE, "element range out of bounds");
    if (count == 0) return FB200_OK;
    if (!out) return fail(ctx, FB200_ERR_SHAPE, "null output");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const int s = op->kind == FB200_LAPLACE ? 1 : ctx->ei.d;
    const uint64_t per = (uint64_t)(s * ctx->ei.n) * (uint64_t)(s * ctx->ei.n);
    double* d_out = nullptr;
    int32_t* d_list = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_out, per * count));
    fb200_status st = FB200_OK;
    if (nonlinear) {
        // K_e depends on the state: the tangent of StVK / NeoHookean at u (materials.rs:232-469), dense from the pair kernel of mass_source.cu
        cudaError_t e = cudaMemsetAsync(d_out, 0, per * count * sizeof(double), ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "memset element matrices");
        if (st == FB200_OK) st = element_matrices_state_dependent(ctx, op, q, u, first, count, d_out);
    } else {
        st = upload_tables(ctx, op, q);
        if (st == FB200_OK) st = dev_alloc(ctx, &d_list, count);
        if (st == FB200_OK) {
            std::vector<int32_t> list(count);
            for (uint64_t k = 0; k < count; ++k) list[k] = (int32_t)(first + k);
            h2d_copy(ctx, d_list, list.data(), count * sizeof(int32_t));
            AssembleParams p;
            fill_params(ctx, p);
            p.elem_list = d_list;
            p.count = count;
            p.dump = d_out;
            p.dump_first = first;
            st = dispatch(ctx, p, op->kind, MODE_DUMP);
        }
    }
    if (st == FB200_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, per * count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H element matrices");
    }
    if (st == FB200_OK) st = read_errword(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (d_list) cudaFree(d_list);
    return st;
}

}  // extern "C"
