// SURVEY 8(f) rank 1 - the neighbours of the stiffness path that complete "assemble a linear system" (examples/poisson2d.rs:33-86):
//   element mass matrix        assemble_element_mass_matrix   src/assembly/local/mass.rs:218-286  (through the CSR scatter)
//   element source vector      assemble_element_source_vector src/assembly/local/source.rs:217-278
//   global vector assembly     VectorAssembler / VectorParAssembler::assemble_vector_into, add_local_to_global
//                              src/assembly/global.rs:569-686, 779-796
//   physical quadrature points FiniteElement::map_reference_coords (what SourceFunction::evaluate receives, source.rs:253-255)
// and SURVEY 8(f) rank 4 - the state-dependent side of the elliptic assembler:
//   element vector / energy    assemble_element_elliptic_vector / _energy   src/assembly/local/elliptic.rs:440-605  (Laplace, linear elastic,
//                              StVK, NeoHookean; compute_volume_u_grad :25-59)
//   tangent stiffness at u     assemble_element_elliptic_matrix with u_grad  elliptic.rs:361-439 for StVKMaterial / NeoHookeanMaterial
//                              (fenris-solid/src/materials.rs:232-469), one lane per node pair I <= J
//   the same with a quadrature rule per element (quadrature_table.rs:57-210, 312-439): one launch per (colour, rule) element list
// Same skeleton as the stiffness kernels (assemble.cu): tables of the uniform quadrature rule staged in shared memory, one warp
// per element, geometry per quadrature point per lane, scatter through the node-block map (matrix) or by node id (vector) with
// f64 reductions (ATOMIC) or plain read-modify-write inside a colour (COLORED = CsrParAssembler / VectorParAssembler semantics).
// These kernels are not tuned like the Hex8 tile kernel; they are HBM/atomic bound at a few percent of the stiffness cost.
#include <algorithm>
#include <cstring>

#include "fb200_internal.h"

namespace fb200 {

struct MsParams {
    const double* vertices;
    const int32_t* conn;
    const int64_t* blk_off;
    const uint16_t* blockmap;
    const int32_t* elem_list;  // colour list or nullptr
    uint64_t count;
    const double* tab;         // w[nq] | rho[nq] | ggeo[nq*ng*d] | phigeo[nq*ng] | phi[nq*n] | mu[nq] | lam[nq] | gref[nq*n*d]
    int nq, n, ng, d, s;
    int plain;                 // COLORED: plain read-modify-write
    double* values;            // CSR values (mass)
    double* vector;            // global vector (source)
    const double* source;      // [count_all or 1][nq][s]
    int source_per_element;
    double* points_out;        // physical points [E][nq][d]
    // elliptic vector / energy (elliptic.rs:440-605): operator kind, nodal solution u [N][s], per-element energies, error word
    int op;
    const double* u;
    double* energies;
    unsigned long long* errword;
    // host side only: launch over exactly this element list (quadrature tables with a rule per element), plain = coloured semantics
    const int32_t* forced_list;
    uint64_t forced_count;
    int forced;
    // WHAT 5 only: dense element matrices instead of the CSR scatter (fb200_element_matrices_u): element first_elem + k -> dump + k (s n)^2,
    // column-major per element; no pattern needed
    double* dump;
    uint64_t first_elem;
};

template <int d>
__device__ __forceinline__ double det_small_dev(const double (&J)[d * d]) {
    if constexpr (d == 2) return J[0] * J[3] - J[2] * J[1];
    // first-row cofactor expansion, as nalgebra's determinant()
    const double c00 = J[4] * J[8] - J[7] * J[5], c01 = J[3] * J[8] - J[6] * J[5], c02 = J[3] * J[7] - J[6] * J[4];
    return J[0] * c00 - J[1] * c01 + J[2] * c02;
}

// inverse from the cofactors and the determinant (nalgebra try_inverse, 2 x 2 and 3 x 3 closed forms)
template <int d>
__device__ __forceinline__ void inverse_small_dev(const double (&J)[d * d], double det, double (&Ji)[d * d]) {
    if constexpr (d == 2) {
        Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
    } else {
        Ji[0] = (J[4] * J[8] - J[7] * J[5]) / det; Ji[1] = (J[2] * J[7] - J[8] * J[1]) / det; Ji[2] = (J[1] * J[5] - J[4] * J[2]) / det;
        Ji[3] = -(J[3] * J[8] - J[6] * J[5]) / det; Ji[4] = (J[0] * J[8] - J[6] * J[2]) / det; Ji[5] = (J[2] * J[3] - J[5] * J[0]) / det;
        Ji[6] = (J[3] * J[7] - J[6] * J[4]) / det; Ji[7] = (J[1] * J[6] - J[7] * J[0]) / det; Ji[8] = (J[0] * J[4] - J[3] * J[1]) / det;
    }
}

// NeoHookeanMaterial (fenris-solid/src/materials.rs:267-318): F^-T and alpha = -mu + lambda log J; NaN for J <= 0 as the reference returns
template <int d>
__device__ __forceinline__ double neo_hookean_state(const double (&F)[d * d], double mu, double lam, double (&FinvT)[d * d]) {
    const double Jd = det_small_dev<d>(F);
    if (!(Jd > 0.0)) {
#pragma unroll
        for (int i = 0; i < d * d; ++i) FinvT[i] = __longlong_as_double(0x7ff8000000000000ll);
        return __longlong_as_double(0x7ff8000000000000ll);
    }
    double Fi[d * d];
    inverse_small_dev<d>(F, Jd, Fi);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) FinvT[i * d + j] = Fi[j * d + i];
    return -mu + lam * log(Jd);
}

// WHAT: 0 mass matrix, 1 source vector, 2 physical points, 3 elliptic vector, 4 elliptic energy, 5 state-dependent elliptic matrix (StVK, NeoHookean)
template <int WHAT, int n, int ng, int d>
__global__ void __launch_bounds__(128) mass_source_kernel(const MsParams p) {
    extern __shared__ double sm[];
    const int nq = p.nq, s = p.s;
    constexpr bool kElliptic = WHAT >= 3;
    const int tab_len = nq * (2 + ng * d + ng + n) + (kElliptic ? nq * (2 + n * d) : 0);
    for (int i = threadIdx.x; i < tab_len; i += blockDim.x) sm[i] = p.tab[i];
    const double* t_w = sm;
    const double* t_rho = t_w + nq;
    const double* t_ggeo = t_rho + nq;
    const double* t_pgeo = t_ggeo + nq * ng * d;
    const double* t_phi = t_pgeo + nq * ng;
    const double* t_mu = t_phi + nq * n;     // (elliptic only)
    const double* t_lam = t_mu + nq;
    const double* t_gref = t_lam + nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    // per warp: X | scale | base (int64) | ids + row lengths (int32) | elliptic: u_e[n][s], A[nq][s][d]
    constexpr int kRec = 4 * d * d + 4;  // WHAT 5, per point: J^-1 | F | E | F F^T | alpha, mu, lambda, tr E
    constexpr int kPairs = n * (n + 1) / 2;
    const int warp_doubles = ng * d + nq + 2 * n + (n & 1) + (kElliptic ? n * s + nq * (WHAT == 5 ? kRec : s * d) : 0);
    // mass matrix of small elements: the products phi_I(q) phi_J(q) do not depend on the element - tabulated once per CTA as
    // pp[q][I n + J] (lane-contiguous: conflict-free), so that M_IJ = sum_q scale_q pp[q][IJ] costs one shared load per term
    constexpr bool kProducts = WHAT == 0 && n * n <= 128;
    const int pp_len = kProducts ? nq * n * n : (WHAT == 5 ? (kPairs + 1) / 2 : 0);
    double* t_pp = sm + ((tab_len + 1) & ~1);
    if constexpr (WHAT == 5) {  // node pairs I <= J of the upper block triangle (elliptic.rs:417-431), a | b << 8
        int* pairs = reinterpret_cast<int*>(t_pp);
        for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
            const int a = i / n, b = i - a * n;
            if (a <= b) pairs[a * n - a * (a - 1) / 2 + (b - a)] = a | (b << 8);
        }
    }
    if constexpr (kProducts) {
        __syncthreads();
        for (int i = threadIdx.x; i < pp_len; i += blockDim.x) {
            const int q = i / (n * n), t = i - q * (n * n);
            const int a = t / n, b = t - a * n;
            t_pp[i] = sm[nq * (2 + ng * d + ng) + q * n + (a < b ? a : b)] * sm[nq * (2 + ng * d + ng) + q * n + (a < b ? b : a)];
        }
    }
    double* w_X = sm + ((tab_len + 1) & ~1) + ((pp_len + 1) & ~1) + warp * warp_doubles;
    double* w_scale = w_X + ng * d;
    long long* w_base = reinterpret_cast<long long*>(w_scale + nq);
    int* w_ids = reinterpret_cast<int*>(w_base + n);
    int* w_len = w_ids + n;
    double* w_u = w_X + ng * d + nq + 2 * n + (n & 1);
    double* w_A = w_u + n * s;
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * warps + warp; k < p.count; k += (uint64_t)gridDim.x * warps) {
        const uint64_t e = p.elem_list ? (uint64_t)p.elem_list[k] : k + p.first_elem;
        for (int a = lane; a < n; a += 32) {
            const int id = p.conn[e * n + a];
            w_ids[a] = id;
            if ((WHAT == 0 || WHAT == 5) && !p.dump) {
                const long long o0 = p.blk_off[id], o1 = p.blk_off[id + 1];
                w_base[a] = (long long)(s * s) * o0;
                w_len[a] = (int)(o1 - o0) * s;
            }
        }
        __syncwarp();
        for (int t = lane; t < ng * d; t += 32) w_X[t] = p.vertices[(uint64_t)w_ids[t / d] * d + (t % d)];
        __syncwarp();
        if constexpr (WHAT == 2) {
            for (int t = lane; t < nq * d; t += 32) {
                const int q = t / d, c = t - q * d;
                double x = 0.0;
                for (int a = 0; a < ng; ++a) x = fma(t_pgeo[q * ng + a], w_X[a * d + c], x);
                p.points_out[(e * nq + q) * d + c] = x;
            }
        }
        if constexpr (kElliptic) {
            // ---- elliptic vector f_I = sum_q w |det J| g^T grad phi_I and energy sum_q w |det J| psi(grad u)  (elliptic.rs:456-605)
            for (int t = lane; t < n * s; t += 32) w_u[t] = p.u[(uint64_t)w_ids[t / s] * s + (t % s)];
            __syncwarp();
            double energy = 0.0;
            for (int q = lane; q < nq; q += 32) {
                double J[d * d], Ji[d * d];
#pragma unroll
                for (int i = 0; i < d * d; ++i) J[i] = 0.0;
#pragma unroll
                for (int a = 0; a < ng; ++a)
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) J[i * d + j] = fma(w_X[a * d + i], t_ggeo[(q * ng + a) * d + j], J[i * d + j]);
                const double det = det_small_dev<d>(J);
                if (det == 0.0) {  // "Singular element Jacobian encountered" (elliptic.rs:493-497)
                    atomicMin(p.errword, ((unsigned long long)e << 8) | (unsigned long long)FB200_ERR_SINGULAR_JACOBIAN);
                    for (int t = 0; t < (WHAT == 5 ? kRec : s * d); ++t) w_A[q * (WHAT == 5 ? kRec : s * d) + t] = 0.0;
                    continue;
                }
                inverse_small_dev<d>(J, det, Ji);
                // grad u = J^{-T} sum_I grad_ref phi_I (x) u_I   (d x s, compute_volume_u_grad, elliptic.rs:25-59)
                double H[d * 3], GU[d * 3];  // s <= 3
#pragma unroll
                for (int i = 0; i < d * 3; ++i) H[i] = 0.0;
                for (int a = 0; a < n; ++a)
#pragma unroll
                    for (int k = 0; k < d; ++k)
                        for (int i = 0; i < s; ++i) H[k * 3 + i] = fma(t_gref[(q * n + a) * d + k], w_u[a * s + i], H[k * 3 + i]);
#pragma unroll
                for (int k = 0; k < d; ++k)
                    for (int i = 0; i < s; ++i) {
                        double acc = 0.0;
#pragma unroll
                        for (int m = 0; m < d; ++m) acc = fma(Ji[m * d + k], H[m * 3 + i], acc);  // (J^{-T})_{km} = Ji[m][k]
                        GU[k * 3 + i] = acc;
                    }
                if constexpr (WHAT == 5) {
                    // StVK state of the point (materials.rs:379-388): F = I + (grad u)^T, E = (F^T F - I) / 2, F F^T
                    double* R = w_A + q * kRec;
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            R[i * d + j] = Ji[i * d + j];
                            R[d * d + i * d + j] = (i == j ? 1.0 : 0.0) + GU[j * 3 + i];
                        }
                    R[4 * d * d] = t_w[q] * fabs(det);
                    R[4 * d * d + 1] = t_mu[q];
                    R[4 * d * d + 2] = t_lam[q];
                    if (p.op == FB200_NEO_HOOKEAN) {  // F^-T replaces E, alpha_nh replaces tr E
                        double F[d * d], T[d * d];
#pragma unroll
                        for (int i = 0; i < d * d; ++i) F[i] = R[d * d + i];
                        R[4 * d * d + 3] = neo_hookean_state<d>(F, t_mu[q], t_lam[q], T);
#pragma unroll
                        for (int i = 0; i < d * d; ++i) R[2 * d * d + i] = T[i];
                        continue;
                    }
                    double trE = 0.0;
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            double c = 0.0, b = 0.0;
#pragma unroll
                            for (int m = 0; m < d; ++m) {
                                c = fma(R[d * d + m * d + i], R[d * d + m * d + j], c);
                                b = fma(R[d * d + i * d + m], R[d * d + j * d + m], b);
                            }
                            const double Eij = 0.5 * (c - (i == j ? 1.0 : 0.0));
                            R[2 * d * d + i * d + j] = Eij;
                            R[3 * d * d + i * d + j] = b;
                            if (i == j) trE += Eij;
                        }
                    R[4 * d * d + 3] = trE;
                    continue;
                }
                // g^T (s x d) and psi
                double GT[3 * d], psi;
                if (p.op == FB200_LAPLACE) {  // g = grad u, psi = |grad u|^2 / 2  (operators/laplace.rs:33-51)
                    psi = 0.0;
#pragma unroll
                    for (int k = 0; k < d; ++k) {
                        GT[k] = GU[k * 3];
                        psi = fma(GU[k * 3], GU[k * 3], psi);
                    }
                    psi *= 0.5;
                } else if (p.op == FB200_STVK) {
                    // StVKMaterial (materials.rs:400-415): psi = mu E:E + lambda tr(E)^2 / 2, P = F (2 mu E + lambda tr(E) I)
                    double F[d * d], E[d * d];
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) F[i * d + j] = (i == j ? 1.0 : 0.0) + GU[j * 3 + i];
                    double tr = 0.0, ee = 0.0;
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            double c = 0.0;
#pragma unroll
                            for (int m = 0; m < d; ++m) c = fma(F[m * d + i], F[m * d + j], c);
                            E[i * d + j] = 0.5 * (c - (i == j ? 1.0 : 0.0));
                            ee = fma(E[i * d + j], E[i * d + j], ee);
                            if (i == j) tr += E[i * d + j];
                        }
                    const double mu = t_mu[q], lam = t_lam[q];
                    psi = mu * ee + 0.5 * lam * tr * tr;
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            double c = 0.0;
#pragma unroll
                            for (int m = 0; m < d; ++m) c = fma(F[i * d + m], E[m * d + j], c);
                            GT[i * d + j] = 2.0 * mu * c + lam * tr * F[i * d + j];
                        }
                } else if (p.op == FB200_NEO_HOOKEAN) {
                    // NeoHookeanMaterial: psi = mu tr(E) - mu log J + lambda (log J)^2 / 2 with log J = log1p(gamma) (materials.rs:251-265,
                    // logdet.rs:37-86; +inf when det F <= 0), P = F^-T (-mu + lambda log J) + mu F (materials.rs:267-289)
                    double F[d * d], T[d * d], U[d * d];
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            U[i * d + j] = GU[j * 3 + i];
                            F[i * d + j] = (i == j ? 1.0 : 0.0) + GU[j * 3 + i];
                        }
                    double gamma, trU = 0.0, uu = 0.0;
                    if constexpr (d == 2) {
                        gamma = U[0] * U[3] + U[0] + U[3] - U[1] * U[2];
                    } else {
                        const double u11 = U[0], u22 = U[4], u33 = U[8], a = 1.0 + u11, e = 1.0 + u22, ii = 1.0 + u33;
                        gamma = u11 * u22 * u33 + u11 * u22 + u11 * u33 + u22 * u33 + u11 + u22 + u33 + U[1] * U[5] * U[6] + U[2] * U[3] * U[7] -
                                U[2] * e * U[6] - U[1] * U[3] * ii - a * U[5] * U[7];
                    }
#pragma unroll
                    for (int i = 0; i < d; ++i) trU += U[i * d + i];
#pragma unroll
                    for (int i = 0; i < d * d; ++i) uu = fma(U[i], U[i], uu);
                    const double mu = t_mu[q], lam = t_lam[q];
                    if (gamma > -1.0) {
                        const double logJ = log1p(gamma);
                        psi = mu * (trU + 0.5 * uu) - mu * logJ + 0.5 * lam * logJ * logJ;
                    } else {
                        psi = __longlong_as_double(0x7ff0000000000000ll);
                    }
                    const double anh = neo_hookean_state<d>(F, mu, lam, T);
#pragma unroll
                    for (int i = 0; i < d * d; ++i) GT[i] = T[i] * anh + F[i] * mu;
                } else {
                    // LinearElasticMaterial through F = I + (grad u)^T, eps = sym(F) - I  (fenris-solid lib.rs:20-29, materials.rs:72-95)
                    double eps[d * d];
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            const double Fij = (i == j ? 1.0 : 0.0) + GU[j * 3 + i], Fji = (i == j ? 1.0 : 0.0) + GU[i * 3 + j];
                            eps[i * d + j] = 0.5 * (Fij + Fji) - (i == j ? 1.0 : 0.0);
                        }
                    double tr = 0.0, ee = 0.0;
#pragma unroll
                    for (int i = 0; i < d; ++i) tr += eps[i * d + i];
#pragma unroll
                    for (int i = 0; i < d * d; ++i) ee = fma(eps[i], eps[i], ee);
                    const double mu = t_mu[q], lam = t_lam[q];
                    psi = mu * ee + 0.5 * lam * tr * tr;
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) GT[i * d + j] = eps[i * d + j] * 2.0 * mu + (i == j ? lam * tr : 0.0);  // P = g^T
                }
                const double alpha = t_w[q] * fabs(det);
                energy = fma(alpha, psi, energy);
                // A_q = alpha g^T J^{-T}  (s x d): f_I += A_q grad_ref phi_I   (elliptic.rs:520-524)
                for (int i = 0; i < s; ++i)
#pragma unroll
                    for (int k = 0; k < d; ++k) {
                        double acc = 0.0;
#pragma unroll
                        for (int m = 0; m < d; ++m) acc = fma(GT[i * d + m], Ji[k * d + m], acc);  // (J^{-T})_{mk} = Ji[k][m]
                        w_A[(q * s + i) * d + k] = alpha * acc;
                    }
            }
            __syncwarp();
            if constexpr (WHAT == 3) {
                for (int t = lane; t < n * s; t += 32) {
                    const int a = t / s, i = t - a * s;
                    double f = 0.0;
                    for (int q = 0; q < nq; ++q)
#pragma unroll
                        for (int k = 0; k < d; ++k) f = fma(w_A[(q * s + i) * d + k], t_gref[(q * n + a) * d + k], f);
                    double* dst = p.vector + (uint64_t)w_ids[a] * s + i;
                    if (p.plain) *dst += f;
                    else atomicAdd(dst, f);
                }
            } else if constexpr (WHAT == 4) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, o);
                if (lane == 0) p.energies[k] = energy;
            } else {
                // K_IJ = sum_q alpha C(F_q; grad phi_I, grad phi_J) for I <= J (operators.rs:176-188 with the contraction of
                // materials.rs:417-437), block (J, I) = K_IJ^T (clone_upper_to_lower, util.rs:38-50); one node pair per lane
                const int* pairs = reinterpret_cast<const int*>(t_pp);
                for (int t = lane; t < kPairs; t += 32) {
                    const int a = pairs[t] & 0xff, b = pairs[t] >> 8;
                    double C[d * d];
#pragma unroll
                    for (int i = 0; i < d * d; ++i) C[i] = 0.0;
                    for (int q = 0; q < nq; ++q) {
                        const double* R = w_A + q * kRec;
                        double ga[d], gb[d], Fa[d], Fb[d], Eb[d];
#pragma unroll
                        for (int c = 0; c < d; ++c) {
                            double xa = 0.0, xb = 0.0;
#pragma unroll
                            for (int m = 0; m < d; ++m) {  // grad phi = J^-T grad_ref phi
                                xa = fma(R[m * d + c], t_gref[(q * n + a) * d + m], xa);
                                xb = fma(R[m * d + c], t_gref[(q * n + b) * d + m], xb);
                            }
                            ga[c] = xa;
                            gb[c] = xb;
                        }
                        const double alpha = R[4 * d * d], mu = R[4 * d * d + 1], lam = R[4 * d * d + 2], trE = R[4 * d * d + 3];
                        if (p.op == FB200_NEO_HOOKEAN) {
                            // C = lambda (F^-T a)(F^-T b)^T - alpha_nh (F^-T b)(F^-T a)^T + mu (a.b) I   (materials.rs:291-318)
                            double Ta[d], Tb[d], dot = 0.0;
#pragma unroll
                            for (int i = 0; i < d; ++i) {
                                double ta = 0.0, tb = 0.0;
#pragma unroll
                                for (int m = 0; m < d; ++m) {
                                    ta = fma(R[2 * d * d + i * d + m], ga[m], ta);
                                    tb = fma(R[2 * d * d + i * d + m], gb[m], tb);
                                }
                                Ta[i] = ta;
                                Tb[i] = tb;
                                dot = fma(ga[i], gb[i], dot);
                            }
#pragma unroll
                            for (int i = 0; i < d; ++i)
#pragma unroll
                                for (int j = 0; j < d; ++j) {
                                    const double c = lam * Ta[i] * Tb[j] - trE * Tb[i] * Ta[j] + (i == j ? mu * dot : 0.0);
                                    C[i * d + j] = fma(alpha, c, C[i * d + j]);
                                }
                            continue;
                        }
                        double ab = 0.0, aEb = 0.0;
#pragma unroll
                        for (int i = 0; i < d; ++i) {
                            double fa = 0.0, fb = 0.0, eb = 0.0;
#pragma unroll
                            for (int m = 0; m < d; ++m) {
                                fa = fma(R[d * d + i * d + m], ga[m], fa);
                                fb = fma(R[d * d + i * d + m], gb[m], fb);
                                eb = fma(R[2 * d * d + i * d + m], gb[m], eb);
                            }
                            Fa[i] = fa;
                            Fb[i] = fb;
                            Eb[i] = eb;
                            ab = fma(ga[i], gb[i], ab);
                        }
#pragma unroll
                        for (int i = 0; i < d; ++i) aEb = fma(ga[i], Eb[i], aEb);
                        const double diag = 2.0 * mu * aEb + lam * trE * ab;
#pragma unroll
                        for (int i = 0; i < d; ++i)
#pragma unroll
                            for (int j = 0; j < d; ++j) {
                                const double c = (i == j ? diag : 0.0) + mu * Fb[i] * Fa[j] + lam * Fa[i] * Fb[j] + mu * ab * R[3 * d * d + i * d + j];
                                C[i * d + j] = fma(alpha, c, C[i * d + j]);
                            }
                    }
                    if (a == b) {  // diagonal block: the scalar upper triangle mirrored (clone_upper_to_lower, util.rs:38-50)
#pragma unroll
                        for (int i = 0; i < d; ++i)
#pragma unroll
                            for (int j = 0; j < i; ++j) C[i * d + j] = C[j * d + i];
                    }
                    if (p.dump) {  // dense K_e, column-major: entry (d a + i, d b + j) and its mirror image
                        double* out = p.dump + k * (uint64_t)(n * d) * (uint64_t)(n * d);
#pragma unroll
                        for (int i = 0; i < d; ++i)
#pragma unroll
                            for (int j = 0; j < d; ++j) {
                                out[(uint64_t)(d * b + j) * (n * d) + (d * a + i)] = C[i * d + j];
                                if (a != b) out[(uint64_t)(d * a + i) * (n * d) + (d * b + j)] = C[i * d + j];
                            }
                        continue;
                    }
                    const int kab = p.blockmap[e * (uint64_t)(n * n) + a * n + b], kba = p.blockmap[e * (uint64_t)(n * n) + b * n + a];
                    double* rab = p.values + (w_base[a] + (long long)(d * kab));
                    double* rba = p.values + (w_base[b] + (long long)(d * kba));
#pragma unroll
                    for (int i = 0; i < d; ++i)
#pragma unroll
                        for (int j = 0; j < d; ++j) {
                            double* d0 = rab + (long long)i * w_len[a] + j;
                            double* d1 = rba + (long long)j * w_len[b] + i;
                            if (p.plain) {
                                *d0 += C[i * d + j];
                                if (a != b) *d1 += C[i * d + j];
                            } else {
                                atomicAdd(d0, C[i * d + j]);
                                if (a != b) atomicAdd(d1, C[i * d + j]);
                            }
                        }
                }
            }
            __syncwarp();
        }
        for (int q = lane; WHAT < 2 && q < nq; q += 32) {  // scale_q = w |det J_q| rho_q  (mass.rs:262-267, source.rs:254-267)
            double J[d * d];
#pragma unroll
            for (int i = 0; i < d * d; ++i) J[i] = 0.0;
#pragma unroll
            for (int a = 0; a < ng; ++a)
#pragma unroll
                for (int i = 0; i < d; ++i)
#pragma unroll
                    for (int j = 0; j < d; ++j) J[i * d + j] = fma(w_X[a * d + i], t_ggeo[(q * ng + a) * d + j], J[i * d + j]);
            w_scale[q] = t_w[q] * fabs(det_small_dev<d>(J)) * (WHAT == 0 ? t_rho[q] : 1.0);
        }
        __syncwarp();
        if constexpr (WHAT == 0) {
            // M_IJ = I_s sum_q scale_q phi_I phi_J; both triangles from the same (min, max)-ordered products: exactly symmetric,
            // like the reference's upper-triangle-then-mirror result (mass.rs:270-283)
            for (int t = lane; t < n * n; t += 32) {
                const int a = t / n, b = t - a * n;
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                double m = 0.0;
                if constexpr (kProducts) {
                    for (int q = 0; q < nq; ++q) m = fma(w_scale[q], t_pp[q * (n * n) + t], m);
                } else {
                    for (int q = 0; q < nq; ++q) m += w_scale[q] * t_phi[q * n + lo] * t_phi[q * n + hi];
                }
                const int kk = p.blockmap[e * (uint64_t)(n * n) + t];
                double* row = p.values + (w_base[a] + (long long)(s * kk));
                for (int i = 0; i < s; ++i) {
                    double* dst = row + (long long)i * w_len[a] + i;
                    if (p.plain) *dst += m;
                    else atomicAdd(dst, m);
                }
            }
        } else if constexpr (WHAT == 1) {
            // f_I = sum_q scale_q phi_I f(x_q)  (source.rs:268-276), added at s I + i (add_local_to_global, global.rs:779-796)
            const double* F = p.source + (p.source_per_element ? e * (uint64_t)(nq * s) : 0);
            for (int t = lane; t < n * s; t += 32) {
                const int a = t / s, i = t - a * s;
                double f = 0.0;
                for (int q = 0; q < nq; ++q) f += (w_scale[q] * F[q * s + i]) * t_phi[q * n + a];
                double* dst = p.vector + (uint64_t)w_ids[a] * s + i;
                if (p.plain) *dst += f;
                else atomicAdd(dst, f);
            }
        }
        __syncwarp();
    }
}

static fb200_status ms_validate(fb200_ctx* ctx, const fb200_quadrature* q) {
    if (!ctx) return FB200_ERR_STATE;
    if (!q) return fail(ctx, FB200_ERR_SHAPE, "null quadrature");
    if (!ctx->has_space) return fail(ctx, FB200_ERR_STATE, "needs fb200_space_upload (a finite element space)");
    if (q->dim != ctx->ei.d) return fail(ctx, FB200_ERR_SHAPE, "quadrature dimension != element reference dimension");
    if (q->num_points < 1 || q->num_points > 64) return fail(ctx, FB200_ERR_UNSUPPORTED, "1..64 quadrature points supported");
    if (!q->weights || !q->points) return fail(ctx, FB200_ERR_SHAPE, "null quadrature arrays");
    return FB200_OK;
}

static fb200_status ms_tables(fb200_ctx* ctx, const fb200_quadrature* q, int with_density /* 0 none, 1 density, 2 Lame pairs */) {
    const int nq = q->num_points, n = ctx->ei.n, ng = ctx->ei.ng, d = ctx->ei.d;
    std::vector<double> h((size_t)nq * (2 + ng * d + ng + n) + (size_t)nq * (2 + n * d), 0.0);
    double* w = h.data();
    double* rho = w + nq;
    double* ggeo = rho + nq;
    double* pgeo = ggeo + (size_t)nq * ng * d;
    double* phi = pgeo + (size_t)nq * ng;
    double* mu = phi + (size_t)nq * n;
    double* lam = mu + nq;
    double* gref = lam + nq;
    for (int k = 0; k < nq; ++k) {
        w[k] = q->weights[k];
        rho[k] = with_density == 1 ? q->data[k] : 1.0;
        if (with_density == 2) {  // Lame parameters per point (elliptic vector / energy of the linear elastic operator)
            mu[k] = q->data[2 * k];
            lam[k] = q->data[2 * k + 1];
        }
        reference_gradients(ctx->elem_type, q->points + (size_t)k * d, gref + (size_t)k * n * d);
        reference_gradients(geometry_type(ctx->elem_type), q->points + (size_t)k * d, ggeo + (size_t)k * ng * d);
        reference_basis(geometry_type(ctx->elem_type), q->points + (size_t)k * d, pgeo + (size_t)k * ng);
        reference_basis(ctx->elem_type, q->points + (size_t)k * d, phi + (size_t)k * n);
    }
    if (ctx->d_ms_tab && ctx->h_ms_tab == h) return FB200_OK;
    if (ctx->ms_tab_capacity < h.size()) {
        dev_free(ctx->d_ms_tab);
        FB200_TRY(dev_alloc(ctx, &ctx->d_ms_tab, h.size()));
        ctx->ms_tab_capacity = h.size();
    }
    FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_ms_tab, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // h is pageable
    ctx->h_ms_tab = h;
    return FB200_OK;
}

template <int WHAT, int N, int NG, int D>
static fb200_status ms_launch_t(fb200_ctx* ctx, MsParams& p, int scatter_mode) {
    p.vertices = ctx->d_vertices;
    p.conn = ctx->d_conn;
    p.blk_off = ctx->d_blk_off;
    p.blockmap = ctx->d_blockmap;
    p.tab = ctx->d_ms_tab;
    p.n = N;
    p.ng = NG;
    p.d = D;
    const int tab_len = p.nq * (2 + NG * D + NG + N) + (WHAT >= 3 ? p.nq * (2 + N * D) : 0);
    const int warp_doubles = NG * D + p.nq + 2 * N + (N & 1) + (WHAT >= 3 ? N * p.s + p.nq * (WHAT == 5 ? 4 * D * D + 4 : p.s * D) : 0);
    // products phi_I phi_J per point (mass) / node pair list (state-dependent matrix): see the kernel
    const int pp_len = (WHAT == 0 && N * N <= 128) ? p.nq * N * N : (WHAT == 5 ? (N * (N + 1) / 2 + 1) / 2 : 0);
    const size_t smem = sizeof(double) * (size_t)(((tab_len + 1) & ~1) + ((pp_len + 1) & ~1) + 4 * warp_doubles);
    auto kernel = mass_source_kernel<WHAT, N, NG, D>;
    if (smem > 48 * 1024) FB200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto run = [&](const int32_t* list, uint64_t count) -> fb200_status {
        if (count == 0) return FB200_OK;
        p.elem_list = list;
        p.count = count;
        const int blocks = (int)std::min<uint64_t>(div_up(count, 4), (uint64_t)ctx->sm_count * 8);
        kernel<<<blocks, 128, smem, ctx->stream>>>(p);
        return check_launch(ctx, "mass_source_kernel");
    };
    if (p.forced) return run(p.forced_list, p.forced_count);
    if (WHAT == 4) {  // energies: every owned element once, no scatter
        p.plain = 0;
        return run(nullptr, ctx->E_owned);
    }
    if (WHAT != 2 && scatter_mode == FB200_SCATTER_COLORED) {
        if (!ctx->has_colors) return fail(ctx, FB200_ERR_STATE, "coloured scatter needs fb200_color_nodes or fb200_colors_adopt");
        p.plain = 1;
        for (size_t c = 0; c + 1 < ctx->h_color_off.size(); ++c)
            FB200_TRY(run(ctx->d_color_elems + ctx->h_color_off[c], ctx->h_color_off[c + 1] - ctx->h_color_off[c]));
        return FB200_OK;
    }
    if (WHAT != 2 && scatter_mode != FB200_SCATTER_ATOMIC)
        return fail(ctx, FB200_ERR_UNSUPPORTED, "mass matrix / vector assembly support the ATOMIC and COLORED scatter");
    p.plain = 0;
    if (WHAT == 2) return run(nullptr, ctx->E);
    // visit the elements in the Morton order of the space (as the stiffness path): all contributions to a CSR row / vector
    // entry arrive while its lines are still in L2
    const bool ordered = ctx->d_order && ctx->order_count == ctx->E_owned;
    return run(ordered ? ctx->d_order : nullptr, ctx->E_owned);
}

template <int WHAT>
static fb200_status ms_launch(fb200_ctx* ctx, MsParams& p, int scatter_mode) {
    switch (ctx->elem_type) {
        case FB200_QUAD4: return ms_launch_t<WHAT, 4, 4, 2>(ctx, p, scatter_mode);
        case FB200_TET4: return ms_launch_t<WHAT, 4, 4, 3>(ctx, p, scatter_mode);
        case FB200_HEX8: return ms_launch_t<WHAT, 8, 8, 3>(ctx, p, scatter_mode);
        case FB200_HEX27: return ms_launch_t<WHAT, 27, 8, 3>(ctx, p, scatter_mode);
        case FB200_TET10: return ms_launch_t<WHAT, 10, 4, 3>(ctx, p, scatter_mode);
        case FB200_HEX20: return ms_launch_t<WHAT, 20, 8, 3>(ctx, p, scatter_mode);
        default: return fail(ctx, FB200_ERR_UNSUPPORTED, "element type has no device specialisation (no CPU fallback)");
    }
}

}  // namespace fb200

using namespace fb200;

extern "C" {

fb200_status fb200_assemble_mass_into_csr_device(fb200_ctx* ctx, const fb200_quadrature* q, int32_t scatter_mode, int32_t accumulate) {
    FB200_TRY(ms_validate(ctx, q));
    if (!q->data) return fail(ctx, FB200_ERR_SHAPE, "the mass matrix needs a density per quadrature point (Density, mass.rs:23-31)");
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern: call fb200_assemble_pattern or fb200_pattern_adopt first");
    if (ctx->ragged || !ctx->d_blockmap) return fail(ctx, FB200_ERR_UNSUPPORTED, "mass assembly needs a uniform-element space");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    FB200_TRY(ms_tables(ctx, q, 1));
    if (!accumulate) FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_values, 0, ctx->nnz * sizeof(double), ctx->stream));
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = ctx->sdim;
    p.values = ctx->d_values;
    return ms_launch<0>(ctx, p, scatter_mode);
}

fb200_status fb200_assemble_mass_into_csr(fb200_ctx* ctx, const fb200_quadrature* q, int32_t scatter_mode, int32_t accumulate, double* values) {
    if (!ctx) return FB200_ERR_STATE;
    if (!values) return fail(ctx, FB200_ERR_SHAPE, "null values");
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern: call fb200_assemble_pattern or fb200_pattern_adopt first");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (accumulate && ctx->nnz)
        FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_values, values, ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_TRY(fb200_assemble_mass_into_csr_device(ctx, q, scatter_mode, accumulate));
    if (ctx->nnz) FB200_CUDA(ctx, cudaMemcpyAsync(values, ctx->d_values, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_errword(ctx);
}

fb200_status fb200_assemble_vector(fb200_ctx* ctx, const fb200_quadrature* q, int32_t solution_dim, const double* source_values,
                                   int32_t per_element, int32_t scatter_mode, int32_t accumulate, double* out) {
    FB200_TRY(ms_validate(ctx, q));
    if (solution_dim < 1 || solution_dim > 3) return fail(ctx, FB200_ERR_SHAPE, "solution_dim must be 1..3");
    if (!source_values || !out) return fail(ctx, FB200_ERR_SHAPE, "null source values / output");
    if (ctx->ragged) return fail(ctx, FB200_ERR_UNSUPPORTED, "vector assembly needs a uniform-element space");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    FB200_TRY(ms_tables(ctx, q, 0));
    const uint64_t len = (uint64_t)solution_dim * ctx->N;
    const uint64_t src_len = (uint64_t)q->num_points * solution_dim * (per_element ? ctx->E : 1);
    if (ctx->vector_capacity < len) {
        dev_free(ctx->d_vector);
        FB200_TRY(dev_alloc(ctx, &ctx->d_vector, len));
        ctx->vector_capacity = len;
    }
    if (ctx->source_capacity < src_len) {
        dev_free(ctx->d_source);
        FB200_TRY(dev_alloc(ctx, &ctx->d_source, src_len));
        ctx->source_capacity = src_len;
    }
    ctx->vector_len = len;
    if (accumulate) {
        if (len) FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_vector, out, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else if (len) {
        FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_vector, 0, len * sizeof(double), ctx->stream));
    }
    if (src_len) FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_source, source_values, src_len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = solution_dim;
    p.vector = ctx->d_vector;
    p.source = ctx->d_source;
    p.source_per_element = per_element ? 1 : 0;
    FB200_TRY(ms_launch<1>(ctx, p, scatter_mode));
    if (len) FB200_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_vector, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_errword(ctx);
}

// ElementEllipticAssembler as ElementVectorAssembler / ElementScalarAssembler (elliptic.rs:342-359, 440-605) through VectorAssembler
// (global.rs:569-686) / assemble_scalar (global.rs:697-722)
static fb200_status elliptic_common(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u, int* s_out) {
    FB200_TRY(ms_validate(ctx, q));
    if (!op || (op->kind != FB200_LAPLACE && op->kind != FB200_LINEAR_ELASTIC && op->kind != FB200_STVK && op->kind != FB200_NEO_HOOKEAN))
        return fail(ctx, FB200_ERR_UNSUPPORTED, "operator has no device specialisation (no CPU fallback)");
    if (op->kind != FB200_LAPLACE && !q->data) return fail(ctx, FB200_ERR_SHAPE, "elastic materials need Lame data per point");
    if (!u) return fail(ctx, FB200_ERR_SHAPE, "null u");
    if (ctx->ragged) return fail(ctx, FB200_ERR_UNSUPPORTED, "needs a uniform-element space");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    FB200_TRY(ms_tables(ctx, q, op->kind != FB200_LAPLACE ? 2 : 0));
    const int s = op->kind == FB200_LAPLACE ? 1 : ctx->ei.d;
    const uint64_t len = (uint64_t)s * ctx->N;
    if (ctx->source_capacity < len) {
        dev_free(ctx->d_source);
        FB200_TRY(dev_alloc(ctx, &ctx->d_source, len));
        ctx->source_capacity = len;
    }
    if (len) FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_source, u, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    *s_out = s;
    return FB200_OK;
}

fb200_status fb200_assemble_elliptic_vector(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                            int32_t scatter_mode, int32_t accumulate, double* out) {
    if (!ctx) return FB200_ERR_STATE;
    if (!out) return fail(ctx, FB200_ERR_SHAPE, "null output");
    int s = 0;
    FB200_TRY(elliptic_common(ctx, op, q, u, &s));
    const uint64_t len = (uint64_t)s * ctx->N;
    if (ctx->vector_capacity < len) {
        dev_free(ctx->d_vector);
        FB200_TRY(dev_alloc(ctx, &ctx->d_vector, len));
        ctx->vector_capacity = len;
    }
    ctx->vector_len = len;
    if (accumulate) {
        if (len) FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_vector, out, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else if (len) {
        FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_vector, 0, len * sizeof(double), ctx->stream));
    }
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = s;
    p.op = op->kind;
    p.u = ctx->d_source;
    p.vector = ctx->d_vector;
    p.errword = ctx->d_errword;
    FB200_TRY(ms_launch<3>(ctx, p, scatter_mode));
    if (len) FB200_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_vector, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_errword(ctx);
}

__global__ void sum_energies_kernel(const double* __restrict__ e, uint64_t n, double* out) {  // one block, fixed order
    __shared__ double sh[256];
    double s = 0.0;
    for (uint64_t i = threadIdx.x; i < n; i += 256) s += e[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

fb200_status fb200_assemble_elliptic_scalar(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u, double* energy) {
    if (!ctx) return FB200_ERR_STATE;
    if (!energy) return fail(ctx, FB200_ERR_SHAPE, "null output");
    int s = 0;
    FB200_TRY(elliptic_common(ctx, op, q, u, &s));
    *energy = 0.0;
    if (ctx->E_owned == 0) return FB200_OK;
    double* d_e = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_e, ctx->E_owned + 1));
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = s;
    p.op = op->kind;
    p.u = ctx->d_source;
    p.energies = d_e;
    p.errword = ctx->d_errword;
    fb200_status st = ms_launch<4>(ctx, p, FB200_SCATTER_ATOMIC);
    if (st == FB200_OK) {
        sum_energies_kernel<<<1, 256, 0, ctx->stream>>>(d_e, ctx->E_owned, d_e + ctx->E_owned);
        st = check_launch(ctx, "sum_energies_kernel");
    }
    if (st == FB200_OK) {
        cudaError_t e = cudaMemcpyAsync(energy, d_e + ctx->E_owned, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H energy");
    }
    if (st == FB200_OK) st = read_errword(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_e);
    return st;
}

}  // extern "C"

// ElementEllipticAssembler as ElementMatrixAssembler with a state-dependent contraction (elliptic.rs:361-439 with u_grad, operators.rs:
// 176-188): the tangent stiffness of StVKMaterial at u.  Called by fb200_assemble_into_csr_device for FB200_STVK.
fb200_status fb200::assemble_state_dependent(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                             int scatter_mode, int accumulate) {
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern: call fb200_assemble_pattern or fb200_pattern_adopt first");
    if (ctx->ragged || !ctx->d_blockmap) return fail(ctx, FB200_ERR_UNSUPPORTED, "needs a uniform-element space");
    if (ctx->ei.d != ctx->sdim) return fail(ctx, FB200_ERR_SHAPE, "pattern solution_dim does not match the operator");
    std::vector<double> zeros;
    if (!u) {  // NULL = zeros (fb200_assemble_into_csr): the tangent at the undeformed state
        zeros.assign((size_t)ctx->sdim * ctx->N + 1, 0.0);
        u = zeros.data();
    }
    int s = 0;
    FB200_TRY(elliptic_common(ctx, op, q, u, &s));
    if (!zeros.empty()) FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the copy reads pageable memory that dies here
    if (!accumulate) FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_values, 0, ctx->nnz * sizeof(double), ctx->stream));
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = s;
    p.op = op->kind;
    p.u = ctx->d_source;
    p.values = ctx->d_values;
    p.errword = ctx->d_errword;
    return ms_launch<5>(ctx, p, scatter_mode);
}

// The ElementMatrixAssembler::assemble_element_matrix view (local.rs:78-80) of a state-dependent operator: dense K_e(u) of the elements
// [first, first + count), column-major per element, on the device buffer d_out.  u = NULL: the tangent at the undeformed state.
fb200_status fb200::element_matrices_state_dependent(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                                     uint64_t first, uint64_t count, double* d_out) {
    std::vector<double> zeros;
    if (!u) {
        zeros.assign((size_t)ctx->ei.d * ctx->N + 1, 0.0);
        u = zeros.data();
    }
    int s = 0;
    FB200_TRY(elliptic_common(ctx, op, q, u, &s));
    if (!zeros.empty()) FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = s;
    p.op = op->kind;
    p.u = ctx->d_source;
    p.errword = ctx->d_errword;
    p.dump = d_out;
    p.first_elem = first;
    p.forced = 1;
    p.forced_list = nullptr;
    p.forced_count = count;
    return ms_launch<5>(ctx, p, FB200_SCATTER_ATOMIC);
}

// One (colour, rule) group of fb200_assemble_into_csr_table_device for a state-dependent operator: u = NULL keeps the state uploaded by
// the previous call of the same assembly; always accumulates (the caller clears the values once).
fb200_status fb200::assemble_state_dependent_list(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                                  const int32_t* d_list, uint64_t count, int plain) {
    if (ctx->ragged || !ctx->d_blockmap) return fail(ctx, FB200_ERR_UNSUPPORTED, "needs a uniform-element space");
    if (ctx->ei.d != ctx->sdim) return fail(ctx, FB200_ERR_SHAPE, "pattern solution_dim does not match the operator");
    if (u) {
        int s = 0;
        FB200_TRY(elliptic_common(ctx, op, q, u, &s));
    } else {
        FB200_TRY(ms_validate(ctx, q));
        if (!q->data) return fail(ctx, FB200_ERR_SHAPE, "elastic materials need Lame data per point");
        FB200_TRY(ms_tables(ctx, q, 2));
    }
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = ctx->sdim;
    p.op = op->kind;
    p.u = ctx->d_source;
    p.values = ctx->d_values;
    p.errword = ctx->d_errword;
    p.forced = 1;
    p.forced_list = d_list;
    p.forced_count = count;
    p.plain = plain;
    return ms_launch<5>(ctx, p, FB200_SCATTER_ATOMIC);
}

extern "C" {

// The same two with a rule per element (CompactQuadratureTable / GeneralQuadratureTable, quadrature_table.rs:57-210, 312-439): elements
// grouped by rule (x colour) on the host, one launch per group over its element list.
static fb200_status elliptic_table_begin(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                         const uint32_t* element_rule, const double* u, int* s_out) {
    if (!rules || !element_rule || num_rules == 0) return fail(ctx, FB200_ERR_SHAPE, "null rules / element map");
    for (uint32_t r = 0; r < num_rules; ++r) {
        FB200_TRY(ms_validate(ctx, &rules[r]));
        if (op && op->kind != FB200_LAPLACE && !rules[r].data) return fail(ctx, FB200_ERR_SHAPE, "elastic materials need Lame data per point");
    }
    return elliptic_common(ctx, op, &rules[0], u, s_out);  // validates the operator, uploads u
}

fb200_status fb200_assemble_elliptic_vector_table(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                                  const uint32_t* element_rule, const double* u, int32_t scatter_mode, int32_t accumulate,
                                                  double* out) {
    if (!ctx) return FB200_ERR_STATE;
    if (!out) return fail(ctx, FB200_ERR_SHAPE, "null output");
    if (scatter_mode != FB200_SCATTER_ATOMIC && scatter_mode != FB200_SCATTER_COLORED)
        return fail(ctx, FB200_ERR_UNSUPPORTED, "vector assembly supports the ATOMIC and COLORED scatter");
    const bool colored = scatter_mode == FB200_SCATTER_COLORED;
    if (colored && !ctx->has_colors) return fail(ctx, FB200_ERR_STATE, "coloured scatter needs fb200_color_nodes or fb200_colors_adopt");
    int s = 0;
    FB200_TRY(elliptic_table_begin(ctx, op, num_rules, rules, element_rule, u, &s));
    std::vector<uint64_t> col_off, off;
    std::vector<int32_t> flat;
    FB200_TRY(group_elements_by_rule(ctx, num_rules, element_rule, colored, col_off, flat, off));
    const uint64_t len = (uint64_t)s * ctx->N;
    if (ctx->vector_capacity < len) {
        dev_free(ctx->d_vector);
        FB200_TRY(dev_alloc(ctx, &ctx->d_vector, len));
        ctx->vector_capacity = len;
    }
    ctx->vector_len = len;
    if (accumulate) {
        if (len) FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_vector, out, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else if (len) {
        FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_vector, 0, len * sizeof(double), ctx->stream));
    }
    int32_t* d_lists = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_lists, flat.size()));
    fb200_status st = FB200_OK;
    if (!flat.empty()) {
        cudaError_t e = cudaMemcpyAsync(d_lists, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "H2D element lists");
    }
    for (uint32_t r = 0; r < num_rules && st == FB200_OK; ++r)
        for (size_t c = 0; c + 1 < col_off.size() && st == FB200_OK; ++c) {
            const size_t i = c * num_rules + r;
            if (off[i + 1] == off[i]) continue;
            st = ms_tables(ctx, &rules[r], op->kind != FB200_LAPLACE ? 2 : 0);
            if (st != FB200_OK) break;
            MsParams p;
            std::memset(&p, 0, sizeof(p));
            p.nq = rules[r].num_points;
            p.s = s;
            p.op = op->kind;
            p.u = ctx->d_source;
            p.vector = ctx->d_vector;
            p.errword = ctx->d_errword;
            p.forced = 1;
            p.forced_list = d_lists + off[i];
            p.forced_count = off[i + 1] - off[i];
            p.plain = colored ? 1 : 0;
            st = ms_launch<3>(ctx, p, FB200_SCATTER_ATOMIC);
        }
    if (st == FB200_OK && len) {
        cudaError_t e = cudaMemcpyAsync(out, ctx->d_vector, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H vector");
    }
    if (st == FB200_OK) st = read_errword(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_lists);
    return st;
}

fb200_status fb200_assemble_elliptic_scalar_table(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                                  const uint32_t* element_rule, const double* u, double* energy) {
    if (!ctx) return FB200_ERR_STATE;
    if (!energy) return fail(ctx, FB200_ERR_SHAPE, "null output");
    int s = 0;
    FB200_TRY(elliptic_table_begin(ctx, op, num_rules, rules, element_rule, u, &s));
    *energy = 0.0;
    if (ctx->E_owned == 0) return FB200_OK;
    std::vector<uint64_t> col_off, off;
    std::vector<int32_t> flat;
    FB200_TRY(group_elements_by_rule(ctx, num_rules, element_rule, false, col_off, flat, off));
    int32_t* d_lists = nullptr;
    double* d_e = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_lists, flat.size()));
    fb200_status st = dev_alloc(ctx, &d_e, ctx->E_owned + 1);
    if (st == FB200_OK) {
        cudaError_t e = cudaMemcpyAsync(d_lists, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "H2D element lists");
    }
    for (uint32_t r = 0; r < num_rules && st == FB200_OK; ++r) {
        if (off[r + 1] == off[r]) continue;
        st = ms_tables(ctx, &rules[r], op->kind != FB200_LAPLACE ? 2 : 0);
        if (st != FB200_OK) break;
        MsParams p;
        std::memset(&p, 0, sizeof(p));
        p.nq = rules[r].num_points;
        p.s = s;
        p.op = op->kind;
        p.u = ctx->d_source;
        p.energies = d_e + off[r];  // the kernel writes energies[position in its list]
        p.errword = ctx->d_errword;
        p.forced = 1;
        p.forced_list = d_lists + off[r];
        p.forced_count = off[r + 1] - off[r];
        st = ms_launch<4>(ctx, p, FB200_SCATTER_ATOMIC);
    }
    if (st == FB200_OK) {
        sum_energies_kernel<<<1, 256, 0, ctx->stream>>>(d_e, ctx->E_owned, d_e + ctx->E_owned);
        st = check_launch(ctx, "sum_energies_kernel");
    }
    if (st == FB200_OK) {
        cudaError_t e = cudaMemcpyAsync(energy, d_e + ctx->E_owned, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H energy");
    }
    if (st == FB200_OK) st = read_errword(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_e);
    cudaFree(d_lists);
    return st;
}

fb200_status fb200_physical_quadrature_points(fb200_ctx* ctx, const fb200_quadrature* q, double* out) {
    FB200_TRY(ms_validate(ctx, q));
    if (!out) return fail(ctx, FB200_ERR_SHAPE, "null output");
    if (ctx->ragged) return fail(ctx, FB200_ERR_UNSUPPORTED, "needs a uniform-element space");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    FB200_TRY(ms_tables(ctx, q, 0));
    const uint64_t len = ctx->E * (uint64_t)q->num_points * ctx->ei.d;
    if (len == 0) return FB200_OK;
    double* d_out = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_out, len));
    MsParams p;
    std::memset(&p, 0, sizeof(p));
    p.nq = q->num_points;
    p.s = 1;
    p.points_out = d_out;
    fb200_status st = ms_launch<2>(ctx, p, FB200_SCATTER_ATOMIC);
    if (st == FB200_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H physical points");
    }
    if (st == FB200_OK) st = read_errword(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    return st;
}

}  // extern "C"
