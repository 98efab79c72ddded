// Hex8 warp-per-element assembly kernel, FP64 tensor-core (DMMA) formulation (included by assemble.cu).
//
// Same mathematics and the same load / stage / scatter skeleton as assemble_hex8_kernel (hex8_kernel.cuh), but the two
// phases that saturated the shared-memory pipe there are rebuilt:
//
//   geometry  lane = (q, i): 4 lanes per quadrature point q, lane i < 3 owns ROW i of the Jacobian
//             J[i][:] = sum_a X_a[i] grad_ref phi_a(q)  (element/hexahedron.rs:101-107).  The reference gradients of the
//             lane's point live in 24 REGISTERS for the whole kernel (the table is uniform over elements,
//             quadrature_table.rs:264-266), so the only shared-memory reads are the 8 broadcast loads of X_a[i].
//             The three rows are exchanged with 6 shuffles; lane i forms the cofactor row c_i = J[i+1] x J[i+2] -
//             which is column i of adj(J) - and det J by the reference's first-row expansion (lane 0's value is broadcast so
//             that "det == 0" is decided exactly as elliptic.rs:400-404 does).  With r = sqrt(w |det J|) / det J the lane then
//             pushes forward component i of ALL 8 nodes: g_a[i] = r (c_i . grad_ref phi_a) = sqrt(alpha) (J^{-T} grad_ref phi_a)_i
//             (elliptic.rs:415-422).
//   transpose the 8 x 3 x 8 scaled gradients go through shared memory once ([node][3 q + i], node stride 28: conflict free
//             for the writers and for the fragment loads).
//   blocks    S_ab = sum_q g_a(q) (x) g_b(q) is the product G G^T of the 24 x 8 matrix G[(i, a)][q].  With the dofs ordered
//             component-major, the 8x8 tile (m, n) of G G^T holds S_ab[m][n] for all 8 x 8 node pairs, and the
//             mma.sync.m8n8k4.f64 accumulator fragment gives lane (a = lane / 4, t = lane % 4) the entries of node pairs
//             (a, 2t) and (a, 2t + 1) of EVERY tile - i.e. two complete 3 x 3 blocks per lane.  The A fragment of tile row
//             m and the B fragment of tile column m are the same register (G[m][a][4 ks + t]), so the whole contraction is
//             6 shared loads + 18 DMMA per element (Laplace: the 3 diagonal tiles accumulated into one, 6 DMMA) instead
//             of 72 broadcast loads + 144 DFMA.
//   epilogue  K_ab = mu [tr(S_ab) I + S_ab^T] + lambda S_ab (fenris-solid/src/materials.rs:108-122), staged row-major and
//             scattered one K_e row per reduction instruction exactly as in hex8_kernel.cuh.
// K_e is exactly symmetric (util.rs:38-50 semantics): S_ab[m][n] and S_ba[n][m] are the same dot product of the same
// operands in the same order.  Requires uniform operator parameters, positive weights and at most 8 quadrature points
// (fewer points are padded with zero columns); everything else takes the DFMA kernels.
#pragma once

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int OP, int MODE, int THREADS, int MINB, bool DYN, bool HINT, int CHUNK = 8>
__global__ void __launch_bounds__(THREADS, MINB) assemble_hex8_mma_kernel(const AssembleParams p) {
    constexpr int N = 8, D = 3;
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    constexpr int SN = S * N;
    constexpr int GS = 28;                                      // node stride of the transposed gradient array
    constexpr int KLEN = S == 1 ? SN * (SN + 1) : SN * SN + 4;  // staged K_e
    constexpr int WARPS = THREADS / 32;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int nq = p.nq;  // <= 8
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARP_DOUBLES = N * D + N * GS + KLEN + (KLEN & 1) + 28;  // + 8 int64 + 8 int32 + 32 uint32
    double* s_X = smem + warp * WARP_DOUBLES;
    double* s_G = s_X + N * D;
    double* s_K = s_G + N * GS;
    long long* s_base = reinterpret_cast<long long*>(s_K + KLEN + (KLEN & 1));
    int* s_rowlen = reinterpret_cast<int*>(s_base + N);
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_rowlen + N);
    const uint16_t* s_pos16 = reinterpret_cast<const uint16_t*>(s_pos);

    // ---- per-lane constants
    const int gq = lane >> 2, s4 = lane & 3;                // geometry role: point gq, Jacobian row gi
    const int gi = s4 == 3 ? 0 : s4;                         // (the 4th lane of a point shadows row 0 and stores nothing)
    const bool gact = gq < nq;
    double R[N][D];                                          // reference gradients of point gq, node-major
    double sqw = 0.0;
    {
        const double* tab_g = p.tab + 3 * nq + (gact ? gq : 0) * (N * D);  // geometry table == basis table for Hex8
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
            for (int j = 0; j < D; ++j) R[a][j] = gact ? tab_g[a * D + j] : 0.0;
        if (gact) sqw = sqrt(p.tab[gq]);
    }
    const int src1 = (lane & ~3) | (gi == 2 ? 0 : gi + 1), src2 = (lane & ~3) | (gi == 0 ? 2 : gi - 1);  // rows i+1, i+2 (mod 3)
    const int ba = lane >> 2, b0 = 2 * (lane & 3);          // block role: blocks (ba, b0), (ba, b0 + 1); fragment row ba, k = lane & 3
    const int frag = ba * GS + 3 * (lane & 3);
    const int col_b = lane / S, col_j = lane - col_b * S;   // scatter role (lane < SN)
    const int xl = lane < N * D ? lane : 0;
    const int x_node = xl / D, x_comp = xl - x_node * D;
    const double mu = p.mu, lam = p.lam;
    constexpr bool hints = HINT;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    const int dbg = p.debug;  // measurement knobs (FB200_DEBUG; results are wrong with 1/2/4 set): 1 skip compute, 2 skip scatter, 4 st for red, 8 L2 prefetch

    // ---- dynamic tickets + two-deep register software pipeline: identical to assemble_hex8_kernel
    const uint32_t nw = (uint32_t)gridDim.x * WARPS;
    const uint32_t count = (uint32_t)p.count;
    constexpr bool dynamic = DYN;
    uint32_t static_pos = (uint32_t)blockIdx.x * WARPS + warp;
    uint32_t gen_base = 0;
    int gen_left = 0;
    unsigned int tick_next = 0;
    if (dynamic && lane == 0) tick_next = atomicAdd(p.ticket32, 1u);
    auto next_pos = [&]() -> uint32_t {
        if constexpr (!dynamic) {
            const uint32_t r = static_pos;
            static_pos = (r >= count) ? r : r + nw;
            return r;
        } else {
            if (gen_left == 0) {
                const unsigned int t = __shfl_sync(FULL, tick_next, 0);
                gen_base = t >= 0x3fffffffu ? 0xfffffff0u : t * CHUNK;
                gen_left = CHUNK;
                if (lane == 0) tick_next = atomicAdd(p.ticket32, 1u);  // consumed CHUNK iterations from now
            }
            const uint32_t r = gen_base + (uint32_t)(CHUNK - gen_left);
            --gen_left;
            return r;
        }
    };
    auto load_ids = [&](uint32_t pos, int& node, uint32_t& mapw) {
        node = 0;
        mapw = 0;
        if (pos < count) {
            if constexpr (hints) {
                if (lane < N) node = (int)ld_u32_hint(p.conn_pos + (uint64_t)pos * N + lane, pol_stream);
                if (MODE != MODE_DUMP) mapw = ld_u32_hint(reinterpret_cast<const uint32_t*>(p.map_pos + pos * (uint64_t)(N * N)) + lane, pol_stream);
            } else {
                if (lane < N) node = p.conn_pos[(uint64_t)pos * N + lane];
                if (MODE != MODE_DUMP) mapw = reinterpret_cast<const uint32_t*>(p.map_pos + pos * (uint64_t)(N * N))[lane];
            }
        }
    };
    auto load_dep = [&](bool ok, int node, long long& a0, long long& a1, double& xv) {
        a0 = 0;
        a1 = 0;
        xv = 0.0;
        const int na = __shfl_sync(FULL, node, x_node);
        if (ok) {
            if (MODE != MODE_DUMP && lane < N) {
                a0 = p.blk_off[node];
                a1 = p.blk_off[node + 1];
            }
            if (lane < N * D) xv = p.vertices[(uint64_t)na * D + x_comp];
        }
    };
    int node0, node1;
    uint32_t mapw, mapw1;
    long long o0, o1;
    double x;
    uint32_t idx = next_pos();
    uint32_t idx_n = next_pos();
    uint32_t idx_n2 = next_pos();
    bool valid = idx < count;
    load_ids(idx, node0, mapw);
    load_ids(idx_n, node1, mapw1);
    load_dep(valid, node0, o0, o1, x);
    while (valid) {  // warp-uniform
        // ---- stage the current element
        if (lane < N * D) s_X[lane] = x;
        if (MODE != MODE_DUMP) {
            if (lane < N) {
                s_base[lane] = (long long)(S * S) * o0;
                s_rowlen[lane] = (int)(o1 - o0) * S;
            }
            s_pos[lane] = mapw;
        }
        // ---- stage A for the element two iterations ahead, stage B for the next one
        const bool valid_n = idx_n < count;
        int node2;
        uint32_t mapw2;
        load_ids(idx_n2, node2, mapw2);
        const uint32_t idx_n3 = next_pos();
        long long o0_n, o1_n;
        double x_n;
        load_dep(valid_n, node1, o0_n, o1_n, x_n);
        if (MODE == MODE_ATOMIC && (dbg & 8)) {
            // pull the next element's CSR row blocks into L2 (first touches otherwise pay the DRAM latency inside the RED path)
            const int pn = lane >> 2;
            const long long b0v = __shfl_sync(FULL, o0_n, pn), b1v = __shfl_sync(FULL, o1_n, pn);
            const char* rb = reinterpret_cast<const char*>(p.values + (long long)(S * S) * b0v);
            const long long bytes = (long long)(S * S) * (b1v - b0v) * 8;
            for (long long off = (long long)(lane & 3) * 128; off < bytes; off += 512)
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(rb + off));
        }
        __syncwarp();

        // ---- geometry: row gi of J at point gq
        if (!(dbg & 1)) {
            double Jr[D];
            {
                double lo[D], hi[D];
                const double x0 = s_X[gi], x4 = s_X[4 * D + gi];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    lo[j] = x0 * R[0][j];
                    hi[j] = x4 * R[4][j];
                }
#pragma unroll
                for (int a = 1; a < 4; ++a) {
                    const double xa = s_X[a * D + gi], xb = s_X[(a + 4) * D + gi];
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        lo[j] = fma(xa, R[a][j], lo[j]);
                        hi[j] = fma(xb, R[a + 4][j], hi[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < D; ++j) Jr[j] = lo[j] + hi[j];
            }
            double r1[D], r2[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                r1[j] = __shfl_sync(FULL, Jr[j], src1);
                r2[j] = __shfl_sync(FULL, Jr[j], src2);
            }
            // cofactors of row gi: c = J[gi+1] x J[gi+2]
            double c[D];
            c[0] = r1[1] * r2[2] - r1[2] * r2[1];
            c[1] = r1[2] * r2[0] - r1[0] * r2[2];
            c[2] = r1[0] * r2[1] - r1[1] * r2[0];
            double det = Jr[0] * c[0] + Jr[1] * c[1] + Jr[2] * c[2];
            det = __shfl_sync(FULL, det, lane & ~3);  // first-row expansion, as the reference's determinant()
            double r = 0.0;
            if (det != 0.0) {
                r = copysign(sqw * rsqrt(fabs(det)), det);  // sqrt(w |det|) / det: gradients come out pre-scaled by sqrt(alpha)
            } else if (gact && s4 == 0) {
                flag_error(p.errword, p.elem_ids ? (uint64_t)p.elem_ids[idx] : p.first_elem + (uint64_t)idx, FB200_ERR_SINGULAR_JACOBIAN);
            }
#pragma unroll
            for (int j = 0; j < D; ++j) c[j] *= r;
            if (s4 < 3) {
                double* go = s_G + 3 * gq + gi;
#pragma unroll
                for (int a = 0; a < N; ++a) go[a * GS] = fma(c[2], R[a][2], fma(c[1], R[a][1], c[0] * R[a][0]));
            }
        }
        __syncwarp();

        // ---- two node blocks per lane: S = G G^T on the FP64 tensor pipe
        double K0[S][S], K1[S][S];
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
            for (int j = 0; j < S; ++j) K0[i][j] = K1[i][j] = 0.0;
        if (!(dbg & 1)) {
            double ga[2][D];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int m = 0; m < D; ++m) ga[ks][m] = s_G[frag + 12 * ks + m];
            if constexpr (S == 1) {
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int m = 0; m < D; ++m) dmma_m8n8k4(t0, t1, ga[ks][m], ga[ks][m]);
                K0[0][0] = t0;
                K1[0][0] = t1;
            } else {
                double M0[D][D], M1[D][D];
#pragma unroll
                for (int m = 0; m < D; ++m)
#pragma unroll
                    for (int n = 0; n < D; ++n) { M0[m][n] = 0.0; M1[m][n] = 0.0; }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int m = 0; m < D; ++m)
#pragma unroll
                        for (int n = 0; n < D; ++n) dmma_m8n8k4(M0[m][n], M1[m][n], ga[ks][m], ga[ks][n]);
                const double tr0 = M0[0][0] + M0[1][1] + M0[2][2], tr1 = M1[0][0] + M1[1][1] + M1[2][2];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        K0[i][j] = mu * ((i == j ? tr0 : 0.0) + M0[j][i]) + lam * M0[i][j];
                        K1[i][j] = mu * ((i == j ? tr1 : 0.0) + M1[j][i]) + lam * M1[i][j];
                    }
            }
        }
        // row r of K_e starts at krow(r): 24 r + r/6 (elasticity) or 9 r (Laplace)
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const int r = S * ba + i;
            const int kr = S == 1 ? r * (SN + 1) : r * SN + (ba >> 1);
#pragma unroll
            for (int j = 0; j < S; ++j) {
                s_K[kr + S * b0 + j] = K0[i][j];
                s_K[kr + S * (b0 + 1) + j] = K1[i][j];
            }
        }
        __syncwarp();

        // ---- scatter: one K_e row per instruction, lane = column
        if (lane < SN && !(dbg & 2) && !((dbg & 16) && col_j != 0)) {
            if (MODE == MODE_DUMP) {
                double* out = p.dump + (uint64_t)idx * (uint64_t)(SN * SN);
#pragma unroll
                for (int r = 0; r < SN; ++r) {
                    const int kr = S == 1 ? r * (SN + 1) : r * SN + r / 6;
                    out[(uint64_t)lane * SN + r] = s_K[kr + lane];
                }
            } else {
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    double* rowp = p.values + (s_base[a] + (long long)(S * (int)s_pos16[a * N + col_b] + col_j));
                    const int rl = s_rowlen[a];
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        if ((dbg & 32) && i != 0) continue;
                        const int r = S * a + i;
                        const int kr = S == 1 ? r * (SN + 1) : r * SN + (a >> 1);
                        const double v = s_K[kr + lane];
                        double* dst = rowp + i * rl;
                        if (MODE == MODE_ATOMIC) {
                            if (dbg & 4) *dst = v;
                            else if constexpr (hints) red_add_f64_hint(dst, v, pol_keep);
                            else atomicAdd(dst, v);
                        }
                        else *dst += v;
                    }
                }
            }
        }
        __syncwarp();
        idx = idx_n;
        idx_n = idx_n2;
        idx_n2 = idx_n3;
        valid = valid_n;
        mapw = mapw1;
        mapw1 = mapw2;
        node0 = node1;
        node1 = node2;
        o0 = o0_n;
        o1 = o1_n;
        x = x_n;
    }
}
