// High-order element assembly kernel - Hex27 (tri-quadratic, 81 x 81 K_e), Hex20 (serendipity, 60 x 60) and Tet10 (quadratic, 30 x 30):
// one CTA of 10 warps per element, node-block contraction on the FP64 tensor pipe (included by assemble.cu; north_star: "tensor cores
// only on the high-order Hex20/Hex27/Tet10 path").  Template parameters N (nodes) and NG (vertices of the embedded linear element that
// carries the geometry: 8 = Hex8, 4 = Tet4); the description below is written for Hex27, the other two differ in the counts only
// (Hex20: 3 node tiles = 6 tile pairs, rows of 60; Tet10: 2 tiles = 3 pairs, 4 points = one k-step, rows of 30, constant Jacobian).
//
// Reference semantics (InteractiveComputerGraphics/fenris @ 7181b15): assemble_element_elliptic_matrix elliptic.rs:361-439;
// Hex27 is SUB-parametric - the Jacobian comes from the embedded Hex8 of the first 8 vertices (hexahedron.rs:318-335) - and the
// basis gradients are the tri-quadratic ones (hexahedron.rs:269-315); contraction materials.rs:108-122 / laplace.rs:60-68.
//
//   geometry  thread (q, i), 4 threads per quadrature point (q < 28): row i of J from the 8 corner vertices with the trilinear
//             reference gradients of point q held in registers, rows exchanged by shuffles, cofactor row c_i, det J (first-row
//             expansion, broadcast), r = sqrt(w |det J|) / det J; then component i of all 27 scaled physical gradients
//             g_a[i](q) = r (c_i . grad_ref phi_a(q)) from the shared-memory table, stored as G[i][a][q] (node rows padded to 32,
//             points to 28, pads stay zero).
//   blocks    S_ab = sum_q g_a(q) (x) g_b(q) = G G^T.  Nodes are split into 4 tiles of 8; warp w owns one of the 10 tile pairs
//             (ta <= tb) - the other triangle is the transpose - and accumulates the 9 component tiles (m, n) over the 7 k-steps
//             with mma.sync.m8n8k4.f64 (63 DMMA, fragments G[m][8 ta + lane/4][4 ks + lane%4] / G[n][8 tb + ..] loaded once):
//             the accumulator fragments hand lane (g, t) the complete 3 x 3 blocks of node pairs (8ta+g, 8tb+2t), (.., 8tb+2t+1).
//   epilogue  K_ab = mu [tr(S_ab) I + S_ab^T] + lambda S_ab; K_ab and (ta < tb) K_ba = K_ab^T staged row-major in shared memory.
//   scatter   one third of a K_e row (27 columns = 9 node blocks) per reduction instruction, lane = column.
// Uniform operator parameters, positive weights, nq <= 28.
#pragma once

constexpr int kH27GQ = 28;                 // point stride of G (== 12 mod 16: conflict-free fragment loads)
constexpr int kH27GC = 32 * kH27GQ + 4;    // component stride of G
constexpr int kH27Threads = 320;

template <int OP, int N = 27>
__host__ __device__ constexpr int hex27_kstride() { return (OP == FB200_LAPLACE ? N : 3 * N) + 2; }  // odd row stride of the staged K_e (Hex27: 29 / 83)

template <int N>
__host__ __device__ constexpr int hex27_mapstride() { return (N * N + 7) & ~7; }  // u16 map of one element, padded (Hex27: 736)

template <int OP, int N = 27>
__host__ __device__ inline size_t hex27_smem_bytes(int nq) {
    constexpr int S = OP == FB200_LAPLACE ? 1 : 3;
    size_t doubles = (size_t)nq * (N * 3) + 3 * kH27GC + (size_t)(N * S) * hex27_kstride<OP, N>() + 2 * 24 /* X */ + 2 * 28 /* base */;
    return doubles * 8 + 2 * 28 * 4 /* rowlen */ + 2 * hex27_mapstride<N>() * 2 /* map */ + 16;
}

// Software pipeline of one CTA over its elements e_0, e_1, ... (dynamic tickets):
//     | P2(j): DMMA + epilogue, all 10 warps | sync A | warps 0-3: P1(j+1) geometry, then the global loads of P0(j+2)   | sync B | P0(j+2) -> smem |
//     |                                      |        | warps 4-9: P3(j) scatter (the long, reduction-throughput-bound phase) |        |                 |
// so the dependent global loads (connectivity -> row offsets / vertices) and the geometry never sit on the critical path.
template <int OP, int MODE, int N = 27, int NG = 8>
__global__ void __launch_bounds__(kH27Threads, 2) assemble_hex27_mma_kernel(const AssembleParams p) {
    constexpr int D = 3;
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    constexpr int SN = S * N;
    constexpr int KST = hex27_kstride<OP, N>();
    constexpr int NT = (N + 7) / 8, NPAIRS = NT * (NT + 1) / 2;  // node tiles of 8 and tile pairs (ta <= tb): one warp each
    constexpr int MAPS = hex27_mapstride<N>();
    constexpr int MAPK = (N * N + 127) / 128;                   // map entries per loader thread
    static_assert(N <= 28 && NPAIRS <= kH27Threads / 32 && NG * D <= 24 && (NG == 8 || NG == 4), "one warp per tile pair, tables sized for <= 28 nodes");
    constexpr int GEO_THREADS = 128, SCAT_WARPS = kH27Threads / 32 - 4;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int nq = p.nq;  // <= 28
    const int ksteps = (nq + 3) >> 2;
    double* s_gref = smem;                       // [nq][N][3]
    double* s_G = s_gref + nq * (N * 3);         // [3][32][28] (+4 per component)
    double* s_K = s_G + 3 * kH27GC;              // [SN][KST]
    double* s_X = s_K + SN * KST;                // [2][8][3]
    long long* s_base = reinterpret_cast<long long*>(s_X + 48);  // [2][28]
    int* s_rowlen = reinterpret_cast<int*>(s_base + 56);          // [2][28]
    uint16_t* s_map = reinterpret_cast<uint16_t*>(s_rowlen + 56); // [2][MAPS]
    __shared__ unsigned int s_tk[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < nq * (N * 3); i += kH27Threads) s_gref[i] = p.tab[3 * nq + nq * NG * D + i];
    for (int i = tid; i < 3 * kH27GC; i += kH27Threads) s_G[i] = 0.0;

    // ---- geometry role (warps 0..3)
    const int gq = tid >> 2, s4 = tid & 3;
    const int gi = s4 == 3 ? 0 : s4;
    const bool gact = gq < nq;                   // threads with gq >= nq only take part in the shuffles
    double R[NG][D];
    double sqw = 0.0;
    {
        const double* tab_g = p.tab + 3 * nq + (gact ? gq : 0) * (NG * D);
#pragma unroll
        for (int a = 0; a < NG; ++a)
#pragma unroll
            for (int j = 0; j < D; ++j) R[a][j] = gact ? tab_g[a * D + j] : 0.0;
        if (gact) sqw = sqrt(p.tab[gq]);
    }
    const int src1 = (lane & ~3) | (gi == 2 ? 0 : gi + 1), src2 = (lane & ~3) | (gi == 0 ? 2 : gi - 1);

    // ---- block role: warp -> tile pair (ta <= tb)
    int ta = 0, tb = warp;
    {
        int rem = NT;
        while (rem > 0 && tb >= rem) {  // pairs in the order (0,0) .. (0,NT-1), (1,1) .. : warp >= NPAIRS has no pair
            tb -= rem;
            ++ta;
            --rem;
        }
        tb += ta;
    }
    const bool has_pair = warp < NPAIRS;
    const int fg = lane >> 2, ft = lane & 3;
    const int na = 8 * ta + fg, nb0 = 8 * tb + 2 * ft;  // this lane's blocks: (na, nb0), (na, nb0 + 1)
    const double mu = p.mu, lam = p.lam;

    // P0 of one element, executed by threads 0..127: global loads into registers, then registers -> buffer `buf`
    struct P0Regs {
        long long base;
        int rowlen;
        double x;
        uint16_t m[MAPK];
    };
    auto p0_load = [&](uint64_t pos, P0Regs& r) {
        const uint64_t e = p.elem_list ? (uint64_t)p.elem_list[pos] : pos;
        const int32_t* en = p.conn + e * N;
        if (tid < N) {
            if (MODE != MODE_DUMP) {
                const int node = en[tid];
                const long long b0 = p.blk_off[node], b1 = p.blk_off[node + 1];
                r.base = (long long)(S * S) * b0;
                r.rowlen = (int)(b1 - b0) * S;
            }
        } else if (tid >= 32 && tid < 32 + NG * D) {
            const int t = tid - 32, a = t / D, i = t - a * D;
            r.x = p.vertices[(uint64_t)en[a] * D + i];
        }
        if (MODE != MODE_DUMP) {
            const uint16_t* mp16 = p.blockmap + e * (uint64_t)(N * N);  // N^2 u16: element rows are 2-byte aligned only
#pragma unroll
            for (int k = 0; k < MAPK; ++k) {
                const int i = tid + k * GEO_THREADS;
                r.m[k] = i < N * N ? mp16[i] : (uint16_t)0;
            }
        }
    };
    auto p0_store = [&](int buf, const P0Regs& r) {
        if (tid < N) {
            if (MODE != MODE_DUMP) {
                s_base[buf * 28 + tid] = r.base;
                s_rowlen[buf * 28 + tid] = r.rowlen;
            }
        } else if (tid >= 32 && tid < 32 + NG * D) {
            s_X[buf * 24 + tid - 32] = r.x;
        }
        if (MODE != MODE_DUMP) {
#pragma unroll
            for (int k = 0; k < MAPK; ++k) {
                const int i = tid + k * GEO_THREADS;
                if (i < N * N) s_map[buf * MAPS + i] = r.m[k];
            }
        }
    };
    auto geometry = [&](int buf, uint64_t pos) {
        const double* sx = s_X + buf * 24;
        double Jr[D];
        if constexpr (NG == 8) {
            double lo[D], hi[D];
            const double x0 = sx[gi], x4 = sx[4 * D + gi];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                lo[j] = x0 * R[0][j];
                hi[j] = x4 * R[4][j];
            }
#pragma unroll
            for (int a = 1; a < 4; ++a) {
                const double xa = sx[a * D + gi], xb = sx[(a + 4) * D + gi];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    lo[j] = fma(xa, R[a][j], lo[j]);
                    hi[j] = fma(xb, R[a + 4][j], hi[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < D; ++j) Jr[j] = lo[j] + hi[j];
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) Jr[j] = sx[gi] * R[0][j];
#pragma unroll
            for (int a = 1; a < NG; ++a) {
                const double xa = sx[a * D + gi];
#pragma unroll
                for (int j = 0; j < D; ++j) Jr[j] = fma(xa, R[a][j], Jr[j]);
            }
        }
        double r1[D], r2[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            r1[j] = __shfl_sync(FULL, Jr[j], src1);
            r2[j] = __shfl_sync(FULL, Jr[j], src2);
        }
        double c[D];
        c[0] = r1[1] * r2[2] - r1[2] * r2[1];
        c[1] = r1[2] * r2[0] - r1[0] * r2[2];
        c[2] = r1[0] * r2[1] - r1[1] * r2[0];
        double det = Jr[0] * c[0] + Jr[1] * c[1] + Jr[2] * c[2];
        det = __shfl_sync(FULL, det, lane & ~3);
        double r = 0.0;
        if (det != 0.0) {
            r = copysign(sqw * rsqrt(fabs(det)), det);
        } else if (gact && s4 == 0) {
            flag_error(p.errword, p.elem_list ? (uint64_t)p.elem_list[pos] : pos, FB200_ERR_SINGULAR_JACOBIAN);
        }
#pragma unroll
        for (int j = 0; j < D; ++j) c[j] *= r;
        if (gact && s4 < 3) {
            const double* tr = s_gref + gq * (N * 3);
            double* go = s_G + gi * kH27GC + gq;
#pragma unroll (N % 9 == 0 ? 9 : 10)
            for (int a = 0; a < N; ++a) go[a * kH27GQ] = fma(c[2], tr[a * D + 2], fma(c[1], tr[a * D + 1], c[0] * tr[a * D]));
        }
    };

    // ---- prologue: tickets of e_0, e_1; P0(0), P0(1); P1(0)
    if (tid == 0) {
        s_tk[0] = atomicAdd(p.ticket32, 1u);
        s_tk[1] = atomicAdd(p.ticket32, 1u);
    }
    __syncthreads();
    uint64_t pos = s_tk[0], pos_n = s_tk[1];
    if (pos >= p.count) return;
    if (tid < GEO_THREADS) {
        P0Regs r;
        p0_load(pos, r);
        p0_store(0, r);
        if (pos_n < p.count) {
            p0_load(pos_n, r);
            p0_store(1, r);
        }
    }
    __syncthreads();
    if (warp < 4) geometry(0, pos);
    __syncthreads();

    for (int cur = 0;; cur ^= 1) {
        // here: s_G = G(e_j), buffer cur = P0(e_j), buffer cur^1 = P0(e_{j+1}) when pos_n is valid
        if (tid == 0) s_tk[cur] = atomicAdd(p.ticket32, 1u);  // e_{j+2}; read after sync A
        // ---- P2: S = G G^T for this warp's tile pair, epilogue, stage K_e
        if (has_pair) {
            double M0[S == 1 ? 1 : D][S == 1 ? 1 : D], M1[S == 1 ? 1 : D][S == 1 ? 1 : D];
#pragma unroll
            for (int m = 0; m < (S == 1 ? 1 : D); ++m)
#pragma unroll
                for (int n = 0; n < (S == 1 ? 1 : D); ++n) { M0[m][n] = 0.0; M1[m][n] = 0.0; }
            const double* fa = s_G + (8 * ta + fg) * kH27GQ + ft;
            const double* fb = s_G + (8 * tb + fg) * kH27GQ + ft;
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
                double A[D], B[D];
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    A[m] = fa[m * kH27GC + 4 * ks];
                    B[m] = fb[m * kH27GC + 4 * ks];
                }
                if constexpr (S == 1) {
#pragma unroll
                    for (int m = 0; m < D; ++m) dmma_m8n8k4(M0[0][0], M1[0][0], A[m], B[m]);
                } else {
#pragma unroll
                    for (int m = 0; m < D; ++m)
#pragma unroll
                        for (int n = 0; n < D; ++n) dmma_m8n8k4(M0[m][n], M1[m][n], A[m], B[n]);
                }
            }
            double K0[S][S], K1[S][S];
            if constexpr (S == 1) {
                K0[0][0] = M0[0][0];
                K1[0][0] = M1[0][0];
            } else {
                const double tr0 = M0[0][0] + M0[1][1] + M0[2][2], tr1 = M1[0][0] + M1[1][1] + M1[2][2];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        K0[i][j] = mu * ((i == j ? tr0 : 0.0) + M0[j][i]) + lam * M0[i][j];
                        K1[i][j] = mu * ((i == j ? tr1 : 0.0) + M1[j][i]) + lam * M1[i][j];
                    }
            }
            if (na < N) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int nb = nb0 + h;
                    if (nb < N) {
#pragma unroll
                        for (int i = 0; i < S; ++i)
#pragma unroll
                            for (int j = 0; j < S; ++j) {
                                const double v = h == 0 ? K0[i][j] : K1[i][j];
                                s_K[(S * na + i) * KST + S * nb + j] = v;
                                if (ta != tb) s_K[(S * nb + j) * KST + S * na + i] = v;  // K_ba = K_ab^T
                            }
                    }
                }
            }
        }
        __syncthreads();  // sync A: K_e staged, s_G free, ticket of e_{j+2} visible
        const uint64_t pos_n2 = s_tk[cur];
        P0Regs pre;
        if (warp < 4) {
            // ---- P1(j+1), then the global loads of P0(j+2)
            if (pos_n < p.count) geometry(cur ^ 1, pos_n);
            if (pos_n2 < p.count) p0_load(pos_n2, pre);
        } else {
            // ---- P3(j): scatter
            const int w6 = warp - 4;
            if (MODE == MODE_DUMP) {
                double* out = p.dump + pos * (uint64_t)(SN * SN);
                for (int t = tid - GEO_THREADS; t < SN * SN; t += kH27Threads - GEO_THREADS) {
                    const int c = t / SN, r = t - c * SN;  // column-major output
                    out[t] = s_K[r * KST + c];
                }
            } else {
                // warp w6 takes K_e rows w6, w6 + 6, ...; lane = column within a third of the row (27 columns = 9 node blocks): the
                // lane's (block, component) split is loop invariant and the row's (node, component) advances without divisions
                constexpr int SEG = (SN + 31) / 32;          // instructions per row (Hex27: 3 x 27 columns, Hex20: 2 x 30, Tet10: 1 x 30)
                static_assert(N % SEG == 0 && SN / SEG <= 32, "a K_e row splits into SEG equal segments of whole node blocks");
                constexpr int LANES = SN / SEG;              // columns per instruction
                const int lb = lane / S, lj = lane - lb * S;  // lane < LANES: block lb (+ N / SEG per segment), component lj
                const long long* sb = s_base + cur * 28;
                const int* sr = s_rowlen + cur * 28;
                const uint16_t* sm = s_map + cur * MAPS;
                int a = w6 / S, i = w6 - a * S;
                for (int r = w6; r < SN; r += SCAT_WARPS) {
                    if (lane < LANES) {
                        double* rowp = p.values + (sb[a] + (long long)i * sr[a] + lj);
                        const uint16_t* mrow = sm + a * N + lb;
                        const double* krow = s_K + r * KST + lane;
#pragma unroll
                        for (int sg = 0; sg < SEG; ++sg) {
                            double* dst = rowp + S * (int)mrow[sg * (N / SEG)];
                            const double v = krow[sg * LANES];
                            if (MODE == MODE_ATOMIC) atomicAdd(dst, v);
                            else *dst += v;
                        }
                    }
                    i += SCAT_WARPS % S;
                    a += SCAT_WARPS / S;
                    if (i >= S) { i -= S; ++a; }
                }
            }
        }
        __syncthreads();  // sync B: buffer cur and s_K are free, s_G = G(e_{j+1})
        if (pos_n >= p.count) break;
        if (warp < 4 && pos_n2 < p.count) p0_store(cur, pre);  // read by P1(j+2) / P3(j+2), both after the next sync A
        pos = pos_n;
        pos_n = pos_n2;
    }
}
