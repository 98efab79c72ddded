// Tet4 chunk-local assembly kernel (included by assemble.cu): ATOMIC scatter for linear tetrahedra with a one-point rule.
//
// Reference semantics: assemble_element_elliptic_matrix elliptic.rs:361-439 with the constant Tet4 gradients
// (tetrahedron.rs:551-590), contraction laplace.rs:60-68 / materials.rs:108-122, scatter global.rs:155-178.
//
// A K_e of a linear tet is 16 rank-one blocks  K_ab = mu [(g_a.g_b) I + g_b g_a^T] + lambda g_a g_b^T,  g_a = sqrt(w |det J|) J^{-T} grad_ref phi_a,
// i.e. almost no arithmetic - the cost of the reference's per-element scatter is the 144 fp64 additions into the CSR.  On the
// device every such addition is one reduction in L2, and a BCC tet mesh sends ~6.4 of them to every CSR value.  This kernel
// removes most of them: a CTA takes a chunk of C consecutive elements of the Morton order and
//   phase 1  one thread per element: gather the 4 vertices, J^{-1}, the 4 scaled gradients -> shared memory (12 doubles / element);
//   phase 2  one thread per SLOT (= node block (I, J) that the chunk touches; lists from chunks.cpp): sums S = sum_e g_a(e) (x) g_b(e)
//            over the chunk's contributors in registers - K is linear in S, so the material law is applied once per slot - and
//            issues the block's s x s reductions (plain stores when the row node is complete inside the chunk and the call
//            overwrites).  Contributors are visited in ascending element order for both (I, J) and (J, I): the result stays
//            exactly symmetric.
// Slots are ordered by contributor count, so the lanes of a warp run loops of equal length.
// Multi-GPU (PEER): the slots of partition-interface rows (slot_flags bit 1) are additionally reduced into the neighbouring rank's values
// through a peer-mapped pointer - the interface exchange of the Tet4 path (config C5) is part of the scatter, as in hex8_tile_kernel.cuh.
// (A variant with one 32-byte record per slot - destination, row length, contributor range and the first eight tags, prefetched one
//  slot ahead - was measured on the C5 share: 1.29 - 1.33 ms against 1.20 ms for these separate arrays; profiles/r02/README.md.)
#pragma once

template <int OP, int T, int C, bool PEER = false>
__global__ void __launch_bounds__(T, (T >= 1024 ? 1 : (C <= 512 ? 4 : 2))) assemble_tet4_chunk_kernel(const AssembleParams p) {
    constexpr int N = 4, D = 3;
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    extern __shared__ double s_g[];  // [12][C]
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.x;
    double gref[N][D];
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) gref[a][j] = p.tab[3 + a * D + j];  // nq == 1: w | mu | lam | ggeo[12] | gref[12] (identical for Tet4)
    const double w = p.tab[0];
    const double mu = p.mu, lam = p.lam;
    const bool overwrite = p.accumulate == 0;

    while (true) {
        __syncthreads();  // phase 2 of the previous chunk has finished reading s_g (and s_ticket)
        if (tid == 0) s_ticket = atomicAdd(p.ticket32, 1u);
        __syncthreads();
        const uint32_t chunk = s_ticket;
        if (chunk >= p.num_chunks) break;
        const uint64_t p0 = (uint64_t)chunk * C;
        const int ne = (int)(p.count - p0 < (uint64_t)C ? p.count - p0 : (uint64_t)C);

        // ---- phase 1: element geometry
        for (int el = tid; el < ne; el += T) {
            const int4 nd = reinterpret_cast<const int4*>(p.conn_pos)[p0 + el];
            const int ids[N] = {nd.x, nd.y, nd.z, nd.w};
            double X[N * D];
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) X[a * D + i] = p.vertices[(uint64_t)ids[a] * D + i];
            double Jinv[D][D], det;
            const bool ok = jacobian_inverse<N, D>(X, &gref[0][0], Jinv, &det);
            double gs = 0.0;
            if (ok) gs = sqrt(w * fabs(det));
            else flag_error(p.errword, p.elem_ids ? (uint64_t)p.elem_ids[p0 + el] : p0 + el, FB200_ERR_SINGULAR_JACOBIAN);
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const double v = ok ? fma(Jinv[2][i], gref[a][2], fma(Jinv[1][i], gref[a][1], Jinv[0][i] * gref[a][0])) : 0.0;  // (J^{-T} g)_i
                    s_g[(a * D + i) * C + el] = v * gs;
                }
        }
        __syncthreads();

        // ---- phase 2: one thread per slot
        const long long so = p.slot_off[chunk];
        const int U = (int)(p.slot_off[chunk + 1] - so);
        const int npairs = ne * (N * N);
        const uint16_t* contrib = p.contrib + p0 * (uint64_t)(N * N);
        for (int u = tid; u < U; u += T) {
            const int cb = p.slot_cbeg[so + u];
            const int ce = u + 1 < U ? (int)p.slot_cbeg[so + u + 1] : npairs;
            // slot metadata first (independent loads: the destination and the row length are precomputed, chunks.cpp): their latency
            // overlaps the contributor loop
            const long long dsti = p.slot_dst[so + u];
            const int rl = p.slot_rl[so + u];
            const unsigned sflags = p.slot_flags[so + u];
            const bool st = overwrite && (sflags & 1u);
            double M[D][D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) M[i][j] = 0.0;
            unsigned tag_next = __ldg(contrib + cb);  // every slot has at least one contributor
            for (int t = cb; t < ce; ++t) {
                const unsigned tag = tag_next;
                if (t + 1 < ce) tag_next = __ldg(contrib + t + 1);
                const int el = tag >> 4, a = (tag >> 2) & 3, b = tag & 3;
                const double* ga = s_g + (a * D) * C + el;
                const double* gb = s_g + (b * D) * C + el;
                const double va[D] = {ga[0], ga[C], ga[2 * C]}, vb[D] = {gb[0], gb[C], gb[2 * C]};
                if constexpr (S == 1) {
                    M[0][0] = fma(va[0], vb[0], fma(va[1], vb[1], fma(va[2], vb[2], M[0][0])));
                } else {
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = 0; j < D; ++j) M[i][j] = fma(va[i], vb[j], M[i][j]);
                }
            }
            double* dst = p.values + dsti;
            double* pdst = nullptr;  // the same block in the neighbouring rank's copy of the row (identical row layout, other offset)
            if constexpr (PEER) {
                if (sflags & 2u) {
                    const int node = p.slot_node[so + u];
                    const uint32_t pw = __ldg(p.peer_row + node);
                    if (pw) pdst = p.peer_values[pw >> 31] + ((long long)(S * S) * (long long)((pw & 0x7fffffffu) - 1u) + (long long)S * (long long)p.slot_k[so + u]);
                }
            }
            if constexpr (S == 1) {
                if (st) dst[0] = M[0][0];
                else atomicAdd(dst, M[0][0]);
                if (PEER && pdst) red_add_f64_sys(pdst, M[0][0]);
            } else {
                const double tr = M[0][0] + M[1][1] + M[2][2];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const double v = mu * ((i == j ? tr : 0.0) + M[j][i]) + lam * M[i][j];
                        if (st) dst[i * rl + j] = v;
                        else atomicAdd(dst + i * rl + j, v);
                        if (PEER && pdst) red_add_f64_sys(pdst + i * rl + j, v);
                    }
            }
        }
    }
}
