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
#pragma once

template <int OP, int T, int C>
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

        // ---- phase 2: one thread per slot.  A slot is one 32-byte record (chunks.cpp) holding everything but the tags beyond the eighth,
        // and the record of the thread's NEXT slot is in flight while the current one is summed: no dependent global load in the loop
        const long long so = p.slot_off[chunk];
        const int U = (int)(p.slot_off[chunk + 1] - so);
        const uint16_t* contrib = p.contrib + p0 * (uint64_t)(N * N);
        const ulonglong2* recs = reinterpret_cast<const ulonglong2*>(p.slot_rec) + 2 * so;
        ulonglong2 ra = make_ulonglong2(0, 0), rb = make_ulonglong2(~0ull, ~0ull);
        if (tid < U) {
            ra = __ldg(recs + 2 * tid);
            rb = __ldg(recs + 2 * tid + 1);
        }
        for (int u = tid; u < U; u += T) {
            const ulonglong2 ca = ra, cb2 = rb;
            if (u + T < U) {
                ra = __ldg(recs + 2 * (u + T));
                rb = __ldg(recs + 2 * (u + T) + 1);
            }
            const long long dsti = (long long)(ca.x & ((1ull << 62) - 1ull));
            const bool st = overwrite && (ca.x >> 62) != 0ull;
            const int rl = (int)(unsigned int)ca.y;
            const int cb = (int)((ca.y >> 32) & 0xffffull), cnt = (int)(ca.y >> 48);
            double M[D][D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) M[i][j] = 0.0;
            unsigned long long tw = cb2.x;  // tags 0-3; then 4-7 (cb2.y); beyond the eighth from the chunk's tag list
            unsigned tag_far = cnt > 8 ? (unsigned)__ldg(contrib + cb + 8) : 0u;
            for (int t = 0; t < cnt; ++t) {
                unsigned tag;
                if (t < 8) {
                    if (t == 4) tw = cb2.y;
                    tag = (unsigned)(tw & 0xffffull);
                    tw >>= 16;
                } else {
                    tag = tag_far;
                    if (t + 1 < cnt) tag_far = __ldg(contrib + cb + t + 1);
                }
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
            if constexpr (S == 1) {
                if (st) dst[0] = M[0][0];
                else atomicAdd(dst, M[0][0]);
            } else {
                const double tr = M[0][0] + M[1][1] + M[2][2];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const double v = mu * ((i == j ? tr : 0.0) + M[j][i]) + lam * M[i][j];
                        if (st) dst[i * rl + j] = v;
                        else atomicAdd(dst + i * rl + j, v);
                    }
            }
        }
    }
}
