// Hex8 warp-per-element assembly kernel (included by assemble.cu).
//
// Same mathematics as assemble_elements_kernel (see assemble.cu for the reference citations), re-mapped so that every
// phase uses all 32 lanes - the FP64 pipe retires a warp instruction in 2 cycles no matter how many lanes are active -
// and so that shared memory and the L2 reduction units see conflict-free / sector-coalesced traffic:
//   load     lanes 0-7 read the node ids (one 32 B sector), every lane reads two uint16 map entries (exactly the two node
//            blocks it will compute); ids are broadcast with shuffles for the coordinate gather (lanes 0-23).  The next
//            element's ids, map, block offsets and coordinates are prefetched into registers while the current element is
//            processed, so no global-load latency sits on the critical path of the loop.
//   geometry lane = (q, s): 4 lanes share a quadrature point; each accumulates J over 2 of the 8 nodes, a 2-step
//            xor-shuffle reduction completes J, all four invert it (closed form) and each pushes 2 nodes' gradients forward,
//            pre-scaled by sqrt(w |det J|), into shared memory.
//   layout   per-point rows of 33 doubles, node a at offset 4a + (a >> 2): the (q, s) lanes of a half-warp hit 16 distinct
//            bank pairs, and the 8 (resp. 4) distinct addresses of the block phase's broadcast loads never share a bank.
//   blocks   lane owns K_{a,b0}, K_{a,b0+1} (a = lane/4, b0 = 2 (lane%4)): per point 9 broadcast LDS.64 (1 wavefront each;
//            a 128-bit load would cost 4 wavefronts regardless of broadcast) and 18 DFMA.
//   stage    K_e row-major in shared memory, row r at 24 r + r/6 (conflict-free for the (a, m) writer layout).
//   scatter  one instruction per K_e row, lane = column: the 3 doubles of a node block are contiguous in the CSR row, so a
//            reduction instruction touches ~12 sectors instead of 32 (the L2 RED units work per sector).
// Requires uniform operator parameters (the per-point case uses assemble_elements_kernel).
#pragma once

// L2 eviction-priority hints (PTX createpolicy / .L2::cache_hint): CSR rows under accumulation are kept (evict_last) until all
// of a node's elements have contributed; the streamed, read-once connectivity / map rows are marked evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void red_add_f64_hint(double* addr, double v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* addr) {
    uint32_t r;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_u32(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_u32_hint(const void* addr, uint64_t pol) {
    uint32_t r;
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(addr), "l"(pol));
    return r;
}

template <int OP, int MODE, int THREADS, int MINB, bool DYN, bool HINT, int CHUNK = 8, bool ZFUSE = false>
__global__ void __launch_bounds__(THREADS + (ZFUSE ? 32 : 0), MINB) assemble_hex8_kernel(const AssembleParams p) {
    constexpr int BLOCK = THREADS + (ZFUSE ? 32 : 0);  // with the fused zero-fill one extra warp per CTA only clears rows
    constexpr int N = 8, D = 3;
    constexpr int S = OP == FB200_LAPLACE ? 1 : D;
    constexpr int SN = S * N;
    constexpr int TS = 33;                                  // row stride of the gradient tables / per-point gradient rows
    constexpr int KLEN = S == 1 ? SN * (SN + 1) : SN * SN + 4;  // staged K_e
    constexpr int WARPS = THREADS / 32;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int nq = p.nq;
    double* s_w = smem;
    double* s_ggeo = s_w + nq;
    double* s_gref = s_ggeo + nq * TS;
    const int tab_len = (nq * (1 + 2 * TS) + 1) & ~1;
    for (int i = threadIdx.x; i < nq; i += BLOCK) s_w[i] = p.tab[i];
    for (int i = threadIdx.x; i < nq * N * D; i += BLOCK) {
        const int q = i / (N * D), r = i - q * (N * D);
        const int a = r / D, j = r - a * D;
        s_ggeo[q * TS + 4 * a + (a >> 2) + j] = p.tab[3 * nq + i];
        s_gref[q * TS + 4 * a + (a >> 2) + j] = p.tab[3 * nq + nq * N * D + i];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp_doubles = N * D + nq * TS + KLEN + (KLEN & 1) + 28;  // + 8 int64 + 8 int32 + 32 uint32
    double* s_X = smem + tab_len + warp * warp_doubles;
    double* s_g = s_X + N * D;
    double* s_K = s_g + nq * TS;
    long long* s_base = reinterpret_cast<long long*>(s_K + KLEN + (KLEN & 1));
    int* s_rowlen = reinterpret_cast<int*>(s_base + N);
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_rowlen + N);
    const uint16_t* s_pos16 = reinterpret_cast<const uint16_t*>(s_pos);
    __syncthreads();

    if constexpr (ZFUSE) {
        // ---- fused zero-fill (values = contributions without a separate memset pass).
        // The extra warp of every CTA is a "clearing" warp: it walks the chunks of the processing order a bounded distance AHEAD
        // of the ticket counter, clears the CSR rows whose first contribution comes from those chunks (coalesced stores that
        // create the lines in L2 without a DRAM read), fences, and publishes each row with the current epoch.  Only these warps
        // pay for the fence; the element warps just check the (prefetched) row flags before they add.  The reductions then hit
        // lines that are already in L2, and DRAM sees every CSR value once - on its final write-back.
        if (warp == WARPS) {
            constexpr uint32_t ZB = 32 / CHUNK > 0 ? 32 / CHUNK : 1;  // chunks per round (~32 rows = one round of list loads)
            for (uint32_t c0 = blockIdx.x * ZB; c0 < p.num_chunks; c0 += gridDim.x * ZB) {
                while (true) {  // stay within p.zero_look chunks of the front so the cleared lines are still in L2 when used
                    const uint32_t t = *reinterpret_cast<volatile unsigned int*>(p.ticket32);
                    if (c0 < t + p.zero_look) break;
                    __nanosleep(256);
                }
                const uint32_t c1 = c0 + ZB < p.num_chunks ? c0 + ZB : p.num_chunks;
                const long long zb = p.zero_off[c0], ze = p.zero_off[c1];
                for (long long k0 = zb; k0 < ze; k0 += 32) {
                    long long rbase = 0;
                    int rlen = 0;
                    if (k0 + lane < ze) {
                        rbase = p.zero_base[k0 + lane];
                        rlen = p.zero_len[k0 + lane];
                    }
                    const int rows = (int)(ze - k0 < 32 ? ze - k0 : 32);
                    for (int j = 0; j < rows; ++j) {
                        double* row = p.values + __shfl_sync(FULL, rbase, j);
                        const int len = __shfl_sync(FULL, rlen, j);
                        for (int w = lane; w < len; w += 32) row[w] = 0.0;
                    }
                }
                __threadfence();
                __syncwarp();
                for (long long k = zb + lane; k < ze; k += 32) st_release_u32(p.row_epoch + p.zero_nodes[k], p.epoch);
            }
            return;
        }
    }

    const int qs = lane >> 2, s4 = lane & 3;                // geometry role
    const int ga0 = 4 * s4, ga1 = 4 * (s4 + 4) + 1;         // row offsets of this lane's two nodes (s4, s4 + 4)
    const int ba = lane >> 2, b0 = 2 * (lane & 3);          // block role
    const int oa = 4 * ba + (ba >> 2);
    const int ob0 = 4 * b0 + (b0 >> 2), ob1 = 4 * (b0 + 1) + ((b0 + 1) >> 2);
    const int col_b = lane / S, col_j = lane - col_b * S;   // scatter role (lane < SN)
    const int xl = lane < N * D ? lane : 0;
    const int x_node = xl / D, x_comp = xl - x_node * D;
    const double mu = p.mu, lam = p.lam;
    constexpr bool hints = HINT;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();

    // Positions (in processing order) are handed out dynamically: a warp takes a ticket = CHUNK consecutive positions from a
    // global counter, so the set of elements in flight stays a compact front of the Morton order no matter how unevenly the
    // warps progress (far-die L2 latency, RED back-pressure).  A compact front is what keeps a node's CSR rows resident in L2
    // from its first to its last contribution.  (Without a ticket counter: static grid-stride assignment.)
    // Software pipeline per warp over its position sequence:
    //   stage A (two elements ahead): node ids + map words          (addresses known in advance)
    //   stage B (one element ahead) : block offsets + coordinates   (addresses depend on stage A's ids, which arrived an iteration ago)
    //   stage C (current element)   : everything is in registers.
    // positions fit 32 bits (the upload rejects meshes with >= 2^31 incidences); saturate instead of wrapping
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
    auto load_dep = [&](bool ok, int node, long long& a0, long long& a1, double& xv, uint32_t& flag) {
        a0 = 0;
        a1 = 0;
        xv = 0.0;
        flag = p.epoch;
        if constexpr (ZFUSE) {
            if (ok && lane < N) flag = ld_acquire_u32(p.row_epoch + node);
        }
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
    uint32_t rflag;
    uint32_t idx = next_pos();
    uint32_t idx_n = next_pos();
    uint32_t idx_n2 = next_pos();
    bool valid = idx < count;
    load_ids(idx, node0, mapw);
    load_ids(idx_n, node1, mapw1);
    load_dep(valid, node0, o0, o1, x, rflag);
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
        uint32_t rflag_n;
        load_dep(valid_n, node1, o0_n, o1_n, x_n, rflag_n);
        __syncwarp();

        // ---- geometry: 8 quadrature points per pass, 4 lanes each
        for (int q0 = 0; q0 < nq; q0 += 8) {
            const int q = q0 + qs;
            const bool act = q < nq;
            const int qq = act ? q : 0;
            const double* tg = s_ggeo + qq * TS;
            double J[D][D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) J[i][j] = s_X[s4 * D + i] * tg[ga0 + j];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) J[i][j] = fma(s_X[(s4 + 4) * D + i], tg[ga1 + j], J[i][j]);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    J[i][j] += __shfl_xor_sync(FULL, J[i][j], 1);
                    J[i][j] += __shfl_xor_sync(FULL, J[i][j], 2);
                }
            const double c00 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
            const double c01 = J[1][0] * J[2][2] - J[2][0] * J[1][2];
            const double c02 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
            const double det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02;
            double r = 0.0;
            if (det != 0.0) {
                r = sqrt(s_w[qq] * fabs(det)) / det;  // (1/det) * sqrt(w |det|): gradients come out pre-scaled
            } else if (act && s4 == 0) {
                flag_error(p.errword, p.elem_ids ? (uint64_t)p.elem_ids[idx] : p.first_elem + (uint64_t)idx, FB200_ERR_SINGULAR_JACOBIAN);
            }
            double Ji[D][D];  // sqrt(alpha) * J^{-1}
            Ji[0][0] = c00 * r;
            Ji[0][1] = (J[0][2] * J[2][1] - J[2][2] * J[0][1]) * r;
            Ji[0][2] = (J[0][1] * J[1][2] - J[1][1] * J[0][2]) * r;
            Ji[1][0] = -c01 * r;
            Ji[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * r;
            Ji[1][2] = (J[0][2] * J[1][0] - J[1][2] * J[0][0]) * r;
            Ji[2][0] = c02 * r;
            Ji[2][1] = (J[0][1] * J[2][0] - J[2][1] * J[0][0]) * r;
            Ji[2][2] = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) * r;
            const double* tr = s_gref + qq * TS;
            double* go = s_g + qq * TS;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int off = h == 0 ? ga0 : ga1;
                const double g0 = tr[off], g1 = tr[off + 1], g2 = tr[off + 2];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const double v = fma(Ji[2][i], g2, fma(Ji[1][i], g1, Ji[0][i] * g0));  // (J^{-T} g)_i
                    if (act) go[off + i] = v;
                }
            }
        }
        __syncwarp();

        // ---- two node blocks per lane
        double K0[S][S], K1[S][S];
        if constexpr (S == 1) {
            double t0 = 0.0, t1 = 0.0;
            for (int q = 0; q < nq; ++q) {
                const double* gq = s_g + q * TS;
                const double a0 = gq[oa], a1 = gq[oa + 1], a2 = gq[oa + 2];
                t0 = fma(a0, gq[ob0], fma(a1, gq[ob0 + 1], fma(a2, gq[ob0 + 2], t0)));
                t1 = fma(a0, gq[ob1], fma(a1, gq[ob1 + 1], fma(a2, gq[ob1 + 2], t1)));
            }
            K0[0][0] = t0;
            K1[0][0] = t1;
        } else {
            double M0[D][D], M1[D][D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) { M0[i][j] = 0.0; M1[i][j] = 0.0; }
#pragma unroll 2
            for (int q = 0; q < nq; ++q) {
                const double* gq = s_g + q * TS;
                const double va[3] = {gq[oa], gq[oa + 1], gq[oa + 2]};
                const double vb[6] = {gq[ob0], gq[ob0 + 1], gq[ob0 + 2], gq[ob1], gq[ob1 + 1], gq[ob1 + 2]};
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        M0[i][j] = fma(va[i], vb[j], M0[i][j]);
                        M1[i][j] = fma(va[i], vb[3 + j], M1[i][j]);
                    }
            }
            const double tr0 = M0[0][0] + M0[1][1] + M0[2][2], tr1 = M1[0][0] + M1[1][1] + M1[2][2];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    K0[i][j] = mu * ((i == j ? tr0 : 0.0) + M0[j][i]) + lam * M0[i][j];
                    K1[i][j] = mu * ((i == j ? tr1 : 0.0) + M1[j][i]) + lam * M1[i][j];
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

        if constexpr (ZFUSE) {
            // all 8 rows must have been cleared (by whichever chunk touches them first) before we add to them; the flags were
            // prefetched an iteration ago and are almost always current - otherwise poll
            while (!__all_sync(FULL, rflag == p.epoch)) {
                if (rflag != p.epoch) rflag = ld_acquire_u32(p.row_epoch + node0);
            }
        }
        // ---- scatter: one K_e row per instruction, lane = column
        if (lane < SN) {
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
                        const int r = S * a + i;
                        const int kr = S == 1 ? r * (SN + 1) : r * SN + (a >> 1);
                        const double v = s_K[kr + lane];
                        double* dst = rowp + i * rl;
                        if (MODE == MODE_ATOMIC) {
                            if constexpr (hints) red_add_f64_hint(dst, v, pol_keep);
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
        rflag = rflag_n;
        o0 = o0_n;
        o1 = o1_n;
        x = x_n;
    }
}
