// Hex8 tile-accumulating assembly kernel (included by assemble.cu): ATOMIC scatter, uniform operator parameters, <= 8 points.
//
// Reference semantics: the per-element loop of CsrAssembler / CsrParAssembler::assemble_into_csr (global.rs:133-182, 314-376) with
// assemble_element_elliptic_matrix (elliptic.rs:361-439); the element mathematics is that of assemble_hex8_mma_kernel
// (hex8_mma_kernel.cuh: Jacobian rows per lane, S = G G^T on the FP64 tensor pipe, K_ab = mu [tr(S_ab) I + S_ab^T] + lambda S_ab,
// fenris-solid/src/materials.rs:108-122).  What changes is the scatter (global.rs:155-178, 504-537).
//
// Measured on B200 the per-element scatter - 24 reductions of 24 doubles per element - is bound by the L2 reduction rate: every CSR
// value receives 2.4 reductions on a structured Hex8 mesh.  Here ONE persistent, warp-specialised CTA per SM owns a TILE of up to 64
// consecutive elements of the Morton order at a time (lists from tiles.cpp) and runs a software pipeline over tiles:
//   compute warps (2 groups of 8)  one element per warp and ROUND: geometry + DMMA contraction in registers (two 3 x 3 node blocks
//            per lane), then the blocks with u_a <= u_b are added to the tile's accumulators in shared memory.  Accumulator
//            positions are chosen on the host so that the 16 lanes of a half-warp hit 16 different banks.  A round holds up to 8
//            elements that share no node; the two groups take the rounds alternately and add strictly in round order - group g
//            waits on a named barrier for the other group's previous round and signals its own (bar.sync / bar.arrive) - so
//            there are no shared-memory atomics, and one group computes while the other one adds.
//   helper warps (8)    everything that is not arithmetic, ahead of / behind the compute warps (they give most of their registers
//            to the compute warps, setmaxnreg):
//            tables  node ids, coordinates and block-row offsets of the coming tiles are fetched into shared memory with
//                    cp.async, one stage of the dependent chain (ticket -> header -> node ids -> coordinates / offsets) per
//                    tile, so neither they nor the compute warps ever wait for a dependent global load;
//            flush   the accumulators are double buffered: while tile i is computed, every node block (u, v) of tile i-1 goes to
//                    the CSR ONCE.  Lanes run over (block, column) in CSR order of row u, so an instruction covers whole runs of
//                    neighbouring node blocks (72 B each); blocks below the diagonal are read transposed from the (v, u)
//                    accumulator.  Rows of nodes whose elements all lie in the tile are complete: they are written with plain
//                    stores (no reduction, no DRAM read of the line) when the call overwrites.
//            ownership  an overwriting call does not zero-fill the values first (the reference's assemble() starts from a zeroed
//                    CsrMatrix, global.rs:124-131 - here that would be a 3.9 GB pass of its own): the flush list of a tile has a
//                    STORE segment (rows complete in the tile, and all entries - zeros included - of shared rows this tile OWNS,
//                    being the lowest-numbered tile that touches the node) and a REDUCE segment.  After its stores a tile publishes
//                    flag[tile] = launch epoch (barrier, fence, release store); before its reductions it waits for the flags of the
//                    owners of the rows it adds to (tiles.cpp `wait`).  A tile only waits for lower-numbered tiles.  With waits the
//                    tiles are dealt round-robin (tile = CTA + k CTAs, cooperative launch: all CTAs resident) - under the atomic
//                    ticket a waiting CTA lets the six tiles it holds age while others take tiles that depend on them (measured
//                    5 - 6 ms instead of 2.5; profiles/r02/README.md); the ticket remains for launches without waits.
//            colours (tile_list) a launch may cover the tiles of one colour only: the deterministic COLORED scatter (assemble.cu).
//            peers   (PEER, multi-GPU) rows of partition-interface nodes are also reduced straight into the neighbouring rank's copy
//                    of the row through a peer-mapped pointer (NVLink): the interface exchange is part of the flush (comm.cu).
// On the structured C3 mesh this is 1.27 CSR updates per value instead of 2.4, 65 % of them plain stores.  Sums differ from the
// per-element kernels by fp reassociation only; (u, v) and (v, u) receive identical per-tile partial sums.
//
// Reference gradients: for the trilinear hexahedron (hexahedron.rs:50-83) d phi_a / d xi = sx_a (1 + sy_a eta)(1 + sz_a zeta) / 8
// depends on the node only through its signs, so a lane keeps 12 table values (P0[sy, sz], P1[sx, sz], P2[sx, sy]) for its
// quadrature point instead of 24 and applies the signs as operand negations.
#pragma once

// a += K, or a = K when this is the first contribution to the accumulator in the tile (tiles.cpp marks it in the element map): the
// accumulators are never cleared between tiles
template <int S>
__device__ __forceinline__ void tile_accumulate(double* __restrict__ a, const double (&K)[S][S], bool first) {
    double v[S * S];
#pragma unroll
    for (int i = 0; i < S * S; ++i) v[i] = first ? 0.0 : a[i];
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) a[i * S + j] = v[i * S + j] + K[i][j];
}

__device__ __forceinline__ void dmma_m8n8k4_zero(double& d0, double& d1, double a, double b) {
    const double z = 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%4};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(z));
}

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// reduction into a peer GPU's memory (NVLink): system scope
__device__ __forceinline__ void red_add_f64_sys(double* addr, double v) {
    asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kTileGroupWarps = 8;   // warps per compute group = elements per round (tiles.cpp: TileShape::warps)
constexpr int kTileHelperWarps = 8;
constexpr int kTileHelperRegs = 48, kTileComputeRegs = 96;  // setmaxnreg: the helpers hand their registers to the compute warps

template <int OP, int MAXN, int MAXP>
struct Hex8TileSmem {
    static constexpr int S = OP == FB200_LAPLACE ? 1 : 3;
    static constexpr int BS = S * S;
    static constexpr int WARPS = 2 * kTileGroupWarps;
    static constexpr int GS = 28;                       // node stride of the transposed gradient array (hex8_mma_kernel.cuh)
    static constexpr int WARP_DOUBLES = 24 + 8 * GS;
    static constexpr int ACC = (MAXP * BS + 1) & ~1;    // one accumulator buffer
    static constexpr int BIG_BYTES = MAXN * 3 * 8 + MAXN * 2 * 8;  // X[MAXN][3] doubles, off[MAXN][2] int64 (ring of 3 tiles)
    static constexpr int SMALL_BYTES = kTileHdrWords * 4 + MAXN * 4;  // header, ids[MAXN] int32 (ring of 8 tiles)
    static constexpr int FROW_BYTES = MAXN * 2 * 8;                // flush row table of one tile
    static constexpr size_t bytes =
        sizeof(double) * (size_t)(2 * ACC + WARPS * WARP_DOUBLES) + 3 * (size_t)BIG_BYTES + 8 * (size_t)SMALL_BYTES + 2 * (size_t)FROW_BYTES + 32;
};

__device__ __forceinline__ void named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_barrier_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// PEER: rows of partition-interface nodes are also reduced into the neighbouring ranks' values (p.peer_row / p.peer_values)
template <int OP, int MAXN, int MAXP, bool PEER = false>
__global__ void __launch_bounds__((2 * kTileGroupWarps + kTileHelperWarps) * 32, 1) assemble_hex8_tile_kernel(const AssembleParams p) {
    using L = Hex8TileSmem<OP, MAXN, MAXP>;
    constexpr int N = 8, D = 3, S = L::S, BS = L::BS, GS = L::GS, WARPS = L::WARPS, GW = kTileGroupWarps;
    constexpr int TC = WARPS * 32;              // compute threads
    constexpr int TH = kTileHelperWarps * 32;   // helper threads
    // named barriers: 0 everybody (prologue), 1 / 2 round hand-over of the compute groups, 3 helpers,
    // 4 / 5 "tile done" (compute arrives, helpers wait; by tile parity), 6 / 7 "accumulators ready" (helpers arrive, compute waits)
    constexpr int BAR_HELPER = 3, BAR_DONE = 4, BAR_READY = 6;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(MAXN <= TH && MAXN <= 128 && GW % 4 == 0 && kTileHelperWarps % 4 == 0, "one helper thread per tile node, 7-bit node index, whole warpgroups");
    extern __shared__ __align__(16) double smem[];
    double* wbase = smem + 2 * L::ACC;
    unsigned char* bufbase = reinterpret_cast<unsigned char*>(wbase + WARPS * L::WARP_DOUBLES);
    auto big_X = [&](int b) { return reinterpret_cast<double*>(bufbase + b * L::BIG_BYTES); };
    auto big_off = [&](int b) { return reinterpret_cast<long long*>(bufbase + b * L::BIG_BYTES + MAXN * 24); };
    unsigned char* smallbase = bufbase + 3 * L::BIG_BYTES;
    auto small_hdr = [&](int sl) { return reinterpret_cast<uint32_t*>(smallbase + sl * L::SMALL_BYTES); };
    auto small_ids = [&](int sl) { return reinterpret_cast<int*>(smallbase + sl * L::SMALL_BYTES + kTileHdrWords * 4); };
    unsigned char* frowbase = smallbase + 8 * L::SMALL_BYTES;
    auto buf_frow = [&](int b) { return reinterpret_cast<long long*>(frowbase + b * L::FROW_BYTES); };
    uint32_t* s_tick = reinterpret_cast<uint32_t*>(frowbase + 2 * L::FROW_BYTES);  // [8]: tile number held by each ring slot
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dbg = p.debug;  // measurement knobs (results are wrong when set): 1 skip compute, 2 skip the global writes, 4 skip accumulate, 8 / 16 / 32 see the flush
    const bool overwrite = p.accumulate == 0;

    for (int i = tid; i < 2 * L::ACC; i += TC + TH) smem[i] = 0.0;

    if (warp >= WARPS) {
        // =========================================================================================== helper warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTileHelperRegs));
        const int ht = tid - TC;
        const uint64_t pol_keep = l2_policy_evict_last();
        // flush lane mapping: a warp handles 32 list entries = SUB x 32 items per step; item 32 m + lane = (entry fe[m], column fj[m])
        constexpr int SUB = S == 1 ? 1 : 3;
        uint32_t fe[SUB];
        int fj[SUB];
#pragma unroll
        for (int m = 0; m < SUB; ++m) {
            fe[m] = S == 1 ? (uint32_t)lane : (uint32_t)(32 * m + lane) / 3u;
            fj[m] = S == 1 ? 0 : (32 * m + lane) % 3;
        }
        // the table pipeline: stage 0 header, stage 1 node ids, stage 2 coordinates + block-row offsets.  A stage reads what the
        // previous one left in shared memory; consecutive stages of a tile run in consecutive iterations (tables_done() between).
        auto stage0 = [&](int sl) {
            const uint32_t tn = s_tick[sl];
            if (tn < p.num_tiles && ht < kTileHdrWords / 4) {
                const uint32_t tile = p.tile_list ? __ldg(p.tile_list + tn) : tn;  // (a launch over the tiles of one colour)
                cp_async_16(small_hdr(sl) + 4 * ht, p.tile_hdr + (size_t)tile * kTileHdrWords + 4 * ht);
            }
        };
        auto stage1 = [&](int sl) {
            if (s_tick[sl] < p.num_tiles) {
                const uint32_t* h = small_hdr(sl);
                if (ht < (int)h[2]) cp_async_4(small_ids(sl) + ht, p.tile_nodes + h[4] + ht);
            }
        };
        auto stage2 = [&](int sl, int bb) {
            if (s_tick[sl] < p.num_tiles && ht < (int)small_hdr(sl)[2]) {
                const int I = small_ids(sl)[ht] & 0x7fffffff;
                const double* v = p.vertices + (uint64_t)I * D;
#pragma unroll
                for (int c = 0; c < D; ++c) cp_async_8(big_X(bb) + ht * D + c, v + c);
                cp_async_8(big_off(bb) + 2 * ht, p.blk_off + I);
                cp_async_8(big_off(bb) + 2 * ht + 1, p.blk_off + I + 1);
            }
        };
        auto tables_done = [&]() {
            cp_async_wait_all();
            named_barrier(BAR_HELPER, TH);
        };
        unsigned int tick_next = 0;
        // tiles are handed out by an atomic ticket (dynamic) or round-robin over the CTAs (p.tile_static: tile = CTA + k * CTAs)
        unsigned int static_next = blockIdx.x;
        auto next_ticket = [&]() -> unsigned int {
            if (p.tile_static) {
                const unsigned int t = static_next;
                static_next += gridDim.x;
                return t;
            }
            return atomicAdd(p.ticket32, 1u);
        };
        if (ht == 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) s_tick[k] = next_ticket();
            tick_next = next_ticket();
        }
        named_barrier(BAR_HELPER, TH);
#pragma unroll
        for (int k = 0; k < 4; ++k) stage0(k);
        tables_done();
#pragma unroll
        for (int k = 0; k < 3; ++k) stage1(k);
        tables_done();
        stage2(0, 0);
        stage2(1, 1);
        tables_done();
        __syncthreads();  // (A) prologue done: tables of tiles 0 and 1 are there, both accumulator buffers are clear
        if (s_tick[0] < p.num_tiles) named_barrier_arrive(BAR_READY + 0, TC + TH);

        // owner stores (see the header comment): flags are compared with this launch's epoch; 0 = the lists carry no ownership or the
        // call accumulates (then every entry is a reduction and nothing is published or waited for)
        const uint32_t epoch = overwrite ? p.tile_epoch : 0u;
        bool gave_up = false;  // this thread has seen a wait time out (or an error reported by anybody): no more waiting
        // iteration `it`: tables of tiles it + 2 .. it + 4, flush of tile it - 1 (after the compute warps have finished it)
        for (uint32_t it = 0;; ++it) {
            const int sl = (int)(it & 7u), b = (int)(it & 1u), nb = b ^ 1;
            if (it > 0) {
                if (s_tick[(it - 1) & 7u] >= p.num_tiles) break;
                named_barrier(BAR_DONE + nb, TC + TH);  // tile it - 1 is complete in accumulator buffer nb
            } else if (s_tick[0] >= p.num_tiles) {
                break;
            }
            const bool valid = s_tick[sl] < p.num_tiles;
            const uint32_t* hdr = small_hdr(sl);
            const int nn = valid ? (int)hdr[2] : 0;
            if (valid) {
                // row table of this tile's flush (used in the next iteration): first value of node u's rows and the row length; for a
                // partition-interface node (PEER) also its block-row offset on the neighbouring rank + 1 | neighbour slot << 31
                if (ht < nn) {
                    const long long* off = big_off((int)(it % 3u));
                    const long long o0 = off[2 * ht], o1 = off[2 * ht + 1];
                    // one 64-bit word per node (a 64-bit shared load costs a quarter of the wavefronts of a 128-bit one): index of the first
                    // value (48 bits) | row length in doubles << 48 (block rows are < 8192 nodes: tiles.cpp); PEER: a second table of words
                    unsigned long long* fr = reinterpret_cast<unsigned long long*>(buf_frow(b));
                    fr[ht] = (unsigned long long)((long long)BS * o0) | ((unsigned long long)(unsigned int)((int)(o1 - o0) * S) << 48);
                    // (only tiles with partition-interface nodes - header flag bit 0 - look the neighbour rows up, here and in the flush)
                    if constexpr (PEER) {
                        if (hdr[11] & 1u) reinterpret_cast<uint32_t*>(fr + MAXN)[ht] = p.peer_row[small_ids(sl)[ht] & 0x7fffffff];
                    }
                }
                // pull this tile's flush list into L2 (it is read in the next iteration)
                const char* f0 = reinterpret_cast<const char*>(p.tile_flush + hdr[5]);
                const uint32_t bytes = hdr[6] * 4u;
                for (uint32_t o = (uint32_t)ht * 128u; o < bytes; o += (uint32_t)TH * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(f0 + o));
            }
            // ---- one stage for each of three coming tiles; the ring slot of tile it - 3 receives the ticket of tile it + 5
            stage2((int)((it + 2) & 7u), (int)((it + 2) % 3u));
            stage1((int)((it + 3) & 7u));
            stage0((int)((it + 4) & 7u));
            if (ht == 0) {
                s_tick[(it + 5) & 7u] = tick_next;
                tick_next = next_ticket();
            }
            // ---- flush of the previous tile (its header is still in ring slot it - 1): every node block goes to the CSR once
            if (it > 0) {
                const uint32_t* ph = small_hdr((int)((it - 1) & 7u));
                const uint32_t* fl = p.tile_flush + ph[5];
                const uint32_t items = ph[6], store_items = ph[8];  // (entries)
                const double* pacc = smem + nb * L::ACC;
                const unsigned long long* prow = reinterpret_cast<const unsigned long long*>(buf_frow(nb));
                const bool tile_iface = PEER && (ph[11] & 1u) != 0u;  // warp-uniform: most tiles have no interface node
                // entries [e0, e1) of the list.  A warp takes blocks of 32 entries = SUB * 32 items; in sub-iteration m lane l handles
                // item 32 m + l of the block = entry fe[m], column fj[m] (per-lane constants: no division in the loop), so that an
                // instruction covers 32 consecutive doubles of a run of neighbouring node blocks.  STORE: plain stores; else reductions
                auto flush_range = [&](const uint32_t e0, const uint32_t e1, auto store_tag) {
                    constexpr bool STORE = decltype(store_tag)::value;
                    constexpr uint32_t STEP = 32u * kTileHelperWarps;
                    uint32_t eb = e0 + 32u * (uint32_t)(ht >> 5);
                    uint32_t w[SUB], wn[SUB];
#pragma unroll
                    for (int m = 0; m < SUB; ++m) w[m] = eb + fe[m] < e1 ? __ldg(fl + eb + fe[m]) : 0xffffffffu;
                    while (eb < e1) {  // warp-uniform
                        const uint32_t en = eb + STEP;
#pragma unroll
                        for (int m = 0; m < SUB; ++m) wn[m] = en + fe[m] < e1 ? __ldg(fl + en + fe[m]) : 0xffffffffu;  // (next block, in flight)
#pragma unroll
                        for (int m = 0; m < SUB; ++m) {
                            const uint32_t a = w[m];
                            if (a == 0xffffffffu) continue;  // past the end of the range (no list word is all ones: u <= 124)
                            const int j = fj[m];
                            const uint32_t un = (a >> 12) & 0x7fu;
                            const unsigned long long rt = prow[un];
                            const int rl = (int)(rt >> 48);
                            const bool tr = (a & 0x800u) != 0u;
                            const uint32_t apos = a & 0x7ffu;
                            const bool zero = apos == kTileZeroPos;  // an owner writes 0.0 where it has no contribution
                            const double* src = pacc + (int)(zero ? 0u : apos) * BS + (tr ? j * S : j);
                            const int sstride = tr ? 1 : S;
                            const int col = S * (int)(a >> 19) + j;
                            double* dst = p.values + ((long long)(rt & 0xffffffffffffull) + (long long)col);
                            double v[S];
#pragma unroll
                            for (int i = 0; i < S; ++i) v[i] = zero ? 0.0 : src[i * sstride];
                            if (dbg & 2) continue;
                            if constexpr (STORE) {
#pragma unroll
                                for (int i = 0; i < S; ++i) dst[(long long)i * rl] = v[i];
                            } else {
                                if (zero) continue;  // (a STORE segment flushed by an accumulating call)
#pragma unroll
                                for (int i = 0; i < S; ++i) red_add_f64_hint(dst + (long long)i * rl, v[i], pol_keep);
                                if constexpr (PEER) {
                                    const uint32_t pw = tile_iface ? reinterpret_cast<const uint32_t*>(prow + MAXN)[un] : 0u;
                                    if (pw) {
                                        double* pd = p.peer_values[pw >> 31] + ((long long)BS * (long long)((pw & 0x7fffffffu) - 1u) + (long long)col);
#pragma unroll
                                        for (int i = 0; i < S; ++i) red_add_f64_sys(pd + (long long)i * rl, v[i]);
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int m = 0; m < SUB; ++m) w[m] = wn[m];
                        eb = en;
                    }
                };
                // STORE segment (nothing of it when the call accumulates): with ownership first the shared rows this tile owns; publish
                // them (barrier: every helper has issued its stores; release store: they are visible before the flag); then the complete
                // rows - while the owners this tile depends on publish theirs - and only then the reductions into other tiles' rows
                // (measurement knobs, results wrong: 8 no publish / no wait, 16 no wait, 32 publish without the barrier)
                const uint32_t n_store = overwrite ? store_items : 0u, n_pub = epoch ? ph[12] : 0u;
                const uint32_t nwait = (!epoch || (dbg & 24)) ? 0u : ph[10];
                const uint32_t* wl = p.tile_wait + ph[9];
                const uint32_t* fp = nullptr;
                uint32_t seen = epoch;
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {
                    flush_range(part ? n_pub : 0u, part ? n_store : n_pub, std::true_type{});
                    if (part == 0 && epoch) {
                        if (!(dbg & 32)) named_barrier(BAR_HELPER, TH);
                        if (ht == 0 && !(dbg & 8)) st_release_u32(p.tile_flag + s_tick[(it - 1) & 7u], epoch);
                        // first poll before the complete rows (its latency hides behind them); lanes beyond the list poll nothing
                        if ((uint32_t)ht < nwait) {
                            fp = p.tile_flag + __ldg(wl + ht);
                            seen = ld_acquire_u32(fp);
                        }
                    }
                }
                {
                    if (nwait) {
                        for (uint32_t w = (uint32_t)ht; w < nwait; w += (uint32_t)TH) {
                            if (w != (uint32_t)ht) {
                                fp = p.tile_flag + wl[w];
                                seen = ld_acquire_u32(fp);
                            }
                            unsigned int spins = 0;
                            // a healthy wait lasts microseconds.  After ~1 s: report and stop waiting - for good, and so does every other
                            // CTA once the error word is set (the result is invalid anyway; the launch must end, not hang the device)
                            while (seen != epoch && !gave_up) {
                                __nanosleep(32);
                                seen = ld_acquire_u32(fp);
                                if (++spins > (1u << 22)) {
                                    flag_error(p.errword, (uint64_t)s_tick[(it - 1) & 7u], FB200_ERR_CUDA);
                                    gave_up = true;
                                } else if ((spins & 1023u) == 0u && *reinterpret_cast<volatile unsigned long long*>(p.errword) != ~0ull) {
                                    gave_up = true;  // somebody else's error (it keeps its own code and element)
                                }
                            }
                            if ((dbg & 64) && spins) {  // diagnostics: blocked waits, their spins, how far back the owner tile is
                                unsigned long long* dc = reinterpret_cast<unsigned long long*>(p.ticket32) + 2;
                                const unsigned long long dist = s_tick[(it - 1) & 7u] - wl[w];
                                atomicAdd(dc + 0, 1ull);
                                atomicAdd(dc + 1, (unsigned long long)spins);
                                atomicMax(dc + 2, (unsigned long long)spins);
                                atomicAdd(dc + 3, dist);
                                atomicMax(dc + 4, dist);
                                atomicAdd(dc + (dist < 8 ? 5 : dist < 64 ? 6 : dist < 512 ? 7 : 8), 1ull);
                                atomicAdd(dc + (dist < 8 ? 9 : dist < 64 ? 10 : dist < 512 ? 11 : 12), (unsigned long long)spins);
                            }
                        }
                        named_barrier(BAR_HELPER, TH);
                    }
                    flush_range(n_store, items, std::false_type{});
                }
            }
            tables_done();  // the stages have landed, and every helper has read its accumulators (no clearing: first contributions store)
            // accumulator buffer nb is clear (again) and the tables of tile it + 2 are there: release tile it + 1
            if (s_tick[(it + 1) & 7u] < p.num_tiles) {
                __threadfence_block();
                named_barrier_arrive(BAR_READY + nb, TC + TH);
            }
        }
        return;
    }

    // =============================================================================================== compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTileComputeRegs));
    const int grp = warp / GW, gwarp = warp - grp * GW;
    double* s_G = wbase + warp * L::WARP_DOUBLES + 24;  // (the first 24 doubles of a warp's block are unused since the coordinates are read in place)
    const int nq = p.nq;  // <= 8
    // lane (gq, gi) owns row gi of the Jacobian at point gq (the 4th lane of a point shadows row 0)
    const int gq = lane >> 2, s4 = lane & 3;
    const int gi = s4 == 3 ? 0 : s4;
    const bool gact = gq < nq;
    double P0[4], P1[4], P2[4];  // index = (second sign > 0) * 2 + (first sign > 0), see the header comment
    double sqw = 0.0;
    {
        const double* tg = p.tab + 3 * nq + (gact ? gq : 0) * (N * D);  // geometry table == basis table for Hex8, node-major
        constexpr int n0[4] = {1, 2, 5, 6}, n1[4] = {3, 2, 7, 6}, n2[4] = {4, 5, 7, 6};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            P0[c] = gact ? tg[n0[c] * D + 0] : 0.0;
            P1[c] = gact ? tg[n1[c] * D + 1] : 0.0;
            P2[c] = gact ? tg[n2[c] * D + 2] : 0.0;
        }
        if (gact) sqw = sqrt(p.tab[gq]);
    }
    const int src1 = (lane & ~3) | (gi == 2 ? 0 : gi + 1), src2 = (lane & ~3) | (gi == 0 ? 2 : gi - 1);
    const int ba = lane >> 2;
    const int frag = ba * GS + 3 * (lane & 3);
    const double mu = p.mu, lam = p.lam;
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint32_t* emap32 = reinterpret_cast<const uint32_t*>(p.tile_emap);

    // element data of a schedule position: tile-local node indices (lanes 0-7, one byte each; byte 0 = 0xff: padding position),
    // accumulator positions of the lane's two blocks
    auto load_elem = [&](uint64_t pos, uint2& ln_, uint32_t& em_) {
        ln_ = __ldg(reinterpret_cast<const uint2*>(p.tile_lnodes) + pos);  // the 8 node bytes, the same 8-byte word for every lane (one transaction)
        em_ = ld_u32_hint(emap32 + pos * 32 + lane, pol_stream);
    };

    __syncthreads();  // (A)
    bool preloaded = false;
    uint2 ln = make_uint2(0xffu, 0u);
    uint32_t em = 0xffffffffu;
    for (uint32_t it = 0;; ++it) {
        const int sl = (int)(it & 7u), b = (int)(it & 1u);
        if (s_tick[sl] >= p.num_tiles) break;
        const uint32_t* hdr = small_hdr(sl);
        const uint32_t p0 = hdr[0];
        const int R = (int)hdr[1];
        const double* tX = big_X((int)(it % 3u));
        double* acc = smem + b * L::ACC;
        if (!preloaded && grp < R) load_elem((uint64_t)p0 + grp * GW + gwarp, ln, em);
        preloaded = false;
        bool ready = false;  // has this warp waited for the helpers to release accumulator buffer b?

        for (int r = grp; r < R; r += 2) {
            const uint32_t em_c = em;
            const bool active = (ln.x & 0xffu) != 0xffu;
            const uint2 ln_c = ln;
            if (r + 2 < R) {
                load_elem((uint64_t)p0 + (r + 2) * GW + gwarp, ln, em);
            } else {
                const int sn = (int)((it + 1) & 7u);
                if (s_tick[sn] < p.num_tiles) {  // the next tile's header landed long ago: this group's first round of it
                    const uint32_t* hn = small_hdr(sn);
                    if (grp < (int)hn[1]) load_elem((uint64_t)hn[0] + grp * GW + gwarp, ln, em);
                    preloaded = true;
                }
            }
            double K0[S][S], K1[S][S];
            if (active && !(dbg & 1)) {
                __syncwarp();  // (every lane has read the previous element's gradients in s_G before anybody overwrites them)
                // ---- geometry: row gi of J at point gq (element/hexahedron.rs:101-107), cofactors, det (elliptic.rs:399-405).  Component
                // gi of the 8 vertices straight from the tile's coordinate table (three distinct addresses per load: broadcast)
                double Jr[D];
                {
                    const double* tXg = tX + gi;
                    auto un = [&](int a) { return (int)(((a < 4 ? ln_c.x : ln_c.y) >> (8 * (a & 3))) & 0x7fu) * D; };
                    const double x0 = tXg[un(0)], x1 = tXg[un(1)], x2 = tXg[un(2)], x3 = tXg[un(3)];
                    const double x4 = tXg[un(4)], x5 = tXg[un(5)], x6 = tXg[un(6)], x7 = tXg[un(7)];
                    Jr[0] = fma(P0[3], x6 - x7, fma(P0[2], x5 - x4, fma(P0[1], x2 - x3, P0[0] * (x1 - x0))));
                    Jr[1] = fma(P1[3], x6 - x5, fma(P1[2], x7 - x4, fma(P1[1], x2 - x1, P1[0] * (x3 - x0))));
                    Jr[2] = fma(P2[3], x6 - x2, fma(P2[2], x7 - x3, fma(P2[1], x5 - x1, P2[0] * (x4 - x0))));
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
                double rs = 0.0;
                if (det != 0.0) {
                    rs = copysign(sqw * rsqrt(fabs(det)), det);
                } else if (gact && s4 == 0) {
                    flag_error(p.errword, (uint64_t)p.tile_elem[(uint64_t)p0 + r * GW + gwarp], FB200_ERR_SINGULAR_JACOBIAN);
                }
#pragma unroll
                for (int j = 0; j < D; ++j) c[j] *= rs;
                if (s4 < 3) {
                    // g_a[gi] = c . grad_ref phi_a,  grad_ref phi_a = (sx P0[sy, sz], sy P1[sx, sz], sz P2[sx, sy])
                    double* go = s_G + 3 * gq + gi;
                    go[0 * GS] = -fma(c[2], P2[0], fma(c[1], P1[0], c[0] * P0[0]));   // node 0 (-,-,-)
                    go[1 * GS] = fma(-c[2], P2[1], fma(-c[1], P1[1], c[0] * P0[0]));  // node 1 (+,-,-)
                    go[2 * GS] = fma(-c[2], P2[3], fma(c[1], P1[1], c[0] * P0[1]));   // node 2 (+,+,-)
                    go[3 * GS] = fma(-c[2], P2[2], fma(c[1], P1[0], -c[0] * P0[1]));  // node 3 (-,+,-)
                    go[4 * GS] = fma(c[2], P2[0], fma(-c[1], P1[2], -c[0] * P0[2]));  // node 4 (-,-,+)
                    go[5 * GS] = fma(c[2], P2[1], fma(-c[1], P1[3], c[0] * P0[2]));   // node 5 (+,-,+)
                    go[6 * GS] = fma(c[2], P2[3], fma(c[1], P1[3], c[0] * P0[3]));    // node 6 (+,+,+)
                    go[7 * GS] = fma(c[2], P2[2], fma(c[1], P1[2], -c[0] * P0[3]));   // node 7 (-,+,+)
                }
                __syncwarp();
                // ---- two node blocks per lane: S = G G^T by DMMA
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
                        for (int n = 0; n < D; ++n) {  // first k-step from a shared zero accumulator (no 18 register clears)
                            dmma_m8n8k4_zero(M0[m][n], M1[m][n], ga[0][m], ga[0][n]);
                            dmma_m8n8k4(M0[m][n], M1[m][n], ga[1][m], ga[1][n]);
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
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i)
#pragma unroll
                    for (int j = 0; j < S; ++j) K0[i][j] = K1[i][j] = 0.0;
            }
            // ---- add the blocks with u_a <= u_b to the tile accumulators, strictly in round order: wait for the other group's
            // round r - 1, add, release its round r + 1
            if (!ready) {
                named_barrier(BAR_READY + b, TC + TH);
                ready = true;
            }
            if (r > 0) named_barrier(1 + grp, TC);
            if (active && !(dbg & 4)) {
                const uint32_t e0 = em_c & 0xffffu, e1 = em_c >> 16;
                if (e0 != 0xffffu) tile_accumulate<S>(acc + (e0 & (kTileFirstTouch - 1u)) * BS, K0, (e0 & kTileFirstTouch) != 0u);
                if (e1 != 0xffffu) tile_accumulate<S>(acc + (e1 & (kTileFirstTouch - 1u)) * BS, K1, (e1 & kTileFirstTouch) != 0u);
            }
            if (r + 1 < R) {
                __threadfence_block();
                named_barrier_arrive(1 + (grp ^ 1), TC);
            }
        }
        if (!ready) named_barrier(BAR_READY + b, TC + TH);  // (a group without a round in this tile)
        __threadfence_block();
        named_barrier_arrive(BAR_DONE + b, TC + TH);  // this warp has added everything it has for tile `it`
    }
}
