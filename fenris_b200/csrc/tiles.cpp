// Host preprocessing for the tile-accumulating Hex8 kernel (hex8_tile_kernel.cuh).
//
// The reference adds every row of every K_e to the CSR separately (global.rs:155-178, 504-537).  On the device each such addition is
// an fp64 reduction in L2 and it is the L2 reduction rate - not HBM, not the FP64 pipe - that bounds the Hex8 assembly
// (profiles/r01/README.md).  Neighbouring elements share most of their node blocks: in a 4 x 4 x 4 block of hexahedra the 64 x 64
// element blocks land in only 2197 distinct CSR node blocks, and 27 of the 125 nodes have ALL their elements inside the block.  The
// lists built here let one CTA sum a tile of consecutive elements of the Morton order in shared memory and touch every CSR node
// block of the tile once - with a plain store when the row node is complete inside the tile.
//
// Per tile (a run of positions of the processing order that share the Morton prefix code >> tile_bits, split further if it would
// exceed the node / accumulator limits):
//   nodes      the distinct nodes, ascending global id (tile-local index u = rank); bit 31 = complete inside the tile
//   slots      one accumulator per node pair (u, v), u <= v, coupled inside the tile: K_e is symmetric block-wise
//              (K_ba = K_ab^T bit for bit, see assemble.cu), so only blocks with u_a <= u_b are accumulated and the flush writes
//              (u, v) and (v, u)^T from the same sums - the assembled matrix is exactly symmetric, as the reference's
//              upper-triangle-then-mirror rule makes it (operators.rs:177-180, util.rs:38-50)
//   positions  accumulator position of a slot, chosen so that the 16 lanes of a half-warp - which hold the blocks (a, b) with
//              a in {0..3} or {4..7} and b in {0,2,4,6} or {1,3,5,7} (the DMMA accumulator fragment) - fall into 16 different
//              shared-memory banks: position mod 16 = 4 alpha(u) + (beta(v) + rot(u)) mod 4, where alpha / beta are 4-colourings
//              of the tile nodes that separate the nodes of every element face {0..3}, {4..7} resp. of {0,2,4,6}, {1,3,5,7}
//              (greedy; on structured meshes they are the coordinate parities), and rot(u) balances the 16 residue classes
//   flush      per (u, v) in CSR order of row u, one word: accumulator position (11 bits) | transposed << 11 | u << 12 | k << 19,
//              k = position of v in the block row of u (from the node-block map; meshes with block rows >= 8192 keep the
//              per-element kernel)
//   schedule   the tile's elements ordered by a greedy node-disjoint colouring (fenris-paradis' idea, coloring.rs:6-70, applied
//              inside the tile) and cut into ROUNDS of at most `warps` elements of one colour (padded to `warps` schedule
//              positions).  The kernel's two groups of compute warps take the rounds alternately and add their blocks strictly
//              in round order (named-barrier hand-over), so no two warps ever update the same accumulator concurrently
#include <algorithm>
#include <atomic>
#include <thread>

#include "fb200_internal.h"

namespace fb200 {

namespace {

struct TileOut {
    uint64_t p0 = 0;
    int ne = 0;
    uint32_t P = 0, rounds = 0;
    std::vector<int32_t> nodes;
    std::vector<uint32_t> flush;
    std::vector<uint8_t> lnodes;   // rounds * warps * 8
    std::vector<uint16_t> emap;    // rounds * warps * 64
    std::vector<int32_t> elem;     // rounds * warps
};

struct Builder {
    const TileShape& shape;
    const int32_t* order;
    const int32_t* conn;
    const uint16_t* blockmap;
    const int32_t* degree;
    HostTiles& out;
    uint64_t conflicts = 0, accesses = 0;
    bool degenerate = false;

    struct RowEntry {
        uint16_t k;
        uint8_t v;
        uint16_t pos;  // accumulator position (v >= u only)
    };
    std::vector<int32_t> nodes;
    std::vector<int> inc;
    std::vector<std::vector<RowEntry>> rows;
    std::vector<std::vector<int>> node_elems;
    std::vector<uint8_t> ln;  // ne * 8

    uint64_t elem_at(uint64_t pos) const { return order ? (uint64_t)order[pos] : pos; }

    // 4-colouring of the tile nodes that separates the members of every group (two groups of 4 local nodes per element)
    void label(int ne, const int (&groups)[2][4], std::vector<uint8_t>& lab) const {
        const int nn = (int)nodes.size();
        lab.assign(nn, 0xff);
        for (int u = 0; u < nn; ++u) {
            int used[4] = {0, 0, 0, 0};
            for (int el : node_elems[u])
                for (int g = 0; g < 2; ++g) {
                    bool member = false;
                    for (int t = 0; t < 4; ++t) member |= ln[el * 8 + groups[g][t]] == u;
                    if (!member) continue;
                    for (int t = 0; t < 4; ++t) {
                        const int w = ln[el * 8 + groups[g][t]];
                        if (w != u && lab[w] != 0xff) ++used[lab[w]];
                    }
                }
            int best = 0;
            for (int c = 1; c < 4; ++c)
                if (used[c] < used[best]) best = c;
            lab[u] = (uint8_t)best;
        }
        (void)ne;
    }

    // Flush reads (hex8_tile_kernel.cuh): item 3 e + j of the list is lane (entry e, column j); it reads the words 9 pos + 3 i + j
    // (transposed: 9 pos + 3 j + i) for i = 0, 1, 2 in three loads, 16 lanes = 16 eight-byte banks per half-warp.  Entries of one row
    // share few bank classes (pos mod 16 = 4 alpha(u) + ...), so consecutive entries collide.  With a rotation r_e the lane reads row
    // (i + r_e) mod 3 in load i: chosen greedily, entry by entry, to minimise the bank maxima of the half-warps the entry touches.
    static void rotate_flush_rows(std::vector<uint32_t>& flush) {
        const size_t nf = flush.size();
        const size_t halves = (3 * nf + 15) / 16;
        std::vector<uint8_t> count(halves * 3 * 16, 0);  // [half-warp][load][bank]
        std::vector<uint8_t> peak(halves * 3, 0);
        for (size_t e = 0; e < nf; ++e) {
            const uint32_t a = flush[e];
            const int pos = (int)(a & 0x7ffu);
            const bool tr = (a >> 11) & 1u;
            int best = 0, best_cost = 1 << 30;
            for (int r = 0; r < 3; ++r) {
                int cost = 0;
                for (int i = 0; i < 3; ++i) {
                    const int ii = (i + r) % 3;
                    uint8_t add[2][16] = {{0}};  // the entry's lanes may straddle two half-warps
                    const size_t h0 = (3 * e) / 16;
                    for (int j = 0; j < 3; ++j) {
                        const size_t h = (3 * e + j) / 16;
                        const int word = pos * 9 + (tr ? j * 3 + ii : ii * 3 + j);
                        ++add[h - h0][word & 15];
                    }
                    for (int hh = 0; hh < 2; ++hh) {
                        if (h0 + hh >= halves) continue;
                        int m = peak[(h0 + hh) * 3 + i];
                        for (int b = 0; b < 16; ++b)
                            if (add[hh][b]) m = std::max<int>(m, count[((h0 + hh) * 3 + i) * 16 + b] + add[hh][b]);
                        cost += m - peak[(h0 + hh) * 3 + i];
                    }
                }
                if (cost < best_cost) {
                    best_cost = cost;
                    best = r;
                }
            }
            for (int i = 0; i < 3; ++i) {
                const int ii = (i + best) % 3;
                for (int j = 0; j < 3; ++j) {
                    const size_t h = (3 * e + j) / 16;
                    const int word = pos * 9 + (tr ? j * 3 + ii : ii * 3 + j);
                    uint8_t& c = count[(h * 3 + i) * 16 + (word & 15)];
                    ++c;
                    peak[h * 3 + i] = std::max(peak[h * 3 + i], c);
                }
            }
            flush[e] = a | ((uint32_t)best << 30);
        }
    }

    bool build_one(uint64_t p0, int ne, TileOut& t) {
        constexpr int n = 8, n2 = 64;
        // ---- nodes
        nodes.clear();
        for (int el = 0; el < ne; ++el) {
            const uint64_t e = elem_at(p0 + el);
            for (int a = 0; a < n; ++a) nodes.push_back(conn[e * n + a]);
        }
        std::sort(nodes.begin(), nodes.end());
        inc.clear();
        {
            size_t w = 0;
            for (size_t i = 0; i < nodes.size();) {
                size_t j = i;
                while (j < nodes.size() && nodes[j] == nodes[i]) ++j;
                nodes[w++] = nodes[i];
                inc.push_back((int)(j - i));
                i = j;
            }
            nodes.resize(w);
        }
        const int nn = (int)nodes.size();
        if (nn > shape.max_nodes) return false;
        ln.assign((size_t)ne * n, 0);
        node_elems.assign(nn, {});
        for (int el = 0; el < ne; ++el) {
            const uint64_t e = elem_at(p0 + el);
            for (int a = 0; a < n; ++a) {
                const int u = (int)(std::lower_bound(nodes.begin(), nodes.end(), conn[e * n + a]) - nodes.begin());
                ln[el * n + a] = (uint8_t)u;
                if (node_elems[u].empty() || node_elems[u].back() != el) node_elems[u].push_back(el);
                else degenerate = true;  // a node repeated inside one element: two lanes would share an accumulator
            }
        }
        if (degenerate) return true;
        // ---- rows: coupled nodes of every tile node, ordered by their position k in the global block row
        rows.assign(nn, {});
        for (int el = 0; el < ne; ++el) {
            const uint64_t e = elem_at(p0 + el);
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) rows[ln[el * n + a]].push_back({blockmap[e * n2 + a * n + b], ln[el * n + b], 0});
        }
        int nslots = 0, nflush = 0;
        for (int u = 0; u < nn; ++u) {
            auto& r = rows[u];
            std::sort(r.begin(), r.end(), [](const RowEntry& x, const RowEntry& y) { return x.k < y.k; });
            r.erase(std::unique(r.begin(), r.end(), [](const RowEntry& x, const RowEntry& y) { return x.k == y.k; }), r.end());
            nflush += (int)r.size();
            for (const RowEntry& x : r) nslots += x.v >= u;
        }
        if (nslots > shape.max_slots) return false;
        // ---- bank-aware accumulator positions
        static const int kFaces[2][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}};
        static const int kStripes[2][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}};
        std::vector<uint8_t> alpha, beta;
        label(ne, kFaces, alpha);
        label(ne, kStripes, beta);
        const int cap = shape.max_slots / 16;
        int fill[16] = {0};
        for (int u = 0; u < nn; ++u) {
            int c[4] = {0, 0, 0, 0};
            for (const RowEntry& x : rows[u])
                if (x.v >= u) ++c[beta[x.v]];
            int best_rot = 0, best_max = 1 << 30;
            for (int rot = 0; rot < 4; ++rot) {
                int m = 0;
                for (int b = 0; b < 4; ++b) m = std::max(m, fill[4 * alpha[u] + ((b + rot) & 3)] + c[b]);
                if (m < best_max) {
                    best_max = m;
                    best_rot = rot;
                }
            }
            for (RowEntry& x : rows[u]) {
                if (x.v < u) continue;
                int r = 4 * alpha[u] + ((beta[x.v] + best_rot) & 3);
                if (fill[r] >= cap) {  // class full: take the emptiest one (costs a bank conflict, never correctness)
                    r = 0;
                    for (int q = 1; q < 16; ++q)
                        if (fill[q] < fill[r]) r = q;
                }
                x.pos = (uint16_t)(r + 16 * fill[r]++);
            }
        }
        int maxfill = 0;
        for (int q = 0; q < 16; ++q) maxfill = std::max(maxfill, fill[q]);
        t.P = (uint32_t)(16 * maxfill);
        auto find_pos = [&](int u, uint16_t k) -> const RowEntry& {
            const auto& r = rows[u];
            return *std::lower_bound(r.begin(), r.end(), k, [](const RowEntry& x, uint16_t kk) { return x.k < kk; });
        };
        // ---- flush list + node list
        t.p0 = p0;
        t.ne = ne;
        t.nodes.resize(nn);
        for (int u = 0; u < nn; ++u) t.nodes[u] = nodes[u] | (inc[u] == degree[nodes[u]] ? (int32_t)0x80000000 : 0);
        t.flush.clear();
        t.flush.reserve(nflush);
        for (int u = 0; u < nn; ++u)
            for (const RowEntry& x : rows[u]) {
                uint32_t pos, tr = 0;
                if (x.v >= u) {
                    pos = x.pos;
                } else {  // mirrored block: find (v, u) in row v
                    pos = 0xffffu;
                    for (const RowEntry& y : rows[x.v])
                        if (y.v == u) {
                            pos = y.pos;
                            break;
                        }
                    tr = 1;
                }
                if (x.k >= (1u << (shape.flush_rot ? kTileKBitsRot : kTileKBits))) degenerate = true;
                t.flush.push_back(pos | (tr << 11) | ((uint32_t)u << 12) | ((uint32_t)x.k << 19));
            }
        if (shape.flush_rot) rotate_flush_rows(t.flush);
        // ---- schedule: greedy node-disjoint colouring inside the tile; every colour class is cut into rounds of at most
        // `warps` elements; rounds are padded to `warps` schedule positions (padding: node byte 0 = 0xff, no accumulators)
        std::vector<uint64_t> node_mask(nn, 0);
        std::vector<int> colour(ne), sched(ne);
        for (int el = 0; el < ne; ++el) {
            uint64_t used = 0;
            for (int a = 0; a < n; ++a) used |= node_mask[ln[el * n + a]];
            int c = 0;
            while (c < 63 && ((used >> c) & 1)) ++c;
            colour[el] = c;
            for (int a = 0; a < n; ++a) node_mask[ln[el * n + a]] |= 1ull << c;
        }
        for (int el = 0; el < ne; ++el) sched[el] = el;
        std::stable_sort(sched.begin(), sched.end(), [&](int x, int y) { return colour[x] < colour[y]; });
        const int gw = shape.warps;
        t.lnodes.clear();
        t.emap.clear();
        t.elem.clear();
        t.rounds = 0;
        auto pad_round = [&]() {
            while (t.elem.size() % gw) {
                t.elem.push_back(-1);
                t.lnodes.insert(t.lnodes.end(), n, (uint8_t)0);
                t.lnodes[t.lnodes.size() - n] = 0xff;
                t.emap.insert(t.emap.end(), n2, (uint16_t)0xffffu);
            }
        };
        for (int s = 0; s < ne; ++s) {
            const int el = sched[s];
            if (s > 0 && colour[el] != colour[sched[s - 1]]) pad_round();
            const uint64_t e = elem_at(p0 + el);
            t.elem.push_back((int32_t)e);
            for (int a = 0; a < n; ++a) t.lnodes.push_back(ln[el * n + a]);
            const size_t eb = t.emap.size();
            t.emap.resize(eb + n2);
            uint16_t* em = &t.emap[eb];
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) {
                    const int u = ln[el * n + a], v = ln[el * n + b];
                    em[a * n + b] = u <= v ? find_pos(u, blockmap[e * n2 + a * n + b]).pos : (uint16_t)0xffffu;
                }
            // diagnostic: bank collisions of the four half-warp access groups
            for (int h = 0; h < 2; ++h)
                for (int r = 0; r < 2; ++r) {
                    int seen[16] = {0};
                    for (int a = 4 * h; a < 4 * h + 4; ++a)
                        for (int tq = 0; tq < 4; ++tq) {
                            const uint16_t pos = em[a * n + 2 * tq + r];
                            if (pos == 0xffffu) continue;
                            ++accesses;
                            if (seen[pos & 15]++) ++conflicts;
                        }
                }
        }
        pad_round();
        t.rounds = (uint32_t)(t.elem.size() / gw);
        return true;
    }
};

}  // namespace

void build_tile_lists(const TileShape& shape, uint64_t count, const int32_t* order, const uint64_t* codes, const int32_t* conn,
                      uint64_t num_elements, uint64_t num_nodes, const uint16_t* blockmap, HostTiles& out) {
    constexpr int n = 8;
    // incidences of every node over ALL elements of the space - also the ghost elements of a partition, which are in the pattern
    // but are not assembled here: a node is complete only if every element it belongs to is processed inside one tile (then
    // the flush writes every entry of its rows)
    std::vector<int32_t> degree(num_nodes, 0);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) ++degree[conn[e * n + a]];
    out.hdr.clear();
    out.nodes.clear();
    out.flush.clear();
    out.lnodes.clear();
    out.emap.clear();
    out.elem.clear();
    out.bank_conflict_share = 0;
    // candidate tiles: runs with a common Morton prefix
    std::vector<std::pair<uint64_t, int>> cand;
    for (uint64_t p0 = 0; p0 < count;) {
        int ne = 1;
        while (ne < shape.max_elems && p0 + ne < count && (codes[p0 + ne] >> shape.tile_bits) == (codes[p0] >> shape.tile_bits)) ++ne;
        cand.emplace_back(p0, ne);
        p0 += ne;
    }
    std::vector<std::vector<TileOut>> results(cand.size());
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> conflicts{0}, accesses{0};
    std::atomic<bool> degenerate{false};
    auto worker = [&]() {
        Builder b{shape, order, conn, blockmap, degree.data(), out};
        std::vector<std::pair<uint64_t, int>> stack;
        for (;;) {
            const uint64_t c0 = next.fetch_add(64);
            if (c0 >= cand.size() || degenerate.load()) break;
            for (uint64_t c = c0; c < std::min<uint64_t>(cand.size(), c0 + 64); ++c) {
                stack.assign(1, cand[c]);
                while (!stack.empty()) {
                    const auto [p0, ne] = stack.back();
                    stack.pop_back();
                    TileOut t;
                    if (b.build_one(p0, ne, t)) {
                        if (b.degenerate) break;
                        results[c].push_back(std::move(t));
                    } else {  // over the limits: halve (a single Hex8 element always fits: 8 nodes, 36 slots)
                        const int h = ne / 2;
                        stack.emplace_back(p0 + h, ne - h);
                        stack.emplace_back(p0, h);
                    }
                }
                if (b.degenerate) {
                    degenerate.store(true);
                    break;
                }
            }
        }
        conflicts += b.conflicts;
        accesses += b.accesses;
    };
    const unsigned hw = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    const unsigned nthreads = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, cand.size() / 64));
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    if (degenerate.load()) {  // elements with repeated nodes (or huge block rows): the caller keeps the per-element kernel
        out.hdr.clear();
        out.bank_conflict_share = -1.0;
        return;
    }
    for (auto& rs : results)
        for (TileOut& t : rs) {
            const uint32_t hdr[kTileHdrWords] = {(uint32_t)out.elem.size(), t.rounds,                    (uint32_t)t.nodes.size(),   t.P,
                                                 (uint32_t)out.nodes.size(), (uint32_t)out.flush.size(), (uint32_t)t.flush.size(), (uint32_t)t.ne};
            out.hdr.insert(out.hdr.end(), hdr, hdr + kTileHdrWords);
            out.nodes.insert(out.nodes.end(), t.nodes.begin(), t.nodes.end());
            out.flush.insert(out.flush.end(), t.flush.begin(), t.flush.end());
            out.lnodes.insert(out.lnodes.end(), t.lnodes.begin(), t.lnodes.end());
            out.emap.insert(out.emap.end(), t.emap.begin(), t.emap.end());
            out.elem.insert(out.elem.end(), t.elem.begin(), t.elem.end());
        }
    out.bank_conflict_share = accesses.load() ? (double)conflicts.load() / (double)accesses.load() : 0.0;
}

}  // namespace fb200

// ---------------------------------------------------------------------------------------------------------------- host self check
extern "C" fb200_status fb200_tile_lists_selftest(uint64_t num_nodes, const double* vertices, uint64_t num_elements, const uint64_t* connectivity,
                                                  uint64_t num_owned, uint64_t stats[8], int32_t* failed_check) {
    uint64_t st[10];
    const fb200_status s = fb200_tile_lists_selftest_ex(num_nodes, vertices, num_elements, connectivity, num_owned, 0, st, failed_check);
    if (s == FB200_OK && stats) std::copy(st, st + 8, stats);
    return s;
}

extern "C" fb200_status fb200_tile_lists_selftest_ex(uint64_t num_nodes, const double* vertices, uint64_t num_elements,
                                                     const uint64_t* connectivity, uint64_t num_owned, int32_t flush_rot, uint64_t stats[10],
                                                     int32_t* failed_check) {
    using namespace fb200;
    constexpr int n = 8, n2 = 64;
    if (failed_check) *failed_check = 0;
    if (!vertices || !connectivity || !stats || num_owned > num_elements) return FB200_ERR_SHAPE;
    for (uint64_t i = 0; i < num_elements * n; ++i)
        if (connectivity[i] >= num_nodes) return FB200_ERR_INDEX_OOB;
    auto fail_check = [&](int id) {
        if (failed_check) *failed_check = id;
        return FB200_ERR_STATE;
    };
    // node-block rows (sorted coupled nodes, global.rs:65-120) and the block map: position of node b in the block row of node a
    std::vector<std::vector<int32_t>> rows(num_nodes);
    std::vector<int32_t> conn(num_elements * n);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) {
            conn[e * n + a] = (int32_t)connectivity[e * n + a];
            for (int b = 0; b < n; ++b) rows[connectivity[e * n + a]].push_back((int32_t)connectivity[e * n + b]);
        }
    for (auto& r : rows) {
        std::sort(r.begin(), r.end());
        r.erase(std::unique(r.begin(), r.end()), r.end());
        if (r.size() >= 65536) return FB200_ERR_UNSUPPORTED;
    }
    std::vector<uint16_t> map(num_elements * (uint64_t)n2);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) {
            const auto& r = rows[conn[e * n + a]];
            for (int b = 0; b < n; ++b) map[e * n2 + a * n + b] = (uint16_t)(std::lower_bound(r.begin(), r.end(), conn[e * n + b]) - r.begin());
        }
    // processing order of the owned elements, as fb200_space_upload / fb200_set_num_owned_elements build it
    std::vector<int32_t> order_all, order;
    std::vector<uint64_t> codes_all, codes;
    morton_order(3, n, num_nodes, vertices, num_elements, connectivity, order_all, codes_all);
    for (size_t i = 0; i < order_all.size(); ++i)
        if ((uint64_t)order_all[i] < num_owned) {
            order.push_back(order_all[i]);
            codes.push_back(codes_all[i]);
        }
    TileShape shape{6, 64, 8, 128, 1216};
    shape.flush_rot = flush_rot ? 1 : 0;
    const int kbits = shape.flush_rot ? kTileKBitsRot : kTileKBits;
    uint64_t flush_loads = 0, flush_wavefronts = 0;  // model of the flush's shared-memory reads (see Builder::rotate_flush_rows)
    HostTiles ht;
    build_tile_lists(shape, order.size(), order.data(), codes.data(), conn.data(), num_elements, num_nodes, map.data(), ht);
    if (ht.bank_conflict_share < 0.0) return FB200_ERR_UNSUPPORTED;
    // ---- checks
    std::vector<int32_t> degree(num_nodes, 0);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) ++degree[conn[e * n + a]];
    std::vector<uint8_t> seen(num_elements, 0);
    const size_t ntiles = ht.hdr.size() / kTileHdrWords;
    uint64_t max_nodes = 0, max_P = 0, complete = 0, max_rounds = 0, scheduled = 0;
    for (size_t t = 0; t < ntiles; ++t) {
        const uint32_t* h = &ht.hdr[t * kTileHdrWords];
        const uint32_t p0 = h[0], R = h[1], nn = h[2], P = h[3], nb = h[4], fb = h[5], nf = h[6], ne = h[7];
        max_nodes = std::max<uint64_t>(max_nodes, nn);
        max_P = std::max<uint64_t>(max_P, P);
        max_rounds = std::max<uint64_t>(max_rounds, R);
        if (nn > (uint32_t)shape.max_nodes || P > (uint32_t)shape.max_slots || (P & 15u) || ne > (uint32_t)shape.max_elems || R > 255) return fail_check(1);
        if ((uint64_t)p0 + (uint64_t)R * shape.warps > ht.elem.size() || (uint64_t)nb + nn > ht.nodes.size() || (uint64_t)fb + nf > ht.flush.size())
            return fail_check(2);
        // nodes ascending, complete flags
        std::vector<int> inc(nn, 0);
        for (uint32_t u = 0; u < nn; ++u)
            if (u && (ht.nodes[nb + u] & 0x7fffffff) <= (ht.nodes[nb + u - 1] & 0x7fffffff)) return fail_check(3);
        // schedule: every position is an element of the tile or padding; rounds are node-disjoint
        std::vector<std::pair<uint32_t, uint32_t>> pairs;  // (u << 8 | v) -> accumulator position, from the element maps
        std::vector<uint32_t> pair_pos;
        uint32_t real = 0;
        for (uint32_t r = 0; r < R; ++r) {
            std::vector<uint8_t> used(nn, 0);
            for (int w = 0; w < shape.warps; ++w) {
                const uint64_t pos = (uint64_t)p0 + r * shape.warps + w;
                const int32_t e = ht.elem[pos];
                const uint8_t* ln = &ht.lnodes[pos * n];
                const uint16_t* em = &ht.emap[pos * n2];
                if (e < 0) {
                    if (ln[0] != 0xff) return fail_check(4);
                    for (int k = 0; k < n2; ++k)
                        if (em[k] != 0xffffu) return fail_check(4);
                    continue;
                }
                if ((uint64_t)e >= num_owned || seen[e]) return fail_check(5);
                seen[e] = 1;
                ++real;
                for (int a = 0; a < n; ++a) {
                    if (ln[a] >= nn || (ht.nodes[nb + ln[a]] & 0x7fffffff) != conn[(uint64_t)e * n + a]) return fail_check(6);
                    if (used[ln[a]]) return fail_check(7);  // two elements of a round share a node (or an element repeats one)
                    used[ln[a]] = 1;
                    ++inc[ln[a]];
                }
                for (int a = 0; a < n; ++a)
                    for (int b = 0; b < n; ++b) {
                        const uint16_t ps = em[a * n + b];
                        if (ln[a] > ln[b]) {
                            if (ps != 0xffffu) return fail_check(8);
                            continue;
                        }
                        if (ps >= P) return fail_check(8);
                        pairs.emplace_back(((uint32_t)ln[a] << 8) | ln[b], ps);
                    }
            }
        }
        if (real != ne) return fail_check(9);
        scheduled += (uint64_t)R * shape.warps;
        // a node pair always maps to the same accumulator, and different pairs to different accumulators
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        std::vector<uint8_t> taken(P, 0);
        for (size_t i = 0; i < pairs.size(); ++i) {
            if (i && pairs[i].first == pairs[i - 1].first) return fail_check(10);
            if (taken[pairs[i].second]) return fail_check(11);
            taken[pairs[i].second] = 1;
        }
        // flush list: every coupled ordered pair (u, v) exactly once, in CSR order of row u, with the position of v in u's block row
        if (nf != 2 * pairs.size() - [&] { size_t d = 0; for (auto& pr : pairs) d += (pr.first >> 8) == (pr.first & 0xffu); return d; }()) return fail_check(12);
        uint32_t last_u = 0, last_k = 0;
        for (uint32_t f = 0; f < nf; ++f) {
            const uint32_t w = ht.flush[fb + f];
            const uint32_t ps = w & 0x7ffu, tr = (w >> 11) & 1u, u = (w >> 12) & 0x7fu, k = (w >> 19) & ((1u << kbits) - 1u);
            if (u >= nn) return fail_check(13);
            if ((w >> 19) >> kbits >= 3u || (!shape.flush_rot && (w >> 19) >> kbits)) return fail_check(21);  // rotation 0..2, none without the flag
            if (f && (u < last_u || (u == last_u && k <= last_k))) return fail_check(14);
            last_u = u;
            last_k = k;
            const auto& r = rows[ht.nodes[nb + u] & 0x7fffffff];
            if (k >= r.size()) return fail_check(15);
            const int32_t vg = r[k];
            const auto it = std::lower_bound(ht.nodes.begin() + nb, ht.nodes.begin() + nb + nn, vg, [](int32_t x, int32_t y) { return (x & 0x7fffffff) < y; });
            if (it == ht.nodes.begin() + nb + nn || (*it & 0x7fffffff) != vg) return fail_check(16);
            const uint32_t v = (uint32_t)(it - (ht.nodes.begin() + nb));
            if (tr != (u > v ? 1u : 0u)) return fail_check(17);
            const uint32_t key = ((std::min(u, v)) << 8) | std::max(u, v);
            const auto pit = std::lower_bound(pairs.begin(), pairs.end(), std::make_pair(key, 0u));
            if (pit == pairs.end() || pit->first != key || pit->second != ps) return fail_check(18);
        }
        // the three 64-bit loads of every 32-item group: two half-warps each, max distinct words per 8-byte bank
        for (uint32_t g = 0; g < 3 * nf; g += 16)
            for (int i = 0; i < 3; ++i) {
                int words[16], cnt = 0, worst = 0;
                for (uint32_t it = g; it < std::min(g + 16, 3 * nf); ++it) {
                    const uint32_t w = ht.flush[fb + it / 3];
                    const int j = (int)(it % 3), ii = (i + (int)(shape.flush_rot ? w >> 30 : 0)) % 3;
                    words[cnt++] = (int)(w & 0x7ffu) * 9 + (((w >> 11) & 1u) ? j * 3 + ii : ii * 3 + j);
                }
                for (int b = 0; b < 16; ++b) {
                    int m = 0;
                    for (int x = 0; x < cnt; ++x) {
                        bool first = (words[x] & 15) == b;
                        for (int y = 0; first && y < x; ++y) first = words[y] != words[x];
                        m += first;
                    }
                    worst = std::max(worst, m);
                }
                flush_wavefronts += worst;
                flush_loads += (g % 32 == 0);
            }
        for (uint32_t u = 0; u < nn; ++u) {
            const bool flag = ht.nodes[nb + u] < 0;
            if (flag != (inc[u] == degree[ht.nodes[nb + u] & 0x7fffffff])) return fail_check(19);
            complete += flag;
        }
    }
    for (uint64_t e = 0; e < num_owned; ++e)
        if (!seen[e]) return fail_check(20);
    stats[0] = ntiles;
    stats[1] = max_nodes;
    stats[2] = max_P;
    stats[3] = ht.flush.size();
    stats[4] = complete;
    stats[5] = (uint64_t)(ht.bank_conflict_share * 1e6);
    stats[6] = scheduled;
    stats[7] = max_rounds;
    stats[8] = flush_loads;
    stats[9] = flush_wavefronts;
    return FB200_OK;
}
