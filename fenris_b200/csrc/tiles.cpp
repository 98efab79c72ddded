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
//              per-element kernel).  The list has two segments: STORE (rows this tile writes with plain stores when the call
//              overwrites) and REDUCE (rows it adds to with reductions).
//   ownership  (TileShape::owner_stores) the reference's assemble() starts from zeroed values (global.rs:124-131); zero-filling 3.9 GB
//              and then reducing into it costs a full extra pass over HBM.  Instead every row has exactly one STORING tile: a node
//              whose elements all lie in one tile is complete there; a node shared by several tiles is owned by the lowest-numbered
//              one, which stores ALL entries of its rows (those it has no contribution for as 0.0: position kTileZeroPos) and
//              publishes a per-tile flag; the other tiles list the owners they depend on (`wait`) and reduce once those flags are
//              up.  Tiles are handed out in index order, so a tile only ever waits for tiles that are already running: no
//              deadlock, no co-residency requirement.  Rows of nodes that ghost elements of a partition touch are never stored
//              (another rank adds to them as well): they and the rows no owned element touches are the `zero_nodes` the caller
//              clears - a fraction of a percent of the values.
//   schedule   the tile's elements ordered by a greedy node-disjoint colouring (fenris-paradis' idea, coloring.rs:6-70, applied
//              inside the tile) and cut into ROUNDS of at most `warps` elements of one colour (padded to `warps` schedule
//              positions).  The kernel's two groups of compute warps take the rounds alternately and add their blocks strictly
//              in round order (named-barrier hand-over), so no two warps ever update the same accumulator concurrently
#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>

#include "fb200_internal.h"

namespace fb200 {

namespace {

struct TileOut {
    uint64_t p0 = 0;
    int ne = 0;
    uint32_t P = 0, rounds = 0;
    std::vector<int32_t> nodes;
    std::vector<uint32_t> flush;   // pass 1: CSR order; pass 2 (finish_flush): STORE segment, then REDUCE segment
    std::vector<uint32_t> wait;    // pass 2
    uint32_t n_store = 0, n_publish = 0, flags = 0, zero_entries = 0;
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
    double sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // FB200_DEBUG_SETUP: seconds per section of build_one

    struct RowEntry {
        uint16_t k;
        uint8_t v;
        uint16_t pos;  // accumulator position (v >= u only)
        uint16_t src;  // el * 64 + a * 8 + b: the block-map entry of the tile that reported k
    };
    // Everything a tile's lists hold except the global numbers (node ids, element ids, the positions k in the block rows) is a function of
    // its LOCAL connectivity `ln` alone - `nodes` is sorted, so the local index is monotone in the global id and the CSR order of a row is
    // the order of the local column indices.  On a structured mesh nearly all tiles share one local connectivity: the lists of a
    // connectivity seen before are copied and re-labelled instead of being rebuilt (byte-identical result: FB200_DEBUG_SETUP checksum;
    // FB200_TILE_MEMO=0 switches the table off).
    struct Memo {
        uint64_t hash;
        int ne, nn;
        std::vector<uint8_t> ln;
        uint32_t P, rounds;
        std::vector<uint32_t> flush;      // pos | tr << 11 | u << 12, k = 0
        std::vector<uint16_t> flush_src;  // where k comes from (RowEntry::src)
        std::vector<uint8_t> lnodes;
        std::vector<uint16_t> emap;
        std::vector<int8_t> elem;  // tile-local element of every schedule position, -1 = padding
        uint64_t conflicts, accesses;
    };
    std::vector<Memo> memo;
    uint64_t memo_hits = 0;
    const bool memo_on = [this] {
        const char* e = std::getenv("FB200_TILE_MEMO");
        return shape.memo != 0 && !(e && e[0] == '0');
    }();
    static constexpr size_t kMemoCap = 64;

    // lists of a known local connectivity; false: not applicable (the caller builds them)
    bool from_memo(const Memo& m, uint64_t p0, int ne, TileOut& t) {
        constexpr int n2 = 64;
        const size_t nf = m.flush.size();
        t.flush.resize(nf);
        uint32_t prev_u = 0xffffffffu, prev_k = 0;
        for (size_t i = 0; i < nf; ++i) {
            const uint32_t src = m.flush_src[i];
            const uint32_t k = blockmap[elem_at(p0 + (src >> 6)) * n2 + (src & 63u)];
            const uint32_t u = (m.flush[i] >> 12) & 0x7fu;
            if (u == prev_u && k <= prev_k) return false;  // not the CSR order of this row: cannot happen with sorted block rows - rebuild
            prev_u = u;
            prev_k = k;
            if (k >= (1u << kTileKBits)) degenerate = true;
            t.flush[i] = m.flush[i] | (k << 19);
        }
        const int nn = (int)nodes.size();
        t.p0 = p0;
        t.ne = ne;
        t.P = m.P;
        t.rounds = m.rounds;
        t.nodes.resize(nn);
        for (int u = 0; u < nn; ++u) t.nodes[u] = nodes[u] | (inc[u] == degree[nodes[u]] ? (int32_t)0x80000000 : 0);
        t.lnodes = m.lnodes;
        t.emap = m.emap;
        t.elem.resize(m.elem.size());
        for (size_t i = 0; i < m.elem.size(); ++i) t.elem[i] = m.elem[i] < 0 ? -1 : (int32_t)elem_at(p0 + (uint64_t)m.elem[i]);
        conflicts += m.conflicts;
        accesses += m.accesses;
        ++memo_hits;
        return true;
    }
    std::vector<int32_t> nodes;
    std::vector<int> inc;
    std::vector<std::vector<RowEntry>> rows;
    std::vector<std::vector<int>> node_elems;
    std::vector<uint8_t> ln;  // ne * 8
    std::vector<uint16_t> flush_src;
    std::vector<uint16_t> pair_pos = std::vector<uint16_t>(128 * 128, (uint16_t)0xffffu);  // accumulator position of the pair (u, v), u <= v
    std::vector<uint8_t> pair_seen = std::vector<uint8_t>(128 * 128, (uint8_t)0);           // (u, v) already has a row entry

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

    bool build_one(uint64_t p0, int ne, TileOut& t) {
        constexpr int n = 8, n2 = 64;
        auto tp = std::chrono::steady_clock::now();
        auto lap = [&](int k) {
            const auto t1 = std::chrono::steady_clock::now();
            sec[k] += std::chrono::duration<double>(t1 - tp).count();
            tp = t1;
        };
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
        if ((int)node_elems.size() < nn) node_elems.resize(nn);  // (inner vectors keep their capacity from tile to tile)
        for (int u = 0; u < nn; ++u) node_elems[u].clear();
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
        uint64_t ln_hash = 1469598103934665603ull;
        if (memo_on) {
            for (uint8_t x : ln) ln_hash = (ln_hash ^ x) * 1099511628211ull;
            lap(0);
            for (const Memo& m : memo)
                if (m.hash == ln_hash && m.ne == ne && m.nn == nn && m.ln == ln && from_memo(m, p0, ne, t)) {
                    lap(5);
                    return true;
                }
        }
        const uint64_t conflicts0 = conflicts, accesses0 = accesses;
        lap(0);
        // ---- rows: coupled nodes of every tile node, ordered by their position k in the global block row
        if ((int)rows.size() < nn) rows.resize(nn);
        for (int u = 0; u < nn; ++u) rows[u].clear();
        // (a pair (u, v) has one position k in the block row of u, whichever element reports it: keep the first report only)
        for (int el = 0; el < ne; ++el) {
            const uint64_t e = elem_at(p0 + el);
            for (int a = 0; a < n; ++a) {
                const int u = ln[el * n + a];
                for (int b = 0; b < n; ++b) {
                    const int v = ln[el * n + b];
                    uint8_t& seen = pair_seen[u * 128 + v];
                    if (!seen) {
                        seen = 1;
                        rows[u].push_back({blockmap[e * n2 + a * n + b], (uint8_t)v, 0, (uint16_t)(el * n2 + a * n + b)});
                    }
                }
            }
        }
        int nslots = 0, nflush = 0;
        for (int u = 0; u < nn; ++u) {
            auto& r = rows[u];
            std::sort(r.begin(), r.end(), [](const RowEntry& x, const RowEntry& y) { return x.k < y.k; });
            nflush += (int)r.size();
            for (const RowEntry& x : r) {
                nslots += x.v >= u;
                pair_seen[u * 128 + x.v] = 0;  // (table clean again for the next tile)
            }
        }
        if (nslots > shape.max_slots) return false;
        lap(1);
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
                pair_pos[u * 128 + x.v] = x.pos;
            }
        }
        int maxfill = 0;
        for (int q = 0; q < 16; ++q) maxfill = std::max(maxfill, fill[q]);
        t.P = (uint32_t)(16 * maxfill);
        lap(2);
        // ---- flush list + node list
        t.p0 = p0;
        t.ne = ne;
        t.nodes.resize(nn);
        for (int u = 0; u < nn; ++u) t.nodes[u] = nodes[u] | (inc[u] == degree[nodes[u]] ? (int32_t)0x80000000 : 0);
        t.flush.clear();
        t.flush.reserve(nflush);
        flush_src.clear();
        for (int u = 0; u < nn; ++u)
            for (const RowEntry& x : rows[u]) {
                flush_src.push_back(x.src);
                uint32_t pos, tr = 0;
                if (x.v >= u) {
                    pos = x.pos;
                } else {  // mirrored block: the accumulator of (v, u)
                    pos = pair_pos[x.v * 128 + u];
                    tr = 1;
                }
                if (x.k >= (1u << kTileKBits)) degenerate = true;
                t.flush.push_back(pos | (tr << 11) | ((uint32_t)u << 12) | ((uint32_t)x.k << 19));
            }
        lap(3);
        // ---- schedule: greedy node-disjoint colouring inside the tile; every colour class is cut into rounds of at most
        // `warps` elements; rounds are padded to `warps` schedule positions (padding: node byte 0 = 0xff, no accumulators)
        std::vector<uint64_t> node_mask(nn, 0);
        std::vector<int> colour(ne), sched(ne);
        for (int el = 0; el < ne; ++el) {
            uint64_t used = 0;
            for (int a = 0; a < n; ++a) used |= node_mask[ln[el * n + a]];
            int c = 0;
            while (c < 64 && ((used >> c) & 1)) ++c;
            if (c == 64) return false;  // 64 mutually conflicting elements: no free colour - the caller halves the tile
            colour[el] = c;
            for (int a = 0; a < n; ++a) node_mask[ln[el * n + a]] |= 1ull << c;
        }
        for (int el = 0; el < ne; ++el) sched[el] = el;
        std::stable_sort(sched.begin(), sched.end(), [&](int x, int y) { return colour[x] < colour[y]; });
        const int gw = shape.warps;
        t.lnodes.clear();
        t.emap.clear();
        t.elem.clear();
        t.lnodes.reserve((size_t)(ne + 8 * gw) * n);  // (one allocation each instead of a dozen reallocations per tile)
        t.emap.reserve((size_t)(ne + 8 * gw) * n2);
        t.elem.reserve((size_t)(ne + 8 * gw));
        t.rounds = 0;
        auto pad_round = [&]() {
            while (t.elem.size() % gw) {
                t.elem.push_back(-1);
                t.lnodes.insert(t.lnodes.end(), n, (uint8_t)0);
                t.lnodes[t.lnodes.size() - n] = 0xff;
                t.emap.insert(t.emap.end(), n2, (uint16_t)0xffffu);
            }
        };
        // bit 15 of a map entry: this is the FIRST contribution to the accumulator in the tile's schedule (rounds run in order and are
        // node-disjoint) - the kernel stores instead of adding, so the accumulators never have to be cleared between tiles
        std::vector<uint8_t> touched(t.P, 0);
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
                    if (u <= v) {
                        const uint16_t pos = pair_pos[u * 128 + v];
                        em[a * n + b] = (uint16_t)(pos | (touched[pos] ? 0u : (unsigned)kTileFirstTouch));
                        touched[pos] = 1;
                    } else {
                        em[a * n + b] = (uint16_t)0xffffu;
                    }
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
        lap(4);
        t.rounds = (uint32_t)(t.elem.size() / gw);
        if (t.rounds > 255) return false;  // bound asserted by the self test (cannot happen with <= 64 elements per tile)
        if (memo_on && !degenerate && memo.size() < kMemoCap) {
            Memo m;
            m.hash = ln_hash;
            m.ne = ne;
            m.nn = nn;
            m.ln = ln;
            m.P = t.P;
            m.rounds = t.rounds;
            m.flush = t.flush;
            for (uint32_t& w : m.flush) w &= (1u << 19) - 1u;
            m.flush_src = flush_src;
            m.lnodes = t.lnodes;
            m.emap = t.emap;
            m.elem.resize(t.elem.size());
            {
                size_t i = 0;
                int s = 0;
                for (int32_t e : t.elem) m.elem[i++] = e < 0 ? (int8_t)-1 : (int8_t)sched[s++];
            }
            m.conflicts = conflicts - conflicts0;
            m.accesses = accesses - accesses0;
            memo.push_back(std::move(m));
        }
        return true;
    }
};

}  // namespace

// pass 2 of a tile: split the CSR-ordered flush list into the STORE and the REDUCE segment (see "ownership" in the header comment)
static void finish_flush(TileOut& t, uint32_t tile, bool owner, const uint32_t* owner_tile, const int32_t* degree, const int32_t* degree_owned,
                         const int64_t* blk_off) {
    const int nn = (int)t.nodes.size();
    // class of every tile node: 0 reduce, 1 store (complete), 2 store all entries of the rows (owner of a shared node)
    uint8_t cls[128];
    t.wait.clear();
    t.flags = 0;
    for (int u = 0; u < nn; ++u) {
        const int32_t id = t.nodes[u] & 0x7fffffff;
        const bool ghosted = degree[id] != degree_owned[id];
        if (ghosted) t.flags |= 1u;  // the tile touches a partition-interface node (the fused exchange looks its neighbour rows up)
        if (t.nodes[u] < 0) {
            cls[u] = 1;
        } else if (!owner) {
            cls[u] = 0;
        } else if (ghosted) {
            cls[u] = 0;  // ghost elements touch the node: other ranks add to its rows too; cleared by the caller, never stored
        } else if (owner_tile[id] == tile) {
            cls[u] = 2;
        } else {
            cls[u] = 0;
            t.wait.push_back(owner_tile[id]);
        }
    }
    std::sort(t.wait.begin(), t.wait.end());
    t.wait.erase(std::unique(t.wait.begin(), t.wait.end()), t.wait.end());
    // STORE segment: first the shared rows this tile owns (the tile publishes its flag right after them, so that the tiles waiting for
    // it are released as early as possible), then the complete rows
    std::vector<uint32_t> store, complete_rows, reduce;
    store.reserve(t.flush.size() + 256);
    complete_rows.reserve(t.flush.size());
    reduce.reserve(t.flush.size());
    t.zero_entries = 0;
    size_t f = 0;
    const size_t nf = t.flush.size();
    while (f < nf) {
        const uint32_t u = (t.flush[f] >> 12) & 0x7fu;
        size_t g = f;
        while (g < nf && ((t.flush[g] >> 12) & 0x7fu) == u) ++g;
        if (cls[u] == 0) {
            reduce.insert(reduce.end(), t.flush.begin() + f, t.flush.begin() + g);
        } else if (cls[u] == 1) {
            complete_rows.insert(complete_rows.end(), t.flush.begin() + f, t.flush.begin() + g);
        } else {
            const int32_t id = t.nodes[u] & 0x7fffffff;
            const uint32_t cnt = (uint32_t)(blk_off[id + 1] - blk_off[id]);
            size_t h = f;
            for (uint32_t k = 0; k < cnt; ++k) {
                if (h < g && (t.flush[h] >> 19) == k) {
                    store.push_back(t.flush[h++]);
                } else {
                    store.push_back(kTileZeroPos | (u << 12) | (k << 19));
                    ++t.zero_entries;
                }
            }
        }
        f = g;
    }
    t.n_publish = (uint32_t)store.size();
    store.insert(store.end(), complete_rows.begin(), complete_rows.end());
    t.n_store = (uint32_t)store.size();
    store.insert(store.end(), reduce.begin(), reduce.end());
    t.flush.swap(store);
}

void build_tile_lists(const TileShape& shape, uint64_t count, const int32_t* order, const uint64_t* codes, const int32_t* conn,
                      uint64_t num_elements, uint64_t num_owned, uint64_t num_nodes, const uint16_t* blockmap, const int64_t* blk_off,
                      HostTiles& out) {
    constexpr int n = 8;
    SetupTimer tm;
    // incidences of every node over ALL elements of the space - also the ghost elements of a partition, which are in the pattern
    // but are not assembled here: a node is complete only if every element it belongs to is processed inside one tile (then
    // the flush writes every entry of its rows)
    const unsigned hw = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    std::atomic<uint64_t> next{0};
    auto run_pool = [&](uint64_t work_items, const std::function<void()>& fn) {
        const unsigned nthreads = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, work_items / 64));
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(fn);
        fn();
        for (auto& th : pool) th.join();
    };
    std::vector<int32_t> degree(num_nodes, 0), degree_owned(num_nodes, 0);
    for (uint64_t e = 0; e < num_elements; ++e)  // (atomic increments from a thread pool: slower than this loop on the 8-core build host)
        for (int a = 0; a < n; ++a) {
            ++degree[conn[e * n + a]];
            if (e < num_owned) ++degree_owned[conn[e * n + a]];
        }
    out.hdr.clear();
    out.nodes.clear();
    out.flush.clear();
    out.wait.clear();
    out.zero_nodes.clear();
    out.lnodes.clear();
    out.emap.clear();
    out.elem.clear();
    out.bank_conflict_share = 0;
    out.zero_entries = 0;
    out.owner_stores = shape.owner_stores != 0 && blk_off != nullptr;
    // candidate tiles: runs with a common Morton prefix
    std::vector<std::pair<uint64_t, int>> cand;
    for (uint64_t p0 = 0; p0 < count;) {
        int ne = 1;
        while (ne < shape.max_elems && p0 + ne < count && (codes[p0 + ne] >> shape.tile_bits) == (codes[p0] >> shape.tile_bits)) ++ne;
        cand.emplace_back(p0, ne);
        p0 += ne;
    }
    std::vector<std::vector<TileOut>> results(cand.size());
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
        if (std::getenv("FB200_DEBUG_SETUP"))
            std::fprintf(stderr, "[fb200 setup]     worker: nodes %.3f rows %.3f positions %.3f flush %.3f schedule %.3f relabel %.3f s\n", b.sec[0], b.sec[1], b.sec[2], b.sec[3], b.sec[4], b.sec[5]),
                std::fprintf(stderr, "[fb200 setup]     worker: %llu tiles relabelled from %zu known local connectivities\n", (unsigned long long)b.memo_hits, b.memo.size());
    };
    tm.lap("  tile lists: degrees + candidates");
    run_pool(cand.size(), worker);
    tm.lap("  tile lists: pass 1 (per tile)");
    if (degenerate.load()) {  // elements with repeated nodes (or huge block rows): the caller keeps the per-element kernel
        out.hdr.clear();
        out.bank_conflict_share = -1.0;
        return;
    }
    // ---- pass 2: tile numbers are final now (the order of `results`); owner of every node = the lowest tile that touches it
    std::vector<TileOut*> tiles;
    for (auto& rs : results)
        for (TileOut& t : rs) tiles.push_back(&t);
    std::vector<uint32_t> owner_tile;
    if (out.owner_stores) {
        owner_tile.assign(num_nodes, 0xffffffffu);
        next.store(0);
        auto claim = [&]() {  // lowest tile number per node: an atomic minimum gives what the in-order loop gives
            for (;;) {
                const uint64_t t0 = next.fetch_add(64);
                if (t0 >= tiles.size()) break;
                for (uint64_t ti = t0; ti < std::min<uint64_t>(tiles.size(), t0 + 64); ++ti)
                    for (int32_t nd : tiles[ti]->nodes) {
                        uint32_t* o = &owner_tile[nd & 0x7fffffff];
                        uint32_t cur = __atomic_load_n(o, __ATOMIC_RELAXED);
                        while ((uint32_t)ti < cur && !__atomic_compare_exchange_n(o, &cur, (uint32_t)ti, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                        }
                    }
            }
        };
        run_pool(tiles.size(), claim);
        for (uint64_t i = 0; i < num_nodes; ++i)
            if (owner_tile[i] == 0xffffffffu || degree[i] != degree_owned[i]) out.zero_nodes.push_back((int32_t)i);
    }
    next.store(0);
    auto finisher = [&]() {
        for (;;) {
            const uint64_t t0 = next.fetch_add(64);
            if (t0 >= tiles.size()) break;
            for (uint64_t ti = t0; ti < std::min<uint64_t>(tiles.size(), t0 + 64); ++ti)
                finish_flush(*tiles[ti], (uint32_t)ti, out.owner_stores, owner_tile.data(), degree.data(), degree_owned.data(), blk_off);
        }
    };
    tm.lap("  tile lists: owners");
    run_pool(tiles.size(), finisher);
    tm.lap("  tile lists: pass 2 (segments)");
    // ---- tile colours for the deterministic coloured scatter (CsrParAssembler's idea, global.rs:322-373, at tile granularity): greedy,
    // in tile order; two tiles of a colour share no node, so a launch over one colour adds at most once to every CSR value and the
    // launches, in colour order, fix the order of all additions.  More than 64 colours: no tile colouring (per-element colours are used)
    {
        std::vector<uint64_t> node_mask(num_nodes, 0);
        std::vector<uint8_t> colour(tiles.size(), 0);
        std::vector<uint64_t> per_colour(64, 0);
        bool ok = true;
        for (size_t ti = 0; ti < tiles.size() && ok; ++ti) {
            uint64_t used = 0;
            for (int32_t nd : tiles[ti]->nodes) used |= node_mask[nd & 0x7fffffff];
            int c = 0;
            while (c < 64 && ((used >> c) & 1)) ++c;
            if (c == 64) {
                ok = false;
                break;
            }
            colour[ti] = (uint8_t)c;
            ++per_colour[c];
            for (int32_t nd : tiles[ti]->nodes) node_mask[nd & 0x7fffffff] |= 1ull << c;
        }
        out.colour_off.clear();
        out.colour_tiles.clear();
        if (ok && !tiles.empty()) {
            int ncol = 0;
            for (int c = 0; c < 64; ++c)
                if (per_colour[c]) ncol = c + 1;
            out.colour_off.assign(ncol + 1, 0);
            for (int c = 0; c < ncol; ++c) out.colour_off[c + 1] = out.colour_off[c] + per_colour[c];
            out.colour_tiles.resize(tiles.size());
            std::vector<uint64_t> cursor(out.colour_off.begin(), out.colour_off.end() - 1);
            for (size_t ti = 0; ti < tiles.size(); ++ti) out.colour_tiles[cursor[colour[ti]]++] = (uint32_t)ti;
        }
    }
    tm.lap("  tile lists: tile colours");
    // ---- concatenate
    uint64_t tot_nodes = 0, tot_flush = 0, tot_wait = 0, tot_pos = 0;
    for (const TileOut* t : tiles) {
        tot_nodes += t->nodes.size();
        tot_flush += t->flush.size();
        tot_wait += t->wait.size();
        tot_pos += t->elem.size();
    }
    // offsets of every tile in the output arrays (prefix sums), then the copies in parallel
    const size_t T = tiles.size();
    std::vector<uint64_t> o_nodes(T + 1, 0), o_flush(T + 1, 0), o_wait(T + 1, 0), o_pos(T + 1, 0);
    for (size_t ti = 0; ti < T; ++ti) {
        o_nodes[ti + 1] = o_nodes[ti] + tiles[ti]->nodes.size();
        o_flush[ti + 1] = o_flush[ti] + tiles[ti]->flush.size();
        o_wait[ti + 1] = o_wait[ti] + tiles[ti]->wait.size();
        o_pos[ti + 1] = o_pos[ti] + tiles[ti]->elem.size();
        out.zero_entries += tiles[ti]->zero_entries;
    }
    (void)tot_nodes, (void)tot_flush, (void)tot_wait, (void)tot_pos;
    out.hdr.resize(T * kTileHdrWords);
    out.nodes.resize(o_nodes[T]);
    out.flush.resize(o_flush[T]);
    out.wait.resize(o_wait[T]);
    out.lnodes.resize(o_pos[T] * n);
    out.emap.resize(o_pos[T] * n * n);
    out.elem.resize(o_pos[T]);
    next.store(0);
    auto copier = [&]() {
        for (;;) {
            const uint64_t t0 = next.fetch_add(64);
            if (t0 >= T) break;
            for (uint64_t ti = t0; ti < std::min<uint64_t>(T, t0 + 64); ++ti) {
                const TileOut& t = *tiles[ti];
                const uint32_t hdr[kTileHdrWords] = {(uint32_t)o_pos[ti],   t.rounds, (uint32_t)t.nodes.size(), t.P, (uint32_t)o_nodes[ti],
                                                     (uint32_t)o_flush[ti], (uint32_t)t.flush.size(), (uint32_t)t.ne, t.n_store,
                                                     (uint32_t)o_wait[ti],  (uint32_t)t.wait.size(), t.flags, t.n_publish, 0u, 0u, 0u};
                std::copy(hdr, hdr + kTileHdrWords, out.hdr.begin() + ti * kTileHdrWords);
                std::copy(t.nodes.begin(), t.nodes.end(), out.nodes.begin() + o_nodes[ti]);
                std::copy(t.flush.begin(), t.flush.end(), out.flush.begin() + o_flush[ti]);
                std::copy(t.wait.begin(), t.wait.end(), out.wait.begin() + o_wait[ti]);
                std::copy(t.lnodes.begin(), t.lnodes.end(), out.lnodes.begin() + o_pos[ti] * n);
                std::copy(t.emap.begin(), t.emap.end(), out.emap.begin() + o_pos[ti] * n * n);
                std::copy(t.elem.begin(), t.elem.end(), out.elem.begin() + o_pos[ti]);
            }
        }
    };
    run_pool(T, copier);
    out.bank_conflict_share = accesses.load() ? (double)conflicts.load() / (double)accesses.load() : 0.0;
    tm.lap("  tile lists: concatenate");
    if (tm.on) {  // FB200_DEBUG_SETUP: a checksum of everything that goes to the device (list builds must be reproducible)
        uint64_t hsh = 1469598103934665603ull;
        auto mix = [&](const void* ptr, size_t bytes) {
            const unsigned char* c = static_cast<const unsigned char*>(ptr);
            for (size_t i = 0; i < bytes; ++i) hsh = (hsh ^ c[i]) * 1099511628211ull;
        };
        mix(out.hdr.data(), out.hdr.size() * 4);
        mix(out.nodes.data(), out.nodes.size() * 4);
        mix(out.flush.data(), out.flush.size() * 4);
        mix(out.wait.data(), out.wait.size() * 4);
        mix(out.lnodes.data(), out.lnodes.size());
        mix(out.emap.data(), out.emap.size() * 2);
        mix(out.elem.data(), out.elem.size() * 4);
        mix(out.colour_tiles.data(), out.colour_tiles.size() * 4);
        mix(out.zero_nodes.data(), out.zero_nodes.size() * 4);
        std::fprintf(stderr, "[fb200 setup]   tile lists: checksum %016llx\n", (unsigned long long)hsh);
    }
}

}  // namespace fb200

// ---------------------------------------------------------------------------------------------------------------- host self check
extern "C" fb200_status fb200_tile_lists_selftest(uint64_t num_nodes, const double* vertices, uint64_t num_elements, const uint64_t* connectivity,
                                                  uint64_t num_owned, uint64_t stats[8], int32_t* failed_check) {
    uint64_t st[10];
    const fb200_status s = fb200_tile_lists_selftest_ex(num_nodes, vertices, num_elements, connectivity, num_owned, 1, st, failed_check);
    if (s == FB200_OK && stats) std::copy(st, st + 8, stats);
    return s;
}

extern "C" fb200_status fb200_tile_lists_selftest_ex(uint64_t num_nodes, const double* vertices, uint64_t num_elements,
                                                     const uint64_t* connectivity, uint64_t num_owned, int32_t owner_stores, uint64_t stats[10],
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
    shape.owner_stores = owner_stores ? 1 : 0;
    std::vector<int64_t> blk_off(num_nodes + 1, 0);
    for (uint64_t i = 0; i < num_nodes; ++i) blk_off[i + 1] = blk_off[i] + (int64_t)rows[i].size();
    HostTiles ht;
    build_tile_lists(shape, order.size(), order.data(), codes.data(), conn.data(), num_elements, num_owned, num_nodes, map.data(), blk_off.data(), ht);
    if (ht.bank_conflict_share < 0.0) return FB200_ERR_UNSUPPORTED;
    if (ht.owner_stores != (owner_stores != 0)) return fail_check(22);
    // ---- checks
    std::vector<int32_t> degree(num_nodes, 0), degree_owned(num_nodes, 0);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) {
            ++degree[conn[e * n + a]];
            if (e < num_owned) ++degree_owned[conn[e * n + a]];
        }
    std::vector<uint8_t> seen(num_elements, 0);
    const size_t ntiles = ht.hdr.size() / kTileHdrWords;
    uint64_t max_nodes = 0, max_P = 0, complete = 0, max_rounds = 0, scheduled = 0, store_entries = 0, zero_entries = 0;
    // ownership: the lowest tile that touches a node; how many tiles store the node's rows (must be exactly one per storable node)
    std::vector<uint32_t> first_tile(num_nodes, 0xffffffffu);
    std::vector<uint8_t> stored_by(num_nodes, 0);
    for (size_t t = 0; t < ntiles; ++t) {
        const uint32_t* h = &ht.hdr[t * kTileHdrWords];
        if ((uint64_t)h[4] + h[2] > ht.nodes.size()) return fail_check(2);
        for (uint32_t u = 0; u < h[2]; ++u) {
            uint32_t& o = first_tile[ht.nodes[h[4] + u] & 0x7fffffff];
            if (o == 0xffffffffu) o = (uint32_t)t;
        }
    }
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
        std::vector<uint8_t> first_seen(P, 0);
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
                        const uint16_t raw = em[a * n + b];
                        if (ln[a] > ln[b]) {
                            if (raw != 0xffffu) return fail_check(8);
                            continue;
                        }
                        const uint16_t ps = raw & (uint16_t)(kTileFirstTouch - 1);
                        if (ps >= P) return fail_check(8);
                        // the first-touch bit is set exactly on the first schedule position that uses the accumulator
                        if (((raw & kTileFirstTouch) != 0) != (first_seen[ps] == 0)) return fail_check(37);
                        first_seen[ps] = 1;
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
        // flush list: every coupled ordered pair (u, v) exactly once with the position of v in u's block row; two segments (STORE, then
        // REDUCE), each in CSR order of its rows; zero words only in the STORE segment of a shared node's owner, completing its rows
        const uint32_t nstore = h[8], wb = h[9], nw = h[10], npub = h[12];
        if (nstore > nf || npub > nstore || (uint64_t)wb + nw > ht.wait.size()) return fail_check(2);
        store_entries += nstore;
        const size_t diag = [&] { size_t d = 0; for (auto& pr : pairs) d += (pr.first >> 8) == (pr.first & 0xffu); return d; }();
        uint32_t last_u = 0, last_k = 0, zeros = 0;
        std::vector<uint8_t> row_seg(nn, 0);       // 1 = row seen in the STORE segment, 2 = in the REDUCE segment
        std::vector<uint32_t> row_entries(nn, 0);
        for (uint32_t f = 0; f < nf; ++f) {
            const uint32_t w = ht.flush[fb + f];
            const uint32_t ps = w & 0x7ffu, tr = (w >> 11) & 1u, u = (w >> 12) & 0x7fu, k = w >> 19;
            const uint32_t seg = f < nstore ? 1u : 2u;
            if (u >= nn) return fail_check(13);
            if (f && f != nstore && f != npub && (u < last_u || (u == last_u && k <= last_k))) return fail_check(14);
            last_u = u;
            last_k = k;
            if (row_seg[u] && row_seg[u] != seg) return fail_check(23);  // a row lives in one segment only
            row_seg[u] = (uint8_t)seg;
            // the leading part of the STORE segment (before the publish) holds exactly the shared rows this tile owns
            if (seg == 1 && (f < npub) != (ht.nodes[nb + u] >= 0)) return fail_check(33);
            ++row_entries[u];
            const auto& r = rows[ht.nodes[nb + u] & 0x7fffffff];
            if (k >= r.size()) return fail_check(15);
            const int32_t vg = r[k];
            const auto it = std::lower_bound(ht.nodes.begin() + nb, ht.nodes.begin() + nb + nn, vg, [](int32_t x, int32_t y) { return (x & 0x7fffffff) < y; });
            const bool v_in_tile = it != ht.nodes.begin() + nb + nn && (*it & 0x7fffffff) == vg;
            const uint32_t v = (uint32_t)(it - (ht.nodes.begin() + nb));
            const uint32_t key = v_in_tile ? ((std::min(u, v)) << 8) | std::max(u, v) : 0xffffffffu;
            const auto pit = std::lower_bound(pairs.begin(), pairs.end(), std::make_pair(key, 0u));
            const bool coupled = v_in_tile && pit != pairs.end() && pit->first == key;
            if (ps == kTileZeroPos) {
                // a zero word: only where the tile has no contribution, only in the STORE segment, only with ownership
                if (!ht.owner_stores || seg != 1 || tr || coupled) return fail_check(24);
                ++zeros;
                continue;
            }
            if (!v_in_tile) return fail_check(16);
            if (tr != (u > v ? 1u : 0u)) return fail_check(17);
            if (!coupled || pit->second != ps) return fail_check(18);
        }
        if (nf - zeros != 2 * pairs.size() - diag) return fail_check(12);
        zero_entries += zeros;
        for (uint32_t u = 0; u < nn; ++u) {
            const int32_t id = ht.nodes[nb + u] & 0x7fffffff;
            const bool is_complete = ht.nodes[nb + u] < 0, ghosted = degree[id] != degree_owned[id];
            const bool owner_here = ht.owner_stores && !ghosted && first_tile[id] == (uint32_t)t;
            const bool must_store = is_complete || owner_here;
            if (row_seg[u] != (must_store ? 1 : 2)) return fail_check(25);
            if (must_store) {
                if (row_entries[u] != rows[id].size()) return fail_check(26);  // a stored row is written completely
                if (++stored_by[id] != 1) return fail_check(27);
            } else if (ht.owner_stores && !ghosted) {
                // reductions into a row another tile stores: that tile is lower-numbered and in the wait list
                const uint32_t o = first_tile[id];
                if (o >= (uint32_t)t || !std::binary_search(ht.wait.begin() + wb, ht.wait.begin() + wb + nw, o)) return fail_check(28);
            }
        }
        for (uint32_t k = 0; k < nw; ++k)
            if (ht.wait[wb + k] >= (uint32_t)t || (k && ht.wait[wb + k] <= ht.wait[wb + k - 1])) return fail_check(29);
        {  // header flag bit 0 <=> the tile touches a node that ghost elements touch too
            bool any = false;
            for (uint32_t u = 0; u < nn; ++u) {
                const int32_t id = ht.nodes[nb + u] & 0x7fffffff;
                any = any || degree[id] != degree_owned[id];
            }
            if (any != ((h[11] & 1u) != 0u)) return fail_check(36);
        }
        if (!ht.owner_stores && (nw || zeros)) return fail_check(30);
        for (uint32_t u = 0; u < nn; ++u) {
            const bool flag = ht.nodes[nb + u] < 0;
            if (flag != (inc[u] == degree[ht.nodes[nb + u] & 0x7fffffff])) return fail_check(19);
            complete += flag;
        }
    }
    for (uint64_t e = 0; e < num_owned; ++e)
        if (!seen[e]) return fail_check(20);
    // with ownership every row is either stored by exactly one tile or listed for the caller to clear
    if (ht.owner_stores) {
        size_t z = 0;
        for (uint64_t i = 0; i < num_nodes; ++i) {
            const bool listed = z < ht.zero_nodes.size() && (uint64_t)ht.zero_nodes[z] == i;
            z += listed;
            if (listed == (stored_by[i] == 1)) return fail_check(31);
            if (listed != (first_tile[i] == 0xffffffffu || degree[i] != degree_owned[i])) return fail_check(31);
        }
        if (z != ht.zero_nodes.size()) return fail_check(31);
        if (zero_entries != ht.zero_entries) return fail_check(32);
    } else if (!ht.zero_nodes.empty()) {
        return fail_check(31);
    }
    // tile colours: every tile exactly once, tiles of a colour node-disjoint
    if (!ht.colour_off.empty()) {
        if (ht.colour_tiles.size() != ntiles || ht.colour_off.back() != ntiles) return fail_check(34);
        std::vector<uint8_t> tseen(ntiles, 0);
        std::vector<uint32_t> stamp(num_nodes, 0xffffffffu);
        for (size_t c = 0; c + 1 < ht.colour_off.size(); ++c)
            for (uint64_t k = ht.colour_off[c]; k < ht.colour_off[c + 1]; ++k) {
                const uint32_t t = ht.colour_tiles[k];
                if (t >= ntiles || tseen[t]++) return fail_check(34);
                const uint32_t* h = &ht.hdr[t * kTileHdrWords];
                for (uint32_t u = 0; u < h[2]; ++u) {
                    uint32_t& st = stamp[ht.nodes[h[4] + u] & 0x7fffffff];
                    if (st == (uint32_t)c) return fail_check(35);
                    st = (uint32_t)c;
                }
            }
    }
    stats[0] = ntiles;
    stats[1] = max_nodes;
    stats[2] = max_P;
    stats[3] = ht.flush.size();
    stats[4] = complete;
    stats[5] = (uint64_t)(ht.bank_conflict_share * 1e6);
    stats[6] = scheduled;
    stats[7] = max_rounds;
    stats[8] = store_entries;
    stats[9] = zero_entries;
    // ---- the lists do not depend on the memo of local connectivities: built tile by tile they are byte-identical
    {
        TileShape plain = shape;
        plain.memo = 0;
        HostTiles h2;
        build_tile_lists(plain, order.size(), order.data(), codes.data(), conn.data(), num_elements, num_owned, num_nodes, map.data(), blk_off.data(), h2);
        auto same = [](const auto& x, const auto& y) { return x.size() == y.size() && std::equal(x.begin(), x.end(), y.begin()); };
        if (!same(ht.hdr, h2.hdr) || !same(ht.nodes, h2.nodes) || !same(ht.flush, h2.flush) || !same(ht.wait, h2.wait) || !same(ht.zero_nodes, h2.zero_nodes) ||
            !same(ht.lnodes, h2.lnodes) || !same(ht.emap, h2.emap) || !same(ht.elem, h2.elem) || !same(ht.colour_off, h2.colour_off) ||
            !same(ht.colour_tiles, h2.colour_tiles) || ht.bank_conflict_share != h2.bank_conflict_share || ht.zero_entries != h2.zero_entries)
            return fail_check(38);
    }
    return FB200_OK;
}
