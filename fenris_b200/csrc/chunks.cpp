// Host preprocessing for the chunk-local scatter (see tet4_chunk_kernel.cuh).
//
// The element loop of the reference adds every local row of every K_e into the CSR separately (global.rs:155-178, 504-537).  On
// the device each of those additions is one fp64 reduction in L2, and the L2 reduction rate - not HBM - bounds the scatter
// (profiles/r01/README.md 6).  Consecutive elements of the processing (Morton) order share most of their nodes, so the lists built
// here group, per chunk of `chunk_elems` consecutive positions, all contributions (element, a, b) that land in the same node block
// (I_a, I_b) of the CSR: the kernel sums them on chip and issues ONE reduction per block entry ("slot").  Blocks of nodes whose
// incident elements all lie inside the chunk are complete after that sum and are flagged so that the kernel can use plain stores.
//
// Per chunk c (positions [c C, min((c+1) C, count))):
//   contrib[c C n^2 + t], t < ne n^2 : contributor tags  e_local n^2 + a n + b, grouped by slot, elements ascending inside a slot
//   slots slot_off[c] .. slot_off[c+1]: node (row node I), k (position of the column node in I's block row), cbeg (first contributor,
//   relative to the chunk), flags (bit 0: row node complete in this chunk, bit 1: partition-interface row); slots are ordered by contributor count, largest first,
//   so that the threads of a warp (one slot each) run loops of equal length.
#include <algorithm>
#include <atomic>
#include <thread>

#include "fb200_internal.h"

namespace fb200 {

void build_chunk_lists(int n, int sdim, uint64_t count, int chunk_elems, const int32_t* order, const int32_t* conn, uint64_t num_elements,
                       uint64_t num_nodes, const int64_t* blk_off, const uint16_t* blockmap, HostChunks& out) {
    const int n2 = n * n;
    const uint64_t num_chunks = (count + chunk_elems - 1) / chunk_elems;
    // incidences of every node over ALL elements of the space - also the ghost elements of a partition, which are in the pattern but are
    // not assembled here: a row is complete in a chunk (plain stores) only if every element of its node is processed inside that chunk.
    // Rows that ghost elements touch are never complete: another rank adds to them too (packed exchange or peer-memory reductions).
    std::vector<int32_t> degree(num_nodes, 0), degree_owned(num_nodes, 0);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) ++degree[conn[e * n + a]];
    for (uint64_t pos = 0; pos < count; ++pos) {
        const uint64_t e = order ? (uint64_t)order[pos] : pos;
        for (int a = 0; a < n; ++a) ++degree_owned[conn[e * n + a]];
    }
    struct Slot {
        int32_t node;
        uint16_t k, cbeg;
        uint8_t flags;
    };
    std::vector<std::vector<Slot>> chunk_slots(num_chunks);
    out.contrib.assign(count * (uint64_t)n2, 0);
    std::atomic<uint64_t> next{0};
    auto worker = [&]() {
        std::vector<uint64_t> pairs;
        std::vector<std::pair<int32_t, int32_t>> local;  // (node, incidences inside the chunk)
        std::vector<int32_t> nodes;
        struct Raw {
            uint64_t key;
            uint32_t begin, cnt;
            int32_t node;
        };
        std::vector<Raw> raw;
        std::vector<uint16_t> tags;
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= num_chunks) break;
            const uint64_t p0 = c * (uint64_t)chunk_elems;
            const int ne = (int)std::min<uint64_t>(chunk_elems, count - p0);
            pairs.clear();
            nodes.clear();
            for (int el = 0; el < ne; ++el) {
                const uint64_t e = order ? (uint64_t)order[p0 + el] : p0 + el;
                for (int a = 0; a < n; ++a) {
                    const int32_t I = conn[e * n + a];
                    nodes.push_back(I);
                    for (int b = 0; b < n; ++b) {
                        const uint64_t key = (uint64_t)blk_off[I] + blockmap[e * n2 + a * n + b];
                        pairs.push_back((key << 16) | (uint64_t)(el * n2 + a * n + b));
                    }
                }
            }
            std::sort(pairs.begin(), pairs.end());
            // nodes whose incident elements all lie inside this chunk
            std::sort(nodes.begin(), nodes.end());
            local.clear();
            for (size_t i = 0; i < nodes.size();) {
                size_t j = i;
                while (j < nodes.size() && nodes[j] == nodes[i]) ++j;
                local.emplace_back(nodes[i], (int32_t)(j - i));
                i = j;
            }
            auto complete = [&](int32_t I) {
                auto it = std::lower_bound(local.begin(), local.end(), std::make_pair(I, (int32_t)0));
                return it != local.end() && it->first == I && it->second == degree[I];
            };
            raw.clear();
            for (size_t i = 0; i < pairs.size();) {
                size_t j = i;
                const uint64_t key = pairs[i] >> 16;
                while (j < pairs.size() && (pairs[j] >> 16) == key) ++j;
                const uint32_t tag = (uint32_t)(pairs[i] & 0xffffu);
                const int el = tag / n2, a = (tag - el * n2) / n;
                const uint64_t e = order ? (uint64_t)order[p0 + el] : p0 + el;
                raw.push_back({key, (uint32_t)i, (uint32_t)(j - i), conn[e * n + a]});
                i = j;
            }
            std::stable_sort(raw.begin(), raw.end(), [](const Raw& x, const Raw& y) { return x.cnt > y.cnt; });
            tags.clear();
            std::vector<Slot>& slots = chunk_slots[c];
            slots.reserve(raw.size());
            for (const Raw& r : raw) {
                slots.push_back({r.node, (uint16_t)(r.key - (uint64_t)blk_off[r.node]), (uint16_t)tags.size(),
                                 (uint8_t)((complete(r.node) ? 1 : 0) | (degree[r.node] != degree_owned[r.node] ? 2 : 0))});
                for (uint32_t t = 0; t < r.cnt; ++t) tags.push_back((uint16_t)(pairs[r.begin + t] & 0xffffu));
            }
            std::copy(tags.begin(), tags.end(), out.contrib.begin() + p0 * (uint64_t)n2);
        }
    };
    const unsigned hw = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    const unsigned nthreads = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, num_chunks));
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();

    out.slot_off.assign(num_chunks + 1, 0);
    for (uint64_t c = 0; c < num_chunks; ++c) out.slot_off[c + 1] = out.slot_off[c] + (int64_t)chunk_slots[c].size();
    const uint64_t total = (uint64_t)out.slot_off[num_chunks];
    out.slot_node.resize(total);
    out.slot_k.resize(total);
    out.slot_cbeg.resize(total);
    out.slot_flags.resize(total);
    out.slot_dst.resize(total);
    out.slot_rl.resize(total);
    for (uint64_t c = 0; c < num_chunks; ++c) {
        uint64_t o = (uint64_t)out.slot_off[c];
        for (const Slot& s : chunk_slots[c]) {
            // where the block's values start and how long the rows are: precomputed, so that the kernel's slot loop has no load that
            // depends on another one (node -> block-row offsets)
            out.slot_dst[o] = (int64_t)(sdim * sdim) * blk_off[s.node] + (int64_t)sdim * s.k;
            out.slot_rl[o] = (int32_t)((blk_off[s.node + 1] - blk_off[s.node]) * sdim);
            out.slot_node[o] = s.node;
            out.slot_k[o] = s.k;
            out.slot_cbeg[o] = s.cbeg;
            out.slot_flags[o] = s.flags;
            ++o;
        }
    }
}

}  // namespace fb200

// ---------------------------------------------------------------------------------------------------------------- host self check
// Builds the node-block rows and the block map the way fb200_assemble_pattern does, the Morton order of the owned elements, the chunk
// lists, and verifies them (no GPU): every contribution (element, a, b) of every owned element appears exactly once, in the slot of its
// node block (row node, k); a slot's contributors are in ascending element order; destination and row length match the block offsets;
// the complete flag is set exactly for rows whose node has ALL its elements - ghost elements included - inside the chunk; the interface
// flag exactly for rows that ghost elements touch.
extern "C" fb200_status fb200_chunk_lists_selftest(uint64_t num_nodes, const double* vertices, uint64_t num_elements, const uint64_t* connectivity,
                                                   uint64_t num_owned, int32_t chunk_elems, int32_t sdim, uint64_t stats[4], int32_t* failed_check) {
    using namespace fb200;
    constexpr int n = 4, n2 = 16;
    if (failed_check) *failed_check = 0;
    if (!vertices || !connectivity || !stats || num_owned > num_elements || chunk_elems < 1 || chunk_elems > 4096 || sdim < 1 || sdim > 3) return FB200_ERR_SHAPE;
    for (uint64_t i = 0; i < num_elements * n; ++i)
        if (connectivity[i] >= num_nodes) return FB200_ERR_INDEX_OOB;
    auto fail_check = [&](int id) {
        if (failed_check) *failed_check = id;
        return FB200_ERR_STATE;
    };
    std::vector<std::vector<int32_t>> rows(num_nodes);
    std::vector<int32_t> conn(num_elements * n);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) {
            conn[e * n + a] = (int32_t)connectivity[e * n + a];
            for (int b = 0; b < n; ++b) rows[connectivity[e * n + a]].push_back((int32_t)connectivity[e * n + b]);
        }
    std::vector<int64_t> blk_off(num_nodes + 1, 0);
    for (uint64_t i = 0; i < num_nodes; ++i) {
        auto& r = rows[i];
        std::sort(r.begin(), r.end());
        r.erase(std::unique(r.begin(), r.end()), r.end());
        if (r.size() >= 65536) return FB200_ERR_UNSUPPORTED;
        blk_off[i + 1] = blk_off[i] + (int64_t)r.size();
    }
    std::vector<uint16_t> map(num_elements * (uint64_t)n2);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) {
            const auto& r = rows[conn[e * n + a]];
            for (int b = 0; b < n; ++b) map[e * n2 + a * n + b] = (uint16_t)(std::lower_bound(r.begin(), r.end(), conn[e * n + b]) - r.begin());
        }
    std::vector<int32_t> order_all, order;
    std::vector<uint64_t> codes_all;
    morton_order(3, n, num_nodes, vertices, num_elements, connectivity, order_all, codes_all);
    for (size_t i = 0; i < order_all.size(); ++i)
        if ((uint64_t)order_all[i] < num_owned) order.push_back(order_all[i]);
    HostChunks hc;
    build_chunk_lists(n, sdim, order.size(), chunk_elems, order.data(), conn.data(), num_elements, num_nodes, blk_off.data(), map.data(), hc);
    const uint64_t count = order.size(), num_chunks = (count + chunk_elems - 1) / chunk_elems;
    if (hc.slot_off.size() != num_chunks + 1 || hc.contrib.size() != count * n2) return fail_check(1);
    std::vector<int32_t> degree(num_nodes, 0), degree_owned(num_nodes, 0);
    for (uint64_t e = 0; e < num_elements; ++e)
        for (int a = 0; a < n; ++a) {
            ++degree[conn[e * n + a]];
            if (e < num_owned) ++degree_owned[conn[e * n + a]];
        }
    uint64_t complete_slots = 0, iface_slots = 0;
    for (uint64_t c = 0; c < num_chunks; ++c) {
        const uint64_t p0 = c * (uint64_t)chunk_elems;
        const int ne = (int)std::min<uint64_t>(chunk_elems, count - p0);
        const int64_t so = hc.slot_off[c], U = hc.slot_off[c + 1] - so;
        if (U < 0 || (uint64_t)(so + U) > hc.slot_node.size()) return fail_check(2);
        std::vector<int32_t> inside(num_nodes > 0 ? 0 : 0);
        std::vector<std::pair<int32_t, int32_t>> local;  // (node, incidences inside the chunk)
        {
            std::vector<int32_t> nodes;
            for (int el = 0; el < ne; ++el)
                for (int a = 0; a < n; ++a) nodes.push_back(conn[(uint64_t)order[p0 + el] * n + a]);
            std::sort(nodes.begin(), nodes.end());
            for (size_t i = 0; i < nodes.size();) {
                size_t j = i;
                while (j < nodes.size() && nodes[j] == nodes[i]) ++j;
                local.emplace_back(nodes[i], (int32_t)(j - i));
                i = j;
            }
        }
        std::vector<uint8_t> seen((size_t)ne * n2, 0);
        for (int64_t u = 0; u < U; ++u) {
            const int32_t node = hc.slot_node[so + u];
            const int k = hc.slot_k[so + u], cb = hc.slot_cbeg[so + u];
            const int ce = u + 1 < U ? (int)hc.slot_cbeg[so + u + 1] : ne * n2;
            if (node < 0 || (uint64_t)node >= num_nodes || k >= (int)rows[node].size() || cb >= ce) return fail_check(3);
            if (hc.slot_dst[so + u] != (int64_t)(sdim * sdim) * blk_off[node] + (int64_t)sdim * k ||
                hc.slot_rl[so + u] != (int32_t)((blk_off[node + 1] - blk_off[node]) * sdim))
                return fail_check(4);
            int last_el = -1;
            for (int t = cb; t < ce; ++t) {
                const unsigned tag = hc.contrib[p0 * n2 + t];
                const int el = tag >> 4, a = (tag >> 2) & 3, b = tag & 3;
                if (el >= ne || el < last_el) return fail_check(5);  // contributors in ascending element order
                last_el = el;
                const uint64_t e = (uint64_t)order[p0 + el];
                if (conn[e * n + a] != node || map[e * n2 + a * n + b] != k) return fail_check(6);
                if (seen[(size_t)el * n2 + a * n + b]++) return fail_check(7);
            }
            const auto it = std::lower_bound(local.begin(), local.end(), std::make_pair(node, (int32_t)0));
            const bool complete = it != local.end() && it->first == node && it->second == degree[node];
            const bool iface = degree[node] != degree_owned[node];
            if (((hc.slot_flags[so + u] & 1) != 0) != complete) return fail_check(8);
            if (((hc.slot_flags[so + u] & 2) != 0) != iface) return fail_check(9);
            if (complete && iface) return fail_check(10);  // a row another rank adds to is never stored
            complete_slots += complete;
            iface_slots += iface;
        }
        for (uint8_t s : seen)
            if (s != 1) return fail_check(11);  // every contribution exactly once
    }
    stats[0] = num_chunks;
    stats[1] = hc.slot_node.size();
    stats[2] = complete_slots;
    stats[3] = iface_slots;
    return FB200_OK;
}
