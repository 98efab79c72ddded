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
