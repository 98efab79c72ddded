// Internal definitions shared by the translation units of libfenris_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/fenris_b200.h"

namespace fb200 {

constexpr int kMaxNodes = 27;
constexpr int kMaxDim = 3;
constexpr unsigned long long kNoError = ~0ull;

struct ElementInfo {
    int n;   // nodes per element
    int ng;  // geometry nodes (sub-parametric high-order elements use the embedded linear element)
    int d;   // geometry == reference dimension
};
bool element_info(int element_type, ElementInfo* out);

// Host restatement of the reference-element gradient tables (see hostgen.cpp).
void reference_gradients(int element_type, const double* xi, double* g /* [n*d], node-major */);
void reference_basis(int element_type, const double* xi, double* phi /* [n] */);
int geometry_type(int element_type);

// Device-side error word: [63:8] smallest offending element index, [7:0] status code. atomicMin keeps the first element.
struct DeviceTables {
    // layout (doubles): w[nq] | mu[nq] | lam[nq] | ggeo[nq*ng*d] | gref[nq*n*d]
    double* d_data = nullptr;
    std::vector<double> host;  // what is currently resident in d_data
    size_t capacity = 0;  // doubles
    int nq = 0;
    bool uniform_params = true;
    double mu0 = 0, lam0 = 0;
};

// connectivity + scatter map gathered into a processing order (see assemble.cu::ensure_ordered)
struct OrderedCopy {
    int32_t* conn = nullptr;
    uint16_t* map = nullptr;
    const int32_t* ids = nullptr;  // the device order array it was built from (not owned)
    uint64_t count = 0;
    bool valid = false;
};

// chunk-local scatter lists (host build: chunks.cpp; consumer: tet4_chunk_kernel.cuh)
struct HostChunks {
    std::vector<int64_t> slot_off;     // num_chunks + 1
    std::vector<uint16_t> contrib;     // count * n^2, chunk c starts at c * chunk_elems * n^2
    std::vector<int32_t> slot_node;
    std::vector<uint16_t> slot_k, slot_cbeg;
    std::vector<uint8_t> slot_flags;   // bit 0: row node complete in the chunk (plain stores), bit 1: partition-interface row
    std::vector<int64_t> slot_dst;     // index of the block's first value: s^2 blk_off[node] + s k
    std::vector<int32_t> slot_rl;      // row length in doubles
};
void build_chunk_lists(int n, int sdim, uint64_t count, int chunk_elems, const int32_t* order, const int32_t* conn, uint64_t num_elements,
                       uint64_t num_nodes, const int64_t* blk_off, const uint16_t* blockmap, HostChunks& out);

// std::vector whose resize() leaves trivially constructible elements uninitialised: the big list arrays are written completely by a pool
// of threads right after being sized, so value-initialising (and page-faulting) hundreds of MB on one thread first only costs time
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = DefaultInitAllocator<U>;
    };
    using std::allocator<T>::allocator;
    template <class U>
    void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
};
template <class T>
using BigVec = std::vector<T, DefaultInitAllocator<T>>;

// tiles of the tile-accumulating Hex8 kernel (host build: tiles.cpp::build_tile_lists; consumer: hex8_tile_kernel.cuh)
constexpr int kTileKBits = 13;     // bits of k (position of the column node in a block row) in a flush word
constexpr uint32_t kTileFirstTouch = 0x8000u;  // bit 15 of an element-map entry: first contribution to the accumulator position in the tile
constexpr uint32_t kTileZeroPos = 0x7ffu;  // accumulator position of a flush word that writes 0.0 (owner rows, see below)
// header of a tile: 0 first schedule position, 1 rounds, 2 n_nodes, 3 n_slots (P), 4 node_begin, 5 flush_begin, 6 n_flush, 7 n_elems,
// 8 n_store (leading flush entries that are plain stores when the call overwrites), 9 wait_begin, 10 n_wait, 11 flags,
// 12 n_publish (leading STORE entries = the shared rows this tile owns; its flag is published after them), 13-15 reserved
constexpr int kTileHdrWords = 16;
struct TileShape {
    int tile_bits;    // low Morton bits dropped to name a tile (5: 4 x 4 x 2 elements, 6: 4 x 4 x 4)
    int max_elems;    // elements per tile
    int warps;        // elements per round = warps per compute group
    int max_nodes;    // distinct nodes per tile (<= 128)
    int max_slots;    // accumulator positions per tile (upper-triangle node blocks + padding)
    int owner_stores = 1;  // 1: rows shared by several tiles are STORED (all of their entries, zeros included) by the lowest-numbered tile
                           // that touches the node and reduced into by the others once that tile has published its stores - an
                           // overwriting assembly then needs no zero-fill of the values.  0: only tile-complete rows are stored, every
                           // other row is reduced into (the caller zero-fills all values first).
    int memo = 1;  // 1: tiles with a local connectivity seen before are re-labelled copies (tiles.cpp Builder::Memo); 0: every tile is built
};
struct HostTiles {
    BigVec<uint32_t> hdr;        // num_tiles * kTileHdrWords
    BigVec<int32_t> nodes;       // global node id | 0x80000000 when all incident elements of the node lie in the tile
    BigVec<uint32_t> flush;      // per tile: first the STORE segment, then the REDUCE segment, each per (row node u, coupled node v) in CSR
                                      // order: position | transposed << 11 | u << 12 | k << 19; position kTileZeroPos = the entry is 0.0
    BigVec<uint32_t> wait;       // per tile: the (lower-numbered) tiles whose stores its reductions must wait for
    std::vector<uint64_t> colour_off;   // tile colours (empty: none): colour c = colour_tiles[colour_off[c] .. colour_off[c + 1])
    std::vector<uint32_t> colour_tiles; // tiles grouped by colour, ascending inside a colour; tiles of one colour share no node
    std::vector<int32_t> zero_nodes;  // owner_stores: nodes whose rows no tile stores (touched by ghost elements, or by no owned element at
                                      // all): the only rows an overwriting call has to clear beforehand
    BigVec<uint8_t> lnodes;      // positions * 8: tile-local node index of each element node (byte 0 = 0xff: padding position)
    BigVec<uint16_t> emap;       // positions * 64: accumulator position of block (a, b) (| kTileFirstTouch), 0xffff when u_a > u_b (mirrored at the flush)
    BigVec<int32_t> elem;        // positions: element id of each schedule position (-1: padding)
    bool owner_stores = false;        // the lists were built with TileShape::owner_stores
    uint64_t zero_entries = 0;        // flush words with position kTileZeroPos
    double bank_conflict_share = 0;   // diagnostic: share of accumulate accesses that collide in a shared-memory bank
};
// blk_off (node-block row offsets, N + 1) is needed for owner_stores only (NULL: built as with owner_stores = 0)
void build_tile_lists(const TileShape& shape, uint64_t count, const int32_t* order, const uint64_t* codes, const int32_t* conn,
                      uint64_t num_elements, uint64_t num_owned, uint64_t num_nodes, const uint16_t* blockmap, const int64_t* blk_off,
                      HostTiles& out);

struct TileLists {
    bool valid = false;
    bool unusable = false;  // the mesh has no usable tile lists (degenerate elements): keep the per-element kernel
    uint64_t count = 0;
    const int32_t* ids = nullptr;
    int tile_bits = 0;
    int owner_stores = 0;   // what was asked for ...
    bool owner = false;     // ... and whether the lists carry it (HostTiles::owner_stores)
    uint32_t num_tiles = 0;
    uint32_t* d_hdr = nullptr;
    int32_t* d_nodes = nullptr;
    uint32_t* d_flush = nullptr;
    uint32_t* d_wait = nullptr;
    uint32_t* d_colour_tiles = nullptr;  // tiles grouped by colour (HostTiles::colour_tiles); offsets on the host
    std::vector<uint64_t> colour_off;
    uint32_t* d_flag = nullptr;        // per tile: epoch of the launch whose stores of this tile are published
    int32_t* d_zero_nodes = nullptr;   // rows to clear before an overwriting launch (owner lists)
    uint64_t zero_node_count = 0;
    uint8_t* d_lnodes = nullptr;
    uint16_t* d_emap = nullptr;
    int32_t* d_elem = nullptr;
};

struct ChunkLists {
    bool valid = false;
    uint64_t count = 0;
    const int32_t* ids = nullptr;  // the device order array it was built from (not owned)
    int chunk_elems = 0;
    uint32_t num_chunks = 0;
    uint64_t total_slots = 0;
    int64_t* d_slot_off = nullptr;
    uint16_t* d_contrib = nullptr;
    int32_t* d_slot_node = nullptr;
    uint16_t* d_slot_k = nullptr;
    uint16_t* d_slot_cbeg = nullptr;
    uint8_t* d_slot_flags = nullptr;
    int64_t* d_slot_dst = nullptr;
    int32_t* d_slot_rl = nullptr;
    int sdim = 0;
    int32_t* d_conn_pos = nullptr;  // connectivity rows in processing order
};

}  // namespace fb200

struct fb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    std::string err;
    int64_t err_elem = -1;
    uint64_t launches = 0;

    // ---- space
    bool has_space = false;       // vertices + uniform connectivity
    bool has_connectivity = false;  // any connectivity (uniform or ragged)
    bool ragged = false;
    int elem_type = 0;
    fb200::ElementInfo ei{0, 0, 0};
    uint64_t N = 0, E = 0, E_owned = 0;
    double* d_vertices = nullptr;
    int32_t* d_conn = nullptr;       // uniform: E*n ; ragged: flat node list
    int64_t* d_elem_off = nullptr;   // ragged only: E+1
    uint64_t conn_len = 0;           // total incidences

    // ---- locality (Morton) visiting order of the owned elements for the atomic scatter
    int32_t* d_order = nullptr;
    uint64_t order_count = 0;
    std::vector<int32_t> h_order;  // over all E elements; filtered to the owned ones on upload
    std::vector<uint64_t> h_order_codes_all, h_order_codes;  // Morton codes of h_order / of the owned elements in d_order

    fb200::OrderedCopy ord_morton, ord_colors;
    fb200::ChunkLists chunks;
    fb200::TileLists tiles;
    int tune_hex8_tile = -1;  // fb200_set_tuning("hex8_tile"); -1 = FB200_HEX8_TILE from the environment, else 64
    int tune_colored_tiles = -1;  // fb200_set_tuning("hex8_colored_tiles"): COLORED scatter of Hex8 by tile colours (deterministic); -1 = on
    int tune_owner = -1;      // fb200_set_tuning("hex8_owner_stores"): first-writer stores instead of zero-fill + reductions (TileShape::owner_stores);
                              // -1 = FB200_HEX8_OWNER from the environment, else on
    uint32_t tile_epoch = 0;  // launch counter of the owner-store flags (TileLists::d_flag)
    bool pending_zero = false;  // an overwriting assembly was requested and no kernel has cleared / overwritten the values yet
    // fused zero-fill lists of the Hex8 atomic kernel (see assemble.cu::ensure_zero_lists)
    int64_t* d_zero_off = nullptr;
    int32_t* d_zero_nodes = nullptr;
    int64_t* d_zero_base = nullptr;
    int32_t* d_zero_len = nullptr;
    uint32_t* d_row_epoch = nullptr;
    uint32_t epoch = 0;
    bool zero_valid = false;
    uint64_t zero_count = 0;

    // ---- adjacency: node -> flat incidence indices k into d_conn (uniform: element = k / n, local node = k % n),
    //      sorted ascending per node (deterministic)
    int64_t* d_adj_off = nullptr;    // N+1
    int32_t* d_adj_inc = nullptr;    // conn_len

    // ---- pattern (node-block form; scalar CSR arrays are derived on download)
    bool has_pattern = false;
    bool adopted = false;
    int sdim = 0;
    uint64_t P = 0;                  // number of coupled node pairs (block nnz)
    uint64_t nrows = 0, nnz = 0;
    int64_t* d_blk_off = nullptr;    // N+1
    int32_t* d_blk_cols = nullptr;   // P, sorted per node
    uint16_t* d_blockmap = nullptr;  // uniform: E*n*n, position of node b in the block row of node a
    int max_row_blocks = 0;

    // ---- colours
    bool has_colors = false;
    std::vector<uint64_t> h_color_off;  // num_colors+1
    std::vector<uint64_t> h_color_elems;
    int32_t* d_color_elems = nullptr;   // E (owned elements only are launched)

    // ---- values
    double* d_values = nullptr;
    uint64_t values_capacity = 0;

    // ---- mass / source vector path: tables w | rho | ggeo | phi_geo | phi, the global vector, source values
    double* d_ms_tab = nullptr;
    std::vector<double> h_ms_tab;
    size_t ms_tab_capacity = 0;
    double* d_vector = nullptr;
    uint64_t vector_len = 0, vector_capacity = 0;
    double* d_source = nullptr;
    uint64_t source_capacity = 0;

    // ---- pinned staging for large D2H copies of the host preprocessing (d2h_staged)
    void* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};

    // ---- tables + deferred error word
    fb200::DeviceTables tab;
    unsigned long long* d_ticket = nullptr;   // dynamic work counter of the element kernels
    unsigned long long* d_errword = nullptr;
    unsigned long long* h_errword = nullptr;  // pinned

    // ---- multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    uint64_t iface_count = 0, iface_packed_len = 0;
    int32_t* d_iface_nodes = nullptr;
    int64_t* d_iface_offsets = nullptr;
    double* d_iface_packed = nullptr;
    // neighbour exchange (fb200_interface_set_peers): per-peer segments of the send / receive buffers
    std::vector<int32_t> peer_ranks;
    std::vector<uint64_t> peer_seg_off;   // num_peers + 1, in doubles
    uint64_t peer_count = 0;              // interface incidences (nodes over all segments)
    int32_t* d_peer_nodes = nullptr;
    int64_t* d_peer_offsets = nullptr;
    double* d_peer_send = nullptr;
    double* d_peer_recv = nullptr;
    std::vector<int32_t> h_peer_nodes;    // the nodes of all segments (host copy, for fb200_interface_enable_p2p)
    std::vector<uint64_t> h_peer_begin;   // num_peers + 1, in nodes

    // fused exchange over peer memory (fb200_interface_enable_p2p, comm.cu): the tile kernel's flush reduces interface rows straight
    // into the neighbouring ranks' values through pointers mapped with CUDA IPC; neighbours synchronise through signal words
    struct P2P {
        bool enabled = false;
        bool pending = false;             // a fused assembly was launched and its closing neighbour barrier has not run yet
        int num_peers = 0;
        int32_t peer_rank[2] = {-1, -1};
        double* values[2] = {nullptr, nullptr};                  // peers' d_values (IPC mappings)
        unsigned long long* signal[2] = {nullptr, nullptr};      // peers' signal words (IPC mappings), one slot per rank
        unsigned long long* d_signal = nullptr;                  // my signal words [nranks]
        uint32_t* d_peer_row = nullptr;                          // [N]: 0, or (block-row offset on the peer + 1) | peer slot << 31
        unsigned long long seq = 0;                              // barrier sequence number
    } p2p;
};

namespace fb200 {

fb200_status fail(fb200_ctx* ctx, fb200_status s, const std::string& msg, int64_t elem = -1);
fb200_status cuda_fail(fb200_ctx* ctx, cudaError_t e, const char* what);

#define FB200_CUDA(ctx, call)                                                   \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return fb200::cuda_fail((ctx), _e, #call);       \
    } while (0)

#define FB200_TRY(call)                         \
    do {                                        \
        fb200_status _s = (call);               \
        if (_s != FB200_OK) return _s;          \
    } while (0)

template <class T>
fb200_status dev_alloc(fb200_ctx* ctx, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
    return FB200_OK;
}
template <class T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// H2D copy of a (pageable) host array that kernels on ctx->stream read next.  ctx->stream is cudaStreamNonBlocking, so a blocking
// cudaMemcpy on the legacy stream is NOT ordered before those kernels (it may return while the DMA from the staging buffer is still
// in flight): copy on the context's own stream and wait for it.
inline cudaError_t h2d_copy(fb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
}

// D2H copy of an array that kernels on ctx->stream produced: ordered on that stream (a blocking cudaMemcpy on the legacy stream is not
// ordered after work on a cudaStreamNonBlocking stream) and complete on return
inline cudaError_t d2h_copy(fb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
}

// D2H copy of a large device array into pageable host memory through two pinned staging buffers (a pageable cudaMemcpy runs at a few
// GB/s; this one at the speed of the host memcpy): chunk k + 1 crosses PCIe while chunk k is copied out of the staging buffer
cudaError_t d2h_staged(fb200_ctx* ctx, void* dst, const void* src, size_t bytes);

// after a kernel launch
fb200_status check_launch(fb200_ctx* ctx, const char* name);

// pattern.cu
fb200_status build_adjacency(fb200_ctx* ctx);
void free_pattern(fb200_ctx* ctx);
void free_space(fb200_ctx* ctx);
fb200_status exclusive_scan_i64(fb200_ctx* ctx, int64_t* d_data, uint64_t count);  // in place, count elements

// context.cu: locality-preserving (Morton) processing order of the elements and the codes it was sorted by
void morton_order(int d, int n, uint64_t N, const double* v, uint64_t E, const uint64_t* conn, std::vector<int32_t>& order,
                  std::vector<uint64_t>& codes);

// comm.cu: neighbour barrier of the fused exchange (no-op unless p2p is enabled); closes a pending fused assembly
fb200_status p2p_neighbour_barrier(fb200_ctx* ctx);
void p2p_disable(fb200_ctx* ctx);

// assemble.cu
void free_ordered(fb200_ctx* ctx);
fb200_status upload_tables(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q);
fb200_status read_errword(fb200_ctx* ctx);  // sync + translate deferred device errors
fb200_status group_elements_by_rule(fb200_ctx* ctx, uint32_t num_rules, const uint32_t* element_rule, bool colored,
                                    std::vector<uint64_t>& col_off, std::vector<int32_t>& flat, std::vector<uint64_t>& off);

// mass_source.cu: CSR assembly of a state-dependent operator (FB200_STVK) at the host vector u (NULL = zeros)
fb200_status assemble_state_dependent(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u, int scatter_mode,
                                      int accumulate);
fb200_status element_matrices_state_dependent(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u, uint64_t first,
                                              uint64_t count, double* d_out);
fb200_status assemble_state_dependent_list(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* q, const double* u,
                                           const int32_t* d_list, uint64_t count, int plain);

inline int div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }

// FB200_DEBUG_SETUP=1: wall-clock of the host-side preprocessing steps on stderr (where does setup time go?)
struct SetupTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool on = std::getenv("FB200_DEBUG_SETUP") != nullptr;
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[fb200 setup] %-40s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// offsets[0 .. count] of a caller's CSR-like structure: starts at 0, never decreases (checked on the host before any kernel indexes with it)
inline bool offsets_well_formed(const uint64_t* offsets, uint64_t count) {
    if (offsets[0] != 0) return false;
    for (uint64_t i = 0; i < count; ++i)
        if (offsets[i + 1] < offsets[i]) return false;
    return true;
}

}  // namespace fb200
