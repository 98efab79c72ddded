// Multi-GPU exchange step: the mesh is partitioned by elements (one rank = one GPU = one slab of elements);
// rows of nodes on a partition interface receive contributions on several ranks and are summed with
// ncclAllReduce over NVLink/NVSwitch on a packed buffer that holds ONLY those interface rows.
// (The reference has no distributed path at all - README.md:58; this is the north-star extension.)
//
// NCCL is resolved at run time with dlopen so that the library shares whatever libnccl.so.2 the host process
// (e.g. torch.distributed) already loaded, and so that single-GPU users need no NCCL at all.
#include <dlfcn.h>

#include <cstring>

#include "fb200_internal.h"

namespace fb200 {

struct Id128 {  // ncclUniqueId (passed by value)
    char bytes[FB200_UNIQUE_ID_BYTES];
};
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static bool load_nccl(std::string* why) {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // prefer the copy the process already has
        if (h) break;
    }
    if (!h)
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
    if (!h) {
        if (why) *why = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        if (why) *why = "libnccl.so.2 lacks required symbols";
        return false;
    }
    g_nccl.handle = h;
    return true;
}

static fb200_status nccl_fail(fb200_ctx* ctx, int code, const char* what) {
    std::string m = std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "?");
    return fail(ctx, FB200_ERR_NCCL, m);
}

// values of node row blocks <-> packed interface buffer
template <bool PACK>
__global__ void iface_copy_kernel(const int32_t* __restrict__ nodes, const int64_t* __restrict__ offsets, uint64_t count,
                                  const int64_t* __restrict__ blk_off, int ss, double* values, double* packed) {
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t k = gwarp; k < count; k += nwarps) {
        const int32_t node = nodes[k];
        const int64_t b = blk_off[node];
        const int64_t len = (blk_off[node + 1] - b) * ss;
        double* v = values + b * ss;
        double* pk = packed + offsets[k];
        for (int64_t t = lane; t < len; t += 32) {
            if (PACK) pk[t] = v[t]; else v[t] = pk[t];
        }
    }
}

// values of node row blocks += received blocks (a node shared with several peers gets several additions: reductions)
__global__ void iface_add_kernel(const int32_t* __restrict__ nodes, const int64_t* __restrict__ offsets, uint64_t count,
                                 const int64_t* __restrict__ blk_off, int ss, double* values, const double* __restrict__ recv) {
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t k = gwarp; k < count; k += nwarps) {
        const int32_t node = nodes[k];
        const int64_t b = blk_off[node];
        const int64_t len = (blk_off[node + 1] - b) * ss;
        double* v = values + b * ss;
        const double* pk = recv + offsets[k];
        for (int64_t t = lane; t < len; t += 32) atomicAdd(v + t, pk[t]);
    }
}

// ---- fused exchange over peer memory: neighbour barrier.  Thread t signals peer t (writes the sequence number into MY slot of the peer's
// signal words, system scope, after a system fence) and waits until the peer's number has arrived in ITS slot of mine.
struct PeerSignals {
    unsigned long long* remote[2];
    int rank[2];
    int n;
};
__global__ void peer_barrier_kernel(unsigned long long* mine, PeerSignals ps, int my_rank, unsigned long long seq, unsigned long long* errword) {
    const int t = threadIdx.x;
    if (t >= ps.n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ps.remote[t] + my_rank), "l"(seq) : "memory");
    const unsigned long long* slot = mine + ps.rank[t];
    unsigned long long v;
    unsigned int spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
        if (v >= seq) break;
        __nanosleep(100);
        if (++spins > (1u << 26)) {  // several seconds: a neighbour is gone - report instead of hanging the stream
            atomicMin(errword, ((unsigned long long)ps.rank[t] << 8) | (unsigned long long)FB200_ERR_NCCL);
            break;
        }
    }
    __threadfence_system();
}

}  // namespace fb200

using namespace fb200;

void fb200::p2p_disable(fb200_ctx* ctx) {
    if (!ctx) return;
    auto& pp = ctx->p2p;
    for (int k = 0; k < 2; ++k) {
        if (pp.values[k]) cudaIpcCloseMemHandle(pp.values[k]);
        if (pp.signal[k]) cudaIpcCloseMemHandle(pp.signal[k]);
        pp.values[k] = nullptr;
        pp.signal[k] = nullptr;
        pp.peer_rank[k] = -1;
    }
    dev_free(pp.d_peer_row);
    pp.enabled = false;
    pp.pending = false;
    pp.num_peers = 0;
    // (d_signal stays allocated: a neighbour may still hold a mapping of it; freed with the context)
}

fb200_status fb200::p2p_neighbour_barrier(fb200_ctx* ctx) {
    auto& pp = ctx->p2p;
    if (!pp.enabled || pp.num_peers == 0) return FB200_OK;
    PeerSignals ps;
    ps.n = pp.num_peers;
    for (int k = 0; k < 2; ++k) {
        ps.remote[k] = pp.signal[k];
        ps.rank[k] = pp.peer_rank[k];
    }
    ++pp.seq;
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(pp.d_signal, ps, ctx->rank, pp.seq, ctx->d_errword);
    return check_launch(ctx, "peer_barrier_kernel");
}

extern "C" {

void fb200_comm_destroy_internal(fb200_ctx* ctx) {
    if (ctx) {
        p2p_disable(ctx);
        dev_free(ctx->p2p.d_signal);
    }
    if (ctx && ctx->nccl_comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
}

fb200_status fb200_comm_unique_id(char id[FB200_UNIQUE_ID_BYTES]) {
    std::string why;
    if (!load_nccl(&why)) return FB200_ERR_NCCL;
    Id128 tmp;
    std::memset(&tmp, 0, sizeof(tmp));
    const int rc = g_nccl.GetUniqueId(&tmp);
    if (rc != 0) return FB200_ERR_NCCL;
    std::memcpy(id, tmp.bytes, FB200_UNIQUE_ID_BYTES);
    return FB200_OK;
}

fb200_status fb200_comm_init(fb200_ctx* ctx, const char id[FB200_UNIQUE_ID_BYTES], int32_t rank, int32_t num_ranks) {
    if (!ctx) return FB200_ERR_STATE;
    std::string why;
    if (!load_nccl(&why)) return fail(ctx, FB200_ERR_NCCL, why);
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    fb200_comm_destroy_internal(ctx);
    Id128 tmp;
    std::memcpy(tmp.bytes, id, FB200_UNIQUE_ID_BYTES);
    void* comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, num_ranks, tmp, rank);
    if (rc != 0) return nccl_fail(ctx, rc, "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = num_ranks;
    return FB200_OK;
}

fb200_status fb200_interface_set(fb200_ctx* ctx, uint64_t count, const uint64_t* local_nodes, const uint64_t* packed_offsets,
                                 uint64_t packed_len) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "interface_set needs a pattern");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    dev_free(ctx->d_iface_nodes);
    dev_free(ctx->d_iface_offsets);
    dev_free(ctx->d_iface_packed);
    ctx->iface_count = 0;
    ctx->iface_packed_len = 0;
    // validate on the host against the block structure
    std::vector<int64_t> off(ctx->N + 1);
    FB200_CUDA(ctx, d2h_copy(ctx, off.data(), ctx->d_blk_off, (ctx->N + 1) * sizeof(int64_t)));
    std::vector<int32_t> nodes(count);
    std::vector<int64_t> offs(count);
    const uint64_t ss = (uint64_t)ctx->sdim * ctx->sdim;
    for (uint64_t k = 0; k < count; ++k) {
        if (local_nodes[k] >= ctx->N) return fail(ctx, FB200_ERR_INDEX_OOB, "interface node out of range");
        const uint64_t len = (uint64_t)(off[local_nodes[k] + 1] - off[local_nodes[k]]) * ss;
        if (packed_offsets[k] + len > packed_len) return fail(ctx, FB200_ERR_SHAPE, "interface row block exceeds the packed buffer");
        nodes[k] = (int32_t)local_nodes[k];
        offs[k] = (int64_t)packed_offsets[k];
    }
    FB200_TRY(dev_alloc(ctx, &ctx->d_iface_nodes, count));
    FB200_TRY(dev_alloc(ctx, &ctx->d_iface_offsets, count));
    FB200_TRY(dev_alloc(ctx, &ctx->d_iface_packed, packed_len));
    if (count) {
        FB200_CUDA(ctx, h2d_copy(ctx, ctx->d_iface_nodes, nodes.data(), count * sizeof(int32_t)));
        FB200_CUDA(ctx, h2d_copy(ctx, ctx->d_iface_offsets, offs.data(), count * sizeof(int64_t)));
    }
    ctx->iface_count = count;
    ctx->iface_packed_len = packed_len;
    return FB200_OK;
}

fb200_status fb200_interface_set_peers(fb200_ctx* ctx, uint64_t num_peers, const int32_t* peer_ranks, const uint64_t* peer_begin,
                                       const uint64_t* nodes) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "interface_set_peers needs a pattern");
    if (num_peers && (!peer_ranks || !peer_begin || !nodes)) return fail(ctx, FB200_ERR_SHAPE, "null peer arrays");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    dev_free(ctx->d_peer_nodes);
    dev_free(ctx->d_peer_offsets);
    dev_free(ctx->d_peer_send);
    dev_free(ctx->d_peer_recv);
    ctx->peer_ranks.clear();
    ctx->peer_seg_off.clear();
    ctx->peer_count = 0;
    ctx->h_peer_nodes.clear();
    ctx->h_peer_begin.clear();
    p2p_disable(ctx);
    if (num_peers == 0) return FB200_OK;
    std::vector<int64_t> off(ctx->N + 1);
    FB200_CUDA(ctx, d2h_copy(ctx, off.data(), ctx->d_blk_off, (ctx->N + 1) * sizeof(int64_t)));
    const uint64_t ss = (uint64_t)ctx->sdim * ctx->sdim, total_nodes = peer_begin[num_peers];
    std::vector<int32_t> h_nodes(total_nodes);
    std::vector<int64_t> h_offs(total_nodes);
    std::vector<uint64_t> seg(num_peers + 1, 0);
    uint64_t pos = 0;
    for (uint64_t pr = 0; pr < num_peers; ++pr) {
        if (peer_begin[pr] > peer_begin[pr + 1]) return fail(ctx, FB200_ERR_SHAPE, "peer_begin must be non-decreasing");
        if (peer_ranks[pr] < 0 || peer_ranks[pr] == ctx->rank || (ctx->nranks > 1 && peer_ranks[pr] >= ctx->nranks))
            return fail(ctx, FB200_ERR_SHAPE, "bad peer rank");
        seg[pr] = pos;
        for (uint64_t k = peer_begin[pr]; k < peer_begin[pr + 1]; ++k) {
            if (nodes[k] >= ctx->N) return fail(ctx, FB200_ERR_INDEX_OOB, "interface node out of range");
            h_nodes[k] = (int32_t)nodes[k];
            h_offs[k] = (int64_t)pos;
            pos += (uint64_t)(off[nodes[k] + 1] - off[nodes[k]]) * ss;
        }
    }
    seg[num_peers] = pos;
    FB200_TRY(dev_alloc(ctx, &ctx->d_peer_nodes, total_nodes));
    FB200_TRY(dev_alloc(ctx, &ctx->d_peer_offsets, total_nodes));
    FB200_TRY(dev_alloc(ctx, &ctx->d_peer_send, pos));
    FB200_TRY(dev_alloc(ctx, &ctx->d_peer_recv, pos));
    if (total_nodes) {
        FB200_CUDA(ctx, h2d_copy(ctx, ctx->d_peer_nodes, h_nodes.data(), total_nodes * sizeof(int32_t)));
        FB200_CUDA(ctx, h2d_copy(ctx, ctx->d_peer_offsets, h_offs.data(), total_nodes * sizeof(int64_t)));
    }
    ctx->peer_ranks.assign(peer_ranks, peer_ranks + num_peers);
    ctx->peer_seg_off = seg;
    ctx->peer_count = total_nodes;
    ctx->h_peer_nodes = h_nodes;
    ctx->h_peer_begin.assign(peer_begin, peer_begin + num_peers + 1);
    return FB200_OK;
}

// neighbour exchange: every rank sends the packed rows it shares with a peer and adds what the peer sends back
static fb200_status interface_exchange_peers(fb200_ctx* ctx) {
    if (!ctx->nccl_comm) return fail(ctx, FB200_ERR_STATE, "fb200_comm_init has not been called");
    if (!g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) return fail(ctx, FB200_ERR_NCCL, "libnccl lacks ncclSend/ncclRecv");
    const int ss = ctx->sdim * ctx->sdim;
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(ctx->peer_count * 32, 256), (uint64_t)ctx->sm_count * 8));
    if (ctx->peer_count) {
        iface_copy_kernel<true><<<blocks, 256, 0, ctx->stream>>>(ctx->d_peer_nodes, ctx->d_peer_offsets, ctx->peer_count, ctx->d_blk_off, ss,
                                                                ctx->d_values, ctx->d_peer_send);
        FB200_TRY(check_launch(ctx, "iface_copy_kernel<pack peers>"));
    }
    int rc = g_nccl.GroupStart();
    if (rc != 0) return nccl_fail(ctx, rc, "ncclGroupStart");
    for (size_t pr = 0; pr < ctx->peer_ranks.size(); ++pr) {
        const uint64_t o = ctx->peer_seg_off[pr], len = ctx->peer_seg_off[pr + 1] - o;
        if (len == 0) continue;
        rc = g_nccl.Send(ctx->d_peer_send + o, len, /*ncclFloat64*/ 8, ctx->peer_ranks[pr], ctx->nccl_comm, ctx->stream);
        if (rc == 0) rc = g_nccl.Recv(ctx->d_peer_recv + o, len, /*ncclFloat64*/ 8, ctx->peer_ranks[pr], ctx->nccl_comm, ctx->stream);
        if (rc != 0) break;
    }
    const int rc_end = g_nccl.GroupEnd();
    if (rc != 0) return nccl_fail(ctx, rc, "ncclSend/ncclRecv");
    if (rc_end != 0) return nccl_fail(ctx, rc_end, "ncclGroupEnd");
    if (ctx->peer_count) {
        iface_add_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_peer_nodes, ctx->d_peer_offsets, ctx->peer_count, ctx->d_blk_off, ss, ctx->d_values,
                                                         ctx->d_peer_recv);
        FB200_TRY(check_launch(ctx, "iface_add_kernel"));
    }
    return FB200_OK;
}

fb200_status fb200_interface_enable_p2p(fb200_ctx* ctx) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "interface_enable_p2p needs a pattern");
    if (!ctx->nccl_comm) return fail(ctx, FB200_ERR_STATE, "fb200_comm_init has not been called");
    if (!g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) return fail(ctx, FB200_ERR_NCCL, "libnccl lacks ncclSend/ncclRecv");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    p2p_disable(ctx);
    auto& pp = ctx->p2p;
    const size_t np = ctx->peer_ranks.size();
    // the fused flush carries ONE neighbour per interface node and at most two neighbours per rank (slab-like partitions); anything else
    // keeps the packed ncclSend / ncclRecv exchange.  The answer must be the same on every rank of a neighbourhood, and it is: a node
    // shared by three ranks is in two peer lists on each of them.
    bool ok = np >= 1 && np <= 2 && ctx->N < (1ull << 31);
    std::vector<uint32_t> owner(ctx->N, 0xffffffffu);
    for (size_t pr = 0; pr < np && ok; ++pr)
        for (uint64_t k = ctx->h_peer_begin[pr]; k < ctx->h_peer_begin[pr + 1]; ++k) {
            uint32_t& o = owner[ctx->h_peer_nodes[k]];
            if (o != 0xffffffffu) ok = false;
            o = (uint32_t)pr;
        }
    // every rank must reach the exchanges below or none: agree on `ok` with the neighbours first (a refusal is sent as a zero-length layout)
    const uint64_t total = ctx->peer_count;
    std::vector<int64_t> off(ctx->N + 1);
    FB200_CUDA(ctx, d2h_copy(ctx, off.data(), ctx->d_blk_off, (ctx->N + 1) * sizeof(int64_t)));
    // message to peer pr: [ok, ipc handle of values (64 B), ipc handle of the signal words (64 B)] as 17 int64, then (offset, count) per node
    if (!pp.d_signal) {
        FB200_TRY(dev_alloc(ctx, &pp.d_signal, (size_t)std::max(ctx->nranks, 1)));
        FB200_CUDA(ctx, cudaMemsetAsync(pp.d_signal, 0, sizeof(unsigned long long) * (size_t)std::max(ctx->nranks, 1), ctx->stream));
        FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    constexpr size_t HDR = 17;
    std::vector<int64_t> send(np * HDR + 2 * total, 0), recv(np * HDR + 2 * total, 0);
    cudaIpcMemHandle_t hv, hs;
    std::memset(&hv, 0, sizeof(hv));
    std::memset(&hs, 0, sizeof(hs));
    if (cudaIpcGetMemHandle(&hv, ctx->d_values) != cudaSuccess || cudaIpcGetMemHandle(&hs, pp.d_signal) != cudaSuccess) {
        cudaGetLastError();
        ok = false;  // (no IPC in this process / container: told to the neighbours as a refusal, decided by the vote below)
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    std::vector<size_t> seg(np + 1, 0);
    for (size_t pr = 0; pr < np; ++pr) {
        const size_t nn = (size_t)(ctx->h_peer_begin[pr + 1] - ctx->h_peer_begin[pr]);
        int64_t* m = send.data() + seg[pr];
        m[0] = ok ? 1 : 0;
        std::memcpy(m + 1, &hv, 64);
        std::memcpy(m + 9, &hs, 64);
        for (size_t k = 0; k < nn; ++k) {
            const int32_t node = ctx->h_peer_nodes[ctx->h_peer_begin[pr] + k];
            m[HDR + 2 * k] = off[node];
            m[HDR + 2 * k + 1] = off[node + 1] - off[node];
        }
        seg[pr + 1] = seg[pr] + HDR + 2 * nn;
    }
    int64_t *d_send = nullptr, *d_recv = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_send, send.size()));
    fb200_status st = dev_alloc(ctx, &d_recv, recv.size());
    if (st == FB200_OK && h2d_copy(ctx, d_send, send.data(), send.size() * sizeof(int64_t)) != cudaSuccess) st = fail(ctx, FB200_ERR_CUDA, "H2D p2p layout");
    if (st == FB200_OK) {
        int rc = g_nccl.GroupStart();
        for (size_t pr = 0; pr < np && rc == 0; ++pr) {
            rc = g_nccl.Send(d_send + seg[pr], seg[pr + 1] - seg[pr], /*ncclInt64*/ 4, ctx->peer_ranks[pr], ctx->nccl_comm, ctx->stream);
            if (rc == 0) rc = g_nccl.Recv(d_recv + seg[pr], seg[pr + 1] - seg[pr], /*ncclInt64*/ 4, ctx->peer_ranks[pr], ctx->nccl_comm, ctx->stream);
        }
        const int rc_end = g_nccl.GroupEnd();
        if (rc != 0 || rc_end != 0) st = nccl_fail(ctx, rc ? rc : rc_end, "p2p layout exchange");
    }
    if (st == FB200_OK && cudaMemcpyAsync(recv.data(), d_recv, recv.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        st = fail(ctx, FB200_ERR_CUDA, "D2H p2p layout");
    if (st == FB200_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FB200_ERR_CUDA, "p2p layout exchange");
    dev_free(d_send);
    dev_free(d_recv);
    FB200_TRY(st);
    // From here on nobody returns before the world has agreed: a rank that enabled the fused exchange next to one that did not would
    // wait for reductions that never come.  local verdict -> ncclAllReduce(min) -> everybody enables or nobody does.
    std::string why;
    for (size_t pr = 0; pr < np; ++pr) ok = ok && recv[seg[pr]] == 1;
    if (np == 0) ok = true;  // a rank without interface has nothing to fuse and no reason to veto
    else if (!ok) why = "fused p2p exchange needs one neighbour per interface node and at most two neighbours per rank";
    std::vector<uint32_t> row(ctx->N, 0u);
    for (size_t pr = 0; pr < np && ok; ++pr) {
        const size_t nn = (size_t)(ctx->h_peer_begin[pr + 1] - ctx->h_peer_begin[pr]);
        const int64_t* m = recv.data() + seg[pr];
        cudaIpcMemHandle_t rv, rs;
        std::memcpy(&rv, m + 1, 64);
        std::memcpy(&rs, m + 9, 64);
        for (size_t k = 0; k < nn && ok; ++k) {
            const int32_t node = ctx->h_peer_nodes[ctx->h_peer_begin[pr] + k];
            if (m[HDR + 2 * k + 1] != off[node + 1] - off[node] || m[HDR + 2 * k] < 0 || m[HDR + 2 * k] >= (int64_t)0x7fffffff) {
                ok = false;
                why = "interface rows have different layouts on the two ranks (ghost elements missing?)";
            }
            row[node] = ((uint32_t)m[HDR + 2 * k] + 1u) | ((uint32_t)pr << 31);
        }
        if (!ok) break;
        void *pv = nullptr, *ps = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&pv, rv, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&ps, rs, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            if (pv) cudaIpcCloseMemHandle(pv);
            cudaGetLastError();
            ok = false;
            why = std::string("cudaIpcOpenMemHandle failed (no peer access between the ranks?): ") + cudaGetErrorString(e);
            break;
        }
        pp.values[pr] = static_cast<double*>(pv);
        pp.signal[pr] = static_cast<unsigned long long*>(ps);
        pp.peer_rank[pr] = ctx->peer_ranks[pr];
    }
    {
        int32_t* d_vote = nullptr;
        int32_t vote = ok ? 1 : 0;
        FB200_TRY(dev_alloc(ctx, &d_vote, 1));
        cudaError_t e = h2d_copy(ctx, d_vote, &vote, sizeof(vote));
        int rc = 0;
        if (e == cudaSuccess) rc = g_nccl.AllReduce(d_vote, d_vote, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, ctx->nccl_comm, ctx->stream);
        if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(&vote, d_vote, sizeof(vote), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(ctx->stream);
        dev_free(d_vote);
        if (rc != 0) {
            p2p_disable(ctx);
            return nccl_fail(ctx, rc, "ncclAllReduce(p2p vote)");
        }
        if (e != cudaSuccess) {
            p2p_disable(ctx);
            return cuda_fail(ctx, e, "p2p vote");
        }
        if (vote != 1) {
            p2p_disable(ctx);
            return fail(ctx, FB200_ERR_UNSUPPORTED, why.empty() ? "another rank cannot use the fused p2p exchange" : why);
        }
    }
    FB200_TRY(dev_alloc(ctx, &pp.d_peer_row, (size_t)ctx->N));
    if (ctx->N) FB200_CUDA(ctx, h2d_copy(ctx, pp.d_peer_row, row.data(), ctx->N * sizeof(uint32_t)));
    pp.num_peers = (int)np;
    pp.enabled = true;
    pp.pending = false;
    // nobody may start reducing into a neighbour before that neighbour has mapped everything: one barrier to finish the set-up
    FB200_TRY(p2p_neighbour_barrier(ctx));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return read_errword(ctx);
}

fb200_status fb200_interface_allreduce(fb200_ctx* ctx) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "interface_allreduce needs a pattern");
    if (ctx->nranks > 1 && !ctx->nccl_comm) return fail(ctx, FB200_ERR_STATE, "fb200_comm_init has not been called");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->p2p.pending) {
        // the last assembly already reduced this rank's interface sums into the neighbours' rows (hex8_tile_kernel.cuh, PEER); the rows
        // are complete once every neighbour's kernel has finished: one neighbour barrier, no data moves here
        ctx->p2p.pending = false;
        return p2p_neighbour_barrier(ctx);
    }
    if (!ctx->peer_ranks.empty()) return interface_exchange_peers(ctx);
    if (ctx->iface_packed_len == 0) return FB200_OK;
    const int ss = ctx->sdim * ctx->sdim;
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(ctx->iface_count * 32, 256), (uint64_t)ctx->sm_count * 8));
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_iface_packed, 0, ctx->iface_packed_len * sizeof(double), ctx->stream));
    if (ctx->iface_count) {
        iface_copy_kernel<true><<<blocks, 256, 0, ctx->stream>>>(ctx->d_iface_nodes, ctx->d_iface_offsets, ctx->iface_count, ctx->d_blk_off, ss,
                                                                ctx->d_values, ctx->d_iface_packed);
        FB200_TRY(check_launch(ctx, "iface_copy_kernel<pack>"));
    }
    if (ctx->nccl_comm) {
        const int rc = g_nccl.AllReduce(ctx->d_iface_packed, ctx->d_iface_packed, ctx->iface_packed_len, /*ncclFloat64*/ 8, /*ncclSum*/ 0,
                                        ctx->nccl_comm, ctx->stream);
        if (rc != 0) return nccl_fail(ctx, rc, "ncclAllReduce");
    }
    if (ctx->iface_count) {
        iface_copy_kernel<false><<<blocks, 256, 0, ctx->stream>>>(ctx->d_iface_nodes, ctx->d_iface_offsets, ctx->iface_count, ctx->d_blk_off, ss,
                                                                 ctx->d_values, ctx->d_iface_packed);
        FB200_TRY(check_launch(ctx, "iface_copy_kernel<unpack>"));
    }
    return FB200_OK;
}

}  // extern "C"
