// CSR sparsity pattern (CsrAssembler::assemble_pattern, src/assembly/global.rs:65-120 / 206-297), the
// node-block scatter map, pattern download/adopt and colour handling.
//
// The reference builds, per node, a hash set of coupled nodes by visiting every element (n^2 inserts each,
// one Mutex per node in the parallel version).  Here the same sets are produced on the device without locks:
//   1. node -> incidence adjacency (count, scan, fill, per-node sort)            [deterministic]
//   2. per node (one warp): gather the nodes of all adjacent elements, drop duplicates, rank -> sorted set
//   3. scan of set sizes -> block offsets; scalar CSR arrays are sdim-expanded views of this block structure
//      (each node contributes sdim identical rows, global.rs:86-109) and are only materialised on download.
//   4. scatter map: for every element and local pair (a,b) the position of node b in the sorted set of node a.
#include <algorithm>

#include "fb200_internal.h"

namespace fb200 {

void greedy_coloring(uint64_t E, uint64_t N, const std::vector<int64_t>& off, const std::vector<int32_t>& nodes,
                     std::vector<uint64_t>& color_off, std::vector<uint64_t>& color_elems);

// ------------------------------------------------------------------------------------------------ scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(int64_t* data, uint64_t count, int64_t* tile_sums) {
    __shared__ int64_t s_warp[kScanThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    int64_t v[kScanItems];
    int64_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < count) ? data[base + i] : 0;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < kScanThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int off = 1; off < kScanThreads / 32; off <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        if (lane < kScanThreads / 32) s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int64_t warp_off = warp > 0 ? s_warp[warp - 1] : 0;
    int64_t run = warp_off + inc - sum;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < count) data[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == kScanThreads - 1) tile_sums[blockIdx.x] = warp_off + inc;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_offsets_kernel(int64_t* data, uint64_t count, const int64_t* tile_offs) {
    const int64_t off = tile_offs[blockIdx.x];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
        if (base + i < count) data[base + i] += off;
}

fb200_status exclusive_scan_i64(fb200_ctx* ctx, int64_t* d_data, uint64_t count) {
    if (count == 0) return FB200_OK;
    const uint64_t tiles = (count + kScanTile - 1) / kScanTile;
    int64_t* d_sums = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_sums, tiles));
    scan_tiles_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(d_data, count, d_sums);
    fb200_status st = check_launch(ctx, "scan_tiles_kernel");
    if (st == FB200_OK && tiles > 1) {
        st = exclusive_scan_i64(ctx, d_sums, tiles);
        if (st == FB200_OK) {
            scan_add_offsets_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(d_data, count, d_sums);
            st = check_launch(ctx, "scan_add_offsets_kernel");
        }
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sums);
    return st;
}

// ------------------------------------------------------------------------------------------------ adjacency
__global__ void count_incidence_kernel(const int32_t* __restrict__ conn, uint64_t len, int64_t* deg) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < len; k += stride)
        atomicAdd((unsigned long long*)&deg[conn[k]], 1ull);
}

__global__ void fill_incidence_kernel(const int32_t* __restrict__ conn, uint64_t len, unsigned long long* cursor, int32_t* adj_inc) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < len; k += stride) {
        const unsigned long long slot = atomicAdd(&cursor[conn[k]], 1ull);
        adj_inc[slot] = (int32_t)k;
    }
}

// one thread per node: insertion sort of its (short) incidence list -> deterministic order
__global__ void sort_incidence_kernel(const int64_t* __restrict__ adj_off, int32_t* adj_inc, uint64_t num_nodes) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t node = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; node < num_nodes; node += stride) {
        const int64_t b = adj_off[node], e = adj_off[node + 1];
        for (int64_t i = b + 1; i < e; ++i) {
            const int32_t key = adj_inc[i];
            int64_t j = i;
            while (j > b && adj_inc[j - 1] > key) {
                adj_inc[j] = adj_inc[j - 1];
                --j;
            }
            adj_inc[j] = key;
        }
    }
}

fb200_status build_adjacency(fb200_ctx* ctx) {
    if (ctx->d_adj_off) return FB200_OK;
    const uint64_t N = ctx->N, len = ctx->conn_len;
    FB200_TRY(dev_alloc(ctx, &ctx->d_adj_off, N + 1));
    FB200_TRY(dev_alloc(ctx, &ctx->d_adj_inc, len));
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_adj_off, 0, (N + 1) * sizeof(int64_t), ctx->stream));
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(len, 256), (uint64_t)ctx->sm_count * 16));
    if (len) {
        count_incidence_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_conn, len, ctx->d_adj_off);
        FB200_TRY(check_launch(ctx, "count_incidence_kernel"));
    }
    FB200_TRY(exclusive_scan_i64(ctx, ctx->d_adj_off, N + 1));
    if (len) {
        unsigned long long* d_cursor = nullptr;
        FB200_TRY(dev_alloc(ctx, &d_cursor, N + 1));
        cudaMemcpyAsync(d_cursor, ctx->d_adj_off, (N + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream);
        fill_incidence_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_conn, len, d_cursor, ctx->d_adj_inc);
        fb200_status st = check_launch(ctx, "fill_incidence_kernel");
        if (st == FB200_OK) {
            const int nb = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(N, 128), (uint64_t)ctx->sm_count * 32));
            sort_incidence_kernel<<<nb, 128, 0, ctx->stream>>>(ctx->d_adj_off, ctx->d_adj_inc, N);
            st = check_launch(ctx, "sort_incidence_kernel");
        }
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_cursor);
        FB200_TRY(st);
    }
    return FB200_OK;
}

// ------------------------------------------------------------------------------------------------ node sets
struct ConnView {
    const int32_t* conn;
    const int64_t* elem_off;  // nullptr => uniform with n nodes per element
    int n;
    uint64_t num_elements;
    // element containing flat incidence k, and its node range
    __device__ __forceinline__ void element_range(int32_t k, int64_t* b, int64_t* e) const {
        if (!elem_off) {
            const int64_t el = k / n;
            *b = el * n;
            *e = *b + n;
        } else {
            // largest el with elem_off[el] <= k
            uint64_t lo = 0, hi = num_elements;
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) >> 1;
                if (elem_off[mid] <= (int64_t)k) lo = mid; else hi = mid;
            }
            *b = elem_off[lo];
            *e = elem_off[lo + 1];
        }
    }
};

__global__ void candidate_count_kernel(ConnView cv, const int64_t* __restrict__ adj_off, const int32_t* __restrict__ adj_inc,
                                       uint64_t num_nodes, int* max_m) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int local_max = 0;
    for (uint64_t node = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; node < num_nodes; node += stride) {
        int64_t m = 0;
        for (int64_t a = adj_off[node]; a < adj_off[node + 1]; ++a) {
            int64_t b, e;
            cv.element_range(adj_inc[a], &b, &e);
            m += e - b;
        }
        local_max = max(local_max, (int)(m < 0x7fffffffLL ? m : 0x7fffffffLL));
    }
    atomicMax(max_m, local_max);
}

constexpr int kSetWarps = 4;        // warps per CTA
constexpr int kSetSmemCap = 1024;   // candidates per warp held in shared memory

// One warp per node. buf: capacity >= m candidates. FILL=false: write the set size; FILL=true: write the sorted set.
template <bool FILL>
__device__ void node_set_warp(const ConnView& cv, const int64_t* adj_off, const int32_t* adj_inc, uint64_t node, int32_t* buf,
                              int64_t* cnt_out, const int64_t* blk_off, int32_t* blk_cols, int lane) {
    // 1. gather candidates (element by element; each element's nodes strided over lanes)
    int m = 0;
    for (int64_t a = adj_off[node]; a < adj_off[node + 1]; ++a) {
        int64_t b, e;
        cv.element_range(adj_inc[a], &b, &e);
        for (int64_t k = b + lane; k < e; k += 32) buf[m + (int)(k - b)] = cv.conn[k];
        m += (int)(e - b);
    }
    __syncwarp();
    // 2. mark later duplicates (keep the first occurrence)
    //    (an element listing the same node twice, or two adjacent elements sharing it)
    int dups = 0;
    for (int base = 0; base < m; base += 32) {
        const int t = base + lane;
        bool dup = false;
        int v = 0;
        if (t < m) {
            v = buf[t];
            for (int u = 0; u < t; ++u)
                if (buf[u] == v) { dup = true; break; }
        }
        __syncwarp();
        // defer the sentinel write until the whole warp has finished reading this far: later chunks only compare
        // against earlier positions, and a sentinel never equals a real node id, so writing now is safe once
        // every lane of THIS chunk has done its scan
        if (dup) buf[t] = 0x7fffffff;
        dups += dup ? 1 : 0;
        __syncwarp();
    }
    if (!FILL) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dups += __shfl_xor_sync(0xffffffffu, dups, off);
        if (lane == 0) cnt_out[node] = (int64_t)(m - dups);
        return;
    }
    // 3. rank of every kept value among the kept values = its position in the sorted set
    const int64_t out0 = blk_off[node];
    for (int t = lane; t < m; t += 32) {
        const int v = buf[t];
        if (v == 0x7fffffff) continue;
        int rank = 0;
        for (int u = 0; u < m; ++u) rank += (buf[u] < v) ? 1 : 0;
        blk_cols[out0 + rank] = v;
    }
    __syncwarp();
}

template <bool FILL>
__global__ void __launch_bounds__(kSetWarps * 32) node_sets_kernel(ConnView cv, const int64_t* __restrict__ adj_off,
                                                                  const int32_t* __restrict__ adj_inc, uint64_t num_nodes,
                                                                  int32_t* scratch, int scratch_per_warp, int64_t* cnt_out,
                                                                  const int64_t* __restrict__ blk_off, int32_t* blk_cols) {
    __shared__ int32_t s_buf[kSetWarps][kSetSmemCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t gwarp = (uint64_t)blockIdx.x * kSetWarps + warp;
    const uint64_t nwarps = (uint64_t)gridDim.x * kSetWarps;
    int32_t* buf = scratch ? scratch + gwarp * (uint64_t)scratch_per_warp : s_buf[warp];
    for (uint64_t node = gwarp; node < num_nodes; node += nwarps) {
        node_set_warp<FILL>(cv, adj_off, adj_inc, node, buf, cnt_out, blk_off, blk_cols, lane);
        __syncwarp();
    }
}

__global__ void max_row_blocks_kernel(const int64_t* __restrict__ blk_off, uint64_t num_nodes, int* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int m = 0;
    for (uint64_t node = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; node < num_nodes; node += stride)
        { const int64_t c = blk_off[node + 1] - blk_off[node]; m = max(m, (int)(c < 0x7fffffffLL ? c : 0x7fffffffLL)); }
    atomicMax(out, m);
}

// ------------------------------------------------------------------------------------------------ scatter map
__global__ void blockmap_kernel(const int32_t* __restrict__ conn, int n, uint64_t num_elements, const int64_t* __restrict__ blk_off,
                                const int32_t* __restrict__ blk_cols, uint16_t* __restrict__ map, unsigned long long* errword) {
    const uint64_t total = num_elements * (uint64_t)(n * n);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t e = t / (uint64_t)(n * n);
        const int r = (int)(t - e * (uint64_t)(n * n));
        const int a = r / n, b = r - a * n;
        const int32_t I = conn[e * n + a], J = conn[e * n + b];
        int64_t lo = blk_off[I], hi = blk_off[I + 1];
        const int64_t base = lo;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (blk_cols[mid] < J) lo = mid + 1; else hi = mid;
        }
        if (lo < blk_off[I + 1] && blk_cols[lo] == J) {
            map[t] = (uint16_t)(lo - base);
        } else {
            map[t] = 0;
            atomicMin(errword, ((unsigned long long)e << 8) | (unsigned long long)FB200_ERR_COLUMN_NOT_IN_PATTERN);
        }
    }
}

static fb200_status build_blockmap(fb200_ctx* ctx) {
    dev_free(ctx->d_blockmap);
    if (ctx->ragged || !ctx->has_space) return FB200_OK;  // pattern-only connectivity
    if (ctx->max_row_blocks > 65535) return fail(ctx, FB200_ERR_UNSUPPORTED, "a node couples to more than 65535 nodes");
    const int n = ctx->ei.n;
    const uint64_t total = ctx->E * (uint64_t)(n * n);
    FB200_TRY(dev_alloc(ctx, &ctx->d_blockmap, total));
    if (total) {
        const int blocks = (int)std::min<uint64_t>(div_up(total, 256), (uint64_t)ctx->sm_count * 32);
        blockmap_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_conn, n, ctx->E, ctx->d_blk_off, ctx->d_blk_cols, ctx->d_blockmap,
                                                         ctx->d_errword);
        FB200_TRY(check_launch(ctx, "blockmap_kernel"));
    }
    return read_errword(ctx);
}

static fb200_status finish_pattern(fb200_ctx* ctx, int sdim) {
    ctx->sdim = sdim;
    ctx->nrows = (uint64_t)sdim * ctx->N;
    ctx->nnz = (uint64_t)sdim * (uint64_t)sdim * ctx->P;
    int* d_max = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_max, 1));
    cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream);
    if (ctx->N) {
        max_row_blocks_kernel<<<(int)std::min<uint64_t>(div_up(ctx->N, 256), 1024), 256, 0, ctx->stream>>>(ctx->d_blk_off, ctx->N, d_max);
        ctx->launches++;
    }
    int h_max = 0;
    cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_max);
    ctx->max_row_blocks = h_max;
    FB200_TRY(build_blockmap(ctx));
    dev_free(ctx->d_values);
    FB200_TRY(dev_alloc(ctx, &ctx->d_values, ctx->nnz));
    ctx->values_capacity = ctx->nnz;
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_values, 0, ctx->nnz * sizeof(double), ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_pattern = true;
    return FB200_OK;
}

// ------------------------------------------------------------------------------------------------ download helpers
__global__ void row_offsets_kernel(const int64_t* __restrict__ blk_off, uint64_t num_nodes, int s, uint64_t* __restrict__ out) {
    const uint64_t rows = num_nodes * (uint64_t)s;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += stride) {
        if (r == rows) {
            out[r] = (uint64_t)(s * s) * (uint64_t)blk_off[num_nodes];
        } else {
            const uint64_t node = r / s;
            const int i = (int)(r - node * s);
            const uint64_t cnt = (uint64_t)(blk_off[node + 1] - blk_off[node]);
            out[r] = (uint64_t)(s * s) * (uint64_t)blk_off[node] + (uint64_t)i * s * cnt;
        }
    }
}

// scalar column indices of the rows of nodes [node0, node1) into out (relative to the first entry of node0)
__global__ void col_indices_kernel(const int64_t* __restrict__ blk_off, const int32_t* __restrict__ blk_cols, uint64_t node0,
                                   uint64_t node1, int s, uint64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t base = (uint64_t)(s * s) * (uint64_t)blk_off[node0];
    for (uint64_t node = node0 + gwarp; node < node1; node += nwarps) {
        const int64_t b = blk_off[node];
        const int64_t cnt = blk_off[node + 1] - b;
        const int64_t per_row = cnt * s;
        const uint64_t o = (uint64_t)(s * s) * (uint64_t)b - base;
        for (int64_t t = lane; t < per_row * s; t += 32) {
            const int64_t within = t % per_row;
            const int64_t k = within / s;
            const int j = (int)(within - k * s);
            out[o + t] = (uint64_t)s * (uint64_t)blk_cols[b + k] + (uint64_t)j;
        }
    }
}

// ------------------------------------------------------------------------------------------------ adopt helpers
// Validate that the caller's scalar CSR is node-block structured and extract per-node block counts.
__global__ void adopt_counts_kernel(const uint64_t* __restrict__ row_off, uint64_t num_nodes, int s, int64_t* cnt, int* bad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t node = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; node < num_nodes; node += stride) {
        const uint64_t len0 = row_off[node * s + 1] - row_off[node * s];
        bool ok = (len0 % (uint64_t)s) == 0;
        for (int i = 1; i < s; ++i) ok = ok && (row_off[node * s + i + 1] - row_off[node * s + i] == len0);
        if (!ok) atomicExch(bad, 1);
        cnt[node] = (int64_t)(len0 / (uint64_t)s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[num_nodes] = 0;
}

__global__ void adopt_cols_kernel(const uint64_t* __restrict__ row_off, const uint64_t* __restrict__ cols, uint64_t num_nodes, int s,
                                  const int64_t* __restrict__ blk_off, int32_t* __restrict__ blk_cols, int* bad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t node = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; node < num_nodes; node += stride) {
        const int64_t cnt = blk_off[node + 1] - blk_off[node];
        bool ok = true;
        for (int i = 0; i < s; ++i) {
            const uint64_t r0 = row_off[node * s + i];
            for (int64_t k = 0; k < cnt; ++k) {
                const uint64_t c0 = cols[r0 + k * s];
                ok = ok && (c0 % (uint64_t)s == 0) && (c0 / (uint64_t)s < num_nodes);
                for (int j = 1; j < s; ++j) ok = ok && (cols[r0 + k * s + j] == c0 + j);
                if (i == 0) {
                    blk_cols[blk_off[node] + k] = (int32_t)(c0 / (uint64_t)s);
                    if (k > 0) ok = ok && (cols[r0 + (k - 1) * s] < c0);
                } else {
                    ok = ok && (cols[row_off[node * s] + k * s] == c0);
                }
            }
        }
        if (!ok) atomicExch(bad, 1);
    }
}

// ------------------------------------------------------------------------------------------------ colour validation
__global__ void check_color_kernel(ConnView cv, const int32_t* __restrict__ elems, uint64_t count, int color, int* stamp,
                                   unsigned long long* errword) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const int32_t e = elems[k];
        int64_t b, en;
        if (cv.elem_off) { b = cv.elem_off[e]; en = cv.elem_off[e + 1]; } else { b = (int64_t)e * cv.n; en = b + cv.n; }
        for (int64_t p = b; p < en; ++p) {
            // a node repeated INSIDE one element is not a conflict between elements: tag with the element too
            const int node = cv.conn[p];
            bool seen_in_self = false;
            for (int64_t q = b; q < p; ++q) seen_in_self = seen_in_self || (cv.conn[q] == node);
            if (seen_in_self) continue;
            const int prev = atomicExch(&stamp[node], color);
            if (prev == color) atomicMin(errword, ((unsigned long long)e << 8) | (unsigned long long)FB200_ERR_COLORING);
        }
    }
}

static ConnView conn_view(const fb200_ctx* ctx) {
    ConnView cv;
    cv.conn = ctx->d_conn;
    cv.elem_off = ctx->ragged ? ctx->d_elem_off : nullptr;
    cv.n = ctx->ragged ? 0 : ctx->ei.n;
    cv.num_elements = ctx->E;
    return cv;
}

static fb200_status upload_colors(fb200_ctx* ctx) {
    free_ordered(ctx);
    dev_free(ctx->d_color_elems);
    const uint64_t total = ctx->h_color_elems.size();
    FB200_TRY(dev_alloc(ctx, &ctx->d_color_elems, total));
    std::vector<int32_t> tmp(total);
    for (uint64_t i = 0; i < total; ++i) tmp[i] = (int32_t)ctx->h_color_elems[i];
    if (total) FB200_CUDA(ctx, h2d_copy(ctx, ctx->d_color_elems, tmp.data(), total * sizeof(int32_t)));
    return FB200_OK;
}

}  // namespace fb200

using namespace fb200;

extern "C" {

fb200_status fb200_assemble_pattern(fb200_ctx* ctx, int32_t sdim, uint64_t* num_rows, uint64_t* nnz) {
    if (!ctx || !ctx->has_connectivity) return fail(ctx, FB200_ERR_STATE, "assemble_pattern needs a space or connectivity");
    if (sdim < 1 || sdim > 3) return fail(ctx, FB200_ERR_SHAPE, "solution_dim must be 1, 2 or 3");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    free_pattern(ctx);
    struct PatternLap {
        SetupTimer tm;
        cudaStream_t s;
        ~PatternLap() {
            if (tm.on) cudaStreamSynchronize(s);
            tm.lap("assemble_pattern (device)");
        }
    } pattern_lap{SetupTimer(), ctx->stream};
    FB200_TRY(build_adjacency(ctx));
    const uint64_t N = ctx->N;
    const ConnView cv = conn_view(ctx);

    int* d_max = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_max, 1));
    cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream);
    if (N) {
        candidate_count_kernel<<<(int)std::min<uint64_t>(div_up(N, 256), 2048), 256, 0, ctx->stream>>>(cv, ctx->d_adj_off, ctx->d_adj_inc, N, d_max);
        ctx->launches++;
    }
    int max_m = 0;
    cudaMemcpyAsync(&max_m, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_max);
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, FB200_ERR_CUDA, "candidate_count_kernel failed");

    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(N, kSetWarps), (uint64_t)ctx->sm_count * 8));
    int32_t* d_scratch = nullptr;
    int per_warp = 0;
    if (max_m > kSetSmemCap) {
        per_warp = max_m;
        FB200_TRY(dev_alloc(ctx, &d_scratch, (uint64_t)blocks * kSetWarps * (uint64_t)per_warp));
    }
    FB200_TRY(dev_alloc(ctx, &ctx->d_blk_off, N + 1));
    FB200_CUDA(ctx, cudaMemsetAsync(ctx->d_blk_off, 0, (N + 1) * sizeof(int64_t), ctx->stream));
    fb200_status st = FB200_OK;
    if (N) {
        node_sets_kernel<false><<<blocks, kSetWarps * 32, 0, ctx->stream>>>(cv, ctx->d_adj_off, ctx->d_adj_inc, N, d_scratch, per_warp,
                                                                           ctx->d_blk_off, nullptr, nullptr);
        st = check_launch(ctx, "node_sets_kernel<count>");
    }
    if (st == FB200_OK) st = exclusive_scan_i64(ctx, ctx->d_blk_off, N + 1);
    int64_t P = 0;
    if (st == FB200_OK) {
        cudaMemcpyAsync(&P, ctx->d_blk_off + N, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FB200_ERR_CUDA, "pattern count failed");
    }
    if (st == FB200_OK) {
        ctx->P = (uint64_t)P;
        st = dev_alloc(ctx, &ctx->d_blk_cols, ctx->P);
    }
    if (st == FB200_OK && N) {
        node_sets_kernel<true><<<blocks, kSetWarps * 32, 0, ctx->stream>>>(cv, ctx->d_adj_off, ctx->d_adj_inc, N, d_scratch, per_warp,
                                                                          nullptr, ctx->d_blk_off, ctx->d_blk_cols);
        st = check_launch(ctx, "node_sets_kernel<fill>");
    }
    cudaStreamSynchronize(ctx->stream);
    if (d_scratch) cudaFree(d_scratch);
    if (st == FB200_OK) st = finish_pattern(ctx, sdim);
    if (st != FB200_OK) {
        free_pattern(ctx);
        return st;
    }
    if (num_rows) *num_rows = ctx->nrows;
    if (nnz) *nnz = ctx->nnz;
    return FB200_OK;
}

fb200_status fb200_pattern_download(fb200_ctx* ctx, uint64_t* row_offsets, uint64_t* col_indices) {
    if (!ctx || !ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const int s = ctx->sdim;
    if (row_offsets) {
        uint64_t* d_ro = nullptr;
        FB200_TRY(dev_alloc(ctx, &d_ro, ctx->nrows + 1));
        row_offsets_kernel<<<(int)std::min<uint64_t>(div_up(ctx->nrows + 1, 256), 4096), 256, 0, ctx->stream>>>(ctx->d_blk_off, ctx->N, s, d_ro);
        fb200_status st = check_launch(ctx, "row_offsets_kernel");
        if (st == FB200_OK) {
            cudaError_t e = cudaMemcpyAsync(row_offsets, d_ro, (ctx->nrows + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H row_offsets");
        }
        cudaFree(d_ro);
        FB200_TRY(st);
    }
    if (col_indices && ctx->nnz) {
        // stream the expansion through a bounded staging buffer
        std::vector<int64_t> h_off(ctx->N + 1);
        FB200_CUDA(ctx, d2h_copy(ctx, h_off.data(), ctx->d_blk_off, (ctx->N + 1) * sizeof(int64_t)));
        const uint64_t chunk_entries = 32ull << 20;  // 256 MiB of u64
        uint64_t* d_stage = nullptr;
        uint64_t max_node_entries = 0;
        for (uint64_t nd = 0; nd < ctx->N; ++nd)
            max_node_entries = std::max<uint64_t>(max_node_entries, (uint64_t)(s * s) * (uint64_t)(h_off[nd + 1] - h_off[nd]));
        const uint64_t cap = std::max(chunk_entries, max_node_entries);
        FB200_TRY(dev_alloc(ctx, &d_stage, cap));
        uint64_t node0 = 0;
        fb200_status st = FB200_OK;
        while (node0 < ctx->N && st == FB200_OK) {
            uint64_t node1 = node0;
            const uint64_t base = (uint64_t)(s * s) * (uint64_t)h_off[node0];
            while (node1 < ctx->N && (uint64_t)(s * s) * (uint64_t)h_off[node1 + 1] - base <= cap) ++node1;
            const uint64_t entries = (uint64_t)(s * s) * (uint64_t)h_off[node1] - base;
            if (entries) {
                const int blocks = (int)std::min<uint64_t>(div_up((node1 - node0) * 32, 256), (uint64_t)ctx->sm_count * 16);
                col_indices_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_blk_off, ctx->d_blk_cols, node0, node1, s, d_stage);
                st = check_launch(ctx, "col_indices_kernel");
                if (st == FB200_OK) {
                    cudaError_t e = cudaMemcpyAsync(col_indices + base, d_stage, entries * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
                    if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H col_indices");
                }
            }
            node0 = node1;
        }
        cudaFree(d_stage);
        FB200_TRY(st);
    }
    return FB200_OK;
}

fb200_status fb200_pattern_adopt(fb200_ctx* ctx, int32_t sdim, uint64_t num_rows, const uint64_t* row_offsets, const uint64_t* col_indices) {
    if (!ctx || !ctx->has_connectivity) return fail(ctx, FB200_ERR_STATE, "pattern_adopt needs a space or connectivity");
    if (sdim < 1 || sdim > 3) return fail(ctx, FB200_ERR_SHAPE, "solution_dim must be 1, 2 or 3");
    if (num_rows != (uint64_t)sdim * ctx->N) return fail(ctx, FB200_ERR_SHAPE, "num_rows != solution_dim * num_nodes");
    if (!row_offsets) return fail(ctx, FB200_ERR_SHAPE, "null row_offsets");
    // a malformed offset array would send the adopt kernels out of bounds: check it on the host first
    if (!offsets_well_formed(row_offsets, num_rows)) return fail(ctx, FB200_ERR_SHAPE, "row_offsets must start at 0 and be non-decreasing");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    free_pattern(ctx);
    const uint64_t N = ctx->N, nnz = row_offsets[num_rows];
    if (nnz && !col_indices) return fail(ctx, FB200_ERR_SHAPE, "null col_indices");
    uint64_t *d_ro = nullptr, *d_ci = nullptr;
    int* d_bad = nullptr;
    fb200_status st = dev_alloc(ctx, &d_ro, num_rows + 1);
    if (st == FB200_OK) st = dev_alloc(ctx, &d_ci, nnz);
    if (st == FB200_OK) st = dev_alloc(ctx, &d_bad, 1);
    if (st == FB200_OK) st = dev_alloc(ctx, &ctx->d_blk_off, N + 1);
    int bad = 0;
    if (st == FB200_OK) {
        cudaMemcpyAsync(d_ro, row_offsets, (num_rows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (nnz) cudaMemcpyAsync(d_ci, col_indices, nnz * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
        cudaMemsetAsync(ctx->d_blk_off, 0, (N + 1) * sizeof(int64_t), ctx->stream);
        if (N) {
            adopt_counts_kernel<<<(int)std::min<uint64_t>(div_up(N, 256), 2048), 256, 0, ctx->stream>>>(d_ro, N, sdim, ctx->d_blk_off, d_bad);
            st = check_launch(ctx, "adopt_counts_kernel");
        }
    }
    if (st == FB200_OK) st = exclusive_scan_i64(ctx, ctx->d_blk_off, N + 1);
    int64_t P = 0;
    if (st == FB200_OK) {
        cudaMemcpyAsync(&P, ctx->d_blk_off + N, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FB200_ERR_CUDA, "adopt failed");
        if (st == FB200_OK && (bad || (uint64_t)P * sdim * sdim != nnz))
            st = fail(ctx, FB200_ERR_UNSUPPORTED, "adopted pattern is not node-block structured");
    }
    if (st == FB200_OK) {
        ctx->P = (uint64_t)P;
        st = dev_alloc(ctx, &ctx->d_blk_cols, ctx->P);
    }
    if (st == FB200_OK && N) {
        adopt_cols_kernel<<<(int)std::min<uint64_t>(div_up(N, 128), 4096), 128, 0, ctx->stream>>>(d_ro, d_ci, N, sdim, ctx->d_blk_off, ctx->d_blk_cols, d_bad);
        st = check_launch(ctx, "adopt_cols_kernel");
        cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FB200_ERR_CUDA, "adopt failed");
        if (st == FB200_OK && bad) st = fail(ctx, FB200_ERR_UNSUPPORTED, "adopted pattern is not node-block structured / sorted");
    }
    cudaStreamSynchronize(ctx->stream);
    if (d_ro) cudaFree(d_ro);
    if (d_ci) cudaFree(d_ci);
    if (d_bad) cudaFree(d_bad);
    if (st == FB200_OK) st = finish_pattern(ctx, sdim);
    if (st != FB200_OK) {
        free_pattern(ctx);
        return st;
    }
    ctx->adopted = true;
    return FB200_OK;
}

// ------------------------------------------------------------------------------------------------ colours
static fb200_status validate_and_upload_colors(fb200_ctx* ctx) {
    FB200_TRY(upload_colors(ctx));
    int* d_stamp = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_stamp, ctx->N));
    cudaMemsetAsync(d_stamp, 0xff, ctx->N * sizeof(int), ctx->stream);
    const ConnView cv = conn_view(ctx);
    fb200_status st = FB200_OK;
    const uint64_t ncol = ctx->h_color_off.size() - 1;
    for (uint64_t c = 0; c < ncol && st == FB200_OK; ++c) {
        const uint64_t b = ctx->h_color_off[c], cnt = ctx->h_color_off[c + 1] - b;
        if (!cnt) continue;
        check_color_kernel<<<(int)std::min<uint64_t>(div_up(cnt, 256), 2048), 256, 0, ctx->stream>>>(cv, ctx->d_color_elems + b, cnt, (int)c,
                                                                                                 d_stamp, ctx->d_errword);
        st = check_launch(ctx, "check_color_kernel");
    }
    if (st == FB200_OK) st = read_errword(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_stamp);
    return st;
}

fb200_status fb200_color_nodes(fb200_ctx* ctx, uint64_t* num_colors) {
    if (!ctx || !ctx->has_connectivity) return fail(ctx, FB200_ERR_STATE, "color_nodes needs a space or connectivity");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    // The reference colours serially on the host (sequential_greedy_coloring); so do we, from the device copy.
    const uint64_t E = ctx->E_owned;
    std::vector<int32_t> nodes(ctx->conn_len);
    if (ctx->conn_len) FB200_CUDA(ctx, d2h_copy(ctx, nodes.data(), ctx->d_conn, ctx->conn_len * sizeof(int32_t)));
    std::vector<int64_t> off(ctx->E + 1);
    if (ctx->ragged) {
        FB200_CUDA(ctx, d2h_copy(ctx, off.data(), ctx->d_elem_off, (ctx->E + 1) * sizeof(int64_t)));
    } else {
        for (uint64_t e = 0; e <= ctx->E; ++e) off[e] = (int64_t)(e * (uint64_t)ctx->ei.n);
    }
    greedy_coloring(E, ctx->N, off, nodes, ctx->h_color_off, ctx->h_color_elems);
    fb200_status st = validate_and_upload_colors(ctx);
    if (st != FB200_OK) return st;
    ctx->has_colors = true;
    if (num_colors) *num_colors = ctx->h_color_off.size() - 1;
    return FB200_OK;
}

fb200_status fb200_colors_download(fb200_ctx* ctx, uint64_t* color_offsets, uint64_t* element_ids) {
    if (!ctx || !ctx->has_colors) return fail(ctx, FB200_ERR_STATE, "no colours");
    if (color_offsets) std::copy(ctx->h_color_off.begin(), ctx->h_color_off.end(), color_offsets);
    if (element_ids) std::copy(ctx->h_color_elems.begin(), ctx->h_color_elems.end(), element_ids);
    return FB200_OK;
}

fb200_status fb200_colors_adopt(fb200_ctx* ctx, uint64_t num_colors, const uint64_t* color_offsets, const uint64_t* element_ids) {
    if (!ctx || !ctx->has_connectivity) return fail(ctx, FB200_ERR_STATE, "colors_adopt needs a space or connectivity");
    if (!color_offsets) return fail(ctx, FB200_ERR_SHAPE, "null colour offsets");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!offsets_well_formed(color_offsets, num_colors)) return fail(ctx, FB200_ERR_SHAPE, "colour offsets must start at 0 and be non-decreasing");
    const uint64_t total = color_offsets[num_colors];
    if (total && !element_ids) return fail(ctx, FB200_ERR_SHAPE, "null colour element list");
    for (uint64_t k = 0; k < total; ++k)
        if (element_ids[k] >= ctx->E_owned) return fail(ctx, FB200_ERR_INDEX_OOB, "colour lists an element that is not owned", (int64_t)element_ids[k]);
    ctx->h_color_off.assign(color_offsets, color_offsets + num_colors + 1);
    ctx->h_color_elems.assign(element_ids, element_ids + total);
    ctx->has_colors = false;
    fb200_status st = validate_and_upload_colors(ctx);
    if (st != FB200_OK) return st;
    ctx->has_colors = true;
    return FB200_OK;
}

}  // extern "C"
