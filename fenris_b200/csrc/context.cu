// Context lifetime, error plumbing, stream/timer helpers and the space (mesh) upload.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "fb200_internal.h"

namespace fb200 {

fb200_status fail(fb200_ctx* ctx, fb200_status s, const std::string& msg, int64_t elem) {
    if (ctx) {
        ctx->err = msg;
        ctx->err_elem = elem;
    }
    return s;
}

fb200_status cuda_fail(fb200_ctx* ctx, cudaError_t e, const char* what) {
    std::string m = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear sticky-less errors
    return fail(ctx, FB200_ERR_CUDA, m);
}

fb200_status check_launch(fb200_ctx* ctx, const char* name) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, name);
    return FB200_OK;
}

void free_pattern(fb200_ctx* ctx) {
    p2p_disable(ctx);  // the neighbours' row offsets and the exported values belong to this pattern
    free_ordered(ctx);
    dev_free(ctx->d_blk_off);
    dev_free(ctx->d_blk_cols);
    dev_free(ctx->d_blockmap);
    dev_free(ctx->d_values);
    ctx->values_capacity = 0;
    ctx->has_pattern = false;
    ctx->adopted = false;
    ctx->P = ctx->nrows = ctx->nnz = 0;
    ctx->sdim = 0;
}

// Locality-preserving visiting order: elements sorted by the Morton code of their centroid (host preprocessing, like
// the colouring).  The assembled sums do not depend on the order; it only decides which CSR rows are live in L2 together and
// which elements share a tile of the tile-accumulating Hex8 kernel.  Centroids are quantised with the mean element size
// h = (volume of the bounding box / E)^(1/d), so that on a structured mesh one quantisation cell is exactly one element and
// aligned 4 x 4 x 4 element blocks are consecutive in the order (the code drops its lowest 2 d bits to name such a block).
void morton_order(int d, int n, uint64_t N, const double* v, uint64_t E, const uint64_t* conn, std::vector<int32_t>& order,
                         std::vector<uint64_t>& codes) {
    order.resize(E);
    codes.resize(E);
    if (E == 0 || N == 0) return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint64_t i = 0; i < N; ++i)
        for (int k = 0; k < d; ++k) {
            lo[k] = std::min(lo[k], v[i * d + k]);
            hi[k] = std::max(hi[k], v[i * d + k]);
        }
    const int bits = d == 2 ? 31 : 21;
    const double qmax = (double)((1ull << bits) - 1);
    double vol = 1.0;
    int dims = 0;
    for (int k = 0; k < d; ++k)
        if (hi[k] > lo[k]) {
            vol *= hi[k] - lo[k];
            ++dims;
        }
    const double h = dims ? std::pow(vol / (double)E, 1.0 / dims) : 1.0;
    const double inv = h > 0.0 ? 1.0 / h : 0.0;
    // codes in parallel (the centroid gather is the expensive part), then a stable LSD radix sort of (code, element) - ties keep the
    // element order, exactly like sorting the pairs (code, e)
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const unsigned nthreads = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, E / 65536));
    auto code_range = [&](uint64_t e0, uint64_t e1) {
        for (uint64_t e = e0; e < e1; ++e) {
            uint64_t q[3] = {0, 0, 0};
            for (int k = 0; k < d; ++k) {
                double c = 0.0;
                for (int a = 0; a < n; ++a) c += v[conn[e * n + a] * d + k];
                c = (c / n - lo[k]) * inv;
                q[k] = (uint64_t)std::min(std::max(c, 0.0), qmax);
            }
            uint64_t code = 0;
            for (int b = bits - 1; b >= 0; --b)
                for (int k = d - 1; k >= 0; --k) code = (code << 1) | ((q[k] >> b) & 1u);
            codes[e] = code;
        }
    };
    {
        std::vector<std::thread> pool;
        const uint64_t per = (E + nthreads - 1) / nthreads;
        for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(code_range, std::min<uint64_t>(E, t * per), std::min<uint64_t>(E, (t + 1) * per));
        code_range(0, std::min<uint64_t>(E, per));
        for (auto& th : pool) th.join();
    }
    uint64_t all_or = 0;
    for (uint64_t e = 0; e < E; ++e) all_or |= codes[e];
    std::vector<uint64_t> ka(codes), kb(E);
    std::vector<int32_t> ia(E), ib(E);
    for (uint64_t e = 0; e < E; ++e) ia[e] = (int32_t)e;
    for (int shift = 0; shift < 64 && (all_or >> shift) != 0; shift += 11) {
        uint64_t count[2049] = {0};
        for (uint64_t e = 0; e < E; ++e) ++count[((ka[e] >> shift) & 2047u) + 1];
        for (int b = 0; b < 2048; ++b) count[b + 1] += count[b];
        for (uint64_t e = 0; e < E; ++e) {
            const uint64_t pos = count[(ka[e] >> shift) & 2047u]++;
            kb[pos] = ka[e];
            ib[pos] = ia[e];
        }
        ka.swap(kb);
        ia.swap(ib);
    }
    order.swap(ia);
    codes.swap(ka);
}

static fb200_status upload_order(fb200_ctx* ctx) {
    free_ordered(ctx);
    dev_free(ctx->d_order);
    ctx->order_count = 0;
    if (ctx->h_order.empty()) return FB200_OK;
    std::vector<int32_t> owned;
    owned.reserve(ctx->E_owned);
    ctx->h_order_codes.clear();
    ctx->h_order_codes.reserve(ctx->E_owned);
    for (size_t i = 0; i < ctx->h_order.size(); ++i) {
        const int32_t e = ctx->h_order[i];
        if ((uint64_t)e < ctx->E_owned) {
            owned.push_back(e);
            ctx->h_order_codes.push_back(ctx->h_order_codes_all[i]);
        }
    }
    FB200_TRY(dev_alloc(ctx, &ctx->d_order, owned.size()));
    if (!owned.empty()) FB200_CUDA(ctx, h2d_copy(ctx, ctx->d_order, owned.data(), owned.size() * sizeof(int32_t)));
    ctx->order_count = owned.size();
    return FB200_OK;
}

void free_space(fb200_ctx* ctx) {
    free_pattern(ctx);
    dev_free(ctx->d_order);
    ctx->order_count = 0;
    ctx->h_order.clear();
    ctx->h_order_codes_all.clear();
    ctx->h_order_codes.clear();
    dev_free(ctx->d_row_epoch);
    dev_free(ctx->d_vertices);
    dev_free(ctx->d_conn);
    dev_free(ctx->d_elem_off);
    dev_free(ctx->d_adj_off);
    dev_free(ctx->d_adj_inc);
    dev_free(ctx->d_color_elems);
    ctx->h_color_off.clear();
    ctx->h_color_elems.clear();
    ctx->has_colors = false;
    ctx->has_space = ctx->has_connectivity = ctx->ragged = false;
    ctx->N = ctx->E = ctx->E_owned = ctx->conn_len = 0;
}

// Narrow the caller's usize connectivity to int32 and validate it against num_nodes.
__global__ void narrow_indices_kernel(const uint64_t* __restrict__ in, int32_t* __restrict__ out, uint64_t count, uint64_t num_nodes,
                                      uint64_t per_element, unsigned long long* errword) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint64_t v = in[i];
        if (v >= num_nodes) {
            const unsigned long long elem = per_element ? i / per_element : i;
            atomicMin(errword, (elem << 8) | (unsigned long long)FB200_ERR_INDEX_OOB);
            out[i] = 0;
        } else {
            out[i] = (int32_t)v;
        }
    }
}

__global__ void narrow_offsets_kernel(const uint64_t* __restrict__ in, int64_t* __restrict__ out, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = (int64_t)in[i];
}

cudaError_t d2h_staged(fb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    constexpr size_t kChunk = 16u << 20;
    if (bytes < 4 * kChunk) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
    }
    for (int b = 0; b < 2; ++b) {
        if (!ctx->h_stage[b]) {
            cudaError_t e = cudaMallocHost(&ctx->h_stage[b], kChunk);
            if (e == cudaSuccess && !ctx->ev_stage[b]) e = cudaEventCreateWithFlags(&ctx->ev_stage[b], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
    }
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    auto issue = [&](size_t k) -> cudaError_t {
        const size_t off = k * kChunk, len = std::min(kChunk, bytes - off);
        cudaError_t e = cudaMemcpyAsync(ctx->h_stage[k & 1], static_cast<const char*>(src) + off, len, cudaMemcpyDeviceToHost, ctx->stream);
        return e != cudaSuccess ? e : cudaEventRecord(ctx->ev_stage[k & 1], ctx->stream);
    };
    cudaError_t e = issue(0);
    for (size_t k = 0; k < nchunks && e == cudaSuccess; ++k) {
        if (k + 1 < nchunks) e = issue(k + 1);  // (buffer (k + 1) & 1 was drained in iteration k - 1)
        if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev_stage[k & 1]);
        if (e == cudaSuccess) {
            // the destination is usually fresh memory: its first touch (page faults) is the slow part, so several threads copy
            const size_t off = k * kChunk, len = std::min(kChunk, bytes - off);
            constexpr int kCopyThreads = 4;
            const size_t part = (len / kCopyThreads + 4095) & ~(size_t)4095;
            std::thread th[kCopyThreads - 1];
            for (int t = 1; t < kCopyThreads; ++t) {
                const size_t b0 = std::min(len, t * part), b1 = std::min(len, (t + 1) * part);
                th[t - 1] = std::thread([=]() { std::memcpy(static_cast<char*>(dst) + off + b0, static_cast<const char*>(ctx->h_stage[k & 1]) + b0, b1 - b0); });
            }
            std::memcpy(static_cast<char*>(dst) + off, ctx->h_stage[k & 1], std::min(len, part));
            for (auto& t : th) t.join();
        }
    }
    return e;
}

fb200_status read_errword(fb200_ctx* ctx) {
    FB200_CUDA(ctx, cudaMemcpyAsync(ctx->h_errword, ctx->d_errword, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long w = *ctx->h_errword;
    if (w == kNoError) return FB200_OK;
    // re-arm
    const unsigned long long init = kNoError;
    FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_errword, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const fb200_status code = (fb200_status)(w & 0xffull);
    const int64_t elem = (int64_t)(w >> 8);
    switch (code) {
        case FB200_ERR_SINGULAR_JACOBIAN: return fail(ctx, code, "Singular element Jacobian encountered", elem);
        case FB200_ERR_INDEX_OOB: return fail(ctx, code, "connectivity index out of bounds (>= num_nodes)", elem);
        case FB200_ERR_COLUMN_NOT_IN_PATTERN:
            return fail(ctx, code, "Could not find column index associated with node in CSR row", elem);
        case FB200_ERR_COLORING: return fail(ctx, code, "colour contains two elements sharing a node", elem);
        default: return fail(ctx, code ? code : FB200_ERR_CUDA, "device reported an error", elem);
    }
}

}  // namespace fb200

using namespace fb200;

extern "C" {

int32_t fb200_abi_version(void) { return FB200_VERSION; }

const char* fb200_status_string(fb200_status s) {
    switch (s) {
        case FB200_OK: return "ok";
        case FB200_ERR_SINGULAR_JACOBIAN: return "singular element Jacobian";
        case FB200_ERR_SHAPE: return "shape mismatch";
        case FB200_ERR_INDEX_OOB: return "index out of bounds";
        case FB200_ERR_COLUMN_NOT_IN_PATTERN: return "column not in pattern";
        case FB200_ERR_UNSUPPORTED: return "unsupported (no CPU fallback)";
        case FB200_ERR_CUDA: return "CUDA error";
        case FB200_ERR_NCCL: return "NCCL error";
        case FB200_ERR_STATE: return "invalid call order";
        case FB200_ERR_COLORING: return "colours are not disjoint";
        case FB200_ERR_NOT_CONVERGED: return "maximum number of iterations reached";
        case FB200_ERR_INDEFINITE: return "indefinite operator or preconditioner";
        default: return "unknown status";
    }
}

fb200_status fb200_create(int32_t device, fb200_ctx** out) {
    if (!out) return FB200_ERR_SHAPE;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return FB200_ERR_CUDA;  // no GPU: there is deliberately no CPU fallback
    }
    fb200_ctx* ctx = new fb200_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return FB200_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaMalloc((void**)&ctx->d_errword, sizeof(unsigned long long));
    cudaMalloc((void**)&ctx->d_ticket, 16 * sizeof(unsigned long long));  // [0] work counter; [1..15] debug counters (FB200_DEBUG & 64)
    cudaMemset(ctx->d_ticket, 0, 16 * sizeof(unsigned long long));
    cudaMallocHost((void**)&ctx->h_errword, sizeof(unsigned long long));
    const unsigned long long init = kNoError;
    h2d_copy(ctx, ctx->d_errword, &init, sizeof(init));
    if (cudaGetLastError() != cudaSuccess) {
        fb200_destroy(ctx);
        return FB200_ERR_CUDA;
    }
    *out = ctx;
    return FB200_OK;
}

// comm.cu
void fb200_comm_destroy_internal(fb200_ctx* ctx);

void fb200_destroy(fb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    fb200_comm_destroy_internal(ctx);
    free_space(ctx);
    dev_free(ctx->tab.d_data);
    for (int b = 0; b < 2; ++b) {
        if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
        if (ctx->ev_stage[b]) cudaEventDestroy(ctx->ev_stage[b]);
    }
    dev_free(ctx->d_errword);
    dev_free(ctx->d_ticket);
    dev_free(ctx->d_iface_nodes);
    dev_free(ctx->d_iface_offsets);
    dev_free(ctx->d_iface_packed);
    dev_free(ctx->d_ms_tab);
    dev_free(ctx->d_vector);
    dev_free(ctx->d_source);
    ctx->ms_tab_capacity = 0;
    ctx->h_ms_tab.clear();
    ctx->vector_len = ctx->vector_capacity = ctx->source_capacity = 0;
    dev_free(ctx->d_peer_nodes);
    dev_free(ctx->d_peer_offsets);
    dev_free(ctx->d_peer_send);
    dev_free(ctx->d_peer_recv);
    ctx->peer_ranks.clear();
    ctx->peer_seg_off.clear();
    ctx->peer_count = 0;
    if (ctx->h_errword) cudaFreeHost(ctx->h_errword);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

fb200_status fb200_last_error(fb200_ctx* ctx, char* buf, size_t len, int64_t* element_index) {
    if (!ctx) return FB200_ERR_STATE;
    if (buf && len) {
        std::strncpy(buf, ctx->err.c_str(), len - 1);
        buf[len - 1] = 0;
    }
    if (element_index) *element_index = ctx->err_elem;
    return FB200_OK;
}

fb200_status fb200_set_stream(fb200_ctx* ctx, void* s) {
    if (!ctx) return FB200_ERR_STATE;
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return FB200_OK;
}

fb200_status fb200_synchronize(fb200_ctx* ctx) {
    if (!ctx) return FB200_ERR_STATE;
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    return read_errword(ctx);
}

fb200_status fb200_timer_begin(fb200_ctx* ctx) {
    if (!ctx) return FB200_ERR_STATE;
    FB200_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return FB200_OK;
}

fb200_status fb200_timer_end(fb200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return FB200_ERR_STATE;
    FB200_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    FB200_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    FB200_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return FB200_OK;
}

uint64_t fb200_launch_count(fb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

fb200_status fb200_set_tuning(fb200_ctx* ctx, const char* name, int32_t value) {
    if (!ctx) return FB200_ERR_STATE;
    if (name && std::strcmp(name, "hex8_tile") == 0) {
        if (value != 0 && value != 64) return fail(ctx, FB200_ERR_SHAPE, "hex8_tile must be 0 or 64");
        ctx->tune_hex8_tile = value;
        return FB200_OK;
    }
    if (name && std::strcmp(name, "hex8_colored_tiles") == 0) {
        if (value != 0 && value != 1) return fail(ctx, FB200_ERR_SHAPE, "hex8_colored_tiles must be 0 or 1");
        ctx->tune_colored_tiles = value;
        return FB200_OK;
    }
    if (name && std::strcmp(name, "hex8_owner_stores") == 0) {
        if (value != 0 && value != 1) return fail(ctx, FB200_ERR_SHAPE, "hex8_owner_stores must be 0 or 1");
        ctx->tune_owner = value;
        return FB200_OK;
    }
    return fail(ctx, FB200_ERR_UNSUPPORTED, "unknown tuning knob");
}

static fb200_status upload_indices(fb200_ctx* ctx, const uint64_t* host, uint64_t count, uint64_t per_element, int32_t** d_out) {
    uint64_t* d_tmp = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_tmp, count));
    fb200_status st = dev_alloc(ctx, d_out, count);
    if (st == FB200_OK && count) {
        cudaError_t e = cudaMemcpyAsync(d_tmp, host, count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "H2D connectivity");
        if (st == FB200_OK) {
            const int blocks = (int)std::min<uint64_t>(div_up(count, 256), 148 * 16);
            narrow_indices_kernel<<<blocks, 256, 0, ctx->stream>>>(d_tmp, *d_out, count, ctx->N, per_element, ctx->d_errword);
            st = check_launch(ctx, "narrow_indices_kernel");
        }
        if (st == FB200_OK) st = read_errword(ctx);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    return st;
}

fb200_status fb200_space_upload(fb200_ctx* ctx, int32_t element_type, uint64_t num_nodes, const double* vertices, uint64_t num_elements,
                                const uint64_t* connectivity) {
    if (!ctx) return FB200_ERR_STATE;
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    ElementInfo ei;
    if (!element_info(element_type, &ei)) return fail(ctx, FB200_ERR_UNSUPPORTED, "unknown element type");
    if ((num_nodes && !vertices) || (num_elements && !connectivity)) return fail(ctx, FB200_ERR_SHAPE, "null mesh arrays");
    if (num_nodes >= (1ull << 31) || num_elements >= (1ull << 31) / (uint64_t)(ei.n))
        return fail(ctx, FB200_ERR_UNSUPPORTED, "mesh too large for 32-bit device indices");
    free_space(ctx);
    SetupTimer tm;
    ctx->elem_type = element_type;
    ctx->ei = ei;
    ctx->N = num_nodes;
    ctx->E = ctx->E_owned = num_elements;
    ctx->conn_len = num_elements * (uint64_t)ei.n;
    FB200_TRY(dev_alloc(ctx, &ctx->d_vertices, num_nodes * ei.d));
    if (num_nodes)
        FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_vertices, vertices, num_nodes * ei.d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    fb200_status st = upload_indices(ctx, connectivity, ctx->conn_len, ei.n, &ctx->d_conn);
    if (st != FB200_OK) {
        free_space(ctx);
        return st;
    }
    ctx->has_space = ctx->has_connectivity = true;
    ctx->ragged = false;
    tm.lap("space_upload: H2D vertices + connectivity");
    morton_order(ei.d, ei.n, num_nodes, vertices, num_elements, connectivity, ctx->h_order, ctx->h_order_codes_all);  // indices were validated above
    tm.lap("space_upload: Morton order (host)");
    const fb200_status so = upload_order(ctx);
    tm.lap("space_upload: upload order");
    return so;
}

fb200_status fb200_space_update_vertices(fb200_ctx* ctx, const double* vertices) {
    if (!ctx || !ctx->has_space) return fail(ctx, FB200_ERR_STATE, "no space uploaded");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->N)
        FB200_CUDA(ctx, cudaMemcpyAsync(ctx->d_vertices, vertices, ctx->N * ctx->ei.d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return FB200_OK;
}

fb200_status fb200_connectivity_upload(fb200_ctx* ctx, uint64_t num_nodes, uint64_t num_elements, const uint64_t* element_offsets,
                                       const uint64_t* element_nodes) {
    if (!ctx) return FB200_ERR_STATE;
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!element_offsets) return fail(ctx, FB200_ERR_SHAPE, "null offsets");
    if (!offsets_well_formed(element_offsets, num_elements)) return fail(ctx, FB200_ERR_SHAPE, "offsets must start at 0 and be non-decreasing");
    const uint64_t total = element_offsets[num_elements];
    if (num_nodes >= (1ull << 31) || num_elements >= (1ull << 31) || total >= (1ull << 31))
        return fail(ctx, FB200_ERR_UNSUPPORTED, "connectivity too large for 32-bit device indices");
    free_space(ctx);
    ctx->elem_type = 0;
    ctx->ei = {0, 0, 0};
    ctx->N = num_nodes;
    ctx->E = ctx->E_owned = num_elements;
    ctx->conn_len = total;
    uint64_t* d_tmp = nullptr;
    FB200_TRY(dev_alloc(ctx, &d_tmp, num_elements + 1));
    fb200_status st = dev_alloc(ctx, &ctx->d_elem_off, num_elements + 1);
    if (st == FB200_OK) {
        cudaMemcpyAsync(d_tmp, element_offsets, (num_elements + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        narrow_offsets_kernel<<<(int)std::min<uint64_t>(div_up(num_elements + 1, 256), 148 * 16), 256, 0, ctx->stream>>>(
            d_tmp, ctx->d_elem_off, num_elements + 1);
        st = check_launch(ctx, "narrow_offsets_kernel");
    }
    if (st == FB200_OK) st = upload_indices(ctx, element_nodes, total, 0, &ctx->d_conn);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    if (st != FB200_OK) {
        // element index for OOB in the ragged case: translate flat position -> element on the host
        if (st == FB200_ERR_INDEX_OOB && ctx->err_elem >= 0) {
            const uint64_t pos = (uint64_t)ctx->err_elem;
            uint64_t e = 0;
            while (e + 1 < num_elements && element_offsets[e + 1] <= pos) ++e;
            ctx->err_elem = (int64_t)e;
        }
        free_space(ctx);
        return st;
    }
    ctx->has_connectivity = true;
    ctx->has_space = false;
    ctx->ragged = true;
    return FB200_OK;
}

fb200_status fb200_set_num_owned_elements(fb200_ctx* ctx, uint64_t num_owned) {
    if (!ctx || !ctx->has_connectivity) return fail(ctx, FB200_ERR_STATE, "no connectivity uploaded");
    if (num_owned > ctx->E) return fail(ctx, FB200_ERR_SHAPE, "num_owned exceeds num_elements");
    ctx->E_owned = num_owned;
    // colours and the visiting order depend on the owned set
    ctx->has_colors = false;
    return upload_order(ctx);
}

}  // extern "C"
