// homogeneous Dirichlet conditions on the device-resident CSR - the step that follows assembly in every
// caller of the reference (examples/poisson2d.rs:82-83, tests/convergence_tests/poisson_mms_common.rs:136-137).
//
// Reference semantics, apply_homogeneous_dirichlet_bc_csr (src/assembly/global.rs:379-451):
//   scale = |first non-zero diagonal entry| in row order (1 if there is none)                          :388-397
//   rows of Dirichlet dofs: diagonal <- scale, everything else <- 0; every column c met there marks row c  :414-431
//   marked rows that are not Dirichlet rows: entries in Dirichlet columns <- 0                          :434-449
// (the "symmetric visit" trick: only rows coupled to a Dirichlet row are touched).  Dirichlet conditions are per NODE
// (all solution_dim dofs), and the CSR is node-block structured, so the kernels work on node blocks: one warp per node.
#include "fb200_internal.h"

namespace fb200 {

// smallest row index whose diagonal entry is non-zero
__global__ void first_diagonal_kernel(const int64_t* __restrict__ blk_off, const int32_t* __restrict__ blk_cols, uint64_t num_nodes, int s,
                                      const double* __restrict__ values, unsigned long long* first_row) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t I = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; I < num_nodes; I += stride) {
        const int64_t b = blk_off[I], e = blk_off[I + 1];
        int64_t lo = b, hi = e;  // position of I in its own (sorted) block row
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (blk_cols[mid] < (int32_t)I) lo = mid + 1; else hi = mid;
        }
        if (lo == e || blk_cols[lo] != (int32_t)I) continue;
        const int64_t cnt = e - b, k = lo - b;
        for (int i = 0; i < s; ++i) {
            const double v = values[(int64_t)(s * s) * b + (int64_t)i * s * cnt + s * k + i];
            if (v != 0.0) {
                atomicMin(first_row, (unsigned long long)(I * s + i));
                break;
            }
        }
    }
}

__global__ void read_diagonal_kernel(const int64_t* __restrict__ blk_off, const int32_t* __restrict__ blk_cols, int s, const double* __restrict__ values,
                                     const unsigned long long* __restrict__ first_row, double* scale) {
    const unsigned long long r = *first_row;
    if (r == ~0ull) {
        *scale = 1.0;
        return;
    }
    const int64_t I = (int64_t)(r / s);
    const int i = (int)(r % s);
    const int64_t b = blk_off[I], e = blk_off[I + 1];
    int64_t k = 0;
    while (blk_cols[b + k] != (int32_t)I) ++k;
    *scale = fabs(values[(int64_t)(s * s) * b + (int64_t)i * s * (e - b) + s * k + i]);
}

// pass 1 (one warp per Dirichlet node): its rows <- scale on the diagonal, 0 elsewhere; coupled nodes are marked for pass 2
__global__ void dirichlet_rows_kernel(const int32_t* __restrict__ nodes, uint64_t count, const int64_t* __restrict__ blk_off,
                                      const int32_t* __restrict__ blk_cols, int s, double* values, const double* __restrict__ scale,
                                      uint8_t* visit) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const double sc = *scale;
    for (uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < count; t += nwarps) {
        const int32_t I = nodes[t];
        const int64_t b = blk_off[I], cnt = blk_off[I + 1] - b;
        double* v = values + (int64_t)(s * s) * b;
        for (int64_t k = lane; k < cnt; k += 32) visit[blk_cols[b + k]] = 1;
        const int64_t len = (int64_t)s * s * cnt;  // s rows of s * cnt entries
        for (int64_t x = lane; x < len; x += 32) {
            const int64_t i = x / (s * cnt), c = x - i * s * cnt;  // row i of the node, column c = s k + j
            const int64_t k = c / s, j = c - k * s;
            v[x] = (blk_cols[b + k] == I && j == i) ? sc : 0.0;
        }
    }
}

// pass 2 (one warp per node): marked, non-Dirichlet rows lose their entries in Dirichlet columns
__global__ void dirichlet_cols_kernel(uint64_t num_nodes, const int64_t* __restrict__ blk_off, const int32_t* __restrict__ blk_cols, int s,
                                      double* values, const uint8_t* __restrict__ member, const uint8_t* __restrict__ visit) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t I = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < num_nodes; I += nwarps) {
        if (!visit[I] || member[I]) continue;
        const int64_t b = blk_off[I], cnt = blk_off[I + 1] - b;
        double* v = values + (int64_t)(s * s) * b;
        for (int64_t k = 0; k < cnt; ++k) {
            if (!member[blk_cols[b + k]]) continue;  // warp-uniform
            for (int x = lane; x < s * s; x += 32) {
                const int i = x / s, j = x - i * s;
                v[(int64_t)i * s * cnt + s * k + j] = 0.0;
            }
        }
    }
}

__global__ void mark_members_kernel(const int32_t* __restrict__ nodes, uint64_t count, uint8_t* member) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) member[nodes[t]] = 1;
}

}  // namespace fb200

using namespace fb200;

extern "C" fb200_status fb200_apply_homogeneous_dirichlet_bc_csr(fb200_ctx* ctx, uint64_t num_dirichlet_nodes, const uint64_t* nodes,
                                                                 double* scale_out) {
    if (!ctx) return FB200_ERR_STATE;
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no pattern / values: assemble first");
    if (num_dirichlet_nodes && !nodes) return fail(ctx, FB200_ERR_SHAPE, "null node list");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> h(num_dirichlet_nodes);
    for (uint64_t k = 0; k < num_dirichlet_nodes; ++k) {
        if (nodes[k] >= ctx->N) return fail(ctx, FB200_ERR_INDEX_OOB, "Dirichlet node out of range", (int64_t)k);
        h[k] = (int32_t)nodes[k];
    }
    const int s = ctx->sdim;
    int32_t* d_nodes = nullptr;
    uint8_t* d_flags = nullptr;  // member[N] | visit[N]
    unsigned long long* d_first = nullptr;
    double* d_scale = nullptr;
    fb200_status st = dev_alloc(ctx, &d_nodes, num_dirichlet_nodes);
    if (st == FB200_OK) st = dev_alloc(ctx, &d_flags, 2 * ctx->N);
    if (st == FB200_OK) st = dev_alloc(ctx, &d_first, 1);
    if (st == FB200_OK) st = dev_alloc(ctx, &d_scale, 1);
    double scale = 1.0;
    if (st == FB200_OK) {
        cudaMemcpyAsync(d_nodes, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(d_flags, 0, 2 * std::max<uint64_t>(ctx->N, 1), ctx->stream);
        cudaMemsetAsync(d_first, 0xff, sizeof(unsigned long long), ctx->stream);
        const int nb = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(ctx->N, 256), (uint64_t)ctx->sm_count * 16));
        if (ctx->N) {
            first_diagonal_kernel<<<nb, 256, 0, ctx->stream>>>(ctx->d_blk_off, ctx->d_blk_cols, ctx->N, s, ctx->d_values, d_first);
            st = check_launch(ctx, "first_diagonal_kernel");
        }
        if (st == FB200_OK) {
            read_diagonal_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_blk_off, ctx->d_blk_cols, s, ctx->d_values, d_first, d_scale);
            st = check_launch(ctx, "read_diagonal_kernel");
        }
        if (st == FB200_OK && num_dirichlet_nodes) {
            const int nd = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(num_dirichlet_nodes * 32, 256), (uint64_t)ctx->sm_count * 16));
            mark_members_kernel<<<nd, 256, 0, ctx->stream>>>(d_nodes, num_dirichlet_nodes, d_flags);
            st = check_launch(ctx, "mark_members_kernel");
            if (st == FB200_OK) {
                dirichlet_rows_kernel<<<nd, 256, 0, ctx->stream>>>(d_nodes, num_dirichlet_nodes, ctx->d_blk_off, ctx->d_blk_cols, s, ctx->d_values,
                                                                   d_scale, d_flags + ctx->N);
                st = check_launch(ctx, "dirichlet_rows_kernel");
            }
            if (st == FB200_OK) {
                const int nw = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(ctx->N * 32, 256), (uint64_t)ctx->sm_count * 16));
                dirichlet_cols_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->N, ctx->d_blk_off, ctx->d_blk_cols, s, ctx->d_values, d_flags, d_flags + ctx->N);
                st = check_launch(ctx, "dirichlet_cols_kernel");
            }
        }
        if (st == FB200_OK) {
            cudaError_t e = cudaMemcpyAsync(&scale, d_scale, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
            if (e != cudaSuccess) st = cuda_fail(ctx, e, "D2H Dirichlet scale");
        }
    }
    cudaStreamSynchronize(ctx->stream);
    if (d_nodes) cudaFree(d_nodes);
    if (d_flags) cudaFree(d_flags);
    if (d_first) cudaFree(d_first);
    if (d_scale) cudaFree(d_scale);
    if (st == FB200_OK && scale_out) *scale_out = scale;
    return st;
}
